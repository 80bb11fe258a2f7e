# round 2, call w: profile ring of the long-query alignment groups sized so that 8 warps x 2 CTAs fit (sw_ring_fit), C4/10 one stream and two streams
mkdir -p gpurun_out
T=${1:-r02w}
run() { n=$1; shift; timeout 300 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu"
run fit0 $B --streams 1 --opt sw_ring_fit=0
run fit1 $B --streams 1 --opt sw_ring_fit=1
run fit1_s2 $B --streams 2
python - <<PY
import json
for n in ("fit0","fit1","fit1_s2"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g busy ms/step %.0f e2e ms %.0f batches %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["contig_batches_per_gpu"]), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d.get("gcups_sw"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_sw_stream -c 10 --csv --log-file gpurun_out/${T}_ncu_sw.csv python bench.py $B --streams 1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit $?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${T}_ncu_sw.csv")) if len(r)>5 and r[0].isdigit()]
byid={}
for r in rows: byid.setdefault(int(r[0]),{"block":r[7]})[r[-3].split('.')[0]]=r[-1]
for i in sorted(byid): print(i, byid[i])
PY
