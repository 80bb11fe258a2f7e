# round 2, call s: DRAM sectors per random 32-byte read (microbenchmark under ncu, three L2 fetch-granularity limits) and the limit in the real kernel
mkdir -p gpurun_out
T=${1:-r02s}
M=dram__sectors_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
for L in 0 32 128; do
  tools/microbench/randgather2 $L > gpurun_out/${T}_rg_$L.txt 2>&1
  timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_rg_ncu_$L.csv tools/microbench/randgather2 $L > /dev/null 2>&1; echo "rg $L exit $?"
done
cat gpurun_out/${T}_rg_*.txt
python - <<PY
import csv
for L in (0,32,128):
    rows=[r for r in csv.reader(open("gpurun_out/${T}_rg_ncu_%d.csv"%L)) if len(r)>5 and r[0].isdigit()]
    byid={}
    for r in rows: byid.setdefault(int(r[0]),{})[r[-3]]=float(r[-1].replace(',',''))
    n=148*4*512*64*4
    for i in sorted(byid):
        d=byid[i]
        if d.get("gpu__time_duration.sum",0)>2e6: print("limit",L,"kernel",i,"sectors/access %.2f"%(d["dram__sectors_read.sum"]/n),"L2 read misses/access %.2f"%(d["lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum"]/n),"ms %.2f"%(d["gpu__time_duration.sum"]/1e6))
PY
run() { n=$1; shift; timeout 500 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1"
run g32 $B --opt l2_fetch_granularity=32
run g128 $B --opt l2_fetch_granularity=128
python - <<PY
import json
for n in ("g32","g128"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
