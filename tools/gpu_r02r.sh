# round 2, call r: index reads (bucket sectors, long entry lists) as streaming loads (L1 no-allocate, L2 evict-first) vs normally cached; device-busy time (union of stage intervals) as the
# denominator of the resident throughput
mkdir -p gpurun_out
T=${1:-r02r}
run() { n=$1; shift; timeout 500 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu"
run h0_s1 $B --streams 1 --phase-stats --opt prefilter_stream_hint=0
run h1_s1 $B --streams 1 --phase-stats --opt prefilter_stream_hint=1
run h1_s2 $B --streams 2
run h0_s2 $B --streams 2 --opt prefilter_stream_hint=0
run c2 --workload c2 --steps 5 --warmup 3 --no-cpu
run c2_h0 --workload c2 --steps 5 --warmup 3 --no-cpu --opt prefilter_stream_hint=0
run c3s --workload c3s --steps 3 --warmup 3 --no-cpu
python - <<PY
import json
for n in ("h0_s1","h1_s1","h1_s2","h0_s2","c2","c2_h0","c3s"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f e2e_ms %.0f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"].get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d.get("prefilter_phase_pct"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
M=dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum
for h in 0 1; do
timeout 400 ncu --metrics $M --clock-control none -k regex:k_prefilter -c 3 --csv --log-file gpurun_out/${T}_ncu_pref_h$h.csv python bench.py $B --streams 1 --opt prefilter_stream_hint=$h > gpurun_out/${T}_ncu_h$h.log 2>&1; echo "ncu h$h exit $?"
done
grep -v "^==" gpurun_out/${T}_ncu_pref_h0.csv | cut -d, -f5,13- | tail -20
grep -v "^==" gpurun_out/${T}_ncu_pref_h1.csv | cut -d, -f5,13- | tail -20
