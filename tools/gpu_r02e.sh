# round 2, call e: entry-parallel gather in k_prefilter; grouped stream alignment launches on side streams; ncu of the big alignment groups and of k_prefilter
mkdir -p gpurun_out
T=${1:-r02e}
run() { n=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu"
run sw6 $B --phase-stats
run sw5 $B --sw-kernel 5
run sw6_priv $B --opt sw_ring=-1
run sw6_s1 $B --streams 1
run sw5_s1 $B --streams 1 --sw-kernel 5
run c3s --workload c3s --steps 3 --warmup 3 --no-cpu
run c2 --workload c2 --steps 5 --warmup 3 --no-cpu
python - <<PY
import json
for n in ("sw6","sw5","sw6_priv","sw6_s1","sw5_s1","c3s","c2"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d.get("gcups_sw"), d.get("prefilter_phase_pct"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sw_stream -s 4 -c 6 -o gpurun_out/${T}_sw_stream_c4m -f python bench.py --workload c4m --steps 1 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_stream.log 2>&1; echo "ncu stream exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_prefilter -s 1 -c 1 -o gpurun_out/${T}_prefilter_c4m -f python bench.py --workload c4m --steps 1 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_pref.log 2>&1; echo "ncu prefilter exit $?"
ls -la gpurun_out | tail -4
