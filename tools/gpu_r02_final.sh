# round 2, final measurement set: GPU suite, default bench line (full C4) + reference arm, C2 / C3-10 lines, launch list and ncu captures of the two top kernels
mkdir -p gpurun_out
T=${1:-r02f}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${T}_pytest_gpu.log
( time python bench.py ) > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "bench exit $?"; tail -4 gpurun_out/${T}_bench_c4.err
( time python bench.py --impl reference ) > gpurun_out/${T}_ref_c4.json 2> gpurun_out/${T}_ref_c4.err; echo "ref exit $?"; tail -4 gpurun_out/${T}_ref_c4.err
run() { n=$1; shift; timeout 500 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
run c2 --workload c2 --steps 5 --warmup 3
run c3s --workload c3s --steps 3 --warmup 3
run c4m_s1 --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1 --phase-stats
python - <<PY
import json
for n in ("bench_c4","ref_c4","c2","c3s","c4m_s1"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0) or d.get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d.get("parity",{}).get("blocks_checked"), (d.get("roofline") or {}).get("frac"), ((d.get("roofline") or {}).get("random_access") or {}).get("frac"), (d.get("roofline_sw") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(n, "no result", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c4m.csv python bench.py --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_launch.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prefilter -s 2 -c 1 -o gpurun_out/${T}_prefilter_c4 -f python bench.py --workload c4 --steps 1 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_pref.log 2>&1; echo "ncu prefilter exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sw_stream -s 5 -c 5 -o gpurun_out/${T}_sw_stream_c4m -f python bench.py --workload c4m --steps 1 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_stream.log 2>&1; echo "ncu stream exit $?"
ls -la gpurun_out | tail -5
