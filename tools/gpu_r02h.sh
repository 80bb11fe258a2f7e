# round 2, call h: the default bench line (full C4), the reference arm, launch list and ncu captures of the two top kernels on the same command
mkdir -p gpurun_out
T=${1:-r02h}
( time python bench.py ) > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "bench exit $?"; tail -4 gpurun_out/${T}_bench_c4.err
( time python bench.py --impl reference ) > gpurun_out/${T}_ref_c4.json 2> gpurun_out/${T}_ref_c4.err; echo "ref exit $?"; tail -4 gpurun_out/${T}_ref_c4.err
python - <<PY
import json
for n in ("bench_c4","ref_c4"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0) or d.get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}), d.get("roofline"), d.get("roofline_sw"), d.get("cpu_baseline"))
    except Exception as e: print(n, "no result", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c4m.csv python bench.py --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_launch.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prefilter -s 2 -c 1 -o gpurun_out/${T}_prefilter_c4 -f python bench.py --workload c4 --steps 1 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_pref.log 2>&1; echo "ncu prefilter exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sw_stream -s 8 -c 2 -o gpurun_out/${T}_sw_stream_c4m -f python bench.py --workload c4m --steps 1 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_stream.log 2>&1; echo "ncu stream exit $?"
ls -la gpurun_out | tail -5
