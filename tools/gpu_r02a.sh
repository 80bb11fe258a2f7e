# round 2, first GPU-box call: gpurun --timeout 1700 -- 'bash tools/gpu_r02a.sh r02a'
mkdir -p gpurun_out
T=${1:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${T}_gpu.txt; nproc >> gpurun_out/${T}_gpu.txt; df -h /tmp | tail -1 >> gpurun_out/${T}_gpu.txt
timeout 700 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
MB200_TRACE=1 timeout 300 python bench.py --workload c4s --steps 3 --warmup 3 > gpurun_out/${T}_bench_c4s.json 2> gpurun_out/${T}_bench_c4s.err; echo "c4s exit $?"
MB200_TRACE=1 timeout 400 python bench.py --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu > gpurun_out/${T}_bench_c4m.json 2> gpurun_out/${T}_bench_c4m.err; echo "c4m exit $?"
MB200_TRACE=1 timeout 600 python bench.py --workload c4 --steps 1 --warmup 1 --allow-short-warmup --no-cpu > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "c4 exit $?"
python - <<PY
import json
for n in ("c4s","c4m","c4"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, "gcups %.0f"%d.get("gcups_sw",0), "roofline %.3f"%d["roofline"]["frac"], d.get("parity"), d.get("cpu_baseline",{}).get("value"), d["index"])
    except Exception as e: print(n, "no result", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_c4m.csv python bench.py --workload c4m --steps 1 --warmup 1 --allow-short-warmup --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1; echo "ncu launch list exit $?"
tail -3 gpurun_out/${T}_bench_c4.err
ls -la gpurun_out | tail -8
