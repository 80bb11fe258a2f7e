# Round 2, first GPU-box call (gpurun --timeout 1500 -- 'bash tools/gpu_round2_first_call.sh r02a'):
#  1. the GPU suite -- tests/test_zz_gpu_adversarial.py (flag sets, second prefilter pass, adversarial fixture) has only run on the SIMT emulation so far
#  2. smoke(), the default bench line, C4/20, C3/10
#  3. launch list of a C4/20 step + one full ncu capture each of k_prefilter (C2) and k_sw_trace2 (C4/20), the two kernels DESIGN.md section 9 names
mkdir -p gpurun_out
T=${1:-r02a}
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "c2 exit $?"
timeout 200 python bench.py --workload c4s --steps 3 --warmup 3 > gpurun_out/${T}_bench_c4s.json 2> gpurun_out/${T}_bench_c4s.err; echo "c4s exit $?"
timeout 200 python bench.py --workload c3s --steps 3 --warmup 3 --cpu-sample 512 > gpurun_out/${T}_bench_c3s.json 2> gpurun_out/${T}_bench_c3s.err; echo "c3s exit $?"
python - <<PY
import json
for n in ("c2","c4s","c3s"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_%s.json"%n)); print(n, "value %.4g e2e %.4g"%(d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, "gcups %.0f"%d.get("gcups_sw",0), "roofline %.3f"%d["roofline"]["frac"])
    except Exception as e: print(n, "no result", e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 80 --csv --log-file gpurun_out/${T}_launches_c4s.csv python bench.py --workload c4s --steps 1 --warmup 1 --cpu-sample 8 > gpurun_out/${T}_ncu_launch.log 2>&1; echo "ncu launch list exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_prefilter -s 3 -c 1 -o gpurun_out/${T}_prefilter_c2 -f python bench.py --steps 1 --warmup 3 --cpu-sample 8 > gpurun_out/${T}_ncu_prefilter.log 2>&1; echo "ncu prefilter exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_sw_trace2 -s 1 -c 1 -o gpurun_out/${T}_trace_c4s -f python bench.py --workload c4s --steps 1 --warmup 1 --cpu-sample 8 > gpurun_out/${T}_ncu_trace.log 2>&1; echo "ncu trace exit $?"
ls -la gpurun_out | tail -12
