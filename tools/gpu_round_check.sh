# one GPU-box call: the whole GPU test suite, the default bench line (C2), the C4/20 and C3/10 lines, the ORF line and smoke()
mkdir -p gpurun_out
T=${1:-r01g}
timeout 480 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "c2 exit $?"
timeout 200 python bench.py --workload c4s --steps 3 --warmup 3 > gpurun_out/${T}_bench_c4s.json 2> gpurun_out/${T}_bench_c4s.err; echo "c4s exit $?"
timeout 200 python bench.py --workload c3s --steps 3 --warmup 3 --cpu-sample 512 > gpurun_out/${T}_bench_c3s.json 2> gpurun_out/${T}_bench_c3s.err; echo "c3s exit $?"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"
python - <<PY
import json
for n in ("c2","c4s","c3s"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_%s.json"%n)); print(n, "value %.4g e2e %.4g"%(d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, "gcups %.0f"%d.get("gcups_sw",0), "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(n, "no result", e)
PY
