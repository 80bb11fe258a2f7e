mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider -k "align_golden or predict_exons or reference_own or larger_synthetic or long_queries" > gpurun_out/r01c_pytest_sw.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01c_pytest_sw.log
tail -4 gpurun_out/r01c_pytest_sw.log
for v in "2 0" "4 2" "4 8" "4 1024"; do set -- $v
  timeout 100 python bench.py --workload c4s --steps 3 --warmup 3 --sw-kernel $1 --lane-max-hits $2 > gpurun_out/r01c_c4s_sw$1_h$2.json 2> gpurun_out/r01c_c4s_sw$1_h$2.err; echo "c4s sw=$1 h=$2 exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r01c_c4s_sw$1_h$2.json")); print("sw=$1 maxhits=$2", "value %.4g e2e %.4g"%(d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["stages_ms"].items()}, "gcups %.0f"%d.get("gcups_sw",0))
except Exception as e: print("no result", e)
PY
done
