mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider -k "align_golden or predict_exons or reference_own or larger_synthetic or long_queries or properties" > gpurun_out/r01d_pytest_sw.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01d_pytest_sw.log
tail -4 gpurun_out/r01d_pytest_sw.log
run() { # workload sw maxhits
  timeout 150 python bench.py --workload $1 --steps 3 --warmup 3 --sw-kernel $2 --lane-max-hits $3 --cpu-sample 256 > gpurun_out/r01d_$1_sw$2_h$3.json 2> gpurun_out/r01d_$1_sw$2_h$3.err; echo "$1 sw=$2 h=$3 exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r01d_$1_sw$2_h$3.json")); print("$1 sw=$2 maxhits=$3", "value %.4g e2e %.4g"%(d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["stages_ms"].items()}, "gcups %.0f"%d.get("gcups_sw",0))
except Exception as e: print("no result", e)
PY
}
run c4s 5 8
run c4s 5 2
run c3s 2 8
run c3s 5 8
