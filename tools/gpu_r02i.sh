# round 2, call i: reverse pass for every block-count group (fallback kernel only for saturated scores / queries > 1024), batch size A/B on the full C4
mkdir -p gpurun_out
T=${1:-r02i}
run() { n=$1; shift; timeout 500 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
run c4m_s1 --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1
B="--workload c4 --steps 2 --warmup 1 --allow-short-warmup --no-cpu"
run c4 $B
run c4_b64 $B --batch-nt 64000000
run c4_b100 $B --batch-nt 100000000
run c4_b16 $B --batch-nt 16000000
python - <<PY
import json
for n in ("c4m_s1","c4","c4_b64","c4_b100","c4_b16"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d["config"].get("contig_batches_per_gpu"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${T}_pytest_gpu.log
