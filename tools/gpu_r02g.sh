# round 2, call g: 256-bit bucket loads, register-resident warp sort, sorted walk for short multi-hit lists (A/B), microbenchmark with 256-bit loads
mkdir -p gpurun_out
T=${1:-r02g}
(cd tools/microbench && timeout 200 ./randgather) > gpurun_out/${T}_randgather.txt 2>&1; cat gpurun_out/${T}_randgather.txt
run() { n=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1 --phase-stats"
run new $B
run bk0 $B --opt prefilter_buckets=0
run walk0 $B --opt prefilter_sort_walk=0
run two --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu
run c2 --workload c2 --steps 5 --warmup 3 --no-cpu --phase-stats
run c3s --workload c3s --steps 3 --warmup 3 --no-cpu --phase-stats
python - <<PY
import json
for n in ("new","bk0","walk0","two","c2","c3s"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d.get("prefilter_phase_pct"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${T}_pytest_gpu.log
