# round 2, call c: prefilter configuration A/B on C4/10 (short-query split), GPU test suite
mkdir -p gpurun_out
T=${1:-r02c}
run() { n=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --phase-stats"
run base $B
run s64c3 $B --opt prefilter_short_len=64 --opt prefilter_short_config=3
run s64c4 $B --opt prefilter_short_len=64 --opt prefilter_short_config=4
run s128c3 $B --opt prefilter_short_len=128 --opt prefilter_short_config=3
run s128c4 $B --opt prefilter_short_len=128 --opt prefilter_short_config=4
run all3 $B --prefilter-config 3
python - <<PY
import json
for n in ("base","s64c3","s64c4","s128c3","s128c4","all3"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d.get("prefilter_phase_pct"))
    except Exception as e: print(n, "no result", e)
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${T}_pytest_gpu.log
