# round 2, call p: kernel configurations again, after the long-list fixes (second occupancy filter, shared-memory double-hit stage, chunked sort)
mkdir -p gpurun_out
T=${1:-r02p}
run() { n=$1; shift; timeout 500 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1 --phase-stats"
run all3 $B --prefilter-config 3
run s128c3 $B --opt prefilter_short_len=128 --opt prefilter_short_config=3
run s64c3 $B --opt prefilter_short_len=64 --opt prefilter_short_config=3
run s64c4 $B --opt prefilter_short_len=64 --opt prefilter_short_config=4
python - <<PY
import json
for n in ("all3","s128c3","s64c3","s64c4"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d.get("prefilter_phase_pct"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
