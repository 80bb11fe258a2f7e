# round 2, call l: L2 prefetch of the index buckets from the expansion phase
mkdir -p gpurun_out
T=${1:-r02l}
run() { n=$1; shift; timeout 500 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
run c4m_s1 --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1 --phase-stats
run c4m --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu
run c2 --workload c2 --steps 5 --warmup 3 --no-cpu --phase-stats
run c3s --workload c3s --steps 3 --warmup 3 --no-cpu --phase-stats
python - <<PY
import json
for n in ("c4m_s1","c4m","c2","c3s"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d.get("prefilter_phase_pct"), d["roofline"].get("random_access",{}).get("frac"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
