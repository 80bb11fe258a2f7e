# round 2, call n: number of worker contexts (host threads + streams) of mb200_predict_exons on the full C4
mkdir -p gpurun_out
T=${1:-r02n}
run() { n=$1; shift; timeout 500 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4 --steps 2 --warmup 1 --allow-short-warmup --no-cpu"
run s2 $B --streams 2
run s3 $B --streams 3
run s4 $B --streams 4
python - <<PY
import json
for n in ("s2","s3","s4"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d["predict_streams"])
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
