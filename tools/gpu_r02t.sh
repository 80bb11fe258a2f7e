# round 2, call t: k_prefilter P4 variants on C4/10, one stream: 64-byte HBM fetch, L2 prefetch of the second bucket pair / of long lists, four buckets per lane in flight (separate build)
mkdir -p gpurun_out
T=${1:-r02t}
run() { n=$1; shift; timeout 300 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1"
for h in 3 5 9 13 15; do run h$h $B --opt prefilter_stream_hint=$h; done
export MB200_LIB_PATH=$PWD/metaeuk_b200/libmetaeuk_b200_quad.so
run quad_h1 $B --opt prefilter_stream_hint=1
run quad_h3 $B --opt prefilter_stream_hint=3
unset MB200_LIB_PATH
python - <<PY
import json
for n in ("h3","h5","h9","h13","h15","quad_h1","quad_h3"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
