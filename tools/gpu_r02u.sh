# round 2, call u: hint 1 vs 3 on the cleaned build (one stream); contig batch size vs overlap on a rank-sized shard (C4/10 = 50 M nucleotides, two streams)
mkdir -p gpurun_out
T=${1:-r02u}
run() { n=$1; shift; timeout 300 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu"
run h1 $B --streams 1 --opt prefilter_stream_hint=1
run h3 $B --streams 1 --opt prefilter_stream_hint=3
run b32 $B --streams 2
run b12 $B --streams 2 --batch-nt 12500000
run b6 $B --streams 2 --batch-nt 6250000
run b4 $B --streams 2 --batch-nt 4000000
run b6s3 $B --streams 3 --batch-nt 6250000
python - <<PY
import json
for n in ("h1","h3","b32","b12","b6","b4","b6s3"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g busy ms/step %.0f e2e ms %.0f batches %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["contig_batches_per_gpu"]), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
