# round 2, call y: persistent host work arrays + parallel compaction in mb200_predict_exons: host phases (MB200_TRACE) on C4/10, then the default full-C4 line
mkdir -p gpurun_out
T=${1:-r02y}
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu"
MB200_TRACE=1 timeout 300 python bench.py $B --streams 1 > gpurun_out/${T}_s1.json 2> gpurun_out/${T}_s1.err; echo "s1 exit $?"
MB200_TRACE=1 timeout 300 python bench.py $B --streams 2 > gpurun_out/${T}_s2.json 2> gpurun_out/${T}_s2.err; echo "s2 exit $?"
grep "host phases" gpurun_out/${T}_s1.err | tail -1
grep "host phases" gpurun_out/${T}_s2.err | tail -2
( time python bench.py ) > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "bench exit $?"; tail -4 gpurun_out/${T}_bench_c4.err
python - <<PY
import json
for n in ("s1","s2","bench_c4"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g busy ms/step %.0f e2e ms %.0f batches %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["contig_batches_per_gpu"]), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity",{}).get("sha256_match"), d.get("parity",{}).get("blocks_checked"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-1500:])
PY
