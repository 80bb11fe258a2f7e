# round 2, call m (2 GPUs): the contig-sharded C4 path under torchrun (BASELINE configs[4] shape), index replicated by one ncclBroadcast
mkdir -p gpurun_out
T=${1:-r02m}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c4m --steps 2 --warmup 1 --allow-short-warmup > gpurun_out/${T}_c4m_2gpu.json 2> gpurun_out/${T}_c4m_2gpu.err; echo "c4m 2gpu exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/${T}_c4_2gpu.json 2> gpurun_out/${T}_c4_2gpu.err; echo "c4 2gpu exit $?"
python - <<PY
import json
for n in ("c4m_2gpu","c4_2gpu"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), d.get("parity",{}).get("sha256_match"), d.get("parity",{}).get("blocks_checked"), d["index"], d["n_gpus"], d["scaling"])
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-2500:])
PY
