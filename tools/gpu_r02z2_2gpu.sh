# round 2, last multi-GPU check (2 GPUs): the contig-sharded full C4 under torchrun with the final library
mkdir -p gpurun_out
T=${1:-r02z2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_c4_2gpu.json 2> gpurun_out/${T}_c4_2gpu.err; echo "c4 2gpu exit $?"
python - <<PY
import json
for n in ("c4_2gpu",):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g busy ms %.0f e2e ms/step %.0f batches %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"].get("ms_per_step",0), d["config"]["contig_batches_per_gpu"]), d.get("parity",{}).get("sha256_match"), d.get("parity",{}).get("blocks_checked"), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["index"].items()}, d["n_gpus"], d["scaling"])
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-2500:])
PY
