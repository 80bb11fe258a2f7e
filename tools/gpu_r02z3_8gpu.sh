# round 2, last 8-GPU measurement: the contig-sharded full C4 under torchrun at N = 8 with the final library
mkdir -p gpurun_out
T=${1:-r02z3}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/${T}_c4_8gpu.json 2> gpurun_out/${T}_c4_8gpu.err; echo "c4 8gpu exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_c4_8gpu.json")); print("c4_8gpu value %.4g e2e %.4g busy ms %.0f e2e ms/step %.0f batches %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"].get("ms_per_step",0), d["config"]["contig_batches_per_gpu"]), d.get("parity",{}).get("sha256_match"), d.get("parity",{}).get("blocks_checked"), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["index"].items()}, d["n_gpus"], d["scaling"])
except Exception as e: print("no result", e); print(open("gpurun_out/${T}_c4_8gpu.err").read()[-2500:])
PY
