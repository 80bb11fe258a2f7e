# round 2, call q (8 GPUs): the contig-sharded full C4 under torchrun at N = 8 and 4 (BASELINE configs[4]), index replicated by one ncclBroadcast
mkdir -p gpurun_out
T=${1:-r02q}
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_c4_${N}gpu.json 2> gpurun_out/${T}_c4_${N}gpu.err; echo "c4 ${N}gpu exit $?"
done
python - <<PY
import json
for n in ("c4_8gpu","c4_4gpu"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), d.get("parity",{}).get("sha256_match"), d.get("parity",{}).get("blocks_checked"), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["index"].items()}, d["n_gpus"], d["scaling"])
    except Exception as e: print(n, "no result", e); print(open("gpurun_out/${T}_%s.err"%n).read()[-2500:])
PY
