# round 2, call x: host phases of mb200_predict_exons (MB200_TRACE) on C4/10, one and two streams
mkdir -p gpurun_out
T=${1:-r02x}
B="--workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu"
MB200_TRACE=1 timeout 300 python bench.py $B --streams 1 > gpurun_out/${T}_s1.json 2> gpurun_out/${T}_s1.err; echo "s1 exit $?"
MB200_TRACE=1 timeout 300 python bench.py $B --streams 2 > gpurun_out/${T}_s2.json 2> gpurun_out/${T}_s2.err; echo "s2 exit $?"
grep "host phases" gpurun_out/${T}_s1.err | tail -2
grep "host phases" gpurun_out/${T}_s2.err | tail -4
nproc
