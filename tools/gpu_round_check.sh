mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r01b_gpu.txt 2>&1
timeout 480 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/r01b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01b_pytest_gpu.log
tail -5 gpurun_out/r01b_pytest_gpu.log
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/r01b_bench_c2.json 2> gpurun_out/r01b_bench_c2.err; echo "c2 exit $?"
timeout 200 python bench.py --workload c4s --steps 3 --warmup 3 > gpurun_out/r01b_bench_c4s.json 2> gpurun_out/r01b_bench_c4s.err; echo "c4s exit $?"
timeout 60 python bench.py --workload orf --steps 5 --warmup 3 > gpurun_out/r01b_bench_orf.json 2> gpurun_out/r01b_bench_orf.err; echo "orf exit $?"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01b_smoke.log 2>&1; echo "smoke exit $?"
cat gpurun_out/r01b_bench_c4s.json | cut -c1-1500
