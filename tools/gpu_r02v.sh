# round 2, call v: GPU suite on the new defaults (index reads evict-first + 64-byte fetch, adaptive contig batches) and the DRAM sectors of k_prefilter with them
mkdir -p gpurun_out
T=${1:-r02v}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${T}_pytest_gpu.log
M=dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:k_prefilter -c 2 --csv --log-file gpurun_out/${T}_ncu_pref.csv python bench.py --workload c4m --steps 1 --warmup 1 --allow-short-warmup --no-cpu --streams 1 --batch-nt 33554432 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit $?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${T}_ncu_pref.csv")) if len(r)>5 and r[0].isdigit()]
byid={}
for r in rows: byid.setdefault(r[0],{})[r[-3]]=float(r[-1].replace(',',''))
for i,d in byid.items(): print("launch",i, {k:("%.4g"%v) for k,v in d.items()})
PY
