# round 2, call b: overlap A/B, prefilter config A/B with phase stats, fragment-workload parity, reference arm, ncu full of the two top kernels
mkdir -p gpurun_out
T=${1:-r02b}
run() { n=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n exit $?"; }
run c4m_s2 --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --phase-stats
run c4m_s1 --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --streams 1
run c4m_cfg2 --workload c4m --steps 2 --warmup 1 --allow-short-warmup --no-cpu --prefilter-config 2 --phase-stats
run c3s --workload c3s --steps 3 --warmup 3
run c2 --workload c2 --steps 5 --warmup 3
run ref_c4 --workload c4 --impl reference --steps 2 --warmup 1
python - <<PY
import json
for n in ("c4m_s2","c4m_s1","c4m_cfg2","c3s","c2","ref_c4"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%n)); print(n, "value %.4g e2e %.4g ms/step %.0f"%(d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), {k:round(v,1) for k,v in d.get("stages_ms",{}).items()}, d.get("parity"), d.get("prefilter_phase_pct"), d.get("cpu_baseline"))
    except Exception as e: print(n, "no result", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sw_wave -s 2 -c 1 -o gpurun_out/${T}_sw_wave_c4m -f python bench.py --workload c4m --steps 1 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_wave.log 2>&1; echo "ncu wave exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_prefilter -s 1 -c 1 -o gpurun_out/${T}_prefilter_c4m -f python bench.py --workload c4m --steps 1 --warmup 1 --allow-short-warmup --no-cpu --streams 1 > gpurun_out/${T}_ncu_pref.log 2>&1; echo "ncu prefilter exit $?"
ls -la gpurun_out | tail -6
