mkdir -p gpurun_out
run() { timeout 100 python bench.py --workload c4s --steps 3 --warmup 3 --lane-max-hits $1 --cpu-sample 64 > gpurun_out/r01f_c4s_h$1.json 2> gpurun_out/r01f_c4s_h$1.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r01f_c4s_h$1.json")); print("c4s maxhits=$1", "value %.4g e2e %.4g"%(d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["stages_ms"].items()}, "gcups %.0f"%d.get("gcups_sw",0))
except Exception as e: print("no result", e)
PY
}
run 8
run 24
# launch list of one C4/20 step (per-launch durations are cold-cache and serialised: shares, not absolutes)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 80 --csv --log-file gpurun_out/r01f_launches_c4s.csv python bench.py --workload c4s --steps 1 --warmup 1 --cpu-sample 8 > gpurun_out/r01f_ncu_launch.log 2>&1; echo "ncu launch list exit $?"
# one full capture of the wavefront alignment kernel (the <= 256-residue launch) and of the per-lane kernel
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_sw_wave|k_sw_lanes' -s 3 -c 3 -o gpurun_out/r01f_sw_wave_lanes -f python bench.py --workload c4s --steps 1 --warmup 1 --cpu-sample 8 > gpurun_out/r01f_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | tail -8
