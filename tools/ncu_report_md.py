#!/usr/bin/env python3
"""Markdown summary of one kernel from an .ncu-rep: headline metrics, stall reasons, hottest source lines.
usage: ncu_report_md.py <report.ncu-rep> <kernel-substring> <lib.so> [launch-index] > profiles/<name>.md"""
import csv, os, subprocess, sys
rep, kname, so = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
L = [dict(zip(hdr, r)) for r in rows[2:] if kname in r[hdr.index("Kernel Name")]][which]
U = dict(zip(hdr, units))
def g(k): return L.get(k, "n/a")
def f(k, nd=2):
    try: return ("%." + str(nd) + "f") % float(L[k].replace(",", ""))
    except Exception: return "n/a"
print("| metric | value |\n|---|---|")
print("| kernel | `%s` grid %s x block %s, %s regs/thread, %s KB dynamic smem/block |" % (L["Kernel Name"][:40], g("launch__grid_size"), g("launch__block_size"),
      g("launch__registers_per_thread"), f("launch__shared_mem_per_block_dynamic")))
print("| gpu__time_duration.sum | %s %s |" % (f("gpu__time_duration.sum", 3), U.get("gpu__time_duration.sum", "")))
print("| dram__bytes_read.sum / dram__bytes_write.sum | %s / %s %s |" % (f("dram__bytes_read.sum", 3), f("dram__bytes_write.sum", 3), U.get("dram__bytes_read.sum", "")))
print("| gpu__dram_throughput (pct of peak) | %s %% |" % f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1))
print("| smsp__issue_active | %s %% |" % f("smsp__issue_active.avg.pct_of_peak_sustained_active", 1))
print("| sm__warps_active | %s %% |" % f("sm__warps_active.avg.pct_of_peak_sustained_active", 1))
print("| ALU pipe / LSU pipe (pct of peak) | %s %% / %s %% |" % (f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1), f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1)))
print("| l1tex / L2 sector hit rate | %s %% / %s %% |" % (f("l1tex__t_sector_hit_rate.pct", 1), f("lts__t_sector_hit_rate.pct", 1)))
print("| smsp__inst_executed.sum | %s |" % g("smsp__inst_executed.sum"))
print("| active threads per warp instruction | %s / 32 |" % f("smsp__thread_inst_executed_per_inst_executed.ratio", 1))
print("| shared-memory bank conflicts | %s |" % g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"))
st = []
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
        try: st.append((float(L[k]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except Exception: pass
st.sort(reverse=True)
print("| stall cycles per issued instruction | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
print("| occupancy limits (blocks/SM) | registers %s, shared memory %s |" % (f("launch__occupancy_limit_registers", 0), f("launch__occupancy_limit_shared_mem", 0)))
print()
print("Hottest source lines (warp-stall samples / executed instructions, `ncu --page source` joined with `nvdisasm -g`):\n\n```")
env = dict(os.environ, TOP=os.environ.get("TOP", "16"))
print(subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_lines.py"), rep, kname, so] + sys.argv[4:5], capture_output=True, text=True, env=env).stdout.rstrip()[:6000])
print("```")
