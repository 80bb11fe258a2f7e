#!/usr/bin/env python3
"""Attribute ncu warp-stall samples to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <kernel-substring> <lib.so> [launch-index]
Joins `ncu --page source --csv` (SASS rows, in order) with `nvdisasm -g` line annotations of the same cubin."""
import csv, os, re, subprocess, sys, tempfile
rep, kname, so = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (tmp, os.path.abspath(so)), shell=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "host_tables" not in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# collect per-kernel instruction -> (file, line)
kern = {}; cur = None; loc = ("?", 0)
for l in sass:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: cur = m.group(1); kern[cur] = []; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: loc = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l): kern[cur].append(loc)
cands = [k for k in kern if kname in k]; name = ([k for k in cands if os.environ.get("MANGLED", "\0") in k] or cands)[0]
locs = kern[name]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split into launches
launches = []; curL = None
for r in rows:
    if r and r[0] == "Kernel Name": curL = {"name": r[1], "hdr": None, "rows": []}; launches.append(curL); continue
    if curL is None: continue
    if r and r[0] == "Address": curL["hdr"] = r; continue
    if curL["hdr"]: curL["rows"].append(r)
L = [x for x in launches if kname in x["name"]][which]
h = L["hdr"]; ci = h.index("Warp Stall Sampling (All Samples)"); ii = h.index("Instructions Executed")
assert len(L["rows"]) == len(locs), (len(L["rows"]), len(locs))
agg = {}; tot = 0; toti = 0
for r, lc in zip(L["rows"], locs):
    v = int(r[ci]); n = int(r[ii]); tot += v; toti += n
    a = agg.setdefault(lc, [0, 0]); a[0] += v; a[1] += n
src = {}
print("kernel", L["name"][:60], "samples", tot, "warp-instr", toti)
for (f, ln), (v, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:int(os.environ.get("TOP", "40"))]:
    if f not in src:
        for d in ("metaeuk_b200/csrc", "."):
            p = os.path.join(d, f)
            if os.path.exists(p): src[f] = open(p).read().splitlines(); break
        else: src[f] = []
    text = src[f][ln - 1].strip()[:110] if 0 < ln <= len(src[f]) else ""
    print("%5.1f%% stall %5.1f%% inst  %s:%d  %s" % (100 * v / tot, 100 * n / toti, f, ln, text))
