#!/usr/bin/env python3
"""bench.py -- predictexons hot-path throughput (query fragments / s) on B200, with the reference's CPU path timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c4s|c3|c3s|c2|orf|reduce] [--impl reference]

Default workload c4 = BASELINE.json configs[3]/[4], the configuration the metric is quoted on: full predictexons (six-frame ORF
extraction + translation, k-mer prefilter, gapped alignment, resultspercontig join, exon-chaining DP) on 10 K synthetic 50 kb contigs
with planted genes against 500 K proteins.  A "step" = one pass of mb200_predict_exons over the whole contig set.  With --gpus N > 1
(torchrun, one rank per GPU) the SAME contig set is sharded by contig across the ranks (strong scaling); the target index is built once on
rank 0 and replicated by ONE ncclBroadcast issued from inside the library (mb200_index_broadcast), no other collective on the data path.
c2 / c3 (configs[1]/[2]) are the prefilter-only and prefilter+align fragment workloads (kept for the kernel rooflines; fragment shards,
weak scaling).

JSON keys: value = fragments / (sum of CUDA-event stage times on the library stream, contigs of a batch resident in HBM), max over ranks;
e2e = fragments / wall clock of the C-ABI call (pinned HOST contigs in, HOST call records out, every copy inside the timed region);
roofline = dominant kernel's algorithmic bytes / its CUDA-event time against MEASURED_PEAKS.json; parity = SHA-256 of the
timed step's output against digests of the UNMODIFIED reference's output on the same data (tests/golden/bench_digests.json, produced in
the build container by tests/golden/make_bench_digests.py) -- a mismatch makes the run exit non-zero; cpu_baseline / --impl reference =
the reference's own AVX2 binary (oracle/_ref/metaeuk-ref) on the box's host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

# torchrun exports OMP_NUM_THREADS=1; the host side of the library (index build, alignment epilogue) is OpenMP code: give every rank its
# share of the cores (must happen before any OpenMP runtime is loaded)
if os.environ.get("OMP_NUM_THREADS") == "1" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

BG = np.array([7.8, 1.9, 5.4, 6.3, 3.9, 7.4, 2.3, 5.1, 5.7, 9.0, 2.2, 4.5, 5.2, 4.3, 5.1, 7.1, 5.8, 6.4, 1.3, 3.2])
BG = BG / BG.sum()
DIGEST_BLOCK = 250     # contigs per parity digest (divides every shard of 10 K contigs over 1/2/4/8 ranks)
WORKLOADS = {
    "c2": dict(n_prot=50_000, n_frag=100_000, align=False, desc="C2: 100K fragments (~100 aa) x 50K proteins (~330 aa), prefilter-only"),
    "c3": dict(n_prot=500_000, n_frag=1_000_000, align=True, desc="C3: 1M fragments x 500K proteins, prefilter + gapped SW align"),
    "c3s": dict(n_prot=500_000, n_frag=100_000, align=True, desc="C3/10: 100K fragments x 500K proteins, prefilter + gapped SW align (profiling size)"),
    "c4": dict(n_prot=500_000, n_contigs=10_000, contig_len=50_000, predict=True, n_frag=0, align=True,
               desc="C4: full predictexons (ORFs + prefilter + align + exon chaining) on 10K contigs x 50 kb with planted genes vs 500K proteins"),
    "c4m": dict(n_prot=500_000, n_contigs=1_000, contig_len=50_000, predict=True, n_frag=0, align=True,
                desc="C4/10: full predictexons on the first 1000 contigs of C4 vs 500K proteins (profiling size)"),
    "c4s": dict(n_prot=50_000, n_contigs=500, contig_len=50_000, predict=True, n_frag=0, align=True,
                desc="C4/20: full predictexons on 500 contigs x 50 kb with planted genes vs 50K proteins (profiling size)"),
    "tiny": dict(n_prot=2_000, n_frag=4_000, align=True, desc="tiny: 4K fragments x 2K proteins (plumbing check)"),
    "reduce": dict(n_contigs=200_000, reduce=True, n_prot=0, n_frag=0, align=False,
                   desc="reduceredundancy stage (SURVEY 8f-2): 200 K contigs x ~24 calls (1-4 exons, exon keys shared inside a contig)"),
    "orf": dict(n_contigs=4_000, contig_len=50_000, orf=True, n_prot=0, n_frag=0, align=False,
                desc="ORF stage (SURVEY 8f-1): 4000 contigs x 50 kb random DNA -> six-frame fragments >= 15 codons, translated"),
}


# ---------------------------------------------------------------------------------------------------------------------
# synthetic inputs, seeded; shared with tests/golden/make_bench_digests.py so both sides see the same bytes
# ---------------------------------------------------------------------------------------------------------------------
def gen_proteins(n, seed=1234):
    rng = np.random.default_rng(seed)
    lens = np.maximum(60, np.rint(rng.normal(330, 120, n))).astype(np.int64)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    res = rng.choice(20, size=int(off[-1]), p=BG).astype(np.uint8)
    return res, off


def gen_fragments(n, prot, poff, seed=4321):
    """half: mutated slices of random proteins (substitution rate in {0.1,0.3,0.5}); half: background"""
    rng = np.random.default_rng(seed)
    lens = np.maximum(15, np.rint(rng.normal(100, 40, n))).astype(np.int64)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    tot = int(off[-1])
    res = rng.choice(20, size=tot, p=BG).astype(np.uint8)
    homolog = (np.arange(n) % 2) == 0
    pidx = rng.integers(0, len(poff) - 1, n)
    plen = (poff[1:] - poff[:-1]).astype(np.int64)[pidx]
    use = np.minimum(lens, plen)
    start = (rng.random(n) * np.maximum(1, plen - use)).astype(np.int64)
    rate = rng.choice([0.1, 0.3, 0.5], n)
    frag_id = np.repeat(np.arange(n), lens)
    within = np.arange(tot) - np.repeat(off[:-1].astype(np.int64), lens)
    src = poff[:-1].astype(np.int64)[pidx][frag_id] + start[frag_id] + np.minimum(within, use[frag_id] - 1)
    take = homolog[frag_id] & (within < use[frag_id]) & (rng.random(tot) >= rate[frag_id])
    res[take] = prot[src[take]]
    return res, off


def gen_contigs(n, length, seed=1234):
    rng = np.random.default_rng(seed)
    nuc = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n * length, dtype=np.uint8)]
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(length))
    return np.ascontiguousarray(nuc), off


_CODON = np.frombuffer(b"GCTTGTGATGAATTTGGTCATATTAAACTGATGAATCCTCAACGTTCTACTGTTTGGTAT", dtype=np.uint8).reshape(20, 3)   # one codon per residue, A..Y in matrix order
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def gen_planted_contigs(c0, c1, length, prot, poff, seed=1234):
    """SURVEY.md 8d contigs [c0, c1) of the workload: i.i.d. ACGT; after 70 % of the inter-genic gaps (300-1500 nt) a gene = a random protein with
    per-residue substitution rate in {0.1, 0.3, 0.5}, back-translated, split into 1-4 exons by introns of 40-600 nt, 50 % reverse-complemented.
    Every contig has its own generator seeded by (seed, contig index): contig c is the same bytes whatever range or rank generates it."""
    n = c1 - c0
    nuc = np.empty(n * length, np.uint8)
    nprot = len(poff) - 1
    genes = 0
    for c in range(c0, c1):
        rng = np.random.default_rng((seed, c))
        ctg = _ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]
        pos = 0
        while True:
            pos += int(rng.integers(300, 1500))
            if rng.random() >= 0.7:
                if pos >= length:
                    break
                continue
            pi = int(rng.integers(0, nprot)); p = prot[int(poff[pi]):int(poff[pi + 1])].copy()
            mut = rng.random(len(p)) < rng.choice([0.1, 0.3, 0.5])
            p[mut] = rng.choice(20, size=int(mut.sum()), p=BG)
            nex = int(rng.integers(1, 5))
            cuts = np.sort(rng.choice(np.arange(20, len(p) - 20), size=nex - 1, replace=False)) if (nex > 1 and len(p) > 60 + nex) else np.zeros(0, np.int64)
            pieces = []
            for a, b in zip(np.concatenate(([0], cuts)), np.concatenate((cuts, [len(p)]))):
                if pieces:
                    pieces.append(_ACGT[rng.integers(0, 4, size=int(rng.integers(40, 600)), dtype=np.uint8)])
                pieces.append(_CODON[p[int(a):int(b)]].reshape(-1))
            g = np.concatenate(pieces)
            if pos + len(g) > length:
                break
            if rng.random() < 0.5:
                g = _COMP[g[::-1]]
            ctg[pos: pos + len(g)] = g
            pos += len(g); genes += 1
        nuc[(c - c0) * length:(c - c0 + 1) * length] = ctg
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(length))
    return nuc, off, genes


# ---------------------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------------------
def pinned(a):
    """copy a numpy array into page-locked host memory (so the C ABI's cudaMemcpyAsync is a real DMA)"""
    import torch
    t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)) if a.dtype != np.uint64 else torch.int64, pin_memory=True)
    v = t.numpy().view(a.dtype)
    v[...] = a
    return v, t


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.dev = dev
        self.p = None
        self.lines = []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        sm, mx, reasons = [], 0, set()
        for l in self.lines:
            w = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(w[1])); mx = max(mx, float(w[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), w[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes_prefilter(st):
    """L + 21 L (profile) + 16 per similar k-mer (two 8-byte offsets) + 6 per index entry
    + per double-hit candidate (overlap length + 6) + 10 per result"""
    return 22 * st["query_residues"] + 16 * st["kmers"] + 6 * st["db_matches"] + st["diag_cells"] + 6 * st["double_hits"] + 10 * st["pref_results"]


def random_access_roofline(st, kernel_ms):
    """The prefilter's index lookups are independent random HBM reads (one 32-byte bucket per similar k-mer): HBM3e limits their RATE, not their bytes.  Ceiling = what
    tools/microbench/randgather.cu measured on a B200 of this pool for 8-byte and single-instruction 32-byte random reads from a 1.5 GB buffer (profiles/r02_randgather.txt)."""
    peak = 40.6
    ach = st["kmers"] / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    return {"what": "index bucket lookups (one random 256-bit HBM read per similar k-mer)", "achieved": ach, "peak": peak, "unit": "G accesses/s", "frac": ach / peak,
            "peak_source": "measured: tools/microbench/randgather.cu on B200, profiles/r02_randgather.txt (40.6 G/s from 1.5 GB; 37.8 G/s from 4 GB)",
            "bytes_note": "by default every random 8- or 32-byte read moves a 128-byte line of DRAM on this GPU (ncu: 3.83-4.07 dram sectors per access; 1.96 with the .L2::64B qualifier) and the rate is "
                          "39.4 G/s either way (tools/microbench/randgather2.cu, profiles/r02s_randgather2.txt): a rate ceiling, not a byte ceiling"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    p = os.path.join(ROOT, "profiles", "prefilter_traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


def golden_digests(workload):
    p = os.path.join(ROOT, "tests", "golden", "bench_digests.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


def emit(obj):
    """the ONE JSON line on the real stdout (libraries that chat on fd 1, e.g. NCCL's version banner, were redirected to stderr)"""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def build_index(ctx, prot, poff):
    """rank 0: tantan-mask the targets as the reference's index builder does (--mask 1), build the k-mer index, upload.  -> seconds per phase"""
    from metaeuk_b200 import lib as mlib
    t0 = time.perf_counter()
    masked, n_masked = mlib.mask_targets(prot, poff)
    t1 = time.perf_counter()
    ctx.index_build(masked, prot, poff)
    return {"mask_s": t1 - t0, "build_upload_s": time.perf_counter() - t1, "masked_residues": int(n_masked)}


def replicate_index(ctx, rank, world):
    """ONE ncclBroadcast of the index arena, issued by the library on its own communicator (the unique id travels through torch.distributed)"""
    import torch
    import torch.distributed as dist
    if world == 1:
        return {"broadcast_s": 0.0, "broadcast_bytes": 0, "comm_init_s": 0.0}
    t0 = time.perf_counter()
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    torch.cuda.synchronize(); dist.barrier()
    t1 = time.perf_counter()
    nbytes, ms = ctx.index_broadcast(0)
    torch.cuda.synchronize(); dist.barrier()
    return {"comm_init_s": t1 - t0, "broadcast_s": time.perf_counter() - t1, "broadcast_ms_device": ms, "broadcast_bytes": int(nbytes),
            "broadcast_gbs": nbytes / max(ms, 1e-6) / 1e6}


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref/metaeuk-ref), bounded samples
# ---------------------------------------------------------------------------------------------------------------------
def ref_available():
    import refrun
    return refrun.ref_bin() is not None


class RefPredict:
    """the unmodified reference on a sample of the c4 workload: index built once (excluded from the step time, reported), every step =
    extractorfs + translatenucs + prefilter + align + resultspercontig + collectoptimalset on the sample contigs, all host threads"""

    def __init__(self, wl, prot, poff, threads):
        import refrun
        self.rr = refrun.RefRun(tempfile.mkdtemp(prefix="mb200_ref_", dir=os.environ.get("MB200_TMP")), threads)
        self.wl = wl; self.prot = prot; self.poff = poff
        self.index = self.rr.create_targets(prot, poff)
        self.n = 0

    def set_sample(self, n_contigs):
        nuc, off, _g = gen_planted_contigs(0, n_contigs, self.wl["contig_len"], self.prot, self.poff)
        self.rr.rm("contigs")
        self.rr.create_contigs(nuc, off)
        self.n = n_contigs

    def step(self):
        r = self.rr.predictexons()
        return r["wall_s"], r["fragments"], r

    def close(self):
        self.rr.cleanup()


def ref_predict_calibrate(rp, wl, target_s):
    """pick the sample so that one step takes about target_s on this box (a first small step measures the rate)"""
    rp.set_sample(min(wl["n_contigs"], 16))
    t, n, _ = rp.step()
    rate = n / t
    per_contig = n / rp.n
    want = int(min(wl["n_contigs"], max(16, rate * target_s / per_contig)))
    if want != rp.n:
        rp.set_sample(want)
    return want


def cpu_oracle_run(prot, poff, frag, foff, n_sample, align, threads):
    """time the CPU oracle (port of the reference algorithm) on the first n_sample fragments of the workload (used only when the reference binary is absent)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reflib
    o = reflib.Oracle()
    o.set_params()
    rc = o.L.mbo_set_targets(o.h, reflib._p(prot, C.c_uint8), reflib._p(prot, C.c_uint8), reflib._p(poff, C.c_uint64), C.c_uint32(len(poff) - 1))
    assert rc == 0
    n = min(n_sample, len(foff) - 1)
    q = np.ascontiguousarray(frag[:int(foff[n])]); qo = np.ascontiguousarray(foff[:n + 1])
    hits = (reflib.Hit * (n * 300))(); cnt = np.zeros(n, np.uint32); st = reflib.Stats()

    def one():
        t0 = time.perf_counter()
        rc = o.L.mbo_prefilter(o.h, reflib._p(q, C.c_uint8), reflib._p(qo, C.c_uint64), C.c_uint32(n), hits, reflib._p(cnt, C.c_uint32), C.byref(st), threads)
        assert rc == 0
        if align:
            hoff = np.zeros(n + 1, np.uint64); hoff[1:] = np.cumsum(cnt)
            ha = np.ctypeslib.as_array(hits)["seqId"].reshape(-1, 300)[:n]
            flat = np.ascontiguousarray(ha[np.arange(300)[None, :] < cnt[:n, None]], np.uint32)
            if len(flat) == 0:
                flat = np.zeros(1, np.uint32)
            alns = (reflib.Aln * max(1, int(hoff[-1])))(); cells = C.c_uint64(0)
            rc = o.L.mbo_align(o.h, reflib._p(q, C.c_uint8), reflib._p(qo, C.c_uint64), C.c_uint32(n), reflib._p(flat, C.c_uint32), reflib._p(hoff, C.c_uint64), alns, C.byref(cells), threads)
            assert rc == 0
        return time.perf_counter() - t0
    return o, one, n


# ---------------------------------------------------------------------------------------------------------------------
# canonical text of the fragment workloads' results (what tools/refrun.digest_entries hashes for the reference's DBs)
# ---------------------------------------------------------------------------------------------------------------------
def pref_text(hits, off):
    out = []
    for i in range(len(off) - 1):
        a, b = int(off[i]), int(off[i + 1])
        if a == b:
            continue
        h = hits[a:b]
        out.append(">%d\n" % i + "".join("%d\t%d\t%d\n" % t for t in zip(h["seq_id"].tolist(), h["score"].tolist(), h["diag"].tolist())))
    return "".join(out).encode()


def aln_text(alns, off, qlens, tlens):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from metaeuk_b200.modules import fmt_seqid
    out = []
    for i in range(len(off) - 1):
        a, b = int(off[i]), int(off[i + 1])
        if a == b:
            continue
        rows = [r for r in alns[a:b] if r["passed"] == 1]
        if not rows:
            continue
        rows.sort(key=lambda r: (r["evalue"], -int(r["bit_score"]), int(tlens[r["seq_id"]]), int(r["seq_id"])))
        out.append(">%d\n" % i + "".join("%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\n" % (r["seq_id"], r["bit_score"], fmt_seqid(r["seq_id_est"]), "%.3E" % r["evalue"], r["q_start"], r["q_end"],
                                                                                  qlens[i], r["db_start"], r["db_end"], tlens[r["seq_id"]]) for r in rows))
    return "".join(out).encode()


def text_digest(b):
    return {"sha256": hashlib.sha256(b).hexdigest(), "entries": b.count(b">"), "lines": b.count(b"\n") - b.count(b">")}


# ---------------------------------------------------------------------------------------------------------------------
# C4: full predictexons, contigs sharded by contig (strong scaling)
# ---------------------------------------------------------------------------------------------------------------------
def bench_predict(args, wl, base, rank, world, local, threads):
    import torch
    import torch.distributed as dist
    from metaeuk_b200.lib import Context
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    ctx.matrices_default()
    for k, v in (("sw_kernel", args.sw_kernel), ("sw_lane_max_hits", args.lane_max_hits), ("prefilter_config", args.prefilter_config), ("predict_batch_nt", args.batch_nt), ("predict_streams", args.streams)):
        if v:
            ctx.option_set(k, v)
    for kv in args.opt:
        ctx.option_set(kv.split("=")[0], int(kv.split("=")[1]))
    t0 = time.perf_counter()
    prot, poff = gen_proteins(wl["n_prot"])
    idx = {"gen_proteins_s": time.perf_counter() - t0}
    if rank == 0:
        idx.update(build_index(ctx, prot, poff))
    idx.update(replicate_index(ctx, rank, world))
    # this rank's contigs: the SAME contig set whatever N is; contigs have equal lengths, so equal counts = equal nucleotides (DBReader.cpp:1216-1256)
    nC = wl["n_contigs"]
    c0, c1 = nC * rank // world, nC * (rank + 1) // world
    t0 = time.perf_counter()
    nuc, off, genes = gen_planted_contigs(c0, c1, wl["contig_len"], prot, poff)
    idx["gen_contigs_s"] = time.perf_counter() - t0
    (nucP, _k1), (offP, _k2) = pinned(nuc), pinned(off)
    STAGES = ("ms_orf", "ms_prep", "ms_prefilter", "ms_sw_fwd", "ms_sw_rev", "ms_chain")
    W = max(3, args.warmup) if not args.allow_short_warmup else max(1, args.warmup)
    for _ in range(W):
        ctx.predict_exons(nucP, offP, copy=False)
    # fragment keys count over the whole contig DB (extractorfs numbers them in contig order): shift this shard's keys by the fragments before it
    nfrag_local = int(ctx.stats()["pe_fragments"])
    frag_counts = [nfrag_local]
    if world > 1:
        t = torch.zeros(world, dtype=torch.int64, device="cuda"); t[rank] = nfrag_local
        dist.all_reduce(t); frag_counts = [int(x) for x in t.tolist()]
    ctx.option_set("orf_key_base", sum(frag_counts[:rank]))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.stats()["launches"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    dev_ms = 0.0; pre_ms = 0.0; sw_ms = 0.0; devsum_ms = 0.0; sts = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call_off, calls, exons = ctx.predict_exons(nucP, offP, copy=False)
        st = ctx.stats(); sts.append(st)
        dev_ms += st["ms_dev_busy"]; pre_ms += st["ms_prefilter_sum"]; sw_ms += st["ms_sw_fwd_sum"] + st["ms_sw_rev_sum"]; devsum_ms += st["ms_dev_sum"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    st = sts[-1]
    vals = [dev_ms * 1e-3, wall]
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in t.tolist()]
    total_frag = sum(frag_counts)
    clocks = sampler.stop() if rank == 0 else None
    # ---- parity: digests of the timed step's output per block of DIGEST_BLOCK contigs, against the unmodified reference's dp_predictions DB
    gold = golden_digests("c4" if args.workload == "c4m" else args.workload)   # c4m = the first 1000 contigs of c4
    mine = {}
    for b0 in range(0, c1 - c0, DIGEST_BLOCK):
        b1 = min(c1 - c0, b0 + DIGEST_BLOCK)
        co = np.ascontiguousarray(call_off[b0:b1 + 1])
        txt = ctx.text_dp_db(co - co[0], calls[int(co[0]):int(co[-1])], exons, np.arange(c0 + b0, c0 + b1, dtype=np.uint32)) if co[-1] > co[0] else b""
        mine[str((c0 + b0) // DIGEST_BLOCK)] = text_digest(txt)
    allmine = [mine]
    stage_all = [{k: st[k] for k in STAGES + ("ms_join", "ms_predict")}]
    if world > 1:
        allmine = [None] * world; dist.all_gather_object(allmine, mine)
        stage_all = [None] * world; dist.all_gather_object(stage_all, {k: st[k] for k in STAGES + ("ms_join", "ms_predict")})
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return 0
    merged = {}
    for m in allmine:
        merged.update(m)
    parity = {"checked": False, "reason": "no digests of the reference for this workload in tests/golden/bench_digests.json"}
    if gold and gold.get("block") == DIGEST_BLOCK:
        common = sorted(set(gold["blocks"]) & set(merged), key=int)
        bad = [b for b in common if gold["blocks"][b]["sha256"] != merged[b]["sha256"]]
        parity = {"checked": bool(common), "sha256_match": not bad and bool(common), "blocks_checked": len(common), "blocks_total": len(merged), "contigs_per_block": DIGEST_BLOCK,
                  "lines": sum(merged[b]["lines"] for b in common), "mismatching_blocks": bad[:8],
                  "against": "dp_predictions of the unmodified reference (metaeuk predictexons module chain, --mask 1) on the same contigs and proteins"}
    peak, peak_src = measured_peak()
    abytes = algorithmic_bytes_prefilter(st); kms = pre_ms / args.steps
    per_kmer = ncu_traffic("dram_bytes_per_similar_kmer_c4") if args.workload in ("c4", "c4m") else None    # (measured on the 500 K-protein index only)
    nb = max(1, int(st["pe_batches"]))
    base["config"].update({"fragments": total_frag, "fragments_per_gpu": nfrag_local, "contigs": nC, "contigs_per_gpu": c1 - c0, "nucleotides_per_gpu": int(nuc.nbytes),
                           "planted_genes_rank0": genes, "stages": "extractorfs+translatenucs+prefilter+align+resultspercontig+collectoptimalset",
                           "contig_batches_per_gpu": int(st["pe_batches"]),
                           "parallelism": "contigs sharded by contig x%d (same contig set at every N), index replicated by one ncclBroadcast" % args.gpus,
                           "l2": "inputs exceed L2 (k-mer offsets 256 MB + index entries 1.3 GB + similarity table >> 126 MB); no explicit flush"})
    base["warmup"] = W
    base["scaling"] = "strong"
    base.update({"value": total_frag * args.steps / vals[0], "ms_per_step": 1000 * vals[0] / args.steps, "clocks": clocks, "gpu_launches": int(st["launches"] - l0),
                 "e2e": {"value": total_frag * args.steps / vals[1], "unit": "fragments/s", "h2d_bytes_per_step": int(nuc.nbytes + off.nbytes),
                         "d2h_bytes_per_step": int(calls.nbytes + exons.nbytes + call_off.nbytes), "ms_per_step": 1000 * vals[1] / args.steps,
                         "last_call_ms": {k: round(st[k], 3) for k in STAGES + ("ms_join", "ms_predict")}},
                 "roofline": {"kernel": "k_prefilter", "bound": "hbm", "achieved": abytes / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": abytes / (kms * 1e-3) / 1e9 / peak,
                              "peak_source": peak_src, "algorithmic_bytes_per_launch": int(abytes // nb), "kernel_ms": kms / nb, "launches_per_step": nb,
                              "share_of_step": pre_ms / devsum_ms if devsum_ms > 0 else None,
                              "traffic": int(per_kmer * st["kmers"] / nb) if per_kmer else None,
                              "traffic_frac": (per_kmer * st["kmers"] / (kms * 1e-3) / 1e9 / peak) if per_kmer and kms > 0 else None,
                              "traffic_note": "DRAM bytes per launch = %.1f B per similar k-mer (ncu, dram__sectors_read + dram__sectors_write of both k_prefilter launches of a C4/10 step; profiles/prefilter_traffic.json) x the similar k-mers of a launch of THIS run; traffic_frac = that / kernel time / peak" % (per_kmer or 0),
                              "kernel_ms_note": "CUDA-event durations of k_prefilter summed over the worker streams of the timed steps (with two streams the kernels of two contig batches share the GPU, so each launch is longer than it would be alone; single-stream figures: DESIGN.md section 8)",
                              "random_access": random_access_roofline(st, kms)},
                 "stages_ms": {k: st[k] for k in STAGES + ("ms_join",)}, "stages_ms_per_rank": stage_all,
                 "work": {k: int(st[k]) for k in ("kmers", "db_matches", "double_hits", "pref_results", "sw_pairs", "sw_cells_fwd", "sw_cells_rev", "pe_fragments", "pe_alignments", "pe_calls", "pe_exons")},
                 "index": idx, "parity": parity,
                 "timing": "value: fragments / device-busy time = length of the union, over the library's worker streams, of the CUDA-event intervals of every stage (ORFs, prep, prefilter, forward / reverse alignment, chaining; contigs of a batch resident) on one device timeline (mb200_stats.ms_dev_busy: no double counting of overlapped streams, host-only gaps excluded), max over ranks; e2e: wall clock of mb200_predict_exons (host contigs in, host call records out) between device synchronisations, max over ranks; stages_ms: per-stage event times of the busier worker stream (the streams share the GPU, so they do not add up to the busy time)"})
    if sw_ms > 0:
        base["gcups_sw"] = (st["sw_cells_fwd"] + st["sw_cells_rev"]) * args.steps / (sw_ms * 1e-3) / 1e9
        base["roofline_sw"] = {"kernel": "k_sw_stream forward + reverse", "bound": "integer ALU (DPX s16x2)", "achieved": base["gcups_sw"] / 1e3, "unit": "TCUPS",
                               "peak": 6.14, "peak_source": "DESIGN.md 4.2: SASS of k_sw_stream has 194 ALU-pipe instructions per column step of 32 rows x 2 targets, the ALU pipe issues one warp instruction per 2 cycles per scheduler: 148 SMs x 4 x 2048 cells / 388 cycles x 1.965 GHz",
                               "frac": base["gcups_sw"] / 1e3 / 6.14, "share_of_step": sw_ms / devsum_ms if devsum_ms > 0 else None,
                               "kernel_ms_note": "stage times summed over the worker streams, as for roofline"}
    base["predict_streams"] = int(st.get("pe_streams", 1))
    if args.phase_stats:     # one extra, untimed, instrumented pass for the per-phase breakdown of k_prefilter
        ctx.option_set("phase_stats", 1)
        ctx.predict_exons(nucP, offP, copy=False)
        ph = ctx.stats().get("pref_phase_cycles") or []
        ctx.option_set("phase_stats", 0)
        if sum(ph) > 0:
            names = ["clear", "P1 windows+scan", "P2 items+scan", "P3 kmers/offsets+scan", "P4 gather hits", "P5 split", "P6-7 double hits", "P8-9 diag scores",
                     "P10 best diag", "P11-12 threshold/sort", "P13-14 output", "", "", "", "", "between queries"]
            base["prefilter_phase_pct"] = {names[i]: round(100.0 * ph[i] / sum(ph), 1) for i in range(16) if ph[i]}
    if args.gpus == 1 and not args.no_cpu:
        base["cpu_baseline"] = cpu_baseline_predict(args, wl, prot, poff, threads)
    emit(base)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if (not parity.get("checked") or parity.get("sha256_match")) else 3


def cpu_baseline_predict(args, wl, prot, poff, threads):
    """the reference binary on a bounded sample (about 10-30 s of CPU work), all host threads"""
    if not ref_available():
        return {"value": None, "unit": "fragments/s", "cores": threads, "kind": "reference", "sample": "unavailable: oracle/_ref/metaeuk-ref not built"}
    rp = RefPredict(wl, prot, poff, threads)
    try:
        n = args.cpu_sample or ref_predict_calibrate(rp, wl, 15.0)
        if rp.n != n:
            rp.set_sample(n)
        t, nfr, r = rp.step()
        return {"value": nfr / t, "unit": "fragments/s", "cores": threads, "kind": "reference",
                "sample": "first %d contigs (%d fragments) of the workload through the unmodified reference binary (extractorfs, translatenucs, prefilter, align, resultspercontig, "
                          "collectoptimalset; --threads %d, AVX2 build); index built beforehand by `indexdb` and excluded" % (n, nfr, threads),
                "wall_s": t, "stages_s": {k: round(v["wall_s"], 3) for k, v in r["stages"].items()}, "index_build_s": rp.index["wall_s"]}
    finally:
        rp.close()


# ---------------------------------------------------------------------------------------------------------------------
# reduceredundancy / ORF stage lines (kept from round 1)
# ---------------------------------------------------------------------------------------------------------------------
def gen_calls(n_contigs, seed=99):
    """synthetic calledExons content: per contig 4..44 calls on both strands; exon keys drawn from a small per-contig pool so calls share exons"""
    from metaeuk_b200.lib import PRED_DTYPE
    rng = np.random.default_rng(seed)
    ncall = rng.integers(4, 45, n_contigs)
    off = np.zeros(n_contigs + 1, np.uint64); off[1:] = np.cumsum(ncall)
    n = int(off[-1])
    P = np.zeros(n, PRED_DTYPE)
    P["target_key"] = rng.integers(0, 500_000, n); P["strand"] = rng.choice(np.array([1, -1], np.int32), n)
    P["total_bit_score"] = rng.integers(40, 900, n); P["n_exons"] = rng.integers(1, 5, n)
    low = rng.integers(0, 48_000, n); P["low_coord"] = low; P["high_coord"] = low + rng.integers(200, 3000, n)
    P["evalue"] = np.float64(10.0) ** (-rng.integers(3, 120, n).astype(np.float64)) * rng.integers(1000, 9999, n) / 1000.0
    eb = np.zeros(n + 1, np.uint64); eb[1:] = np.cumsum(P["n_exons"]); P["exon_begin"] = eb[:-1]
    contig_of = np.repeat(np.arange(n_contigs), ncall)
    pool = np.maximum(2, ncall // 2)
    keys = (np.repeat(contig_of, P["n_exons"]).astype(np.uint64) * 64 + rng.integers(0, 1 << 30, int(eb[-1])) % np.repeat(pool[contig_of], P["n_exons"])).astype(np.uint32)
    return P, off, keys


def bench_reduce(args, wl, base, rank, world, local, threads):
    """the step after the path.  value = calls/s of k_reduce (CUDA events), e2e = the C-ABI call with host buffers."""
    import torch
    import torch.distributed as dist
    from metaeuk_b200.lib import Context
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    P, off, keys = gen_calls(wl["n_contigs"], seed=99 + rank)
    n = len(P)
    for _ in range(max(3, args.warmup)):
        ctx.reduce_redundancy(P, off, keys)
    l0 = ctx.stats()["launches"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    kms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out, mo, ro, cnt = ctx.reduce_redundancy(P, off, keys)
        kms.append(ctx.stats()["ms_reduce"])
    dt = time.perf_counter() - t0
    vals = [dt, float(np.sum(kms)) * 1e-3]
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in t.tolist()]
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak, peak_src = measured_peak()
    abytes = int(P.nbytes + keys.nbytes + out.nbytes + mo.nbytes + ro.nbytes)   # every call read once, every result written once
    kernel_ms = float(np.mean(kms))
    base.update({"metric": "reduceredundancy calls/sec", "unit": "calls/s", "value": world * n * args.steps / vals[1], "ms_per_step": 1000 * vals[1] / args.steps,
                 "gpu_launches": int(ctx.stats()["launches"] - l0), "clocks": None,
                 "e2e": {"value": world * n * args.steps / vals[0], "unit": "calls/s", "h2d_bytes_per_step": int(P.nbytes + off.nbytes + keys.nbytes),
                         "d2h_bytes_per_step": int(out.nbytes + mo.nbytes + ro.nbytes + cnt.nbytes), "ms_per_step": 1000 * vals[0] / args.steps},
                 "roofline": {"kernel": "k_reduce", "bound": "hbm", "achieved": abytes / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": abytes / (kernel_ms * 1e-3) / 1e9 / peak,
                              "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes, "kernel_ms": kernel_ms, "traffic": None},
                 "work": {"contigs": wl["n_contigs"], "calls": n, "representatives": int(cnt.sum())}})
    base["config"].update({"fragments_per_gpu": 0, "calls_per_gpu": n, "stages": "reduceredundancy"})
    if args.gpus == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import reflib
        L = C.CDLL(reflib.ORACLE_SO)
        o2 = np.zeros(n, out.dtype); m2 = np.zeros(n, np.uint32); r2 = np.zeros(n, np.uint32); c2 = np.zeros(wl["n_contigs"], np.uint32)
        t0 = time.perf_counter()
        L.mbo_reduce_redundancy_batch(C.c_void_p(P.ctypes.data), C.c_void_p(off.ctypes.data), C.c_uint32(wl["n_contigs"]), C.c_void_p(keys.ctypes.data), C.c_void_p(o2.ctypes.data),
                                      C.c_void_p(m2.ctypes.data), C.c_void_p(r2.ctypes.data), C.c_void_p(c2.ctypes.data), threads)
        tcpu = time.perf_counter() - t0
        same = bool((o2 == out).all() and (m2 == mo).all() and (c2 == cnt).all())
        base["cpu_baseline"] = {"value": n / tcpu, "unit": "calls/s", "cores": threads, "kind": "port", "sample": "the whole batch, all %d host threads (OpenMP over contigs)" % threads,
                                "identical_to_gpu_result": same}
    emit(base)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def bench_orf(args, wl, base, rank, world, local, threads):
    """the step before the path: contigs -> six-frame fragments -> residues.  value = fragments/s of the kernels with the contigs resident
    (CUDA events), e2e = the C-ABI call with host buffers (H2D of the contigs, kernels, D2H of records + residues)."""
    import torch
    import torch.distributed as dist
    from metaeuk_b200.lib import Context
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    nuc, off = gen_contigs(wl["n_contigs"], wl["contig_len"], seed=1234 + rank)
    (nucP, _k1), (offP, _k2) = pinned(nuc), pinned(off)
    W = max(3, args.warmup)
    for _ in range(W):
        ctx.extract_orfs(nucP, offP)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.stats()["launches"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    kms = []; out_bytes = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        recs, aa = ctx.extract_orfs(nucP, offP)
        kms.append(ctx.stats()["ms_orf"]); out_bytes = int(recs.nbytes + aa.nbytes)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = ctx.stats()
    vals = [dt, float(np.sum(kms)) * 1e-3]
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in t.tolist()]
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return 0
    nfrag = int(st["orf_fragments"])
    peak, peak_src = measured_peak()
    abytes = 2 * int(st["orf_nucleotides"]) + 2 * int(st["orf_residues"]) + 32 * nfrag   # both strands read once, residues + codes written, one record per fragment
    kernel_ms = float(np.mean(kms))
    base["config"].update({"fragments_per_gpu": nfrag, "contigs_per_gpu": wl["n_contigs"], "nucleotides_per_gpu": int(st["orf_nucleotides"]), "stages": "extractorfs+translatenucs",
                           "l2": "contigs (%d MB) exceed L2; no explicit flush" % (int(st["orf_nucleotides"]) >> 20)})
    base.update({"value": world * nfrag * args.steps / vals[1], "ms_per_step": 1000 * vals[1] / args.steps, "clocks": clocks, "gpu_launches": int(st["launches"] - l0),
                 "e2e": {"value": world * nfrag * args.steps / vals[0], "unit": "fragments/s", "h2d_bytes_per_step": int(nuc.nbytes + off.nbytes), "d2h_bytes_per_step": out_bytes,
                         "ms_per_step": 1000 * vals[0] / args.steps},
                 "roofline": {"kernel": "k_orf_scan x2 + k_orf_translate", "bound": "hbm", "achieved": abytes / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": abytes / (kernel_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(abytes), "kernel_ms": kernel_ms, "traffic": None},
                 "work": {"nucleotides": int(st["orf_nucleotides"]), "fragments": nfrag, "residues": int(st["orf_residues"])}})
    if args.gpus == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import reflib
        nsample = min(wl["n_contigs"], 16 * threads)
        sseq = nuc[:int(off[nsample])].tobytes(); soff = np.ascontiguousarray(off[:nsample + 1])
        reflib.oracle_extract_orfs_raw(sseq[:int(off[2])], np.ascontiguousarray(off[:3]))
        t0 = time.perf_counter()
        ncpu, _r, _a = reflib.oracle_extract_orfs_raw(sseq, soff, threads=threads)
        tcpu = time.perf_counter() - t0
        base["cpu_baseline"] = {"value": ncpu / tcpu, "unit": "fragments/s", "cores": threads, "kind": "port",
                                "sample": "first %d contigs of the workload, all %d host threads (OpenMP over contigs)" % (nsample, threads)}
    emit(base)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# C2 / C3: fragment workloads (prefilter-only, prefilter + align), fragment shards, weak scaling
# ---------------------------------------------------------------------------------------------------------------------
def bench_fragments(args, wl, base, rank, world, local, threads):
    import torch
    import torch.distributed as dist
    from metaeuk_b200.lib import Context
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    ctx.matrices_default()
    for k, v in (("sw_kernel", args.sw_kernel), ("sw_lane_max_hits", args.lane_max_hits), ("prefilter_config", args.prefilter_config)):
        if v:
            ctx.option_set(k, v)
    for kv in args.opt:
        ctx.option_set(kv.split("=")[0], int(kv.split("=")[1]))
    t0 = time.perf_counter()
    prot, poff = gen_proteins(wl["n_prot"])
    idx = {"gen_proteins_s": time.perf_counter() - t0}
    if rank == 0:
        idx.update(build_index(ctx, prot, poff))
    idx.update(replicate_index(ctx, rank, world))
    frag, foff = gen_fragments(wl["n_frag"], prot, poff, seed=4321 + rank)
    nq = len(foff) - 1
    (fragP, _k1), (foffP, _k2) = pinned(frag), pinned(foff)

    def step_resident():
        ctx.prefilter_resident()
        if wl["align"]:
            ctx.align_resident()

    e2e_out = {}

    def step_e2e():
        # the call a module makes: host buffers in (pinned), compact host result out
        if wl["align"]:
            res, off = ctx.search(fragP, foffP)
        else:
            res, off = ctx.prefilter_compact(fragP, foffP)
        e2e_out["bytes"] = int(res.nbytes + off.nbytes); e2e_out["res"] = (res, off)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt
    W = max(3, args.warmup)
    ctx.queries_upload(fragP, foffP)
    for _ in range(W):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.stats()["launches"]
    kernel_ms = []
    barrier()
    ctx.timer_start()                      # CUDA events on the library's stream, around exactly K steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
        kernel_ms.append(ctx.stats()["ms_prefilter"])
    dt = ctx.timer_stop() * 1e-3
    wall_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, wall_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, wall_dt = float(t[0].item()), float(t[1].item())
    launches = ctx.stats()["launches"] - l0
    st = ctx.stats()
    for _ in range(2):
        step_e2e()
    dt_e2e = timed(step_e2e, args.steps)
    st_e2e = ctx.stats()
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return 0
    # ---- parity of the timed e2e step's result (rank 0's fragments are the workload's seed) against the reference's pref / search_res DBs
    gold = golden_digests(args.workload)
    parity = {"checked": False, "reason": "no digests of the reference for this workload in tests/golden/bench_digests.json"}
    if gold:
        res, off = e2e_out["res"]
        qlens = (foff[1:] - foff[:-1]).astype(np.int64); tlens = (poff[1:] - poff[:-1]).astype(np.int64)
        d = text_digest(aln_text(res, off, qlens, tlens) if wl["align"] else pref_text(res, off))
        g = gold["aln" if wl["align"] else "pref"]
        parity = {"checked": True, "sha256_match": d["sha256"] == g["sha256"], "lines": d["lines"], "entries": d["entries"], "reference_lines": g["lines"],
                  "against": "%s DB of the unmodified reference (--mask 1) on the same fragments and proteins" % ("search_res (align)" if wl["align"] else "pref (prefilter)")}
    value = world * nq * args.steps / dt
    e2e = world * nq * args.steps / dt_e2e
    peak, peak_src = measured_peak()
    abytes = algorithmic_bytes_prefilter(st)
    kms = float(np.mean(kernel_ms))
    achieved = abytes / (kms * 1e-3) / 1e9
    base["config"].update({"fragments_per_gpu": wl["n_frag"], "parallelism": "fragment shards x%d (every rank its own fragments), index replicated by one ncclBroadcast" % args.gpus})
    base.update({"value": value, "ms_per_step": 1000 * dt / args.steps, "clocks": clocks, "gpu_launches": int(launches),
                 "e2e": {"value": e2e, "unit": "fragments/s", "h2d_bytes_per_step": int(frag.nbytes + foff.nbytes), "d2h_bytes_per_step": e2e_out.get("bytes", 0), "ms_per_step": 1000 * dt_e2e / args.steps,
                         "last_call_ms": {k: round(st_e2e[k], 3) for k in ("ms_h2d", "ms_prep", "ms_prefilter", "ms_sw_fwd", "ms_sw_rev", "ms_d2h", "ms_host")}},
                 "roofline": {"kernel": "k_prefilter", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                              "algorithmic_bytes_per_launch": int(abytes), "kernel_ms": kms, "share_of_step": kms / (1000 * dt / args.steps),
                              "traffic": ncu_traffic("dram_bytes_per_launch_" + args.workload), "random_access": random_access_roofline(st, kms)},
                 "stages_ms": {k: st[k] for k in ("ms_prep", "ms_prefilter", "ms_sw_fwd", "ms_sw_rev")},
                 "work": {k: int(st[k]) for k in ("kmers", "db_matches", "double_hits", "diag_cells", "pref_results", "sw_pairs", "sw_cells_fwd", "sw_cells_rev")},
                 "index": idx, "parity": parity,
                 "timing": "value: CUDA events on the library stream around K steps, max over ranks (wall clock for the same region: %.3f ms/step); e2e: wall clock between device synchronisations (host post-processing is part of the call)" % (1000 * wall_dt / args.steps)})
    ctx.option_set("phase_stats", 1)     # one extra, untimed, instrumented pass for the per-phase breakdown
    ctx.prefilter_resident()
    ph = ctx.stats().get("pref_phase_cycles") or []
    ctx.option_set("phase_stats", 0)
    if sum(ph) > 0:
        names = ["clear", "P1 windows+scan", "P2 items+scan", "P3 kmers/offsets+scan", "P4 gather hits", "P5 split", "P6-7 double hits", "P8-9 diag scores",
                 "P10 best diag", "P11-12 threshold/sort", "P13-14 output", "", "", "", "", "between queries"]
        base["prefilter_phase_pct"] = {names[i]: round(100.0 * ph[i] / sum(ph), 1) for i in range(16) if ph[i]}
    if wl["align"] and (st["ms_sw_fwd"] + st["ms_sw_rev"]) > 0:
        base["gcups_sw"] = (st["sw_cells_fwd"] + st["sw_cells_rev"]) / ((st["ms_sw_fwd"] + st["ms_sw_rev"]) * 1e-3) / 1e9
    if args.gpus == 1 and not args.no_cpu:
        base["cpu_baseline"] = cpu_baseline_fragments(args, wl, prot, poff, frag, foff, threads)
    emit(base)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if (not parity.get("checked") or parity.get("sha256_match")) else 3


def cpu_baseline_fragments(args, wl, prot, poff, frag, foff, threads):
    n = min(len(foff) - 1, args.cpu_sample or (100_000 if not wl["align"] else (100_000 if wl["n_prot"] <= 50_000 else 20_000)))
    if ref_available():
        import refrun
        rr = refrun.RefRun(tempfile.mkdtemp(prefix="mb200_ref_", dir=os.environ.get("MB200_TMP")), threads)
        try:
            index = rr.create_targets(prot, poff)
            r = rr.search(frag[:int(foff[n])], foff[:n + 1], align=wl["align"])
            return {"value": n / r["wall_s"], "unit": "fragments/s", "cores": threads, "kind": "reference",
                    "sample": "first %d fragments of the workload through the unmodified reference binary (prefilter%s; --threads %d, AVX2 build); index built beforehand by `indexdb` and excluded"
                              % (n, " + align" if wl["align"] else "", threads),
                    "wall_s": r["wall_s"], "stages_s": {k: round(v["wall_s"], 3) for k, v in r["stages"].items()}, "index_build_s": index["wall_s"]}
        finally:
            rr.cleanup()
    n = min(n, 8192 if not wl["align"] else 2048)
    o, one, n = cpu_oracle_run(prot, poff, frag, foff, n, wl["align"], threads)
    one()
    tcpu = one()
    return {"value": n / tcpu, "unit": "fragments/s", "cores": threads, "kind": "port",
            "sample": "first %d fragments of the workload, oracle port on all %d host threads (reference binary not built); index build excluded" % (n, threads)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores, rank 0 only
# ---------------------------------------------------------------------------------------------------------------------
def reference_arm(args, wl, base, threads):
    if wl.get("orf") or wl.get("reduce"):
        emit({"impl": "reference", "unavailable": "the reference arm is defined for the predictexons / prefilter / align workloads (c4, c4s, c3, c2)"})
        return 0
    prot, poff = gen_proteins(wl["n_prot"])
    if not ref_available():
        emit({"impl": "reference", "unavailable": "oracle/_ref/metaeuk-ref not built (run oracle/build_ref.sh where /root/reference exists)"})
        return 0
    W = max(0, args.warmup); K = max(1, args.steps)
    if wl.get("predict"):
        rp = RefPredict(wl, prot, poff, threads)
        try:
            # the whole --steps/--warmup run must end within a few minutes: size the sample for about (150 s / (K + W)) per step
            n = args.cpu_sample or ref_predict_calibrate(rp, wl, max(2.0, min(15.0, 150.0 / (K + W))))
            if rp.n != n:
                rp.set_sample(n)
            for _ in range(W):
                rp.step()
            t = 0.0; nfr = 0; last = None
            for _ in range(K):
                dt, nf, last = rp.step(); t += dt; nfr += nf
            sample = ("first %d contigs (%d fragments) of the workload per step through the unmodified reference binary (extractorfs, translatenucs, prefilter, align, "
                      "resultspercontig, collectoptimalset; --threads %d, AVX2 build); index built once by `indexdb` (%.1f s) and excluded" % (n, nfr // K, threads, rp.index["wall_s"]))
            stages = {k: round(v["wall_s"], 3) for k, v in last["stages"].items()}
        finally:
            rp.close()
    else:
        import refrun
        rr = refrun.RefRun(tempfile.mkdtemp(prefix="mb200_ref_", dir=os.environ.get("MB200_TMP")), threads)
        try:
            index = rr.create_targets(prot, poff)
            frag, foff = gen_fragments(wl["n_frag"], prot, poff, seed=4321)
            n = min(len(foff) - 1, args.cpu_sample or 100_000)
            r = rr.search(frag[:int(foff[n])], foff[:n + 1], align=wl["align"])     # calibration step (also warms the page cache)
            if not args.cpu_sample:
                per = max(2.0, min(15.0, 150.0 / (K + W)))
                n = int(min(len(foff) - 1, max(20_000, n * per / r["wall_s"])))
            for _ in range(W):
                rr.search(frag[:int(foff[n])], foff[:n + 1], align=wl["align"])
            t = 0.0; nfr = 0
            for _ in range(K):
                r = rr.search(frag[:int(foff[n])], foff[:n + 1], align=wl["align"]); t += r["wall_s"]; nfr += n
            sample = "first %d fragments of the workload per step through the unmodified reference binary (prefilter%s; --threads %d, AVX2 build); index built once by `indexdb` (%.1f s) and excluded" % (
                n, " + align" if wl["align"] else "", threads, index["wall_s"])
            stages = {k: round(v["wall_s"], 3) for k, v in r["stages"].items()}
        finally:
            rr.cleanup()
    v = nfr / t
    base.update({"impl": "reference", "value": v, "ms_per_step": 1000 * t / K, "gpu_launches": 0, "warmup": W, "stages_s": stages,
                 "cpu_baseline": {"value": v, "unit": "fragments/s", "cores": threads, "kind": "reference", "sample": sample},
                 "e2e": {"value": v, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(base)
    return 0


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="contigs (c4*) / fragments (c2, c3*) in the CPU-baseline sample (0 = sized for about 15 s)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--allow-short-warmup", action="store_true", help="profiling runs only: accept --warmup < 3")
    ap.add_argument("--sw-kernel", type=int, default=0, help="forward alignment kernel variant (diagnostic; 0 = library default)")
    ap.add_argument("--lane-max-hits", type=int, default=0, help="sw_lane_max_hits (diagnostic; 0 = library default)")
    ap.add_argument("--prefilter-config", type=int, default=0, help="prefilter_config (diagnostic; 0 = library default)")
    ap.add_argument("--streams", type=int, default=0, help="predict_streams (diagnostic; 0 = library default = 2 overlapped worker streams)")
    ap.add_argument("--phase-stats", action="store_true", help="add the per-phase cycle breakdown of k_prefilter (one extra untimed pass)")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="mb200_option_set(NAME, VALUE) before the run (diagnostic A/B runs; repeatable)")
    ap.add_argument("--batch-nt", type=int, default=0, help="predict_batch_nt (diagnostic; 0 = library default)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    base = {"metric": "predictexons query-fragments/sec", "unit": "fragments/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(3, args.warmup) if args.impl == "b200" else args.warmup,
            "higher_is_better": True, "scaling": "strong" if wl.get("predict") else "weak", "vs_baseline": None,
            "dtype": "u8/int16x2/int32 (integer DP; float32+float64 only for the composition bias and e-values)",
            "data": "synthetic (seeded; SURVEY.md 8d generator)", "config": {"workload": wl["desc"], "stages": "extractorfs+translatenucs+prefilter+align+resultspercontig+collectoptimalset" if wl.get("predict") else ("prefilter+align" if wl["align"] else "prefilter"),
                                                                   "proteins": wl["n_prot"]}}
    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, wl, base, threads)
    from metaeuk_b200 import lib as _mlib
    if _mlib.build_kind() != "sm_100a":
        raise SystemExit("bench.py measures the sm_100a library only (loaded: %s)" % _mlib.build_kind())
    if wl.get("orf"):
        return bench_orf(args, wl, base, rank, world, local, threads)
    if wl.get("reduce"):
        return bench_reduce(args, wl, base, rank, world, local, threads)
    if wl.get("predict"):
        return bench_predict(args, wl, base, rank, world, local, threads)
    return bench_fragments(args, wl, base, rank, world, local, threads)


if __name__ == "__main__":
    sys.exit(main())
