#!/usr/bin/env python3
"""bench.py -- predictexons hot-path throughput (query fragments / s) on B200, with the CPU path timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c4s|orf|reduce] [--impl reference]

A "step" = one pass of the hot path over one batch of synthetic fragments per GPU.
Workload c2 (default, BASELINE.json configs[1]): 100 K fragments (len max(15,N(100,40)), half mutated protein
slices, half background) x 50 K proteins (len max(60,N(330,120))), prefilter-only.  c3 = 1 M x 500 K, prefilter + align.
Multi-GPU (torchrun, one rank per GPU): rank 0 builds the target index, ONE NCCL broadcast replicates it, every rank then
searches its own shard of fragments (weak scaling, no further collectives).

JSON keys: value = device-resident throughput (inputs already in HBM); e2e = same metric through the host-buffer C ABI call
(H2D of the fragments + kernels + D2H of the hit lists inside the timed region); roofline for the dominant kernel with
algorithmic bytes over the CUDA-event kernel time; cpu_baseline = the CPU oracle (port of the reference
algorithm) on a bounded sample on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1; the host side of the library (index build, alignment epilogue) is OpenMP code: give every rank its
# share of the cores (must happen before any OpenMP runtime is loaded)
if os.environ.get("OMP_NUM_THREADS") == "1" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BG = np.array([7.8, 1.9, 5.4, 6.3, 3.9, 7.4, 2.3, 5.1, 5.7, 9.0, 2.2, 4.5, 5.2, 4.3, 5.1, 7.1, 5.8, 6.4, 1.3, 3.2])
BG = BG / BG.sum()
WORKLOADS = {
    "c2": dict(n_prot=50_000, n_frag=100_000, align=False, desc="C2: 100K fragments (~100 aa) x 50K proteins (~330 aa), prefilter-only"),
    "c3": dict(n_prot=500_000, n_frag=1_000_000, align=True, desc="C3: 1M fragments x 500K proteins, prefilter + gapped SW align"),
    "c3s": dict(n_prot=500_000, n_frag=100_000, align=True, desc="C3/10: 100K fragments x 500K proteins, prefilter + gapped SW align (profiling size)"),
    "c4": dict(n_prot=500_000, n_contigs=10_000, contig_len=50_000, predict=True, n_frag=0, align=True,
               desc="C4: full predictexons (ORFs + prefilter + align + exon chaining) on 10K contigs x 50 kb with planted genes vs 500K proteins"),
    "c4s": dict(n_prot=50_000, n_contigs=500, contig_len=50_000, predict=True, n_frag=0, align=True,
                desc="C4/20: full predictexons on 500 contigs x 50 kb with planted genes vs 50K proteins (profiling size)"),
    "tiny": dict(n_prot=2_000, n_frag=4_000, align=True, desc="tiny: 4K fragments x 2K proteins (plumbing check)"),
    "reduce": dict(n_contigs=200_000, reduce=True, n_prot=0, n_frag=0, align=False,
                   desc="reduceredundancy stage (SURVEY 8f-2): 200 K contigs x ~24 calls (1-4 exons, exon keys shared inside a contig)"),
    "orf": dict(n_contigs=4_000, contig_len=50_000, orf=True, n_prot=0, n_frag=0, align=False,
                desc="ORF stage (SURVEY 8f-1): 4000 contigs x 50 kb random DNA -> six-frame fragments >= 15 codons, translated"),
}


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


def gen_planted_contigs(n, length, prot, poff, seed=1234):
    """SURVEY.md 8d contigs: i.i.d. ACGT; after 70 % of the inter-genic gaps (300-1500 nt) a gene = a random protein with per-residue
    substitution rate in {0.1, 0.3, 0.5}, back-translated, split into 1-4 exons by introns of 40-600 nt, 50 % reverse-complemented"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    nuc = acgt[rng.integers(0, 4, size=n * length, dtype=np.uint8)]
    nprot = len(poff) - 1
    genes = 0
    for c in range(n):
        base = c * length; pos = 0
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
                    pieces.append(acgt[rng.integers(0, 4, size=int(rng.integers(40, 600)), dtype=np.uint8)])
                pieces.append(_CODON[p[int(a):int(b)]].reshape(-1))
            g = np.concatenate(pieces)
            if pos + len(g) > length:
                break
            if rng.random() < 0.5:
                g = _COMP[g[::-1]]
            nuc[base + pos: base + pos + len(g)] = g
            pos += len(g); genes += 1
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(length))
    return np.ascontiguousarray(nuc), off, genes


def cpu_predict_run(prot, poff, nuc, off, n_contigs, threads):
    """the CPU oracle (port of the reference algorithm) on the first n_contigs contigs: ORF extraction + translation, prefilter, gapped alignment
    (the exon-chaining DP, microseconds per contig in the reference too, is left out of the CPU timing)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reflib
    o = reflib.Oracle()
    o.set_params()
    assert o.L.mbo_set_targets(o.h, reflib._p(prot, C.c_uint8), reflib._p(prot, C.c_uint8), reflib._p(poff, C.c_uint64), C.c_uint32(len(poff) - 1)) == 0
    sseq = nuc[:int(off[n_contigs])].tobytes(); soff = np.ascontiguousarray(off[:n_contigs + 1])
    from metaeuk_b200.modules import AA2NUM

    def one():
        t0 = time.perf_counter()
        n, recs, aa = reflib.oracle_extract_orfs_raw(sseq, soff, threads=threads)
        ra = np.ctypeslib.as_array(recs)[:n] if n else np.zeros(0, reflib.OrfRec)
        qo = np.zeros(n + 1, np.uint64); qo[:n] = ra["aaOff"]; qo[n] = (int(ra["aaOff"][-1]) + int(ra["aaLen"][-1])) if n else 0
        q = AA2NUM[np.frombuffer(aa, dtype=np.uint8, count=int(qo[n]))]
        hits = (reflib.Hit * (max(1, n) * 300))(); cnt = np.zeros(max(1, n), np.uint32); st = reflib.Stats()
        assert o.L.mbo_prefilter(o.h, reflib._p(q, C.c_uint8), reflib._p(qo, C.c_uint64), C.c_uint32(n), hits, reflib._p(cnt, C.c_uint32), C.byref(st), threads) == 0
        hoff = np.zeros(n + 1, np.uint64); hoff[1:] = np.cumsum(cnt[:n])
        ha = np.ctypeslib.as_array(hits)["seqId"].reshape(-1, 300)[:n]
        mask = np.arange(300)[None, :] < cnt[:n, None]
        flat = np.ascontiguousarray(ha[mask], np.uint32) if n else np.zeros(1, np.uint32)
        if len(flat) == 0:
            flat = np.zeros(1, np.uint32)
        alns = (reflib.Aln * max(1, int(hoff[-1])))(); cells = C.c_uint64(0)
        assert o.L.mbo_align(o.h, reflib._p(q, C.c_uint8), reflib._p(qo, C.c_uint64), C.c_uint32(n), reflib._p(flat, C.c_uint32), reflib._p(hoff, C.c_uint64), alns, C.byref(cells), threads) == 0
        return time.perf_counter() - t0, n
    return o, one


def bench_predict(args, wl, base, rank, world, local, threads):
    """BASELINE configs[3]/[4]: full predictexons from nucleotide contigs, contigs sharded across ranks (weak scaling: every rank gets
    n_contigs of its own), index replicated by one NCCL broadcast.  value = fragments / sum of the CUDA-event stage times (contigs resident per
    batch), e2e = fragments / wall clock of the C-ABI call mb200_predict_exons with HOST contigs in and HOST call records out."""
    import torch
    import torch.distributed as dist
    from metaeuk_b200.lib import Context
    from metaeuk_b200 import multi
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    ctx.matrices_default()
    if args.sw_kernel:
        ctx.option_set("sw_kernel", args.sw_kernel)
    if args.lane_max_hits:
        ctx.option_set("sw_lane_max_hits", args.lane_max_hits)
    t_idx0 = time.perf_counter()
    prot, poff = gen_proteins(wl["n_prot"])
    if rank == 0:
        ctx.index_build(prot, prot, poff)
    torch.cuda.synchronize()
    t_b0 = time.perf_counter()
    bcast_bytes = multi.replicate_index(ctx, rank, world)
    torch.cuda.synchronize()
    bcast_s = time.perf_counter() - t_b0 if world > 1 else 0.0
    index_s = time.perf_counter() - t_idx0
    nuc, off, genes = gen_planted_contigs(wl["n_contigs"], wl["contig_len"], prot, poff, seed=1234 + rank)
    (nucP, _k1), (offP, _k2) = pinned(nuc), pinned(off)
    STAGES = ("ms_orf", "ms_prep", "ms_prefilter", "ms_sw_fwd", "ms_sw_rev", "ms_chain")
    W = max(3, args.warmup) if wl["n_contigs"] <= 1000 else max(1, min(3, args.warmup))
    for _ in range(W):
        ctx.predict_exons(nucP, offP, copy=False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.stats()["launches"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = 0.0; pre_ms = 0.0; sts = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call_off, calls, exons = ctx.predict_exons(nucP, offP, copy=False)
        st = ctx.stats(); sts.append(st)
        dev_ms += sum(st[k] for k in STAGES); pre_ms += st["ms_prefilter"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    st = sts[-1]
    vals = [dev_ms * 1e-3, wall]
    total_frag = int(st["pe_fragments"])
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in t.tolist()]
        t = torch.tensor([total_frag], dtype=torch.int64, device="cuda")     # every rank has its own contigs (seed + rank): sum the fragments actually processed
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_frag = int(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return
    nfrag = int(st["pe_fragments"])
    peak, peak_src = measured_peak()
    abytes = algorithmic_bytes_prefilter(st); kms = pre_ms / args.steps
    base["config"].update({"fragments_per_gpu": nfrag, "contigs_per_gpu": wl["n_contigs"], "nucleotides_per_gpu": int(nuc.nbytes), "planted_genes_per_gpu": genes,
                           "stages": "extractorfs+translatenucs+prefilter+align+resultspercontig+collectoptimalset", "contig_batches": int(st["pe_batches"]),
                           "parallelism": "contig shards x%d, index replicated by one NCCL broadcast" % args.gpus,
                           "warmup_used": W})
    base["warmup"] = W
    base.update({"value": total_frag * args.steps / vals[0], "ms_per_step": 1000 * vals[0] / args.steps, "clocks": clocks, "gpu_launches": int(st["launches"] - l0),
                 "e2e": {"value": total_frag * args.steps / vals[1], "unit": "fragments/s", "h2d_bytes_per_step": int(nuc.nbytes + off.nbytes),
                         "d2h_bytes_per_step": int(calls.nbytes + exons.nbytes + call_off.nbytes), "ms_per_step": 1000 * vals[1] / args.steps,
                         "last_call_ms": {k: round(st[k], 3) for k in STAGES + ("ms_join", "ms_predict")}},
                 "roofline": {"kernel": "k_prefilter", "bound": "hbm", "achieved": abytes / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": abytes / (kms * 1e-3) / 1e9 / peak,
                              "peak_source": peak_src, "algorithmic_bytes_per_launch": int(abytes // max(1, int(st["pe_batches"]))), "kernel_ms": kms / max(1, int(st["pe_batches"])), "traffic": None},
                 "stages_ms": {k: st[k] for k in STAGES + ("ms_join",)},
                 "work": {k: int(st[k]) for k in ("kmers", "db_matches", "double_hits", "pref_results", "sw_pairs", "sw_cells_fwd", "sw_cells_rev", "pe_fragments", "pe_alignments", "pe_calls", "pe_exons")},
                 "index": {"build_plus_broadcast_s": index_s, "broadcast_s": bcast_s, "broadcast_bytes": int(bcast_bytes), "entries": ctx.index_sizes()[0]},
                 "timing": "value: sum of the CUDA-event stage times of the library stream (contigs of a batch resident), max over ranks; e2e: wall clock of mb200_predict_exons (host contigs in, host call records out) between device synchronisations"})
    if st["ms_sw_fwd"] + st["ms_sw_rev"] > 0:
        base["gcups_sw"] = (st["sw_cells_fwd"] + st["sw_cells_rev"]) / ((st["ms_sw_fwd"] + st["ms_sw_rev"]) * 1e-3) / 1e9
    if args.gpus == 1:
        nsample = min(wl["n_contigs"], args.cpu_sample or 8)
        o, one = cpu_predict_run(prot, poff, nuc, off, nsample, threads)
        tcpu, ncpu = one()
        base["cpu_baseline"] = {"value": ncpu / tcpu, "unit": "fragments/s", "cores": threads, "kind": "port",
                                "sample": "first %d contigs (%d fragments) of the workload: ORF extraction + prefilter + alignment on all %d host threads; index build and the chaining DP excluded" % (nsample, ncpu, threads)}
    emit(base)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


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
        return
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
        return
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


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "prefilter_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch_c2")
        except Exception:
            return None
    return None


def cpu_oracle_run(prot, poff, frag, foff, n_sample, align, threads):
    """time the CPU oracle (port of the reference algorithm) on the first n_sample fragments of the workload"""
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
            flat = np.zeros(max(1, int(hoff[-1])), np.uint32)
            for i in range(n):
                for k in range(cnt[i]):
                    flat[int(hoff[i]) + k] = hits[i * 300 + k].seqId
            alns = (reflib.Aln * max(1, int(hoff[-1])))(); cells = C.c_uint64(0)
            rc = o.L.mbo_align(o.h, reflib._p(q, C.c_uint8), reflib._p(qo, C.c_uint64), C.c_uint32(n), reflib._p(flat, C.c_uint32), reflib._p(hoff, C.c_uint64), alns, C.byref(cells), threads)
            assert rc == 0
        return time.perf_counter() - t0
    return o, one, n


def emit(obj):
    """the ONE JSON line on the real stdout (libraries that chat on fd 1, e.g. NCCL's version banner, were redirected to stderr)"""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="fragments in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--sw-kernel", type=int, default=0, help="forward alignment kernel variant (diagnostic; 0 = library default)")
    ap.add_argument("--lane-max-hits", type=int, default=0, help="sw_lane_max_hits (diagnostic; 0 = library default)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    base = {"metric": "predictexons query-fragments/sec", "unit": "fragments/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(3, args.warmup) if args.impl == "b200" else args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (integer DP; float32+float64 only for the composition bias)",
            "data": "synthetic (seeded; SURVEY.md 8d generator)", "config": {"workload": wl["desc"], "stages": "prefilter+align" if wl["align"] else "prefilter",
                                                                               "fragments_per_gpu": wl["n_frag"], "proteins": wl["n_prot"], "parallelism": "fragment shards x%d, index replicated by one NCCL broadcast" % args.gpus,
                                                                               "l2": "inputs exceed L2 (k-mer offsets 256 MB + index entries + similarity table >> 126 MB); no explicit flush"}}
    if args.impl == "b200":
        from metaeuk_b200 import lib as _mlib
        if _mlib.build_kind() != "sm_100a":
            raise SystemExit("bench.py measures the sm_100a library only (loaded: %s)" % _mlib.build_kind())
    # ------------------------------------------------------------------ reference arm: CPU path on the host cores
    if args.impl == "reference":
        if rank != 0:
            return
        if wl.get("orf"):
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import reflib
            nuc, off = gen_contigs(wl["n_contigs"], wl["contig_len"])
            nsample = min(wl["n_contigs"], 16 * threads)
            sseq = nuc[:int(off[nsample])].tobytes(); soff = np.ascontiguousarray(off[:nsample + 1])
            t0 = time.perf_counter(); nfr = 0
            for _ in range(max(1, args.steps)):
                nfr += reflib.oracle_extract_orfs_raw(sseq, soff, threads=threads)[0]
            t = time.perf_counter() - t0
            base.update({"impl": "reference", "value": nfr / t, "ms_per_step": 1000 * t / max(1, args.steps), "gpu_launches": 0,
                         "cpu_baseline": {"value": nfr / t, "unit": "fragments/s", "cores": threads, "kind": "port", "sample": "first %d contigs per step, all %d host threads (OpenMP over contigs)" % (nsample, threads)},
                         "e2e": {"value": nfr / t, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            emit(base)
            return
        prot, poff = gen_proteins(wl["n_prot"])
        if wl.get("predict"):
            nsample = min(wl["n_contigs"], args.cpu_sample or 8)
            nuc, off, _g = gen_planted_contigs(nsample, wl["contig_len"], prot, poff)
            o, one = cpu_predict_run(prot, poff, nuc, off, nsample, threads)
            for _ in range(args.warmup):
                one()
            t = 0.0; nfr = 0
            for _ in range(max(1, args.steps)):
                dt, n = one(); t += dt; nfr += n
            base["config"]["stages"] = "extractorfs+translatenucs+prefilter+align (chaining DP excluded from the CPU timing)"
            base.update({"impl": "reference", "value": nfr / t, "ms_per_step": 1000 * t / max(1, args.steps), "gpu_launches": 0,
                         "cpu_baseline": {"value": nfr / t, "unit": "fragments/s", "cores": threads, "kind": "port",
                                          "sample": "first %d contigs per step, all %d host threads; index build excluded" % (nsample, threads)},
                         "e2e": {"value": nfr / t, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            emit(base)
            return
        frag, foff = gen_fragments(wl["n_frag"], prot, poff, seed=4321)
        n_sample = args.cpu_sample or (8192 if not wl["align"] else 2048)
        o, one, n = cpu_oracle_run(prot, poff, frag, foff, n_sample, wl["align"], threads)
        for _ in range(args.warmup):
            one()
        t = sum(one() for _ in range(args.steps))
        v = n * args.steps / t
        base.update({"impl": "reference", "value": v, "ms_per_step": 1000 * t / args.steps, "gpu_launches": 0,
                     "cpu_baseline": {"value": v, "unit": "fragments/s", "cores": threads, "kind": "port",
                                      "sample": "first %d fragments of the workload per step, all %d host threads (OpenMP over queries); index build excluded" % (n, threads)},
                     "e2e": {"value": v, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(base)
        return
    # ------------------------------------------------------------------ B200 arm
    if wl.get("orf"):
        return bench_orf(args, wl, base, rank, world, local, threads)
    if wl.get("reduce"):
        return bench_reduce(args, wl, base, rank, world, local, threads)
    if wl.get("predict"):
        return bench_predict(args, wl, base, rank, world, local, threads)
    import torch
    import torch.distributed as dist
    from metaeuk_b200.lib import Context
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    ctx.matrices_default()
    if args.sw_kernel:
        ctx.option_set("sw_kernel", args.sw_kernel)
    if args.lane_max_hits:
        ctx.option_set("sw_lane_max_hits", args.lane_max_hits)
    t_idx0 = time.perf_counter()
    prot, poff = gen_proteins(wl["n_prot"])
    if rank == 0:
        ctx.index_build(prot, prot, poff)   # synthetic background proteins: masked == unmasked (masking is a host preprocessing input)
    from metaeuk_b200 import multi
    torch.cuda.synchronize()
    t_b0 = time.perf_counter()
    bcast_bytes = multi.replicate_index(ctx, rank, world)   # the single collective of the path: replicate the index over NVLink
    torch.cuda.synchronize()
    bcast_s = time.perf_counter() - t_b0 if world > 1 else 0.0
    index_s = time.perf_counter() - t_idx0
    frag, foff = gen_fragments(wl["n_frag"], prot, poff, seed=4321 + rank)
    nq = len(foff) - 1
    (fragP, _k1), (foffP, _k2) = pinned(frag), pinned(foff)
    maxres = ctx.params.max_res_list_len

    def step_resident():
        ctx.prefilter_resident()
        if wl["align"]:
            ctx.align_resident()

    from metaeuk_b200.lib import ALN_DTYPE, HIT_DTYPE
    e2e_out = {}

    def step_e2e():
        # the call a module makes: host buffers in (pinned), compact host result out
        if wl["align"]:
            res, off = ctx.search(fragP, foffP)
            e2e_out["bytes"] = int(res.nbytes + off.nbytes)
        else:
            res, off = ctx.prefilter_compact(fragP, foffP)
            e2e_out["bytes"] = int(res.nbytes + off.nbytes)

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
        return
    value = world * nq * args.steps / dt
    e2e = world * nq * args.steps / dt_e2e
    peak, peak_src = measured_peak()
    abytes = algorithmic_bytes_prefilter(st)
    kms = float(np.mean(kernel_ms))
    achieved = abytes / (kms * 1e-3) / 1e9
    h2d = int(frag.nbytes + foff.nbytes)
    d2h = e2e_out.get("bytes", 0)
    base.update({"value": value, "ms_per_step": 1000 * dt / args.steps, "clocks": clocks, "gpu_launches": int(launches),
                 "e2e": {"value": e2e, "unit": "fragments/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1000 * dt_e2e / args.steps,
                         "last_call_ms": {k: round(st_e2e[k], 3) for k in ("ms_h2d", "ms_prep", "ms_prefilter", "ms_sw_fwd", "ms_sw_rev", "ms_d2h", "ms_host")}},
                 "roofline": {"kernel": "k_prefilter", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                              "algorithmic_bytes_per_launch": int(abytes), "kernel_ms": kms, "traffic": ncu_traffic() if args.workload == "c2" else None},
                 "stages_ms": {k: st[k] for k in ("ms_prep", "ms_prefilter", "ms_sw_fwd", "ms_sw_rev")},
                 "work": {k: int(st[k]) for k in ("kmers", "db_matches", "double_hits", "diag_cells", "pref_results", "sw_pairs", "sw_cells_fwd", "sw_cells_rev")},
                 "index": {"build_plus_broadcast_s": index_s, "broadcast_s": bcast_s, "broadcast_bytes": int(bcast_bytes), "entries": ctx.index_sizes()[0]},
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
    # CPU baseline (oracle port) on a bounded sample of the same workload, rank 0, N=1 only
    if args.gpus == 1:
        n_sample = args.cpu_sample or (8192 if not wl["align"] else 2048)
        o, one, n = cpu_oracle_run(prot, poff, frag, foff, n_sample, wl["align"], threads)
        one()
        tcpu = one()
        base["cpu_baseline"] = {"value": n / tcpu, "unit": "fragments/s", "cores": threads, "kind": "port",
                                "sample": "first %d fragments of the workload, all %d host threads (OpenMP over queries); index build excluded" % (n, threads)}
    emit(base)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
