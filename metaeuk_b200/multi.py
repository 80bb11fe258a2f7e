"""Multi-GPU plumbing of the hot path: one process per GPU (torch.distributed), fragments sharded across ranks, the target
index built once on rank 0 and replicated with ONE broadcast per device buffer (NCCL over NVLink on GPUs; gloo in the CPU
tests).  The search itself has no data-path collective: every fragment is independent of every other fragment
(the reference gets the same independence from `#pragma omp for schedule(dynamic)` over queries, Prefiltering.cpp:790-870,
Alignment.cpp:312-405, and from MPI splits of the query DB, Prefiltering.cpp:560-640).

`ctx` is a metaeuk_b200.lib.Context (or anything with the same four index_* methods -- the gloo test drives this module with
host buffers).
"""
import numpy as np


def shard_bounds(off, world):
    """Contiguous shards of the query batch with (nearly) equal residue counts: rank r owns queries [b[r], b[r+1]).
    off = uint64 offsets[n+1] of the concatenated fragments."""
    off = np.asarray(off, dtype=np.uint64)
    n = len(off) - 1
    if n <= 0:
        return [0] * (world + 1)
    total = int(off[-1]) - int(off[0])
    b = [0]
    for r in range(1, world):
        want = int(off[0]) + (total * r) // world
        i = int(np.searchsorted(off, want, side="left"))
        b.append(min(max(i, b[-1]), n))
    b.append(n)
    return b


def device_view(ptr, nbytes):
    """uint8 torch tensor aliasing a device buffer owned by the library (no copy)"""
    import torch

    class _V:
        pass
    v = _V()
    v.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(v, device="cuda")


def replicate_index(ctx, rank, world, device="cuda", view=device_view):
    """rank 0 has built (or imported) the index; afterwards every rank holds a bit-identical copy.
    Returns the number of bytes broadcast."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return 0
    sz = torch.tensor(ctx.index_sizes() if rank == 0 else [0, 0, 0, 0], dtype=torch.int64, device=device)
    dist.broadcast(sz, 0)
    sizes = [int(x) for x in sz.tolist()]
    if rank != 0:
        ctx.index_alloc(sizes)
    total = 0
    for ptr, nb in ctx.index_buffers():
        if nb == 0:
            continue
        dist.broadcast(view(ptr, nb), 0)
        total += nb
    if rank != 0:
        ctx.index_commit()
    return total


def gather_to_root(obj, rank, world):
    """per-rank python results -> list on rank 0 (None elsewhere); results are small compared with the search itself"""
    import torch.distributed as dist
    if world == 1:
        return [obj]
    out = [None] * world if rank == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out


def search_sharded(run_shard, queries, rank, world):
    """queries: [(key, sequence)].  run_shard(list) -> {key: result text} on this rank's GPU.  Rank 0 returns the merged dict
    (identical to the single-GPU result: per-query results do not depend on the batch they travel in)."""
    off = np.zeros(len(queries) + 1, np.uint64)
    if queries:
        off[1:] = np.cumsum([len(s) for _, s in queries])
    b = shard_bounds(off, world)
    mine = run_shard(queries[b[rank]:b[rank + 1]])
    parts = gather_to_root(mine, rank, world)
    if rank != 0:
        return None
    merged = {}
    for p in parts:
        merged.update(p)
    return merged


def predict_sharded(predict, contig_off, rank, world):
    """predictexons over contigs sharded by nucleotide count (BASELINE configs[4]; the reference's unit of independence is the contig,
    collectoptimalset.cpp:274-418).  predict(c0, c1) runs contigs [c0, c1) on this rank's GPU (lib.Context.predict_exons on that slice) and
    returns (call_off, calls, exons, n_fragments).  Rank 0 returns the merged (call_off, calls, exons), identical to the single-GPU result:
    fragment keys count over the whole contig set in the reference, so every shard's orf_key is shifted by the number of fragments of the
    contigs before it, and exon_begin by the exons before it.  No data-path collective; the gather moves call records only."""
    b = shard_bounds(contig_off, world)
    call_off, calls, exons, nfrag = predict(b[rank], b[rank + 1])
    parts = gather_to_root((np.asarray(call_off, np.uint64), calls, exons, int(nfrag)), rank, world)
    if rank != 0:
        return None
    offs, cs, es = [np.zeros(1, np.uint64)], [], []
    frag_base = call_base = exon_base = 0
    for co, c, e, nf in parts:
        c = c.copy(); e = e.copy()
        c["exon_begin"] += np.uint32(exon_base); e["orf_key"] += np.uint32(frag_base)
        offs.append(co[1:] + np.uint64(call_base)); cs.append(c); es.append(e)
        frag_base += nf; call_base += len(c); exon_base += len(e)
    return np.concatenate(offs), np.concatenate(cs), np.concatenate(es)
