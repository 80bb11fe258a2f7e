"""World-size-2 test of the multi-GPU plumbing (metaeuk_b200/multi.py) on CPU with the gloo backend: rank 0 owns the target DB,
the index buffers reach rank 1 only through replicate_index's broadcasts, each rank searches its own shard of the fragments and
rank 0's merged result must equal the reference's golden single-process result.  The compute on each rank is the CPU oracle (the
GPU library cannot run here); what is under test is the sharding, the replication protocol and the merge."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import reflib
from metaeuk_b200 import multi


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(5)
    for n, world in ((0, 2), (1, 4), (7, 8), (1000, 2), (1000, 3), (12345, 8)):
        lens = rng.integers(15, 300, n)
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        b = multi.shard_bounds(off, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == n
        assert all(b[i] <= b[i + 1] for i in range(world))
        if n >= 1000:
            res = [int(off[b[i + 1]] - off[b[i]]) for i in range(world)]
            assert max(res) - min(res) <= 2 * 300, res


class HostIndexCtx:
    """stands in for lib.Context: same index_* protocol, host buffers, oracle compute"""

    def __init__(self):
        self.bufs = []
        self.o = None

    def load(self, case):
        m, off = reflib.pack(case["masked_targets"])
        u, _ = reflib.pack([s for _, s in case["targets"]])
        self.bufs = [m, u, off.view(np.uint8)]
        self._commit()

    def index_sizes(self):
        return [0, len(self.bufs[2]) // 8 - 1, len(self.bufs[0]), 0]

    def index_alloc(self, sizes):
        self.bufs = [np.zeros(sizes[2], np.uint8), np.zeros(sizes[2], np.uint8), np.zeros(8 * (sizes[1] + 1), np.uint8)]

    def index_buffers(self):
        return [(i, b.nbytes) for i, b in enumerate(self.bufs)]

    def view(self, ptr, nbytes):
        import torch
        assert self.bufs[ptr].nbytes == nbytes
        return torch.from_numpy(self.bufs[ptr])

    def index_commit(self):
        self._commit()

    def _commit(self):
        self.o = reflib.Oracle()
        off = self.bufs[2].view(np.uint64)
        self.nT = len(off) - 1
        rc = self.o.L.mbo_set_targets(self.o.h, reflib._p(self.bufs[0], C.c_uint8), reflib._p(self.bufs[1], C.c_uint8), reflib._p(off, C.c_uint64), C.c_uint32(self.nT))
        assert rc == 0


def _worker(rank, world, port, case_name, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = reflib.load_case(case_name)
        ctx = HostIndexCtx()
        if rank == 0:
            ctx.load(c)
        nbytes = multi.replicate_index(ctx, rank, world, device="cpu", view=ctx.view)
        keys = [k for k, _ in c["targets"]]
        ctx.o.set_params(kmerThr=c["kmer_thr"], maxResListLen=c.get("max_seqs", 300))

        def run_shard(queries):
            res, _ = ctx.o.prefilter([s for _, s in queries], threads=1)
            return {str(k): reflib.fmt_pref(h, keys) for (k, _), h in zip(queries, res)}
        merged = multi.search_sharded(run_shard, [tuple(x) for x in c["queries"]], rank, world)
        if rank == 0:
            bad = [k for k, _ in c["queries"] if merged[str(k)] != c["pref"].get(str(k), "")]
            q.put(("ok" if not bad and nbytes > 0 and len(merged) == len(c["queries"]) else "mismatch %r" % bad[:5]))
    except Exception as e:   # surface the failure instead of hanging the parent
        q.put("rank %d: %r" % (rank, e))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("case_name", ["frag_small", "contigs_sets"])
def test_two_ranks_replicate_shard_and_merge(case_name):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, case_name, q)) for r in range(2)]
    for p in procs:
        p.start()
    msg = q.get(timeout=240)
    for p in procs:
        p.join(60)
    assert msg == "ok", msg
    assert all(p.exitcode == 0 for p in procs)


def _oracle_predict(c, o, contigs):
    """contigs [(key, nt)] -> (call_off, calls, exons, n_fragments) with the CPU oracle's stages (what lib.Context.predict_exons returns)"""
    from metaeuk_b200 import modules as M
    from metaeuk_b200.lib import CALL_DTYPE, PEXON_DTYPE
    hdr, seqs = reflib.oracle_extract_orfs(contigs)
    frags = [seqs[i] for i in range(len(seqs))]
    res, _ = o.prefilter(frags, threads=1)
    alns, _ = o.align(frags, [[t for t, _, _ in h] for h in res], threads=1)
    keys = [k for k, _ in c["targets"]]; tlens = [len(s) for _, s in c["targets"]]
    aln_txt = {i: "".join(reflib.fmt_aln(a, keys, len(frags[i]), tlens) for a in reflib.sort_alns([a for a in al if a["passed"] == 1], keys, tlens)) for i, al in enumerate(alns)}
    est = {(i, keys[a["seqId"]]): (a["seqIdEst"], a["evalue"]) for i, al in enumerate(alns) for a in al if a["passed"] == 1}
    byc = M.resultspercontig(aln_txt, hdr, {k: len(s) for k, s in contigs})
    call_off, calls, exons = [0], [], []
    for ck, _ in contigs:
        cands = []
        for l in byc[ck].splitlines():
            w = l.split("\t")
            cands.append(dict(targetKey=int(w[0]), bitScore=int(w[1]), qStart=int(w[4]), qEnd=int(w[5]), tStart=int(w[7]), tEnd=int(w[8]), tLen=int(w[9]), orfKey=int(w[10]),
                              fragFrom=int(w[17]), fragTo=int(w[18])))
        for cl in o.chain(cands, sum(tlens), max_sets=c["max_exon_sets"]):
            calls.append((cl["targetKey"], cl["strand"], cl["totalBitScore"], len(cl["exons"]), len(exons), cl["lowCoord"], cl["highCoord"], 0, cl["evalue"]))
            for e in cl["exons"]:
                k = cands[e["candIdx"]]; sid, ev = est[(k["orfKey"], k["targetKey"])]
                exons.append((k["orfKey"], k["bitScore"], sid, k["tStart"], k["tEnd"], k["tLen"], e["contigStart"], e["contigEnd"], e["nucLen"], k["fragFrom"], k["fragTo"], 0, ev))
        call_off.append(len(calls))
    return np.array(call_off, np.uint64), np.array(calls, CALL_DTYPE) if calls else np.zeros(0, CALL_DTYPE), np.array(exons, PEXON_DTYPE) if exons else np.zeros(0, PEXON_DTYPE), len(frags)


def _predict_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = reflib.load_case("contigs_sets")
        ctx = HostIndexCtx()
        if rank == 0:
            ctx.load(c)
        multi.replicate_index(ctx, rank, world, device="cpu", view=ctx.view)
        ctx.o.set_params(kmerThr=c["kmer_thr"])
        contigs = [tuple(x) for x in c["contigs"]]
        off = np.zeros(len(contigs) + 1, np.uint64); off[1:] = np.cumsum([len(s) for _, s in contigs])
        merged = multi.predict_sharded(lambda c0, c1: _oracle_predict(c, ctx.o, contigs[c0:c1]), off, rank, world)
        if rank == 0:
            call_off, calls, exons = merged
            emu = C.CDLL(os.path.join(reflib.ROOT, "tests", "emu", "libmb200_emu.so"))
            buf = C.create_string_buffer(1024); bad = []
            for ci, (ck, _) in enumerate(contigs):
                got = ""
                for cl in calls[int(call_off[ci]):int(call_off[ci + 1])]:
                    for e in exons[int(cl["exon_begin"]):int(cl["exon_begin"]) + int(cl["n_exons"])]:
                        n = emu.emu_text_exon_line(C.c_void_p(np.ascontiguousarray(cl).reshape(1).ctypes.data), C.c_void_p(np.ascontiguousarray(e).reshape(1).ctypes.data), buf)
                        got += buf.raw[:n].decode()
                if got != c["dp"][str(ck)]:
                    bad.append(ck)
            q.put("ok" if not bad and len(call_off) == len(contigs) + 1 and len(exons) > 0 else "mismatch %r" % bad[:5])
    except Exception as e:
        q.put("rank %d: %r" % (rank, e))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_shard_contigs_and_merge_calls():
    """BASELINE configs[4] protocol: contigs sharded by nucleotides, index replicated, every rank runs the whole contig -> calls chain on its
    shard (oracle compute here), rank 0 merges the call records with shifted fragment keys; printed, they are the reference's dp_predictions"""
    import subprocess
    import torch.multiprocessing as mp
    subprocess.check_call(["make", "-s", "-C", os.path.join(reflib.ROOT, "tests", "emu"), "CXX=" + ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")])
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_predict_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    msg = q.get(timeout=240)
    for p in procs:
        p.join(60)
    assert msg == "ok", msg
    assert all(p.exitcode == 0 for p in procs)
