"""CPU-only check of the PRODUCT's kernel logic and host preprocessing: the CUDA source
(metaeuk_b200/csrc/prefilter_core.cuh) compiled in serial-emulation mode (tests/emu) must reproduce the
reference's golden outputs.  The real sm_100a kernels are checked on the GPU in test_gpu_parity.py."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

import reflib

EMU_DIR = os.path.join(reflib.ROOT, "tests", "emu")


class PrefHit(C.Structure):
    _fields_ = [("seqId", C.c_uint32), ("score", C.c_int32), ("diag", C.c_int32)]


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "CXX=" + ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")])
    return C.CDLL(os.path.join(EMU_DIR, "libmb200_emu.so"))


def test_host_matrices_and_sim3_match_reference(emu):
    o = reflib.Oracle(need_sim3=False)
    seed, ung, aln, sp, ap = o.matrices()
    s2 = np.zeros(441, np.int16); u2 = np.zeros(441, np.int16); a2 = np.zeros(441, np.int16); sp2 = np.zeros(21); ap2 = np.zeros(21)
    p = reflib._p
    emu.emu_matrices(p(s2, C.c_int16), p(u2, C.c_int16), p(a2, C.c_int16), p(sp2, C.c_double), p(ap2, C.c_double))
    assert (seed == s2).all() and (ung == u2).all() and (aln == a2).all()
    assert sp.tobytes() == sp2.tobytes() and ap.tobytes() == ap2.tobytes()
    n = 8000 * 8000
    sc = np.zeros(n, np.int16); ix = np.zeros(n, np.uint32)
    assert emu.emu_sim3(p(sc, C.c_int16), p(ix, C.c_uint32)) == 0
    c = reflib.load_case("frag_small")
    assert hashlib.sha256(sc.tobytes()).hexdigest() == c["sim3_score_i16_sha256"]
    assert hashlib.sha256(ix.tobytes()).hexdigest() == c["sim3_index_u32_sha256"]
    o.close()


@pytest.mark.parametrize("name", ["frag_small", "contigs_sets"])
def test_host_index_matches_reference_digest(emu, name):
    c = reflib.load_case(name)
    m, off = reflib.pack(c["masked_targets"])
    p = reflib._p
    n = C.c_uint64(0)
    offsets = np.zeros(reflib.TABLE + 1, np.uint64)
    entries = np.zeros(6 * c["idx_n_entries"], np.uint8)
    assert emu.emu_index(p(m, C.c_uint8), p(off, C.c_uint64), len(c["masked_targets"]), c["kmer_thr"], p(offsets, C.c_uint64), p(entries, C.c_uint8), C.byref(n)) == 0
    assert n.value == c["idx_n_entries"]
    assert hashlib.sha256(entries.tobytes()).hexdigest() == c["idx_entries_sha256"]
    assert hashlib.sha256(offsets.tobytes()).hexdigest() == c["idx_offsets_u64_sha256"]


@pytest.mark.parametrize("cfg", [1, 2, 3], ids=["dense", "sparse", "tiny-all-fallbacks"])
@pytest.mark.parametrize("name", ["frag_small", "frag_family", "frag_nomask", "contigs_small", "contigs_sets"])
def test_prefilter_kernel_logic_matches_reference(emu, name, cfg):
    """cfg 1/2 = the two kernel configurations of the library; 3 = a tiny configuration whose shared-memory homes overflow on every
    query, so the global-memory fallbacks of every list (and the sort-based double-hit path) are diffed against the reference too"""
    c = reflib.load_case(name)
    keys = [k for k, _ in c["targets"]]
    m, toff = reflib.pack(c["masked_targets"])
    queries = [s for _, s in c["queries"]]
    qkeys = [k for k, _ in c["queries"]]
    q, qoff = reflib.pack(queries)
    nQ = len(queries)
    maxres = c.get("max_seqs", 300)
    hits = (PrefHit * (nQ * maxres))()
    cnt = np.zeros(nQ, np.uint32); st = np.zeros(nQ, np.uint32); stats = np.zeros(6, np.uint64)
    p = reflib._p
    rc = emu.emu_prefilter(p(m, C.c_uint8), p(toff, C.c_uint64), len(keys), p(q, C.c_uint8), p(qoff, C.c_uint64), nQ, c["kmer_thr"], maxres, 15, 2,
                           hits, p(cnt, C.c_uint32), p(st, C.c_uint32), p(stats, C.c_uint64), cfg)
    assert rc == 0
    assert (st == 0).all()
    bad = 0
    for i, k in enumerate(qkeys):
        got = reflib.fmt_pref([(hits[i * maxres + j].seqId, hits[i * maxres + j].score, hits[i * maxres + j].diag) for j in range(cnt[i])], keys)
        want = c["pref"].get(str(k), "")
        if got != want:
            bad += 1
            if bad < 3:
                print("query", k, "\nwant:\n", want[:300], "\ngot:\n", got[:300])
    assert bad == 0


def test_reduceredundancy_kernel_logic_matches_reference(emu):
    """reduce_core.cuh (one contig per thread on the device) compiled for the host: clustering, overlap exclusion and both output orders"""
    c = reflib.load_case("redundancy_family")
    for k, entry in c["dp"].items():
        preds = reflib.parse_calls(entry)
        n = len(preds)
        P = (reflib.Pred * max(1, n))(); keys = []
        for i, p in enumerate(preds):
            P[i].targetKey = p["targetKey"]; P[i].strand = p["strand"]; P[i].totalBitscore = p["totalBitscore"]; P[i].nExons = len(p["exons"]); P[i].lowCoord = p["lowCoord"]
            P[i].highCoord = p["highCoord"]; P[i].exonBegin = len(keys); P[i].evalue = p["evalue"]
            keys += [int(e[0]) for e in p["exons"]]
        K = np.array(keys + [0], np.uint32)
        O = (reflib.PredOut * max(1, n))(); mo = np.zeros(max(1, n), np.uint32); ro = np.zeros(max(1, n), np.uint32); nrep = C.c_uint32(0)
        assert emu.emu_reduce(P, C.c_uint32(n), reflib._p(K, C.c_uint32), O, reflib._p(mo, C.c_uint32), reflib._p(ro, C.c_uint32), C.byref(nrep)) == 0
        outs = [{"clusterId": O[i].clusterId, "clusterLowCoord": O[i].clusterLowCoord, "noOverlapClusterId": O[i].noOverlapClusterId} for i in range(n)]
        assert reflib.fmt_grouped(preds, outs, [int(x) for x in ro[:nrep.value]]) == c["grouped"][k], k
        assert reflib.fmt_grouped(preds, outs, [int(x) for x in ro[:nrep.value]], allow_overlaps=True) == c["grouped_overlap_allowed"][k], k
        assert reflib.fmt_clusters(preds, outs, [int(x) for x in mo[:n]]) == c["clusters"][k], k
