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
@pytest.mark.parametrize("name", ["frag_small", "frag_family", "frag_nomask", "frag_adversarial", "contigs_small", "contigs_sets"])
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


@pytest.mark.parametrize("cfg", [1, 2, 3], ids=["dense", "sparse", "tiny-all-fallbacks"])
@pytest.mark.parametrize("variant", ["s6", "s1", "s7.5", "kscore", "maxseqs10", "minungapped", "nocomp", "nomask", "reverse", "table6"])
def test_prefilter_kernel_logic_under_non_default_flags(emu, variant, cfg):
    """the prefilter kernel logic under the flag sets of tests/golden/flags.json.gz that reach it (k-mer thresholds 96..154, list length 10, minimal ungapped
    score 30, composition bias off, unmasked targets, reversed fragments, genetic code 6 fragments) against the oracle's hit lists -- the oracle itself
    is pinned on the reference's search_res and calls of every variant (test_predictexons_flags_match_reference)"""
    c = reflib.load_case("flags")
    v = c["variants"][variant]
    par, _ = reflib.FLAG_VARIANTS[variant]
    par = dict(par); nomask = par.pop("nomask", False); comp = par.pop("comp_bias_corr", True)
    keys = [k for k, _ in c["targets"]]; unmasked = [s for _, s in c["targets"]]
    masked = unmasked if nomask else c["masked_targets"]
    qsrc = v if "queries" in v else c["variants"]["default"]
    queries = [(s[::-1] if v.get("reversed") else s) for _, s in qsrc["queries"]]
    o = reflib.Oracle()
    o.set_params(**par); o.set_comp_bias_corr(comp); o.set_targets(masked, unmasked)
    want, _ = o.prefilter(queries)
    o.close()
    m, toff = reflib.pack(masked); q, qoff = reflib.pack(queries)
    nQ = len(queries); maxres = par.get("maxResListLen", 300)
    hits = (PrefHit * (nQ * maxres))()
    cnt = np.zeros(nQ, np.uint32); st = np.zeros(nQ, np.uint32); stats = np.zeros(6, np.uint64)
    p = reflib._p
    emu.emu_set_comp_bias_corr(int(comp))
    try:
        rc = emu.emu_prefilter(p(m, C.c_uint8), p(toff, C.c_uint64), len(keys), p(q, C.c_uint8), p(qoff, C.c_uint64), nQ, par.get("kmerThr", 127), maxres,
                               par.get("minDiagScore", 15), 2, hits, p(cnt, C.c_uint32), p(st, C.c_uint32), p(stats, C.c_uint64), cfg)
        assert rc == 0
        got = [[(hits[i * maxres + j].seqId, hits[i * maxres + j].score, hits[i * maxres + j].diag) for j in range(cnt[i])] for i in range(nQ)]
        again = [i for i in range(nQ) if st[i] == 2]          # MB_ST_OVERFLOW with the first pass' capacities: the library's second pass (larger lists)
        assert (variant in ("s6", "s7.5")) == bool(again) and all(s in (0, 2) for s in st)
        if again:
            emu.emu_set_caps(1 << 22, 1 << 21, 1 << 21, 2 * 1000000 - 1)
            q2, qoff2 = reflib.pack([queries[i] for i in again])
            n2 = len(again); cnt2 = np.zeros(n2, np.uint32); st2 = np.zeros(n2, np.uint32)
            hits2 = (PrefHit * (n2 * maxres))()
            rc = emu.emu_prefilter(p(m, C.c_uint8), p(toff, C.c_uint64), len(keys), p(q2, C.c_uint8), p(qoff2, C.c_uint64), n2, par.get("kmerThr", 127), maxres,
                                   par.get("minDiagScore", 15), 2, hits2, p(cnt2, C.c_uint32), p(st2, C.c_uint32), p(stats, C.c_uint64), cfg)
            assert rc == 0 and (st2 == 0).all()
            for x, i in enumerate(again):
                got[i] = [(hits2[x * maxres + j].seqId, hits2[x * maxres + j].score, hits2[x * maxres + j].diag) for j in range(cnt2[x])]
    finally:
        emu.emu_set_comp_bias_corr(1)
        emu.emu_set_caps(1 << 17, 1 << 18, 1 << 19, 1 << 19)
    assert sum(len(w) for w in want) > (5 if variant == "reverse" else 50)
    for i in range(nQ):
        assert got[i] == [tuple(h) for h in want[i]], (variant, i, i in again)


@pytest.mark.parametrize("cfg", [1, 2], ids=["dense", "sparse"])
def test_very_long_queries_take_the_second_pass(emu, cfg):
    """fragments of 5 000 and 14 000 residues (a stop-free frame may reach 32 734 codons) overflow the first pass' per-CTA work lists at the default sensitivity;
    with the capacities of the library's second pass the kernel logic returns the oracle's hit lists"""
    import random
    rng = random.Random(5)
    AA = "ACDEFGHIKLMNPQRSTVWY"
    W = [7.8, 1.9, 5.4, 6.3, 3.9, 7.4, 2.3, 5.1, 5.7, 9.0, 2.2, 4.5, 5.2, 4.3, 5.1, 7.1, 5.8, 6.4, 1.3, 3.2]      # Robinson-Robinson background
    prots = ["".join(rng.choices(AA, weights=W, k=max(30, int(rng.gauss(330, 120))))) for _ in range(1500)]
    big = "".join(rng.choice(AA) if rng.random() < 0.15 else ch for ch in "".join(rng.sample(prots, 50)))
    queries = [big[:5000], big[:14000]]
    o = reflib.Oracle(); o.set_params(); o.set_targets(prots, prots)
    want, _ = o.prefilter(queries)
    o.close()
    m, toff = reflib.pack(prots); q, qoff = reflib.pack(queries)
    p = reflib._p
    res = {}
    try:
        for name, caps in (("first", (1 << 17, 1 << 18, 1 << 19, 1 << 19)), ("second", (1 << 22, 1 << 21, 1 << 21, 2 * 1000000 - 1))):
            emu.emu_set_caps(*caps)
            hits = (PrefHit * (2 * 300))(); cnt = np.zeros(2, np.uint32); st = np.zeros(2, np.uint32); stats = np.zeros(6, np.uint64)
            assert emu.emu_prefilter(p(m, C.c_uint8), p(toff, C.c_uint64), len(prots), p(q, C.c_uint8), p(qoff, C.c_uint64), 2, 127, 300, 15, 2,
                                     hits, p(cnt, C.c_uint32), p(st, C.c_uint32), p(stats, C.c_uint64), cfg) == 0
            res[name] = (list(st), [[(hits[i * 300 + j].seqId, hits[i * 300 + j].score, hits[i * 300 + j].diag) for j in range(cnt[i])] for i in range(2)])
    finally:
        emu.emu_set_caps(1 << 17, 1 << 18, 1 << 19, 1 << 19)
    assert res["first"][0] == [2, 2]                       # MB_ST_OVERFLOW, nothing returned
    assert res["second"][0] == [0, 0] and res["second"][1] == [[tuple(h) for h in w] for w in want] and min(len(w) for w in want) > 20


def test_orf_stop_codon_sets_of_all_genetic_codes(emu):
    """the ORF scan kernel's stop test (a 64-bit codon set per genetic code, orf_kernels.cuh) agrees with the reference's letter comparison for all
    256^3 byte triples under each of the 25 genetic codes the reference lists"""
    emu.emu_orf_stop_check.restype = C.c_longlong
    assert emu.emu_orf_stop_check() == 0


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


# ---- binary hand-off between the stages and the chaining kernel logic ------------------------------------
class _Orf(C.Structure):
    _fields_ = [("contig", C.c_uint32), ("frm", C.c_uint32), ("to", C.c_uint32), ("flags", C.c_uint32), ("aa_off", C.c_uint64), ("aa_len", C.c_uint32), ("strand_from", C.c_uint32)]


class _Aln(C.Structure):
    _fields_ = [("seq_id", C.c_uint32), ("raw_score", C.c_int32), ("bit_score", C.c_int32), ("q_start", C.c_int32), ("q_end", C.c_int32), ("db_start", C.c_int32),
                ("db_end", C.c_int32), ("aln_len", C.c_int32), ("passed", C.c_int32), ("seq_id_est", C.c_float), ("q_cov", C.c_float), ("db_cov", C.c_float), ("evalue", C.c_double)]


class _PExon(C.Structure):
    _fields_ = [("orf_key", C.c_uint32), ("bit_score", C.c_int32), ("seq_id_est", C.c_float), ("t_start", C.c_int32), ("t_end", C.c_int32), ("t_len", C.c_int32),
                ("contig_start", C.c_int32), ("contig_end", C.c_int32), ("nuc_len", C.c_int32), ("frag_from", C.c_int32), ("frag_to", C.c_int32), ("pad", C.c_int32),
                ("evalue", C.c_double)]


@pytest.mark.parametrize("key_stride", [1, 7], ids=["keys=ids", "keys!=ids"])
@pytest.mark.parametrize("name", ["contigs_small", "contigs_sets", "ref:minus_strand", "ref:multi_exon", "ref:two_contigs", "ref:target_overlap", "ref:cluster_rep", "ref:target_cov"])
def test_binary_handoff_and_chain_logic_match_reference_text(emu, name, key_stride):
    """accepted alignments (binary) + fragment records -> join_by_contig -> mb_chain_contig -> call/exon records must print to the
    reference's search_res_by_contig and dp_predictions entries byte for byte (the product's host_join.cpp + chain_core.cuh on the CPU).
    With key_stride 7 the target DB keys differ from the local ids and their order is reversed, so the (key, ORF) ordering is exercised.
    "ref:*" = the data sets of the reference's own test suite with run.sh's flags (--set-gap-extend, --metaeuk-tcov)."""
    if name.startswith("ref:"):
        c = reflib.load_case("ref_tests")["sets"][name[4:]]
        c["max_exon_sets"] = 1
        c["contig_lens"] = {str(k): len(s) for k, s in c["contigs"]}
    else:
        c = reflib.load_case(name)
    gap_extend, tcov = c.get("set_gap_extend", -1), c.get("tcov", 0.5)
    nT = len(c["targets"])
    keyof = lambda t: t if key_stride == 1 else (nT - 1 - t) * key_stride + 3      # noqa: E731  (local id -> DB key)
    rekey = {k: keyof(i) for i, (k, _) in enumerate(c["targets"])}
    local = {k: i for i, (k, _) in enumerate(c["targets"])}
    tkeys = np.array([keyof(i) for i in range(nT)], np.uint32)
    toff = np.zeros(nT + 1, np.uint64); toff[1:] = np.cumsum([len(s) for _, s in c["targets"]])
    contig_keys = [k for k, _ in c["contigs"]]
    cidx = {k: i for i, k in enumerate(contig_keys)}
    orf_keys = sorted(int(k) for k in c["orf_headers"])
    assert orf_keys == list(range(len(orf_keys)))
    n = len(orf_keys)
    orfs = (_Orf * n)(); qlen = {}
    for k, s in c["queries"]:
        qlen[int(k)] = len(s)
    aln_rows = []; aln_off = np.zeros(n + 1, np.uint64)
    for i in orf_keys:
        ck, frm, to = reflib.parse_orf_header(c["orf_headers"][str(i)])
        orfs[i].contig = cidx[ck]; orfs[i].frm = frm; orfs[i].to = to; orfs[i].aa_len = qlen[i]
        for line in c["aln"].get(str(i), "").splitlines():
            w = line.split("\t")
            assert int(w[6]) == qlen[i]
            aln_rows.append(w)
        aln_off[i + 1] = len(aln_rows)
    alns = (_Aln * max(1, len(aln_rows)))()
    for a, w in zip(alns, aln_rows):
        a.seq_id = local[int(w[0])]; a.bit_score = int(w[1]); a.q_start = int(w[4]); a.q_end = int(w[5]); a.db_start = int(w[7]); a.db_end = int(w[8]); a.passed = 1
        a.seq_id_est = 1.0 if w[2] == "1.00" else (int(w[2][2:]) + 0.5) / 1000.0      # any float the `align` stage prints as w[2]
        a.evalue = float(w[3])
        assert int(w[9]) == int(toff[a.seq_id + 1] - toff[a.seq_id])
    ncand = len(aln_rows)
    cands = (reflib.Cand * max(1, ncand))(); cand_off = np.zeros(len(contig_keys) + 1, np.uint64); cand_aln = np.zeros(max(1, ncand), np.uint64)
    p = reflib._p
    assert emu.emu_join(orfs, C.c_uint64(n), C.c_uint32(len(contig_keys)), C.c_uint64(0), p(aln_off, C.c_uint64), alns, p(tkeys, C.c_uint32), p(toff, C.c_uint64),
                        cands, p(cand_off, C.c_uint64), p(cand_aln, C.c_uint64)) == 0
    assert int(cand_off[-1]) == ncand

    def rekey_lines(text, cols, resort):
        """golden text with the target keys replaced; entries are ordered by (target key, ...) so re-sort stably by the new key"""
        rows = [l.split("\t") for l in text.splitlines()]
        for w in rows:
            for col in cols:
                w[col] = str(rekey[int(w[col])])
        if resort and key_stride != 1:
            rows.sort(key=lambda w: int(w[0]))
        return "".join("\t".join(w) + "\n" for w in rows)

    buf = C.create_string_buffer(1024)
    total_res = int(toff[-1])
    for ci, ck in enumerate(contig_keys):
        b, e = int(cand_off[ci]), int(cand_off[ci + 1])
        got = ""
        for k in range(b, e):
            nb = emu.emu_text_bycontig_line(C.byref(cands[k]), C.byref(alns[int(cand_aln[k])]), C.c_uint32(qlen[cands[k].orfKey]), C.c_uint64(c["contig_lens"][str(ck)]), buf)
            got += buf.raw[:nb].decode()
        assert got == rekey_lines(c["bycontig"][str(ck)], [0], True), ck
        calls = (reflib.Call * max(1, 2 * (e - b)))(); exons = (reflib.Exon * max(1, e - b))()
        sub = (reflib.Cand * max(1, e - b))(*[cands[k] for k in range(b, e)])
        emu.emu_chain.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_double, C.c_void_p, C.c_void_p]
        nc = emu.emu_chain(sub, e - b, 15, 10000, 11, 10, -1, gap_extend, c["max_exon_sets"], tcov, total_res, 0.001, calls, exons)
        assert nc >= 0
        got = ""
        for k in range(nc):
            cl = calls[k]
            for x in range(cl.nExons):
                ex = exons[cl.exonBegin + x]; kd = cands[b + ex.candIdx]; al = alns[int(cand_aln[b + ex.candIdx])]
                pe = _PExon(kd.orfKey, kd.bitScore, al.seq_id_est, kd.tStart, kd.tEnd, kd.tLen, ex.contigStart, ex.contigEnd, ex.nucLen, kd.fragFrom, kd.fragTo, 0, al.evalue)
                nb = emu.emu_text_exon_line(C.byref(cl), C.byref(pe), buf)
                got += buf.raw[:nb].decode()
        if key_stride == 1:
            assert got == c["dp"][str(ck)], ck
        else:   # the calls come in target-key order: compare as multisets of re-keyed lines and check the order of the keys
            assert sorted(got.splitlines()) == sorted(rekey_lines(c["dp"][str(ck)], [0], False).splitlines()), ck
            ks = [int(l.split("\t")[0]) for l in got.splitlines()]
            assert ks == sorted(ks)


# ---- persisted index DB ---------------------------------------------------------------------------------------
def _open_indexdb(emu, path):
    emu.emu_indexdb_open.restype = C.c_void_p
    emu.emu_indexdb_ptr.restype = C.c_void_p
    err = C.create_string_buffer(512)
    h = emu.emu_indexdb_open(os.fsencode(path), err, C.c_size_t(512))
    if not h:
        return None, err.value.decode()
    h = C.c_void_p(h)
    info = (C.c_uint64 * 3)(); meta = (C.c_int * 12)()
    emu.emu_indexdb_info(h, info, meta)
    nent, nseq, nres = (int(x) for x in info)

    def view(which, dtype, count):
        n = count * np.dtype(dtype).itemsize
        if n == 0:
            return np.zeros(0, dtype)
        return np.frombuffer((C.c_char * n).from_address(emu.emu_indexdb_ptr(h, which)), dtype=dtype).copy()
    d = dict(h=h, meta=list(meta), n_entries=nent, n_seq=nseq, offsets=view(0, np.uint64, reflib.TABLE + 1), entries6=view(1, np.uint8, 6 * nent), masked=view(2, np.uint8, nres),
             seq_off=view(3, np.uint64, nseq + 1), keys=view(4, np.uint32, nseq), unmasked=view(5, np.uint8, nres))
    emu.emu_indexdb_close(h)
    return d, ""


@pytest.mark.parametrize("name", ["frag_small", "contigs_sets"])
def test_persisted_index_db_round_trip_and_reference_digests(emu, name, tmp_path):
    """write the index in the reference's index-DB layout, read it back: ENTRIES / ENTRIESOFFSETS must be the bytes the reference's indexdb
    wrote for the same targets (golden digests), masked / unmasked residues, offsets and DB keys must survive, META must say what was built"""
    c = reflib.load_case(name)
    m, off = reflib.pack(c["masked_targets"]); u, _ = reflib.pack([s for _, s in c["targets"]])
    keys = np.array([3 * k + 11 for k, _ in c["targets"]], np.uint32)
    path = str(tmp_path / "targets.idx")
    p = reflib._p
    assert emu.emu_indexdb_write(os.fsencode(path), p(m, C.c_uint8), p(u, C.c_uint8), p(off, C.c_uint64), C.c_uint32(len(keys)), p(keys, C.c_uint32), c["kmer_thr"], c["mask"]) == 0
    assert open(path + ".dbtype", "rb").read() == (9).to_bytes(4, "little")
    d, err = _open_indexdb(emu, path)
    assert d is not None, err
    assert d["n_entries"] == c["idx_n_entries"] and d["n_seq"] == len(keys)
    assert hashlib.sha256(d["entries6"].tobytes()).hexdigest() == c["idx_entries_sha256"]
    assert hashlib.sha256(d["offsets"].tobytes()).hexdigest() == c["idx_offsets_u64_sha256"]
    assert (d["masked"] == m).all() and (d["unmasked"] == u).all() and (d["seq_off"] == off).all() and (d["keys"] == keys).all()
    assert d["meta"][1:7] == [6, 1, 21, 1 if c["mask"] else 0, 1, c["kmer_thr"]] and d["meta"][11] == 1
    # every entry starts on a page boundary and ends with the writer's NUL, as DBWriter::writeData + alignToPageSize leave it
    data = open(path, "rb").read()
    for line in open(path + ".index"):
        k, o, l = (int(x) for x in line.split("\t"))
        if k not in (7, 8):
            assert o % 4096 == 0 and data[o + l - 1] == 0, k
    # damaged / foreign files fail loudly
    with open(path + ".index") as f:
        lines = [l for l in f if not l.startswith("9\t")]
    open(path + ".index", "w").writelines(lines)
    d2, err = _open_indexdb(emu, path)
    assert d2 is None and "ENTRIES" in err
