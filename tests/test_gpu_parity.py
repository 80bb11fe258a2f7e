"""GPU parity tests (run on the B200 box): the sm_100a kernels, called through the C ABI, must reproduce
 (a) the reference's golden result DBs byte for byte, and
 (b) the CPU oracle on seeded synthetic inputs larger than the golden cases, plus size-independent properties."""
import random

import numpy as np
import pytest

import reflib
from metaeuk_b200 import modules as M

pytestmark = pytest.mark.gpu

FRAG = ["frag_small", "frag_family", "frag_nomask"]
CONTIG = ["contigs_small", "contigs_sets"]


def ctx_for(c, **kw):
    tdb = M.TargetDB(c["targets"], c["masked_targets"])
    ctx = M.make_context(tdb, kmer_thr=c["kmer_thr"], max_res_list_len=c.get("max_seqs", 300), **kw)
    return ctx, tdb


def strkeys(d):
    return {str(k): v for k, v in d.items()}


@pytest.mark.parametrize("cfg", [1, 2], ids=["dense", "sparse"])
@pytest.mark.parametrize("name", FRAG + CONTIG)
def test_prefilter_golden(name, cfg):
    """both kernel configurations (the library picks one from the index density; here each is forced)"""
    c = reflib.load_case(name)
    ctx, tdb = ctx_for(c)
    ctx.option_set("prefilter_config", cfg)
    got = strkeys(M.prefilter(ctx, c["queries"], tdb))
    bad = [k for k, _ in c["queries"] if got[str(k)] != c["pref"].get(str(k), "")]
    assert not bad, (bad[:5], got[str(bad[0])][:300], c["pref"].get(str(bad[0]), "")[:300])
    st = ctx.stats()
    assert st["pref_results"] == sum(v.count("\n") for v in c["pref"].values())
    ctx.close()


@pytest.mark.parametrize("hint", [0, 1, 3], ids=["cached", "evict-first", "evict-first-64B"])
@pytest.mark.parametrize("name", FRAG[:2] + CONTIG[:1])
def test_prefilter_index_read_policies(name, hint):
    """the cache policy of the index reads (mb200_option_set "prefilter_stream_hint": LDG.E.ENL2.256 / LDG.E.NA.EFL2.256 / ...LTC64B) must not change a result"""
    c = reflib.load_case(name)
    ctx, tdb = ctx_for(c)
    ctx.option_set("prefilter_stream_hint", hint)
    got = strkeys(M.prefilter(ctx, c["queries"], tdb))
    bad = [k for k, _ in c["queries"] if got[str(k)] != c["pref"].get(str(k), "")]
    assert not bad, (bad[:5], got[str(bad[0])][:300], c["pref"].get(str(bad[0]), "")[:300])
    ctx.close()


@pytest.mark.parametrize("sw", [1, 2, 4, (4, 1), (4, 1024), 5, (5, 1), 6, (6, "sw_ring", -1), (6, "sw_ring", 1), (6, "sw_ring", 16), (6, "sw_ring_fit", 0)],
                         ids=["rows", "rows2-packed", "lanes+rows2", "lanes(1 hit)+rows2", "lanes(any hits)+rows2", "lanes+wave+rows2", "lanes(1 hit)+wave+rows2",
                              "stream", "stream-private-profiles", "stream-ring1", "stream-ring16", "stream-ring-not-capped"])
@pytest.mark.parametrize("name", FRAG + CONTIG)
def test_align_golden(name, sw):
    """every forward alignment kernel variant (mb200_option_set "sw_kernel") must give the reference's result DB; 4 = queries of at most 64
    residues with at most sw_lane_max_hits hits take the one-item-per-lane kernel, the others the chunk kernel; 5 (default) = as 4, and the other
    queries of at most 512 residues take the wavefront kernel (row blocks of one alignment on neighbouring lanes); 6 (default) = streams of
    (query, two targets) pairs over persistent lane groups, end cells recorded by the forward pass, reverse pass only for the accepted ones"""
    c = reflib.load_case(name)
    ctx, tdb = ctx_for(c)
    if isinstance(sw, tuple):
        if len(sw) == 3:
            ctx.option_set(sw[1], sw[2])
        else:
            ctx.option_set("sw_lane_max_hits", sw[1])
        sw = sw[0]
    ctx.option_set("sw_kernel", sw)
    pref = {k: c["pref"].get(str(k), "") for k, _ in c["queries"]}
    got = strkeys(M.align(ctx, c["queries"], tdb, pref))
    bad = [k for k, _ in c["queries"] if got[str(k)] != c["aln"].get(str(k), "")]
    assert not bad, (bad[:5], got[str(bad[0])][:400], c["aln"].get(str(bad[0]), "")[:400])
    ctx.close()


@pytest.mark.parametrize("name", CONTIG)
def test_predictexons_golden(name):
    c = reflib.load_case(name)
    ctx, tdb = ctx_for(c, max_exon_sets=c["max_exon_sets"])
    contig_lens = {int(k): v for k, v in c["contig_lens"].items()}
    hdr = {int(k): v for k, v in c["orf_headers"].items()}
    aln, byc, dp = M.predictexons(ctx, c["queries"], hdr, contig_lens, tdb)
    assert strkeys(aln) == {k: v for k, v in c["aln"].items()}
    assert strkeys(byc) == c["bycontig"]
    assert strkeys(dp) == c["dp"]
    assert sum(v.count("\n") for v in dp.values()) > 0
    ctx.close()


def synth(seed, nprot, nfrag):
    rng = random.Random(seed)
    AA = "ACDEFGHIKLMNPQRSTVWY"
    W = [7.8, 1.9, 5.4, 6.3, 3.9, 7.4, 2.3, 5.1, 5.7, 9.0, 2.2, 4.5, 5.2, 4.3, 5.1, 7.1, 5.8, 6.4, 1.3, 3.2]
    prots = ["".join(rng.choices(AA, weights=W, k=max(60, int(rng.gauss(330, 120))))) for _ in range(nprot)]
    frags = []
    for i in range(nfrag):
        L = max(15, int(rng.gauss(100, 40)))
        if i % 2 == 0:
            p = rng.choice(prots)
            a = rng.randrange(0, max(1, len(p) - L))
            rate = rng.choice([0.1, 0.3, 0.5])
            s = "".join(rng.choice(AA) if rng.random() < rate else ch for ch in p[a:a + L])
        else:
            s = "".join(rng.choices(AA, weights=W, k=L))
        frags.append(s)
    return prots, frags


@pytest.mark.parametrize("cfg", [1, 2], ids=["dense", "sparse"])
def test_search_matches_oracle_on_larger_synthetic(cfg):
    """C2-shaped input scaled to what the scalar oracle finishes in seconds: 3000 proteins x 1500 fragments."""
    prots, frags = synth(2024, 3000, 1500)
    tdb = M.TargetDB(list(enumerate(prots)))
    ctx = M.make_context(tdb)
    ctx.option_set("prefilter_config", cfg)
    queries = list(enumerate(frags))
    got_pref = M.prefilter(ctx, queries, tdb)
    o = reflib.Oracle()
    o.set_params()
    o.set_targets(prots, prots)
    res, stats = o.prefilter(frags)
    keys = list(range(len(prots)))
    for (k, _), hits in zip(queries, res):
        assert got_pref[k] == reflib.fmt_pref(hits, keys), k
    st = ctx.stats()
    for a, b in (("kmers", "kmers"), ("db_matches", "dbMatches"), ("double_hits", "doubleHits"), ("diag_cells", "diagCells"), ("pref_results", "results")):
        assert st[a] == stats[b], (a, st[a], stats[b])
    got_aln = M.search(ctx, queries, tdb)
    alns, _ = o.align(frags, [[t for t, _, _ in h] for h in res])
    tlens = [len(p) for p in prots]
    for (k, q), al in zip(queries, alns):
        want = "".join(reflib.fmt_aln(a, keys, len(q), tlens) for a in reflib.sort_alns([a for a in al if a["passed"] == 1], keys, tlens))
        assert got_aln[k] == want, k
    o.close()
    ctx.close()


def test_prefilter_properties_at_scale():
    """Size-independent properties on an input too large for the scalar oracle: lists are sorted by (score desc, id asc),
    capped at max-seqs, scores >= min-ungapped-score, identical on a second run, and per-query results do not depend on
    which other queries share the batch."""
    prots, frags = synth(7, 20000, 20000)
    tdb = M.TargetDB(list(enumerate(prots)))
    ctx = M.make_context(tdb)
    q, qoff = M.encode_many(frags)
    hits, cnt = ctx.prefilter(q, qoff)
    assert cnt.max() <= 300 and cnt.sum() > 0
    for i in range(0, len(frags), 37):
        h = hits[i, :cnt[i]]
        assert (h["score"] >= 15).all()
        key = list(zip((-h["score"].astype(np.int64)).tolist(), h["seq_id"].tolist()))
        assert key == sorted(key)
        assert len(set(h["seq_id"].tolist())) == len(h)
    hits2, cnt2 = ctx.prefilter(q, qoff)
    assert (cnt == cnt2).all()
    for i in range(len(frags)):
        assert hits[i, :cnt[i]].tobytes() == hits2[i, :cnt[i]].tobytes()
    sub = list(range(0, len(frags), 11))
    q3, qoff3 = M.encode_many([frags[i] for i in sub])
    hits3, cnt3 = ctx.prefilter(q3, qoff3)
    for j, i in enumerate(sub):
        assert cnt3[j] == cnt[i] and hits3[j, :cnt3[j]].tobytes() == hits[i, :cnt[i]].tobytes()
    ctx.close()


def test_edge_cases():
    prots, frags = synth(3, 200, 10)
    tdb = M.TargetDB(list(enumerate(prots)))
    ctx = M.make_context(tdb)
    queries = [(0, "ACDEFGHIK"), (1, "ACDEFGHIKL"), (2, "XXXXXXXXXXXXXXXXXXXX"), (3, prots[5]), (4, prots[7][10:60] + "X" + prots[7][61:120])]
    pref = M.prefilter(ctx, queries, tdb)
    assert pref[0] == "" and pref[2] == ""
    assert pref[3].split("\t")[0] == "5"
    o = reflib.Oracle()
    o.set_params()
    o.set_targets(prots, prots)
    res, _ = o.prefilter([s for _, s in queries])
    for (k, _), hits in zip(queries, res):
        assert pref[k] == reflib.fmt_pref(hits, list(range(len(prots))))
    # empty batch
    h, c = ctx.prefilter(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(c) == 0
    o.close()
    ctx.close()


@pytest.mark.parametrize("name", ["contigs_small", "contigs_sets", "orf_edge"])
def test_extractorfs_golden(name):
    """six-frame ORF fragments + translation on the device == the reference's nucl_6f_h / aa_6f DBs, key for key"""
    c = reflib.load_case(name)
    from metaeuk_b200.lib import Context
    ctx = Context(0)
    queries, headers = M.extractorfs(ctx, [tuple(x) for x in c["contigs"]])
    assert {k: v for k, v in queries} == {int(k): s for k, s in c["queries"]}
    assert headers == {int(k): v for k, v in c["orf_headers"].items()}
    st = ctx.stats()
    assert st["orf_fragments"] == len(c["queries"]) and st["orf_residues"] == sum(len(s) for _, s in c["queries"])
    ctx.close()


def test_extractorfs_matches_oracle_on_larger_contigs():
    """contigs that span many tiles of the scan kernel (tile = 3072 nt), incl. lengths around the tile size and frames cut by the contig end"""
    rng = random.Random(17)
    contigs = []
    for i, L in enumerate([3071, 3072, 3073, 3074, 3075, 6143, 6144, 6149, 50000, 123457, 40, 44, 47, 48, 2]):
        s = "".join(rng.choices("ACGT", k=L))
        if i % 3 == 0:
            s = s[: L // 2] + "N" * min(30, L // 4) + s[L // 2 + min(30, L // 4):].lower()
        contigs.append((100 + i, s))
    from metaeuk_b200.lib import Context
    ctx = Context(0)
    queries, headers = M.extractorfs(ctx, contigs)
    hdr, seqs = reflib.oracle_extract_orfs(contigs)
    assert len(queries) == len(seqs) > 1000
    assert headers == hdr
    assert {k: v for k, v in queries} == seqs
    with pytest.raises(Exception):
        ctx.extract_orfs(b"ACGT" * 10, np.array([0, 40], np.uint64), 15, 32734, 0)   # other start modes are not implemented: loud
    ctx.close()


def test_contigs_to_exon_calls_without_host_round_trip():
    """contigs -> ORF fragments (device) -> resident prefilter + alignment -> exon chaining == the reference's dp_predictions"""
    c = reflib.load_case("contigs_sets")
    ctx, tdb = ctx_for(c, max_exon_sets=c["max_exon_sets"])
    contigs = [tuple(x) for x in c["contigs"]]
    queries, headers = M.extractorfs(ctx, contigs)
    ctx.queries_from_orfs()                      # fragments never leave the GPU
    ctx.prefilter_resident()
    hits, cnt = ctx.results_download()
    keys = tdb.keys
    for (k, _), h, n in zip(queries, hits, cnt):
        got = reflib.fmt_pref([(int(x["seq_id"]), int(x["score"]), int(x["diag"])) for x in h[:n]], keys)
        assert got == c["pref"].get(str(k), ""), k
    contig_lens = {k: len(s) for k, s in contigs}
    aln, byc, dp = M.predictexons(ctx, queries, headers, contig_lens, tdb)
    assert strkeys(dp) == c["dp"]
    ctx.close()


@pytest.mark.parametrize("batch_nt", [0, 9000, 30000], ids=["one-batch", "contig-per-batch", "three-per-batch"])
@pytest.mark.parametrize("name", CONTIG)
def test_predict_exons_binary_handoff_golden(name, batch_nt):
    """mb200_predict_exons: contigs in, call/exon records out, no text between the stages; printed with the reference's
    rules the records must be the reference's dp_predictions DB, whatever the batching"""
    c = reflib.load_case(name)
    ctx, tdb = ctx_for(c, max_exon_sets=c["max_exon_sets"])
    if batch_nt:
        ctx.option_set("predict_batch_nt", batch_nt)
    contigs = [tuple(x) for x in c["contigs"]]
    dp = M.predictexons_contigs(ctx, contigs, tdb)
    assert strkeys(dp) == c["dp"]
    st = ctx.stats()
    assert st["pe_contigs"] == len(contigs) and st["pe_fragments"] == len(c["queries"])
    assert st["pe_batches"] == {0: 1, 9000: len(contigs), 30000: (len(contigs) + 2) // 3}[batch_nt]
    assert st["pe_alignments"] == sum(v.count("\n") for v in c["aln"].values())
    assert st["pe_exons"] == sum(v.count("\n") for v in c["dp"].values())
    ctx.close()


@pytest.mark.parametrize("ds", ["minus_strand", "multi_exon", "two_contigs", "target_overlap", "cluster_rep", "target_cov"])
def test_reference_own_end_to_end_fixtures(ds):
    """the data sets of the reference's own test suite (reference/tests/run.sh; BASELINE configs[0] = two_contigs) with run.sh's flags, contigs in:
    the device pipeline must give the reference's dp_predictions and reduceredundancy DBs, and the calls must print to the CHECKED-IN known
    answers of the reference (header lines of as_should_final_*.fas: bit scores, E-values, exon counts and coordinates)"""
    c = reflib.load_case("ref_tests")["sets"][ds]
    ctx, tdb = ctx_for(c, set_gap_extend=c["set_gap_extend"], tcov=c["tcov"])
    contigs = [tuple(x) for x in c["contigs"]]
    dp = strkeys(M.predictexons_contigs(ctx, contigs, tdb))
    assert dp == c["dp"]
    assert sorted(reflib.fmt_fasta_headers(dp, c["target_acc"], c["contig_acc"])) == sorted(c["as_should"]["as_should_final_united_exons_aa.fas"])
    grouped, clusters = M.reduceredundancy(ctx, {int(k): v for k, v in dp.items()})
    assert strkeys(grouped) == c["grouped"] and strkeys(clusters) == c["clusters"]
    assert sorted(reflib.fmt_fasta_headers(strkeys(grouped), c["target_acc"], c["contig_acc"])) == sorted(c["as_should"]["as_should_final_grouped_predictions_rep.fas"])
    # the stage-by-stage route (text DBs between the stages) agrees too
    queries, headers = M.extractorfs(ctx, contigs)
    assert {str(k): v for k, v in headers.items()} == c["orf_headers"] and dict(queries) == {int(k): s for k, s in c["queries"]}   # (fixture lists are in data-file order)
    aln, byc, dp2 = M.predictexons(ctx, queries, headers, {k: len(s) for k, s in contigs}, tdb)
    assert {str(k): v for k, v in aln.items() if v} == {k: v for k, v in c["aln"].items() if v}
    assert strkeys(byc) == c["bycontig"] and strkeys(dp2) == c["dp"]
    ctx.close()


def test_targets_masked_by_the_library_give_the_reference_calls():
    """raw (unmasked) targets in: mb200_mask_targets (tantan, --mask 1) -> index -> contigs to calls == the reference's dp_predictions"""
    c = reflib.load_case("contigs_small")
    tdb = M.TargetDB.masked_like_the_reference(c["targets"])
    assert tdb.masked == c["masked_targets"]
    ctx = M.make_context(tdb, kmer_thr=c["kmer_thr"], max_exon_sets=c["max_exon_sets"])
    dp = M.predictexons_contigs(ctx, [tuple(x) for x in c["contigs"]], tdb)
    assert strkeys(dp) == c["dp"]
    ctx.close()


def planted_contigs(seed, prots, n_contigs, length):
    """SURVEY 8d contig generator at test size: random ACGT with planted genes = mutated proteins back-translated, split into 1-4 exons,
    introns of 40-600 nt, half of them reverse-complemented"""
    rng = random.Random(seed)
    codon = {"A": "GCT", "C": "TGT", "D": "GAT", "E": "GAA", "F": "TTT", "G": "GGT", "H": "CAT", "I": "ATT", "K": "AAA", "L": "CTG", "M": "ATG", "N": "AAT",
             "P": "CCT", "Q": "CAA", "R": "CGT", "S": "TCT", "T": "ACT", "V": "GTT", "W": "TGG", "Y": "TAT"}
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    AA = "ACDEFGHIKLMNPQRSTVWY"
    out = []
    for ci in range(n_contigs):
        parts, total = [], 0
        while total < length:
            gap = "".join(rng.choice("ACGT") for _ in range(rng.randrange(300, 1500)))
            parts.append(gap); total += len(gap)
            if rng.random() < 0.7:
                p = rng.choice(prots)
                rate = rng.choice([0.05, 0.2, 0.4])
                p = "".join(rng.choice(AA) if rng.random() < rate else ch for ch in p)
                nex = rng.randrange(1, 5)
                cuts = sorted(rng.sample(range(20, max(21, len(p) - 20)), nex - 1)) if len(p) > 60 and nex > 1 else []
                gene = ""
                for a, b in zip([0] + cuts, cuts + [len(p)]):
                    if gene:
                        gene += "GT" + "".join(rng.choice("ACGT") for _ in range(rng.randrange(40, 600))) + "AG"
                    gene += "".join(codon[ch] for ch in p[a:b])
                if rng.random() < 0.5:
                    gene = "".join(comp[ch] for ch in reversed(gene))
                parts.append(gene); total += len(gene)
        out.append((ci, "".join(parts)[:length]))
    return out


def test_predict_exons_matches_oracle_stage_by_stage_on_planted_genes():
    """contigs with planted multi-exon genes on both strands, larger than the golden cases: the oracle's stages (ORF extraction, prefilter,
    alignment, exon chaining) composed through the reference's text formats must equal the device pipeline's binary hand-off"""
    prots, _ = synth(99, 300, 0)
    contigs = planted_contigs(5, prots, 10, 30000)
    tdb = M.TargetDB(list(enumerate(prots)))
    ctx = M.make_context(tdb, max_exon_sets=2)
    ctx.option_set("predict_batch_nt", 100000)
    got = M.predictexons_contigs(ctx, contigs, tdb)
    o = reflib.Oracle(); o.set_params(); o.set_targets(prots, prots)
    want = reflib.oracle_pipeline(o, contigs, list(range(len(prots))), [len(p) for p in prots], max_sets=2)["dp"]
    assert sum(v.count("\n") for v in want.values()) > 20, "the planted genes must produce calls"
    assert any(int(l.split("\t")[4]) > 1 for v in want.values() for l in v.splitlines()), "multi-exon calls expected"
    assert got == want
    o.close(); ctx.close()


def test_cli_reduceredundancy_on_database_files(tmp_path):
    import os
    import subprocess
    from metaeuk_b200 import lib as mlib
    exe = os.environ.get("MB200_CLI_PATH", os.path.join(mlib.PKG, "bin", "mb200"))   # (tests/simt links the same driver against the SIMT emulation)
    c = reflib.load_case("redundancy_family")
    d = str(tmp_path)
    write_db(d + "/dp", [(int(k), v) for k, v in c["dp"].items()], 12)
    subprocess.check_call([exe, "gpureduceredundancy", d + "/dp", d + "/grouped", d + "/clusters"])
    assert read_db(d + "/grouped") == c["grouped"] and read_db(d + "/clusters") == c["clusters"]
    subprocess.check_call([exe, "gpureduceredundancy", d + "/dp", d + "/grouped_ov", d + "/clusters_ov", "--overlap", "1"])
    assert read_db(d + "/grouped_ov") == c["grouped_overlap_allowed"]


def test_reduceredundancy_golden():
    """the step after the path on the device: representatives (with and without same-strand overlaps) and the call -> representative DB"""
    c = reflib.load_case("redundancy_family")
    from metaeuk_b200.lib import Context
    ctx = Context(0)
    grouped, clusters = M.reduceredundancy(ctx, c["dp"])
    assert grouped == c["grouped"] and clusters == c["clusters"]
    grouped_ov, _ = M.reduceredundancy(ctx, c["dp"], allow_overlaps=True)
    assert grouped_ov == c["grouped_overlap_allowed"] and grouped_ov != grouped
    assert M.reduceredundancy(ctx, {}) == ({}, {})
    ctx.close()


def test_long_queries_take_the_32bit_alignment_path():
    """queries longer than 1024 residues skip the packed 16-bit kernel; longer than 512 windows spill the window arrays"""
    prots, _ = synth(11, 300, 1)
    rng = random.Random(5)
    AA = "ACDEFGHIKLMNPQRSTVWY"
    def mutate(s, r):
        return "".join(rng.choice(AA) if rng.random() < r else ch for ch in s)
    queries = [(0, mutate(prots[3] + prots[17] + prots[40] + prots[41], 0.2)), (1, mutate(prots[8], 0.3)), (2, mutate(prots[100] + prots[101] + prots[102] + prots[103] + prots[104], 0.1))]
    assert len(queries[0][1]) > 1024 and len(queries[2][1]) > 1024
    # every profile class of the wavefront kernel: 257..512 (one pair per half-warp), 513..1024 (one pair per warp), and block-edge lengths
    long = mutate(prots[20] + prots[21] + prots[22] + prots[23], 0.15)
    for n in (257, 300, 449, 512, 513, 640, 800, 992, 1000, 1024):
        if len(long) >= n:
            queries.append((len(queries), long[:n]))
    assert len(queries) >= 10
    tdb = M.TargetDB(list(enumerate(prots)))
    ctx = M.make_context(tdb)
    got_pref = M.prefilter(ctx, queries, tdb)
    got_aln = M.search(ctx, queries, tdb)
    o = reflib.Oracle()
    o.set_params()
    o.set_targets(prots, prots)
    res, _ = o.prefilter([s for _, s in queries])
    keys = list(range(len(prots))); tlens = [len(p) for p in prots]
    alns, _ = o.align([s for _, s in queries], [[t for t, _, _ in h] for h in res])
    for (k, q), hits, al in zip(queries, res, alns):
        assert got_pref[k] == reflib.fmt_pref(hits, keys), k
        want = "".join(reflib.fmt_aln(a, keys, len(q), tlens) for a in reflib.sort_alns([a for a in al if a["passed"] == 1], keys, tlens))
        assert got_aln[k] == want and want, k
    o.close()
    ctx.close()


def test_limits_fail_loudly():
    from metaeuk_b200.lib import Mb200Error
    prots, _ = synth(12, 50, 1)
    tdb = M.TargetDB(list(enumerate(prots)))
    ctx = M.make_context(tdb)
    with pytest.raises(Mb200Error):   # >= 32768 residues
        M.prefilter(ctx, [(0, "ACDEFGHIKL" * 3300)], tdb)
    ctx.close()
    ctx = M.make_context(tdb, cap_items_log2=6, cap_kmers_log2=8, cap_hits_log2=8)   # work lists far too small: error, never a silent drop
    with pytest.raises(Mb200Error):
        M.prefilter(ctx, [(0, prots[0])], tdb)
    # overflow policy "skip" (the default of mb200_predict_exons): the query is left without hits, counted and reported -- the other queries of the batch keep theirs
    ctx.option_set("overflow_policy", 1)
    short = prots[1][:12]                                       # 3 windows: fits even these lists
    q, qoff = M.encode_many([prots[0], short])
    hits, cnt = ctx.prefilter(q, qoff)
    assert cnt[0] == 0 and ctx.stats()["overflow_queries"] == 1
    assert ctx.query_status(2).tolist() == [2, 0]
    ctx.close()


def test_persisted_index_save_load_round_trip(tmp_path):
    """the resident index written in the reference's index-DB layout and loaded into a fresh context (mmap + upload, no rebuild)
    must give the reference's prefilter and exon calls; a DB written on the host alone (mb200_index_file_write) must load to the same index"""
    from metaeuk_b200 import lib as mlib
    c = reflib.load_case("contigs_sets")
    ctx, tdb = ctx_for(c, max_exon_sets=c["max_exon_sets"])
    ctx.target_keys_set(np.array(tdb.keys, np.uint32))
    path = str(tmp_path / "targets.idx")
    ctx.index_save(path)
    sizes = ctx.index_sizes()
    ctx.close()
    m, off = M.encode_many(tdb.masked); u, _ = M.encode_many(tdb.seqs)
    path2 = str(tmp_path / "targets_host.idx")
    mlib.index_file_write(path2, m, u, off, keys=np.array(tdb.keys, np.uint32), kmer_thr=c["kmer_thr"], mask_mode=1)
    for pth in (path, path2):
        ctx2 = mlib.Context(0)
        ctx2.set_params(kmer_thr=c["kmer_thr"], max_exon_sets=c["max_exon_sets"])
        ctx2.matrices_default()
        ctx2.index_load(pth)
        assert ctx2.index_sizes() == sizes
        assert list(ctx2.target_keys_get()) == tdb.keys
        got = strkeys(M.prefilter(ctx2, c["queries"][:800], tdb))
        assert all(got[str(k)] == c["pref"].get(str(k), "") for k, _ in c["queries"][:800])
        dp = M.predictexons_contigs(ctx2, [tuple(x) for x in c["contigs"]], tdb)
        assert strkeys(dp) == c["dp"]
        ctx2.close()
    # a context with another k-mer threshold must refuse the file
    ctx3 = mlib.Context(0)
    ctx3.set_params(kmer_thr=c["kmer_thr"] + 5)
    ctx3.matrices_default()
    with pytest.raises(Exception):
        ctx3.index_load(path)
    ctx3.close()


# ------------------------------------------------------------------------------------------------
# C++ host driver (metaeuk_b200/bin/mb200): the reference's module surface on MMseqs database files
# ------------------------------------------------------------------------------------------------
write_db, read_db = reflib.write_db, reflib.read_db


def test_cli_modules_on_database_files(tmp_path):
    import os
    import subprocess
    from metaeuk_b200 import lib as mlib
    exe = os.environ.get("MB200_CLI_PATH", os.path.join(mlib.PKG, "bin", "mb200"))   # (tests/simt links the same driver against the SIMT emulation)
    assert os.path.exists(exe), "build the CLI first (__graft_entry__.build)"
    c = reflib.load_case("contigs_sets")
    d = str(tmp_path)
    write_db(d + "/proteins", [(k, s + "\n") for k, s in c["targets"]], 0)
    write_db(d + "/proteins_masked", [(k, m + "\n") for (k, _), m in zip(c["targets"], c["masked_targets"])], 0)
    write_db(d + "/aa_6f", [(k, s + "\n") for k, s in c["queries"]], 0)
    write_db(d + "/nucl_6f_h", [(int(k), v) for k, v in c["orf_headers"].items()], 12)
    write_db(d + "/contigs", [(int(k), "A" * v + "\n") for k, v in c["contig_lens"].items()], 1)
    common = ["--masked-db", d + "/proteins_masked", "--k-score", str(c["kmer_thr"])]
    subprocess.check_call([exe, "gpuprefilter", d + "/aa_6f", d + "/proteins", d + "/pref"] + common)
    pref = read_db(d + "/pref")
    assert all(pref[str(k)] == c["pref"].get(str(k), "") for k, _ in c["queries"])
    subprocess.check_call([exe, "gpualign", d + "/aa_6f", d + "/proteins", d + "/pref", d + "/aln"] + common)
    aln = read_db(d + "/aln")
    assert all(aln[str(k)] == c["aln"].get(str(k), "") for k, _ in c["queries"])
    subprocess.check_call([exe, "gpuresultspercontig", d + "/contigs", d + "/nucl_6f", d + "/aln", d + "/bycontig"])
    assert read_db(d + "/bycontig") == c["bycontig"]
    subprocess.check_call([exe, "gpucollectoptimalset", d + "/bycontig", d + "/proteins", d + "/dp", "--max-exon-sets", str(c["max_exon_sets"])])
    assert read_db(d + "/dp") == c["dp"]
    subprocess.check_call([exe, "gpupredictexons", d + "/contigs", d + "/aa_6f", d + "/nucl_6f", d + "/proteins", d + "/dp2", "--max-exon-sets",
                           str(c["max_exon_sets"])] + common)
    assert read_db(d + "/dp2") == c["dp"]
    # the step before: contigs -> ORF header DB + translated fragment DB, then the whole chain from those
    write_db(d + "/contigs_nt", [(k, s + "\n") for k, s in c["contigs"]], 1)
    subprocess.check_call([exe, "gpuextractorfs", d + "/contigs_nt", d + "/orf_h", d + "/orf_aa"])
    assert read_db(d + "/orf_h") == {str(k): v for k, v in c["orf_headers"].items()}
    assert read_db(d + "/orf_aa") == {str(k): s + "\n" for k, s in c["queries"]}
    # ... and the whole chain with the contigs as the only input (fragments never leave the GPU between extraction and search)
    subprocess.check_call([exe, "gpupredictexons", d + "/contigs_nt", d + "/proteins", d + "/dp3", "--max-exon-sets", str(c["max_exon_sets"])] + common)
    assert read_db(d + "/dp3") == c["dp"]
    # same, in batches of about 20 kb of contigs, and with the reference's text DBs between the stages instead of the binary records
    subprocess.check_call([exe, "gpupredictexons", d + "/contigs_nt", d + "/proteins", d + "/dp4", "--max-exon-sets", str(c["max_exon_sets"]), "--batch-nt", "20000"] + common)
    assert read_db(d + "/dp4") == c["dp"]
    # persisted index: built on the host (no GPU), then every search module takes it instead of rebuilding
    subprocess.check_call([exe, "gpucreateindex", d + "/proteins", d + "/proteins.idx"] + common)
    subprocess.check_call([exe, "gpuprefilter", d + "/aa_6f", d + "/proteins", d + "/pref_idx", "--index-db", d + "/proteins.idx"] + common)
    assert read_db(d + "/pref_idx") == pref
    subprocess.check_call([exe, "gpupredictexons", d + "/contigs_nt", d + "/proteins", d + "/dp6", "--max-exon-sets", str(c["max_exon_sets"]), "--index-db", d + "/proteins.idx"] + common)
    assert read_db(d + "/dp6") == c["dp"]
    subprocess.check_call([exe, "gpupredictexons", d + "/contigs_nt", d + "/proteins", d + "/dp5", "--max-exon-sets", str(c["max_exon_sets"]), "--text-stages", "1"] + common)
    assert read_db(d + "/dp5") == c["dp"]
