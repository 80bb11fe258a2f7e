"""Host-side mirror of the reference's module surface for the hot path, on in-memory DBs.

Each function corresponds to one reference module (same inputs, same result-DB records, byte for byte):
  prefilter          reference/lib/mmseqs/src/prefiltering/Main.cpp:13     -> pref DB      "tKey\\tscore\\tdiag"
  align              reference/lib/mmseqs/src/alignment/Main.cpp:12        -> alignment DB 10 columns
  resultspercontig   reference/src/exonpredictor/resultspercontig.cpp:34   -> 20 columns per contig
  collectoptimalset  reference/src/exonpredictor/collectoptimalset.cpp:224 -> 19 columns per exon
  predictexons       reference/data/predictexons.sh:42-87 (search -> resultspercontig -> collectoptimalset)
A "DB" here is {key: entry-text}; sequence DBs are lists of (key, sequence) in the reference's local-id order.
All arithmetic on scores/coordinates happens in libmetaeuk_b200 (GPU); this file only converts letters to codes,
orders records and prints them with the reference's formatting rules."""
import numpy as np

from .lib import CAND_DTYPE, Context, pack_codes

_ORDER = "ACDEFGHIKLMNPQRSTVWYX"
AA2NUM = np.full(256, 20, dtype=np.uint8)   # SubstitutionMatrix::setupLetterMapping (SubstitutionMatrix.cpp:262-304)
for _i, _c in enumerate(_ORDER):
    AA2NUM[ord(_c)] = _i
    AA2NUM[ord(_c.lower())] = _i
for _a, _b in (("J", "L"), ("Z", "E"), ("B", "D")):
    AA2NUM[ord(_a)] = AA2NUM[ord(_b)]
    AA2NUM[ord(_a.lower())] = AA2NUM[ord(_b)]


def encode(seq):
    return AA2NUM[np.frombuffer(seq.encode(), dtype=np.uint8)]


def encode_many(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    data = encode("".join(seqs)) if len(seqs) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(data), off


class TargetDB:
    """Target protein DB in local-id order.  `masked` = sequences after the reference's tantan masking (X at masked
    positions); pass None for --mask 0 semantics.  Masking itself is host preprocessing outside the hot path."""

    def __init__(self, entries, masked=None):
        self.keys = [k for k, _ in entries]
        self.seqs = [s for _, s in entries]
        self.masked = list(masked) if masked is not None else self.seqs
        self.lens = [len(s) for s in self.seqs]
        self.total_residues = sum(self.lens)

    @classmethod
    def masked_like_the_reference(cls, entries, mask_prob=0.9):
        """entries [(key, sequence)] -> TargetDB whose masked residues come from mb200_mask_targets (tantan, the reference's --mask 1 default)"""
        from .lib import mask_targets
        u, off = encode_many([s for _, s in entries])
        m, _ = mask_targets(u, off, mask_prob)
        masked = ["".join(_ORDER[c] for c in m[int(off[i]):int(off[i + 1])]) for i in range(len(entries))]
        return cls(entries, masked)

    def upload(self, ctx):
        u, off = encode_many(self.seqs)
        m, off2 = encode_many(self.masked)
        assert (off == off2).all()
        ctx.index_build(m, u, off)


def make_context(tdb, device=0, **params):
    ctx = Context(device)
    if params:
        ctx.set_params(**params)
    ctx.matrices_default()
    tdb.upload(ctx)
    return ctx


def kmer_threshold(sensitivity):
    """`-s` -> k-mer similarity threshold (Prefiltering::getKmerThreshold, sequence search, k = 6); -s 4 -> 127"""
    import ctypes as C
    from .lib import load_library
    L = load_library()
    L.mb200_kmer_threshold.argtypes = [C.c_float]
    return int(L.mb200_kmer_threshold(float(sensitivity)))


_FLAG_PARAMS = {"--max-seqs": ("max_res_list_len", int), "--min-ungapped-score": ("min_diag_score", int), "-e": ("eval_thr", float), "--comp-bias-corr": ("comp_bias_corr", int),
                "--reverse-fragments": ("reverse_fragments", int), "--translation-table": ("translation_table", int), "--metaeuk-eval": ("metaeuk_eval", float),
                "--metaeuk-tcov": ("tcov", float), "--max-intron": ("max_intron", int), "--min-intron": ("min_intron", int), "--min-exon-aa": ("min_exon_aa", int),
                "--max-overlap": ("max_overlap", int), "--max-exon-sets": ("max_exon_sets", int), "--set-gap-open": ("set_gap_open", int), "--set-gap-extend": ("set_gap_extend", int)}


def predictexons_flags(flags):
    """The flags of a `metaeuk predictexons` command line that change what this path computes -> (context parameters for make_context, ORF options for
    extractorfs / predictexons_contigs, mask 0|1).  Same rules as the C++ driver (mb200_cli.cpp:parseOpts): -s or --k-score give the k-mer threshold, the align
    stage's minimal alignment length follows --min-exon-aa (src/workflow/PredictExons.cpp:43-44); a flag this path does not implement raises."""
    w = flags.split() if isinstance(flags, str) else list(flags)
    if len(w) % 2:
        raise ValueError("flags come in (name, value) pairs: %r" % (flags,))
    par, orf, mask, sens, kscore = {}, {}, 1, 4.0, None
    for name, val in zip(w[0::2], w[1::2]):
        if name in _FLAG_PARAMS:
            par[_FLAG_PARAMS[name][0]] = _FLAG_PARAMS[name][1](val)
        elif name == "-s":
            sens = float(val)
        elif name == "--k-score":
            k = int(val.split("seq:")[1].split(",")[0]) if "seq:" in val else int(val)
            kscore = None if k == 2147483647 else k
        elif name == "--min-length":
            orf["min_len"] = int(val)
        elif name == "--max-length":
            orf["max_len"] = int(val)
        elif name == "--mask":
            mask = int(val)
        elif name in ("--threads", "-v", "--compressed", "--remove-tmp-files"):
            pass
        else:
            raise ValueError("flag %s is not implemented on the GPU path" % name)
    par["kmer_thr"] = kscore if kscore is not None else kmer_threshold(sens)
    par["min_aln_len"] = par.get("min_exon_aa", 11)
    return par, orf, mask


# ---- formatting (the reference's own rules) -----------------------------------------------------
def fmt_seqid(x):
    """Util::fastSeqIdToBuffer + Matcher::resultToBuffer's tab overwrite: 1.0 -> '1.00' (Util.cpp:222-251, Matcher.cpp:286-287)"""
    x = np.float32(x)
    if x == np.float32(1.0):
        return "1.00"
    s = "0."
    if x < 0.10:
        s += "0"
    if x < 0.01:
        s += "0"
    return s + str(int(np.float32(x * np.float32(1000))))


def fmt_exon_seqid(text):
    """PotentialExon::exonToBuffer after strtod -> float (PredictionParser.h:21,106-136)"""
    f = np.float32(float(text))
    if f == np.float32(1.0):
        return "1.000"
    s = "0."
    if f < 0.10:
        s += "0"
    if f < 0.01:
        s += "0"
    return s + str(int(np.float32(f * np.float32(1000))))


def prefilter(ctx, queries, tdb):
    """queries: list of (key, seq) -> {key: pref entry}  (QueryMatcher::prefilterHitToBuffer, QueryMatcher.h:118-130)"""
    q, qoff = encode_many([s for _, s in queries])
    hits, cnt = ctx.prefilter(q, qoff)
    out = {}
    for i, (k, _) in enumerate(queries):
        out[k] = "".join("%d\t%d\t%d\n" % (tdb.keys[h["seq_id"]], h["score"], h["diag"]) for h in hits[i, :cnt[i]])
    return out


def _aln_lines(alns, tdb, qlen):
    """accepted records sorted by Matcher::compareHits (Matcher.h:157-168), printed by Matcher::resultToBuffer (Matcher.cpp:280-327)"""
    rows = [a for a in alns if a["passed"] == 1]
    rows.sort(key=lambda a: (a["evalue"], -int(a["bit_score"]), tdb.lens[a["seq_id"]], tdb.keys[a["seq_id"]]))
    return "".join("%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\n" % (tdb.keys[a["seq_id"]], a["bit_score"], fmt_seqid(a["seq_id_est"]), "%.3E" % a["evalue"],
                                                                 a["q_start"], a["q_end"], qlen, a["db_start"], a["db_end"], tdb.lens[a["seq_id"]])
                   for a in rows)


def align(ctx, queries, tdb, pref):
    """pref: {key: pref entry} as written by prefilter -> {key: alignment entry}"""
    key2id = {k: i for i, k in enumerate(tdb.keys)}
    lists = [[key2id[int(l.split("\t")[0])] for l in pref.get(k, "").splitlines()] for k, _ in queries]
    q, qoff = encode_many([s for _, s in queries])
    hoff = np.zeros(len(queries) + 1, np.uint64)
    hoff[1:] = np.cumsum([len(x) for x in lists], dtype=np.uint64)
    flat = np.array([t for x in lists for t in x], dtype=np.uint32)
    alns = ctx.align(q, qoff, flat, hoff)
    return {k: _aln_lines(alns[int(hoff[i]):int(hoff[i + 1])], tdb, len(s)) for i, (k, s) in enumerate(queries)}


def search(ctx, queries, tdb):
    """prefilter + align without the text round trip (blastp.sh:70,85)"""
    q, qoff = encode_many([s for _, s in queries])
    alns, off = ctx.search(q, qoff)
    return {k: _aln_lines(alns[int(off[i]):int(off[i + 1])], tdb, len(s)) for i, (k, s) in enumerate(queries)}


def extractorfs(ctx, contigs, min_len=15, max_len=32734):
    """extractorfs + translatenucs with the flags of `metaeuk predictexons` (extractorfs.cpp:19-159, translatenucs.cpp:13-118):
    contigs [(key, nucleotides)] -> (queries [(orfKey, aa)], orf_headers {orfKey: 'contigKey\\tfrom+len[\\tflags]\\n'}) in the reference's key
    order.  The fragments stay on the device: ctx.queries_from_orfs() makes them the resident query batch."""
    seq = "".join(s for _, s in contigs).encode()
    off = np.zeros(len(contigs) + 1, np.uint64)
    if contigs:
        off[1:] = np.cumsum([len(s) for _, s in contigs], dtype=np.uint64)
    recs, aa = ctx.extract_orfs(seq, off, min_len, max_len, 1)
    aab = aa.tobytes()
    queries, headers = [], {}
    for i, r in enumerate(recs):
        frm, to = int(r["from"]), int(r["to"])
        h = "%d\t%d%s%d" % (contigs[int(r["contig"])][0], frm, "+" if frm < to else "-", abs(frm - to))   # Orf::writeOrfHeader (Orf.cpp:440-457)
        if int(r["flags"]) & 3:
            h += "\t%d" % (int(r["flags"]) & 3)
        headers[i] = h + "\n"
        queries.append((i, aab[int(r["aa_off"]):int(r["aa_off"]) + int(r["aa_len"])].decode()))
    return queries, headers


def predictexons_contigs(ctx, contigs, tdb, min_len=15, max_len=32734):
    """predictexons.sh:40-82 with the contigs as the only input and no text DB between the stages:
    contigs [(key, nucleotides)] -> {contigKey: dp_predictions entry} (19 columns per exon, Prediction::predictionToBuffer).
    Fragment keys are the ranks extractorfs would assign (the key order of nucl_6f / aa_6f)."""
    seq = "".join(s for _, s in contigs).encode()
    off = np.zeros(len(contigs) + 1, np.uint64)
    if contigs:
        off[1:] = np.cumsum([len(s) for _, s in contigs], dtype=np.uint64)
    ctx.target_keys_set(np.array(tdb.keys, np.uint32))
    call_off, calls, exons = ctx.predict_exons(seq, off, min_len, max_len)
    out = {}
    for ci, (k, _) in enumerate(contigs):
        s = ""
        for cl in calls[int(call_off[ci]):int(call_off[ci + 1])]:
            for e in exons[int(cl["exon_begin"]):int(cl["exon_begin"]) + int(cl["n_exons"])]:
                s += ctx.text_exon_line(cl, e)
        out[k] = s
    return out


def _reprint_exon(cols):
    """PotentialExon::setByDPRes -> exonToBuffer (PredictionParser.h:66-150): the identity is parsed to double, narrowed to float and
    re-truncated to three digits; the E-value goes through %.3E again; every other column is an integer"""
    f = np.float32(float(cols[2]))
    if f == np.float32(1.0):
        sid = "1.000"
    else:
        sid = "0." + ("0" if f < np.float32(0.10) else "") + ("0" if f < np.float32(0.01) else "") + str(int(np.float32(f * np.float32(1000))))
    return "\t".join([cols[0], cols[1], sid, "%.3E" % float(cols[3])] + list(cols[4:]))


def reduceredundancy(ctx, dp_by_contig, allow_overlaps=False):
    """reduceredundancy (src/exonpredictor/reduceredundancy.cpp:162-317): {contigKey: calledExons entry} ->
    ({contigKey: representative calls in DP format}, {contigKey: 'repT repS repLow memberT memberS memberLow' lines})"""
    from .lib import PRED_DTYPE
    keys = list(dp_by_contig.keys())
    preds, exon_keys, exon_lines, off = [], [], [], [0]
    for k in keys:
        prev = {1: None, -1: None}
        for line in dp_by_contig[k].split("\n"):
            if not line:
                continue
            w = line.split("\t")
            if len(w) != 19:
                raise ValueError("There should be 19 columns in the input file")
            t, s, low = int(w[0]), int(w[1]), int(w[5])
            j = prev[s]
            if j is None or preds[j][0] != t or preds[j][4] != low:      # a new call per strand when target or low coordinate changes (:237-262)
                preds.append([t, s, int(w[2]), 0, low, int(w[6]), len(exon_keys), 0, float(w[3])])
                exon_lines.append([])
                prev[s] = j = len(preds) - 1
            preds[j][3] += 1
            exon_keys.append(int(w[7])); exon_lines[j].append(w[7:])
        off.append(len(preds))
    P = np.zeros(len(preds), PRED_DTYPE)
    for i, p in enumerate(preds):
        P[i] = tuple(p)
    out, mo, ro, cnt = ctx.reduce_redundancy(P, np.array(off, np.uint64), np.array(exon_keys, np.uint32))
    grouped, clusters = {}, {}
    for c, k in enumerate(keys):
        b, e = off[c], off[c + 1]
        g = ""
        for r in ro[b:b + int(cnt[c])]:
            i = b + int(r); p = preds[i]
            if not allow_overlaps and int(out[i]["no_overlap_cluster_id"]) != p[0]:
                continue
            head = "%d\t%d\t%d\t%s\t%d\t%d\t%d" % (p[0], p[1], p[2], "%.3E" % p[8], p[3], p[4], p[5])      # Prediction::predictionToBuffer (PredictionParser.h:368-396)
            for ex in exon_lines[i]:
                g += head + "\t" + _reprint_exon(ex) + "\n"
        grouped[k] = g
        clusters[k] = "".join("%d\t%d\t%d\t%d\t%d\t%d\n" % (int(out[b + int(m)]["cluster_id"]), preds[b + int(m)][1], int(out[b + int(m)]["cluster_low_coord"]),
                                                              preds[b + int(m)][0], preds[b + int(m)][1], preds[b + int(m)][4]) for m in mo[b:e])
    return grouped, clusters


def parse_orf_header(h):
    """Orf::parseOrfHeader: 'contigKey\\tfrom+len' or 'from-len' (Orf.cpp:351-438)"""
    w = h.split()
    loc = w[1]
    if "+" in loc:
        a, b = loc.split("+")
        return int(w[0]), int(a), int(a) + int(b)
    a, b = loc.split("-")
    return int(w[0]), int(a), int(a) - int(b)


def resultspercontig(aln, orf_headers, contig_lens):
    """Join fragment->target records with the fragment's contig coordinates, regroup per contig, stable order by
    (target key, orf key).  Contigs without records still get an (empty) entry (resultspercontig.cpp:147-196)."""
    per = {k: [] for k in contig_lens}
    for orf, text in aln.items():
        if not text:
            continue
        contig, frm, to = parse_orf_header(orf_headers[orf])
        orf_len = max(frm, to) - min(frm, to) + 1
        for l in text.splitlines():
            w = l.split("\t")
            w[2] = fmt_seqid(float(w[2]))   # parse -> re-print loses up to one unit
            line = "\t".join(w) + "\t%d\t1\t1.00\t0.000E+00\t0\t%d\t%d\t%d\t%d\t%d\n" % (int(orf), orf_len - 1, orf_len, frm, to, contig_lens[contig])
            per[contig].append((int(w[0]), int(orf), line))
    return {c: "".join(r[2] for r in sorted(rows, key=lambda r: (r[0], r[1]))) for c, rows in per.items()}


def collectoptimalset(ctx, bycontig, total_target_residues):
    """{contig: 20-column lines} -> {contig: 19-column exon lines} (Prediction::predictionToBuffer, PredictionParser.h:368-396)"""
    contigs = list(bycontig.keys())
    texts = []
    rows = []
    off = [0]
    for c in contigs:
        for l in bycontig[c].splitlines():
            w = l.split("\t")
            rows.append((int(w[0]), int(w[1]), int(w[4]), int(w[5]), int(w[7]), int(w[8]), int(w[9]), int(w[10]), int(w[17]), int(w[18])))
            texts.append((w[2], w[3]))
        off.append(len(rows))
    cands = np.array(rows, dtype=CAND_DTYPE) if rows else np.zeros(0, CAND_DTYPE)
    calls, cnt, exons = ctx.chain(cands, np.array(off, np.uint64), total_target_residues)
    out = {}
    for ci, c in enumerate(contigs):
        b = off[ci]
        s = ""
        for cl in calls[2 * b: 2 * b + cnt[ci]]:
            head = "%d\t%d\t%d\t%s\t%d\t%d\t%d\t" % (cl["target_key"], cl["strand"], cl["total_bit_score"], "%.3E" % cl["evalue"], cl["n_exons"],
                                                     cl["low_coord"], cl["high_coord"])
            for e in exons[cl["exon_begin"]: cl["exon_begin"] + cl["n_exons"]]:
                k = cands[b + e["cand_idx"]]
                sid, ev = texts[b + e["cand_idx"]]
                s += head + "%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n" % (
                    k["orf_key"], k["bit_score"], fmt_exon_seqid(sid), "%.3E" % float(ev), k["t_start"], k["t_end"], k["t_len"],
                    e["contig_start"], e["contig_end"], e["nuc_len"], k["frag_from"], k["frag_to"])
        out[c] = s
    return out


def predictexons(ctx, fragments, orf_headers, contig_lens, tdb):
    """search -> resultspercontig -> collectoptimalset (predictexons.sh:68-82) on already extracted/translated fragments."""
    aln = search(ctx, fragments, tdb)
    byc = resultspercontig(aln, orf_headers, contig_lens)
    return aln, byc, collectoptimalset(ctx, byc, tdb.total_residues)
