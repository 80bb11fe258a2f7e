"""Test-side helpers: golden fixture loading, residue encoding, the ctypes binding of the CPU oracle
(oracle/libmb200_oracle.so) and the reference's text formatters restated in Python so structured
results can be compared byte-for-byte with the reference's result DBs.

TEST INFRASTRUCTURE ONLY (the product never imports this)."""
import ctypes as C
import gzip
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_SO = os.path.join(ROOT, "oracle", "libmb200_oracle.so")
TABLE = 20 ** 6
ROW3 = 8000

# aa2num after SubstitutionMatrix::setupLetterMapping (reference/lib/mmseqs/src/commons/SubstitutionMatrix.cpp:262-304)
_ORDER = "ACDEFGHIKLMNPQRSTVWYX"
AA2NUM = np.full(256, 20, dtype=np.uint8)
for _i, _c in enumerate(_ORDER):
    AA2NUM[ord(_c)] = _i
    AA2NUM[ord(_c.lower())] = _i
for _a, _b in (("J", "L"), ("Z", "E"), ("B", "D")):
    AA2NUM[ord(_a)] = AA2NUM[ord(_b)]
    AA2NUM[ord(_a.lower())] = AA2NUM[ord(_b)]


def encode(seq):
    return AA2NUM[np.frombuffer(seq.encode(), dtype=np.uint8)]


def pack(seqs):
    """list of str -> (uint8 codes concatenated, uint64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    data = encode("".join(seqs)) if len(seqs) else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(data), off


def load_case(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as f:
        return json.load(f)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "CXX=" + ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")])


class Hit(C.Structure):
    _fields_ = [("seqId", C.c_uint32), ("score", C.c_int32), ("diag", C.c_int32)]


class Aln(C.Structure):
    _fields_ = [("seqId", C.c_uint32), ("score", C.c_int32), ("bitScore", C.c_int32), ("qStart", C.c_int32), ("qEnd", C.c_int32),
                ("dbStart", C.c_int32), ("dbEnd", C.c_int32), ("alnLen", C.c_int32), ("passed", C.c_int32), ("seqIdEst", C.c_float),
                ("pad", C.c_float), ("evalue", C.c_double)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kmers", "lookups", "dbMatches", "doubleHits", "diagCells", "results")]


class Cand(C.Structure):
    _fields_ = [("targetKey", C.c_uint32), ("bitScore", C.c_int32), ("qStart", C.c_int32), ("qEnd", C.c_int32), ("tStart", C.c_int32),
                ("tEnd", C.c_int32), ("tLen", C.c_int32), ("orfKey", C.c_uint32), ("fragFrom", C.c_int32), ("fragTo", C.c_int32)]


class Exon(C.Structure):
    _fields_ = [("candIdx", C.c_uint32), ("contigStart", C.c_int32), ("contigEnd", C.c_int32), ("nucLen", C.c_int32)]


class Call(C.Structure):
    _fields_ = [("targetKey", C.c_uint32), ("strand", C.c_int32), ("totalBitScore", C.c_int32), ("nExons", C.c_int32), ("exonBegin", C.c_uint32),
                ("lowCoord", C.c_int32), ("highCoord", C.c_int32), ("pad", C.c_int32), ("evalue", C.c_double)]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """ctypes binding of oracle/libmb200_oracle.so (the CPU restatement; the checker, never the product)."""

    def __init__(self, need_sim3=True):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        self.L = L
        L.mbo_create.restype = C.c_void_p
        L.mbo_sim3_score.restype = C.POINTER(C.c_int16)
        L.mbo_sim3_index.restype = C.POINTER(C.c_uint32)
        L.mbo_index_entries.restype = C.c_uint64
        L.mbo_evalue.restype = C.c_double
        L.mbo_evalue.argtypes = [C.c_double, C.c_double, C.c_double]
        L.mbo_bitscore.argtypes = [C.c_double]
        self.h = C.c_void_p(L.mbo_create())
        self.max_res = 300
        self.nT = 0
        if need_sim3:
            L.mbo_build_sim3(self.h)

    def close(self):
        if self.h:
            self.L.mbo_destroy(self.h)
            self.h = None

    def set_params(self, kmerThr=127, maxResListLen=300, minDiagScore=15, binSize=2, gapOpen=11, gapExtend=1, evalThr=100.0, minAlnLen=11):
        self.max_res = maxResListLen
        self.L.mbo_set_params(self.h, kmerThr, maxResListLen, minDiagScore, binSize, gapOpen, gapExtend, C.c_double(evalThr), minAlnLen)

    def set_comp_bias_corr(self, on):
        self.L.mbo_set_comp_bias_corr(self.h, int(bool(on)))

    def matrices(self):
        seed = np.zeros(441, np.int16); ung = np.zeros(441, np.int16); aln = np.zeros(441, np.int16)
        sp = np.zeros(21, np.float64); ap = np.zeros(21, np.float64)
        self.L.mbo_get_matrices(self.h, _p(seed, C.c_int16), _p(ung, C.c_int16), _p(aln, C.c_int16), _p(sp, C.c_double), _p(ap, C.c_double))
        return seed, ung, aln, sp, ap

    def sim3(self):
        n = 8000 * ROW3
        sc = np.ctypeslib.as_array(self.L.mbo_sim3_score(self.h), shape=(n,))
        ix = np.ctypeslib.as_array(self.L.mbo_sim3_index(self.h), shape=(n,))
        return sc, ix

    def set_targets(self, masked, unmasked):
        m, off = pack(masked)
        u, off2 = pack(unmasked)
        assert (off == off2).all()
        self.nT = len(masked)
        self.toff = off
        rc = self.L.mbo_set_targets(self.h, _p(m, C.c_uint8), _p(u, C.c_uint8), _p(off, C.c_uint64), C.c_uint32(self.nT))
        assert rc == 0, rc

    def index(self):
        n = self.L.mbo_index_entries(self.h)
        offsets = np.zeros(TABLE + 1, np.uint64)
        entries = np.zeros(6 * n, np.uint8)
        self.L.mbo_index_copy(self.h, _p(offsets, C.c_uint64), _p(entries, C.c_uint8))
        return offsets, entries

    def comp_bias(self, which, seq):
        q = encode(seq)
        out = np.zeros(len(q), np.float32)
        self.L.mbo_comp_bias(self.h, which, _p(q, C.c_uint8), len(q), _p(out, C.c_float))
        return out

    def prefilter(self, queries, threads=0):
        q, off = pack(queries)
        n = len(queries)
        hits = (Hit * (n * self.max_res))()
        cnt = np.zeros(n, np.uint32)
        st = Stats()
        rc = self.L.mbo_prefilter(self.h, _p(q, C.c_uint8), _p(off, C.c_uint64), C.c_uint32(n), hits, _p(cnt, C.c_uint32), C.byref(st), threads)
        assert rc == 0, rc
        out = []
        for i in range(n):
            out.append([(hits[i * self.max_res + k].seqId, hits[i * self.max_res + k].score, hits[i * self.max_res + k].diag) for k in range(cnt[i])])
        return out, {k: getattr(st, k) for k, _ in Stats._fields_}

    def align(self, queries, hit_lists, threads=0):
        """hit_lists[i] = list of target local ids for query i -> list of list of Aln-like dicts"""
        q, off = pack(queries)
        n = len(queries)
        hoff = np.zeros(n + 1, np.uint64)
        hoff[1:] = np.cumsum([len(h) for h in hit_lists], dtype=np.uint64)
        flat = np.array([t for h in hit_lists for t in h], dtype=np.uint32)
        if len(flat) == 0:
            flat = np.zeros(1, np.uint32)
        alns = (Aln * max(1, int(hoff[-1])))()
        cells = C.c_uint64(0)
        rc = self.L.mbo_align(self.h, _p(q, C.c_uint8), _p(off, C.c_uint64), C.c_uint32(n), _p(flat, C.c_uint32), _p(hoff, C.c_uint64), alns, C.byref(cells), threads)
        assert rc == 0, rc
        out = []
        for i in range(n):
            out.append([aln_to_dict(alns[k]) for k in range(int(hoff[i]), int(hoff[i + 1]))])
        return out, cells.value

    def chain(self, cands, total_residues, min_intron=15, max_intron=10000, min_exon_aa=11, max_overlap=10, gap_open=-1, gap_extend=-1,
              tcov=0.5, eval_thr=0.001, max_sets=1):
        return chain_call(self.L.mbo_chain, cands, total_residues, min_intron, max_intron, min_exon_aa, max_overlap, gap_open, gap_extend,
                          tcov, eval_thr, max_sets)


class OrfRec(C.Structure):
    _fields_ = [("contig", C.c_uint32), ("frm", C.c_uint32), ("to", C.c_uint32), ("flags", C.c_uint32), ("aaOff", C.c_uint64), ("aaLen", C.c_uint32), ("pad", C.c_uint32)]


def fmt_orf_header(contig_key, frm, to, flags):
    """Orf::writeOrfHeader (Orf.cpp:440-457): 'key\tfrom+len' / 'from-len', third column only when a side is incomplete"""
    frm, to = int(frm), int(to)
    h = "%d\t%d%s%d" % (contig_key, frm, "+" if frm < to else "-", abs(frm - to))
    if flags & 3:
        h += "\t%d" % (flags & 3)
    return h + "\n"


def oracle_extract_orfs_raw(seq, off, min_len=15, max_len=32734, start_mode=1, threads=1, table=1):
    """seq: bytes of concatenated contigs, off: uint64[n+1] -> (n_orfs, records, residues buffer); the C call only"""
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    L = C.CDLL(ORACLE_SO)
    L.mbo_extract_orfs_tt.restype = C.c_longlong
    cap = max(16, len(seq) // 8)
    recs = (OrfRec * cap)(); aa = C.create_string_buffer(2 * len(seq) + 16); used = C.c_uint64(0)
    n = L.mbo_extract_orfs_tt(seq, _p(off, C.c_uint64), C.c_uint32(len(off) - 1), min_len, max_len, start_mode, int(table), recs, C.c_uint64(cap), aa, C.c_uint64(2 * len(seq) + 16),
                              C.byref(used), int(threads))
    assert n >= 0
    return n, recs, aa


def oracle_extract_orfs(contigs, min_len=15, max_len=32734, start_mode=1, table=1):
    """contigs: [(key, nucleotides)] -> ({orfKey: header}, {orfKey: aa}) through the CPU oracle (extractorfs + translatenucs)"""
    seq = "".join(s for _, s in contigs).encode()
    off = np.zeros(len(contigs) + 1, np.uint64)
    if contigs:
        off[1:] = np.cumsum([len(s) for _, s in contigs])
    n, recs, aa = oracle_extract_orfs_raw(seq, off, min_len, max_len, start_mode, table=table)
    hdr, seqs = {}, {}
    for i in range(n):
        r = recs[i]
        hdr[i] = fmt_orf_header(contigs[r.contig][0], r.frm, r.to, r.flags)
        seqs[i] = aa.raw[r.aaOff:r.aaOff + r.aaLen].decode()
    return hdr, seqs


class Pred(C.Structure):
    _fields_ = [("targetKey", C.c_uint32), ("strand", C.c_int32), ("totalBitscore", C.c_uint32), ("nExons", C.c_uint32), ("lowCoord", C.c_uint32),
                ("highCoord", C.c_uint32), ("exonBegin", C.c_uint32), ("pad", C.c_uint32), ("evalue", C.c_double)]


class PredOut(C.Structure):
    _fields_ = [("clusterId", C.c_uint32), ("clusterLowCoord", C.c_uint32), ("noOverlapClusterId", C.c_uint32), ("isRep", C.c_uint32)]


def parse_calls(entry):
    """one calledExons (dp) DB entry -> (preds [dict], exon lines per pred) as reduceredundancy collects them (reduceredundancy.cpp:222-268):
    a new call starts, per strand, whenever the target key or the low contig coordinate changes"""
    preds, prev = [], {1: None, -1: None}
    for line in entry.split("\n"):
        if not line:
            continue
        w = line.split("\t")
        assert len(w) == 19
        t, s, low = int(w[0]), int(w[1]), int(w[5])
        if prev[s] is None or preds[prev[s]]["targetKey"] != t or preds[prev[s]]["lowCoord"] != low:
            preds.append({"targetKey": t, "strand": s, "totalBitscore": int(w[2]), "evalue": float(w[3]), "evalue_txt": w[3], "nExons": int(w[4]), "lowCoord": low,
                          "highCoord": int(w[6]), "exons": []})
            prev[s] = len(preds) - 1
        preds[prev[s]]["exons"].append(w[7:])
    return preds


def reprint_exon(cols):
    """PotentialExon::setByDPRes -> exonToBuffer (PredictionParser.h:66-150): every field survives except the sequence identity, which is
    parsed to a double, narrowed to float and re-truncated to three digits"""
    f = np.float32(float(cols[2]))
    if f == np.float32(1.0):
        sid = "1.000"
    else:
        sid = "0." + ("0" if f < np.float32(0.10) else "") + ("0" if f < np.float32(0.01) else "") + str(int(np.float32(f * np.float32(1000))))
    ev = "%.3E" % float(cols[3])
    return "\t".join([cols[0], cols[1], sid, ev] + list(cols[4:]))


def fmt_grouped(preds, outs, rep_order, allow_overlaps=False):
    """writeRepPredsInDPFormat + Prediction::predictionToBuffer (reduceredundancy.cpp:149-160, PredictionParser.h:368-396)"""
    s = ""
    for i in rep_order:
        p = preds[i]
        if not allow_overlaps and outs[i]["noOverlapClusterId"] != p["targetKey"]:
            continue
        head = "%d\t%d\t%d\t%s\t%d\t%d\t%d" % (p["targetKey"], p["strand"], p["totalBitscore"], "%.3E" % p["evalue"], p["nExons"], p["lowCoord"], p["highCoord"])
        for e in p["exons"]:
            s += head + "\t" + reprint_exon(e) + "\n"
    return s


def fmt_clusters(preds, outs, member_order):
    """Prediction::predictionClusterToBuffer (PredictionParser.h:398-423)"""
    return "".join("%d\t%d\t%d\t%d\t%d\t%d\n" % (outs[i]["clusterId"], preds[i]["strand"], outs[i]["clusterLowCoord"], preds[i]["targetKey"], preds[i]["strand"],
                                                   preds[i]["lowCoord"]) for i in member_order)


def oracle_reduce_redundancy(entry):
    """dp entry text -> (preds, outs, member_order, rep_order) through the CPU oracle"""
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    L = C.CDLL(ORACLE_SO)
    preds = parse_calls(entry)
    n = len(preds)
    P = (Pred * max(1, n))(); keys = []
    for i, p in enumerate(preds):
        P[i].targetKey = p["targetKey"]; P[i].strand = p["strand"]; P[i].totalBitscore = p["totalBitscore"]; P[i].nExons = len(p["exons"]); P[i].lowCoord = p["lowCoord"]
        P[i].highCoord = p["highCoord"]; P[i].exonBegin = len(keys); P[i].evalue = p["evalue"]
        keys += [int(e[0]) for e in p["exons"]]
    K = np.array(keys + [0], np.uint32)
    O = (PredOut * max(1, n))(); mo = np.zeros(max(1, n), np.uint32); ro = np.zeros(max(1, n), np.uint32); nrep = C.c_uint32(0)
    rc = L.mbo_reduce_redundancy(P, C.c_uint32(n), _p(K, C.c_uint32), O, _p(mo, C.c_uint32), _p(ro, C.c_uint32), C.byref(nrep))
    assert rc == 0
    outs = [{"clusterId": O[i].clusterId, "clusterLowCoord": O[i].clusterLowCoord, "noOverlapClusterId": O[i].noOverlapClusterId, "isRep": O[i].isRep} for i in range(n)]
    return preds, outs, [int(x) for x in mo[:n]], [int(x) for x in ro[:nrep.value]]


def aln_to_dict(a):
    return dict(seqId=a.seqId, score=a.score, bitScore=a.bitScore, qStart=a.qStart, qEnd=a.qEnd, dbStart=a.dbStart, dbEnd=a.dbEnd,
                alnLen=a.alnLen, passed=a.passed, seqIdEst=float(np.float32(a.seqIdEst)), evalue=a.evalue)


def chain_call(fn, cands, total_residues, min_intron, max_intron, min_exon_aa, max_overlap, gap_open, gap_extend, tcov, eval_thr, max_sets):
    """cands: list of dicts with the Cand fields, sorted by (targetKey, orfKey) -> list of calls (dict with 'exons')."""
    n = len(cands)
    arr = (Cand * max(1, n))()
    for i, c in enumerate(cands):
        for k, _ in Cand._fields_:
            setattr(arr[i], k, c[k])
    calls = (Call * max(1, 2 * n * max_sets))()
    exons = (Exon * max(1, n * max_sets))()
    nex = C.c_uint32(0)
    fn.argtypes = [C.POINTER(Cand), C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                   C.c_int, C.POINTER(Call), C.POINTER(Exon), C.POINTER(C.c_uint32)]
    fn.restype = C.c_int
    nc = fn(arr, n, total_residues, min_intron, max_intron, min_exon_aa, max_overlap, gap_open, gap_extend, tcov, eval_thr, max_sets,
            calls, exons, C.byref(nex))
    assert nc >= 0, nc
    out = []
    for i in range(nc):
        c = calls[i]
        ex = [dict(candIdx=exons[k].candIdx, contigStart=exons[k].contigStart, contigEnd=exons[k].contigEnd, nucLen=exons[k].nucLen)
              for k in range(c.exonBegin, c.exonBegin + c.nExons)]
        out.append(dict(targetKey=c.targetKey, strand=c.strand, totalBitScore=c.totalBitScore, evalue=c.evalue, lowCoord=c.lowCoord,
                        highCoord=c.highCoord, exons=ex))
    return out


# ------------------------------------------------------------------------------------------------
# Reference text formatters restated
# ------------------------------------------------------------------------------------------------
def fmt_pref(hits, keys):
    """QueryMatcher::prefilterHitToBuffer (reference/lib/mmseqs/src/prefiltering/QueryMatcher.h:118-130)"""
    return "".join("%d\t%d\t%d\n" % (keys[t], s, d) for t, s, d in hits)


def fmt_seqid(x):
    """Util::fastSeqIdToBuffer + the tab overwrite in Matcher::resultToBuffer
    (reference/lib/mmseqs/src/commons/Util.cpp:222-251, alignment/Matcher.cpp:286-287): 1.0 prints as '1.00'."""
    x = np.float32(x)
    if x == np.float32(1.0):
        return "1.00"
    s = "0."
    if x < 0.10:
        s += "0"
    if x < 0.01:
        s += "0"
    return s + str(int(np.float32(x * np.float32(1000))))


def fmt_eval(e):
    return "%.3E" % e


def sort_alns(alns, keys, tlens):
    """Matcher::compareHits (reference/lib/mmseqs/src/alignment/Matcher.h:157-168): eval asc, score desc, dbLen asc, dbKey asc"""
    return sorted(alns, key=lambda a: (a["evalue"], -a["bitScore"], tlens[a["seqId"]], keys[a["seqId"]]))


def fmt_aln(a, keys, qlen, tlens):
    """Matcher::resultToBuffer, 10 columns (reference/lib/mmseqs/src/alignment/Matcher.cpp:280-327)"""
    return "%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\n" % (keys[a["seqId"]], a["bitScore"], fmt_seqid(a["seqIdEst"]), fmt_eval(a["evalue"]),
                                                         a["qStart"], a["qEnd"], qlen, a["dbStart"], a["dbEnd"], tlens[a["seqId"]])


def parse_orf_header(h):
    """Orf::parseOrfHeader (reference/lib/mmseqs/src/commons/Orf.cpp:351-438): 'contigKey\\tfrom+len' / 'from-len'"""
    w = h.split()
    contig = int(w[0])
    loc = w[1]
    if "+" in loc:
        a, b = loc.split("+")
        return contig, int(a), int(a) + int(b)
    a, b = loc.split("-")
    return contig, int(a), int(a) - int(b)


def fmt_exon_seqid(text):
    """PotentialExon::exonToBuffer seqId chain (reference/src/commons/PredictionParser.h:21,106-136):
    text -> strtod -> double -> float -> *1000 -> int"""
    f = np.float32(float(text))
    if f == np.float32(1.0):
        return "1.000"
    s = "0."
    if f < 0.10:
        s += "0"
    if f < 0.01:
        s += "0"
    return s + str(int(np.float32(f * np.float32(1000))))


def fmt_dp(calls, cands):
    """Prediction::predictionToBuffer, 19 columns per exon line (reference/src/commons/PredictionParser.h:368-396,97-163).
    cands[i] needs the text fields 'seqIdText', 'evalText' of the search_res record it came from."""
    out = ""
    for c in calls:
        head = "%d\t%d\t%d\t%s\t%d\t%d\t%d\t" % (c["targetKey"], c["strand"], c["totalBitScore"], fmt_eval(c["evalue"]), len(c["exons"]),
                                                 c["lowCoord"], c["highCoord"])
        for e in c["exons"]:
            k = cands[e["candIdx"]]
            out += head + "%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n" % (
                k["orfKey"], k["bitScore"], fmt_exon_seqid(k["seqIdText"]), fmt_eval(float(k["evalText"])), k["tStart"], k["tEnd"], k["tLen"],
                e["contigStart"], e["contigEnd"], e["nucLen"], k["fragFrom"], k["fragTo"])
    return out


# ------------------------------------------------------------------------------------------------
# MMseqs database files (data + .index + .dbtype), as the reference's DBWriter / DBReader lay them out
# ------------------------------------------------------------------------------------------------
def write_db(path, entries, dbtype):
    import struct
    pos = 0
    with open(path, "wb") as f, open(path + ".index", "w") as fi:
        for k, e in entries:
            b = e.encode() + b"\0"
            f.write(b)
            fi.write("%d\t%d\t%d\n" % (k, pos, len(b)))
            pos += len(b)
    with open(path + ".dbtype", "wb") as f:
        f.write(struct.pack("<i", dbtype))


def read_db(path):
    data = open(path, "rb").read()
    out = {}
    for line in open(path + ".index"):
        k, o, l = line.split("\t")
        out[k] = data[int(o):int(o) + int(l) - 1].decode()
    return out


# ------------------------------------------------------------------------------------------------
# The whole predictexons chain on the CPU oracle, stage by stage through the reference's text formats
# ------------------------------------------------------------------------------------------------
def bycontig_join(aln, orf_headers, contig_lens):
    """resultspercontig restated (reference/src/exonpredictor/resultspercontig.cpp:64-196): {orfKey: aln entry} -> {contigKey: 20-column lines}"""
    per = {k: [] for k in contig_lens}
    for orf, text in aln.items():
        if not text:
            continue
        contig, frm, to = parse_orf_header(orf_headers[orf])
        orf_len = max(frm, to) - min(frm, to) + 1
        for l in text.splitlines():
            w = l.split("\t")
            w[2] = fmt_seqid(float(w[2]))   # parse -> re-print can lose one unit
            line = "\t".join(w) + "\t%d\t1\t1.00\t0.000E+00\t0\t%d\t%d\t%d\t%d\t%d\n" % (int(orf), orf_len - 1, orf_len, frm, to, contig_lens[contig])
            per[contig].append((int(w[0]), int(orf), line))
    return {c: "".join(r[2] for r in sorted(rows, key=lambda r: (r[0], r[1]))) for c, rows in per.items()}


def cands_of(text):
    out = []
    for l in text.splitlines():
        w = l.split("\t")
        out.append(dict(targetKey=int(w[0]), bitScore=int(w[1]), qStart=int(w[4]), qEnd=int(w[5]), tStart=int(w[7]), tEnd=int(w[8]), tLen=int(w[9]), orfKey=int(w[10]),
                        fragFrom=int(w[17]), fragTo=int(w[18]), seqIdText=w[2], evalText=w[3]))
    return out


def oracle_pipeline(o, contigs, target_keys, target_lens, max_sets=1, gap_open=-1, gap_extend=-1, tcov=0.5, threads=0, min_len=15, max_len=32734, table=1,
                    reverse_fragments=False, **chain_kw):
    """contigs [(key, nucleotides)] -> dict(orf_headers, queries, aln, bycontig, dp) as the reference's DBs (o: Oracle with params and targets set)"""
    hdr, seqs = oracle_extract_orfs(contigs, min_len=min_len, max_len=max_len, table=table)
    if reverse_fragments:       # data/predictexons.sh:52-62 (reverseseq: the null model)
        seqs = {k: v[::-1] for k, v in seqs.items()}
    frags = [seqs[i] for i in range(len(seqs))]
    res, _ = o.prefilter(frags, threads=threads)
    alns, _ = o.align(frags, [[t for t, _, _ in h] for h in res], threads=threads)
    aln = {i: "".join(fmt_aln(a, target_keys, len(frags[i]), target_lens) for a in sort_alns([a for a in al if a["passed"] == 1], target_keys, target_lens))
           for i, al in enumerate(alns)}
    byc = bycontig_join(aln, hdr, {k: len(s) for k, s in contigs})
    dp = {}
    for ck, text in byc.items():
        cands = cands_of(text)
        dp[ck] = fmt_dp(o.chain(cands, sum(target_lens), max_sets=max_sets, gap_open=gap_open, gap_extend=gap_extend, tcov=tcov, **chain_kw), cands)
    return dict(orf_headers=hdr, queries=seqs, aln=aln, bycontig=byc, dp=dp)


def sensitivity_to_kmer_thr(s):
    """Prefiltering::getKmerThreshold for sequence searches with k = 6 (Prefiltering.cpp:1049-1055,1063): float base 163.2, the product s * 8.917 in double,
    the difference stored in a float and truncated"""
    return int(np.float32(float(np.float32(163.2)) - float(np.float32(s)) * 8.917))


# `metaeuk predictexons` flags of tests/golden/flags.json.gz -> (oracle parameters, ORF / pipeline options, exon-chaining options)
FLAG_VARIANTS = {
    "default": ({}, {}),
    "s6": (dict(kmerThr=sensitivity_to_kmer_thr(6)), {}),
    "s1": (dict(kmerThr=sensitivity_to_kmer_thr(1)), {}),
    "s7.5": (dict(kmerThr=sensitivity_to_kmer_thr(7.5)), {}),
    "kscore": (dict(kmerThr=140), {}),
    "maxseqs10": (dict(maxResListLen=10), {}),
    "eval": (dict(evalThr=0.001), {}),
    "minexon": (dict(minAlnLen=20), dict(min_len=30, min_exon_aa=20)),        # PredictExons.cpp:44: the align stage's --min-aln-len follows --min-exon-aa
    "maxlen": ({}, dict(max_len=120)),
    "nocomp": (dict(comp_bias_corr=False), {}),
    "reverse": ({}, dict(reverse_fragments=True)),
    "table4": ({}, dict(table=4)),
    "table6": ({}, dict(table=6)),
    "nomask": (dict(nomask=True), {}),
    "minungapped": (dict(minDiagScore=30), {}),
    "dp": ({}, dict(max_intron=400, min_intron=60, max_overlap=4, gap_open=-3, gap_extend=-2, eval_thr=1e-6, tcov=0.3, max_sets=2)),
}


def fmt_fasta_headers(dp, target_acc, contig_acc):
    """{contigKey: dp_predictions entry} -> the FASTA header of every call as `unitesetstofasta` writes it (default flags; reference/src/
    exonpredictor/unitesetstofasta.cpp:114-190): T|C|strand|bit score|E-value|#exons|low|high|start[adjusted]:end[end]:len[adjusted] per exon.
    The E-value is the dp text parsed to double and printed by an ostream with its default 6 significant digits (= %g)."""
    out = []
    for ck, entry in dp.items():
        calls = []
        for l in entry.splitlines():
            w = l.split("\t")
            head = tuple(w[:7])
            if not calls or calls[-1][0] != head:
                calls.append((head, []))
            calls[-1][1].append(w[7:])
        for (t, s, score, ev, nex, low, high), exons in calls:
            h = "%s|%s|%s|%s|%s|%s|%s|%s" % (target_acc[str(t)], contig_acc[str(ck)], "+" if int(s) == 1 else "-", score, "%g" % float(ev), nex, low, high)
            last = -1
            for e in exons:
                t_start, t_end, c_start, c_end, nlen = int(e[4]), int(e[5]), int(e[7]), int(e[8]), int(e[9])
                adj_start, adj_len = c_start, nlen
                if last >= t_start:
                    d = last - t_start + 1
                    adj_start += 3 * d; adj_len -= 3 * d
                last = t_end
                h += "|%d[%d]:%d[%d]:%d[%d]" % (abs(c_start), abs(adj_start), abs(c_end), abs(c_end), nlen, adj_len)
            out.append(h)
    return out
