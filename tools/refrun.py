#!/usr/bin/env python3
"""Run the UNMODIFIED reference binary (oracle/_ref/metaeuk-ref, built by oracle/build_ref.sh) on bench-scale workloads.

TEST / BASELINE INFRASTRUCTURE -- never imported by the product (metaeuk_b200/).  Two users:
  * tests/golden/make_bench_digests.py (build container): pins the outputs of the workloads bench.py times as SHA-256 digests,
  * bench.py --impl reference and its cpu_baseline leg (GPU box): times the reference's AVX2 CPU path on the host cores.

The module calls and flags are the ones `metaeuk predictexons` passes (reference/data/predictexons.sh:40-82, captured flag
strings as in tests/golden/make_golden.py); stage times are the reference's own "Time for processing" lines
(reference/lib/mmseqs/src/commons/Application.cpp:45-50) next to the wall clock of each process.
"""
import hashlib
import os
import re
import shutil
import struct
import subprocess
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NUM2AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)

PREF_FLAGS = ("--sub-mat aa:blosum62.out,nucl:nucleotide.out --seed-sub-mat aa:VTML80.out,nucl:nucleotide.out "
              "-k 0 --target-search-mode 0 --k-score seq:2147483647,prof:2147483647 --alph-size aa:21,nucl:5 "
              "--max-seq-len 65535 --max-seqs 300 --split 1 --split-mode 2 --split-memory-limit 0 -c 0 --cov-mode 0 "
              "--comp-bias-corr 1 --comp-bias-corr-scale 1 --diag-score 1 --exact-kmer-matching 0 --mask 1 "
              "--mask-prob 0.9 --mask-lower-case 0 --mask-n-repeat 0 --min-ungapped-score 15 --add-self-matches 0 "
              "--spaced-kmer-mode 1 --db-load-mode 0 --pca substitution:1.100,context:1.400 "
              "--pcb substitution:4.100,context:5.800 --threads {thr} --compressed 0 -v 3 -s 4.0")
ALN_FLAGS = ("--sub-mat aa:blosum62.out,nucl:nucleotide.out -a 0 --alignment-mode 2 --alignment-output-mode 0 "
             "--wrapped-scoring 0 -e 100 --min-seq-id 0 --min-aln-len 11 --seq-id-mode 0 --alt-ali 0 -c 0 --cov-mode 0 "
             "--max-seq-len 65535 --comp-bias-corr 1 --comp-bias-corr-scale 1 --max-rejected 2147483647 "
             "--max-accept 2147483647 --add-self-matches 0 --db-load-mode 0 --pca substitution:1.100,context:1.400 "
             "--pcb substitution:4.100,context:5.800 --score-bias 0 --realign 0 --realign-score-bias -0.2 "
             "--realign-max-seqs 2147483647 --corr-score-weight 0 --gap-open aa:11,nucl:5 --gap-extend aa:1,nucl:2 "
             "--zdrop 40 --threads {thr} --compressed 0 -v 3")
ORF_FLAGS = ("--min-length 15 --max-length 32734 --max-gaps 2147483647 --contig-start-mode 2 --contig-end-mode 2 "
             "--orf-start-mode 1 --forward-frames 1,2,3 --reverse-frames 1,2,3 --translation-table 1 --translate 0 "
             "--use-all-table-starts 0 --id-offset 0 --create-lookup 0 --threads {thr} --compressed 0 -v 3")
DP_FLAGS = ("--metaeuk-eval 0.001 --metaeuk-tcov 0.5 --max-intron 10000 --min-intron 15 --min-exon-aa 11 "
            "--max-overlap 10 --max-exon-sets 1 --set-gap-open -1 --set-gap-extend -1 --score-bias 0 "
            "--threads {thr} --compressed 0 -v 3")
IDX_FLAGS = ("--seed-sub-mat aa:VTML80.out,nucl:nucleotide.out -k 6 -s 4 --mask 1 --mask-prob 0.9 "
             "--mask-lower-case 0 --mask-n-repeat 0 --spaced-kmer-mode 1 --comp-bias-corr 1 --split 1 --threads {thr} -v 3")


def ref_bin():
    p = os.environ.get("METAEUK_REF", os.path.join(ROOT, "oracle", "_ref", "metaeuk-ref"))
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


def write_fasta(path, ascii_bytes, off, prefix):
    """ascii_bytes: uint8 array of letters; off: offsets [n+1]; entry i gets the header '>' + prefix + i"""
    ascii_bytes = np.asarray(ascii_bytes, np.uint8)
    with open(path, "wb") as f:
        for i in range(len(off) - 1):
            f.write(b">%s%d\n" % (prefix, i))
            f.write(ascii_bytes[int(off[i]):int(off[i + 1])].tobytes())
            f.write(b"\n")


def write_aa_db(path, codes, off):
    """an amino-acid sequence DB (dbtype 0) with keys 0..n-1 in offset order: entries 'residues\\n\\0' (DBReader.h:211-231) + header DB"""
    n = len(off) - 1
    letters = NUM2AA[np.asarray(codes, np.uint8)]
    lens = (off[1:] - off[:-1]).astype(np.int64)
    pos = np.zeros(n + 1, np.int64); pos[1:] = np.cumsum(lens + 2)
    data = np.empty(int(pos[-1]), np.uint8)
    idx = np.arange(len(letters), dtype=np.int64) + 2 * np.repeat(np.arange(n, dtype=np.int64), lens)
    data[idx] = letters
    data[pos[1:] - 2] = 10; data[pos[1:] - 1] = 0
    data.tofile(path)
    with open(path + ".index", "w") as f:
        f.write("".join("%d\t%d\t%d\n" % (i, pos[i], lens[i] + 2) for i in range(n)))
    with open(path + ".dbtype", "wb") as f:
        f.write(struct.pack("<i", 0))
    hp = 0
    with open(path + "_h", "wb") as f, open(path + "_h.index", "w") as fi:
        for i in range(n):
            b = b"F%d\n\0" % i
            f.write(b); fi.write("%d\t%d\t%d\n" % (i, hp, len(b))); hp += len(b)
    with open(path + "_h.dbtype", "wb") as f:
        f.write(struct.pack("<i", 12))


_TIME_RE = re.compile(r"Time for processing: (\d+)h (\d+)m (\d+)s (\d+)ms")


def run(cmd, log=None):
    """-> (wall seconds, the module's own 'Time for processing' in seconds or None, stdout)"""
    t0 = time.perf_counter()
    r = subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    wall = time.perf_counter() - t0
    if log is not None:
        log.append("$ " + cmd + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("reference command failed (%d): %s\n%s" % (r.returncode, cmd, r.stdout[-4000:]))
    m = _TIME_RE.findall(r.stdout)
    own = None
    if m:
        h, mi, s, ms = (int(x) for x in m[-1]); own = 3600 * h + 60 * mi + s + ms / 1000.0
    return wall, own, r.stdout


class RefRun:
    """one working directory with the reference's DBs"""

    def __init__(self, workdir, threads=None):
        self.ref = ref_bin()
        if not self.ref:
            raise RuntimeError("oracle/_ref/metaeuk-ref is not built (oracle/build_ref.sh needs /root/reference)")
        self.d = workdir; self.thr = threads or os.cpu_count() or 1
        os.makedirs(workdir, exist_ok=True)
        self.log = []; self.times = {}

    def p(self, name):
        return os.path.join(self.d, name)

    def _run(self, stage, cmd):
        w, own, out = run(self.ref + " " + cmd, self.log)
        self.times[stage] = {"wall_s": w, "own_s": own}
        return out

    def rm(self, *names):
        for n in names:
            for f in os.listdir(self.d):
                if f == n or f.startswith(n + ".") or f.startswith(n + "_h"):
                    os.remove(os.path.join(self.d, f))

    def create_targets(self, prot_codes, poff):
        """proteins -> sequence DB (keys = generator order: --shuffle 0) + the prefilter index (indexdb: tantan masking + k-mer index)"""
        write_fasta(self.p("proteins.faa"), NUM2AA[np.asarray(prot_codes, np.uint8)], poff, b"P")
        self._run("createdb_targets", "createdb %s %s --shuffle 0 --dbtype 1 -v 3" % (self.p("proteins.faa"), self.p("proteins")))
        os.remove(self.p("proteins.faa"))
        out = self._run("indexdb", "indexdb %s %s %s" % (self.p("proteins"), self.p("proteins"), IDX_FLAGS.format(thr=self.thr)))
        m = re.search(r"Masked residues: (\d+)", out)
        self.masked_residues = int(m.group(1)) if m else None
        return self.times["indexdb"]

    def create_contigs(self, nuc_ascii, off, name="contigs"):
        write_fasta(self.p(name + ".fna"), nuc_ascii, off, b"C")
        self._run("createdb_" + name, "createdb %s %s --shuffle 0 --dbtype 2 -v 3" % (self.p(name + ".fna"), self.p(name)))
        os.remove(self.p(name + ".fna"))

    def predictexons(self, contigs="contigs", keep=False):
        """the six module calls of data/predictexons.sh:40-82 against the prebuilt index; -> stage times"""
        t = self.thr
        self.rm("nucl_6f", "aa_6f", "pref", "aln", "bycontig", "dp")
        self._run("extractorfs", "extractorfs %s %s %s" % (self.p(contigs), self.p("nucl_6f"), ORF_FLAGS.format(thr=t)))
        self._run("translatenucs", "translatenucs %s %s --translation-table 1 --add-orf-stop 0 -v 3 --compressed 0 --threads %d" % (self.p("nucl_6f"), self.p("aa_6f"), t))
        self._run("prefilter", "prefilter %s %s %s %s" % (self.p("aa_6f"), self.p("proteins.idx"), self.p("pref"), PREF_FLAGS.format(thr=t)))
        self._run("align", "align %s %s %s %s %s" % (self.p("aa_6f"), self.p("proteins.idx"), self.p("pref"), self.p("aln"), ALN_FLAGS.format(thr=t)))
        self._run("resultspercontig", "resultspercontig %s %s %s %s --threads %d --compressed 0 -v 3" % (self.p(contigs), self.p("nucl_6f"), self.p("aln"), self.p("bycontig"), t))
        self._run("collectoptimalset", "collectoptimalset %s %s %s %s" % (self.p("bycontig"), self.p("proteins"), self.p("dp"), DP_FLAGS.format(thr=t)))
        stages = ("extractorfs", "translatenucs", "prefilter", "align", "resultspercontig", "collectoptimalset")
        n_frag = sum(1 for _ in open(self.p("aa_6f.index")))
        return {"fragments": n_frag, "wall_s": sum(self.times[s]["wall_s"] for s in stages), "stages": {s: dict(self.times[s]) for s in stages}}

    def search(self, frag_codes, foff, align=True):
        """prefilter (+ align) of an amino-acid fragment DB (BASELINE configs[1]/[2])"""
        t = self.thr
        self.rm("frags", "pref", "aln")
        write_aa_db(self.p("frags"), frag_codes, foff)
        self._run("prefilter", "prefilter %s %s %s %s" % (self.p("frags"), self.p("proteins.idx"), self.p("pref"), PREF_FLAGS.format(thr=t)))
        stages = ["prefilter"]
        if align:
            self._run("align", "align %s %s %s %s %s" % (self.p("frags"), self.p("proteins.idx"), self.p("pref"), self.p("aln"), ALN_FLAGS.format(thr=t)))
            stages.append("align")
        return {"fragments": len(foff) - 1, "wall_s": sum(self.times[s]["wall_s"] for s in stages), "stages": {s: dict(self.times[s]) for s in stages}}

    def cleanup(self):
        shutil.rmtree(self.d, ignore_errors=True)


# ---------------------------------------------------------------------------------------------------------------
# canonical digests of result DBs: entries in ascending key order, empty entries skipped, ">key\n" + entry text
# ---------------------------------------------------------------------------------------------------------------
def read_result_db(path):
    """-> (keys ascending [int64], list of entry bytes without the trailing NUL); data may be split over path.0 .. path.N"""
    files = [path] if os.path.exists(path) else []
    i = 0
    while not os.path.exists(path) and os.path.exists("%s.%d" % (path, i)):
        files.append("%s.%d" % (path, i)); i += 1
    data = b"".join(open(f, "rb").read() for f in files)
    idx = np.loadtxt(path + ".index", dtype=np.int64, ndmin=2) if os.path.getsize(path + ".index") else np.zeros((0, 3), np.int64)
    idx = idx[np.argsort(idx[:, 0], kind="stable")]
    return idx[:, 0], [data[int(o):int(o) + int(l) - 1] for _, o, l in idx]


def digest_entries(keys, entries, key_lo=None, key_hi=None):
    """-> {"sha256", "entries" (non-empty), "lines"} over keys in [key_lo, key_hi)"""
    h = hashlib.sha256(); ne = 0; nl = 0
    for k, e in zip(keys, entries):
        if (key_lo is not None and k < key_lo) or (key_hi is not None and k >= key_hi) or not e:
            continue
        h.update(b">%d\n" % int(k)); h.update(e); ne += 1; nl += e.count(b"\n")
    return {"sha256": h.hexdigest(), "entries": ne, "lines": nl}


def digest_text(text_with_key_lines):
    """digest of a buffer already in the canonical '>key\\n' + entry form (what mb200_text_dp_db prints)"""
    b = bytes(text_with_key_lines)
    return {"sha256": hashlib.sha256(b).hexdigest(), "entries": b.count(b">"), "lines": b.count(b"\n") - b.count(b">")}
