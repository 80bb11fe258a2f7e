#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs a reference `metaeuk`
binary; env METAEUK_REF, default /tmp/oracle-build/src/metaeuk, built out-of-tree from
/root/reference with the recipe in SURVEY.md section 8c).  Nothing here runs on the GPU box:
the outputs are committed as small fixtures and read by tests/.

For every case it
  1. writes seeded synthetic proteins / contigs / fragments,
  2. runs the reference modules with exactly the flags `metaeuk predictexons` passes
     (captured from a reference run): createdb, extractorfs, translatenucs,
     prefilter, align, resultspercontig, collectoptimalset,
  3. dumps inputs (sequences in the reference's *local id* order = data-offset order,
     reference/lib/mmseqs/src/commons/DBReader.cpp LINEAR_ACCCESS) and every stage's result DB
     per key into one JSON file  tests/golden/<case>.json.gz

Result DB record formats: SURVEY.md Appendix B.
"""
import gzip
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

REF = os.environ.get("METAEUK_REF", "/tmp/oracle-build/src/metaeuk")
HERE = os.path.dirname(os.path.abspath(__file__))
AA = "ACDEFGHIKLMNPQRSTVWY"
W = [7.8, 1.9, 5.4, 6.3, 3.9, 7.4, 2.3, 5.1, 5.7, 9.0, 2.2, 4.5, 5.2, 4.3, 5.1, 7.1, 5.8, 6.4, 1.3, 3.2]
BASES = "TCAG"
CODE = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
CODON = {}
_i = 0
for _a in BASES:
    for _b in BASES:
        for _c in BASES:
            CODON.setdefault(CODE[_i], []).append(_a + _b + _c)
            _i += 1

PREF_FLAGS = ("--sub-mat aa:blosum62.out,nucl:nucleotide.out --seed-sub-mat aa:VTML80.out,nucl:nucleotide.out "
              "-k 0 --target-search-mode 0 --k-score seq:2147483647,prof:2147483647 --alph-size aa:21,nucl:5 "
              "--max-seq-len 65535 --max-seqs {maxseqs} --split 1 --split-mode 2 --split-memory-limit 0 -c 0 --cov-mode 0 "
              "--comp-bias-corr 1 --comp-bias-corr-scale 1 --diag-score 1 --exact-kmer-matching 0 --mask {mask} "
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
            "--max-overlap 10 --max-exon-sets {sets} --set-gap-open -1 --set-gap-extend -1 --score-bias 0 "
            "--threads {thr} --compressed 0 -v 3")


def run(cmd, log, quiet=False):
    r = subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + cmd + "\n" + r.stdout)
    if r.returncode != 0:
        if not quiet:
            sys.stderr.write(r.stdout)
        raise SystemExit("reference command failed: " + cmd)
    return r.stdout


def read_db(path):
    """-> list of (key, bytes) in data-offset order (the reference's LINEAR_ACCCESS local ids)."""
    if os.path.exists(path):
        files = [path]
    else:
        files = []
        i = 0
        while os.path.exists("%s.%d" % (path, i)):
            files.append("%s.%d" % (path, i))
            i += 1
    data = b"".join(open(f, "rb").read() for f in files)
    idx = []
    for line in open(path + ".index"):
        k, o, l = line.split("\t")
        idx.append((int(o), int(k), int(l)))
    idx.sort()
    return [(k, data[o:o + l]) for o, k, l in idx]


def seqdb(path):
    return [(k, e[:-2].decode()) for k, e in read_db(path)]  # strip "\n\0"


def resdb(path):
    return {str(k): e[:-1].decode() for k, e in read_db(path)}  # strip "\0"


IDX_FLAGS = ("--seed-sub-mat aa:VTML80.out,nucl:nucleotide.out -k 6 -s 4 --mask {mask} --mask-prob 0.9 "
             "--mask-lower-case 0 --mask-n-repeat 0 --spaced-kmer-mode 1 --comp-bias-corr 1 --split 1 --threads {thr} -v 3")
NUM2AA = "ACDEFGHIKLMNPQRSTVWYX"


def ref_index(tmp, db, mask, thr, log, extra=None):
    """Run the reference `indexdb` (same IndexBuilder::fillDatabase the prefilter calls,
    reference/lib/mmseqs/src/prefiltering/IndexBuilder.cpp:55) and pull out of the .idx DB
    (entry ids: PrefilteringIndexReader.cpp:10-34): the tantan-MASKED target sequences
    (SEQINDEXDATA=14 / SEQINDEXSEQOFFSET=16), and sha256 digests of the k-mer index
    (ENTRIES=9, ENTRIESOFFSETS=10) and of the 3-mer similarity table (SCOREMATRIX3MER=4)."""
    import hashlib
    import struct
    flags = IDX_FLAGS.format(mask=mask, thr=thr)
    for k, v in (extra or {}).items():
        flags = flags.replace("%s 0" % k, "%s %s" % (k, v))
    out = run("%s indexdb %s/%s %s/%s %s" % (REF, tmp, db, tmp, db, flags), log)
    ent = {}
    for line in open("%s/%s.idx.index" % (tmp, db)):
        k, o, l = line.split("\t")
        ent[int(k)] = (int(o), int(l))
    f = open("%s/%s.idx" % (tmp, db), "rb")

    def get(k):
        o, l = ent[k]
        f.seek(o)
        return f.read(l)
    nseq = struct.unpack("<q", get(13)[:8])[0]
    seqoff = struct.unpack("<%dq" % (nseq + 1), get(16)[:8 * (nseq + 1)])
    seqdata = get(14)
    masked = ["".join(NUM2AA[c] for c in seqdata[seqoff[i]:seqoff[i + 1]]) for i in range(nseq)]
    nent = struct.unpack("<q", get(12)[:8])[0]
    entries = get(9)[:6 * nent]
    offsets = get(10)[:8 * (20 ** 6 + 1)]
    import numpy as np
    sm3 = get(4)
    # ScoreMatrix::serialize: int16 score[8000][rowSize] then uint32 index[8000][rowSize]; rowSize is padded to the
    # build's SIMD width (8064 on AVX2, 8032 on SSE) with -255/0 -- digest only the 8000 real columns.
    row = (len(sm3) - 1) // 6 // 8000
    sc = np.frombuffer(sm3[:2 * 8000 * row], dtype=np.int16).reshape(8000, row)[:, :8000]
    ix = np.frombuffer(sm3[2 * 8000 * row:6 * 8000 * row], dtype=np.uint32).reshape(8000, row)[:, :8000]
    res = {
        "masked_targets": masked, "idx_n_entries": nent,
        "idx_entries_sha256": hashlib.sha256(entries).hexdigest(),
        "idx_offsets_u64_sha256": hashlib.sha256(offsets).hexdigest(),
        "sim3_score_i16_sha256": hashlib.sha256(np.ascontiguousarray(sc).tobytes()).hexdigest(),
        "sim3_index_u32_sha256": hashlib.sha256(np.ascontiguousarray(ix).tobytes()).hexdigest(),
        "idx_masked_residues": int(out.split("Masked residues: ")[1].split()[0]),
    }
    f.close()
    os.remove("%s/%s.idx" % (tmp, db))
    return res


def rand_prot(rng, mean=330, sd=120, lo=60):
    L = max(lo, int(rng.gauss(mean, sd)))
    return "".join(rng.choices(AA, weights=W, k=L))


def mutate(rng, s, rate):
    return "".join(rng.choice(AA) if rng.random() < rate else c for c in s)


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def make_proteins(rng, n, family=0, lowcomplex=0):
    prots = [rand_prot(rng) for _ in range(n)]
    if family:
        base = rand_prot(rng, 220, 10)
        for _ in range(family):
            prots.append(mutate(rng, base, rng.choice([0.05, 0.1, 0.2, 0.3])))
    for _ in range(lowcomplex):
        unit = "".join(rng.choices(AA, k=rng.randint(1, 4)))
        p = rand_prot(rng, 150, 20)
        p = p[:60] + unit * (80 // len(unit)) + p[60:]
        prots.append(p)
    rng.shuffle(prots)
    return prots


def make_contigs(rng, prots, nc, cl):
    out = []
    for _ in range(nc):
        parts = []
        total = 0
        while total < cl:
            inter = "".join(rng.choices("ACGT", k=rng.randint(200, 1500)))
            parts.append(inter)
            total += len(inter)
            if rng.random() < 0.8:
                p = mutate(rng, rng.choice(prots), rng.choice([0.1, 0.3, 0.5]))
                k = rng.randint(1, 4)
                cuts = sorted(rng.sample(range(20, len(p) - 20), k - 1)) if len(p) > 60 and k > 1 else []
                segs = [p[a:b] for a, b in zip([0] + cuts, cuts + [len(p)])]
                gene = ""
                for j, sg in enumerate(segs):
                    gene += "".join(rng.choice(CODON[x]) for x in sg)
                    if j < len(segs) - 1:
                        gene += "".join(rng.choices("ACGT", k=rng.randint(40, 600)))
                if rng.random() < 0.5:
                    gene = rc(gene)
                parts.append(gene)
                total += len(gene)
        out.append("".join(parts)[:cl])
    return out


def make_fragments(rng, prots, n):
    """AA query fragments: half mutated protein slices, half background; a few with X / short / long."""
    out = []
    for i in range(n):
        L = max(15, int(rng.gauss(100, 40)))
        if i % 2 == 0:
            p = rng.choice(prots)
            if len(p) > L:
                a = rng.randrange(0, len(p) - L)
                s = mutate(rng, p[a:a + L], rng.choice([0.1, 0.3, 0.5]))
            else:
                s = mutate(rng, p, 0.2)
        else:
            s = "".join(rng.choices(AA, weights=W, k=L))
        if i % 37 == 5:
            j = rng.randrange(len(s))
            s = s[:j] + "X" + s[j + 1:]
        out.append(s)
    out.append("ACDEFGHIK")          # shorter than the spaced window (no k-mer)
    out.append("ACDEFGHIKL")         # exactly one window
    out.append(mutate(rng, prots[0] + prots[1], 0.1))  # long query spanning two targets
    return out


def write_fasta(path, seqs, prefix):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s%d\n%s\n" % (prefix, i, s))


def adversarial_sets(rng):
    """targets and queries chosen to hit the corners of the path: > 300 near-identical targets (ties at the saturated 8-bit diagonal score -> exact re-score,
    top-300 cut among equals), exact copies (maximal scores), homopolymers / short-period repeats / low complexity (masking, huge k-mer lists, many double hits on
    neighbouring diagonals), X runs, sequences around the window length, very long sequences"""
    base = rand_prot(rng, 180, 10)
    prots = [mutate(rng, base, rng.choice([0.0, 0.0, 0.01, 0.02])) for _ in range(340)]
    prots += [rand_prot(rng) for _ in range(120)]
    for _ in range(30):
        unit = "".join(rng.choices(AA, k=rng.randint(1, 6)))
        p = rand_prot(rng, 120, 30, 40)
        a = rng.randrange(len(p))
        prots.append(p[:a] + unit * rng.randint(5, 40) + p[a:])
    for _ in range(12):
        p = rand_prot(rng, 200, 40)
        a = rng.randrange(len(p) - 20)
        prots.append(p[:a] + "X" * rng.randint(1, 25) + p[a:])
    prots += ["".join(rng.choices(AA, weights=W, k=n)) for n in (10, 11, 12, 15, 20, 1500, 3000)]
    prots += ["A" * 60, "AC" * 40, "ACD" * 30, "W" * 25 + rand_prot(rng, 80, 5)]
    rng.shuffle(prots)
    frags = [base, mutate(rng, base, 0.1), base[:40], base[100:], base + base]
    frags += [p for p in rng.sample(prots, 25)]                                     # exact copies (incl. repeats, X runs, long, short)
    frags += [mutate(rng, p, 0.15) for p in rng.sample(prots, 25)]
    frags += ["A" * 50, "AC" * 30, "ACD" * 25, "ACDE" * 20, "W" * 40, "X" * 30, "ACDEFGHIKX" * 6, "XACDEFGHIKLMNX" * 4]
    frags += ["ACDEFGHIK", "ACDEFGHIKL", "ACDEFGHIKLM", "MKV", "M"]
    frags += ["".join(rng.choices(AA, weights=W, k=n)) for n in (15, 16, 31, 32, 33, 64, 65, 255, 256, 257, 512, 513, 1023, 1024, 1025, 2000)]
    frags += [mutate(rng, "".join(rng.sample(prots, 1)[0] for _ in range(3)), 0.1)[:n] for n in (300, 520, 900)]
    for _ in range(40):
        frags.append("".join(rng.choices(AA, weights=W, k=max(15, int(rng.gauss(100, 40))))))
    return prots, frags


def case_fragments(name, seed, nprot, nfrag, family=0, lowcomplex=0, mask=1, maxseqs=300, thr=8, adversarial=False):
    rng = random.Random(seed)
    tmp = tempfile.mkdtemp(prefix="mb200_golden_")
    log = []
    if adversarial:
        prots, frags = adversarial_sets(rng)
    else:
        prots = make_proteins(rng, nprot, family, lowcomplex)
        frags = make_fragments(rng, prots, nfrag)
    write_fasta(tmp + "/proteins.faa", prots, "P")
    write_fasta(tmp + "/frags.faa", frags, "F")
    run("%s createdb %s/proteins.faa %s/proteins -v 3" % (REF, tmp, tmp), log)
    run("%s createdb %s/frags.faa %s/frags --shuffle 0 -v 3" % (REF, tmp, tmp), log)
    run("%s prefilter %s/frags %s/proteins %s/pref %s" % (REF, tmp, tmp, tmp, PREF_FLAGS.format(mask=mask, thr=thr, maxseqs=maxseqs)), log)
    run("%s align %s/frags %s/proteins %s/pref %s/aln %s" % (REF, tmp, tmp, tmp, tmp, ALN_FLAGS.format(thr=thr)), log)
    preflog = log[-2]
    out = {
        "case": name, "kind": "fragments", "seed": seed, "mask": mask, "max_seqs": maxseqs,
        "targets": seqdb(tmp + "/proteins"), "queries": seqdb(tmp + "/frags"),
        "pref": resdb(tmp + "/pref"), "aln": resdb(tmp + "/aln"),
        "masked_residues": int(preflog.split("Masked residues: ")[1].split()[0]),
        "kmer_thr": int(preflog.split("k-mer similarity threshold: ")[1].split()[0]),
        "ref_version": run("%s version" % REF, log).strip(),
    }
    out.update(ref_index(tmp, "proteins", mask, thr, log))
    assert out["idx_masked_residues"] == out["masked_residues"]
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
        json.dump(out, f)
    shutil.rmtree(tmp)
    print(name, "targets", len(out["targets"]), "queries", len(out["queries"]),
          "masked", out["masked_residues"],
          "pref lines", sum(v.count("\n") for v in out["pref"].values()),
          "aln lines", sum(v.count("\n") for v in out["aln"].values()))


def case_contigs(name, seed, nprot, nc, cl, mask=1, sets=1, thr=8):
    rng = random.Random(seed)
    tmp = tempfile.mkdtemp(prefix="mb200_golden_")
    log = []
    prots = make_proteins(rng, nprot)
    contigs = make_contigs(rng, prots, nc, cl)
    write_fasta(tmp + "/proteins.faa", prots, "P")
    write_fasta(tmp + "/contigs.fna", contigs, "C")
    run("%s createdb %s/proteins.faa %s/proteins -v 3" % (REF, tmp, tmp), log)
    run("%s createdb %s/contigs.fna %s/contigs -v 3" % (REF, tmp, tmp), log)
    run("%s extractorfs %s/contigs %s/nucl_6f %s" % (REF, tmp, tmp, ORF_FLAGS.format(thr=thr)), log)
    run("%s translatenucs %s/nucl_6f %s/aa_6f --translation-table 1 --add-orf-stop 0 -v 3 --compressed 0 --threads %d" % (REF, tmp, tmp, thr), log)
    run("%s prefilter %s/aa_6f %s/proteins %s/pref %s" % (REF, tmp, tmp, tmp, PREF_FLAGS.format(mask=mask, thr=thr, maxseqs=300)), log)
    preflog = log[-1]
    run("%s align %s/aa_6f %s/proteins %s/pref %s/aln %s" % (REF, tmp, tmp, tmp, tmp, ALN_FLAGS.format(thr=thr)), log)
    run("%s resultspercontig %s/contigs %s/nucl_6f %s/aln %s/bycontig --threads %d --compressed 0 -v 3" % (REF, tmp, tmp, tmp, tmp, thr), log)
    run("%s collectoptimalset %s/bycontig %s/proteins %s/dp %s" % (REF, tmp, tmp, tmp, DP_FLAGS.format(thr=thr, sets=sets)), log)
    out = {
        "case": name, "kind": "contigs", "seed": seed, "mask": mask, "max_exon_sets": sets,
        "targets": seqdb(tmp + "/proteins"),
        "contig_lens": {str(k): len(s) for k, s in seqdb(tmp + "/contigs")},
        "contigs": seqdb(tmp + "/contigs"),            # nucleotide contigs: input of extractorfs + translatenucs
        "queries": seqdb(tmp + "/aa_6f"),
        "orf_headers": {str(k): e[:-1].decode() for k, e in read_db(tmp + "/nucl_6f_h")},
        "pref": resdb(tmp + "/pref"), "aln": resdb(tmp + "/aln"),
        "bycontig": resdb(tmp + "/bycontig"), "dp": resdb(tmp + "/dp"),
        "masked_residues": int(preflog.split("Masked residues: ")[1].split()[0]),
        "kmer_thr": int(preflog.split("k-mer similarity threshold: ")[1].split()[0]),
        "ref_version": run("%s version" % REF, log).strip(),
    }
    out.update(ref_index(tmp, "proteins", mask, thr, log))
    assert out["idx_masked_residues"] == out["masked_residues"]
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
        json.dump(out, f)
    shutil.rmtree(tmp)
    print(name, "targets", len(out["targets"]), "fragments", len(out["queries"]),
          "masked", out["masked_residues"],
          "pref lines", sum(v.count("\n") for v in out["pref"].values()),
          "aln lines", sum(v.count("\n") for v in out["aln"].values()),
          "dp lines", sum(v.count("\n") for v in out["dp"].values()))


def case_redundancy(name, seed, nfam=8, famsize=6, nc=10, cl=12000, sets=2, thr=4):
    """reduceredundancy needs calls that share exons: protein FAMILIES as targets, genes planted from family members,
    so every gene is called once per homologous target (and --max-exon-sets 2 adds same-target sub-optimal sets)"""
    rng = random.Random(seed)
    tmp = tempfile.mkdtemp(prefix="mb200_golden_")
    log = []
    prots = [rand_prot(rng) for _ in range(60)]
    fams = []
    for _ in range(nfam):
        base = rand_prot(rng, 260, 40)
        fam = [mutate(rng, base, rng.choice([0.05, 0.1, 0.2, 0.3])) for _ in range(famsize)]
        fams.append(fam); prots += fam
    rng.shuffle(prots)
    famprots = [p for fam in fams for p in fam]
    contigs = make_contigs(rng, famprots, nc, cl)
    # genes of two different families overlapping on the same strand (one sits inside the other's intron, or they interleave):
    # these are the calls excludeSameStrandOverlaps has to choose between
    def enc(p):
        return "".join(rng.choice(CODON[x]) for x in p)
    def spacer(a, b):
        return "".join(rng.choices("ACGT", k=rng.randint(a, b)))
    for t in range(8):
        fa, fb = rng.sample(range(nfam), 2)
        pa = mutate(rng, rng.choice(fams[fa]), 0.15); pb = mutate(rng, rng.choice(fams[fb]), rng.choice([0.1, 0.4]))
        h = len(pa) // 2
        if t % 2 == 0:      # B inside A's intron
            gene = enc(pa[:h]) + spacer(60, 200) + "TAA" + enc(pb) + "TAG" + spacer(60, 200) + enc(pa[h:])
        else:               # interleaved: A1 B1 A2 B2
            hb = len(pb) // 2
            gene = enc(pa[:h]) + spacer(50, 120) + enc(pb[:hb]) + spacer(50, 120) + enc(pa[h:]) + spacer(50, 120) + enc(pb[hb:])
        if t % 3 == 0:
            gene = rc(gene)
        contigs.append(spacer(300, 900) + gene + spacer(300, 900))
    write_fasta(tmp + "/proteins.faa", prots, "P")
    write_fasta(tmp + "/contigs.fna", contigs, "C")
    run("%s createdb %s/proteins.faa %s/proteins -v 3" % (REF, tmp, tmp), log)
    run("%s createdb %s/contigs.fna %s/contigs -v 3" % (REF, tmp, tmp), log)
    run("%s extractorfs %s/contigs %s/nucl_6f %s" % (REF, tmp, tmp, ORF_FLAGS.format(thr=thr)), log)
    run("%s translatenucs %s/nucl_6f %s/aa_6f --translation-table 1 --add-orf-stop 0 -v 3 --compressed 0 --threads %d" % (REF, tmp, tmp, thr), log)
    run("%s prefilter %s/aa_6f %s/proteins %s/pref %s" % (REF, tmp, tmp, tmp, PREF_FLAGS.format(mask=1, thr=thr, maxseqs=300)), log)
    run("%s align %s/aa_6f %s/proteins %s/pref %s/aln %s" % (REF, tmp, tmp, tmp, tmp, ALN_FLAGS.format(thr=thr)), log)
    run("%s resultspercontig %s/contigs %s/nucl_6f %s/aln %s/bycontig --threads %d --compressed 0 -v 3" % (REF, tmp, tmp, tmp, tmp, thr), log)
    run("%s collectoptimalset %s/bycontig %s/proteins %s/dp %s" % (REF, tmp, tmp, tmp, DP_FLAGS.format(thr=thr, sets=sets)), log)
    run("%s reduceredundancy %s/dp %s/grouped %s/clusters --overlap 0 --threads %d --compressed 0 -v 3" % (REF, tmp, tmp, tmp, thr), log)
    run("%s reduceredundancy %s/dp %s/grouped_ov %s/clusters_ov --overlap 1 --threads %d --compressed 0 -v 3" % (REF, tmp, tmp, tmp, thr), log)
    out = {"case": name, "kind": "redundancy", "seed": seed, "dp": resdb(tmp + "/dp"), "grouped": resdb(tmp + "/grouped"), "clusters": resdb(tmp + "/clusters"),
           "grouped_overlap_allowed": resdb(tmp + "/grouped_ov"), "ref_version": run("%s version" % REF, log).strip()}
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
        json.dump(out, f)
    shutil.rmtree(tmp)
    print(name, "dp lines", sum(v.count("\n") for v in out["dp"].values()), "grouped lines", sum(v.count("\n") for v in out["grouped"].values()),
          "grouped(overlap ok) lines", sum(v.count("\n") for v in out["grouped_overlap_allowed"].values()), "cluster lines", sum(v.count("\n") for v in out["clusters"].values()))


def case_orfs(name, seed, thr=4):
    """extractorfs + translatenucs only on contigs with the characters the synthetic genes never contain:
    N runs, lower case, IUPAC ambiguity codes, U, and contigs too short to hold a codon / an ORF"""
    rng = random.Random(seed)
    tmp = tempfile.mkdtemp(prefix="mb200_golden_")
    log = []
    contigs = []
    for i in range(10):
        s = list("".join(rng.choices("ACGT", k=rng.randint(1500, 4000))))
        for _ in range(rng.randint(0, 6)):      # N runs
            a = rng.randrange(len(s)); b = min(len(s), a + rng.randint(1, 40))
            s[a:b] = "N" * (b - a)
        for _ in range(rng.randint(0, 6)):      # soft-masked stretches
            a = rng.randrange(len(s)); b = min(len(s), a + rng.randint(1, 200))
            s[a:b] = [x.lower() for x in s[a:b]]
        for _ in range(rng.randint(0, 60)):     # ambiguity codes, U
            s[rng.randrange(len(s))] = rng.choice("RYKMSWBDHVUnrykmswbdhvu")
        contigs.append("".join(s))
    contigs += ["AC", "ACG", "ATGAAATAG", "A" * 44, "ACGTTGCAAGGCTTAACC" * 3, "TAA" * 20, "GCT" * 15 + "TGA", "GCT" * 15, "G" + "GCT" * 16 + "TA"]
    write_fasta(tmp + "/contigs.fna", contigs, "C")
    run("%s createdb %s/contigs.fna %s/contigs --dbtype 2 -v 3" % (REF, tmp, tmp), log)
    run("%s extractorfs %s/contigs %s/nucl_6f %s" % (REF, tmp, tmp, ORF_FLAGS.format(thr=thr)), log)
    run("%s translatenucs %s/nucl_6f %s/aa_6f --translation-table 1 --add-orf-stop 0 -v 3 --compressed 0 --threads %d" % (REF, tmp, tmp, thr), log)
    out = {"case": name, "kind": "orfs", "seed": seed, "contigs": seqdb(tmp + "/contigs"), "queries": seqdb(tmp + "/aa_6f"),
           "orf_headers": {str(k): e[:-1].decode() for k, e in read_db(tmp + "/nucl_6f_h")},
           "ref_version": run("%s version" % REF, log).strip()}
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
        json.dump(out, f)
    shutil.rmtree(tmp)
    print(name, "contigs", len(out["contigs"]), "fragments", len(out["queries"]))


def case_masking(name="mask_repeats", seed=606, thr=8):
    """tantan masking of the targets (Masker::maskSequence) on proteins built to be masked: tandem repeats with unit lengths 1..60
    (beyond the 50-letter window too), 2-12 copies, 0-30 % divergence per copy, low-complexity stretches, X runs, and sequences around the
    16-letter rescaling step; the masked residues are read back from the reference's `indexdb` output"""
    rng = random.Random(seed)
    tmp = tempfile.mkdtemp(prefix="mb200_golden_")
    log = []
    prots = []
    for i in range(420):
        kind = i % 6
        if kind == 0:
            p = rand_prot(rng)
        elif kind in (1, 2, 3):
            p = rand_prot(rng, 200, 80, 30)
            for _ in range(rng.randint(1, 3)):
                unit = "".join(rng.choices(AA, weights=W, k=rng.choice([1, 2, 3, 4, 5, 7, 10, 15, 24, 33, 49, 50, 51, 60])))
                rep = "".join(mutate(rng, unit, rng.choice([0.0, 0.05, 0.1, 0.2, 0.3])) for _ in range(rng.randint(2, 12)))
                a = rng.randrange(len(p) + 1)
                p = p[:a] + rep + p[a:]
        elif kind == 4:
            p = rand_prot(rng, 150, 50, 30)
            a = rng.randrange(len(p))
            p = p[:a] + "".join(rng.choices(rng.sample(AA, rng.randint(2, 3)), k=rng.randint(10, 80))) + p[a:]
            if rng.random() < 0.5:
                b = rng.randrange(len(p))
                p = p[:b] + "X" * rng.randint(1, 12) + p[b:]
        else:
            p = "".join(rng.choices(AA, weights=W, k=rng.choice([1, 5, 15, 16, 17, 31, 32, 33, 48, 49, 50, 51, 52, 64, 2000])))
            if rng.random() < 0.5 and len(p) > 8:
                p = p[:len(p) // 2] + p[:len(p) // 2]
        prots.append(p)
    write_fasta(tmp + "/proteins.faa", prots, "P")
    run("%s createdb %s/proteins.faa %s/proteins --shuffle 0 -v 3" % (REF, tmp, tmp), log)
    out = {"case": name, "kind": "masking", "seed": seed, "targets": seqdb(tmp + "/proteins")}
    idx = ref_index(tmp, "proteins", 1, thr, log)
    out["masked_targets"] = idx["masked_targets"]; out["masked_residues"] = idx["idx_masked_residues"]
    out["ref_version"] = run("%s version" % REF, log).strip()
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
        json.dump(out, f)
    shutil.rmtree(tmp)
    print(name, "targets", len(out["targets"]), "residues", sum(len(s) for _, s in out["targets"]), "masked", out["masked_residues"])


def case_masking_flags(name="mask_flags", seed=707, thr=8):
    """the other switches of Masker::maskSequence (--mask-lower-case, --mask-n-repeat, with and without tantan) on proteins with soft-masked (lower-case) stretches,
    runs of 2..9 equal letters anywhere -- including runs of A at the very start of a sequence, which the reference's run scan cannot mask -- X runs, and
    sequences that are one single run; the masked residues are read back from the reference's `indexdb` output per variant"""
    rng = random.Random(seed)
    tmp = tempfile.mkdtemp(prefix="mb200_golden_")
    log = []
    prots = []
    for i in range(160):
        p = list(rand_prot(rng, 160, 60, 20))
        for _ in range(rng.randint(0, 4)):
            a = rng.randrange(len(p)); p[a:a] = list(rng.choice(AA + "AAX") * rng.randint(2, 9))
        if i % 5 == 0:
            p[0:0] = list("A" * rng.randint(1, 7))
        if i % 7 == 0:
            p[0:0] = list(rng.choice("CDW") * rng.randint(2, 6))
        for _ in range(rng.randint(0, 3)):
            a = rng.randrange(len(p)); b = min(len(p), a + rng.randint(1, 40)); p[a:b] = [c.lower() for c in p[a:b]]
        prots.append("".join(p))
    prots += ["A" * 12, "AAAAc", "aaaa" + rand_prot(rng, 40, 5), "W" * 9, "ACAC", "x" * 8 + "KLMN" * 5, "MKKKKKKKKKK", "M"]
    write_fasta(tmp + "/proteins.faa", prots, "P")
    run("%s createdb %s/proteins.faa %s/proteins --shuffle 0 -v 3" % (REF, tmp, tmp), log)
    out = {"case": name, "kind": "masking_flags", "seed": seed, "targets": seqdb(tmp + "/proteins"), "variants": {}}
    for vn, mask, lower, nrep in (("tantan+lower", 1, 1, 0), ("lower+nrep3", 0, 1, 3), ("tantan+nrep2", 1, 0, 2), ("nrep1", 0, 0, 1), ("all", 1, 1, 4)):
        idx = ref_index(tmp, "proteins", mask, thr, log, {"--mask-lower-case": lower, "--mask-n-repeat": nrep})
        out["variants"][vn] = {"mask": mask, "mask_lower_case": lower, "mask_n_repeat": nrep, "masked_targets": idx["masked_targets"], "masked_residues": idx["idx_masked_residues"]}
        print(vn, "masked residues", idx["idx_masked_residues"])
    out["ref_version"] = run("%s version" % REF, log).strip()
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
        json.dump(out, f)
    shutil.rmtree(tmp)


REF_TESTS = [("minus_strand", -1, 0), ("multi_exon", -1, 0), ("two_contigs", -1, 0), ("target_overlap", -3, 0), ("cluster_rep", -3, 0), ("target_cov", -1, 0.7)]


def case_reference_tests(name="ref_tests", thr=8):
    """The reference's OWN end-to-end fixtures (reference/tests/run.sh -> test.sh: six data sets with the flags listed there).  For each one:
    the input FASTA files as the reference's createdb stored them, the `predictexons` workflow's intermediate DBs (kept with
    --remove-tmp-files 0), `reduceredundancy`'s outputs, the tantan-masked targets, and -- verbatim -- the header lines of the checked-in
    known answers as_should_final_united_exons_aa.fas / as_should_final_grouped_predictions_rep.fas (target|contig|strand|bit score|E-value|
    #exons|low|high|exon coordinates...), which tests/ rebuild from the call records (unitesetstofasta.cpp:114-190)."""
    tests_dir = os.environ.get("METAEUK_REF_TESTS", "/root/reference/tests")
    out = {"case": name, "kind": "reference_tests", "sets": {}}
    for ds, sge, tcov in REF_TESTS:
        tmp = tempfile.mkdtemp(prefix="mb200_golden_")
        log = []
        src = os.path.join(tests_dir, ds)
        run("%s createdb %s/contigs.fna %s/contigs --dbtype 2 -v 3" % (REF, src, tmp), log)
        run("%s createdb %s/proteins.faa %s/proteins --dbtype 1 -v 3" % (REF, src, tmp), log)
        wf = run("%s predictexons %s/contigs %s/proteins %s/predEx %s/tmp --set-gap-extend %s --metaeuk-tcov %s --remove-tmp-files 0 --threads %d -v 3"
                 % (REF, tmp, tmp, tmp, tmp, sge, tcov, thr), log)
        run("%s reduceredundancy %s/predEx %s/predRed %s/predClust --threads %d -v 3" % (REF, tmp, tmp, tmp, thr), log)
        lt = tmp + "/tmp/latest"
        heads = {}
        for f in ("as_should_final_united_exons_aa.fas", "as_should_final_grouped_predictions_rep.fas"):
            heads[f] = [l[1:].rstrip("\n") for l in open(os.path.join(src, f)) if l.startswith(">")]
        d = {"set_gap_extend": sge, "tcov": tcov, "targets": seqdb(tmp + "/proteins"), "contigs": seqdb(tmp + "/contigs"),
             "target_acc": {str(k): e.decode().split()[0] for k, e in read_db(tmp + "/proteins_h")},
             "contig_acc": {str(k): e.decode().split()[0] for k, e in read_db(tmp + "/contigs_h")},
             "queries": seqdb(lt + "/aa_6f"), "orf_headers": {str(k): e[:-1].decode() for k, e in read_db(lt + "/nucl_6f_h")},
             "aln": resdb(lt + "/search_res"), "bycontig": resdb(lt + "/search_res_by_contig"), "dp": resdb(tmp + "/predEx"),
             "grouped": resdb(tmp + "/predRed"), "clusters": resdb(tmp + "/predClust"),
             "kmer_thr": int(wf.split("k-mer similarity threshold: ")[1].split()[0]),
             "masked_residues": int(wf.split("Masked residues: ")[1].split()[0]), "as_should": heads}
        idx = ref_index(tmp, "proteins", 1, thr, log)
        assert idx["idx_masked_residues"] == d["masked_residues"]
        d["masked_targets"] = idx["masked_targets"]
        out["sets"][ds] = d
        shutil.rmtree(tmp)
        print(ds, "targets", len(d["targets"]), "contigs", len(d["contigs"]), "fragments", len(d["queries"]), "aln lines", sum(v.count("\n") for v in d["aln"].values()),
              "dp lines", sum(v.count("\n") for v in d["dp"].values()), "known answers", len(heads["as_should_final_united_exons_aa.fas"]), len(heads["as_should_final_grouped_predictions_rep.fas"]))
    out["ref_version"] = run("%s version" % REF, []).strip()
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
        json.dump(out, f)

FLAG_VARIANTS = [
    # name, flags handed to `metaeuk predictexons` on top of its defaults
    ("default", ""),
    ("s6", "-s 6"),
    ("s1", "-s 1"),
    ("s7.5", "-s 7.5"),
    ("kscore", "--k-score seq:140,prof:140"),
    ("maxseqs10", "--max-seqs 10"),
    ("eval", "-e 0.001"),
    ("minexon", "--min-exon-aa 20 --min-length 30"),
    ("maxlen", "--max-length 120"),
    ("nocomp", "--comp-bias-corr 0"),
    ("reverse", "--reverse-fragments 1"),
    ("table4", "--translation-table 4"),
    ("table6", "--translation-table 6"),
    ("nomask", "--mask 0"),
    ("minungapped", "--min-ungapped-score 30"),
    ("dp", "--max-intron 400 --min-intron 60 --max-overlap 4 --set-gap-open -3 --set-gap-extend -2 --metaeuk-eval 1e-6 --metaeuk-tcov 0.3 --max-exon-sets 2"),
]


def case_flags(name="flags", seed=808, thr=8):
    """One small contig set through the reference's `predictexons` WORKFLOW (data/predictexons.sh) once per entry of FLAG_VARIANTS: the flags a user of
    `metaeuk predictexons` can change on this path (sensitivity -> k-mer threshold, list length, E-value, ORF length limits, genetic code, composition bias off,
    reversed fragments = the null model, masking off, exon-chaining parameters).  Stored per variant: the k-mer threshold the reference logged, the six-frame
    fragments and their headers (only if they differ from the default run), search_res and the calls."""
    rng = random.Random(seed)
    prots = make_proteins(rng, 150, family=40, lowcomplex=4)
    contigs = make_contigs(rng, prots, 4, 8000)
    out = {"case": name, "kind": "flags", "seed": seed, "variants": {}}
    base_q = None
    for vn, flags in FLAG_VARIANTS:
        tmp = tempfile.mkdtemp(prefix="mb200_golden_")
        log = []
        write_fasta(tmp + "/proteins.faa", prots, "P")
        write_fasta(tmp + "/contigs.fna", contigs, "C")
        run("%s createdb %s/proteins.faa %s/proteins -v 3" % (REF, tmp, tmp), log)
        run("%s createdb %s/contigs.fna %s/contigs -v 3" % (REF, tmp, tmp), log)
        wf = run("%s predictexons %s/contigs %s/proteins %s/predEx %s/tmp %s --remove-tmp-files 0 --threads %d -v 3" % (REF, tmp, tmp, tmp, tmp, flags, thr), log)
        lt = tmp + "/tmp/latest"
        aa = "/aa_6f_reverse" if "--reverse-fragments 1" in flags else "/aa_6f"
        d = {"flags": flags, "kmer_thr": int(wf.split("k-mer similarity threshold: ")[1].split()[0]),
             "aln": {k: t for k, t in resdb(lt + "/search_res").items() if t}, "dp": resdb(tmp + "/predEx")}          # (only the non-empty search_res entries)
        q = {"queries": sorted(seqdb(lt + aa)), "orf_headers": {str(k): e[:-1].decode() for k, e in read_db(lt + "/nucl_6f_h")}}
        if aa.endswith("_reverse"):      # reverseseq = every fragment back to front, headers unchanged: stored as a flag
            assert q["orf_headers"] == base_q["orf_headers"] and q["queries"] == [(k, t[::-1]) for k, t in base_q["queries"]]
            q = base_q; d["reversed"] = True
        if vn == "default":
            base_q = q
            out["targets"] = seqdb(tmp + "/proteins"); out["contigs"] = seqdb(tmp + "/contigs")
            out["masked_targets"] = ref_index(tmp, "proteins", 1, thr, log)["masked_targets"]
            d.update(q)
        elif q != base_q:
            d.update(q)
        out["variants"][vn] = d
        shutil.rmtree(tmp)
        print(vn, "kmer_thr", d["kmer_thr"], "fragments", len(q["queries"]), "aln lines", sum(v.count("\n") for v in d["aln"].values()),
              "dp lines", sum(v.count("\n") for v in d["dp"].values()), "own fragments" if "queries" in d and vn != "default" else "")
    out["ref_version"] = run("%s version" % REF, []).strip()
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
        json.dump(out, f)


if __name__ == "__main__":
    if not os.path.exists(REF):
        raise SystemExit("reference binary not found: " + REF)
    which = sys.argv[1:] or ["frag_small", "frag_family", "frag_nomask", "contigs_small", "contigs_sets"]
    if "frag_small" in which:
        case_fragments("frag_small", 4321, nprot=600, nfrag=400)
    if "frag_family" in which:   # >300 homologous targets: exercises the top-300 cut and its tie order
        case_fragments("frag_family", 99, nprot=300, nfrag=60, family=700)
    if "frag_nomask" in which:   # low-complexity targets with --mask 0
        case_fragments("frag_nomask", 7, nprot=300, nfrag=200, lowcomplex=30, mask=0)
    if "frag_adversarial" in which:
        case_fragments("frag_adversarial", 515, 0, 0, adversarial=True)
    if "contigs_small" in which:
        case_contigs("contigs_small", 1234, nprot=400, nc=24, cl=9000)
    if "contigs_sets" in which:
        case_contigs("contigs_sets", 77, nprot=200, nc=12, cl=9000, sets=3)
    if "orf_edge" in which:
        case_orfs("orf_edge", 31)
    if "redundancy_family" in which:
        case_redundancy("redundancy_family", 2025)
    if "ref_tests" in which:
        case_reference_tests()
    if "mask_repeats" in which:
        case_masking()
    if "flags" in which:
        case_flags()
    if "mask_flags" in which:
        case_masking_flags()
