#!/usr/bin/env python3
"""Differential run of the CPU oracle against the UNMODIFIED reference on fresh random inputs -- TEST INFRASTRUCTURE, build container only
(needs the reference binary, env METAEUK_REF; nothing here runs on the GPU box, nothing is written to the repo).

Every round draws a seed, builds a small protein set (random proteins, a family, low-complexity members) and contigs with planted multi-exon genes
(make_golden.py's generators), picks a random set of `metaeuk predictexons` flags from the ones this path implements, runs the reference workflow and the
oracle pipeline (tests/reflib.py: ORFs -> prefilter -> align -> join -> exon chaining; targets masked by the oracle's tantan restatement, so masking is
part of the comparison) and compares the fragments, search_res and the calls byte for byte.

usage: fuzz_oracle_vs_reference.py [rounds=20] [first_seed=1]"""
import os
import random
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import make_golden as G  # noqa: E402
import reflib            # noqa: E402
import numpy as np       # noqa: E402
import ctypes as C       # noqa: E402


def oracle_mask(seqs):
    L = C.CDLL(reflib.ORACLE_SO)
    L.mbo_mask_targets.restype = C.c_longlong
    u, off = reflib.pack(seqs)
    out = np.zeros_like(u)
    L.mbo_mask_targets(reflib._p(u, C.c_uint8), reflib._p(off, C.c_uint64), C.c_uint32(len(off) - 1), C.c_double(0.9), reflib._p(out, C.c_uint8), 4, 0)
    return ["".join(G.NUM2AA[c] for c in out[int(off[i]):int(off[i + 1])]) for i in range(len(seqs))]


def draw_flags(rng):
    """-> (flag string for the reference, oracle parameters, pipeline options)"""
    f, par, opt = [], {}, {}
    if rng.random() < 0.5:
        s = rng.choice([1, 2, 3, 4.5, 5, 5.7, 6, 7])
        f.append("-s %s" % s); par["kmerThr"] = reflib.sensitivity_to_kmer_thr(s)
    if rng.random() < 0.3:
        n = rng.choice([1, 5, 30, 100]); f.append("--max-seqs %d" % n); par["maxResListLen"] = n
    if rng.random() < 0.3:
        e = rng.choice([10, 1, 1e-3, 1e-10]); f.append("-e %g" % e); par["evalThr"] = e
    if rng.random() < 0.2:
        n = rng.choice([20, 40, 60]); f.append("--min-ungapped-score %d" % n); par["minDiagScore"] = n
    if rng.random() < 0.3:
        n = rng.choice([12, 18, 25]); f.append("--min-exon-aa %d" % n); par["minAlnLen"] = n; opt["min_exon_aa"] = n
    if rng.random() < 0.3:
        n = rng.choice([8, 20, 40]); f.append("--min-length %d" % n); opt["min_len"] = n
    if rng.random() < 0.2:
        n = rng.choice([100, 300]); f.append("--max-length %d" % n); opt["max_len"] = n
    if rng.random() < 0.2:
        f.append("--comp-bias-corr 0"); par["comp_bias_corr"] = False
    if rng.random() < 0.15:
        f.append("--reverse-fragments 1"); opt["reverse_fragments"] = True
    if rng.random() < 0.3:
        t = rng.choice([2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31]); f.append("--translation-table %d" % t); opt["table"] = t
    if rng.random() < 0.2:
        f.append("--mask 0"); par["nomask"] = True
    if rng.random() < 0.4:
        a, b = rng.choice([(15, 10000), (30, 600), (5, 300)]); f.append("--min-intron %d --max-intron %d" % (a, b)); opt["min_intron"] = a; opt["max_intron"] = b
    if rng.random() < 0.3:
        n = rng.choice([0, 3, 10]); f.append("--max-overlap %d" % n); opt["max_overlap"] = n
    if rng.random() < 0.3:
        a, b = rng.choice([(-1, -1), (-2, -1), (-1, -3), (-5, -2)]); f.append("--set-gap-open %d --set-gap-extend %d" % (a, b)); opt["gap_open"] = a; opt["gap_extend"] = b
    if rng.random() < 0.3:
        e = rng.choice([1e-2, 1e-5, 1]); f.append("--metaeuk-eval %g" % e); opt["eval_thr"] = e
    if rng.random() < 0.3:
        t = rng.choice([0, 0.3, 0.8]); f.append("--metaeuk-tcov %g" % t); opt["tcov"] = t
    if rng.random() < 0.3:
        n = rng.choice([2, 3]); f.append("--max-exon-sets %d" % n); opt["max_sets"] = n
    return " ".join(f), par, opt


def one_round(seed, o):
    rng = random.Random(seed)
    prots = G.make_proteins(rng, rng.randint(40, 160), family=rng.choice([0, 10, 30]), lowcomplex=rng.choice([0, 3, 8]))
    contigs = G.make_contigs(rng, prots, rng.randint(2, 5), rng.choice([4000, 9000]))
    flags, par, opt = draw_flags(rng)
    tmp = tempfile.mkdtemp(prefix="mb200_fuzz_")
    log = []
    try:
        G.write_fasta(tmp + "/proteins.faa", prots, "P"); G.write_fasta(tmp + "/contigs.fna", contigs, "C")
        G.run("%s createdb %s/proteins.faa %s/proteins -v 3" % (G.REF, tmp, tmp), log)
        G.run("%s createdb %s/contigs.fna %s/contigs -v 3" % (G.REF, tmp, tmp), log)
        try:
            G.run("%s predictexons %s/contigs %s/proteins %s/predEx %s/tmp %s --remove-tmp-files 0 --threads 4 -v 3" % (G.REF, tmp, tmp, tmp, tmp, flags), log, quiet=True)
        except SystemExit:      # (the reference's own workflow dies on some inputs, e.g. when a stage produces an empty DB: nothing to compare)
            print("seed %d: flags [%s] -> the reference workflow failed, skipped" % (seed, flags), flush=True)
            return True
        lt = tmp + "/tmp/latest"
        want_q = dict(G.seqdb(lt + ("/aa_6f_reverse" if opt.get("reverse_fragments") else "/aa_6f")))
        want_aln = {k: t for k, t in G.resdb(lt + "/search_res").items() if t}
        want_dp = G.resdb(tmp + "/predEx")
        targets = G.seqdb(tmp + "/proteins"); cdb = G.seqdb(tmp + "/contigs")
    finally:
        shutil.rmtree(tmp)
    par = dict(par); nomask = par.pop("nomask", False); comp = par.pop("comp_bias_corr", True)
    keys = [k for k, _ in targets]; unmasked = [s for _, s in targets]
    o.set_params(**par); o.set_comp_bias_corr(comp)
    o.set_targets(unmasked if nomask else oracle_mask(unmasked), unmasked)
    r = reflib.oracle_pipeline(o, cdb, keys, [len(s) for s in unmasked], **opt)
    ok = (r["queries"] == want_q, {str(k): t for k, t in r["aln"].items() if t} == want_aln, {str(k): t for k, t in r["dp"].items()} == want_dp)
    print("seed %d: %d proteins, %d contigs, %d fragments, %d aln lines, %d call lines, flags [%s] -> %s" % (
        seed, len(prots), len(contigs), len(want_q), sum(t.count("\n") for t in want_aln.values()), sum(t.count("\n") for t in want_dp.values()), flags,
        "identical" if all(ok) else "MISMATCH fragments/aln/calls=%s" % (ok,)), flush=True)
    return all(ok)


if __name__ == "__main__":
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    if not os.path.exists(G.REF):
        raise SystemExit("reference binary not found: " + G.REF)
    o = reflib.Oracle()
    bad = [s for s in range(first, first + rounds) if not one_round(s, o)]
    print("%d rounds, %d mismatches %s" % (rounds, len(bad), bad))
    sys.exit(1 if bad else 0)
