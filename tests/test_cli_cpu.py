"""CPU-only: the parts of the C++ driver (metaeuk_b200/bin/mb200) that need no GPU -- `gpucreateindex` builds the target k-mer index on the
host and persists it in the reference's index-DB layout; every command that needs the device must fail loudly here."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

import reflib
from metaeuk_b200 import lib as mlib
from test_kernel_logic_emulated import _open_indexdb, emu  # noqa: F401  (fixture)


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as g
    p = os.path.join(mlib.PKG, "bin", "mb200")
    srcs = [os.path.join(mlib.PKG, "csrc", "mb200_cli.cpp"), os.path.join(reflib.ROOT, "include", "metaeuk_b200.h")]
    stale = os.path.exists(p) and any(os.path.getmtime(f) > os.path.getmtime(p) for f in srcs)     # never test a driver older than its sources
    if stale:
        os.remove(p)
    if not os.path.exists(p) or not os.path.exists(mlib.LIB_PATH):
        g.build()
    return p


def test_gpucreateindex_writes_the_reference_index_entries(exe, emu, tmp_path):  # noqa: F811
    c = reflib.load_case("frag_small")
    d = str(tmp_path)
    reflib.write_db(d + "/proteins", [(k, s + "\n") for k, s in c["targets"]], 0)
    reflib.write_db(d + "/proteins_masked", [(k, m + "\n") for (k, _), m in zip(c["targets"], c["masked_targets"])], 0)
    subprocess.check_call([exe, "gpucreateindex", d + "/proteins", d + "/proteins.idx", "--masked-db", d + "/proteins_masked", "--k-score", str(c["kmer_thr"])])
    x, err = _open_indexdb(emu, d + "/proteins.idx")
    assert x is not None, err
    assert x["n_entries"] == c["idx_n_entries"]
    assert hashlib.sha256(x["entries6"].tobytes()).hexdigest() == c["idx_entries_sha256"]
    assert hashlib.sha256(x["offsets"].tobytes()).hexdigest() == c["idx_offsets_u64_sha256"]
    assert list(x["keys"]) == [k for k, _ in c["targets"]]
    m, _ = reflib.pack(c["masked_targets"]); u, _ = reflib.pack([s for _, s in c["targets"]])
    assert (x["masked"] == m).all() and (x["unmasked"] == u).all()


def test_device_commands_fail_loudly_without_a_gpu(exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = reflib.load_case("frag_small")
    d = str(tmp_path)
    reflib.write_db(d + "/proteins", [(k, s + "\n") for k, s in c["targets"][:20]], 0)
    reflib.write_db(d + "/q", [(k, s + "\n") for k, s in c["queries"][:5]], 0)
    r = subprocess.run([exe, "gpuprefilter", d + "/q", d + "/proteins", d + "/pref"], capture_output=True, text=True)
    assert r.returncode != 0 and not os.path.exists(d + "/pref")
    assert "mb200" in (r.stderr + r.stdout)


@pytest.mark.parametrize("name", ["mask_repeats", "frag_adversarial", "frag_small", "frag_family", "contigs_small", "contigs_sets"])
def test_product_target_masking_matches_reference(name):
    """mb200_mask_targets (host only, product code in host_tables.cpp) against the masked residues the reference stored in its index"""
    c = reflib.load_case(name)
    u, off = reflib.pack([s for _, s in c["targets"]]); m, _ = reflib.pack(c["masked_targets"])
    for width in (4, 2):
        got, n = mlib.mask_targets(u, off, 0.9, width)
        assert n == c["masked_residues"] and (got == m).all(), width
    with pytest.raises(mlib.Mb200Error):
        mlib.mask_targets(np.array([3, 21, 4], np.uint8), np.array([0, 3], np.uint64))


def test_gpucreateindex_masks_like_the_reference_by_default(exe, emu, tmp_path):  # noqa: F811
    """no --masked-db: the driver masks the targets itself (--mask 1, the reference's default) and must arrive at the reference's index"""
    c = reflib.load_case("contigs_small")
    d = str(tmp_path)
    reflib.write_db(d + "/proteins", [(k, s + "\n") for k, s in c["targets"]], 0)
    subprocess.check_call([exe, "gpucreateindex", d + "/proteins", d + "/proteins.idx", "--k-score", str(c["kmer_thr"])])
    x, err = _open_indexdb(emu, d + "/proteins.idx")
    assert x is not None, err
    m, _ = reflib.pack(c["masked_targets"])
    assert (x["masked"] == m).all() and x["meta"][4] == 1
    assert hashlib.sha256(x["entries6"].tobytes()).hexdigest() == c["idx_entries_sha256"]
    subprocess.check_call([exe, "gpucreateindex", d + "/proteins", d + "/unmasked.idx", "--k-score", str(c["kmer_thr"]), "--mask", "0"])
    y, err = _open_indexdb(emu, d + "/unmasked.idx")
    u, _ = reflib.pack([s for _, s in c["targets"]])
    assert (y["masked"] == u).all() and y["meta"][4] == 0 and y["n_entries"] > x["n_entries"]


def test_the_reference_prefilters_and_aligns_from_an_index_written_here(exe, tmp_path):
    """SURVEY 8f-4, the other direction: `metaeuk prefilter` / `align` of the UNMODIFIED reference binary run from a `.idx` that `mb200 gpucreateindex` wrote (score-matrix name,
    spaced pattern, 2-mer and 3-mer similarity tables in ScoreMatrix::serialize layout besides the k-mer index) and give the golden result DBs; build container only (oracle/_ref)"""
    import sys
    sys.path.insert(0, os.path.join(reflib.ROOT, "tools"))
    import refrun
    ref = refrun.ref_bin()
    if not ref:
        pytest.skip("oracle/_ref/metaeuk-ref not built (needs /root/reference: build container)")
    c = reflib.load_case("frag_small")
    d = str(tmp_path)
    reflib.write_db(d + "/proteins", [(k, s + "\n") for k, s in c["targets"]], 0)
    reflib.write_db(d + "/q", [(k, s + "\n") for k, s in c["queries"]], 0)
    subprocess.check_call([exe, "gpucreateindex", d + "/proteins", d + "/proteins.idx", "--k-score", str(c["kmer_thr"])])
    pref = ("--sub-mat aa:blosum62.out,nucl:nucleotide.out --seed-sub-mat aa:VTML80.out,nucl:nucleotide.out -k 0 --target-search-mode 0 --k-score seq:%d,prof:%d "
            "--alph-size aa:21,nucl:5 --max-seq-len 65535 --max-seqs %d --split 1 --split-mode 2 --split-memory-limit 0 -c 0 --cov-mode 0 --comp-bias-corr 1 --comp-bias-corr-scale 1 "
            "--diag-score 1 --exact-kmer-matching 0 --mask 1 --mask-prob 0.9 --mask-lower-case 0 --mask-n-repeat 0 --min-ungapped-score 15 --add-self-matches 0 --spaced-kmer-mode 1 "
            "--db-load-mode 0 --threads 4 --compressed 0 -v 3 -s 4.0") % (c["kmer_thr"], c["kmer_thr"], c.get("max_seqs", 300))
    r = subprocess.run([ref, "prefilter", d + "/q", d + "/proteins.idx", d + "/pref"] + pref.split(), capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "Generated by:  metaeuk_b200" in r.stdout and "ScoreMatrix:  VTML80.out" in r.stdout      # it did read our DB
    keys, ents = refrun.read_result_db(d + "/pref")
    got = {str(int(k)): e.decode() for k, e in zip(keys, ents)}
    bad = [k for k, _ in c["queries"] if got.get(str(k), "") != c["pref"].get(str(k), "")]
    assert not bad, bad[:5]
    aln = ("--sub-mat aa:blosum62.out,nucl:nucleotide.out -a 0 --alignment-mode 2 --alignment-output-mode 0 --wrapped-scoring 0 -e 100 --min-seq-id 0 --min-aln-len 11 --seq-id-mode 0 "
           "--alt-ali 0 -c 0 --cov-mode 0 --max-seq-len 65535 --comp-bias-corr 1 --comp-bias-corr-scale 1 --max-rejected 2147483647 --max-accept 2147483647 --add-self-matches 0 "
           "--db-load-mode 0 --score-bias 0 --realign 0 --gap-open aa:11,nucl:5 --gap-extend aa:1,nucl:2 --zdrop 40 --threads 4 --compressed 0 -v 3")
    r = subprocess.run([ref, "align", d + "/q", d + "/proteins.idx", d + "/pref", d + "/aln"] + aln.split(), capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    keys, ents = refrun.read_result_db(d + "/aln")
    got = {str(int(k)): e.decode() for k, e in zip(keys, ents)}
    bad = [k for k, _ in c["queries"] if got.get(str(k), "") != c["aln"].get(str(k), "")]
    assert not bad, bad[:5]


def test_sensitivity_to_kmer_threshold_matches_the_reference_log():
    """`-s` -> k-mer similarity threshold (mb200_kmer_threshold = Prefiltering::getKmerThreshold, k = 6): the values the reference logged for the
    flags fixture and the whole range of sensitivities in steps of 0.1 against the restated float/double expression"""
    L = mlib.load_library()
    L.mb200_kmer_threshold.argtypes = [C.c_float]
    v = reflib.load_case("flags")["variants"]
    assert L.mb200_kmer_threshold(4.0) == v["default"]["kmer_thr"] == 127
    assert L.mb200_kmer_threshold(6.0) == v["s6"]["kmer_thr"] and L.mb200_kmer_threshold(1.0) == v["s1"]["kmer_thr"] and L.mb200_kmer_threshold(7.5) == v["s7.5"]["kmer_thr"]
    for s in np.arange(1.0, 9.05, 0.1):
        assert L.mb200_kmer_threshold(float(s)) == reflib.sensitivity_to_kmer_thr(float(s)), s


def test_cli_rejects_flags_it_would_otherwise_silently_ignore(exe, emu, tmp_path):  # noqa: F811
    """a reference flag that changes results but is not implemented must stop the run (unless it carries the value `metaeuk predictexons` passes
    itself); flags that never change results are accepted; unknown flags are errors; -s reaches the index builder"""
    c = reflib.load_case("frag_small")
    d = str(tmp_path)
    reflib.write_db(d + "/proteins", [(k, s + "\n") for k, s in c["targets"][:50]], 0)
    base = [exe, "gpucreateindex", d + "/proteins"]
    r = subprocess.run([exe, "gpuprefilter", d + "/proteins", d + "/proteins", d + "/x", "--gpus", "2"], capture_output=True, text=True)     # --gpus is a gpupredictexons option
    assert r.returncode != 0 and "--gpus" in r.stderr
    for bad in (["--alignment-mode", "3"], ["--orf-start-mode", "0"], ["--min-seq-id", "0.3"], ["-c", "0.8"], ["--comp-bias-corr-scale", "0.5"], ["--gap-open", "10"],
                ["--no-such-flag", "1"], ["--exhaustive-search", "1"], ["--sub-mat", "VTML40.out"]):
        r = subprocess.run(base + [d + "/bad.idx"] + bad, capture_output=True, text=True)
        assert r.returncode != 0 and bad[0] in r.stderr and not os.path.exists(d + "/bad.idx"), bad
    ok = ["--threads", "8", "-v", "3", "--compressed", "0", "--alignment-mode", "2", "--gap-open", "aa:11,nucl:5", "--gap-extend", "aa:1,nucl:2", "--sub-mat",
          "aa:blosum62.out,nucl:nucleotide.out", "--seed-sub-mat", "aa:VTML80.out,nucl:nucleotide.out", "-k", "0", "--split", "1", "--orf-start-mode", "1",
          "--k-score", "seq:2147483647,prof:2147483647", "--comp-bias-corr-scale", "1", "--max-seq-len", "65535", "--forward-frames", "1,2,3"]
    subprocess.check_call(base + [d + "/s4.idx"] + ok)
    subprocess.check_call(base + [d + "/s6.idx", "-s", "6"])
    subprocess.check_call(base + [d + "/k109.idx", "--k-score", "seq:109,prof:109", "-s", "2"])
    a, _ = _open_indexdb(emu, d + "/s4.idx"); b, _ = _open_indexdb(emu, d + "/s6.idx"); k, _ = _open_indexdb(emu, d + "/k109.idx")
    assert a["meta"][6] == 127 and b["meta"][6] == 109 and k["meta"][6] == 109      # (the k-mer threshold is part of the index meta data)


def test_predictexons_flags_follow_the_drivers_rules():
    """modules.predictexons_flags (Python mirror of the driver's option rules) on the flag sets of the reference fixture"""
    from metaeuk_b200 import modules as M
    c = reflib.load_case("flags")
    for name, v in c["variants"].items():
        par, orf, mask = M.predictexons_flags(v["flags"])
        assert par["kmer_thr"] == v["kmer_thr"], name
        assert mask == (0 if name == "nomask" else 1)
    par, orf, _ = M.predictexons_flags("--min-exon-aa 20 --min-length 30 --threads 4")
    assert par["min_aln_len"] == 20 and par["min_exon_aa"] == 20 and orf == {"min_len": 30}
    assert M.predictexons_flags("-s 2 --k-score seq:109,prof:109")[0]["kmer_thr"] == 109
    with pytest.raises(ValueError):
        M.predictexons_flags("--orf-start-mode 0")


@pytest.mark.parametrize("variant", ["tantan+lower", "lower+nrep3", "tantan+nrep2", "nrep1", "all"])
def test_product_masking_switches_match_reference(variant):
    """mb200_mask_targets_ex (product host code) on the reference's --mask-lower-case / --mask-n-repeat fixture"""
    from test_oracle_golden import mask_flags_inputs
    c = reflib.load_case("mask_flags"); v = c["variants"][variant]
    u, off, lower = mask_flags_inputs(c)
    want, _ = reflib.pack(v["masked_targets"])
    for width in (4, 2):
        got, n = mlib.mask_targets_ex(u, off, bool(v["mask"]), 0.9, lower if v["mask_lower_case"] else None, v["mask_n_repeat"], width)
        assert (got == want).all() and n == v["masked_residues"], width


def test_gpucreateindex_takes_the_masking_switches(exe, emu, tmp_path):  # noqa: F811
    """--mask-lower-case / --mask-n-repeat reach the index builder of the driver: the residues stored in the index equal the reference's"""
    c = reflib.load_case("mask_flags")
    d = str(tmp_path)
    reflib.write_db(d + "/proteins", [(k, s + "\n") for k, s in c["targets"]], 0)
    for name, v in c["variants"].items():
        out = d + "/" + name.replace("+", "_") + ".idx"
        subprocess.check_call([exe, "gpucreateindex", d + "/proteins", out, "--mask", str(v["mask"]), "--mask-lower-case", str(v["mask_lower_case"]),
                               "--mask-n-repeat", str(v["mask_n_repeat"])])
        x, err = _open_indexdb(emu, out)
        assert x is not None, err
        want, _ = reflib.pack(v["masked_targets"])
        assert (x["masked"] == want).all() and x["meta"][4] == v["mask"], name
