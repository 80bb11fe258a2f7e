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
