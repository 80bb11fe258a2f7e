"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/metaeuk_b200.h declares.
No compute calls (no GPU here); creating a context without a device must fail loudly, not fall back."""
import ctypes as C
import os
import re

import pytest

import metaeuk_b200
from metaeuk_b200 import lib as mlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
    mlib.build_library()
    return mlib.load_library()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "metaeuk_b200.h")).read()
    return sorted(set(re.findall(r"\b(mb200_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(so):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(so, n), n


def test_params_default_are_metaeuk_defaults(so):
    p = mlib.Params()
    so.mb200_params_default(C.byref(p))
    assert (p.kmer_thr, p.max_res_list_len, p.min_diag_score, p.bin_size) == (127, 300, 15, 2)
    assert (p.gap_open, p.gap_extend, p.min_aln_len, p.eval_thr) == (11, 1, 11, 100.0)
    assert (p.min_intron, p.max_intron, p.min_exon_aa, p.max_overlap, p.set_gap_open, p.set_gap_extend, p.max_exon_sets) == (15, 10000, 11, 10, -1, -1, 1)
    assert (p.tcov, p.metaeuk_eval) == (0.5, 0.001)


def test_struct_layouts_match_header():
    assert mlib.HIT_DTYPE.itemsize == 12 and mlib.ALN_DTYPE.itemsize == 56 and mlib.CAND_DTYPE.itemsize == 40
    assert mlib.EXON_DTYPE.itemsize == 16 and mlib.CALL_DTYPE.itemsize == 40
    assert mlib.ALN_DTYPE.fields["evalue"][1] == 48 and mlib.CALL_DTYPE.fields["evalue"][1] == 32


def test_no_cpu_fallback_without_device(so):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(metaeuk_b200.Mb200Error):
        metaeuk_b200.Context(0)
