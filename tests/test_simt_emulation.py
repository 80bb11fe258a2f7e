"""CPU-only: a slice of the GPU suite executed on tests/simt/libmb200_simt.so -- the product's CUDA translation unit (kernels AND the host orchestration in
metaeuk_b200/csrc/mb200.cu) compiled by g++ against a stand-in <cuda_runtime.h> that runs every thread block as cooperative fibers and restates the DPX / PRMT
instructions in C.  TEST INFRASTRUCTURE: it exists so that kernel and launch logic can be diffed against the reference fixtures in a container without a GPU (the
serial emulation in tests/emu covers the prefilter kernel only); it is never shipped, never measured and never a fallback -- the library refuses nothing here, but
`mb200_build_kind()` says what it is and bench.py / the loader only ever use the nvcc build.

The whole GPU suite passes on it (about 20 minutes of CPU time: `MB200_LIB_PATH=tests/simt/libmb200_simt.so pytest -m gpu -k "not cli"`); the slice below is sized for the
CPU suite: every forward alignment layout + the start-position kernel, the ORF kernels with a non-default genetic code, composition bias off, the exon-chaining and
redundancy kernels, and the second prefilter pass for fragments that overflow the default work lists."""
import ctypes as C
import os
import subprocess
import sys

import pytest

import reflib

SIMT_DIR = os.path.join(reflib.ROOT, "tests", "simt")
SIMT_SO = os.path.join(SIMT_DIR, "libmb200_simt.so")

SLICE = [
    "tests/test_gpu_parity.py::test_align_golden[frag_nomask-lanes+wave+rows2]",
    "tests/test_gpu_parity.py::test_align_golden[frag_nomask-rows2-packed]",
    "tests/test_gpu_parity.py::test_extractorfs_golden[orf_edge]",
    "tests/test_gpu_parity.py::test_edge_cases",
    "tests/test_gpu_parity.py::test_reduceredundancy_golden",
    "tests/test_gpu_parity.py::test_reference_own_end_to_end_fixtures[two_contigs]",
    "tests/test_zz_gpu_adversarial.py::test_predictexons_flags_golden[table6]",
    "tests/test_zz_gpu_adversarial.py::test_very_long_fragments_take_the_second_prefilter_pass",
    "tests/test_zz_gpu_adversarial.py::test_cli_gpupredictexons_takes_the_reference_flags[table4]",     # the C++ driver, linked against the emulation
    "tests/test_zz_gpu_adversarial.py::test_index_clone_gives_a_second_context_the_same_index",
    "tests/test_zz_gpu_adversarial.py::test_cli_gpupredictexons_on_several_devices_gives_the_one_device_db",   # --gpus 2 / 3: sharding, key bases, merge (all "devices" = the host here)
]


@pytest.fixture(scope="module")
def simt_lib():
    subprocess.check_call(["make", "-s", "-C", SIMT_DIR, "CXX=" + ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")])
    L = C.CDLL(SIMT_SO)
    L.mb200_build_kind.restype = C.c_char_p
    assert L.mb200_build_kind() == b"simt-emulation"
    return SIMT_SO


def test_product_library_is_the_device_build():
    from metaeuk_b200 import lib as mlib
    L = mlib.load_library()
    L.mb200_build_kind.restype = C.c_char_p
    assert L.mb200_build_kind() == b"sm_100a"


def test_gpu_suite_slice_on_the_simt_emulation(simt_lib):
    env = dict(os.environ, MB200_LIB_PATH=simt_lib, MB200_CLI_PATH=os.path.join(SIMT_DIR, "mb200_simt"), OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + SLICE, cwd=reflib.ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "%d passed" % len(SLICE) in r.stdout, r.stdout[-500:]
