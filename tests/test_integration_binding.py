"""The reference-side binding as compiled code (missing #6): integration/gpupredictexons.cpp is a MetaEuk module with the `Command` signature, built
against the UNMODIFIED reference's headers and libmmseqs-framework.a (oracle/_ref/build) and registered in a tool's command table the way src/metaeuk.cpp does it.

CPU test (build container only -- it needs /root/reference): the tool is linked with the SIMT emulation of the library (tests/simt: the product's kernels and host
orchestration as CPU fibers, TEST INFRASTRUCTURE), run on the reference's own tests/two_contigs data with the flags of tests/run.sh, and its output DB must equal --
data file, index and dbtype, byte for byte -- what the reference's `predictexons` writes for the same DBs.
GPU test: the same tool linked with libmetaeuk_b200.so (built here, travels to the GPU box) on DBs written from the golden fixture; entries must equal the
reference's dp_predictions."""
import os
import subprocess

import pytest

import reflib

INTEG = os.path.join(reflib.ROOT, "integration")
REF = os.environ.get("MB200_REFERENCE_DIR", "/root/reference")
REFBIN = os.path.join(reflib.ROOT, "oracle", "_ref", "metaeuk-ref")
FRAMEWORK = os.path.join(reflib.ROOT, "oracle", "_ref", "build", "lib", "mmseqs", "src", "libmmseqs-framework.a")


def _read_db(path):
    idx = [l.split("\t") for l in open(path + ".index").read().splitlines()]
    files = [path] if os.path.exists(path) else sorted((f for f in os.listdir(os.path.dirname(path)) if f.startswith(os.path.basename(path) + ".") and f.split(".")[-1].isdigit()),
                                                       key=lambda f: int(f.split(".")[-1]))
    data = b"".join(open(f if os.path.isabs(f) else os.path.join(os.path.dirname(path), f), "rb").read() for f in files)
    return {int(k): data[int(o):int(o) + int(n) - 1].decode() for k, o, n in idx}, data, open(path + ".index").read(), open(path + ".dbtype", "rb").read()


@pytest.mark.skipif(not (os.path.isdir(REF) and os.path.exists(FRAMEWORK) and os.path.exists(REFBIN)), reason="needs /root/reference and oracle/_ref (build container)")
def test_compiled_binding_equals_the_reference_module_chain(tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call(["make", "-s", "-C", os.path.join(reflib.ROOT, "tests", "simt"), "CXX=" + gxx])
    subprocess.check_call(["bash", os.path.join(INTEG, "build.sh"), os.path.join(reflib.ROOT, "tests", "simt", "libmb200_simt.so"), "metaeuk-b200-simt"], stdout=subprocess.DEVNULL)
    tool = os.path.join(INTEG, "_build", "metaeuk-b200-simt")
    d = str(tmp_path); data = os.path.join(REF, "tests", "two_contigs")
    env = dict(os.environ, OMP_NUM_THREADS="4")
    q = dict(stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
    subprocess.check_call([tool, "createdb", data + "/contigs.fna", d + "/contigs", "--dbtype", "2"], **q)
    subprocess.check_call([tool, "createdb", data + "/proteins.faa", d + "/proteins", "--dbtype", "1"], **q)
    flags = ["--set-gap-extend", "-1", "--metaeuk-tcov", "0"]      # tests/run.sh: ./test.sh "$1" two_contigs two_contigs_results -1 0
    subprocess.check_call([tool, "gpupredictexons", d + "/contigs", d + "/proteins", d + "/predB200"] + flags, **q)
    subprocess.check_call([REFBIN, "predictexons", d + "/contigs", d + "/proteins", d + "/predRef", d + "/tmp", "--threads", "2"] + flags, **q)
    mine, ref = _read_db(d + "/predB200"), _read_db(d + "/predRef")
    assert mine[0] == ref[0] and sum(len(v) for v in ref[0].values()) > 0
    assert mine[1] == ref[1], "data file differs"
    assert mine[2] == ref[2], "index differs"
    assert mine[3] == ref[3], "dbtype differs"


@pytest.mark.gpu
def test_compiled_binding_on_the_gpu(tmp_path):
    tool = os.path.join(INTEG, "_build", "metaeuk-b200")
    if not os.path.exists(tool):
        pytest.skip("integration/_build/metaeuk-b200 not built (needs /root/reference: build container)")
    t = reflib.load_case("ref_tests")["sets"]["two_contigs"]
    d = str(tmp_path)
    reflib.write_db(d + "/contigs", [(k, s + "\n") for k, s in t["contigs"]], 1)      # Parameters::DBTYPE_NUCLEOTIDES
    reflib.write_db(d + "/proteins", [(k, s + "\n") for k, s in t["targets"]], 0)
    subprocess.check_call([tool, "gpupredictexons", d + "/contigs", d + "/proteins", d + "/pred", "--set-gap-extend", str(t["set_gap_extend"]), "--metaeuk-tcov", str(t["tcov"])],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    got = _read_db(d + "/pred")[0]
    assert {str(k): v for k, v in got.items()} == t["dp"]
