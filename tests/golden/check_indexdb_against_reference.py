#!/usr/bin/env python3
"""One-off pin of the persisted-index reader/writer (metaeuk_b200/csrc/host_indexdb.cpp) against the UNMODIFIED reference (needs a reference
binary, so it is not part of the test suite; the suite pins the same entries through the digests in tests/golden/*.json.gz).

  python tests/golden/check_indexdb_against_reference.py /path/to/metaeuk [workdir]

1. `createdb` + `indexdb` of the reference on the proteins of a golden case -> <work>/prot.idx (about 900 MB: 512 MB k-mer offsets + 387 MB
   similarity tables).
2. open_index_db() on that file must return the sequences, masked residues, offsets and keys of the DB the reference indexed.
3. write_index_db() on the same input must produce entries ENTRIES(9), ENTRIESOFFSETS(10), ENTRIESNUM(12), SEQCOUNT(13), SEQINDEXDATA(14),
   SEQINDEXDATASIZE(15), SEQINDEXSEQOFFSET(16), DBR1DATA(6), VERSION(0), SCOREMATRIX2MER(3), SCOREMATRIX3MER(4), SPACEDPATTERN(23) byte-identical to the reference's.
Result recorded in DESIGN.md section 4.7 (run in the build container, reference built out of tree with the recipe of SURVEY.md 8c)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reflib  # noqa: E402
from test_kernel_logic_emulated import _open_indexdb  # noqa: E402

IDX_FLAGS = ("--seed-sub-mat aa:VTML80.out,nucl:nucleotide.out -k 6 -s 4 --mask 1 --mask-prob 0.9 --mask-lower-case 0 --mask-n-repeat 0 "
             "--spaced-kmer-mode 1 --comp-bias-corr 1 --split 1 --threads 4 -v 1")


def entries(path):
    e = {}
    for line in open(path + ".index"):
        k, o, l = (int(x) for x in line.split("\t"))
        e[k] = (o, l)
    return e


def main():
    ref = sys.argv[1]
    work = sys.argv[2] if len(sys.argv) > 2 else "/tmp/mb200_indexdb_check"
    os.makedirs(work, exist_ok=True)
    c = reflib.load_case("contigs_sets")
    with open(work + "/prot.faa", "w") as f:
        for k, s in c["targets"]:
            f.write(">p%d\n%s\n" % (k, s))
    subprocess.check_call([ref, "createdb", work + "/prot.faa", work + "/prot", "-v", "1"])
    subprocess.check_call([ref, "indexdb", work + "/prot", work + "/prot"] + IDX_FLAGS.split())
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "libmb200_emu.so"))
    d, err = _open_indexdb(emu, work + "/prot.idx")
    assert d is not None, err
    data = open(work + "/prot", "rb").read()
    seqs = {}
    for line in open(work + "/prot.index"):
        k, o, l = (int(x) for x in line.split("\t"))
        seqs[k] = data[o:o + l - 2].decode()
    order = sorted(seqs)
    mask_of = {s: m for (_, s), m in zip(c["targets"], c["masked_targets"])}    # tantan output of the reference, by sequence
    u, off = reflib.pack([seqs[k] for k in order]); m, _ = reflib.pack([mask_of[seqs[k]] for k in order])
    assert (d["unmasked"] == u).all() and (d["seq_off"] == off).all() and (d["masked"] == m).all() and (d["keys"] == np.array(order)).all()
    print("reader: sequences, masked residues, offsets and keys of the reference-written index DB recovered")
    p = reflib._p
    keys = np.array(order, np.uint32)
    assert emu.emu_indexdb_write((work + "/mine.idx").encode(), p(m, C.c_uint8), p(u, C.c_uint8), p(off, C.c_uint64), C.c_uint32(len(keys)), p(keys, C.c_uint32), 127, 1) == 0
    ea, eb = entries(work + "/prot.idx"), entries(work + "/mine.idx")
    fa, fb = open(work + "/prot.idx", "rb"), open(work + "/mine.idx", "rb")
    for k in (0, 3, 4, 23, 6, 9, 10, 12, 13, 14, 15, 16):     # (3, 4 = the 2-mer / 3-mer similarity tables, 23 = the spaced pattern: written since the end of round 2)
        fa.seek(ea[k][0]); fb.seek(eb[k][0])
        a, b = fa.read(ea[k][1]), fb.read(eb[k][1])
        if k == 14:     # the reference writes getDataSize() + 1 bytes: the last one lies past the lookup's data (uninitialised there, 0 here)
            a, b = a[:-2], b[:-2]
        assert ea[k][1] == eb[k][1] and a == b, k
    print("writer: entries 0 3 4 23 6 9 10 12 13 14 15 16 byte-identical to the reference's")


if __name__ == "__main__":
    main()
