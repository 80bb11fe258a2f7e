#!/usr/bin/env python3
"""Pin the outputs of the workloads bench.py TIMES against the UNMODIFIED reference (build container only).

  python tests/golden/make_bench_digests.py c4s c2 c3s c4:1000     # c4:N = the first N contigs of the c4 workload (N a multiple of 250)

TEST INFRASTRUCTURE.  Generates the same seeded inputs as bench.py (it imports bench.py's generators), runs the reference binary
oracle/_ref/metaeuk-ref (oracle/build_ref.sh) through the module chain of `metaeuk predictexons` (tools/refrun.py: the flags the workflow
passes, --mask 1, --split 1) and stores SHA-256 digests + line counts of the result DBs in tests/golden/bench_digests.json:
  c4 / c4s : dp_predictions, one digest per block of 250 contigs (contig keys block*250 ..), plus line counts of pref / search_res
  c2       : pref             c3s : pref and search_res (align)
Digest form: entries in ascending key order, empty entries skipped, ">key\\n" + entry text (tools/refrun.py: digest_entries).
bench.py hashes the result of its timed step the same way and exits non-zero on a mismatch; tests/test_gpu_bench_parity.py does the same.
"""
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import refrun  # noqa: E402

OUT = os.path.join(HERE, "bench_digests.json")


def count_lines(rr, name):
    keys, ents = refrun.read_result_db(rr.p(name))
    return refrun.digest_entries(keys, ents)


def main():
    todo = sys.argv[1:] or ["c4s", "c2"]
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    thr = int(os.environ.get("MB200_REF_THREADS", os.cpu_count() or 1))
    for spec in todo:
        name, _, limit = spec.partition(":")
        wl = bench.WORKLOADS[name]
        t0 = time.time()
        prot, poff = bench.gen_proteins(wl["n_prot"])
        rr = refrun.RefRun(tempfile.mkdtemp(prefix="mb200_digest_", dir=os.environ.get("MB200_TMP")), thr)
        try:
            rr.create_targets(prot, poff)
            rec = {"proteins": wl["n_prot"], "masked_residues": rr.masked_residues, "reference": "unmodified soedinglab/metaeuk at /root/reference, AVX2 build (oracle/build_ref.sh)",
                   "threads": thr}
            if wl.get("predict"):
                n = int(limit) if limit else wl["n_contigs"]
                assert n % bench.DIGEST_BLOCK == 0 or n == wl["n_contigs"]
                nuc, off, genes = bench.gen_planted_contigs(0, n, wl["contig_len"], prot, poff)
                rr.create_contigs(nuc, off)
                r = rr.predictexons()
                keys, ents = refrun.read_result_db(rr.p("dp"))
                blocks = {}
                for b in range((n + bench.DIGEST_BLOCK - 1) // bench.DIGEST_BLOCK):
                    blocks[str(b)] = refrun.digest_entries(keys, ents, b * bench.DIGEST_BLOCK, (b + 1) * bench.DIGEST_BLOCK)
                rec.update({"block": bench.DIGEST_BLOCK, "contigs": n, "fragments": r["fragments"], "planted_genes": genes, "blocks": blocks,
                            "pref": count_lines(rr, "pref"), "aln": count_lines(rr, "aln"), "dp": refrun.digest_entries(keys, ents),
                            "reference_stage_wall_s": {k: round(v["wall_s"], 2) for k, v in r["stages"].items()}})
            else:
                frag, foff = bench.gen_fragments(wl["n_frag"], prot, poff, seed=4321)
                r = rr.search(frag, foff, align=wl["align"])
                rec.update({"fragments": wl["n_frag"], "pref": count_lines(rr, "pref"), "reference_stage_wall_s": {k: round(v["wall_s"], 2) for k, v in r["stages"].items()}})
                if wl["align"]:
                    rec["aln"] = count_lines(rr, "aln")
            rec["index_build_wall_s"] = round(rr.times["indexdb"]["wall_s"], 2)
            out[name] = rec
            json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
            print(spec, "done in %.0f s:" % (time.time() - t0), {k: v for k, v in rec.items() if k != "blocks"}, flush=True)
        finally:
            rr.cleanup()


if __name__ == "__main__":
    main()
