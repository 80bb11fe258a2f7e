#!/usr/bin/env bash
# TEST INFRASTRUCTURE / CPU BASELINE, never linked or called by the product.
# Builds the UNMODIFIED reference (soedinglab/metaeuk at /root/reference, vendored lib/mmseqs) out of tree with the
# recipe of SURVEY.md §8c; outputs go ONLY to oracle/_ref/ (git-ignored, not gpurun-ignored: the binary travels to the
# GPU box like our own .so).  No reference source is copied into the repo.
#   oracle/_ref/metaeuk-ref            the AVX2 binary (cpu_baseline.kind = "reference", bench.py --impl reference)
#   oracle/_ref/build/                 cmake tree (libmmseqs-framework.a for integration/, build container only)
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${MB200_REFERENCE_DIR:-/root/reference}"
out="$here/_ref"
if [ ! -d "$ref" ]; then echo "build_ref: $ref absent (GPU box) - using prebuilt $out/metaeuk-ref" >&2; test -x "$out/metaeuk-ref"; exit $?; fi
mkdir -p "$out/build"
cd "$out/build"
if [ ! -f Makefile ]; then
  cmake -G "Unix Makefiles" -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ \
        -DCMAKE_BUILD_TYPE=Release -DHAVE_AVX2=1 -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
        -DVERSION_OVERRIDE=oracle "$ref" > "$out/cmake.log" 2>&1
fi
make -j"$(nproc)" metaeuk > "$out/make.log" 2>&1
cp -f src/metaeuk "$out/metaeuk-ref"
strip "$out/metaeuk-ref" || true
echo "build_ref: $out/metaeuk-ref"
