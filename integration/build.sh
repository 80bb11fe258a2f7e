#!/usr/bin/env bash
# integration/build.sh [path/to/lib<name>.so [output name]]
# Compiles the binding against the UNMODIFIED reference (headers from /root/reference, libmmseqs-framework.a from oracle/_ref/build -- run
# oracle/build_ref.sh first) and links it with the C-ABI library given (default: the sm_100a build -> integration/_build/metaeuk-b200; the tests
# also link it with tests/simt/libmb200_simt.so -> metaeuk-b200-simt to run it without a GPU).  Build-container only: outputs go to
# integration/_build/ (git-ignored, not gpurun-ignored: the binary travels to the GPU box); nothing here is linked into libmetaeuk_b200.so.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
root="$(dirname "$here")"
ref="${MB200_REFERENCE_DIR:-/root/reference}"
rb="$root/oracle/_ref/build"
lib="${1:-$root/metaeuk_b200/libmetaeuk_b200.so}"
exe="${2:-metaeuk-b200}"
test -d "$ref" || { echo "integration/build.sh: $ref absent" >&2; exit 2; }
test -f "$rb/lib/mmseqs/src/libmmseqs-framework.a" || { echo "integration/build.sh: run oracle/build_ref.sh first" >&2; exit 2; }
out="$here/_build"; mkdir -p "$out"
M="$ref/lib/mmseqs"
INC="-I$ref/lib -I$ref/src/commons -I$M/lib/zstd/lib -I$M/lib/tinyexpr -I$M/lib/microtar -I$M/lib/tantan -I$M/lib/simde -I$M/lib -I$M/lib/simd -I$M/lib/alp -I$M/lib/ksw2 -I$M/lib/xxhash -I$M/lib/ips4o \
 -I$rb/generated -I$rb/lib/mmseqs/generated -I$M/src/alignment -I$M/src/clustering -I$M/src/commons -I$M/src/multihit -I$M/src/prefiltering -I$M/src/linclust -I$M/src/taxonomy -I$M/src/util -I$M/src -I$root/include"
CXX="${CXX:-/usr/bin/g++}"
FLAGS="-O2 -DNDEBUG -fsigned-char -mavx2 -mfma -mcx16 -std=c++1y -fno-exceptions -fopenmp -DENABLE_IPS4O=1 -DHAVE_POSIX_FADVISE=1 -DHAVE_POSIX_MADVISE=1 -DHAVE_ZLIB=1 -DOPENMP=1 -w"
$CXX $FLAGS $INC -c "$here/gpupredictexons.cpp" -o "$out/gpupredictexons.o"
$CXX $FLAGS $INC -c "$here/metaeuk_b200_tool.cpp" -o "$out/metaeuk_b200_tool.o"
libdir="$(cd "$(dirname "$lib")" && pwd)"; libname="$(basename "$lib")"; libname="${libname#lib}"; libname="${libname%.so}"
$CXX -o "$out/$exe" "$out/gpupredictexons.o" "$out/metaeuk_b200_tool.o" \
  "$rb/lib/mmseqs/src/libmmseqs-framework.a" "$rb/src/version/libversion.a" "$(ls /usr/lib/gcc/x86_64-linux-gnu/*/libatomic.so | head -1)" "$rb/lib/mmseqs/lib/tinyexpr/libtinyexpr.a" \
  "$rb/lib/mmseqs/lib/zstd/build/cmake/lib/libzstd.a" "$rb/lib/mmseqs/lib/microtar/libmicrotar.a" "$rb/lib/mmseqs/lib/tantan/libtantan.a" -lz "$(ls /usr/lib/gcc/x86_64-linux-gnu/*/libgomp.so | head -1)" -lpthread \
  -L"$libdir" -l"$libname" -Wl,-rpath,"$libdir" -Wl,-rpath,'$ORIGIN/../../metaeuk_b200'
echo "integration/build.sh: $out/$exe (linked with $lib)"
