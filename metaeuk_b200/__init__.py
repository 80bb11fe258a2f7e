"""metaeuk_b200 -- B200-native (sm_100a) implementation of MetaEuk's homology hot path:
MMseqs2 k-mer prefilter -> gapped Smith-Waterman -> exon-chaining DP, behind a C ABI (include/metaeuk_b200.h).

`metaeuk_b200.lib` is the ctypes binding of csrc (libmetaeuk_b200.so); `metaeuk_b200.modules` mirrors the
reference's module surface (prefilter / align / resultspercontig / collectoptimalset / predictexons) on
in-memory DBs and emits the reference's result-DB records byte for byte.
There is NO CPU fallback: importing works anywhere, creating a Context requires a CUDA device and the built library."""
from .lib import Context, LIB_PATH, Mb200Error, build_library, load_library  # noqa: F401
