// cta_prims.cuh -- block-level ("CTA") programming primitives used by the prefilter / chain kernels.
//
// The kernels are written as a sequence of block-wide parallel-for regions separated by barriers
// (CTA_FOR ... CTA_SYNC) plus three collective primitives (exclusive scan, 64-bit key sort, append).
// Under nvcc these compile to the real sm_100a code.  Under a host compiler with -DMB200_EMULATE the
// same source compiles to a serial single-"thread" execution (CTA_FOR = plain loop); that build exists
// ONLY so tests/ can check the kernel logic against the oracle in a container without a GPU.  It is not
// part of libmetaeuk_b200.so and is never used as a fallback.
#pragma once
#include <stdint.h>

#if (defined(__CUDACC__) || defined(MB200_SIMT_EMU)) && !defined(MB200_EMULATE)
#define MB_DEVICE_BUILD 1
#define MB_DEV __device__ __forceinline__
#define MB_DEV_CALL __device__ __noinline__   // block-wide collectives used at several sites: one copy of the code
#define CTA_TID (threadIdx.x)
#define CTA_NT (blockDim.x)
#define CTA_FOR(i, n) for (uint32_t i = threadIdx.x, i##_end = (uint32_t)(n); i < i##_end; i += blockDim.x)
#define CTA_SYNC() __syncthreads()
#define CTA_SINGLE if (threadIdx.x == 0)
// one warp per work item / contiguous chunk of a range per thread
#define CTA_WARP_FOR(i, n) for (uint32_t i = threadIdx.x >> 5, i##_end = (uint32_t)(n); i < i##_end; i += blockDim.x >> 5)
#define CTA_CHUNK(b, e, n) uint32_t b, e; { const uint32_t n_ = (uint32_t)(n), c_ = (n_ + blockDim.x - 1) / blockDim.x; b = threadIdx.x * c_; if (b > n_) b = n_; e = b + c_; if (e > n_) e = n_; }
MB_DEV uint32_t mb_atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
MB_DEV uint32_t mb_atomic_or(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
MB_DEV uint32_t mb_atomic_max(uint32_t *p, uint32_t v) { return atomicMax(p, v); }
MB_DEV uint32_t mb_atomic_cas(uint32_t *p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
// IEEE double ops that must never be contracted into FMAs (composition-bias parity with the CPU path)
MB_DEV double mb_dmul(double a, double b) { return __dmul_rn(a, b); }
MB_DEV double mb_dadd(double a, double b) { return __dadd_rn(a, b); }
MB_DEV double mb_ddiv(double a, double b) { return __ddiv_rn(a, b); }
MB_DEV float mb_fadd(float a, float b) { return __fadd_rn(a, b); }
MB_DEV float mb_fmul(float a, float b) { return __fmul_rn(a, b); }
MB_DEV float mb_fdiv(float a, float b) { return __fdiv_rn(a, b); }
#else
#define MB_DEVICE_BUILD 0
#define MB_DEV static inline
#define MB_DEV_CALL static inline
#define CTA_TID (0u)
#define CTA_NT (1u)
#define CTA_FOR(i, n) for (uint32_t i = 0, i##_end = (uint32_t)(n); i < i##_end; i++)
#define CTA_SYNC() ((void)0)
#define CTA_SINGLE if (true)
#define CTA_WARP_FOR(i, n) for (uint32_t i = 0, i##_end = (uint32_t)(n); i < i##_end; i++)
#define CTA_CHUNK(b, e, n) uint32_t b = 0, e = (uint32_t)(n);
#include <algorithm>
struct uint2 { unsigned int x, y; };
struct uint4 { unsigned int x, y, z, w; };
MB_DEV uint32_t mb_atomic_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
MB_DEV uint32_t mb_atomic_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
MB_DEV uint32_t mb_atomic_max(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
MB_DEV uint32_t mb_atomic_cas(uint32_t *p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = v; return o; }
MB_DEV double mb_dmul(double a, double b) { return a * b; }   // emulation is built with -ffp-contract=off
MB_DEV double mb_dadd(double a, double b) { return a + b; }
MB_DEV double mb_ddiv(double a, double b) { return a / b; }
MB_DEV float mb_fadd(float a, float b) { return a + b; }
MB_DEV float mb_fmul(float a, float b) { return a * b; }
MB_DEV float mb_fdiv(float a, float b) { return a / b; }
#endif

#ifndef MB_CTA_THREADS
#define MB_CTA_THREADS 512
#endif

// Reserve one slot of a list whose length lives at *ctr (shared memory) for every calling thread with pred == true and return its
// index (~0u otherwise).  On the device the threads of a warp that append in the same instruction share ONE atomic.
MB_DEV uint32_t cta_append_slot(uint32_t *ctr, bool pred) {
#if MB_DEVICE_BUILD
    const unsigned act = __activemask();
    const unsigned b = __ballot_sync(act, pred);
    if (!pred) return ~0u;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(b) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(ctr, (uint32_t)__popc(b));
    base = __shfl_sync(b, base, leader);
    return base + (uint32_t)__popc(b & ((1u << lane) - 1u));
#else
    if (!pred) return ~0u;
    const uint32_t o = *ctr; *ctr = o + 1; return o;
#endif
}

// Scratch every collective needs (lives in shared memory on the device).
struct CtaScratch {
    uint32_t part[MB_CTA_THREADS + 1];
    uint32_t bcast[4];
};

// In-place exclusive prefix sum of a[0..n); also writes a[n] = total and returns the total to every thread.
// a may live in shared or global memory.
MB_DEV_CALL uint32_t cta_exclusive_scan(uint32_t *a, uint32_t n, CtaScratch *s) {
#if MB_DEVICE_BUILD
    // every warp owns one contiguous range and sweeps it 32 elements at a time (coalesced, shuffle scans, no block barrier
    // inside the sweep); one barrier to exchange warp totals; a second sweep adds the warp's prefix.
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5, nW = blockDim.x >> 5;
    const uint32_t per = ((n + nW - 1) / nW + 31u) & ~31u;
    uint32_t rb = wid * per; if (rb > n) rb = n;
    uint32_t re = rb + per; if (re > n) re = n;
    uint32_t carry = 0;
    for (uint32_t base = rb; base < re; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t v = (i < re) ? a[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += y; }
        if (i < re) a[i] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s->part[wid] = carry;
    __syncthreads();
    uint32_t wt = (lane < nW) ? s->part[lane] : 0u, winc = wt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, winc, d); if ((int)lane >= d) winc += y; }
    const uint32_t myPrefix = __shfl_sync(0xffffffffu, winc - wt, wid);
    const uint32_t total = __shfl_sync(0xffffffffu, winc, nW - 1);
    if (myPrefix) for (uint32_t i = rb + lane; i < re; i += 32) a[i] += myPrefix;
    if (tid == 0) a[n] = total;
    __syncthreads();
    return total;
#else
    (void)s;
    uint32_t run = 0;
    for (uint32_t i = 0; i < n; i++) { uint32_t v = a[i]; a[i] = run; run += v; }
    a[n] = run;
    return run;
#endif
}

// largest idx in [0,n) with a[idx] <= x, for a non-decreasing a[0..n] with a[0] == 0 <= x < a[n]
MB_DEV uint32_t cta_find_segment(const uint32_t *a, uint32_t n, uint32_t x) {
    uint32_t lo = 0, hi = n;   // invariant: a[lo] <= x < a[hi]
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (a[mid] <= x) lo = mid; else hi = mid; }
    return lo;
}

MB_DEV uint32_t mb_next_pow2(uint32_t n) { uint32_t p = 1; while (p < n) p <<= 1; return p; }

// Ascending sort of 64-bit keys a[0..n).  The buffer must have room for next_pow2(n) elements (padding is
// written as ~0).  Bitonic network with block barriers; fine for the few-hundred-element lists the prefilter sorts.
#if MB_DEVICE_BUILD
// Up to 32 * R keys sorted by ONE warp: R keys per lane in registers (element r * 32 + lane), bitonic network with the compare-exchanges between
// lanes through shuffles and those between registers in place -- no block barrier and no shared memory inside the sort.
template <int R> __device__ __forceinline__ void warp_sort_u64(uint64_t *a, uint32_t n) {
    const uint32_t lane = threadIdx.x & 31u;
    uint64_t key[R];
#pragma unroll
    for (int r = 0; r < R; r++) { const uint32_t idx = (uint32_t)r * 32u + lane; key[r] = (idx < n) ? a[idx] : ~0ull; }
#pragma unroll
    for (uint32_t k = 2; k <= 32u * R; k <<= 1) {
#pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32u) {
                const int rj = (int)(j >> 5);
#pragma unroll
                for (int r = 0; r < R; r++) if ((r & rj) == 0) {
                    const bool asc = ((((uint32_t)r * 32u) & k) == 0);      // k >= 64 here: the direction bit lies in the register index
                    const uint64_t x = key[r], y = key[r | rj];
                    const bool sw = (x > y) == asc;
                    key[r] = sw ? y : x; key[r | rj] = sw ? x : y;
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const uint64_t other = __shfl_xor_sync(0xffffffffu, key[r], (int)j);
                    const bool asc = (((uint32_t)r * 32u + lane) & k) == 0, lower = (lane & j) == 0;
                    const uint64_t mn = key[r] < other ? key[r] : other, mx = key[r] < other ? other : key[r];
                    key[r] = (asc == lower) ? mn : mx;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) { const uint32_t idx = (uint32_t)r * 32u + lane; if (idx < n) a[idx] = key[r]; }
}
#endif

MB_DEV_CALL void cta_sort_u64(uint64_t *a, uint32_t n) {
#if MB_DEVICE_BUILD
    const uint32_t n2 = mb_next_pow2(n);
    if (n2 <= 256) {   // short lists (the usual case): one warp, keys in registers, a single block barrier
        if (threadIdx.x < 32) {
            if (n2 <= 32) warp_sort_u64<1>(a, n);
            else if (n2 <= 64) warp_sort_u64<2>(a, n);
            else if (n2 <= 128) warp_sort_u64<4>(a, n);
            else warp_sort_u64<8>(a, n);
        }
        __syncthreads();
        return;
    }
    // Longer lists: bitonic network with every warp keeping a 64-key chunk in registers (two keys per lane).  Network steps between keys less than 64 apart stay inside a
    // chunk (shuffles / register pairs, no barrier); only the steps with a distance of 64 or more go through memory with a block barrier: 10 barriers for 512 keys, not 45.
    for (uint32_t i = n + threadIdx.x; i < n2; i += blockDim.x) a[i] = ~0ull;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5, nChunks = n2 >> 6;
    for (uint32_t c = warp; c < nChunks; c += nWarps) {       // chunks sorted on their own: network stages k = 2..64
        const uint32_t base = c << 6;
        uint64_t k0 = a[base + lane], k1 = a[base + 32u + lane];
#pragma unroll
        for (uint32_t k = 2; k <= 64u; k <<= 1) {
#pragma unroll
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                if (j == 32u) {
                    const bool asc = ((base & 64u) == 0);     // k == 64: the direction alternates from chunk to chunk
                    const bool sw = (k0 > k1) == asc; const uint64_t x = k0; k0 = sw ? k1 : k0; k1 = sw ? x : k1;
                } else {
                    const uint64_t o0 = __shfl_xor_sync(0xffffffffu, k0, (int)j), o1 = __shfl_xor_sync(0xffffffffu, k1, (int)j);
                    const bool lower = (lane & j) == 0;
                    const bool asc0 = k == 64u ? ((base & 64u) == 0) : ((lane & k) == 0), asc1 = k == 64u ? asc0 : (((32u + lane) & k) == 0);
                    k0 = (asc0 == lower) ? (k0 < o0 ? k0 : o0) : (k0 < o0 ? o0 : k0);
                    k1 = (asc1 == lower) ? (k1 < o1 ? k1 : o1) : (k1 < o1 ? o1 : k1);
                }
            }
        }
        a[base + lane] = k0; a[base + 32u + lane] = k1;
    }
    __syncthreads();
    for (uint32_t k = 128; k <= n2; k <<= 1) {
        for (uint32_t j = k >> 1; j >= 64u; j >>= 1) {         // distances of 64 and more: through memory
            for (uint32_t t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
                const uint32_t l = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));   // j is a power of two: (t / j) * 2j + t % j
                const uint32_t r = l + j;
                const bool asc = ((l & k) == 0);
                const uint64_t x = a[l], y = a[r];
                if ((x > y) == asc) { a[l] = y; a[r] = x; }
            }
            __syncthreads();
        }
        for (uint32_t c = warp; c < nChunks; c += nWarps) {   // distances 32..1: inside the chunk, one direction for the whole chunk
            const uint32_t base = c << 6;
            const bool asc = ((base & k) == 0);
            uint64_t k0 = a[base + lane], k1 = a[base + 32u + lane];
            { const bool sw = (k0 > k1) == asc; const uint64_t x = k0; k0 = sw ? k1 : k0; k1 = sw ? x : k1; }
#pragma unroll
            for (uint32_t j = 16; j > 0; j >>= 1) {
                const uint64_t o0 = __shfl_xor_sync(0xffffffffu, k0, (int)j), o1 = __shfl_xor_sync(0xffffffffu, k1, (int)j);
                const bool lower = (lane & j) == 0;
                k0 = (asc == lower) ? (k0 < o0 ? k0 : o0) : (k0 < o0 ? o0 : k0);
                k1 = (asc == lower) ? (k1 < o1 ? k1 : o1) : (k1 < o1 ? o1 : k1);
            }
            a[base + lane] = k0; a[base + 32u + lane] = k1;
        }
        __syncthreads();
    }
#else
    std::sort(a, a + n);
#endif
}
