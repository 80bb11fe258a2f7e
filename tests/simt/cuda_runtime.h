// tests/simt/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.  NOT the CUDA runtime, not shipped, never a fallback.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the product's CUDA translation unit (metaeuk_b200/csrc/mb200.cu: kernels AND host
// orchestration) into tests/simt/libmb200_simt.so, so that the GPU test-suite's logic can be exercised in a container without a GPU:
//   * "device memory" is host memory, streams/events are no-ops with wall-clock timers;
//   * a kernel launch runs the blocks of the grid one after the other; the threads of a block are cooperative fibers (ucontext) that run
//     until they reach a synchronisation point -- __syncthreads, __syncwarp, a __shfl_*/__ballot/__any/__all/__reduce collective -- and are
//     resumed when every participant has arrived, which is the SIMT contract the kernels are written against (no assumption of lock-step
//     execution between those points);
//   * the integer SIMD (DPX) intrinsics and the few PTX instructions the kernels use are restated in C.
// The library built this way reports itself through mb200_build_kind() = "simt-emulation"; bench.py and the product loader never use it.
#pragma once
#define MB200_SIMT_EMU 1
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <chrono>
#include <functional>

// ---------------------------------------------------------------- vector types, qualifiers
struct uint2 { unsigned int x, y; };
struct uint3 { unsigned int x, y, z; };
struct uint4 { unsigned int x, y, z, w; };
struct int2 { int x, y; };
struct dim3 { unsigned int x, y, z; dim3(unsigned int a = 1, unsigned int b = 1, unsigned int c = 1) : x(a), y(b), z(c) {} };
static inline uint2 make_uint2(unsigned int x, unsigned int y) { uint2 r; r.x = x; r.y = y; return r; }
static inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---------------------------------------------------------------- the fiber scheduler (simt.cpp)
namespace simt {
struct Thread { uint3 tid, bid; dim3 bdim, gdim; };
extern Thread *cur;                       // the fiber that is running
void launch(dim3 grid, dim3 block, size_t dynSmem, const std::function<void()> &body);
unsigned char *dyn_smem();
void sync_block();
enum Op { SHFL_IDX, SHFL_UP, SHFL_DOWN, SHFL_XOR, BALLOT, REDUCE_MAX, REDUCE_ADD, SYNC };
uint64_t warp_collective(unsigned mask, Op op, uint64_t value, int param, int width);
}
#define threadIdx (simt::cur->tid)
#define blockIdx (simt::cur->bid)
#define blockDim (simt::cur->bdim)
#define gridDim (simt::cur->gdim)
#define MB_UNPAREN(...) __VA_ARGS__
#define MB_LAUNCH(kern, grid, block, smem, stream, ...) simt::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { MB_UNPAREN kern(__VA_ARGS__); })
#define MB_DYN_SHARED(name) unsigned char *name = simt::dyn_smem()

static inline void __syncthreads() { simt::sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_collective(mask, simt::SYNC, 0, 0, 32); }
template <class T> static inline uint64_t simt_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes"); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T simt_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) { return simt_unbits<T>(simt::warp_collective(m, simt::SHFL_IDX, simt_bits(v), src, width)); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) { return simt_unbits<T>(simt::warp_collective(m, simt::SHFL_UP, simt_bits(v), (int)d, width)); }
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) { return simt_unbits<T>(simt::warp_collective(m, simt::SHFL_DOWN, simt_bits(v), (int)d, width)); }
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) { return simt_unbits<T>(simt::warp_collective(m, simt::SHFL_XOR, simt_bits(v), x, width)); }
static inline unsigned __ballot_sync(unsigned m, int pred) { return (unsigned)simt::warp_collective(m, simt::BALLOT, pred ? 1 : 0, 0, 32); }
static inline int __any_sync(unsigned m, int pred) { return (__ballot_sync(m, pred) & m) != 0; }
static inline int __all_sync(unsigned m, int pred) { return (__ballot_sync(m, pred) & m) == m; }
static inline int __reduce_max_sync(unsigned m, int v) { return (int)(int64_t)simt::warp_collective(m, simt::REDUCE_MAX, (uint64_t)(int64_t)v, 0, 32); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)simt::warp_collective(m, simt::REDUCE_MAX, (uint64_t)v, 0, 32); }
// lanes that happen to be converged: any subset containing the caller is a legal answer; the smallest one keeps the emulation simple
static inline unsigned __activemask() { return 1u << (threadIdx.x & 31u); }

// ---------------------------------------------------------------- atomics (one OS thread runs all fibers: plain read-modify-write)
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T *p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ---------------------------------------------------------------- scalar intrinsics
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline long long clock64() { return (long long)std::chrono::steady_clock::now().time_since_epoch().count(); }
static inline double __dmul_rn(double a, double b) { return a * b; }     // built with -ffp-contract=off
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
template <class A, class B> static inline auto max(A a, B b) -> decltype(a + b) { typedef decltype(a + b) R; return (R)a > (R)b ? (R)a : (R)b; }
template <class A, class B> static inline auto min(A a, B b) -> decltype(a + b) { typedef decltype(a + b) R; return (R)a < (R)b ? (R)a : (R)b; }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t v = ((uint64_t)y << 32) | x; unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
// prmt.b32 in its default mode: selector nibble bit 3 = replicate the sign bit of the selected byte
static inline unsigned simt_prmt(unsigned a, unsigned b, unsigned sel) {
    const uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned n = (sel >> (4 * i)) & 0xFu; unsigned byte = (unsigned)((v >> (8 * (n & 7u))) & 0xFFu);
        if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0x00u;
        r |= byte << (8 * i);
    }
    return r;
}
// DPX: three-operand integer max/add with optional clamp at zero; the s16x2 forms work on the two halves independently (wrapping add)
static inline int __viaddmax_s32(int a, int b, int c) { const int s = (int)((unsigned)a + (unsigned)b); return s > c ? s : c; }
static inline int __viaddmax_s32_relu(int a, int b, int c) { const int r = __viaddmax_s32(a, b, c); return r > 0 ? r : 0; }
static inline int __vimax_s32_relu(int a, int b) { const int r = a > b ? a : b; return r > 0 ? r : 0; }
static inline int __vimax3_s32(int a, int b, int c) { const int r = a > b ? a : b; return r > c ? r : c; }
static inline unsigned simt_pack2(int lo, int hi) { return ((unsigned)lo & 0xFFFFu) | ((unsigned)hi << 16); }
static inline int simt_lo(unsigned v) { return (int)(int16_t)(v & 0xFFFFu); }
static inline int simt_hi(unsigned v) { return (int)(int16_t)(v >> 16); }
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) {
    const int l = (int)(int16_t)(simt_lo(a) + simt_lo(b)), h = (int)(int16_t)(simt_hi(a) + simt_hi(b));
    return simt_pack2(l > simt_lo(c) ? l : simt_lo(c), h > simt_hi(c) ? h : simt_hi(c));
}
static inline unsigned __viaddmax_s16x2_relu(unsigned a, unsigned b, unsigned c) {
    const unsigned r = __viaddmax_s16x2(a, b, c);
    return simt_pack2(simt_lo(r) > 0 ? simt_lo(r) : 0, simt_hi(r) > 0 ? simt_hi(r) : 0);
}
static inline unsigned __vmaxs2(unsigned a, unsigned b) { return simt_pack2(simt_lo(a) > simt_lo(b) ? simt_lo(a) : simt_lo(b), simt_hi(a) > simt_hi(b) ? simt_hi(a) : simt_hi(b)); }
static inline unsigned __vimax_s16x2_relu(unsigned a, unsigned b) { const unsigned r = __vmaxs2(a, b); return simt_pack2(simt_lo(r) > 0 ? simt_lo(r) : 0, simt_hi(r) > 0 ? simt_hi(r) : 0); }

// ---------------------------------------------------------------- runtime API (host memory, synchronous "streams")
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef struct simtStream *cudaStream_t;
typedef struct simtEvent { std::chrono::steady_clock::time_point t; } *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };
static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "simt emulation error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { const char *e = getenv("MB200_SIMT_DEVICES"); *n = e && atoi(e) > 0 ? atoi(e) : 1; return cudaSuccess; }   // (all emulated devices are the host)
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 2; strcpy(p->name, "simt emulation"); return cudaSuccess; }
// device allocations are filled with a poison pattern: cudaMalloc does not zero memory, and a kernel that silently relies on zeros should fail here
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
    const size_t m = (n + 255) & ~(size_t)255; *p = (T *)aligned_alloc(256, m ? m : 256);
    if (*p) memset((void *)*p, getenv("MB200_SIMT_POISON") ? atoi(getenv("MB200_SIMT_POISON")) : 0xCD, m);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { *p = (T *)aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new simtEvent(); return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new simtEvent(); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }
