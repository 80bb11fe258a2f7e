// randgather2.cu -- DRAM sectors per random 32-byte read (diagnostic for k_prefilter's bucket lookups; not part of the library).
// usage: randgather2 [L2 fetch granularity limit: 0 = leave the default | 32 | 64 | 128]      run under
//   ncu --metrics dram__sectors_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
// Variants of one 256-bit load per access from a 2 GB table (the size of the bucket table): default policy, L1 no-allocate + L2 evict-first, and the
// prefetch-size qualifiers .L2::64B / .L2::128B; plus 8-byte loads.  sectors / access = dram__sectors_read.sum / (blocks * threads * rounds * U).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
template <int V> __device__ __forceinline__ void ld256(const uint4 *p, uint32_t *v) {
    if (V == 0) asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
    if (V == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
    if (V == 2) { const uint4 a = __ldg(p), b = __ldg(p + 1); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
    if (V == 3) { const uint2 a = __ldg(reinterpret_cast<const uint2 *>(p)); v[0] = a.x; v[1] = a.y; v[2] = v[3] = v[4] = v[5] = v[6] = 0; v[7] = a.y; }
    if (V == 4) { uint32_t a, b; asm volatile("ld.global.nc.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p)); v[0] = a; v[1] = b; v[2] = v[3] = v[4] = v[5] = v[6] = 0; v[7] = b; }
    if (V == 5) { uint32_t a, b; asm volatile("ld.global.nc.L2::128B.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p)); v[0] = a; v[1] = b; v[2] = v[3] = v[4] = v[5] = v[6] = 0; v[7] = b; }
}
template <int V, int U> __global__ void __launch_bounds__(512) k_gather(const uint4 *__restrict__ buf, uint64_t nUnits, int rounds, unsigned long long *sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (int r = 0; r < rounds; r++) {
        uint64_t idx[U]; uint32_t v[U][8];
#pragma unroll
        for (int u = 0; u < U; u++) idx[u] = mix(tid * 1315423911ull + (uint64_t)r * U + u) % nUnits;
#pragma unroll
        for (int u = 0; u < U; u++) ld256<V>(buf + idx[u] * 2, v[u]);
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u][0] ^ v[u][7];
    }
    if (acc == 0x1234567) atomicAdd(sink, acc);
}
template <int V> static void run(const uint4 *buf, size_t bytes, unsigned long long *sink, const char *tag) {
    const int blocks = 148 * 4, threads = 512, rounds = 64, U = 4;
    const uint64_t nUnits = bytes / 32;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_gather<V, U><<<blocks, threads>>>(buf, nUnits, 4, sink);
    cudaEventRecord(a);
    k_gather<V, U><<<blocks, threads>>>(buf, nUnits, rounds, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double n = (double)blocks * threads * rounds * U;
    printf("%-34s %7.2f G accesses/s   (%.0f accesses in the timed launch)\n", tag, n / ms * 1e-6, n);
}
int main(int argc, char **argv) {
    const int lim = argc > 1 ? atoi(argv[1]) : 0;
    if (lim) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)lim); printf("cudaDeviceSetLimit(MaxL2FetchGranularity, %d): %s\n", lim, cudaGetErrorString(e)); }
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("--- L2 fetch granularity limit %zu B\n", got);
    const size_t bytes = 2ull << 30;
    uint4 *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    unsigned long long *sink; cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
    run<0>(buf, bytes, sink, "256-bit load, default policy");
    run<1>(buf, bytes, sink, "256-bit load, NA + L2 evict-first");
    run<2>(buf, bytes, sink, "two 128-bit loads");
    run<3>(buf, bytes, sink, "8-byte load");
    run<4>(buf, bytes, sink, "8-byte load .L2::64B");
    run<5>(buf, bytes, sink, "8-byte load .L2::128B");
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
