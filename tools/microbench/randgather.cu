// randgather.cu -- how many independent random reads per second does one B200 sustain from HBM?  (diagnostic for k_prefilter's index lookups:
// offsets[k-mer] -> entries[...]; not part of the library).  Every thread issues U independent loads per round at hashed addresses of a buffer of
// S bytes; access sizes 8 B (one offsets / entry read) and 32 B (one inline bucket = one sector, two 128-bit loads).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
template <int BYTES, int U> __global__ void __launch_bounds__(512) k_gather(const uint4 *__restrict__ buf, uint64_t nUnits, int rounds, unsigned long long *sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (int r = 0; r < rounds; r++) {
        uint64_t idx[U];
#pragma unroll
        for (int u = 0; u < U; u++) idx[u] = mix(tid * 1315423911ull + (uint64_t)r * U + u) % nUnits;
        if (BYTES == 8) {
            const uint2 *b2 = reinterpret_cast<const uint2 *>(buf);
            uint2 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = b2[idx[u]];
#pragma unroll
            for (int u = 0; u < U; u++) acc += v[u].x ^ v[u].y;
        } else if (BYTES == 33) {     // 32 bytes in ONE 256-bit load (LDG.E.256, sm_100a)
            uint32_t v[U][8];
#pragma unroll
            for (int u = 0; u < U; u++)
                asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]), "=r"(v[u][4]), "=r"(v[u][5]), "=r"(v[u][6]), "=r"(v[u][7])
                             : "l"(buf + idx[u] * 2));
#pragma unroll
            for (int u = 0; u < U; u++) acc += v[u][0] ^ v[u][7];
        } else {
            uint4 v[U][2];
#pragma unroll
            for (int u = 0; u < U; u++) { v[u][0] = buf[idx[u] * 2]; v[u][1] = buf[idx[u] * 2 + 1]; }
#pragma unroll
            for (int u = 0; u < U; u++) acc += v[u][0].x ^ v[u][1].w;
        }
    }
    if (acc == 0x1234567) atomicAdd(sink, acc);
}
template <int BYTES, int U> static void run(const uint4 *buf, size_t bytes, unsigned long long *sink, const char *tag) {
    const int blocks = 148 * 4, threads = 512, rounds = 64;
    const uint64_t nUnits = bytes / (BYTES == 33 ? 32 : BYTES);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_gather<BYTES, U><<<blocks, threads>>>(buf, nUnits, 4, sink);
    cudaEventRecord(a);
    k_gather<BYTES, U><<<blocks, threads>>>(buf, nUnits, rounds, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double n = (double)blocks * threads * rounds * U;
    printf("%-10s buffer %6.0f MB  access %2d B  in flight/thread %d : %7.2f G accesses/s  (%6.1f GB/s useful)\n", tag, bytes / 1048576.0, BYTES, U, n / ms * 1e-6, n * BYTES / ms * 1e-6);
}
int main() {
    const size_t maxBytes = 4ull << 30;
    uint4 *buf; cudaMalloc(&buf, maxBytes); cudaMemset(buf, 1, maxBytes);
    unsigned long long *sink; cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
    int lim[1] = {0};
    for (int g = 0; g < 1; g++) {
        if (lim[g]) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, lim[g]); }
        size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("--- L2 fetch granularity %zu B\n", got);
        const size_t sizes[4] = {64ull << 20, 256ull << 20, 1536ull << 20, 4ull << 30};
        for (int s = 0; s < 4; s++) {
            run<8, 4>(buf, sizes[s], sink, "u8x4"); run<8, 8>(buf, sizes[s], sink, "u8x8"); run<32, 4>(buf, sizes[s], sink, "b32x4"); run<33, 4>(buf, sizes[s], sink, "v256x4"); run<33, 2>(buf, sizes[s], sink, "v256x2");
        }
    }
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
