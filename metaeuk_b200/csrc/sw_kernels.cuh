// sw_kernels.cuh -- gapped Smith-Waterman on the device (sm_100a), exact affine-gap local alignment.
//
// What the reference computes per (query, target) hit (M/ = reference/lib/mmseqs/):
//   SmithWaterman::ssw_align_private<SEQ_SEQ>   M/src/alignment/StripedSmithWaterman.cpp:309-545
//     forward  sw_sse2_byte / sw_sse2_word (:638-940, :942-1214): score, FIRST target column reaching it, smallest query row
//     e-value gate (:389-398) -- only alignments that can pass -e get start coordinates
//     reverse  on the reversed query prefix / target prefix, stopping at the first column whose maximum == score (:400-476, :900)
// Recurrence:  H = max(0, Hdiag + s, E, F);  E' = max(E - ge, H - go);  F' = max(F - ge, H - go);  s = mat[t][q_j] + compBias8[j].
//
// Kernels (forward = score + e-value gate for all hits; trace = start coordinates for the few % that pass):
//   k_sw_rows   forward, one target per lane, 32-bit cells, register-resident 32-row blocks (also used for queries > 1024 residues)
//   k_sw_rows2  forward, two targets per lane in packed s16x2 cells (DPX), one (query, chunk of 32 hits) per half-warp  [default]
//   k_sw_trace2 pass list, the same with 16-row register blocks                                                         [default]
#pragma once
#include <stdint.h>

#define MB_SW_PADRES 21   // matrix row/column used for columns beyond the target end (very negative)

struct SwTraceOut { uint32_t slot; int32_t score, qStart, qEnd, dbStart, dbEnd, flag, pad; };

__device__ __forceinline__ int sw_max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }

// per-query permutation of the hit list by target length, so the 32 lanes of a chunk run for (nearly) the same number of columns
__global__ void __launch_bounds__(128) k_perm_by_len(const MbPrefHit *__restrict__ hits, const uint32_t *__restrict__ hitCount, const uint64_t *__restrict__ toff,
                                                     uint32_t nQ, uint32_t maxHits, uint16_t *__restrict__ perm) {
    __shared__ uint32_t keys[1024];
    for (uint32_t q = blockIdx.x; q < nQ; q += gridDim.x) {
        const uint32_t n = hitCount[q];
        uint32_t n2 = 1; while (n2 < n) n2 <<= 1;
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < n2; k += blockDim.x) {
            uint32_t key = 0xFFFFFFFFu;
            if (k < n) { const uint32_t t = hits[(size_t)q * maxHits + k].seqId; key = ((uint32_t)(toff[t + 1] - toff[t]) << 16) | k; }
            keys[k] = key;
        }
        __syncthreads();
        for (uint32_t kk = 2; kk <= n2; kk <<= 1) for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
                const uint32_t l = ((t / j) * (j << 1)) + (t % j), r = l + j;
                const bool asc = ((l & kk) == 0);
                const uint32_t x = keys[l], y = keys[r];
                if ((x > y) == asc) { keys[l] = y; keys[r] = x; }
            }
            __syncthreads();
        }
        for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) perm[(size_t)q * maxHits + k] = (uint16_t)(keys[k] & 0xFFFFu);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_sw_rows: score-only SW with the DP state in REGISTERS.  One warp per (query, chunk of <= 32 hits), one alignment per lane.
// The query is cut into blocks of RB rows; for one block a lane sweeps all target columns keeping H(previous column) and E for
// the RB rows in registers and F / the diagonal in scalars.  The only state that leaves the registers is the block boundary
// (H of the block's last row and the F entering the next block, one packed word per target column, lane-interleaved in global
// memory, touched once per RB cells).  Substitution scores come from a per-warp query profile in shared memory,
// P[t][j] = mat[t][q_j] + compBias8[j] as int8, fetched 16 rows at a time with 128-bit loads -> no per-cell shared-memory access.
// ---------------------------------------------------------------------------------------------------------------------
#define MB_SWR_WARPS 8
// one block of RBK query rows against all columns of this lane's target; state in registers, block boundary through myBnd
template <int RBK>
__device__ __forceinline__ void sw_sweep_block(const int8_t *profB, uint32_t profStride, const uint8_t *__restrict__ t, int tLen, uint32_t *myBnd,
                                               bool hasTop, bool hasBottom, int gapO, int negE, int &mx) {
    int Hp[RBK], E[RBK];
#pragma unroll
    for (int r = 0; r < RBK; r++) { Hp[r] = 0; E[r] = 0; }
    int prevTop = 0;
    uint32_t topNext = (hasTop && tLen > 0) ? myBnd[0] : 0u;
    uint32_t tcNext = tLen > 0 ? (uint32_t)t[0] : 0u;
    for (int i = 0; i < tLen; i++) {
        const uint32_t top = topNext, tc = tcNext;
        if (i + 1 < tLen) { tcNext = (uint32_t)t[i + 1]; if (hasTop) topNext = myBnd[(size_t)(i + 1) * 32u]; }
        const uint4 *pp = reinterpret_cast<const uint4 *>(profB + tc * profStride);
        uint32_t pw[RBK / 4];
#pragma unroll
        for (int v = 0; v < RBK / 16; v++) { const uint4 x = pp[v]; pw[4 * v] = x.x; pw[4 * v + 1] = x.y; pw[4 * v + 2] = x.z; pw[4 * v + 3] = x.w; }
        const int hTop = (int)(top & 0xFFFFu);
        int F = (int)(top >> 16);          // F entering the block's first row (already max(F-ge, H-go, 0) of the row above)
        int diag = prevTop; prevTop = hTop;
        int gUp = 0;
#pragma unroll
        for (int r = 0; r < RBK; r++) {
            const int s = (int)(int8_t)(pw[r >> 2] >> (8 * (r & 3)));
            const int f = (r == 0) ? F : __viaddmax_s32(F, negE, gUp);
            const int t1 = __viaddmax_s32(diag, s, E[r]);
            const int h = __vimax_s32_relu(t1, f);
            diag = Hp[r]; Hp[r] = h;
            const int g = h - gapO;
            E[r] = __viaddmax_s32_relu(E[r], negE, g);
            F = f; gUp = g;
            mx = max(mx, h);
        }
        if (hasBottom) myBnd[(size_t)i * 32u] = (uint32_t)Hp[RBK - 1] | ((uint32_t)__viaddmax_s32_relu(F, negE, gUp) << 16);
    }
}
template <int RB>
__global__ void __launch_bounds__(MB_SWR_WARPS * 32) k_sw_rows(
    const uint32_t *__restrict__ order, const uint32_t *__restrict__ chunkOff, uint32_t pBegin, uint32_t pEnd,
    const MbPrefHit *__restrict__ hits, const uint32_t *__restrict__ hitCount, const uint16_t *__restrict__ perm, uint32_t maxHits,
    const uint8_t *__restrict__ qres, const uint64_t *__restrict__ qoff, const int8_t *__restrict__ cb,
    const uint8_t *__restrict__ tseq, const uint64_t *__restrict__ toff, const int16_t *__restrict__ alnMat, int gapO, int gapE,
    uint32_t profStride, const int32_t *__restrict__ minScore, int32_t *__restrict__ scoreOut, uint32_t *__restrict__ passList, uint32_t *nPass,
    uint32_t passCap, uint32_t *counter, unsigned long long *cellCount, uint32_t *errFlag, uint32_t *__restrict__ bnd, uint32_t maxT) {
    MB_DYN_SHARED(dynProf);
    __shared__ int8_t sMat8[448];
    for (int i = threadIdx.x; i < 441; i += blockDim.x) sMat8[i] = (int8_t)alnMat[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    int8_t *prof = reinterpret_cast<int8_t *>(dynProf) + (size_t)wib * 21u * profStride;
    uint32_t *myBnd = bnd + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * maxT * 32u + lane;
    const uint32_t chunkBase = chunkOff[pBegin], chunkEnd = chunkOff[pEnd];
    const int negE = -gapE;
    unsigned long long cells = 0;
    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = chunkBase + atomicAdd(counter, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= chunkEnd) break;
        uint32_t lo = pBegin, hi = pEnd;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (chunkOff[mid] <= chunk) lo = mid; else hi = mid; }
        const uint32_t qi = order[lo];
        const uint32_t kk = (chunk - chunkOff[lo]) * 32u + lane;
        const uint32_t cnt = hitCount[qi];
        const bool active = kk < cnt;
        const uint64_t qb = qoff[qi]; const int L = (int)(qoff[qi + 1] - qb);
        const int nb = (L + RB - 1) / RB;
        __syncwarp();
        for (uint32_t idx = lane; idx < 21u * profStride; idx += 32u) {
            const uint32_t t = idx / profStride, j = idx - t * profStride;
            prof[idx] = (j < (uint32_t)L) ? (int8_t)(sMat8[t * 21 + qres[qb + j]] + cb[qb + j]) : (int8_t)-64;   // rows past the query end can only lower scores
        }
        __syncwarp();
        const uint32_t k = (active && perm) ? (uint32_t)perm[(size_t)qi * maxHits + kk] : kk;
        const size_t slot = (size_t)qi * maxHits + k;
        int tLen = 0; const uint8_t *t = tseq;
        if (active) { const uint32_t tid = hits[slot].seqId; const uint64_t tb = toff[tid]; tLen = (int)(toff[tid + 1] - tb); t = tseq + tb; }
        int mx = 0;
        for (int b = 0; b < nb; b++) {
            const bool hasTop = b > 0, hasBottom = b < nb - 1;
            if (L - b * RB <= RB / 2) sw_sweep_block<RB / 2>(prof + b * RB, profStride, t, tLen, myBnd, hasTop, hasBottom, gapO, negE, mx);   // short tail block
            else sw_sweep_block<RB>(prof + b * RB, profStride, t, tLen, myBnd, hasTop, hasBottom, gapO, negE, mx);
            cells += (unsigned long long)tLen * ((L - b * RB <= RB / 2) ? RB / 2 : RB);
        }
        if (active) {
            scoreOut[slot] = mx;
            if (mx >= 65000) atomicExch(errFlag, 1u);
            if (mx >= minScore[L]) { const uint32_t idx = atomicAdd(nPass, 1u); if (idx < passCap) passList[idx] = (uint32_t)slot; else atomicExch(errFlag, 2u); }
        }
    }
    if (cells) atomicAdd(cellCount, cells);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_sw_rows2: as k_sw_rows, but every lane aligns TWO targets at once in packed signed 16-bit halves (DPX s16x2 instructions:
// VIADDMNMX.S16x2 etc.), and a warp serves two (query, chunk of 32 hits) work items, one per half-warp, so short hit lists
// waste fewer lanes.  Hits k = 2*lane16, 2*lane16+1 of the length-sorted chunk share a lane (similar target lengths).
// Scores must stay below 32000 (checked, fail loudly).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sw_pack_scores(uint32_t wa, uint32_t wb, int k) {
    // halfword 0 = sign-extended byte k of wa, halfword 1 = sign-extended byte k of wb  (PRMT with msb replication)
    // (__byte_perm masks the selector nibbles to 3 bits, so the sign-replicating form needs the PTX instruction itself)
    const uint32_t sel = (uint32_t)k | ((uint32_t)(k | 8) << 4) | ((uint32_t)(4 + k) << 8) | ((uint32_t)((4 + k) | 8) << 12);
    uint32_t d;
#ifdef MB200_SIMT_EMU
    d = simt_prmt(wa, wb, sel);      // (tests/simt: the same instruction restated in C)
#else
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(wa), "r"(wb), "r"(sel));
#endif
    return d;
}

template <int RBK>
__device__ __forceinline__ void sw_sweep_block2(const int8_t *profB, uint32_t profStride, const uint8_t *__restrict__ tA, int lenA,
                                                const uint8_t *__restrict__ tB, int lenB, uint2 *myBnd, bool hasTop, bool hasBottom,
                                                uint32_t go2, uint32_t negE2, uint32_t &mx) {
    uint32_t Hp[RBK], E[RBK];
#pragma unroll
    for (int r = 0; r < RBK; r++) { Hp[r] = 0; E[r] = 0; }
    const int n = lenA > lenB ? lenA : lenB;
    const uint32_t MIN2 = 0x80008000u;
    uint32_t prevTop = 0;
    uint2 topNext = make_uint2(0u, 0u);
    if (hasTop && n > 0) topNext = myBnd[0];
    uint32_t taNext = lenA > 0 ? (uint32_t)tA[0] : 21u, tbNext = lenB > 0 ? (uint32_t)tB[0] : 21u;
    for (int i = 0; i < n; i++) {
        const uint2 top = topNext; const uint32_t ta = taNext, tb = tbNext;
        if (i + 1 < n) {
            taNext = (i + 1 < lenA) ? (uint32_t)tA[i + 1] : 21u; tbNext = (i + 1 < lenB) ? (uint32_t)tB[i + 1] : 21u;
            if (hasTop) topNext = myBnd[(size_t)(i + 1) * 32u];
        }
        const uint4 *pa = reinterpret_cast<const uint4 *>(profB + ta * profStride);
        const uint4 *pb = reinterpret_cast<const uint4 *>(profB + tb * profStride);
        uint32_t wa[RBK / 4], wb[RBK / 4];
#pragma unroll
        for (int v = 0; v < RBK / 16; v++) {
            const uint4 x = pa[v], y = pb[v];
            wa[4 * v] = x.x; wa[4 * v + 1] = x.y; wa[4 * v + 2] = x.z; wa[4 * v + 3] = x.w;
            wb[4 * v] = y.x; wb[4 * v + 1] = y.y; wb[4 * v + 2] = y.z; wb[4 * v + 3] = y.w;
        }
        const uint32_t hTop = top.x;
        uint32_t F = top.y;
        uint32_t diag = prevTop; prevTop = hTop;
        uint32_t gUp = 0;
#pragma unroll
        for (int r = 0; r < RBK; r++) {
            const uint32_t s = sw_pack_scores(wa[r >> 2], wb[r >> 2], r & 3);
            const uint32_t f = (r == 0) ? F : __viaddmax_s16x2(F, negE2, gUp);
            const uint32_t t1 = __viaddmax_s16x2(diag, s, E[r]);
            const uint32_t h = __vimax_s16x2_relu(t1, f);
            diag = Hp[r]; Hp[r] = h;
            const uint32_t g = __viaddmax_s16x2(h, go2, MIN2);      // h - gapOpen (go2 holds -gapOpen in both halves)
            E[r] = __viaddmax_s16x2_relu(E[r], negE2, g);
            F = f; gUp = g;
            mx = __vmaxs2(mx, h);
        }
        if (hasBottom) myBnd[(size_t)i * 32u] = make_uint2(Hp[RBK - 1], __viaddmax_s16x2_relu(F, negE2, gUp));
    }
}

template <int RB>
__global__ void __launch_bounds__(MB_SWR_WARPS * 32) k_sw_rows2(
    const uint32_t *__restrict__ order, const uint32_t *__restrict__ chunkOff, uint32_t pBegin, uint32_t pEnd,
    const MbPrefHit *__restrict__ hits, const uint32_t *__restrict__ hitCount, const uint16_t *__restrict__ perm, uint32_t maxHits,
    const uint8_t *__restrict__ qres, const uint64_t *__restrict__ qoff, const int8_t *__restrict__ cb,
    const uint8_t *__restrict__ tseq, const uint64_t *__restrict__ toff, const int16_t *__restrict__ alnMat, int gapO, int gapE,
    uint32_t profStride, const int32_t *__restrict__ minScore, int32_t *__restrict__ scoreOut, uint32_t *__restrict__ passList, uint32_t *nPass,
    uint32_t passCap, uint32_t *counter, unsigned long long *cellCount, uint32_t *errFlag, uint2 *__restrict__ bnd, uint32_t maxT) {
    MB_DYN_SHARED(dynProf);
    __shared__ int8_t sMat8[448];
    for (int i = threadIdx.x; i < 441; i += blockDim.x) sMat8[i] = (int8_t)alnMat[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15u;
    int8_t *prof = reinterpret_cast<int8_t *>(dynProf) + (size_t)(wib * 2u + half) * 22u * profStride;   // row 21 = padding residue
    uint2 *myBnd = bnd + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * maxT * 32u + lane;
    const uint32_t chunkBase = chunkOff[pBegin], chunkEnd = chunkOff[pEnd];
    const uint32_t negE2 = (uint32_t)(uint16_t)(int16_t)(-gapE) * 0x10001u, go2 = (uint32_t)(uint16_t)(int16_t)(-gapO) * 0x10001u;
    unsigned long long cells = 0;
    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = chunkBase + atomicAdd(counter, 2u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0) + half;          // half-warp 0 / 1 take consecutive work items
        if (__all_sync(0xffffffffu, chunk >= chunkEnd)) break;
        const bool haveItem = chunk < chunkEnd;
        uint32_t qi = 0, kk0 = 0, cnt = 0; int L = 0; uint64_t qb = 0;
        if (haveItem) {
            uint32_t lo = pBegin, hi = pEnd;
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (chunkOff[mid] <= chunk) lo = mid; else hi = mid; }
            qi = order[lo]; kk0 = (chunk - chunkOff[lo]) * 32u + l16 * 2u; cnt = hitCount[qi];
            qb = qoff[qi]; L = (int)(qoff[qi + 1] - qb);
        }
        const int nb = (L + RB - 1) / RB;
        __syncwarp();
        if (haveItem) for (uint32_t idx = l16; idx < 22u * profStride; idx += 16u) {
            const uint32_t t = idx / profStride, j = idx - t * profStride;
            prof[idx] = (t < 21u && j < (uint32_t)L) ? (int8_t)(sMat8[t * 21 + qres[qb + j]] + cb[qb + j]) : (int8_t)-64;
        }
        __syncwarp();
        const bool actA = haveItem && kk0 < cnt, actB = haveItem && kk0 + 1 < cnt;
        size_t slotA = 0, slotB = 0; int lenA = 0, lenB = 0; const uint8_t *tA = tseq, *tB = tseq;
        if (actA) { const uint32_t k = perm ? (uint32_t)perm[(size_t)qi * maxHits + kk0] : kk0; slotA = (size_t)qi * maxHits + k;
                    const uint32_t tid = hits[slotA].seqId; const uint64_t tb = toff[tid]; lenA = (int)(toff[tid + 1] - tb); tA = tseq + tb; }
        if (actB) { const uint32_t k = perm ? (uint32_t)perm[(size_t)qi * maxHits + kk0 + 1] : kk0 + 1; slotB = (size_t)qi * maxHits + k;
                    const uint32_t tid = hits[slotB].seqId; const uint64_t tb = toff[tid]; lenB = (int)(toff[tid + 1] - tb); tB = tseq + tb; }
        uint32_t mx = 0;
        for (int b = 0; b < nb; b++) {
            const bool hasTop = b > 0, hasBottom = b < nb - 1;
            if (L - b * RB <= RB / 2) sw_sweep_block2<RB / 2>(prof + b * RB, profStride, tA, lenA, tB, lenB, myBnd, hasTop, hasBottom, go2, negE2, mx);
            else sw_sweep_block2<RB>(prof + b * RB, profStride, tA, lenA, tB, lenB, myBnd, hasTop, hasBottom, go2, negE2, mx);
            cells += (unsigned long long)(lenA + lenB) * ((L - b * RB <= RB / 2) ? RB / 2 : RB);
        }
        const int sA = (int)(int16_t)(mx & 0xFFFFu), sB = (int)(int16_t)(mx >> 16);
        if (actA) { scoreOut[slotA] = sA; if (sA >= 32000) atomicExch(errFlag, 1u);
                    if (sA >= minScore[L]) { const uint32_t idx = atomicAdd(nPass, 1u); if (idx < passCap) passList[idx] = (uint32_t)slotA; else atomicExch(errFlag, 2u); } }
        if (actB) { scoreOut[slotB] = sB; if (sB >= 32000) atomicExch(errFlag, 1u);
                    if (sB >= minScore[L]) { const uint32_t idx = atomicAdd(nPass, 1u); if (idx < passCap) passList[idx] = (uint32_t)slotB; else atomicExch(errFlag, 2u); } }
    }
    if (cells) atomicAdd(cellCount, cells);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_sw_lanes: the packed arithmetic of k_sw_rows2 with ONE work item per LANE: (query, two of its hits).  Six-frame fragments of
// metagenomic contigs are short (mean 37 residues) and most of them have a handful of hits, so a half-warp per query leaves 14 of 16 lanes
// idle (C4: 2 hits per fragment on average).  Here every lane builds its own query profile in shared memory (22 x STRIDE bytes, STRIDE = 32
// or 64 rows) and sweeps its own pair of targets; a warp takes 32 consecutive items of the length-sorted item list, so its lanes run the
// same number of row blocks.  Queries with more than laneMaxHits hits or more than 64 residues stay with k_sw_rows2.
// ---------------------------------------------------------------------------------------------------------------------
template <int STRIDE>
__global__ void __launch_bounds__(STRIDE == 32 ? 256 : 128) k_sw_lanes(
    const uint32_t *__restrict__ order, const uint32_t *__restrict__ laneOff, uint32_t pBegin, uint32_t pEnd,
    const MbPrefHit *__restrict__ hits, const uint32_t *__restrict__ hitCount, const uint16_t *__restrict__ perm, uint32_t maxHits,
    const uint8_t *__restrict__ qres, const uint64_t *__restrict__ qoff, const int8_t *__restrict__ cb,
    const uint8_t *__restrict__ tseq, const uint64_t *__restrict__ toff, const int16_t *__restrict__ alnMat, int gapO, int gapE,
    const int32_t *__restrict__ minScore, int32_t *__restrict__ scoreOut, uint32_t *__restrict__ passList, uint32_t *nPass,
    uint32_t passCap, uint32_t *counter, unsigned long long *cellCount, uint32_t *errFlag, uint2 *__restrict__ bnd, uint32_t maxT) {
    constexpr int RB = 32;
    constexpr uint32_t LANE_BYTES = 22u * STRIDE + 16u;       // an odd multiple of 16 bytes: the 128-bit profile loads of 8 neighbouring lanes fall into 8 different bank groups
    MB_DYN_SHARED(dynProf);
    __shared__ int8_t sMat8[448];
    for (int i = threadIdx.x; i < 441; i += blockDim.x) sMat8[i] = (int8_t)alnMat[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    int8_t *prof = reinterpret_cast<int8_t *>(dynProf) + (size_t)threadIdx.x * LANE_BYTES;
    uint2 *myBnd = bnd + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * maxT * 32u + lane;
    const uint32_t itemBase = laneOff[pBegin], itemEnd = laneOff[pEnd];
    const uint32_t negE2 = (uint32_t)(uint16_t)(int16_t)(-gapE) * 0x10001u, go2 = (uint32_t)(uint16_t)(int16_t)(-gapO) * 0x10001u;
    unsigned long long cells = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = itemBase + atomicAdd(counter, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= itemEnd) break;
        const uint32_t item = base + lane;
        if (item < itemEnd) {
            uint32_t lo = pBegin, hi = pEnd;
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (laneOff[mid] <= item) lo = mid; else hi = mid; }
            const uint32_t qi = order[lo], kk0 = (item - laneOff[lo]) * 2u, cnt = hitCount[qi];
            const uint64_t qb = qoff[qi]; const int L = (int)(qoff[qi + 1] - qb);
            // this lane's profile: prof[t][j] = mat[t][q_j] + compBias8[j]; rows past the query end and the padding residue can only lower scores
            for (int j4 = 0; j4 < STRIDE; j4 += 4) {
                uint32_t qv[4]; int cv[4];
#pragma unroll
                for (int x = 0; x < 4; x++) { const bool in = j4 + x < L; qv[x] = in ? (uint32_t)qres[qb + j4 + x] : 0u; cv[x] = in ? (int)cb[qb + j4 + x] : 0; }
                for (uint32_t t = 0; t < 22u; t++) {
                    uint32_t w = 0;
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        const int v = (t < 21u && j4 + x < L) ? (int)(int8_t)(sMat8[t * 21u + qv[x]] + cv[x]) : -64;
                        w |= (uint32_t)(uint8_t)(int8_t)v << (8 * x);
                    }
                    *reinterpret_cast<uint32_t *>(prof + t * STRIDE + j4) = w;
                }
            }
            const bool actB = kk0 + 1 < cnt;
            size_t slotA, slotB = 0; int lenA, lenB = 0; const uint8_t *tA, *tB = tseq;
            { const uint32_t k = perm ? (uint32_t)perm[(size_t)qi * maxHits + kk0] : kk0; slotA = (size_t)qi * maxHits + k;
              const uint32_t tid = hits[slotA].seqId; const uint64_t tb = toff[tid]; lenA = (int)(toff[tid + 1] - tb); tA = tseq + tb; }
            if (actB) { const uint32_t k = perm ? (uint32_t)perm[(size_t)qi * maxHits + kk0 + 1] : kk0 + 1; slotB = (size_t)qi * maxHits + k;
                        const uint32_t tid = hits[slotB].seqId; const uint64_t tb = toff[tid]; lenB = (int)(toff[tid + 1] - tb); tB = tseq + tb; }
            uint32_t mx = 0;
            const int nb = (L + RB - 1) / RB;
            for (int b = 0; b < nb; b++) {
                const bool hasTop = b > 0, hasBottom = b < nb - 1;
                if (L - b * RB <= RB / 2) sw_sweep_block2<RB / 2>(prof + b * RB, STRIDE, tA, lenA, tB, lenB, myBnd, hasTop, hasBottom, go2, negE2, mx);
                else sw_sweep_block2<RB>(prof + b * RB, STRIDE, tA, lenA, tB, lenB, myBnd, hasTop, hasBottom, go2, negE2, mx);
                cells += (unsigned long long)(lenA + lenB) * ((L - b * RB <= RB / 2) ? RB / 2 : RB);
            }
            const int sA = (int)(int16_t)(mx & 0xFFFFu), sB = (int)(int16_t)(mx >> 16);
            { scoreOut[slotA] = sA; if (sA >= 32000) atomicExch(errFlag, 1u);
              if (sA >= minScore[L]) { const uint32_t idx = atomicAdd(nPass, 1u); if (idx < passCap) passList[idx] = (uint32_t)slotA; else atomicExch(errFlag, 2u); } }
            if (actB) { scoreOut[slotB] = sB; if (sB >= 32000) atomicExch(errFlag, 1u);
                        if (sB >= minScore[L]) { const uint32_t idx = atomicAdd(nPass, 1u); if (idx < passCap) passList[idx] = (uint32_t)slotB; else atomicExch(errFlag, 2u); } }
        }
        __syncwarp();
    }
    if (cells) atomicAdd(cellCount, cells);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_sw_wave: the packed arithmetic with the ROW BLOCKS of one alignment spread over neighbouring lanes.  k_sw_rows2 gives a lane one pair of
// targets and lets it sweep the query's 32-row blocks one after the other (block boundaries through global memory); a query with few hits then
// occupies few of the 16 lanes of its half-warp.  Here a pair of targets is aligned by nb = ceil(L / 32) lanes at once: lane b owns rows
// [32 b, 32 b + 32) and runs b columns behind lane b - 1, whose bottom row (H, F) of the column it needs arrives by one shuffle per step -- a
// systolic wavefront along the query.  A half-warp (one query profile in shared memory) therefore carries 16 / nb pairs = 2 * (16 / nb) hits per
// item: 16 pairs for L <= 32 (the k_sw_rows2 layout), 4 pairs for L = 100, one pair for L in (256, 512]; queries of 513..1024 residues take a whole warp
// per item (G = 32 lanes, one pair).  No boundary traffic to global memory.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MB_SWR_WARPS * 32) k_sw_wave(
    const uint32_t *__restrict__ order, const uint32_t *__restrict__ waveOff, uint32_t pBegin, uint32_t pEnd,
    const MbPrefHit *__restrict__ hits, const uint32_t *__restrict__ hitCount, const uint16_t *__restrict__ perm, uint32_t maxHits,
    const uint8_t *__restrict__ qres, const uint64_t *__restrict__ qoff, const int8_t *__restrict__ cb,
    const uint8_t *__restrict__ tseq, const uint64_t *__restrict__ toff, const int16_t *__restrict__ alnMat, int gapO, int gapE,
    uint32_t profStride, const int32_t *__restrict__ minScore, int32_t *__restrict__ scoreOut, uint32_t *__restrict__ passList, uint32_t *nPass,
    uint32_t passCap, uint32_t *counter, unsigned long long *cellCount, uint32_t *errFlag, uint32_t G /* lanes per item: 16 (L <= 512) or 32 */) {
    constexpr int RB = 32;
    MB_DYN_SHARED(dynProf);
    __shared__ int8_t sMat8[448];
    for (int i = threadIdx.x; i < 441; i += blockDim.x) sMat8[i] = (int8_t)alnMat[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5, perWarp = 32u / G, half = G == 16u ? lane >> 4 : 0u, l16 = lane & (G - 1u);   // l16: lane within its group
    int8_t *prof = reinterpret_cast<int8_t *>(dynProf) + (size_t)(wib * perWarp + half) * 22u * profStride;   // row 21 = padding residue
    const uint32_t itemBase = waveOff[pBegin], itemEnd = waveOff[pEnd];
    const uint32_t negE2 = (uint32_t)(uint16_t)(int16_t)(-gapE) * 0x10001u, go2 = (uint32_t)(uint16_t)(int16_t)(-gapO) * 0x10001u;
    const uint32_t MIN2 = 0x80008000u;
    unsigned long long cells = 0;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = itemBase + atomicAdd(counter, perWarp);
        item = __shfl_sync(0xffffffffu, item, 0) + half;           // half-warp 0 / 1 take consecutive work items (G = 16)
        if (__all_sync(0xffffffffu, item >= itemEnd)) break;
        const bool haveItem = item < itemEnd;
        uint32_t qi = 0, cnt = 0, nb = 1, kk0 = 0; int L = 0; uint64_t qb = 0;
        if (haveItem) {
            uint32_t lo = pBegin, hi = pEnd;
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (waveOff[mid] <= item) lo = mid; else hi = mid; }
            qi = order[lo]; cnt = hitCount[qi]; qb = qoff[qi]; L = (int)(qoff[qi + 1] - qb);
            nb = (uint32_t)(L + RB - 1) / RB; if (nb == 0) nb = 1;
            kk0 = (item - waveOff[lo]) * 2u * (G / nb);
        }
        __syncwarp();
        if (haveItem) for (uint32_t idx = l16; idx < 22u * profStride; idx += G) {
            const uint32_t t = idx / profStride, j = idx - t * profStride;
            prof[idx] = (t < 21u && j < (uint32_t)L) ? (int8_t)(sMat8[t * 21 + qres[qb + j]] + cb[qb + j]) : (int8_t)-64;
        }
        __syncwarp();
        const uint32_t slotInItem = l16 / nb, b = l16 - slotInItem * nb;     // which pair of the item, which row block of the pair
        const bool valid = haveItem && slotInItem < G / nb;
        const uint32_t kA = kk0 + 2u * slotInItem;
        const bool actA = valid && kA < cnt, actB = valid && kA + 1 < cnt;
        size_t slotA = 0, slotB = 0; int lenA = 0, lenB = 0; const uint8_t *tA = tseq, *tB = tseq;
        if (actA) { const uint32_t k = perm ? (uint32_t)perm[(size_t)qi * maxHits + kA] : kA; slotA = (size_t)qi * maxHits + k;
                    const uint32_t tid = hits[slotA].seqId; const uint64_t tb = toff[tid]; lenA = (int)(toff[tid + 1] - tb); tA = tseq + tb; }
        if (actB) { const uint32_t k = perm ? (uint32_t)perm[(size_t)qi * maxHits + kA + 1] : kA + 1; slotB = (size_t)qi * maxHits + k;
                    const uint32_t tid = hits[slotB].seqId; const uint64_t tb = toff[tid]; lenB = (int)(toff[tid + 1] - tb); tB = tseq + tb; }
        const int n = lenA > lenB ? lenA : lenB;
        const int nSteps = __reduce_max_sync(0xffffffffu, n > 0 ? n + (int)nb - 1 : 0);
        const int8_t *profB = prof + b * RB;
        uint32_t Hp[RB], E[RB];
#pragma unroll
        for (int r = 0; r < RB; r++) { Hp[r] = 0; E[r] = 0; }
        uint32_t prevTop = 0, mx = 0; uint2 bottom = make_uint2(0u, 0u);
        uint32_t taNext = 21u, tbNext = 21u;
        if (b == 0 && n > 0) { taNext = lenA > 0 ? (uint32_t)tA[0] : 21u; tbNext = lenB > 0 ? (uint32_t)tB[0] : 21u; }
        for (int s = 0; s < nSteps; s++) {
            uint2 top;
            top.x = __shfl_up_sync(0xffffffffu, bottom.x, 1); top.y = __shfl_up_sync(0xffffffffu, bottom.y, 1);
            if (b == 0) top = make_uint2(0u, 0u);
            const int i = s - (int)b;
            if (i >= 0 && i < n) {
                const uint32_t ta = taNext, tb = tbNext;
                taNext = (i + 1 < lenA) ? (uint32_t)tA[i + 1] : 21u; tbNext = (i + 1 < lenB) ? (uint32_t)tB[i + 1] : 21u;
                const uint4 *pa = reinterpret_cast<const uint4 *>(profB + ta * profStride);
                const uint4 *pb = reinterpret_cast<const uint4 *>(profB + tb * profStride);
                uint32_t wa[RB / 4], wb[RB / 4];
#pragma unroll
                for (int v = 0; v < RB / 16; v++) {
                    const uint4 x = pa[v], y = pb[v];
                    wa[4 * v] = x.x; wa[4 * v + 1] = x.y; wa[4 * v + 2] = x.z; wa[4 * v + 3] = x.w;
                    wb[4 * v] = y.x; wb[4 * v + 1] = y.y; wb[4 * v + 2] = y.z; wb[4 * v + 3] = y.w;
                }
                uint32_t F = top.y, diag = prevTop, gUp = 0; prevTop = top.x;
#pragma unroll
                for (int r = 0; r < RB; r++) {
                    const uint32_t sc = sw_pack_scores(wa[r >> 2], wb[r >> 2], r & 3);
                    const uint32_t f = (r == 0) ? F : __viaddmax_s16x2(F, negE2, gUp);
                    const uint32_t t1 = __viaddmax_s16x2(diag, sc, E[r]);
                    const uint32_t h = __vimax_s16x2_relu(t1, f);
                    diag = Hp[r]; Hp[r] = h;
                    const uint32_t g = __viaddmax_s16x2(h, go2, MIN2);
                    E[r] = __viaddmax_s16x2_relu(E[r], negE2, g);
                    F = f; gUp = g;
                    mx = __vmaxs2(mx, h);
                }
                bottom = make_uint2(Hp[RB - 1], __viaddmax_s16x2_relu(F, negE2, gUp));
            } else if (i == -1 && n > 0) {      // the step before this lane's first column: fetch its first residues
                taNext = lenA > 0 ? (uint32_t)tA[0] : 21u; tbNext = lenB > 0 ? (uint32_t)tB[0] : 21u;
            }
        }
        if (n > 0) { const int rows = L - (int)b * RB; cells += (unsigned long long)(lenA + lenB) * (unsigned)(rows > RB ? RB : (rows > 0 ? rows : 0)); }
        // maximum over the row blocks of the pair -> the lane of block 0
#pragma unroll
        for (int d = 1; d < 32; d++) { const uint32_t y = __shfl_down_sync(0xffffffffu, mx, d); if (b + (uint32_t)d < nb) mx = __vmaxs2(mx, y); }
        if (b == 0) {
            const int sA = (int)(int16_t)(mx & 0xFFFFu), sB = (int)(int16_t)(mx >> 16);
            if (actA) { scoreOut[slotA] = sA; if (sA >= 32000) atomicExch(errFlag, 1u);
                        if (sA >= minScore[L]) { const uint32_t idx = atomicAdd(nPass, 1u); if (idx < passCap) passList[idx] = (uint32_t)slotA; else atomicExch(errFlag, 2u); } }
            if (actB) { scoreOut[slotB] = sB; if (sB >= 32000) atomicExch(errFlag, 1u);
                        if (sB >= minScore[L]) { const uint32_t idx = atomicAdd(nPass, 1u); if (idx < passCap) passList[idx] = (uint32_t)slotB; else atomicExch(errFlag, 2u); } }
        }
    }
    if (cells) atomicAdd(cellCount, cells);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_sw_trace2: end / start coordinates for the alignments on the pass list, one alignment per lane (every lane its own query and
// target), same register-resident row-block sweep as k_sw_rows (16 rows per block).  A pass returns the best cell under the
// reference's tie rule -- maximum score, then the FIRST column reaching it, then the smallest row in that column
// (StripedSmithWaterman.cpp:873-912) -- tracked as a packed key (score << 16 | 0xFFFF - column).  The reverse pass is the same sweep
// on the reversed query prefix / target prefix: the first reversed column holding the forward score is where the reference's
// early termination stops (:900), so no early exit is needed for identical coordinates.
// ---------------------------------------------------------------------------------------------------------------------
#define MB_TR_RB 16
__device__ __forceinline__ void sw_trace_pass(const int8_t *sMatT /* [22][32] : [q][t] */, const uint8_t *__restrict__ q, const int8_t *__restrict__ cb, int qFirst, int qStep,
                                              int qn, const uint8_t *__restrict__ t, int tFirst, int tStep, int tn, uint32_t *myBnd, int gapO, int negE,
                                              int &bestScore, int &bestCol, int &bestRow) {
    uint32_t bestKey = 0; int bRow = 0;
    const int nb = (qn + MB_TR_RB - 1) / MB_TR_RB;
    for (int b = 0; b < nb; b++) {
        uint32_t QR[MB_TR_RB];   // (q_r * 32) | cb_r << 16 ; rows past the end use the padding row 21
#pragma unroll
        for (int r = 0; r < MB_TR_RB; r++) {
            const int row = b * MB_TR_RB + r;
            uint32_t v = 21u * 32u;
            if (row < qn) { const int qi = qFirst + qStep * row; v = ((uint32_t)q[qi] * 32u) | ((uint32_t)(uint8_t)cb[qi] << 16); }
            QR[r] = v;
        }
        int Hp[MB_TR_RB], E[MB_TR_RB];
#pragma unroll
        for (int r = 0; r < MB_TR_RB; r++) { Hp[r] = 0; E[r] = 0; }
        const bool hasTop = b > 0, hasBottom = b < nb - 1;
        int prevTop = 0;
        for (int i = 0; i < tn; i++) {
            const uint32_t tc = (uint32_t)t[tFirst + tStep * i];
            const uint32_t top = hasTop ? myBnd[(size_t)i * 32u] : 0u;
            const int hTop = (int)(top & 0xFFFFu);
            int F = (int)(top >> 16);
            int diag = prevTop; prevTop = hTop;
            int gUp = 0;
            const uint32_t colKey = 0xFFFFu - (uint32_t)i;
#pragma unroll
            for (int r = 0; r < MB_TR_RB; r++) {
                const int s = (int)sMatT[(QR[r] & 0xFFFFu) + tc] + (int)(int8_t)(QR[r] >> 16);
                const int f = (r == 0) ? F : __viaddmax_s32(F, negE, gUp);
                const int t1 = __viaddmax_s32(diag, s, E[r]);
                const int h = min(__vimax_s32_relu(t1, f), 32767);   // the reference's 16-bit cells saturate (_mm_adds_epi16, StripedSmithWaterman.cpp:942-1214)
                diag = Hp[r]; Hp[r] = h;
                const int g = h - gapO;
                E[r] = __viaddmax_s32_relu(E[r], negE, g);
                F = f; gUp = g;
                const uint32_t key = ((uint32_t)h << 16) | colKey;
                if (key > bestKey) { bestKey = key; bRow = b * MB_TR_RB + r; }
            }
            if (hasBottom) myBnd[(size_t)i * 32u] = (uint32_t)Hp[MB_TR_RB - 1] | ((uint32_t)__viaddmax_s32_relu(F, negE, gUp) << 16);
        }
    }
    bestScore = (int)(bestKey >> 16); bestCol = (int)(0xFFFFu - (bestKey & 0xFFFFu)); bestRow = bRow;
}

__global__ void __launch_bounds__(128) k_sw_trace2(const uint32_t *__restrict__ passList, uint32_t nPass, const MbPrefHit *__restrict__ hits, uint32_t maxHits,
                                                   const uint8_t *__restrict__ qres, const uint64_t *__restrict__ qoff, const int8_t *__restrict__ cb,
                                                   const uint8_t *__restrict__ tseq, const uint64_t *__restrict__ toff, const int16_t *__restrict__ alnMat,
                                                   int gapO, int gapE, uint32_t *__restrict__ bnd /* [warp][maxT][32] */, uint32_t maxT,
                                                   SwTraceOut *__restrict__ outs, uint32_t *counter, unsigned long long *cellCount) {
    __shared__ int8_t sMatT[22 * 32];
    for (int i = threadIdx.x; i < 22 * 32; i += blockDim.x) { const int qq = i >> 5, tt = i & 31; sMatT[i] = (qq < 21 && tt < 21) ? (int8_t)alnMat[tt * 21 + qq] : (int8_t)-64; }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t *myBnd = bnd + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * maxT * 32u + lane;
    const int negE = -gapE;
    unsigned long long cells = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counter, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= nPass) break;
        const uint32_t idx = base + lane;
        if (idx >= nPass) continue;
        const uint32_t slot = passList[idx];
        const uint32_t qi = slot / maxHits;
        const uint32_t tid = hits[slot].seqId;
        const uint8_t *q = qres + qoff[qi]; const int8_t *c = cb + qoff[qi]; const int L = (int)(qoff[qi + 1] - qoff[qi]);
        const uint8_t *t = tseq + toff[tid]; const int tLen = (int)(toff[tid + 1] - toff[tid]);
        int sc, col, row;
        sw_trace_pass(sMatT, q, c, 0, 1, L, t, 0, 1, tLen, myBnd, gapO, negE, sc, col, row);
        cells += (unsigned long long)L * tLen;
        SwTraceOut o; o.slot = slot; o.score = sc; o.qEnd = row; o.dbEnd = col; o.qStart = -1; o.dbStart = -1; o.flag = 0; o.pad = 0;
        if (sc <= 0) { o.dbEnd = -1; o.qEnd = L - 1; }
        else {
            int rs, rcol, rrow;
            sw_trace_pass(sMatT, q, c, row, -1, row + 1, t, col, -1, col + 1, myBnd, gapO, negE, rs, rcol, rrow);
            cells += (unsigned long long)(row + 1) * (col + 1);
            o.dbStart = col - rcol; o.qStart = row - rrow;
            if (rs != sc) o.flag = 1;
        }
        outs[idx] = o;
    }
    if (cells) atomicAdd(cellCount, cells);
}

// exclusive scans over the prefilter result: pairOff[i] (natural query order) and, in length-sorted order, the work lists of the three
// forward kernels.  A query with L residues and n hits goes to
//   k_sw_lanes  (laneOff, items of 2 hits)                     when laneMaxHits != 0, n <= laneMaxHits and L <= laneMaxL,
//   k_sw_wave   (waveOff, items of 2 * (G / ceil(L/32)) hits, G = 16 lanes for L <= 512, else 32) when L <= waveMaxL (0 = never; <= 1024),
//   k_sw_rows*  (chunkOff, chunks of 32 hits)                  otherwise.            laneOff / waveOff may be NULL.
__device__ __forceinline__ int sw_query_class(uint32_t L, uint32_t n, uint32_t laneMaxL, uint32_t laneMaxHits, uint32_t waveMaxL) {
    if (laneMaxHits && n <= laneMaxHits && L <= laneMaxL) return 1;
    if (L <= waveMaxL) return 2;
    return 0;
}
__global__ void __launch_bounds__(MB_CTA_THREADS) k_scan_counts(const uint32_t *__restrict__ hitCount, const uint32_t *__restrict__ order, uint32_t nQ,
                                                                uint32_t *pairOff, uint32_t *chunkOff, uint32_t *laneOff, uint32_t *waveOff, const uint64_t *__restrict__ qoff,
                                                                uint32_t laneMaxL, uint32_t laneMaxHits, uint32_t waveMaxL) {
    __shared__ CtaScratch cs;
    if (blockIdx.x == 0) {
        for (uint32_t i = threadIdx.x; i < nQ; i += blockDim.x) pairOff[i] = hitCount[i];
        __syncthreads();
        cta_exclusive_scan(pairOff, nQ, &cs);
    } else {
        const int which = (int)blockIdx.x - 1;      // 0 chunk, 1 lane, 2 wave
        uint32_t *out = which == 0 ? chunkOff : (which == 1 ? laneOff : waveOff);
        if (!out) return;
        for (uint32_t p = threadIdx.x; p < nQ; p += blockDim.x) {
            const uint32_t q = order[p], n = hitCount[q], L = (uint32_t)(qoff[q + 1] - qoff[q]);
            const int cls = sw_query_class(L, n, laneMaxL, laneMaxHits, waveMaxL);
            uint32_t v = 0;
            if (cls == which) {
                if (which == 0) v = (n + 31u) >> 5;
                else if (which == 1) v = (n + 1u) >> 1;
                else { const uint32_t nb = L ? (L + 31u) >> 5 : 1u, per = 2u * ((nb <= 16u ? 16u : 32u) / nb); v = (n + per - 1u) / per; }
            }
            out[p] = v;
        }
        __syncthreads();
        cta_exclusive_scan(out, nQ, &cs);
    }
}

__global__ void k_compact_hits(const MbPrefHit *__restrict__ hits, const uint32_t *__restrict__ hitCount, const uint32_t *__restrict__ pairOff, uint32_t nQ,
                               uint32_t maxHits, MbPrefHit *__restrict__ out) {
    const uint32_t q = blockIdx.x;
    if (q >= nQ) return;
    const uint32_t n = hitCount[q], b = pairOff[q];
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) out[b + k] = hits[(size_t)q * maxHits + k];
}

// caller-supplied hit lists -> the dense [query][maxHits] layout the SW kernels read
__global__ void k_scatter_hits(const uint32_t *__restrict__ flatSeq, const uint64_t *__restrict__ hitOff, uint32_t nQ, uint32_t maxHits,
                               MbPrefHit *__restrict__ hits, uint32_t *__restrict__ hitCount) {
    const uint32_t q = blockIdx.x;
    if (q >= nQ) return;
    const uint64_t b = hitOff[q]; const uint32_t n = (uint32_t)(hitOff[q + 1] - b);
    if (threadIdx.x == 0) hitCount[q] = n;
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) { MbPrefHit h; h.seqId = flatSeq[b + k]; h.score = 0; h.diag = 0; hits[(size_t)q * maxHits + k] = h; }
}
