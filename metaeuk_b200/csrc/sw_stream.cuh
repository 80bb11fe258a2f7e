// sw_stream.cuh -- gapped Smith-Waterman as a stream of (query, two targets) pairs over persistent lane groups ("slots").
//
// Same recurrence and results as sw_kernels.cuh (StripedSmithWaterman.cpp:309-545 of the reference: score, FIRST target column reaching it,
// smallest query row in that column; reverse pass to the first column whose maximum equals the score).  What changes is how work is laid on
// the machine.  k_sw_wave binds a half-warp to ONE query for the duration of an item: lanes idle when the query has fewer pairs than the
// half-warp has slots, and every lane waits for the longest target of the item (protein lengths N(330, 120): mean / max of an item ~ 0.6).
// Here a warp is 32 / nb independent slots (nb = ceil(L / 32) neighbouring lanes, lane b owns query rows [32 b, 32 b + 32) and runs b columns
// behind lane b - 1: the systolic wavefront of k_sw_wave).  A slot that finishes its pair immediately takes the NEXT pair of the kernel's
// flat pair list -- whatever query it belongs to -- so slots are busy until the list is empty and never wait for a neighbour's longer target.
//   * pair descriptors (query, two hits of similar length, target offsets / lengths, gate score) are produced by k_sw_desc_fwd, one thread
//     per pair, so a refill is one coalesced 32-byte load per pair (a warp fetches 32 descriptors at a time into shared memory);
//   * query profiles (22 x stride bytes, composition bias folded in) live in a small per-warp RING in shared memory, shared by all slots
//     working on that query (MODE 0: queries with many hits -- a handful of queries cover the warp's slots), or one private profile per
//     slot rebuilt at every refill (MODE 1: queries with few hits; MODE 2: the reverse pass);
//   * the forward pass also records the END CELL of every alignment that can pass the e-value gate (column maximum per block; only a strict
//     increase over the slot's best so far that also reaches the gate score scans the 32 rows still in registers), so the start-position
//     kernel no longer recomputes the forward pass: MODE 2 runs the reverse pass only, on a profile of the reversed query prefix
//     [0..qEnd], walking the target backwards from dbEnd and stopping at the first column whose maximum equals the forward score (:900).
// Arithmetic: packed s16x2 DPX as before, with two changes that cut ALU-pipe instructions per cell pair from 7 to 5.5: all H/E/F values carry
// a bias of +16 in both halves, so "H - gapOpen" is one plain 32-bit subtraction (no borrow between the halves: H + 16 >= 16 > gapOpen) that
// the compiler may place on the FMA pipe, the clamp at 0 becomes the third operand of a 3-input max, and the running maximum takes two rows
// per instruction.
#pragma once
#include <stdint.h>

#define MB_SWS_BIAS 16
#define MB_SWS_MAXWARPS 8
#define MB_SWS_QUEUE_BYTES 1024u   // 32 pair descriptors per warp
#define MB_SWS_INFO_BYTES 1024u    // descriptor of the pair every slot is working on
#define MB_SWS_TAG_BYTES 64u       // ring tags (query + 1; 0 = empty)

struct SwPair { uint32_t qi, slotA, slotB, offA, offB, lens, rows, aux; };   // 32 bytes; lens = lenA | lenB << 16; aux = gate score (forward) / forward score (reverse)
struct SwPassRec { uint32_t slot; int32_t score, qEnd, dbEnd; };              // an alignment that passed the gate, with its end cell

__device__ __forceinline__ uint32_t sws_max3(uint32_t a, uint32_t b, uint32_t c) {
#ifdef MB200_SIMT_EMU
    return __vmaxs2(__vmaxs2(a, b), c);
#else
    return __vimax3_s16x2(a, b, c);
#endif
}
__device__ __forceinline__ int sws_lo(uint32_t v) { return (int)(int16_t)(v & 0xFFFFu); }
__device__ __forceinline__ int sws_hi(uint32_t v) { return (int)(int16_t)(v >> 16); }
__device__ __forceinline__ uint32_t sws_pack(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

// ---- pair descriptors of the forward pass: positions [pB, pE) of the length-sorted query order, pairOff = exclusive scan of ceil(hits / 2)
//      over the positions that belong to this kernel (0 elsewhere); descriptor g - pairOff[pB] goes to desc[]
__global__ void __launch_bounds__(256) k_sw_desc_fwd(const uint32_t *__restrict__ order, const uint32_t *__restrict__ pairOff, uint32_t pB, uint32_t pE,
                                                     const MbPrefHit *__restrict__ hits, const uint32_t *__restrict__ hitCount, const uint16_t *__restrict__ perm, uint32_t maxHits,
                                                     const uint64_t *__restrict__ qoff, const uint64_t *__restrict__ toff, const int32_t *__restrict__ minScore,
                                                     SwPair *__restrict__ desc) {
    const uint32_t base = pairOff[pB], end = pairOff[pE];
    for (uint32_t g = base + blockIdx.x * blockDim.x + threadIdx.x; g < end; g += gridDim.x * blockDim.x) {
        uint32_t lo = pB, hi = pE;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (pairOff[mid] <= g) lo = mid; else hi = mid; }
        const uint32_t qi = order[lo], k0 = (g - pairOff[lo]) * 2u, cnt = hitCount[qi];
        const uint32_t L = (uint32_t)(qoff[qi + 1] - qoff[qi]);
        SwPair d; d.qi = qi; d.rows = L; d.aux = (uint32_t)minScore[L];
        { const uint32_t k = perm ? (uint32_t)perm[(size_t)qi * maxHits + k0] : k0; d.slotA = qi * maxHits + k;
          const uint32_t tid = hits[d.slotA].seqId; const uint64_t tb = toff[tid]; d.offA = (uint32_t)tb; d.lens = (uint32_t)(toff[tid + 1] - tb); }
        d.slotB = 0xFFFFFFFFu; d.offB = 0;
        if (k0 + 1 < cnt) { const uint32_t k = perm ? (uint32_t)perm[(size_t)qi * maxHits + k0 + 1] : k0 + 1; d.slotB = qi * maxHits + k;
                            const uint32_t tid = hits[d.slotB].seqId; const uint64_t tb = toff[tid]; d.offB = (uint32_t)tb; d.lens |= (uint32_t)(toff[tid + 1] - tb) << 16; }
        desc[g - base] = d;
    }
}

// ---- descriptors of the reverse pass, one per pass record with a known end cell, binned by the number of 32-row blocks of the query prefix [0..qEnd]
//      into the launch groups of the forward pass (all block counts with the same number of slots per warp share a launch): classes of 1, 2, 3, 4, 5, 6,
//      7..8, 9..10, 11..16 and 17..32 blocks.  Records without an end cell (16-bit range exhausted in the forward kernel) go to the fallback list
//      (k_sw_trace2 recomputes them with 32-bit cells).  desc[c * cap + atomic index]; slotA = index of the record (= index of the output).
#define MB_SWS_REV_CLASSES 10
__device__ __forceinline__ uint32_t sws_group_of(uint32_t nb) { return nb <= 6u ? nb - 1u : (nb <= 8u ? 6u : (nb <= 10u ? 7u : (nb <= 16u ? 8u : 9u))); }
__global__ void __launch_bounds__(256) k_sw_desc_rev(const SwPassRec *__restrict__ recs, uint32_t nRec, const MbPrefHit *__restrict__ hits, uint32_t maxHits,
                                                     const uint64_t *__restrict__ toff, SwPair *__restrict__ desc, uint32_t cap, uint32_t *__restrict__ classCount,
                                                     uint32_t *__restrict__ fallback, uint32_t *__restrict__ nFallback, SwTraceOut *__restrict__ outs) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nRec; r += gridDim.x * blockDim.x) {
        const SwPassRec p = recs[r];
        const uint32_t rows = (uint32_t)p.qEnd + 1u;
        if (p.dbEnd < 0 || rows > 1024u) {     // end cell unknown (16-bit overflow in the forward kernel)
            fallback[atomicAdd(nFallback, 1u)] = p.slot;
            SwTraceOut o; o.slot = 0xFFFFFFFFu; o.score = 0; o.qStart = o.qEnd = o.dbStart = o.dbEnd = -1; o.flag = 0; o.pad = 0; outs[r] = o;   // (dropped by the host; the fallback kernel writes its own)
            continue;
        }
        const uint32_t c = sws_group_of((rows + 31u) >> 5);
        SwPair d; d.qi = p.slot / maxHits; d.slotA = r; d.slotB = p.slot; d.offA = (uint32_t)toff[hits[p.slot].seqId]; d.offB = 0;
        d.lens = (uint32_t)p.dbEnd + 1u; d.rows = rows; d.aux = (uint32_t)p.score;
        desc[(size_t)c * cap + atomicAdd(&classCount[c], 1u)] = d;
    }
}

// profile of one query in shared memory, built by the whole warp: prof[t][j] = mat[t][q_j] + compBias8[j] (int8); rows past the query end and the
// padding residue (row 21) hold -64.  rev: row j is query position rows - 1 - j (the reversed prefix [0..qEnd] of the reverse pass).
__device__ __forceinline__ void sws_build_profile(int8_t *prof, uint32_t stride, uint32_t qi, uint32_t rows, bool rev, const uint8_t *__restrict__ qres,
                                                  const uint64_t *__restrict__ qoff, const int8_t *__restrict__ cb, const int8_t *sMat8, uint32_t lane) {
    const uint64_t qb = qoff[qi];
    for (uint32_t j0 = 0; j0 < stride; j0 += 32u) {
        const uint32_t j = j0 + lane;
        uint32_t q = 0; int c = 0; const bool in = j < rows;
        if (in) { const uint32_t src = rev ? rows - 1u - j : j; q = qres[qb + src]; c = cb[qb + src]; }
        for (uint32_t t = 0; t < 22u; t++) prof[t * stride + j] = (in && t < 21u) ? (int8_t)(sMat8[t * 21u + q] + c) : (int8_t)-64;
    }
}

// merge the candidate end cell (oBest, oCol, oRow) of a LATER row block into (best, col, row), per 16-bit half: higher score wins, then the smaller column
// (equal columns keep the earlier block = the smaller row)
__device__ __forceinline__ void sws_merge_end(uint32_t &best, uint32_t &col, uint32_t &row, uint32_t oBest, uint32_t oCol, uint32_t oRow) {
    const bool tl = sws_lo(oBest) > sws_lo(best) || (sws_lo(oBest) == sws_lo(best) && (oCol & 0xFFFFu) < (col & 0xFFFFu));
    const bool th = sws_hi(oBest) > sws_hi(best) || (sws_hi(oBest) == sws_hi(best) && (oCol >> 16) < (col >> 16));
    const uint32_t m = (tl ? 0xFFFFu : 0u) | (th ? 0xFFFF0000u : 0u);
    best = (oBest & m) | (best & ~m); col = (oCol & m) | (col & ~m); row = (oRow & m) | (row & ~m);
}

template <int MODE>   // 0: forward, shared profile ring;  1: forward, private profile per slot;  2: reverse pass (private profiles, one alignment per slot)
__global__ void __launch_bounds__(MB_SWS_MAXWARPS * 32, 2) k_sw_stream(
    const SwPair *__restrict__ desc, uint32_t nDesc, uint32_t nb, uint32_t stride, uint32_t P, uint32_t warpBytes,
    const uint8_t *__restrict__ qres, const uint64_t *__restrict__ qoff, const int8_t *__restrict__ cb, const uint8_t *__restrict__ tseq,
    const int16_t *__restrict__ alnMat, int gapO, int gapE,
    SwPassRec *__restrict__ passRec, uint32_t *nPass, uint32_t passCap, uint32_t *counter, unsigned long long *cellCount, uint32_t *errFlag,
    SwTraceOut *__restrict__ outs) {
    constexpr int RB = 32;
    const unsigned FULL = 0xffffffffu;
    MB_DYN_SHARED(dyn);
    __shared__ int8_t sMat8[448];
    for (int i = threadIdx.x; i < 441; i += blockDim.x) sMat8[i] = (int8_t)alnMat[i];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    unsigned char *wbase = reinterpret_cast<unsigned char *>(dyn) + (size_t)wib * warpBytes;
    SwPair *queue = reinterpret_cast<SwPair *>(wbase);
    SwPair *slotInfo = reinterpret_cast<SwPair *>(wbase + MB_SWS_QUEUE_BYTES);
    uint32_t *tags = reinterpret_cast<uint32_t *>(wbase + MB_SWS_QUEUE_BYTES + MB_SWS_INFO_BYTES);
    int8_t *profs = reinterpret_cast<int8_t *>(wbase + MB_SWS_QUEUE_BYTES + MB_SWS_INFO_BYTES + MB_SWS_TAG_BYTES);
    if (lane < 16u) tags[lane] = 0u;
    __syncthreads();
    const uint32_t slots = 32u / nb, mySlot = lane / nb, b = lane - mySlot * nb;
    const bool spare = mySlot >= slots;
    const uint32_t entryBytes = 22u * stride + 16u;
    const uint32_t B2 = (uint32_t)MB_SWS_BIAS * 0x10001u;
    const uint32_t negE2 = (uint32_t)(uint16_t)(int16_t)(-gapE) * 0x10001u, GO2 = (uint32_t)gapO * 0x10001u;   // gapO <= MB_SWS_BIAS (checked by the host)
    unsigned long long cells = 0;
    // slot state (uniform over the nb lanes of a slot unless noted)
    bool busy = false;
    uint32_t s = 0, myEntry = 0, offA = 0, offB = 0;
    int n = 0, lenA = 0, lenB = 0, stopAt = 0x7fffffff, lastStep = 0;
    const int8_t *profB = profs;
    uint32_t Hp[RB], E[RB];       // per lane: H of the previous column and E for the 32 rows of this lane's block
#pragma unroll
    for (int r = 0; r < RB; r++) { Hp[r] = B2; E[r] = B2; }
    uint32_t prevTop = B2, botH = B2, botF = B2, mx = B2, best2 = 0, col2 = 0, row2 = 0, taNext = 21u, tbNext = 21u;
    // warp-local queue of descriptors
    uint32_t qN = 0, qPos = 0; bool exhausted = false;
    for (;;) {
        // ---------------- refill: every idle slot takes the next pair of the list
        unsigned idle = __ballot_sync(FULL, !busy && b == 0u && !spare);
        if (idle && !exhausted) {
            while (idle) {
                if (qPos == qN) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(counter, slots);      // as many pairs as the warp has slots: the tail of the list spreads over all warps
                    base = __shfl_sync(FULL, base, 0);
                    if (base >= nDesc) { exhausted = true; break; }
                    qN = nDesc - base < slots ? nDesc - base : slots; qPos = 0;
                    __syncwarp();
                    if (lane < qN) queue[lane] = desc[base + lane];
                    __syncwarp();
                }
                const int sl = __ffs((int)idle) - 1;             // first lane of the slot to fill
                const SwPair d = queue[qPos];
                uint32_t e;
                if (MODE == 0) {
                    unsigned inUse = 0;
                    for (uint32_t k = 0; k < P; k++) if (__ballot_sync(FULL, busy && myEntry == k)) inUse |= 1u << k;
                    int found = -1;
                    for (uint32_t k = 0; k < P; k++) if (tags[k] == d.qi + 1u) found = (int)k;
                    if (found < 0) {
                        const unsigned freeM = ~inUse & ((1u << P) - 1u);
                        if (!freeM) break;                       // every profile of the ring is in use: this slot waits for another one to finish
                        found = __ffs((int)freeM) - 1;
                        __syncwarp();
                        sws_build_profile(profs + (size_t)found * entryBytes, stride, d.qi, d.rows, false, qres, qoff, cb, sMat8, lane);
                        if (lane == 0) tags[found] = d.qi + 1u;
                        __syncwarp();
                    }
                    e = (uint32_t)found;
                } else {
                    e = (uint32_t)sl / nb;
                    __syncwarp();
                    sws_build_profile(profs + (size_t)e * entryBytes, stride, d.qi, d.rows, MODE == 2, qres, qoff, cb, sMat8, lane);
                    __syncwarp();
                }
                qPos++; idle &= idle - 1u;
                if ((int)lane == sl) slotInfo[(uint32_t)sl / nb] = d;
                if ((int)lane >= sl && lane < (uint32_t)sl + nb) {
                    busy = true; s = 0; myEntry = e; stopAt = 0x7fffffff;
                    lenA = (int)(d.lens & 0xFFFFu); lenB = (int)(d.lens >> 16); n = lenA > lenB ? lenA : lenB;
                    lastStep = n + (int)((d.rows + 31u) >> 5) - 1;   // the pair's own number of row blocks (a launch may serve several block counts)
                    offA = d.offA; offB = d.offB;
                    profB = profs + (size_t)e * entryBytes + b * RB;
#pragma unroll
                    for (int r = 0; r < RB; r++) { Hp[r] = B2; E[r] = B2; }
                    prevTop = B2; botH = B2; botF = B2; mx = B2; col2 = 0; row2 = 0;
                    if (MODE == 2) best2 = sws_pack((int)d.aux + MB_SWS_BIAS, 0x7FFF);          // reverse: the score to find (half B never matches)
                    else { const int g = (int)d.aux + MB_SWS_BIAS - 1; best2 = sws_pack(g, g); }   // forward: only column maxima >= the gate score are end-cell candidates
                    taNext = 21u; tbNext = 21u;
                    if (b == 0u) {
                        if (MODE == 2) taNext = tseq[offA + (uint32_t)(lenA - 1)];
                        else { taNext = tseq[offA]; if (lenB > 0) tbNext = tseq[offB]; }
                    }
                }
            }
            __syncwarp();
        }
        if (!__any_sync(FULL, busy)) break;                      // list exhausted and every slot done
        // ---------------- one column step of every busy slot (lane b works on column s - b of its pair)
        uint32_t topH = __shfl_up_sync(FULL, botH, 1), topF = __shfl_up_sync(FULL, botF, 1);
        if (b == 0u) { topH = B2; topF = B2; }
        const int i = (int)s - (int)b;
        bool hit = false;                                        // reverse: this lane saw the forward score in this column for the first time
        if (busy && i >= 0 && i < n) {
            const uint32_t ta = taNext, tb = tbNext;
            if (MODE == 2) { taNext = (i + 1 < lenA) ? (uint32_t)tseq[offA + (uint32_t)(lenA - 2 - i)] : 21u; }
            else { taNext = (i + 1 < lenA) ? (uint32_t)tseq[offA + (uint32_t)(i + 1)] : 21u; tbNext = (i + 1 < lenB) ? (uint32_t)tseq[offB + (uint32_t)(i + 1)] : 21u; }
            const uint4 *pa = reinterpret_cast<const uint4 *>(profB + ta * stride);
            const uint4 *pb = reinterpret_cast<const uint4 *>(profB + tb * stride);
            uint32_t wa[RB / 4], wb[RB / 4];
#pragma unroll
            for (int v = 0; v < RB / 16; v++) {
                const uint4 x = pa[v], y = pb[v];
                wa[4 * v] = x.x; wa[4 * v + 1] = x.y; wa[4 * v + 2] = x.z; wa[4 * v + 3] = x.w;
                wb[4 * v] = y.x; wb[4 * v + 1] = y.y; wb[4 * v + 2] = y.z; wb[4 * v + 3] = y.w;
            }
            uint32_t F = topF, diag = prevTop, gUp = 0, cm = B2; prevTop = topH;
#pragma unroll
            for (int r = 0; r < RB; r++) {
                const uint32_t sc = sw_pack_scores(wa[r >> 2], wb[r >> 2], r & 3);
                const uint32_t f = (r == 0) ? F : __viaddmax_s16x2(F, negE2, gUp);
                const uint32_t t1 = __viaddmax_s16x2(diag, sc, E[r]);
                const uint32_t h = sws_max3(t1, f, B2);
                diag = Hp[r]; Hp[r] = h;
                const uint32_t g = h - GO2;                       // both halves >= 16 > gapO: no borrow across the halves
                E[r] = __viaddmax_s16x2(E[r], negE2, g);
                F = f; gUp = g;
                if (r & 1) cm = sws_max3(cm, Hp[r - 1], h);
            }
            botH = Hp[RB - 1]; botF = __viaddmax_s16x2(F, negE2, gUp);
            mx = __vmaxs2(mx, cm);
            if (MODE == 2) {
                if (stopAt == 0x7fffffff && sws_lo(cm) == sws_lo(best2) && (col2 >> 31) == 0u) {   // first column of this block holding the forward score
                    uint32_t row = 0;
#pragma unroll
                    for (int r = RB - 1; r >= 0; r--) if (((Hp[r] ^ cm) & 0xFFFFu) == 0u) row = (uint32_t)r;
                    col2 = 0x80000000u | (uint32_t)i; row2 = b * RB + row; hit = true;
                }
            } else {
                const bool tl = sws_lo(cm) > sws_lo(best2), th = sws_hi(cm) > sws_hi(best2);
                if (tl | th) {                                   // a new best column maximum that reaches the gate: remember the end cell (smallest row holding it)
                    if (tl) {
                        uint32_t row = 0;
#pragma unroll
                        for (int r = RB - 1; r >= 0; r--) if (((Hp[r] ^ cm) & 0xFFFFu) == 0u) row = (uint32_t)r;
                        best2 = (best2 & 0xFFFF0000u) | (cm & 0xFFFFu); col2 = (col2 & 0xFFFF0000u) | (uint32_t)i; row2 = (row2 & 0xFFFF0000u) | (b * RB + row);
                    }
                    if (th) {
                        uint32_t row = 0;
#pragma unroll
                        for (int r = RB - 1; r >= 0; r--) if (((Hp[r] ^ cm) >> 16) == 0u) row = (uint32_t)r;
                        best2 = (best2 & 0xFFFFu) | (cm & 0xFFFF0000u); col2 = (col2 & 0xFFFFu) | ((uint32_t)i << 16); row2 = (row2 & 0xFFFFu) | ((b * RB + row) << 16);
                    }
                }
            }
        } else if (busy && i == -1) {                            // the step before this lane's first column: fetch its first residues
            if (MODE == 2) taNext = tseq[offA + (uint32_t)(lenA - 1)];
            else { taNext = tseq[offA]; if (lenB > 0) tbNext = tseq[offB]; }
        }
        if (MODE == 2) {
            // the first column holding the score may be found by a later block at a later step: run nb - 1 more steps, then take the smallest column
            const unsigned hb = __ballot_sync(FULL, hit);
            const unsigned mine = (hb >> (lane - b)) & ((nb >= 32u ? 0u : (1u << nb)) - 1u);
            if (busy && mine && stopAt == 0x7fffffff) stopAt = (int)s + (int)nb - 1;
        }
        if (busy) s++;
        // ---------------- finished pairs: combine the row blocks, emit
        const bool fin = busy && ((int)s >= lastStep || (MODE == 2 && (int)s > stopAt));
        if (__any_sync(FULL, fin)) {
            if (busy) { const int done = (int)s - (int)b; const int colsDone = done < 0 ? 0 : (done > n ? n : done);
                        if (fin) { const SwPair &d = slotInfo[mySlot]; const int rows = (int)d.rows - (int)(b * RB);
                                   cells += (unsigned long long)(MODE == 2 ? colsDone : lenA + lenB) * (unsigned)(rows > RB ? RB : (rows > 0 ? rows : 0)); } }
            uint32_t rBest = best2, rCol = col2, rRow = row2, rMx = mx;
            if (MODE == 2 && (col2 >> 31) == 0u) rCol = 0x7FFFFFFFu;      // reverse: no hit in this block
            for (uint32_t dlt = 1; dlt < nb; dlt++) {
                const uint32_t oBest = __shfl_down_sync(FULL, rBest, dlt), oCol = __shfl_down_sync(FULL, rCol, dlt), oRow = __shfl_down_sync(FULL, rRow, dlt);
                const uint32_t oMx = __shfl_down_sync(FULL, rMx, dlt);
                if (b + dlt < nb) {
                    rMx = __vmaxs2(rMx, oMx);
                    if (MODE == 2) { if ((oCol & 0x7FFFFFFFu) < (rCol & 0x7FFFFFFFu)) { rCol = oCol; rRow = oRow; } }   // smallest column; equal columns keep the earlier block (smaller reversed row)
                    else sws_merge_end(rBest, rCol, rRow, oBest, oCol, oRow);
                }
            }
            if (fin && b == 0u) {
                const SwPair d = slotInfo[mySlot];
                if (MODE == 2) {
                    SwTraceOut o; o.slot = d.slotB; o.score = (int)d.aux; o.qEnd = (int)d.rows - 1; o.dbEnd = (int)(d.lens & 0xFFFFu) - 1; o.pad = 0;
                    const bool ok = (rCol >> 31) != 0u && rCol != 0x7FFFFFFFu;
                    o.qStart = ok ? o.qEnd - (int)rRow : -1; o.dbStart = ok ? o.dbEnd - (int)(rCol & 0x7FFFFFFFu) : -1;
                    o.flag = ok ? 0 : 1;                        // the reference EXITs when the reverse pass does not reproduce the forward score (StripedSmithWaterman.cpp:466-473)
                    outs[d.slotA] = o;
                } else {
                    const int sA = sws_lo(rMx) - MB_SWS_BIAS, sB = sws_hi(rMx) - MB_SWS_BIAS, gate = (int)d.aux;
                    if (sA >= gate || sA >= 32000 - MB_SWS_BIAS) {
                        const uint32_t idx = atomicAdd(nPass, 1u);
                        if (idx < passCap) { SwPassRec p; p.slot = d.slotA; p.score = sA; p.qEnd = (int)(rRow & 0xFFFFu); p.dbEnd = (int)(rCol & 0xFFFFu);
                                             if (sA >= 32000 - MB_SWS_BIAS) p.dbEnd = -1;     // 16-bit range exhausted: recomputed with 32-bit cells by the fallback kernel
                                             passRec[idx] = p; }
                        else atomicExch(errFlag, 2u);
                    }
                    if (d.slotB != 0xFFFFFFFFu && (sB >= gate || sB >= 32000 - MB_SWS_BIAS)) {
                        const uint32_t idx = atomicAdd(nPass, 1u);
                        if (idx < passCap) { SwPassRec p; p.slot = d.slotB; p.score = sB; p.qEnd = (int)(rRow >> 16); p.dbEnd = (int)(rCol >> 16);
                                             if (sB >= 32000 - MB_SWS_BIAS) p.dbEnd = -1;
                                             passRec[idx] = p; }
                        else atomicExch(errFlag, 2u);
                    }
                }
            }
            if (fin) busy = false;
        }
    }
    if (cells) atomicAdd(cellCount, cells);
}
