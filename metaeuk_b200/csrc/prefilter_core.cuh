// prefilter_core.cuh -- the per-query k-mer prefilter, one CTA per query fragment.
//
// Reproduces, bit for bit, what the reference computes per query in
//   QueryMatcher::matchQuery / match          reference/lib/mmseqs/src/prefiltering/QueryMatcher.cpp:85-211,213-346
//   KmerGenerator::generateKmerList           .../KmerGenerator.cpp:107-183,186-216
//   CacheFriendlyOperations::findDuplicates   .../CacheFriendlyOperations.cpp:185-274   (order-dependent double-hit rule)
//   UngappedAlignment::computeScores          .../UngappedAlignment.cpp:30-43,331-362
//   keepMaxScoreElementOnly / top-N / sort    .../CacheFriendlyOperations.cpp:350-380, QueryMatcher.cpp:153-209,363-420
// but organised for a GPU: the reference's per-target byte array scan in *arrival order* becomes
//   (1) a 2-bit-per-slot shared-memory occupancy map that separates targets hit once from targets hit >= 2 times,
//   (2) a sort of the (few) multi-hit entries by (target, arrival index) and a per-target run walk.
// Arrival order = order of the reference's databaseHits buffer: (query position, similar-k-mer rank, list order).  An index list holds
// every target at most once and is sorted by target id, so that order is the lexicographic order of (k, seqId) with k = the ordinal of
// the similar k-mer inside the query.  The hit word therefore carries k as its "arrival" field and no per-hit rank is ever computed.
#pragma once
#include "cta_prims.cuh"

#define MB_AL 21
#define MB_XRES 20
#define MB_N3 8000
#define MB_CNT_MIN (-192)
#define MB_CNT_W 384

// hit / candidate word:  seqId[27] | arrival[21] | diag16[16]
#define MB_MAX_SEQID_BITS 27
#define MB_MAX_ARRIVAL_BITS 21
#define MB_W_SEQ(w) ((uint32_t)((w) >> 37))
#define MB_W_ARR(w) ((uint32_t)(((w) >> 16) & 0x1FFFFFu))
#define MB_W_DIAG(w) ((uint16_t)((w) & 0xFFFFu))

enum { MB_ST_OK = 0, MB_ST_TOO_LONG = 1, MB_ST_OVERFLOW = 2 };

struct MbPrefHit { uint32_t seqId; int32_t score; int32_t diag; };

struct MbPrefTables {
    const uint32_t *sim3;      // [8000][8000]  (neighbour 3-mer index << 16) | (uint16)score, descending score, reference order
    const uint16_t *cnt3;      // [8000][MB_CNT_W]  #neighbours with score >= (x + MB_CNT_MIN)
    const uint32_t *offsets;   // [20^6 + 1]
    const uint32_t *kmerBits;  // optional [20^6 / 32] bitmap of k-mers with a non-empty list (L2 resident), or NULL
    const uint2 *entries;      // (seqId, pos_j)
    const uint4 *buckets;      // optional [20^6][2]: one 32-byte sector per k-mer = u16 count | up to 5 inline (seqId lo, seqId hi, pos_j) entries; longer lists: count, first entry index, length
    const uint8_t *tseq;       // masked target residues
    const uint64_t *toff;      // [nT + 1]
    uint32_t nT;
    int kmerThr, maxHits, minDiag;
    uint32_t binMask;          // BINSIZE - 1 of the reference's CacheFriendlyOperations<BINSIZE>
    uint32_t capP, capI, capK, capM;
    uint32_t phaseStats;       // 1 = per-phase cycle accounting (diagnostic runs only)
    uint32_t streamHint;       // bit 0 = the index reads (bucket sectors, long entry lists: random, never reused) bypass L1 and are marked evict-first in L2, so that 100+ GB of them per launch do not
                               //     push the per-CTA work lists out of L2 (ncu: every hit word was written back to DRAM before the next query overwrote it);
                               // bit 1 = they fetch 64 bytes instead of the default 128-byte line from HBM (a random read moves a whole line otherwise: 4.07 DRAM sectors per access)
    uint32_t sortWalkMax;      // multi-hit lists of at most this many entries are sorted (one warp, registers) and walked instead of grouped by hashing (0: measured slower on C2, r02g)
};

struct MbPrefScratch {        // per-CTA global scratch
    uint32_t *itemOff;  // capP + 1
    int16_t *wThr; uint16_t *wI0, *wI1;   // capP
    uint32_t *kmerOff;  // capI + 1
    uint32_t *itemInfo; // capI          p << 13 | i
    uint64_t *A, *B, *C;   // capM each (power of two)
    int32_t *S;            // capM
};

// shared-memory homes of the expansion arrays (queries that exceed them use the global scratch instead)
#define MB_SM_P 512

// Kernel configuration: threads per CTA, CTAs per SM and the shared-memory homes of the per-query work lists.  Two are instantiated
// (mb200.cu): "dense" for indexes whose k-mer lists are long (many hits per query: big occupancy map, big lists) and "sparse" for
// small / sparse indexes (half the threads, twice the CTAs per SM: block barriers of one query overlap with work of three others).
template <int NT_, int CTAS_, int NSLOT_LOG_, int SM_I_, int SM_NZ_LOG_> struct MbPrefCfg {
    static constexpr uint32_t NT = NT_, CTAS = CTAS_;
    static constexpr uint32_t NSLOT = 1u << NSLOT_LOG_, NSLOT_LOG = NSLOT_LOG_;     // 2-bit occupancy slots (target & (NSLOT - 1))
    static constexpr uint32_t SM_I = SM_I_;                 // items (window, first-half neighbour)
    static constexpr uint32_t SM_NZ = 1u << SM_NZ_LOG_;     // non-empty similar k-mers; the same space then holds the hash table / sort buffer
    static constexpr uint32_t SM_S = SM_NZ, HT_LOG_MAX = SM_NZ_LOG_, MUL_CAP = SM_NZ / 4;
};
typedef MbPrefCfg<512, 2, 17, 4096, 12> MbPrefCfgDense;
typedef MbPrefCfg<256, 4, 15, 2048, 10> MbPrefCfgSparse;
typedef MbPrefCfg<256, 4, 16, 2048, 11> MbPrefCfgMedium;   // short fragments against a dense index: four barrier domains per SM, occupancy map and hash table sized for a few thousand hits
typedef MbPrefCfg<128, 8, 15, 1024, 10> MbPrefCfgSmall;    // eight queries per SM, four warps each

template <class CFG> struct MbPrefSharedT {
    typedef CFG Cfg;
    CtaScratch cs;
    uint32_t hist[256];
    uint32_t cnt[8];       // 0 nMulti, 1 nCand, 2 nSurv, 3 thr (overflow flag before P11), 4 nPass, 5 rescale, 6 hits, 7 non-empty k-mers
    long long phaseT; unsigned long long phaseAcc[16];
    int8_t ungMat[448];
    uint32_t slots[CFG::NSLOT / 16];
    uint32_t sItemOff[MB_SM_P + 1]; int16_t sThr[MB_SM_P]; uint16_t sI0[MB_SM_P], sI1[MB_SM_P];
    uint32_t sKmerOff[CFG::SM_I + 1];
    uint16_t sItemIdx0[CFG::SM_I];   // first-half neighbour 3-mer of every item (read once in P2, used by every k-mer of the item in P3)
    union {   // the list of similar k-mers with a non-empty index list is dead once the hits are gathered; the sorts that follow reuse the space
        uint64_t sNz[CFG::SM_NZ];
        uint32_t sHitOff[2 * CFG::SM_NZ];   // hash table of the double-hit stage (keys | group offsets)
        uint64_t sSort[CFG::SM_S];
        struct {                            // double-hit stage entirely in shared memory (multi-hit lists of at most MUL_CAP entries: the usual case)
            uint64_t sMul[CFG::SM_NZ / 4], sGrp[CFG::SM_NZ / 4];
            uint32_t sHt[CFG::SM_NZ];       // keys | group offsets, SM_NZ / 2 slots each
        };
    };
};
typedef MbPrefSharedT<MbPrefCfgDense> MbPrefShared;

// sort a (usually short) global list: through shared memory when it fits
template <class SH>
MB_DEV_CALL void mb_sort_list(uint64_t *g, uint32_t n, SH *sh) {
    if (n <= 256) {
        cta_sort_u64(g, n);      // one warp sorts in registers, straight from and to the list
    } else if (n <= SH::Cfg::SM_S) {
        CTA_FOR(i, n) sh->sSort[i] = g[i];
        CTA_SYNC();
        cta_sort_u64(sh->sSort, n);
        CTA_FOR(i, n) g[i] = sh->sSort[i];
        CTA_SYNC();
    } else {
        cta_sort_u64(g, n);
    }
}

MB_DEV uint32_t mb_ht_find(const uint32_t *htKey, uint32_t htLog, uint32_t id1) {
    const uint32_t mask = (1u << htLog) - 1u;
    uint32_t slot = (id1 * 2654435761u) >> (32 - htLog);
    while (htKey[slot] != id1) slot = (slot + 1u) & mask;
    return slot;
}

// ---- inline buckets: the index entries of a k-mer sit in the same 32-byte sector as their count, so a lookup is ONE random HBM access
//      instead of offsets[k-mer] followed by entries[...] (the gather phase is bound by the rate of random DRAM accesses, not by bytes).
//      16 halfwords: h[0] = list length (0xFFFF: 65535 or more), then either up to MB_BUCKET_INLINE entries of three halfwords
//      (seqId low, seqId high, position) or, for longer lists, h[2..3] = index of the first entry in the classic array, h[4..5] = length.
#define MB_BUCKET_INLINE 5
#if MB_DEVICE_BUILD
// one 256-bit load (LDG.E.256 on sm_100a): the whole sector in a single request
// hint bits (MbPrefTables::streamHint): 1 = L1 no-allocate + L2 evict-first, 2 = fetch 64 B instead of the default 128-byte line from HBM (.L2::64B; needs bit 1).
// Measured and dropped (call t, C4/10): L2 prefetch of a warp's second pair of buckets while the first is in flight, L2 prefetch of a long list as soon as its bucket arrives, all four
// buckets of a lane in flight at once (232 bytes of spills at 64 registers): the kernel time does not move or gets worse.
__device__ __forceinline__ void mb_load_bucket(const uint4 *p, uint4 &a, uint4 &b, uint32_t hint) {
#ifdef MB200_SIMT_EMU
    (void)hint; a = p[0]; b = p[1];
#else
    if ((hint & 3u) == 3u)      // LDG.E.NA.EFL2.LTC64B.256.CONSTANT
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
    else if (hint & 1u)         // LDG.E.NA.EFL2.256.CONSTANT
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
    else
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
#endif
}
// one index entry (seqId, position) of a long list: same policy
__device__ __forceinline__ uint2 mb_load_entry(const uint2 *p, uint32_t hint) {
#ifdef MB200_SIMT_EMU
    (void)hint; return *p;
#else
    uint2 e;
    if (hint & 1u) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(e.x), "=r"(e.y) : "l"(p), "l"(pol));
    } else e = __ldg(p);
    return e;
#endif
}
#endif
MB_DEV uint32_t mb_bucket_h(const uint4 &a, const uint4 &b, int i) {   // halfword i of the 32-byte bucket (i is a compile-time constant at every call site)
    const uint32_t w = (i < 8) ? ((i >> 1) == 0 ? a.x : (i >> 1) == 1 ? a.y : (i >> 1) == 2 ? a.z : a.w)
                               : (((i - 8) >> 1) == 0 ? b.x : ((i - 8) >> 1) == 1 ? b.y : ((i - 8) >> 1) == 2 ? b.z : b.w);
    return (i & 1) ? (w >> 16) : (w & 0xFFFFu);
}
MB_DEV void mb_bucket_entry(const uint4 &a, const uint4 &b, int j, uint32_t &seqId, uint32_t &pos) {
    seqId = mb_bucket_h(a, b, 1 + 3 * j) | (mb_bucket_h(a, b, 2 + 3 * j) << 16); pos = mb_bucket_h(a, b, 3 + 3 * j);
}

struct MbPrefStats { unsigned long long kmers, lookups, dbMatches, doubleHits, diagCells, results; unsigned long long phase[16]; };

MB_DEV uint32_t mb_count_ge(const uint16_t *cnt3, uint32_t row, int x) {
    if (x <= MB_CNT_MIN) return MB_N3;
    if (x >= MB_CNT_MIN + MB_CNT_W) return 0;
    return cnt3[(size_t)row * MB_CNT_W + (x - MB_CNT_MIN)];
}

// UngappedAlignment::computeSingelSequenceScores + scalarDiagonalScoring (UngappedAlignment.cpp:30-43,416-430)
MB_DEV int mb_diag_score(const int8_t *ungMat, const uint8_t *q, const int8_t *corr, uint32_t L, const uint8_t *t, uint32_t tLen, uint16_t diag16,
                         uint32_t *cells) {
    const int diagonal = (int16_t)diag16;
    const uint16_t d1 = (uint16_t)(0 - diag16), d2 = diag16;
    const uint32_t dist = d1 < d2 ? d1 : d2;
    uint32_t n, qo, to;
    if (diagonal >= 0 && dist < L) { n = (tLen < L - dist) ? tLen : (L - dist); qo = dist; to = 0; }
    else if (diagonal < 0 && dist < tLen) { n = (tLen - dist < L) ? (tLen - dist) : L; qo = 0; to = dist; }
    else { *cells = 0; return 0; }
    int mx = 0, score = 0;
    for (uint32_t pos = 0; pos < n; pos++) {
        const int curr = (int8_t)(ungMat[q[qo + pos] * MB_AL + t[to + pos]] + corr[qo + pos]);
        score += curr; score = score < 0 ? 0 : score; mx = score > mx ? score : mx;
    }
    *cells = n;
    return mx;
}

#if MB_DEVICE_BUILD
// The same recurrence evaluated by a whole warp: every lane folds a contiguous piece of the diagonal into the transfer
// function of "running sum clamped at 0" -- (S, R, P, B): out(x) = max(x + S, R), best(x) = max(x + P, B) -- and the pieces are
// combined left to right with the associative rule  S = Sa+Sb, R = max(Ra+Sb, Rb), P = max(Pa, Sa+Pb), B = max(Ba, Ra+Pb, Bb).
__device__ __forceinline__ int mb_diag_score_warp(const int8_t *ungMat, const uint8_t *q, const int8_t *corr, uint32_t L, const uint8_t *t, uint32_t tLen,
                                                  uint16_t diag16, uint32_t *cells) {
    const int diagonal = (int16_t)diag16;
    const uint16_t d1 = (uint16_t)(0 - diag16), d2 = diag16;
    const uint32_t dist = d1 < d2 ? d1 : d2;
    uint32_t n, qo, to;
    if (diagonal >= 0 && dist < L) { n = (tLen < L - dist) ? tLen : (L - dist); qo = dist; to = 0; }
    else if (diagonal < 0 && dist < tLen) { n = (tLen - dist < L) ? (tLen - dist) : L; qo = 0; to = dist; }
    else { *cells = 0; return 0; }
    const uint32_t lane = threadIdx.x & 31u, chunk = (n + 31u) >> 5;
    uint32_t b = lane * chunk; if (b > n) b = n;
    uint32_t e = b + chunk; if (e > n) e = n;
    const int NEG = -(1 << 28);
    int S = 0, R = 0, P = NEG, B = 0;
    for (uint32_t pos = b; pos < e; pos++) {
        const int v = (int8_t)(ungMat[q[qo + pos] * MB_AL + t[to + pos]] + corr[qo + pos]);
        S += v; R += v; R = R < 0 ? 0 : R; P = S > P ? S : P; B = R > B ? R : B;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int Sb = __shfl_down_sync(0xffffffffu, S, d), Rb = __shfl_down_sync(0xffffffffu, R, d);
        const int Pb = __shfl_down_sync(0xffffffffu, P, d), Bb = __shfl_down_sync(0xffffffffu, B, d);
        if ((lane & (2 * d - 1)) == 0) {
            const int nB = max(max(B, R + Pb), Bb), nP = max(P, S + Pb), nR = max(R + Sb, Rb);
            B = nB; P = nP < NEG ? NEG : nP; R = nR; S += Sb;
        }
    }
    *cells = n;
    return __shfl_sync(0xffffffffu, B, 0);
}
#endif

// per-phase cycle accounting (thread 0 of the CTA; sums over all queries land in MbPrefStats::phase)
#if MB_DEVICE_BUILD
#define MB_PHASE(n) do { if (T.phaseStats && threadIdx.x == 0) { const long long t_ = clock64(); sh->phaseAcc[n] += (unsigned long long)(t_ - sh->phaseT); sh->phaseT = t_; } } while (0)
#else
#define MB_PHASE(n) ((void)0)
#endif

// P1..P3 are instantiated per home of the expansion arrays (SMP: window arrays in shared memory, SMI: item arrays in shared memory) so that
// every access compiles to a shared (LDS/STS) or a global (LDG/STG) instruction instead of a generic one.
// ---- P1 windows: threshold and first-level neighbour count  (QueryMatcher.cpp:225-244, KmerGenerator.cpp:107-140); returns #items
template <bool SMP, class SH>
MB_DEV uint32_t mb_pref_windows(const MbPrefTables &T, const MbPrefScratch &W, SH *sh, const uint8_t *q, const float *bias, uint32_t P) {
    uint32_t *const itemOff = SMP ? sh->sItemOff : W.itemOff; int16_t *const wThr = SMP ? sh->sThr : W.wThr;
    uint16_t *const wI0 = SMP ? sh->sI0 : W.wI0; uint16_t *const wI1 = SMP ? sh->sI1 : W.wI1;
    CTA_FOR(p, P) {
        const uint8_t *w = q + p;
        const uint32_t r0 = w[0], r1 = w[1], r2 = w[3], r3 = w[5], r4 = w[8], r5 = w[9];
        float b = 0.0f;
        b = mb_fadd(b, bias[p + 0]); b = mb_fadd(b, bias[p + 1]); b = mb_fadd(b, bias[p + 3]);
        b = mb_fadd(b, bias[p + 5]); b = mb_fadd(b, bias[p + 8]); b = mb_fadd(b, bias[p + 9]);
        uint32_t n0 = 0; int thr = 0; uint32_t i0 = 0, i1 = 0;
        const bool hasX = (r0 == MB_XRES) | (r1 == MB_XRES) | (r2 == MB_XRES) | (r3 == MB_XRES) | (r4 == MB_XRES) | (r5 == MB_XRES);
        if (!hasX) {
            const short bs = (short)((b < 0.0) ? (double)b - 0.5 : (double)b + 0.5);
            thr = T.kmerThr - bs; if (thr < 0) thr = 0;
            i0 = r0 + 20 * r1 + 400 * r2; i1 = r3 + 20 * r4 + 400 * r5;
            const int top1 = (int16_t)(T.sim3[(size_t)i1 * MB_N3] & 0xFFFFu);
            const short cutoff1 = (short)(thr - top1);
            n0 = mb_count_ge(T.cnt3, i0, cutoff1);
        }
        itemOff[p] = n0; wThr[p] = (int16_t)thr; wI0[p] = (uint16_t)i0; wI1[p] = (uint16_t)i1;
    }
    CTA_SYNC();
    return cta_exclusive_scan(itemOff, P, &sh->cs);
}

// ---- P2 + P3; returns K (similar k-mers); the non-empty ones are listed in sh->sNz / W.B (count in sh->cnt[7]).  K > capK: nothing listed.
template <bool SMP, bool SMI, class SH>
MB_DEV uint32_t mb_pref_expand(const MbPrefTables &T, const MbPrefScratch &W, SH *sh, uint32_t P, uint32_t I) {
    uint32_t *const itemOff = SMP ? sh->sItemOff : W.itemOff; int16_t *const wThr = SMP ? sh->sThr : W.wThr;
    uint16_t *const wI0 = SMP ? sh->sI0 : W.wI0; uint16_t *const wI1 = SMP ? sh->sI1 : W.wI1;
    uint32_t *const kmerOff = SMI ? sh->sKmerOff : W.kmerOff;
    uint16_t *const itemIdx0 = SMI ? sh->sItemIdx0 : (uint16_t *)W.itemInfo;
    // ---- P2 items (p, i): second-level neighbour count  (calculateArrayProduct, KmerGenerator.cpp:186-216).  Each thread owns a
    //      contiguous run of items (one search for the first window, then a cursor); 4 per round so the two dependent table loads overlap.
    {
        CTA_CHUNK(ib, ie, I);
        uint32_t pw = (ib < ie) ? cta_find_segment(itemOff, P, ib) : 0;
        for (uint32_t i0 = ib; i0 < ie; i0 += 4) {
            uint32_t pv[4], a0[4]; bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const uint32_t it = i0 + u; ok[u] = it < ie; pv[u] = 0; if (ok[u]) { while (itemOff[pw + 1] <= it) pw++; pv[u] = pw; } }
#pragma unroll
            for (int u = 0; u < 4; u++) a0[u] = ok[u] ? T.sim3[(uint32_t)wI0[pv[u]] * MB_N3 + (i0 + u - itemOff[pv[u]])] : 0;
#pragma unroll
            for (int u = 0; u < 4; u++) if (ok[u]) {
                const short cutoff2 = (short)((int)wThr[pv[u]] - (int)(int16_t)(a0[u] & 0xFFFFu));
                kmerOff[i0 + u] = mb_count_ge(T.cnt3, wI1[pv[u]], cutoff2);
                itemIdx0[i0 + u] = (uint16_t)(a0[u] >> 16);
            }
        }
    }
    CTA_SYNC();
    const uint32_t K = cta_exclusive_scan(kmerOff, I, &sh->cs);
    MB_PHASE(2);
    if (K > T.capK) return K;

    // ---- P3 similar k-mers in the reference's enumeration order k = 0..K-1 (window, first-half rank, second-half rank); those whose
    //      index list is non-empty (bitmap test; IndexTable::getDBSeqList, IndexTable.h:176-180) go to an unordered list of
    //      (window | k | k-mer) words.  Each thread owns a contiguous run of k and keeps the (item, window) cursor in registers; the run is
    //      handled in batches of 4 with uniform control flow, so the dependent loads of a batch (sim3 row -> bitmap) are in flight
    //      together and the list append is one shared-memory atomic per warp and batch slot.
    {
        CTA_CHUNK(kb, ke, K);
        uint32_t it = 0, pw = 0, o = 0, oe = 0, idx0 = 0, rowBase = 0;
        if (kb < ke) {
            it = cta_find_segment(kmerOff, I, kb); pw = cta_find_segment(itemOff, P, it);
            o = kmerOff[it]; oe = kmerOff[it + 1]; idx0 = itemIdx0[it]; rowBase = (uint32_t)wI1[pw] * MB_N3;
        }
        for (uint32_t k0 = kb; k0 < ke; k0 += 4) {
            uint32_t ad[4], i0v[4], pv[4], a1[4], km[4]; bool ok[4], nz[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t k = k0 + u; ok[u] = k < ke; ad[u] = 0; i0v[u] = 0; pv[u] = 0;
                if (ok[u]) {
                    while (k >= oe) {
                        it++; o = oe; oe = kmerOff[it + 1]; idx0 = itemIdx0[it];
                        while (itemOff[pw + 1] <= it) pw++;
                        rowBase = (uint32_t)wI1[pw] * MB_N3;
                    }
                    ad[u] = rowBase + (k - o); i0v[u] = idx0; pv[u] = pw;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) a1[u] = ok[u] ? T.sim3[ad[u]] : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                km[u] = i0v[u] + (a1[u] >> 16) * 8000u; nz[u] = ok[u];
                if (ok[u] && T.kmerBits) nz[u] = (T.kmerBits[km[u] >> 5] >> (km[u] & 31u)) & 1u;
            }
            // sparse index: only the k-mers whose list is non-empty (bitmap) are listed, in no particular order.  Dense index (no bitmap: nearly every
            // k-mer has a list): every k-mer is listed at its own ordinal and the gather phase finds the few empty lists itself.
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t n = T.kmerBits ? cta_append_slot(&sh->cnt[7], nz[u]) : k0 + u;
                if (nz[u]) {
                    const uint64_t w = ((uint64_t)pv[u] << 47) | ((uint64_t)(k0 + u) << 26) | km[u];   // window[17] | k[21] | k-mer[26]
                    if (n < SH::Cfg::SM_NZ) sh->sNz[n] = w;
                    else if (n - SH::Cfg::SM_NZ < T.capM) W.B[n - SH::Cfg::SM_NZ] = w;
                    else sh->cnt[3] = 1u;                                                               // list capacity exhausted (checked by the caller)
                }
            }
        }
    }
    return K;
}

// One query.  q/bias/corr point at this query's residues, float composition bias (seed matrix) and int8 profile correction.
template <class SH>
MB_DEV void mb_prefilter_query(const MbPrefTables &T, const MbPrefScratch &W, SH *sh, const uint8_t *q, const float *bias,
                               const int8_t *corr, uint32_t L, MbPrefHit *out, uint32_t *outCount, uint32_t *outStatus, MbPrefStats *stats) {
    const uint32_t P = (L >= 10) ? (L - 9) : 0;   // Sequence::hasNextKmer, spaced pattern length 10 (Sequence.h:23,92-94)
    MB_PHASE(15);   // time between queries (queue pop, previous epilogue)
    CTA_SINGLE { for (int i = 0; i < 8; i++) sh->cnt[i] = 0; }
    CTA_FOR(i, 256) sh->hist[i] = 0;
    CTA_FOR(i, SH::Cfg::NSLOT / 16) sh->slots[i] = 0;
    CTA_SYNC();
    MB_PHASE(0);    // clearing
    if (L >= 32768 || P > T.capP) { CTA_SINGLE { *outCount = 0; *outStatus = MB_ST_TOO_LONG; } return; }
    if (P == 0) { CTA_SINGLE { *outCount = 0; *outStatus = MB_ST_OK; } return; }

    const bool smP = P <= MB_SM_P;
    const uint32_t I = smP ? mb_pref_windows<true>(T, W, sh, q, bias, P) : mb_pref_windows<false>(T, W, sh, q, bias, P);
    MB_PHASE(1);
    if (I > T.capI) { CTA_SINGLE { *outCount = 0; *outStatus = MB_ST_OVERFLOW; } return; }
    const uint32_t K = (smP && I <= SH::Cfg::SM_I) ? mb_pref_expand<true, true>(T, W, sh, P, I)
                     : smP ? mb_pref_expand<true, false>(T, W, sh, P, I) : mb_pref_expand<false, false>(T, W, sh, P, I);
    if (K > T.capK) { CTA_SINGLE { *outCount = 0; *outStatus = MB_ST_OVERFLOW; } return; }
    CTA_SYNC();
    const uint32_t nNz = T.kmerBits ? sh->cnt[7] : K;
    MB_PHASE(3);
    if (sh->cnt[3]) { CTA_SINGLE { *outCount = 0; *outStatus = MB_ST_OVERFLOW; } return; }   // more non-empty k-mers than the list holds

    // ---- P4 gather the index entries of the listed k-mers; mark slot occupancy (0 / 1 / >=2 hits).
    //      A warp takes 128 lists at a time: their (first entry, length) come from the offsets table with four loads in flight per lane, a warp scan lays
    //      the lists end to end, and the entries of that flattened range are then fetched one per lane and step -- four steps (128 independent loads per
    //      warp) in flight before the first is consumed -- whatever list they belong to (owner found by binary search over the 128 prefix sums, staged
    //      in the warp's slice of the k-mer offset array, which is dead after P3).  Output slots are reserved with one shared-memory atomic per warp and round.
#if MB_DEVICE_BUILD
    {
        const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
        uint32_t *const wPre = sh->sKmerOff + warp * 256u, *const wLo = wPre + 128u;
        for (uint32_t n0 = warp * 128u; n0 < nNz; n0 += nWarps * 128u) {
            uint32_t lo[4], len[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { lo[u] = 0; len[u] = 0; }
            if (T.buckets) {
                // inline buckets: the entries of a short list arrive with the lookup itself (two 128-bit loads per k-mer, two k-mers per lane in flight);
                // only the lists of more than MB_BUCKET_INLINE entries go through the flattened gather below
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    uint4 ba[2], bb[2]; uint32_t nIn[2], inl = 0;
#pragma unroll
                    for (int v = 0; v < 2; v++) {
                        const uint32_t n = n0 + (uint32_t)(2 * hh + v) * 32u + lane; nIn[v] = 0;
                        if (n < nNz) {
                            const uint32_t km = (uint32_t)((n < SH::Cfg::SM_NZ) ? sh->sNz[n] : W.B[n - SH::Cfg::SM_NZ]) & 0x3FFFFFFu;
                            mb_load_bucket(T.buckets + 2 * (size_t)km, ba[v], bb[v], T.streamHint);
                        }
                    }
#pragma unroll
                    for (int v = 0; v < 2; v++) if (n0 + (uint32_t)(2 * hh + v) * 32u + lane < nNz) {
                        const uint32_t cnt = mb_bucket_h(ba[v], bb[v], 0);
                        if (cnt <= MB_BUCKET_INLINE) { nIn[v] = cnt; inl += cnt; }
                        else { lo[2 * hh + v] = ba[v].y; len[2 * hh + v] = ba[v].z; }     // (halfwords 2..3 and 4..5)
                    }
                    uint32_t inc = inl;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += y; }
                    const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
                    uint32_t ob = 0, ofits = 1;
                    if (lane == 0 && tot) {
                        if (*(volatile uint32_t *)&sh->cnt[3] == 0u) {
                            ob = atomicAdd(&sh->cnt[6], tot);
                            ofits = (ob <= T.capM && tot <= T.capM - ob) ? 1u : 0u;
                            if (!ofits) sh->cnt[3] = 1u;
                        } else ofits = 0;
                    }
                    ob = __shfl_sync(0xffffffffu, ob, 0); ofits = __shfl_sync(0xffffffffu, ofits, 0);
                    if (ofits && tot) {
                        uint32_t o = ob + inc - inl;
#pragma unroll
                        for (int v = 0; v < 2; v++) if (nIn[v]) {
                            const uint32_t n = n0 + (uint32_t)(2 * hh + v) * 32u + lane;
                            const uint64_t wv = (n < SH::Cfg::SM_NZ) ? sh->sNz[n] : W.B[n - SH::Cfg::SM_NZ];
                            const uint32_t pp = (uint32_t)(wv >> 47), kk = (uint32_t)(wv >> 26) & 0x1FFFFFu;
#pragma unroll
                            for (int j = 0; j < MB_BUCKET_INLINE; j++) if ((uint32_t)j < nIn[v]) {
                                uint32_t seqId, pos; mb_bucket_entry(ba[v], bb[v], j, seqId, pos);
                                const uint16_t diag16 = (uint16_t)(pp - pos);                  // position_i - position_j (CacheFriendlyOperations.cpp:342)
                                W.A[o++] = ((uint64_t)seqId << 37) | ((uint64_t)kk << 16) | diag16;
                                const uint32_t slot = seqId & (SH::Cfg::NSLOT - 1), sft = (slot & 15u) * 2u;
                                const uint32_t old = mb_atomic_or(&sh->slots[slot >> 4], 1u << sft);
                                if ((old >> sft) & 1u) mb_atomic_or(&sh->slots[slot >> 4], 2u << sft);
                            }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t n = n0 + (uint32_t)u * 32u + lane;
                    if (n < nNz) {
                        const uint32_t km = (uint32_t)((n < SH::Cfg::SM_NZ) ? sh->sNz[n] : W.B[n - SH::Cfg::SM_NZ]) & 0x3FFFFFFu;
                        lo[u] = T.offsets[km]; len[u] = T.offsets[km + 1] - lo[u];
                    }
                }
            }
            uint32_t run = 0; bool big = false;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                uint32_t l = len[u]; if (l > T.capM) { big = true; l = 0; }
                uint32_t inc = l;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += y; }
                wPre[u * 32 + lane] = run + inc - l; wLo[u * 32 + lane] = lo[u];
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            const uint32_t Tn = run;
            big = __any_sync(0xffffffffu, big);
            uint32_t b0 = 0, fits = 1;
            if (lane == 0) {
                if (big) { sh->cnt[3] = 1u; fits = 0; }
                else if (Tn) {
                    if (*(volatile uint32_t *)&sh->cnt[3] == 0u) {   // cnt[3] = overflow flag until P11 reuses it
                        b0 = atomicAdd(&sh->cnt[6], Tn);
                        fits = (b0 <= T.capM && Tn <= T.capM - b0) ? 1u : 0u;
                        if (!fits) sh->cnt[3] = 1u;
                    } else fits = 0;
                }
            }
            b0 = __shfl_sync(0xffffffffu, b0, 0); fits = __shfl_sync(0xffffffffu, fits, 0);
            __syncwarp();
            if (fits) for (uint32_t e0 = 0; e0 < Tn; e0 += 128u) {
                uint32_t own[4]; uint2 en[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t e = e0 + (uint32_t)j * 32u + lane; own[j] = 0xFFFFFFFFu; en[j].x = 0; en[j].y = 0;
                    if (e < Tn) {
                        uint32_t a = 0, b = 128;   // wPre[a] <= e < wPre[b]  (wPre[128] = Tn)
#pragma unroll
                        for (int it = 0; it < 7; it++) { const uint32_t mid = (a + b) >> 1; if (wPre[mid] <= e) a = mid; else b = mid; }
                        own[j] = a; en[j] = mb_load_entry(T.entries + (size_t)wLo[a] + (e - wPre[a]), T.streamHint);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) if (own[j] != 0xFFFFFFFFu) {
                    const uint32_t e = e0 + (uint32_t)j * 32u + lane, n = n0 + own[j];
                    const uint64_t wv = (n < SH::Cfg::SM_NZ) ? sh->sNz[n] : W.B[n - SH::Cfg::SM_NZ];
                    const uint32_t pp = (uint32_t)(wv >> 47), kk = (uint32_t)(wv >> 26) & 0x1FFFFFu;
                    const uint16_t diag16 = (uint16_t)(pp - (en[j].y & 0xFFFFu));   // position_i - position_j (CacheFriendlyOperations.cpp:342)
                    W.A[b0 + e] = ((uint64_t)en[j].x << 37) | ((uint64_t)kk << 16) | diag16;
                    const uint32_t slot = en[j].x & (SH::Cfg::NSLOT - 1), sft = (slot & 15u) * 2u;
                    const uint32_t old = mb_atomic_or(&sh->slots[slot >> 4], 1u << sft);
                    if ((old >> sft) & 1u) mb_atomic_or(&sh->slots[slot >> 4], 2u << sft);
                }
            }
            __syncwarp();
        }
    }
#else
    for (uint32_t n = 0; n < nNz; n++) {   // (serial restatement for tests/emu)
        const uint64_t wv = (n < SH::Cfg::SM_NZ) ? sh->sNz[n] : W.B[n - SH::Cfg::SM_NZ];
        const uint32_t km = (uint32_t)wv & 0x3FFFFFFu, lo = T.offsets[km], len = T.offsets[km + 1] - lo;
        if (!len) continue;
        if (sh->cnt[3] != 0u) break;
        const uint32_t b0 = sh->cnt[6];
        if (len > T.capM || b0 > T.capM || len > T.capM - b0) { sh->cnt[3] = 1u; break; }
        sh->cnt[6] += len;
        const uint32_t pp = (uint32_t)(wv >> 47), kk = (uint32_t)(wv >> 26) & 0x1FFFFFu;
        for (uint32_t x = 0; x < len; x++) {
            const uint2 en = T.entries[(size_t)lo + x];
            const uint16_t diag16 = (uint16_t)(pp - (en.y & 0xFFFFu));   // position_i - position_j (CacheFriendlyOperations.cpp:342)
            W.A[b0 + x] = ((uint64_t)en.x << 37) | ((uint64_t)kk << 16) | diag16;
            const uint32_t slot = en.x & (SH::Cfg::NSLOT - 1), sft = (slot & 15u) * 2u;
            const uint32_t old = mb_atomic_or(&sh->slots[slot >> 4], 1u << sft);
            if ((old >> sft) & 1u) mb_atomic_or(&sh->slots[slot >> 4], 2u << sft);
        }
    }
#endif
    CTA_SYNC();
    const uint32_t M = sh->cnt[6];
    MB_PHASE(4);
    if (sh->cnt[3]) { CTA_SINGLE { *outCount = 0; *outStatus = MB_ST_OVERFLOW; } return; }
    // ---- P5 split: targets seen once -> candidate iff 8-bit diagonal == 0 (initial state of the reference's byte array);
    //               targets possibly seen more than once -> multi list
    for (uint32_t h0 = 0; h0 < M; h0 += 4 * CTA_NT) {       // four hit words per thread in flight (the list lives in L2)
        uint64_t wv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const uint32_t h = h0 + u * CTA_NT + CTA_TID; wv[u] = (h < M) ? W.A[h] : 0; }
#pragma unroll
        for (int u = 0; u < 4; u++) if (h0 + u * CTA_NT + CTA_TID < M) {
            const uint64_t w = wv[u];
            const uint32_t slot = MB_W_SEQ(w) & (SH::Cfg::NSLOT - 1), sft = (slot & 15u) * 2u;
            const bool multi = (sh->slots[slot >> 4] >> sft) & 2u;
            const uint32_t ib = cta_append_slot(&sh->cnt[0], multi);
            if (multi) { W.B[ib] = w; if (ib < SH::Cfg::MUL_CAP) sh->sMul[ib] = w; }
            const bool single0 = !multi && (w & 0xFFu) == 0;
            const uint32_t ic = cta_append_slot(&sh->cnt[1], single0);
            if (single0) W.C[ic] = w;
        }
    }
    CTA_SYNC();
    uint32_t nMulti = sh->cnt[0];
    // A long query has so many hits that unrelated targets share occupancy slots (M^2 / (2 NSLOT) colliding pairs: thousands for M = 30 000): filter the "possibly seen twice" list
    // again with another hash of the target id -- the hits of a target seen once that collide neither time leave as singles -- until it fits the shared-memory stage below.
    for (uint32_t round = 1; round <= 2u && nMulti > SH::Cfg::MUL_CAP; round++) {
        const uint32_t mul = round == 1u ? 0x9E3779B1u : 0x85EBCA6Bu;
        CTA_SYNC();      // (every thread has read the list length before it is reset)
        CTA_FOR(i, SH::Cfg::NSLOT / 16) sh->slots[i] = 0;
        CTA_SINGLE { sh->cnt[0] = 0; }
        CTA_SYNC();
        CTA_FOR(e, nMulti) {
            const uint32_t slot = (MB_W_SEQ(W.B[e]) * mul) >> (32u - SH::Cfg::NSLOT_LOG), sft = (slot & 15u) * 2u;
            const uint32_t old = mb_atomic_or(&sh->slots[slot >> 4], 1u << sft);
            if ((old >> sft) & 1u) mb_atomic_or(&sh->slots[slot >> 4], 2u << sft);
        }
        CTA_SYNC();
        for (uint32_t h0 = 0; h0 < nMulti; h0 += 4 * CTA_NT) {
            uint64_t wv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const uint32_t h = h0 + u * CTA_NT + CTA_TID; wv[u] = (h < nMulti) ? W.B[h] : 0; }
#pragma unroll
            for (int u = 0; u < 4; u++) if (h0 + u * CTA_NT + CTA_TID < nMulti) {
                const uint64_t w = wv[u];
                const uint32_t slot = (MB_W_SEQ(w) * mul) >> (32u - SH::Cfg::NSLOT_LOG), sft = (slot & 15u) * 2u;
                const bool multi = (sh->slots[slot >> 4] >> sft) & 2u;
                const uint32_t ib = cta_append_slot(&sh->cnt[0], multi);
                if (multi) W.A[ib] = w;
                const bool single0 = !multi && (w & 0xFFu) == 0;
                const uint32_t ic = cta_append_slot(&sh->cnt[1], single0);
                if (single0) W.C[ic] = w;
            }
        }
        CTA_SYNC();
        nMulti = sh->cnt[0];
        CTA_FOR(e, nMulti) { const uint64_t w = W.A[e]; W.B[e] = w; if (e < SH::Cfg::MUL_CAP) sh->sMul[e] = w; }
        CTA_SYNC();
    }
    MB_PHASE(5);
    // ---- P6/P7 per-target replay of findDuplicates in arrival order  (CacheFriendlyOperations.cpp:193-262)
    if (nMulti == 0) {
    } else if (nMulti > T.sortWalkMax && nMulti <= SH::Cfg::MUL_CAP) {
        // The usual case: the whole stage in shared memory (list, group-ordered copy, hash table), same three passes as the global-memory variant below.
        uint32_t htLog = SH::Cfg::HT_LOG_MAX - 1u < 6u ? SH::Cfg::HT_LOG_MAX - 1u : 6u; while ((1u << htLog) < 2u * nMulti) htLog++;
        const uint32_t htN = 1u << htLog;                    // <= SM_NZ / 2
        uint32_t *htKey = sh->sHt, *htOff = sh->sHt + htN;
        CTA_FOR(i, 2 * htN) sh->sHt[i] = 0;
        CTA_SYNC();
        CTA_FOR(e, nMulti) {
            const uint32_t id1 = MB_W_SEQ(sh->sMul[e]) + 1u;
            uint32_t slot = (id1 * 2654435761u) >> (32 - htLog);
            for (;;) {
                const uint32_t old = mb_atomic_cas(&htKey[slot], 0u, id1);
                if (old == 0u || old == id1) break;
                slot = (slot + 1u) & (htN - 1u);
            }
            mb_atomic_add(&htOff[slot], 1u);
        }
        CTA_SYNC();
        cta_exclusive_scan(htOff, htN - 1, &sh->cs);
        CTA_FOR(e, nMulti) { const uint64_t w = sh->sMul[e]; const uint32_t slot = mb_ht_find(htKey, htLog, MB_W_SEQ(w) + 1u); sh->sGrp[mb_atomic_add(&htOff[slot], 1u)] = w; }   // afterwards htOff[s] = END of group s
        CTA_SYNC();
        CTA_FOR(e, nMulti) {      // pass X (see below); the flagged words go back into sMul, in group order
            const uint64_t w = sh->sGrp[e];
            const uint32_t slot = mb_ht_find(htKey, htLog, MB_W_SEQ(w) + 1u);
            const uint32_t gb = slot ? htOff[slot - 1] : 0u, ge = htOff[slot];
            uint64_t pred = 0; bool has = false;
            for (uint32_t x = gb; x < ge; x++) { const uint64_t y = sh->sGrp[x]; if (y < w && (!has || y > pred)) { pred = y; has = true; } }
            const uint32_t d8 = (uint32_t)(w & 0xFFu), last = has ? (uint32_t)(pred & 0xFFu) : 0u;
            sh->sMul[e] = w | ((d8 == last) ? (1ull << 63) : 0ull);
        }
        CTA_SYNC();
        CTA_FOR(e, nMulti) {      // pass Y
            const uint64_t wf = sh->sMul[e];
            if (wf >> 63) {
                const uint64_t w = wf & ~(1ull << 63);
                const uint32_t slot = mb_ht_find(htKey, htLog, MB_W_SEQ(w) + 1u);
                const uint32_t gb = slot ? htOff[slot - 1] : 0u, ge = htOff[slot];
                uint64_t pc = 0; bool has = false;
                for (uint32_t x = gb; x < ge; x++) { const uint64_t yf = sh->sMul[x]; const uint64_t y = yf & ~(1ull << 63); if ((yf >> 63) && y < w && (!has || y > pc)) { pc = y; has = true; } }
                const bool accept = !has || ((uint32_t)(pc & 0xFFu) != (uint32_t)(w & 0xFFu));
                const uint32_t ic = cta_append_slot(&sh->cnt[1], accept);
                if (accept) W.C[ic] = w;
            } else { cta_append_slot(&sh->cnt[1], false); }
        }
    } else if (nMulti > T.sortWalkMax && nMulti <= (1u << (SH::Cfg::HT_LOG_MAX - 1))) {   // (up to 256 entries one warp sorts in registers: the sorted walk below is cheaper)
        // Sort-free: group the multi list by target with a shared-memory hash table (the k-mer list's space is free now), then every
        // hit looks up its predecessor *of the same target* (largest smaller arrival key) inside its group.
        uint32_t htLog = SH::Cfg::HT_LOG_MAX < 6u ? SH::Cfg::HT_LOG_MAX : 6u; while ((1u << htLog) < 2u * nMulti) htLog++;
        const uint32_t htN = 1u << htLog;
        uint32_t *htKey = sh->sHitOff, *htOff = sh->sHitOff + htN;
        CTA_FOR(i, 2 * htN) sh->sHitOff[i] = 0;
        CTA_SYNC();
        CTA_FOR(e, nMulti) {
            const uint32_t id1 = MB_W_SEQ(W.B[e]) + 1u;
            uint32_t slot = (id1 * 2654435761u) >> (32 - htLog);
            for (;;) {
                const uint32_t old = mb_atomic_cas(&htKey[slot], 0u, id1);
                if (old == 0u || old == id1) break;
                slot = (slot + 1u) & (htN - 1u);
            }
            mb_atomic_add(&htOff[slot], 1u);
            W.S[e] = (int32_t)slot;
        }
        CTA_SYNC();
        cta_exclusive_scan(htOff, htN - 1, &sh->cs);   // (the last slot's count is covered by the total; offsets only need [0, N-1))
        CTA_FOR(e, nMulti) { const uint32_t slot = (uint32_t)W.S[e]; W.A[mb_atomic_add(&htOff[slot], 1u)] = W.B[e]; }   // afterwards htOff[s] = END of group s
        CTA_SYNC();
        // pass X: is this hit a duplicate of the immediately preceding hit of its target?  (byte array starts at 0: CacheFriendlyOperations.cpp:213-218)
        CTA_FOR(e, nMulti) {
            const uint64_t w = W.A[e];
            const uint32_t slot = mb_ht_find(htKey, htLog, MB_W_SEQ(w) + 1u);
            const uint32_t gb = slot ? htOff[slot - 1] : 0u, ge = htOff[slot];
            uint64_t pred = 0; bool has = false;
            for (uint32_t x = gb; x < ge; x++) { const uint64_t y = W.A[x]; if (y < w && (!has || y > pred)) { pred = y; has = true; } }
            const uint32_t d8 = (uint32_t)(w & 0xFFu), last = has ? (uint32_t)(pred & 0xFFu) : 0u;
            W.B[e] = w | ((d8 == last) ? (1ull << 63) : 0ull);   // bit 63 = candidate flag (seqId uses 27 bits)
        }
        CTA_SYNC();
        // pass Y: runs of candidates with equal 8-bit diagonal collapse to their first (phases 2/3, CacheFriendlyOperations.cpp:236-262)
        CTA_FOR(e, nMulti) {
            const uint64_t wf = W.B[e];
            if (wf >> 63) {
                const uint64_t w = wf & ~(1ull << 63);
                const uint32_t slot = mb_ht_find(htKey, htLog, MB_W_SEQ(w) + 1u);
                const uint32_t gb = slot ? htOff[slot - 1] : 0u, ge = htOff[slot];
                uint64_t pc = 0; bool has = false;
                for (uint32_t x = gb; x < ge; x++) { const uint64_t yf = W.B[x]; const uint64_t y = yf & ~(1ull << 63); if ((yf >> 63) && y < w && (!has || y > pc)) { pc = y; has = true; } }
                const bool accept = !has || ((uint32_t)(pc & 0xFFu) != (uint32_t)(w & 0xFFu));
                const uint32_t ic = cta_append_slot(&sh->cnt[1], accept);
                if (accept) W.C[ic] = w;
            } else { cta_append_slot(&sh->cnt[1], false); }
        }
    } else {
    mb_sort_list(W.B, nMulti, sh);
    CTA_FOR(e, nMulti) {
        const uint32_t id = MB_W_SEQ(W.B[e]);
        if (e == 0 || MB_W_SEQ(W.B[e - 1]) != id) {
            uint32_t last = 0; int prevCand = -1;
            for (uint32_t x = e; x < nMulti && MB_W_SEQ(W.B[x]) == id; x++) {
                const uint64_t w = W.B[x]; const uint32_t d8 = (uint32_t)(w & 0xFFu);
                if (d8 == last) {                         // duplicate of the immediately preceding hit of this target
                    if (prevCand != (int)d8) W.C[mb_atomic_add(&sh->cnt[1], 1u)] = w;   // runs of equal diagonals collapse to their first
                    prevCand = (int)d8;
                }
                last = d8;
            }
        }
    }
    }
    CTA_SYNC();
    const uint32_t D = sh->cnt[1];
    MB_PHASE(6);
    if (D > (1u << 21)) { CTA_SINGLE { *outCount = 0; *outStatus = MB_ST_OVERFLOW; } return; }   // candidate indices are 21-bit fields of the sort keys below
    // ---- P8/P9 ungapped diagonal scores
    mb_sort_list(W.C, D, sh);
    uint32_t myCells = 0;
    CTA_WARP_FOR(c, D) {   // one warp per candidate diagonal
        const uint64_t w = W.C[c]; const uint32_t id = MB_W_SEQ(w);
        const uint64_t to = T.toff[id]; const uint32_t tLen = (uint32_t)(T.toff[id + 1] - to);
        uint32_t cells;
#if MB_DEVICE_BUILD
        const int sc = mb_diag_score_warp(sh->ungMat, q, corr, L, T.tseq + to, tLen, MB_W_DIAG(w), &cells);
        if ((threadIdx.x & 31u) == 0) { W.S[c] = sc; myCells += cells; }
#else
        W.S[c] = mb_diag_score(sh->ungMat, q, corr, L, T.tseq + to, tLen, MB_W_DIAG(w), &cells);
        myCells += cells;
#endif
    }
    CTA_SYNC();
    // ---- P10 best diagonal per target: first maximum in arrival order (keepMaxElement, CacheFriendlyOperations.cpp:350-380)
    MB_PHASE(7);
    CTA_FOR(c, D) {
        const uint32_t id = MB_W_SEQ(W.C[c]);
        if (c == 0 || MB_W_SEQ(W.C[c - 1]) != id) {
            uint32_t best = 0, bestIdx = c;
            for (uint32_t x = c; x < D && MB_W_SEQ(W.C[x]) == id; x++) {
                const int s = W.S[x]; const uint32_t cnt = s > 255 ? 255u : (uint32_t)s;
                if (cnt > best) { best = cnt; bestIdx = x; }
            }
            if (best > 0) {   // count==0 entries never reach the output (threshold >= 1 is enforced by the host API)
                const uint64_t bw = W.C[bestIdx];
                W.B[mb_atomic_add(&sh->cnt[2], 1u)] = ((uint64_t)(255u - best) << 53) | ((uint64_t)(id & T.binMask) << 42) | ((uint64_t)MB_W_ARR(bw) << 21) | bestIdx;
                mb_atomic_add(&sh->hist[best], 1u);
            }
        }
    }
    CTA_SYNC();
    const uint32_t nSurv = sh->cnt[2];
    MB_PHASE(8);
    // ---- P11 score threshold (computeScoreThreshold, QueryMatcher.h:206-216; QueryMatcher.cpp:153-155)
#if MB_DEVICE_BUILD
    if (threadIdx.x < 32) {   // one warp: lane l owns score bins [8l, 8l+8); S(t) = number of targets with 8-bit score >= t
        const uint32_t maxHits = (uint32_t)T.maxHits < T.nT ? (uint32_t)T.maxHits : T.nT;
        const uint32_t lane = threadIdx.x;
        uint32_t hb[8], mine = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) { hb[b] = sh->hist[lane * 8 + b]; mine += hb[b]; }
        uint32_t incl = mine;   // sum over lanes >= lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_down_sync(0xffffffffu, incl, d); if (lane + d < 32) incl += o; }
        uint32_t run = incl - mine; int cand = 0;   // largest t in this lane's bins (t >= 1) with S(t) >= maxHits
#pragma unroll
        for (int b = 7; b >= 0; b--) { run += hb[b]; const int t = (int)lane * 8 + b; if (cand == 0 && t >= 1 && run >= maxHits) cand = t; }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, cand, d); cand = o > cand ? o : cand; }
        uint32_t thr = (uint32_t)cand;
        if (thr < (uint32_t)T.minDiag) thr = (uint32_t)T.minDiag;
        uint32_t part = 0;   // S(thr)
#pragma unroll
        for (int b = 0; b < 8; b++) if (lane * 8 + b >= thr) part += hb[b];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        if (lane == 0) { sh->cnt[3] = thr; sh->cnt[4] = part; }
    }
#else
    CTA_SINGLE {
        const uint32_t maxHits = (uint32_t)T.maxHits < T.nT ? (uint32_t)T.maxHits : T.nT;
        uint32_t found = 0, thr = 0;
        for (thr = 255; thr > 0; thr--) { found += sh->hist[thr]; if (found >= maxHits) break; }
        if (thr < (uint32_t)T.minDiag) thr = (uint32_t)T.minDiag;
        uint32_t nPass = 0;
        for (uint32_t s = thr; s < 256; s++) nPass += sh->hist[s];
        sh->cnt[3] = thr; sh->cnt[4] = nPass;
    }
#endif
    // ---- P12 stable order of the reference's counting sort: score desc, then bin (seqId & (BINSIZE-1)), then arrival
    mb_sort_list(W.B, nSurv, sh);
    const uint32_t thr = sh->cnt[3];
    uint32_t nPass = sh->cnt[4];
    const uint32_t maxHits = (uint32_t)T.maxHits < T.nT ? (uint32_t)T.maxHits : T.nT;
    const bool truncated = thr >= 255;
    if (truncated) {
        // ---- rescoreHits (QueryMatcher.cpp:160-170,525-544): >= maxHits targets saturate the 8-bit score
        CTA_SINGLE {
            uint32_t cells;
            int self = mb_diag_score(sh->ungMat, q, corr, L, q, L, 0, &cells);
            self -= 255; if (self < 1) self = 1; if (self > 65535) self = 65535;
            sh->cnt[5] = (uint32_t)self;
        }
        CTA_SYNC();
        const float fltMax = (float)(int)sh->cnt[5];
        CTA_FOR(i, nPass) {
            const uint32_t ci = (uint32_t)(W.B[i] & 0x1FFFFFu);
            uint32_t ns = (uint32_t)W.S[ci] - 255u; if (ns > 65535u) ns = 65535u;
            const float sc = (float)ns;
            const uint32_t nc = (uint32_t)(unsigned char)((double)mb_fmul(mb_fdiv(sc, fltMax), 255.0f) + 0.5);
            W.A[i] = ((uint64_t)(255u - nc) << 53) | ((uint64_t)i << 21) | ci;
        }
        CTA_SYNC();
        CTA_FOR(i, nPass) W.B[i] = W.A[i];
        CTA_SYNC();
        mb_sort_list(W.B, nPass, sh);
    }
    const uint32_t nSel = nPass < maxHits ? nPass : maxHits;
    // ---- P13 getResult<UNGAPPED_DIAGONAL_SCORE> (QueryMatcher.cpp:363-420): exact score when the 8-bit score saturated
    MB_PHASE(9);
    CTA_FOR(i, nSel) {
        const uint64_t w = W.B[i]; const uint32_t ci = (uint32_t)(w & 0x1FFFFFu);
        const uint32_t cnt = 255u - (uint32_t)(w >> 53);
        uint32_t score;
        if (truncated) score = 255u + (cnt * sh->cnt[5] / 255u);
        else score = (cnt >= 255u) ? (uint32_t)W.S[ci] : cnt;
        const uint64_t cw = W.C[ci];
        W.A[i] = ((uint64_t)(0xFFFFFu - (score & 0xFFFFFu)) << 43) | ((uint64_t)MB_W_SEQ(cw) << 16) | MB_W_DIAG(cw);
    }
    CTA_SYNC();
    // ---- P14 final order: score desc, seqId asc (hit_t::compareHitsByScoreAndId, QueryMatcher.h:38-48)
    mb_sort_list(W.A, nSel, sh);
    CTA_FOR(i, nSel) {
        const uint64_t w = W.A[i];
        MbPrefHit hres; hres.seqId = (uint32_t)((w >> 16) & 0x7FFFFFFu); hres.score = (int32_t)(0xFFFFFu - (uint32_t)(w >> 43)); hres.diag = (int16_t)(w & 0xFFFFu);
        out[i] = hres;
    }
    // stats (algorithmic-byte numerators for the roofline)
#if MB_DEVICE_BUILD
    if (stats) {
        if (myCells) atomicAdd(&stats->diagCells, (unsigned long long)myCells);
        if (threadIdx.x == 0) { atomicAdd(&stats->kmers, (unsigned long long)K); atomicAdd(&stats->lookups, (unsigned long long)K);
            atomicAdd(&stats->dbMatches, (unsigned long long)M); atomicAdd(&stats->doubleHits, (unsigned long long)D); atomicAdd(&stats->results, (unsigned long long)nSel); }
    }
#else
    if (stats) { stats->diagCells += myCells; stats->kmers += K; stats->lookups += K; stats->dbMatches += M; stats->doubleHits += D; stats->results += nSel; }
#endif
    CTA_SINGLE { *outCount = nSel; *outStatus = MB_ST_OK; }
    CTA_SYNC();
    MB_PHASE(10);   // P13/P14 + output
}

// ---------------------------------------------------------------------------------------------------
// Per-residue preparation: float composition bias (SubstitutionMatrix::calcLocalAaBiasCorrection,
// reference/lib/mmseqs/src/commons/SubstitutionMatrix.cpp:79-109) for the seed matrix and the alignment matrix,
// the int8 ungapped-profile correction (UngappedAlignment.cpp:392-397) and the int8 SW composition bias
// (StripedSmithWaterman.cpp:1229-1237).  Exact IEEE op order, no FMA contraction.
// ---------------------------------------------------------------------------------------------------
MB_DEV float mb_comp_bias_at(const int16_t *mat /*[441]*/, const double *pBack /*[21]*/, const uint8_t *seq, int N, int i) {
    const int minPos = (i - 20) > 0 ? (i - 20) : 0;
    const int maxPos = (i + 20) < N ? (i + 20) : N;
    const int windowLength = maxPos - minPos;
    int sum = 0;
    const int16_t *row = mat + seq[i] * MB_AL;
    for (int j = minPos; j < maxPos; j++) sum += row[seq[j]];
    sum -= row[seq[i]];
    float d = (float)sum;
    d = (float)mb_ddiv((double)d, mb_dmul(-1.0, (double)(float)windowLength));
    for (int a = 0; a < MB_AL; a++) d = (float)mb_dadd((double)d, mb_dmul(pBack[a], (double)(float)row[a]));
    return d;   // scale == 1.0f
}
MB_DEV int8_t mb_round_bias_quarter(float b) {   // UngappedAlignment.cpp:394-396
    float t = mb_fdiv(b, 4.0f);
    t = (float)((b < 0.0) ? mb_dadd((double)t, -0.5) : mb_dadd((double)t, 0.5));
    return (int8_t)t;
}
MB_DEV int8_t mb_round_bias(float b) {           // StripedSmithWaterman.cpp:1233
    const double t = (b < 0.0) ? mb_dadd((double)b, -0.5) : mb_dadd((double)b, 0.5);
    return (int8_t)t;
}
