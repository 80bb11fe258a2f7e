// orf_kernels.cuh -- six-frame ORF fragments of nucleotide contigs and their translation, the step that feeds the
// prefilter: what `extractorfs` + `translatenucs` do with the flags `metaeuk predictexons` passes
//   --min-length 15 --max-length 32734 --orf-start-mode 1 (ANY_TO_STOP) --contig-start-mode 2 --contig-end-mode 2 --translation-table 1
// (--min-length, --max-length and --translation-table are parameters; the other three are fixed)
// Reference behaviour reproduced (not its code):
//   Orf::setSequence / findForward     reference/lib/mmseqs/src/commons/Orf.cpp:125-157,230-349   (stop-to-stop fragments per frame,
//                                      first fragment of a frame has an incomplete start, a fragment cut by the contig end an incomplete end)
//   extractorfs header coordinates     reference/lib/mmseqs/src/util/extractorfs.cpp:82-87, Orf.cpp:440-457
//   TranslateNucl::translate           reference/lib/mmseqs/src/commons/TranslateNucl.h:383-503  (IUPAC-aware codon table, case preserved)
//   key order                          DBWriter::createRenumberedDB (DBWriter.cpp:690-738): contig, then plus strand before minus, then
//                                      the order in which fragments END along the strand (all three frames interleaved)
// GPU organisation: in ANY_TO_STOP mode a fragment is fully determined by its terminating codon (a stop, or the last complete codon of
// its frame) and the previous terminator of the same frame, so one CTA per (contig, strand) streams the strand in tiles, finds
// terminators with one thread per 24 positions, and gets "previous terminator" (three frames at once) and the output rank from two block scans.  Pass 1 counts,
// a scan over (contig, strand) gives every CTA its output base, pass 2 writes the records in the reference's order, and a third
// kernel translates with one thread per residue.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MB_ORF_HD __host__ __device__ __forceinline__
#else
#define MB_ORF_HD inline
#endif

// Stop codons of the genetic code in use (--translation-table; Orf::Orf takes them from the table's amino-acid string, Orf.cpp:54-64) as a 64-bit set.
// The reference compares the upper-cased letters with the codon strings (Orf.cpp:266-269, isInCodons :206-228), so only plain A/C/G/T codons can match.
// Letter code m(u) = (u >> 1) & 3 for u in {A, C, G, T}: A 0, C 1, T 2, G 3; bit 16 m(u0) + 4 m(u1) + m(u2) of the set.
MB_ORF_HD bool orf_is_acgt(int u) { const unsigned d = (unsigned)(u - 'A'); return d < 20u && ((0x80045u >> d) & 1u); }   // bits 0 (A), 2 (C), 6 (G), 19 (T)
MB_ORF_HD unsigned orf_codon_bit(int u0, int u1, int u2) { return (((unsigned)u0 >> 1) & 3u) * 16u + (((unsigned)u1 >> 1) & 3u) * 4u + (((unsigned)u2 >> 1) & 3u); }
MB_ORF_HD bool orf_is_stop(int u0, int u1, int u2, unsigned long long stopSet) {   // u* = upper-cased letters (c & ~0x20)
    return orf_is_acgt(u0) && orf_is_acgt(u1) && orf_is_acgt(u2) && ((stopSet >> orf_codon_bit(u0, u1, u2)) & 1ull);
}
// the set from an NCBI amino-acid string in TCAG order (64 letters, '*' = stop)
inline unsigned long long orf_stop_set(const char *ncbieaa) {
    const char L[4] = {'T', 'C', 'A', 'G'}; unsigned long long m = 0;
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) for (int c = 0; c < 4; c++) if (ncbieaa[16 * a + 4 * b + c] == '*') m |= 1ull << orf_codon_bit(L[a], L[b], L[c]);
    return m;
}

#if defined(__CUDACC__) || defined(MB200_SIMT_EMU)    // everything below is device code (the helpers above also compile for the host: tests/emu checks them exhaustively)
#define MB_ORF_THREADS 256
#define MB_ORF_PER_THREAD 24                      // 8 codons of each frame
#define MB_ORF_TILE (MB_ORF_THREADS * MB_ORF_PER_THREAD)

struct MbOrfRec { uint32_t contig, from, to, flags; uint64_t aaOff; uint32_t aaLen; uint32_t strandFrom; };   // from/to: header coordinates

struct MbOrfTables {
    unsigned char comp[256];     // complement with '.' -> 'N' (Orf.cpp:47-51,146-150)
    unsigned char base[256];     // IUPAC letter -> 4-bit base set (TranslateNucl.h:334-362)
    char aaOf[4096];             // codon (three 4-bit codes) -> residue
    unsigned char aa2num[256];   // residue letter -> prefilter alphabet code (Sequence::mapSequence)
};

// character p of strand s of a contig (s = 0 plus, 1 minus); positions >= L read as 127 (the reference pads with CHAR_MAX)
__device__ __forceinline__ int orf_char(const unsigned char *seq, uint32_t L, int s, uint32_t p, const unsigned char *comp) {
    if (p >= L) return 127;
    unsigned char c = s ? seq[L - 1u - p] : seq[p];
    if (c == 'u') c = 't';                        // (only the lower-case branch survives in Orf::setSequence, Orf.cpp:141-144)
    return s ? comp[c] : c;
}

// exclusive block scans over MB_ORF_THREADS threads, several values per call so that one pair of barriers serves all of them
template <int N> __device__ __forceinline__ void orf_block_scan_add(uint32_t (&v)[N], uint32_t *sm /* N * warps */, uint32_t (&total)[N]) {
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint32_t inc[N];
#pragma unroll
    for (int k = 0; k < N; k++) inc[k] = v[k];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int k = 0; k < N; k++) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc[k], d); if ((int)lane >= d) inc[k] += y; }
    }
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < N; k++) sm[k * (MB_ORF_THREADS / 32) + w] = inc[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; k++) {
        uint32_t wp = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < MB_ORF_THREADS / 32; i++) { const uint32_t x = sm[k * (MB_ORF_THREADS / 32) + i]; if (i < (int)w) wp += x; tot += x; }
        total[k] = tot; v[k] = wp + inc[k] - v[k];
    }
    __syncthreads();
}

// exclusive "last terminator seen so far" for the three frames at once: max-scan of (position + 1), 0 = none
__device__ __forceinline__ void orf_block_scan_max3(uint32_t (&v)[3], uint32_t *sm /* 3 * warps */, uint32_t (&total)[3]) {
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint32_t inc[3] = {v[0], v[1], v[2]};
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc[k], d); if ((int)lane >= d) inc[k] = max(inc[k], y); }
    }
    uint32_t excl[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { excl[k] = __shfl_up_sync(0xffffffffu, inc[k], 1); if (lane == 0) excl[k] = 0; }
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < 3; k++) sm[k * (MB_ORF_THREADS / 32) + w] = inc[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint32_t wp = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < MB_ORF_THREADS / 32; i++) { const uint32_t x = sm[k * (MB_ORF_THREADS / 32) + i]; if (i < (int)w) wp = max(wp, x); tot = max(tot, x); }
        total[k] = tot; v[k] = max(wp, excl[k]);
    }
    __syncthreads();
}

// EMIT = false: counts[(contig, strand)] = (#fragments, #residues).  EMIT = true: writes the records at recBase / aaBase.
template <bool EMIT>
__global__ void __launch_bounds__(MB_ORF_THREADS) k_orf_scan(const unsigned char *__restrict__ nuc, const uint64_t *__restrict__ off, uint32_t nContigs,
                                                             const MbOrfTables *__restrict__ tab, uint32_t minLen, uint32_t maxLen, unsigned long long stopSet,
                                                             uint32_t *__restrict__ cntOut, uint64_t *__restrict__ aaOut,
                                                             const uint64_t *__restrict__ recBase, const uint64_t *__restrict__ aaBase, MbOrfRec *__restrict__ recs) {
    __shared__ unsigned char tile[MB_ORF_TILE + 8];
    __shared__ unsigned char comp[256];
    __shared__ uint32_t sm[3 * (MB_ORF_THREADS / 32)];
    comp[threadIdx.x] = tab->comp[threadIdx.x];
    for (uint32_t unit = blockIdx.x; unit < 2u * nContigs; unit += gridDim.x) {
        const uint32_t c = unit >> 1; const int s = (int)(unit & 1u);
        const unsigned char *seq = nuc + off[c]; const uint64_t L64 = off[c + 1] - off[c]; const uint32_t L = (uint32_t)L64;
        uint32_t carryTerm[3] = {0, 0, 0};      // (position + 1) of the last terminator of each frame in earlier tiles
        uint32_t carryCnt = 0; uint64_t carryAa = 0;
        __syncthreads();
        if (L >= 3) for (uint32_t tb = 0; tb < L; tb += MB_ORF_TILE) {
            for (uint32_t i = threadIdx.x; i < MB_ORF_TILE + 8; i += MB_ORF_THREADS) tile[i] = (unsigned char)orf_char(seq, L, s, tb + i, comp);
            __syncthreads();
            // ---- terminators among my 12 positions
            const uint32_t lb = threadIdx.x * MB_ORF_PER_THREAD;
            uint32_t termMask = 0, stopMask = 0, myLast[3] = {0, 0, 0};
#pragma unroll
            for (int j = 0; j < MB_ORF_PER_THREAD; j++) {
                const uint32_t p = tb + lb + j;
                const int c0 = tile[lb + j], c1 = tile[lb + j + 1], c2 = tile[lb + j + 2];
                const bool complete = p + 2 < L;            // (127 never occurs inside a contig: createdb stores printable characters)
                const int u0 = c0 & ~0x20, u1 = c1 & ~0x20, u2 = c2 & ~0x20;
                const bool stop = complete && orf_is_stop(u0, u1, u2, stopSet);
                const bool isLast = complete && (p + 5 >= L);
                if (stop) stopMask |= 1u << j;
                if (stop || isLast) { termMask |= 1u << j; myLast[j % 3] = p + 1; }
            }
            uint32_t prev[3] = {myLast[0], myLast[1], myLast[2]}, tot3[3];
            orf_block_scan_max3(prev, sm, tot3);
#pragma unroll
            for (int f = 0; f < 3; f++) { prev[f] = max(prev[f], carryTerm[f]); carryTerm[f] = max(carryTerm[f], tot3[f]); }
            // ---- fragments that end in my positions
            uint32_t myCnt = 0, myAa = 0;
            uint32_t eFrom[6], eLen[6], eInfo[6];     // fragments of one frame end >= 3 * minLen nucleotides apart: with minLen >= 4 at most 2 per frame in 24 positions
            int ne = 0;
#pragma unroll
            for (int j = 0; j < MB_ORF_PER_THREAD; j++) if ((termMask >> j) & 1u) {
                const int f = j % 3; const uint32_t p = tb + lb + j;
                const bool stop = (stopMask >> j) & 1u;
                const uint32_t from = prev[f] ? prev[f] - 1u + 3u : (uint32_t)f;
                const uint32_t len = (p - from) / 3u + (stop ? 0u : 1u);
                if (len >= minLen && len <= maxLen && len > 0) {
                    if (ne < 6) { eFrom[ne] = from; eLen[ne] = len; eInfo[ne] = (prev[f] ? 0u : 1u) | (stop ? 0u : 2u) | ((uint32_t)j << 8); ne++; }
                    myCnt++; myAa += len;
                }
                prev[f] = p + 1;
            }
            uint32_t ranks[2] = {myCnt, myAa}, tots[2];
            orf_block_scan_add<2>(ranks, sm, tots);
            const uint32_t rank = ranks[0], aaRank = ranks[1], totCnt = tots[0], totAa = tots[1];
            if (EMIT) {
                uint64_t a = aaBase[unit] + carryAa + aaRank;
                for (int e = 0; e < ne; e++) {
                    MbOrfRec r; r.contig = c; r.flags = (eInfo[e] & 3u) | (s ? 4u : 0u); r.aaOff = a; r.aaLen = eLen[e]; r.strandFrom = eFrom[e];
                    const uint32_t to = eFrom[e] + 3u * eLen[e] - 1u;   // last nucleotide of the last counted codon
                    r.from = s ? (L - 1u) - eFrom[e] : eFrom[e]; r.to = s ? (L - 1u) - to : to;
                    recs[recBase[unit] + carryCnt + rank + e] = r;
                    a += eLen[e];
                }
            }
            carryCnt += totCnt; carryAa += totAa;
            __syncthreads();
        }
        if (!EMIT && threadIdx.x == 0) { cntOut[unit] = carryCnt; aaOut[unit] = carryAa; }
    }
}

// one warp per fragment, lanes over its residues: codon -> amino acid letter (case preserved) and prefilter alphabet code
__global__ void __launch_bounds__(256) k_orf_translate(const unsigned char *__restrict__ nuc, const uint64_t *__restrict__ off, const MbOrfTables *__restrict__ tab,
                                                       const MbOrfRec *__restrict__ recs, uint64_t nRecs, int reversed, char *__restrict__ aa, uint8_t *__restrict__ codes) {
    __shared__ char aaOf[4096];
    __shared__ unsigned char comp[256], base[256], a2n[256];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) aaOf[i] = tab->aaOf[i];
    comp[threadIdx.x] = tab->comp[threadIdx.x]; base[threadIdx.x] = tab->base[threadIdx.x]; a2n[threadIdx.x] = tab->aa2num[threadIdx.x];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t nWarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < nRecs; w += nWarps) {
        const MbOrfRec r = recs[w];
        const unsigned char *seq = nuc + off[r.contig]; const uint32_t L = (uint32_t)(off[r.contig + 1] - off[r.contig]);
        const int s = (r.flags & 4u) ? 1 : 0;
        for (uint32_t k = lane; k < r.aaLen; k += 32u) {
            const uint32_t p = r.strandFrom + 3u * k;
            const int c0 = orf_char(seq, L, s, p, comp), c1 = orf_char(seq, L, s, p + 1, comp), c2 = orf_char(seq, L, s, p + 2, comp);
            char res = aaOf[256 * base[c0 & 255] + 16 * base[c1 & 255] + base[c2 & 255]];
            const bool lower = (c0 >= 'a' && c0 <= 'z') || (c1 >= 'a' && c1 <= 'z') || (c2 >= 'a' && c2 <= 'z');
            if (lower && res >= 'A' && res <= 'Z') res = (char)(res + 32);
            const uint64_t o = r.aaOff + (reversed ? r.aaLen - 1u - k : k);   // --reverse-fragments: the fragment back to front (reverseseq)
            aa[o] = res; codes[o] = a2n[(unsigned char)res];
        }
    }
}
#endif  // __CUDACC__
