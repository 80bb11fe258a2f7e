// oracle/mb200_oracle.cpp
//
// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Scalar CPU restatement of the reference's predictexons hot path:
// MMseqs2 k-mer prefilter -> gapped Smith-Waterman -> MetaEuk exon-chaining DP, plus the two neighbouring steps of section 8f:
// extractorfs + translatenucs (six-frame ORF fragments) before it and reduceredundancy after it.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// Parity: PINNED.  tests/test_oracle_golden.py checks every stage of this file against outputs of the
// unmodified reference (tests/golden/*.json.gz, produced by tests/golden/make_golden.py).
//
// Each function cites the reference file:line it restates ("M/" = reference/lib/mmseqs/).
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: the float composition-bias code must not be fused).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <cctype>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "submat_data.h"

namespace {

const int AL = 21;          // alphabet incl. X (=20)
const int XRES = 20;
const int KSZ = 6;          // k (M/src/prefiltering/IndexTable.h:439-449 -> 6 below 3.35 G residues)
const int SPACED_LEN = 10;  // spaced_seed_6 = 1101010011 (M/src/commons/Sequence.h:23)
const int SPACED_OFF[KSZ] = {0, 1, 3, 5, 8, 9};
const size_t TABLE = 64000000;  // 20^6
const int ROW3 = 8000;      // the reference pads rows to its SIMD width with -255 (ExtendedSubstitutionMatrix.cpp:23-24,56-59); padding is never read
const int N3 = 8000;

// ---------------------------------------------------------------------------------------------
// a1  SubstitutionMatrix ctor / readProbMatrix / BaseMatrix::generateSubMatrix
//     M/src/commons/SubstitutionMatrix.cpp:12-57,326-419 ; M/src/commons/BaseMatrix.cpp:100-159
// ---------------------------------------------------------------------------------------------
void build_submat(const double S[21][21], const double bg[21], double lambda, float bitFactorF, float scoreBiasF,
                  int16_t out[441], double pBackOut[21]) {
    double pBack[21];
    for (int i = 0; i < AL; i++) pBack[i] = bg[i];
    bool xIsPositive = false;   // SubstitutionMatrix.cpp:369-377
    for (int j = 0; j < AL; j++) if (S[XRES][j] > 0 || S[j][XRES] > 0) { xIsPositive = true; break; }
    if (!xIsPositive) for (int i = 0; i < AL - 1; i++) pBack[i] = pBack[i] * (1.0 - pBack[XRES]);  // :404-408
    double prob[21][21];
    for (int i = 0; i < AL; i++) for (int j = 0; j < AL; j++) prob[i][j] = std::exp(lambda * S[i][j]) * pBack[i] * pBack[j];  // :410-416
    // BaseMatrix::computeBackground (BaseMatrix.cpp:96-108): row sums, X forced to ANY_BACK
    double pb2[21];
    for (int i = 0; i < AL; i++) { pb2[i] = 0; for (int j = 0; j < AL; j++) pb2[i] += prob[i][j]; }
    pb2[AL - 1] = 1E-5;
    double bitFactor = bitFactorF, scoreBias = scoreBiasF;  // float ctor args widened (SubstitutionMatrix.cpp:12, BaseMatrix.cpp:141)
    for (int i = 0; i < AL; i++) for (int j = 0; j < AL; j++) {
        double sm = std::log2(prob[i][j] / (pb2[i] * pb2[j]));                 // BaseMatrix.cpp:125-127
        double v = bitFactor * sm + scoreBias;                                 // :149
        out[i * AL + j] = (short)((v < 0.0) ? v - 0.5 : v + 0.5);               // :150
    }
    for (int i = 0; i < AL; i++) pBackOut[i] = pBack[i];   // the member pBack calcLocalAaBiasCorrection reads
}

// ---------------------------------------------------------------------------------------------
// a5  SubstitutionMatrix::calcLocalAaBiasCorrection   M/src/commons/SubstitutionMatrix.cpp:79-109
// ---------------------------------------------------------------------------------------------
void comp_bias(const int16_t mat[441], const double pBack[21], const uint8_t *seq, int N, float *bias, float scale) {
    const int windowSize = 40;
    for (int i = 0; i < N; i++) {
        const int minPos = std::max(0, (i - windowSize / 2));
        const int maxPos = std::min(N, (i + windowSize / 2));
        const int windowLength = maxPos - minPos;
        int sumSubScores = 0;
        const int16_t *subMat = mat + seq[i] * AL;
        for (int j = minPos; j < maxPos; j++) sumSubScores += subMat[seq[j]];
        sumSubScores -= subMat[seq[i]];
        float deltaS_i = (float)sumSubScores;
        deltaS_i /= -1.0 * static_cast<float>(windowLength);
        for (int a = 0; a < AL; a++) deltaS_i += pBack[a] * static_cast<float>(subMat[a]);
        bias[i] = scale * deltaS_i;
    }
}

// ---------------------------------------------------------------------------------------------
// a2  ExtendedSubstitutionMatrix::calcScoreMatrix (k=3)  M/src/prefiltering/ExtendedSubstitutionMatrix.cpp:20-69
//     neighbours of every 3-mer sorted by descending score, stable w.r.t. cartesian order (first letter outermost)
// ---------------------------------------------------------------------------------------------
void build_sim3(const int16_t mat[441], int16_t *score, uint32_t *index) {
#pragma omp parallel
    {
        std::vector<std::pair<short, unsigned int>> tmp(N3);
#pragma omp for schedule(static)
        for (int a0 = 0; a0 < 20; a0++) for (int a1 = 0; a1 < 20; a1++) for (int a2 = 0; a2 < 20; a2++) {
            const unsigned int i_index = a0 + a1 * 20 + a2 * 400;   // Indexer::int2index, little-endian (Indexer.h:21-84)
            int j = 0;
            for (int b0 = 0; b0 < 20; b0++) for (int b1 = 0; b1 < 20; b1++) for (int b2 = 0; b2 < 20; b2++) {
                short s = 0;
                s += mat[a0 * AL + b0]; s += mat[a1 * AL + b1]; s += mat[a2 * AL + b2];
                tmp[j].first = s; tmp[j].second = b0 + b1 * 20 + b2 * 400; j++;
            }
            std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<short, unsigned int> &l, const std::pair<short, unsigned int> &r) { return l.first > r.first; });
            for (int z = 0; z < N3; z++) { score[(size_t)i_index * ROW3 + z] = tmp[z].first; index[(size_t)i_index * ROW3 + z] = tmp[z].second; }
        }
    }
}

struct Entry { uint32_t seqId; uint16_t pos; };   // IndexEntryLocal, 6 B packed in the reference (IndexTable.h:25-39)

struct Oracle {
    int16_t seedMat[441], ungMat[441], alnMat[441];
    double seedPBack[21], ungPBack[21], alnPBack[21];
    std::vector<int16_t> sim3Score; std::vector<uint32_t> sim3Index;
    // index
    std::vector<uint64_t> offsets; std::vector<Entry> entries;
    std::vector<uint8_t> masked; std::vector<uint64_t> moff;      // masked target residues (SequenceLookup)
    std::vector<uint8_t> unmasked; std::vector<uint64_t> uoff;    // unmasked target residues (align stage)
    uint32_t nT = 0; uint64_t totalResidues = 0;
    // parameters (MetaEuk defaults)
    int kmerThr = 127; int maxResListLen = 300; int minDiagScore = 15; int binSize = 2;
    int gapOpen = 11, gapExtend = 1; double evalThr = 100; int minAlnLen = 11;
    bool compBiasCorr = true;    // --comp-bias-corr (Prefiltering.cpp:52, Alignment.cpp:26)
};

// ---------------------------------------------------------------------------------------------
// a3  IndexBuilder::fillDatabase + IndexTable::{addKmerCount,addSequence,sortDBSeqLists}
//     M/src/prefiltering/IndexBuilder.cpp:55-239 ; IndexTable.h:133-173,348-401,182-189
//     input = ALREADY MASKED residues (tantan masking, Masker.cpp:14-58, is a host preprocessing input)
// ---------------------------------------------------------------------------------------------
void index_build(Oracle &o) {
    o.offsets.assign(TABLE + 1, 0);
    char idScore[21];
    for (int a = 0; a < AL; a++) idScore[a] = (char)o.seedMat[a * AL + a];   // getScoreLookup, IndexBuilder.cpp:10-21
    std::vector<std::vector<std::pair<uint32_t, uint16_t>>> perSeq(o.nT);     // (kmer, first pos) sorted by kmer
#pragma omp parallel for schedule(dynamic, 100)
    for (uint32_t t = 0; t < o.nT; t++) {
        const uint8_t *s = &o.masked[o.moff[t]];
        const int L = (int)(o.moff[t + 1] - o.moff[t]);
        std::vector<std::pair<uint32_t, uint16_t>> &v = perSeq[t];
        for (int p = 0; p + SPACED_LEN <= L; p++) {                          // Sequence::hasNextKmer, Sequence.h:92-94
            bool hasX = false; int sc = 0; uint32_t idx = 0; uint32_t pw = 1;
            for (int k = 0; k < KSZ; k++) { uint8_t c = s[p + SPACED_OFF[k]]; if (c == XRES) hasX = true; sc += idScore[c]; idx += c * pw; pw *= 20; }
            if (hasX) continue;                                               // IndexTable.h:141-143
            if (o.kmerThr > 0 && sc < o.kmerThr) continue;                    // IndexTable.h:144-152
            v.push_back(std::make_pair(idx, (uint16_t)p));
        }
        std::sort(v.begin(), v.end());                                        // by (kmer, pos)  IndexTable.h:386-388
        size_t w = 0;
        for (size_t i = 0; i < v.size(); i++) if (i == 0 || v[i].first != v[i - 1].first) v[w++] = v[i];   // first position only, :390-399
        v.resize(w);
    }
    for (uint32_t t = 0; t < o.nT; t++) for (auto &e : perSeq[t]) o.offsets[e.first + 1]++;
    for (size_t i = 0; i < TABLE; i++) o.offsets[i + 1] += o.offsets[i];
    o.entries.resize(o.offsets[TABLE]);
    std::vector<uint64_t> cur(o.offsets.begin(), o.offsets.end() - 1);
    // inserting targets in ascending seqId order yields lists sorted by (seqId,pos) == sortDBSeqLists (IndexTable.h:182-189)
    for (uint32_t t = 0; t < o.nT; t++) for (auto &e : perSeq[t]) { Entry &x = o.entries[cur[e.first]++]; x.seqId = t; x.pos = e.second; }
}

// ---------------------------------------------------------------------------------------------
// a9  UngappedAlignment::scalarDiagonalScoring / computeSingelSequenceScores
//     M/src/prefiltering/UngappedAlignment.cpp:30-43,416-430  (sequences < 32768 only)
// ---------------------------------------------------------------------------------------------
int diag_score(const int8_t *prof /*[L][21]*/, int qLen, const uint8_t *t, int tLen, uint16_t diag16) {
    int diagonal = (short)diag16;
    unsigned short dist = std::min((unsigned short)(0 - diag16), (unsigned short)(diag16 - 0));   // distanceFromDiagonal :364-369
    const int8_t *p; const uint8_t *d; unsigned int n;
    if (diagonal >= 0 && dist < (unsigned)qLen) { n = std::min((unsigned)tLen, (unsigned)qLen - dist); p = prof + dist * AL; d = t; }
    else if (diagonal < 0 && dist < (unsigned)tLen) { n = std::min((unsigned)tLen - dist, (unsigned)qLen); p = prof; d = t + dist; }
    else return 0;
    int max = 0, score = 0;
    for (unsigned int pos = 0; pos < n; pos++) {
        int curr = p[pos * AL + d[pos]];
        score = curr + score; score = (score < 0) ? 0 : score; max = (score > max) ? score : max;
    }
    return max;
}

struct Cand { uint32_t id; uint16_t diag; uint8_t count; };   // CounterResult (CacheFriendlyOperations.h:46-66)
struct PrefHit { uint32_t seqId; int32_t score; uint16_t diag; };   // hit_t (QueryMatcher.h:33-48)
struct PrefStats { uint64_t kmers, lookups, dbMatches, doubleHits, diagCells, results; };

// ---------------------------------------------------------------------------------------------
// a4,a6,a7,a8,a10  QueryMatcher::matchQuery / match  M/src/prefiltering/QueryMatcher.cpp:85-211,213-346
// ---------------------------------------------------------------------------------------------
int prefilter_query(const Oracle &o, const uint8_t *q, int L, std::vector<PrefHit> &out, PrefStats &st) {
    out.clear();
    if (L >= 32768) return -2;
    std::vector<float> bias(L > 0 ? L : 1);
    if (o.compBiasCorr) comp_bias(o.seedMat, o.seedPBack, q, L, bias.data(), 1.0f);   // QueryMatcher.cpp:91-93 (kmerSubMat)
    else std::fill(bias.begin(), bias.end(), 0.0f);                             // QueryMatcher.cpp:97-99
    // UngappedAlignment::createProfile  UngappedAlignment.cpp:385-414
    std::vector<int8_t> prof((size_t)(L > 0 ? L : 1) * AL);
    for (int pos = 0; pos < L; pos++) {
        float aaCorrBias = bias[pos];
        aaCorrBias = (aaCorrBias < 0.0) ? aaCorrBias / 4 - 0.5 : aaCorrBias / 4 + 0.5;
        char corr = static_cast<char>(aaCorrBias);
        for (int a = 0; a < AL; a++) prof[(size_t)pos * AL + a] = (char)(o.ungMat[q[pos] * AL + a] + corr);
    }
    // ---- match(): similar k-mers + index lists in arrival order (QueryMatcher.cpp:225-322)
    struct Hit { uint32_t id; uint16_t diag; };
    std::vector<Hit> hits;
    const size_t maxDbMatches = std::max((size_t)1000000, (size_t)o.nT) * 2;    // QueryMatcher.cpp:44
    for (int i = 0; i + SPACED_LEN <= L; i++) {
        float biasCorrection = 0;
        for (int k = 0; k < KSZ; k++) biasCorrection += bias[i + SPACED_OFF[k]];
        bool hasX = false; uint8_t w[KSZ];
        for (int k = 0; k < KSZ; k++) { w[k] = q[i + SPACED_OFF[k]]; if (w[k] == XRES) hasX = true; }
        if (hasX) continue;                                                     // :234-238
        short b = static_cast<short>((biasCorrection < 0.0) ? biasCorrection - 0.5 : biasCorrection + 0.5);
        short thr = std::max(o.kmerThr - b, 0);                                 // :240-241
        // KmerGenerator::generateKmerList / calculateArrayProduct  KmerGenerator.cpp:107-183,186-216
        const uint32_t idx0 = w[0] + w[1] * 20 + w[2] * 400, idx1 = w[3] + w[4] * 20 + w[5] * 400;
        const int16_t *s0 = &o.sim3Score[(size_t)idx0 * ROW3]; const uint32_t *k0 = &o.sim3Index[(size_t)idx0 * ROW3];
        const int16_t *s1 = &o.sim3Score[(size_t)idx1 * ROW3]; const uint32_t *k1 = &o.sim3Index[(size_t)idx1 * ROW3];
        const short possibleRest0 = s1[0];
        const short cutoff1 = thr - possibleRest0;
        for (int a = 0; a < N3; a++) {
            const short score_i = s0[a];
            if (score_i < cutoff1) break;
            const short cutoff2 = thr - score_i - 0;
            for (int c = 0; c < N3 && s1[c] >= cutoff2; c++) {
                const size_t kmer = (size_t)k0[a] + (size_t)k1[c] * 8000;
                st.kmers++;
                const uint64_t lo = o.offsets[kmer], hi = o.offsets[kmer + 1];   // IndexTable::getDBSeqList IndexTable.h:176-180
                st.lookups++;
                if (hits.size() + (hi - lo) >= maxDbMatches) return -3;         // reference overflow path (QueryMatcher.cpp:281-316) not restated
                for (uint64_t e = lo; e < hi; e++) { Hit h; h.id = o.entries[e].seqId; h.diag = (uint16_t)(i - o.entries[e].pos); hits.push_back(h); }
            }
        }
    }
    st.dbMatches += hits.size();
    // ---- a8 CacheFriendlyOperations<BIN>::findDuplicates  CacheFriendlyOperations.cpp:38-49,185-274,337-347
    const unsigned BIN = o.binSize; unsigned shift = 0; while ((1u << shift) < BIN) shift++;
    std::vector<std::vector<Hit>> bins(BIN);
    for (auto &h : hits) bins[h.id & (BIN - 1)].push_back(h);
    std::vector<uint8_t> last((o.nT >> shift) + 2, 0);
    std::vector<Cand> found;
    for (unsigned bin = 0; bin < BIN; bin++) {
        std::vector<Hit> tmp;
        for (auto &e : bins[bin]) {
            const unsigned hb = e.id >> shift; const uint8_t cur = (uint8_t)e.diag; const uint8_t prev = last[hb];
            if (cur == prev) tmp.push_back(e);
            last[hb] = cur;
        }
        for (size_t n = tmp.size(); n-- > 0;) last[tmp[n].id >> shift] = (uint8_t)((uint8_t)tmp[n].diag + 1);       // :236-241
        size_t base = found.size(); size_t cnt = 0; found.resize(base + tmp.size() + 1);
        for (size_t n = 0; n < tmp.size(); n++) {                                                                   // :244-262
            const unsigned hb = tmp[n].id >> shift;
            found[base + cnt].id = tmp[n].id; found[base + cnt].count = 0; found[base + cnt].diag = tmp[n].diag;
            cnt += (last[hb] != (uint8_t)tmp[n].diag) ? 1 : 0;
            last[hb] = (uint8_t)tmp[n].diag;
        }
        found.resize(base + cnt);
        for (auto &e : bins[bin]) last[e.id >> shift] = 0;                                                          // :264-271
    }
    st.doubleHits += found.size();
    // ---- a9 ungapped diagonal scores  UngappedAlignment.cpp:331-362 (count==0 entries only; all are 0 here)
    for (auto &c : found) {
        const uint8_t *t = &o.masked[o.moff[c.id]]; const int tLen = (int)(o.moff[c.id + 1] - o.moff[c.id]);
        if (tLen >= 32768) return -2;
        int mx = diag_score(prof.data(), L, t, tLen, c.diag);
        c.count = (uint8_t)std::min(255, mx);
        int diagonal = (short)c.diag; unsigned short dist = std::min((unsigned short)(0 - c.diag), (unsigned short)(c.diag - 0));
        st.diagCells += (diagonal >= 0) ? (dist < (unsigned)L ? std::min((unsigned)tLen, (unsigned)L - dist) : 0) : (dist < (unsigned)tLen ? std::min((unsigned)tLen - dist, (unsigned)L) : 0);
    }
    // ---- a10 keepMaxScoreElementOnly  CacheFriendlyOperations.cpp:74-82,350-380
    {
        std::vector<std::vector<Cand>> cb(BIN);
        for (auto &c : found) cb[c.id & (BIN - 1)].push_back(c);
        size_t n = 0;
        for (unsigned bin = 0; bin < BIN; bin++) {
            for (auto &e : cb[bin]) { uint8_t &d = last[e.id >> shift]; d = std::max(d, e.count); }
            for (auto &e : cb[bin]) {
                uint8_t &d = last[e.id >> shift];
                found[n] = e;
                bool keep = (d == e.count);
                n += keep; d = d * (1 - keep);
            }
        }
        found.resize(n);
        std::fill(last.begin(), last.end(), 0);   // the reference memsets at the start of the next keepMaxElement (:353)
    }
    // ---- histogram, threshold, stable counting sort (QueryMatcher.cpp:153-176,498-523; QueryMatcher.h:206-216)
    unsigned int scoreSizes[256]; memset(scoreSizes, 0, sizeof(scoreSizes));
    for (auto &c : found) scoreSizes[c.count]++;
    const size_t maxHits = std::min((size_t)o.maxResListLen, (size_t)o.nT);
    auto computeScoreThreshold = [&](const unsigned int *ss) { size_t fh = 0; size_t thr = 0; for (thr = 255; thr > 0; thr--) { fh += ss[thr]; if (fh >= maxHits) break; } return (unsigned)thr; };
    unsigned int diagonalThr = std::max((unsigned)o.minDiagScore, computeScoreThreshold(scoreSizes));
    auto radixSort = [&](const unsigned int *ss, const std::vector<Cand> &in, unsigned thr, std::vector<Cand> &res) {
        size_t N = in.size(); res.assign(N + 1, Cand());
        std::vector<size_t> ptr(256); size_t prev = N;
        for (int i = 0; i < 256; i++) { ptr[i] = prev - ss[i]; prev = ptr[i]; }
        size_t above = 0;
        for (size_t i = 0; i < N; i++) if (in[i].count >= thr) { above++; res[ptr[in[i].count]++] = in[i]; }
        return above;
    };
    auto exactScore = [&](const Cand &c) {   // scoreSingelSequenceByCounterResult  UngappedAlignment.cpp:432-446
        const uint8_t *t = &o.masked[o.moff[c.id]]; const int tLen = (int)(o.moff[c.id + 1] - o.moff[c.id]);
        return diag_score(prof.data(), L, t, tLen, c.diag);
    };
    std::vector<Cand> sorted;
    size_t above = radixSort(scoreSizes, found, diagonalThr, sorted);
    int rescale = 0; unsigned getThr = diagonalThr;
    if (diagonalThr >= 255) {   // scoreIsTruncated: rescoreHits  QueryMatcher.cpp:160-170,525-544
        memset(scoreSizes, 0, sizeof(scoreSizes));
        int maxSelfScore = diag_score(prof.data(), L, q, L, 0);
        maxSelfScore = maxSelfScore - 255; maxSelfScore = std::max(1, maxSelfScore); maxSelfScore = std::min(maxSelfScore, 65535);
        float fltMax = (float)maxSelfScore; size_t elements = 0;
        for (size_t i = 0; i < above && sorted[i].count >= 255; i++) {
            unsigned int ns = (unsigned)exactScore(sorted[i]); ns -= 255;
            float sc = (float)std::min(ns, 65535u);
            sorted[i].count = (unsigned char)((sc / fltMax) * 255.0f + 0.5);
            scoreSizes[sorted[i].count]++; elements++;
        }
        std::vector<Cand> in(sorted.begin(), sorted.begin() + elements), res2;
        above = radixSort(scoreSizes, in, 0, res2); sorted.swap(res2);
        rescale = maxSelfScore; getThr = 0;
    }
    // ---- getResult<UNGAPPED_DIAGONAL_SCORE>  QueryMatcher.cpp:363-420
    for (size_t i = 0; i < above && out.size() < maxHits; i++) {
        const Cand &c = sorted[i];
        if (c.count >= getThr) {
            PrefHit h; h.seqId = c.id; h.score = c.count; h.diag = c.diag;
            if (rescale != 0) h.score = (int)(255u + ((unsigned)c.count * (unsigned)rescale / 255u));
            else if ((int)c.count >= 255) h.score = exactScore(c);
            out.push_back(h);
        }
    }
    // ---- final order  QueryMatcher.cpp:203-209, QueryMatcher.h:38-48
    std::sort(out.begin(), out.end(), [](const PrefHit &a, const PrefHit &b) {
        if (abs(a.score) > abs(b.score)) return true; if (abs(b.score) > abs(a.score)) return false; return a.seqId < b.seqId; });
    st.results += out.size();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// a11-a13  SmithWaterman::ssw_init / sw_sse2_byte|word / ssw_align_private  -- restated as exact Gotoh
//     M/src/alignment/StripedSmithWaterman.cpp:1216-1345, 638-940, 309-545  
// ---------------------------------------------------------------------------------------------
struct SwEnd { int score, ref, read; };
// one pass: columns = target residues in order given by (tBegin, tStep, nCols); rows = query profile rows 0..qn-1.
// Returns max score, FIRST column reaching it (strictly greater than before), smallest row equal to max in that column.
// terminate>0: stop after the first column whose column maximum == terminate (reverse pass, :900).
SwEnd sw_pass(const int16_t *alnMat, const uint8_t *qres, const int8_t *qbias, int qn, const uint8_t *t, int tBegin, int tStep, int nCols,
              int gapO, int gapE, int terminate) {
    std::vector<int> H(qn + 1, 0), E(qn + 1, 0), Hmax;   // H[j+1] = H(prev column, row j)
    int max = 0, end_db = -1, end_q = qn - 1;
    for (int c = 0, ti = tBegin; c < nCols; c++, ti += tStep) {
        const int16_t *row = alnMat + t[ti] * AL;
        int diagPrev = 0;   // H(col-1,row-1)
        int F = 0, hUp = 0; // within this column
        int colMax = 0;
        std::vector<int> Hnew(qn + 1, 0);
        for (int j = 0; j < qn; j++) {
            int s = row[qres[j]] + qbias[j];
            int e = E[j + 1];                                   // E(col,j) prepared from previous column
            int f = (j == 0) ? 0 : std::max(F - gapE, hUp - gapO);
            int h = diagPrev + s; if (h < 0) h = 0; if (e > h) h = e; if (f > h) h = f;
            diagPrev = H[j + 1];
            Hnew[j + 1] = h; F = f; hUp = h;
            int en = std::max(e - gapE, h - gapO); if (en < 0) en = 0;   // unsigned-saturating in the reference; equivalent
            E[j + 1] = en;
            if (h > colMax) colMax = h;
        }
        H.swap(Hnew);
        if (colMax > max) { max = colMax; end_db = ti; Hmax = H; }
        if (terminate > 0 && colMax == terminate) break;
    }
    if (end_db >= 0) for (int j = 0; j < qn; j++) if (Hmax[j + 1] == max) { if (j < end_q) end_q = j; break; }
    SwEnd r; r.score = max; r.ref = end_db; r.read = end_q; return r;
}

struct AlnRes { int score; int qStart, qEnd, dbStart, dbEnd; double evalue; int bitScore; float seqId; float qCov, tCov; int alnLen; int passed; };

// ---------------------------------------------------------------------------------------------
// a14  EvalueComputation (ALP, fixed BLOSUM62 11/1 parameter set)  M/src/alignment/EvalueComputation.h:18-40,56-69
//      Sls::AlignmentEvaluer::area -> pvalues::get_appr_tail_prob_with_cov_without_errors
//      M/lib/alp/sls_alignment_evaluer.cpp:989-1028 ; M/lib/alp/sls_pvalues.cpp:342-364,366-532 ; sls_basic.cpp:107-165
// ---------------------------------------------------------------------------------------------
struct Alp {
    double lambda = 0.27359865037097330642, K = 0.044620920658722244834;
    double a1 = 1.5938724404943873658, b1 = -19.959867650284412122, a2 = 1.5938724404943873658, b2 = -19.959867650284412122;
    double alpha1 = 30.455610143099914211, beta1 = -622.28684628915891608, alpha2 = 30.455610143099914211, beta2 = -622.28684628915891608;
    double sigma = 29.602444874818868215, tau = -601.81087985041381216;
    double vi_y_thr, vj_y_thr, c_y_thr, logK;
    Alp() {
        const double nat_cut_off_in_max = 2.0;   // sls_pvalues.hpp
        vi_y_thr = std::max(nat_cut_off_in_max * alpha2 / lambda, 0.0);   // alpha_I = d_alpha2
        vj_y_thr = std::max(nat_cut_off_in_max * alpha1 / lambda, 0.0);   // alpha_J = d_alpha1
        c_y_thr = std::max(nat_cut_off_in_max * sigma / lambda, 0.0);
        logK = std::log(K);
    }
    // sls_basic::normal_probability(double) -- the one-argument overload the area code calls (M/lib/alp/sls_basic.hpp:195-198)
    static double normal_probability(double x_) { return 0.5 * erfc(-sqrt(0.5) * x_); }
    double area(double y_, double m_ /*seqlen2 = db residues*/, double n_ /*seqlen1 = query len*/) const {
        // ai_hat = a_I = d_a2 ; aj_hat = a_J = d_a1 (sls_alignment_evaluer.cpp:679-683)
        const double pi = 3.1415926535897932384626433832795;
        const double const_val = 1 / sqrt(2.0 * pi);   // M/lib/alp/sls_basic.hpp:58-59
        double m_li_y = m_ - (a2 * y_ + b2);
        double vi_y = std::max(vi_y_thr, alpha2 * y_ + beta2);
        double sqrt_vi_y = sqrt(vi_y);
        double m_F = (sqrt_vi_y == 0.0) ? 1e100 : m_li_y / sqrt_vi_y;
        double P_m_F = normal_probability(m_F);
        double E_m_F = -const_val * exp(-0.5 * m_F * m_F);
        double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;
        double n_lj_y = n_ - (a1 * y_ + b1);
        double vj_y = std::max(vj_y_thr, alpha1 * y_ + beta1);
        double sqrt_vj_y = sqrt(vj_y);
        double n_F = (sqrt_vj_y == 0.0) ? 1e100 : n_lj_y / sqrt_vj_y;
        double P_n_F = normal_probability(n_F);
        double E_n_F = -const_val * exp(-0.5 * n_F * n_F);
        double p2 = n_lj_y * P_n_F - sqrt_vj_y * E_n_F;
        double c_y = std::max(c_y_thr, sigma * y_ + tau);
        return p1 * p2 + c_y * (P_m_F * P_n_F);
    }
    double evalue(double score, double qLen, double dbRes) const { return (K * exp(-lambda * score)) * area(score, qLen, dbRes); }
    double bitScore(double score) const { return (lambda * score - logK) / log(2.0); }
};

void align_pair(const Oracle &o, const Alp &alp, const uint8_t *q, int qLen, const int8_t *cb, const uint8_t *t, int tLen, AlnRes &r) {
    memset(&r, 0, sizeof(r)); r.qStart = r.dbStart = -1;
    SwEnd f = sw_pass(o.alnMat, q, cb, qLen, t, 0, 1, tLen, o.gapOpen, o.gapExtend, 0);
    r.score = f.score; r.dbEnd = f.ref; r.qEnd = f.read;
    if (f.ref == -1) { r.evalue = 0; r.passed = 0; r.qEnd = qLen - 1; return; }
    r.evalue = alp.evalue(f.score, qLen, (double)o.totalResidues);            // StripedSmithWaterman.cpp:389
    r.bitScore = (int)(alp.bitScore(f.score) + 0.5);                          // Matcher.cpp:130
    if (r.evalue > o.evalThr) { r.passed = 0; return; }                       // :390-398 (covThr 0 always passes)
    // reverse pass on the reversed query prefix / target prefix  :400-476
    std::vector<uint8_t> qr(f.read + 1); std::vector<int8_t> cbr(f.read + 1);
    for (int j = 0; j <= f.read; j++) { qr[j] = q[f.read - j]; cbr[j] = cb[f.read - j]; }
    SwEnd b = sw_pass(o.alnMat, qr.data(), cbr.data(), f.read + 1, t, f.ref, -1, f.ref + 1, o.gapOpen, o.gapExtend, f.score);
    if (b.score != f.score) { r.passed = -1; return; }                        // the reference EXITs here (:466-473)
    r.dbStart = b.ref; r.qStart = f.read - b.read;
    auto cov = [](unsigned s, unsigned e, unsigned len) { return (std::min(len, std::max(s, e)) - std::min(s, e) + 1) / (float)len; };   // :1671-1673
    r.qCov = cov(r.qStart, r.qEnd, qLen); r.tCov = cov(r.dbStart, r.dbEnd, tLen);
    r.alnLen = std::max(abs(r.qEnd - r.qStart), abs(r.dbEnd - r.dbStart)) + 1;                     // Matcher.cpp:156-158
    unsigned qa = std::max((unsigned)(r.qEnd - r.qStart), 1u), ta = std::max((unsigned)(r.dbEnd - r.dbStart), 1u);
    float sid = ((uint16_t)r.score / static_cast<float>(std::max(qa, ta))) * 0.1656 + 0.1141;      // Matcher.cpp:160-164
    sid = std::min(sid, 1.0f); r.seqId = std::max(0.0f, sid);
    r.passed = (r.evalue <= o.evalThr && r.alnLen >= o.minAlnLen) ? 1 : 0;                         // Alignment.cpp:548-567
}

// int8 composition bias for the align stage  StripedSmithWaterman.cpp:1229-1237
void aln_comp_bias(const Oracle &o, const uint8_t *q, int L, int8_t *cb) {
    std::vector<float> tmp(L > 0 ? L : 1);
    if (!o.compBiasCorr) { memset(cb, 0, (size_t)(L > 0 ? L : 0)); return; }               // StripedSmithWaterman.cpp:1238-1240
    comp_bias(o.alnMat, o.alnPBack, q, L, tmp.data(), 1.0f);
    for (int i = 0; i < L; i++) cb[i] = (int8_t)((tmp[i] < 0.0) ? tmp[i] - 0.5 : tmp[i] + 0.5);
}

}  // namespace

// =============================================================================================
//  C API (ctypes).  All arrays caller-owned unless stated.
// =============================================================================================
extern "C" {

struct mbo_hit { uint32_t seqId; int32_t score; int32_t diag; };   // diag = (short)diagonal as written to pref DB
struct mbo_aln { uint32_t seqId; int32_t score; int32_t bitScore; int32_t qStart, qEnd, dbStart, dbEnd, alnLen, passed; float seqId_est; float pad; double evalue; };
struct mbo_stats { uint64_t kmers, lookups, dbMatches, doubleHits, diagCells, results; };

void *mbo_create(void) {
    Oracle *o = new Oracle();
    build_submat(MB200_VTML80_HALFBITS, MB200_VTML80_BACKGROUND, MB200_VTML80_LAMBDA, 8.0f, -0.2f, o->seedMat, o->seedPBack);     // Prefiltering.cpp:68, :990
    build_submat(MB200_BLOSUM62_HALFBITS, MB200_BLOSUM62_BACKGROUND, MB200_BLOSUM62_LAMBDA, 2.0f, -0.2f, o->ungMat, o->ungPBack);  // Prefiltering.cpp:69
    build_submat(MB200_BLOSUM62_HALFBITS, MB200_BLOSUM62_BACKGROUND, MB200_BLOSUM62_LAMBDA, 2.0f, 0.0f, o->alnMat, o->alnPBack);   // Alignment.cpp:152
    return o;
}
void mbo_destroy(void *h) { delete (Oracle *)h; }
void mbo_set_params(void *h, int kmerThr, int maxResListLen, int minDiagScore, int binSize, int gapOpen, int gapExtend, double evalThr, int minAlnLen) {
    Oracle *o = (Oracle *)h; o->kmerThr = kmerThr; o->maxResListLen = maxResListLen; o->minDiagScore = minDiagScore; o->binSize = binSize;
    o->gapOpen = gapOpen; o->gapExtend = gapExtend; o->evalThr = evalThr; o->minAlnLen = minAlnLen;
}
void mbo_set_comp_bias_corr(void *h, int on) { ((Oracle *)h)->compBiasCorr = on != 0; }
void mbo_get_matrices(void *h, int16_t *seed, int16_t *ung, int16_t *aln, double *seedPBack, double *alnPBack) {
    Oracle *o = (Oracle *)h; memcpy(seed, o->seedMat, 882); memcpy(ung, o->ungMat, 882); memcpy(aln, o->alnMat, 882);
    memcpy(seedPBack, o->seedPBack, 168); memcpy(alnPBack, o->alnPBack, 168);
}
void mbo_build_sim3(void *h) { Oracle *o = (Oracle *)h; o->sim3Score.resize((size_t)N3 * ROW3); o->sim3Index.resize((size_t)N3 * ROW3); build_sim3(o->seedMat, o->sim3Score.data(), o->sim3Index.data()); }
const int16_t *mbo_sim3_score(void *h) { return ((Oracle *)h)->sim3Score.data(); }
const uint32_t *mbo_sim3_index(void *h) { return ((Oracle *)h)->sim3Index.data(); }

// targets: masked + unmasked residue codes (0..20), offsets[nT+1]
int mbo_set_targets(void *h, const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t nT) {
    Oracle *o = (Oracle *)h; o->nT = nT; o->moff.assign(off, off + nT + 1); o->uoff = o->moff;
    o->masked.assign(masked, masked + off[nT]); o->unmasked.assign(unmasked, unmasked + off[nT]); o->totalResidues = off[nT];
    for (uint32_t t = 0; t < nT; t++) if (off[t + 1] - off[t] >= 32768) return -2;
    index_build(*o);
    return 0;
}
uint64_t mbo_index_entries(void *h) { return ((Oracle *)h)->entries.size(); }
// copy out: offsets as u64[20^6+1], entries packed 6 B (u32 seqId, u16 pos) exactly as IndexEntryLocal
void mbo_index_copy(void *h, uint64_t *offsets, uint8_t *entries6) {
    Oracle *o = (Oracle *)h; memcpy(offsets, o->offsets.data(), (TABLE + 1) * 8);
    for (size_t i = 0; i < o->entries.size(); i++) { memcpy(entries6 + i * 6, &o->entries[i].seqId, 4); memcpy(entries6 + i * 6 + 4, &o->entries[i].pos, 2); }
}
void mbo_comp_bias(void *h, int which, const uint8_t *q, int L, float *out) {
    Oracle *o = (Oracle *)h; if (which == 0) comp_bias(o->seedMat, o->seedPBack, q, L, out, 1.0f); else comp_bias(o->alnMat, o->alnPBack, q, L, out, 1.0f);
}

// prefilter a batch: q residues concatenated, qOff[nQ+1].  hits: caller buffer of nQ*maxResListLen, hitCount[nQ].
int mbo_prefilter(void *h, const uint8_t *q, const uint64_t *qOff, uint32_t nQ, mbo_hit *hits, uint32_t *hitCount, mbo_stats *stats, int threads) {
    Oracle *o = (Oracle *)h; int err = 0; PrefStats tot; memset(&tot, 0, sizeof(tot));
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        std::vector<PrefHit> out; PrefStats st; memset(&st, 0, sizeof(st));
#pragma omp for schedule(dynamic, 1)
        for (uint32_t i = 0; i < nQ; i++) {
            int rc = prefilter_query(*o, q + qOff[i], (int)(qOff[i + 1] - qOff[i]), out, st);
            if (rc != 0) {
#pragma omp critical
                err = rc;
                hitCount[i] = 0; continue;
            }
            hitCount[i] = (uint32_t)out.size();
            for (size_t k = 0; k < out.size(); k++) { mbo_hit &x = hits[(size_t)i * o->maxResListLen + k]; x.seqId = out[k].seqId; x.score = out[k].score; x.diag = (short)out[k].diag; }
        }
#pragma omp critical
        { tot.kmers += st.kmers; tot.lookups += st.lookups; tot.dbMatches += st.dbMatches; tot.doubleHits += st.doubleHits; tot.diagCells += st.diagCells; tot.results += st.results; }
    }
    if (stats) { stats->kmers = tot.kmers; stats->lookups = tot.lookups; stats->dbMatches = tot.dbMatches; stats->doubleHits = tot.doubleHits; stats->diagCells = tot.diagCells; stats->results = tot.results; }
    return err;
}

// align: for query i, its prefilter hits hitSeqId[hitOff[i]..hitOff[i+1]) in prefilter order. alns has the same layout.
// (Alignment::run hot loop, M/src/alignment/Alignment.cpp:312-405; sorting by Matcher::compareHits is left to the caller.)
int mbo_align(void *h, const uint8_t *q, const uint64_t *qOff, uint32_t nQ, const uint32_t *hitSeqId, const uint64_t *hitOff, mbo_aln *alns, uint64_t *cells, int threads) {
    Oracle *o = (Oracle *)h; Alp alp; int err = 0; uint64_t totCells = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 5) reduction(+ : totCells)
    for (uint32_t i = 0; i < nQ; i++) {
        const uint8_t *qs = q + qOff[i]; const int L = (int)(qOff[i + 1] - qOff[i]);
        if (hitOff[i + 1] == hitOff[i]) continue;
        std::vector<int8_t> cb(L > 0 ? L : 1); aln_comp_bias(*o, qs, L, cb.data());
        for (uint64_t k = hitOff[i]; k < hitOff[i + 1]; k++) {
            const uint32_t t = hitSeqId[k]; AlnRes r;
            const uint8_t *ts = &o->unmasked[o->uoff[t]]; const int tLen = (int)(o->uoff[t + 1] - o->uoff[t]);
            align_pair(*o, alp, qs, L, cb.data(), ts, tLen, r);
            if (r.passed < 0) err = -4;
            mbo_aln &a = alns[k]; a.seqId = t; a.score = r.score; a.bitScore = r.bitScore; a.qStart = r.qStart; a.qEnd = r.qEnd; a.dbStart = r.dbStart; a.dbEnd = r.dbEnd;
            a.alnLen = r.alnLen; a.passed = r.passed; a.seqId_est = r.seqId; a.pad = 0; a.evalue = r.evalue;
            totCells += (uint64_t)L * tLen + ((r.dbStart >= 0) ? (uint64_t)(r.qEnd + 1) * (r.dbEnd + 1) : 0);
        }
    }
    if (cells) *cells = totCells;
    return err;
}

double mbo_evalue(double score, double qLen, double dbResidues) { Alp a; return a.evalue(score, qLen, dbResidues); }
int mbo_bitscore(double score) { Alp a; return (int)(a.bitScore(score) + 0.5); }

// ---------------------------------------------------------------------------------------------
// a16-a18  exon-chaining DP   reference/src/exonpredictor/collectoptimalset.cpp:36-222,224-424 ;
//          reference/src/commons/PredictionParser.h:15-65,166-182
// One call = one contig.  cands sorted by (targetKey, orfKey) as resultspercontig emits them (resultspercontig.cpp:17-32,182).
// ---------------------------------------------------------------------------------------------
struct mbo_cand { uint32_t targetKey; int32_t bitScore; int32_t qStart, qEnd; int32_t tStart, tEnd, tLen; uint32_t orfKey; int32_t fragFrom, fragTo; };
struct mbo_exon { uint32_t candIdx; int32_t contigStart, contigEnd, nucLen; };   // contigStart/End signed as in PotentialExon
struct mbo_call { uint32_t targetKey; int32_t strand; int32_t totalBitScore; int32_t nExons; uint32_t exonBegin; int32_t lowCoord, highCoord; int32_t pad; double evalue; };

struct PE { uint32_t cand; int strand; int bitScore; int tStart, tEnd, tLen; double tCov; int cStart, cEnd, nucLen, aaLen; bool used; };

static int dp_find(std::vector<PE> &c, std::vector<PE> &opt, size_t minIntron, size_t maxIntron, size_t maxOv, int gapO, int gapE, double tcovThr) {
    size_t n = c.size(); if (n == 0) return 0;
    std::stable_sort(c.begin(), c.end(), [](const PE &a, const PE &b) {       // PredictionParser.h:166-182
        if (a.used < b.used) return true; if (a.used > b.used) return false;
        if (a.cStart < b.cStart) return true; if (a.cStart > b.cStart) return false;
        if (a.cEnd < b.cEnd) return true; if (a.cEnd > b.cEnd) return false; return false; });
    size_t firstUsed = n; for (size_t i = 0; i < n; i++) if (c[i].used) { firstUsed = i; break; }
    c.resize(firstUsed); n = c.size();
    if (n == 0) return 0;   // (the reference would read c[0] of an empty vector here; unreachable with max-exon-sets loops guarded by size()>0 only before the call)
    int targetLength = c[0].tLen;
    struct Row { size_t prev; int score; size_t nEx; double cov; int aaLen; };
    std::vector<Row> R(n);
    for (size_t i = 0; i < n; i++) { R[i].prev = i; R[i].score = c[i].bitScore; R[i].nEx = 1; R[i].cov = c[i].tCov; R[i].aaLen = c[i].aaLen; }
    int best = 0; size_t last = 0;
    for (size_t cur = 0; cur < n; cur++) {
        for (size_t prev = 0; prev < cur; prev++) {
            const PE &a = c[prev], &b = c[cur];
            // isPairCompatible  collectoptimalset.cpp:36-82
            if (a.strand != b.strand) continue;
            if (b.cEnd < a.cEnd) continue;
            int diffOnContig = b.cStart - a.cEnd - 1; if (diffOnContig < 0) continue;
            size_t dnn = abs(diffOnContig); if (dnn < minIntron || dnn > maxIntron) continue;
            int diffAAs = b.tStart - a.tEnd - 1; size_t ov = 0;
            if (diffAAs < 0) { ov = abs(diffAAs); if (ov > maxOv) continue; }
            if (b.tStart < a.tStart) continue;
            // getPenaltyForProtCoords  :84-106
            int pen; if (diffAAs < 0) pen = gapO + gapE * (abs(diffAAs) - 1); else if (diffAAs <= 1) pen = 0; else pen = gapO + gapE * (diffAAs - 1);
            size_t nEx = R[prev].nEx + 1; int bonus = (int)log2(nEx);
            int sc = R[prev].score + pen + b.bitScore + bonus;
            double cov = R[prev].cov + b.tCov; int aal = R[prev].aaLen + b.aaLen - (int)ov;
            if (sc > R[cur].score) { R[cur].prev = prev; R[cur].score = sc; R[cur].nEx = nEx; R[cur].cov = cov; R[cur].aaLen = aal; }
        }
        if ((double)R[cur].aaLen / (double)targetLength >= tcovThr) if (R[cur].score > best) { last = cur; best = R[cur].score; }
    }
    if (best == 0) return 0;
    size_t e = last;
    while (R[e].prev != e) { opt.push_back(c[e]); c[e].used = true; e = R[e].prev; }
    opt.push_back(c[e]); c[e].used = true;
    std::reverse(opt.begin(), opt.end());
    return best;
}

// returns number of calls; exons appended to exonsOut (caller sizes both to nCands*maxSets ; calls to 2*nTargets*maxSets)
int mbo_chain(const mbo_cand *cands, uint32_t n, uint64_t totalTargetResidues, int minIntron, int maxIntron, int minExonAa, int maxOverlap,
              int setGapOpen, int setGapExtend, double tcov, double evalThr, int maxSets, mbo_call *calls, mbo_exon *exonsOut, uint32_t *nExonsOut) {
    uint32_t nCalls = 0, nEx = 0;
    uint32_t i = 0;
    while (i < n) {
        uint32_t j = i; while (j < n && cands[j].targetKey == cands[i].targetKey) j++;
        std::vector<PE> plus, minus;
        for (uint32_t k = i; k < j; k++) {
            const mbo_cand &c = cands[k]; PE e; e.cand = k; e.bitScore = c.bitScore; e.tStart = c.tStart; e.tEnd = c.tEnd; e.tLen = c.tLen; e.used = false;
            if (c.fragFrom < c.fragTo) { e.cStart = c.fragFrom + c.qStart * 3; e.cEnd = c.fragFrom + c.qEnd * 3 + 2; e.strand = 1; }      // PredictionParser.h:40-51
            else { e.cStart = -1 * (c.fragFrom - c.qStart * 3); e.cEnd = -1 * (c.fragFrom - c.qEnd * 3 - 2); e.strand = -1; }
            e.nucLen = e.cEnd - e.cStart + 1; if (e.nucLen % 3 != 0) return -1; e.aaLen = e.nucLen / 3;
            e.tCov = (double)(e.tEnd - e.tStart + 1) / e.tLen;
            size_t aal = std::abs(e.nucLen) / 3;
            if (aal >= (size_t)minExonAa) { if (e.strand == 1) plus.push_back(e); else minus.push_back(e); }                              // collectoptimalset.cpp:360-367
        }
        int it = 0;
        while (it < maxSets && (plus.size() > 0 || minus.size() > 0)) {                                                                   // :313-317
            std::vector<PE> op, om;
            int sp = dp_find(plus, op, minIntron, maxIntron, maxOverlap, setGapOpen, setGapExtend, tcov);
            int sm = dp_find(minus, om, minIntron, maxIntron, maxOverlap, setGapOpen, setGapExtend, tcov);
            for (int s = 0; s < 2; s++) {
                std::vector<PE> &os = s == 0 ? op : om; int tot = s == 0 ? sp : sm;
                if (os.size() == 0) continue;
                double l2 = log2(totalTargetResidues) + log2(2) - tot; double ev = pow(2, l2);                                            // :323-325
                if (ev <= evalThr) {
                    mbo_call &cl = calls[nCalls++]; cl.targetKey = cands[i].targetKey; cl.strand = s == 0 ? 1 : -1; cl.totalBitScore = tot; cl.nExons = (int)os.size();
                    cl.exonBegin = nEx; cl.evalue = ev; cl.pad = 0;
                    cl.lowCoord = (s == 0) ? os.front().cStart : (-1 * os.back().cEnd); cl.highCoord = (s == 0) ? os.back().cEnd : (-1 * os.front().cStart);   // PredictionParser.h:236-237
                    for (auto &e : os) { mbo_exon &x = exonsOut[nEx++]; x.candIdx = e.cand; x.contigStart = e.cStart; x.contigEnd = e.cEnd; x.nucLen = e.nucLen; }
                }
            }
            it++;
        }
        i = j;
    }
    *nExonsOut = nEx;
    return (int)nCalls;
}

}  // extern "C"

// =====================================================================================================
// extractorfs + translatenucs: six-frame ORF fragments of a contig and their translation.
//   Orf::setSequence / findAll / findForward      reference/lib/mmseqs/src/commons/Orf.cpp:125-157,171-190,230-349
//   extractorfs driver, header coordinates        reference/lib/mmseqs/src/util/extractorfs.cpp:70-100, Orf.cpp:440-457
//   TranslateNucl (NCBI state machine)            reference/lib/mmseqs/src/commons/TranslateNucl.h:334-503
//   translatenucs driver                          reference/lib/mmseqs/src/util/translatenucs.cpp:47-104
// =====================================================================================================
static const char *IUPAC_BASES = "-ACMGRSVTWYHKDBN";   // 4-bit code = bit set of {A=1, C=2, G=4, T=8}
static int base_code(unsigned char c) {                  // TranslateNucl::initConversionTable (TranslateNucl.h:334-362)
    if (c < 16) return c;                                // ncbi4na values map to themselves
    const char *u = strchr(IUPAC_BASES, toupper(c));
    if (c == 'U' || c == 'u') return 8;
    if (c == 'X' || c == 'x') return 15;
    if (c == '-' ) return 0;
    return (u && *u) ? (int)(u - IUPAC_BASES) : 0;       // illegal characters -> gap
}
static char iupac_complement(char c) {                   // Orf::iupacReverseComplementTable (Orf.cpp:47-51); '.' = no complement
    static const char *from = "ATUGCYRSWKMBDHVN", *to = "TAACGRYSWMKVHDBN";
    const char *p = strchr(from, toupper((unsigned char)c));
    if (!p || !*p || !isalpha((unsigned char)c)) return '.';
    return islower((unsigned char)c) ? (char)tolower(to[p - from]) : to[p - from];
}
// amino acid of every (possibly ambiguous) codon: consensus over the expansions, with B/Z/J merging (TranslateNucl.h:383-485)
// NCBI genetic codes (amino acid per codon, TCAG order) as TranslateNucl's constructor lists them (TranslateNucl.h:110-246); NULL = not a table id
static const char *genetic_code(int table) {
    switch (table) {
    case 1: case 11: return "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 2:  return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG";
    case 3:  return "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 4:  return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 5:  return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG";
    case 6:  return "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 9:  return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG";
    case 10: return "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 12: return "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 13: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG";
    case 14: return "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG";
    case 15: return "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 16: return "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 21: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG";
    case 22: return "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 23: return "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 24: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG";
    case 25: return "FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 26: return "FFLLSSSSYY**CC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 27: case 28: return "FFLLSSSSYYQQCCWWLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 29: return "FFLLSSSSYYYYCC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 30: return "FFLLSSSSYYEECC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    case 31: return "FFLLSSSSYYEECCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    default: return nullptr;
    }
}
static void build_codon_table(char aaOf[4096], const char *code) {
    const int exp[4] = {1, 2, 4, 8}; const int tcag[9] = {0, 2, 1, 0, 3, 0, 0, 0, 0};
    for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) for (int k = 0; k < 16; k++) {
        char aa = 0; bool go = true;
        for (int p = 0; p < 4 && go; p++) if (exp[p] & i) for (int q = 0; q < 4 && go; q++) if (exp[q] & j) for (int r = 0; r < 4 && go; r++) if (exp[r] & k) {
            const char ch = code[16 * tcag[exp[p]] + 4 * tcag[exp[q]] + tcag[exp[r]]];
            if (!aa) aa = ch;
            else if (aa != ch) {
                if ((aa == 'B' || aa == 'D' || aa == 'N') && (ch == 'D' || ch == 'N')) aa = 'B';
                else if ((aa == 'Z' || aa == 'E' || aa == 'Q') && (ch == 'E' || ch == 'Q')) aa = 'Z';
                else if ((aa == 'J' || aa == 'I' || aa == 'L') && (ch == 'I' || ch == 'L')) aa = 'J';
                else { aa = 'X'; go = false; }
            }
        }
        aaOf[256 * i + 16 * j + k] = aa ? aa : 'X';
    }
}
struct mbo_orf { uint32_t contig; uint32_t from, to; uint32_t flags; uint64_t aaOff; uint32_t aaLen; uint32_t pad; };   // from/to = header coordinates; flags: 1 incomplete start, 2 incomplete end, 4 minus strand

static void find_forward(const std::string &seq, int strand, size_t minLen, size_t maxLen, size_t maxGaps, int startMode, const char *code, std::vector<std::array<size_t, 4>> &res) {
    const size_t L = seq.size();
    // stop codons of the table = the '*' entries of its amino-acid string (TranslateNucl.h:428-434 -> Orf::Orf, Orf.cpp:54-64), compared letter by letter with the
    // upper-cased sequence (Orf.cpp:266-269, isInCodons :206-228): ambiguity codes never match
    auto tcag = [](int c) { return c == 'T' ? 0 : c == 'C' ? 1 : c == 'A' ? 2 : c == 'G' ? 3 : -1; };
    auto at = [&](size_t p) -> int { return p < L ? (unsigned char)seq[p] : 127; };   // CHAR_MAX padding (Orf.cpp:151-154)
    auto up = [&](int c) { return c == 127 ? 127 : (c & ~0x20); };
    auto isStop = [&](int a, int b, int c) { const int x = tcag(a), y = tcag(b), z = tcag(c); return x >= 0 && y >= 0 && z >= 0 && code[16 * x + 4 * y + z] == '*'; };
    auto isStart = [&](int a, int b, int c) { return a == 'A' && b == 'T' && c == 'G'; };
    auto gapOrN = [&](int c) { return c == 'N' || c == 127 || iupac_complement((char)c) == '.'; };
    bool inside[3] = {true, true, true}, hasStart[3] = {false, false, false};
    size_t gaps[3] = {0, 0, 0}, len[3] = {0, 0, 0}, from[3] = {0, 1, 2};
    for (size_t i = 0; i + 2 < L; i += 3) for (size_t pos = i; pos < i + 3; pos++) {      // Orf.cpp:263-264
        const int c0 = up(at(pos)), c1 = up(at(pos + 1)), c2 = up(at(pos + 2));
        const size_t f = pos % 3;
        const bool incomplete = c0 == 127 || c1 == 127 || c2 == 127;
        const bool isLast = !incomplete && (at(pos + 3) == 127 || at(pos + 4) == 127 || at(pos + 5) == 127);
        bool shouldStart;
        if (startMode == 0) shouldStart = !inside[f] && isStart(c0, c1, c2);
        else if (startMode == 1) shouldStart = !inside[f];
        else shouldStart = isStart(c0, c1, c2);
        if (shouldStart) { inside[f] = true; hasStart[f] = true; from[f] = pos; gaps[f] = 0; len[f] = 0; }
        const bool stop = isStop(c0, c1, c2);
        if (inside[f]) { if (!stop) len[f]++; if (gapOrN(c0) || gapOrN(c1) || gapOrN(c2)) gaps[f]++; }
        if (inside[f] && (stop || isLast)) {
            inside[f] = false;
            if (len[f] == 0 && stop) continue;
            const size_t to = (isLast && !stop) ? pos + 2 : pos - 1;
            if (gaps[f] > maxGaps || len[f] > maxLen || len[f] < minLen) continue;
            res.push_back({from[f], to, (size_t)((hasStart[f] ? 0 : 1) | (stop ? 0 : 2)), (size_t)strand});
        }
    }
}

extern "C" {
// Contigs: concatenated ASCII nucleotides.  Output in the reference's key order: contigs ascending, plus-strand ORFs in emission order,
// then minus-strand ORFs (DBWriter::createRenumberedDB sorts by (contig key, data offset), DBWriter.cpp:690-738).  aa = ASCII residues.
// Returns the number of ORFs or -1 if a capacity is exceeded.
long long mbo_extract_orfs_tt(const char *nuc, const uint64_t *off, uint32_t nContigs, int minLen, int maxLen, int startMode, int table,
                              mbo_orf *recs, uint64_t capRecs, char *aa, uint64_t capAa, uint64_t *aaUsed, int threads) {
    const char *code = genetic_code(table); if (!code) return -2;
    char aaOf[4096]; build_codon_table(aaOf, code);
    std::vector<std::vector<mbo_orf>> perRec(nContigs); std::vector<std::string> perAa(nContigs);
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (uint32_t c = 0; c < nContigs; c++) {
        const size_t L = off[c + 1] - off[c];
        if (L < 3) continue;                                                              // Orf::setSequence (Orf.cpp:126-128)
        std::string fwd(nuc + off[c], L), rev(L, 'N');
        for (size_t i = 0; i < L; i++) if (fwd[i] == 'u') fwd[i] = 't';                  // (the 'U' branch is overwritten, Orf.cpp:141-144)
        for (size_t i = 0; i < L; i++) { char x = iupac_complement(fwd[L - 1 - i]); rev[i] = (x == '.') ? 'N' : x; }
        std::vector<std::array<size_t, 4>> res;
        find_forward(fwd, +1, minLen, maxLen, (size_t)-1, startMode, code, res);
        find_forward(rev, -1, minLen, maxLen, (size_t)-1, startMode, code, res);
        for (auto &o : res) {
            const std::string &s = (int)o[3] > 0 ? fwd : rev;
            size_t nucLen = o[1] - o[0] + 1;
            if (nucLen < 3) continue;                                                     // translatenucs.cpp:75-78
            mbo_orf r; r.contig = c; r.flags = (uint32_t)o[2] | ((int)o[3] < 0 ? 4u : 0u); r.pad = 0;
            r.from = (uint32_t)((int)o[3] > 0 ? o[0] : (L - 1) - o[0]); r.to = (uint32_t)((int)o[3] > 0 ? o[1] : (L - 1) - o[1]);   // extractorfs.cpp:82-87
            r.aaOff = perAa[c].size(); r.aaLen = (uint32_t)(nucLen / 3);
            for (size_t k = 0; k + 2 < nucLen; k += 3) {                                  // TranslateNucl::translate (TranslateNucl.h:488-503)
                const unsigned char a = s[o[0] + k], b = s[o[0] + k + 1], d = s[o[0] + k + 2];
                char res1 = aaOf[256 * base_code(a) + 16 * base_code(b) + base_code(d)];
                if (islower(a) || islower(b) || islower(d)) res1 = (char)tolower(res1);
                perAa[c].push_back(res1);
            }
            perRec[c].push_back(r);
        }
    }
    uint64_t n = 0, na = 0;
    for (uint32_t c = 0; c < nContigs; c++) {
        if (n + perRec[c].size() > capRecs || na + perAa[c].size() > capAa) return -1;
        for (auto r : perRec[c]) { r.aaOff += na; recs[n++] = r; }
        memcpy(aa + na, perAa[c].data(), perAa[c].size()); na += perAa[c].size();
    }
    *aaUsed = na;
    return (long long)n;
}
long long mbo_extract_orfs_mt(const char *nuc, const uint64_t *off, uint32_t nContigs, int minLen, int maxLen, int startMode,
                              mbo_orf *recs, uint64_t capRecs, char *aa, uint64_t capAa, uint64_t *aaUsed, int threads) {
    return mbo_extract_orfs_tt(nuc, off, nContigs, minLen, maxLen, startMode, 1, recs, capRecs, aa, capAa, aaUsed, threads);
}
long long mbo_extract_orfs(const char *nuc, const uint64_t *off, uint32_t nContigs, int minLen, int maxLen, int startMode,
                           mbo_orf *recs, uint64_t capRecs, char *aa, uint64_t capAa, uint64_t *aaUsed) {
    return mbo_extract_orfs_mt(nuc, off, nContigs, minLen, maxLen, startMode, recs, capRecs, aa, capAa, aaUsed, 1);
}
}  // extern "C"

// =====================================================================================================
// reduceredundancy: greedy clustering of the calls of one contig + same-strand overlap exclusion
//   clusterPredictions / excludeSameStrandOverlaps   reference/src/exonpredictor/reduceredundancy.cpp:24-109,111-147
//   comparators                                      reference/src/commons/PredictionParser.h:318-366
//   driver (per contig, plus then minus, final order) reference/src/exonpredictor/reduceredundancy.cpp:273-296
// =====================================================================================================
struct mbo_pred { uint32_t targetKey; int32_t strand; uint32_t totalBitscore, nExons, lowCoord, highCoord, exonBegin, pad; double evalue; };
struct mbo_pred_out { uint32_t clusterId, clusterLowCoord, noOverlapClusterId, isRep; };

static void rr_strand(const mbo_pred *P, const uint32_t *exonKeys, std::vector<uint32_t> &idx, mbo_pred_out *out, std::vector<uint32_t> &reps) {
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {                // comparePredictionsByContigStart
        if (P[a].lowCoord != P[b].lowCoord) return P[a].lowCoord < P[b].lowCoord;
        if (P[a].highCoord != P[b].highCoord) return P[a].highCoord > P[b].highCoord;
        if (P[a].totalBitscore != P[b].totalBitscore) return P[a].totalBitscore > P[b].totalBitscore;
        return P[a].targetKey < P[b].targetKey; });
    std::vector<char> clustered(idx.size(), 0);
    for (size_t i = 0; i < idx.size(); i++) {
        if (clustered[i]) continue;
        const mbo_pred &pi = P[idx[i]];
        uint32_t finalId = pi.targetKey, finalLow = pi.lowCoord, maxScore = pi.totalBitscore;
        clustered[i] = 1;
        std::vector<size_t> members(1, i);
        for (size_t j = i + 1; j < idx.size(); j++) {
            const mbo_pred &pj = P[idx[j]];
            if (pj.lowCoord >= pi.highCoord) break;                                       // reduceredundancy.cpp:52-55
            bool share = false;
            for (uint32_t a = 0; a < pi.nExons && !share; a++) for (uint32_t b = 0; b < pj.nExons; b++)
                if (exonKeys[pi.exonBegin + a] == exonKeys[pj.exonBegin + b]) { share = true; break; }
            if (share && !clustered[j]) {
                clustered[j] = 1;
                if (pj.totalBitscore > maxScore) { maxScore = pj.totalBitscore; finalId = pj.targetKey; finalLow = pj.lowCoord; }
                members.push_back(j);
            }
        }
        for (size_t m : members) {
            mbo_pred_out &o = out[idx[m]]; o.clusterId = finalId; o.clusterLowCoord = finalLow;
            if (P[idx[m]].targetKey == finalId && P[idx[m]].lowCoord == finalLow) { o.isRep = 1; reps.push_back(idx[m]); }   // :93-97
        }
    }
}
static void rr_overlaps(const mbo_pred *P, std::vector<uint32_t> &reps, mbo_pred_out *out) {
    std::stable_sort(reps.begin(), reps.end(), [&](uint32_t a, uint32_t b) {              // comparePredictionsByEvalue
        if (P[a].evalue != P[b].evalue) return P[a].evalue < P[b].evalue;
        return P[a].lowCoord < P[b].lowCoord; });
    std::vector<char> done(reps.size(), 0);
    for (size_t i = 0; i < reps.size(); i++) {
        if (done[i]) continue;
        done[i] = 1; const mbo_pred &a = P[reps[i]]; out[reps[i]].noOverlapClusterId = a.targetKey;
        for (size_t j = i + 1; j < reps.size(); j++) {
            const mbo_pred &b = P[reps[j]];
            if ((b.highCoord < a.highCoord && b.highCoord > a.lowCoord) || (b.lowCoord < a.highCoord && b.lowCoord > a.lowCoord) ||
                (b.highCoord < a.highCoord && b.lowCoord > a.lowCoord) || (b.highCoord > a.highCoord && b.lowCoord < a.lowCoord)) {
                done[j] = 1; out[reps[j]].noOverlapClusterId = a.targetKey;               // :139-141 (assigned even if already assigned)
            }
        }
    }
}
extern "C" {
// one contig: preds in the order of the calledExons entry.  memberOrder = every call, plus strand (sorted by contig start) then minus;
// repOrder = the representatives in output order (stable by (target, strand) over [plus reps by E-value, minus reps by E-value]).
int mbo_reduce_redundancy(const mbo_pred *P, uint32_t n, const uint32_t *exonKeys, mbo_pred_out *out, uint32_t *memberOrder, uint32_t *repOrder, uint32_t *nRep) {
    for (uint32_t i = 0; i < n; i++) { out[i].clusterId = 0; out[i].clusterLowCoord = 0; out[i].noOverlapClusterId = 0; out[i].isRep = 0; }
    std::vector<uint32_t> plus, minus, rp, rm;
    for (uint32_t i = 0; i < n; i++) (P[i].strand == 1 ? plus : minus).push_back(i);
    rr_strand(P, exonKeys, plus, out, rp); rr_overlaps(P, rp, out);
    rr_strand(P, exonKeys, minus, out, rm); rr_overlaps(P, rm, out);
    uint32_t k = 0;
    for (uint32_t x : plus) memberOrder[k++] = x;
    for (uint32_t x : minus) memberOrder[k++] = x;
    std::vector<uint32_t> reps(rp); reps.insert(reps.end(), rm.begin(), rm.end());
    std::stable_sort(reps.begin(), reps.end(), [&](uint32_t a, uint32_t b) {              // comparePredictionsByTarget
        if (P[a].targetKey != P[b].targetKey) return P[a].targetKey < P[b].targetKey;
        return P[a].strand < P[b].strand; });
    for (size_t i = 0; i < reps.size(); i++) repOrder[i] = reps[i];
    *nRep = (uint32_t)reps.size();
    return 0;
}
// batch form (same layout as mb200_reduce_redundancy: exon ranges are absolute, orders are relative to the contig's range); OpenMP over contigs
int mbo_reduce_redundancy_batch(const mbo_pred *P, const uint64_t *off, uint32_t nContigs, const uint32_t *exonKeys, mbo_pred_out *out, uint32_t *memberOrder,
                                uint32_t *repOrder, uint32_t *repCount, int threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (uint32_t c = 0; c < nContigs; c++)
        mbo_reduce_redundancy(P + off[c], (uint32_t)(off[c + 1] - off[c]), exonKeys, out + off[c], memberOrder + off[c], repOrder + off[c], repCount + c);
    return 0;
}
}  // extern "C"

// =============================================================================================
//  a3 (masking)  Masker::maskSequence with the prefilter's flags (--mask 1 --mask-prob 0.9 --mask-lower-case 0 --mask-n-repeat 0)
//     M/src/commons/Masker.cpp:14-58,121-128  ->  tantan::maskSequences  M/lib/tantan/tantan.cpp:478-523
//     (repeat HMM without gap states: maxCycleLength 50, repeatProb 0.005, repeatEndProb 0.05, repeatOffsetProbDecay 0.9; forward-backward in
//      double with a rescale every 16 letters, calcRepeatProbs :401-433); likelihood ratios from the SEED matrix (IndexBuilder.cpp:84-86):
//      ProbabilityMatrix, BaseMatrix.h:85-96 = P(a,b) / (pBack[a] * pBack[b]).  A letter is masked (-> X) when P(repeat) >= maskProb.
//  The reference's inner loops run over 4 (AVX2) or 2 (SSE4.1/NEON) doubles at a time; `lanes` restates their summation order
//  (tantan.cpp:300-335,347-382, mcf_simd.h:175-179).
// =============================================================================================
namespace {
void tantan_lr(double lr[21][21]) {
    double pBack[21];
    for (int i = 0; i < AL; i++) pBack[i] = MB200_VTML80_BACKGROUND[i];
    bool xIsPositive = false;
    for (int j = 0; j < AL; j++) if (MB200_VTML80_HALFBITS[XRES][j] > 0 || MB200_VTML80_HALFBITS[j][XRES] > 0) { xIsPositive = true; break; }
    if (!xIsPositive) for (int i = 0; i < AL - 1; i++) pBack[i] = pBack[i] * (1.0 - pBack[XRES]);
    for (int i = 0; i < AL; i++) for (int j = 0; j < AL; j++) {
        const double prob = std::exp(MB200_VTML80_LAMBDA * MB200_VTML80_HALFBITS[i][j]) * pBack[i] * pBack[j];   // SubstitutionMatrix.cpp:410-416
        lr[i][j] = prob / (pBack[i] * pBack[j]);                                                                  // BaseMatrix.h:93
    }
}
double lane_sum(const double *s, int lanes) { return lanes == 4 ? (s[0] + s[2]) + (s[1] + s[3]) : s[0] + s[1]; }   // simdHorizontalAddDbl

int tantan_mask(uint8_t *seq, int n, const double lr[21][21], double minMaskProb, int lanes) {
    const int W = 50; const double repeatProb = 0.005, repeatEndProb = 0.05, decay = 0.9;
    const double b2b = 1 - repeatProb, f2b = repeatEndProb, f2f0 = 1 - repeatEndProb;
    const double b2fFirst = repeatProb * ((1 - decay) / (1 - std::pow(decay, W)));       // firstRepeatOffsetProb, tantan.cpp:26-31
    double b2f[W], fg[W];
    { double p = b2fFirst; for (int i = 0; i < W; i++) { b2f[i] = p; p *= decay; } }
    std::vector<double> scale(n / 16 + 1, 1.0);
    std::vector<float> prob(n > 0 ? n : 1);
    double bg = 1.0;
    for (int i = 0; i < W; i++) fg[i] = 0.0;
    for (int pos = 0; pos < n; pos++) {                                                  // forward, tantan.cpp:404-410
        const double *row = lr[seq[pos]]; const int maxOffset = pos < W ? pos : W; const double b = bg;
        double s[4] = {0, 0, 0, 0}; int i = 0;
        for (; i <= maxOffset - lanes; i += lanes)
            for (int j = 0; j < lanes; j++) { const double f = fg[i + j]; s[j] += f; fg[i + j] = (b * b2f[i + j] + f * f2f0) * row[seq[pos - i - j - 1]]; }
        double fromFg = lane_sum(s, lanes);
        for (; i < maxOffset; i++) { const double f = fg[i]; fromFg += f; fg[i] = (b * b2f[i] + f * f2f0) * row[seq[pos - i - 1]]; }
        bg = b * b2b + fromFg * f2b;
        if (pos % 16 == 15) { const double sc = 1 / bg; scale[pos / 16] = sc; bg *= sc; for (int k = 0; k < W; k++) fg[k] *= sc; }
        prob[pos] = static_cast<float>(bg);
    }
    double acc = 0.0; for (int k = 0; k < W; k++) acc += fg[k];
    const double z = bg * b2b + acc * f2b;                                               // forwardTotal
    bg = b2b; for (int k = 0; k < W; k++) fg[k] = f2b;                                   // initializeBackwardAlgorithm
    for (int pos = n - 1; pos >= 0; pos--) {                                             // backward, :416-426
        const double nonRepeat = prob[pos] * bg / z;
        prob[pos] = 1 - static_cast<float>(nonRepeat);
        if (pos % 16 == 15) { const double sc = scale[pos / 16]; bg *= sc; for (int k = 0; k < W; k++) fg[k] *= sc; }
        const double *row = lr[seq[pos]]; const int maxOffset = pos < W ? pos : W; const double toBg = f2b * bg;
        double s[4] = {0, 0, 0, 0}; int i = 0;
        for (; i <= maxOffset - lanes; i += lanes)
            for (int j = 0; j < lanes; j++) { const double f = fg[i + j] * row[seq[pos - i - j - 1]]; s[j] += b2f[i + j] * f; fg[i + j] = toBg + f2f0 * f; }
        double toFg = lane_sum(s, lanes);
        for (; i < maxOffset; i++) { const double f = fg[i] * row[seq[pos - i - 1]]; toFg += b2f[i] * f; fg[i] = toBg + f2f0 * f; }
        bg = b2b * bg + toFg;
    }
    int masked = 0;
    for (int pos = 0; pos < n; pos++) if (prob[pos] >= minMaskProb) { seq[pos] = XRES; masked++; }   // maskProbableLetters :525-540; finalizeMasking keeps X
    return masked;
}
}  // namespace

extern "C" {
// out = residues after masking; returns the number of masked residues (the reference's "Masked residues:" line) or -1
// every switch of Masker::maskSequence (Masker.cpp:14-58): tantan, then maskRepeats (:83-116) on the codes tantan left, then lower-case letters
long long mbo_mask_targets_ex(const uint8_t *res, const uint64_t *off, uint32_t nT, int tantan, double maskProb, const uint8_t *lowerCase, int maskNrepeating, uint8_t *out, int lanes) {
    if (lanes != 2 && lanes != 4) return -1;
    double lr[21][21]; tantan_lr(lr);
    long long total = 0;
    for (uint32_t t = 0; t < nT; t++) {
        const uint64_t b = off[t]; const unsigned int seqLen = (unsigned int)(off[t + 1] - b);
        unsigned char *numSequence = out + b;
        for (unsigned int i = 0; i < seqLen; i++) numSequence[i] = res[b + i];
        if (tantan) total += tantan_mask(numSequence, (int)seqLen, lr, maskProb, lanes);
        if (maskNrepeating > 0) {                                   // Masker::maskRepeats, statement by statement (previousChar '\0' is the code of 'A')
            unsigned int repeatCount = 0; int startOfRepeat = -1; char previousChar = '\0';
            for (unsigned int pos = 0; pos < seqLen; ++pos) {
                char currentChar = (char)numSequence[pos];
                if (currentChar == previousChar) repeatCount++;
                else {
                    if (repeatCount > (unsigned int)maskNrepeating) for (unsigned int i = (unsigned int)startOfRepeat; i < pos; ++i) { numSequence[i] = XRES; total++; }
                    repeatCount = 1; startOfRepeat = (int)pos; previousChar = currentChar;
                }
            }
            if (repeatCount > (unsigned int)maskNrepeating) for (unsigned int i = (unsigned int)startOfRepeat; i < seqLen; ++i) { numSequence[i] = XRES; total++; }
        }
        if (lowerCase) for (unsigned int i = 0; i < seqLen; i++) if (lowerCase[b + i]) { numSequence[i] = XRES; total++; }
    }
    return total;
}
long long mbo_mask_targets(const uint8_t *res, const uint64_t *off, uint32_t nT, double maskProb, uint8_t *out, int lanes, int threads) {
    if (lanes != 2 && lanes != 4) return -1;
    double lr[21][21]; tantan_lr(lr);
    long long total = 0;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : total) num_threads(threads)
    for (uint32_t t = 0; t < nT; t++) {
        const uint64_t b = off[t]; const int n = (int)(off[t + 1] - b);
        for (int i = 0; i < n; i++) out[b + i] = res[b + i];
        total += tantan_mask(out + b, n, lr, maskProb, lanes);
    }
    return total;
}
}
