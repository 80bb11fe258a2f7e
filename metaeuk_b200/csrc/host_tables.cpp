// host_tables.cpp -- see host_tables.h.  Plain C++17 + OpenMP; compiled into libmetaeuk_b200.so by nvcc's host
// compiler and (for CPU-only logic tests) directly by g++.
#include "host_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <parallel/algorithm>

#include "submat_data.h"

namespace mb200 {

// ---- matrices -------------------------------------------------------------------------------------
// Scores are round-half-away( bitFactor * log2( p_ab / (p_a p_b) ) + bias ) where the joint probabilities are rebuilt
// from the shipped half-bit table: p_ab = exp(lambda * S_ab) * f_a * f_b  (SubstitutionMatrix.cpp:404-416), the
// marginals p_a are the row sums of p_ab with X pinned to 1e-5 (BaseMatrix.cpp:96-108), and the background f used by
// the composition-bias code is the shipped background scaled by (1 - f_X) (SubstitutionMatrix.cpp:404-408).
static void make_matrix(const double S[21][21], const double f0[21], double lambda, float bitFactor, float bias, int16_t *out, double *fOut) {
    double f[21];
    bool xPositive = false;
    for (int a = 0; a < AL; a++) { f[a] = f0[a]; if (S[AL - 1][a] > 0 || S[a][AL - 1] > 0) xPositive = true; }
    if (!xPositive) for (int a = 0; a < AL - 1; a++) f[a] = f[a] * (1.0 - f[AL - 1]);
    std::vector<double> joint(AL * AL), marg(AL, 0.0);
    for (int a = 0; a < AL; a++) for (int b = 0; b < AL; b++) joint[a * AL + b] = std::exp(lambda * S[a][b]) * f[a] * f[b];
    for (int a = 0; a < AL; a++) { double s = 0; for (int b = 0; b < AL; b++) s += joint[a * AL + b]; marg[a] = s; }
    marg[AL - 1] = 1E-5;
    const double bf = bitFactor, sb = bias;   // float arguments widened, as in the reference's constructor chain
    for (int a = 0; a < AL; a++) for (int b = 0; b < AL; b++) {
        const double bits = std::log2(joint[a * AL + b] / (marg[a] * marg[b]));
        const double v = bf * bits + sb;
        out[a * AL + b] = (short)((v < 0.0) ? v - 0.5 : v + 0.5);
    }
    if (fOut) for (int a = 0; a < AL; a++) fOut[a] = f[a];
}

void default_matrices(Matrices &m) {
    make_matrix(MB200_VTML80_HALFBITS, MB200_VTML80_BACKGROUND, MB200_VTML80_LAMBDA, 8.0f, -0.2f, m.seed, m.seedPBack);   // Prefiltering.cpp:68,990
    make_matrix(MB200_BLOSUM62_HALFBITS, MB200_BLOSUM62_BACKGROUND, MB200_BLOSUM62_LAMBDA, 2.0f, -0.2f, m.ung, nullptr);   // Prefiltering.cpp:69
    make_matrix(MB200_BLOSUM62_HALFBITS, MB200_BLOSUM62_BACKGROUND, MB200_BLOSUM62_LAMBDA, 2.0f, 0.0f, m.aln, m.alnPBack); // Alignment.cpp:152
}

// ---- target masking (tantan) --------------------------------------------------------------------------
namespace {
// The HMM: one background state and W = 50 repeat states (state k copies the letter k + 1 positions back).  Per letter the reference does
// transition + emission in one pass over the states, in double precision, rescaling all states by 1 / background every 16 letters.
class RepeatHmm {
public:
    static const int W = 50;
    explicit RepeatHmm(int simdWidth) : width(simdWidth) {
        // likelihood ratios P(a,b) / (f_a f_b), P rebuilt from the shipped half-bit table as in make_matrix
        double f[21]; bool xPositive = false;
        for (int a = 0; a < AL; a++) { f[a] = MB200_VTML80_BACKGROUND[a]; if (MB200_VTML80_HALFBITS[AL - 1][a] > 0 || MB200_VTML80_HALFBITS[a][AL - 1] > 0) xPositive = true; }
        if (!xPositive) for (int a = 0; a < AL - 1; a++) f[a] = f[a] * (1.0 - f[AL - 1]);
        for (int a = 0; a < AL; a++) for (int b = 0; b < AL; b++)
            ratio[a][b] = (std::exp(MB200_VTML80_LAMBDA * MB200_VTML80_HALFBITS[a][b]) * f[a] * f[b]) / (f[a] * f[b]);
        const double repeatProb = 0.005, decay = 0.9;
        stay = 1 - repeatProb; leave = 0.05; keep = 1 - leave;
        double p = repeatProb * ((1 - decay) / (1 - std::pow(decay, W)));
        for (int k = 0; k < W; k++) { enter[k] = p; p *= decay; }
    }
    // posterior P(letter i is in a repeat) for one sequence
    void posteriors(const uint8_t *x, int n, std::vector<float> &post, std::vector<double> &scales) const {
        post.resize(n); scales.assign(n / 16 + 1, 1.0);
        double rep[W], bg = 1.0;
        std::fill(rep, rep + W, 0.0);
        for (int i = 0; i < n; i++) {
            const double from = forward(x, i, rep, bg);
            bg = bg * stay + from * leave;
            if ((i & 15) == 15) { const double sc = 1 / bg; scales[i >> 4] = sc; bg *= sc; for (int k = 0; k < W; k++) rep[k] *= sc; }
            post[i] = static_cast<float>(bg);
        }
        double tail = 0.0; for (int k = 0; k < W; k++) tail += rep[k];
        const double total = bg * stay + tail * leave;
        bg = stay; std::fill(rep, rep + W, leave);
        for (int i = n - 1; i >= 0; i--) {
            const double notRepeat = post[i] * bg / total;
            post[i] = 1 - static_cast<float>(notRepeat);
            if ((i & 15) == 15) { const double sc = scales[i >> 4]; bg *= sc; for (int k = 0; k < W; k++) rep[k] *= sc; }
            const double to = backward(x, i, rep, leave * bg);
            bg = stay * bg + to;
        }
    }
private:
    int width; double ratio[21][21], enter[W], stay, leave, keep;
    double fold(const double *part) const { return width == 4 ? (part[0] + part[2]) + (part[1] + part[3]) : part[0] + part[1]; }
    // one letter of the forward pass over the reachable states; returns the probability mass that was in repeat states
    double forward(const uint8_t *x, int i, double *rep, double bg) const {
        const double *r = ratio[x[i]]; const int reach = i < W ? i : W;
        double part[4] = {0, 0, 0, 0}; int k = 0;
        for (; k + width <= reach; k += width)
            for (int j = 0; j < width; j++) { const double v = rep[k + j]; part[j] += v; rep[k + j] = (bg * enter[k + j] + v * keep) * r[x[i - 1 - k - j]]; }
        double sum = fold(part);
        for (; k < reach; k++) { const double v = rep[k]; sum += v; rep[k] = (bg * enter[k] + v * keep) * r[x[i - 1 - k]]; }
        return sum;
    }
    double backward(const uint8_t *x, int i, double *rep, double toBg) const {
        const double *r = ratio[x[i]]; const int reach = i < W ? i : W;
        double part[4] = {0, 0, 0, 0}; int k = 0;
        for (; k + width <= reach; k += width)
            for (int j = 0; j < width; j++) { const double v = rep[k + j] * r[x[i - 1 - k - j]]; part[j] += enter[k + j] * v; rep[k + j] = toBg + keep * v; }
        double sum = fold(part);
        for (; k < reach; k++) { const double v = rep[k] * r[x[i - 1 - k]]; sum += enter[k] * v; rep[k] = toBg + keep * v; }
        return sum;
    }
};
}  // namespace

uint64_t mask_targets(const uint8_t *res, const uint64_t *off, uint32_t nT, double maskProb, uint8_t *out, int simdWidth) {
    const RepeatHmm hmm(simdWidth == 2 ? 2 : 4);
    uint64_t masked = 0;
#pragma omp parallel reduction(+ : masked)
    {
        std::vector<float> post; std::vector<double> scales;
#pragma omp for schedule(dynamic, 64)
        for (uint32_t t = 0; t < nT; t++) {
            const uint64_t b = off[t]; const int n = (int)(off[t + 1] - b);
            hmm.posteriors(res + b, n, post, scales);
            for (int i = 0; i < n; i++) { const bool m = post[i] >= maskProb; out[b + i] = m ? (uint8_t)(AL - 1) : res[b + i]; masked += m; }
        }
    }
    return masked;
}

// Masker::maskSequence with all of its switches (Masker.cpp:14-58): tantan (--mask), then runs of more than `maskNRepeat` equal letters (--mask-n-repeat,
// Masker::maskRepeats :83-116, applied to the codes tantan left behind, so a run of X counts too), then letters that were lower case in the DB
// (--mask-lower-case).  The reference's run scan starts with previousChar = 0 and startOfRepeat = -1: code 0 is 'A', so a run of A at the very start of a
// sequence is never masked (its masking loop starts at (unsigned) -1) and is counted one letter longer than it is -- reproduced here.  Returns what the
// reference adds up as "Masked residues" (a letter masked by two rules counts twice).
uint64_t mask_targets_ex(const uint8_t *res, const uint64_t *off, uint32_t nT, bool tantan, double maskProb, int simdWidth, const uint8_t *lowerCase, int maskNRepeat, uint8_t *out) {
    const uint8_t X = (uint8_t)(AL - 1);
    uint64_t masked = tantan ? mask_targets(res, off, nT, maskProb, out, simdWidth) : 0;
    if (!tantan) memcpy(out + off[0], res + off[0], off[nT] - off[0]);
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : masked)
    for (uint32_t t = 0; t < nT; t++) {
        uint8_t *s = out + off[t]; const uint32_t n = (uint32_t)(off[t + 1] - off[t]);
        if (maskNRepeat > 0) {
            uint32_t count = 0; long start = -1; uint8_t prev = 0;
            for (uint32_t pos = 0; pos <= n; pos++) {
                if (pos < n && s[pos] == prev) { count++; continue; }
                if (count > (uint32_t)maskNRepeat && start >= 0) for (uint32_t i = (uint32_t)start; i < pos; i++) { s[i] = X; masked++; }
                if (pos < n) { count = 1; start = (long)pos; prev = s[pos]; }
            }
        }
        if (lowerCase) for (uint32_t i = 0; i < n; i++) if (lowerCase[off[t] + i]) { s[i] = X; masked++; }
    }
    return masked;
}

// ---- 3-mer similarity table -----------------------------------------------------------------------
// Stable descending order w.r.t. the cartesian enumeration (first letter outermost) == a counting sort over the
// (small) score range that visits neighbours in cartesian order.
void build_sim3(const int16_t seed[441], std::vector<uint32_t> &sim3, std::vector<uint16_t> &cnt3, int cntMin, int cntW) {
    sim3.assign((size_t)N3 * N3, 0);
    cnt3.assign((size_t)N3 * cntW, 0);
    const int lo = cntMin + 1, hi = cntMin + cntW - 1;   // every score must lie in [lo, hi]
#pragma omp parallel for schedule(static)
    for (int r = 0; r < N3; r++) {
        const int a0 = r % 20, a1 = (r / 20) % 20, a2 = r / 400;
        std::vector<int> hist(cntW + 1, 0);
        std::vector<int16_t> sc(N3);
        int z = 0;
        for (int b0 = 0; b0 < 20; b0++) for (int b1 = 0; b1 < 20; b1++) for (int b2 = 0; b2 < 20; b2++) {
            int s = seed[a0 * AL + b0] + seed[a1 * AL + b1] + seed[a2 * AL + b2];
            if (s < lo) s = lo; if (s > hi) s = hi;   // (never triggers for VTML80 x8; keeps the table in range for any matrix)
            sc[z++] = (int16_t)s; hist[s - cntMin]++;
        }
        // start[s] = number of neighbours with a strictly higher score
        std::vector<int> start(cntW + 1, 0);
        int run = 0;
        for (int x = cntW - 1; x >= 0; x--) { start[x] = run; run += hist[x]; cnt3[(size_t)r * cntW + x] = (uint16_t)run; }
        z = 0;
        uint32_t *row = &sim3[(size_t)r * N3];
        for (int b0 = 0; b0 < 20; b0++) for (int b1 = 0; b1 < 20; b1++) for (int b2 = 0; b2 < 20; b2++) {
            const int s = sc[z++];
            row[start[s - cntMin]++] = ((uint32_t)(b0 + 20 * b1 + 400 * b2) << 16) | (uint16_t)(int16_t)s;
        }
    }
}

// ---- target k-mer index ---------------------------------------------------------------------------
static const int SPACED[6] = {0, 1, 3, 5, 8, 9};   // 1101010011 (Sequence.h:23)

int build_index(const uint8_t *masked, const uint64_t *off, uint32_t nT, const int16_t seed[441], int kmerThr, TargetIndex &idx) {
    const uint64_t total = off[nT];
    if (total >= (1ull << 32)) return -1;
    int8_t self[AL];
    for (int a = 0; a < AL; a++) self[a] = (int8_t)seed[a * AL + a];
    // key = kmer << 32 | global residue position of the window start; sorting the keys orders every k-mer's occurrences
    // by (target, position) -- the order IndexTable::sortDBSeqLists establishes.
    std::vector<uint64_t> keys;
    {
        std::vector<uint64_t> cnt(nT + 1, 0);
#pragma omp parallel for schedule(dynamic, 256)
        for (uint32_t t = 0; t < nT; t++) {
            const uint8_t *s = masked + off[t]; const int64_t L = (int64_t)(off[t + 1] - off[t]); uint64_t c = 0;
            for (int64_t p = 0; p + 10 <= L; p++) {
                int sc = 0; bool x = false;
                for (int k = 0; k < 6; k++) { const uint8_t r = s[p + SPACED[k]]; x |= (r >= 20); sc += self[r < AL ? r : 20]; }
                if (!x && !(kmerThr > 0 && sc < kmerThr)) c++;
            }
            cnt[t + 1] = c;
        }
        for (uint32_t t = 0; t < nT; t++) cnt[t + 1] += cnt[t];
        keys.resize(cnt[nT]);
#pragma omp parallel for schedule(dynamic, 256)
        for (uint32_t t = 0; t < nT; t++) {
            const uint8_t *s = masked + off[t]; const int64_t L = (int64_t)(off[t + 1] - off[t]); uint64_t w = cnt[t];
            for (int64_t p = 0; p + 10 <= L; p++) {
                int sc = 0; bool x = false; uint64_t km = 0, pw = 1;
                for (int k = 0; k < 6; k++) { const uint8_t r = s[p + SPACED[k]]; x |= (r >= 20); sc += self[r < AL ? r : 20]; km += r * pw; pw *= 20; }
                if (!x && !(kmerThr > 0 && sc < kmerThr)) keys[w++] = (km << 32) | (off[t] + (uint64_t)p);
            }
        }
    }
    __gnu_parallel::sort(keys.begin(), keys.end());
    // keep the first position of a k-mer inside a target (IndexTable.h:390-399)
    idx.offsets.assign(KMER_TABLE + 1, 0);
    idx.entries.clear(); idx.entries.reserve(keys.size());
    std::vector<uint32_t> tid(keys.size());
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < keys.size(); i++) {
        const uint64_t g = keys[i] & 0xFFFFFFFFull;
        tid[i] = (uint32_t)(std::upper_bound(off, off + nT + 1, g) - off - 1);
    }
    uint64_t prevKmer = ~0ull; uint32_t prevT = ~0u;
    for (size_t i = 0; i < keys.size(); i++) {
        const uint64_t km = keys[i] >> 32, g = keys[i] & 0xFFFFFFFFull; const uint32_t t = tid[i];
        if (km == prevKmer && t == prevT) continue;
        prevKmer = km; prevT = t;
        idx.entries.push_back((uint64_t)t | ((g - off[t]) << 32));
        idx.offsets[km + 1]++;
    }
    for (size_t k = 0; k < KMER_TABLE; k++) idx.offsets[k + 1] += idx.offsets[k];
    return 0;
}

int import_index(const uint64_t *offsets, const uint8_t *entries6, uint64_t nEntries, TargetIndex &idx) {
    if (nEntries >= (1ull << 32)) return -1;
    idx.offsets.resize(KMER_TABLE + 1);
    for (size_t k = 0; k <= KMER_TABLE; k++) idx.offsets[k] = (uint32_t)offsets[k];
    idx.entries.resize(nEntries);
    for (uint64_t i = 0; i < nEntries; i++) { uint32_t s; uint16_t p; memcpy(&s, entries6 + 6 * i, 4); memcpy(&p, entries6 + 6 * i + 4, 2); idx.entries[i] = (uint64_t)s | ((uint64_t)p << 32); }
    return 0;
}

// ---- e-value model --------------------------------------------------------------------------------
// Constants: the reference's fixed BLOSUM62 11/1 parameter set (EvalueComputation.h:64-69).  Formulas: ALP's finite-size
// corrected area (reference/lib/mmseqs/lib/alp/sls_pvalues.cpp:366-532 with sls_basic.hpp:195-198 for the normal CDF).
EvalueModel::EvalueModel() {
    lambda = 0.27359865037097330642; K = 0.044620920658722244834;
    a1 = 1.5938724404943873658; b1 = -19.959867650284412122; a2 = 1.5938724404943873658; b2 = -19.959867650284412122;
    alpha1 = 30.455610143099914211; beta1 = -622.28684628915891608; alpha2 = 30.455610143099914211; beta2 = -622.28684628915891608;
    sigma = 29.602444874818868215; tau = -601.81087985041381216;
    viThr = std::max(2.0 * alpha2 / lambda, 0.0); vjThr = std::max(2.0 * alpha1 / lambda, 0.0); cThr = std::max(2.0 * sigma / lambda, 0.0);
    logK = std::log(K);
}
static inline double normal_cdf(double x) { return 0.5 * erfc(-sqrt(0.5) * x); }
double EvalueModel::evalue(double y, double qLen, double dbRes) const {
    const double pi = 3.1415926535897932384626433832795;
    const double cv = 1 / sqrt(2.0 * pi);
    // "I" dimension = database residues (seqlen2), "J" dimension = query length (seqlen1)  (sls_alignment_evaluer.cpp:1011-1016)
    const double mli = dbRes - (a2 * y + b2);
    const double svi = sqrt(std::max(viThr, alpha2 * y + beta2));
    const double mF = (svi == 0.0) ? 1e100 : mli / svi;
    const double PmF = normal_cdf(mF), EmF = -cv * exp(-0.5 * mF * mF);
    const double p1 = mli * PmF - svi * EmF;
    const double nlj = qLen - (a1 * y + b1);
    const double svj = sqrt(std::max(vjThr, alpha1 * y + beta1));
    const double nF = (svj == 0.0) ? 1e100 : nlj / svj;
    const double PnF = normal_cdf(nF), EnF = -cv * exp(-0.5 * nF * nF);
    const double p2 = nlj * PnF - svj * EnF;
    const double cy = std::max(cThr, sigma * y + tau);
    const double area = p1 * p2 + cy * (PmF * PnF);
    return (K * exp(-lambda * y)) * area;
}
int EvalueModel::bitScore(double score) const { return (int)((lambda * score - logK) / log(2.0) + 0.5); }
int EvalueModel::minScoreFor(double evalThr, double qLen, double dbRes) const {
    const int SMAX = 4096;
    int first = SMAX + 1;
    for (int s = 1; s <= SMAX; s++) {
        const bool ok = !(evalue(s, qLen, dbRes) > evalThr);
        if (ok && first > SMAX) first = s;
        if (!ok && first <= SMAX) return -1;   // not monotone: caller must fail loudly
    }
    return first;
}

// ---- genetic codes (TranslateNucl.h:110-246) and the sensitivity -> threshold rule (Prefiltering.cpp:1005-1064) -----------------------
const char *genetic_code(int table) {
    static const struct { int id; const char *aa; } codes[] = {
        {1,  "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"}, {2,  "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG"},
        {3,  "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG"}, {4,  "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
        {5,  "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG"}, {6,  "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
        {9,  "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG"}, {10, "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
        {11, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"}, {12, "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
        {13, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG"}, {14, "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
        {15, "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"}, {16, "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
        {21, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG"}, {22, "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
        {23, "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"}, {24, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG"},
        {25, "FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"}, {26, "FFLLSSSSYY**CC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
        {27, "FFLLSSSSYYQQCCWWLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"}, {28, "FFLLSSSSYYQQCCWWLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
        {29, "FFLLSSSSYYYYCC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"}, {30, "FFLLSSSSYYEECC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
        {31, "FFLLSSSSYYEECCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"}};
    for (const auto &c : codes) if (c.id == table) return c.aa;
    return nullptr;
}

int kmer_threshold_for_sensitivity(float sensitivity) {
    const float base = 163.2;                                  // float base; the product in double; the difference back in a float; truncation
    const float thr = base - (sensitivity * 8.917);
    return static_cast<int>(thr);
}

}  // namespace mb200
