// host_tables.h -- host-side preprocessing for the B200 hot path (plain C++, no CUDA):
// substitution matrices, 3-mer similarity tables, target k-mer index, e-value model.
// These are the one-time "setup" steps the reference also does on the host before its query loop
// (SURVEY.md section 8a rows a1-a3, a14); the per-query work is in the CUDA kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace mb200 {

const int AL = 21;
const int N3 = 8000;
const size_t KMER_TABLE = 64000000;   // 20^6

struct Matrices {
    int16_t seed[441];   // VTML80  x8  bias -0.2  (k-mer similarity)
    int16_t ung[441];    // BLOSUM62 x2 bias -0.2  (ungapped diagonal score)
    int16_t aln[441];    // BLOSUM62 x2 bias  0.0  (gapped alignment)
    double seedPBack[21];
    double alnPBack[21];
};
// SubstitutionMatrix ctor semantics (reference/lib/mmseqs/src/commons/SubstitutionMatrix.cpp:12-57, BaseMatrix.cpp:110-159)
void default_matrices(Matrices &m);

// sim3[r][z] = (neighbour 3-mer index << 16) | (uint16_t)score, z-th best neighbour of 3-mer r in the reference's
// order (ExtendedSubstitutionMatrix::calcScoreMatrix, reference/lib/mmseqs/src/prefiltering/ExtendedSubstitutionMatrix.cpp:20-69);
// cnt3[r][x - cntMin] = number of neighbours of r scoring >= x.
void build_sim3(const int16_t seed[441], std::vector<uint32_t> &sim3, std::vector<uint16_t> &cnt3, int cntMin, int cntW);

struct TargetIndex {
    std::vector<uint32_t> offsets;          // [20^6 + 1]
    std::vector<uint64_t> entries;          // low 32 bits seqId, next 16 bits pos_j  (device layout uint2)
};
// IndexBuilder::fillDatabase semantics on ALREADY MASKED residues
// (reference/lib/mmseqs/src/prefiltering/IndexBuilder.cpp:55-239, IndexTable.h:133-173,348-401,182-189)
int build_index(const uint8_t *masked, const uint64_t *off, uint32_t nT, const int16_t seed[441], int kmerThr, TargetIndex &idx);
// same layout from a reference IndexTable (offsets size_t[20^6+1], packed 6-byte IndexEntryLocal entries)
int import_index(const uint64_t *offsets, const uint8_t *entries6, uint64_t nEntries, TargetIndex &idx);

// Repeat / low-complexity masking of the targets before indexing = Masker::maskSequence with the prefilter's flags (--mask 1, --mask-lower-case 0,
// --mask-n-repeat 0; reference/lib/mmseqs/src/commons/Masker.cpp:14-58 -> tantan::maskSequences, reference/lib/mmseqs/lib/tantan/tantan.cpp:478-540):
// posterior repeat probability of every letter under tantan's gap-free repeat HMM (window 50, repeatProb 0.005, repeatEndProb 0.05, offset decay 0.9),
// likelihood ratios from the SEED matrix (ProbabilityMatrix, BaseMatrix.h:85-96); letters with P(repeat) >= maskProb become X (code 20).
// res / out: numeric residue codes, off[nT + 1].  simdWidth = number of doubles the reference build's tantan loops add side by side (4 = AVX2 build,
// 2 = SSE4.1 / NEON build): it fixes the summation order, so the double-precision sums -- and the result -- are those of that build.  Returns #masked.
uint64_t mask_targets(const uint8_t *res, const uint64_t *off, uint32_t nT, double maskProb, uint8_t *out, int simdWidth = 4);
// all switches of Masker::maskSequence: tantan on/off (--mask), runs of more than maskNRepeat equal letters (--mask-n-repeat, 0 = off), letters that were
// lower case in the DB (--mask-lower-case; lowerCase[i] != 0 per residue, or NULL)
uint64_t mask_targets_ex(const uint8_t *res, const uint64_t *off, uint32_t nT, bool tantan, double maskProb, int simdWidth, const uint8_t *lowerCase, int maskNRepeat, uint8_t *out);

// Gumbel/finite-size e-value model of the reference for BLOSUM62 gap 11/1
// (reference/lib/mmseqs/src/alignment/EvalueComputation.h:18-40,56-69 -> Sls::AlignmentEvaluer::area)
struct EvalueModel {
    double lambda, K, a1, b1, a2, b2, alpha1, beta1, alpha2, beta2, sigma, tau, viThr, vjThr, cThr, logK;
    EvalueModel();
    double evalue(double score, double qLen, double dbResidues) const;
    int bitScore(double score) const;   // int(bits + 0.5)  (Matcher.cpp:130)
    // smallest raw score whose e-value is <= evalThr for this query length (e-value is decreasing in score)
    int minScoreFor(double evalThr, double qLen, double dbResidues) const;
};

// NCBI genetic code `table` as the amino acid of each codon in TCAG order ('*' = stop), NULL if the reference does not know the id
// (TranslateNucl's constructor, TranslateNucl.h:110-246: 1-6, 9-16, 21-31).  The ORF scan takes its stop codons from it, the translation its residues.
const char *genetic_code(int table);

// Prefiltering::getKmerThreshold for a sequence search with k = 6 (Prefiltering.cpp:1049-1055,1063): what `-s` means for the k-mer similarity threshold
int kmer_threshold_for_sensitivity(float sensitivity);

}  // namespace mb200
