// host_join.h -- the binary hand-off between the stages of predictexons, plain C++ (no CUDA).
//
// The reference passes text DBs from stage to stage: `align` prints 10 columns per accepted alignment, `resultspercontig`
// re-parses them, appends the ORF -> contig pseudo-record and regroups per contig, `collectoptimalset` re-parses the 20
// columns again (reference/src/exonpredictor/resultspercontig.cpp:64-196, collectoptimalset.cpp:273-300).  Here the accepted
// alignments (mb200_aln) and the fragment records (mb200_orf) are joined directly into the chaining input (mb200_cand),
// and the call/exon records carry every column of the reference's dp_predictions line, so the text DBs can still be printed
// on request -- byte for byte, including what the reference loses on its parse/re-print round trips (text_seq_id_*).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/metaeuk_b200.h"

namespace mb200 {

// resultspercontig as a join.  orfs[0..nOrfs) are the fragments of contigs [0..nContigs) in key order (sorted by contig, as
// mb200_extract_orfs returns them); the accepted alignments of fragment i are alns[alnOff[i] .. alnOff[i+1]).  tKeys maps a target's
// local id to its DB key (NULL = identity), tOff are the target offsets (lengths).  The fragment's DB key is orfKeyBase + i.
// Output: cands of contig c = cands[candOff[c] .. candOff[c+1]) in the order of the reference's search_res_by_contig entry
// (stable by (target key, ORF key), resultspercontig.cpp:180-190); candAln[k] = index into alns of candidate k.
// Returns 0, or -1 when an ORF key would not fit 32 bits / a record names a contig >= nContigs.
int join_by_contig(const mb200_orf *orfs, uint64_t nOrfs, uint32_t nContigs, uint64_t orfKeyBase, const uint64_t *alnOff, const mb200_aln *alns,
                   const uint32_t *tKeys, const uint64_t *tOff, std::vector<mb200_cand> &cands, std::vector<uint64_t> &candOff, std::vector<uint64_t> &candAln);

// Host epilogue of the chaining kernel for one contig (collectoptimalset.cpp:323-330): set E-value = 2^(log2(target residues) + log2(2) - score)
// in double precision, keep the calls with E <= --metaeuk-eval, compact calls and exons.  in[0..n) / exIn are the kernel's calls and exons of
// the contig (exon_begin relative to exIn); out / exOut receive the kept ones, exon_begin rewritten to exBase + position in exOut.  Returns #kept.
uint32_t finish_calls(const mb200_call *in, uint32_t n, const mb200_exon *exIn, uint64_t exBase, uint64_t totalTargetResidues, double evalThr,
                      mb200_call *out, mb200_exon *exOut);

// The sequence-identity column as each stage of the reference prints it: `align` truncates the float to three digits
// and prints 1.0 as "1.00" (Util::fastSeqIdToBuffer + Matcher::resultToBuffer, Util.cpp:222-251, Matcher.cpp:286-287); `resultspercontig`
// parses that text with strtof and prints it the same way again; `collectoptimalset` parses with strtod, narrows to float and prints
// three digits ("1.000").  Every parse/print pair can lose one unit in the last digit, so the later texts are derived from the earlier ones.
std::string text_seq_id_align(float seqIdEst);
std::string text_seq_id_bycontig(float seqIdEst);
std::string text_seq_id_exon(float seqIdEst);
std::string text_evalue(double e);   // "%.3E"

// one 20-column search_res_by_contig line / one 19-column dp_predictions line from the binary records
std::string text_bycontig_line(const mb200_cand &k, const mb200_aln &a, uint32_t qLen, uint64_t contigLen);
std::string text_exon_line(const mb200_call &cl, const mb200_pexon &e);

}  // namespace mb200
