/* metaeuk_b200.h -- C ABI of libmetaeuk_b200.so: the B200-native predictexons hot path.
 *
 * One context per GPU.  Plain pointers and sizes only; every function returns 0 on success or a negative
 * MB200_E_* code and records a message retrievable with mb200_last_error() (the reference's modules print
 * via Debug(Debug::ERROR) and EXIT(EXIT_FAILURE); a caller keeps that convention by exiting on != 0).
 *
 * What each entry point replaces in the reference (paths relative to /root/reference, M/ = lib/mmseqs/):
 *   mb200_matrices_*     SubstitutionMatrix ctor                M/src/commons/SubstitutionMatrix.cpp:12-57
 *                        (Prefiltering.cpp:67-70 kmerSubMat/ungappedSubMat, Alignment.cpp:152 m)
 *   mb200_mask_targets   Masker::maskSequence (tantan)          M/src/commons/Masker.cpp:14-58, M/lib/tantan/tantan.cpp:401-540
 *   mb200_index_build    IndexBuilder::fillDatabase             M/src/prefiltering/IndexBuilder.cpp:55-239
 *   mb200_index_import   IndexTable::initTableByExternalData    M/src/prefiltering/IndexTable.h:223-240
 *   mb200_prefilter      Prefiltering::runSplit query loop      M/src/prefiltering/Prefiltering.cpp:790-870
 *                        = QueryMatcher::matchQuery per query   M/src/prefiltering/QueryMatcher.cpp:85-211
 *   mb200_align          Alignment::run hot loop                M/src/alignment/Alignment.cpp:312-405
 *                        = Matcher::getSWResult per hit         M/src/alignment/Matcher.cpp:60-142
 *   mb200_chain          collectoptimalset per-contig loop      src/exonpredictor/collectoptimalset.cpp:273-418
 *   mb200_search         blastp.sh prefilter -> align           M/data/workflow/blastp.sh:70,85
 *   mb200_predict_exons  predictexons.sh:40-82 (extractorfs, translatenucs, search, resultspercontig, collectoptimalset)
 *                        with binary records handed from stage to stage instead of text DBs
 *   mb200_index_save / mb200_index_load   the persisted index of `createindex` (M/src/prefiltering/PrefilteringIndexReader.cpp:60-330)
 *
 * Residues are the reference's numeric codes (Sequence::mapSequence): A C D E F G H I K L M N P Q R S T V W Y = 0..19, X = 20.
 * Target ids are LOCAL ids (rank in the target DB as the reference's DBReader enumerates it); dbKeys map them back.
 * Limits (checked, fail loudly): sequences < 32768 residues, < 2^27 targets, < 2^32 target residues, k = 6 spaced 1101010011.
 */
#ifndef METAEUK_B200_H
#define METAEUK_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MB200_OK 0
#define MB200_E_ARG (-1)
#define MB200_E_CUDA (-2)
#define MB200_E_LIMIT (-3)     /* an input exceeds a documented limit */
#define MB200_E_OVERFLOW (-4)  /* a per-query work list exceeded its configured capacity */
#define MB200_E_STATE (-5)     /* call order (e.g. prefilter before index) */
#define MB200_E_INTERNAL (-6)

typedef struct mb200_ctx mb200_ctx;

typedef struct mb200_params {
    /* prefilter (MetaEuk defaults) */
    int32_t kmer_thr;          /* 127 = -s 4, k 6 (Prefiltering::getKmerThreshold) */
    int32_t max_res_list_len;  /* --max-seqs 300 */
    int32_t min_diag_score;    /* --min-ungapped-score 15 (must be >= 1) */
    int32_t bin_size;          /* BINSIZE of CacheFriendlyOperations<BINSIZE>, 2 for <= 2.6 M targets (QueryMatcher.cpp:422-450) */
    /* align */
    int32_t gap_open, gap_extend;   /* 11, 1 */
    int32_t min_aln_len;            /* --min-aln-len 11 */
    int32_t pad0;
    double eval_thr;                /* -e 100 */
    /* exon chaining (LocalParameters.h:138-146) */
    int32_t min_intron, max_intron, min_exon_aa, max_overlap, set_gap_open, set_gap_extend, max_exon_sets, pad1;
    double tcov, metaeuk_eval;
    /* capacities of the per-CTA work lists, log2 (0 = default) */
    int32_t cap_items_log2, cap_kmers_log2, cap_hits_log2, pad2;
    /* flags of `metaeuk predictexons` that change what the path computes (defaults 1, 0, 1) */
    int32_t comp_bias_corr;     /* --comp-bias-corr: 0 = no composition-bias correction in the prefilter and the alignment (QueryMatcher.cpp:91-99,
                                   StripedSmithWaterman.cpp:1228-1240) */
    int32_t reverse_fragments;  /* --reverse-fragments 1: every six-frame fragment is searched back to front (the null model; data/predictexons.sh:52-62,
                                   `reverseseq`); headers and coordinates stay those of the forward fragment */
    int32_t translation_table;  /* --translation-table: NCBI genetic code of the ORF scan (its stop codons, Orf.cpp:54-64) and of the translation
                                   (TranslateNucl.h:110-246): 1-6, 9-16, 21-31 */
    int32_t pad3;
} mb200_params;

typedef struct mb200_hit { uint32_t seq_id; int32_t score; int32_t diag; } mb200_hit;       /* = hit_t written to pref_N */
typedef struct mb200_aln {                                                                  /* = Matcher::result_t (SCORE_COV mode) */
    uint32_t seq_id; int32_t raw_score; int32_t bit_score;
    int32_t q_start, q_end, db_start, db_end, aln_len;
    int32_t passed;            /* 1 = accepted by Alignment::checkCriteria (e-value, min-aln-len) */
    float seq_id_est;          /* Matcher::estimateSeqIdByScorePerCol */
    float q_cov, db_cov;
    double evalue;
} mb200_aln;
typedef struct mb200_cand {   /* one 20-column search_res_by_contig line */
    uint32_t target_key; int32_t bit_score; int32_t q_start, q_end; int32_t t_start, t_end, t_len; uint32_t orf_key; int32_t frag_from, frag_to;
} mb200_cand;
typedef struct mb200_exon { uint32_t cand_idx; int32_t contig_start, contig_end, nuc_len; } mb200_exon;
typedef struct mb200_call {
    uint32_t target_key; int32_t strand; int32_t total_bit_score; int32_t n_exons; uint32_t exon_begin; int32_t low_coord, high_coord, pad; double evalue;
} mb200_call;
typedef struct mb200_pexon {  /* one exon of a call = columns 8-19 of a dp_predictions line (PotentialExon::exonToBuffer, PredictionParser.h:97-163) */
    uint32_t orf_key; int32_t bit_score; float seq_id_est;   /* seq_id_est / evalue as `align` computed them; the text columns derive from them */
    int32_t t_start, t_end, t_len, contig_start, contig_end, nuc_len, frag_from, frag_to, pad;
    double evalue;
} mb200_pexon;
typedef struct mb200_pred {   /* one call of collectoptimalset = Prediction::setByDPRes (PredictionParser.h:253-272) */
    uint32_t target_key; int32_t strand; uint32_t total_bit_score, n_exons, low_coord, high_coord;
    uint32_t exon_begin, pad;  /* exon keys of the call: exon_keys[exon_begin .. exon_begin + n_exons) */
    double evalue;             /* atof of the printed %.3E value, as the reference parses it */
} mb200_pred;
typedef struct mb200_pred_out { uint32_t cluster_id, cluster_low_coord, no_overlap_cluster_id, is_rep; } mb200_pred_out;
typedef struct mb200_orf {   /* one six-frame fragment = one entry of the reference's nucl_6f_h / aa_6f DBs (key = index in the array) */
    uint32_t contig;           /* index of the contig in the input batch */
    uint32_t from, to;         /* header coordinates on the contig (Orf::writeOrfHeader): from > to on the minus strand */
    uint32_t flags;            /* 1 = incomplete start, 2 = incomplete end (header column 3), 4 = minus strand */
    uint64_t aa_off; uint32_t aa_len;   /* residues aa[aa_off .. aa_off + aa_len) */
    uint32_t strand_from;      /* first nucleotide on the strand the fragment was read from */
} mb200_orf;
typedef struct mb200_stats {   /* algorithmic work counters of the last prefilter / align call (roofline numerators) */
    uint64_t kmers, lookups, db_matches, double_hits, diag_cells, pref_results;
    uint64_t sw_pairs, sw_cells_fwd, sw_cells_rev, sw_target_bytes;
    uint64_t query_residues, n_queries;
    double ms_prep, ms_prefilter, ms_sw_fwd, ms_sw_rev, ms_chain, ms_h2d, ms_d2h, ms_host;
    uint32_t launches;
    uint32_t overflow_queries;        /* queries the prefilter could not process (work lists exhausted even in the second pass, or >= 32768 residues) and SKIPPED with an empty hit list --
                                         only under overflow policy "skip" (the default of mb200_predict_exons; mb200_option_set "overflow_policy"); mb200_query_status_get tells which */
    uint64_t pref_phase_cycles[16];   /* k_prefilter: CTA cycles summed over queries, per phase; [15] = between queries */
    double ms_orf, ms_orf_host;       /* mb200_extract_orfs: kernels with the contigs resident (CUDA events) / whole call incl. H2D and D2H (wall clock) */
    uint64_t orf_nucleotides, orf_fragments, orf_residues;
    double ms_reduce;                 /* k_reduce (mb200_reduce_redundancy) */
    /* mb200_predict_exons: sums over its contig batches (the ms_* and counter fields above are then sums too) */
    double ms_join, ms_predict;       /* host join of alignments and fragments per contig / the whole call (wall clock) */
    uint64_t pe_batches, pe_contigs, pe_fragments, pe_alignments, pe_calls, pe_exons;
    uint64_t pe_streams;              /* worker streams the last mb200_predict_exons call used (2: stage times are those of the busier stream) */
    double ms_prefilter_sum, ms_sw_fwd_sum, ms_sw_rev_sum, ms_dev_sum;   /* mb200_predict_exons: CUDA-event stage times summed over ALL worker streams (kernels of two streams share the GPU, so each
                                                                             is longer than it would be alone): the denominators of the rooflines in bench.py; ms_dev_sum = all device stages */
    double ms_dev_busy;               /* mb200_predict_exons: length of the UNION over all worker streams of the stage intervals [start event, end event] on one device timeline, i.e. the time at
                                         least one stage of the call had work in flight on the GPU.  With one stream it equals the sum of the stage times; with two it is the device time of the
                                         call without double counting (and without the host-only gaps): the denominator of bench.py's resident throughput */
} mb200_stats;

int mb200_ctx_create(int device, mb200_ctx **out);
void mb200_ctx_destroy(mb200_ctx *ctx);
const char *mb200_last_error(void);
const char *mb200_build_kind(void);   /* "sm_100a" for the library; "simt-emulation" only for the CPU test build under tests/simt (kernels as fibers; never shipped) */
void mb200_params_default(mb200_params *p);
/* what `-s <sensitivity>` means for kmer_thr: Prefiltering::getKmerThreshold for a sequence search with k = 6 (M/src/prefiltering/Prefiltering.cpp:1005-1064);
 * -s 4 (MetaEuk's default) -> 127 */
int mb200_kmer_threshold(float sensitivity);
/* per-query status of the last prefilter pass over the resident batch: 0 ok, 1 sequence too long, 2 work lists exhausted (reference: QueryMatcher.cpp:280-316 takes its own
 * overflow path there, which changes the result and is not reproduced).  Policy (mb200_option_set "overflow_policy"): 0 = such a query fails the call (MB200_E_OVERFLOW /
 * MB200_E_LIMIT; default of the prefilter / search entry points: never a silent drop), 1 = it is skipped with an empty hit list and counted in mb200_stats.overflow_queries
 * (default of mb200_predict_exons: one pathological fragment must not abort millions), -1 = those defaults. */
int mb200_query_status_get(mb200_ctx *ctx, uint32_t *status /* n_queries */);
int mb200_params_set(mb200_ctx *ctx, const mb200_params *p);

/* matrices: built like the reference builds them, or supplied by a caller that already holds the reference's BaseMatrix objects */
int mb200_matrices_default(mb200_ctx *ctx);
int mb200_matrices_set(mb200_ctx *ctx, const int16_t seed21x21[441], const int16_t ungapped21x21[441], const int16_t aln21x21[441],
                       const double seed_pback[21], const double aln_pback[21]);
int mb200_matrices_get(mb200_ctx *ctx, int16_t seed21x21[441], int16_t ungapped21x21[441], int16_t aln21x21[441]);

/* target side.  masked = residues after the reference's tantan masking (Masker::maskSequence; masked positions = 20),
 * unmasked = raw residues (the align stage reads the unmasked target, Alignment.cpp:346-367). off[n_targets+1]. */
int mb200_index_build(mb200_ctx *ctx, const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t n_targets);
int mb200_index_import(mb200_ctx *ctx, const uint64_t *offsets, const uint8_t *entries6, uint64_t n_entries,
                       const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t n_targets);
/* multi-GPU: rank 0 builds, every rank allocates with mb200_index_alloc (sizes from rank 0), the caller broadcasts the
 * device buffers listed by mb200_index_buffers (NCCL over NVLink), then each rank calls mb200_index_commit. */
int mb200_index_sizes(mb200_ctx *ctx, uint64_t sizes[4] /* n_entries, n_targets, total_residues, reserved */);
int mb200_index_alloc(mb200_ctx *ctx, const uint64_t sizes[4]);
int mb200_index_buffers(mb200_ctx *ctx, void *dev_ptrs[8], uint64_t nbytes[8], int *n);
int mb200_index_commit(mb200_ctx *ctx);
/* Multi-GPU: the query side shards by contig with no data-path collective; the ONE collective of the path replicates the
 * target index.  mb200_index_broadcast is collective over `nccl_comm` (an ncclComm_t; NULL = the communicator mb200_comm_init created for this
 * context): the root's index arena (k-mer offsets | index entries | masked residues | unmasked residues | target offsets, one device
 * allocation) reaches every rank with ONE ncclBroadcast over NVLink/NVSwitch, preceded by a 32-byte size header; communicator setup is not
 * part of it.  nbytes / ms (optional) return the arena size and the CUDA-event time of the arena broadcast.  NCCL is resolved at run time
 * (dlopen libnccl.so.2, an already loaded copy is joined), so the library has no link dependency on it.
 * mb200_comm_unique_id: ncclGetUniqueId into a caller buffer of 128 bytes (rank 0; ship it to the other ranks by any means);
 * mb200_comm_init: ncclCommInitRank, one rank per context / process;  mb200_comm_init_all: ncclCommInitAll for n contexts of ONE process. */
int mb200_comm_unique_id(void *id128);
int mb200_comm_init(mb200_ctx *ctx, const void *id128, int rank, int world);
int mb200_comm_init_all(mb200_ctx **ctxs, int n);
int mb200_index_broadcast(mb200_ctx *ctx, void *nccl_comm, int root, uint64_t *nbytes, double *ms);
/* one process, several contexts on several devices: copy the index (and the target keys) of `src` device to device into `dst` (cudaMemcpyPeerAsync), commit it there.
 * Replaces, for a single-process host driver (bin/mb200 gpupredictexons --gpus N), what mb200_index_broadcast does for one process per GPU. */
int mb200_index_clone(mb200_ctx *dst, mb200_ctx *src, uint64_t *nbytes, double *ms);

/* Repeat masking of the targets before indexing = Masker::maskSequence with the prefilter's default flags (--mask 1 --mask-prob 0.9;
 * M/src/commons/Masker.cpp:14-58 -> tantan, M/lib/tantan/tantan.cpp:401-433,478-540; called from IndexBuilder::fillDatabase, IndexBuilder.cpp:84-131).
 * Host only (no GPU, no context): unmasked / masked are numeric residue codes with the same offsets off[n_targets+1]; masked letters become X (20).
 * simd_width 4 reproduces the summation order of the reference's AVX2 build, 2 that of its SSE4.1 / NEON build (0 = 4). */
int mb200_mask_targets(const uint8_t *unmasked, const uint64_t *off, uint32_t n_targets, double mask_prob, int simd_width, uint8_t *masked, uint64_t *n_masked);
/* the same with every switch of Masker::maskSequence: mask_tantan = --mask, mask_n_repeat = --mask-n-repeat (runs of more than that many equal letters, 0 = off;
 * Masker::maskRepeats, Masker.cpp:83-116, including its blind spot for a run of A at the start of a sequence), lower_case = --mask-lower-case 1: one byte per
 * residue, != 0 where the DB holds a lower-case letter (NULL = off).  n_masked = the reference's "Masked residues" count. */
int mb200_mask_targets_ex(const uint8_t *unmasked, const uint64_t *off, uint32_t n_targets, int mask_tantan, double mask_prob, int simd_width,
                          const uint8_t *lower_case, int mask_n_repeat, uint8_t *masked, uint64_t *n_masked);

/* persisted index: an MMseqs index database in the layout `createindex` / `indexdb` write (<path>, <path>.index, <path>.dbtype;
 * entries keyed as in M/src/prefiltering/PrefilteringIndexReader.cpp:10-34).  mb200_index_load mmaps such a DB -- written by
 * mb200_index_save / mb200_index_file_write or by the reference itself -- checks its META entry (k = 6, alphabet 21, spaced pattern, one split,
 * amino acids, k-mer threshold equal to params.kmer_thr), uploads k-mer table, masked and unmasked residues and takes the DB keys
 * (mb200_target_keys_set).  mb200_index_file_write builds the index on the host (no GPU needed; default seed matrix) and writes it.
 * A DB written here also carries SCOREMATRIXNAME, SPACEDPATTERN and the 2-mer / 3-mer similarity tables (ScoreMatrix::serialize layout for AVX2 builds), so the
 * unmodified reference runs `prefilter` / `align` from it as it does from its own `.idx` (tests/test_cli_cpu.py). */
int mb200_index_save(mb200_ctx *ctx, const char *path);
int mb200_index_load(mb200_ctx *ctx, const char *path);
int mb200_index_file_write(const char *path, const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t n_targets,
                           const uint32_t *keys /* NULL = local ids */, int kmer_thr, int mask_mode);
/* copies the DB keys of the targets (identity if none were set); keys: caller buffer of n_targets */
int mb200_target_keys_get(mb200_ctx *ctx, uint32_t *keys);

/* query side: host buffers in, host buffers out (the call a module makes).  q_off[n_q+1]. */
int mb200_prefilter(mb200_ctx *ctx, const uint8_t *q, const uint64_t *q_off, uint32_t n_q, mb200_hit *hits /* n_q*max_res_list_len */, uint32_t *hit_count);
/* alns[k] for every listed hit k (<= max_res_list_len hits per query): hits whose score can pass -e carry all fields and `passed`;
 * the others (which the reference rejects before computing start coordinates) only carry seq_id and passed = 0. */
int mb200_align(mb200_ctx *ctx, const uint8_t *q, const uint64_t *q_off, uint32_t n_q, const uint32_t *hit_seq_id, const uint64_t *hit_off,
                mb200_aln *alns /* hit_off[n_q] */);
/* same as mb200_prefilter with a compact result: hits of query i = (*hits)[hit_off[i]..hit_off[i+1]).  hit_off is a caller buffer of
 * n_q+1; *hits is owned by the context and valid until the next call on it. */
int mb200_prefilter_compact(mb200_ctx *ctx, const uint8_t *q, const uint64_t *q_off, uint32_t n_q, uint64_t *hit_off, const mb200_hit **hits);
/* prefilter + align in one call (what `search` writes to search_res): the ACCEPTED alignments of query i are
 * (*alns)[aln_off[i]..aln_off[i+1]) in prefilter order (sort with Matcher::compareHits to print).  aln_off: caller buffer n_q+1;
 * *alns is owned by the context and valid until the next call on it. */
int mb200_search(mb200_ctx *ctx, const uint8_t *q, const uint64_t *q_off, uint32_t n_q, uint64_t *aln_off, const mb200_aln **alns);
/* exon chaining for a batch of contigs: cands of contig c = cands[cand_off[c]..cand_off[c+1]) sorted by (target_key, orf_key).
 * calls/exons are caller buffers sized 2*n_cands*max_exon_sets and n_cands*max_exon_sets; call_off/exon ranges are per contig. */
int mb200_chain(mb200_ctx *ctx, const mb200_cand *cands, const uint64_t *cand_off, uint32_t n_contigs, uint64_t total_target_residues,
                mb200_call *calls, uint32_t *call_count /* n_contigs */, mb200_exon *exons);

/* Six-frame ORF fragments of nucleotide contigs and their translation = `extractorfs` + `translatenucs` with the flags of
 * `metaeuk predictexons` (M/src/util/extractorfs.cpp:19-159, M/src/commons/Orf.cpp:230-349, M/src/util/translatenucs.cpp:13-118,
 * M/src/commons/TranslateNucl.h:383-503).  nuc = concatenated ASCII nucleotides (as stored in the contig DB, any case, IUPAC codes),
 * off[n_contigs+1]; min_len / max_len in codons (15 / 32734), start_mode 1.  Fragments come back in the reference's key order;
 * *orfs and *aa are owned by the context (valid until the next call).  mb200_queries_from_orfs() then makes the fragments the resident
 * query batch without a host round trip (follow with mb200_prefilter_resident / mb200_align_resident). */
int mb200_extract_orfs(mb200_ctx *ctx, const char *nuc, const uint64_t *off, uint32_t n_contigs, int min_len, int max_len, int start_mode,
                       uint64_t *n_orfs, const mb200_orf **orfs, uint64_t *n_aa, const char **aa);
int mb200_queries_from_orfs(mb200_ctx *ctx);
/* mb200_search on the resident query batch (mb200_queries_upload or mb200_queries_from_orfs); aln_off: caller buffer of n_queries + 1 */
int mb200_search_resident(mb200_ctx *ctx, uint64_t *aln_off, const mb200_aln **alns);

/* predictexons with the contigs as the only input and binary records between the stages (data/predictexons.sh:40-82):
 * six-frame fragments (device) -> resident query batch -> prefilter -> gapped alignment -> join per contig (resultspercontig.cpp:64-196,
 * as a binary join) -> exon chaining -> calls.  Contigs are processed in batches of about "predict_batch_nt" nucleotides (mb200_option_set;
 * results do not depend on it; default 32 M, less for small calls so that every worker stream gets at least four batches, never below 4 M).  Fragment keys (orf_key) count over the whole call in the reference's key order.  Target keys are the
 * keys given to mb200_target_keys_set (default: local ids).  Calls of contig c = (*calls)[call_off[c] .. call_off[c+1]) in the order of the
 * reference's dp_predictions entry; the exons of a call = (*exons)[exon_begin .. exon_begin + n_exons).  call_off: caller buffer of
 * n_contigs + 1; *calls and *exons are owned by the context (valid until the next call on it). */
int mb200_predict_exons(mb200_ctx *ctx, const char *nuc, const uint64_t *off, uint32_t n_contigs, int min_len, int max_len,
                        uint64_t *call_off, const mb200_call **calls, const mb200_pexon **exons, uint64_t *n_exons);
int mb200_target_keys_set(mb200_ctx *ctx, const uint32_t *keys, uint32_t n_targets);   /* DB keys of the targets in local-id order */
/* one 19-column dp_predictions line (Prediction::predictionToBuffer, PredictionParser.h:368-396) from the binary records, with the
 * digits the reference's text round trips leave; returns the length written (without the NUL) or MB200_E_ARG if cap is too small */
int mb200_text_exon_line(const mb200_call *call, const mb200_pexon *exon, char *buf, size_t cap);
/* the dp_predictions DB of contigs [0, n_contigs) as one text buffer: for every contig with at least one call ">contigKey\n" followed by the
 * entry the reference writes for it (the 19-column lines of its calls in order, DBWriter entry without the trailing NUL;
 * collectoptimalset.cpp:399-418).  contig_keys NULL = keys 0..n-1.  Returns the number of bytes the text needs (written only if <= cap);
 * this is what bench.py / the tests hash against the reference's DB (tools/refrun.py: digest_entries). */
int64_t mb200_text_dp_db(const uint64_t *call_off, uint32_t n_contigs, const uint32_t *contig_keys, const mb200_call *calls, const mb200_pexon *exons,
                         char *buf, size_t cap);

/* reduceredundancy (src/exonpredictor/reduceredundancy.cpp:24-147,273-296) for a batch of contigs: calls of contig c =
 * preds[pred_off[c] .. pred_off[c+1]) in the order of the calledExons entry.  Per call: cluster representative (target key, low coordinate),
 * the winner of the same-strand overlap exclusion, and whether the call is a representative.  member_order (per contig, same ranges) = the
 * order of the call -> representative DB; rep_order[pred_off[c] .. + rep_count[c]) = the representatives in output order (a representative is
 * printed when overlaps are allowed or its no_overlap_cluster_id equals its own target key).  Indices are relative to the contig's range. */
int mb200_reduce_redundancy(mb200_ctx *ctx, const mb200_pred *preds, const uint64_t *pred_off, uint32_t n_contigs, const uint32_t *exon_keys, uint64_t n_exon_keys,
                            mb200_pred_out *out, uint32_t *member_order, uint32_t *rep_order, uint32_t *rep_count);

/* device-resident variant used for kernel-only timing: upload once, run many times, download once */
int mb200_queries_upload(mb200_ctx *ctx, const uint8_t *q, const uint64_t *q_off, uint32_t n_q);
int mb200_prefilter_resident(mb200_ctx *ctx);
int mb200_align_resident(mb200_ctx *ctx);             /* aligns the resident prefilter result */
/* dense layouts [n_q][max_res_list_len]; alns holds accepted alignments only, aln_count[i] of them */
int mb200_results_download(mb200_ctx *ctx, mb200_hit *hits, uint32_t *hit_count, mb200_aln *alns, uint32_t *aln_count);

/* tuning / diagnostic switches (results never depend on them, except "overflow_policy" and "orf_key_base" which are documented above / below):
 *   "phase_stats" 0|1            per-phase cycle accounting in k_prefilter (mb200_stats.pref_phase_cycles)
 *   "prefilter_config" 0..4      0 = chosen from the index density, 1 dense (512 threads x 2 CTAs/SM), 2 sparse, 3 medium, 4 small;  "prefilter_short_len" / "prefilter_short_config": a
 *                                separate launch with another configuration for queries of at most that many residues (0 = off)
 *   "prefilter_buckets" 0|1      inline 32-byte bucket per k-mer (default) or the offsets + entries arrays
 *   "prefilter_stream_hint" 0..3 cache policy of the index reads: bit 0 = L1 no-allocate + L2 evict-first, bit 1 = 64-byte instead of 128-byte HBM fetch (default 3)
 *   "prefilter_sort_walk" 0..256 multi-hit lists of at most this many entries are sorted and walked instead of grouped by hashing (default 0)
 *   "l2_fetch_granularity" 32|64|128   cudaLimitMaxL2FetchGranularity of the device (diagnostic; measured: no effect on sm_100)
 *   "sw_kernel" 1|2|4|5|6        alignment kernel family (6 = k_sw_stream, default; 1, 2, 4, 5 = the round-1 layouts kept for A/B);  "sw_lane_max_hits": sub-variant of 4 / 5
 *   "sw_ring" -1|0|1..16         query profiles per warp of k_sw_stream: 0 = sized from the pairs per query (default), -1 = one private profile per slot
 *   "sw_ring_fit" 0|1            cap that ring so that 8 warps x 2 CTAs fit on an SM (default 1: long queries no longer run at 8 warps per SM)
 *   "predict_batch_nt" n         nucleotides per contig batch of mb200_predict_exons (default: 32 M, fewer for small calls);  "predict_streams" 1..4: worker streams (default 2)
 *   "orf_key_base" n             fragment key of the first fragment of the next mb200_predict_exons call (a rank's shard of a contig DB) */
int mb200_option_set(mb200_ctx *ctx, const char *name, int64_t value);
int mb200_stats_get(mb200_ctx *ctx, mb200_stats *out);
int mb200_profile_dump(mb200_ctx *ctx, const char *json_path);   /* the same counters and stage times as one JSON object in a file */
/* CUDA-event stopwatch on the context's stream (the stream every kernel and copy of this library is issued on) */
int mb200_timer_start(mb200_ctx *ctx);
int mb200_timer_stop(mb200_ctx *ctx, double *ms);
double mb200_evalue(mb200_ctx *ctx, int raw_score, int q_len);   /* EvalueComputation::computeEvalue with the ctx's target residues */
int mb200_bit_score(mb200_ctx *ctx, int raw_score);

#ifdef __cplusplus
}
#endif
#endif
