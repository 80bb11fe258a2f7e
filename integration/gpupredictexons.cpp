// integration/gpupredictexons.cpp -- the reference-side binding of libmetaeuk_b200, as a real translation unit.
//
// A MetaEuk module with the `Command` signature (reference: lib/mmseqs/src/commons/Command.h:94-105; MetaEuk's own table: src/metaeuk.cpp:21-96):
//     metaeuk-b200 gpupredictexons <i:contigsDB> <i:targetsDB> <o:calledExonsDB> [the flags of `metaeuk predictexons`]
// It replaces the module chain of data/predictexons.sh:40-87 (extractorfs, translatenucs, prefilter, align, resultspercontig, collectoptimalset).
// Everything around the hot path is the reference's own code, compiled from /root/reference: flag parsing (LocalParameters), DB input (DBReader),
// residue codes (SubstitutionMatrix::aa2num), DB output (DBWriter).  The hot path is reached only through the C ABI of include/metaeuk_b200.h
// -- plain pointers and sizes, no MMseqs2 type crosses the boundary.
//
// Built only where the reference is present (integration/build.sh; the build container) -- it is the compiled form of the binding a maintainer
// adds, not part of libmetaeuk_b200.so.
#include "Command.h"
#include "DBReader.h"
#include "DBWriter.h"
#include "Debug.h"
#include "LocalParameters.h"
#include "SubstitutionMatrix.h"
#include "Util.h"

#include "metaeuk_b200.h"

#include <climits>
#include <string>
#include <vector>

#define MB(x) do { if ((x) != MB200_OK) { Debug(Debug::ERROR) << "libmetaeuk_b200: " << mb200_last_error() << "\n"; EXIT(EXIT_FAILURE); } } while (0)

// the defaults `predictexons` sets before parsing (src/workflow/PredictExons.cpp:8-16)
static void setGpuPredictExonsDefaults(Parameters *p) {
    p->orfStartMode = 1;
    p->orfMinLength = 15;
    p->alignmentMode = Parameters::ALIGNMENT_MODE_SCORE_COV;
    p->evalThr = 100;
}

int gpupredictexons(int argc, const char **argv, const Command &command) {
    LocalParameters &par = LocalParameters::getLocalInstance();
    setGpuPredictExonsDefaults(&par);
    par.parseParameters(argc, argv, command, true, 0, 0);

    // ---- inputs through the reference's readers (local ids = key rank, DBReader::sortIndex)
    DBReader<unsigned int> contigs(par.db1.c_str(), par.db1Index.c_str(), par.threads, DBReader<unsigned int>::USE_DATA | DBReader<unsigned int>::USE_INDEX);
    contigs.open(DBReader<unsigned int>::NOSORT);
    DBReader<unsigned int> targets(par.db2.c_str(), par.db2Index.c_str(), par.threads, DBReader<unsigned int>::USE_DATA | DBReader<unsigned int>::USE_INDEX);
    targets.open(DBReader<unsigned int>::NOSORT);
    if (Parameters::isEqualDbtype(targets.getDbtype(), Parameters::DBTYPE_AMINO_ACIDS) == false) {
        Debug(Debug::ERROR) << "gpupredictexons: the target DB must hold amino-acid sequences (profile targets stay with predictexons)\n";
        EXIT(EXIT_FAILURE);
    }
    SubstitutionMatrix subMat(par.scoringMatrixFile.values.aminoacid().c_str(), 2.0, 0.0);

    const uint32_t nT = (uint32_t) targets.getSize();
    std::vector<uint64_t> toff(1, 0); std::vector<uint8_t> unmasked; std::vector<uint8_t> lower; std::vector<uint32_t> tkeys(nT);
    for (size_t id = 0; id < nT; id++) {
        const char *s = targets.getData(id, 0); const size_t L = targets.getSeqLen(id);
        for (size_t k = 0; k < L; k++) {
            unmasked.push_back((uint8_t) subMat.aa2num[(int) s[k]]);
            if (par.maskLowerCaseMode) lower.push_back(islower((unsigned char) s[k]) ? 1 : 0);
        }
        toff.push_back(unmasked.size()); tkeys[id] = targets.getDbKey(id);
    }
    std::string nuc; std::vector<uint64_t> coff(1, 0);
    for (size_t id = 0; id < contigs.getSize(); id++) { nuc.append(contigs.getData(id, 0), contigs.getSeqLen(id)); coff.push_back(nuc.size()); }

    // ---- flags -> mb200_params (the values `predictexons` hands to its modules, src/workflow/PredictExons.cpp:38-48)
    mb200_ctx *ctx = NULL; MB(mb200_ctx_create(0, &ctx));
    mb200_params p; mb200_params_default(&p);
    const int kScore = par.kmerScore.values.sequence();
    p.kmer_thr = (kScore != INT_MAX) ? kScore : mb200_kmer_threshold(par.sensitivity);     // Prefiltering::getKmerThreshold (Prefiltering.cpp:1005-1011)
    p.max_res_list_len = (int) par.maxResListLen; p.min_diag_score = par.minDiagScoreThr; p.eval_thr = par.evalThr;
    p.gap_open = par.gapOpen.values.aminoacid(); p.gap_extend = par.gapExtend.values.aminoacid();
    p.min_aln_len = (int) par.minExonAaLength;                                            // "align module should return alignments of at least a minimal exon length"
    p.comp_bias_corr = par.compBiasCorrection; p.reverse_fragments = par.reverseFragments; p.translation_table = par.translationTable;
    p.metaeuk_eval = par.metaeukEvalueThr; p.tcov = par.metaeukTargetCovThr;
    p.max_intron = (int) par.maxIntronLength; p.min_intron = (int) par.minIntronLength; p.min_exon_aa = (int) par.minExonAaLength; p.max_overlap = (int) par.maxAaOverlap;
    p.max_exon_sets = (int) par.maxExonSets; p.set_gap_open = par.setGapOpenPenalty; p.set_gap_extend = par.setGapExtendPenalty;
    MB(mb200_params_set(ctx, &p));
    MB(mb200_matrices_default(ctx));

    // ---- target index: masking as IndexBuilder::fillDatabase does it (tantan, lower case, N-repeats), then the k-mer index, both inside the library
    std::vector<uint8_t> masked(unmasked);
    if (par.maskMode || par.maskLowerCaseMode || par.maskNrepeats)
        MB(mb200_mask_targets_ex(unmasked.data(), toff.data(), nT, par.maskMode, par.maskProb, 0, par.maskLowerCaseMode ? lower.data() : NULL, par.maskNrepeats, masked.data(), NULL));
    MB(mb200_index_build(ctx, masked.data(), unmasked.data(), toff.data(), nT));
    MB(mb200_target_keys_set(ctx, tkeys.data(), nT));

    // ---- the hot path: contigs in, exon calls out
    std::vector<uint64_t> callOff(contigs.getSize() + 1); const mb200_call *calls = NULL; const mb200_pexon *exons = NULL; uint64_t nExons = 0;
    MB(mb200_predict_exons(ctx, nuc.data(), coff.data(), (uint32_t) contigs.getSize(), par.orfMinLength, par.orfMaxLength, callOff.data(), &calls, &exons, &nExons));

    // ---- output DB: one entry per contig key, the lines collectoptimalset prints (collectoptimalset.cpp:224-279)
    DBWriter writer(par.db3.c_str(), par.db3Index.c_str(), 1, par.compressed, Parameters::DBTYPE_GENERIC_DB);
    writer.open();
    char line[512]; std::string entry;
    for (size_t id = 0; id < contigs.getSize(); id++) {
        entry.clear();
        for (uint64_t k = callOff[id]; k < callOff[id + 1]; k++)
            for (int x = 0; x < calls[k].n_exons; x++) {
                if (mb200_text_exon_line(&calls[k], &exons[calls[k].exon_begin + x], line, sizeof(line)) < 0) { Debug(Debug::ERROR) << mb200_last_error() << "\n"; EXIT(EXIT_FAILURE); }
                entry.append(line);
            }
        writer.writeData(entry.c_str(), entry.size(), contigs.getDbKey(id), 0);
    }
    writer.close(true);
    mb200_ctx_destroy(ctx);
    contigs.close(); targets.close();
    return EXIT_SUCCESS;
}
