// tests/emu/emu_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the product's kernel source (prefilter_core.cuh) with MB200_EMULATE (serial CTA emulation, see
// cta_prims.cuh) plus the product's host preprocessing (host_tables.cpp) into a CPU library so the kernel
// LOGIC can be diffed against the oracle in a container without a GPU.  Not shipped, not a fallback.
#define MB200_EMULATE 1
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../metaeuk_b200/csrc/host_tables.h"
#include "../../metaeuk_b200/csrc/prefilter_core.cuh"
#include "../../metaeuk_b200/csrc/reduce_core.cuh"
#include "../../metaeuk_b200/csrc/chain_core.cuh"
#include "../../metaeuk_b200/csrc/host_join.h"
#include "../../metaeuk_b200/csrc/host_indexdb.h"
#include "../../metaeuk_b200/csrc/orf_kernels.cuh"    // (host-callable parts only: the stop-codon set of a genetic code)

static uint32_t g_capI = 1u << 17, g_capK = 1u << 18, g_capMAlloc = 1u << 19, g_capM = 1u << 19;   // the library's default per-CTA capacities (mb200.cu: ctx_alloc_scratch)
static int g_compBiasCorr = 1;   // --comp-bias-corr: the product zeroes the per-residue arrays instead of running k_prep (mb200.cu: queries_install)

template <class SH>
static int emu_prefilter_t(const uint8_t *masked, const uint64_t *toff, uint32_t nT, const uint8_t *q, const uint64_t *qOff, uint32_t nQ,
                           int kmerThr, int maxHits, int minDiag, int binSize, MbPrefHit *out, uint32_t *outCount, uint32_t *outStatus, unsigned long long *statsOut) {
    mb200::Matrices m; mb200::default_matrices(m);
    std::vector<uint32_t> sim3; std::vector<uint16_t> cnt3;
    mb200::build_sim3(m.seed, sim3, cnt3, MB_CNT_MIN, MB_CNT_W);
    mb200::TargetIndex idx;
    if (mb200::build_index(masked, toff, nT, m.seed, kmerThr, idx)) return -1;
    MbPrefTables T; memset(&T, 0, sizeof(T));
    std::vector<uint32_t> bits(mb200::KMER_TABLE / 32, 0);   // same bitmap k_kmer_bits builds on the device
    for (size_t k = 0; k < mb200::KMER_TABLE; k++) if (idx.offsets[k + 1] != idx.offsets[k]) bits[k >> 5] |= 1u << (k & 31);
    T.kmerBits = bits.data();
    T.sim3 = sim3.data(); T.cnt3 = cnt3.data(); T.offsets = idx.offsets.data(); T.entries = (const uint2 *)idx.entries.data();
    T.tseq = masked; T.toff = toff; T.nT = nT; T.kmerThr = kmerThr; T.maxHits = maxHits; T.minDiag = minDiag; T.binMask = (uint32_t)binSize - 1;
    T.capP = 32768; T.capI = g_capI; T.capK = g_capK; T.capM = g_capM;
    std::vector<uint32_t> itemOff(T.capP + 1), kmerOff(T.capI + 1), itemInfo(T.capI);
    std::vector<int16_t> wThr(T.capP); std::vector<uint16_t> wI0(T.capP), wI1(T.capP);
    std::vector<uint64_t> A(g_capMAlloc), B(g_capMAlloc), C(g_capMAlloc); std::vector<int32_t> S(g_capMAlloc);
    MbPrefScratch W; W.itemOff = itemOff.data(); W.wThr = wThr.data(); W.wI0 = wI0.data(); W.wI1 = wI1.data(); W.kmerOff = kmerOff.data();
    W.itemInfo = itemInfo.data(); W.A = A.data(); W.B = B.data(); W.C = C.data(); W.S = S.data();
    SH *sh = new SH();
    for (int i = 0; i < 441; i++) sh->ungMat[i] = (int8_t)m.ung[i];
    MbPrefStats st; memset(&st, 0, sizeof(st));
    std::vector<float> bias; std::vector<int8_t> corr;
    for (uint32_t i = 0; i < nQ; i++) {
        const uint8_t *qs = q + qOff[i]; const uint32_t L = (uint32_t)(qOff[i + 1] - qOff[i]);
        bias.resize(L + 1); corr.resize(L + 1);
        for (uint32_t p = 0; p < L; p++) { bias[p] = g_compBiasCorr ? mb_comp_bias_at(m.seed, m.seedPBack, qs, (int)L, (int)p) : 0.0f; corr[p] = g_compBiasCorr ? mb_round_bias_quarter(bias[p]) : 0; }
        mb_prefilter_query(T, W, sh, qs, bias.data(), corr.data(), L, out + (size_t)i * maxHits, outCount + i, outStatus + i, &st);
    }
    delete sh;
    if (statsOut) { statsOut[0] = st.kmers; statsOut[1] = st.lookups; statsOut[2] = st.dbMatches; statsOut[3] = st.doubleHits; statsOut[4] = st.diagCells; statsOut[5] = st.results; }
    return 0;
}



extern "C" {

int emu_matrices(int16_t *seed, int16_t *ung, int16_t *aln, double *seedPBack, double *alnPBack) {
    mb200::Matrices m; mb200::default_matrices(m);
    memcpy(seed, m.seed, 882); memcpy(ung, m.ung, 882); memcpy(aln, m.aln, 882); memcpy(seedPBack, m.seedPBack, 168); memcpy(alnPBack, m.alnPBack, 168);
    return 0;
}

// sim3 unpacked for comparison with the oracle / reference digest
int emu_sim3(int16_t *score, uint32_t *index) {
    mb200::Matrices m; mb200::default_matrices(m);
    std::vector<uint32_t> sim3; std::vector<uint16_t> cnt3;
    mb200::build_sim3(m.seed, sim3, cnt3, MB_CNT_MIN, MB_CNT_W);
    for (size_t i = 0; i < sim3.size(); i++) { score[i] = (int16_t)(sim3[i] & 0xFFFF); index[i] = sim3[i] >> 16; }
    // cnt3 self-check
    for (int r = 0; r < 8000; r += 97) for (int x = -150; x < 180; x += 7) {
        uint32_t c = 0; for (int z = 0; z < 8000; z++) c += ((int16_t)(sim3[(size_t)r * 8000 + z] & 0xFFFF) >= x);
        if (c != mb_count_ge(cnt3.data(), r, x)) return -1;
    }
    return 0;
}

// index in the reference's serialised form (u64 offsets, packed 6-byte entries) for digest comparison
int emu_index(const uint8_t *masked, const uint64_t *off, uint32_t nT, int kmerThr, uint64_t *offsets, uint8_t *entries6, uint64_t *nEntries) {
    mb200::Matrices m; mb200::default_matrices(m);
    mb200::TargetIndex idx;
    if (mb200::build_index(masked, off, nT, m.seed, kmerThr, idx)) return -1;
    *nEntries = idx.entries.size();
    if (offsets) for (size_t k = 0; k <= mb200::KMER_TABLE; k++) offsets[k] = idx.offsets[k];
    if (entries6) for (size_t i = 0; i < idx.entries.size(); i++) { uint32_t s = (uint32_t)idx.entries[i]; uint16_t p = (uint16_t)(idx.entries[i] >> 32); memcpy(entries6 + 6 * i, &s, 4); memcpy(entries6 + 6 * i + 4, &p, 2); }
    return 0;
}

// cfg 1 = dense, 2 = sparse (the two device configurations), 3 = a deliberately tiny one that pushes every list through its
// global-memory fallback (items, non-empty k-mers, hash table -> sort path, sorts in global memory)
// k_orf_scan's stop test (orf_is_stop over the 64-bit set of the genetic code) against the reference's rule -- the upper-cased three letters equal one of the
// table's stop codons (Orf.cpp:54-64,206-228,266-269) -- for every byte triple and every genetic code; returns the number of disagreements (-1: unknown table id)
long long emu_orf_stop_check(void) {
    long long bad = 0; int tables = 0;
    for (int t = 0; t < 40; t++) {
        const char *code = mb200::genetic_code(t);
        if (!code) continue;
        tables++;
        const unsigned long long set = orf_stop_set(code);
        const char L[4] = {'T', 'C', 'A', 'G'};
        for (int a = 0; a < 256; a++) for (int b = 0; b < 256; b++) for (int c = 0; c < 256; c++) {
            const int u0 = a & ~0x20, u1 = b & ~0x20, u2 = c & ~0x20;
            bool want = false;
            for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++) for (int z = 0; z < 4; z++)
                if (code[16 * x + 4 * y + z] == '*' && u0 == L[x] && u1 == L[y] && u2 == L[z]) want = true;
            if (orf_is_stop(u0, u1, u2, set) != want) bad++;
        }
    }
    return tables == 25 ? bad : -1;
}
void emu_set_comp_bias_corr(int on) { g_compBiasCorr = on; }
// capacities of the work lists: (2^17, 2^18, 2^19, 2^19) = the library's first pass; its second pass for overflowed queries uses
// (2^22, 2^21, power-of-two allocation >= 2 max(10^6, #targets), that bound - 1) -- see mb200_prefilter_resident
void emu_set_caps(uint32_t capI, uint32_t capK, uint32_t capMAlloc, uint32_t capM) { g_capI = capI; g_capK = capK; g_capMAlloc = capMAlloc; g_capM = capM; }

int emu_prefilter(const uint8_t *masked, const uint64_t *toff, uint32_t nT, const uint8_t *q, const uint64_t *qOff, uint32_t nQ, int kmerThr, int maxHits,
                  int minDiag, int binSize, MbPrefHit *out, uint32_t *outCount, uint32_t *outStatus, unsigned long long *statsOut, int cfg) {
    if (cfg == 2) return emu_prefilter_t<MbPrefSharedT<MbPrefCfgSparse>>(masked, toff, nT, q, qOff, nQ, kmerThr, maxHits, minDiag, binSize, out, outCount, outStatus, statsOut);
    if (cfg == 3) return emu_prefilter_t<MbPrefSharedT<MbPrefCfg<64, 1, 8, 64, 5>>>(masked, toff, nT, q, qOff, nQ, kmerThr, maxHits, minDiag, binSize, out, outCount, outStatus, statsOut);
    return emu_prefilter_t<MbPrefSharedT<MbPrefCfgDense>>(masked, toff, nT, q, qOff, nQ, kmerThr, maxHits, minDiag, binSize, out, outCount, outStatus, statsOut);
}

int emu_reduce(const MbPred *P, uint32_t n, const uint32_t *exonKeys, MbPredOut *out, uint32_t *memberOrder, uint32_t *repOrder, uint32_t *nRep) {
    std::vector<uint32_t> flag(n + 1), tmp(n + 1);
    *nRep = mb_reduce_contig(P, n, exonKeys, out, memberOrder, repOrder, flag.data(), tmp.data());
    return 0;
}

// ---- binary hand-off between the stages (host_join.cpp) and the chaining kernel logic (chain_core.cuh) ----
// cands / candAln: caller buffers of alnOff[nOrfs] entries, candOff of nContigs + 1
int emu_join(const mb200_orf *orfs, uint64_t nOrfs, uint32_t nContigs, uint64_t orfKeyBase, const uint64_t *alnOff, const mb200_aln *alns, const uint32_t *tKeys,
             const uint64_t *tOff, mb200_cand *cands, uint64_t *candOff, uint64_t *candAln) {
    std::vector<mb200_cand> c; std::vector<uint64_t> co, ca;
    const int rc = mb200::join_by_contig(orfs, nOrfs, nContigs, orfKeyBase, alnOff, alns, tKeys, tOff, c, co, ca);
    if (rc) return rc;
    memcpy(candOff, co.data(), co.size() * 8);
    if (!c.empty()) { memcpy(cands, c.data(), c.size() * sizeof(mb200_cand)); memcpy(candAln, ca.data(), ca.size() * 8); }
    return 0;
}

// one contig through mb_chain_contig (what one thread of k_chain runs) + the host epilogue; calls: 2n entries, exons: n entries
int emu_chain(const mb200_cand *cands, uint32_t n, int minIntron, int maxIntron, int minExonAa, int maxOverlap, int gapOpen, int gapExtend, int maxSets, double tcov,
              uint64_t totalTargetResidues, double evalThr, mb200_call *calls, mb200_exon *exons) {
    static_assert(sizeof(MbCand) == sizeof(mb200_cand) && sizeof(MbCall) == sizeof(mb200_call) && sizeof(MbExon) == sizeof(mb200_exon), "ABI structs");
    MbChainParams P; P.minIntron = minIntron; P.maxIntron = maxIntron; P.minExonAa = minExonAa; P.maxOverlap = maxOverlap; P.gapOpen = gapOpen; P.gapExtend = gapExtend;
    P.maxSets = maxSets; P.tcov = tcov;
    std::vector<MbPE> pe(n + 1); std::vector<int32_t> op(n + 1), om(n + 1); std::vector<MbRow> rows(n + 1); std::vector<MbCall> cl(2 * n + 2); std::vector<MbExon> ex(n + 1);
    uint32_t nEx = 0;
    const int nc = mb_chain_contig(reinterpret_cast<const MbCand *>(cands), n, P, pe.data(), op.data(), om.data(), rows.data(), cl.data(), ex.data(), &nEx);
    if (nc < 0) return nc;
    return (int)mb200::finish_calls(reinterpret_cast<const mb200_call *>(cl.data()), (uint32_t)nc, reinterpret_cast<const mb200_exon *>(ex.data()), 0, totalTargetResidues, evalThr, calls, exons);
}

int emu_text_bycontig_line(const mb200_cand *k, const mb200_aln *a, uint32_t qLen, uint64_t contigLen, char *buf) { const std::string l = mb200::text_bycontig_line(*k, *a, qLen, contigLen); memcpy(buf, l.c_str(), l.size() + 1); return (int)l.size(); }
int emu_text_exon_line(const mb200_call *c, const mb200_pexon *e, char *buf) { const std::string l = mb200::text_exon_line(*c, *e); memcpy(buf, l.c_str(), l.size() + 1); return (int)l.size(); }

// ---- persisted index DB (host_indexdb.cpp) ----
int emu_indexdb_write(const char *path, const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t nT, const uint32_t *keys, int kmerThr, int maskMode) {
    mb200::Matrices m; mb200::default_matrices(m);
    mb200::TargetIndex idx;
    if (mb200::build_index(masked, off, nT, m.seed, kmerThr, idx)) return -1;
    std::string err;
    const int rc = mb200::write_index_db(path, idx, masked, unmasked, off, nT, keys, kmerThr, maskMode, err);
    if (rc) fprintf(stderr, "emu_indexdb_write: %s\n", err.c_str());
    return rc;
}
void *emu_indexdb_open(const char *path, char *errOut, size_t errCap) {
    mb200::IndexDb *db = new mb200::IndexDb(); std::string err;
    if (mb200::open_index_db(path, *db, err)) { if (errOut && errCap) { strncpy(errOut, err.c_str(), errCap - 1); errOut[errCap - 1] = 0; } delete db; return nullptr; }
    return db;
}
void emu_indexdb_info(void *h, uint64_t out[3], int meta[12]) { auto *db = (mb200::IndexDb *)h; out[0] = db->nEntries; out[1] = db->nSeq; out[2] = db->seqOff[db->nSeq]; memcpy(meta, &db->meta, 48); }
const void *emu_indexdb_ptr(void *h, int which) {
    auto *db = (mb200::IndexDb *)h;
    switch (which) { case 0: return db->offsets; case 1: return db->entries6; case 2: return db->masked; case 3: return db->seqOff; case 4: return db->keys.data(); case 5: return db->unmasked.data(); }
    return nullptr;
}
void emu_indexdb_close(void *h) { delete (mb200::IndexDb *)h; }

}  // extern "C"
