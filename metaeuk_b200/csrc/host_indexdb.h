// host_indexdb.h -- the persisted target index, plain C++ (no CUDA).
//
// The reference persists its prefilter index with `createindex` / `indexdb` as an MMseqs database `<targetDB>.idx` (+ `.idx.index`,
// `.idx.dbtype` = DBTYPE_INDEX_DB) whose entries are keyed by the constants of PrefilteringIndexReader
// (reference/lib/mmseqs/src/prefiltering/PrefilteringIndexReader.cpp:10-34, written at :60-307, read at :330-520).  This file reads
// that layout (so an index built once by the reference -- or by write_index_db below -- is mmapped and uploaded on later runs instead of
// being rebuilt) and writes it with the entries the GPU path needs:
//   VERSION(0) META(1) GENERATOR(22) DBR1INDEX(5) DBR1DATA(6) DBR2INDEX(7)/DBR2DATA(8) (aliases of DBR1, as the reference writes them for a
//   DB without a source DB) ENTRIES(9) ENTRIESOFFSETS(10) ENTRIESNUM(12) SEQCOUNT(13) SEQINDEXDATA(14) SEQINDEXDATASIZE(15) SEQINDEXSEQOFFSET(16)
// ENTRIES / ENTRIESOFFSETS / SEQINDEXDATA / SEQINDEXSEQOFFSET are byte-identical to the reference's entries for the same target DB and
// parameters (the k-mer index digests of tests/golden are taken over exactly these bytes).  Not written: the 2-mer/3-mer similarity tables
// (SCOREMATRIX2MER/3MER, 386 MB that depend only on the seed matrix and are rebuilt in about a second), SCOREMATRIXNAME, headers.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "host_tables.h"

namespace mb200 {

struct IndexDbMeta { int maxSeqLen, kmerSize, biasCorr, alphabetSize, mask, spacedKmer, kmerThr, seqType, srcSeqType, headers1, headers2, splits; };

// An opened (mmapped) index DB.  Pointers stay valid until close().
struct IndexDb {
    void *map = nullptr; size_t mapSize = 0;
    IndexDbMeta meta{};
    std::string version, generator;
    const uint64_t *offsets = nullptr;      // [20^6 + 1]   (ENTRIESOFFSETS)
    const uint8_t *entries6 = nullptr;      // packed IndexEntryLocal {uint32 seqId; uint16 position_j}  (ENTRIES)
    uint64_t nEntries = 0;
    const uint8_t *masked = nullptr;        // numeric residue codes after masking (SEQINDEXDATA)
    const uint64_t *seqOff = nullptr;       // [nSeq + 1]  (SEQINDEXSEQOFFSET)
    uint64_t nSeq = 0;
    std::vector<uint32_t> keys;             // DB key of local id i (DBR1INDEX order)
    std::vector<uint8_t> unmasked;          // numeric codes of the stored sequences (DBR1DATA through aa2num), same offsets as seqOff
    void close();
    ~IndexDb() { close(); }
};

// 0 on success; on failure returns -1 and sets err (missing entry, wrong version, split index, profile/nucleotide DB, inconsistent sizes)
int open_index_db(const std::string &path, IndexDb &db, std::string &err);

// targets as numeric codes (masked / unmasked share off[nT + 1]); keys[nT] or NULL (= local ids); idx = build_index() of the masked residues.
// Also writes what the reference's own prefilter needs to run from this DB: SCOREMATRIXNAME, SPACEDPATTERN and the 2-mer / 3-mer similarity tables of seedMatrix (NULL = the default
// VTML80 x 8) with rows padded to rowAlign int16 (0 = 64, the MAX_ALIGN_INT of AVX2 builds of the reference; 32 for SSE builds).
int write_index_db(const std::string &path, const TargetIndex &idx, const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t nT,
                   const uint32_t *keys, int kmerThr, int maskMode, std::string &err, const int16_t *seedMatrix = nullptr, int rowAlign = 0);

extern const char NUM2AA[22];               // "ACDEFGHIKLMNPQRSTVWYX"
void aa2num_table(uint8_t table[256]);      // SubstitutionMatrix::setupLetterMapping (SubstitutionMatrix.cpp:262-304)

}  // namespace mb200
