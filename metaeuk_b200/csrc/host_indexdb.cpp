// host_indexdb.cpp -- see host_indexdb.h.  Plain C++17; part of libmetaeuk_b200.so and (for CPU-only tests) of tests/emu.
#include "host_indexdb.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace mb200 {

const char NUM2AA[22] = "ACDEFGHIKLMNPQRSTVWYX";

void aa2num_table(uint8_t t[256]) {
    memset(t, 20, 256);
    for (int i = 0; i < 21; i++) { t[(unsigned char)NUM2AA[i]] = (uint8_t)i; t[(unsigned char)(NUM2AA[i] | 0x20)] = (uint8_t)i; }
    const char *alias = "JLZEBD";
    for (int i = 0; i < 6; i += 2) { t[(unsigned char)alias[i]] = t[(unsigned char)alias[i + 1]]; t[(unsigned char)(alias[i] | 0x20)] = t[(unsigned char)alias[i + 1]]; }
}

// entry ids (PrefilteringIndexReader.cpp:10-34)
enum { E_VERSION = 0, E_META = 1, E_SCOREMATRIXNAME = 2, E_SCOREMATRIX2MER = 3, E_SCOREMATRIX3MER = 4, E_SPACEDPATTERN = 23, E_DBR1INDEX = 5, E_DBR1DATA = 6, E_DBR2INDEX = 7, E_DBR2DATA = 8, E_ENTRIES = 9, E_ENTRIESOFFSETS = 10, E_ENTRIESNUM = 12,
       E_SEQCOUNT = 13, E_SEQINDEXDATA = 14, E_SEQINDEXDATASIZE = 15, E_SEQINDEXSEQOFFSET = 16, E_GENERATOR = 22 };
static const char *INDEX_VERSION = "16";   // MMSEQS_CURRENT_INDEX_VERSION (reference/lib/mmseqs/src/MMseqsBase.cpp:6)
static const int DBTYPE_INDEX_DB = 9, DBTYPE_AMINO_ACIDS = 0;   // Parameters.h:68-77
static const size_t REC = 24, HDR = 28;   // DBReader<unsigned int>::Index {uint32 id; size_t offset; uint32 length} with padding; serialize() header (DBReader.cpp:944-972)

void IndexDb::close() {
    if (map) munmap(map, mapSize);
    map = nullptr; mapSize = 0; offsets = nullptr; entries6 = nullptr; masked = nullptr; seqOff = nullptr;
}

static int bad(std::string &err, const std::string &m) { err = m; return -1; }

int open_index_db(const std::string &path, IndexDb &db, std::string &err) {
    db.close(); db.keys.clear(); db.unmasked.clear();
    // .index: "key \t offset \t length \n" in any order
    FILE *fi = fopen((path + ".index").c_str(), "r");
    if (!fi) return bad(err, "cannot open " + path + ".index");
    std::map<unsigned, std::pair<uint64_t, uint64_t>> ent;
    unsigned k; unsigned long long o, l;
    while (fscanf(fi, "%u %llu %llu", &k, &o, &l) == 3) ent[k] = {o, l};
    fclose(fi);
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return bad(err, "cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size == 0) { ::close(fd); return bad(err, "cannot stat " + path); }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (m == MAP_FAILED) return bad(err, "cannot mmap " + path);
    db.map = m; db.mapSize = (size_t)st.st_size;
    const uint8_t *base = (const uint8_t *)m;
    auto get = [&](unsigned key, uint64_t minLen, const uint8_t *&p, uint64_t &len) -> bool {
        auto it = ent.find(key);
        if (it == ent.end() || it->second.first + it->second.second > db.mapSize || it->second.second < minLen) return false;
        p = base + it->second.first; len = it->second.second; return true;
    };
    const uint8_t *p; uint64_t len;
    if (!get(E_VERSION, 1, p, len)) return bad(err, "index DB has no VERSION entry");
    db.version.assign((const char *)p, strnlen((const char *)p, len));
    if (strncmp(db.version.c_str(), INDEX_VERSION, strlen(INDEX_VERSION)) != 0) return bad(err, "index version '" + db.version + "' is not " + INDEX_VERSION);
    if (!get(E_META, sizeof(IndexDbMeta), p, len)) return bad(err, "index DB has no META entry");
    memcpy(&db.meta, p, sizeof(IndexDbMeta));
    if (get(E_GENERATOR, 1, p, len)) db.generator.assign((const char *)p, strnlen((const char *)p, len));
    const IndexDbMeta &M = db.meta;
    if (M.splits != 1) return bad(err, "split index DBs are not supported (the top-300 cut of a target split differs, Prefiltering.cpp:356-361)");
    if (M.kmerSize != 6 || M.alphabetSize != 21 || M.spacedKmer != 1 || (M.seqType & 0xFFFF) != DBTYPE_AMINO_ACIDS)
        return bad(err, "index DB is not an amino-acid index with k = 6, alphabet 21 and the spaced pattern");
    if (!get(E_ENTRIESNUM, 8, p, len)) return bad(err, "index DB has no ENTRIESNUM entry (built without a k-mer index?)");
    memcpy(&db.nEntries, p, 8);
    if (!get(E_ENTRIES, db.nEntries * 6, p, len)) return bad(err, "ENTRIES entry missing or too short");
    db.entries6 = p;
    if (!get(E_ENTRIESOFFSETS, (KMER_TABLE + 1) * 8, p, len)) return bad(err, "ENTRIESOFFSETS entry missing or too short");
    db.offsets = (const uint64_t *)p;
    if (db.offsets[KMER_TABLE] != db.nEntries) return bad(err, "ENTRIESOFFSETS does not end at ENTRIESNUM");
    if (!get(E_SEQCOUNT, 8, p, len)) return bad(err, "SEQCOUNT entry missing");
    memcpy(&db.nSeq, p, 8);
    if (!get(E_SEQINDEXSEQOFFSET, (db.nSeq + 1) * 8, p, len)) return bad(err, "SEQINDEXSEQOFFSET entry missing or too short");
    db.seqOff = (const uint64_t *)p;
    // a corrupt lookup must not turn into out-of-bounds writes below: offsets start at 0 and never decrease; the last one is checked against SEQINDEXDATA next
    if (db.seqOff[0] != 0) return bad(err, "SEQINDEXSEQOFFSET does not start at 0");
    for (uint64_t i = 0; i < db.nSeq; i++) if (db.seqOff[i + 1] < db.seqOff[i]) return bad(err, "SEQINDEXSEQOFFSET is not monotonic at sequence " + std::to_string(i));
    if (!get(E_SEQINDEXDATA, db.seqOff[db.nSeq], p, len)) return bad(err, "SEQINDEXDATA entry missing or too short");
    db.masked = p;
    // the stored sequence DB: keys and unmasked residues
    const uint8_t *ix; uint64_t ixLen; const uint8_t *dat; uint64_t datLen;
    if (!get(E_DBR1INDEX, HDR, ix, ixLen) || !get(E_DBR1DATA, 1, dat, datLen)) return bad(err, "DBR1INDEX / DBR1DATA entry missing");
    uint64_t n; memcpy(&n, ix, 8);
    if (n != db.nSeq || ixLen < HDR + n * REC) return bad(err, "DBR1INDEX does not list SEQCOUNT sequences");
    uint8_t a2n[256]; aa2num_table(a2n);
    db.keys.resize(n); db.unmasked.resize(db.seqOff[n]);
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *r = ix + HDR + i * REC; uint32_t id, length; uint64_t off;
        memcpy(&id, r, 4); memcpy(&off, r + 8, 8); memcpy(&length, r + 16, 4);
        const uint64_t L = db.seqOff[i + 1] - db.seqOff[i];
        if (length < 2 || (uint64_t)length - 2 != L || off + length > datLen) return bad(err, "DBR1 entry " + std::to_string(i) + " does not match the sequence lookup");
        db.keys[i] = id;
        for (uint64_t x = 0; x < L; x++) db.unmasked[db.seqOff[i] + x] = a2n[dat[off + x]];
    }
    return 0;
}

namespace {
struct Writer {
    FILE *f = nullptr; uint64_t off = 0; std::map<unsigned, std::pair<uint64_t, uint64_t>> ent; bool ok = true;
    void raw(const void *p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; off += n; }
    void begin(unsigned key) { ent[key] = {off, 0}; }
    void end(unsigned key) { const char z = 0; raw(&z, 1); ent[key].second = off - ent[key].first; align(); }   // DBWriter::writeEnd appends a NUL; length counts it
    void data(unsigned key, const void *p, size_t n) { begin(key); raw(p, n); end(key); }
    void align() { static const char zeros[4096] = {0}; const uint64_t pad = (4096 - (off & 4095)) & 4095; raw(zeros, pad); }   // DBWriter::alignToPageSize
};
}  // namespace

int write_index_db(const std::string &path, const TargetIndex &idx, const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t nT,
                   const uint32_t *keys, int kmerThr, int maskMode, std::string &err, const int16_t *seedMatrix, int rowAlign) {
    if (idx.offsets.size() != KMER_TABLE + 1) return bad(err, "index not built");
    Writer w; w.f = fopen(path.c_str(), "wb");
    if (!w.f) return bad(err, "cannot create " + path);
    uint32_t maxLen = 0, lastKey = 0;
    for (uint32_t i = 0; i < nT; i++) { maxLen = std::max<uint32_t>(maxLen, (uint32_t)(off[i + 1] - off[i])); lastKey = std::max(lastKey, keys ? keys[i] : i); }
    w.data(E_VERSION, INDEX_VERSION, strlen(INDEX_VERSION));
    IndexDbMeta M{}; M.maxSeqLen = 65535; M.kmerSize = 6; M.biasCorr = 1; M.alphabetSize = 21; M.mask = maskMode > 0; M.spacedKmer = 1; M.kmerThr = kmerThr;
    M.seqType = DBTYPE_AMINO_ACIDS; M.srcSeqType = DBTYPE_AMINO_ACIDS; M.headers1 = 0; M.headers2 = 0; M.splits = 1;
    w.data(E_META, &M, sizeof(M));
    const char *gen = "metaeuk_b200";
    w.data(E_GENERATOR, gen, strlen(gen));
    {   // What the reference's own prefilter needs besides the k-mer index to run from this DB (PrefilteringIndexReader.cpp:91-99,196-215,536-590): the name of the k-mer matrix
        // ("<name>.out:<text>"; Prefiltering.cpp:156 stores it and printSummary prints the name -- the text is never parsed on that path, and the matrix file itself is not part of this
        // library), the empty spaced pattern (= the default 1101010011), and the 2-mer / 3-mer similarity tables in ScoreMatrix::serialize layout (ScoreMatrix.h:26-41): int16 score[n][row]
        // then uint32 index[n][row], neighbours in descending score order, row = n rounded up to the next multiple of the reader's MAX_ALIGN_INT (64 in AVX2 builds, 32 in SSE builds:
        // the reader recomputes it, so an index is only readable by builds of the width it was written for), padded with -255 / 0.
        static const char *name = "VTML80.out:# VTML80, bit factor 8, score bias -0.2: the k-mer similarity matrix of this index (its 2-mer / 3-mer tables are entries 3 and 4).\n"
                                  "# Written by metaeuk_b200 gpucreateindex, which does not embed the matrix file.\n";
        w.data(E_SCOREMATRIXNAME, name, strlen(name));
        w.data(E_SPACEDPATTERN, "", 0);
        Matrices mats; default_matrices(mats);
        const int16_t *seed = seedMatrix ? seedMatrix : mats.seed;
        const size_t align = rowAlign > 0 ? (size_t)rowAlign : 64;
        auto writeTable = [&](unsigned key, size_t n, const std::vector<uint32_t> &tab) {   // tab[n][n] = (neighbour << 16) | (uint16) score
            const size_t row = (n / align + 1) * align;
            std::vector<int16_t> sc(row, (int16_t)-255); std::vector<uint32_t> ix(row, 0u);
            w.begin(key);
            for (size_t r = 0; r < n; r++) { for (size_t j = 0; j < n; j++) sc[j] = (int16_t)(uint16_t)(tab[r * n + j] & 0xFFFFu); w.raw(sc.data(), row * 2); }
            for (size_t r = 0; r < n; r++) { for (size_t j = 0; j < n; j++) ix[j] = tab[r * n + j] >> 16; w.raw(ix.data(), row * 4); }
            w.end(key);
        };
        {
            std::vector<uint32_t> sim3; std::vector<uint16_t> cnt3;
            build_sim3(seed, sim3, cnt3, -192, 384);
            writeTable(E_SCOREMATRIX3MER, (size_t)N3, sim3);
        }
        {   // 2-mers: the same stable descending order over the cartesian enumeration (first letter outermost, lowest digit)
            std::vector<uint32_t> sim2((size_t)400 * 400);
            for (int r = 0; r < 400; r++) {
                const int a0 = r % 20, a1 = r / 20;
                std::vector<std::pair<int, uint32_t>> v; v.reserve(400);
                for (int b0 = 0; b0 < 20; b0++) for (int b1 = 0; b1 < 20; b1++) v.push_back({seed[a0 * AL + b0] + seed[a1 * AL + b1], (uint32_t)(b0 + 20 * b1)});
                std::stable_sort(v.begin(), v.end(), [](const std::pair<int, uint32_t> &x, const std::pair<int, uint32_t> &y) { return x.first > y.first; });
                for (int j = 0; j < 400; j++) sim2[(size_t)r * 400 + j] = (v[j].second << 16) | (uint16_t)(int16_t)v[j].first;
            }
            writeTable(E_SCOREMATRIX2MER, 400, sim2);
        }
    }
    {   // DBR1INDEX: DBReader<unsigned int>::serialize layout
        const uint64_t total = off[nT]; const uint64_t dataSize = total + 2ull * nT;
        std::vector<uint8_t> b(HDR + (size_t)nT * REC, 0);
        const uint64_t n64 = nT; const int dbtype = DBTYPE_AMINO_ACIDS;
        memcpy(&b[0], &n64, 8); memcpy(&b[8], &dataSize, 8); memcpy(&b[16], &lastKey, 4); memcpy(&b[20], &dbtype, 4); const uint32_t maxSeqLen = maxLen + 2; memcpy(&b[24], &maxSeqLen, 4);   // DBReader counts the trailing "\\n\\0"
        uint64_t o = 0;
        for (uint32_t i = 0; i < nT; i++) {
            uint8_t *r = &b[HDR + (size_t)i * REC]; const uint32_t id = keys ? keys[i] : i, length = (uint32_t)(off[i + 1] - off[i]) + 2;
            memcpy(r, &id, 4); memcpy(r + 8, &o, 8); memcpy(r + 16, &length, 4);
            o += length;
        }
        w.data(E_DBR1INDEX, b.data(), b.size());
        // DBR1DATA: "residues \n \0" per sequence
        w.begin(E_DBR1DATA);
        std::vector<char> line;
        for (uint32_t i = 0; i < nT; i++) {
            const uint64_t L = off[i + 1] - off[i]; line.resize(L + 2);
            for (uint64_t x = 0; x < L; x++) line[x] = NUM2AA[std::min<uint8_t>(unmasked[off[i] + x], 20)];
            line[L] = '\n'; line[L + 1] = '\0';
            w.raw(line.data(), line.size());
        }
        w.end(E_DBR1DATA);
        // a DB without a source DB: DBR2 entries alias DBR1 with length + 1 (PrefilteringIndexReader.cpp:128-130)
        w.ent[E_DBR2INDEX] = {w.ent[E_DBR1INDEX].first, b.size() + 1};
        w.ent[E_DBR2DATA] = {w.ent[E_DBR1DATA].first, dataSize + 1};
    }
    {   // ENTRIES: packed 6-byte IndexEntryLocal; ENTRIESOFFSETS: size_t[20^6 + 1]
        w.begin(E_ENTRIES);
        std::vector<uint8_t> chunk; chunk.reserve(6u << 16);
        for (size_t i = 0; i < idx.entries.size(); i++) {
            const uint32_t s = (uint32_t)idx.entries[i]; const uint16_t p = (uint16_t)(idx.entries[i] >> 32); uint8_t e[6]; memcpy(e, &s, 4); memcpy(e + 4, &p, 2);
            chunk.insert(chunk.end(), e, e + 6);
            if (chunk.size() >= (6u << 16)) { w.raw(chunk.data(), chunk.size()); chunk.clear(); }
        }
        w.raw(chunk.data(), chunk.size());
        w.end(E_ENTRIES);
        w.begin(E_ENTRIESOFFSETS);
        std::vector<uint64_t> o64(1 << 16);
        for (size_t k = 0; k <= KMER_TABLE; k += o64.size()) {
            const size_t n = std::min(o64.size(), KMER_TABLE + 1 - k);
            for (size_t x = 0; x < n; x++) o64[x] = idx.offsets[k + x];
            w.raw(o64.data(), n * 8);
        }
        w.end(E_ENTRIESOFFSETS);
        const uint64_t ne = idx.entries.size();
        w.data(E_ENTRIESNUM, &ne, 8);
    }
    const uint64_t cnt = nT; const int64_t dsz = (int64_t)off[nT];
    w.data(E_SEQCOUNT, &cnt, 8);
    w.data(E_SEQINDEXDATASIZE, &dsz, 8);
    w.data(E_SEQINDEXSEQOFFSET, off, (size_t)(nT + 1) * 8);
    w.begin(E_SEQINDEXDATA); w.raw(masked, off[nT]); { const char z = 0; w.raw(&z, 1); } w.end(E_SEQINDEXDATA);   // getDataSize() + 1 bytes, then the writer's NUL
    const bool ok = w.ok && fclose(w.f) == 0;
    if (!ok) return bad(err, "write error on " + path);
    FILE *fi = fopen((path + ".index").c_str(), "w");
    if (!fi) return bad(err, "cannot create " + path + ".index");
    for (auto &kv : w.ent) fprintf(fi, "%u\t%llu\t%llu\n", kv.first, (unsigned long long)kv.second.first, (unsigned long long)kv.second.second);
    if (fclose(fi) != 0) return bad(err, "write error on " + path + ".index");
    FILE *ft = fopen((path + ".dbtype").c_str(), "wb");
    if (!ft) return bad(err, "cannot create " + path + ".dbtype");
    const int t = DBTYPE_INDEX_DB; fwrite(&t, 4, 1, ft);
    if (fclose(ft) != 0) return bad(err, "write error on " + path + ".dbtype");
    return 0;
}

}  // namespace mb200
