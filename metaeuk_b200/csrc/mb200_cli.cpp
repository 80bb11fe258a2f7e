// mb200_cli.cpp -- C++ host driver with MetaEuk's module surface over libmetaeuk_b200 (C ABI only).
//
//   mb200 gpuextractorfs       <contigsDB> <orfHeaderDB> <orfAaDB>      == extractorfs + translatenucs (M/src/util/extractorfs.cpp:19, translatenucs.cpp:13)
//   mb200 gpuprefilter         <queryDB> <targetDB> <prefDB>            == prefilter          (M/src/prefiltering/Main.cpp:13)
//   mb200 gpualign             <queryDB> <targetDB> <prefDB> <alnDB>    == align              (M/src/alignment/Main.cpp:12)
//   mb200 gpusearch            <queryDB> <targetDB> <alnDB>             == search (blastp.sh:70,85 without the pref DB round trip)
//   mb200 gpuresultspercontig  <contigsDB> <orfDB> <alnDB> <byContigDB> == resultspercontig   (src/exonpredictor/resultspercontig.cpp:34)
//   mb200 gpucollectoptimalset <byContigDB> <targetDB> <predDB>         == collectoptimalset  (src/exonpredictor/collectoptimalset.cpp:224)
//   mb200 gpureduceredundancy  <calledExonsDB> <predsDB> <predToCallDB> == reduceredundancy   (src/exonpredictor/reduceredundancy.cpp:162)
//   mb200 gpupredictexons      <contigsDB> <targetDB> <predDB>          == predictexons.sh:40-82 with the contigs as the only input (ORFs stay on the GPU)
//   mb200 gpupredictexons      <contigsDB> <orfAaDB> <orfNuclDB> <targetDB> <predDB>
//   mb200 gpucreateindex       <targetDB> <indexDB>                     == createindex / indexdb (M/src/util/indexdb.cpp); every search module takes --index-db <indexDB>
//                                                   == predictexons.sh:68-82 (search -> resultspercontig -> collectoptimalset)
// DBs are MMseqs databases: data file(s) name[.N], name.index "key\toffset\tlength", name.dbtype.
// Result DBs are written byte-compatible with the reference's (same records, same keys; one data file).
// Options: -s X (sensitivity -> k-mer threshold as Prefiltering::getKmerThreshold)  --k-score N  --max-seqs N  --min-ungapped-score N  -e X  --min-aln-len N
// --comp-bias-corr 0|1  --min-length N --max-length N --translation-table N (ORF extraction)  --reverse-fragments 0|1 (null model)  --mask 0|1 (default 1: tantan masking of the targets as the reference's
// index builder does it, --mask-prob 0.9)  --masked-db <db> (take the masked target residues from a DB instead: lower case or X = masked)  --metaeuk-eval X
// --metaeuk-tcov X --max-intron N --min-intron N --min-exon-aa N --max-overlap N --max-exon-sets N --set-gap-open N --set-gap-extend N
// --device N  --gpus N (gpupredictexons <contigsDB> <targetDB> <predDB>: one context per device N .. N + gpus - 1 in this process, index copied device to device, contigs sharded by
// nucleotide count, the merged DB equals the one-GPU DB).  Everything numeric happens in the library (GPU); this file is DB I/O and the reference's text formatting.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/metaeuk_b200.h"

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(EXIT_FAILURE); } while (0)
#define MB(x) do { if ((x) != MB200_OK) DIE("mb200: %s", mb200_last_error()); } while (0)

// ---------------------------------------------------------------- MMseqs DB I/O
struct Db {
    std::vector<char> data; std::vector<uint32_t> keys; std::vector<uint64_t> off, len;   // entries in data-offset order (= local id)
    int dbtype = 0;
    std::map<uint32_t, size_t> byKey;
    const char *entry(size_t id) const { return data.data() + off[id]; }
    size_t find(uint32_t key) const { auto it = byKey.find(key); return it == byKey.end() ? (size_t)-1 : it->second; }
};
static bool readFile(const std::string &p, std::vector<char> &out, bool append) {
    FILE *f = fopen(p.c_str(), "rb"); if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    size_t o = append ? out.size() : 0; out.resize(o + n);
    if (n && fread(out.data() + o, 1, n, f) != (size_t)n) DIE("short read on %s", p.c_str());
    fclose(f); return true;
}
static Db openDb(const std::string &name) {
    Db d;
    if (!readFile(name, d.data, false)) {   // split data files name.0, name.1, ... concatenated in order (DBReader)
        int i = 0; bool any = false;
        while (readFile(name + "." + std::to_string(i), d.data, true)) { any = true; i++; }
        if (!any) DIE("cannot open database %s", name.c_str());
    }
    std::vector<char> idx; if (!readFile(name + ".index", idx, false)) DIE("cannot open %s.index", name.c_str());
    struct E { uint64_t off, len; uint32_t key; };
    std::vector<E> es;
    const char *p = idx.data(), *end = p + idx.size();
    while (p < end) {
        char *q; E e; e.key = (uint32_t)strtoul(p, &q, 10); e.off = strtoull(q, &q, 10); e.len = strtoull(q, &q, 10);
        es.push_back(e); while (q < end && *q != '\n') q++; p = q + 1;
    }
    // local ids = key rank, as DBReader::sortIndex assigns them (M/src/commons/DBReader.cpp): the target ids in the index -- and with them the tie-breaks at the --max-seqs cut --
    // must not depend on where an entry sits in the data file (createsubdb / merged multi-thread output are not in offset order)
    std::stable_sort(es.begin(), es.end(), [](const E &a, const E &b) { return a.key < b.key; });
    for (auto &e : es) { d.byKey[e.key] = d.keys.size(); d.keys.push_back(e.key); d.off.push_back(e.off); d.len.push_back(e.len); }
    std::vector<char> t; if (readFile(name + ".dbtype", t, false) && t.size() >= 4) memcpy(&d.dbtype, t.data(), 4);
    return d;
}
struct DbOut {
    FILE *f, *fi; uint64_t pos = 0; std::string name;
    DbOut(const std::string &n, int dbtype) : name(n) {
        f = fopen(n.c_str(), "wb"); fi = fopen((n + ".index").c_str(), "wb"); if (!f || !fi) DIE("cannot write %s", n.c_str());
        FILE *t = fopen((n + ".dbtype").c_str(), "wb"); fwrite(&dbtype, 4, 1, t); fclose(t);
    }
    void write(uint32_t key, const std::string &s) { fwrite(s.data(), 1, s.size(), f); fputc('\0', f); fprintf(fi, "%u\t%llu\t%zu\n", key, (unsigned long long)pos, s.size() + 1); pos += s.size() + 1; }
    ~DbOut() { fclose(f); fclose(fi); }
};

// ---------------------------------------------------------------- residue codes (SubstitutionMatrix::setupLetterMapping, SubstitutionMatrix.cpp:262-304)
static uint8_t AA2NUM[256];
static void initCodes() {
    const char *order = "ACDEFGHIKLMNPQRSTVWYX";
    for (int i = 0; i < 256; i++) AA2NUM[i] = 20;
    for (int i = 0; order[i]; i++) { AA2NUM[(int)order[i]] = i; AA2NUM[tolower(order[i])] = i; }
    const char *al[3][2] = {{"J", "L"}, {"Z", "E"}, {"B", "D"}};
    for (auto &a : al) { AA2NUM[(int)a[0][0]] = AA2NUM[(int)a[1][0]]; AA2NUM[tolower(a[0][0])] = AA2NUM[(int)a[1][0]]; }
}
struct Seqs { std::vector<uint8_t> res; std::vector<uint64_t> off; };
static Seqs encodeDb(const Db &d, bool lowerIsMasked) {
    Seqs s; s.off.push_back(0);
    for (size_t i = 0; i < d.keys.size(); i++) {
        const char *e = d.entry(i); size_t n = d.len[i] >= 2 ? d.len[i] - 2 : 0;   // "residues\n\0" (DBReader.h:211-231)
        for (size_t k = 0; k < n; k++) s.res.push_back((lowerIsMasked && islower((unsigned char)e[k])) ? 20 : AA2NUM[(unsigned char)e[k]]);
        s.off.push_back(s.res.size());
    }
    return s;
}

// ---------------------------------------------------------------- the reference's text formatting
static std::string fmtSeqId(float x, bool dropLastZero) {   // Util::fastSeqIdToBuffer (Util.cpp:222-251); resultToBuffer overwrites the last 0 of "1.000"
    char b[32];
    if (x == 1.0f) return dropLastZero ? "1.00" : "1.000";
    std::string s = "0."; if (x < 0.10) s += "0"; if (x < 0.01) s += "0";
    snprintf(b, sizeof(b), "%d", (int)(x * 1000)); return s + b;
}
static std::string fmtE(double e) { char b[64]; snprintf(b, sizeof(b), "%.3E", e); return b; }

struct Opt {
    mb200_params p; std::string maskedDb; int device = 0, gpus = 1; int overlap = 0; int textStages = 0; long long batchNt = 0; std::string indexDb; int mask = 1; double maskProb = 0.9;
    int maskLowerCase = 0, maskNRepeat = 0;   // --mask-lower-case, --mask-n-repeat (Masker::maskSequence, Masker.cpp:14-58)
    int minLen = 15, maxLen = 32734;   // --min-length / --max-length of extractorfs in codons (setPredictExonsDefaults, src/workflow/PredictExons.cpp:8-16)
};

// Flags of the reference's modules that this driver takes.  Three groups: flags that change what the path computes and are implemented (parseOpts); flags that
// never change a result (threads, verbosity, ...), accepted and ignored; flags of behaviour that is NOT implemented, accepted only with the value `metaeuk
// predictexons` itself passes -- anything else stops the run instead of silently computing something different.  Unknown flags stop it too.
static const char *IGNORED_FLAGS[] = {"--threads", "-v", "--compressed", "--remove-tmp-files", "--db-load-mode", "--split-memory-limit", "--force-reuse", "--mpi-runner",
                                      "--local-tmp", "--disk-space-limit", "--pca", "--pcb", "--zdrop", "--strand", "--realign-score-bias", "--realign-max-seqs",
                                      "--e-profile", "--mask-profile", "--filter-msa", "--max-seq-id", "--qid", "--qsc", "--cov", "--diff", "--pseudo-cnt-mode", "--wg", nullptr};
struct FixedFlag { const char *name; const char *ok[4]; };
static const FixedFlag FIXED_FLAGS[] = {
    {"--alignment-mode", {"2"}}, {"--alignment-output-mode", {"0"}}, {"-a", {"0"}}, {"--gap-open", {"11", "aa:11,nucl:5"}}, {"--gap-extend", {"1", "aa:1,nucl:2"}},
    {"--sub-mat", {"blosum62.out", "aa:blosum62.out,nucl:nucleotide.out"}}, {"--seed-sub-mat", {"VTML80.out", "aa:VTML80.out,nucl:nucleotide.out"}},
    {"-k", {"0", "6"}}, {"--spaced-kmer-mode", {"1"}}, {"--alph-size", {"21", "aa:21,nucl:5"}}, {"--max-seq-len", {"65535"}}, {"--split", {"0", "1"}}, {"--split-mode", {"2"}},
    {"-c", {"0"}}, {"--cov-mode", {"0"}}, {"--comp-bias-corr-scale", {"1"}}, {"--diag-score", {"1"}}, {"--exact-kmer-matching", {"0"}},
    {"--add-self-matches", {"0"}}, {"--min-seq-id", {"0"}}, {"--seq-id-mode", {"0"}}, {"--alt-ali", {"0"}}, {"--max-rejected", {"2147483647"}},
    {"--max-accept", {"2147483647"}}, {"--score-bias", {"0"}}, {"--realign", {"0"}}, {"--wrapped-scoring", {"0"}}, {"--corr-score-weight", {"0"}},
    {"--orf-start-mode", {"1"}}, {"--contig-start-mode", {"2"}}, {"--contig-end-mode", {"2"}}, {"--forward-frames", {"1,2,3"}}, {"--reverse-frames", {"1,2,3"}},
    {"--max-gaps", {"2147483647"}}, {"--translate", {"0"}}, {"--use-all-table-starts", {"0"}}, {"--id-offset", {"0"}}, {"--create-lookup", {"0"}}, {"--add-orf-stop", {"0"}},
    {"--num-iterations", {"1"}}, {"--exhaustive-search", {"0"}}, {"--target-search-mode", {"0"}}, {"--gpu", {"0"}}, {"--gpu-server", {"0"}}, {"--prefilter-mode", {"0"}},
    {"--sens-steps", {"1"}}, {"--lca-search", {"0"}}, {"--slice-search", {"0"}}, {"--rescore-mode", {"0"}}, {"--chain-alignments", {"0"}}, {"--merge-query", {"1"}},
    {"--taxon-list", {""}}, {"--spaced-kmer-pattern", {""}}, {nullptr, {}}};
static bool sameValue(const std::string &v, const char *ok) {
    if (v == ok) return true;
    char *e1, *e2; const double a = strtod(v.c_str(), &e1), b = strtod(ok, &e2);
    return *v.c_str() && *ok && !*e1 && !*e2 && a == b;
}
static int seqValue(const std::string &v) {   // "140" or the reference's printed form "seq:140,prof:140" (MultiParam)
    const size_t p = v.find("seq:"); return atoi(v.c_str() + (p == std::string::npos ? 0 : p + 4));
}

static Opt parseOpts(int argc, char **argv, int first, std::vector<std::string> &pos) {
    Opt o; mb200_params_default(&o.p);
    float sens = 4.0f; int kScore = 2147483647;     // -s 4: MetaEuk's default; --k-score INT_MAX = derive the threshold from -s
    for (int i = first; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { if (i + 1 >= argc) DIE("missing value for %s", a.c_str()); return std::string(argv[++i]); };
        if (a == "--max-seqs") o.p.max_res_list_len = atoi(val().c_str());
        else if (a == "--min-ungapped-score") o.p.min_diag_score = atoi(val().c_str());
        else if (a == "--k-score") kScore = seqValue(val());
        else if (a == "-s") sens = strtof(val().c_str(), nullptr);
        else if (a == "-e") o.p.eval_thr = atof(val().c_str());
        else if (a == "--min-aln-len") o.p.min_aln_len = atoi(val().c_str());
        else if (a == "--min-length") o.minLen = atoi(val().c_str());
        else if (a == "--max-length") o.maxLen = atoi(val().c_str());
        else if (a == "--comp-bias-corr") o.p.comp_bias_corr = atoi(val().c_str());
        else if (a == "--reverse-fragments") o.p.reverse_fragments = atoi(val().c_str());
        else if (a == "--translation-table") o.p.translation_table = atoi(val().c_str());
        else if (a == "--masked-db") o.maskedDb = val();
        else if (a == "--metaeuk-eval") o.p.metaeuk_eval = atof(val().c_str());
        else if (a == "--metaeuk-tcov") o.p.tcov = atof(val().c_str());
        else if (a == "--max-intron") o.p.max_intron = atoi(val().c_str());
        else if (a == "--min-intron") o.p.min_intron = atoi(val().c_str());
        else if (a == "--min-exon-aa") o.p.min_exon_aa = atoi(val().c_str());
        else if (a == "--max-overlap") o.p.max_overlap = atoi(val().c_str());
        else if (a == "--max-exon-sets") o.p.max_exon_sets = atoi(val().c_str());
        else if (a == "--set-gap-open") o.p.set_gap_open = atoi(val().c_str());
        else if (a == "--set-gap-extend") o.p.set_gap_extend = atoi(val().c_str());
        else if (a == "--device") o.device = atoi(val().c_str());
        else if (a == "--gpus") { o.gpus = atoi(val().c_str()); if (o.gpus < 1 || o.gpus > 64) DIE("--gpus: 1..64"); }
        else if (a == "--overlap") o.overlap = atoi(val().c_str());
        else if (a == "--index-db") o.indexDb = val();
        else if (a == "--mask") o.mask = atoi(val().c_str());
        else if (a == "--mask-prob") o.maskProb = atof(val().c_str());
        else if (a == "--mask-lower-case") o.maskLowerCase = atoi(val().c_str());
        else if (a == "--mask-n-repeat") o.maskNRepeat = atoi(val().c_str());
        else if (a == "--text-stages") o.textStages = atoi(val().c_str());
        else if (a == "--batch-nt") o.batchNt = atoll(val().c_str());
        else if (a.size() > 1 && a[0] == '-' && !isdigit((unsigned char)a[1])) {
            bool known = false;
            for (int k = 0; IGNORED_FLAGS[k] && !known; k++) if (a == IGNORED_FLAGS[k]) { val(); known = true; }
            for (int k = 0; FIXED_FLAGS[k].name && !known; k++) if (a == FIXED_FLAGS[k].name) {
                const std::string v = val(); bool ok = false;
                for (const char *alt : FIXED_FLAGS[k].ok) if (alt && sameValue(v, alt)) ok = true;
                if (!ok) DIE("%s %s: only the value `metaeuk predictexons` uses (%s) is implemented on the GPU path", a.c_str(), v.c_str(), FIXED_FLAGS[k].ok[0]);
                known = true;
            }
            if (!known) DIE("unknown option %s", a.c_str());
        }
        else pos.push_back(a);
    }
    o.p.kmer_thr = kScore != 2147483647 ? kScore : mb200_kmer_threshold(sens);     // Prefiltering::getKmerThreshold (Prefiltering.cpp:1005-1011)
    return o;
}

// the masked target residues the index is built from, as IndexBuilder::fillDatabase gets them from Masker::maskSequence (IndexBuilder.cpp:84-131)
static void maskTargets(const Opt &o, const Db &tdb, const Seqs &t, Seqs &m) {
    if (!o.mask && !o.maskLowerCase && !o.maskNRepeat) return;                        // m == t
    std::vector<uint8_t> lower;
    if (o.maskLowerCase) {
        lower.reserve(t.res.size());
        for (size_t i = 0; i < tdb.keys.size(); i++) { const char *e = tdb.entry(i); const size_t n = tdb.len[i] >= 2 ? tdb.len[i] - 2 : 0; for (size_t k = 0; k < n; k++) lower.push_back(islower((unsigned char)e[k]) ? 1 : 0); }
    }
    MB(mb200_mask_targets_ex(t.res.data(), t.off.data(), (uint32_t)tdb.keys.size(), o.mask, o.maskProb, 0, o.maskLowerCase ? lower.data() : nullptr, o.maskNRepeat, m.res.data(), nullptr));
}

struct Session {
    mb200_ctx *ctx = nullptr; Db tdb; Seqs t; std::vector<uint64_t> tlen;
    void open(const Opt &o, const std::string &targetDb) {
        MB(mb200_ctx_create(o.device, &ctx)); MB(mb200_params_set(ctx, &o.p)); MB(mb200_matrices_default(ctx));
        tdb = openDb(targetDb); t = encodeDb(tdb, false);
        Seqs m = t;
        if (o.maskedDb.empty() && o.indexDb.empty()) maskTargets(o, tdb, t, m);   // --mask 1 (the reference's default): tantan masking on the host before indexing
        if (!o.maskedDb.empty()) {
            Db md = openDb(o.maskedDb); Seqs raw = encodeDb(md, true);
            for (size_t i = 0; i < tdb.keys.size(); i++) {   // align by key
                size_t j = md.find(tdb.keys[i]); if (j == (size_t)-1 || raw.off[j + 1] - raw.off[j] != t.off[i + 1] - t.off[i]) DIE("masked DB does not match target DB at key %u", tdb.keys[i]);
                memcpy(&m.res[t.off[i]], &raw.res[raw.off[j]], t.off[i + 1] - t.off[i]);
            }
        }
        if (!o.indexDb.empty()) {
            // persisted index (written by gpucreateindex or by the reference's createindex / indexdb): mmap + upload instead of a rebuild
            MB(mb200_index_load(ctx, o.indexDb.c_str()));
            uint64_t sz[4]; MB(mb200_index_sizes(ctx, sz));
            if (sz[1] != tdb.keys.size() || sz[2] != t.res.size()) DIE("index DB %s does not belong to target DB %s", o.indexDb.c_str(), targetDb.c_str());
            std::vector<uint32_t> k(tdb.keys.size() + 1); MB(mb200_target_keys_get(ctx, k.data()));
            for (size_t i = 0; i < tdb.keys.size(); i++) if (k[i] != tdb.keys[i]) DIE("index DB %s lists the targets in another order than %s", o.indexDb.c_str(), targetDb.c_str());
        } else {
            MB(mb200_index_build(ctx, m.res.data(), t.res.data(), t.off.data(), (uint32_t)tdb.keys.size()));
        }
        for (size_t i = 0; i < tdb.keys.size(); i++) tlen.push_back(t.off[i + 1] - t.off[i]);
    }
    ~Session() { if (ctx) mb200_ctx_destroy(ctx); }
};

static std::string prefEntry(const Session &s, const mb200_hit *h, uint64_t n) {   // QueryMatcher::prefilterHitToBuffer (QueryMatcher.h:118-130)
    std::string e; char b[64];
    for (uint64_t k = 0; k < n; k++) { snprintf(b, sizeof(b), "%u\t%d\t%d\n", s.tdb.keys[h[k].seq_id], h[k].score, h[k].diag); e += b; }
    return e;
}
static std::string alnEntry(const Session &s, std::vector<mb200_aln> v, uint32_t qLen) {   // Matcher::compareHits + resultToBuffer (Matcher.h:157-168, Matcher.cpp:280-327)
    v.erase(std::remove_if(v.begin(), v.end(), [](const mb200_aln &a) { return !a.passed; }), v.end());
    std::sort(v.begin(), v.end(), [&](const mb200_aln &a, const mb200_aln &b) {
        if (a.evalue != b.evalue) return a.evalue < b.evalue; if (a.bit_score != b.bit_score) return a.bit_score > b.bit_score;
        if (s.tlen[a.seq_id] != s.tlen[b.seq_id]) return s.tlen[a.seq_id] < s.tlen[b.seq_id]; return s.tdb.keys[a.seq_id] < s.tdb.keys[b.seq_id]; });
    std::string e; char b[256];
    for (auto &a : v) {
        snprintf(b, sizeof(b), "%u\t%d\t%s\t%s\t%d\t%d\t%u\t%d\t%d\t%llu\n", s.tdb.keys[a.seq_id], a.bit_score, fmtSeqId(a.seq_id_est, true).c_str(), fmtE(a.evalue).c_str(),
                 a.q_start, a.q_end, qLen, a.db_start, a.db_end, (unsigned long long)s.tlen[a.seq_id]);
        e += b;
    }
    return e;
}

static std::vector<std::string> splitTab(const std::string &l) { std::vector<std::string> w; size_t a = 0; for (;;) { size_t b = l.find('\t', a); if (b == std::string::npos) { w.push_back(l.substr(a)); break; } w.push_back(l.substr(a, b - a)); a = b + 1; } return w; }
static std::vector<std::string> lines(const char *p) { std::vector<std::string> v; while (*p) { const char *q = strchr(p, '\n'); if (!q) { v.push_back(p); break; } v.push_back(std::string(p, q - p)); p = q + 1; } return v; }

// resultspercontig (resultspercontig.cpp:64-196): ORF header "contigKey\tfrom+len" / "from-len" (Orf::parseOrfHeader, Orf.cpp:351-438)
static std::map<uint32_t, std::string> byContig(const Db &contigs, const Db &orfHeaders, const std::map<uint32_t, std::string> &aln) {
    struct Row { uint32_t t, orf; std::string line; };
    std::map<uint32_t, std::vector<Row>> per;
    for (size_t i = 0; i < contigs.keys.size(); i++) per[contigs.keys[i]];
    for (auto &kv : aln) {
        if (kv.second.empty()) continue;
        size_t h = orfHeaders.find(kv.first); if (h == (size_t)-1) continue;
        std::string hdr(orfHeaders.entry(h)); char *q; uint32_t contig = (uint32_t)strtoul(hdr.c_str(), &q, 10);
        long from = strtol(q, &q, 10); char sign = *q; long len = strtol(q + 1, &q, 10); long to = sign == '+' ? from + len : from - len;
        size_t c = contigs.find(contig); if (c == (size_t)-1) continue;
        long orfLen = std::max(from, to) - std::min(from, to) + 1; unsigned long long contigLen = contigs.len[c] >= 2 ? contigs.len[c] - 2 : 0;
        for (auto &l : lines(kv.second.c_str())) {
            auto w = splitTab(l); w[2] = fmtSeqId(strtof(w[2].c_str(), nullptr), true);   // parse + re-print
            std::string line; for (size_t x = 0; x < w.size(); x++) { line += w[x]; line += '\t'; }
            char b[256]; snprintf(b, sizeof(b), "%u\t1\t1.00\t0.000E+00\t0\t%ld\t%ld\t%ld\t%ld\t%llu\n", kv.first, orfLen - 1, orfLen, from, to, contigLen);
            per[contig].push_back({(uint32_t)strtoul(w[0].c_str(), nullptr, 10), kv.first, line + b});
        }
    }
    std::map<uint32_t, std::string> out;
    for (auto &kv : per) { auto &v = kv.second; std::stable_sort(v.begin(), v.end(), [](const Row &a, const Row &b) { return a.t != b.t ? a.t < b.t : a.orf < b.orf; }); std::string e; for (auto &r : v) e += r.line; out[kv.first] = e; }
    return out;
}

// collectoptimalset on {contig: 20-column lines} (collectoptimalset.cpp:273-418 ; Prediction::predictionToBuffer PredictionParser.h:368-396)
static std::map<uint32_t, std::string> chainAll(mb200_ctx *ctx, const std::map<uint32_t, std::string> &byc, uint64_t totalRes) {
    std::vector<mb200_cand> cands; std::vector<uint64_t> off(1, 0); std::vector<std::pair<std::string, std::string>> txt; std::vector<uint32_t> keys;
    for (auto &kv : byc) {
        keys.push_back(kv.first);
        for (auto &l : lines(kv.second.c_str())) {
            auto w = splitTab(l); if (w.size() != 20) DIE("there should be 20 columns in the input file");
            mb200_cand c; c.target_key = strtoul(w[0].c_str(), 0, 10); c.bit_score = atoi(w[1].c_str()); c.q_start = atoi(w[4].c_str()); c.q_end = atoi(w[5].c_str());
            c.t_start = atoi(w[7].c_str()); c.t_end = atoi(w[8].c_str()); c.t_len = atoi(w[9].c_str()); c.orf_key = strtoul(w[10].c_str(), 0, 10); c.frag_from = atoi(w[17].c_str()); c.frag_to = atoi(w[18].c_str());
            cands.push_back(c); txt.push_back({w[2], w[3]});
        }
        off.push_back(cands.size());
    }
    std::vector<mb200_call> calls(std::max<size_t>(1, 2 * cands.size())); std::vector<mb200_exon> exons(std::max<size_t>(1, cands.size())); std::vector<uint32_t> cnt(std::max<size_t>(1, keys.size()));
    if (cands.empty()) cands.resize(1);
    MB(mb200_chain(ctx, cands.data(), off.data(), (uint32_t)keys.size(), totalRes, calls.data(), cnt.data(), exons.data()));
    std::map<uint32_t, std::string> out;
    for (size_t ci = 0; ci < keys.size(); ci++) {
        std::string e; char b[512]; const uint64_t base = off[ci];
        for (uint32_t k = 0; k < cnt[ci]; k++) {
            const mb200_call &cl = calls[2 * base + k];
            for (int x = 0; x < cl.n_exons; x++) {
                const mb200_exon &ex = exons[cl.exon_begin + x]; const mb200_cand &c = cands[base + ex.cand_idx]; auto &t = txt[base + ex.cand_idx];
                snprintf(b, sizeof(b), "%u\t%d\t%d\t%s\t%d\t%d\t%d\t%u\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", cl.target_key, cl.strand, cl.total_bit_score, fmtE(cl.evalue).c_str(),
                         cl.n_exons, cl.low_coord, cl.high_coord, c.orf_key, c.bit_score, fmtSeqId((float)strtod(t.first.c_str(), nullptr), false).c_str(), fmtE(strtod(t.second.c_str(), nullptr)).c_str(),
                         c.t_start, c.t_end, c.t_len, ex.contig_start, ex.contig_end, ex.nuc_len, c.frag_from, c.frag_to);
                e += b;
            }
        }
        out[keys[ci]] = e;
    }
    return out;
}

static std::map<uint32_t, std::string> searchAll(Session &s, const Db &qdb, const Seqs &q) {
    std::vector<uint64_t> off(qdb.keys.size() + 1); const mb200_aln *alns = nullptr;
    MB(mb200_search(s.ctx, q.res.data(), q.off.data(), (uint32_t)qdb.keys.size(), off.data(), &alns));
    std::map<uint32_t, std::string> out;
    for (size_t i = 0; i < qdb.keys.size(); i++) out[qdb.keys[i]] = alnEntry(s, std::vector<mb200_aln>(alns + off[i], alns + off[i + 1]), (uint32_t)(q.off[i + 1] - q.off[i]));
    return out;
}
static void writeDb(const std::string &name, int dbtype, const std::map<uint32_t, std::string> &m) { DbOut o(name, dbtype); for (auto &kv : m) o.write(kv.first, kv.second); }

int main(int argc, char **argv) {
    if (argc < 2) DIE("usage: mb200 <gpucreateindex|gpuextractorfs|gpureduceredundancy|gpuprefilter|gpualign|gpusearch|gpuresultspercontig|gpucollectoptimalset|gpupredictexons> <DBs...> [options]");
    initCodes();
    std::string cmd = argv[1]; std::vector<std::string> pos; Opt o = parseOpts(argc, argv, 2, pos);
    if (o.gpus > 1 && !(cmd == "gpupredictexons" && pos.size() == 3 && !o.textStages))
        DIE("--gpus N is implemented for `gpupredictexons <contigsDB> <targetDB> <predDB>` (binary hand-off) only; the other modules run on one device (--device N)");
    if (cmd == "gpuprefilter") {
        if (pos.size() != 3) DIE("gpuprefilter <queryDB> <targetDB> <prefDB>");
        Session s; s.open(o, pos[1]); Db qdb = openDb(pos[0]); Seqs q = encodeDb(qdb, false);
        std::vector<uint64_t> off(qdb.keys.size() + 1); const mb200_hit *hits = nullptr;
        MB(mb200_prefilter_compact(s.ctx, q.res.data(), q.off.data(), (uint32_t)qdb.keys.size(), off.data(), &hits));
        DbOut out(pos[2], 7);   // DBTYPE_PREFILTER_RES
        for (size_t i = 0; i < qdb.keys.size(); i++) out.write(qdb.keys[i], prefEntry(s, hits + off[i], off[i + 1] - off[i]));
    } else if (cmd == "gpualign") {
        if (pos.size() != 4) DIE("gpualign <queryDB> <targetDB> <prefDB> <alnDB>");
        Session s; s.open(o, pos[1]); Db qdb = openDb(pos[0]); Seqs q = encodeDb(qdb, false); Db pref = openDb(pos[2]);
        std::vector<uint32_t> flat; std::vector<uint64_t> off(1, 0);
        for (size_t i = 0; i < qdb.keys.size(); i++) {
            size_t p = pref.find(qdb.keys[i]);
            if (p != (size_t)-1) for (auto &l : lines(pref.entry(p))) { size_t t = s.tdb.find((uint32_t)strtoul(l.c_str(), nullptr, 10)); if (t == (size_t)-1) DIE("prefilter hit not in target DB"); flat.push_back((uint32_t)t); }
            off.push_back(flat.size());
        }
        std::vector<mb200_aln> alns(std::max<size_t>(1, flat.size())); if (flat.empty()) flat.push_back(0);
        MB(mb200_align(s.ctx, q.res.data(), q.off.data(), (uint32_t)qdb.keys.size(), flat.data(), off.data(), alns.data()));
        DbOut out(pos[3], 5);   // DBTYPE_ALIGNMENT_RES
        for (size_t i = 0; i < qdb.keys.size(); i++) out.write(qdb.keys[i], alnEntry(s, std::vector<mb200_aln>(alns.begin() + off[i], alns.begin() + off[i + 1]), (uint32_t)(q.off[i + 1] - q.off[i])));
    } else if (cmd == "gpusearch") {
        if (pos.size() != 3) DIE("gpusearch <queryDB> <targetDB> <alnDB>");
        Session s; s.open(o, pos[1]); Db qdb = openDb(pos[0]); Seqs q = encodeDb(qdb, false);
        writeDb(pos[2], 5, searchAll(s, qdb, q));
    } else if (cmd == "gpuresultspercontig") {
        if (pos.size() != 4) DIE("gpuresultspercontig <contigsDB> <orfDB> <alnDB> <byContigDB>");
        Db contigs = openDb(pos[0]), hdr = openDb(pos[1] + "_h"), aln = openDb(pos[2]);
        std::map<uint32_t, std::string> a; for (size_t i = 0; i < aln.keys.size(); i++) a[aln.keys[i]] = aln.entry(i);
        writeDb(pos[3], 5, byContig(contigs, hdr, a));
    } else if (cmd == "gpucollectoptimalset") {
        if (pos.size() != 3) DIE("gpucollectoptimalset <byContigDB> <targetDB> <predDB>");
        Db byc = openDb(pos[0]), tdb = openDb(pos[1]);
        uint64_t total = 0; for (size_t i = 0; i < tdb.keys.size(); i++) total += tdb.len[i] >= 2 ? tdb.len[i] - 2 : 0;
        mb200_ctx *ctx; MB(mb200_ctx_create(o.device, &ctx)); MB(mb200_params_set(ctx, &o.p));
        std::map<uint32_t, std::string> m; for (size_t i = 0; i < byc.keys.size(); i++) m[byc.keys[i]] = byc.entry(i);
        writeDb(pos[2], 12, chainAll(ctx, m, total));   // DBTYPE_GENERIC_DB
        mb200_ctx_destroy(ctx);
    } else if (cmd == "gpupredictexons" && pos.size() == 3) {
        o.p.min_aln_len = o.p.min_exon_aa;      // "align module should return alignments of at least a minimal exon length" (src/workflow/PredictExons.cpp:43-44)
        // contigs are the only input: ORF fragments are extracted and translated on the device, become the resident query batch
        // (no aa_6f / nucl_6f DBs, no host round trip), are searched, regrouped per contig and chained
        Session s; s.open(o, pos[1]); Db contigs = openDb(pos[0]);
        std::vector<size_t> order(contigs.keys.size()); for (size_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return contigs.keys[a] < contigs.keys[b]; });
        std::string nuc; std::vector<uint64_t> off(1, 0);
        for (size_t id : order) { const size_t n = contigs.len[id] >= 2 ? contigs.len[id] - 2 : 0; nuc.append(contigs.entry(id), n); off.push_back(nuc.size()); }
        if (o.textStages) {
            // the same chain with the reference's text DBs between the stages (in memory), for parity checks of the binary hand-off
            uint64_t n = 0, na = 0; const mb200_orf *orfs = nullptr; const char *aa = nullptr;
            MB(mb200_extract_orfs(s.ctx, nuc.data(), off.data(), (uint32_t)order.size(), o.minLen, o.maxLen, 1, &n, &orfs, &na, &aa));
            Db hdr; hdr.dbtype = 12; char buf[96];
            for (uint64_t i = 0; i < n; i++) {          // in-memory nucl_6f_h (Orf::writeOrfHeader, Orf.cpp:440-457)
                const mb200_orf &r = orfs[i]; const long long d = (long long)r.from - (long long)r.to;
                int w = snprintf(buf, sizeof(buf), "%u\t%u%c%lld", contigs.keys[order[r.contig]], r.from, r.from < r.to ? '+' : '-', d < 0 ? -d : d);
                if (r.flags & 3u) w += snprintf(buf + w, sizeof(buf) - w, "\t%u", r.flags & 3u);
                buf[w++] = '\n'; buf[w++] = '\0';
                hdr.byKey[(uint32_t)i] = hdr.keys.size(); hdr.keys.push_back((uint32_t)i); hdr.off.push_back(hdr.data.size()); hdr.len.push_back(w);
                hdr.data.insert(hdr.data.end(), buf, buf + w);
            }
            std::vector<uint32_t> qlen(n); for (uint64_t i = 0; i < n; i++) qlen[i] = orfs[i].aa_len;   // (the record buffer is reused by later calls)
            MB(mb200_queries_from_orfs(s.ctx));
            std::vector<uint64_t> aoff(n + 1); const mb200_aln *alns = nullptr;
            MB(mb200_search_resident(s.ctx, aoff.data(), &alns));
            std::map<uint32_t, std::string> aln;
            for (uint64_t i = 0; i < n; i++) aln[(uint32_t)i] = alnEntry(s, std::vector<mb200_aln>(alns + aoff[i], alns + aoff[i + 1]), qlen[i]);
            auto byc = byContig(contigs, hdr, aln);
            writeDb(pos[2], 12, chainAll(s.ctx, byc, s.t.res.size()));
        } else {
            // binary records from stage to stage (mb200_predict_exons); only the final dp_predictions entries are printed
            MB(mb200_target_keys_set(s.ctx, s.tdb.keys.data(), (uint32_t)s.tdb.keys.size()));
            if (o.batchNt > 0) MB(mb200_option_set(s.ctx, "predict_batch_nt", o.batchNt));
            std::vector<uint64_t> coff(order.size() + 1); const mb200_call *calls = nullptr; const mb200_pexon *ex = nullptr; uint64_t nEx = 0;
            const uint32_t nC = (uint32_t)order.size();
            std::vector<mb200_call> mCalls; std::vector<mb200_pexon> mEx;        // merged records of the shards (--gpus N)
            const int G = (int)std::min<uint32_t>((uint32_t)o.gpus, std::max<uint32_t>(nC, 1u));
            if (G <= 1) {
                MB(mb200_predict_exons(s.ctx, nuc.data(), off.data(), nC, o.minLen, o.maxLen, coff.data(), &calls, &ex, &nEx));
            } else {
                // --gpus N: one context per device (--device d .. d + N - 1), the index copied device to device from the first one, the contigs sharded by nucleotide count
                // (the reference splits a DB over its workers the same way, DBReader.cpp:1216-1256), every shard searched by its own host thread, the records concatenated in
                // contig order.  Fragment keys count over the whole contig DB: a counting pass of the ORF stage gives every shard the number of fragments before it.
                std::vector<mb200_ctx *> ctxs(G, nullptr); ctxs[0] = s.ctx;
                for (int g = 1; g < G; g++) {
                    MB(mb200_ctx_create(o.device + g, &ctxs[g])); MB(mb200_params_set(ctxs[g], &o.p)); MB(mb200_matrices_default(ctxs[g]));
                    uint64_t nb = 0; double ms = 0; MB(mb200_index_clone(ctxs[g], s.ctx, &nb, &ms));
                    if (o.batchNt > 0) MB(mb200_option_set(ctxs[g], "predict_batch_nt", o.batchNt));
                }
                std::vector<uint32_t> cb(G + 1, 0); cb[G] = nC;
                for (int g = 1; g < G; g++) {
                    const uint64_t want = off[nC] * (uint64_t)g / (uint64_t)G;
                    uint32_t c = (uint32_t)(std::lower_bound(off.begin(), off.end(), want) - off.begin());
                    cb[g] = std::min(std::max(c, cb[g - 1]), nC);
                }
                std::vector<std::string> errs(G);
                auto eachShard = [&](const std::function<int(int)> &fn) {
                    std::vector<std::thread> th;
                    for (int g = 0; g < G; g++) th.emplace_back([&, g] { if (fn(g) != MB200_OK) errs[g] = mb200_last_error(); });
                    for (auto &t : th) t.join();
                    for (int g = 0; g < G; g++) if (!errs[g].empty()) DIE("%s (device %d)", errs[g].c_str(), o.device + g);
                };
                std::vector<uint64_t> nFrag(G, 0);
                eachShard([&](int g) {
                    if (cb[g + 1] == cb[g]) return (int)MB200_OK;
                    std::vector<uint64_t> so(off.begin() + cb[g], off.begin() + cb[g + 1] + 1); const uint64_t base = so[0]; for (uint64_t &x : so) x -= base;
                    uint64_t na = 0; const mb200_orf *orfs = nullptr; const char *aa = nullptr;
                    return mb200_extract_orfs(ctxs[g], nuc.data() + base, so.data(), cb[g + 1] - cb[g], o.minLen, o.maxLen, 1, &nFrag[g], &orfs, &na, &aa);
                });
                std::vector<std::vector<uint64_t>> sOff(G); std::vector<const mb200_call *> sCalls(G, nullptr); std::vector<const mb200_pexon *> sEx(G, nullptr); std::vector<uint64_t> sNEx(G, 0);
                uint64_t keyBase = 0; std::vector<uint64_t> keyBases(G, 0);
                for (int g = 0; g < G; g++) { keyBases[g] = keyBase; keyBase += nFrag[g]; }
                eachShard([&](int g) {
                    sOff[g].assign(cb[g + 1] - cb[g] + 1, 0);
                    if (cb[g + 1] == cb[g]) return (int)MB200_OK;
                    std::vector<uint64_t> so(off.begin() + cb[g], off.begin() + cb[g + 1] + 1); const uint64_t base = so[0]; for (uint64_t &x : so) x -= base;
                    int rc = mb200_option_set(ctxs[g], "orf_key_base", (int64_t)keyBases[g]);
                    if (rc == MB200_OK) rc = mb200_predict_exons(ctxs[g], nuc.data() + base, so.data(), cb[g + 1] - cb[g], o.minLen, o.maxLen, sOff[g].data(), &sCalls[g], &sEx[g], &sNEx[g]);
                    return rc;
                });
                coff[0] = 0;
                for (int g = 0; g < G; g++) {
                    const uint64_t callBase = mCalls.size(); const uint32_t exBase = (uint32_t)mEx.size();
                    for (uint64_t k = 0; k < sOff[g].back(); k++) { mb200_call cl = sCalls[g][k]; cl.exon_begin += exBase; mCalls.push_back(cl); }
                    mEx.insert(mEx.end(), sEx[g], sEx[g] + sNEx[g]);
                    for (uint32_t ci = cb[g]; ci < cb[g + 1]; ci++) coff[ci + 1] = callBase + sOff[g][ci - cb[g] + 1];
                    mb200_stats st; MB(mb200_stats_get(ctxs[g], &st));
                    if (st.overflow_queries) fprintf(stderr, "mb200: warning: %u fragment(s) of shard %d exceeded the prefilter's work lists and were searched with an empty hit list\n", st.overflow_queries, g);
                }
                calls = mCalls.data(); ex = mEx.data(); nEx = mEx.size();
                for (int g = 1; g < G; g++) mb200_ctx_destroy(ctxs[g]);
            }
            if (G <= 1) { mb200_stats st; MB(mb200_stats_get(s.ctx, &st));
              if (st.overflow_queries) fprintf(stderr, "mb200: warning: %u fragment(s) exceeded the prefilter's work lists (where the reference takes its own overflow path) and were searched with an empty hit list\n", st.overflow_queries); }
            std::map<uint32_t, std::string> out; char line[512];
            for (size_t ci = 0; ci < order.size(); ci++) {
                std::string e;
                for (uint64_t k = coff[ci]; k < coff[ci + 1]; k++)
                    for (int x = 0; x < calls[k].n_exons; x++) { if (mb200_text_exon_line(&calls[k], &ex[calls[k].exon_begin + x], line, sizeof(line)) < 0) DIE("%s", mb200_last_error()); e += line; }
                out[contigs.keys[order[ci]]] = e;
            }
            writeDb(pos[2], 12, out);
        }
    } else if (cmd == "gpupredictexons") {
        if (pos.size() != 5) DIE("gpupredictexons <contigsDB> <orfAaDB> <orfNuclDB> <targetDB> <predDB>   |   gpupredictexons <contigsDB> <targetDB> <predDB>");
        o.p.min_aln_len = o.p.min_exon_aa;      // src/workflow/PredictExons.cpp:43-44
        Session s; s.open(o, pos[3]); Db contigs = openDb(pos[0]), qdb = openDb(pos[1]), hdr = openDb(pos[2] + "_h"); Seqs q = encodeDb(qdb, false);
        auto aln = searchAll(s, qdb, q);
        auto byc = byContig(contigs, hdr, aln);
        writeDb(pos[4], 12, chainAll(s.ctx, byc, s.t.res.size()));
    } else if (cmd == "gpureduceredundancy") {
        // reduceredundancy: calledExonsDB -> representative calls (DP format) + call -> representative DB
        if (pos.size() != 3) DIE("gpureduceredundancy <i:calledExonsDB> <o:predictionsExonsDB> <o:predToCall> [--overlap 0|1]");
        Db dp = openDb(pos[0]);
        std::vector<mb200_pred> preds; std::vector<uint32_t> exonKeys; std::vector<std::vector<std::vector<std::string>>> exonCols; std::vector<uint64_t> off(1, 0);
        for (size_t id = 0; id < dp.keys.size(); id++) {
            long prevIdx[2] = {-1, -1};
            for (auto &l : lines(dp.entry(id))) {
                if (l.empty()) continue;
                auto w = splitTab(l); if (w.size() != 19) DIE("There should be 19 columns in the input file");
                const uint32_t t = (uint32_t)atoi(w[0].c_str()), low = (uint32_t)atoi(w[5].c_str()); const int s = atoi(w[1].c_str()); const int si = s == 1 ? 0 : 1;
                long j = prevIdx[si];
                if (j < 0 || preds[j].target_key != t || preds[j].low_coord != low) {   // reduceredundancy.cpp:237-262
                    mb200_pred p; memset(&p, 0, sizeof(p)); p.target_key = t; p.strand = s; p.total_bit_score = (uint32_t)atoi(w[2].c_str()); p.evalue = atof(w[3].c_str());
                    p.low_coord = low; p.high_coord = (uint32_t)atoi(w[6].c_str()); p.exon_begin = (uint32_t)exonKeys.size();
                    preds.push_back(p); exonCols.emplace_back(); prevIdx[si] = j = (long)preds.size() - 1;
                }
                preds[j].n_exons++; exonKeys.push_back((uint32_t)atoi(w[7].c_str())); exonCols[j].push_back(std::vector<std::string>(w.begin() + 7, w.end()));
            }
            off.push_back(preds.size());
        }
        mb200_ctx *ctx; MB(mb200_ctx_create(o.device, &ctx));
        std::vector<mb200_pred_out> out(preds.size() + 1); std::vector<uint32_t> mo(preds.size() + 1), ro(preds.size() + 1), cnt(dp.keys.size() + 1), keys1(exonKeys); keys1.push_back(0);
        mb200_pred dummy; memset(&dummy, 0, sizeof(dummy));
        MB(mb200_reduce_redundancy(ctx, preds.empty() ? &dummy : preds.data(), off.data(), (uint32_t)dp.keys.size(), keys1.data(), exonKeys.size(), out.data(), mo.data(), ro.data(), cnt.data()));
        DbOut grouped(pos[1], 12), clusters(pos[2], 12);
        char buf[256];
        for (size_t id = 0; id < dp.keys.size(); id++) {
            const uint64_t b = off[id], e = off[id + 1]; std::string g, cl;
            for (uint32_t r = 0; r < cnt[id]; r++) {
                const uint64_t i = b + ro[b + r]; const mb200_pred &p = preds[i];
                if (!o.overlap && out[i].no_overlap_cluster_id != p.target_key) continue;
                for (auto &ex : exonCols[i]) {       // Prediction::predictionToBuffer + PotentialExon::exonToBuffer (PredictionParser.h:368-396,99-150)
                    snprintf(buf, sizeof(buf), "%u\t%d\t%u\t%.3E\t%u\t%u\t%u\t", p.target_key, p.strand, p.total_bit_score, p.evalue, p.n_exons, p.low_coord, p.high_coord); g += buf;
                    const float sid = (float)strtod(ex[2].c_str(), NULL);
                    std::string sidTxt;
                    if (sid == 1.0f) sidTxt = "1.000"; else { sidTxt = "0."; if (sid < 0.10f) sidTxt += "0"; if (sid < 0.01f) sidTxt += "0"; sidTxt += std::to_string((int)(sid * 1000)); }
                    snprintf(buf, sizeof(buf), "%.3E", strtod(ex[3].c_str(), NULL));
                    g += ex[0] + "\t" + ex[1] + "\t" + sidTxt + "\t" + buf;
                    for (size_t k = 4; k < ex.size(); k++) g += "\t" + ex[k];
                    g += "\n";
                }
            }
            for (uint64_t m = b; m < e; m++) {
                const uint64_t i = b + mo[m]; snprintf(buf, sizeof(buf), "%u\t%d\t%u\t%u\t%d\t%u\n", out[i].cluster_id, preds[i].strand, out[i].cluster_low_coord, preds[i].target_key, preds[i].strand, preds[i].low_coord); cl += buf;
            }
            grouped.write(dp.keys[id], g); clusters.write(dp.keys[id], cl);
        }
        mb200_ctx_destroy(ctx);
    } else if (cmd == "gpucreateindex") {
        // createindex / indexdb (M/src/util/indexdb.cpp): build the k-mer index on the host and persist it in the reference's index-DB layout
        if (pos.size() != 2) DIE("gpucreateindex <targetDB> <o:indexDB> [--masked-db <DB>] [--k-score N]");
        initCodes();
        Db tdb = openDb(pos[0]); Seqs t = encodeDb(tdb, false); Seqs m = t;
        if (o.maskedDb.empty()) maskTargets(o, tdb, t, m);
        if (!o.maskedDb.empty()) {
            Db md = openDb(o.maskedDb); Seqs raw = encodeDb(md, true);
            for (size_t i = 0; i < tdb.keys.size(); i++) {
                size_t j = md.find(tdb.keys[i]); if (j == (size_t)-1 || raw.off[j + 1] - raw.off[j] != t.off[i + 1] - t.off[i]) DIE("masked DB does not match target DB at key %u", tdb.keys[i]);
                memcpy(&m.res[t.off[i]], &raw.res[raw.off[j]], t.off[i + 1] - t.off[i]);
            }
        }
        MB(mb200_index_file_write(pos[1].c_str(), m.res.data(), t.res.data(), t.off.data(), (uint32_t)tdb.keys.size(), tdb.keys.data(), o.p.kmer_thr, (o.maskedDb.empty() && !o.mask) ? 0 : 1));   // META's mask field = --mask only (PrefilteringIndexReader.cpp:78)
    } else if (cmd == "gpuextractorfs") {
        // extractorfs + translatenucs in one step (data/predictexons.sh: contigs -> nucl_6f(_h) -> aa_6f): writes the ORF header DB and the
        // translated fragment DB with the reference's renumbered keys
        if (pos.size() != 3) DIE("gpuextractorfs <contigsDB> <o:orfHeaderDB (nucl_6f_h)> <o:aaDB (aa_6f)>");
        Db contigs = openDb(pos[0]);
        // the reference processes contigs in index order and renumbers by (contig key, offset): feed them sorted by key
        std::vector<size_t> order(contigs.keys.size()); for (size_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return contigs.keys[a] < contigs.keys[b]; });
        std::string nuc; std::vector<uint64_t> off(1, 0);
        for (size_t id : order) { const size_t n = contigs.len[id] >= 2 ? contigs.len[id] - 2 : 0; nuc.append(contigs.entry(id), n); off.push_back(nuc.size()); }
        mb200_ctx *ctx; MB(mb200_ctx_create(o.device, &ctx)); MB(mb200_params_set(ctx, &o.p));
        uint64_t n = 0, na = 0; const mb200_orf *orfs = nullptr; const char *aa = nullptr;
        MB(mb200_extract_orfs(ctx, nuc.data(), off.data(), (uint32_t)order.size(), o.minLen, o.maxLen, 1, &n, &orfs, &na, &aa));
        DbOut hdr(pos[1], 12), seq(pos[2], 0);   // DBTYPE_GENERIC_DB, DBTYPE_AMINO_ACIDS
        char buf[96];
        for (uint64_t i = 0; i < n; i++) {
            const mb200_orf &r = orfs[i]; const long long d = (long long)r.from - (long long)r.to;
            int w = snprintf(buf, sizeof(buf), "%u\t%u%c%lld", contigs.keys[order[r.contig]], r.from, r.from < r.to ? '+' : '-', d < 0 ? -d : d);   // Orf::writeOrfHeader (Orf.cpp:440-457)
            if (r.flags & 3u) w += snprintf(buf + w, sizeof(buf) - w, "\t%u", r.flags & 3u);
            buf[w++] = '\n';
            hdr.write((uint32_t)i, std::string(buf, w));
            seq.write((uint32_t)i, std::string(aa + r.aa_off, r.aa_len) + "\n");
        }
        mb200_ctx_destroy(ctx);
    } else DIE("unknown command %s", cmd.c_str());
    return EXIT_SUCCESS;
}
