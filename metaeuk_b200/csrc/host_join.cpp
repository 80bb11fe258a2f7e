// host_join.cpp -- see host_join.h.  Plain C++17 + OpenMP; part of libmetaeuk_b200.so and (for CPU-only logic tests) of tests/emu.
#include "host_join.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace mb200 {

int join_by_contig(const mb200_orf *orfs, uint64_t nOrfs, uint32_t nContigs, uint64_t orfKeyBase, const uint64_t *alnOff, const mb200_aln *alns,
                   const uint32_t *tKeys, const uint64_t *tOff, std::vector<mb200_cand> &cands, std::vector<uint64_t> &candOff, std::vector<uint64_t> &candAln) {
    candOff.assign((size_t)nContigs + 1, 0);
    if (orfKeyBase + nOrfs > (1ull << 32)) return -1;
    // fragment range of every contig (records are sorted by contig)
    std::vector<uint64_t> fb((size_t)nContigs + 1, 0);
    for (uint64_t i = 0; i < nOrfs; i++) {
        if (orfs[i].contig >= nContigs || (i && orfs[i].contig < orfs[i - 1].contig)) return -1;
        fb[orfs[i].contig + 1]++;
    }
    for (uint32_t c = 0; c < nContigs; c++) fb[c + 1] += fb[c];
    for (uint32_t c = 0; c < nContigs; c++) candOff[c + 1] = candOff[c] + (alnOff[fb[c + 1]] - alnOff[fb[c]]);
    const uint64_t n = candOff[nContigs];
    if (cands.size() < n) cands.resize(n);          // grow only: the caller may pass the arrays of the previous batch (every element of [0, n) is written below)
    if (candAln.size() < n) candAln.resize(n);
#pragma omp parallel
    {
        std::vector<std::pair<uint64_t, uint64_t>> order;   // ((target key << 32) | ORF key, alignment index): keys are unique inside a contig
#pragma omp for schedule(dynamic, 16)
        for (uint32_t c = 0; c < nContigs; c++) {
            order.clear();
            for (uint64_t i = fb[c]; i < fb[c + 1]; i++)
                for (uint64_t a = alnOff[i]; a < alnOff[i + 1]; a++) {
                    const uint32_t t = alns[a].seq_id;
                    order.push_back({((uint64_t)(tKeys ? tKeys[t] : t) << 32) | (uint32_t)(orfKeyBase + i), a});
                }
            std::sort(order.begin(), order.end());
            uint64_t w = candOff[c];
            for (const auto &o : order) {
                const mb200_aln &al = alns[o.second]; const uint64_t i = (uint32_t)o.first - orfKeyBase; const mb200_orf &f = orfs[i];
                mb200_cand &k = cands[w]; candAln[w] = o.second; w++;
                k.target_key = (uint32_t)(o.first >> 32); k.bit_score = al.bit_score; k.q_start = al.q_start; k.q_end = al.q_end;
                k.t_start = al.db_start; k.t_end = al.db_end; k.t_len = (int32_t)(tOff[al.seq_id + 1] - tOff[al.seq_id]);
                k.orf_key = (uint32_t)o.first; k.frag_from = (int32_t)f.from; k.frag_to = (int32_t)f.to;
            }
        }
    }
    return 0;
}

uint32_t finish_calls(const mb200_call *in, uint32_t n, const mb200_exon *exIn, uint64_t exBase, uint64_t totalTargetResidues, double evalThr,
                      mb200_call *out, mb200_exon *exOut) {
    uint32_t w = 0, we = 0;
    for (uint32_t k = 0; k < n; k++) {
        mb200_call cl = in[k];
        const double l2 = log2(totalTargetResidues) + log2(2) - cl.total_bit_score; const double ev = pow(2, l2);
        if (!(ev <= evalThr)) continue;
        const uint32_t eb = cl.exon_begin;
        cl.evalue = ev; cl.exon_begin = (uint32_t)(exBase + we);
        for (int x = 0; x < cl.n_exons; x++) exOut[we + x] = exIn[eb + x];
        we += (uint32_t)cl.n_exons;
        out[w++] = cl;
    }
    return w;
}

static std::string seq_id_digits(float x, bool dropLastZero) {   // Util::fastSeqIdToBuffer (Util.cpp:222-251)
    if (x == 1.0f) return dropLastZero ? "1.00" : "1.000";
    std::string s = "0."; if (x < 0.10) s += "0"; if (x < 0.01) s += "0";
    char b[32]; snprintf(b, sizeof(b), "%d", (int)(x * 1000));
    return s + b;
}
std::string text_seq_id_align(float x) { return seq_id_digits(x, true); }
std::string text_seq_id_bycontig(float x) { return seq_id_digits(strtof(text_seq_id_align(x).c_str(), nullptr), true); }
std::string text_seq_id_exon(float x) { return seq_id_digits((float)strtod(text_seq_id_bycontig(x).c_str(), nullptr), false); }
std::string text_evalue(double e) { char b[64]; snprintf(b, sizeof(b), "%.3E", e); return b; }

std::string text_bycontig_line(const mb200_cand &k, const mb200_aln &a, uint32_t qLen, uint64_t contigLen) {
    const long from = k.frag_from, to = k.frag_to, orfLen = std::max(from, to) - std::min(from, to) + 1;
    char b[512];
    snprintf(b, sizeof(b), "%u\t%d\t%s\t%s\t%d\t%d\t%u\t%d\t%d\t%d\t%u\t1\t1.00\t0.000E+00\t0\t%ld\t%ld\t%ld\t%ld\t%llu\n", k.target_key, k.bit_score,
             text_seq_id_bycontig(a.seq_id_est).c_str(), text_evalue(a.evalue).c_str(), k.q_start, k.q_end, qLen, k.t_start, k.t_end, k.t_len,
             k.orf_key, orfLen - 1, orfLen, from, to, (unsigned long long)contigLen);
    return b;
}

std::string text_exon_line(const mb200_call &cl, const mb200_pexon &e) {   // Prediction::predictionToBuffer (PredictionParser.h:368-396)
    char b[512];
    // the exon's E-value was printed with %.3E by `align` and goes through strtod + %.3E twice more: idempotent after the first print
    const double ev = strtod(text_evalue(e.evalue).c_str(), nullptr);
    snprintf(b, sizeof(b), "%u\t%d\t%d\t%s\t%d\t%d\t%d\t%u\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", cl.target_key, cl.strand, cl.total_bit_score,
             text_evalue(cl.evalue).c_str(), cl.n_exons, cl.low_coord, cl.high_coord, e.orf_key, e.bit_score, text_seq_id_exon(e.seq_id_est).c_str(),
             text_evalue(ev).c_str(), e.t_start, e.t_end, e.t_len, e.contig_start, e.contig_end, e.nuc_len, e.frag_from, e.frag_to);
    return b;
}

}  // namespace mb200
