// reduce_core.cuh -- reduceredundancy, the step after the path: per contig and strand, greedy clustering of calls that
// share an exon with a (temporary) representative, the best-scoring member becomes the representative, then E-value ordered exclusion of
// same-strand overlaps among representatives.  One contig per thread, integers and one double comparison only (the E-value is parsed
// from the `%.3E` text on the host, exactly as the reference does with atof).
//   clusterPredictions / excludeSameStrandOverlaps   reference/src/exonpredictor/reduceredundancy.cpp:24-109,111-147
//   comparators                                      reference/src/commons/PredictionParser.h:318-366
//   per-contig driver and output orders              reference/src/exonpredictor/reduceredundancy.cpp:273-296
// std::stable_sort is reproduced with stable insertion sorts on index arrays.
#pragma once
#include "cta_prims.cuh"

struct MbPred { uint32_t targetKey; int32_t strand; uint32_t totalBitscore, nExons, lowCoord, highCoord, exonBegin, pad; double evalue; };
struct MbPredOut { uint32_t clusterId, clusterLowCoord, noOverlapClusterId, isRep; };

// comparePredictionsByContigStart: start asc, end desc, bit score desc, target asc
MB_DEV bool mb_rr_before_start(const MbPred &a, const MbPred &b) {
    if (a.lowCoord != b.lowCoord) return a.lowCoord < b.lowCoord;
    if (a.highCoord != b.highCoord) return a.highCoord > b.highCoord;
    if (a.totalBitscore != b.totalBitscore) return a.totalBitscore > b.totalBitscore;
    return a.targetKey < b.targetKey;
}
MB_DEV bool mb_rr_before_evalue(const MbPred &a, const MbPred &b) {     // comparePredictionsByEvalue
    if (a.evalue != b.evalue) return a.evalue < b.evalue;
    return a.lowCoord < b.lowCoord;
}
MB_DEV bool mb_rr_before_target(const MbPred &a, const MbPred &b) {     // comparePredictionsByTarget (minus before plus)
    if (a.targetKey != b.targetKey) return a.targetKey < b.targetKey;
    return a.strand < b.strand;
}
template <class CMP> MB_DEV void mb_rr_stable_sort(uint32_t *idx, uint32_t n, const MbPred *P, CMP before) {
    for (uint32_t i = 1; i < n; i++) {
        const uint32_t x = idx[i]; uint32_t j = i;
        while (j > 0 && before(P[x], P[idx[j - 1]])) { idx[j] = idx[j - 1]; j--; }
        idx[j] = x;
    }
}

// One strand: idx[0..n) = the calls of the strand in input order (sorted in place = the order of the call -> representative DB);
// reps[0..*nRep) = representatives, sorted by E-value afterwards.  flag / tmp: scratch of n entries.
MB_DEV void mb_rr_strand(const MbPred *P, const uint32_t *exonKeys, uint32_t *idx, uint32_t n, MbPredOut *out, uint32_t *reps, uint32_t *nRepOut,
                         uint32_t *flag, uint32_t *tmp) {
    mb_rr_stable_sort(idx, n, P, mb_rr_before_start);
    for (uint32_t i = 0; i < n; i++) flag[i] = 0;
    uint32_t nRep = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (flag[i]) continue;
        const MbPred pi = P[idx[i]];
        uint32_t finalId = pi.targetKey, finalLow = pi.lowCoord, maxScore = pi.totalBitscore, nm = 0;
        flag[i] = 1; tmp[nm++] = i;
        for (uint32_t j = i + 1; j < n; j++) {
            const MbPred pj = P[idx[j]];
            if (pj.lowCoord >= pi.highCoord) break;
            if (flag[j]) continue;                 // (sharing an exon only matters for calls that are still free)
            bool share = false;
            for (uint32_t a = 0; a < pi.nExons && !share; a++) { const uint32_t ka = exonKeys[pi.exonBegin + a]; for (uint32_t b = 0; b < pj.nExons; b++) if (exonKeys[pj.exonBegin + b] == ka) { share = true; break; } }
            if (share) {
                flag[j] = 1; tmp[nm++] = j;
                if (pj.totalBitscore > maxScore) { maxScore = pj.totalBitscore; finalId = pj.targetKey; finalLow = pj.lowCoord; }
            }
        }
        for (uint32_t m = 0; m < nm; m++) {
            const uint32_t g = idx[tmp[m]];
            out[g].clusterId = finalId; out[g].clusterLowCoord = finalLow;
            if (P[g].targetKey == finalId && P[g].lowCoord == finalLow) { out[g].isRep = 1; reps[nRep++] = g; }
        }
    }
    // same-strand overlaps among representatives, best E-value first
    mb_rr_stable_sort(reps, nRep, P, mb_rr_before_evalue);
    for (uint32_t i = 0; i < nRep; i++) flag[i] = 0;
    for (uint32_t i = 0; i < nRep; i++) {
        if (flag[i]) continue;
        flag[i] = 1; const MbPred a = P[reps[i]]; out[reps[i]].noOverlapClusterId = a.targetKey;
        for (uint32_t j = i + 1; j < nRep; j++) {
            const MbPred b = P[reps[j]];
            if ((b.highCoord < a.highCoord && b.highCoord > a.lowCoord) || (b.lowCoord < a.highCoord && b.lowCoord > a.lowCoord) ||
                (b.highCoord < a.highCoord && b.lowCoord > a.lowCoord) || (b.highCoord > a.highCoord && b.lowCoord < a.lowCoord)) {
                flag[j] = 1; out[reps[j]].noOverlapClusterId = a.targetKey;
            }
        }
    }
    *nRepOut = nRep;
}

// One contig: P[0..n) in input order.  memberOrder[0..n): plus-strand calls (contig-start order) then minus; repOrder[0..return):
// representatives in output order.  flag, tmp: scratch of n entries each.
MB_DEV uint32_t mb_reduce_contig(const MbPred *P, uint32_t n, const uint32_t *exonKeys, MbPredOut *out, uint32_t *memberOrder, uint32_t *repOrder,
                                 uint32_t *flag, uint32_t *tmp) {
    uint32_t np = 0;
    for (uint32_t i = 0; i < n; i++) { out[i].clusterId = 0; out[i].clusterLowCoord = 0; out[i].noOverlapClusterId = 0; out[i].isRep = 0; if (P[i].strand == 1) memberOrder[np++] = i; }
    uint32_t k = np;
    for (uint32_t i = 0; i < n; i++) if (P[i].strand != 1) memberOrder[k++] = i;
    uint32_t rp = 0, rm = 0;
    mb_rr_strand(P, exonKeys, memberOrder, np, out, repOrder, &rp, flag, tmp);
    mb_rr_strand(P, exonKeys, memberOrder + np, n - np, out, repOrder + rp, &rm, flag, tmp);
    mb_rr_stable_sort(repOrder, rp + rm, P, mb_rr_before_target);
    return rp + rm;
}
