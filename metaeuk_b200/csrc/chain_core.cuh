// chain_core.cuh -- MetaEuk's per-(target, contig, strand) exon-chaining dynamic program, one contig per thread.
//
// Restates, with integer arithmetic only on the device,
//   PotentialExon::setByAln                 reference/src/commons/PredictionParser.h:15-65
//   isPairCompatible / getPenaltyForProtCoords / findoptimalsetbydp
//                                           reference/src/exonpredictor/collectoptimalset.cpp:36-82,84-106,108-222
//   the per-target grouping and --max-exon-sets loop    collectoptimalset.cpp:288-418
// The set e-value (double pow/log2, collectoptimalset.cpp:323-325) and its <= --metaeuk-eval filter are applied
// by the host wrapper from the integer total score this kernel returns, so no libm call happens on the device.
#pragma once
#include "cta_prims.cuh"

struct MbCand { uint32_t targetKey; int32_t bitScore; int32_t qStart, qEnd; int32_t tStart, tEnd, tLen; uint32_t orfKey; int32_t fragFrom, fragTo; };
struct MbExon { uint32_t candIdx; int32_t contigStart, contigEnd, nucLen; };
struct MbCall { uint32_t targetKey; int32_t strand; int32_t totalBitScore; int32_t nExons; uint32_t exonBegin; int32_t lowCoord, highCoord, pad; double evalue; };
struct MbChainParams { int minIntron, maxIntron, minExonAa, maxOverlap, gapOpen, gapExtend, maxSets; double tcov; };
struct MbPE { int32_t cStart, cEnd, tStart, tEnd, tLen, bitScore, aaLen, nucLen; uint32_t cand; int32_t used; };
struct MbRow { int32_t prev, score, nEx, aaLen; };

MB_DEV int mb_ilog2(uint32_t v) { int r = 0; while (v >>= 1) r++; return r; }   // (int)log2(n) for n >= 1 (collectoptimalset.cpp:174)

// findoptimalsetbydp on the list ord[0..n) of indices into pe[].  Returns the best path score (0 = none);
// the chosen exons (in contig order) are appended to exOut[*nEx...].  Shrinks *pn to the not-yet-used prefix.
MB_DEV int mb_dp_find(MbPE *pe, int32_t *ord, int32_t *pn, MbRow *rows, const MbChainParams &P, MbExon *exOut, uint32_t *nEx, uint32_t candBase) {
    int n = *pn;
    if (n == 0) return 0;
    // stable sort by (used, contigStart, contigEnd)  (PredictionParser.h:166-182); insertion sort keeps equal keys in order
    for (int i = 1; i < n; i++) {
        const int32_t v = ord[i]; const MbPE &a = pe[v]; int j = i - 1;
        while (j >= 0) {
            const MbPE &b = pe[ord[j]];
            const bool less = (a.used < b.used) || (a.used == b.used && (a.cStart < b.cStart || (a.cStart == b.cStart && a.cEnd < b.cEnd)));
            if (!less) break;
            ord[j + 1] = ord[j]; j--;
        }
        ord[j + 1] = v;
    }
    int firstUsed = n;
    for (int i = 0; i < n; i++) if (pe[ord[i]].used) { firstUsed = i; break; }
    n = firstUsed; *pn = n;
    if (n == 0) return 0;
    const int targetLength = pe[ord[0]].tLen;
    for (int i = 0; i < n; i++) { rows[i].prev = i; rows[i].score = pe[ord[i]].bitScore; rows[i].nEx = 1; rows[i].aaLen = pe[ord[i]].aaLen; }
    int best = 0, last = 0;
    for (int cur = 0; cur < n; cur++) {
        const MbPE &b = pe[ord[cur]];
        for (int prev = 0; prev < cur; prev++) {
            const MbPE &a = pe[ord[prev]];
            if (b.cEnd < a.cEnd) continue;                                   // first must not contain second
            const int diffOnContig = b.cStart - a.cEnd - 1;
            if (diffOnContig < 0) continue;                                  // no overlap on the contig
            if (diffOnContig < P.minIntron || diffOnContig > P.maxIntron) continue;
            const int diffAAs = b.tStart - a.tEnd - 1;
            int ov = 0;
            if (diffAAs < 0) { ov = -diffAAs; if (ov > P.maxOverlap) continue; }
            if (b.tStart < a.tStart) continue;                               // contig order must follow target order
            int pen;
            if (diffAAs < 0) pen = P.gapOpen + P.gapExtend * (-diffAAs - 1);
            else if (diffAAs <= 1) pen = 0;
            else pen = P.gapOpen + P.gapExtend * (diffAAs - 1);
            const int nEx2 = rows[prev].nEx + 1;
            const int sc = rows[prev].score + pen + b.bitScore + mb_ilog2((uint32_t)nEx2);
            if (sc > rows[cur].score) { rows[cur].prev = prev; rows[cur].score = sc; rows[cur].nEx = nEx2; rows[cur].aaLen = rows[prev].aaLen + b.aaLen - ov; }
        }
        if ((double)rows[cur].aaLen / (double)targetLength >= P.tcov) if (rows[cur].score > best) { last = cur; best = rows[cur].score; }
    }
    if (best == 0) return 0;
    // traceback (last exon first), then emit in contig order
    int cnt = 0; int e = last;
    while (true) { cnt++; if (rows[e].prev == e) break; e = rows[e].prev; }
    e = last; int k = cnt;
    while (true) {
        MbPE &x = pe[ord[e]]; x.used = 1; k--;
        MbExon &o = exOut[*nEx + k]; o.candIdx = x.cand - candBase; o.contigStart = x.cStart; o.contigEnd = x.cEnd; o.nucLen = x.nucLen;
        if (rows[e].prev == e) break; e = rows[e].prev;
    }
    *nEx += cnt;
    return best;
}

// One (contig, target) group: cands[0..m) all have the same target key (m candidates of one contig, sorted by orfKey); every array is the group's own slice.
// idxBase = index of cands[0] inside its contig (candidate / exon indices in the output are contig-relative).  Returns the number of calls (appended to calls[0..)),
// or -1 on an illegal codon segment length; the exons of the calls are appended to exons[0..) (at most m of each).
MB_DEV int mb_chain_group(const MbCand *cands, uint32_t m, uint32_t idxBase, const MbChainParams &P, MbPE *pe, int32_t *ordPlus, int32_t *ordMinus, MbRow *rows,
                          MbCall *calls, MbExon *exons) {
    uint32_t nCalls = 0, nEx = 0;
    int32_t np = 0, nm = 0;
    for (uint32_t k = 0; k < m; k++) {
        const MbCand &c = cands[k]; MbPE &e = pe[k];
        e.cand = k; e.bitScore = c.bitScore; e.tStart = c.tStart; e.tEnd = c.tEnd; e.tLen = c.tLen; e.used = 0;
        int strand;
        if (c.fragFrom < c.fragTo) { e.cStart = c.fragFrom + c.qStart * 3; e.cEnd = c.fragFrom + c.qEnd * 3 + 2; strand = 1; }
        else { e.cStart = -1 * (c.fragFrom - c.qStart * 3); e.cEnd = -1 * (c.fragFrom - c.qEnd * 3 - 2); strand = -1; }
        e.nucLen = e.cEnd - e.cStart + 1;
        if (e.nucLen % 3 != 0) return -1;
        e.aaLen = e.nucLen / 3;
        const int aal = (e.nucLen < 0 ? -e.nucLen : e.nucLen) / 3;
        if (aal >= P.minExonAa) { if (strand == 1) ordPlus[np++] = (int32_t)k; else ordMinus[nm++] = (int32_t)k; }
    }
    int it = 0;
    while (it < P.maxSets && (np > 0 || nm > 0)) {
        for (int s = 0; s < 2; s++) {
            const uint32_t exBegin = nEx;
            const int tot = (s == 0) ? mb_dp_find(pe, ordPlus, &np, rows, P, exons, &nEx, 0u - idxBase) : mb_dp_find(pe, ordMinus, &nm, rows, P, exons, &nEx, 0u - idxBase);
            if (nEx > exBegin) {
                MbCall &cl = calls[nCalls++]; cl.targetKey = cands[0].targetKey; cl.strand = (s == 0) ? 1 : -1; cl.totalBitScore = tot; cl.nExons = (int32_t)(nEx - exBegin);
                cl.exonBegin = idxBase + exBegin; cl.pad = 0; cl.evalue = 0.0;
                const MbExon &f = exons[exBegin], &l = exons[nEx - 1];
                cl.lowCoord = (s == 0) ? f.contigStart : (-1 * l.contigEnd); cl.highCoord = (s == 0) ? l.contigEnd : (-1 * f.contigStart);   // PredictionParser.h:236-237
            }
        }
        it++;
    }
    return (int)nCalls;
}

// One contig: cands[0..n) sorted by (targetKey, orfKey).  Scratch arrays have n entries each.  Emits every optimal set
// (host applies the e-value filter).  Returns the number of calls, or -1 on an illegal codon segment length.  Calls land in calls[0..), in target order; the exons of a
// group stay in the group's own slice of exons[] (calls carry contig-relative exon indices).
MB_DEV int mb_chain_contig(const MbCand *cands, uint32_t n, const MbChainParams &P, MbPE *pe, int32_t *ordPlus, int32_t *ordMinus, MbRow *rows,
                           MbCall *calls, MbExon *exons, uint32_t *nExonsOut) {
    uint32_t nCalls = 0, nEx = 0, i = 0;
    while (i < n) {
        uint32_t j = i; while (j < n && cands[j].targetKey == cands[i].targetKey) j++;
        // (the group writes its calls behind the contig's earlier ones: at most one call per candidate and strand round, so calls[nCalls..] never runs into a later group's slice)
        const int r = mb_chain_group(cands + i, j - i, i, P, pe + i, ordPlus + i, ordMinus + i, rows + i, calls + nCalls, exons + i);
        if (r < 0) return -1;
        for (int k = 0; k < r; k++) nEx += (uint32_t)calls[nCalls + k].nExons;
        nCalls += (uint32_t)r;
        i = j;
    }
    *nExonsOut = nEx;
    return (int)nCalls;
}
