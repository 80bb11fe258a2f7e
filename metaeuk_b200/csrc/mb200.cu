// mb200.cu -- CUDA kernels (sm_100a) and the C ABI of libmetaeuk_b200.so.  See include/metaeuk_b200.h.
//
// Data layout in HBM (per context = per GPU), all resident for the life of the index:
//   sim3     u32 [8000][8000]   3-mer neighbour lists (index<<16 | score), 256 MB, only row prefixes are touched
//   cnt3     u16 [8000][384]    "how many neighbours score >= x", 6 MB, L2 resident
//   offsets  u32 [20^6+1]       k-mer -> first index entry, 256 MB, random 8-byte reads
//   entries  u32x2 [nEntries]   (target id, position), streamed per list
//   tseqM/U  u8  [residues]     masked / unmasked target residues, read along diagonals / rows
// Per-CTA scratch lists (window/item/k-mer/hit expansions) live in global memory but are sized to stay in L2.
#include <cuda_runtime.h>
#ifndef MB200_SIMT_EMU
// kernel launch / dynamic shared memory as macros: tests/simt compiles this file with a stand-in <cuda_runtime.h> that runs kernels as cooperative fibers on the CPU
// (test infrastructure for the GPU suite's logic; the library that ships is the nvcc build of this file and has no CPU path)
#define MB_UNPAREN(...) __VA_ARGS__
#define MB_LAUNCH(kern, grid, block, smem, stream, ...) MB_UNPAREN kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define MB_DYN_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#include <algorithm>
#include <parallel/algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "../../include/metaeuk_b200.h"
#include "chain_core.cuh"
#include "reduce_core.cuh"
#include "host_tables.h"
#include "host_join.h"
#include "host_indexdb.h"
#include "prefilter_core.cuh"

static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_err = buf; return code;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(MB200_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); } while (0)

// =====================================================================================================
// kernels
// =====================================================================================================

// ---- per-residue preparation (one thread per query residue) ---------------------------------------
__global__ void __launch_bounds__(256) k_prep(const uint8_t *__restrict__ qres, const uint64_t *__restrict__ qoff, uint32_t nQ, uint64_t total,
                                              const int16_t *__restrict__ seedMat, const double *__restrict__ seedPBack,
                                              const int16_t *__restrict__ alnMat, const double *__restrict__ alnPBack,
                                              float *__restrict__ bias, int8_t *__restrict__ corr, int8_t *__restrict__ cb) {
    __shared__ int16_t sSeed[441], sAln[441];
    __shared__ double sPs[21], sPa[21];
    for (int i = threadIdx.x; i < 441; i += blockDim.x) { sSeed[i] = seedMat[i]; sAln[i] = alnMat[i]; }
    if (threadIdx.x < 21) { sPs[threadIdx.x] = seedPBack[threadIdx.x]; sPa[threadIdx.x] = alnPBack[threadIdx.x]; }
    __syncthreads();
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= total) return;
    uint32_t lo = 0, hi = nQ;   // qoff[lo] <= r < qoff[hi]
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (qoff[mid] <= r) lo = mid; else hi = mid; }
    const uint64_t b = qoff[lo]; const int N = (int)(qoff[lo + 1] - b); const int i = (int)(r - b);
    const float bs = mb_comp_bias_at(sSeed, sPs, qres + b, N, i);
    bias[r] = bs; corr[r] = mb_round_bias_quarter(bs);
    cb[r] = mb_round_bias(mb_comp_bias_at(sAln, sPa, qres + b, N, i));
}

// ---- prefilter: persistent CTAs, one query at a time, dynamic work distribution --------------------
template <class SH>
__global__ void __launch_bounds__(SH::Cfg::NT, SH::Cfg::CTAS) k_prefilter(MbPrefTables T, const MbPrefScratch *__restrict__ scratch, const uint8_t *__restrict__ qres,
                                                              const uint64_t *__restrict__ qoff, const float *__restrict__ bias, const int8_t *__restrict__ corr,
                                                              const uint32_t *__restrict__ order, uint32_t nQ, const int8_t *__restrict__ ungMat8,
                                                              MbPrefHit *out, uint32_t *outCount, uint32_t *outStatus, MbPrefStats *stats, uint32_t *nextQ) {
    MB_DYN_SHARED(smemRaw);
    SH *sh = reinterpret_cast<SH *>(smemRaw);
    for (int i = threadIdx.x; i < 441; i += blockDim.x) sh->ungMat[i] = ungMat8[i];
    if (T.phaseStats && threadIdx.x == 0) { for (int i = 0; i < 16; i++) sh->phaseAcc[i] = 0; sh->phaseT = clock64(); }
    const MbPrefScratch W = scratch[blockIdx.x];
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) sh->cs.bcast[0] = atomicAdd(nextQ, 1u);
        __syncthreads();
        const uint32_t slot = sh->cs.bcast[0];
        if (slot >= nQ) break;
        const uint32_t qi = order[slot];
        const uint64_t b = qoff[qi]; const uint32_t L = (uint32_t)(qoff[qi + 1] - b);
        mb_prefilter_query(T, W, sh, qres + b, bias + b, corr + b, L, out + (size_t)qi * T.maxHits, outCount + qi, outStatus + qi, stats);
        __syncthreads();
    }
    if (T.phaseStats && threadIdx.x == 0 && stats) for (int i = 0; i < 16; i++) atomicAdd(&stats->phase[i], sh->phaseAcc[i]);
}

#include "sw_kernels.cuh"
#include "sw_stream.cuh"
#include "orf_kernels.cuh"

// ---- exon chaining: one thread per (contig, target) group -- a contig has hundreds of groups of a few candidates each, so a batch of 500 contigs gives ~10^5 independent DPs
//      (round 1 ran one contig per thread: 500 threads on 4 CTAs, 8.8 ms per C4 batch).  A thread whose candidate is the first of its group runs the group; calls and exons
//      land in the group's own slices (at most one call / exon per candidate), k_chain_compact then lines the calls of a contig up in target order.
__global__ void __launch_bounds__(128) k_chain_groups(const MbCand *__restrict__ cands, const uint64_t *__restrict__ candOff, uint32_t nContigs, uint64_t n, MbChainParams P,
                                                      MbPE *pe, int32_t *ordPlus, int32_t *ordMinus, MbRow *rows, MbCall *grpCalls, MbExon *exons, int32_t *grpCnt) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    uint32_t lo = 0, hi = nContigs;   // candOff[lo] <= g < candOff[hi]
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (candOff[mid] <= g) lo = mid; else hi = mid; }
    const uint64_t b = candOff[lo], e = candOff[lo + 1];
    const uint32_t key = cands[g].targetKey;
    if (g > b && cands[g - 1].targetKey == key) { grpCnt[g] = 0; return; }        // not the first candidate of its group
    uint64_t j = g + 1; while (j < e && cands[j].targetKey == key) j++;
    grpCnt[g] = mb_chain_group(cands + g, (uint32_t)(j - g), (uint32_t)(g - b), P, pe + g, ordPlus + g, ordMinus + g, rows + g, grpCalls + g, exons + g);
}
__global__ void __launch_bounds__(128) k_chain_compact(const uint64_t *__restrict__ candOff, uint32_t nContigs, const int32_t *__restrict__ grpCnt, const MbCall *__restrict__ grpCalls,
                                                       MbCall *calls, int32_t *callCount) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nContigs) return;
    const uint64_t b = candOff[c], e = candOff[c + 1];
    int32_t nC = 0;
    for (uint64_t g = b; g < e; g++) {
        const int32_t cnt = grpCnt[g];
        if (cnt < 0) { nC = -1; break; }                                           // illegal codon segment length
        for (int32_t k = 0; k < cnt; k++) calls[2 * b + (uint64_t)nC++] = grpCalls[g + (uint64_t)k];
    }
    callCount[c] = nC;
}
// ---- reduceredundancy: one contig per thread ---------------------------------------------------------
__global__ void __launch_bounds__(128) k_reduce(const MbPred *__restrict__ preds, const uint64_t *__restrict__ predOff, uint32_t nContigs, const uint32_t *__restrict__ exonKeys,
                                                MbPredOut *out, uint32_t *memberOrder, uint32_t *repOrder, uint32_t *repCount, uint32_t *flag, uint32_t *tmp) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nContigs) return;
    const uint64_t b = predOff[c]; const uint32_t n = (uint32_t)(predOff[c + 1] - b);
    repCount[c] = mb_reduce_contig(preds + b, n, exonKeys, out + b, memberOrder + b, repOrder + b, flag + b, tmp + b);
}

// =====================================================================================================
// context
// =====================================================================================================
template <class T> struct DevBuf {
    T *p = nullptr; size_t n = 0; bool view = false;   // view: a slice of another allocation (the index arena), never freed here
    void set_view(void *ptr, size_t count) { if (p && !view) cudaFree(p); p = reinterpret_cast<T *>(ptr); n = count; view = true; }
    int alloc(size_t count) {
        if (count <= n && p) return 0;
        if (p && !view) cudaFree(p);
        p = nullptr; n = 0; view = false;
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) return fail(MB200_E_CUDA, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        n = count; return 0;
    }
    void release() { if (p && !view) cudaFree(p); p = nullptr; n = 0; view = false; }
};

template <class T> struct PinBuf {   // grow-only page-locked host staging buffer
    T *p = nullptr; size_t n = 0;
    int alloc(size_t count) {
        if (count <= n && p) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        if (count == 0) count = 1;
        count += count / 4;
        cudaError_t e = cudaHostAlloc((void **)&p, count * sizeof(T), cudaHostAllocDefault);
        if (e != cudaSuccess) return fail(MB200_E_CUDA, "cudaHostAlloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        n = count; return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};


// ---- NCCL, resolved at run time: the library has no link-time dependency on libnccl (a process that already loaded one -- e.g. torch's bundled
//      copy -- is joined through RTLD_NOLOAD so communicators and ours come from the same instance; MB200_NCCL_LIB overrides the name) ----
#ifndef MB200_SIMT_EMU
#include <dlfcn.h>
#include <nccl.h>
struct NcclApi {
    void *h = nullptr; bool tried = false;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
    NcclApi &N = g_nccl;
    if (N.h) return 0;
    if (N.tried) return fail(MB200_E_STATE, "libnccl.so.2 could not be loaded (set MB200_NCCL_LIB)");
    N.tried = true;
    const char *env = getenv("MB200_NCCL_LIB");
    if (env) N.h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!N.h) N.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!N.h) N.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!N.h) return fail(MB200_E_STATE, "libnccl.so.2 could not be loaded: %s", dlerror());
#define MB_NCCL_SYM(f) do { *(void **)(&N.f) = dlsym(N.h, "nccl" #f); if (!N.f) { N.h = nullptr; return fail(MB200_E_STATE, "nccl" #f " not found in libnccl"); } } while (0)
    MB_NCCL_SYM(GetUniqueId); MB_NCCL_SYM(CommInitRank); MB_NCCL_SYM(CommInitAll); MB_NCCL_SYM(CommDestroy); MB_NCCL_SYM(CommCount); MB_NCCL_SYM(CommUserRank);
    MB_NCCL_SYM(Broadcast); MB_NCCL_SYM(GetErrorString);
#undef MB_NCCL_SYM
    return 0;
}
#define NC(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail(MB200_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #x, g_nccl.GetErrorString(r_)); } while (0)

#else

#endif

struct mb200_ctx;
static void mb200_comm_release(mb200_ctx *c);

struct mb200_ctx {
    int device = 0; cudaStream_t stream = nullptr; int nSM = 148;
    mb200_params par; mb200::Matrices mat; bool haveMat = false; mb200::EvalueModel ev;
    // tables
    DevBuf<uint32_t> dSim3; DevBuf<uint16_t> dCnt3; DevBuf<uint32_t> dOffsets, dKmerBits; DevBuf<uint2> dEntries; bool useBits = false; DevBuf<uint4> dBuckets; bool haveBuckets = false; int optBuckets = 1, optStreamHint = 3, optSortWalk = 0, optOverflowPolicy = -1; bool inPredict = false;
    DevBuf<uint8_t> dTseqM, dTseqU; DevBuf<uint64_t> dToff;
    DevBuf<unsigned char> dIndexArena; size_t arenaBytes = 0;   // offsets | entries | masked | unmasked residues | target offsets in ONE allocation (one ncclBroadcast replicates it)
    void *comm = nullptr;                                        // ncclComm_t created by mb200_comm_init (owned)
    DevBuf<int16_t> dSeedMat, dAlnMat; DevBuf<double> dSeedPB, dAlnPB; DevBuf<int8_t> dUng8;
    uint64_t nEntries = 0, totalRes = 0; uint32_t nT = 0; std::vector<uint64_t> hToff; bool haveIndex = false, haveSim3 = false;
    // prefilter scratch
    int nCta = 0; DevBuf<MbPrefScratch> dScratch, dScratchShort; DevBuf<unsigned char> dScratchMem, dScratchShortMem; uint32_t capP = 0, capI = 0, capK = 0, capM = 0;
    int nCtaRescue = 0; DevBuf<MbPrefScratch> dScratchRescue; DevBuf<unsigned char> dScratchRescueMem; DevBuf<uint32_t> dOrderRescue; uint32_t rescueCapM = 0;   // second pass for queries whose work lists overflowed
    uint64_t rescuedQueries = 0;
    // resident query batch
    DevBuf<uint8_t> dQ; DevBuf<uint64_t> dQoff; DevBuf<float> dBias; DevBuf<int8_t> dCorr, dCb; DevBuf<uint32_t> dOrder;
    std::vector<uint64_t> hQoff; uint32_t nQ = 0; uint64_t totalQ = 0; uint32_t maxQLen = 0;
    DevBuf<MbPrefHit> dHits; DevBuf<uint32_t> dHitCount, dStatus, dCounter; DevBuf<MbPrefStats> dStats; DevBuf<unsigned long long> dCells;
    bool havePref = false;
    // align
    DevBuf<uint32_t> dPairOff, dChunkOff, dLaneOff, dWaveOff, dPassList, dHE, dFlatSeq; DevBuf<int32_t> dScore, dMinScore; DevBuf<SwTraceOut> dTrace; DevBuf<MbPrefHit> dFlatHits;
    DevBuf<uint64_t> dHitOff64; DevBuf<uint16_t> dPerm; DevBuf<uint32_t> dBnd;
    DevBuf<SwPair> dSwDesc, dSwDescRev; DevBuf<SwPassRec> dPassRec; DevBuf<SwTraceOut> dTrace2; DevBuf<uint32_t> dFallback;   // sw_kernel 6 (sw_stream.cuh)
    std::vector<int> minScore;   // per query length, -2 = not computed
    std::vector<uint32_t> hOrder;   // queries by descending length
    std::vector<SwTraceOut> hTrace;  // last alignment pass, sorted by slot
    std::vector<mb200_aln> hAlns; std::vector<uint64_t> hAlnOff; std::vector<uint64_t> hHitOff;
    // host work arrays of mb200_predict_exons / collect_accepted: owned by the context and never shrunk, so that a batch does not pay the page faults and the zero fill of
    // 100+ MB of fresh vectors (they were 40 % of the host-only time of a batch)
    std::vector<mb200_aln> hAlnTmp; std::vector<mb200_cand> peCands; std::vector<uint64_t> peCandOff, peCandAln; std::vector<mb200_call> peBCalls; std::vector<mb200_exon> peBExons; std::vector<uint32_t> peBCnt;
    // ORF extraction
    DevBuf<unsigned char> dNuc; DevBuf<uint64_t> dNucOff, dOrfAaCnt, dOrfRecBase, dOrfAaBase; DevBuf<uint32_t> dOrfCnt; DevBuf<MbOrfTables> dOrfTab; bool haveOrfTab = false;
    DevBuf<MbOrfRec> dOrfRecs; DevBuf<char> dOrfAa; DevBuf<uint8_t> dOrfCodes; PinBuf<MbOrfRec> pOrfRecs; PinBuf<char> pOrfAa; uint64_t nOrfs = 0, nOrfAa = 0;
    PinBuf<mb200_hit> pFlatHits; PinBuf<uint32_t> pOff, pStatus; size_t nFlatHits = 0;
    bool haveAln = false;
    // predictexons with binary hand-off
    DevBuf<MbCand> dChCand; DevBuf<uint64_t> dChOff; DevBuf<MbPE> dChPE; DevBuf<int32_t> dChOp, dChOm, dChCallCnt; DevBuf<MbRow> dChRows; DevBuf<MbCall> dChCalls; DevBuf<MbExon> dChEx;
    DevBuf<uint32_t> dChExCnt; DevBuf<MbCall> dChGrpCalls; DevBuf<int32_t> dChGrpCnt; PinBuf<MbCall> pChCalls; PinBuf<MbExon> pChEx; PinBuf<int32_t> pChCnt;   // exon chaining (mb200_chain)
    std::vector<uint32_t> tKeys; std::vector<mb200_call> peCalls; std::vector<mb200_pexon> peExons; uint64_t optBatchNt = 32ull << 20, optOrfKeyBase = 0; bool optBatchNtSet = false; int optStreams = 2; mb200_ctx *workers[4] = {nullptr, nullptr, nullptr, nullptr};
    int optPrefCfg = 0 /* 0 auto, 1 dense, 2 sparse, 3 medium, 4 small */, prefCfg = 0 /* config the scratch was sized for */, optPrefShortLen = 0, optPrefShortCfg = 3, prefCfgShort = 0, nCtaShort = 0, optPhaseStats = 0, optSwKernel = 6, optSwRing = 0 /* ring profiles per warp: 0 auto, -1 private profile per slot */, optSwRingFit = 1, optTraceKernel = 2, optLaneMaxHits = 8;   // mb200_option_set
    mb200_stats st;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evT0 = nullptr, evT1 = nullptr;
    cudaEvent_t evOrigin = nullptr, originRef = nullptr;    // mb200_predict_exons: origin of the call's device timeline (owned by the caller's context; workers hold a reference)
    std::vector<std::pair<float, float>> busyIv;            // stage intervals of this context on that timeline (ms_dev_busy)
    cudaStream_t swSide[4] = {nullptr, nullptr, nullptr, nullptr}; cudaEvent_t swFork = nullptr, swJoin[4] = {nullptr, nullptr, nullptr, nullptr};   // the alignment classes of one batch run side by side
};

static int upload(void *dst, const void *src, size_t bytes, cudaStream_t s) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s)); return 0;
}
#define TRY(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

static int ctx_upload_matrices(mb200_ctx *c) {
    TRY(c->dSeedMat.alloc(441)); TRY(c->dAlnMat.alloc(441)); TRY(c->dSeedPB.alloc(21)); TRY(c->dAlnPB.alloc(21)); TRY(c->dUng8.alloc(448));
    int8_t u8[448]; memset(u8, 0, sizeof(u8)); for (int i = 0; i < 441; i++) u8[i] = (int8_t)c->mat.ung[i];
    TRY(upload(c->dSeedMat.p, c->mat.seed, 882, c->stream)); TRY(upload(c->dAlnMat.p, c->mat.aln, 882, c->stream));
    TRY(upload(c->dSeedPB.p, c->mat.seedPBack, 168, c->stream)); TRY(upload(c->dAlnPB.p, c->mat.alnPBack, 168, c->stream)); TRY(upload(c->dUng8.p, u8, 448, c->stream));
    // similarity tables depend on the seed matrix only
    std::vector<uint32_t> sim3; std::vector<uint16_t> cnt3;
    mb200::build_sim3(c->mat.seed, sim3, cnt3, MB_CNT_MIN, MB_CNT_W);
    TRY(c->dSim3.alloc(sim3.size())); TRY(c->dCnt3.alloc(cnt3.size()));
    TRY(upload(c->dSim3.p, sim3.data(), sim3.size() * 4, c->stream)); TRY(upload(c->dCnt3.p, cnt3.data(), cnt3.size() * 2, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->haveMat = true; c->haveSim3 = true;
    return 0;
}

// per-CTA work lists of k_prefilter for `n` CTAs with the given capacities, carved out of one allocation
static int pref_scratch_build(mb200_ctx *c, int n, uint32_t capP, uint32_t capI, uint32_t capM, DevBuf<unsigned char> &mem, DevBuf<MbPrefScratch> &dS) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t per = al((capP + 1) * 4) + al(capP * 2) * 3 + al(((size_t)capI + 1) * 4) + al((size_t)capI * 4) + al((size_t)capM * 8) * 3 + al((size_t)capM * 4);
    TRY(mem.alloc(per * n)); TRY(dS.alloc(n));
    std::vector<MbPrefScratch> hs(n);
    for (int i = 0; i < n; i++) {
        unsigned char *b = mem.p + per * i; MbPrefScratch &W = hs[i];
        W.itemOff = (uint32_t *)b; b += al((capP + 1) * 4); W.wThr = (int16_t *)b; b += al(capP * 2); W.wI0 = (uint16_t *)b; b += al(capP * 2); W.wI1 = (uint16_t *)b; b += al(capP * 2);
        W.kmerOff = (uint32_t *)b; b += al(((size_t)capI + 1) * 4); W.itemInfo = (uint32_t *)b; b += al((size_t)capI * 4);
        W.A = (uint64_t *)b; b += al((size_t)capM * 8); W.B = (uint64_t *)b; b += al((size_t)capM * 8); W.C = (uint64_t *)b; b += al((size_t)capM * 8); W.S = (int32_t *)b;
    }
    TRY(upload(dS.p, hs.data(), hs.size() * sizeof(MbPrefScratch), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- prefilter kernel configurations (prefilter_core.cuh): 1 dense, 2 sparse, 3 medium, 4 small
template <class SH> static int pref_cfg_ctas(int nSM, int *nCta) {
    int perSM = 0;
    CU(cudaFuncSetAttribute(k_prefilter<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SH)));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_prefilter<SH>, SH::Cfg::NT, sizeof(SH)));
    if (perSM > (int)SH::Cfg::CTAS) perSM = SH::Cfg::CTAS;
    if (perSM < 1) return fail(MB200_E_INTERNAL, "prefilter kernel does not fit on an SM");
    *nCta = nSM * perSM; return 0;
}
static int pref_cfg_ctas(int cfg, int nSM, int *nCta) {
    switch (cfg) {
    case 1: return pref_cfg_ctas<MbPrefSharedT<MbPrefCfgDense>>(nSM, nCta);
    case 2: return pref_cfg_ctas<MbPrefSharedT<MbPrefCfgSparse>>(nSM, nCta);
    case 3: return pref_cfg_ctas<MbPrefSharedT<MbPrefCfgMedium>>(nSM, nCta);
    default: return pref_cfg_ctas<MbPrefSharedT<MbPrefCfgSmall>>(nSM, nCta);
    }
}
template <class SH> static void pref_launch_t(mb200_ctx *c, const MbPrefTables &T, const MbPrefScratch *scratch, int nCta, const uint32_t *order, uint32_t n, uint32_t *counter) {
    MB_LAUNCH((k_prefilter<SH>), nCta, SH::Cfg::NT, sizeof(SH), c->stream, T, scratch, c->dQ.p, c->dQoff.p, c->dBias.p, c->dCorr.p, order, n, c->dUng8.p, c->dHits.p, c->dHitCount.p,
              c->dStatus.p, c->dStats.p, counter);
}
static void pref_launch(mb200_ctx *c, int cfg, const MbPrefTables &T, const MbPrefScratch *scratch, int nCta, const uint32_t *order, uint32_t n, uint32_t *counter) {
    switch (cfg) {
    case 1: pref_launch_t<MbPrefSharedT<MbPrefCfgDense>>(c, T, scratch, nCta, order, n, counter); break;
    case 2: pref_launch_t<MbPrefSharedT<MbPrefCfgSparse>>(c, T, scratch, nCta, order, n, counter); break;
    case 3: pref_launch_t<MbPrefSharedT<MbPrefCfgMedium>>(c, T, scratch, nCta, order, n, counter); break;
    default: pref_launch_t<MbPrefSharedT<MbPrefCfgSmall>>(c, T, scratch, nCta, order, n, counter); break;
    }
}

static int ctx_alloc_scratch(mb200_ctx *c) {
    const uint32_t capP = 32768, capI = 1u << (c->par.cap_items_log2 ? c->par.cap_items_log2 : 17), capK = 1u << (c->par.cap_kmers_log2 ? c->par.cap_kmers_log2 : 18),
                   capM = 1u << (c->par.cap_hits_log2 ? c->par.cap_hits_log2 : 19);
    if (capK > (1u << 21)) return fail(MB200_E_ARG, "cap_kmers_log2 > 21 (the k-mer ordinal is a 21-bit field)");
    if (capI > (1u << 24) || capM > (1u << 24)) return fail(MB200_E_ARG, "cap_items_log2 / cap_hits_log2 must be <= 24");
    // dense config for indexes with long k-mer lists (>= 1 entry per k-mer slot on average), sparse config otherwise; short queries may take their own
    // configuration in a second launch (prefilter_short_len / prefilter_short_config)
    const bool denseIndex = (double)c->nEntries >= 1.0 * (double)mb200::KMER_TABLE;
    const int cfg = c->optPrefCfg ? c->optPrefCfg : (denseIndex ? 1 : 2);
    const int cfgS = (c->optPrefShortLen > 0 && c->optPrefShortCfg != cfg) ? c->optPrefShortCfg : 0;
    if (c->nCta && cfg == c->prefCfg && cfgS == c->prefCfgShort && capI == c->capI && capK == c->capK && capM == c->capM) return 0;
    TRY(pref_cfg_ctas(cfg, c->nSM, &c->nCta));
    TRY(pref_scratch_build(c, c->nCta, capP, capI, capM, c->dScratchMem, c->dScratch));
    c->nCtaShort = 0;
    if (cfgS) {
        // short queries: work lists sized for them (a query longer than prefilter_short_len never gets here; an overflow still takes the second pass)
        TRY(pref_cfg_ctas(cfgS, c->nSM, &c->nCtaShort));
        TRY(pref_scratch_build(c, c->nCtaShort, capP, std::min(capI, 1u << 15), std::min(capM, 1u << 17), c->dScratchShortMem, c->dScratchShort));
    }
    c->capP = capP; c->capI = capI; c->capK = capK; c->capM = capM; c->prefCfg = cfg; c->prefCfgShort = cfgS;
    return 0;
}

static int sync_check(mb200_ctx *c, const char *what) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(MB200_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}
// a finished stage [a, b] (both events complete) on the device timeline of the running mb200_predict_exons call
static void stage_interval(mb200_ctx *c, cudaEvent_t a, cudaEvent_t b) {
    if (!c->originRef) return;
    float s0 = 0, s1 = 0;
    if (cudaEventElapsedTime(&s0, c->originRef, a) != cudaSuccess || cudaEventElapsedTime(&s1, c->originRef, b) != cudaSuccess) { cudaGetLastError(); return; }
    c->busyIv.emplace_back(s0, s1);
}

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char *mb200_last_error(void) { return g_err.c_str(); }

void mb200_params_default(mb200_params *p) {
    memset(p, 0, sizeof(*p));
    p->kmer_thr = 127; p->max_res_list_len = 300; p->min_diag_score = 15; p->bin_size = 2;
    p->gap_open = 11; p->gap_extend = 1; p->min_aln_len = 11; p->eval_thr = 100.0;
    p->min_intron = 15; p->max_intron = 10000; p->min_exon_aa = 11; p->max_overlap = 10; p->set_gap_open = -1; p->set_gap_extend = -1; p->max_exon_sets = 1;
    p->tcov = 0.5; p->metaeuk_eval = 0.001;
    p->comp_bias_corr = 1; p->reverse_fragments = 0; p->translation_table = 1;
}

const char *mb200_build_kind(void) {
#ifdef MB200_SIMT_EMU
    return "simt-emulation";     // tests/simt: kernels run as CPU fibers -- test infrastructure, never measured, never shipped
#else
    return "sm_100a";
#endif
}

int mb200_kmer_threshold(float sensitivity) { return mb200::kmer_threshold_for_sensitivity(sensitivity); }

int mb200_ctx_create(int device, mb200_ctx **out) {
    if (!out) return fail(MB200_E_ARG, "out == NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(MB200_E_CUDA, "no CUDA device (%s): libmetaeuk_b200 has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(MB200_E_ARG, "device %d out of range (%d devices)", device, n);
    CU(cudaSetDevice(device));
    mb200_ctx *c = new mb200_ctx();
    c->device = device;
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device)); c->nSM = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1)); CU(cudaEventCreate(&c->evT0)); CU(cudaEventCreate(&c->evT1));
    mb200_params_default(&c->par); memset(&c->st, 0, sizeof(c->st));
    *out = c; return 0;
}

void mb200_ctx_destroy(mb200_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int k = 0; k < 4; k++) { mb200_ctx_destroy(c->workers[k]); c->workers[k] = nullptr; }   // (their index / matrix buffers are views of ours)
    for (int k = 0; k < 4; k++) { if (c->swSide[k]) cudaStreamDestroy(c->swSide[k]); if (c->swJoin[k]) cudaEventDestroy(c->swJoin[k]); c->swSide[k] = nullptr; c->swJoin[k] = nullptr; }
    if (c->swFork) { cudaEventDestroy(c->swFork); c->swFork = nullptr; }
    mb200_comm_release(c); c->dIndexArena.release(); c->dSim3.release(); c->dCnt3.release(); c->dOffsets.release(); c->dKmerBits.release(); c->dEntries.release(); c->dBuckets.release(); c->dTseqM.release(); c->dTseqU.release(); c->dToff.release();
    c->dSeedMat.release(); c->dAlnMat.release(); c->dSeedPB.release(); c->dAlnPB.release(); c->dUng8.release(); c->dScratch.release(); c->dScratchMem.release(); c->dScratchShort.release(); c->dScratchShortMem.release(); c->dScratchRescue.release(); c->dScratchRescueMem.release(); c->dOrderRescue.release();
    c->dQ.release(); c->dQoff.release(); c->dBias.release(); c->dCorr.release(); c->dCb.release(); c->dOrder.release(); c->dHits.release(); c->dHitCount.release();
    c->dStatus.release(); c->dCounter.release(); c->dStats.release(); c->dCells.release(); c->dHE.release(); c->dPairOff.release(); c->dChunkOff.release(); c->dLaneOff.release(); c->dWaveOff.release(); c->dPassList.release();
    c->dFlatSeq.release(); c->dScore.release(); c->dMinScore.release(); c->dTrace.release(); c->dFlatHits.release(); c->dHitOff64.release(); c->dPerm.release(); c->dBnd.release();
    c->dSwDesc.release(); c->dSwDescRev.release(); c->dPassRec.release(); c->dTrace2.release(); c->dFallback.release();
    c->pFlatHits.release(); c->pOff.release(); c->pStatus.release(); c->pOrfRecs.release(); c->pOrfAa.release();
    c->dChCand.release(); c->dChOff.release(); c->dChPE.release(); c->dChOp.release(); c->dChOm.release(); c->dChCallCnt.release(); c->dChRows.release(); c->dChCalls.release();
    c->dChEx.release(); c->dChExCnt.release(); c->dChGrpCalls.release(); c->dChGrpCnt.release(); c->pChCalls.release(); c->pChEx.release(); c->pChCnt.release();
    c->dNuc.release(); c->dNucOff.release(); c->dOrfAaCnt.release(); c->dOrfRecBase.release(); c->dOrfAaBase.release(); c->dOrfCnt.release(); c->dOrfTab.release(); c->dOrfRecs.release(); c->dOrfAa.release(); c->dOrfCodes.release();
    if (c->evOrigin) cudaEventDestroy(c->evOrigin); if (c->ev0) cudaEventDestroy(c->ev0); if (c->ev1) cudaEventDestroy(c->ev1); if (c->evT0) cudaEventDestroy(c->evT0); if (c->evT1) cudaEventDestroy(c->evT1); if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int mb200_option_set(mb200_ctx *c, const char *name, int64_t value) {
    if (!c || !name) return fail(MB200_E_ARG, "null argument");
    const std::string n(name);
    if (n == "phase_stats") c->optPhaseStats = value ? 1 : 0;
    else if (n == "overflow_policy") { if (value < -1 || value > 1) return fail(MB200_E_ARG, "overflow_policy: -1 (default per entry point), 0 (fail the call) or 1 (skip the query, count it)"); c->optOverflowPolicy = (int)value; }
    else if (n == "prefilter_stream_hint") { if (value < 0 || value > 3) return fail(MB200_E_ARG, "prefilter_stream_hint: bit 0 = index reads L1 no-allocate + L2 evict-first, bit 1 = 64-byte HBM fetch (.L2::64B)"); c->optStreamHint = (int)value; }
    else if (n == "l2_fetch_granularity") {   // diagnostic: cudaLimitMaxL2FetchGranularity of the context's device (32, 64 or 128 bytes; a hint the driver may ignore)
        if (value != 32 && value != 64 && value != 128) return fail(MB200_E_ARG, "l2_fetch_granularity: 32, 64 or 128");
#ifndef MB200_SIMT_EMU
        CU(cudaSetDevice(c->device)); CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
#endif
    }
    else if (n == "prefilter_sort_walk") { if (value < 0 || value > 256) return fail(MB200_E_ARG, "prefilter_sort_walk: 0..256"); c->optSortWalk = (int)value; }
    else if (n == "prefilter_buckets") { if (value != 0 && value != 1) return fail(MB200_E_ARG, "prefilter_buckets: 0 (offsets + entries arrays) or 1 (inline bucket table)"); c->optBuckets = (int)value; }
    else if (n == "prefilter_config") { if (value < 0 || value > 4) return fail(MB200_E_ARG, "prefilter_config: 0 (auto), 1 (dense), 2 (sparse), 3 (medium) or 4 (small)"); c->optPrefCfg = (int)value; }
    else if (n == "prefilter_short_len") { if (value < 0 || value > 32767) return fail(MB200_E_ARG, "prefilter_short_len: 0 (off) .. 32767"); c->optPrefShortLen = (int)value; }
    else if (n == "prefilter_short_config") { if (value < 1 || value > 4) return fail(MB200_E_ARG, "prefilter_short_config: 1..4"); c->optPrefShortCfg = (int)value; }
    else if (n == "sw_kernel") { if (value < 1 || value > 6 || value == 3) return fail(MB200_E_ARG, "sw_kernel: 1 (one target per lane), 2 (two, packed 16-bit), 4 (2 + one short query with few hits per lane), 5 (4 + row blocks of an alignment spread over lanes) or 6 (streams of pairs over persistent lane groups, end cells recorded by the forward pass); 3 (per-lane state machine) lost its A/B and was removed"); c->optSwKernel = (int)value; }
    else if (n == "sw_ring_fit") { if (value != 0 && value != 1) return fail(MB200_E_ARG, "sw_ring_fit: 0 or 1"); c->optSwRingFit = (int)value; }
    else if (n == "sw_ring") { if (value < -1 || value > 16) return fail(MB200_E_ARG, "sw_ring: -1 (private profile per slot), 0 (auto) or 1..16 ring profiles per warp"); c->optSwRing = (int)value; }
    else if (n == "sw_lane_max_hits") { if (value < 1 || value > 1024) return fail(MB200_E_ARG, "sw_lane_max_hits: 1..1024"); c->optLaneMaxHits = (int)value; }
    else if (n == "trace_kernel") { if (value != 2) return fail(MB200_E_ARG, "trace_kernel: 2 (k_sw_trace2; variant 1 was removed)"); c->optTraceKernel = (int)value; }
    else if (n == "predict_batch_nt") { if (value < 1) return fail(MB200_E_ARG, "predict_batch_nt must be >= 1"); c->optBatchNt = (uint64_t)value; c->optBatchNtSet = true; }
    else if (n == "predict_streams") { if (value < 1 || value > 4) return fail(MB200_E_ARG, "predict_streams: 1..4"); c->optStreams = (int)value; }
    else if (n == "orf_key_base") { if (value < 0 || value >= (1ll << 32)) return fail(MB200_E_ARG, "orf_key_base must fit 32 bits"); c->optOrfKeyBase = (uint64_t)value; }   // contig shards: fragment keys count over the whole contig DB
    else return fail(MB200_E_ARG, "unknown option '%s'", name);
    return 0;
}

int mb200_timer_start(mb200_ctx *c) {
    if (!c) return fail(MB200_E_ARG, "null argument");
    CU(cudaSetDevice(c->device)); CU(cudaStreamSynchronize(c->stream)); CU(cudaEventRecord(c->evT0, c->stream)); return 0;
}

int mb200_timer_stop(mb200_ctx *c, double *ms) {
    if (!c || !ms) return fail(MB200_E_ARG, "null argument");
    CU(cudaSetDevice(c->device)); CU(cudaEventRecord(c->evT1, c->stream)); CU(cudaEventSynchronize(c->evT1));
    float f = 0; CU(cudaEventElapsedTime(&f, c->evT0, c->evT1)); *ms = f; return 0;
}

int mb200_params_set(mb200_ctx *c, const mb200_params *p) {
    if (!c || !p) return fail(MB200_E_ARG, "null argument");
    if (p->min_diag_score < 1) return fail(MB200_E_ARG, "min_diag_score must be >= 1");
    if (p->bin_size < 1 || (p->bin_size & (p->bin_size - 1)) || p->bin_size > 2048) return fail(MB200_E_ARG, "bin_size must be a power of two <= 2048");
    if (p->max_res_list_len < 1 || p->max_res_list_len > 65535) return fail(MB200_E_ARG, "max_res_list_len out of range");
    if (p->gap_open != c->par.gap_open || p->gap_extend != c->par.gap_extend || p->eval_thr != c->par.eval_thr) c->minScore.clear();
    if ((p->gap_open != 11 || p->gap_extend != 1)) return fail(MB200_E_ARG, "only the reference's fixed BLOSUM62 11/1 e-value model is implemented");
    if (!mb200::genetic_code(p->translation_table)) return fail(MB200_E_ARG, "translation_table %d is not an NCBI genetic code the reference knows (1-6, 9-16, 21-31)", p->translation_table);
    if (p->translation_table != c->par.translation_table) c->haveOrfTab = false;
    c->par = *p; return 0;
}

int mb200_matrices_default(mb200_ctx *c) {
    if (!c) return fail(MB200_E_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    mb200::default_matrices(c->mat); return ctx_upload_matrices(c);
}
int mb200_matrices_set(mb200_ctx *c, const int16_t *seed, const int16_t *ung, const int16_t *aln, const double *spb, const double *apb) {
    if (!c || !seed || !ung || !aln || !spb || !apb) return fail(MB200_E_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    memcpy(c->mat.seed, seed, 882); memcpy(c->mat.ung, ung, 882); memcpy(c->mat.aln, aln, 882); memcpy(c->mat.seedPBack, spb, 168); memcpy(c->mat.alnPBack, apb, 168);
    return ctx_upload_matrices(c);
}
int mb200_matrices_get(mb200_ctx *c, int16_t *seed, int16_t *ung, int16_t *aln) {
    if (!c || !c->haveMat) return fail(MB200_E_STATE, "matrices not set");
    memcpy(seed, c->mat.seed, 882); memcpy(ung, c->mat.ung, 882); memcpy(aln, c->mat.aln, 882); return 0;
}

static int check_targets(const uint64_t *off, uint32_t nT) {
    if (nT == 0) return fail(MB200_E_ARG, "no targets");
    if (nT >= (1u << MB_MAX_SEQID_BITS)) return fail(MB200_E_LIMIT, "more than 2^%d targets", MB_MAX_SEQID_BITS);
    if (off[nT] >= (1ull << 32)) return fail(MB200_E_LIMIT, "more than 2^32 target residues");
    for (uint32_t t = 0; t < nT; t++) if (off[t + 1] - off[t] >= 32768) return fail(MB200_E_LIMIT, "target %u has >= 32768 residues", t);
    return 0;
}
// The replicated index lives in one device allocation so that ONE ncclBroadcast moves it (mb200_index_broadcast); the typed buffers are views.
static int index_arena_alloc(mb200_ctx *c, uint64_t nEntries, uint32_t nT, uint64_t totalRes) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t bOff = al((mb200::KMER_TABLE + 1) * 4), bEnt = al((size_t)std::max<uint64_t>(nEntries, 1) * 8), bSeq = al(totalRes + 16), bTo = al((size_t)(nT + 1) * 8);
    const size_t total = bOff + bEnt + 2 * bSeq + bTo;
    TRY(c->dIndexArena.alloc(total));
    unsigned char *b = c->dIndexArena.p;
    c->dOffsets.set_view(b, mb200::KMER_TABLE + 1); b += bOff;
    c->dEntries.set_view(b, std::max<uint64_t>(nEntries, 1)); b += bEnt;
    c->dTseqM.set_view(b, totalRes + 16); b += bSeq;
    c->dTseqU.set_view(b, totalRes + 16); b += bSeq;
    c->dToff.set_view(b, nT + 1);
    c->arenaBytes = total;
    return 0;
}
static int upload_targets(mb200_ctx *c, const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t nT) {
    const uint64_t tot = off[nT];
    TRY(upload(c->dTseqM.p, masked, tot, c->stream)); TRY(upload(c->dTseqU.p, unmasked, tot, c->stream)); TRY(upload(c->dToff.p, off, (nT + 1) * 8, c->stream));
    c->hToff.assign(off, off + nT + 1); c->nT = nT; c->totalRes = tot; c->minScore.clear();
    return 0;
}
// bitmap of k-mers whose index list is non-empty: 8 MB, L2 resident; consulted before the random read of offsets[]
__global__ void k_kmer_bits(const uint32_t *__restrict__ offsets, uint32_t nWords, uint32_t *__restrict__ bits) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nWords) return;
    uint32_t m = 0;
    for (uint32_t b = 0; b < 32; b++) { const uint32_t k = w * 32u + b; if (offsets[k + 1] != offsets[k]) m |= 1u << b; }
    bits[w] = m;
}
// inline bucket table (prefilter_core.cuh): 2 GB, built on the device from offsets / entries when the prefilter first needs it
__global__ void k_build_buckets(const uint32_t *__restrict__ offsets, const uint2 *__restrict__ entries, uint32_t nKmers, uint4 *__restrict__ buckets) {
    for (uint32_t km = blockIdx.x * blockDim.x + threadIdx.x; km < nKmers; km += gridDim.x * blockDim.x) {
    const uint32_t lo = offsets[km], len = offsets[km + 1] - lo;
    uint32_t h[16];
#pragma unroll
    for (int i = 0; i < 16; i++) h[i] = 0;
    h[0] = len < 0xFFFFu ? len : 0xFFFFu;
    if (len <= MB_BUCKET_INLINE) {
#pragma unroll
        for (int j = 0; j < MB_BUCKET_INLINE; j++) if ((uint32_t)j < len) { const uint2 e = entries[(size_t)lo + j]; h[1 + 3 * j] = e.x & 0xFFFFu; h[2 + 3 * j] = e.x >> 16; h[3 + 3 * j] = e.y & 0xFFFFu; }
    } else { h[2] = lo & 0xFFFFu; h[3] = lo >> 16; h[4] = len & 0xFFFFu; h[5] = len >> 16; }
    uint4 a, b;
    a.x = h[0] | (h[1] << 16); a.y = h[2] | (h[3] << 16); a.z = h[4] | (h[5] << 16); a.w = h[6] | (h[7] << 16);
    b.x = h[8] | (h[9] << 16); b.y = h[10] | (h[11] << 16); b.z = h[12] | (h[13] << 16); b.w = h[14] | (h[15] << 16);
    buckets[2 * (size_t)km] = a; buckets[2 * (size_t)km + 1] = b;
    }
}
static int build_buckets(mb200_ctx *c) {
    if (c->haveBuckets || !c->optBuckets) return 0;
    TRY(c->dBuckets.alloc(2 * (size_t)mb200::KMER_TABLE));
    MB_LAUNCH((k_build_buckets), c->nSM * 16, 256, 0, c->stream, c->dOffsets.p, c->dEntries.p, (uint32_t)mb200::KMER_TABLE, c->dBuckets.p);
    CU(cudaGetLastError()); CU(cudaStreamSynchronize(c->stream));
    c->haveBuckets = true; return 0;
}
static int build_kmer_bits(mb200_ctx *c) {
    c->useBits = false; c->haveBuckets = false;
    if (c->nEntries >= mb200::KMER_TABLE) return 0;   // dense index: nearly every k-mer has a list, the extra lookup would not pay
    const uint32_t nWords = (uint32_t)(mb200::KMER_TABLE / 32);
    TRY(c->dKmerBits.alloc(nWords));
    MB_LAUNCH((k_kmer_bits), (nWords + 255) / 256, 256, 0, c->stream, c->dOffsets.p, nWords, c->dKmerBits.p);
    CU(cudaGetLastError()); CU(cudaStreamSynchronize(c->stream));
    c->useBits = true; return 0;
}
static int upload_index(mb200_ctx *c, const mb200::TargetIndex &idx) {
    TRY(upload(c->dOffsets.p, idx.offsets.data(), idx.offsets.size() * 4, c->stream));
    TRY(upload(c->dEntries.p, idx.entries.data(), idx.entries.size() * 8, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->nEntries = idx.entries.size(); c->haveIndex = true;
    return build_kmer_bits(c);
}

int mb200_index_build(mb200_ctx *c, const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t nT) {
    if (!c || !masked || !unmasked || !off) return fail(MB200_E_ARG, "null argument");
    if (!c->haveMat) return fail(MB200_E_STATE, "set the matrices before building the index");
    CU(cudaSetDevice(c->device));
    TRY(check_targets(off, nT));
    mb200::TargetIndex idx;
    if (mb200::build_index(masked, off, nT, c->mat.seed, c->par.kmer_thr, idx)) return fail(MB200_E_LIMIT, "index build failed");
    TRY(index_arena_alloc(c, idx.entries.size(), nT, off[nT]));
    TRY(upload_targets(c, masked, unmasked, off, nT));
    return upload_index(c, idx);
}

int mb200_index_import(mb200_ctx *c, const uint64_t *offsets, const uint8_t *entries6, uint64_t nEntries, const uint8_t *masked, const uint8_t *unmasked,
                       const uint64_t *off, uint32_t nT) {
    if (!c || !offsets || !entries6 || !masked || !unmasked || !off) return fail(MB200_E_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    TRY(check_targets(off, nT));
    mb200::TargetIndex idx;
    if (mb200::import_index(offsets, entries6, nEntries, idx)) return fail(MB200_E_LIMIT, "index has >= 2^32 entries");
    TRY(index_arena_alloc(c, idx.entries.size(), nT, off[nT]));
    TRY(upload_targets(c, masked, unmasked, off, nT));
    return upload_index(c, idx);
}

// ---- target masking (host only) ----
int mb200_mask_targets(const uint8_t *unmasked, const uint64_t *off, uint32_t nT, double maskProb, int simdWidth, uint8_t *masked, uint64_t *nMasked) {
    if (!unmasked || !off || !masked) return fail(MB200_E_ARG, "null argument");
    if (simdWidth != 0 && simdWidth != 2 && simdWidth != 4) return fail(MB200_E_ARG, "simd_width: 0, 2 or 4");
    for (uint64_t i = off[0]; i < off[nT]; i++) if (unmasked[i] > 20) return fail(MB200_E_ARG, "residue code %u at position %llu is not in 0..20", unmasked[i], (unsigned long long)i);
    const uint64_t n = mb200::mask_targets(unmasked, off, nT, maskProb, masked, simdWidth ? simdWidth : 4);
    if (nMasked) *nMasked = n;
    return 0;
}

int mb200_mask_targets_ex(const uint8_t *unmasked, const uint64_t *off, uint32_t nT, int maskTantan, double maskProb, int simdWidth, const uint8_t *lowerCase, int maskNRepeat,
                          uint8_t *masked, uint64_t *nMasked) {
    if (!unmasked || !off || !masked) return fail(MB200_E_ARG, "null argument");
    if (simdWidth != 0 && simdWidth != 2 && simdWidth != 4) return fail(MB200_E_ARG, "simd_width: 0, 2 or 4");
    if (maskNRepeat < 0) return fail(MB200_E_ARG, "mask_n_repeat < 0");
    for (uint64_t i = off[0]; i < off[nT]; i++) if (unmasked[i] > 20) return fail(MB200_E_ARG, "residue code %u at position %llu is not in 0..20", unmasked[i], (unsigned long long)i);
    const uint64_t n = mb200::mask_targets_ex(unmasked, off, nT, maskTantan != 0, maskProb, simdWidth ? simdWidth : 4, lowerCase, maskNRepeat, masked);
    if (nMasked) *nMasked = n;
    return 0;
}

// ---- persisted index ----
int mb200_index_file_write(const char *path, const uint8_t *masked, const uint8_t *unmasked, const uint64_t *off, uint32_t nT, const uint32_t *keys, int kmerThr, int maskMode) {
    if (!path || !masked || !unmasked || !off) return fail(MB200_E_ARG, "null argument");
    TRY(check_targets(off, nT));
    mb200::Matrices m; mb200::default_matrices(m);
    mb200::TargetIndex idx;
    if (mb200::build_index(masked, off, nT, m.seed, kmerThr, idx)) return fail(MB200_E_LIMIT, "index build failed");
    std::string err;
    if (mb200::write_index_db(path, idx, masked, unmasked, off, nT, keys, kmerThr, maskMode, err)) return fail(MB200_E_ARG, "%s", err.c_str());
    return 0;
}

int mb200_index_save(mb200_ctx *c, const char *path) {
    if (!c || !path) return fail(MB200_E_ARG, "null argument");
    if (!c->haveIndex) return fail(MB200_E_STATE, "no index");
    CU(cudaSetDevice(c->device));
    mb200::TargetIndex idx; idx.offsets.resize(mb200::KMER_TABLE + 1); idx.entries.resize(c->nEntries);
    std::vector<uint8_t> m(c->totalRes + 1), u(c->totalRes + 1);
    CU(cudaMemcpy(idx.offsets.data(), c->dOffsets.p, idx.offsets.size() * 4, cudaMemcpyDeviceToHost));
    if (c->nEntries) CU(cudaMemcpy(idx.entries.data(), c->dEntries.p, c->nEntries * 8, cudaMemcpyDeviceToHost));
    if (c->totalRes) { CU(cudaMemcpy(m.data(), c->dTseqM.p, c->totalRes, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(u.data(), c->dTseqU.p, c->totalRes, cudaMemcpyDeviceToHost)); }
    bool masked = false; for (uint64_t i = 0; i < c->totalRes && !masked; i++) masked = m[i] != u[i];
    std::string err;
    if (mb200::write_index_db(path, idx, m.data(), u.data(), c->hToff.data(), c->nT, c->tKeys.size() == c->nT ? c->tKeys.data() : nullptr, c->par.kmer_thr, masked ? 1 : 0, err, c->haveMat ? c->mat.seed : nullptr))
        return fail(MB200_E_ARG, "%s", err.c_str());
    return 0;
}

int mb200_index_load(mb200_ctx *c, const char *path) {
    if (!c || !path) return fail(MB200_E_ARG, "null argument");
    mb200::IndexDb db; std::string err;
    if (mb200::open_index_db(path, db, err)) return fail(MB200_E_ARG, "%s: %s", path, err.c_str());
    if (db.meta.kmerThr != c->par.kmer_thr) return fail(MB200_E_ARG, "%s was built with k-mer threshold %d, the context uses %d", path, db.meta.kmerThr, c->par.kmer_thr);
    if (db.nSeq >= (1ull << 27)) return fail(MB200_E_LIMIT, "too many targets");
    TRY(mb200_index_import(c, db.offsets, db.entries6, db.nEntries, db.masked, db.unmasked.data(), db.seqOff, (uint32_t)db.nSeq));
    c->tKeys = db.keys;
    return 0;
}

int mb200_target_keys_get(mb200_ctx *c, uint32_t *keys) {
    if (!c || !keys) return fail(MB200_E_ARG, "null argument");
    if (!c->haveIndex) return fail(MB200_E_STATE, "no index");
    for (uint32_t i = 0; i < c->nT; i++) keys[i] = c->tKeys.size() == c->nT ? c->tKeys[i] : i;
    return 0;
}

int mb200_index_sizes(mb200_ctx *c, uint64_t sizes[4]) {
    if (!c || !c->haveIndex) return fail(MB200_E_STATE, "no index");
    sizes[0] = c->nEntries; sizes[1] = c->nT; sizes[2] = c->totalRes; sizes[3] = 0; return 0;
}
int mb200_index_alloc(mb200_ctx *c, const uint64_t sizes[4]) {
    if (!c) return fail(MB200_E_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    c->nEntries = sizes[0]; c->nT = (uint32_t)sizes[1]; c->totalRes = sizes[2];
    return index_arena_alloc(c, c->nEntries, c->nT, c->totalRes);
}
int mb200_index_buffers(mb200_ctx *c, void *ptrs[8], uint64_t nbytes[8], int *n) {
    if (!c || !c->dOffsets.p) return fail(MB200_E_STATE, "index buffers not allocated");
    ptrs[0] = c->dOffsets.p; nbytes[0] = (mb200::KMER_TABLE + 1) * 4;
    ptrs[1] = c->dEntries.p; nbytes[1] = c->nEntries * 8;
    ptrs[2] = c->dTseqM.p; nbytes[2] = c->totalRes;
    ptrs[3] = c->dTseqU.p; nbytes[3] = c->totalRes;
    ptrs[4] = c->dToff.p; nbytes[4] = (uint64_t)(c->nT + 1) * 8;
    *n = 5; return 0;
}
int mb200_index_commit(mb200_ctx *c) {
    if (!c || !c->dToff.p) return fail(MB200_E_STATE, "index buffers not allocated");
    CU(cudaSetDevice(c->device));
    c->hToff.resize(c->nT + 1);
    CU(cudaMemcpy(c->hToff.data(), c->dToff.p, (size_t)(c->nT + 1) * 8, cudaMemcpyDeviceToHost));
    c->haveIndex = true; c->minScore.clear();
    return build_kmer_bits(c);
}
// One process, several contexts on several devices (bin/mb200 gpupredictexons --gpus N): the index of src -- offsets, entries, masked / unmasked residues, target offsets --
// is copied device to device (cudaMemcpyPeerAsync: NVLink when peer access is possible, staged through the host otherwise) into dst's arena, together with the target keys.
// The NCCL path (mb200_index_broadcast) is for one process per GPU.
int mb200_index_clone(mb200_ctx *dst, mb200_ctx *src, uint64_t *nbytes, double *ms) {
    if (!dst || !src || dst == src) return fail(MB200_E_ARG, "mb200_index_clone: two different contexts");
    if (!src->haveIndex) return fail(MB200_E_STATE, "source context has no index");
    uint64_t sizes[4]; TRY(mb200_index_sizes(src, sizes)); TRY(mb200_index_alloc(dst, sizes));
    void *ps[8], *pd[8]; uint64_t ns[8], nd[8]; int n1 = 0, n2 = 0;
    TRY(mb200_index_buffers(src, ps, ns, &n1)); TRY(mb200_index_buffers(dst, pd, nd, &n2));
    CU(cudaSetDevice(src->device)); CU(cudaStreamSynchronize(src->stream));
    CU(cudaSetDevice(dst->device));
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t total = 0;
    for (int i = 0; i < n1; i++) {
        if (ns[i] != nd[i]) return fail(MB200_E_INTERNAL, "index arenas differ");
        if (ns[i] && src->device != dst->device) CU(cudaMemcpyPeerAsync(pd[i], dst->device, ps[i], src->device, ns[i], dst->stream));
        else if (ns[i]) CU(cudaMemcpyAsync(pd[i], ps[i], ns[i], cudaMemcpyDeviceToDevice, dst->stream));      // two contexts on one device
        total += ns[i];
    }
    CU(cudaStreamSynchronize(dst->stream));
    if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (nbytes) *nbytes = total;
    dst->tKeys = src->tKeys;
    return mb200_index_commit(dst);
}


// ---- multi-GPU: replicate the index with ONE broadcast of the index arena --------------------------------------------
#ifndef MB200_SIMT_EMU
}  // extern "C"
static void mb200_comm_release(mb200_ctx *c) { if (c->comm && g_nccl.h) { g_nccl.CommDestroy((ncclComm_t)c->comm); } c->comm = nullptr; }
extern "C" {
int mb200_comm_unique_id(void *id128) {
    if (!id128) return fail(MB200_E_ARG, "null argument");
    TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NC(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId *>(id128)));
    return 0;
}
int mb200_comm_init(mb200_ctx *c, const void *id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(MB200_E_ARG, "bad argument");
    TRY(nccl_load());
    CU(cudaSetDevice(c->device));
    mb200_comm_release(c);
    ncclUniqueId id; memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    NC(g_nccl.CommInitRank(&comm, world, id, rank));
    c->comm = comm;
    return 0;
}
int mb200_comm_init_all(mb200_ctx **ctxs, int n) {   // one process, n contexts on n devices (bin/mb200 --gpus n)
    if (!ctxs || n < 1) return fail(MB200_E_ARG, "bad argument");
    TRY(nccl_load());
    std::vector<int> dev(n); std::vector<ncclComm_t> comms(n, nullptr);
    for (int i = 0; i < n; i++) { if (!ctxs[i]) return fail(MB200_E_ARG, "null ctx"); dev[i] = ctxs[i]->device; mb200_comm_release(ctxs[i]); }
    NC(g_nccl.CommInitAll(comms.data(), n, dev.data()));
    for (int i = 0; i < n; i++) ctxs[i]->comm = comms[i];
    return 0;
}
// Collective over the communicator (NULL = the one mb200_comm_init created): the root's index (k-mer offsets, index entries, masked and unmasked
// target residues, target offsets) is replicated on every rank.  Two device-side steps: a 32-byte header (sizes) so the other ranks can size
// their arena, then ONE ncclBroadcast of the arena itself over NVLink.  Communicator setup is not part of it.  ms (optional) = CUDA-event time
// of the arena broadcast on this rank's stream.
int mb200_index_broadcast(mb200_ctx *c, void *ncclComm, int root, uint64_t *nbytes, double *ms) {
    if (!c) return fail(MB200_E_ARG, "null ctx");
    TRY(nccl_load());
    ncclComm_t comm = ncclComm ? (ncclComm_t)ncclComm : (ncclComm_t)c->comm;
    if (!comm) return fail(MB200_E_STATE, "no communicator (mb200_comm_init, or pass an ncclComm_t)");
    CU(cudaSetDevice(c->device));
    int rank = -1, world = 0; NC(g_nccl.CommUserRank(comm, &rank)); NC(g_nccl.CommCount(comm, &world));
    if (root < 0 || root >= world) return fail(MB200_E_ARG, "root %d out of range", root);
    if (rank == root && !c->haveIndex) return fail(MB200_E_STATE, "the root has no index");
    uint64_t hdr[4] = {c->nEntries, c->nT, c->totalRes, (uint64_t)c->par.kmer_thr};
    TRY(c->dCells.alloc(4));
    if (rank == root) TRY(upload(c->dCells.p, hdr, sizeof(hdr), c->stream));
    NC(g_nccl.Broadcast(c->dCells.p, c->dCells.p, sizeof(hdr), ncclUint8, root, comm, c->stream));
    CU(cudaMemcpyAsync(hdr, c->dCells.p, sizeof(hdr), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (rank != root) {
        if ((int)hdr[3] != c->par.kmer_thr) return fail(MB200_E_STATE, "the root's index was built with k-mer threshold %d, this context uses %d", (int)hdr[3], c->par.kmer_thr);
        c->nEntries = hdr[0]; c->nT = (uint32_t)hdr[1]; c->totalRes = hdr[2];
        TRY(index_arena_alloc(c, c->nEntries, c->nT, c->totalRes));
    }
    CU(cudaEventRecord(c->ev0, c->stream));
    NC(g_nccl.Broadcast(c->dIndexArena.p, c->dIndexArena.p, c->arenaBytes, ncclUint8, root, comm, c->stream));
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float f = 0; CU(cudaEventElapsedTime(&f, c->ev0, c->ev1));
    if (ms) *ms = f;
    if (nbytes) *nbytes = c->arenaBytes;
    if (rank != root) return mb200_index_commit(c);
    return 0;
}
#else
}  // extern "C"
static void mb200_comm_release(mb200_ctx *) {}
extern "C" {
int mb200_comm_unique_id(void *) { return fail(MB200_E_STATE, "no NCCL in the SIMT emulation"); }
int mb200_comm_init(mb200_ctx *, const void *, int, int) { return fail(MB200_E_STATE, "no NCCL in the SIMT emulation"); }
int mb200_comm_init_all(mb200_ctx **, int) { return fail(MB200_E_STATE, "no NCCL in the SIMT emulation"); }
int mb200_index_broadcast(mb200_ctx *, void *, int, uint64_t *, double *) { return fail(MB200_E_STATE, "no NCCL in the SIMT emulation"); }
#endif

// ---- query upload + per-residue preparation ----------------------------------------------------------
// q: host residues, or NULL with qDev = residues already in device memory (fragments produced by mb200_extract_orfs)
static int queries_install(mb200_ctx *c, const uint8_t *q, const uint8_t *qDev, const uint64_t *qoff, uint32_t nQ) {
    if (!c || (!q && !qDev) || !qoff) return fail(MB200_E_ARG, "null argument");
    if (!c->haveMat) return fail(MB200_E_STATE, "matrices not set");
    CU(cudaSetDevice(c->device));
    const uint64_t tot = qoff[nQ];
    uint32_t maxL = 0;
    for (uint32_t i = 0; i < nQ; i++) { const uint64_t l = qoff[i + 1] - qoff[i]; if (l >= 32768) return fail(MB200_E_LIMIT, "query %u has >= 32768 residues", i); if (l > maxL) maxL = (uint32_t)l; }
    TRY(c->dQ.alloc(tot + 16)); TRY(c->dQoff.alloc(nQ + 1)); TRY(c->dBias.alloc(tot + 16)); TRY(c->dCorr.alloc(tot + 16)); TRY(c->dCb.alloc(tot + 16)); TRY(c->dOrder.alloc(nQ + 1));
    // longest queries first: the persistent prefilter CTAs pull work from this order
    std::vector<uint32_t> order(nQ);
    {   // stable counting sort by descending length (lengths < 32768)
        std::vector<uint32_t> start(maxL + 2, 0);
        for (uint32_t i = 0; i < nQ; i++) start[maxL - (uint32_t)(qoff[i + 1] - qoff[i]) + 1]++;
        for (uint32_t l = 0; l <= maxL; l++) start[l + 1] += start[l];
        for (uint32_t i = 0; i < nQ; i++) order[start[maxL - (uint32_t)(qoff[i + 1] - qoff[i])]++] = i;
    }
    CU(cudaEventRecord(c->ev0, c->stream));
    if (q) { TRY(upload(c->dQ.p, q, tot, c->stream)); } else if (tot) { CU(cudaMemcpyAsync(c->dQ.p, qDev, tot, cudaMemcpyDeviceToDevice, c->stream)); }
    TRY(upload(c->dQoff.p, qoff, (size_t)(nQ + 1) * 8, c->stream)); TRY(upload(c->dOrder.p, order.data(), (size_t)nQ * 4, c->stream));
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0; CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->st.ms_h2d = ms;
    c->hQoff.assign(qoff, qoff + nQ + 1); c->nQ = nQ; c->totalQ = tot; c->maxQLen = maxL; c->havePref = false; c->haveAln = false; c->hOrder = order;
    if (tot && !c->par.comp_bias_corr) {
        // --comp-bias-corr 0: the reference zeroes all three arrays (QueryMatcher.cpp:97-99, StripedSmithWaterman.cpp:1238-1240)
        CU(cudaMemsetAsync(c->dBias.p, 0, tot * sizeof(float), c->stream)); CU(cudaMemsetAsync(c->dCorr.p, 0, tot, c->stream)); CU(cudaMemsetAsync(c->dCb.p, 0, tot, c->stream));
        CU(cudaStreamSynchronize(c->stream)); c->st.ms_prep = 0;
    } else if (tot) {
        CU(cudaEventRecord(c->ev0, c->stream));
        MB_LAUNCH((k_prep), (unsigned)((tot + 255) / 256), 256, 0, c->stream, c->dQ.p, c->dQoff.p, nQ, tot, c->dSeedMat.p, c->dSeedPB.p, c->dAlnMat.p, c->dAlnPB.p, c->dBias.p, c->dCorr.p, c->dCb.p);
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->st.ms_prep = ms; c->st.launches++; stage_interval(c, c->ev0, c->ev1);
    }
    c->st.query_residues = tot; c->st.n_queries = nQ;
    return 0;
}

int mb200_queries_upload(mb200_ctx *c, const uint8_t *q, const uint64_t *qoff, uint32_t nQ) {
    if (!q) return fail(MB200_E_ARG, "null argument");
    return queries_install(c, q, nullptr, qoff, nQ);
}


// ---- ORF extraction + translation (the step before the path: extractorfs + translatenucs) ----------------
static void orf_build_tables(MbOrfTables &T, const char *code) {
    // complement (Orf::iupacReverseComplementTable, Orf.cpp:47-51) with "no complement" -> 'N' (Orf.cpp:146-150)
    const char *from = "ATUGCYRSWKMBDHVN", *to = "TAACGRYSWMKVHDBN";
    for (int c = 0; c < 256; c++) T.comp[c] = 'N';
    for (int i = 0; from[i]; i++) { T.comp[(unsigned char)from[i]] = (unsigned char)to[i]; T.comp[(unsigned char)tolower(from[i])] = (unsigned char)tolower(to[i]); }
    // IUPAC letter -> base set (TranslateNucl::initConversionTable, TranslateNucl.h:334-362); illegal characters -> gap
    const char *bases = "-ACMGRSVTWYHKDBN";
    for (int c = 0; c < 256; c++) T.base[c] = 0;
    for (int i = 0; i < 16; i++) { T.base[(unsigned char)bases[i]] = (unsigned char)i; T.base[(unsigned char)tolower(bases[i])] = (unsigned char)i; T.base[i] = (unsigned char)i; }
    T.base['U'] = T.base['u'] = 8; T.base['X'] = T.base['x'] = 15;
    // codon table of the genetic code over base sets (TranslateNucl::initTranslationTable, TranslateNucl.h:383-485)
    const int ex[4] = {1, 2, 4, 8}; const int tcag[9] = {0, 2, 1, 0, 3, 0, 0, 0, 0};
    for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) for (int k = 0; k < 16; k++) {
        char aa = 0; bool go = true;
        for (int p = 0; p < 4 && go; p++) if (ex[p] & i) for (int q = 0; q < 4 && go; q++) if (ex[q] & j) for (int r = 0; r < 4 && go; r++) if (ex[r] & k) {
            const char ch = code[16 * tcag[ex[p]] + 4 * tcag[ex[q]] + tcag[ex[r]]];
            if (!aa) aa = ch;
            else if (aa != ch) {
                if ((aa == 'B' || aa == 'D' || aa == 'N') && (ch == 'D' || ch == 'N')) aa = 'B';
                else if ((aa == 'Z' || aa == 'E' || aa == 'Q') && (ch == 'E' || ch == 'Q')) aa = 'Z';
                else if ((aa == 'J' || aa == 'I' || aa == 'L') && (ch == 'I' || ch == 'L')) aa = 'J';
                else { aa = 'X'; go = false; }
            }
        }
        T.aaOf[256 * i + 16 * j + k] = aa ? aa : 'X';
    }
    // residue letter -> prefilter alphabet (SubstitutionMatrix::setupLetterMapping, SubstitutionMatrix.cpp:262-304)
    const char *order = "ACDEFGHIKLMNPQRSTVWYX";
    for (int c = 0; c < 256; c++) T.aa2num[c] = 20;
    for (int i = 0; order[i]; i++) { T.aa2num[(unsigned char)order[i]] = (unsigned char)i; T.aa2num[(unsigned char)tolower(order[i])] = (unsigned char)i; }
    const char *alias = "JLZEBD";
    for (int i = 0; i < 6; i += 2) { T.aa2num[(unsigned char)alias[i]] = T.aa2num[(unsigned char)alias[i + 1]]; T.aa2num[(unsigned char)tolower(alias[i])] = T.aa2num[(unsigned char)alias[i + 1]]; }
}

int mb200_extract_orfs(mb200_ctx *c, const char *nuc, const uint64_t *off, uint32_t nContigs, int minLen, int maxLen, int startMode,
                       uint64_t *nOrfs, const mb200_orf **orfs, uint64_t *nAa, const char **aa) {
    if (!c || !nuc || !off || !nOrfs || !orfs || !nAa || !aa) return fail(MB200_E_ARG, "null argument");
    if (startMode != 1) return fail(MB200_E_ARG, "only --orf-start-mode 1 (any-to-stop, what metaeuk predictexons uses) is implemented");
    if (minLen < 4 || maxLen < minLen) return fail(MB200_E_ARG, "need 4 <= min_len <= max_len (codons)");
    if (nContigs >= (1u << 31)) return fail(MB200_E_LIMIT, "too many contigs");
    static_assert(sizeof(MbOrfRec) == sizeof(mb200_orf), "ABI struct");
    const auto t0 = std::chrono::steady_clock::now();
    CU(cudaSetDevice(c->device));
    const uint64_t total = off[nContigs] - off[0];
    for (uint32_t i = 0; i < nContigs; i++) if (off[i + 1] - off[i] >= (1ull << 31)) return fail(MB200_E_LIMIT, "contig %u has >= 2^31 nucleotides", i);
    const char *code = mb200::genetic_code(c->par.translation_table);
    if (!code) return fail(MB200_E_ARG, "unknown translation table %d", c->par.translation_table);
    const unsigned long long stopSet = orf_stop_set(code);
    if (!c->haveOrfTab) { MbOrfTables T; orf_build_tables(T, code); TRY(c->dOrfTab.alloc(1)); TRY(upload(c->dOrfTab.p, &T, sizeof(T), c->stream)); CU(cudaStreamSynchronize(c->stream)); c->haveOrfTab = true; }
    const uint32_t units = 2u * nContigs;
    TRY(c->dNuc.alloc(total + 16)); TRY(c->dNucOff.alloc(nContigs + 1)); TRY(c->dOrfCnt.alloc(units + 1)); TRY(c->dOrfAaCnt.alloc(units + 1));
    TRY(c->dOrfRecBase.alloc(units + 1)); TRY(c->dOrfAaBase.alloc(units + 1));
    std::vector<uint64_t> rel(nContigs + 1);
    for (uint32_t i = 0; i <= nContigs; i++) rel[i] = off[i] - off[0];
    TRY(upload(c->dNuc.p, nuc + off[0], total, c->stream)); TRY(upload(c->dNucOff.p, rel.data(), (size_t)(nContigs + 1) * 8, c->stream));
    CU(cudaEventRecord(c->ev0, c->stream));   // ms_orf = the two scan passes + translation (inputs resident), incl. the host scan between them
    c->nOrfs = 0; c->nOrfAa = 0; *nOrfs = 0; *nAa = 0; *orfs = nullptr; *aa = nullptr;
    if (nContigs == 0) return 0;
    const int grid = (int)std::min<uint32_t>(units, (uint32_t)c->nSM * 8u);
    MB_LAUNCH((k_orf_scan<false>), grid, MB_ORF_THREADS, 0, c->stream, c->dNuc.p, c->dNucOff.p, nContigs, c->dOrfTab.p, (uint32_t)minLen, (uint32_t)maxLen, stopSet,
                                                              c->dOrfCnt.p, c->dOrfAaCnt.p, nullptr, nullptr, nullptr);
    CU(cudaGetLastError());
    std::vector<uint32_t> cnt(units); std::vector<uint64_t> acnt(units), rb(units + 1, 0), ab(units + 1, 0);
    CU(cudaMemcpyAsync(cnt.data(), c->dOrfCnt.p, (size_t)units * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(acnt.data(), c->dOrfAaCnt.p, (size_t)units * 8, cudaMemcpyDeviceToHost, c->stream));
    TRY(sync_check(c, "k_orf_scan (count)"));
    for (uint32_t u = 0; u < units; u++) { rb[u + 1] = rb[u] + cnt[u]; ab[u + 1] = ab[u] + acnt[u]; }   // key order: contig, plus strand, minus strand
    const uint64_t n = rb[units], na = ab[units];
    TRY(c->dOrfRecs.alloc(n + 1)); TRY(c->dOrfAa.alloc(na + 16)); TRY(c->dOrfCodes.alloc(na + 16)); TRY(c->pOrfRecs.alloc(n + 1)); TRY(c->pOrfAa.alloc(na + 16));
    TRY(upload(c->dOrfRecBase.p, rb.data(), (size_t)(units + 1) * 8, c->stream)); TRY(upload(c->dOrfAaBase.p, ab.data(), (size_t)(units + 1) * 8, c->stream));
    if (n) {
        MB_LAUNCH((k_orf_scan<true>), grid, MB_ORF_THREADS, 0, c->stream, c->dNuc.p, c->dNucOff.p, nContigs, c->dOrfTab.p, (uint32_t)minLen, (uint32_t)maxLen, stopSet,
                                                                 nullptr, nullptr, c->dOrfRecBase.p, c->dOrfAaBase.p, c->dOrfRecs.p);
        CU(cudaGetLastError());
        MB_LAUNCH((k_orf_translate), c->nSM * 8, 256, 0, c->stream, c->dNuc.p, c->dNucOff.p, c->dOrfTab.p, c->dOrfRecs.p, n, c->par.reverse_fragments ? 1 : 0, c->dOrfAa.p, c->dOrfCodes.p);
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaMemcpyAsync(c->pOrfRecs.p, c->dOrfRecs.p, (size_t)n * sizeof(MbOrfRec), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(c->pOrfAa.p, c->dOrfAa.p, (size_t)na, cudaMemcpyDeviceToHost, c->stream));
    } else {
        CU(cudaEventRecord(c->ev1, c->stream));
    }
    TRY(sync_check(c, "k_orf_scan (emit) / k_orf_translate"));
    float ms = 0; CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->st.ms_orf = ms; stage_interval(c, c->ev0, c->ev1); c->st.orf_nucleotides = total; c->st.orf_fragments = n; c->st.orf_residues = na; c->st.launches += 3;
    c->st.ms_orf_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->nOrfs = n; c->nOrfAa = na;
    *nOrfs = n; *nAa = na; *orfs = reinterpret_cast<const mb200_orf *>(c->pOrfRecs.p); *aa = c->pOrfAa.p;
    return 0;
}

int mb200_queries_from_orfs(mb200_ctx *c) {
    if (!c) return fail(MB200_E_ARG, "null ctx");
    if (c->nOrfs >= (1ull << 32)) return fail(MB200_E_LIMIT, "more than 2^32 fragments in one batch");
    std::vector<uint64_t> qoff(c->nOrfs + 1);
    for (uint64_t i = 0; i < c->nOrfs; i++) qoff[i] = c->pOrfRecs.p[i].aaOff;
    qoff[c->nOrfs] = c->nOrfAa;
    return queries_install(c, nullptr, c->dOrfCodes.p, qoff.data(), (uint32_t)c->nOrfs);
}


// ---- reduceredundancy (the step after the path) ---------------------------------------------------------
int mb200_reduce_redundancy(mb200_ctx *c, const mb200_pred *preds, const uint64_t *predOff, uint32_t nContigs, const uint32_t *exonKeys, uint64_t nExonKeys,
                            mb200_pred_out *out, uint32_t *memberOrder, uint32_t *repOrder, uint32_t *repCount) {
    if (!c || !preds || !predOff || !exonKeys || !out || !memberOrder || !repOrder || !repCount) return fail(MB200_E_ARG, "null argument");
    static_assert(sizeof(MbPred) == sizeof(mb200_pred) && sizeof(MbPredOut) == sizeof(mb200_pred_out), "ABI structs");
    CU(cudaSetDevice(c->device));
    const uint64_t n = predOff[nContigs];
    if (nContigs == 0 || n == 0) { for (uint32_t i = 0; i < nContigs; i++) repCount[i] = 0; return 0; }
    for (uint64_t i = 0; i < n; i++) if ((uint64_t)preds[i].exon_begin + preds[i].n_exons > nExonKeys) return fail(MB200_E_ARG, "call %llu: exon range outside exon_keys", (unsigned long long)i);
    for (uint32_t i = 0; i < nContigs; i++) if (predOff[i + 1] - predOff[i] >= (1ull << 31)) return fail(MB200_E_LIMIT, "too many calls on contig %u", i);
    DevBuf<MbPred> dP; DevBuf<uint64_t> dOff; DevBuf<uint32_t> dKeys, dMem, dRep, dCnt, dFlag, dTmp; DevBuf<MbPredOut> dOut;
    int rc = 0;
    auto done = [&](int r) { dP.release(); dOff.release(); dKeys.release(); dMem.release(); dRep.release(); dCnt.release(); dFlag.release(); dTmp.release(); dOut.release(); return r; };
    if ((rc = dP.alloc(n)) || (rc = dOff.alloc(nContigs + 1)) || (rc = dKeys.alloc(nExonKeys + 1)) || (rc = dMem.alloc(n)) || (rc = dRep.alloc(n)) || (rc = dCnt.alloc(nContigs)) ||
        (rc = dFlag.alloc(n)) || (rc = dTmp.alloc(n)) || (rc = dOut.alloc(n))) return done(rc);
    if ((rc = upload(dP.p, preds, n * sizeof(MbPred), c->stream)) || (rc = upload(dOff.p, predOff, (size_t)(nContigs + 1) * 8, c->stream)) ||
        (rc = upload(dKeys.p, exonKeys, (size_t)nExonKeys * 4, c->stream))) return done(rc);
    cudaEventRecord(c->ev0, c->stream);
    MB_LAUNCH((k_reduce), (nContigs + 127) / 128, 128, 0, c->stream, dP.p, dOff.p, nContigs, dKeys.p, dOut.p, dMem.p, dRep.p, dCnt.p, dFlag.p, dTmp.p);
    cudaEventRecord(c->ev1, c->stream);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dOut.p, n * sizeof(MbPredOut), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(memberOrder, dMem.p, n * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(repOrder, dRep.p, n * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(repCount, dCnt.p, (size_t)nContigs * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return done(fail(MB200_E_CUDA, "k_reduce: %s", cudaGetErrorString(e)));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); c->st.ms_reduce = ms; c->st.launches++;
    return done(0);
}

int mb200_prefilter_resident(mb200_ctx *c) {
    if (!c) return fail(MB200_E_ARG, "null ctx");
    if (!c->haveIndex || !c->haveSim3) return fail(MB200_E_STATE, "index / matrices not set");
    CU(cudaSetDevice(c->device));
    TRY(ctx_alloc_scratch(c));
    TRY(build_buckets(c));
    const uint32_t nQ = c->nQ; const uint32_t maxHits = (uint32_t)c->par.max_res_list_len;
    TRY(c->dHits.alloc((size_t)nQ * maxHits)); TRY(c->dHitCount.alloc(nQ + 1)); TRY(c->dStatus.alloc(nQ + 1)); TRY(c->dCounter.alloc(4)); TRY(c->dStats.alloc(1));
    CU(cudaMemsetAsync(c->dCounter.p, 0, 16, c->stream)); CU(cudaMemsetAsync(c->dStats.p, 0, sizeof(MbPrefStats), c->stream));
    CU(cudaMemsetAsync(c->dHitCount.p, 0, (size_t)(nQ + 1) * 4, c->stream)); CU(cudaMemsetAsync(c->dStatus.p, 0xFF, (size_t)(nQ + 1) * 4, c->stream));
    MbPrefTables T; memset(&T, 0, sizeof(T));
    T.sim3 = c->dSim3.p; T.cnt3 = c->dCnt3.p; T.offsets = c->dOffsets.p; T.kmerBits = c->useBits ? c->dKmerBits.p : nullptr; T.entries = c->dEntries.p; T.buckets = (c->optBuckets && c->haveBuckets) ? c->dBuckets.p : nullptr; T.tseq = c->dTseqM.p; T.toff = c->dToff.p; T.nT = c->nT;
    T.kmerThr = c->par.kmer_thr; T.maxHits = c->par.max_res_list_len; T.minDiag = c->par.min_diag_score; T.binMask = (uint32_t)c->par.bin_size - 1;
    T.capP = c->capP; T.capI = c->capI; T.capK = c->capK; T.phaseStats = c->optPhaseStats; T.sortWalkMax = (uint32_t)c->optSortWalk; T.streamHint = (uint32_t)c->optStreamHint;
    T.capM = (uint32_t)std::min<uint64_t>(c->capM, 2ull * std::max<uint64_t>(1000000ull, c->nT) - 1);   // beyond that the reference switches to its own overflow path (QueryMatcher.cpp:43,280-305)
    if (nQ) {
        CU(cudaEventRecord(c->ev0, c->stream));
        // queries are taken longest first (dOrder); with a short-query configuration the order splits at prefilter_short_len
        uint32_t pSplit = nQ;
        if (c->prefCfgShort) { uint32_t lo = 0, hi = nQ; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; const uint32_t q = c->hOrder[mid];
                                                                              if ((uint32_t)(c->hQoff[q + 1] - c->hQoff[q]) <= (uint32_t)c->optPrefShortLen) hi = mid; else lo = mid + 1; } pSplit = lo; }
        if (pSplit) pref_launch(c, c->prefCfg, T, c->dScratch.p, c->nCta, c->dOrder.p, pSplit, c->dCounter.p);
        if (pSplit < nQ) {
            MbPrefTables TS = T; TS.capI = std::min(T.capI, 1u << 15); TS.capM = std::min(T.capM, 1u << 17);
            pref_launch(c, c->prefCfgShort, TS, c->dScratchShort.p, c->nCtaShort, c->dOrder.p + pSplit, nQ - pSplit, c->dCounter.p + 1); c->st.launches++;
        }
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        float ms = 0; CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->st.ms_prefilter = ms; c->st.launches++; stage_interval(c, c->ev0, c->ev1);
    }
    // status check: any query that could not be processed is an error, never a silent drop
    TRY(c->pStatus.alloc(nQ + 1));
    uint32_t *status = c->pStatus.p;
    CU(cudaMemcpyAsync(status, c->dStatus.p, (size_t)(nQ + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    // Queries whose work lists overflowed the default per-CTA capacities (long fragments at a high sensitivity, low-complexity queries against dense indexes)
    // take a second pass on a few CTAs with the largest capacities the kernel's word layouts allow: 2^22 (window, neighbour) items, 2^21 similar k-mers,
    // and as many index hits as the reference itself can hold before ITS overflow path starts (maxDbMatches = 2 max(10^6, #targets), QueryMatcher.cpp:43,
    // 280-305 -- that path changes the result and is not reproduced, so reaching it stays an error).  Capacities set explicitly by the caller are hard limits.
    if (nQ && !c->par.cap_items_log2 && !c->par.cap_kmers_log2 && !c->par.cap_hits_log2) {
        std::vector<uint32_t> again;
        for (uint32_t i = 0; i < nQ; i++) if (status[i] == MB_ST_OVERFLOW) again.push_back(i);
        if (!again.empty()) {
            const uint64_t refMax = 2ull * std::max<uint64_t>(1000000ull, c->nT);
            uint32_t allocM = 1u << 21; while (allocM < (1u << 24) && allocM < refMax) allocM <<= 1;          // the lists are sorted in place: power-of-two room
            const uint32_t capI2 = 1u << 22, capK2 = 1u << 21, capM2 = (uint32_t)std::min<uint64_t>(allocM, refMax - 1);
            const int nR = (int)std::min<size_t>(again.size(), 32);
            if (c->nCtaRescue < nR || c->rescueCapM != allocM) {
                TRY(pref_scratch_build(c, nR, 32768, capI2, allocM, c->dScratchRescueMem, c->dScratchRescue)); c->nCtaRescue = nR; c->rescueCapM = allocM;
            }
            TRY(c->dOrderRescue.alloc(again.size()));
            TRY(upload(c->dOrderRescue.p, again.data(), again.size() * 4, c->stream));
            CU(cudaMemsetAsync(c->dCounter.p, 0, 16, c->stream));
            MbPrefTables T2 = T; T2.capI = capI2; T2.capK = capK2; T2.capM = capM2;
            CU(cudaEventRecord(c->ev0, c->stream));
            pref_launch(c, c->prefCfg, T2, c->dScratchRescue.p, nR, c->dOrderRescue.p, (uint32_t)again.size(), c->dCounter.p);
            CU(cudaGetLastError());
            CU(cudaEventRecord(c->ev1, c->stream));
            CU(cudaMemcpyAsync(status, c->dStatus.p, (size_t)(nQ + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            float ms2 = 0; CU(cudaEventElapsedTime(&ms2, c->ev0, c->ev1)); c->st.ms_prefilter += ms2; c->st.launches++; stage_interval(c, c->ev0, c->ev1);
            c->rescuedQueries += again.size();
        }
    }
    const bool skipBad = c->optOverflowPolicy == 1 || (c->optOverflowPolicy < 0 && c->inPredict);
    uint32_t nBad = 0;
    for (uint32_t i = 0; i < nQ; i++) if (status[i] != MB_ST_OK) {
        if (skipBad) { nBad++; continue; }     // hit count 0 (written by the kernel), status readable through mb200_query_status_get, counted in the stats
        return status[i] == MB_ST_OVERFLOW
            ? fail(MB200_E_OVERFLOW, "query %u (%llu residues): more work than the prefilter's lists hold (explicit cap_*_log2, or the second pass' limits: 2^22 window/neighbour items, 2^21 similar k-mers, "
                   "%llu index hits = where the reference's own overflow path starts) -- lower the sensitivity or shorten the fragment", i,
                   (unsigned long long)(c->hQoff[i + 1] - c->hQoff[i]), (unsigned long long)(2ull * std::max<uint64_t>(1000000ull, c->nT) - 1))
            : fail(MB200_E_LIMIT, "query %u: prefilter status %u (sequence of 32768 or more residues)", i, status[i]);
    }
    c->st.overflow_queries = nBad;
    MbPrefStats hs; CU(cudaMemcpy(&hs, c->dStats.p, sizeof(hs), cudaMemcpyDeviceToHost));
    c->st.kmers = hs.kmers; c->st.lookups = hs.lookups; c->st.db_matches = hs.dbMatches; c->st.double_hits = hs.doubleHits; c->st.diag_cells = hs.diagCells; c->st.pref_results = hs.results;
    for (int i = 0; i < 16; i++) c->st.pref_phase_cycles[i] = hs.phase[i];
    c->havePref = true;
    return 0;
}

static int min_score_for(mb200_ctx *c, uint32_t L) {
    if (c->minScore.size() < 32769) c->minScore.assign(32769, -2);
    if (c->minScore[L] == -2) c->minScore[L] = c->ev.minScoreFor(c->par.eval_thr, (double)L, (double)c->totalRes);
    return c->minScore[L];
}


// ---- sw_kernel 6 (default): forward and reverse passes as streams of (query, two targets) pairs over persistent lane groups (sw_stream.cuh).
// Shared-memory geometry of one launch: every warp has a 32-descriptor queue, the descriptors of the pairs in flight, ring tags and P query profiles of
// 22 * stride + 16 bytes (stride = 32 * nb rows).
struct SwStreamGeom { uint32_t nb, stride, slots, P, warpBytes; int W, mode; size_t smem; };
static SwStreamGeom sw_stream_geom(uint32_t nb, int mode, double pairsPerQuery, int ringOpt, int ringFit = 0) {
    SwStreamGeom g; g.nb = nb; g.stride = 32u * nb; g.slots = 32u / nb; g.mode = mode;
    const uint32_t entry = 22u * g.stride + 16u, fixed = MB_SWS_QUEUE_BYTES + MB_SWS_INFO_BYTES + MB_SWS_TAG_BYTES;
    if (mode == 0) {
        // ring: the pairs of a query are neighbours in the list, so slots / (pairs per query) profiles are live at a time; a few spare ones keep
        // a slot from waiting for a free profile
        uint32_t P = ringOpt > 0 ? (uint32_t)ringOpt : (uint32_t)std::ceil((double)g.slots / std::max(1.0, pairsPerQuery)) + 3u;
        P = std::min(std::min(P, g.slots), 16u); if (P < 1) P = 1;
        // Long queries: a profile of 22 x 32 nb bytes is 5.6 KB at nb = 8 and 11 KB at nb = 16, and a ring sized for comfort leaves room for 4 warps per CTA only (8 warps per SM: one or two
        // per scheduler, ALU pipe half idle).  ringFit: no more profiles than still let 8 warps x 2 CTAs fit -- the slots of a warp then work on one or two queries at a time, which is
        // what they do anyway (a warp fetches as many pairs as it has slots, neighbours in the list, targets of a query sorted by length).
        if (ringFit && ringOpt == 0) { const uint32_t fit = (13824u - fixed) / entry; P = std::min(P, std::max(fit, 1u)); }
        g.P = P;
    } else g.P = g.slots;
    g.warpBytes = fixed + g.P * entry;
    int W = (int)((108u * 1024u) / g.warpBytes);                 // two CTAs per SM when they fit
    if (W < 2) W = (int)((216u * 1024u) / g.warpBytes);
    g.W = std::max(1, std::min(W, MB_SWS_MAXWARPS));
    g.smem = (size_t)g.W * g.warpBytes;
    return g;
}
static int sw_stream_launch(mb200_ctx *c, cudaStream_t st, const SwStreamGeom &g, const SwPair *desc, uint32_t nDesc, SwPassRec *passRec, uint32_t *nPass, uint32_t passCap,
                            uint32_t *counter, unsigned long long *cells, uint32_t *err, SwTraceOut *outs) {
    int perSM = 1;
    if (g.mode == 0) { CU(cudaFuncSetAttribute(k_sw_stream<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sw_stream<0>, g.W * 32, g.smem)); }
    else if (g.mode == 1) { CU(cudaFuncSetAttribute(k_sw_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sw_stream<1>, g.W * 32, g.smem)); }
    else { CU(cudaFuncSetAttribute(k_sw_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sw_stream<2>, g.W * 32, g.smem)); }
    if (perSM < 1) perSM = 1;
    const uint64_t perCta = (uint64_t)g.slots * g.W;               // pairs a CTA works on at a time
    uint64_t blocks = (uint64_t)c->nSM * perSM; const uint64_t need = (nDesc + perCta - 1) / perCta;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
#define MB_SWS_ARGS desc, nDesc, g.nb, g.stride, g.P, g.warpBytes, c->dQ.p, c->dQoff.p, c->dCb.p, c->dTseqU.p, c->dAlnMat.p, c->par.gap_open, c->par.gap_extend, passRec, nPass, passCap, counter, cells, err, outs
    if (g.mode == 0) MB_LAUNCH((k_sw_stream<0>), (unsigned)blocks, g.W * 32, g.smem, st, MB_SWS_ARGS);
    else if (g.mode == 1) MB_LAUNCH((k_sw_stream<1>), (unsigned)blocks, g.W * 32, g.smem, st, MB_SWS_ARGS);
    else MB_LAUNCH((k_sw_stream<2>), (unsigned)blocks, g.W * 32, g.smem, st, MB_SWS_ARGS);
#undef MB_SWS_ARGS
    CU(cudaGetLastError()); c->st.launches++;
    return 0;
}
static int align_stream(mb200_ctx *c, uint32_t nPairs, uint32_t pLong, uint32_t maxT, bool usePerm) {
    const uint32_t nQ = c->nQ, maxHits = (uint32_t)c->par.max_res_list_len;
    auto lenAt = [&](uint32_t p) { const uint32_t q = c->hOrder[p]; return (uint32_t)(c->hQoff[q + 1] - c->hQoff[q]); };
    auto firstLE = [&](uint32_t x) { uint32_t lo = 0, hi = nQ; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (lenAt(mid) <= x) hi = mid; else lo = mid + 1; } return lo; };
    // classes by number of 32-row blocks: class nb = positions [B[nb], B[nb - 1]) of the length-sorted order (lengths 32 (nb - 1) + 1 .. 32 nb); B[32] = pLong
    uint32_t B[33], hOff[33];
    for (uint32_t nb = 1; nb <= 32; nb++) B[nb] = firstLE(32u * nb);
    B[0] = nQ;
    for (uint32_t k = 0; k <= 32; k++) CU(cudaMemcpyAsync(&hOff[k], c->dLaneOff.p + B[k], 4, cudaMemcpyDeviceToHost, c->stream));
    TRY(sync_check(c, "scan"));
    const uint32_t nDescTotal = hOff[0] - hOff[32];
    uint32_t *dNPass = c->dCounter.p + 32, *dErr = c->dCounter.p + 33;
    TRY(c->dSwDesc.alloc(nDescTotal)); TRY(c->dPassRec.alloc(nPairs));
    CU(cudaEventRecord(c->ev0, c->stream));   // ms_sw_fwd = hit permutation + descriptors + forward kernels
    if (usePerm) {
        TRY(c->dPerm.alloc((size_t)nQ * maxHits));
        MB_LAUNCH((k_perm_by_len), c->nSM * 8, 128, 0, c->stream, c->dHits.p, c->dHitCount.p, c->dToff.p, nQ, maxHits, c->dPerm.p);
        CU(cudaGetLastError()); c->st.launches++;
    }
    // Launch groups: all block counts that give the same number of slots per warp (32 / nb) share one launch with the geometry of the largest -- a query
    // of 17..32 blocks has the warp to itself either way.  The groups are independent, so they run side by side on up to four streams (few long queries:
    // a launch of its own would leave most SMs idle behind the last long alignment).
    if (!c->swFork) {
        CU(cudaEventCreateWithFlags(&c->swFork, cudaEventDisableTiming));
        for (int k = 0; k < 4; k++) { CU(cudaStreamCreateWithFlags(&c->swSide[k], cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&c->swJoin[k], cudaEventDisableTiming)); }
    }
    static const uint32_t grpHi[10] = {32, 16, 10, 8, 6, 5, 4, 3, 2, 1}, grpLo[10] = {17, 11, 9, 7, 6, 5, 4, 3, 2, 1};
    for (int gI = 0; gI < 10; gI++) {
        const uint32_t pB = B[grpHi[gI]], pE = B[grpLo[gI] - 1], nDesc = hOff[grpLo[gI] - 1] - hOff[grpHi[gI]];
        if (pB >= pE || nDesc == 0) continue;
        MB_LAUNCH((k_sw_desc_fwd), (unsigned)std::min<uint64_t>((uint64_t)c->nSM * 8, (nDesc + 255) / 256), 256, 0, c->stream, c->dOrder.p, c->dLaneOff.p, pB, pE, c->dHits.p, c->dHitCount.p,
                  usePerm ? c->dPerm.p : nullptr, maxHits, c->dQoff.p, c->dToff.p, c->dMinScore.p, c->dSwDesc.p + (hOff[grpHi[gI]] - hOff[32]));
        CU(cudaGetLastError()); c->st.launches++;
    }
    CU(cudaEventRecord(c->swFork, c->stream));
    bool used[4] = {false, false, false, false}; int nLaunch = 0;
    for (int gI = 0; gI < 10; gI++) {
        const uint32_t pB = B[grpHi[gI]], pE = B[grpLo[gI] - 1], base = hOff[grpHi[gI]] - hOff[32], nDesc = hOff[grpLo[gI] - 1] - hOff[grpHi[gI]];
        if (pB >= pE || nDesc == 0) continue;
        const double ppq = (double)nDesc / (double)(pE - pB);
        const SwStreamGeom g = sw_stream_geom(grpHi[gI], (c->optSwRing >= 0 && ppq >= 2.5) ? 0 : 1, ppq, c->optSwRing, c->optSwRingFit);
        const int k = nLaunch++ & 3;
        if (!used[k]) { CU(cudaStreamWaitEvent(c->swSide[k], c->swFork, 0)); used[k] = true; }
        TRY(sw_stream_launch(c, c->swSide[k], g, c->dSwDesc.p + base, nDesc, c->dPassRec.p, dNPass, nPairs, c->dCounter.p + 64 + gI, c->dCells.p, dErr, nullptr));
    }
    for (int k = 0; k < 4; k++) if (used[k]) { CU(cudaEventRecord(c->swJoin[k], c->swSide[k])); CU(cudaStreamWaitEvent(c->stream, c->swJoin[k], 0)); }
    CU(cudaEventRecord(c->ev1, c->stream));
    uint32_t hn[2] = {0, 0}; unsigned long long hc[2] = {0, 0};
    CU(cudaMemcpyAsync(hn, dNPass, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hc, c->dCells.p, 16, cudaMemcpyDeviceToHost, c->stream));
    TRY(sync_check(c, "forward alignment kernel"));
    float ms1 = 0; CU(cudaEventElapsedTime(&ms1, c->ev0, c->ev1)); c->st.ms_sw_fwd = ms1; c->st.sw_cells_fwd = hc[0]; stage_interval(c, c->ev0, c->ev1);
    if (hn[1] == 2) return fail(MB200_E_INTERNAL, "pass list overflow");
    const uint32_t nPass = hn[0];
    // queries longer than 1024 residues: every hit goes straight to the 32-bit kernel (forward + reverse; the host applies the e-value test)
    std::vector<uint32_t> extra;
    if (pLong > 0) {
        std::vector<uint32_t> hcnt(nQ);
        CU(cudaMemcpy(hcnt.data(), c->dHitCount.p, (size_t)nQ * 4, cudaMemcpyDeviceToHost));
        for (uint32_t p = 0; p < pLong; p++) { const uint32_t q = c->hOrder[p]; for (uint32_t k = 0; k < hcnt[q]; k++) extra.push_back(q * maxHits + k); }
    }
    if (!nPass && extra.empty()) return 0;
    CU(cudaEventRecord(c->ev0, c->stream));
    uint32_t hcls[MB_SWS_REV_CLASSES + 1]; memset(hcls, 0, sizeof(hcls));
    uint32_t *dCls = c->dCounter.p + 100, *dNFb = c->dCounter.p + 100 + MB_SWS_REV_CLASSES;
    TRY(c->dFallback.alloc((size_t)nPass + extra.size()));
    if (nPass) {
        // reverse pass: the forward kernels recorded the end cell; descriptors binned by the number of row blocks of the query prefix [0..qEnd]
        TRY(c->dTrace.alloc(nPass)); TRY(c->dSwDescRev.alloc((size_t)MB_SWS_REV_CLASSES * nPass));
        MB_LAUNCH((k_sw_desc_rev), (unsigned)std::min<uint64_t>((uint64_t)c->nSM * 8, ((uint64_t)nPass + 255) / 256), 256, 0, c->stream, c->dPassRec.p, nPass, c->dHits.p, maxHits, c->dToff.p,
                  c->dSwDescRev.p, nPass, dCls, c->dFallback.p, dNFb, c->dTrace.p);
        CU(cudaGetLastError()); c->st.launches++;
        CU(cudaMemcpyAsync(hcls, dCls, sizeof(hcls), cudaMemcpyDeviceToHost, c->stream));
        TRY(sync_check(c, "reverse descriptors"));
        CU(cudaEventRecord(c->swFork, c->stream));
        bool usedR[4] = {false, false, false, false};
        static const uint32_t revNb[MB_SWS_REV_CLASSES] = {1, 2, 3, 4, 5, 6, 8, 10, 16, 32};
        for (int cl = 0; cl < MB_SWS_REV_CLASSES; cl++) {
            if (!hcls[cl]) continue;
            const SwStreamGeom g = sw_stream_geom(revNb[cl], 2, 1.0, 0);
            const int k = cl & 3;
            if (!usedR[k]) { CU(cudaStreamWaitEvent(c->swSide[k], c->swFork, 0)); usedR[k] = true; }
            TRY(sw_stream_launch(c, c->swSide[k], g, c->dSwDescRev.p + (size_t)cl * nPass, hcls[cl], nullptr, nullptr, 0, c->dCounter.p + 112 + cl, c->dCells.p + 1, dErr, c->dTrace.p));
        }
        for (int k = 0; k < 4; k++) if (usedR[k]) { CU(cudaEventRecord(c->swJoin[k], c->swSide[k])); CU(cudaStreamWaitEvent(c->stream, c->swJoin[k], 0)); }
    }
    uint32_t nFb = hcls[MB_SWS_REV_CLASSES];
    if (!extra.empty()) { TRY(upload(c->dFallback.p + nFb, extra.data(), extra.size() * 4, c->stream)); nFb += (uint32_t)extra.size(); }
    if (nFb) {
        // 32-bit cells, forward + reverse: query prefixes of more than 256 rows, scores beyond the packed 16-bit range, queries of more than 1024 residues
        const int blocks = c->nSM * 4; const size_t nWarps = (size_t)blocks * 4;
        TRY(c->dTrace2.alloc(nFb)); TRY(c->dHE.alloc(nWarps * maxT * 32));
        MB_LAUNCH((k_sw_trace2), blocks, 128, 0, c->stream, c->dFallback.p, nFb, c->dHits.p, maxHits, c->dQ.p, c->dQoff.p, c->dCb.p, c->dTseqU.p, c->dToff.p, c->dAlnMat.p,
                  c->par.gap_open, c->par.gap_extend, c->dHE.p, maxT, c->dTrace2.p, c->dCounter.p + 40, c->dCells.p + 1);
        CU(cudaGetLastError()); c->st.launches++;
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    c->hTrace.resize((size_t)nPass + nFb);
    if (nPass) CU(cudaMemcpyAsync(c->hTrace.data(), c->dTrace.p, (size_t)nPass * sizeof(SwTraceOut), cudaMemcpyDeviceToHost, c->stream));
    if (nFb) CU(cudaMemcpyAsync(c->hTrace.data() + nPass, c->dTrace2.p, (size_t)nFb * sizeof(SwTraceOut), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hc, c->dCells.p, 16, cudaMemcpyDeviceToHost, c->stream));
    TRY(sync_check(c, "reverse alignment kernels"));
    CU(cudaEventElapsedTime(&ms1, c->ev0, c->ev1)); c->st.ms_sw_rev = ms1; c->st.sw_cells_rev = hc[1]; stage_interval(c, c->ev0, c->ev1);
    size_t w = 0;
    for (size_t r = 0; r < c->hTrace.size(); r++) {
        const SwTraceOut &o = c->hTrace[r];
        if (o.slot == 0xFFFFFFFFu) continue;     // handed to the 32-bit kernel (its own record follows)
        if (o.flag & 1) return fail(MB200_E_INTERNAL, "forward/backward SW scores differ (query %u, hit %u)", o.slot / maxHits, o.slot % maxHits);
        c->hTrace[w++] = o;
    }
    c->hTrace.resize(w);
    __gnu_parallel::sort(c->hTrace.begin(), c->hTrace.end(), [](const SwTraceOut &a, const SwTraceOut &b) { return a.slot < b.slot; });
    return 0;
}


// Gapped alignment of the resident hit lists (dHits / dHitCount).  Leaves c->hTrace = every alignment whose score passes the
// e-value gate (score, start/end coordinates), sorted by slot = query * maxHits + rank.
static int align_pipeline(mb200_ctx *c) {
    const uint32_t nQ = c->nQ, maxHits = (uint32_t)c->par.max_res_list_len;
    c->hTrace.clear(); c->haveAln = false;
    c->st.sw_pairs = 0; c->st.sw_cells_fwd = c->st.sw_cells_rev = 0; c->st.ms_sw_fwd = c->st.ms_sw_rev = 0; c->st.sw_target_bytes = 0;
    if (nQ == 0) { c->haveAln = true; return 0; }
    if ((uint64_t)nQ * maxHits >= (1ull << 32)) return fail(MB200_E_LIMIT, "n_queries * max_res_list_len must be < 2^32 per call");
    TRY(c->dPairOff.alloc(nQ + 1)); TRY(c->dChunkOff.alloc(nQ + 1)); TRY(c->dLaneOff.alloc(nQ + 1)); TRY(c->dWaveOff.alloc(nQ + 1)); TRY(c->dCounter.alloc(128)); TRY(c->dCells.alloc(2));
    CU(cudaMemsetAsync(c->dCounter.p, 0, 128 * 4, c->stream)); CU(cudaMemsetAsync(c->dCells.p, 0, 16, c->stream));
    // sw_kernel 4: queries of at most 64 residues with at most sw_lane_max_hits hits go to k_sw_lanes (one (query, 2 hits) item per lane), the rest to k_sw_rows2
    // sw_kernel 5: the other queries of at most 1024 residues go to k_sw_wave (the row blocks of one alignment on neighbouring lanes)
    // sw_kernel 6 (default): every query of at most 1024 residues goes to the stream kernels (sw_stream.cuh); needs 32-bit target offsets and a gap-open penalty within the cell bias
    const bool streamMode = c->optSwKernel == 6 && c->totalRes < (1ull << 32) && c->par.gap_open >= 0 && c->par.gap_open <= MB_SWS_BIAS && c->par.gap_extend >= 0 && c->par.gap_extend <= MB_SWS_BIAS;
    const bool laneMode = !streamMode && c->optSwKernel >= 4 && maxHits <= 1024, waveMode = !streamMode && c->optSwKernel >= 5 && maxHits <= 1024;
    if (streamMode)   // laneOff = scan of ceil(hits / 2) over the queries of at most 1024 residues, in length-sorted order
        MB_LAUNCH((k_scan_counts), 4, MB_CTA_THREADS, 0, c->stream, c->dHitCount.p, c->dOrder.p, nQ, c->dPairOff.p, c->dChunkOff.p, c->dLaneOff.p, c->dWaveOff.p, c->dQoff.p, 1024u, 0xFFFFFFFFu, 0u);
    else
    MB_LAUNCH((k_scan_counts), 4, MB_CTA_THREADS, 0, c->stream, c->dHitCount.p, c->dOrder.p, nQ, c->dPairOff.p, c->dChunkOff.p, c->dLaneOff.p, c->dWaveOff.p, c->dQoff.p, 64u,
                                                       laneMode ? (uint32_t)c->optLaneMaxHits : 0u, waveMode ? 1024u : 0u);
    CU(cudaGetLastError());
    uint32_t nPairs = 0;
    CU(cudaMemcpyAsync(&nPairs, c->dPairOff.p + nQ, 4, cudaMemcpyDeviceToHost, c->stream));
    TRY(sync_check(c, "scan"));
    c->st.launches++; c->st.sw_pairs = nPairs;
    if (nPairs == 0) { c->haveAln = true; return 0; }
    // e-value gate as a per-length minimum raw score (StripedSmithWaterman.cpp:389-398)
    std::vector<int32_t> ms(c->maxQLen + 2, 0x7fffffff);
    for (uint32_t i = 0; i < nQ; i++) {
        const uint32_t L = (uint32_t)(c->hQoff[i + 1] - c->hQoff[i]);
        if (ms[L] == 0x7fffffff) { const int v = min_score_for(c, L); if (v < 0) return fail(MB200_E_INTERNAL, "e-value not monotone in the score for query length %u", L); ms[L] = v; }
    }
    TRY(c->dMinScore.alloc(ms.size())); TRY(upload(c->dMinScore.p, ms.data(), ms.size() * 4, c->stream));
    const bool usePerm = maxHits <= 1024;
    if (streamMode) {
        auto lenAt0 = [&](uint32_t p) { const uint32_t q = c->hOrder[p]; return (uint32_t)(c->hQoff[q + 1] - c->hQoff[q]); };
        uint32_t lo = 0, hi = nQ; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (lenAt0(mid) <= 1024u) hi = mid; else lo = mid + 1; }
        uint32_t maxT0 = 1; for (uint32_t t = 0; t < c->nT; t++) maxT0 = std::max<uint32_t>(maxT0, (uint32_t)(c->hToff[t + 1] - c->hToff[t]));
        TRY(align_stream(c, nPairs, lo, maxT0, usePerm));
        c->haveAln = true;
        return 0;
    }
    TRY(c->dScore.alloc((size_t)nQ * maxHits)); TRY(c->dPassList.alloc(nPairs));
    CU(cudaEventRecord(c->ev0, c->stream));   // ms_sw_fwd = hit permutation + forward kernels
    if (usePerm) {
        TRY(c->dPerm.alloc((size_t)nQ * maxHits));
        MB_LAUNCH((k_perm_by_len), c->nSM * 8, 128, 0, c->stream, c->dHits.p, c->dHitCount.p, c->dToff.p, nQ, maxHits, c->dPerm.p);
        CU(cudaGetLastError()); c->st.launches++;
    }
    uint32_t *dNPass = c->dCounter.p + 32, *dErr = c->dCounter.p + 33;
    // two launches over the length-sorted query order (descending): queries <= 256 residues and 257..1024 (profile size differs)
    auto lenAt = [&](uint32_t p) { const uint32_t q = c->hOrder[p]; return (uint32_t)(c->hQoff[q + 1] - c->hQoff[q]); };
    auto firstLE = [&](uint32_t x) { uint32_t lo = 0, hi = nQ; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (lenAt(mid) <= x) hi = mid; else lo = mid + 1; } return lo; };
    const uint32_t pLong = firstLE(1024);
    uint32_t maxT = 1; for (uint32_t t = 0; t < c->nT; t++) maxT = std::max<uint32_t>(maxT, (uint32_t)(c->hToff[t + 1] - c->hToff[t]));
    const uint32_t bounds[3] = {pLong, firstLE(256), nQ}; const uint32_t strides[2] = {1024 + 16, 256 + 16};
    const bool packed = c->optSwKernel >= 2;   // (4 = 2 for the queries that stay with the chunk kernel)   // 1 = one target per lane (32-bit cells), 2 = two targets per lane (packed 16-bit cells)
    for (int k = 0; k < 2; k++) {
        const uint32_t pB = bounds[k], pE = bounds[k + 1];
        if (pB >= pE) continue;
        if (packed) {
            int W = (int)((100u * 1024u) / (2u * 22u * strides[k])); if (W > MB_SWR_WARPS) W = MB_SWR_WARPS; if (W < 1) W = 1;
            const size_t smem = (size_t)W * 2u * 22u * strides[k];
            CU(cudaFuncSetAttribute(k_sw_rows2<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            int perSM = 1; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sw_rows2<32>, W * 32, smem)); if (perSM < 1) perSM = 1;
            const int blocks = c->nSM * perSM;
            TRY(c->dBnd.alloc((size_t)blocks * W * maxT * 32u * 2u));
            MB_LAUNCH((k_sw_rows2<32>), blocks, W * 32, smem, c->stream, c->dOrder.p, c->dChunkOff.p, pB, pE, c->dHits.p, c->dHitCount.p, usePerm ? c->dPerm.p : nullptr, maxHits,
                                                             c->dQ.p, c->dQoff.p, c->dCb.p, c->dTseqU.p, c->dToff.p, c->dAlnMat.p, c->par.gap_open, c->par.gap_extend,
                                                             strides[k], c->dMinScore.p, c->dScore.p, c->dPassList.p, dNPass, nPairs, c->dCounter.p + k, c->dCells.p, dErr,
                                                             reinterpret_cast<uint2 *>(c->dBnd.p), maxT);
        } else {
            const size_t smem = (size_t)MB_SWR_WARPS * 21u * strides[k];
            CU(cudaFuncSetAttribute(k_sw_rows<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            int perSM = 1; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sw_rows<32>, MB_SWR_WARPS * 32, smem)); if (perSM < 1) perSM = 1;
            const int blocks = c->nSM * perSM;
            TRY(c->dBnd.alloc((size_t)blocks * MB_SWR_WARPS * maxT * 32u));
            MB_LAUNCH((k_sw_rows<32>), blocks, MB_SWR_WARPS * 32, smem, c->stream, c->dOrder.p, c->dChunkOff.p, pB, pE, c->dHits.p, c->dHitCount.p, usePerm ? c->dPerm.p : nullptr, maxHits,
                                                                         c->dQ.p, c->dQoff.p, c->dCb.p, c->dTseqU.p, c->dToff.p, c->dAlnMat.p, c->par.gap_open, c->par.gap_extend,
                                                                         strides[k], c->dMinScore.p, c->dScore.p, c->dPassList.p, dNPass, nPairs, c->dCounter.p + k, c->dCells.p, dErr,
                                                                         c->dBnd.p, maxT);
        }
        CU(cudaGetLastError()); c->st.launches++;
    }
    if (waveMode) {
        // by profile stride: positions [pLong, firstLE(512)) hold the queries of 513..1024 residues (one item per warp), [firstLE(512), firstLE(256)) those of
        // 257..512, [firstLE(256), nQ) those of <= 256 (one item per half-warp)
        const uint32_t wbnd[4] = {pLong, firstLE(512), firstLE(256), nQ}; const uint32_t wstr[3] = {1024 + 16, 512 + 16, 256 + 16}, wG[3] = {32, 16, 16};
        for (int k = 0; k < 3; k++) {
            if (wbnd[k] >= wbnd[k + 1]) continue;
            const uint32_t perWarpBytes = (32u / wG[k]) * 22u * wstr[k];
            int W = (int)((100u * 1024u) / perWarpBytes); if (W > MB_SWR_WARPS) W = MB_SWR_WARPS; if (W < 1) W = 1;
            const size_t smem = (size_t)W * perWarpBytes;
            CU(cudaFuncSetAttribute(k_sw_wave, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            int perSM = 1; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sw_wave, W * 32, smem)); if (perSM < 1) perSM = 1;
            MB_LAUNCH((k_sw_wave), c->nSM * perSM, W * 32, smem, c->stream, c->dOrder.p, c->dWaveOff.p, wbnd[k], wbnd[k + 1], c->dHits.p, c->dHitCount.p, usePerm ? c->dPerm.p : nullptr, maxHits,
                                                                 c->dQ.p, c->dQoff.p, c->dCb.p, c->dTseqU.p, c->dToff.p, c->dAlnMat.p, c->par.gap_open, c->par.gap_extend,
                                                                 wstr[k], c->dMinScore.p, c->dScore.p, c->dPassList.p, dNPass, nPairs, c->dCounter.p + 4 + k, c->dCells.p, dErr, wG[k]);
            CU(cudaGetLastError()); c->st.launches++;
        }
    }
    if (laneMode) {
        // positions [firstLE(64), firstLE(32)) hold the queries of 33..64 residues (two row blocks), [firstLE(32), nQ) those of <= 32 (one block)
        const uint32_t lb[3] = {firstLE(64), firstLE(32), nQ};
        for (int k = 0; k < 2; k++) {
            if (lb[k] >= lb[k + 1]) continue;
            const int threads = k == 0 ? 128 : 256; const size_t smem = (size_t)threads * (22u * (k == 0 ? 64u : 32u) + 16u);
            if (k == 0) CU(cudaFuncSetAttribute(k_sw_lanes<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else CU(cudaFuncSetAttribute(k_sw_lanes<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int perSM = 1;
            if (k == 0) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sw_lanes<64>, threads, smem));
            else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sw_lanes<32>, threads, smem));
            if (perSM < 1) perSM = 1;
            const int blocks = c->nSM * perSM;
            TRY(c->dBnd.alloc((size_t)blocks * (threads / 32) * maxT * 32u * 2u));
            if (k == 0)
                MB_LAUNCH((k_sw_lanes<64>), blocks, threads, smem, c->stream, c->dOrder.p, c->dLaneOff.p, lb[k], lb[k + 1], c->dHits.p, c->dHitCount.p, usePerm ? c->dPerm.p : nullptr, maxHits,
                                                                   c->dQ.p, c->dQoff.p, c->dCb.p, c->dTseqU.p, c->dToff.p, c->dAlnMat.p, c->par.gap_open, c->par.gap_extend,
                                                                   c->dMinScore.p, c->dScore.p, c->dPassList.p, dNPass, nPairs, c->dCounter.p + 2 + k, c->dCells.p, dErr,
                                                                   reinterpret_cast<uint2 *>(c->dBnd.p), maxT);
            else
                MB_LAUNCH((k_sw_lanes<32>), blocks, threads, smem, c->stream, c->dOrder.p, c->dLaneOff.p, lb[k], lb[k + 1], c->dHits.p, c->dHitCount.p, usePerm ? c->dPerm.p : nullptr, maxHits,
                                                                   c->dQ.p, c->dQoff.p, c->dCb.p, c->dTseqU.p, c->dToff.p, c->dAlnMat.p, c->par.gap_open, c->par.gap_extend,
                                                                   c->dMinScore.p, c->dScore.p, c->dPassList.p, dNPass, nPairs, c->dCounter.p + 2 + k, c->dCells.p, dErr,
                                                                   reinterpret_cast<uint2 *>(c->dBnd.p), maxT);
            CU(cudaGetLastError()); c->st.launches++;
        }
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    uint32_t hn[2] = {0, 0};
    CU(cudaMemcpyAsync(hn, dNPass, 8, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long hc[2] = {0, 0};
    CU(cudaMemcpyAsync(hc, c->dCells.p, 16, cudaMemcpyDeviceToHost, c->stream));
    TRY(sync_check(c, "forward alignment kernel"));
    float ms1 = 0; CU(cudaEventElapsedTime(&ms1, c->ev0, c->ev1)); c->st.ms_sw_fwd = ms1; c->st.sw_cells_fwd = hc[0]; stage_interval(c, c->ev0, c->ev1);
    if (hn[1] == 1) return fail(MB200_E_LIMIT, "alignment score overflows 16 bits (scores must stay below 32000)");
    if (hn[1] == 2) return fail(MB200_E_INTERNAL, "pass list overflow");
    // queries longer than 1024 residues: every hit goes straight to the trace kernel (host applies the e-value test)
    std::vector<uint32_t> extra;
    if (pLong > 0) {
        std::vector<uint32_t> hcnt(nQ);
        CU(cudaMemcpy(hcnt.data(), c->dHitCount.p, (size_t)nQ * 4, cudaMemcpyDeviceToHost));
        for (uint32_t p = 0; p < pLong; p++) { const uint32_t q = c->hOrder[p]; for (uint32_t k = 0; k < hcnt[q]; k++) extra.push_back(q * maxHits + k); }
    }
    uint32_t nPass = hn[0];
    if (!extra.empty()) {
        if ((uint64_t)nPass + extra.size() > nPairs) return fail(MB200_E_INTERNAL, "pass list overflow");
        TRY(upload(c->dPassList.p + nPass, extra.data(), extra.size() * 4, c->stream)); nPass += (uint32_t)extra.size();
    }
    if (nPass) {
        const int blocks = c->nSM * 4; const size_t nWarps = (size_t)blocks * 4;
        CU(cudaEventRecord(c->ev0, c->stream));
        TRY(c->dTrace.alloc(nPass));
        TRY(c->dHE.alloc(nWarps * maxT * 32));
        MB_LAUNCH((k_sw_trace2), blocks, 128, 0, c->stream, c->dPassList.p, nPass, c->dHits.p, maxHits, c->dQ.p, c->dQoff.p, c->dCb.p, c->dTseqU.p, c->dToff.p, c->dAlnMat.p,
                                                   c->par.gap_open, c->par.gap_extend, c->dHE.p, maxT, c->dTrace.p, c->dCounter.p + 40, c->dCells.p + 1);
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->ev1, c->stream));
        c->hTrace.resize(nPass);
        CU(cudaMemcpyAsync(c->hTrace.data(), c->dTrace.p, (size_t)nPass * sizeof(SwTraceOut), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(hc, c->dCells.p, 16, cudaMemcpyDeviceToHost, c->stream));
        TRY(sync_check(c, "k_sw_trace"));
        CU(cudaEventElapsedTime(&ms1, c->ev0, c->ev1)); c->st.ms_sw_rev = ms1; c->st.sw_cells_rev = hc[1]; stage_interval(c, c->ev0, c->ev1); c->st.launches++;
        for (const SwTraceOut &o : c->hTrace) {
            if (o.flag & 1) return fail(MB200_E_INTERNAL, "forward/backward SW scores differ (slot %u)", o.slot);
            if (o.flag & 2) return fail(MB200_E_LIMIT, "alignment score overflows 16 bits");
        }
        __gnu_parallel::sort(c->hTrace.begin(), c->hTrace.end(), [](const SwTraceOut &a, const SwTraceOut &b) { return a.slot < b.slot; });
    }
    c->haveAln = true;
    return 0;
}

// host epilogue for one traced alignment: e-value, bit score, coverage, estimated identity, acceptance
// (Matcher::getSWResult, Matcher.cpp:100-142; Alignment::checkCriteria, Alignment.cpp:548-567)
static void finish_aln(const mb200_ctx *c, const SwTraceOut &o, uint32_t tid, uint32_t L, mb200_aln &a) {
    memset(&a, 0, sizeof(a));
    a.seq_id = tid; a.raw_score = o.score; a.q_end = o.qEnd; a.db_end = o.dbEnd; a.q_start = o.qStart; a.db_start = o.dbStart;
    if (o.dbEnd < 0) { a.q_end = (int)L - 1; return; }
    a.evalue = c->ev.evalue(a.raw_score, L, (double)c->totalRes);
    a.bit_score = c->ev.bitScore(a.raw_score);
    if (a.evalue > c->par.eval_thr) { a.q_start = a.db_start = -1; return; }   // the reference returns before the reverse pass
    const uint32_t tLen = (uint32_t)(c->hToff[tid + 1] - c->hToff[tid]);
    a.q_cov = (std::min(L, (uint32_t)std::max(a.q_start, a.q_end)) - (uint32_t)std::min(a.q_start, a.q_end) + 1) / (float)L;
    a.db_cov = (std::min(tLen, (uint32_t)std::max(a.db_start, a.db_end)) - (uint32_t)std::min(a.db_start, a.db_end) + 1) / (float)tLen;
    a.aln_len = std::max(abs(a.q_end - a.q_start), abs(a.db_end - a.db_start)) + 1;
    const unsigned qa = std::max((unsigned)(a.q_end - a.q_start), 1u), ta = std::max((unsigned)(a.db_end - a.db_start), 1u);
    float sid = ((uint16_t)a.raw_score / static_cast<float>(std::max(qa, ta))) * 0.1656 + 0.1141;
    sid = std::min(sid, 1.0f); a.seq_id_est = std::max(0.0f, sid);
    a.passed = (a.evalue <= c->par.eval_thr && a.aln_len >= c->par.min_aln_len) ? 1 : 0;
}

static int trace_tids(mb200_ctx *c, std::vector<uint32_t> &tids);

int mb200_align_resident(mb200_ctx *c) {
    if (!c || !c->havePref) return fail(MB200_E_STATE, "no resident prefilter result");
    CU(cudaSetDevice(c->device));
    return align_pipeline(c);
}

// compact device->host copy of the resident prefilter result: hit_off[n_q+1], hits in query order
static int download_hits_compact(mb200_ctx *c) {
    const uint32_t nQ = c->nQ, maxHits = (uint32_t)c->par.max_res_list_len;
    c->hHitOff.assign(nQ + 1, 0); c->nFlatHits = 0;
    if (nQ == 0) return 0;
    TRY(c->dPairOff.alloc(nQ + 1)); TRY(c->dChunkOff.alloc(nQ + 1)); TRY(c->pOff.alloc(nQ + 1));
    MB_LAUNCH((k_scan_counts), 2, MB_CTA_THREADS, 0, c->stream, c->dHitCount.p, c->dOrder.p, nQ, c->dPairOff.p, c->dChunkOff.p, nullptr, nullptr, c->dQoff.p, 0u, 0u, 0u);
    CU(cudaGetLastError());
    uint32_t *off = c->pOff.p;
    CU(cudaEventRecord(c->ev0, c->stream));
    CU(cudaMemcpyAsync(off, c->dPairOff.p, (size_t)(nQ + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    TRY(sync_check(c, "scan"));
    const uint32_t n = off[nQ];
    TRY(c->dFlatHits.alloc(n)); TRY(c->pFlatHits.alloc(n));
    MB_LAUNCH((k_compact_hits), nQ, 64, 0, c->stream, c->dHits.p, c->dHitCount.p, c->dPairOff.p, nQ, maxHits, c->dFlatHits.p);
    CU(cudaGetLastError());
    if (n) CU(cudaMemcpyAsync(c->pFlatHits.p, c->dFlatHits.p, (size_t)n * sizeof(MbPrefHit), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev1, c->stream));
    for (uint32_t i = 0; i <= nQ; i++) c->hHitOff[i] = off[i];
    TRY(sync_check(c, "compact hits"));
    float ms = 0; CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->st.ms_d2h = ms; c->st.launches += 2;
    c->nFlatHits = n;
    return 0;
}

int mb200_results_download(mb200_ctx *c, mb200_hit *hits, uint32_t *hitCount, mb200_aln *alns, uint32_t *alnCount) {
    if (!c || !c->havePref) return fail(MB200_E_STATE, "no resident result");
    CU(cudaSetDevice(c->device));
    const uint32_t nQ = c->nQ, maxHits = (uint32_t)c->par.max_res_list_len;
    if (hits && hitCount) {
        CU(cudaMemcpy(hits, c->dHits.p, (size_t)nQ * maxHits * sizeof(MbPrefHit), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(hitCount, c->dHitCount.p, (size_t)nQ * 4, cudaMemcpyDeviceToHost));
    }
    if (alns && alnCount) {
        if (!c->haveAln) return fail(MB200_E_STATE, "no resident alignment result");
        for (uint32_t i = 0; i < nQ; i++) alnCount[i] = 0;
        std::vector<uint32_t> tids; TRY(trace_tids(c, tids));
        for (size_t r = 0; r < c->hTrace.size(); r++) {
            const uint32_t q = c->hTrace[r].slot / maxHits; const uint32_t L = (uint32_t)(c->hQoff[q + 1] - c->hQoff[q]);
            mb200_aln a; finish_aln(c, c->hTrace[r], tids[r], L, a);
            if (a.passed) alns[(size_t)q * maxHits + alnCount[q]++] = a;
        }
    }
    return 0;
}

int mb200_prefilter(mb200_ctx *c, const uint8_t *q, const uint64_t *qoff, uint32_t nQ, mb200_hit *hits, uint32_t *hitCount) {
    if (!hits || !hitCount) return fail(MB200_E_ARG, "null output");
    TRY(mb200_queries_upload(c, q, qoff, nQ));
    TRY(mb200_prefilter_resident(c));
    return mb200_results_download(c, hits, hitCount, nullptr, nullptr);
}

int mb200_prefilter_compact(mb200_ctx *c, const uint8_t *q, const uint64_t *qoff, uint32_t nQ, uint64_t *hitOff, const mb200_hit **hits) {
    if (!hitOff || !hits) return fail(MB200_E_ARG, "null output");
    const auto t0 = std::chrono::steady_clock::now();
    TRY(mb200_queries_upload(c, q, qoff, nQ));
    TRY(mb200_prefilter_resident(c));
    TRY(download_hits_compact(c));
    memcpy(hitOff, c->hHitOff.data(), (size_t)(nQ + 1) * 8);
    c->st.ms_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *hits = c->nFlatHits ? c->pFlatHits.p : nullptr;
    return 0;
}

// accepted alignments of the resident pass, compact: fills c->hAlns / c->hAlnOff (prefilter order inside a query)
static int collect_accepted(mb200_ctx *c, const std::vector<uint32_t> &tidOfTrace) {
    const uint32_t nQ = c->nQ, maxHits = (uint32_t)c->par.max_res_list_len;
    const size_t n = c->hTrace.size();
    if (c->hAlnTmp.size() < n) c->hAlnTmp.resize(n);
    mb200_aln *tmp = c->hAlnTmp.data();
    // records are sorted by slot (query-major): the accepted ones keep that order.  Chunks of the record list are finished and counted in parallel, then copied to their place.
    const int nChunk = (int)std::min<size_t>(256, std::max<size_t>(1, n / 4096));
    std::vector<size_t> cBeg(nChunk + 1), cOut(nChunk + 1, 0);
    for (int k = 0; k <= nChunk; k++) cBeg[k] = n * (size_t)k / (size_t)nChunk;
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < nChunk; k++) {
        size_t cnt = 0;
        for (size_t r = cBeg[k]; r < cBeg[k + 1]; r++) {
            const uint32_t q = c->hTrace[r].slot / maxHits; const uint32_t L = (uint32_t)(c->hQoff[q + 1] - c->hQoff[q]);
            finish_aln(c, c->hTrace[r], tidOfTrace[r], L, tmp[r]);
            cnt += tmp[r].passed ? 1 : 0;
        }
        cOut[k + 1] = cnt;
    }
    for (int k = 0; k < nChunk; k++) cOut[k + 1] += cOut[k];
    c->hAlns.resize(cOut[nChunk]);
    c->hAlnOff.assign((size_t)nQ + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < nChunk; k++) {
        size_t w = cOut[k];
        for (size_t r = cBeg[k]; r < cBeg[k + 1]; r++) if (tmp[r].passed) c->hAlns[w++] = tmp[r];
    }
    for (size_t r = 0; r < n; r++) if (tmp[r].passed) c->hAlnOff[c->hTrace[r].slot / maxHits + 1]++;
    for (uint32_t i = 0; i < nQ; i++) c->hAlnOff[i + 1] += c->hAlnOff[i];
    return 0;
}

// target ids of the traced slots, gathered on the device
__global__ void k_gather_tid(const MbPrefHit *__restrict__ hits, const SwTraceOut *__restrict__ tr, uint32_t n, uint32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hits[tr[i].slot].seqId;
}
static int trace_tids(mb200_ctx *c, std::vector<uint32_t> &tids) {
    const size_t n = c->hTrace.size(); tids.resize(n);
    if (!n) return 0;
    // hTrace was sorted on the host; re-upload so the gather follows the same order
    TRY(c->dTrace.alloc(n)); TRY(c->dFlatSeq.alloc(n));
    TRY(upload(c->dTrace.p, c->hTrace.data(), n * sizeof(SwTraceOut), c->stream));
    MB_LAUNCH((k_gather_tid), (unsigned)((n + 255) / 256), 256, 0, c->stream, c->dHits.p, c->dTrace.p, (uint32_t)n, c->dFlatSeq.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(tids.data(), c->dFlatSeq.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    TRY(sync_check(c, "gather tid")); c->st.launches++;
    return 0;
}

int mb200_align(mb200_ctx *c, const uint8_t *q, const uint64_t *qoff, uint32_t nQ, const uint32_t *hitSeq, const uint64_t *hitOff, mb200_aln *alns) {
    if (!hitSeq || !hitOff || !alns) return fail(MB200_E_ARG, "null argument");
    if (!c || !c->haveIndex) return fail(MB200_E_STATE, "index not set");
    TRY(mb200_queries_upload(c, q, qoff, nQ));
    const uint32_t maxHits = (uint32_t)c->par.max_res_list_len;
    const uint64_t n = hitOff[nQ];
    for (uint32_t i = 0; i < nQ; i++) if (hitOff[i + 1] - hitOff[i] > maxHits) return fail(MB200_E_LIMIT, "query %u has more than max_res_list_len hits", i);
    for (uint64_t k = 0; k < n; k++) if (hitSeq[k] >= c->nT) return fail(MB200_E_ARG, "hit target id %u out of range", hitSeq[k]);
    TRY(c->dHits.alloc((size_t)nQ * maxHits)); TRY(c->dHitCount.alloc(nQ + 1)); TRY(c->dFlatSeq.alloc(n ? n : 1)); TRY(c->dHitOff64.alloc(nQ + 1));
    if (n) TRY(upload(c->dFlatSeq.p, hitSeq, n * 4, c->stream));
    TRY(upload(c->dHitOff64.p, hitOff, (size_t)(nQ + 1) * 8, c->stream));
    if (nQ) { MB_LAUNCH((k_scatter_hits), nQ, 64, 0, c->stream, c->dFlatSeq.p, c->dHitOff64.p, nQ, maxHits, c->dHits.p, c->dHitCount.p); CU(cudaGetLastError()); c->st.launches++; }
    c->havePref = true;
    TRY(align_pipeline(c));
    for (uint32_t i = 0; i < nQ; i++) for (uint64_t k = hitOff[i]; k < hitOff[i + 1]; k++) { memset(&alns[k], 0, sizeof(mb200_aln)); alns[k].seq_id = hitSeq[k]; alns[k].q_start = alns[k].db_start = -1; }
    for (const SwTraceOut &o : c->hTrace) {
        const uint32_t qi = o.slot / maxHits, k = o.slot % maxHits; const uint64_t pos = hitOff[qi] + k;
        finish_aln(c, o, hitSeq[pos], (uint32_t)(c->hQoff[qi + 1] - c->hQoff[qi]), alns[pos]);
    }
    return 0;
}

int mb200_search(mb200_ctx *c, const uint8_t *q, const uint64_t *qoff, uint32_t nQ, uint64_t *alnOff, const mb200_aln **alns) {
    if (!alnOff || !alns) return fail(MB200_E_ARG, "null output");
    TRY(mb200_queries_upload(c, q, qoff, nQ));
    TRY(mb200_prefilter_resident(c));
    TRY(align_pipeline(c));
    std::vector<uint32_t> tids; TRY(trace_tids(c, tids));
    TRY(collect_accepted(c, tids));
    memcpy(alnOff, c->hAlnOff.data(), (size_t)(nQ + 1) * 8);
    *alns = c->hAlns.empty() ? nullptr : c->hAlns.data();
    return 0;
}

int mb200_search_resident(mb200_ctx *c, uint64_t *alnOff, const mb200_aln **alns) {
    if (!c || !alnOff || !alns) return fail(MB200_E_ARG, "null argument");
    TRY(mb200_prefilter_resident(c));
    TRY(align_pipeline(c));
    std::vector<uint32_t> tids; TRY(trace_tids(c, tids));
    TRY(collect_accepted(c, tids));
    memcpy(alnOff, c->hAlnOff.data(), (size_t)(c->nQ + 1) * 8);
    *alns = c->hAlns.empty() ? nullptr : c->hAlns.data();
    return 0;
}

int mb200_chain(mb200_ctx *c, const mb200_cand *cands, const uint64_t *candOff, uint32_t nContigs, uint64_t totalTargetResidues, mb200_call *calls,
                uint32_t *callCount, mb200_exon *exons) {
    if (!c || !cands || !candOff || !calls || !callCount || !exons) return fail(MB200_E_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    const uint64_t n = candOff[nContigs];
    for (uint32_t i = 0; i < nContigs; i++) callCount[i] = 0;
    if (n == 0) return 0;
    static_assert(sizeof(MbCand) == sizeof(mb200_cand) && sizeof(MbCall) == sizeof(mb200_call) && sizeof(MbExon) == sizeof(mb200_exon), "ABI structs");
    // grow-only buffers owned by the context (a cudaMalloc / cudaFree pair per call costs more than the kernel)
    TRY(c->dChCand.alloc(n)); TRY(c->dChOff.alloc(nContigs + 1)); TRY(c->dChPE.alloc(n)); TRY(c->dChOp.alloc(n)); TRY(c->dChOm.alloc(n)); TRY(c->dChRows.alloc(n));
    TRY(c->dChCalls.alloc(2 * n)); TRY(c->dChEx.alloc(n)); TRY(c->dChCallCnt.alloc(nContigs)); TRY(c->dChExCnt.alloc(nContigs));
    TRY(c->pChCalls.alloc(2 * n)); TRY(c->pChEx.alloc(n)); TRY(c->pChCnt.alloc(nContigs));
    TRY(upload(c->dChCand.p, cands, n * sizeof(MbCand), c->stream)); TRY(upload(c->dChOff.p, candOff, (size_t)(nContigs + 1) * 8, c->stream));
    MbChainParams P; P.minIntron = c->par.min_intron; P.maxIntron = c->par.max_intron; P.minExonAa = c->par.min_exon_aa; P.maxOverlap = c->par.max_overlap;
    P.gapOpen = c->par.set_gap_open; P.gapExtend = c->par.set_gap_extend; P.maxSets = c->par.max_exon_sets; P.tcov = c->par.tcov;
    CU(cudaEventRecord(c->ev0, c->stream));
    TRY(c->dChGrpCalls.alloc(n)); TRY(c->dChGrpCnt.alloc(n));
    MB_LAUNCH((k_chain_groups), (unsigned)((n + 127) / 128), 128, 0, c->stream, c->dChCand.p, c->dChOff.p, nContigs, n, P, c->dChPE.p, c->dChOp.p, c->dChOm.p, c->dChRows.p,
              c->dChGrpCalls.p, c->dChEx.p, c->dChGrpCnt.p);
    CU(cudaGetLastError());
    MB_LAUNCH((k_chain_compact), (nContigs + 127) / 128, 128, 0, c->stream, c->dChOff.p, nContigs, c->dChGrpCnt.p, c->dChGrpCalls.p, c->dChCalls.p, c->dChCallCnt.p);
    CU(cudaGetLastError()); c->st.launches++;
    CU(cudaEventRecord(c->ev1, c->stream));
    // only the used prefix of every contig's call / exon range is meaningful, but the ranges are interleaved with unused slots: copy whole arrays
    CU(cudaMemcpyAsync(c->pChCalls.p, c->dChCalls.p, 2 * n * sizeof(MbCall), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->pChEx.p, c->dChEx.p, n * sizeof(MbExon), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->pChCnt.p, c->dChCallCnt.p, (size_t)nContigs * 4, cudaMemcpyDeviceToHost, c->stream));
    TRY(sync_check(c, "chain kernel"));
    float ms = 0; CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->st.ms_chain = ms; c->st.launches++; stage_interval(c, c->ev0, c->ev1);
    // host epilogue: set e-value and --metaeuk-eval filter (collectoptimalset.cpp:323-330); compact per contig
    const MbCall *hCalls = c->pChCalls.p; const MbExon *hEx = c->pChEx.p; const int32_t *hCnt = c->pChCnt.p;
    for (uint32_t ci = 0; ci < nContigs; ci++) if (hCnt[ci] < 0) return fail(MB200_E_ARG, "contig %u: illegal codon segment length", ci);
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t ci = 0; ci < nContigs; ci++) {
        const uint64_t b = candOff[ci];
        callCount[ci] = mb200::finish_calls(reinterpret_cast<const mb200_call *>(&hCalls[2 * b]), (uint32_t)hCnt[ci], reinterpret_cast<const mb200_exon *>(&hEx[b]), b, totalTargetResidues,
                                            c->par.metaeuk_eval, calls + 2 * b, exons + b);
    }
    return 0;
}

// ---- predictexons from contigs, binary records between the stages -----------------------------------------
int mb200_target_keys_set(mb200_ctx *c, const uint32_t *keys, uint32_t nT) {
    if (!c || (!keys && nT)) return fail(MB200_E_ARG, "null argument");
    if (c->haveIndex && nT != c->nT && nT != 0) return fail(MB200_E_ARG, "%u keys for %u targets", nT, c->nT);
    c->tKeys.assign(keys, keys + nT);
    return 0;
}

int mb200_text_exon_line(const mb200_call *call, const mb200_pexon *exon, char *buf, size_t cap) {
    if (!call || !exon || !buf) return fail(MB200_E_ARG, "null argument");
    const std::string l = mb200::text_exon_line(*call, *exon);
    if (l.size() + 1 > cap) return fail(MB200_E_ARG, "buffer too small for one exon line (%zu bytes needed)", l.size() + 1);
    memcpy(buf, l.c_str(), l.size() + 1);
    return (int)l.size();
}

int64_t mb200_text_dp_db(const uint64_t *callOff, uint32_t nContigs, const uint32_t *contigKeys, const mb200_call *calls, const mb200_pexon *exons, char *buf, size_t cap) {
    if (!callOff || (callOff[nContigs] && (!calls || !exons))) return fail(MB200_E_ARG, "null argument");
    std::vector<std::string> part(nContigs);
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t ci = 0; ci < nContigs; ci++) {
        if (callOff[ci + 1] == callOff[ci]) continue;
        std::string &s = part[ci]; s = ">" + std::to_string(contigKeys ? contigKeys[ci] : ci) + "\n";
        for (uint64_t k = callOff[ci]; k < callOff[ci + 1]; k++)
            for (int x = 0; x < calls[k].n_exons; x++) s += mb200::text_exon_line(calls[k], exons[calls[k].exon_begin + x]);
    }
    size_t need = 0; for (const std::string &s : part) need += s.size();
    if (buf && need <= cap) { size_t w = 0; for (const std::string &s : part) { memcpy(buf + w, s.data(), s.size()); w += s.size(); } }
    return (int64_t)need;
}

// One batch of contigs [b0, b1) on context w (the caller's context, or one of its two workers): ORFs -> prefilter -> alignment -> join -> chaining.
// Fragment keys count over the whole call, so a batch needs the number of fragments of all batches before it: it publishes its own count right after
// its ORF stage and asks for its base just before the join (by then the earlier batches have long published theirs).
struct PeBatchOut { std::vector<mb200_call> calls; std::vector<mb200_pexon> exons; std::vector<uint64_t> callEnd; uint64_t nFrag = 0, nAln = 0; mb200_stats st; };
struct PeShared {
    std::mutex mu; std::condition_variable cv; std::vector<int64_t> nFragOf; uint64_t base = 0; bool failed = false;
    void publish(uint32_t b, uint64_t n) { { std::lock_guard<std::mutex> l(mu); nFragOf[b] = (int64_t)n; } cv.notify_all(); }
    void abort() { { std::lock_guard<std::mutex> l(mu); failed = true; } cv.notify_all(); }
    bool orfBase(uint32_t b, uint64_t &out) {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [&] { if (failed) return true; for (uint32_t i = 0; i < b; i++) if (nFragOf[i] < 0) return false; return true; });
        if (failed) return false;
        out = base; for (uint32_t i = 0; i < b; i++) out += (uint64_t)nFragOf[i];
        return true;
    }
};

static int predict_batch(mb200_ctx *c, const char *nuc, const uint64_t *off, uint32_t b0, uint32_t b1, int minLen, int maxLen, PeShared &sh, uint32_t batchIdx, PeBatchOut &out, double *ph) {
    const uint32_t nb = b1 - b0;
    uint64_t n = 0, na = 0; const mb200_orf *orfs = nullptr; const char *aa = nullptr;
    auto tp = std::chrono::steady_clock::now();
    auto lap = [&](int k) { const auto now = std::chrono::steady_clock::now(); ph[k] += std::chrono::duration<double, std::milli>(now - tp).count(); tp = now; };
    memset(&c->st, 0, sizeof(c->st));
    struct InPredict { mb200_ctx *c; InPredict(mb200_ctx *x) : c(x) { c->inPredict = true; } ~InPredict() { c->inPredict = false; } } inPredictGuard(c);
    TRY(mb200_extract_orfs(c, nuc, off + b0, nb, minLen, maxLen, 1, &n, &orfs, &na, &aa)); lap(0);
    sh.publish(batchIdx, n);
    TRY(mb200_queries_from_orfs(c)); lap(1);
    TRY(mb200_prefilter_resident(c)); lap(2);
    TRY(align_pipeline(c)); lap(3);
    std::vector<uint32_t> tids; TRY(trace_tids(c, tids)); lap(4);
    TRY(collect_accepted(c, tids)); lap(5);
    uint64_t orfBase = 0;
    if (!sh.orfBase(batchIdx, orfBase)) return fail(MB200_E_INTERNAL, "another contig batch failed");
    std::vector<mb200_cand> &cands = c->peCands; std::vector<uint64_t> &candOff = c->peCandOff, &candAln = c->peCandAln;
    std::vector<mb200_call> &bCalls = c->peBCalls; std::vector<mb200_exon> &bExons = c->peBExons; std::vector<uint32_t> &bCnt = c->peBCnt;
    const auto tj = std::chrono::steady_clock::now();
    if (mb200::join_by_contig(orfs, n, nb, orfBase, c->hAlnOff.data(), c->hAlns.data(), c->tKeys.empty() ? nullptr : c->tKeys.data(), c->hToff.data(), cands, candOff, candAln))
        return fail(MB200_E_LIMIT, "more than 2^32 fragments in one mb200_predict_exons call");
    c->st.ms_join = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tj).count();
    const uint64_t nc = candOff[nb];
    if (bCalls.size() < std::max<uint64_t>(1, 2 * nc)) bCalls.resize(std::max<uint64_t>(1, 2 * nc));      // (grow only: mb200_chain writes every record that is read below)
    if (bExons.size() < std::max<uint64_t>(1, nc)) bExons.resize(std::max<uint64_t>(1, nc));
    bCnt.assign(nb, 0);
    c->st.ms_chain = 0;
    lap(6);
    if (nc) TRY(mb200_chain(c, cands.data(), candOff.data(), nb, c->totalRes, bCalls.data(), bCnt.data(), bExons.data()));
    lap(7);
    out.calls.clear(); out.exons.clear(); out.callEnd.assign(nb, 0);
    for (uint32_t ci = 0; ci < nb; ci++) {
        const uint64_t base = candOff[ci];
        for (uint32_t k = 0; k < bCnt[ci]; k++) {
            mb200_call cl = bCalls[2 * base + k]; const uint32_t eb = cl.exon_begin;
            cl.exon_begin = (uint32_t)out.exons.size();     // relative to the batch; shifted when the batches are concatenated
            for (int x = 0; x < cl.n_exons; x++) {
                const mb200_exon &ex = bExons[eb + x]; const mb200_cand &k2 = cands[base + ex.cand_idx]; const mb200_aln &al = c->hAlns[candAln[base + ex.cand_idx]];
                mb200_pexon pe; memset(&pe, 0, sizeof(pe));
                pe.orf_key = k2.orf_key; pe.bit_score = k2.bit_score; pe.seq_id_est = al.seq_id_est; pe.t_start = k2.t_start; pe.t_end = k2.t_end; pe.t_len = k2.t_len;
                pe.contig_start = ex.contig_start; pe.contig_end = ex.contig_end; pe.nuc_len = ex.nuc_len; pe.frag_from = k2.frag_from; pe.frag_to = k2.frag_to; pe.evalue = al.evalue;
                out.exons.push_back(pe);
            }
            out.calls.push_back(cl);
        }
        out.callEnd[ci] = out.calls.size();
    }
    out.nFrag = n; out.nAln = c->hAlns.size(); out.st = c->st;
    lap(8);
    return 0;
}

// A worker context shares the caller's read-only device tables (views) and owns its stream, query batch, hit lists and scratch: two of them take
// the contig batches of one mb200_predict_exons call alternately, each on its own host thread, so the host phases of one batch (ORF record scan,
// alignment epilogue, join, call assembly, PCIe copies) run while the kernels of the other batch occupy the GPU.
static int worker_sync(mb200_ctx *p, mb200_ctx *&w) {
    if (!w) {
        w = new mb200_ctx(); w->device = p->device; w->nSM = p->nSM;
        CU(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&w->ev0)); CU(cudaEventCreate(&w->ev1)); CU(cudaEventCreate(&w->evT0)); CU(cudaEventCreate(&w->evT1));
        memset(&w->st, 0, sizeof(w->st));
    }
    if (w->par.eval_thr != p->par.eval_thr || w->totalRes != p->totalRes || w->par.translation_table != p->par.translation_table) { w->minScore.clear(); w->haveOrfTab = false; }
    w->par = p->par; w->mat = p->mat; w->haveMat = p->haveMat; w->ev = p->ev;
    w->dSim3.set_view(p->dSim3.p, p->dSim3.n); w->dCnt3.set_view(p->dCnt3.p, p->dCnt3.n); w->dOffsets.set_view(p->dOffsets.p, p->dOffsets.n);
    w->dKmerBits.set_view(p->dKmerBits.p, p->dKmerBits.n); w->useBits = p->useBits; w->dEntries.set_view(p->dEntries.p, p->dEntries.n);
    w->optBuckets = p->optBuckets; w->optStreamHint = p->optStreamHint; w->optSortWalk = p->optSortWalk; w->optOverflowPolicy = p->optOverflowPolicy; w->haveBuckets = p->haveBuckets; if (p->haveBuckets) w->dBuckets.set_view(p->dBuckets.p, p->dBuckets.n);
    w->dTseqM.set_view(p->dTseqM.p, p->dTseqM.n); w->dTseqU.set_view(p->dTseqU.p, p->dTseqU.n); w->dToff.set_view(p->dToff.p, p->dToff.n);
    w->dSeedMat.set_view(p->dSeedMat.p, p->dSeedMat.n); w->dAlnMat.set_view(p->dAlnMat.p, p->dAlnMat.n); w->dSeedPB.set_view(p->dSeedPB.p, p->dSeedPB.n);
    w->dAlnPB.set_view(p->dAlnPB.p, p->dAlnPB.n); w->dUng8.set_view(p->dUng8.p, p->dUng8.n);
    w->nEntries = p->nEntries; w->totalRes = p->totalRes; w->nT = p->nT; w->hToff = p->hToff; w->haveIndex = p->haveIndex; w->haveSim3 = p->haveSim3; w->tKeys = p->tKeys;
    w->optPrefCfg = p->optPrefCfg; w->optPrefShortLen = p->optPrefShortLen; w->optPrefShortCfg = p->optPrefShortCfg; w->optPhaseStats = p->optPhaseStats; w->optSwKernel = p->optSwKernel; w->optSwRing = p->optSwRing; w->optSwRingFit = p->optSwRingFit; w->optTraceKernel = p->optTraceKernel; w->optLaneMaxHits = p->optLaneMaxHits;
    return 0;
}

int mb200_predict_exons(mb200_ctx *c, const char *nuc, const uint64_t *off, uint32_t nContigs, int minLen, int maxLen,
                        uint64_t *callOff, const mb200_call **calls, const mb200_pexon **exons, uint64_t *nExons) {
    if (!c || !nuc || !off || !callOff || !calls || !exons || !nExons) return fail(MB200_E_ARG, "null argument");
    if (!c->haveIndex || !c->haveSim3) return fail(MB200_E_STATE, "index / matrices not set");
    if (!c->tKeys.empty() && c->tKeys.size() != c->nT) return fail(MB200_E_STATE, "mb200_target_keys_set: %zu keys for %u targets", c->tKeys.size(), c->nT);
    const auto t0 = std::chrono::steady_clock::now();
    CU(cudaSetDevice(c->device));
    c->peCalls.clear(); c->peExons.clear(); callOff[0] = 0; *calls = nullptr; *exons = nullptr; *nExons = 0;
    // contig batches of about predict_batch_nt nucleotides
    std::vector<uint32_t> bnd(1, 0);
    // Default size: 32 M nucleotides, but at least four batches per worker stream when the call is small (a rank's shard of a contig DB split over 4 or 8 GPUs), so that the host phases
    // of one batch still hide under the kernels of another -- C4/10 (50 M nucleotides) on two streams: 866 -> 830 ms per call with 6 M-nucleotide batches (call u); never below 4 M.
    uint64_t batchNt = c->optBatchNt;
    if (!c->optBatchNtSet && c->optStreams > 1 && nContigs > 0) {
        const uint64_t per = (off[nContigs] - off[0]) / (4ull * (uint64_t)c->optStreams);
        batchNt = std::min<uint64_t>(batchNt, std::max<uint64_t>(per, 4ull << 20));
    }
    for (uint32_t b0 = 0; b0 < nContigs;) { uint32_t b1 = b0 + 1; while (b1 < nContigs && off[b1 + 1] - off[b0] <= batchNt) b1++; bnd.push_back(b1); b0 = b1; }
    const uint32_t nBatches = (uint32_t)bnd.size() - 1;
    std::vector<PeBatchOut> outs(nBatches);
    PeShared sh; sh.nFragOf.assign(nBatches, -1); sh.base = c->optOrfKeyBase;
    const int nW = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)c->optStreams, nBatches));
    // origin of the call's device timeline: every stage of every worker stream records its [start, end] relative to it (ms_dev_busy = length of their union)
    if (!c->evOrigin) CU(cudaEventCreate(&c->evOrigin));
    CU(cudaEventRecord(c->evOrigin, c->stream)); CU(cudaEventSynchronize(c->evOrigin));
    struct OriginGuard { mb200_ctx *c; int n; ~OriginGuard() { c->originRef = nullptr; for (int k = 0; k < n; k++) if (c->workers[k]) c->workers[k]->originRef = nullptr; } } originGuard{c, 4};
    c->originRef = c->evOrigin; c->busyIv.clear();
    double ph[4][9] = {{0}};
    mb200_stats wst[4]; memset(wst, 0, sizeof(wst));
    int rcs[4] = {0, 0, 0, 0}; std::string errs[4];
    const mb200_stats before = c->st;
    auto addStats = [](mb200_stats &a, const mb200_stats &s) {
        a.kmers += s.kmers; a.lookups += s.lookups; a.db_matches += s.db_matches; a.double_hits += s.double_hits; a.diag_cells += s.diag_cells; a.pref_results += s.pref_results;
        a.sw_pairs += s.sw_pairs; a.sw_cells_fwd += s.sw_cells_fwd; a.sw_cells_rev += s.sw_cells_rev; a.sw_target_bytes += s.sw_target_bytes;
        a.query_residues += s.query_residues; a.n_queries += s.n_queries;
        a.ms_prep += s.ms_prep; a.ms_prefilter += s.ms_prefilter; a.ms_sw_fwd += s.ms_sw_fwd; a.ms_sw_rev += s.ms_sw_rev; a.ms_chain += s.ms_chain; a.ms_h2d += s.ms_h2d; a.ms_join += s.ms_join;
        a.ms_orf += s.ms_orf; a.ms_orf_host += s.ms_orf_host; a.orf_nucleotides += s.orf_nucleotides; a.orf_fragments += s.orf_fragments; a.orf_residues += s.orf_residues;
        for (int i = 0; i < 16; i++) a.pref_phase_cycles[i] += s.pref_phase_cycles[i];
        a.launches += s.launches; a.overflow_queries += s.overflow_queries;
    };
    if (nW == 1) {
        for (uint32_t b = 0; b < nBatches && !rcs[0]; b++) {
            rcs[0] = predict_batch(c, nuc, off, bnd[b], bnd[b + 1], minLen, maxLen, sh, b, outs[b], ph[0]);
            if (!rcs[0]) addStats(wst[0], outs[b].st);
        }
        if (rcs[0]) return rcs[0];
    } else {
        TRY(build_buckets(c));      // (shared with the workers as a view)
        for (int k = 0; k < nW; k++) { TRY(worker_sync(c, c->workers[k])); c->workers[k]->originRef = c->evOrigin; c->workers[k]->busyIv.clear(); }
        std::atomic<uint32_t> next(0);
        auto run = [&](int k) {
            mb200_ctx *w = c->workers[k];
            cudaSetDevice(w->device);
            for (;;) {
                const uint32_t b = next.fetch_add(1);
                if (b >= nBatches) break;
                { std::lock_guard<std::mutex> l(sh.mu); if (sh.failed) break; }
                const int rc = predict_batch(w, nuc, off, bnd[b], bnd[b + 1], minLen, maxLen, sh, b, outs[b], ph[k]);
                if (rc) { rcs[k] = rc; errs[k] = mb200_last_error(); sh.abort(); break; }
                addStats(wst[k], outs[b].st);
            }
        };
        std::vector<std::thread> th;
        for (int k = 1; k < nW; k++) th.emplace_back(run, k);
        run(0);
        for (std::thread &t : th) t.join();
        for (int k = 0; k < nW; k++) if (rcs[k]) return fail(rcs[k], "%s", errs[k].c_str());
    }
    // concatenate the batches in contig order
    size_t nCalls = 0, nEx = 0;
    for (const PeBatchOut &o : outs) { nCalls += o.calls.size(); nEx += o.exons.size(); }
    if (nEx >= (1ull << 32)) return fail(MB200_E_LIMIT, "more than 2^32 exons in one mb200_predict_exons call");
    c->peCalls.reserve(nCalls); c->peExons.reserve(nEx);
    mb200_stats acc; memset(&acc, 0, sizeof(acc));
    for (uint32_t b = 0; b < nBatches; b++) {
        const PeBatchOut &o = outs[b];
        const uint64_t callBase = c->peCalls.size(); const uint32_t exBase = (uint32_t)c->peExons.size();
        for (mb200_call cl : o.calls) { cl.exon_begin += exBase; c->peCalls.push_back(cl); }
        c->peExons.insert(c->peExons.end(), o.exons.begin(), o.exons.end());
        for (uint32_t ci = 0; ci < bnd[b + 1] - bnd[b]; ci++) callOff[bnd[b] + ci + 1] = callBase + o.callEnd[ci];
        acc.pe_batches++; acc.pe_contigs += bnd[b + 1] - bnd[b]; acc.pe_fragments += o.nFrag; acc.pe_alignments += o.nAln;
    }
    // work counters: sums over all batches; stage times (CUDA events): those of the busier worker stream -- with two streams the kernels of two batches share
    // the GPU, so the per-stream times overlap in wall-clock terms and their sum over both streams would count the same interval twice
    addStats(acc, wst[0]);
    for (int k = 0; k < nW; k++) {
        acc.ms_prefilter_sum += wst[k].ms_prefilter; acc.ms_sw_fwd_sum += wst[k].ms_sw_fwd; acc.ms_sw_rev_sum += wst[k].ms_sw_rev;
        acc.ms_dev_sum += wst[k].ms_orf + wst[k].ms_prep + wst[k].ms_prefilter + wst[k].ms_sw_fwd + wst[k].ms_sw_rev + wst[k].ms_chain;
    }
    if (nW >= 2) {
        auto dev = [](const mb200_stats &s) { return s.ms_orf + s.ms_prep + s.ms_prefilter + s.ms_sw_fwd + s.ms_sw_rev + s.ms_chain; };
        int busyK = 0;
        for (int k = 1; k < nW; k++) { addStats(acc, wst[k]); if (dev(wst[k]) > dev(wst[busyK])) busyK = k; }
        const mb200_stats &busy = wst[busyK];
        acc.ms_orf = busy.ms_orf; acc.ms_prep = busy.ms_prep; acc.ms_prefilter = busy.ms_prefilter; acc.ms_sw_fwd = busy.ms_sw_fwd; acc.ms_sw_rev = busy.ms_sw_rev; acc.ms_chain = busy.ms_chain;
        acc.ms_join = busy.ms_join; acc.ms_h2d = busy.ms_h2d; acc.ms_orf_host = busy.ms_orf_host;
    }
    {   // union of the stage intervals of all streams
        std::vector<std::pair<float, float>> iv(c->busyIv);
        if (nW >= 2) for (int k = 0; k < nW; k++) iv.insert(iv.end(), c->workers[k]->busyIv.begin(), c->workers[k]->busyIv.end());
        std::sort(iv.begin(), iv.end());
        double busy = 0; float curB = 0, curE = -1;
        for (const auto &x : iv) {
            if (curE < curB || x.first > curE) { if (curE >= curB) busy += (double)curE - curB; curB = x.first; curE = x.second; }
            else if (x.second > curE) curE = x.second;
        }
        if (curE >= curB) busy += (double)curE - curB;
        acc.ms_dev_busy = busy;
    }
    if (getenv("MB200_TRACE"))
        for (int k = 0; k < nW; k++)
            fprintf(stderr, "mb200_predict_exons worker %d host phases (ms, wall): orfs %.1f install %.1f prefilter %.1f align %.1f tids %.1f accept %.1f join %.1f chain %.1f output %.1f\n",
                    k, ph[k][0], ph[k][1], ph[k][2], ph[k][3], ph[k][4], ph[k][5], ph[k][6], ph[k][7], ph[k][8]);
    acc.pe_calls = c->peCalls.size(); acc.pe_exons = c->peExons.size();
    acc.launches += before.launches; acc.ms_reduce = before.ms_reduce; acc.ms_d2h = before.ms_d2h; acc.ms_host = before.ms_host;
    acc.pe_streams = (uint64_t)nW;
    acc.ms_predict = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->st = acc;
    *calls = c->peCalls.empty() ? nullptr : c->peCalls.data(); *exons = c->peExons.empty() ? nullptr : c->peExons.data(); *nExons = c->peExons.size();
    return 0;
}

int mb200_query_status_get(mb200_ctx *c, uint32_t *status) {
    if (!c || !status) return fail(MB200_E_ARG, "null argument");
    if (!c->havePref || !c->pStatus.p) return fail(MB200_E_STATE, "no resident prefilter result");
    memcpy(status, c->pStatus.p, (size_t)c->nQ * 4);
    return 0;
}
int mb200_stats_get(mb200_ctx *c, mb200_stats *out) { if (!c || !out) return fail(MB200_E_ARG, "null argument"); *out = c->st; return 0; }
// the work counters and stage times of the last call as one JSON object (what bench.py reads through mb200_stats_get, for callers without a struct mirror)
int mb200_profile_dump(mb200_ctx *c, const char *jsonPath) {
    if (!c || !jsonPath) return fail(MB200_E_ARG, "null argument");
    FILE *f = fopen(jsonPath, "w");
    if (!f) return fail(MB200_E_ARG, "cannot create %s", jsonPath);
    const mb200_stats &s = c->st;
    fprintf(f, "{\"build\": \"%s\", \"device\": %d, \"sms\": %d,\n", mb200_build_kind(), c->device, c->nSM);
#define MB_U(k) fprintf(f, " \"" #k "\": %llu,", (unsigned long long)s.k)
#define MB_D(k) fprintf(f, " \"" #k "\": %.6f,", (double)s.k)
    MB_U(kmers); MB_U(lookups); MB_U(db_matches); MB_U(double_hits); MB_U(diag_cells); MB_U(pref_results); fputc('\n', f);
    MB_U(sw_pairs); MB_U(sw_cells_fwd); MB_U(sw_cells_rev); MB_U(sw_target_bytes); MB_U(query_residues); MB_U(n_queries); fputc('\n', f);
    MB_D(ms_orf); MB_D(ms_orf_host); MB_D(ms_prep); MB_D(ms_prefilter); MB_D(ms_sw_fwd); MB_D(ms_sw_rev); MB_D(ms_chain); MB_D(ms_reduce); MB_D(ms_h2d); MB_D(ms_d2h); MB_D(ms_host); MB_D(ms_join); MB_D(ms_predict); fputc('\n', f);
    MB_D(ms_prefilter_sum); MB_D(ms_sw_fwd_sum); MB_D(ms_sw_rev_sum); MB_D(ms_dev_sum); MB_D(ms_dev_busy); fputc('\n', f);
    MB_U(orf_nucleotides); MB_U(orf_fragments); MB_U(orf_residues); MB_U(pe_batches); MB_U(pe_contigs); MB_U(pe_fragments); MB_U(pe_alignments); MB_U(pe_calls); MB_U(pe_exons); MB_U(pe_streams); fputc('\n', f);
    MB_U(launches); MB_U(overflow_queries);
#undef MB_U
#undef MB_D
    fprintf(f, " \"pref_phase_cycles\": [");
    for (int i = 0; i < 16; i++) fprintf(f, "%s%llu", i ? ", " : "", (unsigned long long)s.pref_phase_cycles[i]);
    fprintf(f, "]}\n");
    if (fclose(f) != 0) return fail(MB200_E_ARG, "write error on %s", jsonPath);
    return 0;
}
double mb200_evalue(mb200_ctx *c, int rawScore, int qLen) { return c->ev.evalue(rawScore, qLen, (double)c->totalRes); }
int mb200_bit_score(mb200_ctx *c, int rawScore) { return c->ev.bitScore(rawScore); }

}  // extern "C"
