// tests/simt/simt.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory).
// Cooperative-fiber execution of a CUDA thread block on one OS thread: every thread of the block is a ucontext fiber that runs until it reaches a
// synchronisation point and is resumed once all participants of that point have arrived.  Blocks of a grid run one after the other.
#include "cuda_runtime.h"

#include <setjmp.h>
#include <ucontext.h>

#include <mutex>
#include <vector>

namespace simt {

enum State { READY, WAIT_BLOCK, WAIT_WARP, DONE };
struct Fiber { ucontext_t ctx; jmp_buf jb; bool started; Thread th; State state; uint64_t result; unsigned char *stack; };
struct Warp { unsigned arrived, exited; uint64_t val[32]; int param[32], width[32]; Op op[32]; };

Thread *cur = nullptr;
static const size_t STACK = 512 * 1024;
static jmp_buf schedJb;      // fibers are entered once with setcontext and switched with _setjmp/_longjmp afterwards (no signal-mask system call per switch)
static std::vector<Fiber *> fibers;
static Fiber *curF = nullptr;
static std::vector<Warp> warps;
static int nLive = 0;
static unsigned barCount = 0;
static const std::function<void()> *body = nullptr;
static unsigned char *dynMem = nullptr; static size_t dynCap = 0;
static std::mutex launchMutex;

unsigned char *dyn_smem() { return dynMem; }

static void yield() { if (!_setjmp(curF->jb)) _longjmp(schedJb, 1); }

static void release_block() { barCount = 0; for (Fiber *f : fibers) if (f->state == WAIT_BLOCK) f->state = READY; }

void sync_block() {
    if (++barCount == (unsigned)nLive) { release_block(); return; }
    curF->state = WAIT_BLOCK; yield();
}

static void finish_collective(Warp &w, unsigned group, unsigned wid) {
    unsigned ballot = 0; int64_t mx = INT64_MIN; uint64_t sum = 0;
    for (int l = 0; l < 32; l++) if ((group >> l) & 1u) { if (w.val[l]) ballot |= 1u << l; if ((int64_t)w.val[l] > mx) mx = (int64_t)w.val[l]; sum += w.val[l]; }
    for (int l = 0; l < 32; l++) if ((group >> l) & 1u) {
        const int wd = w.width[l], lo = l & ~(wd - 1), hi = lo + wd - 1; int src = l; uint64_t r = 0;
        switch (w.op[l]) {
        case SHFL_IDX: src = lo | (w.param[l] & (wd - 1)); r = w.val[src]; break;
        case SHFL_UP: src = l - w.param[l]; r = w.val[src < lo ? l : src]; break;
        case SHFL_DOWN: src = l + w.param[l]; r = w.val[src > hi ? l : src]; break;
        case SHFL_XOR: src = l ^ w.param[l]; r = w.val[(src < lo || src > hi) ? l : src]; break;
        case BALLOT: r = ballot; break;
        case REDUCE_MAX: r = (uint64_t)mx; break;
        case REDUCE_ADD: r = sum; break;
        case SYNC: r = 0; break;
        }
        Fiber *f = fibers[wid * 32 + l];
        f->result = r;
        if (f->state == WAIT_WARP) f->state = READY;
    }
    w.arrived &= ~group;
}

uint64_t warp_collective(unsigned mask, Op op, uint64_t value, int param, int width) {
    const unsigned tid = curF->th.tid.x, lane = tid & 31u, wid = tid >> 5;
    Warp &w = warps[wid];
    w.val[lane] = value; w.param[lane] = param; w.width[lane] = width; w.op[lane] = op;
    w.arrived |= 1u << lane;
    const unsigned group = mask & ~w.exited;            // lanes that have left the kernel cannot arrive any more
    if ((w.arrived & group) == group) { finish_collective(w, group, wid); return curF->result; }
    curF->state = WAIT_WARP; yield();
    return curF->result;
}

static void entry() {
    (*body)();
    curF->state = DONE;
    _longjmp(schedJb, 1);
}

void launch(dim3 grid, dim3 block, size_t dynSmem, const std::function<void()> &fn) {
    std::lock_guard<std::mutex> lock(launchMutex);
    if (block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) { fprintf(stderr, "simt: only 1-D launches are emulated\n"); abort(); }
    const unsigned nt = block.x;
    if (nt == 0 || grid.x == 0) return;
    if (dynSmem + 256 > dynCap) { free(dynMem); dynCap = dynSmem + 256; dynMem = (unsigned char *)aligned_alloc(256, (dynCap + 255) & ~(size_t)255); memset(dynMem, 0, dynCap); }
    while (fibers.size() < ((nt + 31u) & ~31u)) { Fiber *f = new Fiber(); f->stack = (unsigned char *)malloc(STACK); fibers.push_back(f); }
    warps.assign((nt + 31) / 32, Warp());
    body = &fn;
    for (unsigned b = 0; b < grid.x; b++) {
        for (unsigned w = 0; w < warps.size(); w++) { warps[w].arrived = 0; warps[w].exited = 0; for (unsigned l = 0; l < 32; l++) if (w * 32 + l >= nt) warps[w].exited |= 1u << l; }
        for (unsigned t = 0; t < nt; t++) {
            Fiber *f = fibers[t];
            f->th.tid = uint3{t, 0, 0}; f->th.bid = uint3{b, 0, 0}; f->th.bdim = block; f->th.gdim = grid; f->state = READY; f->result = 0; f->started = false;
            getcontext(&f->ctx); f->ctx.uc_stack.ss_sp = f->stack; f->ctx.uc_stack.ss_size = STACK; f->ctx.uc_link = nullptr;
            makecontext(&f->ctx, entry, 0);
        }
        for (size_t t = nt; t < fibers.size(); t++) fibers[t]->state = DONE;
        nLive = (int)nt; barCount = 0;
        while (nLive > 0) {
            bool progressed = false;
            for (unsigned t = 0; t < nt; t++) {
                Fiber *f = fibers[t];
                if (f->state != READY) continue;
                progressed = true;
                curF = f; cur = &f->th;
                if (!_setjmp(schedJb)) { if (f->started) _longjmp(f->jb, 1); f->started = true; setcontext(&f->ctx); }
                if (f->state == DONE) {
                    nLive--;
                    Warp &w = warps[t >> 5]; w.exited |= 1u << (t & 31u);
                    if (barCount && barCount == (unsigned)nLive) release_block();
                    // a warp collective the exited lane was expected at can complete without it
                    const unsigned waiting = w.arrived & ~w.exited;
                    if (waiting && nLive > 0) {
                        bool all = true;
                        for (int l = 0; l < 32 && all; l++) if (!((w.exited >> l) & 1u) && !((w.arrived >> l) & 1u) && fibers[(t & ~31u) + l]->state != DONE) all = false;
                        if (all) finish_collective(w, waiting, t >> 5);
                    }
                }
            }
            if (!progressed) { fprintf(stderr, "simt: deadlock in block %u (%d threads waiting at different synchronisation points)\n", b, nLive); abort(); }
        }
    }
    cur = nullptr; curF = nullptr; body = nullptr;
}

}  // namespace simt
