// cuda_emu.cpp -- scheduler of the lock-step SIMT emulator (see cuda_emu.h). TEST INFRASTRUCTURE.
#include "cuda_emu.h"

namespace emu {

Block* W = nullptr;
thread_local dim3emu g_threadIdx, g_blockIdx, g_blockDim, g_gridDim = {1, 1, 1};
static char g_errbuf[256];
const char* last_error() { return g_errbuf; }

struct Tramp { LaneFn fn; void* arg; };
static Tramp g_tramp;

static void lane_entry() {
    Block* w = W;
    g_tramp.fn(g_tramp.arg);
    int l = w->cur;
    w->done[l] = true;
    w->kind[l] = K_DONE;
    swapcontext(&w->ctx[l], &w->sched);
}

static inline uint64_t next_rand(uint64_t& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

// resolve a warp whose live lanes are all parked at the same warp-level collective
static void resolve_warp(Block* w, int base, int kind) {
    uint32_t bal = 0, red = 0;
    for (int l = 0; l < 32; l++) { if (w->in[base + l]) bal |= 1u << l; red |= (uint32_t)w->in[base + l]; }
    for (int l = 0; l < 32; l++) {
        int i = base + l;
        switch (kind) {
            case K_SHFL: w->out[i] = w->in[base + (w->arg[i] & 31)]; break;
            case K_SHFL_XOR: w->out[i] = w->in[base + ((l ^ w->arg[i]) & 31)]; break;
            case K_SHFL_UP: w->out[i] = (l >= (int)w->arg[i]) ? w->in[i - w->arg[i]] : w->in[i]; break;
            case K_BALLOT: w->out[i] = bal; break;
            case K_REDOR: w->out[i] = red; break;
            case K_MATCH: { uint32_t m = 0; for (int j = 0; j < 32; j++) if (w->in[base + j] == w->in[i]) m |= 1u << j; w->out[i] = m; break; }
            default: w->out[i] = 0; break;
        }
        w->kind[i] = K_NONE;
    }
}

int run_block(LaneFn fn, void* arg, int block, int nthreads, uint64_t seed) {
    static const size_t STACK = 256 * 1024;
    if (nthreads <= 0 || nthreads > MAX_LANES || nthreads % 32) { snprintf(g_errbuf, sizeof g_errbuf, "bad block size %d", nthreads); return -1; }
    Block* w = (Block*)calloc(1, sizeof(Block));
    W = w;
    w->nlanes = nthreads;
    w->rng = seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    g_tramp.fn = fn; g_tramp.arg = arg;
    g_errbuf[0] = 0;
    for (int l = 0; l < nthreads; l++) {
        w->stack[l] = (char*)malloc(STACK);
        getcontext(&w->ctx[l]);
        w->ctx[l].uc_stack.ss_sp = w->stack[l];
        w->ctx[l].uc_stack.ss_size = STACK;
        w->ctx[l].uc_link = &w->sched;
        makecontext(&w->ctx[l], lane_entry, 0);
        w->kind[l] = K_NONE;
    }
    int rc = 0;
    int* order = (int*)malloc(sizeof(int) * nthreads);
    for (;;) {
        // one round: resume every runnable lane once, in random order
        for (int i = 0; i < nthreads; i++) order[i] = i;
        for (int i = nthreads - 1; i > 0; i--) { int j = (int)(next_rand(w->rng) % (uint64_t)(i + 1)); int t = order[i]; order[i] = order[j]; order[j] = t; }
        int ran = 0, live = 0;
        for (int i = 0; i < nthreads; i++) {
            int l = order[i];
            if (w->done[l] || w->kind[l] != K_NONE) continue;
            w->cur = l;
            g_threadIdx.x = (unsigned)l; g_threadIdx.y = g_threadIdx.z = 0;
            g_blockIdx.x = (unsigned)block; g_blockIdx.y = g_blockIdx.z = 0;
            g_blockDim.x = (unsigned)nthreads; g_blockDim.y = g_blockDim.z = 1;
            swapcontext(&w->sched, &w->ctx[l]);
            ran++;
        }
        for (int l = 0; l < nthreads; l++) if (!w->done[l]) live++;
        if (live == 0) break;
        // resolve warp collectives
        int progressed = 0;
        for (int base = 0; base < nthreads && !rc; base += 32) {
            int kind = -1; bool same = true, any_done = false, any_live = false, all_parked = true;
            for (int l = base; l < base + 32; l++) {
                if (w->done[l]) { any_done = true; continue; }
                any_live = true;
                if (w->kind[l] == K_NONE) { all_parked = false; continue; }
                if (kind < 0) kind = w->kind[l]; else if (kind != w->kind[l]) same = false;
            }
            if (!any_live || !all_parked || kind == K_BAR) { if (any_live && all_parked && kind == K_BAR && !same) { snprintf(g_errbuf, sizeof g_errbuf, "warp %d: lanes split between __syncthreads and a warp collective", base / 32); rc = -1; } continue; }
            if (!same) { snprintf(g_errbuf, sizeof g_errbuf, "warp %d: divergent collectives", base / 32); rc = -1; break; }
            if (any_done) { snprintf(g_errbuf, sizeof g_errbuf, "warp %d: full-mask collective (kind %d) after some lanes exited", base / 32, kind); rc = -1; break; }
            resolve_warp(w, base, kind);
            w->collectives++;
            progressed++;
        }
        if (rc) break;
        // block barrier: every live lane parked at K_BAR
        bool all_bar = true;
        for (int l = 0; l < nthreads; l++) if (!w->done[l] && w->kind[l] != K_BAR) { all_bar = false; break; }
        if (all_bar) { for (int l = 0; l < nthreads; l++) if (!w->done[l]) w->kind[l] = K_NONE; progressed++; }
        if (!progressed && !ran) { snprintf(g_errbuf, sizeof g_errbuf, "deadlock: no lane can make progress"); rc = -1; break; }
    }
    free(order);
    for (int l = 0; l < nthreads; l++) free(w->stack[l]);
    free(w);
    W = nullptr;
    return rc;
}

}  // namespace emu
