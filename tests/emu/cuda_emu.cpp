// cuda_emu.cpp -- scheduler of the lock-step warp emulator (see cuda_emu.h). TEST INFRASTRUCTURE.
#include "cuda_emu.h"

namespace emu {

Warp* W = nullptr;
thread_local dim3emu g_threadIdx, g_blockIdx;
static char g_errbuf[256];
const char* last_error() { return g_errbuf; }

struct Tramp { LaneFn fn; void* arg; };
static Tramp g_tramp;

static void lane_entry() {
    Warp* w = W;
    g_tramp.fn(g_tramp.arg);
    int l = w->cur;
    w->done[l] = true;
    w->kind[l] = K_DONE;
    swapcontext(&w->ctx[l], &w->sched);
}

static inline uint64_t next_rand(uint64_t& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

int run_warp(LaneFn fn, void* arg, int block, int warp, uint64_t seed) {
    static const size_t STACK = 256 * 1024;
    Warp* w = (Warp*)calloc(1, sizeof(Warp));
    W = w;
    w->rng = seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    g_tramp.fn = fn; g_tramp.arg = arg;
    g_errbuf[0] = 0;
    for (int l = 0; l < 32; l++) {
        w->stack[l] = (char*)malloc(STACK);
        getcontext(&w->ctx[l]);
        w->ctx[l].uc_stack.ss_sp = w->stack[l];
        w->ctx[l].uc_stack.ss_size = STACK;
        w->ctx[l].uc_link = &w->sched;
        makecontext(&w->ctx[l], lane_entry, 0);
        w->kind[l] = K_NONE;
    }
    int rc = 0;
    for (;;) {
        // one round: resume every live lane once, in random order
        int order[32];
        for (int i = 0; i < 32; i++) order[i] = i;
        for (int i = 31; i > 0; i--) { int j = (int)(next_rand(w->rng) % (uint64_t)(i + 1)); int t = order[i]; order[i] = order[j]; order[j] = t; }
        int live = 0;
        for (int i = 0; i < 32; i++) {
            int l = order[i];
            if (w->done[l]) continue;
            w->cur = l;
            g_threadIdx.x = (unsigned)(warp * 32 + l); g_threadIdx.y = g_threadIdx.z = 0;
            g_blockIdx.x = (unsigned)block; g_blockIdx.y = g_blockIdx.z = 0;
            swapcontext(&w->sched, &w->ctx[l]);
            if (!w->done[l]) live++;
        }
        if (live == 0) break;
        // all live lanes are parked at a collective: they must agree, and nobody may have exited
        int kind = -1;
        bool any_done = false;
        for (int l = 0; l < 32; l++) {
            if (w->done[l]) { any_done = true; continue; }
            if (kind < 0) kind = w->kind[l];
            else if (kind != w->kind[l]) { snprintf(g_errbuf, sizeof g_errbuf, "divergent collectives: kind %d vs %d", kind, w->kind[l]); rc = -1; }
        }
        if (any_done) { snprintf(g_errbuf, sizeof g_errbuf, "full-mask collective (kind %d) reached after some lanes exited", kind); rc = -1; }
        if (rc) break;
        w->collectives++;
        uint32_t bal = 0, red = 0;
        for (int l = 0; l < 32; l++) { if (w->in[l]) bal |= 1u << l; red |= w->in[l]; }
        for (int l = 0; l < 32; l++) {
            switch (kind) {
                case K_SHFL: w->out[l] = w->in[w->arg[l] & 31]; break;
                case K_SHFL_XOR: w->out[l] = w->in[(l ^ w->arg[l]) & 31]; break;
                case K_SHFL_UP: w->out[l] = (l >= (int)w->arg[l]) ? w->in[l - w->arg[l]] : w->in[l]; break;
                case K_BALLOT: w->out[l] = bal; break;
                case K_REDOR: w->out[l] = red; break;
                case K_MATCH: { uint32_t m = 0; for (int j = 0; j < 32; j++) if (w->in[j] == w->in[l]) m |= 1u << j; w->out[l] = m; break; }
                default: w->out[l] = 0; break;
            }
        }
    }
    for (int l = 0; l < 32; l++) free(w->stack[l]);
    free(w);
    W = nullptr;
    return rc;
}

}  // namespace emu
