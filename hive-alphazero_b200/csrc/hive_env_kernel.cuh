// hive_env_kernel.cuh -- the environment step kernel (included by hive_env.cu and, verbatim, by the
// CPU lock-step SIMT emulator under tests/emu that checks it against the oracle without a GPU).
#pragma once
#include "hive_core.cuh"

#ifndef HIVE_NOOP
#define HIVE_NOOP (-2)
#endif
#ifndef HIVE_RESET
#define HIVE_RESET (-3)
#endif
#ifndef HIVE_PLANES_ELEMS
#define HIVE_PLANES_ELEMS (56 * 144)
#endif

namespace hive {

#ifndef HIVE_ENCODE_WARPS
#define HIVE_ENCODE_WARPS 8                        // games per CTA of the encode kernel
#endif
enum Op { OP_RESET = 0, OP_STEP = 1, OP_EVAL = 2, OP_RANDOM = 3, OP_INIT = 4 };   // INIT = first reset, zeroes the counters

struct EnvArgs {
    GameRec* recs;
    uint32_t* legal;       // [n][50]
    int32_t* count;        // [n]
    uint32_t* status;      // [n] turn | winner<<8 | done<<16
    uint16_t* planes;      // [n][56*144] bf16
    GameScratch* scratch;  // [n] kernel-to-kernel intermediates (L2 resident)
    uint32_t* bits;        // [n][BITS_WORDS] bit planes, encode kernel -> plane-store kernel (this step's buffer of two)
    BatchQueues bq;        // batch-wide work queues
    const int32_t* actions;
    const uint8_t* mask;
    int32_t* chosen;
    const uint32_t* hop_lines;   // GEO_* tables (hive_core.cuh): hop lines, neighbour ranks, neighbour cells
    uint64_t seed;
    int n, op, max_turn, auto_reset;
    int g_offset, n_total;   // this launch covers games [g_offset, g_offset+n) of a batch of n_total (pointers are pre-offset)
};

// ---- optional per-CTA timeline (builds with -DHIVE_TRACE only; profiles/trace_probe.py reads it)
#ifdef HIVE_TRACE
struct TraceRec { unsigned long long t0, t1; uint32_t sm, kernel, g_offset, block; };
__device__ TraceRec* g_trace;
__device__ unsigned int g_trace_n, g_trace_cap;
__device__ __forceinline__ unsigned long long trace_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
struct TraceScope {
    TraceRec* r;
    __device__ TraceScope(int kernel, int g_offset) : r(nullptr) {
        if (threadIdx.x == 0 && g_trace) {
            const unsigned i = atomicAdd(&g_trace_n, 1u);
            if (i < g_trace_cap) {
                r = g_trace + i;
                uint32_t sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
                r->sm = sm; r->kernel = kernel; r->g_offset = g_offset; r->block = blockIdx.x; r->t1 = 0; r->t0 = trace_now();
            }
        }
    }
    __device__ ~TraceScope() { if (r) r->t1 = trace_now(); }
};
#define HIVE_TRACE_SCOPE(k, a) TraceScope trace_scope_(k, (a).g_offset)
#else
#define HIVE_TRACE_SCOPE(k, a)
#endif

// Programmatic dependent launch (hive_env.cu launches the kernels of a slice's chain with
// cudaLaunchAttributeProgrammaticStreamSerialization): a kernel's CTAs may become resident while the kernel before it in
// the stream is still running; they wait here until that grid has completed and its writes are visible, and only then
// let their own dependent start launching (so at most one kernel of a chain is pre-launched).  Without the launch
// attribute both instructions return at once.
__device__ __forceinline__ void chain_wait_then_release() {
#ifndef HIVE_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// ---- kernel 1: decode the operation, apply the action, analyse the new position (warp <-> game)
__global__ void __launch_bounds__(GROUP * 32, 64 / GROUP) hive_analyse_kernel(EnvArgs a) {
    __shared__ GroupQueues q;
    __shared__ uint32_t occ_s[GROUP][8];
    chain_wait_then_release();
    HIVE_TRACE_SCOPE(0, a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x * GROUP + warp;
    if (tid < 5) (&q.n_flood)[tid] = 0;
    __syncthreads();

    bool live = g < a.n;
    if (live) {
        GameRec* rec = a.recs + g;
        int cell = HAND, level = 0;
        if (lane < N_PIECE) { cell = rec->cell[lane]; level = rec->level[lane]; }
        // header words: [11] = turn|winner|done|flags, [12] episode, [13] steps, [14] n_legal
        uint32_t* hw = reinterpret_cast<uint32_t*>(rec);
        const uint32_t h11 = hw[11];
        int turn = h11 & 0xFF, winner = (h11 >> 8) & 0xFF;
        const int done = (h11 >> 16) & 0xFF;
        uint32_t episode = hw[12], steps = hw[13];
        const uint32_t n_legal_prev = hw[14];

        bool do_reset = false, push = false;
        int action = HIVE_NOOP;
        if (a.op == OP_RESET) {
            if (a.mask && !a.mask[g]) live = false; else do_reset = true;
        } else if (a.op == OP_INIT) {
            do_reset = true; episode = 0xFFFFFFFFu; steps = 0;      // first episode of the slot is number 0
        } else if (a.op == OP_STEP) {
            action = a.actions[g];
            if (action == HIVE_NOOP) live = false;
            if (action == HIVE_RESET) do_reset = true;
        } else if (a.op == OP_EVAL) {
            if (a.mask && !a.mask[g]) live = false;
            push = (h11 >> 24) & 1u;          // "push history when this position is evaluated" (set by the search)
        } else {   // OP_RANDOM
            if (done || turn >= a.max_turn) {
                if (!a.auto_reset) live = false; else do_reset = true;
            } else if (n_legal_prev == 0) {
                action = -1;
            } else {
                const uint64_t gid = (uint64_t)(g + a.g_offset) + (uint64_t)a.n_total * episode;
                const uint64_t x = splitmix64(a.seed ^ (gid << 32) ^ (uint64_t)turn);
                action = select_kth_action(a.legal + (size_t)g * LEGAL_WORDS, lane, (int)(x % n_legal_prev));
            }
            if (a.chosen && lane == 0) a.chosen[g] = (do_reset || !live) ? HIVE_NOOP : action;
        }
        __syncwarp();
        if (live) {
            if (do_reset) {                                     // GamePlay.new_game, env_hive.py:61-97
                cell = HAND; level = 0; turn = 1; winner = 0; episode++;
                if (lane < 20) reinterpret_cast<uint4*>(rec->hist)[lane] = make_uint4(0u, 0u, 0u, 0u);
                push = true;                                    // add_history starts True (env_hive.py:51)
            } else if (action >= 0) {                           // env_hive.py:105-148
                const int side = (turn & 1) ? 0 : 1;
                const int k = action % 11, end = action / 11, p = side * 11 + k;
                const int h_end = __popc(__ballot_sync(FULL, cell == end));
                if (lane == p) { cell = end; level = h_end; }   // level = len(end_tile.pieces) before the move
                turn++; steps++; push = true;
            } else if (action == -1) {                          // pass, env_hive.py:100-103
                turn++; steps++;
            }
            __syncwarp();
            if (lane < N_PIECE) { rec->cell[lane] = (uint8_t)cell; rec->level[lane] = (uint8_t)level; }
            if (lane == 0) { hw[12] = episode; hw[13] = steps; }
            eval_analyse(a.scratch[g], q, occ_s[warp], warp, lane, cell, level, turn, push, winner, a.hop_lines);
        } else if (lane == 0) {
            a.scratch[g].head[2] = 0;                           // not evaluated in this launch
        }
    }
    __syncthreads();
    // publish this group's work: reserve a slice of every batch-wide queue (one atomic per class and CTA)
    if (tid < 5) { const uint32_t c = (&q.n_flood)[tid]; q.base[tid] = c ? atomicAdd(a.bq.counters + tid, c) : 0u; }
    __syncthreads();
    {
        const uint32_t g0 = (uint32_t)blockIdx.x * GROUP;
        const int nf = (int)q.n_flood;
        for (int i = tid; i < nf; i += GROUP * 32) {
            const uint32_t it = q.flood[i];
            a.bq.flood[q.base[0] + i] = ((g0 + (it & 15u)) << 6) | (((it >> 4) & 31u) << 1) | ((it >> 9) & 1u);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int nm = (int)q.n_mv[c];
            for (int i = tid; i < nm; i += GROUP * 32) {
                const uint32_t it = q.mv[c][i];
                a.bq.mv[c][q.base[1 + c] + i] = ((g0 + (it & 15u)) << 6) | (((it >> 4) & 31u) << 1) | 1u;
            }
        }
    }
}

// ---- kernel 2: one-hive floods over the batch-wide flood queue (thread <-> queued piece)
constexpr int SEARCH_THREADS = 128;
__global__ void __launch_bounds__(SEARCH_THREADS) hive_flood_kernel(EnvArgs a) {
    chain_wait_then_release();
    HIVE_TRACE_SCOPE(1, a);
    const int lane = threadIdx.x & 31;
    const int nf = (int)a.bq.counters[0];
    const int stride = gridDim.x * SEARCH_THREADS;
    for (int t0 = (blockIdx.x * SEARCH_THREADS + threadIdx.x) - lane; t0 < nf; t0 += stride) {   // warp-uniform loop
        const int t = t0 + lane;
        bool push = false;
        uint32_t item = 0;
        int cls = 0;
        if (t < nf) {
            item = a.bq.flood[t];
            const int p = (item >> 1) & 31;
            GameScratch& gs = a.scratch[item >> 6];
            const bool pinned = eval_flood(gs, p);
            if (pinned) atomicOr(&gs.head[1], 1u << p);
            else if (item & 1u) { push = true; cls = move_class(piece_type_of(p >= 11 ? p - 11 : p)); }
        }
        // survivors that want a move search join the move queues (one atomic per class and warp)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const unsigned m = __ballot_sync(FULL, push && cls == c);
            if (m) {
                uint32_t base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(a.bq.counters + 1 + c, (uint32_t)__popc(m));
                base = __shfl_sync(FULL, base, leader);
                if (push && cls == c) a.bq.mv[c][base + __popc(m & ((1u << lane) - 1u))] = item;
            }
        }
    }
}

// ---- kernel 3: move searches, warps homogeneous in piece type (thread <-> queued piece)
__global__ void __launch_bounds__(SEARCH_THREADS) hive_moves_kernel(EnvArgs a) {
    chain_wait_then_release();
    HIVE_TRACE_SCOPE(2, a);
    // move classes start at warp boundaries so that warps stay homogeneous
    const int n0 = (int)a.bq.counters[1], n1 = (int)a.bq.counters[2], n2 = (int)a.bq.counters[3], n3 = (int)a.bq.counters[4];
    const int s1 = (n0 + 31) & ~31, s2 = s1 + ((n1 + 31) & ~31), s3 = s2 + ((n2 + 31) & ~31), total = s3 + n3;
    const int stride = gridDim.x * SEARCH_THREADS;
    for (int t = blockIdx.x * SEARCH_THREADS + threadIdx.x; t < total; t += stride) {
        int cls, idx, cnt;
        if (t < s1) { cls = 0; idx = t; cnt = n0; }
        else if (t < s2) { cls = 1; idx = t - s1; cnt = n1; }
        else if (t < s3) { cls = 2; idx = t - s2; cnt = n2; }
        else { cls = 3; idx = t - s3; cnt = n3; }
        if (idx < cnt) {
            const uint32_t item = a.bq.mv[cls][idx];
            eval_moves(a.scratch[item >> 6], (item >> 1) & 31, a.hop_lines);
        }
    }
}

// ---- kernel 4: legal mask, bit planes, history, terminal test, small outputs (warp <-> game).  Everything the
// next step needs is written here; the 16 KB of bf16 planes per game are left to kernel 5, which runs on its own
// stream beside the next step's kernels.
__global__ void __launch_bounds__(HIVE_ENCODE_WARPS * 32, 6) hive_encode_kernel(EnvArgs a) {
    __shared__ WarpScratch scratch[HIVE_ENCODE_WARPS];
    chain_wait_then_release();
    HIVE_TRACE_SCOPE(3, a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x * HIVE_ENCODE_WARPS + warp;
    if (blockIdx.x == 0 && tid < 8) a.bq.counters[tid] = 0;     // the queues are consumed: reset for the next step
    if (g >= a.n) return;
    // every global read of this game is issued before the first use (also for games this launch skips)
    GameRec* rec = a.recs + g;
    const EncodeIn in = encode_fetch(a.scratch[g], lane);
    const uint4 h4 = reinterpret_cast<const uint4*>(rec->hist)[lane < 20 ? lane : 0];
    WarpScratch& sm = scratch[warp];
    uint32_t* bits = a.bits + (size_t)g * BITS_WORDS;
    if (lane < 20) reinterpret_cast<uint4*>(&sm.hist[0][0][0][0])[lane] = h4;    // consumed before the exit test: keeps the load up here
    if (!(in.head.z & 1u)) {
        if (lane == 0) bits[BITS_LIVE] = 0u;                    // the plane store leaves this game's planes alone
        return;
    }
    __syncwarp();
    const EvalResult r = eval_encode(sm, in, lane, a.hop_lines);
    const int turn = in.head.x & 0xFF;
    if (lane == 0) {
        uint32_t* w = reinterpret_cast<uint32_t*>(rec);
        const uint32_t st = (uint32_t)turn | ((uint32_t)r.winner << 8) | ((uint32_t)r.done << 16);
        w[11] = st; w[14] = (uint32_t)r.n_legal;
        a.count[g] = r.n_legal;
        a.status[g] = st;
        sm.planes[31][0] = (uint32_t)turn; sm.planes[31][1] = 1u;
    }
    if (lane < 20) reinterpret_cast<uint4*>(rec->hist)[lane] = reinterpret_cast<const uint4*>(&sm.hist[0][0][0][0])[lane];
    if (lane < 25) reinterpret_cast<uint2*>(a.legal + (size_t)g * LEGAL_WORDS)[lane] = reinterpret_cast<const uint2*>(sm.legal)[lane];
    __syncwarp();
    {   // 1120 B of bit planes -> L2 (70 x 16 B)
        const uint4* src = reinterpret_cast<const uint4*>(&sm.planes[0][0]);
        uint4* dst = reinterpret_cast<uint4*>(bits);
#pragma unroll
        for (int i = 0; i < 3; i++) { const int t = lane + 32 * i; if (t < BITS_WORDS / 4) dst[t] = src[t]; }
    }
}

// ---- kernel 5: bit planes -> bf16 CHW planes [56][144] per game, through the TMA (persistent: the grid is capped in
// hive_env.cu so that every CTA is resident at once and other kernels are placed beside it; a warp walks over games)
#ifndef HIVE_STORE_WARPS
#define HIVE_STORE_WARPS 4
#endif
constexpr int STORE_STAGE_BYTES = HIVE_STORE_WARPS * STAGE_BUFS * STAGE_BYTES;     // dynamic shared memory
#ifndef HIVE_STORE_MIN_CTAS
#define HIVE_STORE_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(HIVE_STORE_WARPS * 32, HIVE_STORE_MIN_CTAS) hive_planes_kernel(EnvArgs a) {
    __shared__ uint4 bf16_lut[256];
    __shared__ __align__(16) uint32_t planes_s[HIVE_STORE_WARPS][BITS_WORDS];
#ifdef HIVE_EMU
    __shared__ uint4 stage_ring[STORE_STAGE_BYTES / 16];
#else
    extern __shared__ __align__(128) uint4 stage_ring[];
#endif
    HIVE_TRACE_SCOPE(4, a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int t = tid; t < 256; t += HIVE_STORE_WARPS * 32) fill_bf16_lut(bf16_lut, t);
    __shared__ uint32_t lut_addr;
#ifndef HIVE_EMU
    if (tid == 0) lut_addr = (uint32_t)__cvta_generic_to_shared(bf16_lut);
#endif
    __syncthreads();
    const uint32_t lut_s = *reinterpret_cast<volatile uint32_t*>(&lut_addr);
    constexpr int NV = BITS_WORDS / 4;                           // 70 uint4 per game
    const int g_first = blockIdx.x * HIVE_STORE_WARPS + warp, g_stride = gridDim.x * HIVE_STORE_WARPS;
    uint4 v[3];
    auto fetch = [&](int g) {
        const uint4* src = reinterpret_cast<const uint4*>(a.bits + (size_t)g * BITS_WORDS);
#pragma unroll
        for (int i = 0; i < 3; i++) { const int t = lane + 32 * i; v[i] = src[t < NV ? t : 0]; }
    };
    if (g_first < a.n) fetch(g_first);
    uint32_t* mine = planes_s[warp];
    for (int g = g_first; g < a.n; g += g_stride) {
        __syncwarp();                                           // the previous game's planes have been expanded
#pragma unroll
        for (int i = 0; i < 3; i++) { const int t = lane + 32 * i; if (t < NV) reinterpret_cast<uint4*>(mine)[t] = v[i]; }
        __syncwarp();
        const uint32_t live = mine[BITS_LIVE], turn = mine[BITS_TURN];
        if (g + g_stride < a.n) fetch(g + g_stride);            // next game's bit planes arrive during the expansion
        if (live)
            store_planes_bulk(reinterpret_cast<const uint8_t*>(mine), bf16_lut, lut_s, stage_ring + warp * (STAGE_BUFS * STAGE_CHUNKS), lane,
                              (int)turn, a.planes + (size_t)g * HIVE_PLANES_ELEMS);
    }
    if (lane == 0) bulk_wait_read<0>();                         // shared memory must outlive the copy engine's reads
}

}  // namespace hive
