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

constexpr int GAMES_PER_CTA = 8;                  // one warp per game, 8 games share a CTA
#ifndef HIVE_MIN_CTAS
#define HIVE_MIN_CTAS 3                           // resident CTAs per SM the register budget is set for
#endif
enum Op { OP_RESET = 0, OP_STEP = 1, OP_EVAL = 2, OP_RANDOM = 3, OP_INIT = 4 };   // INIT = first reset, zeroes the counters

struct EnvArgs {
    GameRec* recs;
    uint32_t* legal;       // [n][50]
    int32_t* count;        // [n]
    uint32_t* status;      // [n] turn | winner<<8 | done<<16
    uint16_t* planes;      // [n][56*144] bf16
    const int32_t* actions;
    const uint8_t* mask;
    int32_t* chosen;
    const uint32_t* hop_lines;
    uint64_t seed;
    int n, op, max_turn, auto_reset;
};

template <int G>
__global__ void __launch_bounds__(G * 32, HIVE_MIN_CTAS) hive_env_kernel(EnvArgs a) {
    static_assert(G <= MAX_GAMES_PER_CTA, "queue item encoding holds 4 bits of game slot");
    __shared__ WarpScratch scratch[G];
    __shared__ CtaQueues queues;
    __shared__ uint2 bf16_lut[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x * G + warp;
    WarpScratch& sm = scratch[warp];
    fill_bf16_lut(bf16_lut, tid);
    if (tid < 5) (&queues.n_flood)[tid] = 0;
    __syncthreads();

    // ---------------- per-game prologue: decode the operation, apply the action (warp <-> game)
    bool live = g < a.n;
    int cell = HAND, level = 0, turn = 1, winner = 0;
    uint32_t episode = 0, steps = 0;
    bool push = false;
    GameRec* rec = a.recs + (live ? g : 0);
    if (live) {
        if (lane < N_PIECE) { cell = rec->cell[lane]; level = rec->level[lane]; }
        // header words: [11] = turn|winner|done|flags, [12] episode, [13] steps, [14] n_legal
        const uint32_t* hw = reinterpret_cast<const uint32_t*>(rec);
        const uint32_t h11 = hw[11];
        turn = h11 & 0xFF; winner = (h11 >> 8) & 0xFF;
        const int done = (h11 >> 16) & 0xFF;
        episode = hw[12]; steps = hw[13];
        const uint32_t n_legal_prev = hw[14];
        if (lane < 20) reinterpret_cast<uint4*>(&sm.hist[0][0][0][0])[lane] = reinterpret_cast<const uint4*>(rec->hist)[lane];

        bool do_reset = false;
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
        } else {   // OP_RANDOM
            if (done || turn >= a.max_turn) {
                if (!a.auto_reset) live = false; else do_reset = true;
            } else if (n_legal_prev == 0) {
                action = -1;
            } else {
                const uint64_t gid = (uint64_t)g + (uint64_t)a.n * episode;
                const uint64_t x = splitmix64(a.seed ^ (gid << 32) ^ (uint64_t)turn);
                action = select_kth_action(a.legal + (size_t)g * LEGAL_WORDS, lane, (int)(x % n_legal_prev));
            }
            if (a.chosen && lane == 0) a.chosen[g] = (do_reset || !live) ? HIVE_NOOP : action;
        }
        __syncwarp();
        if (live) {
            if (do_reset) {                                     // GamePlay.new_game, env_hive.py:61-97
                cell = HAND; level = 0; turn = 1; winner = 0; episode++;
                uint32_t* hz = &sm.hist[0][0][0][0];
                for (int i = lane; i < 80; i += 32) hz[i] = 0;
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
        }
    }

    // ---------------- phase A (warp <-> game)
    EvalResult r;
    r.n_legal = 0; r.done = 0; r.winner = 0;
    if (live) r = eval_phase_a(sm, queues, warp, lane, cell, level, turn, winner);
    __syncthreads();

    // ---------------- phase B (thread <-> queued piece task; floods, then moves grouped by piece type)
    eval_phase_b(scratch, queues, tid, G * 32, a.hop_lines);
    __syncthreads();

    // ---------------- phase C (warp <-> game) + write back
    if (!live) return;
    r.n_legal = eval_phase_c(sm, lane, cell, turn, push);
    if (lane < N_PIECE) { rec->cell[lane] = (uint8_t)cell; rec->level[lane] = (uint8_t)level; }
    if (lane == 0) {
        uint32_t* w = reinterpret_cast<uint32_t*>(rec);
        w[11] = (uint32_t)turn | ((uint32_t)r.winner << 8) | ((uint32_t)r.done << 16);
        w[12] = episode; w[13] = steps; w[14] = (uint32_t)r.n_legal;
        a.count[g] = r.n_legal;
        a.status[g] = w[11];
    }
    if (lane < 20) reinterpret_cast<uint4*>(rec->hist)[lane] = reinterpret_cast<const uint4*>(&sm.hist[0][0][0][0])[lane];
    if (lane < 25) reinterpret_cast<uint2*>(a.legal + (size_t)g * LEGAL_WORDS)[lane] = reinterpret_cast<const uint2*>(sm.legal)[lane];
    store_planes_bf16(sm, bf16_lut, lane, turn, a.planes + (size_t)g * HIVE_PLANES_ELEMS);
}

}  // namespace hive
