// hive_mcts_kernel.cuh -- PUCT search kernels, one warp per tree (sm_100a).
//
// Bit-exact restatement of the reference's sequential search (woker/solo_play.py:110-374,
// HivePlayer with none_queue=False): transposition table keyed by the position (== state_key,
// env_hive.py:150-168: no turn, no history), first visit evaluates and returns, lazy prior
// normalisation in float32, PUCT in the reference's mixed float32/float64 order, fresh Dirichlet
// row per root visit, virtual loss applied on the way down and undone on the way up, draw / turn
// limit sentinel 5.
//
// A search advances in waves: `mcts_descend_kernel` runs every tree until its simulation reaches
// a position that is not in the tree (terminal simulations are finished in place), the leaf
// positions are evaluated by the environment kernels (legal mask + planes) and the network, and
// `mcts_expand_kernel` creates the node and backs the value up.
#pragma once
#include "hive_core.cuh"

#ifndef HIVE_PLANES_ELEMS
#define HIVE_PLANES_ELEMS (56 * 144)
#endif

namespace hive {

constexpr int MCTS_MAX_DEPTH = 64;
constexpr int MCTS_WARPS = 4;
enum TreeState { TREE_IDLE = 0, TREE_NEED_EVAL = 1, TREE_DONE = 2 };

struct __align__(16) MctsNode {
    uint8_t cell[N_PIECE];
    uint8_t level[N_PIECE];
    uint8_t player;            // 0 white to move, 1 black to move
    uint8_t normalized;        // priors already divided by their float32 sum (solo_play.py:304-313)
    uint8_t pad[2];
    int32_t sum_n;
    int32_t n_edges;           // edges stored (>= 1: a position without legal actions holds the pass edge -1)
    int32_t edge_off;
    int32_t n_legal;
};
static_assert(sizeof(MctsNode) == 64, "MctsNode layout");

struct __align__(16) MctsTree {
    int32_t n_nodes, edges_used, sims_done, depth;
    int32_t root_selects, state, error, last_real;
    int32_t path_node[MCTS_MAX_DEPTH];
    int32_t path_edge[MCTS_MAX_DEPTH];
};

struct MctsArgs {
    int n, sims, max_turn, node_cap, edge_cap, ht_size;     // ht_size: power of two
    int noise_rows, noise_cols;                             // recorded root noise (0 = generate on device)
    uint64_t noise_seed;
    const GameRec* root_recs;                               // the real games
    GameRec* sim_recs;                                      // working positions (a hive_env batch)
    uint32_t* sim_legal;                                    // [n][50] legal mask of the evaluated leaf
    int32_t* sim_count;                                     // [n]
    uint16_t* sim_planes;                                   // [n][56*144] planes of the evaluated leaf
    const uint32_t* root_legal; const int32_t* root_count; const uint16_t* root_planes;   // the real games' outputs
    const uint32_t* root_shadow; uint32_t* sim_shadow;      // delta plane store: bit images of root_planes / sim_planes (or null)
    uint8_t* env_mask;                                      // [n] leaf positions the environment kernels must evaluate
    const float* leaf_p;                                    // [n][1584] network policy of the leaf
    const double* leaf_v;                                   // [n] network value of the leaf
    uint8_t* need_eval;                                     // [n]
    const uint8_t* tree_mask;                               // [n] or null: trees taking part in this search
    int32_t* pending;                                       // [1] number of trees waiting for an evaluation
    MctsTree* trees;
    MctsNode* nodes;                                        // [n][node_cap]
    int32_t* htab;                                          // [n][ht_size]
    int16_t* e_action; int32_t* e_n; double* e_w; double* e_q; float* e_p;   // [n][edge_cap]
    const double* noise;                                    // [n][noise_rows][noise_cols]
    double* pi;                                             // [n][1584] output policy
    int32_t* out_action; int32_t* out_sum_n;                // [n]
    uint32_t* search_no;                                    // [n] searches begun in this slot (device noise stream position)
    uint32_t* error_any;                                    // [1] OR of 1 << T.error over all trees since mcts_begin
};

__device__ __forceinline__ uint32_t key_hash(int lane, int cell, int level, int player) {
    uint32_t h = 0;
    if (lane < N_PIECE) {
        h = ((uint32_t)cell | ((uint32_t)level << 8) | ((uint32_t)(lane + 1) << 16)) * 0x9E3779B1u;
        h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h ^= __shfl_xor_sync(FULL, h, o);
    h = (h ^ (uint32_t)player) * 0xC2B2AE3Du;
    return h ^ (h >> 16);
}

// index of the node holding this position, or -1 (all lanes converged)
__device__ __forceinline__ int tree_lookup(const MctsArgs& a, int t, int lane, int cell, int level, int player) {
    const int32_t* ht = a.htab + (size_t)t * a.ht_size;
    const MctsNode* nodes = a.nodes + (size_t)t * a.node_cap;
    uint32_t idx = key_hash(lane, cell, level, player) & (uint32_t)(a.ht_size - 1);
    for (;;) {
        const int id = ht[idx];
        if (id < 0) return -1;
        const MctsNode& nd = nodes[id];
        bool same = true;
        if (lane < N_PIECE) same = nd.cell[lane] == cell && nd.level[lane] == level;
        if (lane == N_PIECE) same = nd.player == player;
        if (__ballot_sync(FULL, same) == FULL) return id;
        idx = (idx + 1) & (uint32_t)(a.ht_size - 1);
    }
}

// Dirichlet(alpha) row generated on the device (used when no recorded noise is supplied):
// Gamma(alpha<1) = Gamma(alpha+1) * U^(1/alpha), Marsaglia-Tsang for the shape alpha+1.
__device__ __forceinline__ double u01(uint64_t& s) {
    s = splitmix64(s);
    return ((double)(s >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double gamma_sample(double alpha, uint64_t seed) {
    uint64_t s = seed;
    const double d = alpha + 1.0 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = d;
    for (int it = 0; it < 64; it++) {
        const double u1 = u01(s), u2 = u01(s);
        const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        const double v0 = 1.0 + c * x;
        if (v0 <= 0.0) continue;
        const double v = v0 * v0 * v0, u = u01(s);
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) { g = d * v; break; }
    }
    return g * pow(u01(s), 1.0 / alpha);
}

// the value a parent books for a child's return r, and what it returns itself (solo_play.py:217-247)
__device__ __forceinline__ void backup_path(const MctsArgs& a, int t, MctsTree& T, int depth, double r, int lane) {
    if (lane != 0) return;
    MctsNode* nodes = a.nodes + (size_t)t * a.node_cap;
    const size_t eb = (size_t)t * a.edge_cap;
    for (int d = depth - 1; d >= 0; d--) {
        MctsNode& nd = nodes[T.path_node[d]];
        const size_t e = eb + nd.edge_off + T.path_edge[d];
        const bool reach_max = (r == 5.0);
        double lv = reach_max ? 1.0 : r;
        lv = -lv;
        nd.sum_n += -1 + 1;
        a.e_n[e] += -1 + 1;
        a.e_w[e] = a.e_w[e] + (1.0 + lv);                  // my_stats.w += virtual_loss + leaf_v
        a.e_q[e] = a.e_w[e] / (double)a.e_n[e];
        r = reach_max ? 5.0 : lv;
    }
}

// A simulation that cannot go on (arena full, path too deep): take its virtual losses back (solo_play.py:205-208
// reversed: N-1, W+1) so that the statistics hold only finished simulations, flag the tree and the batch.
// The search of this tree stops (mcts_descend_kernel marks it TREE_DONE); the host raises on the flag.
__device__ __forceinline__ void abort_simulation(const MctsArgs& a, int t, MctsTree& T, int depth, int code, int lane) {
    if (lane != 0) return;
    MctsNode* nodes = a.nodes + (size_t)t * a.node_cap;
    const size_t eb = (size_t)t * a.edge_cap;
    for (int d = depth - 1; d >= 0; d--) {
        MctsNode& nd = nodes[T.path_node[d]];
        const size_t e = eb + nd.edge_off + T.path_edge[d];
        nd.sum_n -= 1;
        a.e_n[e] -= 1;
        a.e_w[e] = a.e_w[e] + 1.0;
        a.e_q[e] = a.e_n[e] > 0 ? a.e_w[e] / (double)a.e_n[e] : 0.0;
    }
    T.error = code;
    atomicOr(a.error_any, 1u << code);
}

// seed of the device Dirichlet row of root visit `visit` of search number `search` of tree t
__device__ __forceinline__ uint64_t noise_row_seed(uint64_t seed, int t, uint32_t search, int visit) {
    uint64_t s = splitmix64(seed ^ ((uint64_t)(uint32_t)t * 0xD1342543DE82EF95ULL));
    return splitmix64(s ^ ((uint64_t)search << 32) ^ (uint64_t)(uint32_t)visit);
}

__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_reset_kernel(MctsArgs a) {
    const int t = blockIdx.x;
    if (a.tree_mask && !a.tree_mask[t]) return;
    int32_t* ht = a.htab + (size_t)t * a.ht_size;
    for (int i = threadIdx.x; i < a.ht_size; i += blockDim.x) ht[i] = -1;
    for (int i = threadIdx.x; i < (int)(sizeof(MctsTree) / 4); i += blockDim.x) reinterpret_cast<int32_t*>(a.trees + t)[i] = 0;
    if (threadIdx.x == 0) { a.need_eval[t] = 0; a.env_mask[t] = 0; a.search_no[t] += 1u; }
}

__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_descend_kernel(MctsArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * MCTS_WARPS + warp;
    if (t >= a.n) return;
    if (a.tree_mask && !a.tree_mask[t]) return;
    MctsTree& T = a.trees[t];
    if (T.state != TREE_IDLE) return;
    MctsNode* nodes = a.nodes + (size_t)t * a.node_cap;
    const size_t eb = (size_t)t * a.edge_cap;
    const GameRec* root = a.root_recs + t;
    GameRec* sim = a.sim_recs + t;

    for (;;) {
        if (T.sims_done >= a.sims || T.error) { if (lane == 0) T.state = TREE_DONE; return; }
        // ---- a new simulation starts from a copy of the root position (deepcopy(env), solo_play.py:162)
        int cell = HAND, level = 0;
        if (lane < N_PIECE) { cell = root->cell[lane]; level = root->level[lane]; }
        int turn = reinterpret_cast<const uint32_t*>(root)[11] & 0xFF;
        if (lane < 20) reinterpret_cast<uint4*>(sim->hist)[lane] = reinterpret_cast<const uint4*>(root->hist)[lane];
        __syncwarp();
        int depth = 0;
        bool last_real = false;
        double r = 0.0;
        for (;;) {
            const int side = (turn & 1) ? 0 : 1;
            const bool valid = lane < N_PIECE, on_board = valid && cell != HAND;
            const bool own = valid && ((lane >= 11 ? 1 : 0) == side);
            const BB src = on_board ? bb_bit(cell) : bb_zero();
            const BB own_all = warp_or(own ? src : bb_zero());
            const BB opp_all = warp_or((valid && !own) ? src : bb_zero());
            const BB occ = own_all | opp_all;
            // game_is_over (move_checker.py:140-165): a queen with six occupied neighbours
            bool full = false;
            if (on_board && (lane == 0 || lane == 11)) {
                full = true;
#pragma unroll
                for (int i = 0; i < 6; i++) full = full && bb_test(occ, cell_nbr(cell, i));
            }
            const unsigned sur = __ballot_sync(FULL, full);
            const bool ws = sur & 1u, bs = (sur >> 11) & 1u;
            if (ws || bs) {                                        // solo_play.py:169-180
                const int winner = (ws && bs) ? 0 : ws ? 2 : 1;    // 1 white, 2 black, 0 none
                if (winner == 0) r = 5.0;
                else r = ((side == 0) == (winner == 1)) ? 1.0 : -1.0;
                break;
            }
            if (turn >= a.max_turn) { r = 5.0; break; }            // solo_play.py:181-183
            const int id = tree_lookup(a, t, lane, cell, level, side);
            if (id < 0) {
                // ---- not in the tree: hand the position to the evaluator (solo_play.py:188-197)
                if (lane < N_PIECE) { sim->cell[lane] = (uint8_t)cell; sim->level[lane] = (uint8_t)level; }
                if (depth == 0) {
                    // the root itself: its legal set and planes were produced by the step that created the
                    // position (its history entry is already pushed, so it must not be encoded again)
                    const uint4* ps = reinterpret_cast<const uint4*>(a.root_planes + (size_t)t * HIVE_PLANES_ELEMS);
                    uint4* pd = reinterpret_cast<uint4*>(a.sim_planes + (size_t)t * HIVE_PLANES_ELEMS);
                    for (int i = lane; i < HIVE_PLANES_ELEMS * 2 / 16; i += 32) pd[i] = ps[i];
                    if (a.sim_shadow)                              // ... and the bit image the delta plane store keeps of them
                        for (int i = lane; i < BITS_WORDS; i += 32) a.sim_shadow[(size_t)t * BITS_WORDS + i] = a.root_shadow[(size_t)t * BITS_WORDS + i];
                    for (int i = lane; i < LEGAL_WORDS; i += 32) a.sim_legal[(size_t)t * LEGAL_WORDS + i] = a.root_legal[(size_t)t * LEGAL_WORDS + i];
                    if (lane == 0) a.sim_count[t] = a.root_count[t];
                }
                if (lane == 0) {
                    uint32_t* w = reinterpret_cast<uint32_t*>(sim);
                    w[11] = (uint32_t)turn | ((uint32_t)last_real << 24);      // bit 24: push history when evaluated
                    T.depth = depth; T.state = TREE_NEED_EVAL;
                    a.need_eval[t] = 1;
                    a.env_mask[t] = depth > 0 ? 1 : 0;
                    atomicAdd(a.pending, 1);
                }
                return;
            }
            if (depth >= MCTS_MAX_DEPTH) {                         // the path arrays are full: give the simulation up
                abort_simulation(a, t, T, depth, 3, lane);
                __syncwarp();
                if (lane == 0) T.state = TREE_DONE;
                return;
            }
            // ---- the move that led here pushes (own-any, opp-any) of this position (env_hive.py:436-445)
            if (last_real) {
                uint32_t* h = &sim->hist[side][0][0][0];
                uint32_t keep = (lane < 30) ? h[lane] : 0;
                __syncwarp();
                if (lane < 30) h[10 + lane] = keep;
                if (lane < 5) { h[lane] = own_all.w[lane]; h[5 + lane] = opp_all.w[lane]; }
                __syncwarp();
            }
            // ---- select (solo_play.py:294-335)
            MctsNode& nd = nodes[id];
            const int ne = nd.n_edges;
            const size_t e0 = eb + nd.edge_off;
            int best_i = 0;
            if (nd.n_legal > 0) {
                if (!nd.normalized) {
                    float tot = 1e-8f;
                    if (lane == 0) for (int i = 0; i < ne; i++) tot = tot + a.e_p[e0 + i];
                    tot = __shfl_sync(FULL, tot, 0);
                    for (int i = lane; i < ne; i += 32) a.e_p[e0 + i] = a.e_p[e0 + i] / tot;
                    if (lane == 0) nd.normalized = 1;
                    __syncwarp();
                }
                const double xx = sqrt((double)(nd.sum_n + 1));
                const bool is_root = depth == 0;
                const double* noise = nullptr;
                double gsum = 1.0;
                if (is_root && a.noise) {
                    const int row = T.root_selects < a.noise_rows ? T.root_selects : a.noise_rows - 1;
                    noise = a.noise + ((size_t)t * a.noise_rows + row) * a.noise_cols;
                }
                double best = -999.0;
                int bi = 0x7fffffff;
                // device noise: a fresh Dirichlet(0.3) row per root visit (solo_play.py:323) = normalised Gamma(0.3) draws
                // from a counter-based stream keyed by (seed, tree, search number of this slot, root visit, edge).  The
                // draws of the row are parked in this tree's output-policy row (unused until mcts_finalize_kernel
                // rewrites it); every lane reads back only what it wrote itself.
                double* gam = a.pi + (size_t)t * 1584;
                if (is_root && !a.noise) {
                    const uint64_t row_seed = noise_row_seed(a.noise_seed, t, a.search_no[t], T.root_selects);
                    double s = 0.0;
                    for (int i = lane; i < ne; i += 32) {
                        const double g = gamma_sample(0.3, row_seed ^ ((uint64_t)i * 0x9E3779B97F4A7C15ULL));
                        gam[i] = g;
                        s += g;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
                    gsum = s;
                }
                for (int i = lane; i < ne; i += 32) {
                    const float p = a.e_p[e0 + i];
                    const double q = a.e_q[e0 + i];
                    const int n = a.e_n[e0 + i];
                    double u;
                    if (is_root) {
                        const double nz = noise ? noise[i] : gam[i] / gsum;
                        const double pp = (double)(0.75f * p) + 0.25 * nz;        // (1-e)*p_ + e*noise[i]
                        u = 0.7 * pp * xx / (double)(1 + n);
                    } else {
                        const float cp = 0.7f * p;                                 // c_puct * p_ stays float32 (NEP 50)
                        u = (double)cp * xx / (double)(1 + n);
                    }
                    const double b = q + u;
                    if (b > best) { best = b; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(FULL, best, o);
                    const int oi = __shfl_xor_sync(FULL, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                best_i = bi;
                if (is_root && lane == 0) T.root_selects++;
            }
            // ---- virtual loss (solo_play.py:205-208), remember the path, play the move
            const int action = a.e_action[e0 + best_i];
            if (lane == 0) {
                nd.sum_n += 1;
                a.e_n[e0 + best_i] += 1;
                a.e_w[e0 + best_i] = a.e_w[e0 + best_i] + (-1.0);
                a.e_q[e0 + best_i] = a.e_w[e0 + best_i] / (double)a.e_n[e0 + best_i];
                T.path_node[depth] = id; T.path_edge[depth] = best_i;
            }
            __syncwarp();
            depth++;
            if (action >= 0) {                                     // env_hive.py:105-148
                const int k = action % 11, end = action / 11, p = side * 11 + k;
                const int h_end = __popc(__ballot_sync(FULL, cell == end));
                if (lane == p) { cell = end; level = h_end; }
                last_real = true;
            } else {
                last_real = false;                                  // pass: no history push (env_hive.py:100-103)
            }
            turn++;
        }
        // ---- terminal / sentinel simulation: back up now and start the next one
        backup_path(a, t, T, depth, r, lane);
        if (lane == 0) T.sims_done++;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_expand_kernel(MctsArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * MCTS_WARPS + warp;
    if (t >= a.n) return;
    MctsTree& T = a.trees[t];
    if (T.state != TREE_NEED_EVAL) return;
    MctsNode* nodes = a.nodes + (size_t)t * a.node_cap;
    const size_t eb = (size_t)t * a.edge_cap;
    const GameRec* sim = a.sim_recs + t;
    const int count = a.sim_count[t];
    const int ne = count > 0 ? count : 1;
    const int id = T.n_nodes, off = T.edges_used;
    if (id >= a.node_cap || off + ne > a.edge_cap) {
        // arena full: the simulation is given up (virtual losses taken back), the tree stops and the batch is flagged
        abort_simulation(a, t, T, T.depth, (id >= a.node_cap) ? 1 : 2, lane);
        if (lane == 0) { T.state = TREE_IDLE; a.need_eval[t] = 0; a.env_mask[t] = 0; }
        return;
    }
    int cell = HAND, level = 0;
    if (lane < N_PIECE) { cell = sim->cell[lane]; level = sim->level[lane]; }
    const int turn = reinterpret_cast<const uint32_t*>(sim)[11] & 0xFF;
    const int player = (turn & 1) ? 0 : 1;
    MctsNode& nd = nodes[id];
    if (lane < N_PIECE) { nd.cell[lane] = (uint8_t)cell; nd.level[lane] = (uint8_t)level; }
    if (lane == 0) {
        nd.player = (uint8_t)player; nd.normalized = count > 0 ? 0 : 1;
        nd.sum_n = 0; nd.n_edges = ne; nd.edge_off = off; nd.n_legal = count;
    }
    // edges in ascending action order with the raw network priors (tree[state].p = leaf_p)
    if (count > 0) {
        const uint32_t* words = a.sim_legal + (size_t)t * LEGAL_WORDS;
        uint32_t w0 = 0, w1 = 0;
        if (lane < 25) { w0 = words[2 * lane]; w1 = words[2 * lane + 1]; }
        const int c = __popc(w0) + __popc(w1);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        int pos = incl - c;
        uint64_t m = ((uint64_t)w1 << 32) | w0;
        while (m) {
            const int b = __ffsll((long long)m) - 1; m &= m - 1;
            const int act = lane * 64 + b;
            const size_t e = eb + off + pos++;
            a.e_action[e] = (int16_t)act; a.e_n[e] = 0; a.e_w[e] = 0.0; a.e_q[e] = 0.0;
            a.e_p[e] = a.leaf_p[(size_t)t * 1584 + act];
        }
    } else if (lane == 0) {
        const size_t e = eb + off;                                  // the pass edge (solo_play.py:298-300,203)
        a.e_action[e] = -1; a.e_n[e] = 0; a.e_w[e] = 0.0; a.e_q[e] = 0.0; a.e_p[e] = 0.f;
    }
    // insert into the transposition table
    {
        int32_t* ht = a.htab + (size_t)t * a.ht_size;
        uint32_t idx = key_hash(lane, cell, level, player) & (uint32_t)(a.ht_size - 1);
        if (lane == 0) {
            while (ht[idx] >= 0) idx = (idx + 1) & (uint32_t)(a.ht_size - 1);
            ht[idx] = id;
        }
    }
    __syncwarp();
    backup_path(a, t, T, T.depth, a.leaf_v[t], lane);
    if (lane == 0) {
        T.n_nodes = id + 1; T.edges_used = off + ne; T.sims_done++;
        T.state = TREE_IDLE; a.need_eval[t] = 0; a.env_mask[t] = 0;
    }
}

// calc_policy + apply_temperature (solo_play.py:337-374): pi = N / sum N, or the priors when every
// W is negative; the move is the first maximum of pi.
__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_finalize_kernel(MctsArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * MCTS_WARPS + warp;
    if (t >= a.n) return;
    if (a.tree_mask && !a.tree_mask[t]) return;
    const MctsTree& T = a.trees[t];
    double* pi = a.pi + (size_t)t * 1584;
    for (int i = lane; i < 1584; i += 32) pi[i] = 0.0;
    __syncwarp();
    if (T.n_nodes == 0) { if (lane == 0) { a.out_action[t] = -1; a.out_sum_n[t] = 0; } return; }
    const MctsNode& nd = (a.nodes + (size_t)t * a.node_cap)[0];     // the first node created is the root
    const size_t e0 = (size_t)t * a.edge_cap + nd.edge_off;
    const int ne = nd.n_edges;
    double sum = 0.0, maxw = -1e300;
    for (int i = lane; i < ne; i += 32) { sum += (double)a.e_n[e0 + i]; maxw = fmax(maxw, a.e_w[e0 + i]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(FULL, sum, o); maxw = fmax(maxw, __shfl_xor_sync(FULL, maxw, o)); }
    const bool use_prior = maxw < 0.0;
    double best = -1.0; int bi = 0x7fffffff;
    for (int i = lane; i < ne; i += 32) {
        int act = a.e_action[e0 + i];
        if (act < 0) act = 1583;                                    // policy[-1] (solo_play.py:360-362)
        const double v = use_prior ? (double)a.e_p[e0 + i] : (double)a.e_n[e0 + i] / sum;
        pi[act] = v;
    }
    __syncwarp();
    for (int i = lane; i < 1584; i += 32) { const double v = pi[i]; if (v > best) { best = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(FULL, best, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    // a root without legal actions holds only the pass edge: the batched driver passes (-1); the policy
    // vector keeps the reference's policy[-1] quirk
    if (lane == 0) { a.out_action[t] = nd.n_legal > 0 ? bi : -1; a.out_sum_n[t] = (int)sum; }
}

}  // namespace hive
