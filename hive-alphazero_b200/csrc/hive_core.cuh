// hive_core.cuh -- device building blocks of the Hive position evaluator for sm_100a.
//
// Boards are whole 144-cell bit strings held in registers (cell = q*12+r, 5 x u32), so translating a
// board along one of the six torus directions (tile.py:111-123) is a handful of funnel shifts and no
// shuffles.  A position is evaluated in three stages (kernels in hive_env_kernel.cuh):
//   analyse -- one warp per game, lane p (0..21) owns piece p (white Q,B0,B1,S0,S1,G0,G1,G2,A0,A1,A2
//              then the same for black: reference order, inventory_frame.py:47-99 / env_hive.py:71-87).
//              Board-wide facts (occupancy, colour masks) are OR-reduced across lanes with REDUX
//              (__reduce_or_sync), stack heights come from __match_any_sync; pieces that need a one-hive
//              flood or a move search are queued.
//   search  -- one thread per queued piece, warps homogeneous in piece type: one-hive flood, Ant flood
//              over slide gates, exact 3-step Spider walk, Grasshopper line flood, Queen/Beetle ring logic.
//   encode  -- one warp per game again: dense 1584-bit legal mask and the 56 network planes assembled in
//              shared memory and written as bf16 with 16-byte stores.
//
// What is computed is exactly what the reference computes in GamePlay.move()'s tail
// (hive_engine/env_hive.py:170-171): pre_actions() (env_hive.py:196-304, move_checker.py:9-55,
// pieces.py:35-158) and make_state_value() (env_hive.py:320-485), bug-for-bug (SURVEY.md
// Appendix A/B).  No code is shared with oracle/hive_oracle.c.
#pragma once
#ifndef HIVE_EMU
#include <cuda_runtime.h>
#endif
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace hive {

constexpr unsigned FULL = 0xffffffffu;
constexpr int HAND = 255;
constexpr int N_PIECE = 22;
constexpr int N_PLANE = 56;
constexpr int N_PLANE_C = 56;
constexpr int LEGAL_WORDS = 50;      // 1584 bits -> 49.5 u32 (25 u64)
// bit-plane record handed from the encode kernel to the plane-store kernel: 56 planes x 5 words; the slot of
// plane 31 (the turn plane holds no bits) carries the turn in word 0 and "evaluated in this launch" in word 1
constexpr int BITS_WORDS = N_PLANE_C * 5, BITS_TURN = 31 * 5, BITS_LIVE = 31 * 5 + 1;
// constant geometry tables behind EnvArgs::hop_lines (built on the host by hive_tables.h): is_straight_line masks,
// per-cell neighbour ranks (3 bits per direction) and per-cell neighbour cells (six bytes in two words)
constexpr int GEO_HOP = 0, GEO_RANK = 144 * 5, GEO_NBR = GEO_RANK + 144, GEO_WORDS = GEO_NBR + 288;
constexpr int START_CELL = 6 * 12 + 6;   // tile.py:156,188 Start_Tile
constexpr int TURN2_CELL = 5 * 12 + 6;   // core_index ('M','13'), env_hive.py:157-159

enum PieceType { T_QUEEN = 0, T_BEETLE = 1, T_SPIDER = 2, T_HOPPER = 3, T_ANT = 4 };

// ------------------------------------------------------------------------------------------
// 384-byte game record in HBM (DESIGN.md "state record").
struct __align__(16) GameRec {
    uint8_t cell[N_PIECE];     // q*12+r, HAND = in inventory
    uint8_t level[N_PIECE];    // index in its stack (env_hive.py:119,125)
    uint8_t turn;              // 1-based (game_state.py:38)
    uint8_t winner;            // 0 none, 1 white, 2 black
    uint8_t done;              // game_is_over() (move_checker.py:140-165)
    uint8_t flags;
    uint32_t episode;          // bumped by every reset of this slot
    uint32_t steps;            // env steps taken in this slot
    uint32_t n_legal;
    uint32_t pad;
    uint32_t hist[2][4][2][5]; // [side][age][own-any, opp-any] boards (env_hive.py:431-445)
};
static_assert(sizeof(GameRec) == 384, "GameRec must stay 384 bytes");

// per-warp shared-memory scratch of the encode kernel (hist / legal move with vector accesses)
struct __align__(16) WarpScratch {
    uint32_t hist[2][4][2][5];     // 320 B, 16-byte aligned
    uint32_t planes[N_PLANE][5];   // bit boards of the 56 planes (plane 31 unused: it is the turn)
    uint32_t legal[LEGAL_WORDS + 2];
    uint32_t qn[12];               // queen neighbours in tile.adjacent_tiles order: cell | rank<<8 | empty<<12 (white 0..5, black 6..11)
    uint32_t moves[N_PIECE][5];    // the per-piece sets of GameScratch, staged so that other lanes can index them
    uint32_t occ[6];
};
static_assert(sizeof(WarpScratch) % 16 == 0, "WarpScratch must keep 16-byte alignment in arrays");
static_assert(offsetof(WarpScratch, hist) % 16 == 0 && offsetof(WarpScratch, legal) % 8 == 0, "vector access alignment");
static_assert(offsetof(WarpScratch, planes) % 16 == 0 && offsetof(WarpScratch, legal) == offsetof(WarpScratch, planes) + N_PLANE * 20, "planes+legal are zeroed as one 16-byte aligned run");

// Per-game intermediate record handed from kernel to kernel through L2 (656 B, never leaves the
// 126 MB L2 at 16,384 games):
//   info[p] = cell | height<<8 | top<<12 | level<<13 | ring<<16 (ring = occupancy of the six neighbours)
//   head[0] = turn | cq_w<<8 | cq_b<<16, head[1] = pinned-piece mask,
//   head[2] = live | push_history<<1 | prev_winner<<8
struct __align__(16) GameScratch {
    uint32_t info[24];
    uint32_t head[4];
    uint32_t occ[8];
    uint32_t own_all[8];
    uint32_t opp_all[8];
    uint32_t moves[N_PIECE][5];    // action list per piece (own) / mobility set (opponent)
    uint32_t pad[2];
};
static_assert(sizeof(GameScratch) == 656, "GameScratch layout");

// Work queues.  The analyse kernel first collects the items of its 16 games in shared memory
// (GroupQueues), then reserves a slice of the batch-wide queues with one atomic per class and CTA.
// item = game<<6 | piece<<1 | wants_moves
#ifndef HIVE_GROUP
#define HIVE_GROUP 8
#endif
constexpr int GROUP = HIVE_GROUP;  // games per CTA of the analyse kernel (<= 16: queue items carry the slot in 4 bits)
static_assert(GROUP >= 1 && GROUP <= 16, "GROUP");
struct __align__(16) GroupQueues {
    uint32_t n_flood;
    uint32_t n_mv[4];              // move classes: 0 Ant, 1 Grasshopper, 2 Spider, 3 Queen/Beetle
    uint32_t base[5];              // reserved offsets in the batch-wide queues
    uint32_t pad[2];
    uint16_t flood[GROUP * N_PIECE];
    uint16_t mv[4][GROUP * 6];
};
struct BatchQueues {
    uint32_t* counters;            // [8]: n_flood, n_mv[4]
    uint32_t* flood;               // capacity n*22
    uint32_t* mv[4];               // capacity n*6 each
};

// ------------------------------------------------------------------------------------------
// 144-bit boards
struct BB { uint32_t w[5]; };

// column-0 / column-11 positions of each 12-cell row inside the linear bit string
#define HIVE_COLFIRST(i) ((i) == 0 ? 0x01001001u : (i) == 1 ? 0x10010010u : (i) == 2 ? 0x00100100u : (i) == 3 ? 0x01001001u : 0x00000010u)
#define HIVE_COLLAST(i)  ((i) == 0 ? 0x00800800u : (i) == 1 ? 0x08008008u : (i) == 2 ? 0x80080080u : (i) == 3 ? 0x00800800u : 0x00008008u)

__device__ __forceinline__ BB bb_zero() { BB r; r.w[0] = r.w[1] = r.w[2] = r.w[3] = r.w[4] = 0; return r; }
__device__ __forceinline__ BB bb_bit(int c) {
    BB r; uint32_t b = 1u << (c & 31); int wi = c >> 5;
#pragma unroll
    for (int i = 0; i < 5; i++) r.w[i] = (wi == i) ? b : 0u;
    return r;
}
__device__ __forceinline__ BB operator|(const BB& a, const BB& b) { BB r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.w[i] = a.w[i] | b.w[i]; return r; }
__device__ __forceinline__ BB operator&(const BB& a, const BB& b) { BB r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.w[i] = a.w[i] & b.w[i]; return r; }
__device__ __forceinline__ BB operator^(const BB& a, const BB& b) { BB r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.w[i] = a.w[i] ^ b.w[i]; return r; }
__device__ __forceinline__ BB bb_andn(const BB& a, const BB& b) { BB r;   // a & ~b
#pragma unroll
    for (int i = 0; i < 5; i++) r.w[i] = a.w[i] & ~b.w[i]; return r; }
__device__ __forceinline__ bool bb_any(const BB& a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3] | a.w[4]) != 0; }
__device__ __forceinline__ bool bb_eq(const BB& a, const BB& b) {
    return ((a.w[0] ^ b.w[0]) | (a.w[1] ^ b.w[1]) | (a.w[2] ^ b.w[2]) | (a.w[3] ^ b.w[3]) | (a.w[4] ^ b.w[4])) == 0; }
__device__ __forceinline__ int bb_popc(const BB& a) {
    return __popc(a.w[0]) + __popc(a.w[1]) + __popc(a.w[2]) + __popc(a.w[3]) + __popc(a.w[4]); }
__device__ __forceinline__ bool bb_test(const BB& a, int c) {
    uint32_t w = c < 64 ? (c < 32 ? a.w[0] : a.w[1]) : (c < 96 ? a.w[2] : (c < 128 ? a.w[3] : a.w[4]));
    return (w >> (c & 31)) & 1u;
}
__device__ __forceinline__ bool words_test(const uint32_t* w, int c) { return (w[c >> 5] >> (c & 31)) & 1u; }

// (q, r) -> (q, r+1 mod 12)
__device__ __forceinline__ BB bb_colL(const BB& x) {
    BB r;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        uint32_t shl = (i == 0) ? (x.w[0] << 1) : __funnelshift_l(x.w[i - 1], x.w[i], 1);
        uint32_t shr = (i == 4) ? (x.w[4] >> 11) : __funnelshift_r(x.w[i], x.w[i + 1], 11);
        uint32_t f = HIVE_COLFIRST(i);
        r.w[i] = (shl & ~f) | (shr & f);
    }
    r.w[4] &= 0xFFFFu;
    return r;
}
// (q, r) -> (q, r-1 mod 12)
__device__ __forceinline__ BB bb_colR(const BB& x) {
    BB r;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        uint32_t shr = (i == 4) ? (x.w[4] >> 1) : __funnelshift_r(x.w[i], x.w[i + 1], 1);
        uint32_t shl = (i == 0) ? (x.w[0] << 11) : __funnelshift_l(x.w[i - 1], x.w[i], 11);
        uint32_t l = HIVE_COLLAST(i);
        r.w[i] = (shr & ~l) | (shl & l);
    }
    r.w[4] &= 0xFFFFu;
    return r;
}
// (q, r) -> (q+1 mod 12, r)
__device__ __forceinline__ BB bb_up(const BB& x) {
    BB r;
    r.w[0] = (x.w[0] << 12) | (x.w[4] >> 4);
    r.w[1] = __funnelshift_l(x.w[0], x.w[1], 12);
    r.w[2] = __funnelshift_l(x.w[1], x.w[2], 12);
    r.w[3] = __funnelshift_l(x.w[2], x.w[3], 12);
    r.w[4] = __funnelshift_l(x.w[3], x.w[4], 12) & 0xFFFFu;
    return r;
}
// (q, r) -> (q-1 mod 12, r)
__device__ __forceinline__ BB bb_down(const BB& x) {
    BB r;
    r.w[0] = __funnelshift_r(x.w[0], x.w[1], 12);
    r.w[1] = __funnelshift_r(x.w[1], x.w[2], 12);
    r.w[2] = __funnelshift_r(x.w[2], x.w[3], 12);
    r.w[3] = __funnelshift_r(x.w[3], x.w[4], 12);
    r.w[4] = (x.w[4] >> 12) | ((x.w[0] & 0xFFFu) << 4);
    return r;
}
// union of the six torus neighbours of every set cell (tile.py:114-121)
__device__ __forceinline__ BB bb_nbrs(const BB& x) {
    BB a = bb_colL(x), b = bb_colR(x);
    BB u = bb_up(x | a), d = bb_down(x | b);
    return a | b | u | d;
}

// direction ring d0=(+1,0) d1=(+1,+1) d2=(0,+1) d3=(-1,0) d4=(-1,-1) d5=(0,-1): consecutive
// directions are adjacent, so the two common neighbours of c and c+d_i are c+d_{i-1}, c+d_{i+1}.
__device__ __forceinline__ int cell_up(int c)   { c += 12; return c >= 144 ? c - 144 : c; }
__device__ __forceinline__ int cell_down(int c) { c -= 12; return c < 0 ? c + 144 : c; }
__device__ __forceinline__ int cell_colL(int c) { return (c % 12 == 11) ? c - 11 : c + 1; }
__device__ __forceinline__ int cell_colR(int c) { return (c % 12 == 0) ? c + 11 : c - 1; }
__device__ __forceinline__ int cell_nbr(int c, int i) {
    switch (i) {
        case 0: return cell_up(c);
        case 1: return cell_up(cell_colL(c));
        case 2: return cell_colL(c);
        case 3: return cell_down(c);
        case 4: return cell_down(cell_colR(c));
        default: return cell_colR(c);
    }
}
__device__ __forceinline__ uint32_t rot6l(uint32_t x) { return ((x << 1) | (x >> 5)) & 63u; }
__device__ __forceinline__ uint32_t rot6r(uint32_t x) { return ((x >> 1) | (x << 5)) & 63u; }

__device__ __forceinline__ BB warp_or(const BB& x) {
    BB r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.w[i] = __reduce_or_sync(FULL, x.w[i]);
    return r;
}

// ------------------------------------------------------------------------------------------
// Slide relation R of Ant / Spider (move_checker.py:189-214 inside path_exists :235-240), in
// destination form: arriving at n along d_i is allowed iff n is empty and exactly one of
// n+d_{i-2}, n+d_{i+2} (the two flanks of the step) is occupied.  `occp` has the mover lifted.
struct Slide { BB g[6]; };

__device__ __forceinline__ void slide_init(Slide& s, const BB& occp) {
    BB cL = bb_colL(occp), cR = bb_colR(occp);
    BB o0 = bb_down(occp), o3 = bb_up(occp), o1 = bb_down(cR), o4 = bb_up(cL);
    const BB& o2 = cR; const BB& o5 = cL;
    s.g[0] = bb_andn(o4 ^ o2, occp);
    s.g[1] = bb_andn(o5 ^ o3, occp);
    s.g[2] = bb_andn(o0 ^ o4, occp);
    s.g[3] = bb_andn(o1 ^ o5, occp);
    s.g[4] = bb_andn(o2 ^ o0, occp);
    s.g[5] = bb_andn(o3 ^ o1, occp);
}
__device__ __forceinline__ BB slide_step(const Slide& s, const BB& x) {
    BB a = bb_colL(x), b = bb_colR(x);
    BB r = (bb_up(x) & s.g[0]) | (bb_up(a) & s.g[1]) | (a & s.g[2]);
    r = r | (bb_down(x) & s.g[3]) | (bb_down(b) & s.g[4]) | (b & s.g[5]);
    return r;
}

// move_checker.py:106-137; nq/first-queen colour derived from the two queen cells
__device__ __forceinline__ bool obeys_queen_by_4(int turn, bool wq_on, bool bq_on, bool mover_queen, int mover_color) {
    int nq = (int)wq_on + (int)bq_on;
    if (nq == 2) return true;
    if (nq == 0) return mover_queen && ((turn == 7 && mover_color == 0) || (turn == 8 && mover_color == 1));
    bool c_white = wq_on;
    if (c_white) return turn == 7 || mover_queen;      // (white,7) -> True ; (white,8) -> mover is a Queen
    return turn == 8 || mover_queen;                   // (black,8) -> True ; (black,7) -> mover is a Queen
}

struct EvalResult { int n_legal; int done; int winner; };

// ------------------------------------------------------------------------------------------
// A position is evaluated by three kernels (hive_env_kernel.cuh):
//   analyse (warp <-> game, lane p <-> piece p): stacks, occupancy, ring occupancy, placements,
//            turn gates; publishes one info word per piece and queues the pieces that need a
//            one-hive flood and/or a move search;
//   search  (thread <-> queued piece of a 16-game group, move searches grouped BY PIECE TYPE so a
//            warp runs Ant floods, or Spider walks ... of several games at once);
//   encode  (warp <-> game again): dense legal mask, the 56 planes, history push, terminal test.
__device__ __forceinline__ int move_class(int type) {
    return type == T_ANT ? 0 : type == T_HOPPER ? 1 : type == T_SPIDER ? 2 : 3;
}
__device__ __forceinline__ int piece_type_of(int k) {
    return (k == 0) ? T_QUEEN : (k <= 2) ? T_BEETLE : (k <= 4) ? T_SPIDER : (k <= 7) ? T_HOPPER : T_ANT;
}

__device__ __forceinline__ void eval_analyse(GameScratch& gs, GroupQueues& q, uint32_t* occ_s, int game_slot, int lane, int cell,
                                             int level, int turn, bool push_history, int prev_winner,
                                             const uint32_t* __restrict__ geo) {
    const int side = (turn & 1) ? 0 : 1;                 // game_state.py:58-62
    const bool valid = lane < N_PIECE;
    const int color = lane >= 11 ? 1 : 0;
    const int k = lane - 11 * color;
    const int type = piece_type_of(k);
    const bool own = valid && (color == side);
    const bool on_board = valid && cell != HAND;

    // stacks: pieces sharing a cell (tile.pieces); top piece <=> level+1 == len (env_hive.py:213)
    const unsigned peers = __match_any_sync(FULL, on_board ? cell : 256 + lane);
    const int height = __popc(peers);
    const bool top = on_board && (level == height - 1);

    const BB src = on_board ? bb_bit(cell) : bb_zero();
    const BB own_all = warp_or(own ? src : bb_zero());
    const BB opp_all = warp_or((valid && !own) ? src : bb_zero());
    const BB occ = own_all | opp_all;
    const BB top_opp = warp_or((top && !own) ? src : bb_zero());

    const int cq_w = __shfl_sync(FULL, cell, 0), cq_b = __shfl_sync(FULL, cell, 11);
    const bool wq_on = cq_w != HAND, bq_on = cq_b != HAND;
    const bool ownq_on = (side == 0 ? cq_w : cq_b) != HAND;
    const unsigned in_hand = __ballot_sync(FULL, own && !on_board);

    if (lane < 5) occ_s[lane] = occ.w[lane];             // per-warp shared copy for the random-access ring tests
    __syncwarp();
    uint32_t ring = 0;                                   // occupancy of the six neighbours (cells from the GEO_NBR table)
    if (on_board) {
        const uint32_t n03 = __ldg(geo + GEO_NBR + 2 * cell), n45 = __ldg(geo + GEO_NBR + 2 * cell + 1);
#pragma unroll
        for (int i = 0; i < 6; i++) ring |= (uint32_t)words_test(occ_s, (int)(((i < 4 ? n03 : n45) >> (8 * (i & 3))) & 0xFFu)) << i;
    }

    // turn gates shared by every candidate of a piece (move_checker.py:38-55)
    bool gate = true;
    if (turn <= 2) gate = false;                                             // no on-board mover can exist / matter
    else if (turn <= 6) gate = ownq_on;                                      // queen_is_on_board: colour by turn parity
    else if (turn <= 8) gate = obeys_queen_by_4(turn, wq_on, bq_on, type == T_QUEEN, color);
    // opponent mobility is only consumed through the own queen's neighbourhood (env_hive.py:459-478)
    const bool wants_moves = top && gate && (own || ownq_on);
    bool pinned_now = false, need_flood = false;
    if (top && height == 1) {
        if (ring == 0) pinned_now = true;                                    // nothing left on the board -> `return False`
        else need_flood = __popc(ring & ~rot6l(ring)) > 1;                   // >1 arc of neighbours: may be an articulation point
    }
    const unsigned pin_mask = __ballot_sync(FULL, pinned_now);

    // moves row: zero for on-board pieces (the search kernel fills movers), placements for the hand
    BB mv = bb_zero();
    if (valid && !on_board) {
        // placements (env_hive.py:217-225; move_checker.py:168-179): first in-hand piece per type
        const unsigned same_type_before = in_hand & ((1u << lane) - 1u) &
            (type == T_QUEEN ? 0x00000801u : type == T_BEETLE ? 0x00003006u : type == T_SPIDER ? 0x0000C018u
             : type == T_HOPPER ? 0x000700E0u : 0x00380700u);
        if (own && same_type_before == 0) {
            if (turn == 1) mv = bb_bit(START_CELL);
            else if (turn == 2) mv = bb_andn(bb_nbrs(occ), occ) & bb_bit(TURN2_CELL);
            else {
                bool ok = true;
                if (turn == 7 || turn == 8) ok = obeys_queen_by_4(turn, wq_on, bq_on, type == T_QUEEN, color);
                if (ok) mv = bb_andn(bb_andn(bb_nbrs(occ), occ), bb_nbrs(top_opp));
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int i = 0; i < 5; i++) gs.moves[lane][i] = mv.w[i];
        gs.info[lane] = (uint32_t)cell | ((uint32_t)height << 8) | ((uint32_t)top << 12) | ((uint32_t)level << 13) | (ring << 16);
    }
    if (lane < 5) { gs.occ[lane] = occ.w[lane]; gs.own_all[lane] = own_all.w[lane]; gs.opp_all[lane] = opp_all.w[lane]; }
    if (lane == 0) {
        gs.head[0] = (uint32_t)turn | ((uint32_t)cq_w << 8) | ((uint32_t)cq_b << 16);
        gs.head[1] = pin_mask;
        gs.head[2] = 1u | ((uint32_t)push_history << 1) | ((uint32_t)prev_winner << 8);
    }
    if (on_board) {
        const uint32_t item = (uint32_t)game_slot | ((uint32_t)lane << 4) | ((uint32_t)wants_moves << 9);
        if (need_flood) q.flood[atomicAdd(&q.n_flood, 1u)] = (uint16_t)item;
        else if (wants_moves && !pinned_now) {
            const int cls = move_class(type);
            q.mv[cls][atomicAdd(&q.n_mv[cls], 1u)] = (uint16_t)item;
        }
    }
}

// search, part 1: one thread, one one-hive flood (move_checker.py:58-83 / env_hive.py:509-530):
// lift the top piece and test that the rest of the hive stays connected.  Returns true if pinned.
__device__ __forceinline__ bool eval_flood(GameScratch& gs, int p) {
    const uint32_t info = gs.info[p];
    const int cell = info & 0xFF;
    const uint32_t ring = (info >> 16) & 63u;
    BB occp;
#pragma unroll
    for (int i = 0; i < 5; i++) occp.w[i] = gs.occ[i];
    const BB src = bb_bit(cell);
    occp = occp ^ src;
    const BB goal = bb_nbrs(src) & occp;
    BB x = bb_bit(cell_nbr(cell, __ffs(ring) - 1));
    for (;;) {
        BB nx = x | (bb_nbrs(x) & occp);
        if (bb_eq(nx & goal, goal)) return false;
        if (bb_eq(nx, x)) return true;
        x = nx;
    }
}

// search, part 2: one thread, the move set of one unpinned top piece whose turn gates are open.
// `hop_lines` = 144x5 u32 table of is_straight_line masks (move_checker.py:249-265).
__device__ __forceinline__ void eval_moves(GameScratch& gs, int p, const uint32_t* __restrict__ hop_lines) {
    const uint32_t info = gs.info[p];
    const int cell = info & 0xFF;
    const int height = (info >> 8) & 0xF;
    const uint32_t ring = (info >> 16) & 63u;
    const int type = piece_type_of(p >= 11 ? p - 11 : p);
    uint32_t* row = gs.moves[p];
    BB occ;
#pragma unroll
    for (int i = 0; i < 5; i++) occ.w[i] = gs.occ[i];

    if (type == T_QUEEN || type == T_BEETLE) {
        uint32_t ok;
        if (type == T_QUEEN) {                                           // pieces.py:35-44
            ok = ~ring & (rot6l(ring) ^ rot6r(ring)) & 63u;
        } else {                                                         // pieces.py:100-113
            uint32_t fl = rot6l(ring), fr = rot6r(ring);                 // bit i: flank c+d_{i-1} / c+d_{i+1}
            uint32_t k1 = fl ^ fr, k0 = ~(fl | fr) & 63u;
            ok = k1 | ring | (height > 1 ? 63u : 0u);
            if (k0 & ~ok) {
                // k==0: allowed iff the target has an occupied neighbour other than `old`
                // (len(new_adjacents_with_pieces) - 1 != 0, move_checker.py:205-207)
                const BB hns = bb_nbrs(occ ^ bb_bit(cell));
#pragma unroll
                for (int i = 0; i < 6; i++)
                    if (((k0 & ~ok) >> i) & 1u) { if (bb_test(hns, cell_nbr(cell, i))) ok |= 1u << i; }
            }
        }
        uint32_t w[5] = {0u, 0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 6; i++)
            if ((ok >> i) & 1u) {
                const int c = cell_nbr(cell, i); const uint32_t bit = 1u << (c & 31); const int wi = c >> 5;
#pragma unroll
                for (int j = 0; j < 5; j++) w[j] |= (wi == j) ? bit : 0u;
            }
#pragma unroll
        for (int i = 0; i < 5; i++) row[i] = w[i];
        return;
    }

    const BB src = bb_bit(cell);
    BB mv;
    if (type == T_HOPPER) {                                              // pieces.py:128-158
        BB line;
#pragma unroll
        for (int i = 0; i < 5; i++) line.w[i] = __ldg(hop_lines + cell * 5 + i);
        const BB walk = occ & line;
        BB v = src;
        for (;;) {
            BB nv = v | (bb_nbrs(v) & walk);
            if (bb_eq(nv, v)) break;
            v = nv;
        }
        mv = bb_andn(bb_andn(bb_nbrs(v) & line, occ), bb_nbrs(src));
    } else {                                                             // Ant / Spider
        const BB occp = occ ^ src;
        Slide sl;
        slide_init(sl, occp);
        if (type == T_ANT) {                                             // pieces.py:59-63
            BB x = src;
            for (;;) {
                BB nx = x | slide_step(sl, x);
                if (bb_eq(nx, x)) break;
                x = nx;
            }
            mv = bb_andn(x, src);
        } else {                                                         // pieces.py:78-85
            mv = bb_zero();
#ifdef HIVE_SPIDER_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int i = 0; i < 6; i++) {
                const int c1 = cell_nbr(cell, i);
                if (!bb_test(sl.g[i], c1)) continue;
                const BB a = bb_bit(c1);
                const BB b = bb_andn(slide_step(sl, a), src);
                const BB c = bb_andn(slide_step(sl, b), src | a);
                mv = mv | c;
            }
            // end check with the spider back on `old`: adjacent target with both flanks occupied
            const uint32_t k2 = rot6l(ring) & rot6r(ring);
#pragma unroll
            for (int i = 0; i < 6; i++)
                if ((k2 >> i) & 1u) mv = bb_andn(mv, bb_bit(cell_nbr(cell, i)));
        }
    }
#pragma unroll
    for (int i = 0; i < 5; i++) row[i] = mv.w[i];
}

// encode: legal mask + all 56 planes (env_hive.py:287-304, 320-447) into shared
// memory, history push, terminal test.  "own" = side to move.
// Everything the encode step reads of a game's GameScratch, fetched by encode_fetch with all loads issued
// back to back (one L2 round trip; the profile of the first version showed five serial ones).
struct EncodeIn {
    uint4 head;                    // GameScratch::head
    uint32_t info;                 // info[lane] (HAND for lanes >= 22)
    BB mv;                         // moves[lane]
    uint32_t occ_w, own_w, opp_w;  // word `lane` of the three boards (lanes 0..4)
};
__device__ __forceinline__ EncodeIn encode_fetch(const GameScratch& gs, int lane) {
    EncodeIn in;
    in.head = *reinterpret_cast<const uint4*>(gs.head);
    const int l = lane < N_PIECE ? lane : 0, w = lane < 5 ? lane : 0;
    in.info = gs.info[l];
#pragma unroll
    for (int i = 0; i < 5; i++) in.mv.w[i] = gs.moves[l][i];
    in.occ_w = gs.occ[w]; in.own_w = gs.own_all[w]; in.opp_w = gs.opp_all[w];
    if (lane >= N_PIECE) { in.info = (uint32_t)HAND; in.mv = bb_zero(); }
    if (lane >= 5) in.occ_w = in.own_w = in.opp_w = 0;
    return in;
}

__device__ __forceinline__ EvalResult eval_encode(WarpScratch& sm, const EncodeIn& in, int lane, const uint32_t* __restrict__ geo) {
    const uint32_t head = in.head.x, flags = in.head.z;
    const int turn = head & 0xFF, cq_w = (head >> 8) & 0xFF, cq_b = (head >> 16) & 0xFF;
    const bool push_history = (flags >> 1) & 1u;
    const int prev_winner = (flags >> 8) & 0xFF;
    const int side = (turn & 1) ? 0 : 1;
    const bool valid = lane < N_PIECE;
    const int color = lane >= 11 ? 1 : 0;
    const int k = lane - 11 * color;
    const int type = piece_type_of(k);
    const bool own = valid && (color == side);
    const uint32_t info = in.info;
    const int cell = info & 0xFF, level = (info >> 13) & 7;
    const bool on_board = valid && cell != HAND;
    const bool top = (info >> 12) & 1u;
    const uint32_t ring = (info >> 16) & 63u;
    const bool pinned = (in.head.y >> lane) & 1u;

    {   // zero the scratch outputs: planes (1120 B) and legal (208 B) are contiguous and 16-byte aligned
        uint4* pz = reinterpret_cast<uint4*>(&sm.planes[0][0]);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < 3; i++) { const int t = lane + 32 * i; if (t < (N_PLANE * 20 + (LEGAL_WORDS + 2) * 4) / 16) pz[t] = z; }
    }
    const BB mv = in.mv;
    const uint32_t occ_w = in.occ_w, own_w = in.own_w, opp_w = in.opp_w;
    if (valid) {
#pragma unroll
        for (int i = 0; i < 5; i++) sm.moves[lane][i] = mv.w[i];
    }
    if (lane < 5) sm.occ[lane] = occ_w;
    __syncwarp();

    // dense legal mask a = cell*11 + k: the 11 own pieces x 5 board words are 55 work items spread
    // over the 32 lanes, so no lane walks more than two words
    int n_mine = 0;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int item = lane + 32 * r;
        if (item < 55) {
            const int kk = item / 5, w = item - kk * 5;
            uint32_t m = sm.moves[side * 11 + kk][w];
            n_mine += __popc(m);
            while (m) {
                const int b = __ffs(m) - 1; m &= m - 1;
                const int a = (w * 32 + b) * 11 + kk;
                atomicOr(&sm.legal[a >> 5], 1u << (a & 31));
            }
        }
    }
    int n_legal = n_mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_legal += __shfl_xor_sync(FULL, n_legal, o);

    // terminal test (move_checker.py:140-165)
    const unsigned surrounded = __ballot_sync(FULL, on_board && type == T_QUEEN && ring == 63u);
    const bool ws = surrounded & 1u, bs = (surrounded >> 11) & 1u;
    EvalResult res;
    res.n_legal = n_legal;
    res.done = ws || bs;
    res.winner = (ws && bs) ? prev_winner : ws ? 2 : bs ? 1 : prev_winner;

    {   // neighbours of both queens, ranked in tile.adjacent_tiles order (board_tiles order: q descending, then
        // r ascending; tile.py:111-123; cells and ranks from the GEO tables): lanes 0..5 white queen, 6..11 black queen.
        // The occupied ones are planes 32 (own queen) / 33 (opponent queen).
        const int which = lane >= 6 ? 1 : 0, qcell = which ? cq_b : cq_w, dir = lane - 6 * which;
        const bool use = lane < 12 && qcell != HAND;
        uint32_t nb = 0, rank = 0;
        if (use) {
            nb = (__ldg(geo + GEO_NBR + 2 * qcell + (dir >> 2)) >> (8 * (dir & 3))) & 0xFFu;
            rank = (__ldg(geo + GEO_RANK + qcell) >> (3 * dir)) & 7u;
        }
        const bool empty = use && !words_test(sm.occ, (int)nb);
        if (lane < 12) sm.qn[lane] = use ? (nb | (rank << 8) | ((uint32_t)empty << 12)) : 0u;
        if (use && !empty) atomicOr(&sm.planes[which == side ? 32 : 33][nb >> 5], 1u << (nb & 31));
    }
    __syncwarp();
    if (on_board) {
        const uint32_t bit = 1u << (cell & 31); const int wi = cell >> 5;
        sm.planes[(own ? 0 : 12) + k][wi] = bit;                          // 0-10 / 12-22
        if (type == T_BEETLE && level >= 2)                               // 24-26 / 27-29
            atomicOr(&sm.planes[(own ? 24 : 27) + level - 2][wi], bit);
        // 34: own pieces without a legal action; 35: opponent pieces covered or pinned
        if (!top || (own ? !bb_any(mv) : pinned)) atomicOr(&sm.planes[own ? 34 : 35][wi], bit);
        // 44+j: opponent pieces able to reach the j-th empty neighbour of the own queen;
        // 50+j: own on-board pieces whose action list holds the j-th empty neighbour of the opponent queen
        // (every piece looks at the queen of the other colour; table built once per game above).
        if (bb_any(mv)) {
            const uint32_t* tbl = &sm.qn[(1 - color) * 6];
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const uint32_t e = tbl[i];
                if (((e >> 12) & 1u) && words_test(sm.moves[lane], (int)(e & 0xFFu)))
                    atomicOr(&sm.planes[(own ? 50 : 44) + ((e >> 8) & 7u)][wi], bit);
            }
        }
    }
    if (lane < 5) {
        sm.planes[11][lane] = own_w;
        sm.planes[23][lane] = opp_w;
        sm.planes[30][lane] = occ_w;
    }
    {   // 36..43 history of the side to move
        const uint32_t* h = &sm.hist[side][0][0][0];
        for (int i = lane; i < 40; i += 32) (&sm.planes[36][0])[i] = h[i];
    }
    __syncwarp();
    // history push (env_hive.py:436-445): only after a real move / at reset
    if (push_history) {
        uint32_t* h = &sm.hist[side][0][0][0];
        uint32_t keep = (lane < 30) ? h[lane] : 0;         // ages 0..2 -> 1..3
        __syncwarp();
        if (lane < 30) h[10 + lane] = keep;
        if (lane < 5) { h[lane] = own_w; h[5 + lane] = opp_w; }
        __syncwarp();
    }
    return res;
}

// ------------------------------------------------------------------------------------------
// `lut` = 256 x uint4 in shared memory: byte of eight {0,1} cells -> eight bf16 values.
__device__ __forceinline__ void fill_bf16_lut(uint4* lut, int t) {
    if (t < 256) {
        uint4 v;
        v.x = ((t & 1) ? 0x3F80u : 0u) | ((t & 2) ? 0x3F800000u : 0u);
        v.y = ((t & 4) ? 0x3F80u : 0u) | ((t & 8) ? 0x3F800000u : 0u);
        v.z = ((t & 16) ? 0x3F80u : 0u) | ((t & 32) ? 0x3F800000u : 0u);
        v.w = ((t & 64) ? 0x3F80u : 0u) | ((t & 128) ? 0x3F800000u : 0u);
        lut[t] = v;
    }
}
// ------------------------------------------------------------------------------------------
// Plane store through the TMA: a warp expands its game's bit planes into a small bf16 staging ring in
// shared memory (STAGE_PLANES planes at a time) and lane 0 hands every filled stage to the bulk-copy
// engine (cp.async.bulk shared -> global).  The 16 KB per game leave the SM without occupying the
// load/store queue the position evaluation of the other warps needs for its shared-memory traffic.
#ifndef HIVE_STAGE_PLANES
#define HIVE_STAGE_PLANES 28
#endif
#ifndef HIVE_STAGE_BUFS
#define HIVE_STAGE_BUFS 2
#endif
constexpr int STAGE_PLANES = HIVE_STAGE_PLANES, STAGE_BUFS = HIVE_STAGE_BUFS;
constexpr int STAGE_CHUNKS = STAGE_PLANES * 18, STAGE_BYTES = STAGE_CHUNKS * 16, N_STAGE = N_PLANE / STAGE_PLANES;
static_assert(N_PLANE % STAGE_PLANES == 0 && N_STAGE % STAGE_BUFS == 0 && STAGE_BYTES % 16 == 0, "stages tile the 56 planes");

__device__ __forceinline__ void fence_proxy_async_smem() {
#ifndef HIVE_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_store_s2g(void* dst, const void* src_smem, uint32_t bytes) {
#ifdef HIVE_EMU
    memcpy(dst, src_smem, bytes);
#else
    // the planes are a pure output stream: mark their lines evict-first so that they do not push the step's small
    // working set (records, scratch, queues, legal masks) out of the L2
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(bytes), "l"(policy) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read() {      // at most PENDING of this thread's bulk stores still read shared memory
#ifndef HIVE_EMU
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
#endif
}

// `bytes`: the game's bit planes in shared memory, 20 B per plane of which 18 are used.
// A stage covers STAGE_PLANES planes = STAGE_PLANES*20 source bytes; lane l looks at source bytes l, l+32, ... of
// the stage.  Which plane and byte that is, whether it is one of the two padding bytes, and the 16-byte chunk of the
// staging buffer its eight bf16 values go to are the same for every stage and every game, so they are worked out
// once per call (SRC_ROUNDS offsets kept in registers) and the inner loop is: byte load at an immediate offset,
// LUT row load, 16-byte store.
constexpr int STAGE_SRC_BYTES = STAGE_PLANES * 20, SRC_ROUNDS = (STAGE_SRC_BYTES + 31) / 32;
constexpr int TURN_STAGE = 31 / STAGE_PLANES, TURN_CHUNK0 = (31 % STAGE_PLANES) * 18;   // plane 31 = the turn number
// `lut_s`: the LUT's shared-window address, read back from memory by the caller so that it stays in a register (ptxas
// otherwise re-materialises the window base -- S2R + LEA -- in front of every LUT load).
__device__ __forceinline__ void store_planes_bulk(const uint8_t* bytes, const uint4* lut, uint32_t lut_s, uint4* stage, int lane, int turn,
                                                  uint16_t* __restrict__ out) {
    const uint32_t tb = __float_as_uint((float)turn) >> 16;     // bf16(turn): turn <= 255 is exact
    const uint32_t tt = tb | (tb << 16);
    const uint4 turn4 = make_uint4(tt, tt, tt, tt);
    int dst[SRC_ROUNDS];                                        // chunk of the stage buffer, -1: padding byte / past the stage
#pragma unroll
    for (int i = 0; i < SRC_ROUNDS; i++) {
        const int sb = lane + 32 * i, pl = sb / 20, j = sb - 20 * pl;
        dst[i] = (sb < STAGE_SRC_BYTES && j < 18) ? pl * 18 + j : -1;
    }
    const uint8_t* mine = bytes + lane;
#ifdef HIVE_EMU
    (void)lut_s;
#endif
#pragma unroll
    for (int q = 0; q < N_STAGE; q++) {
        uint4* buf = stage + (q % STAGE_BUFS) * STAGE_CHUNKS;
        // the stage that used this buffer last (of this game or of the warp's previous game) has left shared memory
        if (lane == 0) bulk_wait_read<STAGE_BUFS - 1>();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < SRC_ROUNDS; i++)
            if (dst[i] >= 0) {
#ifdef HIVE_EMU
                buf[dst[i]] = lut[mine[q * STAGE_SRC_BYTES + 32 * i]];
#else
                const uint32_t byte = mine[q * STAGE_SRC_BYTES + 32 * i];
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(lut_s + byte * 16u));
                buf[dst[i]] = v;
#endif
            }
        if (q == TURN_STAGE) {                                  // plane 31 holds the turn number, not bits
            __syncwarp();                                       // (its chunks were written by other lanes above)
            if (lane < 18) buf[TURN_CHUNK0 + lane] = turn4;
        }
        fence_proxy_async_smem();                               // generic-proxy writes -> visible to the bulk-copy engine
        __syncwarp();
        if (lane == 0) bulk_store_s2g(reinterpret_cast<uint8_t*>(out) + q * STAGE_BYTES, buf, STAGE_BYTES);
    }
}

// splitmix64 -- counter-based action choice of SURVEY 8d Config 2
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// index of the kth (0-based) set bit of a 50-word mask held in memory readable by the warp
__device__ __forceinline__ int select_kth_action(const uint32_t* words, int lane, int kth) {
    uint32_t w0 = 0, w1 = 0;
    if (lane < 25) { w0 = words[2 * lane]; w1 = words[2 * lane + 1]; }
    int c = __popc(w0) + __popc(w1), incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    const unsigned hit = __ballot_sync(FULL, incl > kth);
    const int owner = __ffs(hit) - 1;
    int ans = -1;
    if (lane == owner) {
        // r-th set bit of the 64-bit mask w1:w0 by halving on popcounts (a clear-lowest-bit loop here was compiled
        // into a 64-bit division for its trip count)
        int r = kth - (incl - c), pos = 0;
        uint32_t w = w0;
        const int c0 = __popc(w0);
        if (r >= c0) { r -= c0; w = w1; pos = 32; }
#pragma unroll
        for (int h = 16; h > 0; h >>= 1) {
            const int t = __popc(w & ((1u << h) - 1u));
            if (r >= t) { r -= t; w >>= h; pos += h; }
        }
        ans = lane * 64 + pos;
    }
    return __shfl_sync(FULL, ans, owner < 0 ? 0 : owner);
}

}  // namespace hive
