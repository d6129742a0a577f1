// hive_core.cuh -- device building blocks of the Hive position evaluator for sm_100a.
//
// Boards are whole 144-cell bit strings held in registers (cell = q*12+r, 5 x u32), so translating a
// board along one of the six torus directions (tile.py:111-123) is a handful of funnel shifts and no
// shuffles.  Pieces are numbered in the reference's order: white Q,B0,B1,S0,S1,G0,G1,G2,A0,A1,A2, then
// the same for black (inventory_frame.py:47-99 / env_hive.py:71-87).  This header holds the pieces the
// step kernel (hive_step_kernel.cuh) is assembled from: the 384-byte game record, board algebra, the
// slide relation, the per-piece searches (one-hive flood, Ant flood over slide gates, exact 3-step
// Spider walk, Grasshopper line flood, Queen/Beetle ring logic) and the TMA plane store.
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
// per-cell neighbour ranks (3 bits per direction), per-cell neighbour cells (six bytes in two words) and per-cell
// neighbourhood boards (the six neighbours as a 144-bit board)
constexpr int GEO_HOP = 0, GEO_RANK = 144 * 5, GEO_NBR = GEO_RANK + 144, GEO_NBRMASK = GEO_NBR + 288, GEO_WORDS = GEO_NBRMASK + 144 * 5;
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

__device__ __forceinline__ int move_class(int type) {
    return type == T_ANT ? 0 : type == T_HOPPER ? 1 : type == T_SPIDER ? 2 : 3;
}
__device__ __forceinline__ int piece_type_of(int k) {
    return (k == 0) ? T_QUEEN : (k <= 2) ? T_BEETLE : (k <= 4) ? T_SPIDER : (k <= 7) ? T_HOPPER : T_ANT;
}

// info word of a piece, written by the analyse phase of the step kernel:
//   cell | height<<8 | top<<12 | level<<13 | ring<<16   (ring = occupancy of the six neighbours, direction order d0..d5)

// search, part 1: one thread, one one-hive flood (move_checker.py:58-83 / env_hive.py:509-530):
// lift the top piece and test that the rest of the hive stays connected.  Returns true if pinned.
__device__ __forceinline__ bool eval_flood(uint32_t info, BB occp) {
    const int cell = info & 0xFF;
    const uint32_t ring = (info >> 16) & 63u;
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

// search, part 2: one thread, the move set of one unpinned top piece `p` whose turn gates are open.
// `hop_lines` = 144x5 u32 table of is_straight_line masks (move_checker.py:249-265).
__device__ __forceinline__ BB eval_moves(uint32_t info, const BB& occ, int p, const uint32_t* __restrict__ hop_lines) {
    const int cell = info & 0xFF;
    const int height = (info >> 8) & 0xF;
    const uint32_t ring = (info >> 16) & 63u;
    const int type = piece_type_of(p >= 11 ? p - 11 : p);

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
        BB r = bb_zero();
#pragma unroll
        for (int i = 0; i < 6; i++)
            if ((ok >> i) & 1u) {
                const int c = cell_nbr(cell, i); const uint32_t bit = 1u << (c & 31); const int wi = c >> 5;
#pragma unroll
                for (int j = 0; j < 5; j++) r.w[j] |= (wi == j) ? bit : 0u;
            }
        return r;
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
            // the first steps the gates allow (usually two: along the hive either way), then one trip of a rolled loop per
            // first step -- a lane walks only its own paths, and the warp as many trips as its busiest lane
            uint32_t starts = 0, nb03 = 0, nb45 = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const int c1 = cell_nbr(cell, i);
                if (bb_test(sl.g[i], c1)) starts |= 1u << i;
                if (i < 4) nb03 |= (uint32_t)c1 << (8 * i); else nb45 |= (uint32_t)c1 << (8 * (i - 4));
            }
            while (starts) {
                const int i = __ffs(starts) - 1;
                starts &= starts - 1;
                const int c1 = (int)(((i < 4 ? nb03 : nb45) >> (8 * (i & 3))) & 0xFFu);
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
    return mv;
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
// A stage covers SP planes = SP*20 source bytes; lane l looks at source bytes l, l+32, ... of the stage.  Which plane
// and byte that is, whether it is one of the two padding bytes, and the 16-byte chunk of the staging buffer its eight
// bf16 values go to are the same for every stage and every game, so they are worked out once per call (offsets kept
// in registers) and the inner loop is: byte load at an immediate offset, LUT row load, 16-byte store.
// `lut_s`: the LUT's shared-window address, read back from memory by the caller so that it stays in a register (ptxas
// otherwise re-materialises the window base -- S2R + LEA -- in front of every LUT load).
// SP = planes per stage (the ring holds STAGE_BUFS stages of SP*288 bytes per warp).
template <int SP>
__device__ __forceinline__ void store_planes_bulk_t(const uint8_t* bytes, const uint4* lut, uint32_t lut_s, uint4* stage, int lane, int turn,
                                                    uint16_t* __restrict__ out) {
    constexpr int CHUNKS = SP * 18, BYTES = CHUNKS * 16, NST = N_PLANE / SP, SRC_BYTES = SP * 20, ROUNDS = (SRC_BYTES + 31) / 32;
    constexpr int T_STAGE = 31 / SP, T_CHUNK0 = (31 % SP) * 18;      // plane 31 = the turn number
    static_assert(N_PLANE % SP == 0 && NST % STAGE_BUFS == 0, "stages tile the 56 planes");
    const uint32_t tb = __float_as_uint((float)turn) >> 16;     // bf16(turn): turn <= 255 is exact
    const uint32_t tt = tb | (tb << 16);
    const uint4 turn4 = make_uint4(tt, tt, tt, tt);
    int dst[ROUNDS];                                            // chunk of the stage buffer, -1: padding byte / past the stage
#pragma unroll
    for (int i = 0; i < ROUNDS; i++) {
        const int sb = lane + 32 * i, pl = sb / 20, j = sb - 20 * pl;
        dst[i] = (sb < SRC_BYTES && j < 18) ? pl * 18 + j : -1;
    }
    const uint8_t* mine = bytes + lane;
#ifdef HIVE_EMU
    (void)lut_s;
#endif
#pragma unroll
    for (int q = 0; q < NST; q++) {
        uint4* buf = stage + (q % STAGE_BUFS) * CHUNKS;
        // the stage that used this buffer last (of this game or of the warp's previous game) has left shared memory
        if (lane == 0) bulk_wait_read<STAGE_BUFS - 1>();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < ROUNDS; i++)
            if (dst[i] >= 0) {
#ifdef HIVE_EMU
                buf[dst[i]] = lut[mine[q * SRC_BYTES + 32 * i]];
#else
                const uint32_t byte = mine[q * SRC_BYTES + 32 * i];
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(lut_s + byte * 16u));
                buf[dst[i]] = v;
#endif
            }
        if (q == T_STAGE) {                                     // plane 31 holds the turn number, not bits
            __syncwarp();                                       // (its chunks were written by other lanes above)
            if (lane < 18) buf[T_CHUNK0 + lane] = turn4;
        }
        fence_proxy_async_smem();                               // generic-proxy writes -> visible to the bulk-copy engine
        __syncwarp();
        if (lane == 0) bulk_store_s2g(reinterpret_cast<uint8_t*>(out) + q * BYTES, buf, BYTES);
    }
}
__device__ __forceinline__ void store_planes_bulk(const uint8_t* bytes, const uint4* lut, uint32_t lut_s, uint4* stage, int lane, int turn,
                                                  uint16_t* __restrict__ out) {
    store_planes_bulk_t<STAGE_PLANES>(bytes, lut, lut_s, stage, lane, turn, out);
}

// splitmix64 -- counter-based action choice of SURVEY 8d Config 2
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// A game's 50-word legal mask, fetched with all loads in flight at once (rows are 200 B: 8-byte aligned), and the
// index of its kth (0-based) set bit (one thread).
struct LegalRow { uint2 v[LEGAL_WORDS / 2]; };
__device__ __forceinline__ LegalRow load_legal_row(const uint32_t* words) {
    LegalRow r;
    const uint2* w2 = reinterpret_cast<const uint2*>(words);
#pragma unroll
    for (int i = 0; i < LEGAL_WORDS / 2; i++) r.v[i] = w2[i];
    return r;
}
__device__ __forceinline__ int kth_legal_action(const LegalRow& r, int kth) {
    uint32_t lo = 0, hi = 0;
    int idx = -1, acc = 0, rem = 0;
#pragma unroll
    for (int i = 0; i < LEGAL_WORDS / 2; i++) {
        const int c = __popc(r.v[i].x) + __popc(r.v[i].y);
        if (idx < 0 && kth < acc + c) { idx = i; lo = r.v[i].x; hi = r.v[i].y; rem = kth - acc; }
        acc += c;
    }
    if (idx < 0) return -1;
    // rem-th set bit of the 64-bit mask hi:lo by halving on popcounts
    uint32_t w = lo; int pos = 0;
    const int c0 = __popc(lo);
    if (rem >= c0) { rem -= c0; w = hi; pos = 32; }
#pragma unroll
    for (int h = 16; h > 0; h >>= 1) {
        const int t = __popc(w & ((1u << h) - 1u));
        if (rem >= t) { rem -= t; w >>= h; pos += h; }
    }
    return idx * 64 + pos;
}

}  // namespace hive
