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

enum Op { OP_RESET = 0, OP_STEP = 1, OP_EVAL = 2, OP_RANDOM = 3, OP_INIT = 4 };   // INIT = first reset, zeroes the counters

struct EnvArgs {
    GameRec* recs;
    uint32_t* legal;       // [n][50]
    int32_t* count;        // [n]
    uint32_t* status;      // [n] turn | winner<<8 | done<<16
    uint16_t* planes;      // [n][56*144] bf16
    uint32_t* bits;        // [n][BITS_WORDS] bit planes, step kernel -> plane-store kernel (this step's buffer of two)
    uint32_t* shadow;      // [n][BITS_WORDS] the bit planes the planes arena holds right now (delta plane store)
    uint8_t* lists;        // [ceil(n/SG)][LIST_BLOCK_BYTES] compact legal lists of the host-driven path, or null (see LIST_* below)
    const int32_t* actions;
    const uint8_t* mask;
    int32_t* chosen;
    const uint32_t* hop_lines;   // GEO_* tables (hive_core.cuh): hop lines, neighbour ranks, neighbour cells, neighbourhood boards
    uint64_t seed;
    int n, op, max_turn, auto_reset;
    int g_offset, n_total;   // this launch covers games [g_offset, g_offset+n) of a batch of n_total (pointers are pre-offset)
    int stagger_ns, stagger_div;   // rollout kernel: start offset between the CTAs of one SM, CTAs per wave (= #SMs)
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

// ---- optional per-phase clocks of the step kernel (builds with -DHIVE_PHASE_CLOCKS only; profiles/phase_probe.py):
// thread 0 of every CTA adds the SM clocks it spent between the phase barriers to g_phase_clk[phase], CTAs to [7]
#ifdef HIVE_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[8];
#define HIVE_PHASE_MARK(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_clk[i], (unsigned long long)(t_ - phase_t_)); phase_t_ = t_; } } while (0)
#define HIVE_PHASE_BEGIN() long long phase_t_ = clock64(); if (threadIdx.x == 0) atomicAdd(&g_phase_clk[7], 1ull)
#else
#define HIVE_PHASE_MARK(i)
#define HIVE_PHASE_BEGIN()
#endif

// ==========================================================================================================
// The step kernel: GamePlay.move() of 32 games per CTA (env_hive.py:99-171), in five phases separated by block
// barriers.  In the analyse and encode phases LANE <-> GAME (a warp's 32 lanes work on 32 different games, so board
// algebra, hashing, the legal-mask scatter ... are plain per-thread code without a single shuffle), and the warps of
// the CTA split the work by PIECE (analyse) or by OUTPUT (encode: legal mask halves, plane groups, history,
// mobility planes).  In the search phases THREAD <-> QUEUED PIECE: the pieces that need a one-hive flood or a move
// search are collected in CTA-local queues, the move queues one per piece class, so that a warp runs 32 Ant floods,
// or 32 Spider walks ... of different games side by side.  Per-game fields live in shared memory as [field][lane]
// (bank = lane: conflict-free for any per-lane index).
//   1a (warp 0)   decode the operation, pick / apply the action
//   1b (all)      per piece: stack height, top, occupancy boards; then ring occupancy, hive graph, turn gates, queues
//                 (the last warp also derives the placements)
//   2  (all)      one-hive floods; survivors join the move queues
//   3  (all)      move searches by class; move sets -> shared rows
//   4  (all)      legal mask (shared), 56 bit planes -> a.bits, history push
//   5  (warps 0,1) legal mask, count, status, record header out
#ifndef HIVE_STEP_WARPS
#define HIVE_STEP_WARPS 8
#endif
constexpr int SG = 32;                                   // games per CTA
constexpr int SG_GAMES = SG;
constexpr int SW = HIVE_STEP_WARPS, STEP_THREADS = SW * 32;
static_assert(SW >= 2 && SW <= 16, "warps per CTA of the step kernel");
#ifndef HIVE_STEP_MIN_CTAS
#define HIVE_STEP_MIN_CTAS (768 / (HIVE_STEP_WARPS * 32))
#endif

// Compact legal lists (EnvArgs::lists; what the host-driven loop downloads instead of the 198-byte masks: 96 B per game).
// One block of LIST_BLOCK_BYTES per group of SG games: SG headers of 12 bytes, then the games' action ids packed back to back,
// ascending (GamePlay.actions() is an ascending list, env_hive.py:182,301-304), ONE BYTE per action: the low 8 bits of the id.
// Header of a game: u16 offset of its ids inside the ids area; u8 cum[7], cum[p] = number of its actions with id < 256 (p+1)
// (cum[6] = the count), so action k is 256 p + ids[offset + k] with p the first page whose cum exceeds k; u8 flags, bit 0 =
// the group's lists did not fit (more than LIST_IDS_CAP actions in the group or 256+ in one game: use the mask).
constexpr int LIST_HDR_BYTES = 12, LIST_BLOCK_BYTES = 3072, LIST_IDS_CAP = LIST_BLOCK_BYTES - SG_GAMES * LIST_HDR_BYTES;

struct __align__(16) StepShared {
    union {
        uint32_t legal[LEGAL_WORDS][SG];     // phases 3..5 (zeroed at the start of phase 3)
        struct {                             // phases 1b..2: the hive as a graph of occupied cells, node = top piece of a cell
            uint32_t cmap[36][SG];           // byte c: the top piece standing on cell c
            uint32_t adj[N_PIECE][SG];       // per top piece: the top pieces of its occupied neighbour cells
        } graph;
    };
    uint32_t rows[N_PIECE][5][SG];   // move set of every searched piece (own: its action list; opponent: its mobility set)
    uint32_t info[N_PIECE][SG];      // cell | height<<8 | top<<12 | level<<13 | ring<<16
    uint32_t pc[11][SG];             // record bytes 0..43 after the action: cell[22], level[22]
    uint32_t occ[5][SG], own[5][SG], opp[5][SG], place[5][SG];
    uint32_t head[SG];               // turn | cq_w<<8 | cq_b<<16
    uint32_t flags[SG];              // live | push_history<<1 | prev_winner<<8
    uint32_t pin[SG];                // pinned pieces (lifting them breaks the hive)
    uint32_t placeable[SG];          // in-hand pieces of the side to move that may be placed on `place`
    uint32_t nonempty[SG];           // searched pieces with a non-empty move set
    uint32_t nlegal[SG], episode[SG], steps[SG];
    uint32_t n_flood, n_mv[4];       // queue fill: floods; move classes 0 Ant, 1 Grasshopper, 2 Spider, 3 Queen/Beetle
    uint32_t any_live, next_task, next_pre;
    uint16_t q_flood[SG * N_PIECE];  // item = slot | piece<<5 | wants_moves<<10
    uint16_t q_mv[4][SG * 6];
};

__device__ __forceinline__ BB bb_onehot(int c) { return bb_bit(c); }   // HAND (255) selects no word: empty board

// append `item` to a CTA queue for every lane with `cond` (one shared-memory atomic per warp)
__device__ __forceinline__ void queue_push(uint32_t* counter, uint16_t* q, bool cond, uint32_t item, int lane) {
    const unsigned m = __ballot_sync(FULL, cond);
    if (m) {
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
        base = __shfl_sync(FULL, base, leader);
        if (cond) q[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)item;
    }
}

// ---- phase 1a (one warp, lane <-> game)
__device__ __forceinline__ void step_prologue(StepShared& s, const EnvArgs& a, int lane, int g) {
#ifdef HIVE_PROLOGUE_SPLIT
    const long long pt0_ = clock64(); long long pta_ = pt0_, ptb_ = pt0_;
#endif
    bool live = g < a.n;
    uint32_t pc[11];
#pragma unroll
    for (int i = 0; i < 11; i++) pc[i] = 0;
    int turn = 1, winner = 0;
    uint32_t episode = 0, steps = 0;
    bool push = false;
    if (live) {
        GameRec* rec = a.recs + g;
        const uint4* r4 = reinterpret_cast<const uint4*>(rec);
        const uint4 v0 = r4[0], v1 = r4[1], v2 = r4[2], v3 = r4[3];
        LegalRow lrow;                                          // the random policy picks from the current legal mask: fetched with the record
        if (a.op == OP_RANDOM) lrow = load_legal_row(a.legal + (size_t)g * LEGAL_WORDS);
        pc[0] = v0.x; pc[1] = v0.y; pc[2] = v0.z; pc[3] = v0.w; pc[4] = v1.x; pc[5] = v1.y; pc[6] = v1.z; pc[7] = v1.w;
        pc[8] = v2.x; pc[9] = v2.y; pc[10] = v2.z;
        // header words: [11] = turn|winner|done|flags, [12] episode, [13] steps, [14] n_legal
        const uint32_t h11 = v2.w;
        turn = h11 & 0xFF; winner = (h11 >> 8) & 0xFF;
        const int done = (h11 >> 16) & 0xFF;
        episode = v3.x; steps = v3.y;
        uint32_t n_legal_prev = v3.z;
#ifdef HIVE_PROLOGUE_SPLIT
        if (a.op == OP_RANDOM) asm volatile("" : "+r"(lrow.v[24].y), "+r"(lrow.v[0].x));
        asm volatile("" : "+r"(n_legal_prev), "+r"(turn));
        pta_ = clock64();
#endif

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
            push = (h11 >> 24) & 1u;          // "push history when this position is evaluated" (set by the search)
        } else {   // OP_RANDOM
            if (done || turn >= a.max_turn) {
                if (!a.auto_reset) live = false; else do_reset = true;
            } else if (n_legal_prev == 0) {
                action = -1;
            } else {
                const uint64_t gid = (uint64_t)(g + a.g_offset) + (uint64_t)a.n_total * episode;
                const uint64_t x = splitmix64(a.seed ^ (gid << 32) ^ (uint64_t)turn);
                action = kth_legal_action(lrow, (int)(x % n_legal_prev));
            }
            if (a.chosen) a.chosen[g] = (do_reset || !live) ? HIVE_NOOP : action;
        }
#ifdef HIVE_PROLOGUE_SPLIT
        asm volatile("" : "+r"(action));
        ptb_ = clock64();
#endif
        if (live) {
            if (do_reset) {                                     // GamePlay.new_game, env_hive.py:61-97
#pragma unroll
                for (int i = 0; i < 11; i++) pc[i] = i < 5 ? 0xFFFFFFFFu : i == 5 ? 0x0000FFFFu : 0u;   // 22 x HAND, levels 0
                turn = 1; winner = 0; episode++;
                uint4* h4 = reinterpret_cast<uint4*>(rec->hist);
#pragma unroll
                for (int i = 0; i < 20; i++) h4[i] = make_uint4(0u, 0u, 0u, 0u);
                push = true;                                    // add_history starts True (env_hive.py:51)
            } else if (action >= 0) {                           // env_hive.py:105-148
                const int side = (turn & 1) ? 0 : 1;
                const int k = action % 11, end = action / 11, p = side * 11 + k;
                int h_end = 0;                                  // level = len(end_tile.pieces) before the move
#pragma unroll
                for (int q = 0; q < N_PIECE; q++) h_end += (int)(((pc[q >> 2] >> (8 * (q & 3))) & 0xFFu) == (uint32_t)end);
                const int wc = p >> 2, shc = 8 * (p & 3), wl = (22 + p) >> 2, shl = 8 * ((22 + p) & 3);
                const uint32_t mc = 0xFFu << shc, vc = (uint32_t)end << shc, ml = 0xFFu << shl, vl = (uint32_t)h_end << shl;
#pragma unroll
                for (int i = 0; i < 11; i++) {
                    uint32_t w = pc[i];
                    w = wc == i ? (w & ~mc) | vc : w;
                    w = wl == i ? (w & ~ml) | vl : w;
                    pc[i] = w;
                }
                turn++; steps++; push = true;
            } else if (action == -1) {                          // pass, env_hive.py:100-103
                turn++; steps++;
            }
        } else {
            a.bits[(size_t)g * BITS_WORDS + BITS_LIVE] = 0u;    // not evaluated in this launch: the plane store leaves this game's planes alone
        }
    }

    const int cq_w = pc[0] & 0xFF, cq_b = (pc[2] >> 24) & 0xFF;
#pragma unroll
    for (int i = 0; i < 11; i++) s.pc[i][lane] = pc[i];
#pragma unroll
    for (int i = 0; i < 5; i++) { s.occ[i][lane] = 0u; s.own[i][lane] = 0u; s.opp[i][lane] = 0u; }     // filled by step_stacks
    s.head[lane] = (uint32_t)turn | ((uint32_t)cq_w << 8) | ((uint32_t)cq_b << 16);
    s.flags[lane] = (live ? 1u : 0u) | ((uint32_t)push << 1) | ((uint32_t)winner << 8);
    s.pin[lane] = 0; s.nonempty[lane] = 0; s.nlegal[lane] = 0;
    s.episode[lane] = episode; s.steps[lane] = steps;
    const unsigned lv = __ballot_sync(FULL, live);
    if (lane == 0) s.any_live = lv;
#ifdef HIVE_PROLOGUE_SPLIT
    if (lane == 0) atomicAdd(&g_phase_clk[6], (unsigned long long)(pta_ - pt0_) | ((unsigned long long)(ptb_ - pta_) << 32));
#endif
}

// ---- phase 1b (all warps; lane <-> game, the warps take the pieces in turn): stacks; which piece tops which cell
__device__ __forceinline__ void step_stacks(StepShared& s, int warp, int lane) {
    const bool live = s.flags[lane] & 1u;
    const int side = (s.head[lane] & 1u) ? 0 : 1;           // turn odd: white to move (game_state.py:58-62)
    int bc[4];                                              // cells of the beetles above level 0 (else: no cell)
    {
        const uint32_t w0 = s.pc[0][lane], w3 = s.pc[3][lane], w5 = s.pc[5][lane], w6 = s.pc[6][lane], w8 = s.pc[8][lane];
        const int c1 = (w0 >> 8) & 0xFF, c2 = (w0 >> 16) & 0xFF, c12 = w3 & 0xFF, c13 = (w3 >> 8) & 0xFF;
        const int l1 = (w5 >> 24) & 0xFF, l2 = w6 & 0xFF, l12 = (w8 >> 16) & 0xFF, l13 = (w8 >> 24) & 0xFF;
        bc[0] = (l1 >= 1 && c1 != HAND) ? c1 : 0x100; bc[1] = (l2 >= 1 && c2 != HAND) ? c2 : 0x101;
        bc[2] = (l12 >= 1 && c12 != HAND) ? c12 : 0x102; bc[3] = (l13 >= 1 && c13 != HAND) ? c13 : 0x103;
    }
    for (int p = warp; p < N_PIECE; p += SW) {              // warp-uniform
        const int c = (s.pc[p >> 2][lane] >> (8 * (p & 3))) & 0xFF;
        const int lv = (s.pc[(22 + p) >> 2][lane] >> (8 * ((22 + p) & 3))) & 0xFF;
        const bool on = live && c != HAND;
        // pieces sharing a cell (tile.pieces); top piece <=> level+1 == len (env_hive.py:213)
        const int height = 1 + (int)(bc[0] == c) + (int)(bc[1] == c) + (int)(bc[2] == c) + (int)(bc[3] == c);
        const bool top = on && lv == height - 1;
        s.info[p][lane] = (uint32_t)c | ((uint32_t)height << 8) | ((uint32_t)top << 12) | ((uint32_t)lv << 13);
        if (top) reinterpret_cast<uint8_t*>(&s.graph.cmap[c >> 2][lane])[c & 3] = (uint8_t)p;
        if (on) {                                           // occupancy boards (own / opponent by the side to move), one word each
            const uint32_t bit = 1u << (c & 31);
            atomicOr(&s.occ[c >> 5][lane], bit);
            atomicOr(((p >= 11) == (side == 1)) ? &s.own[c >> 5][lane] : &s.opp[c >> 5][lane], bit);
        }
    }
}

// ---- placements (env_hive.py:217-225; move_checker.py:168-179) -- one warp, lane <-> game, beside the pieces pass: the
// board of cells where the side to move may place a piece and the first in-hand piece of each type that may go there
__device__ __forceinline__ void step_placements(StepShared& s, int lane) {
    const uint32_t hd = s.head[lane];
    const int turn = hd & 0xFF, cq_w = (hd >> 8) & 0xFF, cq_b = (hd >> 16) & 0xFF;
    const int side = (turn & 1) ? 0 : 1;
    const bool wq_on = cq_w != HAND, bq_on = cq_b != HAND;
    uint32_t in_hand = 0;                                   // colour-relative: bit k = own piece k still in hand
    BB top_opp = bb_zero();
    for (int k = 0; k < 11; k++) {
        in_hand |= (uint32_t)((s.info[side * 11 + k][lane] & 0xFFu) == (uint32_t)HAND) << k;
        const uint32_t io = s.info[(1 - side) * 11 + k][lane];
        if ((io >> 12) & 1u) top_opp = top_opp | bb_onehot((int)(io & 0xFFu));
    }
    uint32_t first = 0;
    { uint32_t t;
      t = in_hand & 0x001u; first |= t & (0u - t);
      t = in_hand & 0x006u; first |= t & (0u - t);
      t = in_hand & 0x018u; first |= t & (0u - t);
      t = in_hand & 0x0E0u; first |= t & (0u - t);
      t = in_hand & 0x700u; first |= t & (0u - t); }
    BB occ;
#pragma unroll
    for (int i = 0; i < 5; i++) occ.w[i] = s.occ[i][lane];
    BB place;
    if (turn == 1) place = bb_bit(START_CELL);
    else {
        const BB frontier = bb_andn(bb_nbrs(occ), occ);
        if (turn == 2) place = frontier & bb_bit(TURN2_CELL);
        else {
            place = bb_andn(frontier, bb_nbrs(top_opp));
            if (turn == 7 || turn == 8) {
                const bool ok_q = obeys_queen_by_4(turn, wq_on, bq_on, true, side), ok_n = obeys_queen_by_4(turn, wq_on, bq_on, false, side);
                first = (ok_q ? first & 1u : 0u) | (ok_n ? first & ~1u : 0u);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 5; i++) s.place[i][lane] = place.w[i];
    s.placeable[lane] = first << (11 * side);
}

// ---- phase 1c (all warps; lane <-> game): ring occupancy, hive graph, turn gates; queue floods and move searches
__device__ __forceinline__ void step_pieces(StepShared& s, int warp, int lane, const uint32_t* __restrict__ geo) {
    const uint32_t hd = s.head[lane];
    const int turn = hd & 0xFF, cq_w = (hd >> 8) & 0xFF, cq_b = (hd >> 16) & 0xFF;
    const int side = (turn & 1) ? 0 : 1;
    const bool wq_on = cq_w != HAND, bq_on = cq_b != HAND, ownq_on = (side == 0 ? cq_w : cq_b) != HAND;
    uint32_t pin_bits = 0;
    for (int p = warp; p < N_PIECE; p += SW) {              // warp-uniform
        const int color = p >= 11 ? 1 : 0, k = p - 11 * color, type = piece_type_of(k);
        const uint32_t info = s.info[p][lane];
        const int c = info & 0xFF, height = (info >> 8) & 0xF;
        const bool top = (info >> 12) & 1u;
        const bool on = (s.flags[lane] & 1u) && c != HAND;
        const bool own = color == side;
        uint32_t ring = 0, adj = 0;                         // occupancy of the six neighbours (cells from the GEO_NBR table)
        if (on && (top || type == T_QUEEN)) {
            const uint32_t n03 = __ldg(geo + GEO_NBR + 2 * c), n45 = __ldg(geo + GEO_NBR + 2 * c + 1);
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const uint32_t nb = ((i < 4 ? n03 : n45) >> (8 * (i & 3))) & 0xFFu;
                if ((s.occ[nb >> 5][lane] >> (nb & 31)) & 1u) {
                    ring |= 1u << i;
                    adj |= 1u << reinterpret_cast<const uint8_t*>(&s.graph.cmap[nb >> 2][lane])[nb & 3];
                }
            }
            s.graph.adj[p][lane] = adj;
            s.info[p][lane] = info | (ring << 16);
        }
        // turn gates shared by every candidate of a piece (move_checker.py:38-55)
        bool gate = true;
        if (turn <= 2) gate = false;                                             // no on-board mover can exist / matter
        else if (turn <= 6) gate = ownq_on;                                      // queen_is_on_board: colour by turn parity
        else if (turn <= 8) gate = obeys_queen_by_4(turn, wq_on, bq_on, type == T_QUEEN, color);
        // opponent mobility is only consumed through the own queen's neighbourhood (env_hive.py:459-478)
        const bool wants = top && gate && (own || ownq_on);
        bool pinned_now = false, need_flood = false;
        if (top && height == 1) {
            if (ring == 0) pinned_now = true;                                    // nothing left on the board -> `return False`
            else need_flood = __popc(ring & ~rot6l(ring)) > 1;                   // >1 arc of neighbours: may be an articulation point
        }
        if (pinned_now) pin_bits |= 1u << p;
        const uint32_t item = (uint32_t)lane | ((uint32_t)p << 5) | ((uint32_t)wants << 10);
        queue_push(&s.n_flood, s.q_flood, need_flood, item, lane);
        const int cls = move_class(type);
        queue_push(&s.n_mv[cls], s.q_mv[cls], on && !need_flood && wants && !pinned_now, item, lane);
    }
    if (pin_bits) atomicOr(&s.pin[lane], pin_bits);
}

// one thread, one one-hive test (move_checker.py:58-83 / env_hive.py:509-530) on the graph of occupied cells: lift the
// top piece `p` (alone on its cell) and test that its neighbour cells stay connected.  Returns true if pinned.
__device__ __forceinline__ bool flood_graph(const StepShared& s, int slot, int p) {
    const uint32_t goal = s.graph.adj[p][slot], removed = 1u << p;
    uint32_t reach = goal & (0u - goal), todo = reach;       // reached cells; reached cells whose neighbours are still to be added
    while (todo) {                                           // one flat loop, two nodes per trip: their two loads are in flight together
        if ((reach & goal) == goal) return false;
        const int i0 = __ffs(todo) - 1; todo &= todo - 1;
        const int i1 = todo ? __ffs(todo) - 1 : i0; todo &= todo - 1;      // (todo == 0 stays 0: 0 & 0xFFFFFFFF)
        const uint32_t nw = (s.graph.adj[i0][slot] | s.graph.adj[i1][slot]) & ~(reach | removed);
        reach |= nw; todo |= nw;
    }
    return (reach & goal) != goal;
}

// ---- phase 4 helpers (lane <-> game)
struct EncodeCtx {
    int lane, side, turn, cq_w, cq_b;
    bool live, push;         // every lane runs every task (warp-uniform loops); the lanes without a live game store nothing
    uint32_t has_row, placeable, pin;   // has_row: searched, unpinned pieces with a non-empty move set
    uint32_t* bits;
    uint32_t* legal_out;     // the game's row of the dense legal mask in global memory
    GameRec* rec;
};
__device__ __forceinline__ BB piece_row(const StepShared& s, const EncodeCtx& e, int p) {
    BB r = bb_zero();
    if ((e.has_row >> p) & 1u) {
#pragma unroll
        for (int i = 0; i < 5; i++) r.w[i] = s.rows[p][i][e.lane];
    } else if ((e.placeable >> p) & 1u) {
#pragma unroll
        for (int i = 0; i < 5; i++) r.w[i] = s.place[i][e.lane];
    }
    return r;
}
// four consecutive planes (20 words) -> bits[first_plane*5 ...] as five 16-byte stores (first_plane % 4 == 0)
__device__ __forceinline__ void store_plane_quad(const EncodeCtx& e, int first_plane, const uint32_t (&r)[20]) {
    if (!e.live) return;
    uint4* dst = reinterpret_cast<uint4*>(e.bits + first_plane * 5);
#pragma unroll
    for (int i = 0; i < 5; i++) dst[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
}

// dense legal mask a = cell*11 + k (env_hive.py:287-304) of ALL own pieces for the 32 cells of bitboard word `w`, as whole
// words: cells 32w .. 32w+31 are exactly the actions 352w .. 352w+351 = mask words 11w .. 11w+10 (word 4 holds 16 cells:
// words 44 .. 49), so the tasks (two per bitboard word) write disjoint words with plain stores -- no atomics, no zeroing, no data-dependent
// loop.  A mask word takes two or three consecutive cells of every piece k (bits 11c + k - 32o of word o): those bits of
// the piece's move-set word are spread to stride 11 by one multiplication (x * (1 + 2^10 + 2^20) puts bit b at b + 10b)
// and masked; which cells, which shift and which mask are compile-time constants of (o, k).
// `half`: 0 = mask words 11w .. 11w+5 (and the count), 1 = words 11w+6 .. 11w+10 (not for w = 4).
__device__ __forceinline__ void encode_legal_words(StepShared& s, const EncodeCtx& e, int w, int half) {
    uint32_t r[11];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 11; k++) {
        const int p = e.side * 11 + k;
        r[k] = ((e.has_row >> p) & 1u) ? s.rows[p][w][e.lane] : ((e.placeable >> p) & 1u) ? s.place[w][e.lane] : 0u;
        cnt += __popc(r[k]);
    }
#pragma unroll
    for (int o = 0; o < 11; o++) {
        if ((half >= 0 && (o >= 6) != (half != 0)) || (o >= 6 && w == 4)) continue;      // (warp-uniform; half < 0: all words)
        uint32_t out = 0u;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const int lo = 32 * o - k, c0 = lo <= 0 ? 0 : (lo + 10) / 11, c1x = (32 * o + 31 - k) / 11, c1 = c1x > 31 ? 31 : c1x;
            const int n = c1 - c0 + 1, off = 11 * c0 + k - 32 * o;       // n = 2 or 3 cells; bit of cell c0 in this word
            const uint32_t sel = (1u << n) - 1u, keep = (1u | (n > 1 ? 1u << 11 : 0u) | (n > 2 ? 1u << 22 : 0u)) << off;
            const uint32_t x = (r[k] >> c0) & sel;
            out |= (x * (0x00100401u << off)) & keep;
        }
        s.legal[11 * w + o][e.lane] = out;                  // (the compact lists are built from the shared copy)
        if (e.live) e.legal_out[11 * w + o] = out;
    }
    if (cnt && half <= 0) atomicAdd(&s.nlegal[e.lane], (uint32_t)cnt);
}

// planes 0-11 (which = 0: pieces of the side to move + their union) or 12-23 (which = 1: the opponent's)
__device__ __forceinline__ void encode_piece_planes(const StepShared& s, const EncodeCtx& e, int which) {
    const int base = (which ? 1 - e.side : e.side) * 11;
#pragma unroll
    for (int grp = 0; grp < 3; grp++) {
        uint32_t r[20];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int kk = grp * 4 + j;
            if (kk < 11) {
                const BB b = bb_onehot((int)(s.info[base + kk][e.lane] & 0xFFu));
#pragma unroll
                for (int i = 0; i < 5; i++) r[j * 5 + i] = b.w[i];
            } else {
#pragma unroll
                for (int i = 0; i < 5; i++) r[j * 5 + i] = which ? s.opp[i][e.lane] : s.own[i][e.lane];
            }
        }
        store_plane_quad(e, (which ? 12 : 0) + grp * 4, r);
    }
}

// planes 24-31: beetle levels, occupancy, the turn slot (nothing here depends on the one-hive tests or the move searches)
__device__ __forceinline__ void encode_level_planes(const StepShared& s, const EncodeCtx& e) {
    const int own0 = e.side * 11, opp0 = (1 - e.side) * 11;
    uint32_t r[20];
    BB lvl[6];                                               // 24-26 own beetles at level 2,3,4; 27-29 the opponent's
#pragma unroll
    for (int i = 0; i < 6; i++) lvl[i] = bb_zero();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t info = s.info[(j < 2 ? own0 : opp0) + 1 + (j & 1)][e.lane];
        const int c = info & 0xFF, lv = (info >> 13) & 7;
        if (c != HAND && lv >= 2) {
            const BB b = bb_onehot(c);
#pragma unroll
            for (int t = 0; t < 3; t++) if (lv == t + 2) lvl[(j < 2 ? 0 : 3) + t] = lvl[(j < 2 ? 0 : 3) + t] | b;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 5; i++) r[j * 5 + i] = lvl[j].w[i];
    store_plane_quad(e, 24, r);
#pragma unroll
    for (int i = 0; i < 5; i++) {
        r[i] = lvl[4].w[i]; r[5 + i] = lvl[5].w[i];
        r[10 + i] = s.occ[i][e.lane];                        // 30
        r[15 + i] = i == 0 ? (uint32_t)e.turn : i == 1 ? 1u : 0u;   // slot of plane 31: turn, "evaluated in this launch"
    }
    store_plane_quad(e, 28, r);
}
// planes 32-35: occupied queen neighbours, stuck pieces
__device__ __forceinline__ void encode_misc_planes(const StepShared& s, const EncodeCtx& e, const uint32_t* __restrict__ geo) {
    const int own0 = e.side * 11, opp0 = (1 - e.side) * 11;
    uint32_t r[20];
    // 32 / 33: occupied neighbours of the own / the opponent's queen
    const int q_own = e.side ? e.cq_b : e.cq_w, q_opp = e.side ? e.cq_w : e.cq_b;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        r[i] = q_own != HAND ? (__ldg(geo + GEO_NBRMASK + q_own * 5 + i) & s.occ[i][e.lane]) : 0u;
        r[5 + i] = q_opp != HAND ? (__ldg(geo + GEO_NBRMASK + q_opp * 5 + i) & s.occ[i][e.lane]) : 0u;
    }
    // 34: own pieces without a legal action; 35: opponent pieces covered or pinned
    BB stuck_own = bb_zero(), stuck_opp = bb_zero();
    for (int kk = 0; kk < 11; kk++) {
        const uint32_t io = s.info[own0 + kk][e.lane], ip = s.info[opp0 + kk][e.lane];
        const int co = io & 0xFF, cp = ip & 0xFF;
        if (co != HAND) {
            if (!((io >> 12) & 1u) || !((e.has_row >> (own0 + kk)) & 1u)) stuck_own = stuck_own | bb_onehot(co);
        }
        if (cp != HAND && (!((ip >> 12) & 1u) || ((e.pin >> (opp0 + kk)) & 1u))) stuck_opp = stuck_opp | bb_onehot(cp);
    }
#pragma unroll
    for (int i = 0; i < 5; i++) { r[10 + i] = stuck_own.w[i]; r[15 + i] = stuck_opp.w[i]; }
    store_plane_quad(e, 32, r);
}

// planes 36-43: the 4-step history of the side to move (env_hive.py:431-445), then the push of this position
__device__ __forceinline__ void encode_history(const StepShared& s, const EncodeCtx& e) {
    if (!e.live) return;
    uint4* h4 = reinterpret_cast<uint4*>(&e.rec->hist[e.side][0][0][0]);     // 40 words = 10 x 16 B
    uint32_t h[40];
#pragma unroll
    for (int i = 0; i < 10; i++) { const uint4 v = h4[i]; h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = v.z; h[4 * i + 3] = v.w; }
    uint4* dst = reinterpret_cast<uint4*>(e.bits + 36 * 5);
#pragma unroll
    for (int i = 0; i < 10; i++) dst[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
    if (e.push) {                                            // only after a real move / at reset: ages 0..2 -> 1..3
        uint32_t nh[40];
#pragma unroll
        for (int i = 0; i < 5; i++) { nh[i] = s.own[i][e.lane]; nh[5 + i] = s.opp[i][e.lane]; }
#pragma unroll
        for (int i = 0; i < 30; i++) nh[10 + i] = h[i];
#pragma unroll
        for (int i = 0; i < 10; i++) h4[i] = make_uint4(nh[4 * i], nh[4 * i + 1], nh[4 * i + 2], nh[4 * i + 3]);
    }
}

// N words f[0..N) -> bits[OFF ...) with the widest stores the alignment of OFF allows (compile-time peeling)
template <int OFF, int N>
__device__ __forceinline__ void store_words(uint32_t* bits, const uint32_t (&f)[N]) {
    int i = 0;
    if ((OFF + i) % 2 && i < N) { bits[OFF + i] = f[i]; i += 1; }
    if ((OFF + i) % 4 && i + 1 < N) { *reinterpret_cast<uint2*>(bits + OFF + i) = make_uint2(f[i], f[i + 1]); i += 2; }
#pragma unroll
    for (int j = 0; j < N / 4; j++)
        if (i + 3 < N) { *reinterpret_cast<uint4*>(bits + OFF + i) = make_uint4(f[i], f[i + 1], f[i + 2], f[i + 3]); i += 4; }
    if (i + 1 < N) { *reinterpret_cast<uint2*>(bits + OFF + i) = make_uint2(f[i], f[i + 1]); i += 2; }
    if (i < N) { bits[OFF + i] = f[i]; i += 1; }
}

// three of the planes 44-49 (WHICH = 1: opponent pieces able to reach the j-th empty neighbour of the own queen) or 50-55
// (WHICH = 0: own on-board pieces whose action list holds the j-th empty neighbour of the opponent's queen); j = rank of
// the neighbour in tile.adjacent_tiles order (env_hive.py:448-478; GEO_RANK); HALF = 0: ranks 0-2, 1: ranks 3-5
// (which, half are run-time values: ONE copy of the code, which the four warps that take these tasks run side by side)
__device__ __forceinline__ void encode_mobility(const StepShared& s, const EncodeCtx& e, const uint32_t* __restrict__ geo, int which, int half) {
    const int col = which ? 1 - e.side : e.side, base = col * 11;
    const int qc = col ? e.cq_w : e.cq_b;                    // the queen of the other colour than the pieces
    uint32_t nbr_by_rank[3];
    bool any_empty = false;
#pragma unroll
    for (int r = 0; r < 3; r++) nbr_by_rank[r] = 0xFFFFu;    // no empty neighbour of that rank
    if (qc != HAND) {
        const uint32_t n03 = __ldg(geo + GEO_NBR + 2 * qc), n45 = __ldg(geo + GEO_NBR + 2 * qc + 1), ranks = __ldg(geo + GEO_RANK + qc);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const uint32_t nb = ((i < 4 ? n03 : n45) >> (8 * (i & 3))) & 0xFFu, rank = (ranks >> (3 * i)) & 7u;
            const bool empty = !((s.occ[nb >> 5][e.lane] >> (nb & 31)) & 1u);
#pragma unroll
            for (int r = 0; r < 3; r++) if (empty && rank == (uint32_t)(r + 3 * half)) { nbr_by_rank[r] = nb; any_empty = true; }
        }
    }
    BB pl[3];
#pragma unroll
    for (int r = 0; r < 3; r++) pl[r] = bb_zero();
    uint32_t todo = any_empty ? (e.has_row >> base) & 0x7FFu : 0u;   // the searched pieces of that colour with a non-empty move set
    uint32_t nw[3], nsh[3], nok[3];                          // word and bit of the neighbour in a move-set row; 0 / ~0: there is one
#pragma unroll
    for (int r = 0; r < 3; r++) {
        nok[r] = nbr_by_rank[r] == 0xFFFFu ? 0u : 0xFFFFFFFFu;
        nw[r] = nok[r] ? nbr_by_rank[r] >> 5 : 0u; nsh[r] = nbr_by_rank[r] & 31u;
    }
    while (__ballot_sync(FULL, todo != 0u)) {                // warp-uniform trip count: the lanes stay together
        if (todo) {
            const int kk = __ffs(todo) - 1; todo &= todo - 1;
            const int p = base + kk;                         // (a searched piece with a non-empty set: its row is in s.rows)
            const BB b = bb_onehot((int)(s.info[p][e.lane] & 0xFFu));
#pragma unroll
            for (int r = 0; r < 3; r++) {                    // one row word per rank instead of the whole row
                const uint32_t hit = (0u - ((s.rows[p][nw[r]][e.lane] >> nsh[r]) & 1u)) & nok[r];
#pragma unroll
                for (int i = 0; i < 5; i++) pl[r].w[i] |= b.w[i] & hit;
            }
        }
    }
    uint32_t f[15];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int i = 0; i < 5; i++) f[r * 5 + i] = pl[r].w[i];
    if (e.live) {                                            // planes 44-46, 47-49, 50-52, 53-55: four alignments of the 15 words
        if (which) { if (half) store_words<47 * 5, 15>(e.bits, f); else store_words<44 * 5, 15>(e.bits, f); }
        else { if (half) store_words<53 * 5, 15>(e.bits, f); else store_words<50 * 5, 15>(e.bits, f); }
    }
}

// the outputs that depend neither on the one-hive tests nor on the move searches (piece planes, beetle levels / occupancy /
// turn slot, history + push), taken dynamically by the warps that have finished their share of phase 3 -- a phase as long
// as its longest Ant flood, in which most warps would otherwise wait at the barrier
__device__ __forceinline__ void step_early_outputs(StepShared& s, const EnvArgs& a, int lane, int g) {
    const uint32_t flags = s.flags[lane], hd = s.head[lane];
    EncodeCtx e;
    e.lane = lane; e.turn = hd & 0xFF; e.cq_w = (hd >> 8) & 0xFF; e.cq_b = (hd >> 16) & 0xFF;
    e.side = (e.turn & 1) ? 0 : 1;
    e.live = flags & 1u; e.push = (flags >> 1) & 1u;
    e.pin = 0u; e.has_row = 0u; e.placeable = 0u;           // (not final yet, and not read by these tasks)
    e.bits = a.bits + (size_t)g * BITS_WORDS; e.rec = a.recs + g;
    e.legal_out = nullptr;
    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(&s.next_pre, 1u);
        task = __shfl_sync(FULL, task, 0);
        if (task >= 4u) break;
        switch (task) {
            case 0: encode_level_planes(s, e); break;
            case 1: case 2: encode_piece_planes(s, e, (int)task - 1); break;
            default: encode_history(s, e); break;
        }
    }
}

// phases 1..5 of one step of the CTA's 32 games (all threads; ends without a trailing barrier).  Returns the mask of the
// games evaluated in this step (identical in every thread).
__device__ __forceinline__ unsigned step_phases(StepShared& s, const EnvArgs& a, int blk_index) {      // blk_index: which group of SG games
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blk_index * SG + lane;
    const uint32_t* geo = a.hop_lines;
    HIVE_PHASE_BEGIN();
    if (warp == 0) step_prologue(s, a, lane, g);
    else if (warp == SW - 1 && lane < 7) (&s.n_flood)[lane >= 5 ? lane + 1 : lane] = 0u;      // queue fills, next_task, next_pre (never in the prologue's warp: it must stay converged)
    __syncthreads();
    HIVE_PHASE_MARK(0);
    const unsigned live_mask = s.any_live;
    if (!live_mask && !a.lists) return 0u;                  // (with lists the idle games are listed again: the host reads nothing else)

    step_stacks(s, warp, lane);
    __syncthreads();
    step_pieces(s, warp, lane, geo);
    if (warp == SW - 1) step_placements(s, lane);           // (the last warp has the fewest pieces)
    __syncthreads();
    HIVE_PHASE_MARK(1);

    {   // ---- phase 2: one-hive tests (thread <-> queued piece)
        const int nf = (int)s.n_flood;
        for (int t0 = tid - lane; t0 < nf; t0 += STEP_THREADS) {               // warp-uniform loop
            const int t = t0 + lane;
            bool pushm = false;
            uint32_t item = 0;
            int cls = 0;
            if (t < nf) {
                item = s.q_flood[t];
                const int slot = item & 31, p = (item >> 5) & 31;
                if (flood_graph(s, slot, p)) atomicOr(&s.pin[slot], 1u << p);
                else if ((item >> 10) & 1u) { pushm = true; cls = move_class(piece_type_of(p >= 11 ? p - 11 : p)); }
            }
            // survivors that want a move search join the move queues
#pragma unroll
            for (int c = 0; c < 4; c++) queue_push(&s.n_mv[c], s.q_mv[c], pushm && cls == c, item, lane);
        }
    }
    __syncthreads();
    HIVE_PHASE_MARK(2);
    // (the hive graph is dead: its space becomes the legal mask, every word of which a phase-4 task writes)

    {   // ---- phase 3: move searches, warps homogeneous in piece class (thread <-> queued piece)
        const int n0 = (int)s.n_mv[0], n1 = (int)s.n_mv[1], n2 = (int)s.n_mv[2], n3 = (int)s.n_mv[3];
        const int s1 = (n0 + 31) & ~31, s2 = s1 + ((n1 + 31) & ~31), s3 = s2 + ((n2 + 31) & ~31), total = s3 + n3;
        for (int t = tid; t < total; t += STEP_THREADS) {
            int cls, idx, cnt;
            if (t < s1) { cls = 0; idx = t; cnt = n0; }
            else if (t < s2) { cls = 1; idx = t - s1; cnt = n1; }
            else if (t < s3) { cls = 2; idx = t - s2; cnt = n2; }
            else { cls = 3; idx = t - s3; cnt = n3; }
            if (idx < cnt) {
                const uint32_t item = s.q_mv[cls][idx];
                const int slot = item & 31, p = (item >> 5) & 31;
                BB occ;
#pragma unroll
                for (int i = 0; i < 5; i++) occ.w[i] = s.occ[i][slot];
                const BB mv = eval_moves(s.info[p][slot], occ, p, geo);
                if (bb_any(mv)) {
#pragma unroll
                    for (int i = 0; i < 5; i++) s.rows[p][i][slot] = mv.w[i];
                    atomicOr(&s.nonempty[slot], 1u << p);
                }
            }
        }
        __syncwarp();                                           // (the item loop's trip count differs between the lanes)
        step_early_outputs(s, a, lane, g);
    }
    __syncthreads();
    HIVE_PHASE_MARK(3);

    const uint32_t flags = s.flags[lane];
    const bool live = flags & 1u;
    const uint32_t hd = s.head[lane];
    EncodeCtx e;
    {   // ---- phase 4: outputs (lane <-> game; the warps split the outputs)
        e.lane = lane; e.turn = hd & 0xFF; e.cq_w = (hd >> 8) & 0xFF; e.cq_b = (hd >> 16) & 0xFF;
        e.side = (e.turn & 1) ? 0 : 1;
        e.live = live; e.push = (flags >> 1) & 1u;
        e.pin = s.pin[lane];
        e.has_row = live ? s.nonempty[lane] : 0u;
        e.placeable = live ? s.placeable[lane] : 0u;
        e.bits = a.bits + (size_t)g * BITS_WORDS; e.rec = a.recs + g;
        e.legal_out = a.legal + (size_t)g * LEGAL_WORDS;
    }
    // the outputs that need the move sets are cut into 10 tasks of different length (longest first); a warp takes the next
    // one when it is free (the four that do not were taken by the warps that left phase 3 early: step_early_outputs)
    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(&s.next_task, 1u);
        task = __shfl_sync(FULL, task, 0);
        if (task >= 10u) break;
        {
            switch (task) {
                case 0: case 1: case 2: case 3: encode_mobility(s, e, geo, task < 2u, (int)(task & 1u)); break;
                case 4: encode_misc_planes(s, e, geo); break;
                default: encode_legal_words(s, e, (int)task - 5, -1); break;                               // 5 .. 9: bitboard word w
            }
        }
    }
    __syncthreads();
    HIVE_PHASE_MARK(4);

    if (live) {   // ---- phase 5: count, status, record header (the legal mask went out word by word in phase 4)
        if (warp == 0) {
            const int turn = hd & 0xFF, prev_winner = (flags >> 8) & 0xFF;
            const uint32_t iw = s.info[0][lane], ib = s.info[11][lane];
            // terminal test (move_checker.py:140-165): a queen on the board with six occupied neighbours
            const bool ws = (iw & 0xFFu) != HAND && ((iw >> 16) & 63u) == 63u, bs = (ib & 0xFFu) != HAND && ((ib >> 16) & 63u) == 63u;
            const int done = ws || bs;
            const int winner = (ws && bs) ? prev_winner : ws ? 2 : bs ? 1 : prev_winner;
            const uint32_t n_legal = s.nlegal[lane];
            const uint32_t st = (uint32_t)turn | ((uint32_t)winner << 8) | ((uint32_t)done << 16);
            uint4* r4 = reinterpret_cast<uint4*>(a.recs + g);
            r4[0] = make_uint4(s.pc[0][lane], s.pc[1][lane], s.pc[2][lane], s.pc[3][lane]);
            r4[1] = make_uint4(s.pc[4][lane], s.pc[5][lane], s.pc[6][lane], s.pc[7][lane]);
            r4[2] = make_uint4(s.pc[8][lane], s.pc[9][lane], s.pc[10][lane], st);
            r4[3] = make_uint4(s.episode[lane], s.steps[lane], n_legal, 0u);
            a.count[g] = (int32_t)n_legal;
            a.status[g] = st;
        }
    }
    if (a.lists) {   // ---- compact legal lists (warps SW-4 .. SW-1, beside phase 5; lane <-> game, a warp per range of mask words)
        uint8_t* blk = reinterpret_cast<uint8_t*>(&s.rows[0][0][0]);             // the move rows are dead: headers, then ids
        static_assert(sizeof(s.rows) >= LIST_BLOCK_BYTES && SW >= 6, "list block fits the dead move rows; list warps beside warps 0, 1");
        if (warp >= SW - 4) {
            const int j = warp - (SW - 4);
            const int w0 = j == 0 ? 0 : j == 1 ? 13 : j == 2 ? 25 : 38, w1 = j == 0 ? 13 : j == 1 ? 25 : j == 2 ? 38 : 50;
            const bool in_batch = g < a.n;
            const uint32_t* gw = a.legal + (size_t)g * LEGAL_WORDS;              // an idle game keeps the mask of an earlier step
            const uint32_t cnt = live ? s.nlegal[lane] : (in_batch ? (uint32_t)a.count[g] : 0u);
            uint32_t off = cnt;                                                   // inclusive prefix over the CTA's games
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL, off, d); if (lane >= d) off += o; }
            const uint32_t total = __shfl_sync(FULL, off, 31);
            off -= cnt;
            const bool overflow = total > (uint32_t)LIST_IDS_CAP || __ballot_sync(FULL, cnt > 255u) != 0u;
            uint32_t r = 0;                                                       // rank of the first action of this warp's word range
            for (int i = 0; i < w0; i++) r += (uint32_t)__popc(live ? s.legal[i][lane] : (in_batch ? gw[i] : 0u));
            uint8_t* hdr = blk + lane * LIST_HDR_BYTES;
            uint8_t* ids = blk + SG * LIST_HDR_BYTES + off;
            for (int i = w0; i < w1; i++) {
                uint32_t w = live ? s.legal[i][lane] : (in_batch ? gw[i] : 0u);
                while (__ballot_sync(FULL, w != 0u)) {                            // warp-uniform trip count
                    if (w) {
                        const int b = __ffs(w) - 1; w &= w - 1;
                        if (!overflow) ids[r] = (uint8_t)((i * 32 + b) & 255);
                        r++;
                    }
                }
                if ((i & 7) == 7) hdr[2 + (i >> 3)] = (uint8_t)(r < 255u ? r : 255u);      // end of a 256-id page (8 mask words)
            }
            if (j == 3) hdr[8] = (uint8_t)(r < 255u ? r : 255u);                  // cum[6] = the count
            if (j == 0) { hdr[0] = (uint8_t)(off & 255u); hdr[1] = (uint8_t)(off >> 8); hdr[9] = overflow ? 1u : 0u; hdr[10] = 0; hdr[11] = 0; }
        }
        __syncthreads();
        if (tid < LIST_BLOCK_BYTES / 16)
            reinterpret_cast<uint4*>(a.lists + (size_t)blk_index * LIST_BLOCK_BYTES)[tid] = reinterpret_cast<const uint4*>(blk)[tid];
    }
    HIVE_PHASE_MARK(5);
    return live_mask;
}

__global__ void __launch_bounds__(STEP_THREADS, HIVE_STEP_MIN_CTAS) hive_step_kernel(EnvArgs a) {
    __shared__ StepShared s;
    HIVE_TRACE_SCOPE(0, a);
    step_phases(s, a, blockIdx.x);
}

// ---- the rollout kernel (experiment, HIVE_B200_ROLLOUT_KERNEL=1; measured SLOWER than the per-step kernels, 102 vs 66 us
// per step: a CTA's own plane store -- 512 KB per step through a 4 KB staging ring per warp -- sits on its critical path):
// n_steps consecutive OP_RANDOM steps of the CTA's 32 games in ONE launch.  Games are independent
// and a CTA owns its games' records, legal masks and bit planes, so nothing orders the CTAs: each walks through its steps
// at its own pace (no kernel boundary per step to wait at for the slowest CTA of a launch, no launch gap), and after
// every step it stores its games' planes itself -- phase 6: the bit planes it has just written (L2) are expanded to
// bf16 through the TMA, SG / SW games per warp, with the staging ring laid over the step's shared memory; the bulk
// stores drain to HBM while the next step is computed.
#ifndef HIVE_ROLL_STAGE_PLANES
#define HIVE_ROLL_STAGE_PLANES 7
#endif
constexpr int RSP = HIVE_ROLL_STAGE_PLANES;
struct __align__(128) RollStoreShared {
    uint4 stage[SW][STAGE_BUFS * RSP * 18];
    uint4 lut[256];
    uint32_t planes[SW][BITS_WORDS];
    uint32_t lut_addr;
};
union __align__(128) RollShared {
    StepShared step;
    RollStoreShared store;
};

#ifndef HIVE_ROLL_MIN_CTAS
#define HIVE_ROLL_MIN_CTAS (1024 / (HIVE_STEP_WARPS * 32))     // every CTA of a 16,384-game batch resident at once (4 per SM)
#endif
__global__ void __launch_bounds__(STEP_THREADS, HIVE_ROLL_MIN_CTAS) hive_rollout_kernel(EnvArgs a, int n_steps) {
    __shared__ RollShared sm;
    HIVE_TRACE_SCOPE(5, a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifndef HIVE_EMU
    {   // The CTAs that share an SM (blocks i, i + #SMs, i + 2 #SMs ... of a launch) start a fraction of a step apart, so that
        // one of them stores planes while the others compute: all CTAs do the same work per step and would otherwise
        // stay in lock step, computing together and then queueing for HBM together.
        const unsigned wave = blockIdx.x / a.stagger_div;
        if (a.stagger_ns && wave) {
            unsigned long long t0, t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
            do { __nanosleep(1000); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); } while (t - t0 < (unsigned long long)a.stagger_ns * (wave & 3u));
        }
    }
#endif
    for (int step = 0; step < n_steps; step++) {
        const unsigned live_mask = step_phases(sm.step, a, blockIdx.x);
        __syncthreads();                                        // the step's shared memory is dead; its global writes are visible to the CTA
        if (live_mask) {   // ---- phase 6: bit planes -> bf16 CHW planes [56][144] per game (warp <-> game)
            RollStoreShared& st = sm.store;
            for (int t = tid; t < 256; t += STEP_THREADS) fill_bf16_lut(st.lut, t);
#ifndef HIVE_EMU
            if (tid == 0) st.lut_addr = (uint32_t)__cvta_generic_to_shared(st.lut);
#endif
            __syncthreads();
            const uint32_t lut_s = *reinterpret_cast<volatile uint32_t*>(&st.lut_addr);
            constexpr int NV = BITS_WORDS / 4;                  // 70 uint4 per game
            uint32_t* mine = st.planes[warp];
            uint4 v[3];
            auto fetch = [&](int slot) {
                const uint4* src = reinterpret_cast<const uint4*>(a.bits + (size_t)(blockIdx.x * SG + slot) * BITS_WORDS);
#pragma unroll
                for (int i = 0; i < 3; i++) { const int t = lane + 32 * i; v[i] = src[t < NV ? t : 0]; }
            };
            unsigned todo = 0;                                  // this warp's evaluated games: slots warp, warp + SW, ...
            for (int slot = warp; slot < SG; slot += SW) if ((live_mask >> slot) & 1u) todo |= 1u << slot;
            if (todo) fetch(__ffs(todo) - 1);
            while (todo) {
                const int slot = __ffs(todo) - 1; todo &= todo - 1;
                __syncwarp();                                   // the previous game's planes have been expanded
#pragma unroll
                for (int i = 0; i < 3; i++) { const int t = lane + 32 * i; if (t < NV) reinterpret_cast<uint4*>(mine)[t] = v[i]; }
                __syncwarp();
                const uint32_t turn = mine[BITS_TURN];
                if (todo) fetch(__ffs(todo) - 1);               // the next game's bit planes arrive during the expansion
                store_planes_bulk_t<RSP>(reinterpret_cast<const uint8_t*>(mine), st.lut, lut_s, st.stage[warp], lane, (int)turn,
                                         a.planes + (size_t)(blockIdx.x * SG + slot) * HIVE_PLANES_ELEMS);
            }
            if (lane == 0) bulk_wait_read<0>();                 // the copy engine has read the staging ring (the writes to HBM go on)
            __syncthreads();                                    // before the next step reuses the shared memory
        }
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

// ==========================================================================================================
// The queue-driven rollout (hive_step_random_multi, the default for n_steps >= 2): n_steps consecutive OP_RANDOM steps of
// the whole batch WITHOUT a kernel boundary per step.  Games are independent, so the only order that matters is per
// group of 32 games: step k+1 of a group after its step k.  Two persistent kernels share the SMs:
//   hive_rollout_q_kernel  CTAs take tickets t = 0, 1, 2 ... from a global counter; ticket t is step k = t / G of group
//                          g = t % G (G groups).  A CTA waits until done[g] == k (the group's previous step has left
//                          whichever CTA ran it) and until the store of the group's step k-2 has read the bit-plane buffer
//                          this step overwrites, runs the five phases, and publishes done[g] = k + 1.
//   hive_planes_q_kernel   CTAs take tickets in the same order, wait for done[g] > k, expand the group's bit planes to
//                          bf16 through the TMA staging ring, and publish stored[g] = k + 1.
// Against one kernel launch per step and slice: no launch waits for its slowest CTA, no launch gap, no wave of CTAs
// that finds the SM slots taken -- every resident CTA starts its next group-step the moment it is free.
// MEASURED SLOWER than the per-step graphs (87 vs 65 us per 16,384-game step; 70 us without the store's work; the
// flag-chained launches below: 79 / 60 us) and therefore behind HIVE_B200_ROLLOUT_QUEUE: an SM finishes a group-step
// every ~40 k clocks whether two or three of these CTAs share it (51 k for one CTA alone), the per-step kernels of the
// graph pipeline one every ~28 k -- what bounds the step is the SM's throughput for this code, not idle SM slots.
// Synchronisation: thread 0 spins on an acquire load, a block barrier hands the result to the CTA; the publisher's
// threads have passed a block barrier before thread 0 fences and stores the flag (the split-K semaphore pattern).
struct RollSync {
    unsigned ticket_step, ticket_store, error, pad_[29];   // error: 0, or wait kind << 28 | ticket of the first wait that timed out
    unsigned flags[1];                   // done[G] then stored[G]
};
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
#ifdef HIVE_EMU
    return *p;
#else
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
#ifdef HIVE_EMU
    *p = v;
#else
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
// false: gave up (another CTA has failed, or ~0.3 s without progress: a lost CTA must not hang the GPU); the caller then
// reports through sync->error and every CTA leaves at its next ticket
__device__ __forceinline__ bool spin_until_at_least(const unsigned* p, unsigned want, RollSync* sync, unsigned code) {
#ifdef HIVE_EMU
    (void)p; (void)want; (void)sync; (void)code;   // (the emulator runs the CTAs one after the other, in ticket order)
    return true;
#else
    for (unsigned spins = 0; ld_acquire_u32(p) < want; spins++) {
        __nanosleep(64);
        if ((spins & 1023u) == 1023u && (spins > (1u << 18) || ld_acquire_u32(&sync->error))) {
            atomicCAS(&sync->error, 0u, code);
            return false;
        }
    }
    return true;
#endif
}
// a.bits / bits_alt: the two bit-plane buffers (step k of a launch uses buffer k & 1)
#ifdef HIVE_ROLLQ_MAXNREG
__global__ void __maxnreg__(HIVE_ROLLQ_MAXNREG) hive_rollout_q_kernel(
#else
__global__ void __launch_bounds__(STEP_THREADS, HIVE_STEP_MIN_CTAS) hive_rollout_q_kernel(
#endif
    EnvArgs a, uint32_t* bits_alt, int n_steps, RollSync* sync) {
    __shared__ StepShared s;
    __shared__ unsigned s_ticket;
    HIVE_TRACE_SCOPE(6, a);
    const int tid = threadIdx.x;
    const unsigned G = (unsigned)((a.n + SG - 1) / SG), total = G * (unsigned)n_steps;
    unsigned* done = sync->flags;
    const unsigned* stored = sync->flags + G;
    for (;;) {
        __syncthreads();                                        // the previous group-step of this CTA has left shared memory
        if (tid == 0) {
            const unsigned t = atomicAdd(&sync->ticket_step, 1u);
            s_ticket = t;
            if (t < total) {
                const unsigned g = t % G, k = t / G;
                bool ok = spin_until_at_least(done + g, k, sync, (1u << 28) | t);          // the group's step k-1 is complete (records, legal masks)
                if (ok && k >= 2) ok = spin_until_at_least(stored + g, k - 1, sync, (2u << 28) | t);   // ... and its step k-2 has left this bit-plane buffer
                if (!ok || ld_acquire_u32(&sync->error)) s_ticket = 0xFFFFFFFFu;
            }
        }
        __syncthreads();
        const unsigned t = s_ticket;
        if (t >= total) break;
        const unsigned g = t % G, k = t / G;
        EnvArgs b = a;
        if (k & 1u) b.bits = bits_alt;
        step_phases(s, b, (int)g);
        __syncthreads();                                        // every thread's global writes of this group-step are issued
        if (tid == 0) st_release_u32(done + g, k + 1u);
    }
}
// The same order kept by FLAGS between ordinary launches (hive_step_random_multi, HIVE_B200_ROLLOUT_QUEUE=2): one launch
// per step over all groups, each launched with programmatic stream serialization -- the next step's CTAs become resident
// as soon as every CTA of this step has STARTED (griddepcontrol.launch_dependents is the first thing a CTA does), and a
// CTA of step k waits for done[its group] == k instead of the whole previous grid: the launch gap and the wait for a
// step's slowest CTA are gone, a freed SM slot goes to the next step at once.  The store is hive_planes_q_kernel.
#ifdef HIVE_ROLLQ_MAXNREG
__global__ void __maxnreg__(HIVE_ROLLQ_MAXNREG) hive_step_flow_kernel(EnvArgs a, int k, RollSync* sync) {
#else
__global__ void __launch_bounds__(STEP_THREADS, HIVE_STEP_MIN_CTAS) hive_step_flow_kernel(EnvArgs a, int k, RollSync* sync) {
#endif
    __shared__ StepShared s;
    __shared__ unsigned s_ok;
    HIVE_TRACE_SCOPE(8, a);
#ifndef HIVE_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    const int tid = threadIdx.x;
    const unsigned g = blockIdx.x, G = gridDim.x;
    if (tid == 0) {
        bool ok = spin_until_at_least(sync->flags + g, (unsigned)k, sync, (1u << 28) | ((unsigned)k * G + g));
        if (ok && k >= 2) ok = spin_until_at_least(sync->flags + G + g, (unsigned)k - 1u, sync, (2u << 28) | ((unsigned)k * G + g));
        s_ok = ok && !ld_acquire_u32(&sync->error);
    }
    __syncthreads();
    if (!s_ok) return;
    step_phases(s, a, (int)g);
    __syncthreads();                                            // every thread's global writes of this group-step are issued
    if (tid == 0) st_release_u32(sync->flags + g, (unsigned)k + 1u);
}

#ifndef HIVE_STORE_Q_MIN_CTAS
#define HIVE_STORE_Q_MIN_CTAS (640 / (HIVE_STORE_WARPS * 32))      // <= 96 registers: two of these fit beside two step CTAs of 80 registers
#endif
__global__ void __launch_bounds__(HIVE_STORE_WARPS * 32, HIVE_STORE_Q_MIN_CTAS) hive_planes_q_kernel(EnvArgs a, uint32_t* bits_alt, int n_steps, RollSync* sync) {
    __shared__ uint4 bf16_lut[256];
    __shared__ __align__(16) uint32_t planes_s[HIVE_STORE_WARPS][BITS_WORDS];
    __shared__ unsigned s_ticket;
    __shared__ uint32_t lut_addr;
#ifdef HIVE_EMU
    __shared__ uint4 stage_ring[STORE_STAGE_BYTES / 16];
#else
    extern __shared__ __align__(128) uint4 stage_ring[];
#endif
    HIVE_TRACE_SCOPE(7, a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int t = tid; t < 256; t += HIVE_STORE_WARPS * 32) fill_bf16_lut(bf16_lut, t);
#ifndef HIVE_EMU
    if (tid == 0) lut_addr = (uint32_t)__cvta_generic_to_shared(bf16_lut);
#endif
    __syncthreads();
    const uint32_t lut_s = *reinterpret_cast<volatile uint32_t*>(&lut_addr);
    constexpr int NV = BITS_WORDS / 4;                           // 70 uint4 per game
    const unsigned G = (unsigned)((a.n + SG - 1) / SG), total = G * (unsigned)n_steps;
    const unsigned* done = sync->flags;
    unsigned* stored = sync->flags + G;
    uint32_t* mine = planes_s[warp];
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned t = atomicAdd(&sync->ticket_store, 1u);
            s_ticket = t;
            if (t < total) {                                    // the group's step k has written its bit planes
                if (!spin_until_at_least(done + t % G, t / G + 1u, sync, (3u << 28) | t) || ld_acquire_u32(&sync->error)) s_ticket = 0xFFFFFFFFu;
            }
        }
        __syncthreads();
        const unsigned t = s_ticket;
        if (t >= total) break;
        const unsigned g = t % G, k = t / G;
        const uint32_t* bits = (k & 1u) ? bits_alt : a.bits;
        uint4 v[3];
        auto fetch = [&](int game) {
            const uint4* src = reinterpret_cast<const uint4*>(bits + (size_t)game * BITS_WORDS);
#pragma unroll
            for (int i = 0; i < 3; i++) { const int tt = lane + 32 * i; v[i] = src[tt < NV ? tt : 0]; }
        };
        const int g0 = (int)g * SG;
        if (g0 + warp < a.n) fetch(g0 + warp);
        for (int slot = warp; slot < SG && g0 + slot < a.n; slot += HIVE_STORE_WARPS) {
            const int game = g0 + slot;
            __syncwarp();                                       // the previous game's planes have been expanded
#pragma unroll
            for (int i = 0; i < 3; i++) { const int tt = lane + 32 * i; if (tt < NV) reinterpret_cast<uint4*>(mine)[tt] = v[i]; }
            __syncwarp();
            const uint32_t live = mine[BITS_LIVE], turn = mine[BITS_TURN];
            if (slot + HIVE_STORE_WARPS < SG && game + HIVE_STORE_WARPS < a.n) fetch(game + HIVE_STORE_WARPS);
            if (live)
                store_planes_bulk(reinterpret_cast<const uint8_t*>(mine), bf16_lut, lut_s, stage_ring + warp * (STAGE_BUFS * STAGE_CHUNKS), lane,
                                  (int)turn, a.planes + (size_t)game * HIVE_PLANES_ELEMS);
        }
        __syncthreads();                                        // every warp has read its games' bit planes: the buffer is free
        if (tid == 0) st_release_u32(stored + g, k + 1u);
    }
    if (lane == 0) bulk_wait_read<0>();                         // shared memory must outlive the copy engine's reads
}

// ---- kernel 5, delta form (HIVE_B200_DELTA_STORE=1; measured at the same speed as the full store, see profiles/README.md):
// bit planes -> bf16 CHW planes, writing only the 32-byte sectors of a game's planes whose 16 cells differ from what the
// planes arena already holds.  The arena is private to the handle and a game's planes are only ever written by the store
// kernel, so a copy of the bit planes it stored last (a.shadow, 1,120 B per game) says exactly what is there: consecutive
// positions of a game differ in about a fifth of their 504 sectors (every piece plane holds one bit, the history planes
// shift by a ply, plane 31 is the turn number), so ~3.6 KB per game reach HBM instead of 16 KB -- but as scattered 32-byte
// writes, which cost the HBM as much time as the 16 KB stream (30-35 us per 16,384 games either way).
// Warp <-> game; the next game's rows arrive by cp.async while this game's sectors are written.
#ifndef HIVE_DELTA_WARPS
#define HIVE_DELTA_WARPS 2
#endif
__device__ __forceinline__ void store_sector(uint16_t* dst, uint4 lo, uint4 hi) {       // 16 cells = 32 bytes of bf16
#ifdef HIVE_EMU
    reinterpret_cast<uint4*>(dst)[0] = lo; reinterpret_cast<uint4*>(dst)[1] = hi;
#else
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x),
                 "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");     // one 256-bit streaming store (an output, not the step's working set)
#endif
}
// 16 bytes global -> shared without passing through registers (LDGSTS); groups are per thread
__device__ __forceinline__ void async_copy16(void* dst_smem, const void* src) {
#ifdef HIVE_EMU
    memcpy(dst_smem, src, 16);
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void async_commit() {
#ifndef HIVE_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int PENDING>
__device__ __forceinline__ void async_wait() {
#ifndef HIVE_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#endif
}
__global__ void __launch_bounds__(HIVE_DELTA_WARPS * 32, 1024 / (HIVE_DELTA_WARPS * 32)) hive_planes_delta_kernel(EnvArgs a) {
    __shared__ uint4 bf16_lut[256];
    // per warp, double-buffered: the new bit planes and the shadow of the game in work and of the warp's next game (the
    // next game's 2 x 1,120 B arrive while this game's sectors are written: a warp never waits for a load it has just issued)
    __shared__ __align__(16) uint32_t rows[HIVE_DELTA_WARPS][2][2][BITS_WORDS];
    HIVE_TRACE_SCOPE(4, a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int t = tid; t < 256; t += HIVE_DELTA_WARPS * 32) fill_bf16_lut(bf16_lut, t);
    __syncthreads();
    constexpr int ROUNDS = (BITS_WORDS + 31) / 32;               // 9 words per lane
    constexpr int NV = BITS_WORDS / 4;                           // 70 uint4 per row
    const int g_stride = gridDim.x * HIVE_DELTA_WARPS;
    auto fetch = [&](int g, int b) {
        const uint4* nb = reinterpret_cast<const uint4*>(a.bits + (size_t)g * BITS_WORDS);
        const uint4* sh = reinterpret_cast<const uint4*>(a.shadow + (size_t)g * BITS_WORDS);
        for (int t = lane; t < 2 * NV; t += 32) {
            const bool second = t >= NV;
            async_copy16(reinterpret_cast<uint4*>(rows[warp][b][second ? 1 : 0]) + (second ? t - NV : t), second ? sh + (t - NV) : nb + t);
        }
        async_commit();
    };
    int g = blockIdx.x * HIVE_DELTA_WARPS + warp, b = 0;
    if (g < a.n) fetch(g, 0);
    for (; g < a.n; g += g_stride, b ^= 1) {
        if (g + g_stride < a.n) { fetch(g + g_stride, b ^ 1); async_wait<1>(); } else async_wait<0>();
        __syncwarp();                                            // every lane's copies of this game have landed
        const uint32_t* nv = rows[warp][b][0];
        const uint32_t* ov = rows[warp][b][1];
        const uint32_t live = nv[BITS_LIVE], turn = nv[BITS_TURN], turn_old = ov[BITS_TURN];
        if (live) {                                              // (else: not evaluated in this step, its planes stay as they are)
            uint32_t* sh = a.shadow + (size_t)g * BITS_WORDS;
            const uint32_t tb = __float_as_uint((float)turn) >> 16, tt = tb | (tb << 16);      // bf16(turn): turn <= 255 is exact
            const uint4 turn4 = make_uint4(tt, tt, tt, tt);
            uint16_t* out = a.planes + (size_t)g * HIVE_PLANES_ELEMS;
            // lane <-> sector: a round covers 32 consecutive sectors = 1 KB of the game's planes, so the lanes that store
            // touch at most 8 lines (scattered sectors per lane would cost the load/store unit a pass per lane)
#pragma unroll 4
            for (int j = 0; j < 16; j++) {
                const int sct = lane + 32 * j;                   // 504 sectors: plane = sct / 9, 9 sectors (4.5 words) per plane
                const int pl = (sct * 7282) >> 16, k = sct - 9 * pl, w = pl * 5 + (k >> 1), sft = (k & 1) * 16;
                if (sct < 504) {
                    const uint32_t hb = (nv[w] >> sft) & 0xFFFFu, ho = (ov[w] >> sft) & 0xFFFFu;
                    // plane 31 = the turn number in every cell (its words hold the turn and the live flag, not bits)
                    if (pl == 31 ? turn != turn_old : hb != ho)
                        store_sector(out + sct * 16, pl == 31 ? turn4 : bf16_lut[hb & 0xFFu], pl == 31 ? turn4 : bf16_lut[hb >> 8]);
                }
            }
#pragma unroll
            for (int i = 0; i < ROUNDS; i++) {                   // the shadow follows (lane <-> word, coalesced)
                const int w = lane + 32 * i;
                if (w < BITS_WORDS && nv[w] != ov[w]) sh[w] = nv[w];
            }
        }
        __syncwarp();                                            // the rows of this buffer are free for the game after the next
    }
}

}  // namespace hive
