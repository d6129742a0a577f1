/*
 * hive_b200.h -- C ABI of the B200-native Hive environment / search hot path.
 *
 * The reference (HaiDangDang/hive-Alphazero) has no FFI layer: its hot path is the Python
 * classes GamePlay (hive_engine/env_hive.py:24) and HivePlayer (woker/solo_play.py:69).  These
 * entry points are what a binding for that path replaces; each cites the reference interface it
 * stands in for.  Plain pointers and sizes only; no torch / CUDA types in any signature.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (HIVE_E_*); hive_last_error() gives text.
 *   - a handle owns `n_games` independent games on one GPU; all work is stream-ordered.
 *   - `_host` entry points take HOST buffers and copy inside the call (the reference-facing
 *     plugin path); the others take DEVICE pointers (or expose library-owned device arenas) so a
 *     resident pipeline never leaves HBM.
 *   - action ids are the reference's: a = q*132 + r*11 + piece_idx = cell*11 + piece_idx
 *     (hive_engine/config.py:10, env_hive.py:107-114); -1 = pass (env_hive.py:100-103);
 *     HIVE_NOOP leaves that game untouched, HIVE_RESET starts a new game in that slot.
 *   - there is no CPU fallback: without a CUDA device hive_create fails with HIVE_E_CUDA.
 */
#ifndef HIVE_B200_H
#define HIVE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIVE_ACTION_SPACE 1584      /* hive_engine/config.py:10 */
#define HIVE_STATE_FEATURES 56      /* hive_engine/config.py:21 */
#define HIVE_CELLS 144
#define HIVE_PIECES 22
#define HIVE_LEGAL_U64 25           /* 1584-bit mask */
#define HIVE_HAND 255
#define HIVE_NOOP (-2)
#define HIVE_RESET (-3)                /* as an action: new_game() for that game inside the step */
#define HIVE_PLANES_ELEMS (HIVE_STATE_FEATURES * HIVE_CELLS)   /* bf16 per game, CHW */
#define HIVE_STATE_BYTES 384

#define HIVE_E_ARG (-1)
#define HIVE_E_CUDA (-2)
#define HIVE_E_HANDLE (-3)
#define HIVE_E_SEARCH (-4)          /* a search tree ran out of its arena / depth budget (mcts_policy_host) */

typedef struct hive_env hive_env_t;

const char* hive_last_error(void);
int hive_abi_version(void);

/* GamePlay(HEIGHT_MAP, WIDTH_MAP) x n_games  (env_hive.py:26-97).  `stream` is a cudaStream_t
 * passed as void* (NULL = library-created non-blocking stream).  Games start reset. */
int hive_create(int n_games, int device, void* stream, hive_env_t** out);
int hive_destroy(hive_env_t* h);
int hive_num_games(const hive_env_t* h);
int hive_sync(hive_env_t* h);

/* GamePlay.new_game (env_hive.py:61-97) for games with mask[g] != 0 (NULL = all). Host mask. */
int hive_reset(hive_env_t* h, const uint8_t* game_mask);

/* GamePlay.move(action) (env_hive.py:99-171) for every game: apply, regenerate the legal set of
 * the new side to move and encode its 56 planes. */
int hive_step_host(hive_env_t* h, const int32_t* actions);          /* host int32[n] */
int hive_step(hive_env_t* h, const int32_t* actions_dev);           /* device int32[n] */
/* Fully asynchronous form for pipelined callers: queues the H2D copy of `actions`, the step and the
 * D2H copies of the new legal masks / counts / packed status (any may be NULL) on the handle's
 * stream and returns at once.  All buffers should be pinned and must stay untouched until hive_sync.
 * From the second call with the same set of page-locked buffers on, the whole sequence (upload, the kernels of
 * every slice, downloads) is replayed as one CUDA graph launch; pageable buffers stay on the plain path. */
int hive_step_host_async(hive_env_t* h, const int32_t* actions, uint64_t* mask, int32_t* count, uint32_t* packed_status);
/* The same step, but the new legal sets come down as COMPACT LISTS instead of 198-byte masks: per group of 32 games one
 * block of HIVE_LIST_BLOCK_BYTES -- 32 headers of 12 bytes (u16 offset of the game's ids in the ids area; u8 cum[7], cum[p] =
 * number of its legal actions with id < 256 (p+1), cum[6] = the count; u8 flags, bit 0: the group did not fit, use the mask),
 * then the games' ids back to back, ascending (as GamePlay.actions(), env_hive.py:182,301-304), one byte (id & 255) each:
 * action k of a game = 256 p + ids[offset + k], p = the first page with cum[p] > k.  96 B per game instead of 208 on the
 * wire: the host-driven loop is bound by this download.  lists: page-locked, [ceil(n/32)][3072]. */
#define HIVE_LIST_BLOCK_BYTES 3072
int hive_step_host_async_lists(hive_env_t* h, const int32_t* actions, uint8_t* lists, uint32_t* packed_status);
/* Blocks until the downloads of the last hive_step_host_async have landed in the caller's buffers.  The step's
 * 16 KB/game of planes may still be in flight (they stay in HBM for the network; hive_sync waits for them too), so a
 * host loop that only needs masks / counts / status picks its next actions beside the plane store. */
int hive_wait_results(hive_env_t* h);

/* On-device rollout policy of the benchmark (SURVEY 8d Config 2): a = A[x % len(A)],
 * x = splitmix64(seed ^ game_id<<32 ^ turn), game_id = slot + n_games*episode; pass when A is
 * empty.  With auto_reset, a finished game (game_is_over() or turn >= max_turn, the
 * MAX_GAME_LENGTH cut of self_play.py:162) is reset instead of stepped.  `chosen_dev` (optional,
 * device int32[n]) receives the action taken (HIVE_NOOP for a reset / idle slot). */
int hive_step_random(hive_env_t* h, uint64_t seed, int max_turn, int auto_reset, int32_t* chosen_dev);
/* n_steps consecutive hive_step_random steps replayed as one CUDA graph (the slices of the batch run
 * their step chains independently inside it, so kernels of different steps overlap). */
int hive_step_random_multi(hive_env_t* h, uint64_t seed, int max_turn, int auto_reset, int n_steps);

/* GamePlay.actions() (env_hive.py:182): bit a of mask[g] set <=> action a legal; count = len. */
int hive_legal_host(hive_env_t* h, uint64_t* mask /*[n][25]*/, int32_t* count /*[n]*/);
/* GamePlay.encode_board() (env_hive.py:306-318), bf16, CHW [n][56][144] (the layout
 * api_hive.py:61 feeds the net after transpose(2,0,1)). */
int hive_encode_host(hive_env_t* h, uint16_t* planes_bf16);
/* The same planes as BITS, 1,120 B per game instead of 16 KB: [n][56 planes][5 words], bit c of a plane's 160-bit row =
 * cell c (144 used); plane 31's row holds the turn number in word 0 and in word 1 the flag "this game was evaluated by
 * the last step / reset / load" -- only rows with that flag set describe the game's current planes (a game the last
 * launch left alone keeps its planes, but its row here is stale).  What a sample writer wants: the 55 binary planes of
 * a training row are these bits (self_play.py:160 stores the planes; packing them is ours). */
int hive_bits_host(hive_env_t* h, uint32_t* bits /*[n][280]*/);
/* state.turn, winner (0 none / 1 white / 2 black, settings.py:3-4), game_is_over()
 * (move_checker.py:140-165).  Any pointer may be NULL. */
int hive_status_host(hive_env_t* h, int32_t* turn, int8_t* winner, uint8_t* done);
/* the same three fields packed per game: turn | winner<<8 | done<<16 (one 4-byte D2H per game) */
int hive_status_packed_host(hive_env_t* h, uint32_t* packed);
/* Host-side twin of hive_step_random's action rule for callers that drive hive_step_host from
 * host buffers: actions[g] = HIVE_RESET if the game is over or turn >= max_turn (episodes[g] is
 * then incremented), -1 if count[g]==0, else the (x % count[g])-th set bit of mask[g] with
 * x = splitmix64(seed ^ (g + n*episodes[g])<<32 ^ turn).  Pure host code, no GPU work. */
int hive_host_pick_actions(int n, const uint64_t* mask, const int32_t* count, const uint32_t* packed_status,
                           uint32_t* episodes, uint64_t seed, int max_turn, int32_t* actions);
/* The same rule read from the compact lists of hive_step_host_async_lists (identical actions).  *n_overflow (optional) = games
 * whose group's lists did not fit: their entries of `actions` are left untouched -- fetch the masks (hive_legal_host) and pick
 * those with hive_host_pick_actions. */
int hive_host_pick_actions_lists(int n, const uint8_t* lists, const uint32_t* packed_status, uint32_t* episodes, uint64_t seed,
                                 int max_turn, int32_t* actions, int* n_overflow);
/* ---- the host-driven game loop inside the library ----------------------------------------------------------------
 * The reference's self-play worker is a host loop around GamePlay: actions() -> policy -> move()
 * (woker/self_play.py:54-56,116-193).  hive_host_loop_* runs that loop natively for a whole batch: the batch is cut
 * into n_parts parts (each a hive_env of its own with page-locked staging buffers; every step of a part is ONE CUDA
 * graph launch: H2D actions -> step -> D2H legal masks / counts / status), and n_threads driver threads walk over
 * their parts: wait for the part's downloads, call the policy, launch the part's next step -- while the GPU steps the
 * other parts.  policy == NULL: the host twin of the on-device random policy (hive_host_pick_actions, per-part seed
 * seed + 77*(part+1)).  A policy gets the part's host buffers (mask [n][25] u64, count [n], status [n] packed
 * turn | winner<<8 | done<<16) and writes actions[n] (an action id, -1 pass, HIVE_RESET, HIVE_NOOP); it is called
 * concurrently for different parts from different threads. */
typedef struct hive_host_loop hive_host_loop_t;
typedef void (*hive_policy_fn)(void* user, int part, int first_game, int n, const uint64_t* mask, const int32_t* count,
                               const uint32_t* packed_status, int32_t* actions);
int hive_host_loop_create(int n_games, int device, int n_parts, int n_threads, hive_host_loop_t** out);
int hive_host_loop_destroy(hive_host_loop_t* l);
int hive_host_loop_parts(const hive_host_loop_t* l);
int hive_host_loop_threads(const hive_host_loop_t* l);
/* the environment of one part (its games are [first_game, first_game + hive_num_games) of the batch) */
hive_env_t* hive_host_loop_part(hive_host_loop_t* l, int part, int* first_game);
/* n_steps steps of every part; returns when all results and planes have landed.  seconds: wall time of the loop;
 * policy_seconds / wait_seconds: mean time a driver thread spent inside the policy / waiting for downloads. */
int hive_host_loop_run(hive_host_loop_t* l, int n_steps, uint64_t seed, int max_turn, hive_policy_fn policy, void* user,
                       double* seconds, double* policy_seconds, double* wait_seconds);
/* sum of the env-step counters of all parts (difference across a run = the steps it made) */
long long hive_host_loop_env_steps(hive_host_loop_t* l);

/* counters: env steps and episodes per slot */
int hive_counters_host(hive_env_t* h, uint32_t* steps, uint32_t* episodes);

/* GamePlay.state_key (env_hive.py:150-168) of one game, NUL-terminated; returns length. */
int hive_state_key(hive_env_t* h, int game, char* buf, int buflen);

/* position injection / extraction: turn + per-piece cell (HIVE_HAND = inventory) and stack
 * level, white pieces 0..10 then black, order Q,B0,B1,S0,S1,G0,G1,G2,A0,A1,A2
 * (env_hive.py:71-87).  History is cleared by load. */
int hive_load_state(hive_env_t* h, int game, int turn, const uint8_t* cells, const uint8_t* levels);
int hive_dump_state(hive_env_t* h, int game, int32_t* turn, uint8_t* cells, uint8_t* levels);
/* the raw HIVE_STATE_BYTES record of one game (cells, levels, turn / winner / done, counters, the 4-step plane
 * history of both sides: history_white / history_black of env_hive.py:38-39,436-445) */
int hive_record_host(hive_env_t* h, int game, void* rec384);
/* whole-record snapshot (HIVE_STATE_BYTES each) -- the deepcopy(env) of solo_play.py:158 */
int hive_copy_state(hive_env_t* dst, int dst_game, hive_env_t* src, int src_game);

/* library-owned device arenas (valid until hive_destroy): */
void* hive_dev_state(hive_env_t* h);     /* [n] x 384-byte records                    */
void* hive_dev_legal(hive_env_t* h);     /* [n][25] uint64                            */
void* hive_dev_count(hive_env_t* h);     /* [n] int32                                 */
void* hive_dev_status(hive_env_t* h);    /* [n] uint32 turn | winner<<8 | done<<16    */
void* hive_dev_planes(hive_env_t* h);    /* [n][56][144] bf16                         */

/* launches issued by this handle since creation (bench.py's gpu_launches) */
long long hive_launch_count(const hive_env_t* h);
/* last kernel's device time in ms measured with events on the handle's stream (0 if timing is
 * off); hive_set_timing(h,1) enables per-launch events. */
/* one rollout step with events around its two kernels: ms[2] = step kernel, plane store */
int hive_profile_step(hive_env_t* h, uint64_t seed, int max_turn, float* ms);
int hive_set_timing(hive_env_t* h, int on);
/* Roofline aid: GB/s of a write-only stream (16-byte stores, nothing read) over this batch's planes arena,
 * averaged over `reps` launches.  Overwrites the planes arena with a probe pattern. */
int hive_probe_write_stream(hive_env_t* h, int reps, double* gbs);
float hive_last_kernel_ms(hive_env_t* h);


/* ------------------------------------------------------------------------------------------
 * Batched PUCT search -- HivePlayer (woker/solo_play.py:69-385), sequential mode, one tree per
 * game of `env`.  Bit-exact visit counts given identical network outputs and root noise.
 * A search runs in waves:
 *     mcts_begin(m, mask);                                    // HivePlayer.reset + roots = env states
 *     for (;;) { mcts_descend(m, &pending); if (!pending) break;   // search_my_move down to a new position
 *                <network on mcts_dev_leaf_planes -> mcts_dev_leaf_policy / mcts_dev_leaf_value>
 *                mcts_expand(m); }                                 // tree[state].p = p; backup of v
 *     mcts_policy_host(m, pi, action, sum_n);                 // calc_policy + apply_temperature
 */
typedef struct hive_mcts hive_mcts_t;
/* capacity for `sims` simulations per move (simulation_num_per_move, solo_play.py:23,98);
 * edges_per_sim <= 0 picks the default edge arena (160 edges per simulation and tree; the reference's games peak
 * at 130 legal actions).  A tree that runs out of nodes / edges / path depth gives the running simulation up
 * (virtual losses taken back), stops, and makes mcts_policy_host fail with HIVE_E_SEARCH -- never a silent
 * short search. */
int mcts_create(hive_env_t* env, int sims, int edges_per_sim, hive_mcts_t** out);
int mcts_destroy(hive_mcts_t* m);
int mcts_set_params(hive_mcts_t* m, int sims, int max_turn /*MAX_GAME_LENGTH*/, uint64_t noise_seed);
/* recorded Dirichlet rows for the root (np.random.dirichlet([0.3]*n_edges), solo_play.py:323), one
 * row per root visit: noise[n][rows][cols] doubles.  NULL switches back to on-device sampling. */
int mcts_set_root_noise_host(hive_mcts_t* m, const double* noise, int rows, int cols);
int mcts_begin(hive_mcts_t* m, const uint8_t* tree_mask /*host, NULL = all*/);
int mcts_descend(hive_mcts_t* m, int* n_pending /*host, may be NULL*/);
int mcts_expand(hive_mcts_t* m);
/* number of trees that asked for an evaluation in the last mcts_descend (synchronises) */
int mcts_pending_host(hive_mcts_t* m, int* n_pending);
/* leaf evaluation interface (device): planes bf16 [n][56][144] in, policy float [n][1584] and value
 * double [n] out; only rows whose pending-mask byte is 1 are read. */
void* mcts_dev_leaf_planes(hive_mcts_t* m);
void* mcts_dev_leaf_policy(hive_mcts_t* m);
void* mcts_dev_leaf_value(hive_mcts_t* m);
void* mcts_dev_pending_mask(hive_mcts_t* m);
/* the same through host buffers (expand_and_evaluate, solo_play.py:260-291) */
int mcts_leaf_planes_host(hive_mcts_t* m, uint16_t* planes_bf16, uint8_t* pending_mask);
int mcts_set_leaf_eval_host(hive_mcts_t* m, const float* policy, const double* value);
/* HivePlayer.action's return: pi[n][1584] (float64 like the reference), move, sum N.  Returns HIVE_E_SEARCH (outputs
 * still written) if any tree of this search was flagged (see mcts_create). */
int mcts_policy_host(hive_mcts_t* m, double* pi, int32_t* action, int32_t* sum_n);
/* OR of (1 << error code) over the trees of the running search: 2 node arena, 4 edge arena, 8 depth (synchronises) */
int mcts_error_host(hive_mcts_t* m, uint32_t* flags);
/* Shape of the trees of the last search over the first max_trees trees (measurement call: reads the trees back):
 * out[0] mean edges per node, out[1] mean select depth per simulation, out[2] mean nodes per tree, out[3] mean
 * simulations per tree -- the E and D of the per-simulation byte budget. */
int mcts_tree_stats_host(hive_mcts_t* m, int max_trees, double* out);
/* Deterministic stand-in for the network ON THE DEVICE, for parity tests of the device leaf-evaluation path
 * (expand_and_evaluate, solo_play.py:260-291, with a reproducible evaluator): policy / value are a pure hash of the
 * row's bf16 planes and `salt` (formula at the kernel in csrc/hive_mcts.cu).  Rows with mask_dev[row] == 0 are left
 * untouched (mask_dev may be NULL).  All pointers are device pointers; `stream` is a cudaStream_t. */
int mcts_hash_eval_dev(const uint16_t* planes_dev, float* policy_dev, double* value_dev, const uint8_t* mask_dev, int n,
                       uint64_t salt, void* stream);
/* the cudaStream_t the search (and its environment handle) queues its work on */
void* mcts_stream(hive_mcts_t* m);
/* root edges of one tree in ascending action order; info[6] = n_edges, sum_n, n_nodes, sims_done,
 * error (1 node arena, 2 edge arena, 3 depth), root_selects */
int mcts_root_stats_host(hive_mcts_t* m, int tree, int max_edges, int32_t* action, int32_t* N, double* W, double* Q,
                         float* P, int32_t* info);
long long mcts_launch_count(const hive_mcts_t* m);

/* ------------------------------------------------------------------------------------------
 * Network trunk on the tensor cores -- the 39 3x3 convolutions of ChessNet
 * (alpha_zero/alpha_net.py:25-54: ConvBlock + 19 ResBlocks, 98 % of the 6.56 GFLOP per position),
 * bf16 tcgen05 implicit GEMM with folded BatchNorm, residual add and ReLU fused.  The two small
 * heads (alpha_net.py:56-80) stay with the caller (plain library GEMMs).
 */
typedef struct hive_net hive_net_t;
int net_create(int device, void* stream, int max_boards, hive_net_t** out);
int net_destroy(hive_net_t* n);
/* layer 0 = stem (cin 56); 1+2i / 2+2i = conv1 / conv2 of residual block i (cin 256).
 * w [256][cin][3][3] fp32 with BatchNorm folded, bias [256] fp32 (host pointers). */
int net_load_conv_host(hive_net_t* n, int layer, const float* w, const float* bias, int cin);
/* planes: device bf16 [n_boards][56][144]; *out_nhwc: device bf16 [n_boards][144][256], valid until the
 * next call.  Stream-ordered on the handle's stream. */
int net_trunk_forward(hive_net_t* n, const uint16_t* planes_chw_dev, int n_boards, uint16_t** out_nhwc);
/* The heads (alpha_net.py:56-80), BatchNorm folded into the two 1x1 convolutions by the caller: pconv_w [128][256],
 * pconv_b [128]; vconv_w [256], vconv_b [1]; fc_w [1584][18432] with CELL-MAJOR columns (k = cell*128 + channel: the
 * reference's flatten c*144 + cell permuted once), fc_b [1584]; fc1_w [64][144], fc1_b [64]; fc2_w [64], fc2_b [1]. */
int net_load_heads_host(hive_net_t* n, const float* pconv_w, const float* pconv_b, const float* vconv_w, const float* vconv_b,
                        const float* fc_w, const float* fc_b, const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b);
/* The same two loads from DEVICE fp32 tensors (the module's own parameters after BatchNorm folding): packed by kernels on
 * the handle's stream, nothing crosses PCIe; reloads keep the device addresses of the packed operands (captured graphs
 * stay valid).  net_load_heads_dev takes fc_w in the REFERENCE's layout [1584][128*144] (channel-major columns). */
int net_load_conv_dev(hive_net_t* n, int layer, const float* w_dev, const float* bias_dev, int cin);
int net_load_heads_dev(hive_net_t* n, const float* pconv_w, const float* pconv_b, const float* vconv_w, const float* vconv_b,
                       const float* fc_w_ref, const float* fc_b, const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b);
/* ChessNet.forward (alpha_net.py:82-95; api_hive.py:60-74 is the server loop it replaces) for n_boards positions:
 * planes_chw_dev device bf16 [n][56][144] -> policy_dev device float32 [n][1584] (softmax), value_dev device float64 [n]
 * (tanh).  Trunk (39 tcgen05 convolutions) + heads (two tcgen05 GEMMs + one finishing kernel), all hand-written
 * sm_100a kernels queued on the handle's stream; the outputs may be the search's leaf arenas. */
int net_forward(hive_net_t* n, const uint16_t* planes_chw_dev, int n_boards, float* policy_dev, double* value_dev);
long long net_launch_count(const hive_net_t* n);

#ifdef __cplusplus
}
#endif
#endif /* HIVE_B200_H */
