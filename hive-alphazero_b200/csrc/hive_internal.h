// hive_internal.h -- private to the library: the handle behind hive_env_t and the launch helper the
// MCTS code shares with the environment code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/hive_b200.h"
#include "hive_core.cuh"

struct hive_env {
    int n = 0, device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = false;
    hive::GameRec* recs = nullptr;
    uint32_t* legal = nullptr;
    int32_t* count = nullptr;
    uint32_t* status = nullptr;
    uint8_t* lists = nullptr;       // [ceil(n/32)][LIST_BLOCK_BYTES] compact legal lists (written by steps that ask for them)
    bool want_lists = false;        // the launch being issued writes the lists (hive_step_host_async_lists)
    uint16_t* planes = nullptr;
    uint32_t* bits[2] = {nullptr, nullptr};   // bit planes, encode kernel -> plane-store kernel (double-buffered over steps)
    int last_bits = 0;              // which of the two buffers the last launch wrote (hive_bits_host)
    uint32_t* shadow = nullptr;     // delta plane store: the bit planes whose expansion the planes arena holds right now, [game][word]
    size_t shadow_bytes = 0;
    bool full_store = true;         // every step rewrites all 16 KB of a game's planes (hive_planes_kernel, the TMA store); HIVE_B200_DELTA_STORE=1: only the changed sectors
    int delta_ctas_per_sm = 4;      // HIVE_B200_DELTA_CTAS: cap of the delta store's grid (all concurrent store launches together)
    static constexpr int MAX_SUB = 16;
    int n_sub = 1;                  // the batch is cut into n_sub slices whose kernel chains overlap on side streams
    int sm_count = 148, store_ctas_per_sm = 3;   // the persistent plane-store kernels together keep this many CTAs per SM
    int host_slices = 2;            // slices of a step the host launches kernel by kernel (graph replays use n_sub)
    cudaStream_t sub_stream[MAX_SUB] = {}, store_stream[MAX_SUB] = {};
    cudaEvent_t encoded_ev[MAX_SUB] = {}, stored_ev[MAX_SUB][2] = {};
    cudaEvent_t fork_ev = nullptr, join_ev[MAX_SUB] = {};
    cudaEvent_t results_ev = nullptr;   // behind the result downloads of the last hive_step_host_async (hive_wait_results)
    // the step of the resident rollout loop is replayed from a CUDA graph (same arguments every step)
    struct StepGraph { int op = -1; const void* actions = nullptr; const void* mask = nullptr; void* chosen = nullptr;
                       uint64_t seed = 0; int max_turn = 0, auto_reset = 0; cudaGraphExec_t exec = nullptr; } graph, multi_graph;
    // the host-driven step (hive_step_host_async) from one fixed set of page-locked buffers: upload, kernels and
    // downloads replayed as one graph
    struct HostGraph { const void* actions = nullptr; const void* mask = nullptr; const void* count = nullptr; const void* status = nullptr;
                       const void* lists = nullptr; const void* seen_lists = nullptr;
                       const void* seen_actions = nullptr; const void* seen_mask = nullptr; const void* seen_count = nullptr;
                       const void* seen_status = nullptr; int launches = 0; cudaGraphExec_t exec = nullptr; } host_graph;
    int async_slices = 4;           // slices of a graph-replayed host-driven step
    cudaGraphExec_t slice_exec[MAX_SUB] = {};   // multi-step rollout: one graph per slice, each on its own stream
    bool skip_planes = false;       // measurement aid (HIVE_B200_EXPERIMENT_SKIP_PLANES): results are then incomplete
    int split_graphs = 1;
    int use_graph = 1;
    int stagger_ns = 25000;         // rollout kernel: start offset between the CTAs sharing an SM (HIVE_B200_STAGGER_US)
    int use_rollout_kernel = 0;     // HIVE_B200_ROLLOUT_KERNEL=1: hive_step_random_multi as ONE persistent launch (measured slower; default: per-step kernels in graphs)
    int rollout_queue = 0;          // HIVE_B200_ROLLOUT_QUEUE: hive_step_random_multi as two persistent kernels whose CTAs take (group, step) tickets
    int roll_ctas_per_sm = 2;       // step CTAs per SM of the queue-driven rollout (<= 2: the store CTAs must fit beside them)
    int roll_store_ctas_per_sm = 2; // store CTAs per SM of the queue-driven rollout
    void* roll_sync = nullptr;      // RollSync: ticket counters, done[G], stored[G]
    size_t roll_sync_bytes = 0;
    bool roll_pending = false;      // a queue-driven rollout was launched since the last hive_sync (its error word is unread)
    int32_t* d_actions[2] = {nullptr, nullptr};
    int act_flip = 0;
    cudaEvent_t act_read_ev[2] = {nullptr, nullptr};   // behind the step that read d_actions[b] (hive_step_host)
    bool act_used[2] = {false, false};
    uint8_t* d_mask = nullptr;
    uint32_t* hop_lines = nullptr;
    cudaEvent_t copy_done = nullptr, t0 = nullptr, t1 = nullptr;
    bool timing = false;
    long long launches = 0;
};


namespace hive {
// hive_create with an explicit slice count (<= 0: default)
int create_env(int n_games, int device, void* stream, int slices, hive_env** out);
// the random policy's host twin for games [g0, g1) of a batch of n (hive_host_pick_actions), on the calling thread
void pick_range(int g0, int g1, int n, const uint64_t* mask, const int32_t* count, const uint32_t* packed_status,
                uint32_t* episodes, uint64_t seed, int max_turn, int32_t* actions);
// sets the thread-local error text and returns `code`
int fail(int code, const std::string& msg);
// one environment step / evaluation over the whole batch (see hive_env_kernel.cuh for `op`)
int launch_env(hive_env* h, int op, const int32_t* actions, const uint8_t* mask, uint64_t seed, int max_turn,
               int auto_reset, int32_t* chosen);
}  // namespace hive

#define CUDA_TRY(x)                                                                                 \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) return hive::fail(HIVE_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)
