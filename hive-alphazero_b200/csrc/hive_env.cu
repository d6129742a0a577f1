// hive_env.cu -- environment kernels + C ABI (include/hive_b200.h) for sm_100a.
//
// One step of a batch of games is the whole GamePlay.move() of the reference (hive_engine/env_hive.py:99-171): apply
// the action, regenerate the legal set of the new side to move, encode its 56 planes, push history, test for the end
// of the game.  Two kernels per step (hive_env_kernel.cuh): hive_step_kernel (32 games per CTA: analyse -> one-hive
// floods -> move searches -> legal mask + bit planes) and hive_planes_kernel (bit planes -> bf16 planes through the
// TMA, persistent, on its own stream beside the next step).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/hive_b200.h"
#include "hive_env_kernel.cuh"
#include "hive_internal.h"
#include "hive_tables.h"

#ifndef HIVE_DEFAULT_SLICES
#define HIVE_DEFAULT_SLICES 2
#endif

using namespace hive;

namespace {

// ------------------------------------------------------------------------------------------
thread_local std::string g_err;

}  // namespace

namespace hive {

int fail(int code, const std::string& msg) { g_err = msg; return code; }

static int launch_env_kernels(hive_env* h, int op, const int32_t* actions, const uint8_t* mask, uint64_t seed, int max_turn,
                              int auto_reset, int32_t* chosen, int repeat = 1, int force_slices = 0, int* deferred_stores = nullptr);

int launch_env(hive_env* h, int op, const int32_t* actions, const uint8_t* mask, uint64_t seed, int max_turn,
               int auto_reset, int32_t* chosen) {
    h->last_bits = 0;                                        // a single step writes bit-plane buffer 0
    if (h->timing) CUDA_TRY(cudaEventRecord(h->t0, h->stream));
    // (host-driven steps alternate between two staging buffers, so they would never hit the cache)
    const bool graphable = h->use_graph && (op == OP_RANDOM || (op == OP_STEP && actions != h->d_actions[0] && actions != h->d_actions[1]));
    if (!graphable) {
        int rc = launch_env_kernels(h, op, actions, mask, seed, max_turn, auto_reset, chosen);
        if (rc) return rc;
    } else {
        hive_env::StepGraph& g = h->graph;
        const bool hit = g.exec && g.op == op && g.actions == actions && g.mask == mask && g.chosen == chosen && g.seed == seed &&
                         g.max_turn == max_turn && g.auto_reset == auto_reset;
        if (!hit) {
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            cudaGraph_t graph = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
            const long long l0 = h->launches;
            int rc = launch_env_kernels(h, op, actions, mask, seed, max_turn, auto_reset, chosen);
            cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
            h->launches = l0;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(HIVE_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
            e = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { g.exec = nullptr; return fail(HIVE_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
            g.op = op; g.actions = actions; g.mask = mask; g.chosen = chosen; g.seed = seed; g.max_turn = max_turn; g.auto_reset = auto_reset;
        }
        CUDA_TRY(cudaGraphLaunch(g.exec, h->stream));
        const int S = h->n_sub;
        h->launches += 2 * S;
    }
    if (h->timing) CUDA_TRY(cudaEventRecord(h->t1, h->stream));
    return 0;
}

static EnvArgs slice_args(hive_env* h, int s, int per, int op, const int32_t* actions, const uint8_t* mask, uint64_t seed, int max_turn,
                          int auto_reset, int32_t* chosen) {
    const int off = s * per;
    const int cnt = (off + per <= h->n) ? per : h->n - off;
    EnvArgs a;
    a.recs = h->recs + off; a.legal = h->legal + (size_t)off * LEGAL_WORDS; a.count = h->count + off;
    a.status = h->status + off; a.planes = h->planes + (size_t)off * HIVE_PLANES_ELEMS;
    a.actions = actions ? actions + off : nullptr; a.mask = mask ? mask + off : nullptr;
    a.chosen = chosen ? chosen + off : nullptr; a.hop_lines = h->hop_lines;
    a.seed = seed; a.n = cnt > 0 ? cnt : 0; a.op = op; a.max_turn = max_turn; a.auto_reset = auto_reset;
    a.g_offset = off; a.n_total = h->n;
    a.shadow = h->shadow ? h->shadow + (size_t)off * BITS_WORDS : nullptr;
    a.lists = h->want_lists ? h->lists + (size_t)(off / SG) * LIST_BLOCK_BYTES : nullptr;
    a.stagger_ns = h->stagger_ns; a.stagger_div = h->sm_count > 0 ? h->sm_count : 148;
    return a;
}
// The plane-store kernel is persistent: the store launches that run at the same time (`concurrent` slices) share
// store_ctas_per_sm CTAs per SM, so each grid is resident at once and never queues in front of other kernels.
static void launch_planes_part(hive_env* h, const EnvArgs& a, cudaStream_t st, int concurrent) {
    if (!h->full_store) {                                      // delta store: only the sectors that changed (hive_planes_delta_kernel)
        int blocks = (a.n + HIVE_DELTA_WARPS - 1) / HIVE_DELTA_WARPS;
        if (concurrent > 0) {
            int cap = h->sm_count * h->delta_ctas_per_sm / concurrent;
            if (cap < 1) cap = 1;
            if (blocks > cap) blocks = cap;
        }
        hive_planes_delta_kernel<<<blocks, HIVE_DELTA_WARPS * 32, 0, st>>>(a);
        return;
    }
    int blocks = (a.n + HIVE_STORE_WARPS - 1) / HIVE_STORE_WARPS;
    if (concurrent > 0) {                                      // 0: alone on the GPU (profiling), one warp per game
        int cap = h->sm_count * h->store_ctas_per_sm / concurrent;
        if (cap < 1) cap = 1;
        if (blocks > cap) blocks = cap;
    }
    hive_planes_kernel<<<blocks, HIVE_STORE_WARPS * 32, STORE_STAGE_BYTES, st>>>(a);
}

// `repeat` steps of one slice: the step kernel on stream `st`, and the plane store of every step on stream `ss`, so
// that it runs beside the next step's kernel (the bit planes it reads are double-buffered; step k+2's kernel waits
// for step k's store).  With st != ss the chain ends joined on `st`.
// `defer_store_join` (single steps only): `st` is NOT joined with the plane store; the caller joins
// stored_ev[s][0] itself, after whatever it wants to run beside the store (the result downloads of the host-driven step).
static int launch_slice_chain(hive_env* h, int s, EnvArgs a, cudaStream_t st, cudaStream_t ss, int concurrent, int repeat,
                              bool defer_store_join = false) {
    const bool side = st != ss;
    for (int rep = 0; rep < repeat; rep++) {
        a.bits = h->bits[rep & 1] + (size_t)a.g_offset * BITS_WORDS;
        const bool waits_store = side && rep >= 2;
        if (waits_store) CUDA_TRY(cudaStreamWaitEvent(st, h->stored_ev[s][rep & 1], 0));    // this bit-plane buffer is free again
        hive_step_kernel<<<(a.n + SG - 1) / SG, STEP_THREADS, 0, st>>>(a);
        if (side) {
            CUDA_TRY(cudaEventRecord(h->encoded_ev[s], st));
            CUDA_TRY(cudaStreamWaitEvent(ss, h->encoded_ev[s], 0));
        }
        if (!h->skip_planes) launch_planes_part(h, a, ss, concurrent);
        if (side) CUDA_TRY(cudaEventRecord(h->stored_ev[s][rep & 1], ss));
    }
    if (side && !(defer_store_join && repeat == 1)) {
        CUDA_TRY(cudaStreamWaitEvent(st, h->stored_ev[s][(repeat - 1) & 1], 0));
        if (repeat > 1) CUDA_TRY(cudaStreamWaitEvent(st, h->stored_ev[s][repeat & 1], 0));
    }
    CUDA_TRY(cudaGetLastError());
    h->launches += 2 * repeat;
    return 0;
}

// One or `repeat` steps of the whole batch: every slice's chain on its own pair of streams, joined into h->stream.
// `deferred_stores` (repeat == 1): if given, h->stream is joined with the encode kernels only and *deferred_stores
// receives the number of slices whose plane store (stored_ev[s][0]) the caller still has to join.
static int launch_env_kernels(hive_env* h, int op, const int32_t* actions, const uint8_t* mask, uint64_t seed, int max_turn,
                              int auto_reset, int32_t* chosen, int repeat, int force_slices, int* deferred_stores) {
    // launches the host issues one by one (not graph replays) are launch-bound: they use at most host_slices slices
    cudaStreamCaptureStatus cap_state = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(h->stream, &cap_state));
    int S = (cap_state == cudaStreamCaptureStatusActive || h->n_sub < h->host_slices) ? h->n_sub : h->host_slices;
    if (force_slices > 0) S = force_slices < h->n_sub ? force_slices : h->n_sub;
    // slices are multiples of SG games so that CTAs never straddle two slices
    const int per = ((h->n + S - 1) / S + SG - 1) / SG * SG;
    const bool side = S > 1 || repeat > 1;                 // side streams in use (else everything goes down h->stream)
    const bool defer = deferred_stores && side && repeat == 1 && !h->skip_planes;
    if (deferred_stores) *deferred_stores = 0;
    if (side) CUDA_TRY(cudaEventRecord(h->fork_ev, h->stream));
    for (int s = 0; s < S; s++) {
        const EnvArgs a = slice_args(h, s, per, op, actions, mask, seed, max_turn, auto_reset, chosen);
        if (a.n <= 0) break;
        cudaStream_t st = side ? h->sub_stream[s] : h->stream, ss = side ? h->store_stream[s] : h->stream;
        if (side) CUDA_TRY(cudaStreamWaitEvent(st, h->fork_ev, 0));
        int rc = launch_slice_chain(h, s, a, st, ss, S, repeat, defer);
        if (rc) return rc;
        if (side) {
            CUDA_TRY(cudaEventRecord(h->join_ev[s], st));
            CUDA_TRY(cudaStreamWaitEvent(h->stream, h->join_ev[s], 0));
        }
        if (defer) *deferred_stores = s + 1;
    }
    return 0;
}

}  // namespace hive

namespace {
int check(const hive_env* h) { return h && h->n > 0 ? 0 : fail(HIVE_E_HANDLE, "bad handle"); }
}  // namespace

extern "C" {

#ifdef HIVE_TRACE
// experiment builds only: per-CTA timeline of the four kernels (records of 32 bytes, see TraceRec)
int hive_trace_start(int cap) {
    TraceRec* buf = nullptr;
    unsigned zero = 0, ucap = (unsigned)cap;
    if (cudaMalloc(&buf, (size_t)cap * sizeof(TraceRec)) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)); cudaMemcpyToSymbol(g_trace_n, &zero, 4); cudaMemcpyToSymbol(g_trace_cap, &ucap, 4);
    return 0;
}
int hive_trace_read(void* host, int cap) {
    TraceRec* buf = nullptr; unsigned n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&buf, g_trace, sizeof(buf)); cudaMemcpyFromSymbol(&n, g_trace_n, 4);
    if ((int)n > cap) n = cap;
    cudaMemcpy(host, buf, (size_t)n * sizeof(TraceRec), cudaMemcpyDeviceToHost);
    TraceRec* none = nullptr; cudaMemcpyToSymbol(g_trace, &none, sizeof(none)); cudaFree(buf);
    return (int)n;
}
#endif

#ifdef HIVE_PHASE_CLOCKS
// experiment builds only: summed SM clocks per phase of the step kernel ([7] = CTAs), optionally zeroed after the read
int hive_phase_clocks(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    if (out) cudaMemcpyFromSymbol(out, g_phase_clk, 64);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_phase_clk, z, 64); }
    return 0;
}
#endif

const char* hive_last_error(void) { return g_err.c_str(); }
int hive_abi_version(void) { return 1; }

int hive_create(int n_games, int device, void* stream, hive_env_t** out) { return hive::create_env(n_games, device, stream, 0, out); }

}  // extern "C"

// slices <= 0: HIVE_B200_SLICES / the default; the search's working batch asks for 1 (its evaluations are
// not graph-replayed, so fewer launches win)
int hive::create_env(int n_games, int device, void* stream, int slices, hive_env** out) {
    if (!out || n_games <= 0) return fail(HIVE_E_ARG, "hive_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(HIVE_E_CUDA, std::string("hive_create: no CUDA device (") + cudaGetErrorString(e) +
                                     "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(HIVE_E_ARG, "hive_create: bad device index");
    CUDA_TRY(cudaSetDevice(device));
    hive_env* h = new hive_env();
    h->n = n_games; h->device = device;
    if (stream) { h->stream = (cudaStream_t)stream; } else {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
    CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->copy_done, cudaEventDisableTiming));
    for (int b = 0; b < 2; b++) CUDA_TRY(cudaEventCreateWithFlags(&h->act_read_ev[b], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&h->t0));
    CUDA_TRY(cudaEventCreate(&h->t1));
    const size_t n = (size_t)n_games;
    // static + dynamic shared memory of the plane-store kernel exceeds the 48 KB default
    CUDA_TRY(cudaFuncSetAttribute(hive_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STORE_STAGE_BYTES));
    // the step kernels and the persistent plane stores share the SMs: both ask for the largest shared-memory carve-out, so
    // that an SM never has to drain to be re-partitioned before a CTA of the other kernel fits
    if (!getenv("HIVE_B200_NO_CARVEOUT")) {
        CUDA_TRY(cudaFuncSetAttribute(hive_planes_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(hive_step_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(hive_rollout_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        // (for the two persistent kernels of the queue-driven rollout this is a matter of progress, not of speed: an SM
        // configured for one of them alone could never take a CTA of the other)
        CUDA_TRY(cudaFuncSetAttribute(hive_rollout_q_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(hive_planes_q_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(hive_step_flow_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    for (int b = 0; b < 2; b++) {
        CUDA_TRY(cudaMalloc(&h->bits[b], n * BITS_WORDS * 4));
        CUDA_TRY(cudaMemsetAsync(h->bits[b], 0, n * BITS_WORDS * 4, h->stream));
    }
    CUDA_TRY(cudaMalloc(&h->recs, n * sizeof(GameRec)));
    // legal masks, counts and status words sit in ONE arena (208 B per game) so that the host-driven step brings them
    // down with a single copy
    CUDA_TRY(cudaMalloc(&h->legal, n * (LEGAL_WORDS * 4 + 8)));
    h->count = reinterpret_cast<int32_t*>(h->legal + n * LEGAL_WORDS);
    h->status = reinterpret_cast<uint32_t*>(h->count + n);
    CUDA_TRY(cudaMalloc(&h->planes, n * HIVE_PLANES_ELEMS * 2));
    CUDA_TRY(cudaMalloc(&h->lists, (n + SG - 1) / SG * LIST_BLOCK_BYTES));
    {
        const char* rk0 = getenv("HIVE_B200_ROLLOUT_KERNEL");
        const char* ds = getenv("HIVE_B200_DELTA_STORE");          // 1: hive_planes_delta_kernel (measured: same speed, 4x less HBM traffic)
        h->full_store = !(ds && atoi(ds)) || (rk0 && atoi(rk0));   // (the rollout kernel stores full planes itself)
        const char* dc = getenv("HIVE_B200_DELTA_CTAS");
        h->delta_ctas_per_sm = dc && atoi(dc) > 0 ? atoi(dc) : 4;
        if (!h->full_store) {   // planes arena and shadow start out equal: all zero
            h->shadow_bytes = n * BITS_WORDS * 4;
            CUDA_TRY(cudaMalloc(&h->shadow, h->shadow_bytes));
            CUDA_TRY(cudaMemsetAsync(h->shadow, 0, h->shadow_bytes, h->stream));
            CUDA_TRY(cudaMemsetAsync(h->planes, 0, n * HIVE_PLANES_ELEMS * 2, h->stream));
        }
    }
    {
        const char* e = getenv("HIVE_B200_SLICES");
        // measured with three store CTAs per SM (16,384 games): 2 slices 53.8, 4 slices 53.0, 5: 52.9, 6: 55.9, 8: 53.3 us per step;
        // the 4,096-game parts of the host-driven path are faster with 2 (183 against 178 M env-steps/s)
        int S = slices > 0 ? slices : (e ? atoi(e) : (n_games >= 16384 ? 4 : HIVE_DEFAULT_SLICES));
        if (S < 1) S = 1;
        if (S > hive_env::MAX_SUB) S = hive_env::MAX_SUB;
        while (S > 1 && n_games < S * SG * 2) S--;            // small batches are not worth slicing
        h->n_sub = S;
        const char* ec = getenv("HIVE_B200_STORE_CTAS");
        h->store_ctas_per_sm = ec && atoi(ec) > 0 ? atoi(ec) : 3;      // measured: 2 -> 55.3, 3 -> 54.4, 4 -> 56.6 us per 16,384-game step
        const char* hs = getenv("HIVE_B200_HOST_SLICES");
        h->host_slices = hs && atoi(hs) > 0 ? atoi(hs) : 2;
        const char* as = getenv("HIVE_B200_ASYNC_SLICES");
        h->async_slices = as && atoi(as) > 0 ? atoi(as) : 4;
        if (h->async_slices > h->n_sub) h->async_slices = h->n_sub;
        h->skip_planes = getenv("HIVE_B200_EXPERIMENT_SKIP_PLANES") != nullptr;   // measurement aid: the step without its plane store
        const char* sg = getenv("HIVE_B200_SPLIT_GRAPHS");
        h->split_graphs = sg ? atoi(sg) : 1;
        const char* sg2 = getenv("HIVE_B200_STAGGER_US");
        h->stagger_ns = (sg2 ? atoi(sg2) : 25) * 1000;
        const char* rk = getenv("HIVE_B200_ROLLOUT_KERNEL");
        h->use_rollout_kernel = rk ? atoi(rk) : 0;         // measured slower than the per-step kernels (profiles/README.md): off
        const char* rq = getenv("HIVE_B200_ROLLOUT_QUEUE");
        h->rollout_queue = rq ? atoi(rq) : 0;
        const char* rc = getenv("HIVE_B200_ROLL_CTAS");
        h->roll_ctas_per_sm = rc && atoi(rc) >= 1 && atoi(rc) <= 4 ? atoi(rc) : 2;    // default 2: two store CTAs fit beside them (a configuration without room for the store ends in the waits' time-out, not in a hang)
        const char* qs = getenv("HIVE_B200_ROLL_STORE_CTAS");
        h->roll_store_ctas_per_sm = qs && atoi(qs) >= 1 ? atoi(qs) : 2;
        h->roll_sync_bytes = 128 + 2 * ((n + SG - 1) / SG) * 4;
        CUDA_TRY(cudaMalloc(&h->roll_sync, h->roll_sync_bytes));
        CUDA_TRY(cudaFuncSetAttribute(hive_planes_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STORE_STAGE_BYTES));
        const char* ug = getenv("HIVE_B200_GRAPH");
        h->use_graph = ug ? atoi(ug) : 1;
    }
    CUDA_TRY(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->results_ev, cudaEventDisableTiming));
    for (int s = 0; s < h->n_sub; s++) {
        {   // HIVE_B200_PRIO (experiment): 1 = step kernels above the plane stores, 2 = the other way round
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            const char* pe = getenv("HIVE_B200_PRIO");
            const int mode = pe ? atoi(pe) : 0;
            CUDA_TRY(cudaStreamCreateWithPriority(&h->sub_stream[s], cudaStreamNonBlocking, mode == 1 ? hi : mode == 2 ? lo : 0));
            CUDA_TRY(cudaStreamCreateWithPriority(&h->store_stream[s], cudaStreamNonBlocking, mode == 1 ? lo : mode == 2 ? hi : 0));
        }
        CUDA_TRY(cudaEventCreateWithFlags(&h->join_ev[s], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->encoded_ev[s], cudaEventDisableTiming));
        for (int b = 0; b < 2; b++) CUDA_TRY(cudaEventCreateWithFlags(&h->stored_ev[s][b], cudaEventDisableTiming));
    }
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        h->sm_count = sms;
    }
    CUDA_TRY(cudaMalloc(&h->d_actions[0], n * 4));
    CUDA_TRY(cudaMalloc(&h->d_actions[1], n * 4));
    CUDA_TRY(cudaMalloc(&h->d_mask, n));
    CUDA_TRY(cudaMalloc(&h->hop_lines, GEO_WORDS * 4));       // hop lines + neighbour rank / neighbour cell tables (hive_tables.h)
    std::vector<uint32_t> lines;
    build_geometry_tables(lines);
    CUDA_TRY(cudaMemcpyAsync(h->hop_lines, lines.data(), lines.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->recs, 0, n * sizeof(GameRec), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    int rc = launch_env(h, OP_INIT, nullptr, nullptr, 0, 0, 0, nullptr);
    if (rc) { hive_destroy(h); return rc; }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *out = h;
    return 0;
}

extern "C" {

int hive_destroy(hive_env_t* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->graph.exec) cudaGraphExecDestroy(h->graph.exec);
    if (h->multi_graph.exec && h->multi_graph.exec != h->slice_exec[0]) cudaGraphExecDestroy(h->multi_graph.exec);
    for (int s = 0; s < hive_env::MAX_SUB; s++) if (h->slice_exec[s]) cudaGraphExecDestroy(h->slice_exec[s]);
    if (h->host_graph.exec) cudaGraphExecDestroy(h->host_graph.exec);
    cudaFree(h->recs); cudaFree(h->legal); cudaFree(h->planes); cudaFree(h->lists); cudaFree(h->bits[0]); cudaFree(h->bits[1]); cudaFree(h->shadow); cudaFree(h->roll_sync);
    for (int s = 0; s < h->n_sub; s++) {
        if (h->sub_stream[s]) cudaStreamDestroy(h->sub_stream[s]);
        if (h->store_stream[s]) cudaStreamDestroy(h->store_stream[s]);
        if (h->encoded_ev[s]) cudaEventDestroy(h->encoded_ev[s]);
        for (int b = 0; b < 2; b++) if (h->stored_ev[s][b]) cudaEventDestroy(h->stored_ev[s][b]);
        if (h->join_ev[s]) cudaEventDestroy(h->join_ev[s]);
    }
    if (h->fork_ev) cudaEventDestroy(h->fork_ev);
    if (h->results_ev) cudaEventDestroy(h->results_ev);
    cudaFree(h->d_actions[0]); cudaFree(h->d_actions[1]); cudaFree(h->d_mask); cudaFree(h->hop_lines);
    if (h->copy_done) cudaEventDestroy(h->copy_done);
    for (int b = 0; b < 2; b++) if (h->act_read_ev[b]) cudaEventDestroy(h->act_read_ev[b]);
    if (h->t0) cudaEventDestroy(h->t0);
    if (h->t1) cudaEventDestroy(h->t1);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int hive_num_games(const hive_env_t* h) { return h ? h->n : HIVE_E_HANDLE; }

int hive_sync(hive_env_t* h) {
    if (check(h)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->roll_pending) {   // a queue-driven rollout ran since the last sync: did one of its waits give up?
        h->roll_pending = false;
        unsigned err = 0;
        CUDA_TRY(cudaMemcpy(&err, reinterpret_cast<const char*>(h->roll_sync) + 8, 4, cudaMemcpyDeviceToHost));
        if (err) {
            char msg[160];
            snprintf(msg, sizeof msg, "queue-driven rollout: wait kind %u of ticket %u timed out (the batch is in an undefined state: reset it)",
                     err >> 28, err & 0x0FFFFFFFu);
            return fail(HIVE_E_CUDA, msg);
        }
    }
    return 0;
}

int hive_reset(hive_env_t* h, const uint8_t* game_mask) {
    if (check(h)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaSetDevice(h->device));
    const uint8_t* dm = nullptr;
    if (game_mask) {
        CUDA_TRY(cudaMemcpyAsync(h->d_mask, game_mask, h->n, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        dm = h->d_mask;
    }
    return launch_env(h, OP_RESET, nullptr, dm, 0, 0, 0, nullptr);
}

int hive_step(hive_env_t* h, const int32_t* actions_dev) {
    if (check(h)) return HIVE_E_HANDLE;
    if (!actions_dev) return fail(HIVE_E_ARG, "hive_step: null actions");
    CUDA_TRY(cudaSetDevice(h->device));
    return launch_env(h, OP_STEP, actions_dev, nullptr, 0, 0, 0, nullptr);
}

int hive_step_host(hive_env_t* h, const int32_t* actions) {
    if (check(h)) return HIVE_E_HANDLE;
    if (!actions) return fail(HIVE_E_ARG, "hive_step_host: null actions");
    CUDA_TRY(cudaSetDevice(h->device));
    // double-buffered device copy of the actions on a side stream: the caller's buffer is free again when this
    // returns, and the previous step may still be reading the other buffer.  The buffer used now was read by the step
    // of two calls ago: the copy waits for that step's event before it overwrites it.
    const int flip = h->act_flip;
    int32_t* d = h->d_actions[flip];
    h->act_flip ^= 1;
    if (h->act_used[flip]) CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->act_read_ev[flip], 0));
    CUDA_TRY(cudaMemcpyAsync(d, actions, (size_t)h->n * 4, cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(cudaEventRecord(h->copy_done, h->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->copy_done, 0));
    int rc = launch_env(h, OP_STEP, d, nullptr, 0, 0, 0, nullptr);
    CUDA_TRY(cudaEventRecord(h->act_read_ev[flip], h->stream));
    h->act_used[flip] = true;
    CUDA_TRY(cudaEventSynchronize(h->copy_done));
    return rc;
}

// true if `p` is page-locked host memory (cudaHostAlloc / cudaHostRegister): only such buffers can be the
// endpoints of copy nodes in a captured graph
static bool is_pinned_host(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// H2D actions -> the step's kernels -> D2H results, queued on h->stream.  The downloads need the encode kernels only:
// they are queued BEFORE the join with the plane stores, and results_ev is recorded behind them, so that
// hive_wait_results returns while the 16 KB/game of planes are still being written (hive_sync waits for those too).
static int queue_host_step(hive_env* h, int32_t* d, const int32_t* actions, uint64_t* mask, int32_t* count, uint32_t* packed_status,
                           int slices, uint8_t* lists = nullptr) {
    // page-locked actions are read by the step kernel straight from host memory (4 B per game, coalesced: one PCIe read
    // per warp) instead of through a copy node of their own; pageable ones are staged
    const int32_t* act = actions;
    if (!is_pinned_host(actions)) {
        CUDA_TRY(cudaMemcpyAsync(d, actions, (size_t)h->n * 4, cudaMemcpyHostToDevice, h->stream));
        act = d;
    }
    int deferred = 0;
    h->last_bits = 0;
    h->want_lists = lists != nullptr;
    int rc = launch_env_kernels(h, OP_STEP, act, nullptr, 0, 0, 0, nullptr, 1, slices > 0 ? slices : h->host_slices, &deferred);
    h->want_lists = false;
    if (rc) return rc;
    const size_t n = (size_t)h->n;
    if (lists) CUDA_TRY(cudaMemcpyAsync(lists, h->lists, (n + SG - 1) / SG * LIST_BLOCK_BYTES, cudaMemcpyDeviceToHost, h->stream));
    if (mask && count && packed_status && reinterpret_cast<uint8_t*>(count) == reinterpret_cast<uint8_t*>(mask) + n * LEGAL_WORDS * 4 &&
        reinterpret_cast<uint8_t*>(packed_status) == reinterpret_cast<uint8_t*>(count) + n * 4) {
        // the caller's buffers are laid out like the device arena: one download
        CUDA_TRY(cudaMemcpyAsync(mask, h->legal, n * (LEGAL_WORDS * 4 + 8), cudaMemcpyDeviceToHost, h->stream));
    } else {
        if (mask) CUDA_TRY(cudaMemcpyAsync(mask, h->legal, n * LEGAL_WORDS * 4, cudaMemcpyDeviceToHost, h->stream));
        if (count) CUDA_TRY(cudaMemcpyAsync(count, h->count, n * 4, cudaMemcpyDeviceToHost, h->stream));
        if (packed_status) CUDA_TRY(cudaMemcpyAsync(packed_status, h->status, n * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    cudaStreamCaptureStatus cap_state = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(h->stream, &cap_state));
    CUDA_TRY(cudaEventRecordWithFlags(h->results_ev, h->stream,
                                      cap_state == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
    for (int s = 0; s < deferred; s++) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->stored_ev[s][0], 0));
    return 0;
}

// The host-driven step.  A caller that steps from the same page-locked buffers every time (the normal loop: one
// set of pinned staging buffers per handle) gets the whole sequence -- action upload, the five kernels of every
// slice with their fork/join events, the three result downloads -- replayed as ONE CUDA graph launch: the host
// thread then spends a few microseconds per step instead of ~30 driver calls, which is what bounded this path.
static int step_host_async(hive_env_t* h, const int32_t* actions, uint64_t* mask, int32_t* count, uint32_t* packed_status, uint8_t* lists);
int hive_step_host_async(hive_env_t* h, const int32_t* actions, uint64_t* mask, int32_t* count, uint32_t* packed_status) {
    return step_host_async(h, actions, mask, count, packed_status, nullptr);
}
int hive_step_host_async_lists(hive_env_t* h, const int32_t* actions, uint8_t* lists, uint32_t* packed_status) {
    if (!lists) return fail(HIVE_E_ARG, "hive_step_host_async_lists: null lists");
    return step_host_async(h, actions, nullptr, nullptr, packed_status, lists);
}
static int step_host_async(hive_env_t* h, const int32_t* actions, uint64_t* mask, int32_t* count, uint32_t* packed_status, uint8_t* lists) {
    if (check(h)) return HIVE_E_HANDLE;
    if (!actions) return fail(HIVE_E_ARG, "hive_step_host_async: null actions");
    CUDA_TRY(cudaSetDevice(h->device));
    hive_env::HostGraph& g = h->host_graph;
    const bool same = g.exec && g.actions == actions && g.mask == mask && g.count == count && g.status == packed_status && g.lists == lists;
    if (!same && h->use_graph) {
        // a second call with the same buffers builds the graph; one-off callers stay on the plain path
        const bool repeat_caller = g.seen_actions == actions && g.seen_mask == mask && g.seen_count == count && g.seen_status == packed_status && g.seen_lists == lists;
        g.seen_actions = actions; g.seen_mask = mask; g.seen_count = count; g.seen_status = packed_status; g.seen_lists = lists;
        if (repeat_caller && is_pinned_host(actions) && (!mask || is_pinned_host(mask)) && (!count || is_pinned_host(count)) &&
            (!packed_status || is_pinned_host(packed_status)) && (!lists || is_pinned_host(lists))) {
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            cudaGraph_t graph = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
            const long long l0 = h->launches;
            int rc = queue_host_step(h, h->d_actions[0], actions, mask, count, packed_status, h->async_slices, lists);
            g.launches = (int)(h->launches - l0);
            h->launches = l0;
            cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(HIVE_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
            e = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { g.exec = nullptr; return fail(HIVE_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
            g.actions = actions; g.mask = mask; g.count = count; g.status = packed_status; g.lists = lists;
        }
    }
    if (g.exec && g.actions == actions && g.mask == mask && g.count == count && g.status == packed_status && g.lists == lists) {
        CUDA_TRY(cudaGraphLaunch(g.exec, h->stream));
        h->launches += g.launches;
        return 0;
    }
    // stream-ordered on h->stream, so one device-side action buffer would do; the two alternate with hive_step_host's
    int32_t* d = h->d_actions[h->act_flip];
    h->act_flip ^= 1;
    return queue_host_step(h, d, actions, mask, count, packed_status, 0, lists);
}

int hive_wait_results(hive_env_t* h) {
    if (check(h)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventSynchronize(h->results_ev));
    return 0;
}

int hive_step_random(hive_env_t* h, uint64_t seed, int max_turn, int auto_reset, int32_t* chosen_dev) {
    if (check(h)) return HIVE_E_HANDLE;
    if (max_turn < 1 || max_turn > 250) return fail(HIVE_E_ARG, "hive_step_random: max_turn out of range");
    CUDA_TRY(cudaSetDevice(h->device));
    return launch_env(h, OP_RANDOM, nullptr, nullptr, seed, max_turn, auto_reset, chosen_dev);
}

// n_steps consecutive rollout steps as ONE CUDA graph: inside the graph every slice of the batch runs its
// own chain of n_steps x (analyse, flood, moves, encode), so the memory-bound encode of one slice
// overlaps the issue-bound kernels of the others across step boundaries.
int hive_step_random_multi(hive_env_t* h, uint64_t seed, int max_turn, int auto_reset, int n_steps) {
    if (check(h)) return HIVE_E_HANDLE;
    if (max_turn < 1 || max_turn > 250 || n_steps < 1 || n_steps > 4096) return fail(HIVE_E_ARG, "hive_step_random_multi: bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->timing) CUDA_TRY(cudaEventRecord(h->t0, h->stream));
    h->last_bits = h->use_rollout_kernel ? 0 : (n_steps - 1) & 1;      // step k of a multi-step rollout writes buffer k & 1
    if (h->use_rollout_kernel) {
        // one launch: every CTA walks through the n_steps steps of its 32 games and stores their planes itself
        EnvArgs a = slice_args(h, 0, h->n, OP_RANDOM, nullptr, nullptr, seed, max_turn, auto_reset, nullptr);
        a.bits = h->bits[0];
        hive_rollout_kernel<<<(h->n + SG - 1) / SG, STEP_THREADS, 0, h->stream>>>(a, n_steps);
        CUDA_TRY(cudaGetLastError());
        h->launches += 1;
        if (h->timing) CUDA_TRY(cudaEventRecord(h->t1, h->stream));
        return 0;
    }
    if (h->rollout_queue && n_steps >= 2 && h->full_store && !h->skip_planes) {
        // the queue-driven rollout: two persistent kernels, CTAs take (group, step) tickets (hive_rollout_q_kernel)
        const int G = (h->n + SG - 1) / SG;
        EnvArgs a = slice_args(h, 0, h->n, OP_RANDOM, nullptr, nullptr, seed, max_turn, auto_reset, nullptr);
        a.bits = h->bits[0];
        cudaStream_t ss = h->store_stream[0];
        if (h->roll_pending) { int rc = hive_sync(h); if (rc) return rc; }      // (reads the previous rollout's error word before it is zeroed)
        CUDA_TRY(cudaMemsetAsync(h->roll_sync, 0, h->roll_sync_bytes, h->stream));
        CUDA_TRY(cudaEventRecord(h->fork_ev, h->stream));
        CUDA_TRY(cudaStreamWaitEvent(ss, h->fork_ev, 0));
        int store_blocks = h->sm_count * h->roll_store_ctas_per_sm, step_blocks = h->sm_count * h->roll_ctas_per_sm;
        if (store_blocks > G) store_blocks = G;
        if (step_blocks > G) step_blocks = G;
        // the store kernel first: its CTAs must find room (the step CTAs wait for it from their third step on; with at most
        // two step CTAs per SM there is always room, whatever the placement)
        hive_planes_q_kernel<<<store_blocks, HIVE_STORE_WARPS * 32, STORE_STAGE_BYTES, ss>>>(a, h->bits[1], n_steps, reinterpret_cast<RollSync*>(h->roll_sync));
        if (h->rollout_queue == 2) {
            // one launch per step, chained by programmatic stream serialization + per-group flags (hive_step_flow_kernel)
            for (int k = 0; k < n_steps; k++) {
                EnvArgs ak = a;
                ak.bits = h->bits[k & 1];
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)G); cfg.blockDim = dim3(STEP_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = k > 0 ? 1 : 0;
                CUDA_TRY(cudaLaunchKernelEx(&cfg, hive_step_flow_kernel, ak, k, reinterpret_cast<RollSync*>(h->roll_sync)));
            }
            h->launches += n_steps - 1;
        } else {
            hive_rollout_q_kernel<<<step_blocks, STEP_THREADS, 0, h->stream>>>(a, h->bits[1], n_steps, reinterpret_cast<RollSync*>(h->roll_sync));
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(h->stored_ev[0][0], ss));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->stored_ev[0][0], 0));
        h->launches += 2;
        h->roll_pending = true;
        if (h->timing) CUDA_TRY(cudaEventRecord(h->t1, h->stream));
        return 0;
    }
    hive_env::StepGraph& g = h->multi_graph;
    const bool hit = g.exec && g.seed == seed && g.max_turn == max_turn && g.auto_reset == auto_reset && g.op == n_steps;
    const int S = h->n_sub;
    const bool split = h->split_graphs && S > 1;
    if (!hit) {
        if (g.exec && g.exec != h->slice_exec[0]) cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
        for (int s = 0; s < hive_env::MAX_SUB; s++) if (h->slice_exec[s]) { cudaGraphExecDestroy(h->slice_exec[s]); h->slice_exec[s] = nullptr; }
        const long long l0 = h->launches;
        if (!split) {
            cudaGraph_t graph = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
            int rc = launch_env_kernels(h, OP_RANDOM, nullptr, nullptr, seed, max_turn, auto_reset, nullptr, n_steps);
            cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
            h->launches = l0;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(HIVE_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
            e = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { g.exec = nullptr; return fail(HIVE_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
        } else {
            // One graph PER SLICE, each launched into its own stream.  The nodes of a single graph holding all
            // slices are issued level by level (all analyse kernels, then all floods, ...: profiles/README.md,
            // v8 trace), which lines the slices' phases up and leaves the SMs to one kernel type at a time;
            // separate graphs are separate launch queues, so the chains drift apart and their phases mix.
            const int per = ((h->n + S - 1) / S + SG - 1) / SG * SG;
            for (int s = 0; s < S; s++) {
                const EnvArgs a = slice_args(h, s, per, OP_RANDOM, nullptr, nullptr, seed, max_turn, auto_reset, nullptr);
                if (a.n <= 0) break;
                cudaGraph_t graph = nullptr;
                CUDA_TRY(cudaStreamBeginCapture(h->sub_stream[s], cudaStreamCaptureModeRelaxed));
                int rc = launch_slice_chain(h, s, a, h->sub_stream[s], h->store_stream[s], S, n_steps);
                cudaError_t e = cudaStreamEndCapture(h->sub_stream[s], &graph);
                if (rc) { h->launches = l0; if (graph) cudaGraphDestroy(graph); return rc; }
                if (e != cudaSuccess) { h->launches = l0; return fail(HIVE_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); }
                e = cudaGraphInstantiate(&h->slice_exec[s], graph, 0);
                cudaGraphDestroy(graph);
                if (e != cudaSuccess) { h->slice_exec[s] = nullptr; h->launches = l0; return fail(HIVE_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
            }
            h->launches = l0;
            g.exec = h->slice_exec[0];       // marks the cache valid (owned through slice_exec)
        }
        g.op = n_steps; g.seed = seed; g.max_turn = max_turn; g.auto_reset = auto_reset;
    }
    if (!split) {
        CUDA_TRY(cudaGraphLaunch(g.exec, h->stream));
    } else {
        CUDA_TRY(cudaEventRecord(h->fork_ev, h->stream));
        for (int s = 0; s < S && h->slice_exec[s]; s++) {
            CUDA_TRY(cudaStreamWaitEvent(h->sub_stream[s], h->fork_ev, 0));
            CUDA_TRY(cudaGraphLaunch(h->slice_exec[s], h->sub_stream[s]));
            CUDA_TRY(cudaEventRecord(h->join_ev[s], h->sub_stream[s]));
            CUDA_TRY(cudaStreamWaitEvent(h->stream, h->join_ev[s], 0));
        }
    }
    h->launches += 2LL * h->n_sub * n_steps;
    if (h->timing) CUDA_TRY(cudaEventRecord(h->t1, h->stream));
    return 0;
}

// One rollout step with CUDA events around its kernels (whole batch as one slice, no graph):
// ms[0] = step kernel, ms[1] = plane store.  For bench.py's per-kernel roofline; it advances the games.
int hive_profile_step(hive_env_t* h, uint64_t seed, int max_turn, float* ms) {
    if (check(h)) return HIVE_E_HANDLE;
    if (!ms || max_turn < 1 || max_turn > 250) return fail(HIVE_E_ARG, "hive_profile_step: bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaEvent_t ev[3];
    for (int i = 0; i < 3; i++) CUDA_TRY(cudaEventCreate(&ev[i]));
    EnvArgs a = slice_args(h, 0, h->n, OP_RANDOM, nullptr, nullptr, seed, max_turn, 1, nullptr);
    a.bits = h->bits[0];
    CUDA_TRY(cudaEventRecord(ev[0], h->stream));
    hive_step_kernel<<<(h->n + SG - 1) / SG, STEP_THREADS, 0, h->stream>>>(a);
    CUDA_TRY(cudaEventRecord(ev[1], h->stream));
    if (!h->skip_planes) launch_planes_part(h, a, h->stream, 0);   // alone on the GPU: uncapped
    CUDA_TRY(cudaEventRecord(ev[2], h->stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    for (int i = 0; i < 3; i++) cudaEventDestroy(ev[i]);
    h->launches += 2;
    return 0;
}

// Measurement aid for the roofline: a write-only stream over the planes arena (16-byte stores, one per thread,
// nothing read) -- the ceiling of a kernel that, like the step, only writes.  GB/s over `reps` launches.
__global__ void hive_write_stream_kernel(uint4* __restrict__ dst, size_t n_vec, uint32_t tag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_vec) dst[i] = make_uint4(tag, tag ^ (uint32_t)i, tag + 2, threadIdx.x);
}
int hive_probe_write_stream(hive_env_t* h, int reps, double* gbs) {
    if (check(h)) return HIVE_E_HANDLE;
    if (!gbs || reps < 1) return fail(HIVE_E_ARG, "hive_probe_write_stream: bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t n_vec = (size_t)h->n * HIVE_PLANES_ELEMS * 2 / 16;
    const unsigned blocks = (unsigned)((n_vec + 255) / 256);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    for (int i = 0; i < 2; i++) hive_write_stream_kernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<uint4*>(h->planes), n_vec, 1u);
    CUDA_TRY(cudaEventRecord(e0, h->stream));
    for (int i = 0; i < reps; i++) hive_write_stream_kernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<uint4*>(h->planes), n_vec, 2u + i);
    CUDA_TRY(cudaEventRecord(e1, h->stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gbs = (double)n_vec * 16.0 * reps / (ms * 1e-3) / 1e9;
    if (h->shadow) {   // the arena holds the probe pattern: back to the all-zero state the delta store's shadow describes
        CUDA_TRY(cudaMemsetAsync(h->shadow, 0, h->shadow_bytes, h->stream));
        CUDA_TRY(cudaMemsetAsync(h->planes, 0, (size_t)h->n * HIVE_PLANES_ELEMS * 2, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return 0;    // the planes arena no longer holds the games' planes: step or reset the batch before reading planes
}

int hive_legal_host(hive_env_t* h, uint64_t* mask, int32_t* count) {
    if (check(h)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaSetDevice(h->device));
    if (mask) CUDA_TRY(cudaMemcpyAsync(mask, h->legal, (size_t)h->n * LEGAL_WORDS * 4, cudaMemcpyDeviceToHost, h->stream));
    if (count) CUDA_TRY(cudaMemcpyAsync(count, h->count, (size_t)h->n * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

int hive_bits_host(hive_env_t* h, uint32_t* bits) {
    if (check(h)) return HIVE_E_HANDLE;
    if (!bits) return fail(HIVE_E_ARG, "hive_bits_host: null output");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(bits, h->bits[h->last_bits], (size_t)h->n * BITS_WORDS * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

int hive_encode_host(hive_env_t* h, uint16_t* planes_bf16) {
    if (check(h)) return HIVE_E_HANDLE;
    if (!planes_bf16) return fail(HIVE_E_ARG, "hive_encode_host: null output");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(planes_bf16, h->planes, (size_t)h->n * HIVE_PLANES_ELEMS * 2, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

static int fetch_recs(hive_env_t* h, std::vector<GameRec>& out, int first, int cnt) {
    out.resize(cnt);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(out.data(), h->recs + first, (size_t)cnt * sizeof(GameRec), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

int hive_status_packed_host(hive_env_t* h, uint32_t* packed) {
    if (check(h)) return HIVE_E_HANDLE;
    if (!packed) return fail(HIVE_E_ARG, "hive_status_packed_host: null output");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(packed, h->status, (size_t)h->n * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

int hive_status_host(hive_env_t* h, int32_t* turn, int8_t* winner, uint8_t* done) {
    if (check(h)) return HIVE_E_HANDLE;
    std::vector<uint32_t> st(h->n);
    int rc = hive_status_packed_host(h, st.data());
    if (rc) return rc;
    for (int i = 0; i < h->n; i++) {
        if (turn) turn[i] = st[i] & 0xFF;
        if (winner) winner[i] = (int8_t)((st[i] >> 8) & 0xFF);
        if (done) done[i] = (uint8_t)((st[i] >> 16) & 0xFF);
    }
    return 0;
}

static inline uint64_t host_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

}  // extern "C"
void hive::pick_range(int g0, int g1, int n, const uint64_t* mask, const int32_t* count, const uint32_t* packed_status,
                      uint32_t* episodes, uint64_t seed, int max_turn, int32_t* actions) {
    for (int g = g0; g < g1; g++) {
        const uint32_t st = packed_status[g];
        const int turn = st & 0xFF, done = (st >> 16) & 0xFF;
        if (done || turn >= max_turn) { actions[g] = HIVE_RESET; episodes[g]++; continue; }
        if (count[g] <= 0) { actions[g] = -1; continue; }
        const uint64_t gid = (uint64_t)g + (uint64_t)n * episodes[g];
        int k = (int)(host_splitmix64(seed ^ (gid << 32) ^ (uint64_t)turn) % (uint64_t)count[g]);
        const uint64_t* m = mask + (size_t)g * HIVE_LEGAL_U64;
        int a = -1;
        for (int w = 0; w < HIVE_LEGAL_U64; w++) {
            const int c = __builtin_popcountll(m[w]);
            if (k < c) {
                uint64_t x = m[w];
                for (int i = 0; i < k; i++) x &= x - 1;
                a = w * 64 + __builtin_ctzll(x);
                break;
            }
            k -= c;
        }
        actions[g] = a;
    }
}
extern "C" {

// A small persistent worker pool for the host-side policy twin (the caller's thread takes a share too).
// The host-driven loop calls it every few tens of microseconds, far below the wake-up latency of a sleeping
// thread, so idle workers first spin on the ticket word (HIVE_B200_HOST_SPIN_US, default 200 us) and only then
// block on the condition variable.  ticket = epoch<<40 | parts<<20 | next: a part is claimed by a CAS on the
// whole word, so a worker late from the previous job can never claim (or skip) a part of the next one.
namespace {
struct PickPool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work;
    std::function<void(int)> job;          // job(part); written by run() before the ticket is published
    std::atomic<uint64_t> ticket{0};
    std::atomic<int> pending{0}, sleepers{0};
    std::atomic<bool> stop{false};
    long spin_ns = 200000;
    static uint64_t epoch_of(uint64_t t) { return t >> 40; }
    static int parts_of(uint64_t t) { return (int)((t >> 20) & 0xFFFFF); }
    static int next_of(uint64_t t) { return (int)(t & 0xFFFFF); }
    // claims and runs parts of the job published as epoch `e` until none is left
    void drain(uint64_t e) {
        uint64_t t = ticket.load(std::memory_order_acquire);
        for (;;) {
            if (epoch_of(t) != e || next_of(t) >= parts_of(t)) return;
            if (ticket.compare_exchange_weak(t, t + 1, std::memory_order_acq_rel, std::memory_order_acquire)) {
                job(next_of(t));
                pending.fetch_sub(1, std::memory_order_acq_rel);
                t = ticket.load(std::memory_order_acquire);
            }
        }
    }
    explicit PickPool(int nthreads) {
        const char* e = getenv("HIVE_B200_HOST_SPIN_US");
        if (e) spin_ns = atol(e) * 1000L;
        for (int i = 0; i < nthreads; i++)
            workers.emplace_back([this] {
                uint64_t seen = 0;
                for (;;) {
                    // wait for a new epoch: spin first, then sleep
                    const auto t0 = std::chrono::steady_clock::now();
                    uint64_t t;
                    int polls = 0;
                    while (epoch_of(t = ticket.load(std::memory_order_acquire)) == seen && !stop.load(std::memory_order_relaxed)) {
#if defined(__x86_64__) || defined(__i386__)
                        __builtin_ia32_pause();
#endif
                        if ((++polls & 255) == 0 &&
                            std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count() > spin_ns) {
                            std::unique_lock<std::mutex> lk(mu);
                            sleepers.fetch_add(1);
                            cv_work.wait(lk, [&] { return stop.load() || epoch_of(ticket.load(std::memory_order_acquire)) != seen; });
                            sleepers.fetch_sub(1);
                        }
                    }
                    if (stop.load()) return;
                    seen = epoch_of(t);
                    drain(seen);
                }
            });
    }
    ~PickPool() {
        { std::lock_guard<std::mutex> lk(mu); stop.store(true); }
        cv_work.notify_all();
        for (auto& t : workers) t.join();
    }
    // one caller at a time (the library's handles are single-owner; calls from several threads serialise here)
    std::mutex run_mu;
    void run(int nparts, std::function<void(int)> f) {
        std::lock_guard<std::mutex> serial(run_mu);
        job = std::move(f);
        pending.store(nparts, std::memory_order_relaxed);
        const uint64_t e = (epoch_of(ticket.load(std::memory_order_relaxed)) + 1) & 0xFFFFFF;
        ticket.store((e << 40) | ((uint64_t)nparts << 20), std::memory_order_release);
        if (sleepers.load(std::memory_order_acquire) > 0) { std::lock_guard<std::mutex> lk(mu); cv_work.notify_all(); }
        drain(e);                            // the calling thread works too
        while (pending.load(std::memory_order_acquire) > 0) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
    }
};
PickPool* pick_pool() {
    static PickPool* pool = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        unsigned hc = std::thread::hardware_concurrency();
        int nt = hc > 2 ? (int)(hc < 16 ? hc : 16) - 1 : 0;
        const char* e = getenv("HIVE_B200_HOST_THREADS");
        if (e) nt = atoi(e) - 1;
        if (nt < 0) nt = 0;
        pool = new PickPool(nt);
    });
    return pool;
}
}  // namespace

int hive_host_pick_actions(int n, const uint64_t* mask, const int32_t* count, const uint32_t* packed_status,
                           uint32_t* episodes, uint64_t seed, int max_turn, int32_t* actions) {
    if (n <= 0 || !mask || !count || !packed_status || !episodes || !actions)
        return fail(HIVE_E_ARG, "hive_host_pick_actions: bad arguments");
    if (n <= 512) { pick_range(0, n, n, mask, count, packed_status, episodes, seed, max_turn, actions); return 0; }
    // one chunk per thread of the pool (a fixed 512-game chunk left most of the 16 threads idle on a 5,461-game part)
    PickPool* pool = pick_pool();
    const int threads = (int)pool->workers.size() + 1;
    int chunk = (n + threads - 1) / threads;
    if (chunk < 128) chunk = 128;
    const int parts = (n + chunk - 1) / chunk;
    pool->run(parts, [=](int part) {
        const int g0 = part * chunk, g1 = g0 + chunk < n ? g0 + chunk : n;
        pick_range(g0, g1, n, mask, count, packed_status, episodes, seed, max_turn, actions);
    });
    return 0;
}

// The random policy's host twin read from the compact lists (LIST_* in hive_env_kernel.cuh).  Returns in *n_overflow the number
// of games whose group's lists did not fit (their actions are left untouched: pick them from the mask).
int hive_host_pick_actions_lists(int n, const uint8_t* lists, const uint32_t* packed_status, uint32_t* episodes, uint64_t seed,
                                 int max_turn, int32_t* actions, int* n_overflow) {
    if (n <= 0 || !lists || !packed_status || !episodes || !actions) return fail(HIVE_E_ARG, "hive_host_pick_actions_lists: bad arguments");
    std::atomic<int> over{0};
    auto range = [&](int g0, int g1) {
        int ov = 0;
        for (int g = g0; g < g1; g++) {
            const uint8_t* blk = lists + (size_t)(g / SG) * LIST_BLOCK_BYTES;
            const uint8_t* hdr = blk + (g % SG) * LIST_HDR_BYTES;
            const uint32_t st = packed_status[g];
            const int turn = st & 0xFF, done = (st >> 16) & 0xFF;
            if (done || turn >= max_turn) { actions[g] = HIVE_RESET; episodes[g]++; continue; }
            if (hdr[9] & 1u) { ov++; continue; }
            const int count = hdr[8];
            if (count == 0) { actions[g] = -1; continue; }
            const uint64_t gid = (uint64_t)g + (uint64_t)n * episodes[g];
            const int k = (int)(host_splitmix64(seed ^ (gid << 32) ^ (uint64_t)turn) % (uint64_t)count);
            int p = 0;
            while (hdr[2 + p] <= k) p++;                         // the 256-id page of the k-th action (cum[6] = count > k: the scan ends)
            const int off = hdr[0] | (hdr[1] << 8);
            actions[g] = p * 256 + blk[SG * LIST_HDR_BYTES + off + k];
        }
        if (ov) over.fetch_add(ov);
    };
    if (n <= 512) range(0, n);
    else {
        PickPool* pool = pick_pool();
        const int threads = (int)pool->workers.size() + 1;
        int chunk = (n + threads - 1) / threads;
        if (chunk < 128) chunk = 128;
        const int parts = (n + chunk - 1) / chunk;
        pool->run(parts, [&](int part) { const int g0 = part * chunk; range(g0, g0 + chunk < n ? g0 + chunk : n); });
    }
    if (n_overflow) *n_overflow = over.load();
    return 0;
}

int hive_counters_host(hive_env_t* h, uint32_t* steps, uint32_t* episodes) {
    if (check(h)) return HIVE_E_HANDLE;
    std::vector<GameRec> r;
    int rc = fetch_recs(h, r, 0, h->n);
    if (rc) return rc;
    for (int i = 0; i < h->n; i++) {
        if (steps) steps[i] = r[i].steps;
        if (episodes) episodes[i] = r[i].episode;
    }
    return 0;
}

int hive_dump_state(hive_env_t* h, int game, int32_t* turn, uint8_t* cells, uint8_t* levels) {
    if (check(h)) return HIVE_E_HANDLE;
    if (game < 0 || game >= h->n) return fail(HIVE_E_ARG, "hive_dump_state: bad game index");
    std::vector<GameRec> r;
    int rc = fetch_recs(h, r, game, 1);
    if (rc) return rc;
    if (turn) *turn = r[0].turn;
    if (cells) memcpy(cells, r[0].cell, N_PIECE);
    if (levels) memcpy(levels, r[0].level, N_PIECE);
    return 0;
}

int hive_load_state(hive_env_t* h, int game, int turn, const uint8_t* cells, const uint8_t* levels) {
    if (check(h)) return HIVE_E_HANDLE;
    if (game < 0 || game >= h->n || !cells || !levels || turn < 1 || turn > 250)
        return fail(HIVE_E_ARG, "hive_load_state: bad arguments");
    // validate stacks: levels at each cell must be 0..height-1 without gaps
    int height[144] = {0};
    for (int lvl = 0; lvl < 5; lvl++)
        for (int p = 0; p < N_PIECE; p++)
            if (cells[p] != HAND && levels[p] == lvl) {
                if (cells[p] >= 144 || height[cells[p]] != lvl) return fail(HIVE_E_ARG, "hive_load_state: inconsistent stacks");
                height[cells[p]]++;
            }
    for (int p = 0; p < N_PIECE; p++)
        if (cells[p] != HAND && levels[p] > 4) return fail(HIVE_E_ARG, "hive_load_state: level > 4");
    GameRec rec;
    memset(&rec, 0, sizeof rec);
    memcpy(rec.cell, cells, N_PIECE);
    memcpy(rec.level, levels, N_PIECE);
    rec.turn = (uint8_t)turn;
    CUDA_TRY(cudaSetDevice(h->device));
    std::vector<GameRec> old;
    int rc = fetch_recs(h, old, game, 1);
    if (rc) return rc;
    rec.episode = old[0].episode; rec.steps = old[0].steps;
    std::vector<uint8_t> mask(h->n, 0);
    mask[game] = 1;
    CUDA_TRY(cudaMemcpyAsync(h->recs + game, &rec, sizeof rec, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_mask, mask.data(), h->n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return launch_env(h, OP_EVAL, nullptr, h->d_mask, 0, 0, 0, nullptr);
}

int hive_record_host(hive_env_t* h, int game, void* rec384) {
    if (check(h)) return HIVE_E_HANDLE;
    if (game < 0 || game >= h->n || !rec384) return fail(HIVE_E_ARG, "hive_record_host: bad arguments");
    std::vector<GameRec> r;
    int rc = fetch_recs(h, r, game, 1);
    if (rc) return rc;
    memcpy(rec384, &r[0], sizeof(GameRec));
    return 0;
}

int hive_copy_state(hive_env_t* dst, int dst_game, hive_env_t* src, int src_game) {
    if (check(dst) || check(src)) return HIVE_E_HANDLE;
    if (dst_game < 0 || dst_game >= dst->n || src_game < 0 || src_game >= src->n)
        return fail(HIVE_E_ARG, "hive_copy_state: bad game index");
    if (dst->device != src->device) return fail(HIVE_E_ARG, "hive_copy_state: handles on different devices");
    CUDA_TRY(cudaSetDevice(dst->device));
    CUDA_TRY(cudaStreamSynchronize(src->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->recs + dst_game, src->recs + src_game, sizeof(GameRec), cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->legal + (size_t)dst_game * LEGAL_WORDS, src->legal + (size_t)src_game * LEGAL_WORDS,
                             LEGAL_WORDS * 4, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->count + dst_game, src->count + src_game, 4, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->status + dst_game, src->status + src_game, 4, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->planes + (size_t)dst_game * HIVE_PLANES_ELEMS, src->planes + (size_t)src_game * HIVE_PLANES_ELEMS,
                             HIVE_PLANES_ELEMS * 2, cudaMemcpyDeviceToDevice, dst->stream));
    if (dst->shadow) {   // the copied planes' bit image: the source's shadow row, or (full-store source) "unknown" = rewrite all
        if (src->shadow) CUDA_TRY(cudaMemcpyAsync(dst->shadow + (size_t)dst_game * BITS_WORDS, src->shadow + (size_t)src_game * BITS_WORDS,
                                                  BITS_WORDS * 4, cudaMemcpyDeviceToDevice, dst->stream));
        else return fail(HIVE_E_ARG, "hive_copy_state: handles with different plane-store modes");
    }
    return 0;
}

// GamePlay.state_key (env_hive.py:150-168): board_tiles order (q descending, r ascending), '.' or
// the 2-char ids of the stack bottom->top, then the player digit.
int hive_state_key(hive_env_t* h, int game, char* buf, int buflen) {
    if (check(h)) return HIVE_E_HANDLE;
    if (game < 0 || game >= h->n || !buf || buflen < 200) return fail(HIVE_E_ARG, "hive_state_key: bad arguments (buflen >= 200)");
    std::vector<GameRec> r;
    int rc = fetch_recs(h, r, game, 1);
    if (rc) return rc;
    static const char tch[11] = {'Q', 'B', 'B', 'S', 'S', 'G', 'G', 'G', 'A', 'A', 'A'};
    static const char num[11] = {'0', '0', '1', '0', '1', '0', '1', '2', '0', '1', '2'};
    int n = 0;
    for (int q = 11; q >= 0; q--)
        for (int c12 = 0; c12 < 12; c12++) {
            const int c = q * 12 + c12;
            bool any = false;
            for (int lvl = 0; lvl < 5; lvl++)
                for (int p = 0; p < N_PIECE; p++)
                    if (r[0].cell[p] == c && r[0].level[p] == lvl) {
                        char ch = tch[p % 11];
                        if (p >= 11) ch = (char)(ch - 'A' + 'a');
                        buf[n++] = ch; buf[n++] = num[p % 11];
                        any = true;
                    }
            if (!any) buf[n++] = '.';
        }
    buf[n++] = (char)('0' + ((r[0].turn & 1) ? 0 : 1));
    buf[n] = 0;
    return n;
}

void* hive_dev_state(hive_env_t* h) { return h ? h->recs : nullptr; }
void* hive_dev_legal(hive_env_t* h) { return h ? h->legal : nullptr; }
void* hive_dev_count(hive_env_t* h) { return h ? h->count : nullptr; }
void* hive_dev_status(hive_env_t* h) { return h ? h->status : nullptr; }
void* hive_dev_planes(hive_env_t* h) { return h ? h->planes : nullptr; }
long long hive_launch_count(const hive_env_t* h) { return h ? h->launches : 0; }

int hive_set_timing(hive_env_t* h, int on) {
    if (check(h)) return HIVE_E_HANDLE;
    h->timing = on != 0;
    return 0;
}

float hive_last_kernel_ms(hive_env_t* h) {
    if (!h || !h->timing) return 0.f;
    float ms = 0.f;
    if (cudaEventSynchronize(h->t1) != cudaSuccess) return 0.f;
    if (cudaEventElapsedTime(&ms, h->t0, h->t1) != cudaSuccess) return 0.f;
    return ms;
}

}  // extern "C"
