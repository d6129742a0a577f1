// hive_mcts.cu -- C ABI of the batched PUCT search (include/hive_b200.h, mcts_* entry points).
// One tree per game of a hive_env handle; kernels in hive_mcts_kernel.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "hive_internal.h"
#include "hive_mcts_kernel.cuh"

using namespace hive;

struct hive_mcts {
    hive_env* env = nullptr;       // the real games (roots)
    hive_env* sim = nullptr;       // working positions + leaf evaluation (legal mask, planes)
    int n = 0, sims = 0, max_turn = 55, node_cap = 0, edge_cap = 0, ht_size = 0;
    int noise_rows = 0, noise_cols = 0;
    uint64_t noise_seed = 0x5EEDull;
    MctsTree* trees = nullptr;
    MctsNode* nodes = nullptr;
    int32_t* htab = nullptr;
    int16_t* e_action = nullptr; int32_t* e_n = nullptr; double* e_w = nullptr; double* e_q = nullptr; float* e_p = nullptr;
    float* leaf_p = nullptr; double* leaf_v = nullptr;
    uint8_t* need_eval = nullptr; uint8_t* env_mask = nullptr; uint8_t* tree_mask = nullptr; bool use_mask = false;
    int32_t* pending = nullptr;
    double* noise = nullptr;
    double* pi = nullptr; int32_t* out_action = nullptr; int32_t* out_sum_n = nullptr;
    uint32_t* search_no = nullptr; uint32_t* error_any = nullptr;
    long long launches = 0;
};

namespace {

MctsArgs make_args(hive_mcts* m) {
    MctsArgs a;
    a.n = m->n; a.sims = m->sims; a.max_turn = m->max_turn; a.node_cap = m->node_cap; a.edge_cap = m->edge_cap;
    a.ht_size = m->ht_size; a.noise_rows = m->noise_rows; a.noise_cols = m->noise_cols; a.noise_seed = m->noise_seed;
    a.root_recs = m->env->recs; a.sim_recs = m->sim->recs; a.sim_legal = m->sim->legal; a.sim_count = m->sim->count; a.sim_planes = m->sim->planes;
    a.root_legal = m->env->legal; a.root_count = m->env->count; a.root_planes = m->env->planes; a.env_mask = m->env_mask;
    a.root_shadow = m->env->shadow; a.sim_shadow = m->env->shadow ? m->sim->shadow : nullptr;
    a.leaf_p = m->leaf_p; a.leaf_v = m->leaf_v; a.need_eval = m->need_eval;
    a.tree_mask = m->use_mask ? m->tree_mask : nullptr; a.pending = m->pending;
    a.trees = m->trees; a.nodes = m->nodes; a.htab = m->htab;
    a.e_action = m->e_action; a.e_n = m->e_n; a.e_w = m->e_w; a.e_q = m->e_q; a.e_p = m->e_p;
    a.noise = m->noise; a.pi = m->pi; a.out_action = m->out_action; a.out_sum_n = m->out_sum_n;
    a.search_no = m->search_no; a.error_any = m->error_any;
    return a;
}
int check(const hive_mcts* m) { return m && m->n > 0 ? 0 : fail(HIVE_E_HANDLE, "bad mcts handle"); }
int blocks_for(int n) { return (n + MCTS_WARPS - 1) / MCTS_WARPS; }
int search_failed(uint32_t flags) {
    return fail(HIVE_E_SEARCH, std::string("search gave up on at least one tree:") + ((flags & 2u) ? " node arena full;" : "") +
                                   ((flags & 4u) ? " edge arena full (raise edges_per_sim in mcts_create);" : "") +
                                   ((flags & 8u) ? " simulation deeper than MCTS_MAX_DEPTH;" : "") +
                                   " the statistics of the flagged trees hold fewer simulations than asked for");
}

}  // namespace

extern "C" {

static int mcts_alloc(hive_mcts* m) {
    hive_env* env = m->env;
    const size_t n = (size_t)m->n;
    CUDA_TRY(cudaMalloc(&m->trees, n * sizeof(MctsTree)));
    CUDA_TRY(cudaMalloc(&m->nodes, n * m->node_cap * sizeof(MctsNode)));
    CUDA_TRY(cudaMalloc(&m->htab, n * m->ht_size * 4));
    CUDA_TRY(cudaMalloc(&m->e_action, n * m->edge_cap * 2));
    CUDA_TRY(cudaMalloc(&m->e_n, n * m->edge_cap * 4));
    CUDA_TRY(cudaMalloc(&m->e_w, n * m->edge_cap * 8));
    CUDA_TRY(cudaMalloc(&m->e_q, n * m->edge_cap * 8));
    CUDA_TRY(cudaMalloc(&m->e_p, n * m->edge_cap * 4));
    CUDA_TRY(cudaMalloc(&m->leaf_p, n * 1584 * 4));
    CUDA_TRY(cudaMalloc(&m->leaf_v, n * 8));
    CUDA_TRY(cudaMalloc(&m->need_eval, n));
    CUDA_TRY(cudaMalloc(&m->env_mask, n));
    CUDA_TRY(cudaMemsetAsync(m->env_mask, 0, n, env->stream));
    CUDA_TRY(cudaMalloc(&m->tree_mask, n));
    CUDA_TRY(cudaMalloc(&m->pending, 4));
    CUDA_TRY(cudaMalloc(&m->pi, n * 1584 * 8));
    CUDA_TRY(cudaMalloc(&m->out_action, n * 4));
    CUDA_TRY(cudaMalloc(&m->out_sum_n, n * 4));
    CUDA_TRY(cudaMalloc(&m->search_no, n * 4));
    CUDA_TRY(cudaMalloc(&m->error_any, 4));
    CUDA_TRY(cudaMemsetAsync(m->search_no, 0, n * 4, env->stream));
    CUDA_TRY(cudaMemsetAsync(m->error_any, 0, 4, env->stream));
    CUDA_TRY(cudaMemsetAsync(m->trees, 0, n * sizeof(MctsTree), env->stream));
    CUDA_TRY(cudaMemsetAsync(m->need_eval, 0, n, env->stream));
    CUDA_TRY(cudaMemsetAsync(m->leaf_p, 0, n * 1584 * 4, env->stream));
    CUDA_TRY(cudaMemsetAsync(m->leaf_v, 0, n * 8, env->stream));
    CUDA_TRY(cudaStreamSynchronize(env->stream));
    return 0;
}

int mcts_create(hive_env_t* env, int sims, int edges_per_sim, hive_mcts_t** out) {
    if (!env || env->n <= 0 || sims < 1 || !out) return fail(HIVE_E_ARG, "mcts_create: bad arguments");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(env->device));
    hive_mcts* m = new hive_mcts();
    m->env = env; m->n = env->n; m->sims = sims;
    int rc = create_env(env->n, env->device, env->stream, 1, &m->sim);   // same stream: strict ordering; one slice
    if (rc) { delete m; return rc; }
    m->node_cap = sims + 1;
    // Edge arena: a node stores one edge per legal action.  The reference's own 1,000-game run peaks at 130 legal
    // actions and averages ~51; 160 per simulation keeps a tree whose EVERY node sits at that peak inside
    // the arena.  A tree that still runs out stops, is flagged, and mcts_policy_host fails loudly (HIVE_E_SEARCH).
    if (edges_per_sim <= 0) edges_per_sim = 160;
    m->edge_cap = sims * edges_per_sim + 256;
    int ht = 64;
    while (ht < 2 * m->node_cap) ht <<= 1;
    m->ht_size = ht;
    rc = mcts_alloc(m);
    if (rc) { const std::string why = hive_last_error(); mcts_destroy(m); return fail(rc, why); }
    *out = m;
    return 0;
}

int mcts_destroy(hive_mcts_t* m) {
    if (!m) return 0;
    cudaSetDevice(m->env->device);
    cudaStreamSynchronize(m->env->stream);
    if (m->sim) hive_destroy(m->sim);
    cudaFree(m->trees); cudaFree(m->nodes); cudaFree(m->htab); cudaFree(m->e_action); cudaFree(m->e_n);
    cudaFree(m->e_w); cudaFree(m->e_q); cudaFree(m->e_p); cudaFree(m->leaf_p); cudaFree(m->leaf_v);
    cudaFree(m->need_eval); cudaFree(m->env_mask); cudaFree(m->tree_mask); cudaFree(m->pending); cudaFree(m->noise); cudaFree(m->pi);
    cudaFree(m->out_action); cudaFree(m->out_sum_n); cudaFree(m->search_no); cudaFree(m->error_any);
    delete m;
    return 0;
}

int mcts_set_params(hive_mcts_t* m, int sims, int max_turn, uint64_t noise_seed) {
    if (check(m)) return HIVE_E_HANDLE;
    if (sims < 1 || sims + 1 > m->node_cap || max_turn < 1 || max_turn > 250)
        return fail(HIVE_E_ARG, "mcts_set_params: sims exceeds the capacity given to mcts_create, or bad max_turn");
    m->sims = sims; m->max_turn = max_turn; m->noise_seed = noise_seed;
    return 0;
}

int mcts_set_root_noise_host(hive_mcts_t* m, const double* noise, int rows, int cols) {
    if (check(m)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaSetDevice(m->env->device));
    CUDA_TRY(cudaStreamSynchronize(m->env->stream));
    cudaFree(m->noise); m->noise = nullptr; m->noise_rows = m->noise_cols = 0;
    if (!noise) return 0;
    if (rows < 1 || cols < 1) return fail(HIVE_E_ARG, "mcts_set_root_noise_host: bad shape");
    const size_t bytes = (size_t)m->n * rows * cols * 8;
    CUDA_TRY(cudaMalloc(&m->noise, bytes));
    CUDA_TRY(cudaMemcpy(m->noise, noise, bytes, cudaMemcpyHostToDevice));
    m->noise_rows = rows; m->noise_cols = cols;
    return 0;
}

int mcts_begin(hive_mcts_t* m, const uint8_t* tree_mask) {
    if (check(m)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaSetDevice(m->env->device));
    cudaStream_t st = m->env->stream;
    m->use_mask = tree_mask != nullptr;
    if (tree_mask) {
        CUDA_TRY(cudaMemcpyAsync(m->tree_mask, tree_mask, m->n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    CUDA_TRY(cudaMemsetAsync(m->error_any, 0, 4, st));
    MctsArgs a = make_args(m);
    mcts_reset_kernel<<<m->n, MCTS_WARPS * 32, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    m->launches++;
    return 0;
}

int mcts_descend(hive_mcts_t* m, int* n_pending) {
    if (check(m)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaSetDevice(m->env->device));
    cudaStream_t st = m->env->stream;
    CUDA_TRY(cudaMemsetAsync(m->pending, 0, 4, st));
    MctsArgs a = make_args(m);
    mcts_descend_kernel<<<blocks_for(m->n), MCTS_WARPS * 32, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    m->launches++;
    // evaluate the leaf positions: legal mask + planes (only the trees that asked)
    int rc = launch_env(m->sim, /*OP_EVAL*/ 2, nullptr, m->env_mask, 0, 0, 0, nullptr);
    if (rc) return rc;
    if (n_pending) {
        CUDA_TRY(cudaMemcpyAsync(n_pending, m->pending, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 0;
}

int mcts_expand(hive_mcts_t* m) {
    if (check(m)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaSetDevice(m->env->device));
    MctsArgs a = make_args(m);
    mcts_expand_kernel<<<blocks_for(m->n), MCTS_WARPS * 32, 0, m->env->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    m->launches++;
    return 0;
}

void* mcts_dev_leaf_planes(hive_mcts_t* m) { return m ? m->sim->planes : nullptr; }
void* mcts_dev_leaf_policy(hive_mcts_t* m) { return m ? m->leaf_p : nullptr; }
void* mcts_dev_leaf_value(hive_mcts_t* m) { return m ? m->leaf_v : nullptr; }
void* mcts_dev_pending_mask(hive_mcts_t* m) { return m ? m->need_eval : nullptr; }
long long mcts_launch_count(const hive_mcts_t* m) { return m ? m->launches + m->sim->launches : 0; }

int mcts_leaf_planes_host(hive_mcts_t* m, uint16_t* planes_bf16, uint8_t* pending_mask) {
    if (check(m)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaSetDevice(m->env->device));
    cudaStream_t st = m->env->stream;
    if (planes_bf16) CUDA_TRY(cudaMemcpyAsync(planes_bf16, m->sim->planes, (size_t)m->n * HIVE_PLANES_ELEMS * 2, cudaMemcpyDeviceToHost, st));
    if (pending_mask) CUDA_TRY(cudaMemcpyAsync(pending_mask, m->need_eval, m->n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int mcts_set_leaf_eval_host(hive_mcts_t* m, const float* policy, const double* value) {
    if (check(m)) return HIVE_E_HANDLE;
    if (!policy || !value) return fail(HIVE_E_ARG, "mcts_set_leaf_eval_host: null input");
    CUDA_TRY(cudaSetDevice(m->env->device));
    cudaStream_t st = m->env->stream;
    CUDA_TRY(cudaMemcpyAsync(m->leaf_p, policy, (size_t)m->n * 1584 * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(m->leaf_v, value, (size_t)m->n * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int mcts_pending_host(hive_mcts_t* m, int* n_pending) {
    if (check(m)) return HIVE_E_HANDLE;
    if (!n_pending) return fail(HIVE_E_ARG, "mcts_pending_host: null output");
    CUDA_TRY(cudaSetDevice(m->env->device));
    CUDA_TRY(cudaMemcpyAsync(n_pending, m->pending, 4, cudaMemcpyDeviceToHost, m->env->stream));
    CUDA_TRY(cudaStreamSynchronize(m->env->stream));
    return 0;
}

int mcts_policy_host(hive_mcts_t* m, double* pi, int32_t* action, int32_t* sum_n) {
    if (check(m)) return HIVE_E_HANDLE;
    CUDA_TRY(cudaSetDevice(m->env->device));
    cudaStream_t st = m->env->stream;
    MctsArgs a = make_args(m);
    mcts_finalize_kernel<<<blocks_for(m->n), MCTS_WARPS * 32, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    m->launches++;
    if (pi) CUDA_TRY(cudaMemcpyAsync(pi, m->pi, (size_t)m->n * 1584 * 8, cudaMemcpyDeviceToHost, st));
    if (action) CUDA_TRY(cudaMemcpyAsync(action, m->out_action, (size_t)m->n * 4, cudaMemcpyDeviceToHost, st));
    if (sum_n) CUDA_TRY(cudaMemcpyAsync(sum_n, m->out_sum_n, (size_t)m->n * 4, cudaMemcpyDeviceToHost, st));
    uint32_t flags = 0;
    CUDA_TRY(cudaMemcpyAsync(&flags, m->error_any, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (flags) return search_failed(flags);
    return 0;
}

int mcts_error_host(hive_mcts_t* m, uint32_t* flags) {
    if (check(m)) return HIVE_E_HANDLE;
    if (!flags) return fail(HIVE_E_ARG, "mcts_error_host: null output");
    CUDA_TRY(cudaSetDevice(m->env->device));
    CUDA_TRY(cudaMemcpyAsync(flags, m->error_any, 4, cudaMemcpyDeviceToHost, m->env->stream));
    CUDA_TRY(cudaStreamSynchronize(m->env->stream));
    return 0;
}

// ---- deterministic stand-in network on the device (parity tests of the device leaf-evaluation path): a pure function
// of the leaf's bf16 planes and a salt.  The NumPy twin is oracle/mcts_oracle.py::device_hash_net.
//   key  = splitmix64(salt ^ sum_j splitmix64(j<<16 | planes[j]))          (sum mod 2^64, order-free)
//   p[a] = x^6, x = (splitmix64(key ^ (a+1)*0x9E3779B97F4A7C15) >> 40) * 2^-24   (float32 products x2 = x*x, x4 = x2*x2, x4*x2)
//   v    = (splitmix64(key ^ 0x5bf03635) >> 11) * 2^-52 - 1                    (double)
__global__ void __launch_bounds__(128) mcts_hash_eval_kernel(const uint16_t* __restrict__ planes, float* __restrict__ policy,
                                                             double* __restrict__ value, const uint8_t* __restrict__ mask, uint64_t salt) {
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    __shared__ uint64_t part[4];
    const uint16_t* src = planes + (size_t)b * HIVE_PLANES_ELEMS;
    uint64_t h = 0;
    for (int j = threadIdx.x; j < HIVE_PLANES_ELEMS; j += 128) h += splitmix64(((uint64_t)j << 16) | (uint64_t)src[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(FULL, h, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = h;
    __syncthreads();
    const uint64_t key = splitmix64(salt ^ (part[0] + part[1] + part[2] + part[3]));
    for (int a = threadIdx.x; a < 1584; a += 128) {
        const float x = (float)(uint32_t)(splitmix64(key ^ ((uint64_t)(a + 1) * 0x9E3779B97F4A7C15ULL)) >> 40) * (1.0f / 16777216.0f);
        const float x2 = __fmul_rn(x, x), x4 = __fmul_rn(x2, x2);
        policy[(size_t)b * 1584 + a] = __fmul_rn(x4, x2);
    }
    if (threadIdx.x == 0) value[b] = __dsub_rn((double)(splitmix64(key ^ 0x5bf03635ULL) >> 11) * (1.0 / 4503599627370496.0), 1.0);
}

int mcts_hash_eval_dev(const uint16_t* planes_dev, float* policy_dev, double* value_dev, const uint8_t* mask_dev, int n,
                       uint64_t salt, void* stream) {
    if (!planes_dev || !policy_dev || !value_dev || n < 1) return fail(HIVE_E_ARG, "mcts_hash_eval_dev: bad arguments");
    mcts_hash_eval_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(planes_dev, policy_dev, value_dev, mask_dev, salt);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Shape of the trees of the last search, averaged over the first `max_trees` trees that hold nodes (read back and
// reduced on the host; a measurement call, not on any hot path): out[0] = mean edges per node (E), out[1] = mean
// select depth per simulation (D: every select bumps the sum_n of the node it leaves, so sum over nodes of sum_n = all
// selects), out[2] = mean nodes per tree, out[3] = mean simulations per tree.
int mcts_tree_stats_host(hive_mcts_t* m, int max_trees, double* out) {
    if (!m || !out || max_trees < 1) return fail(HIVE_E_ARG, "mcts_tree_stats_host: bad arguments");
    CUDA_TRY(cudaSetDevice(m->env->device));
    CUDA_TRY(cudaStreamSynchronize(m->env->stream));
    const int nt = max_trees < m->n ? max_trees : m->n;
    std::vector<MctsTree> trees(nt);
    std::vector<MctsNode> nodes((size_t)nt * m->node_cap);
    CUDA_TRY(cudaMemcpy(trees.data(), m->trees, (size_t)nt * sizeof(MctsTree), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(nodes.data(), m->nodes, nodes.size() * sizeof(MctsNode), cudaMemcpyDeviceToHost));
    double n_nodes = 0, n_edges = 0, selects = 0, sims = 0, used = 0;
    for (int t = 0; t < nt; t++) {
        if (trees[t].n_nodes <= 0) continue;
        used += 1; n_nodes += trees[t].n_nodes; sims += trees[t].sims_done;
        for (int i = 0; i < trees[t].n_nodes && i < m->node_cap; i++) {
            const MctsNode& nd = nodes[(size_t)t * m->node_cap + i];
            n_edges += nd.n_edges; selects += nd.sum_n;
        }
    }
    out[0] = n_nodes > 0 ? n_edges / n_nodes : 0.0;
    out[1] = sims > 0 ? selects / sims : 0.0;
    out[2] = used > 0 ? n_nodes / used : 0.0;
    out[3] = used > 0 ? sims / used : 0.0;
    return 0;
}

void* mcts_stream(hive_mcts_t* m) { return m ? (void*)m->env->stream : nullptr; }

int mcts_root_stats_host(hive_mcts_t* m, int tree, int max_edges, int32_t* action, int32_t* N, double* W, double* Q,
                         float* P, int32_t* info /*[6]: n_edges, sum_n, n_nodes, sims_done, error, root_selects*/) {
    if (check(m)) return HIVE_E_HANDLE;
    if (tree < 0 || tree >= m->n || !info) return fail(HIVE_E_ARG, "mcts_root_stats_host: bad arguments");
    CUDA_TRY(cudaSetDevice(m->env->device));
    CUDA_TRY(cudaStreamSynchronize(m->env->stream));
    MctsTree T; MctsNode nd;
    CUDA_TRY(cudaMemcpy(&T, m->trees + tree, sizeof T, cudaMemcpyDeviceToHost));
    memset(info, 0, 6 * 4);
    info[2] = T.n_nodes; info[3] = T.sims_done; info[4] = T.error; info[5] = T.root_selects;
    if (T.n_nodes == 0) return 0;
    CUDA_TRY(cudaMemcpy(&nd, m->nodes + (size_t)tree * m->node_cap, sizeof nd, cudaMemcpyDeviceToHost));
    info[0] = nd.n_edges; info[1] = nd.sum_n;
    const int k = nd.n_edges < max_edges ? nd.n_edges : max_edges;
    const size_t e0 = (size_t)tree * m->edge_cap + nd.edge_off;
    std::vector<int16_t> act(k);
    if (k > 0) {
        CUDA_TRY(cudaMemcpy(act.data(), m->e_action + e0, k * 2, cudaMemcpyDeviceToHost));
        if (action) for (int i = 0; i < k; i++) action[i] = act[i];
        if (N) CUDA_TRY(cudaMemcpy(N, m->e_n + e0, k * 4, cudaMemcpyDeviceToHost));
        if (W) CUDA_TRY(cudaMemcpy(W, m->e_w + e0, k * 8, cudaMemcpyDeviceToHost));
        if (Q) CUDA_TRY(cudaMemcpy(Q, m->e_q + e0, k * 8, cudaMemcpyDeviceToHost));
        if (P) CUDA_TRY(cudaMemcpy(P, m->e_p + e0, k * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

}  // extern "C"
