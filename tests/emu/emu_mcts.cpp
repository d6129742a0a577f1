// emu_mcts.cpp -- runs the PUCT search kernels (verbatim device source) on the CPU SIMT emulator.
// TEST INFRASTRUCTURE.
#include "cuda_emu.h"
#include <vector>
#include "../../hive-alphazero_b200/csrc/hive_mcts_kernel.cuh"

using namespace hive;

extern "C" int emu_env_run(void* recs, uint32_t* legal, int32_t* count, uint32_t* status, uint16_t* planes, int n, int op,
                           const int32_t* actions, const uint8_t* mask, uint64_t seed, int max_turn, int auto_reset,
                           int32_t* chosen, uint64_t sched_seed, uint32_t* shadow);

struct EmuMcts {
    int n, sims, node_cap, edge_cap, ht_size;
    std::vector<GameRec> sim_recs;
    std::vector<uint32_t> sim_legal, sim_status;
    std::vector<int32_t> sim_count;
    std::vector<uint16_t> sim_planes;
    std::vector<uint32_t> sim_shadow;                       // delta plane store: bit image of sim_planes
    const uint32_t* root_shadow = nullptr;
    std::vector<MctsTree> trees;
    std::vector<MctsNode> nodes;
    std::vector<int32_t> htab, e_n, out_action, out_sum_n;
    std::vector<int16_t> e_action;
    std::vector<double> e_w, e_q, leaf_v, noise, pi;
    std::vector<float> e_p, leaf_p;
    std::vector<uint8_t> need_eval, env_mask;
    std::vector<uint32_t> search_no;
    uint32_t error_any = 0;
    const uint32_t* root_legal = nullptr; const int32_t* root_count = nullptr; const uint16_t* root_planes = nullptr;
    int32_t pending;
    int noise_rows = 0, noise_cols = 0;
    const GameRec* root = nullptr;
    uint64_t sched = 1;
};

static MctsArgs args_of(EmuMcts* m) {
    MctsArgs a;
    a.n = m->n; a.sims = m->sims; a.max_turn = 55; a.node_cap = m->node_cap; a.edge_cap = m->edge_cap; a.ht_size = m->ht_size;
    a.noise_rows = m->noise_rows; a.noise_cols = m->noise_cols; a.noise_seed = 7;
    a.root_recs = m->root; a.sim_recs = m->sim_recs.data(); a.sim_legal = m->sim_legal.data(); a.sim_count = m->sim_count.data(); a.sim_planes = m->sim_planes.data();
    a.root_legal = m->root_legal; a.root_count = m->root_count; a.root_planes = m->root_planes; a.env_mask = m->env_mask.data();
    a.root_shadow = m->root_shadow; a.sim_shadow = m->root_shadow ? m->sim_shadow.data() : nullptr;
    a.leaf_p = m->leaf_p.data(); a.leaf_v = m->leaf_v.data(); a.need_eval = m->need_eval.data(); a.tree_mask = nullptr;
    a.pending = &m->pending; a.trees = m->trees.data(); a.nodes = m->nodes.data(); a.htab = m->htab.data();
    a.e_action = m->e_action.data(); a.e_n = m->e_n.data(); a.e_w = m->e_w.data(); a.e_q = m->e_q.data(); a.e_p = m->e_p.data();
    a.noise = m->noise_rows ? m->noise.data() : nullptr; a.pi = m->pi.data();
    a.out_action = m->out_action.data(); a.out_sum_n = m->out_sum_n.data();
    a.search_no = m->search_no.data(); a.error_any = &m->error_any;
    return a;
}
static void k_reset(void* p) { mcts_reset_kernel(*(MctsArgs*)p); }
static void k_descend(void* p) { mcts_descend_kernel(*(MctsArgs*)p); }
static void k_expand(void* p) { mcts_expand_kernel(*(MctsArgs*)p); }
static void k_final(void* p) { mcts_finalize_kernel(*(MctsArgs*)p); }

extern "C" {

void* emu_mcts_create(int n, int sims, int edges_per_sim) {
    EmuMcts* m = new EmuMcts();
    m->n = n; m->sims = sims; m->node_cap = sims + 1; m->edge_cap = sims * edges_per_sim + 256;
    int ht = 64; while (ht < 2 * m->node_cap) ht <<= 1; m->ht_size = ht;
    m->sim_recs.resize(n); m->sim_legal.resize(n * 50); m->sim_status.resize(n); m->sim_count.resize(n);
    m->sim_planes.resize((size_t)n * 56 * 144); m->sim_shadow.resize((size_t)n * BITS_WORDS); m->trees.resize(n); m->nodes.resize((size_t)n * m->node_cap);
    m->htab.resize((size_t)n * ht); m->e_n.resize((size_t)n * m->edge_cap); m->e_action.resize((size_t)n * m->edge_cap);
    m->e_w.resize((size_t)n * m->edge_cap); m->e_q.resize((size_t)n * m->edge_cap); m->e_p.resize((size_t)n * m->edge_cap);
    m->leaf_p.resize((size_t)n * 1584); m->leaf_v.resize(n); m->need_eval.resize(n); m->env_mask.resize(n); m->pi.resize((size_t)n * 1584);
    m->out_action.resize(n); m->out_sum_n.resize(n); m->search_no.assign(n, 0);
    memset(m->sim_recs.data(), 0, n * sizeof(GameRec));
    return m;
}
void emu_mcts_destroy(void* h) { delete (EmuMcts*)h; }
void emu_mcts_set_noise(void* h, const double* noise, int rows, int cols) {
    EmuMcts* m = (EmuMcts*)h;
    m->noise.assign(noise, noise + (size_t)m->n * rows * cols); m->noise_rows = rows; m->noise_cols = cols;
}
int emu_mcts_begin(void* h, const void* root_recs, const uint32_t* root_legal, const int32_t* root_count, const uint16_t* root_planes,
                   const uint32_t* root_shadow) {
    EmuMcts* m = (EmuMcts*)h;
    m->root_shadow = root_shadow;
    m->root = (const GameRec*)root_recs; m->root_legal = root_legal; m->root_count = root_count; m->root_planes = root_planes;
    m->error_any = 0;
    MctsArgs a = args_of(m);
    for (int b = 0; b < m->n; b++) { int rc = emu::run_block(k_reset, &a, b, MCTS_WARPS * 32, m->sched++); if (rc) return rc; }
    return 0;
}
int emu_mcts_descend(void* h, int* pending) {
    EmuMcts* m = (EmuMcts*)h;
    m->pending = 0;
    MctsArgs a = args_of(m);
    const int blocks = (m->n + MCTS_WARPS - 1) / MCTS_WARPS;
    for (int b = 0; b < blocks; b++) { int rc = emu::run_block(k_descend, &a, b, MCTS_WARPS * 32, m->sched++); if (rc) return rc; }
    int rc = emu_env_run(m->sim_recs.data(), m->sim_legal.data(), m->sim_count.data(), m->sim_status.data(), m->sim_planes.data(),
                         m->n, /*OP_EVAL*/ 2, nullptr, m->env_mask.data(), 0, 55, 0, nullptr, m->sched++, m->root_shadow ? m->sim_shadow.data() : nullptr);
    if (rc) return rc;
    *pending = m->pending;
    return 0;
}
const uint16_t* emu_mcts_planes(void* h) { return ((EmuMcts*)h)->sim_planes.data(); }
const uint8_t* emu_mcts_pending_mask(void* h) { return ((EmuMcts*)h)->need_eval.data(); }
void emu_mcts_set_leaf(void* h, const float* p, const double* v) {
    EmuMcts* m = (EmuMcts*)h;
    memcpy(m->leaf_p.data(), p, (size_t)m->n * 1584 * 4); memcpy(m->leaf_v.data(), v, (size_t)m->n * 8);
}
int emu_mcts_expand(void* h) {
    EmuMcts* m = (EmuMcts*)h;
    MctsArgs a = args_of(m);
    const int blocks = (m->n + MCTS_WARPS - 1) / MCTS_WARPS;
    for (int b = 0; b < blocks; b++) { int rc = emu::run_block(k_expand, &a, b, MCTS_WARPS * 32, m->sched++); if (rc) return rc; }
    return 0;
}
int emu_mcts_finalize(void* h, double* pi, int32_t* action, int32_t* sum_n) {
    EmuMcts* m = (EmuMcts*)h;
    MctsArgs a = args_of(m);
    const int blocks = (m->n + MCTS_WARPS - 1) / MCTS_WARPS;
    for (int b = 0; b < blocks; b++) { int rc = emu::run_block(k_final, &a, b, MCTS_WARPS * 32, m->sched++); if (rc) return rc; }
    memcpy(pi, m->pi.data(), (size_t)m->n * 1584 * 8); memcpy(action, m->out_action.data(), m->n * 4); memcpy(sum_n, m->out_sum_n.data(), m->n * 4);
    return 0;
}
unsigned emu_mcts_error(void* h) { return ((EmuMcts*)h)->error_any; }
int emu_mcts_root(void* h, int t, int max_edges, int32_t* action, int32_t* N, double* W, double* Q, float* P, int32_t* info) {
    EmuMcts* m = (EmuMcts*)h;
    const MctsTree& T = m->trees[t];
    memset(info, 0, 24);
    info[2] = T.n_nodes; info[3] = T.sims_done; info[4] = T.error; info[5] = T.root_selects;
    if (!T.n_nodes) return 0;
    const MctsNode& nd = m->nodes[(size_t)t * m->node_cap];
    info[0] = nd.n_edges; info[1] = nd.sum_n;
    size_t e0 = (size_t)t * m->edge_cap + nd.edge_off;
    for (int i = 0; i < nd.n_edges && i < max_edges; i++) {
        action[i] = m->e_action[e0 + i]; N[i] = m->e_n[e0 + i]; W[i] = m->e_w[e0 + i]; Q[i] = m->e_q[e0 + i]; P[i] = m->e_p[e0 + i];
    }
    return 0;
}
}
