// emu_env.cpp -- runs hive_env_kernel (verbatim device source) on the CPU warp emulator.
// TEST INFRASTRUCTURE: lets `pytest -m "not gpu"` differential-test the kernel logic against the
// oracle.  Never loaded by the product package.
#include "cuda_emu.h"
#include <vector>
#include "../../hive-alphazero_b200/csrc/hive_env_kernel.cuh"
#include "../../hive-alphazero_b200/csrc/hive_tables.h"

using namespace hive;

static std::vector<uint32_t> g_lines;
static std::vector<uint8_t> g_lists;        // compact legal lists written by the last emu_env_run (EnvArgs::lists)
static void build_lines() {
    if (g_lines.empty()) build_geometry_tables(g_lines);
}

static void k_step(void* p) { hive_step_kernel(*(EnvArgs*)p); }
static void k_planes(void* p) { hive_planes_kernel(*(EnvArgs*)p); }
static void k_planes_delta(void* p) { hive_planes_delta_kernel(*(EnvArgs*)p); }

extern "C" {

// state: n x 384 B records; legal: n x 50 u32; count: n; planes: n x 8064 u16; shadow: n x 280 u32 (the bit image
// of `planes`, owned by the caller: the delta plane store, the product's default) or null (full plane store)
int emu_env_run(void* recs, uint32_t* legal, int32_t* count, uint32_t* status, uint16_t* planes, int n, int op,
                const int32_t* actions, const uint8_t* mask, uint64_t seed, int max_turn, int auto_reset,
                int32_t* chosen, uint64_t sched_seed, uint32_t* shadow) {
    build_lines();
    EnvArgs a;
    a.recs = (GameRec*)recs; a.legal = legal; a.count = count; a.status = status; a.planes = planes;
    a.actions = actions; a.mask = mask; a.chosen = chosen; a.hop_lines = g_lines.data();
    a.seed = seed; a.n = n; a.op = op; a.max_turn = max_turn; a.auto_reset = auto_reset;
    a.g_offset = 0; a.n_total = n; a.stagger_ns = 0; a.stagger_div = 148;
    static std::vector<uint32_t> bits;
    if ((int)bits.size() < n * BITS_WORDS) bits.resize((size_t)n * BITS_WORDS);
    a.bits = bits.data(); a.shadow = shadow;
    g_lists.assign((size_t)((n + SG - 1) / SG) * LIST_BLOCK_BYTES, 0xEE);
    a.lists = (op == OP_STEP || op == OP_RANDOM || op == OP_RESET) ? g_lists.data() : nullptr;   // (the search's evaluations run without)
    const int blocks = (n + SG - 1) / SG;
    emu::g_gridDim.x = blocks;
    for (int b = 0; b < blocks; b++) {
        int rc = emu::run_block(k_step, &a, b, STEP_THREADS, sched_seed + (uint64_t)b);
        if (rc) return rc;
    }
    if (shadow) {
        int delta_blocks = (n + HIVE_DELTA_WARPS - 1) / HIVE_DELTA_WARPS;
        if (delta_blocks > 3) delta_blocks = 3;     // capped grid, as on the GPU
        emu::g_gridDim.x = delta_blocks;
        for (int b = 0; b < delta_blocks; b++) {
            int rc = emu::run_block(k_planes_delta, &a, b, HIVE_DELTA_WARPS * 32, sched_seed + 3000 + (uint64_t)b);
            if (rc) return rc;
        }
        return 0;
    }
    int store_blocks = (n + HIVE_STORE_WARPS - 1) / HIVE_STORE_WARPS;
    if (store_blocks > 2) store_blocks = 2;     // capped grid: warps walk over several games like the persistent launch on the GPU
    emu::g_gridDim.x = store_blocks;
    for (int b = 0; b < store_blocks; b++) {
        int rc = emu::run_block(k_planes, &a, b, HIVE_STORE_WARPS * 32, sched_seed + 3000 + (uint64_t)b);
        if (rc) return rc;
    }
    return 0;
}
// n_steps OP_RANDOM steps through the rollout kernel (every CTA loops over its steps and stores its planes itself)
int emu_env_rollout(void* recs, uint32_t* legal, int32_t* count, uint32_t* status, uint16_t* planes, int n, uint64_t seed,
                    int max_turn, int auto_reset, int n_steps, uint64_t sched_seed) {
    build_lines();
    EnvArgs a;
    a.recs = (GameRec*)recs; a.legal = legal; a.count = count; a.status = status; a.planes = planes;
    a.actions = nullptr; a.mask = nullptr; a.chosen = nullptr; a.hop_lines = g_lines.data();
    a.seed = seed; a.n = n; a.op = OP_RANDOM; a.max_turn = max_turn; a.auto_reset = auto_reset;
    a.g_offset = 0; a.n_total = n; a.stagger_ns = 0; a.stagger_div = 148; a.lists = nullptr;
    static std::vector<uint32_t> bits;
    if ((int)bits.size() < n * BITS_WORDS) bits.resize((size_t)n * BITS_WORDS);
    a.bits = bits.data(); a.shadow = nullptr;
    struct Call { EnvArgs a; int n_steps; } call = {a, n_steps};
    const int blocks = (n + SG - 1) / SG;
    emu::g_gridDim.x = blocks;
    for (int b = 0; b < blocks; b++) {
        int rc = emu::run_block([](void* p) { Call* c = (Call*)p; hive_rollout_kernel(c->a, c->n_steps); }, &call, b, STEP_THREADS, sched_seed + (uint64_t)b);
        if (rc) return rc;
    }
    return 0;
}
// n_steps (<= 2 here: the emulator runs the CTAs one after the other) OP_RANDOM steps through the queue-driven rollout:
// the step CTAs take (group, step) tickets, then the store CTAs take them in the same order
int emu_env_rollout_q(void* recs, uint32_t* legal, int32_t* count, uint32_t* status, uint16_t* planes, int n, uint64_t seed,
                      int max_turn, int auto_reset, int n_steps, uint64_t sched_seed) {
    build_lines();
    if (n_steps < 1 || n_steps > 2) return -1;
    EnvArgs a;
    a.recs = (GameRec*)recs; a.legal = legal; a.count = count; a.status = status; a.planes = planes;
    a.actions = nullptr; a.mask = nullptr; a.chosen = nullptr; a.hop_lines = g_lines.data();
    a.seed = seed; a.n = n; a.op = OP_RANDOM; a.max_turn = max_turn; a.auto_reset = auto_reset;
    a.g_offset = 0; a.n_total = n; a.stagger_ns = 0; a.stagger_div = 148; a.shadow = nullptr; a.lists = nullptr;
    std::vector<uint32_t> bits0((size_t)n * BITS_WORDS), bits1((size_t)n * BITS_WORDS);
    a.bits = bits0.data();
    const int G = (n + SG - 1) / SG;
    std::vector<unsigned> sync(32 + 2 * G, 0u);
    struct Call { EnvArgs a; uint32_t* alt; int n_steps; RollSync* sync; } call = {a, bits1.data(), n_steps, (RollSync*)sync.data()};
    const int step_blocks = G > 2 ? 2 : G;
    emu::g_gridDim.x = step_blocks;
    for (int b = 0; b < step_blocks; b++) {
        int rc = emu::run_block([](void* p) { Call* c = (Call*)p; hive_rollout_q_kernel(c->a, c->alt, c->n_steps, c->sync); }, &call, b, STEP_THREADS, sched_seed + (uint64_t)b);
        if (rc) return rc;
    }
    for (int b = 0; b < 2; b++) {
        int rc = emu::run_block([](void* p) { Call* c = (Call*)p; hive_planes_q_kernel(c->a, c->alt, c->n_steps, c->sync); }, &call, b, HIVE_STORE_WARPS * 32, sched_seed + 5000 + (uint64_t)b);
        if (rc) return rc;
    }
    return 0;
}
const uint8_t* emu_env_lists() { return g_lists.data(); }
const char* emu_last_error() { return emu::last_error(); }
int emu_rec_bytes() { return (int)sizeof(GameRec); }
}
