// hive_host_loop.cu -- the host-driven game loop inside the library (C ABI in include/hive_b200.h).
//
// The reference's self-play worker is a host loop: read the legal actions, let a policy pick, call GamePlay.move,
// repeat (woker/self_play.py:54-56,116-193).  Driving the batched environment that way from Python costs three
// interpreter round trips per step and part (round-1 e2e probe: wait 9.5 us + policy 18.5 us + launch 14.9 us, all on
// one thread).  Here the batch is cut into parts, every part is a hive_env of its own (own streams, own page-locked
// staging buffers, the whole upload -> step -> download sequence one CUDA-graph launch, see hive_step_host_async), and
// a fixed set of native driver threads walks over the parts: wait for a part's downloads, run the policy on its
// legal masks / counts / status, launch its next step -- while the GPU steps the other parts.  No interpreter, no
// lock and no allocation in the loop; a thread's parts are its own.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hive_b200.h"
#include "hive_internal.h"

struct hive_host_loop {
    int device = 0, n_games = 0, n_parts = 0, n_threads = 0;
    bool use_lists = true;                   // HIVE_B200_HOST_LISTS=0: download the 198-byte masks every step (the round-1/2 form)
    struct Part {
        hive_env_t* env = nullptr;
        int n = 0, first = 0;                // games [first, first + n) of the batch
        uint64_t* mask = nullptr;            // page-locked staging buffers
        int32_t* count = nullptr;
        uint32_t* status = nullptr;
        int32_t* actions = nullptr;
        uint8_t* lists = nullptr;            // page-locked: compact legal lists of the last step (hive_step_host_async_lists)
        bool have_lists = false;             // false until the first step of this part has run (the first pick reads the masks)
        std::vector<uint32_t> scratch_ep;    // overflow groups: episode counters for the mask-based re-pick
        std::vector<int32_t> scratch_act;
        std::vector<uint32_t> episodes;
    };
    std::vector<Part> parts;
};

extern "C" {

int hive_host_loop_destroy(hive_host_loop_t* l) {
    if (!l) return 0;
    cudaSetDevice(l->device);
    for (auto& p : l->parts) {
        if (p.env) hive_destroy(p.env);
        if (p.mask) cudaFreeHost(p.mask);
        if (p.actions) cudaFreeHost(p.actions);
        if (p.lists) cudaFreeHost(p.lists);
    }
    delete l;
    return 0;
}

int hive_host_loop_create(int n_games, int device, int n_parts, int n_threads, hive_host_loop_t** out) {
    if (!out || n_games <= 0 || n_parts < 1 || n_threads < 1 || n_parts > n_games)
        return hive::fail(HIVE_E_ARG, "hive_host_loop_create: bad arguments");
    *out = nullptr;
    hive_host_loop* l = new hive_host_loop();
    l->device = device; l->n_games = n_games; l->n_parts = n_parts; l->n_threads = n_threads < n_parts ? n_threads : n_parts;
    { const char* e = getenv("HIVE_B200_HOST_LISTS"); l->use_lists = !(e && atoi(e) == 0); }
    l->parts.resize(n_parts);
    int first = 0;
    for (int i = 0; i < n_parts; i++) {
        hive_host_loop::Part& p = l->parts[i];
        p.n = n_games / n_parts + (i < n_games % n_parts ? 1 : 0);
        p.first = first; first += p.n;
        int rc = hive_create(p.n, device, nullptr, &p.env);
        if (rc) { hive_host_loop_destroy(l); return rc; }
        // masks, counts and status in one page-locked arena laid out like the device's: one download per step
        cudaError_t e = cudaHostAlloc(&p.mask, (size_t)p.n * (HIVE_LEGAL_U64 * 8 + 8), cudaHostAllocDefault);
        if (e == cudaSuccess) {
            p.count = reinterpret_cast<int32_t*>(p.mask + (size_t)p.n * HIVE_LEGAL_U64);
            p.status = reinterpret_cast<uint32_t*>(p.count + p.n);
            e = cudaHostAlloc(&p.actions, (size_t)p.n * 4, cudaHostAllocDefault);
        }
        if (e == cudaSuccess) e = cudaHostAlloc(&p.lists, (size_t)((p.n + 31) / 32) * HIVE_LIST_BLOCK_BYTES, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            hive_host_loop_destroy(l);
            return hive::fail(HIVE_E_CUDA, std::string("hive_host_loop_create: cudaHostAlloc: ") + cudaGetErrorString(e));
        }
        p.episodes.assign(p.n, 0);
        rc = hive_legal_host(p.env, p.mask, p.count);
        if (!rc) rc = hive_status_packed_host(p.env, p.status);
        if (rc) { hive_host_loop_destroy(l); return rc; }
    }
    *out = l;
    return 0;
}

int hive_host_loop_parts(const hive_host_loop_t* l) { return l ? l->n_parts : HIVE_E_HANDLE; }
int hive_host_loop_threads(const hive_host_loop_t* l) { return l ? l->n_threads : HIVE_E_HANDLE; }
hive_env_t* hive_host_loop_part(hive_host_loop_t* l, int part, int* first_game) {
    if (!l || part < 0 || part >= l->n_parts) return nullptr;
    if (first_game) *first_game = l->parts[part].first;
    return l->parts[part].env;
}

int hive_host_loop_run(hive_host_loop_t* l, int n_steps, uint64_t seed, int max_turn, hive_policy_fn policy, void* user,
                       double* seconds, double* policy_seconds, double* wait_seconds) {
    if (!l) return hive::fail(HIVE_E_HANDLE, "bad handle");
    if (n_steps < 1 || max_turn < 1 || max_turn > 250) return hive::fail(HIVE_E_ARG, "hive_host_loop_run: bad arguments");
    const int T = l->n_threads;
    std::vector<int> rcs(T, 0);
    std::vector<std::string> errs(T);
    std::vector<double> t_policy(T, 0.0), t_wait(T, 0.0);
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    auto body = [&](int t) {
        cudaSetDevice(l->device);
        ready.fetch_add(1);
        while (!go.load(std::memory_order_acquire)) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        using clk = std::chrono::steady_clock;
        double tp = 0.0, tw = 0.0;
        for (int step = 0; step < n_steps && !rcs[t]; step++)
            for (int i = t; i < l->n_parts; i += T) {
                hive_host_loop::Part& p = l->parts[i];
                const auto a = clk::now();
                int rc = hive_wait_results(p.env);          // this part's last downloads have landed (its planes may still be in flight)
                const auto b = clk::now();
                const bool lists = l->use_lists && !policy;        // (a caller's policy is handed the masks: its signature says so)
                const uint64_t pseed = seed + 77ull * (uint64_t)(i + 1);
                if (!rc) {
                    if (policy) policy(user, i, p.first, p.n, p.mask, p.count, p.status, p.actions);
                    else if (lists && p.have_lists) {
                        // the k-th legal action from the compact lists (96 B per game came down instead of 208)
                        int overflow = 0;
                        p.scratch_ep = p.episodes;                 // (the counters before this pick, for a possible re-pick)
                        rc = hive_host_pick_actions_lists(p.n, p.lists, p.status, p.episodes.data(), pseed, max_turn, p.actions, &overflow);
                        if (!rc && overflow) {
                            // a group's lists did not fit its block (very rare: > 84 legal actions per game on average in a
                            // group): fetch the masks and pick those games from them
                            rc = hive_legal_host(p.env, p.mask, p.count);
                            p.scratch_act.resize(p.n);
                            if (!rc) rc = hive_host_pick_actions(p.n, p.mask, p.count, p.status, p.scratch_ep.data(), pseed, max_turn, p.scratch_act.data());
                            for (int g = 0; g < p.n && !rc; g++)
                                if (p.lists[(size_t)(g / 32) * HIVE_LIST_BLOCK_BYTES + (g % 32) * 12 + 9] & 1u) p.actions[g] = p.scratch_act[g];
                        }
                    } else rc = hive_host_pick_actions(p.n, p.mask, p.count, p.status, p.episodes.data(), pseed, max_turn, p.actions);   // small parts on this thread, large ones over the policy pool
                }
                const auto c = clk::now();
                // H2D 4 B/game -> kernels -> D2H, all queued; the thread's other parts are handled meanwhile
                if (!rc) {
                    if (lists) { rc = hive_step_host_async_lists(p.env, p.actions, p.lists, p.status); p.have_lists = true; }
                    else { rc = hive_step_host_async(p.env, p.actions, p.mask, p.count, p.status); p.have_lists = false; }
                }
                tw += std::chrono::duration<double>(b - a).count();
                tp += std::chrono::duration<double>(c - b).count();
                if (rc) { rcs[t] = rc; errs[t] = hive_last_error(); break; }
            }
        for (int i = t; i < l->n_parts; i += T) {            // the last step's results and planes have landed
            int rc = hive_sync(l->parts[i].env);
            if (rc && !rcs[t]) { rcs[t] = rc; errs[t] = hive_last_error(); }
        }
        t_policy[t] = tp; t_wait[t] = tw;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(body, t);
    while (ready.load() < T - 1) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    body(0);                                                 // the caller's thread drives a share too
    for (auto& x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (seconds) *seconds = dt;
    double sp = 0.0, sw = 0.0;
    for (int t = 0; t < T; t++) { sp += t_policy[t]; sw += t_wait[t]; }
    if (policy_seconds) *policy_seconds = sp / T;            // mean per driver thread
    if (wait_seconds) *wait_seconds = sw / T;
    for (int t = 0; t < T; t++) if (rcs[t]) return hive::fail(rcs[t], "hive_host_loop_run: " + errs[t]);
    return 0;
}

long long hive_host_loop_env_steps(hive_host_loop_t* l) {
    if (!l) return -1;
    long long total = 0;
    std::vector<uint32_t> steps;
    for (auto& p : l->parts) {
        steps.resize(p.n);
        if (hive_counters_host(p.env, steps.data(), nullptr)) return -1;
        for (int i = 0; i < p.n; i++) total += steps[i];
    }
    return total;
}

}  // extern "C"
