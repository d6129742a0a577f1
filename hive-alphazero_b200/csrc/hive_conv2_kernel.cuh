// hive_conv2_kernel.cuh -- the trunk's 3x3 convolution on CTA PAIRS (tcgen05.mma.cta_group::2, M = 256).
// Same op, operand layouts, weight packing and TMA box as hive_conv_kernel.cuh (alpha_zero/alpha_net.py:29-54); what
// changes is who reads what from shared memory.
//
// hive_conv3x3_kernel: a CTA multiplies ONE out-channel half (128 rows of W) with TWO boards; every MMA reads a 4 KB
// weight slice and a 5 KB board slice from the CTA's shared memory -- 18 KB per 2 x (128 x 160 x 16) MACs, and the
// operand reads (~115 of the 128 B/clk an SM's shared memory delivers) are what bounds it (r02_net_ncu_full_summary.csv:
// tensor pipe 69.5 % active).
//
// Here two CTAs on the SMs of one TPC form a cluster and work on one board pair together:
//   * CTA r holds out-channel half r of the weights (its 128 rows of A) and board r of the pair (its tile of B);
//   * one MMA of the pair is D[256][160] += W[256][16] * X[160][16]^T where the 160 columns of X are 80 pixel slots of
//     board 0 (read from CTA 0) followed by the SAME 80 slots of board 1 (read from CTA 1): the K-major no-swizzle
//     board tile makes "slots s0 .. s0+79 shifted by tap (dy,dx)" a start-address offset, identical in both CTAs.
//     Two such MMAs (s0 = 0, 80) per tap and 16-channel step cover both boards; accumulator j of the pair holds
//     [board 0 slots 80j..80j+79 | board 1 slots 80j..80j+79] for the CTA's own 128 out-channels;
//   * per CTA and pair of MMAs: 2 x 4 KB of weights + 2 x 2.5 KB of board = 13 KB for 2 x (128 x 160 x 16) MACs of its
//     own tensor core: 28 % fewer operand bytes per MAC.
// Only the leader CTA (rank 0) issues MMAs; both CTAs run a TMA producer (own weights half, own board), CTA 1 forwards
// "stage full" to the leader with remote mbarrier arrives, the leader's commits release the stages and publish the
// accumulators in both CTAs (multicast commit), and both CTAs run the epilogue on their own TMEM (set s of the
// epilogue warps = board s of the pair).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "hive_conv_kernel.cuh"

namespace hive {

#ifndef HIVE_CONV2_A_STAGES
#define HIVE_CONV2_A_STAGES 10          // one chunk's nine taps: the weight stream's latency (L2 -> shared memory, then a hop to the leader) is what starves the MMAs (5 stages: 1,221 TFLOP/s, 9: 1,353)
#endif
#ifndef HIVE_CONV2_B_STAGES
#define HIVE_CONV2_B_STAGES 2
#endif
constexpr int C2_A_STAGES = HIVE_CONV2_A_STAGES, C2_B_STAGES = HIVE_CONV2_B_STAGES;
constexpr int C2_SLOTS = 3;                        // accumulator slots of 160 columns, used round-robin (two per board pair)
constexpr int C2_TMEM_COLS = 512;
constexpr int C2_EPI_WARPS = 8;                    // two sets of four (one per TMEM lane quarter): set = board of the pair
constexpr int C2_THREADS = 64 + 32 * C2_EPI_WARPS;
constexpr int C2_STAGE_ROWS = 8;                   // rows of a warp's transposition buffer (half a 16-slot column group at a time: the
                                                   // shared memory saved is the tenth weight stage)
constexpr int C2_HALF_N = CONV_N / 2;              // pixel slots per board and MMA
constexpr int C2_SMEM_BYTES = C2_A_STAGES * CONV_A_BYTES + C2_B_STAGES * CONV_BOARD_BYTES + C2_EPI_WARPS * C2_STAGE_ROWS * CONV_STAGE_STRIDE * 4 + 1024;
static_assert(C2_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(CONV_SHIFTS == 1 && CONV_N == 160, "the pair kernel is written for the padded N = 160 tile");

namespace umma2 {
using umma::smem_u32;
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// true in exactly one lane of a converged warp (the compiler then knows that a single thread runs the guarded code and
// keeps its operands in uniform registers instead of broadcasting every descriptor before every MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp of EACH CTA of the pair (same warp id, same smem offset)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// USE: 0 = plain; 1 = keep this A slice in the collector for the next MMA (fill); 2 = the A slice is the one kept (lastuse)
template <int USE>
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (USE == 1)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else if (USE == 2)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs when all previously issued MMAs have completed
__device__ __forceinline__ void mma_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// arrive on the mbarrier at this shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    // (default semantics, release at CTA scope, as CUTLASS's ClusterBarrier::arrive(cta_id): a cluster-scope release would
    // make an epilogue lane wait for its outstanding global stores before the TMEM slot is handed back)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {     // the arrivals come from the other CTA
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
}  // namespace umma2

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C2_THREADS, 1) hive_conv3x3_pair_kernel(const __grid_constant__ CUtensorMap in_map, ConvArgs a) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + C2_A_STAGES * CONV_A_BYTES;
    float* sStage = reinterpret_cast<float*>(sB + C2_B_STAGES * CONV_BOARD_BYTES);
    __shared__ uint64_t a_full[C2_A_STAGES], a_empty[C2_A_STAGES], b_full[C2_B_STAGES], b_empty[C2_B_STAGES];
    __shared__ uint64_t peer_a_full[C2_A_STAGES], peer_b_full[C2_B_STAGES];      // used in the leader: CTA 1's stage is full
    __shared__ uint64_t acc_full[C2_SLOTS], acc_empty[C2_SLOTS];                 // acc_empty is used in the leader (both CTAs' epilogues arrive)
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = umma2::cluster_rank();

    if (tid == 0) {
        for (int i = 0; i < C2_A_STAGES; i++) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&peer_a_full[i], 1); }
        for (int i = 0; i < C2_B_STAGES; i++) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); mbar_init(&peer_b_full[i], 1); }
        for (int i = 0; i < C2_SLOTS; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * C2_EPI_WARPS); }
        mbar_fence_init();
        tma_prefetch_desc(&in_map);
    }
    if (warp == 1) umma2::tmem_alloc(&tmem_base, C2_TMEM_COLS);
    tc_fence_before();
    umma2::cluster_sync();                                   // both CTAs' barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem = tmem_base;

    const int n_pairs = (a.n_boards + 1) / 2;
    const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ------------------------------------------------------------ producer (both CTAs: own weights half, own board)
        if (lane == 0) {
            int as = 0, aph = 0, bs = 0, bph = 0;
            const uint8_t* wbase = a.weights + (size_t)rank * 9 * a.n_chunks * CONV_A_BYTES;
            for (int item = cluster; item < n_pairs; item += n_clusters) {
                int b = item * 2 + (int)rank;
                if (b >= a.n_boards) b = a.n_boards - 1;                  // odd tail: reload the last board (result discarded)
                for (int c = 0; c < a.n_chunks; c++) {
                    mbar_wait(&b_empty[bs], bph ^ 1);
                    mbar_expect_tx(&b_full[bs], CONV_BOARD_BYTES);
                    tma_load_5d(sB + bs * CONV_BOARD_BYTES, &in_map, &b_full[bs], 0, -1, -1, c * CONV_KG, b);
                    if (++bs == C2_B_STAGES) { bs = 0; bph ^= 1; }
                    for (int t = 0; t < 9; t++) {
                        mbar_wait(&a_empty[as], aph ^ 1);
                        mbar_expect_tx(&a_full[as], CONV_A_BYTES);
                        bulk_load(sA + as * CONV_A_BYTES, wbase + (size_t)(t * a.n_chunks + c) * CONV_A_BYTES, CONV_A_BYTES, &a_full[as]);
                        if (++as == C2_A_STAGES) { as = 0; aph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 1) {
        // ------------------------------------------------------------ CTA 1: tell the leader when a stage of this CTA is full
        if (lane == 0) {
            int as = 0, aph = 0, bs = 0, bph = 0;
            for (int item = cluster; item < n_pairs; item += n_clusters)
                for (int c = 0; c < a.n_chunks; c++) {
                    mbar_wait(&b_full[bs], bph);
                    umma2::mbar_arrive_cta(&peer_b_full[bs], 0);
                    if (++bs == C2_B_STAGES) { bs = 0; bph ^= 1; }
                    for (int t = 0; t < 9; t++) {
                        mbar_wait(&a_full[as], aph);
                        umma2::mbar_arrive_cta(&peer_a_full[as], 0);
                        if (++as == C2_A_STAGES) { as = 0; aph ^= 1; }
                    }
                }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ leader: MMA issuer for the pair
        if (umma2::elect_one()) {
            const uint32_t idesc = idesc_bf16(2 * CONV_OC_TILE, CONV_N);
            const uint64_t a_desc0 = smem_desc(smem_u32(sA), CONV_OC_TILE * 16, 128, 0);     // LBO = k-group stride, SBO = 8 rows
            const uint64_t b_desc0 = smem_desc(smem_u32(sB), CONV_PLANE_BYTES, 128, 0);
            int as = 0, aph = 0, bs = 0, bph = 0;
            uint32_t empty_ph = 0;                                     // bit s: parity of the next wait on acc_empty[s]
            int k = 0;
            for (int item = cluster; item < n_pairs; item += n_clusters, k++) {
                int slot[2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    slot[j] = (2 * k + j) % C2_SLOTS;
                    umma2::mbar_wait_cluster(&acc_empty[slot[j]], ((empty_ph >> slot[j]) & 1u) ^ 1u);
                    empty_ph ^= 1u << slot[j];
                }
                tc_fence_after();
                for (int c = 0; c < a.n_chunks; c++) {
                    mbar_wait(&b_full[bs], bph);
                    umma2::mbar_wait_cluster(&peer_b_full[bs], bph);
#pragma unroll
                    for (int t = 0; t < 9; t++) {
                        mbar_wait(&a_full[as], aph);
                        umma2::mbar_wait_cluster(&peer_a_full[as], aph);
                        tc_fence_after();
                        const uint32_t kShift = (uint32_t)((t / 3) * CONV_PADW + (t % 3));      // 16-byte units (constant after unrolling)
                        const uint64_t a_lo = a_desc0 + (uint64_t)((uint32_t)(as * CONV_A_BYTES) >> 4);
                        const uint64_t b_lo0 = b_desc0 + (uint64_t)(((uint32_t)(bs * CONV_BOARD_BYTES) >> 4) + kShift);
#ifdef HIVE_CONV2_COLLECT
                        // the two MMAs of a 16-channel step share their weight slice: the second takes it from the collector
#pragma unroll
                        for (int ks = 0; ks < CONV_KG / 2; ks++) {
                            const uint64_t a_k = a_lo + (uint64_t)(ks * ((2 * CONV_OC_TILE * 16) >> 4)), b_k = b_lo0 + (uint64_t)(ks * ((2 * CONV_PLANE_BYTES) >> 4));
                            umma2::mma_bf16<1>(tmem + slot[0] * CONV_SLOT_COLS, a_k, b_k, idesc, (c | t | ks) != 0);
                            umma2::mma_bf16<2>(tmem + slot[1] * CONV_SLOT_COLS, a_k, b_k + (uint64_t)C2_HALF_N, idesc, (c | t | ks) != 0);
                        }
#else
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const uint64_t b_lo = b_lo0 + (uint64_t)(j * C2_HALF_N);
#pragma unroll
                            for (int ks = 0; ks < CONV_KG / 2; ks++)
                                umma2::mma_bf16<0>(tmem + slot[j] * CONV_SLOT_COLS, a_lo + (uint64_t)(ks * ((2 * CONV_OC_TILE * 16) >> 4)),
                                                   b_lo + (uint64_t)(ks * ((2 * CONV_PLANE_BYTES) >> 4)), idesc, (c | t | ks) != 0);
                        }
#endif
                        umma2::mma_commit_both(&a_empty[as]);          // both CTAs' weight stages are free when these MMAs retire
                        if (++as == C2_A_STAGES) { as = 0; aph ^= 1; }
                    }
                    umma2::mma_commit_both(&b_empty[bs]);
                    if (++bs == C2_B_STAGES) { bs = 0; bph ^= 1; }
                }
#pragma unroll
                for (int j = 0; j < 2; j++) umma2::mma_commit_both(&acc_full[slot[j]]);
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
        // accumulator j: columns [0,80) = board 0 of the pair, slots 80j..80j+79; columns [80,160) = board 1, same slots.
        // Warp set s (4 warps = the 4 TMEM lane quarters) takes the 5 column groups of board s.
        const int q = warp & 3;                                    // TMEM lane quarter this warp may read
        const int ew = warp - 2, set = ew >> 2;
        float* stage = sStage + ew * C2_STAGE_ROWS * CONV_STAGE_STRIDE;
        const int sl = lane >> 2, ch8 = (lane & 3) * 8;            // phase-2 role: slots sl and sl+8, channels ch8..ch8+7
        constexpr int G = 5;                                       // column groups (16 slots) per board and accumulator
        const int oc0 = (int)rank * CONV_OC_TILE + q * 32;
        const float bias = a.bias[oc0 + lane];
        uint32_t full_ph = 0;
        int it = 0;
        for (int item = cluster; item < n_pairs; item += n_clusters, it++) {
            const int b = item * 2 + set;
            const bool valid = b < a.n_boards;
            const size_t bbase = (size_t)(valid ? b : 0) * 144 * 256 + oc0 + ch8;
            for (int j = 0; j < 2; j++) {
                const int slot = (2 * it + j) % C2_SLOTS;
                auto slot_off = [&](int g, int k) -> long long {      // element offset of (group g, k-th slot of this lane) or -1 for a padding slot
                    const int n = j * C2_HALF_N + g * 16 + sl + 8 * k, y = n / CONV_PADW, x = n - y * CONV_PADW;
                    return (valid && x < 12 && y < 12) ? (long long)(bbase + (size_t)(y * 12 + x) * 256) : -1;
                };
                // residual rows are requested BEFORE the accumulator is awaited, so their latency hides under the MMAs
                uint4 rq[G][2];
                if (a.residual) {
#pragma unroll
                    for (int g = 0; g < G; g++)
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const long long o = slot_off(g, k);
                            rq[g][k] = o >= 0 ? *reinterpret_cast<const uint4*>(a.residual + o) : make_uint4(0u, 0u, 0u, 0u);
                        }
                }
                mbar_wait(&acc_full[slot], (full_ph >> slot) & 1u);
                full_ph ^= 1u << slot;
                tc_fence_after();
                uint32_t v[G][16];
#pragma unroll
                for (int g = 0; g < G; g++)
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + slot * CONV_SLOT_COLS + (set * G + g) * 16, v[g]);
                tmem_ld_wait();
                // the accumulator now lives in registers: hand the TMEM slot back to the leader's MMA issuer
                tc_fence_before();
                __syncwarp();
                if (lane == 0) umma2::mbar_arrive_cta(&acc_empty[slot], 0);
                if (!valid) continue;
#pragma unroll
                for (int g = 0; g < G; g++) {
#pragma unroll
                    for (int k = 0; k < 2; k++) {                       // slots 8k .. 8k+7 of the group
                        if (k) __syncwarp();                            // the first half has been read
#pragma unroll
                        for (int i = 0; i < 8; i++) stage[i * CONV_STAGE_STRIDE + lane] = __uint_as_float(v[g][8 * k + i]) + bias;
                        __syncwarp();
                        const long long o = slot_off(g, k);
                        const float4 f0 = *reinterpret_cast<const float4*>(stage + sl * CONV_STAGE_STRIDE + ch8);
                        const float4 f1 = *reinterpret_cast<const float4*>(stage + sl * CONV_STAGE_STRIDE + ch8 + 4);
                        float r[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                        if (a.residual) {
                            const uint4 rv = rq[g][k];
                            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                r[2 * e] += __uint_as_float(rw[e] << 16);
                                r[2 * e + 1] += __uint_as_float(rw[e] & 0xFFFF0000u);
                            }
                        }
                        if (a.relu) {
#pragma unroll
                            for (int e = 0; e < 8; e++) r[e] = fmaxf(r[e], 0.f);
                        }
                        if (o >= 0) {
                            uint4 pk;
                            __nv_bfloat162 h;
                            h = __floats2bfloat162_rn(r[0], r[1]); pk.x = *reinterpret_cast<uint32_t*>(&h);
                            h = __floats2bfloat162_rn(r[2], r[3]); pk.y = *reinterpret_cast<uint32_t*>(&h);
                            h = __floats2bfloat162_rn(r[4], r[5]); pk.z = *reinterpret_cast<uint32_t*>(&h);
                            h = __floats2bfloat162_rn(r[6], r[7]); pk.w = *reinterpret_cast<uint32_t*>(&h);
                            *reinterpret_cast<uint4*>(a.out + o) = pk;
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    tc_fence_before();
    umma2::cluster_sync();                                   // the leader's MMAs read CTA 1's shared memory; both TMEMs are drained
    if (warp == 1) umma2::tmem_dealloc(tmem, C2_TMEM_COLS);
}

}  // namespace hive
