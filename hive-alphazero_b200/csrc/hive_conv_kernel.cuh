// hive_conv_kernel.cuh -- 3x3 convolution (pad 1) over 12x12 boards as an implicit GEMM on the
// 5th-generation tensor cores (tcgen05 + TMEM), bf16 in / fp32 accumulate / bf16 out, with the
// folded-BatchNorm bias, the residual add and the ReLU fused into the epilogue.
// Reference op: alpha_zero/alpha_net.py:29-54 (ConvBlock / ResBlock), 98 % of the network's FLOPs.
//
// GEMM view per CTA:  D[128 out-channels][160 pixel slots] += W[128][64 in-ch] * X[160][64 in-ch]^T
//   * X is the zero-padded board placed ONCE in shared memory by a 5-D TMA box {8 ch, 13, 15, 8
//     ch-groups, 1 board} starting at pixel (-1,-1) whose out-of-range pixels are zero-filled: rows of
//     13 slots = one left pad + 12 pixels (the right pad of a row is the left pad of the next one).
//     The nine filter taps are nine VIEWS of that tile: the K-major no-swizzle operand layout
//     [ch-group][slot][8 ch] makes "shift by (dy,dx)" a start-address offset of (13*dy+dx)*16 bytes in
//     the shared-memory descriptor.  Output slot n = 13*y + x; slots with x == 12 (and n >= 156) are
//     padding columns of D that are never read back (144 of 160 columns are useful).
//   * W tiles are pre-packed on the host in the operand layout [k-group 8][128 rows][8 ch] (16 KB per
//     tap and 64-channel chunk) and streamed with 1-D bulk copies.
//   * CONV_BOARDS_PER_PASS boards share every weight tile (one accumulator each in TMEM); with one board
//     per pass two CTAs are resident per SM (256 TMEM columns each) so that the epilogue of one
//     overlaps the MMAs of the other.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 = epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "umma.cuh"

namespace hive {

constexpr int CONV_OC_TILE = 128;                 // UMMA M
#ifdef HIVE_CONV_N144
// Variant without padding columns: N = 144 = the 12 x 12 pixels, row pitch 12.  A horizontal shift can then no longer
// be an address offset (it would wrap into the neighbouring row), so the three column shifts dx = -1, 0, +1 are three
// separate zero-filled tiles (three TMA boxes starting at x = dx); the vertical shift stays an offset of 12 slots.
// Three times the operand bytes per board: 32-channel chunks keep two stages of two boards in shared memory.
// Measured SLOWER than the padded form (941 vs 1,245 TFLOP/s at 2,048 boards: the tripled operand writes compete with the
// MMA operand reads for shared-memory bandwidth); kept as a checked variant, not the default.
constexpr int CONV_N = 144;
constexpr int CONV_PADW = 12;
constexpr int CONV_PADH = 14;
constexpr int CONV_KG = 4;                        // 8-channel groups per chunk (32 input channels)
constexpr int CONV_SHIFTS = 3;
#else
constexpr int CONV_N = 160;                       // UMMA N (12 rows x 13 slots = 156, rounded to 16)
constexpr int CONV_PADW = 13;
constexpr int CONV_PADH = 15;
constexpr int CONV_KG = 8;                        // 8-channel groups per chunk (64 input channels)
constexpr int CONV_SHIFTS = 1;
#endif
constexpr int CONV_CHUNK_CH = CONV_KG * 8;        // input channels per chunk
constexpr int CONV_SLOT_COLS = 160;               // TMEM columns between accumulator slots
constexpr int CONV_PLANE_BYTES = CONV_PADH * CONV_PADW * 16;    // one 8-channel group of a padded board
constexpr int CONV_TILE_BYTES = CONV_KG * CONV_PLANE_BYTES;     // one column shift of one chunk of a board
constexpr int CONV_BOARD_BYTES = CONV_SHIFTS * CONV_TILE_BYTES; // 24,960 B (N = 160) / 32,256 B (N = 144)
constexpr int CONV_A_BYTES = CONV_KG * CONV_OC_TILE * 16;       // 16,384 B / 8,192 B
#ifndef CONV_BOARDS_PER_PASS
#define CONV_BOARDS_PER_PASS 2
#endif
constexpr int CONV_BOARDS = CONV_BOARDS_PER_PASS;  // boards per weight pass (accumulators in TMEM)
constexpr int CONV_A_STAGES = (CONV_BOARDS == 1 ? 3 : 4) * (8 / CONV_KG);
constexpr int CONV_B_STAGES = 2;
constexpr int CONV_CTAS_PER_SM = CONV_BOARDS == 1 ? 2 : 1;   // 1-board CTAs run two per SM: one's epilogue hides under the other's MMAs
constexpr int CONV_TMEM_COLS = CONV_BOARDS == 1 ? 256 : 512;
constexpr int CONV_SLOTS = CONV_BOARDS == 1 ? 1 : 3;          // accumulator slots of CONV_N columns, used round-robin
constexpr int CONV_EPI_WARPS = CONV_BOARDS == 1 ? 4 : 8;      // epilogue warps (multiple of 4: one per TMEM lane quarter)
constexpr int CONV_THREADS = 64 + 32 * CONV_EPI_WARPS;
constexpr int CONV_TAIL_PAD = 0;                  // last slot + largest shift stays inside the tile (187 < 195 / 167 < 168)
constexpr int CONV_STAGE_STRIDE = 36;             // floats per staging row (32 + 4: keeps 16-byte alignment, spreads banks)
constexpr int CONV_SMEM_BYTES = CONV_A_STAGES * CONV_A_BYTES + CONV_B_STAGES * CONV_BOARDS * CONV_BOARD_BYTES + CONV_TAIL_PAD +
                                CONV_EPI_WARPS * 16 * CONV_STAGE_STRIDE * 4 + 1024;
static_assert(CONV_SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct ConvArgs {
    const uint8_t* weights;        // [2 halves][9 taps][n_chunks][16 KB] packed operand tiles of this layer
    const float* bias;             // [256]
    const __nv_bfloat16* __restrict__ residual; // [B][144][256] or null (never aliases `out`)
    __nv_bfloat16* __restrict__ out;            // [B][144][256]
    int n_boards, n_chunks, relu;  // n_chunks = input channels / CONV_CHUNK_CH
};

__device__ __forceinline__ void bulk_load(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(umma::smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(CONV_THREADS, CONV_CTAS_PER_SM) hive_conv3x3_kernel(const __grid_constant__ CUtensorMap in_map, ConvArgs a) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + CONV_A_STAGES * CONV_A_BYTES;
    float* sStage = reinterpret_cast<float*>(sB + CONV_B_STAGES * CONV_BOARDS * CONV_BOARD_BYTES + CONV_TAIL_PAD);   // [4 warps][16][36]
    __shared__ uint64_t a_full[CONV_A_STAGES], a_empty[CONV_A_STAGES], b_full[CONV_B_STAGES], b_empty[CONV_B_STAGES];
    __shared__ uint64_t acc_full[CONV_SLOTS], acc_empty[CONV_SLOTS];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < CONV_A_STAGES; i++) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < CONV_B_STAGES; i++) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < CONV_SLOTS; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], CONV_EPI_WARPS); }
        mbar_fence_init();
        tma_prefetch_desc(&in_map);
    }
    if (warp == 1) tmem_alloc(&tmem_base, CONV_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;

    const int n_pairs = (a.n_boards + CONV_BOARDS - 1) / CONV_BOARDS;
    const int n_items = 2 * n_pairs;                         // (out-channel half, board pair)

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (elect_one()) {
            int as = 0, aph = 0, bs = 0, bph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int half = item & 1, pair = item >> 1;
                const uint8_t* wbase = a.weights + (size_t)half * 9 * a.n_chunks * CONV_A_BYTES;
                for (int c = 0; c < a.n_chunks; c++) {
                    mbar_wait(&b_empty[bs], bph ^ 1);
                    mbar_expect_tx(&b_full[bs], CONV_BOARDS * CONV_BOARD_BYTES);
                    for (int j = 0; j < CONV_BOARDS; j++) {
                        int b = pair * CONV_BOARDS + j;
                        if (b >= a.n_boards) b = a.n_boards - 1;          // odd tail: reload the last board (result discarded)
#pragma unroll
                        for (int sh = 0; sh < CONV_SHIFTS; sh++)      // one box per column shift (a single box from x = -1 when the pad column is kept)
                            tma_load_5d(sB + (bs * CONV_BOARDS + j) * CONV_BOARD_BYTES + sh * CONV_TILE_BYTES, &in_map, &b_full[bs], 0,
                                        CONV_SHIFTS == 1 ? -1 : sh - 1, -1, c * CONV_KG, b);
                    }
                    if (++bs == CONV_B_STAGES) { bs = 0; bph ^= 1; }
                    for (int t = 0; t < 9; t++) {
                        mbar_wait(&a_empty[as], aph ^ 1);
                        mbar_expect_tx(&a_full[as], CONV_A_BYTES);
                        bulk_load(sA + as * CONV_A_BYTES, wbase + (size_t)(t * a.n_chunks + c) * CONV_A_BYTES, CONV_A_BYTES, &a_full[as]);
                        if (++as == CONV_A_STAGES) { as = 0; aph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // (elect.sync instead of lane == 0: the compiler then knows that one thread runs this and keeps the descriptors in
        // uniform registers -- 6 instead of 21 instructions per MMA)
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16(CONV_OC_TILE, CONV_N);
            const uint64_t a_desc0 = smem_desc(smem_u32(sA), CONV_OC_TILE * 16, 128, 0);     // LBO = k-group stride, SBO = 8 rows
            const uint64_t b_desc0 = smem_desc(smem_u32(sB), CONV_PLANE_BYTES, 128, 0);
            int as = 0, aph = 0, bs = 0, bph = 0;
            uint32_t empty_ph = 0;                                     // bit s: parity of the next wait on acc_empty[s]
            int k = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, k++) {
                int slot[CONV_BOARDS];
#pragma unroll
                for (int j = 0; j < CONV_BOARDS; j++) {
                    slot[j] = (CONV_BOARDS * k + j) % CONV_SLOTS;
                    mbar_wait(&acc_empty[slot[j]], ((empty_ph >> slot[j]) & 1u) ^ 1u);
                    empty_ph ^= 1u << slot[j];
                }
                tc_fence_after();
                for (int c = 0; c < a.n_chunks; c++) {
                    mbar_wait(&b_full[bs], bph);
#pragma unroll
                    for (int t = 0; t < 9; t++) {
                        mbar_wait(&a_full[as], aph);
                        tc_fence_after();
                        // descriptors differ only in their 14-bit start-address field: add to the low word
                        const uint32_t kShift = CONV_SHIFTS == 1 ? (uint32_t)((t / 3) * CONV_PADW + (t % 3))   // 16-byte units (constant after unrolling)
                                                                 : (uint32_t)((t / 3) * CONV_PADW + (t % 3) * (CONV_TILE_BYTES >> 4));
                        const uint64_t a_lo = a_desc0 + (uint64_t)((uint32_t)(as * CONV_A_BYTES) >> 4);
#pragma unroll
                        for (int j = 0; j < CONV_BOARDS; j++) {
                            const uint64_t b_lo = b_desc0 + (uint64_t)(((uint32_t)((bs * CONV_BOARDS + j) * CONV_BOARD_BYTES) >> 4) + kShift);
#pragma unroll
                            for (int ks = 0; ks < CONV_KG / 2; ks++)
                                mma_bf16(tmem + slot[j] * CONV_SLOT_COLS, a_lo + (uint64_t)(ks * ((2 * CONV_OC_TILE * 16) >> 4)),
                                         b_lo + (uint64_t)(ks * ((2 * CONV_PLANE_BYTES) >> 4)), idesc, (c | t | ks) != 0);
                        }
                        mma_commit(&a_empty[as]);                      // weight stage free when these MMAs retire
                        if (++as == CONV_A_STAGES) { as = 0; aph ^= 1; }
                    }
                    mma_commit(&b_empty[bs]);
                    if (++bs == CONV_B_STAGES) { bs = 0; bph ^= 1; }
                }
#pragma unroll
                for (int j = 0; j < CONV_BOARDS; j++) mma_commit(&acc_full[slot[j]]);
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        // TMEM -> registers (thread = out-channel row, 16 pixel slots at a time) -> +bias -> shared staging
        // [slot][32 ch] -> 16-byte vectors (thread = 8 channels of one slot) -> +residual, ReLU -> bf16 NHWC.
        const int q = warp & 3;                                    // TMEM lane quarter this warp may read
        const int ew = warp - 2;                                   // epilogue warp index 0..CONV_EPI_WARPS-1
        constexpr int SETS = CONV_EPI_WARPS / 4;                   // warp sets sharing a board: each takes G/SETS column groups
        const int set = ew >> 2;
        float* stage = sStage + ew * 16 * CONV_STAGE_STRIDE;
        const int sl = lane >> 2, ch8 = (lane & 3) * 8;            // phase-2 role: slots sl and sl+8, channels ch8..ch8+7
        constexpr int NG = CONV_N / 16;                            // column groups of an accumulator (10 / 9)
        constexpr int G = (NG + SETS - 1) / SETS;                  // column groups per warp (the last set may own a dummy one)
        constexpr int CH = 5;                                      // column groups held in registers at a time
        static_assert(G % CH == 0, "column groups per warp must be a multiple of the register chunk");
        uint32_t full_ph = 0;
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
            const int half = item & 1, pair = item >> 1;
            const int oc0 = half * CONV_OC_TILE + q * 32;
            const float bias = a.bias[oc0 + lane];
            for (int j = 0; j < CONV_BOARDS; j++) {
                const int slot = (CONV_BOARDS * it + j) % CONV_SLOTS;
                const int b = pair * CONV_BOARDS + j;
                const bool valid = b < a.n_boards;
                const int g0 = set * G;                                // first column group of this warp
                const size_t bbase = (size_t)(valid ? b : 0) * 144 * 256 + oc0 + ch8;
                // element offset of (group g, k-th slot of this lane) or -1 for a padding slot
                auto slot_off = [&](int g, int k) -> long long {      // g counts from this warp's first group
                    const int n = (g0 + g) * 16 + sl + 8 * k, y = n / CONV_PADW, x = n - y * CONV_PADW;
                    return (valid && g0 + g < NG && x < 12 && y < 12) ? (long long)(bbase + (size_t)(y * 12 + x) * 256) : -1;
                };
#pragma unroll
                for (int c0 = 0; c0 < G; c0 += CH) {
                    // residual rows of this chunk are requested BEFORE the accumulator is awaited, so their
                    // latency hides under the MMAs that are still running
                    uint4 rq[CH][2];
                    if (a.residual) {
#pragma unroll
                        for (int g = 0; g < CH; g++)
#pragma unroll
                            for (int k = 0; k < 2; k++) {
                                const long long o = slot_off(c0 + g, k);
                                rq[g][k] = o >= 0 ? *reinterpret_cast<const uint4*>(a.residual + o) : make_uint4(0u, 0u, 0u, 0u);
                            }
                    }
                    if (c0 == 0) {
                        mbar_wait(&acc_full[slot], (full_ph >> slot) & 1u);
                        full_ph ^= 1u << slot;
                        tc_fence_after();
                    }
                    // all TMEM reads of the chunk in flight at once, one wait
                    uint32_t v[CH][16];
#pragma unroll
                    for (int g = 0; g < CH; g++)
                        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + slot * CONV_SLOT_COLS + (g0 + c0 + g) * 16, v[g]);
                    tmem_ld_wait();
                    if (c0 + CH >= G) {
                        // the accumulator now lives in registers: hand the TMEM slot back to the MMA issuer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[slot]);
                    }
                    if (!valid) continue;
#pragma unroll
                    for (int g = 0; g < CH; g++) {
#pragma unroll
                        for (int i = 0; i < 16; i++) stage[i * CONV_STAGE_STRIDE + lane] = __uint_as_float(v[g][i]) + bias;
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const long long o = slot_off(c0 + g, k);
                            const float4 f0 = *reinterpret_cast<const float4*>(stage + (sl + 8 * k) * CONV_STAGE_STRIDE + ch8);
                            const float4 f1 = *reinterpret_cast<const float4*>(stage + (sl + 8 * k) * CONV_STAGE_STRIDE + ch8 + 4);
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, CONV_TMEM_COLS);
}

}  // namespace hive
