// hive_heads_kernel.cuh -- the policy / value heads of the network on the 5th-generation tensor cores (tcgen05 + TMEM),
// reference alpha_zero/alpha_net.py:56-80 (OutBlock): 1x1 conv 256->128 + BN + ReLU, flatten, Linear 18432->1584,
// softmax; 1x1 conv 256->1 + BN + ReLU, Linear 144->64 + ReLU, Linear 64->1, tanh.  Three kernels per wave, fed by the
// trunk's NHWC activations and writing straight into the search's leaf arenas (policy float32, value float64):
//
//   hive_head_conv_kernel    both 1x1 convolutions as ONE GEMM  D[128 (board,cell) rows][144 cols] = X[rows][256] * W1[129][256]^T
//                            (cols 0..127 policy channels, col 128 the value channel, 129..143 padding), bias + ReLU in
//                            the epilogue; the policy half is written as bf16 directly in the operand layout the next
//                            GEMM streams, the value half as float [board][cell].
//   hive_head_fc_kernel      D[128 boards][176 cols] = P[boards][18432] * Wfc[1584][18432]^T (nine 176-column tiles), + bias
//                            -> float logits.
//   hive_head_finish_kernel  one warp per board: softmax over its 1,584 logits -> policy arena; 144->64->1 MLP + tanh
//                            -> value arena.
//
// Operand tiles are K-major without swizzle: [k-group of 8 elements][rows][8 elements] (16-byte core rows), so a
// descriptor's leading-dimension byte offset is rows*16 and its stride-dimension offset 128 (eight rows).  Weights are
// packed on the host per 64-wide K chunk, every tile one contiguous bulk copy; X is gathered into that layout by a
// 3-D TMA box {8 elements, 128 rows, 8 k-groups}.
// Warp roles in both GEMM kernels: warp 0 = producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 = epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "hive_conv_kernel.cuh"      // bulk_load
#include "umma.cuh"

namespace hive {

constexpr int HEAD_M = 128;                          // rows per tile (UMMA M, TMEM lanes)
constexpr int HEAD_KC = 64;                          // K elements per chunk (8 k-groups)
constexpr int HEAD_A_BYTES = HEAD_M * HEAD_KC * 2;   // 16,384 B: one A tile chunk
constexpr int HEAD_THREADS = 192;

// ---- 1x1 convolutions
constexpr int HC_N = 144;                            // 128 policy channels + value channel + padding (multiple of 16)
constexpr int HC_CHUNKS = 256 / HEAD_KC;             // 4
constexpr int HC_W_BYTES = HC_CHUNKS * HC_N * HEAD_KC * 2;      // 73,728 B, resident in shared memory
constexpr int HC_STAGES = 8;
constexpr int HC_SMEM_BYTES = HC_W_BYTES + HC_STAGES * HEAD_A_BYTES + 1024;
constexpr int HC_TMEM_COLS = 512;                    // two accumulator slots at columns 0 and 256
// ---- policy fc
constexpr int FC_K = 144 * 128;                      // 18,432
constexpr int FC_CHUNKS = FC_K / HEAD_KC;            // 288
constexpr int FC_N = 176;                            // columns per tile: 1584 = 9 x 176
constexpr int FC_NT = 1584 / FC_N;                   // 9
constexpr int FC_B_BYTES = FC_N * HEAD_KC * 2;       // 22,528 B
constexpr int FC_STAGES = 5;
constexpr int FC_SMEM_BYTES = FC_STAGES * (HEAD_A_BYTES + FC_B_BYTES) + 1024;
constexpr int FC_TMEM_COLS = 256;

// element offset of (board, k) in the policy-fc A operand: [board tile of 128][chunk of 64][k-group][128 boards][8]
__host__ __device__ __forceinline__ size_t fc_a_offset(int board, int k) {
    const int mt = board >> 7, r = board & 127, kc = k >> 6, kg = (k >> 3) & 7, e = k & 7;
    return ((((size_t)mt * FC_CHUNKS + kc) * 8 + kg) * HEAD_M + r) * 8 + e;
}

struct HeadConvArgs {
    const uint8_t* w;            // [4 chunks][8 k-groups][144 rows][8] bf16 (rows 0..127 policy conv, 128 value conv, rest 0)
    const float* bias;           // [144] (folded BatchNorm)
    __nv_bfloat16* fc_a;         // policy activations in the fc's A-operand layout (fc_a_offset)
    float* value_cells;          // [boards][144] value-conv output after ReLU
    int n_rows;                  // boards * 144
};

__global__ void __launch_bounds__(HEAD_THREADS, 1) hive_head_conv_kernel(const __grid_constant__ CUtensorMap x_map, HeadConvArgs a) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;
    uint8_t* sX = smem + HC_W_BYTES;
    __shared__ uint64_t w_full, x_full[HC_STAGES], x_empty[HC_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&w_full, 1);
        for (int i = 0; i < HC_STAGES; i++) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        mbar_fence_init();
        tma_prefetch_desc(&x_map);
    }
    if (warp == 1) tmem_alloc(&tmem_base, HC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    const int n_tiles = (a.n_rows + HEAD_M - 1) / HEAD_M;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(&w_full, HC_W_BYTES);
            for (int c = 0; c < HC_CHUNKS; c++) bulk_load(sW + c * (HC_W_BYTES / HC_CHUNKS), a.w + c * (HC_W_BYTES / HC_CHUNKS), HC_W_BYTES / HC_CHUNKS, &w_full);
            int st = 0, ph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int c = 0; c < HC_CHUNKS; c++) {
                    mbar_wait(&x_empty[st], ph ^ 1);
                    mbar_expect_tx(&x_full[st], HEAD_A_BYTES);
                    tma_load_3d(sX + st * HEAD_A_BYTES, &x_map, &x_full[st], 0, tile * HEAD_M, c * 8);   // rows past the end: zero-filled
                    if (++st == HC_STAGES) { st = 0; ph ^= 1; }
                }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16(HEAD_M, HC_N);
            const uint64_t a_desc0 = smem_desc(smem_u32(sX), HEAD_M * 16, 128, 0);
            const uint64_t b_desc0 = smem_desc(smem_u32(sW), HC_N * 16, 128, 0);
            mbar_wait(&w_full, 0);
            int st = 0, ph = 0, it = 0;
            uint32_t empty_ph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
                const int slot = it & 1;
                mbar_wait(&acc_empty[slot], ((empty_ph >> slot) & 1u) ^ 1u);
                empty_ph ^= 1u << slot;
                tc_fence_after();
                for (int c = 0; c < HC_CHUNKS; c++) {
                    mbar_wait(&x_full[st], ph);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < HEAD_KC / 16; ks++)
                        mma_bf16(tmem + slot * 256, a_desc0 + (uint64_t)((uint32_t)(st * HEAD_A_BYTES + ks * 2 * HEAD_M * 16) >> 4),
                                 b_desc0 + (uint64_t)((uint32_t)(c * (HC_W_BYTES / HC_CHUNKS) + ks * 2 * HC_N * 16) >> 4), idesc, (c | ks) != 0);
                    mma_commit(&x_empty[st]);
                    if (++st == HC_STAGES) { st = 0; ph ^= 1; }
                }
                mma_commit(&acc_full[slot]);
            }
        }
    } else {
        const int q = warp & 3;                                     // TMEM lane quarter this warp may read
        uint32_t full_ph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int slot = it & 1;
            const int row = tile * HEAD_M + q * 32 + lane;
            const bool valid = row < a.n_rows;
            const int board = row / 144, cell = row - board * 144;
            mbar_wait(&acc_full[slot], (full_ph >> slot) & 1u);
            full_ph ^= 1u << slot;
            tc_fence_after();
#pragma unroll
            for (int g0 = 0; g0 < 9; g0 += 3) {
                uint32_t v[3][16];
#pragma unroll
                for (int g = 0; g < 3; g++) tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + slot * 256 + (g0 + g) * 16, v[g]);
                tmem_ld_wait();
                if (g0 == 6) {                                       // the accumulator is in registers: hand the slot back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[slot]);
                }
                if (!valid) continue;
#pragma unroll
                for (int g = 0; g < 3; g++) {
                    const int c0 = (g0 + g) * 16;
                    if (c0 < 128) {
                        float r[16];
#pragma unroll
                        for (int i = 0; i < 16; i++) r[i] = fmaxf(__uint_as_float(v[g][i]) + __ldg(a.bias + c0 + i), 0.f);
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            uint4 pk;
                            __nv_bfloat162 t;
                            t = __floats2bfloat162_rn(r[8 * h + 0], r[8 * h + 1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
                            t = __floats2bfloat162_rn(r[8 * h + 2], r[8 * h + 3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
                            t = __floats2bfloat162_rn(r[8 * h + 4], r[8 * h + 5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
                            t = __floats2bfloat162_rn(r[8 * h + 6], r[8 * h + 7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
                            *reinterpret_cast<uint4*>(a.fc_a + fc_a_offset(board, cell * 128 + c0 + 8 * h)) = pk;   // cell-major flatten
                        }
                    } else {
                        a.value_cells[(size_t)board * 144 + cell] = fmaxf(__uint_as_float(v[g][0]) + __ldg(a.bias + 128), 0.f);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, HC_TMEM_COLS);
}

struct HeadFcArgs {
    const __nv_bfloat16* fc_a;   // [board tiles][288 chunks][16 KB]
    const uint8_t* w;            // [9 column tiles][288 chunks][8 k-groups][176 rows][8] bf16
    const float* bias;           // [1584]
    float* logits;               // [boards][1584]
    int n_boards;
};

// grid = (9 column tiles, board tiles)
__global__ void __launch_bounds__(HEAD_THREADS, 1) hive_head_fc_kernel(HeadFcArgs a) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + FC_STAGES * HEAD_A_BYTES;
    __shared__ uint64_t full[FC_STAGES], empty[FC_STAGES], acc_full;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = blockIdx.x, mt = blockIdx.y;
    if (tid == 0) {
        for (int i = 0; i < FC_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base, FC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;

    if (warp == 0) {
        if (elect_one()) {
            const uint8_t* ga = reinterpret_cast<const uint8_t*>(a.fc_a) + (size_t)mt * FC_CHUNKS * HEAD_A_BYTES;
            const uint8_t* gb = a.w + (size_t)nt * FC_CHUNKS * FC_B_BYTES;
            int st = 0, ph = 0;
            for (int c = 0; c < FC_CHUNKS; c++) {
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&full[st], HEAD_A_BYTES + FC_B_BYTES);
                bulk_load(sA + st * HEAD_A_BYTES, ga + (size_t)c * HEAD_A_BYTES, HEAD_A_BYTES, &full[st]);
                bulk_load(sB + st * FC_B_BYTES, gb + (size_t)c * FC_B_BYTES, FC_B_BYTES, &full[st]);
                if (++st == FC_STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16(HEAD_M, FC_N);
            const uint64_t a_desc0 = smem_desc(smem_u32(sA), HEAD_M * 16, 128, 0);
            const uint64_t b_desc0 = smem_desc(smem_u32(sB), FC_N * 16, 128, 0);
            int st = 0, ph = 0;
            for (int c = 0; c < FC_CHUNKS; c++) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < HEAD_KC / 16; ks++)
                    mma_bf16(tmem, a_desc0 + (uint64_t)((uint32_t)(st * HEAD_A_BYTES + ks * 2 * HEAD_M * 16) >> 4),
                             b_desc0 + (uint64_t)((uint32_t)(st * FC_B_BYTES + ks * 2 * FC_N * 16) >> 4), idesc, (c | ks) != 0);
                mma_commit(&empty[st]);
                if (++st == FC_STAGES) { st = 0; ph ^= 1; }
            }
            mma_commit(&acc_full);
        }
    } else {
        const int q = warp & 3;
        const int board = mt * HEAD_M + q * 32 + lane;
        mbar_wait(&acc_full, 0);
        tc_fence_after();
        float* out = a.logits + (size_t)board * 1584 + nt * FC_N;
        const float* bias = a.bias + nt * FC_N;
        for (int g = 0; g < FC_N / 16; g++) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + g * 16, v);
            tmem_ld_wait();
            if (board < a.n_boards) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(out + g * 16 + i) =
                        make_float4(__uint_as_float(v[i]) + __ldg(bias + g * 16 + i), __uint_as_float(v[i + 1]) + __ldg(bias + g * 16 + i + 1),
                                    __uint_as_float(v[i + 2]) + __ldg(bias + g * 16 + i + 2), __uint_as_float(v[i + 3]) + __ldg(bias + g * 16 + i + 3));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, FC_TMEM_COLS);
}

struct HeadFinishArgs {
    const float* logits;         // [boards][1584]
    const float* value_cells;    // [boards][144]
    const float* fc1_w; const float* fc1_b;   // [64][144], [64]
    const float* fc2_w; const float* fc2_b;   // [64], [1]
    float* policy;               // [boards][1584] softmax (the search's leaf policy arena)
    double* value;               // [boards] tanh (the search's leaf value arena)
    const uint8_t* mask;         // optional [boards]: rows to produce
    int n_boards;
};

constexpr int FIN_WARPS = 8;
__global__ void __launch_bounds__(FIN_WARPS * 32) hive_head_finish_kernel(HeadFinishArgs a) {
    __shared__ float w1[64 * 145];                               // fc1 weights, rows padded to 145 floats (bank spread)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 64 * 144; i += FIN_WARPS * 32) w1[(i / 144) * 145 + i % 144] = a.fc1_w[i];
    __syncthreads();
    const int b = blockIdx.x * FIN_WARPS + warp;
    if (b >= a.n_boards || (a.mask && !a.mask[b])) return;
    // softmax over the board's logits (alpha_net.py:79-80: logsoftmax(...).exp())
    const float* lg = a.logits + (size_t)b * 1584;
    float x[50];
    float mx = -3.4e38f;
#pragma unroll
    for (int i = 0; i < 50; i++) { const int j = lane + 32 * i; x[i] = j < 1584 ? lg[j] : -3.4e38f; mx = fmaxf(mx, x[i]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 50; i++) { x[i] = lane + 32 * i < 1584 ? __expf(x[i] - mx) : 0.f; sum += x[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    float* p = a.policy + (size_t)b * 1584;
#pragma unroll
    for (int i = 0; i < 50; i++) { const int j = lane + 32 * i; if (j < 1584) p[j] = x[i] * inv; }
    // value head (alpha_net.py:70-73): relu(fc1(v)) -> tanh(fc2(.))
    const float* vc = a.value_cells + (size_t)b * 144;
    float vin[5];
#pragma unroll
    for (int i = 0; i < 5; i++) vin[i] = lane + 32 * i < 144 ? vc[lane + 32 * i] : 0.f;
    float acc = 0.f;
    for (int h = 0; h < 64; h++) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 5; i++) { const int j = lane + 32 * i; if (j < 144) d += w1[h * 145 + j] * vin[i]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        acc += fmaxf(d + a.fc1_b[h], 0.f) * a.fc2_w[h];
    }
    if (lane == 0) a.value[b] = (double)tanhf(acc + a.fc2_b[0]);
}

}  // namespace hive
