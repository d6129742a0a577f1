// hive_conv_host.h -- host helpers of the tensor-core convolution: weight packing into the UMMA operand
// layout and the 5-D TMA descriptor of an NHWC board tensor.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>

namespace hive {

// fn(i) for i in [0, n) on a handful of host threads (weight packing after a broadcast sits in self-play's timed region)
template <typename F>
inline void parallel_for(int n, F fn) {
    unsigned hc = std::thread::hardware_concurrency();
    int nt = hc > 8 ? 8 : (hc > 1 ? (int)hc : 1);
    if (nt > n) nt = n;
    if (nt <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([=] { for (int i = t; i < n; i += nt) fn(i); });
    for (auto& x : th) x.join();
}

// w: [256 oc][C ic][3][3] fp32 (BatchNorm already folded) -> [2 halves][9 taps][C/(8 KG) chunks][KG k-groups][128 rows][8 ch] bf16
inline void pack_conv_weights(const float* w, int C, int KG, std::vector<uint8_t>& out) {
    const int nC = C / (8 * KG);
    out.assign((size_t)2 * 9 * nC * KG * 128 * 16, 0);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out.data());
    parallel_for(2 * 9, [=](int ht) {
        const int half = ht / 9, t = ht % 9;
        for (int c = 0; c < nC; c++)
            for (int kg = 0; kg < KG; kg++)
                for (int r = 0; r < 128; r++)
                    for (int e = 0; e < 8; e++) {
                        const int oc = half * 128 + r, ic = (c * KG + kg) * 8 + e;
                        const size_t dst = ((((size_t)(half * 9 + t) * nC + c) * KG + kg) * 128 + r) * 8 + e;
                        o[dst] = __float2bfloat16(w[((size_t)oc * C + ic) * 9 + t]);
                    }
    });
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NHWC bf16 tensor [B][12][12][C] seen as 5-D {8 ch, W, H, C/8 ch-groups, B}; box = one zero-padded board (box_w slots x box_h rows,
// out-of-range pixels zero-filled) of box_groups x 8 channels
inline int make_board_tensor_map(CUtensorMap* map, const void* base, int B, int C, int box_w, int box_h, int box_groups) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return -1;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t dims[5] = {8, 12, 12, (cuuint64_t)(C / 8), (cuuint64_t)B};
    const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)12 * C * 2, 16, (cuuint64_t)144 * C * 2};   // bytes, dims 1..4
    const cuuint32_t box[5] = {8, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_groups, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// Row-major bf16 matrix [rows][K] (K multiple of 8) seen as 3-D {8 elements, rows, K/8 k-groups}; box = {8, box_rows,
// box_groups}: lands in shared memory as [k-group][row][8 elements], the K-major no-swizzle UMMA operand layout.
// Rows past the end are zero-filled.
inline int make_rows_tensor_map(CUtensorMap* map, const void* base, uint64_t rows, int K, int box_rows, int box_groups) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return -1;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t dims[3] = {8, (cuuint64_t)rows, (cuuint64_t)(K / 8)};
    const cuuint64_t strides[2] = {(cuuint64_t)K * 2, 16};       // bytes, dims 1..2
    const cuuint32_t box[3] = {8, (cuuint32_t)box_rows, (cuuint32_t)box_groups};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace hive
