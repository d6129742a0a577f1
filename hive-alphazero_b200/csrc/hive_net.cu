// hive_net.cu -- C ABI of the network on the tensor cores (include/hive_b200.h, net_*): trunk and heads.
// Stem + 19 residual blocks of alpha_zero/alpha_net.py (ConvBlock :25-34, ResBlock :36-54) = 39 3x3
// convolutions with folded BatchNorm, each one launch of hive_conv3x3_kernel (hive_conv_kernel.cuh).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "hive_conv_host.h"
#include "hive_conv_kernel.cuh"
#include "hive_conv2_kernel.cuh"
#include "hive_heads_kernel.cuh"
#include "hive_internal.h"

using namespace hive;

constexpr int NET_LAYERS = 39;

struct hive_net {
    int device = 0, max_boards = 0, sms = 148;
    cudaStream_t stream = nullptr;
    __nv_bfloat16* x0 = nullptr;          // [max][144][64]  stem input, NHWC, channels 56..63 zero
    __nv_bfloat16* act[3] = {nullptr, nullptr, nullptr};   // [max][144][256] ping / temp / pong
    CUtensorMap map_x0, map_act[3];
    uint8_t* weights[NET_LAYERS] = {};
    float* bias[NET_LAYERS] = {};
    int n_chunks[NET_LAYERS] = {};
    int loaded = 0;
    long long launches = 0;
    // heads (alpha_net.py:56-80; kernels in hive_heads_kernel.cuh)
    CUtensorMap map_rows;                 // act[2] (the trunk's output after 19 blocks) as [boards*144][256] rows
    uint8_t* head_w1 = nullptr;           // both 1x1 convolutions packed as one 144-row operand
    float* head_b1 = nullptr;             // [144]
    uint8_t* head_wfc = nullptr;          // policy fc, nine 176-row column tiles x 288 chunks
    float* head_bfc = nullptr;            // [1584]
    float* head_fc1_w = nullptr; float* head_fc1_b = nullptr; float* head_fc2_w = nullptr; float* head_fc2_b = nullptr;
    __nv_bfloat16* fc_a = nullptr;        // policy activations in the fc's A-operand layout
    float* value_cells = nullptr;         // [max][144]
    float* logits = nullptr;              // [max][1584]
    bool heads_loaded = false;
};

namespace {

// planes CHW bf16 [B][56][144] (the encoder's output) -> NHWC [B][144][64], zero channels 56..63
__global__ void __launch_bounds__(256) chw_to_nhwc64_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int n_boards) {
    __shared__ uint16_t tile[56][144 + 2];
    const int b = blockIdx.x;
    if (b >= n_boards) return;
    const uint16_t* src = in + (size_t)b * 56 * 144;
    for (int i = threadIdx.x; i < 56 * 144; i += 256) tile[i / 144][i % 144] = src[i];
    __syncthreads();
    uint16_t* dst = out + (size_t)b * 144 * 64;
    for (int i = threadIdx.x; i < 144 * 64; i += 256) {
        const int px = i >> 6, c = i & 63;
        dst[i] = c < 56 ? tile[c][px] : (uint16_t)0;
    }
}

// ---- weight packing on the device (a reload after a weight broadcast sits in self-play's timed region)
// w: [256][cin][3][3] fp32 -> [2 halves][9 taps][cpad/64 chunks][8 k-groups][128 rows][8 ch] bf16 (pack_conv_weights' layout)
__global__ void pack_conv_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin, int cpad) {
    const int nC = cpad / (8 * CONV_KG);
    const size_t total = (size_t)2 * 9 * nC * CONV_KG * 128 * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = r % 8; r /= 8;
        const int row = r % 128; r /= 128;
        const int kg = r % CONV_KG; r /= CONV_KG;
        const int c = r % nC; r /= nC;
        const int t = r % 9; const int half = (int)(r / 9);
        const int oc = half * 128 + row, ic = (c * CONV_KG + kg) * 8 + e;
        out[i] = __float2bfloat16(ic < cin ? w[((size_t)oc * cin + ic) * 9 + t] : 0.f);
    }
}
// both 1x1 convolutions -> [4 chunks][8 k-groups][144 rows][8] bf16; bias -> [144]
__global__ void pack_head_conv_kernel(const float* __restrict__ pw, const float* __restrict__ pb, const float* __restrict__ vw,
                                      const float* __restrict__ vb, __nv_bfloat16* __restrict__ out, float* __restrict__ bias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < HC_W_BYTES / 2) {
        int r = i;
        const int e = r % 8; r /= 8;
        const int row = r % HC_N; r /= HC_N;
        const int kg = r % 8; const int c = r / 8;
        const int k = (c * 8 + kg) * 8 + e;
        out[i] = __float2bfloat16(row < 128 ? pw[(size_t)row * 256 + k] : row == 128 ? vw[k] : 0.f);
    }
    if (i < HC_N) bias[i] = i < 128 ? pb[i] : i == 128 ? vb[0] : 0.f;
}
// policy fc in the REFERENCE's layout [1584][128 channels * 144 cells] (alpha_net.py:77 flattens channel-major) ->
// [9 column tiles][288 chunks][8 k-groups][176 rows][8] bf16 with cell-major K (k = cell*128 + channel)
__global__ void pack_head_fc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
    const size_t total = (size_t)1584 * FC_K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = r % 8; r /= 8;
        const int row = r % FC_N; r /= FC_N;
        const int kg = r % 8; r /= 8;
        const int kc = r % FC_CHUNKS; const int nt = (int)(r / FC_CHUNKS);
        const int k = (kc * 8 + kg) * 8 + e, cell = k >> 7, ch = k & 127;
        out[i] = __float2bfloat16(w[(size_t)(nt * FC_N + row) * FC_K + ch * 144 + cell]);
    }
}

int check(const hive_net* n) { return n && n->max_boards > 0 ? 0 : fail(HIVE_E_HANDLE, "bad net handle"); }

int launch_conv(hive_net* n, const CUtensorMap& in_map, int layer, const __nv_bfloat16* residual, __nv_bfloat16* out, int boards) {
    ConvArgs a;
    a.weights = n->weights[layer]; a.bias = n->bias[layer]; a.residual = residual; a.out = out;
    a.n_boards = boards; a.n_chunks = n->n_chunks[layer]; a.relu = 1;
    static const int pair_mode = getenv("HIVE_B200_CONV_PAIR") ? atoi(getenv("HIVE_B200_CONV_PAIR")) : 1;     // 0: the single-CTA kernel
    if (pair_mode) {   // CTA pairs (cta_group::2): one cluster of two CTAs per board pair, both out-channel halves
        const int n_pairs = (boards + 1) / 2, max_clusters = n->sms / 2;
        const int clusters = n_pairs < max_clusters ? n_pairs : max_clusters;
        hive_conv3x3_pair_kernel<<<2 * clusters, C2_THREADS, C2_SMEM_BYTES, n->stream>>>(in_map, a);
        CUDA_TRY(cudaGetLastError());
        n->launches++;
        return 0;
    }
    const int items = 2 * ((boards + CONV_BOARDS - 1) / CONV_BOARDS);
    const int cap = n->sms * CONV_CTAS_PER_SM;
    hive_conv3x3_kernel<<<items < cap ? items : cap, CONV_THREADS, CONV_SMEM_BYTES, n->stream>>>(in_map, a);
    CUDA_TRY(cudaGetLastError());
    n->launches++;
    return 0;
}

}  // namespace

extern "C" {

int net_create(int device, void* stream, int max_boards, hive_net_t** out) {
    if (!out || max_boards < 1) return fail(HIVE_E_ARG, "net_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(HIVE_E_CUDA, "net_create: no such CUDA device (there is no CPU fallback)");
    CUDA_TRY(cudaSetDevice(device));
    hive_net* n = new hive_net();
    n->device = device; n->max_boards = max_boards; n->stream = (cudaStream_t)stream;
    cudaDeviceGetAttribute(&n->sms, cudaDevAttrMultiProcessorCount, device);
    const size_t B = (size_t)max_boards;
    CUDA_TRY(cudaMalloc(&n->x0, B * 144 * 64 * 2));
    for (int i = 0; i < 3; i++) CUDA_TRY(cudaMalloc(&n->act[i], B * 144 * 256 * 2));
    if (make_board_tensor_map(&n->map_x0, n->x0, max_boards, 64, CONV_PADW, CONV_PADH, CONV_KG)) return fail(HIVE_E_CUDA, "net_create: cuTensorMapEncodeTiled failed (stem input)");
    for (int i = 0; i < 3; i++)
        if (make_board_tensor_map(&n->map_act[i], n->act[i], max_boards, 256, CONV_PADW, CONV_PADH, CONV_KG)) return fail(HIVE_E_CUDA, "net_create: cuTensorMapEncodeTiled failed");
    CUDA_TRY(cudaFuncSetAttribute(hive_conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(hive_conv3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES));
    // heads
    if (make_rows_tensor_map(&n->map_rows, n->act[2], (uint64_t)B * 144, 256, HEAD_M, 8)) return fail(HIVE_E_CUDA, "net_create: cuTensorMapEncodeTiled failed (head rows)");
    const size_t mt = (B + HEAD_M - 1) / HEAD_M;
    CUDA_TRY(cudaMalloc(&n->head_w1, HC_W_BYTES));
    CUDA_TRY(cudaMalloc(&n->head_b1, HC_N * 4));
    CUDA_TRY(cudaMalloc(&n->head_wfc, (size_t)FC_NT * FC_CHUNKS * FC_B_BYTES));
    CUDA_TRY(cudaMalloc(&n->head_bfc, 1584 * 4));
    CUDA_TRY(cudaMalloc(&n->head_fc1_w, 64 * 144 * 4)); CUDA_TRY(cudaMalloc(&n->head_fc1_b, 64 * 4));
    CUDA_TRY(cudaMalloc(&n->head_fc2_w, 64 * 4)); CUDA_TRY(cudaMalloc(&n->head_fc2_b, 4));
    CUDA_TRY(cudaMalloc(&n->fc_a, mt * FC_CHUNKS * HEAD_A_BYTES));
    CUDA_TRY(cudaMemset(n->fc_a, 0, mt * FC_CHUNKS * HEAD_A_BYTES));      // rows of boards past the batch stay finite
    CUDA_TRY(cudaMalloc(&n->value_cells, B * 144 * 4));
    CUDA_TRY(cudaMalloc(&n->logits, B * 1584 * 4));
    CUDA_TRY(cudaFuncSetAttribute(hive_head_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HC_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(hive_head_fc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM_BYTES));
    *out = n;
    return 0;
}

int net_destroy(hive_net_t* n) {
    if (!n) return 0;
    cudaSetDevice(n->device);
    cudaStreamSynchronize(n->stream);
    cudaFree(n->x0);
    for (int i = 0; i < 3; i++) cudaFree(n->act[i]);
    for (int i = 0; i < NET_LAYERS; i++) { cudaFree(n->weights[i]); cudaFree(n->bias[i]); }
    cudaFree(n->head_w1); cudaFree(n->head_b1); cudaFree(n->head_wfc); cudaFree(n->head_bfc);
    cudaFree(n->head_fc1_w); cudaFree(n->head_fc1_b); cudaFree(n->head_fc2_w); cudaFree(n->head_fc2_b);
    cudaFree(n->fc_a); cudaFree(n->value_cells); cudaFree(n->logits);
    delete n;
    return 0;
}

// layer 0 = stem (cin = 56), layers 1+2i / 2+2i = conv1 / conv2 of residual block i (cin = 256).
// w: [256][cin][3][3] fp32 with BatchNorm folded in, bias: [256] fp32.
int net_load_conv_host(hive_net_t* n, int layer, const float* w, const float* bias, int cin) {
    if (check(n)) return HIVE_E_HANDLE;
    if (layer < 0 || layer >= NET_LAYERS || !w || !bias || (cin != 56 && cin != 256) || (layer == 0) != (cin == 56))
        return fail(HIVE_E_ARG, "net_load_conv_host: bad arguments");
    CUDA_TRY(cudaSetDevice(n->device));
    const int cpad = cin == 56 ? 64 : 256;
    std::vector<float> wp((size_t)256 * cpad * 9, 0.f);
    for (int oc = 0; oc < 256; oc++)
        for (int ic = 0; ic < cin; ic++)
            memcpy(&wp[((size_t)oc * cpad + ic) * 9], &w[((size_t)oc * cin + ic) * 9], 9 * sizeof(float));
    std::vector<uint8_t> packed;
    pack_conv_weights(wp.data(), cpad, CONV_KG, packed);
    CUDA_TRY(cudaStreamSynchronize(n->stream));
    // a reload (new weights after a broadcast) goes into the SAME device buffers: a layer's packed size depends on
    // its shape only, and CUDA graphs captured over this trunk keep pointing at valid, current weights
    if (!n->weights[layer]) CUDA_TRY(cudaMalloc(&n->weights[layer], packed.size()));
    if (!n->bias[layer]) CUDA_TRY(cudaMalloc(&n->bias[layer], 256 * 4));
    CUDA_TRY(cudaMemcpy(n->weights[layer], packed.data(), packed.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(n->bias[layer], bias, 256 * 4, cudaMemcpyHostToDevice));
    n->n_chunks[layer] = cpad / CONV_CHUNK_CH;
    n->loaded |= 0;   // counted below
    int cnt = 0;
    for (int i = 0; i < NET_LAYERS; i++) cnt += n->weights[i] != nullptr;
    n->loaded = cnt;
    return 0;
}

// planes: device bf16 [n_boards][56][144] (hive_dev_planes / mcts_dev_leaf_planes).  On return
// *out_nhwc is a device pointer to the trunk output [n_boards][144][256] bf16 (NHWC), valid until the
// next call; the work is queued on the handle's stream.
int net_trunk_forward(hive_net_t* n, const uint16_t* planes_chw_dev, int n_boards, uint16_t** out_nhwc) {
    if (check(n)) return HIVE_E_HANDLE;
    if (!planes_chw_dev || !out_nhwc || n_boards < 1 || n_boards > n->max_boards) return fail(HIVE_E_ARG, "net_trunk_forward: bad arguments");
    if (n->loaded != NET_LAYERS) return fail(HIVE_E_ARG, "net_trunk_forward: not all 39 convolutions are loaded");
    CUDA_TRY(cudaSetDevice(n->device));
    chw_to_nhwc64_kernel<<<n_boards, 256, 0, n->stream>>>(planes_chw_dev, reinterpret_cast<uint16_t*>(n->x0), n_boards);
    CUDA_TRY(cudaGetLastError());
    n->launches++;
    int cur = 0;                                            // act[cur] holds the block input
    int rc = launch_conv(n, n->map_x0, 0, nullptr, n->act[cur], n_boards);
    if (rc) return rc;
    for (int blk = 0; blk < 19; blk++) {
        const int tmp = 1, nxt = cur == 0 ? 2 : 0;
        rc = launch_conv(n, n->map_act[cur], 1 + 2 * blk, nullptr, n->act[tmp], n_boards);             // relu(bn1(conv1(x)))
        if (rc) return rc;
        rc = launch_conv(n, n->map_act[tmp], 2 + 2 * blk, n->act[cur], n->act[nxt], n_boards);         // relu(bn2(conv2(.)) + x)
        if (rc) return rc;
        cur = nxt;
    }
    *out_nhwc = reinterpret_cast<uint16_t*>(n->act[cur]);
    return 0;
}

// The same two loads from DEVICE tensors (fp32, e.g. the torch module's own parameters after folding): packing runs as
// kernels on the handle's stream, nothing crosses PCIe -- the reload after a weight broadcast takes a few milliseconds.
// net_load_heads_dev takes fc_w in the REFERENCE's layout ([1584][128*144], channel-major columns).
int net_load_conv_dev(hive_net_t* n, int layer, const float* w_dev, const float* bias_dev, int cin) {
    if (check(n)) return HIVE_E_HANDLE;
    if (layer < 0 || layer >= NET_LAYERS || !w_dev || !bias_dev || (cin != 56 && cin != 256) || (layer == 0) != (cin == 56))
        return fail(HIVE_E_ARG, "net_load_conv_dev: bad arguments");
    CUDA_TRY(cudaSetDevice(n->device));
    const int cpad = cin == 56 ? 64 : 256;
    const size_t bytes = (size_t)2 * 9 * cpad * 128 * 2;
    if (!n->weights[layer]) CUDA_TRY(cudaMalloc(&n->weights[layer], bytes));
    if (!n->bias[layer]) CUDA_TRY(cudaMalloc(&n->bias[layer], 256 * 4));
    pack_conv_kernel<<<256, 256, 0, n->stream>>>(w_dev, reinterpret_cast<__nv_bfloat16*>(n->weights[layer]), cin, cpad);
    CUDA_TRY(cudaMemcpyAsync(n->bias[layer], bias_dev, 256 * 4, cudaMemcpyDeviceToDevice, n->stream));
    CUDA_TRY(cudaGetLastError());
    n->n_chunks[layer] = cpad / CONV_CHUNK_CH;
    int cnt = 0;
    for (int i = 0; i < NET_LAYERS; i++) cnt += n->weights[i] != nullptr;
    n->loaded = cnt;
    return 0;
}

int net_load_heads_dev(hive_net_t* n, const float* pconv_w, const float* pconv_b, const float* vconv_w, const float* vconv_b,
                       const float* fc_w_ref, const float* fc_b, const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b) {
    if (check(n)) return HIVE_E_HANDLE;
    if (!pconv_w || !pconv_b || !vconv_w || !vconv_b || !fc_w_ref || !fc_b || !fc1_w || !fc1_b || !fc2_w || !fc2_b)
        return fail(HIVE_E_ARG, "net_load_heads_dev: null argument");
    CUDA_TRY(cudaSetDevice(n->device));
    pack_head_conv_kernel<<<(HC_W_BYTES / 2 + 255) / 256, 256, 0, n->stream>>>(pconv_w, pconv_b, vconv_w, vconv_b,
                                                                               reinterpret_cast<__nv_bfloat16*>(n->head_w1), n->head_b1);
    pack_head_fc_kernel<<<1184, 256, 0, n->stream>>>(fc_w_ref, reinterpret_cast<__nv_bfloat16*>(n->head_wfc));
    CUDA_TRY(cudaMemcpyAsync(n->head_bfc, fc_b, 1584 * 4, cudaMemcpyDeviceToDevice, n->stream));
    CUDA_TRY(cudaMemcpyAsync(n->head_fc1_w, fc1_w, 64 * 144 * 4, cudaMemcpyDeviceToDevice, n->stream));
    CUDA_TRY(cudaMemcpyAsync(n->head_fc1_b, fc1_b, 64 * 4, cudaMemcpyDeviceToDevice, n->stream));
    CUDA_TRY(cudaMemcpyAsync(n->head_fc2_w, fc2_w, 64 * 4, cudaMemcpyDeviceToDevice, n->stream));
    CUDA_TRY(cudaMemcpyAsync(n->head_fc2_b, fc2_b, 4, cudaMemcpyDeviceToDevice, n->stream));
    CUDA_TRY(cudaGetLastError());
    n->heads_loaded = true;
    return 0;
}

// The heads (alpha_net.py:56-80), BatchNorm folded into the two 1x1 convolutions by the caller:
//   pconv_w [128][256], pconv_b [128]   policy conv;   vconv_w [256], vconv_b [1]   value conv
//   fc_w [1584][18432] with CELL-MAJOR columns (k = cell*128 + channel: the reference's channel-major flatten
//   c*144 + cell of alpha_net.py:77 permuted once), fc_b [1584];  fc1 [64][144] + [64], fc2 [64] + [1]   value MLP.
// A reload writes into the same device buffers (captured graphs stay valid).
int net_load_heads_host(hive_net_t* n, const float* pconv_w, const float* pconv_b, const float* vconv_w, const float* vconv_b,
                        const float* fc_w, const float* fc_b, const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b) {
    if (check(n)) return HIVE_E_HANDLE;
    if (!pconv_w || !pconv_b || !vconv_w || !vconv_b || !fc_w || !fc_b || !fc1_w || !fc1_b || !fc2_w || !fc2_b)
        return fail(HIVE_E_ARG, "net_load_heads_host: null argument");
    CUDA_TRY(cudaSetDevice(n->device));
    CUDA_TRY(cudaStreamSynchronize(n->stream));
    {   // [4 chunks][8 k-groups][144 rows][8]: rows 0..127 policy conv, row 128 value conv, the rest zero
        std::vector<__nv_bfloat16> w1((size_t)HC_W_BYTES / 2, __float2bfloat16(0.f));
        std::vector<float> b1(HC_N, 0.f);
        for (int r = 0; r <= 128; r++)
            for (int k = 0; k < 256; k++) {
                const float v = r < 128 ? pconv_w[(size_t)r * 256 + k] : vconv_w[k];
                w1[((((size_t)(k >> 6) * 8 + ((k >> 3) & 7)) * HC_N + r) * 8) + (k & 7)] = __float2bfloat16(v);
            }
        for (int r = 0; r < 128; r++) b1[r] = pconv_b[r];
        b1[128] = vconv_b[0];
        CUDA_TRY(cudaMemcpy(n->head_w1, w1.data(), HC_W_BYTES, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(n->head_b1, b1.data(), HC_N * 4, cudaMemcpyHostToDevice));
    }
    {   // [9 column tiles][288 chunks][8 k-groups][176 rows][8]
        std::vector<__nv_bfloat16> wf((size_t)FC_NT * FC_CHUNKS * FC_B_BYTES / 2);
        __nv_bfloat16* wfp = wf.data();
        parallel_for(1584, [=](int o) {
            const int nt = o / FC_N, rr = o - nt * FC_N;
            const float* src = fc_w + (size_t)o * FC_K;
            for (int k = 0; k < FC_K; k += 8) {                  // eight consecutive k share a 16-byte destination row
                __nv_bfloat16* dst = wfp + ((((size_t)nt * FC_CHUNKS + (k >> 6)) * 8 + ((k >> 3) & 7)) * FC_N + rr) * 8;
                for (int e = 0; e < 8; e++) dst[e] = __float2bfloat16(src[k + e]);
            }
        });
        CUDA_TRY(cudaMemcpy(n->head_wfc, wf.data(), wf.size() * 2, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMemcpy(n->head_bfc, fc_b, 1584 * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(n->head_fc1_w, fc1_w, 64 * 144 * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(n->head_fc1_b, fc1_b, 64 * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(n->head_fc2_w, fc2_w, 64 * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(n->head_fc2_b, fc2_b, 4, cudaMemcpyHostToDevice));
    n->heads_loaded = true;
    return 0;
}

// ChessNet.forward (alpha_net.py:82-95) for n_boards positions: planes (device bf16 [n][56][144]) -> policy_dev (device
// float32 [n][1584], softmax) and value_dev (device float64 [n], tanh) -- e.g. mcts_dev_leaf_planes / _policy / _value, so a
// search wave runs without a single copy or library kernel.  Queued on the handle's stream.
int net_forward(hive_net_t* n, const uint16_t* planes_chw_dev, int n_boards, float* policy_dev, double* value_dev) {
    if (check(n)) return HIVE_E_HANDLE;
    if (!policy_dev || !value_dev) return fail(HIVE_E_ARG, "net_forward: null output");
    if (!n->heads_loaded) return fail(HIVE_E_ARG, "net_forward: the heads are not loaded (net_load_heads_host)");
    uint16_t* nhwc = nullptr;
    int rc = net_trunk_forward(n, planes_chw_dev, n_boards, &nhwc);
    if (rc) return rc;
    if (reinterpret_cast<__nv_bfloat16*>(nhwc) != n->act[2]) return fail(HIVE_E_CUDA, "net_forward: unexpected trunk output buffer");
    HeadConvArgs hc;
    hc.w = n->head_w1; hc.bias = n->head_b1; hc.fc_a = n->fc_a; hc.value_cells = n->value_cells; hc.n_rows = n_boards * 144;
    const int tiles = (hc.n_rows + HEAD_M - 1) / HEAD_M;
    hive_head_conv_kernel<<<tiles < n->sms ? tiles : n->sms, HEAD_THREADS, HC_SMEM_BYTES, n->stream>>>(n->map_rows, hc);
    HeadFcArgs hf;
    hf.fc_a = n->fc_a; hf.w = n->head_wfc; hf.bias = n->head_bfc; hf.logits = n->logits; hf.n_boards = n_boards;
    hive_head_fc_kernel<<<dim3(FC_NT, (n_boards + HEAD_M - 1) / HEAD_M), HEAD_THREADS, FC_SMEM_BYTES, n->stream>>>(hf);
    HeadFinishArgs fin;
    fin.logits = n->logits; fin.value_cells = n->value_cells; fin.fc1_w = n->head_fc1_w; fin.fc1_b = n->head_fc1_b;
    fin.fc2_w = n->head_fc2_w; fin.fc2_b = n->head_fc2_b; fin.policy = policy_dev; fin.value = value_dev; fin.mask = nullptr;
    fin.n_boards = n_boards;
    hive_head_finish_kernel<<<(n_boards + FIN_WARPS - 1) / FIN_WARPS, FIN_WARPS * 32, 0, n->stream>>>(fin);
    CUDA_TRY(cudaGetLastError());
    n->launches += 3;
    return 0;
}

long long net_launch_count(const hive_net_t* n) { return n ? n->launches : 0; }

}  // extern "C"
