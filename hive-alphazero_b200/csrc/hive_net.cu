// hive_net.cu -- C ABI of the network trunk on the tensor cores (include/hive_b200.h, net_*).
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

int check(const hive_net* n) { return n && n->max_boards > 0 ? 0 : fail(HIVE_E_HANDLE, "bad net handle"); }

int launch_conv(hive_net* n, const CUtensorMap& in_map, int layer, const __nv_bfloat16* residual, __nv_bfloat16* out, int boards) {
    ConvArgs a;
    a.weights = n->weights[layer]; a.bias = n->bias[layer]; a.residual = residual; a.out = out;
    a.n_boards = boards; a.n_chunks = n->n_chunks[layer]; a.relu = 1;
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

long long net_launch_count(const hive_net_t* n) { return n ? n->launches : 0; }

}  // extern "C"
