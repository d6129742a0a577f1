// conv_selftest.cu -- standalone check of hive_conv3x3_kernel against a CPU convolution, plus timing.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../hive_conv_kernel.cuh"
#include "../hive_conv_host.h"

using namespace hive;

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 5, C = argc > 2 ? atoi(argv[2]) : 256, check = argc > 3 ? atoi(argv[3]) : 1;
    const int nC = C / CONV_CHUNK_CH;
    srand(3);
    std::vector<float> x((size_t)B * 144 * C), w((size_t)256 * C * 9), bias(256), res((size_t)B * 144 * 256);
    auto q = [](float v) { return __bfloat162float(__float2bfloat16(v)); };
    for (auto& v : x) v = q((rand() % 9 - 4) / 4.0f) * ((rand() % 4) == 0);
    for (auto& v : w) v = q((rand() % 15 - 7) / 32.0f);
    for (auto& v : bias) v = (rand() % 11 - 5) / 8.0f;
    for (auto& v : res) v = q((rand() % 9 - 4) / 2.0f);
    std::vector<__nv_bfloat16> hx(x.size()), hres(res.size());
    for (size_t i = 0; i < x.size(); i++) hx[i] = __float2bfloat16(x[i]);
    for (size_t i = 0; i < res.size(); i++) hres[i] = __float2bfloat16(res[i]);
    std::vector<uint8_t> packed;
    pack_conv_weights(w.data(), C, CONV_KG, packed);                       // w is [oc][ic][3][3] fp32
    __nv_bfloat16 *dx, *dres, *dout; uint8_t* dw; float* dbias;
    cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dres, hres.size() * 2); cudaMalloc(&dout, (size_t)B * 144 * 256 * 2);
    cudaMalloc(&dw, packed.size()); cudaMalloc(&dbias, 256 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dres, hres.data(), hres.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, packed.data(), packed.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dbias, bias.data(), 256 * 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xFF, (size_t)B * 144 * 256 * 2);
    CUtensorMap map;
    if (make_board_tensor_map(&map, dx, B, C, CONV_PADW, CONV_PADH, CONV_KG)) { printf("tensor map encode failed\n"); return 2; }
    ConvArgs a; a.weights = dw; a.bias = dbias; a.residual = dres; a.out = dout; a.n_boards = B; a.n_chunks = nC; a.relu = 1;
    cudaFuncSetAttribute(hive_conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM_BYTES);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int items = 2 * ((B + CONV_BOARDS - 1) / CONV_BOARDS), cap = sms * CONV_CTAS_PER_SM, grid = items < cap ? items : cap;
    hive_conv3x3_kernel<<<grid, CONV_THREADS, CONV_SMEM_BYTES>>>(map, a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    if (check) {
        std::vector<__nv_bfloat16> hout((size_t)B * 144 * 256);
        cudaMemcpy(hout.data(), dout, hout.size() * 2, cudaMemcpyDeviceToHost);
        double maxerr = 0; long bad = 0;
        for (int b = 0; b < B; b++)
            for (int y = 0; y < 12; y++)
                for (int xx = 0; xx < 12; xx++)
                    for (int oc = 0; oc < 256; oc++) {
                        double acc = bias[oc];
                        for (int dy = 0; dy < 3; dy++)
                            for (int dx2 = 0; dx2 < 3; dx2++) {
                                int yy = y + dy - 1, xq = xx + dx2 - 1;
                                if (yy < 0 || yy >= 12 || xq < 0 || xq >= 12) continue;
                                const float* xp = &x[((size_t)b * 144 + yy * 12 + xq) * C];
                                for (int ic = 0; ic < C; ic++) acc += (double)xp[ic] * w[((size_t)oc * C + ic) * 9 + dy * 3 + dx2];
                            }
                        size_t o = ((size_t)b * 144 + y * 12 + xx) * 256 + oc;
                        acc += res[o];
                        if (acc < 0) acc = 0;
                        double got = __bfloat162float(hout[o]);
                        double err = fabs(got - acc), tol = 0.01 * fabs(acc) + 0.02;
                        if (!(err <= tol)) { if (bad < 5) printf("mismatch b%d y%d x%d oc%d: got %g want %g\n", b, y, xx, oc, got, acc); bad++; }
                        if (err > maxerr) maxerr = err;
                    }
        printf("B=%d C=%d: max|err| = %g, mismatches = %ld\n", B, C, maxerr, bad);
        if (bad) return 3;
    }
    // timing
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) hive_conv3x3_kernel<<<grid, CONV_THREADS, CONV_SMEM_BYTES>>>(map, a);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; i++) hive_conv3x3_kernel<<<grid, CONV_THREADS, CONV_SMEM_BYTES>>>(map, a);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("B=%d C=%d: %.3f ms/layer, %.1f TFLOP/s (useful flops)\n", B, C, ms, 2.0 * B * 144 * 256 * (double)C * 9 / ms / 1e9);
    return 0;
}
