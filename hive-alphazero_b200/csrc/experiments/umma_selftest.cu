// umma_selftest.cu -- standalone check of the tcgen05 descriptor conventions used by the conv kernel
// (K-major, no-swizzle operands laid out [k-chunk of 8][row][16 B]); prints max |error| vs a CPU GEMM
// for a few (LBO, SBO) interpretations.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_selftest umma_selftest.cu
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../umma.cuh"

using namespace umma;
constexpr int M = 128, K = 64;

// A: [M][K] row-major bf16 in global; B: [NROWS][K]; D: [M][N] fp32.  shift = first B row used (row-shift test)
__global__ void __launch_bounds__(128) gemm_test(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int N, int nrows,
                                                 int shift, uint32_t lboA, uint32_t sboA, uint32_t lboB, uint32_t sboB) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t* sA = smem;                       // [8 chunks][128 rows][16 B]
    uint8_t* sB = smem + 8 * M * 16;          // [8 chunks][nrows][16 B]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * 8; i += 128) {
        int r = i % M, c = i / M;
        *reinterpret_cast<uint4*>(sA + (c * M + r) * 16) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
    }
    for (int i = tid; i < nrows * 8; i += 128) {
        int r = i % nrows, c = i / nrows;
        *reinterpret_cast<uint4*>(sB + (c * nrows + r) * 16) = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
    }
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t id = idesc_bf16(M, N);
        for (int s = 0; s < K / 16; s++) {
            const uint64_t ad = smem_desc(smem_u32(sA) + 2 * s * M * 16, lboA, sboA, 0);
            const uint64_t bd = smem_desc(smem_u32(sB) + 2 * s * nrows * 16 + shift * 16, lboB, sboB, 0);
            mma_bf16(tm, ad, bd, id, s > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

int main() {
    const int N = 176, nrows = 208, shift = 17;
    std::vector<__nv_bfloat16> hA(M * K), hB(nrows * K);
    std::vector<float> fA(M * K), fB(nrows * K);
    srand(1);
    for (int i = 0; i < M * K; i++) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < nrows * K; i++) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = 8 * M * 16 + 8 * nrows * 16 + 1024;
    cudaFuncSetAttribute(gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct Cfg { const char* name; uint32_t lboA, sboA, lboB, sboB; int shift; };
    Cfg cfgs[] = {
        {"LBO=chunk-plane stride, SBO=128 (8 rows x 16B), shift 0", (uint32_t)M * 16, 128, (uint32_t)nrows * 16, 128, 0},
        {"same, B rows shifted by 17", (uint32_t)M * 16, 128, (uint32_t)nrows * 16, 128, shift},
        {"swapped: LBO=128, SBO=chunk-plane stride, shift 0", 128, (uint32_t)M * 16, 128, (uint32_t)nrows * 16, 0},
    };
    for (const Cfg& c : cfgs) {
        cudaMemset(dD, 0xFF, M * N * 4);
        gemm_test<<<1, 128, smem>>>(dA, dB, dD, N, nrows, c.shift, c.lboA, c.sboA, c.lboB, c.sboB);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
        std::vector<float> hD(M * N);
        cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < M; m++)
            for (int n = 0; n < N; n++) {
                double ref = 0;
                for (int k = 0; k < K; k++) ref += (double)fA[m * K + k] * fB[(n + c.shift) * K + k];
                double err = fabs(ref - hD[m * N + n]);
                if (!(err <= 1e30)) err = 1e30;
                if (err > maxerr) maxerr = err;
            }
        printf("%-60s max|err| = %g  D[0][0..3] = %g %g %g %g\n", c.name, maxerr, hD[0], hD[1], hD[2], hD[3]);
    }
    return 0;
}
