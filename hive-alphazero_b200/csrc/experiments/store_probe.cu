// Write-only HBM stream probe: which store form gets closest to the pin bandwidth?  The env step is 96% writes
// (DESIGN.md section 5), so this number -- not the read+write copy figure -- bounds the encode kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

enum { PLAIN = 0, CS = 1, WT = 2, CG = 3, NOALLOC = 4 };

template <int MODE, int UNROLL>
__global__ void store_kernel(uint4* __restrict__ dst, size_t n_vec, uint32_t tag) {
    const uint4 v = make_uint4(tag, tag + 1, tag + 2, threadIdx.x);
    const size_t stride = (size_t)gridDim.x * blockDim.x * UNROLL;
    for (size_t base = (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; base < n_vec; base += stride) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint4* p = dst + base + (size_t)u * blockDim.x;
            if (base + (size_t)u * blockDim.x >= n_vec) break;
            if (MODE == PLAIN) *p = v;
            else if (MODE == CS) __stcs(p, v);
            else if (MODE == WT) __stwt(p, v);
            else if (MODE == CG) __stcg(p, v);
            else asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
}

// warp-contiguous variant: every warp owns CHUNK bytes at a time (like one game's planes)
template <int CHUNK>
__global__ void store_chunk_kernel(uint4* __restrict__ dst, size_t n_chunks, uint32_t tag) {
    const uint4 v = make_uint4(tag, tag + 1, tag + 2, threadIdx.x);
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t c = warp; c < n_chunks; c += n_warps) {
        uint4* p = dst + c * (CHUNK / 16);
#pragma unroll 8
        for (int i = lane; i < CHUNK / 16; i += 32) p[i] = v;
    }
}

// TMA bulk store: one elected thread per CTA streams the same smem tile to successive global chunks
template <int CHUNK, int DEPTH>
__global__ void store_bulk_kernel(uint8_t* __restrict__ dst, size_t n_chunks, uint32_t tag) {
    extern __shared__ __align__(128) uint8_t smem[];
    for (int i = threadIdx.x; i < CHUNK / 4; i += blockDim.x) ((uint32_t*)smem)[i] = tag + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * CHUNK), "r"(s), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH) : "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// encode-like: non-persistent grid, one warp per 16128-B chunk, optional LUT prologue / LUT reads / small side writes
template <int FLAGS>   // 1 = LUT fill + barrier, 2 = data through smem byte -> LUT lookups, 4 = small side writes, 8 = global reads first
__global__ void __launch_bounds__(256) store_like_encode(uint4* __restrict__ dst, size_t n_chunks, uint32_t tag,
                                                          uint32_t* __restrict__ side, const uint4* __restrict__ src) {
    __shared__ uint4 lut[256];
    __shared__ uint8_t bytes[8][1136];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (FLAGS & 1) {
        uint4 v; v.x = (tid & 1) ? 0x3F80u : 0u; v.y = (tid & 4) ? 0x3F80u : 0u; v.z = (tid & 16) ? 0x3F80u : 0u; v.w = (tid & 64) ? 0x3F80u : 0u;
        lut[tid] = v;
        for (int i = lane; i < 1136; i += 32) bytes[warp][i] = (uint8_t)((i * 37 + tag) & ((i % 7 == 0) ? 0xFF : 0));
        __syncthreads();
    }
    const size_t c = (size_t)blockIdx.x * 8 + warp;
    if (c >= n_chunks) return;
    uint4 v = make_uint4(tag, tag + 1, tag + 2, threadIdx.x);
    if (FLAGS & 8) {
        const uint4 r = src[c * 64 + lane];
        if (r.x == 0x12345678u) return;
        v.w ^= r.y;
    }
    uint4* p = dst + c * 1008;
    int pl = lane / 18, j = lane - 18 * pl;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        if (i < 31 || lane < 16) {
            if (FLAGS & 2) v = lut[bytes[warp][pl * 20 + j]];
            p[i * 32 + lane] = v;
        }
        j += 14; pl += 1;
        if (j >= 18) { j -= 18; pl += 1; }
    }
    if (FLAGS & 4) {
        if (lane == 0) { side[c] = tag; side[n_chunks + c] = tag; side[2 * n_chunks + c * 96 + 11] = tag; side[2 * n_chunks + c * 96 + 14] = tag; }
        if (lane < 20) reinterpret_cast<uint4*>(side + 2 * n_chunks + c * 96 + 16)[lane] = v;
        if (lane < 25) reinterpret_cast<uint2*>(side + 98 * n_chunks + c * 50)[lane] = make_uint2(v.x, v.y);
    }
    if ((FLAGS & 16) && lane == 0) { side[c] = tag; side[n_chunks + c] = tag; }                                   // count/status scalars
    if ((FLAGS & 32) && lane == 0) { side[2 * n_chunks + c * 96 + 11] = tag; side[2 * n_chunks + c * 96 + 14] = tag; }   // record header words
    if ((FLAGS & 64) && lane < 20) reinterpret_cast<uint4*>(side + 2 * n_chunks + c * 96 + 16)[lane] = v;          // history (full sectors)
    if ((FLAGS & 128) && lane < 25) reinterpret_cast<uint2*>(side + 98 * n_chunks + c * 50)[lane] = make_uint2(v.x, v.y);   // legal mask
    if ((FLAGS & 256) && lane < 24) reinterpret_cast<uint4*>(side + 2 * n_chunks + c * 96)[lane] = v;              // whole 384-B record
}

template <class F>
static double time_it(F launch, size_t bytes, int reps = 20) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return (double)bytes * reps / (ms * 1e-3) / 1e9;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t sizes[2] = {(size_t)20480 * 16128, (size_t)1 << 30};
    uint8_t* buf; CK(cudaMalloc(&buf, sizes[1]));
    for (int si = 0; si < 2; ++si) {
        const size_t bytes = sizes[si], n_vec = bytes / 16;
        printf("== buffer %.1f MB (%s)\n", bytes / 1e6, si == 0 ? "the planes arena at 20480 games" : "1 GiB");
        printf("memset                        %7.0f GB/s\n", time_it([&] { cudaMemsetAsync(buf, 1, bytes); }, bytes));
#define RUN(MODE, U, BLK, CTAS)                                                                                        \
    printf("%-8s unroll %d blk %4d ctas/sm %2d  %7.0f GB/s\n", #MODE, U, BLK, CTAS,                                     \
           time_it([&] { store_kernel<MODE, U><<<sms * CTAS, BLK>>>((uint4*)buf, n_vec, 7); }, bytes))
        RUN(PLAIN, 1, 256, 8); RUN(PLAIN, 4, 256, 8); RUN(PLAIN, 8, 256, 4); RUN(PLAIN, 4, 1024, 2); RUN(PLAIN, 1, 256, 64);
        RUN(CS, 4, 256, 8); RUN(WT, 4, 256, 8); RUN(CG, 4, 256, 8); RUN(NOALLOC, 4, 256, 8);
        printf("full grid, one uint4/thread   %7.0f GB/s\n",
               time_it([&] { store_kernel<PLAIN, 1><<<(unsigned)((n_vec + 255) / 256), 256>>>((uint4*)buf, n_vec, 7); }, bytes));
        printf("warp-chunk 16128 B            %7.0f GB/s\n",
               time_it([&] { store_chunk_kernel<16128><<<sms * 8, 256>>>((uint4*)buf, bytes / 16128, 7); }, bytes / 16128 * 16128));
        printf("warp-chunk 16128 B, 16 w/cta  %7.0f GB/s\n",
               time_it([&] { store_chunk_kernel<16128><<<sms * 4, 512>>>((uint4*)buf, bytes / 16128, 7); }, bytes / 16128 * 16128));
#define BULK(CH, D, CTAS)                                                                                              \
    printf("bulk %6d B depth %d ctas/sm %d   %7.0f GB/s\n", CH, D, CTAS,                                                \
           time_it([&] { store_bulk_kernel<CH, D><<<sms * CTAS, 128, CH>>>(buf, bytes / CH, 7); }, bytes / CH * CH))
        {
            const size_t nc = bytes / 16128;
            uint32_t* side; CK(cudaMalloc(&side, nc * 160 * 4));
            uint4* src; CK(cudaMalloc(&src, nc * 64 * 16)); CK(cudaMemset(src, 0, nc * 64 * 16));
#define LIKE(F)                                                                                                        \
    printf("encode-like flags %3d            %7.0f GB/s\n", F,                                                          \
           time_it([&] { store_like_encode<F><<<(unsigned)((nc + 7) / 8), 256>>>((uint4*)buf, nc, 7, side, src); }, nc * 16128))
            LIKE(0); LIKE(3); LIKE(4); LIKE(16); LIKE(32); LIKE(64); LIKE(128); LIKE(256); LIKE(8); LIKE(15); LIKE(0);
            CK(cudaFree(side)); CK(cudaFree(src));
        }
        BULK(16128, 4, 4); BULK(16128, 8, 8); BULK(4096, 8, 8); BULK(32256, 4, 4); BULK(2048, 16, 16);
    }
    return 0;
}
