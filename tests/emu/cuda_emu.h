// cuda_emu.h -- TEST INFRASTRUCTURE: a lock-step SIMT emulator for one warp, so that the CUDA
// device code under hive-alphazero_b200/csrc (included VERBATIM) can be checked against the
// oracle on a machine without a GPU.
//
// Each of the 32 lanes is a ucontext coroutine.  A lane runs until it reaches a warp
// collective (__shfl_sync, __ballot_sync, __match_any_sync, __reduce_or_sync, __syncwarp ...),
// parks there, and the scheduler resumes the next lane -- in a fresh random order every round,
// so a missing __syncwarp between a shared-memory write and another lane's read shows up as a
// result mismatch.  When every live lane is parked the scheduler checks that they all sit at
// the SAME collective (divergent use of a full-mask collective is an error on the GPU too),
// computes the results and starts the next round.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#define HIVE_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 v = {x, y}; return v; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = {x, y, z, w}; return v; }
struct dim3emu { unsigned x, y, z; };

namespace emu {

enum Kind { K_NONE = 0, K_SHFL, K_SHFL_XOR, K_SHFL_UP, K_BALLOT, K_MATCH, K_REDOR, K_SYNC, K_BAR, K_DONE };

constexpr int MAX_LANES = 1024;
struct Block {
    ucontext_t sched;
    ucontext_t ctx[MAX_LANES];
    char* stack[MAX_LANES];
    int nlanes;
    int cur;
    int kind[MAX_LANES];          // collective the lane is parked at (K_NONE = runnable)
    uint64_t in[MAX_LANES], out[MAX_LANES];
    uint32_t arg[MAX_LANES];
    bool done[MAX_LANES];
    uint64_t rng;
    long collectives;
};
extern Block* W;
extern thread_local dim3emu g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

inline void park(int kind, uint64_t in, uint32_t arg) {
    Block* w = W;
    int l = w->cur;
    w->kind[l] = kind; w->in[l] = in; w->arg[l] = arg;
    swapcontext(&w->ctx[l], &w->sched);
}
inline uint64_t result() { return W->out[W->cur]; }

}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline void emu_check_mask(unsigned m) { if (m != 0xffffffffu) { fprintf(stderr, "emu: only full-mask collectives supported\n"); abort(); } }
template <typename T> static inline uint64_t emu_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T emu_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <typename T> static inline T __shfl_sync(unsigned m, T v, int src) { emu_check_mask(m); emu::park(emu::K_SHFL, emu_bits(v), (uint32_t)src & 31); return emu_unbits<T>(emu::result()); }
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { emu_check_mask(m); emu::park(emu::K_SHFL_XOR, emu_bits(v), (uint32_t)x); return emu_unbits<T>(emu::result()); }
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, int d) { emu_check_mask(m); emu::park(emu::K_SHFL_UP, emu_bits(v), (uint32_t)d); return emu_unbits<T>(emu::result()); }
static inline unsigned __ballot_sync(unsigned m, bool p) { emu_check_mask(m); emu::park(emu::K_BALLOT, p ? 1u : 0u, 0); return (unsigned)emu::result(); }
static inline unsigned __match_any_sync(unsigned m, int v) { emu_check_mask(m); emu::park(emu::K_MATCH, (uint32_t)v, 0); return (unsigned)emu::result(); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { emu_check_mask(m); emu::park(emu::K_REDOR, v, 0); return (unsigned)emu::result(); }
static inline void __syncwarp(unsigned m = 0xffffffffu) { emu_check_mask(m); emu::park(emu::K_SYNC, 0, 0); }
static inline void __syncthreads() { emu::park(emu::K_BAR, 0, 0); }

static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
#include <math.h>
static inline double cospi(double x) { return cos(M_PI * x); }

namespace emu {
// Run `fn(arg)` as one thread block of `nthreads` threads (multiple of 32), block index `block`.
typedef void (*LaneFn)(void*);
int run_block(LaneFn fn, void* arg, int block, int nthreads, uint64_t seed);
const char* last_error();
}  // namespace emu
