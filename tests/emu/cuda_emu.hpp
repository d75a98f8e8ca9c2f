// tests/emu/cuda_emu.hpp -- TEST INFRASTRUCTURE ONLY: the sliver of the CUDA device programming model
// that csrc/nbody_kernels.cuh uses, for the host compiler.  One OS thread plays one CUDA thread, the blocks
// of a grid run one after the other (so function-local `static` stands in for `__shared__`), and
// __syncthreads() is a pthread barrier over the block.  It exists so that the CPU test-suite can run the
// REAL kernel source -- indexing, staging, ragged tails, segment lists, arrival counters, fused combine +
// integrate -- and compare it bit for bit with the oracle's MIRRORED flavour.  What it cannot show is
// anything the hardware adds: MUFU.RSQ's rounding (emulated as the correctly rounded 1/sqrt the MIRRORED
// flavour uses), warp scheduling, memory-model races between concurrently resident blocks, PDL, and the real
// asynchrony of TMA (the bulk copy is a memcpy here; what IS checked is its addresses, sizes and phase logic).
// Nothing outside tests/ includes this file; the product build never defines MAPC_HOST_EMULATION.
#pragma once

#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <functional>

// ---- qualifiers ----------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types ----------------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- built-in variables: per emulated thread ---------------------------------------------------------------
namespace cuda_emu {
struct Block {
    pthread_barrier_t barrier;            // __syncthreads()
    pthread_barrier_t warp_barrier[32];   // one per warp of the block, for the shuffle emulation
    float xchg[32][32];                   // [warp][lane]: values being shuffled
};
extern thread_local uint3 t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
extern thread_local Block *t_block;
}  // namespace cuda_emu
#define threadIdx (cuda_emu::t_threadIdx)
#define blockIdx (cuda_emu::t_blockIdx)
#define blockDim (cuda_emu::t_blockDim)
#define gridDim (cuda_emu::t_gridDim)

static inline void __syncthreads() { pthread_barrier_wait(&cuda_emu::t_block->barrier); }

// ---- arithmetic intrinsics: IEEE round-to-nearest, no contraction (build with -ffp-contract=off) -------------
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }

// ---- memory and synchronisation -------------------------------------------------------------------------------
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T __ldcg(const T *p) { return *p; }
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
// Warp shuffle: the 32 OS threads of a warp meet at a per-warp barrier, publish their values, read the source
// lane's, and meet again before the slots are reused.  Every lane of the warp must execute the shuffle (true of
// the SHFL staging variant: a uniform loop over the 64-body tile).
static inline float __shfl_sync(unsigned, float v, int src_lane)
{
    cuda_emu::Block *b = cuda_emu::t_block;
    const unsigned w = cuda_emu::t_threadIdx.x / 32, lane = cuda_emu::t_threadIdx.x % 32;
    b->xchg[w][lane] = v;
    pthread_barrier_wait(&b->warp_barrier[w]);
    const float r = b->xchg[w][src_lane & 31];
    pthread_barrier_wait(&b->warp_barrier[w]);
    return r;
}

// ---- stand-ins for the PTX of nbody_kernels.cuh -----------------------------------------------------------------
namespace mapc {
// the MIRRORED oracle flavour's `1.0f / sqrtf(x)` (oracle/oracle.c): the GPU's MUFU.RSQ differs by <= 2 ulp
static inline float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
static inline unsigned long long global_timer_ns()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
static inline unsigned smem_u32(const void *) { return 0; }
// 1-D TMA staging: the mbarrier word counts completed phases.  The kernel arms a phase with one arrival
// (expect_tx by the issuing thread) and completes it with the bulk copy's bytes, so here the copy itself
// (a memcpy) completes the phase; a waiter for parity P spins until the current phase's parity differs from
// P.  Bytes must be what the hardware accepts: a non-zero multiple of 16, both addresses 16-byte aligned.
static inline void mbar_init(unsigned long long *bar, unsigned count)
{
    if (count != 1) abort();
    __atomic_store_n(bar, 0ull, __ATOMIC_RELEASE);
}
static inline void mbar_expect_tx(unsigned long long *, unsigned bytes)
{
    if (bytes == 0 || bytes % 16 != 0) abort();
}
static inline void mbar_wait(unsigned long long *bar, unsigned parity)
{
    while ((__atomic_load_n(bar, __ATOMIC_ACQUIRE) & 1ull) == (unsigned long long)(parity & 1u)) sched_yield();
}
static inline void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    if (bytes == 0 || bytes % 16 != 0 || ((uintptr_t)smem_dst & 15) || ((uintptr_t)gmem_src & 15)) abort();
    memcpy(smem_dst, gmem_src, bytes);
    __atomic_fetch_add(bar, 1ull, __ATOMIC_RELEASE);
}
static inline void fence_mbarrier_init() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned long long load_acquire_sys(const unsigned long long *p)
{
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
static inline unsigned load_acquire_gpu(const unsigned *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void store_release_gpu(unsigned *p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline void pdl_launch_dependents() {}
static inline void pdl_wait() {}
}  // namespace mapc

// ---- launch: blocks in grid order (x fastest), one OS thread per CUDA thread ------------------------------------
namespace cuda_emu {
struct ThreadCtx {
    const std::function<void()> *body;
    unsigned tid;
    dim3 bdim, gdim;
    int order;
    Block *block;
};

// One OS thread plays CUDA thread `tid` of every block in turn; the barrier after each block keeps a fast
// thread from entering the next block (and its `static` shared memory) while a slow one is still in this one.
inline void *thread_main(void *p)
{
    auto *c = static_cast<ThreadCtx *>(p);
    const unsigned long long nblocks = (unsigned long long)c->gdim.x * c->gdim.y;
    t_threadIdx = uint3{c->tid, 0, 0};
    t_blockDim = c->bdim;
    t_gridDim = c->gdim;
    t_block = c->block;
    for (unsigned long long k = 0; k < nblocks; ++k) {
        const unsigned long long lin = c->order == 0 ? k : nblocks - 1 - k;
        t_blockIdx = uint3{(unsigned)(lin % c->gdim.x), (unsigned)(lin / c->gdim.x), 0};
        (*c->body)();
        pthread_barrier_wait(&c->block->barrier);
    }
    return nullptr;
}

// body = one CUDA thread's call of the kernel, e.g. [&] { kernel(args); }
// order: 0 = blocks in ascending linear index, 1 = descending (to show results do not depend on which
// block arrives last at a target block's counter)
inline void launch(dim3 grid, dim3 block, int order, const std::function<void()> &body)
{
    const unsigned nthreads = block.x;
    ThreadCtx *ctx = new ThreadCtx[nthreads];
    pthread_t *th = new pthread_t[nthreads];
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    Block blk;
    pthread_barrier_init(&blk.barrier, nullptr, nthreads);
    const unsigned nwarps = (nthreads + 31) / 32;
    for (unsigned w = 0; w < nwarps; ++w) {
        const unsigned lanes = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
        pthread_barrier_init(&blk.warp_barrier[w], nullptr, lanes);
    }
    for (unsigned t = 0; t < nthreads; ++t) {
        ctx[t] = ThreadCtx{&body, t, block, grid, order, &blk};
        if (pthread_create(&th[t], &attr, thread_main, &ctx[t]) != 0) abort();
    }
    for (unsigned t = 0; t < nthreads; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&blk.barrier);
    for (unsigned w = 0; w < nwarps; ++w) pthread_barrier_destroy(&blk.warp_barrier[w]);
    pthread_attr_destroy(&attr);
    delete[] th;
    delete[] ctx;
}
}  // namespace cuda_emu
