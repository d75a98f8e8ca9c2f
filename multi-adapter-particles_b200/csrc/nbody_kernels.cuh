// nbody_kernels.cuh -- sm_100a kernels of the n-body step.
//
// What they reproduce (paths relative to the reference checkout):
//   bodyBodyInteraction   Particles/nBodyGravityCS.hlsl:44-57   -> pair_interaction()
//   CSMain epilogue       Particles/nBodyGravityCS.hlsl:103-108 -> integrate_body()
//   CSMain as shipped     Particles/nBodyGravityCS.hlsl:86-109  -> well_step_kernel
//   constants             Particles/nBodyGravityCS.hlsl:37-38, Particles/defines.h:37
//
// Design (B200-first, not a translation of the HLSL):
//  * The pair math runs on packed fp32 pairs (FADD2 / FFMA2 / FMUL2, PTX *.f32x2): one thread
//    owns 2*P target bodies as P register pairs, so every FMA-pipe instruction advances two
//    interactions and issue slots stop being the limit; the j body is a scalar-broadcast operand
//    (SASS `R.F32`), so one LDS.128 per source body feeds 2*P interactions.
//  * MUFU.RSQ (rsqrt.approx.ftz) replaces 1/sqrt: fxc lowers `1.0f/sqrt(x)` to `rsq` as well.
//  * Canonical summation order (frozen): sources are cut into S = 32 segments
//    (mapc_plan_segments), a segment into chains of MAPC_CHAIN_SOURCES = 2,048 sources; a chain is one
//    sequential fp32 accumulation, the chain sums of a segment are folded left to right into the
//    segment's partial (in shared memory), and the 32 partials left to right into the acceleration.
//    A block owns one (target block, segment) cell and writes one partial per target, so the result
//    does not depend on which launch / stream / GPU evaluated a segment, nor on P or the block size.
//  * No tensor cores: the work is not a contraction (d^2 via a GEMM cancels catastrophically).
#pragma once

// MAPC_HOST_EMULATION (tests/emu/ only, never defined by the product build): the same kernel source
// compiled by g++ over a thread-per-CUDA-thread shim, so the CPU suite can check the kernels' indexing,
// staging, tails and fused combine bit for bit against the oracle without a GPU.  The PTX below then has
// host stand-ins (tests/emu/cuda_emu.hpp); nothing else in this file knows about it.
#ifdef MAPC_HOST_EMULATION
#include "cuda_emu.hpp"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/mapc.h"

namespace mapc {

constexpr int kTileBodies = 256;  // source bodies per shared-memory stage (4 reference tiles of 64)

struct SegList {
    int count;
    int ids[MAPC_MAX_SEGMENTS];
};

// j range [j0, j1) of canonical segment s: tile-aligned (64 bodies, Particles/defines.h:37),
// tiles = dimx of Particles/Compute.cpp:544.
__host__ __device__ inline void segment_range(int n_sources, int S, int s, int &j0, int &j1)
{
    const long long tiles = (n_sources + MAPC_BLOCK_SIZE - 1) / MAPC_BLOCK_SIZE;
    long long a = (tiles * s) / S * MAPC_BLOCK_SIZE;
    long long b = (tiles * (s + 1)) / S * MAPC_BLOCK_SIZE;
    if (a > n_sources) a = n_sources;
    if (b > n_sources) b = n_sources;
    j0 = (int)a;
    j1 = (int)b;
}

#ifndef MAPC_HOST_EMULATION
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif

// Two bodyBodyInteraction calls (nBodyGravityCS.hlsl:44-57) at once: lanes .x/.y are two
// target bodies, b is the source body.  Operation order per lane:
//   r = bj - bi                      (:46)   FADD2
//   distSqr = dot(r,r) + 25          (:48-49) 3 FFMA2, softening folded into the first
//   invDist = rsqrt(distSqr)         (:51)   MUFU.RSQ x2
//   invDistCube = (inv*inv)*inv      (:52)   2 FMUL2
//   s = mass * invDistCube [* 1]     (:54)   FMUL2 (particles == 1 is an exact no-op)   [MASS_IN_LOOP]
//   ai += r * s                      (:56)   3 FFMA2
// MASS_IN_LOOP = false (default): the uniform factor g_fParticleMass is left out of the pair and
// applied once to each segment partial instead (lever (i)): 11 instead of 12
// FMA-pipe lane-ops per interaction, 77-78 % instead of 72 % of the FP32 peak.  Same formula; each term
// loses one rounding, so results differ from the per-pair multiply by ~1 ulp per partial -- far inside
// the 1e-5 tolerance; the oracle's MIRRORED flavour does the same, LITERAL keeps the shader's order.
template <bool MASS_IN_LOOP>
__device__ __forceinline__ void pair_interaction(const float4 b, const float2 nxi, const float2 nyi,
                                                 const float2 nzi, float2 &ax, float2 &ay, float2 &az)
{
    const float2 dx = __fadd2_rn(make_float2(b.x, b.x), nxi);
    const float2 dy = __fadd2_rn(make_float2(b.y, b.y), nyi);
    const float2 dz = __fadd2_rn(make_float2(b.z, b.z), nzi);
    float2 d2 = __ffma2_rn(dx, dx, make_float2(MAPC_SOFTENING_SQUARED, MAPC_SOFTENING_SQUARED));
    d2 = __ffma2_rn(dy, dy, d2);
    d2 = __ffma2_rn(dz, dz, d2);
    float2 inv;
    inv.x = rsqrt_approx(d2.x);
    inv.y = rsqrt_approx(d2.y);
    const float2 inv2 = __fmul2_rn(inv, inv);
    const float2 inv3 = __fmul2_rn(inv2, inv);
    const float2 s = MASS_IN_LOOP ? __fmul2_rn(inv3, make_float2(MAPC_PARTICLE_MASS, MAPC_PARTICLE_MASS)) : inv3;
    ax = __ffma2_rn(dx, s, ax);
    ay = __ffma2_rn(dy, s, ay);
    az = __ffma2_rn(dz, s, az);
}

// The same arithmetic for all P pairs of a thread against one source body, written operation-major
// (every step looped over the pairs) so neighbouring instructions share operands: the broadcast
// source coordinate across the subtractions and the scale s across the three accumulations, which is
// what lets the register-reuse cache feed the 3-operand FFMA2s.  Rounding is untouched.
template <int P, bool MASS_IN_LOOP>
__device__ __forceinline__ void group_interaction(const float4 b, const float2 *nxi, const float2 *nyi,
                                                  const float2 *nzi, float2 *ax, float2 *ay, float2 *az)
{
    float2 dx[P], dy[P], dz[P], d2[P], s[P];
#pragma unroll
    for (int p = 0; p < P; ++p) dx[p] = __fadd2_rn(make_float2(b.x, b.x), nxi[p]);
#pragma unroll
    for (int p = 0; p < P; ++p) dy[p] = __fadd2_rn(make_float2(b.y, b.y), nyi[p]);
#pragma unroll
    for (int p = 0; p < P; ++p) dz[p] = __fadd2_rn(make_float2(b.z, b.z), nzi[p]);
#pragma unroll
    for (int p = 0; p < P; ++p)
        d2[p] = __ffma2_rn(dx[p], dx[p], make_float2(MAPC_SOFTENING_SQUARED, MAPC_SOFTENING_SQUARED));
#pragma unroll
    for (int p = 0; p < P; ++p) d2[p] = __ffma2_rn(dy[p], dy[p], d2[p]);
#pragma unroll
    for (int p = 0; p < P; ++p) d2[p] = __ffma2_rn(dz[p], dz[p], d2[p]);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        d2[p].x = rsqrt_approx(d2[p].x);
        d2[p].y = rsqrt_approx(d2[p].y);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) s[p] = __fmul2_rn(d2[p], d2[p]);
#pragma unroll
    for (int p = 0; p < P; ++p) s[p] = __fmul2_rn(s[p], d2[p]);
    if (MASS_IN_LOOP) {
#pragma unroll
        for (int p = 0; p < P; ++p) s[p] = __fmul2_rn(s[p], make_float2(MAPC_PARTICLE_MASS, MAPC_PARTICLE_MASS));
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        ax[p] = __ffma2_rn(dx[p], s[p], ax[p]);
        ay[p] = __ffma2_rn(dy[p], s[p], ay[p]);
        az[p] = __ffma2_rn(dz[p], s[p], az[p]);
    }
}

// nBodyGravityCS.hlsl:103-108 with the contractions pinned (the CPU oracle's MIRRORED flavour
// uses the same fmaf placement):
//   vel += accel*dt (:103) FFMA; vel *= damping (:104) FMUL; pos += vel*dt (:105) FFMA;
//   pos.w = length(accel) (:107); velocity stored as a float3, .w = 0 (:108).
__device__ __forceinline__ void integrate_body(const float4 pos_in, const float4 vel_in,
                                               const float ax, const float ay, const float az,
                                               const float dt, const float damping, float4 &pos_out,
                                               float4 &vel_out)
{
    float vx = __fmaf_rn(ax, dt, vel_in.x);
    float vy = __fmaf_rn(ay, dt, vel_in.y);
    float vz = __fmaf_rn(az, dt, vel_in.z);
    vx = __fmul_rn(vx, damping);
    vy = __fmul_rn(vy, damping);
    vz = __fmul_rn(vz, damping);
    pos_out.x = __fmaf_rn(vx, dt, pos_in.x);
    pos_out.y = __fmaf_rn(vy, dt, pos_in.y);
    pos_out.z = __fmaf_rn(vz, dt, pos_in.z);
    pos_out.w = __fsqrt_rn(__fmaf_rn(az, az, __fmaf_rn(ay, ay, __fmul_rn(ax, ax))));
    vel_out = make_float4(vx, vy, vz, 0.f);
}

// Everything one launch of the force kernel needs (passed by value).
struct StepArgs {
    const float4 *pos;         // packed positions of the read side, global indexing (targets and sources)
    // partials scratch: [slot][S][T*2P] float4, slot = target block % scratch_blocks.  scratch_blocks ==
    // n_iblocks: every target block has its own slot.  scratch_blocks < n_iblocks (unsharded fused steps):
    // a small ring that stays in L2 -- a cell may overwrite a slot only after the target block that used it
    // scratch_blocks blocks earlier has been combined (slot_gen), and cells are handed out through a ticket
    // counter in (target block, segment) order, so every cell a block can wait for is already running.
    float4 *partial;
    int scratch_blocks;
    int group_blocks;          // > 0 (ring only): tickets map to cells in groups of this many target blocks, segment major inside
    unsigned *ticket;          // next cell of this launch, or null: cell = blockIdx.x
    unsigned *slot_gen;        // [scratch_blocks] target blocks combined out of each slot this step, or null
    int i_first, i_cnt;        // local targets are bodies [i_first, i_first + i_cnt)
    int n_sources, S;          // canonical segmentation of the sources
    SegList segs;              // the segments this launch evaluates
    int n_iblocks;             // target blocks of T*2P bodies
    // fused combine + integrate (FUSE): the block that completes the last segment of a target block
    // sums its S partials in canonical order and applies the integration step
    unsigned *counters;        // [n_iblocks] arrivals, zero between steps
    const mapc_posvelo *in;    // PosVelo read side (local indexing)
    mapc_posvelo *out;         // PosVelo write side
    float4 *pos_next;          // packed positions of the write side (global indexing)
    float dt, damping;
    // PEER (collective-free multi-GPU): per segment of this launch (indexed like segs.ids) the packed
    // array of the rank that owns those sources -- a peer-mapped pointer read over NVLink -- and that
    // rank's step flag, which must reach flag_expect before its positions of this step may be read
    const float4 *seg_src[MAPC_MAX_SEGMENTS];
    const unsigned long long *seg_flag[MAPC_MAX_SEGMENTS];
    unsigned long long flag_expect;
    // "simulate ms" timer without stream operations (the reference brackets its Dispatch with in-queue
    // timestamp queries, D3D12GpuTimer.h:117-129): block (0,0) stamps %globaltimer when it starts, the
    // block that integrates the LAST target block of the step stamps it when it is done
    unsigned long long *stamp_begin;   // pinned host memory, or null
    unsigned long long *stamp_end;     // pinned host memory, or null
    unsigned *done;                    // device, never null for FUSE: [0] target blocks integrated so far this step,
                                       // [1] the cell ticket, [2] id of the last step whose host-visible completion
                                       // (stamp, fence word) has been written
    // Every in-kernel wait (a peer's step flag, a scratch-ring slot) is bounded: after wait_timeout_ns the
    // waiting block stores MAPC_ERR_TIMEOUT-style evidence to error_word (pinned host memory: [0] = 1 peer
    // flag / 2 ring slot, [1] = what it waited for) and carries on, so the grid always terminates and
    // WaitForGpu reports the step as failed instead of the GPU hanging.
    unsigned long long wait_timeout_ns;
    unsigned long long *error_word;
    // Step-to-step dataflow (batched small-N steps): block_step[ib] = id of the last step
    // whose target block ib has been integrated (published by every fused step when not null).  A launch with
    // wait_prev != 0 does NOT wait for the whole previous grid (griddepcontrol.wait): each cell waits only for
    // the target blocks of step step_id - 1 it reads -- its own targets and the bodies of its source segment --
    // so its math overlaps the drain of the previous step.  Requires the previous step to have had the same
    // launch shape and body counts (the host decides).
    unsigned *block_step;
    unsigned step_id;
    int wait_prev;
    // fence signal from inside the kernel (ID3D12CommandQueue::Signal after the Dispatch, Compute.cpp:999):
    // the same last block stores fence_value to the fence word (pinned host memory) -- or null
    unsigned long long *fence_word;
    unsigned long long fence_value;
};

#ifndef MAPC_HOST_EMULATION
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---- 1-D TMA (cp.async.bulk) staging helpers: shared::cluster destination, mbarrier completion ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ unsigned long long load_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned load_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void store_release_gpu(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// programmatic dependent launch (no-ops for a launch without the programmatic-serialization attribute)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fence_mbarrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif  // !MAPC_HOST_EMULATION

// Force kernel.  Work is cut into cells = (target block of T*2P bodies) x (canonical segment); one
// thread block evaluates one cell.  Cells are numbered target block major, segment fastest
// (cell = ib * segs.count + k), so the S cells of a target block run next to each other in time and the
// combine finds their partials in L2; the cell of a thread block is blockIdx.x, or -- a.ticket != null --
// the next ticket of an atomic counter (then a cell with a smaller number is guaranteed to be running or
// finished, which is what makes the scratch ring's wait deadlock-free without any assumption about the
// order in which the hardware dispatches blocks).  (A persistent variant that walked several cells per
// block was measured 25 % slower: with the targets reloaded inside a loop ptxas re-pairs them with ~170
// MOVs per 8 sources instead of keeping the register pairs live.)  A thread owns 2P targets as P register
// pairs and streams the segment's sources through shared memory in stages of TJ bodies (register prefetch
// of the next stage, one barrier per stage); the math runs in 64-body tiles -- the reference tile
// (Particles/defines.h:37) -- so only the globally last tile is ever ragged.
// CHAIN: sources per sequential accumulation chain (MAPC_CHAIN_SOURCES in the product; the emulation tests
// also instantiate 256 so that small problems have several chains per segment).  At the end of a chain the
// chain sum -- scaled by the mass like the oracle's MIRRORED flavour -- is folded into the segment's partial,
// which lives in shared memory (thread-private slots, no barrier): partial = c0, then partial += c1, ...
// U = unroll of the source loop, MINB = resident blocks per SM asked of ptxas.  None of P, T, TJ, U,
// MINB or ORDER changes a rounding: each target's chains are the same ops in ascending j.
// PEER: the cell's sources are read straight from the owning GPU's memory (no all-gather): the block
// first waits until the owner's step flag says its positions of this step are published.  Waiting
// cannot deadlock: a peer's step k-1 never depends on this GPU's step k.
// TMA: the source stages are filled by 1-D bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) issued
// by one thread instead of LDG/STS by all; measured A/B in profiles/ -- it changes nothing, because
// staging is ~0.02 % of the instruction stream either way.
// SHFL: the A/B for "broadcast the tile with warp shuffles": each lane reads ONE body of a 32-body
// group from shared memory and the group is handed round with SHFL.IDX (3 per source) instead of one
// uniform-address LDS.128 per source.  Same bodies in the same order, so bit-identical; measured
// slower (profiles/), because the shared-memory broadcast read is already a single instruction per
// source and the shuffles triple the non-FMA issue slots.  Off by default (MAPC_SHFL=1).
// Instructions of padding in front of the hot loop of one instantiation (see the comment at its use).
// MAPC_LOOP_PAD=k (compile time) adds k to every instantiation: the alignment sweep of tools/sassprobe/pad_sweep.sh.
template <int P, int T, int TJ, int U, int ORDER, bool FUSE, bool PEER, bool TMA, bool MASS_IN_LOOP, bool SHFL>
__host__ __device__ constexpr int hot_loop_pad()
{
    int pad = 0;
#define MAPC_HOT_LOOP_PAD(P_, T_, TJ_, U_, ORDER_, FUSE_, PEER_, TMA_, MASS_, SHFL_, PAD_)                          \
    if (P == P_ && T == T_ && TJ == TJ_ && U == U_ && ORDER == ORDER_ && FUSE == FUSE_ && PEER == PEER_ &&        \
        TMA == TMA_ && MASS_IN_LOOP == MASS_ && SHFL == SHFL_)                                                    \
        pad = PAD_;
#include "hot_loop_pad.inc"
#undef MAPC_HOT_LOOP_PAD
#ifdef MAPC_LOOP_PAD
    pad += MAPC_LOOP_PAD;
#endif
    return pad;
}

template <int P, int T, int TJ, int U, int MINB, int ORDER, bool FUSE, bool PEER = false, bool TMA = false,
          bool MASS_IN_LOOP = false, bool SHFL = false, int CHAIN = MAPC_CHAIN_SOURCES>
__global__ void __launch_bounds__(T, MINB) force_cells_kernel(const __grid_constant__ StepArgs a)
{
    constexpr int kLoads = TJ / T;  // staging loads per thread per stage
    static_assert(TJ % T == 0 && TJ % MAPC_BLOCK_SIZE == 0, "stage must be a multiple of block and tile size");
    static_assert(CHAIN > 0 && CHAIN % TJ == 0, "a chain is a whole number of stages");
    constexpr int kStagesPerChain = CHAIN / TJ;
    constexpr int kBlockTargets = T * 2 * P;
    __shared__ __align__(128) float4 tile[2][TJ];
    __shared__ float seg_acc[6 * P][T];   // the segment's partial while its chains are folded in: [3*q + axis][tid]
    __shared__ __align__(8) unsigned long long full_bar[2];
    __shared__ int s_is_last;
    __shared__ unsigned s_cell;

    const int tid = threadIdx.x;
    if (TMA) {
        if (tid == 0) {
            mbar_init(&full_bar[0], 1);
            mbar_init(&full_bar[1], 1);
            fence_mbarrier_init();
        }
        __syncthreads();
    }
    const float4 *__restrict__ pos = a.pos;
    // Programmatic dependent launch (batched steps): let the next step's grid start filling SMs as this
    // one drains, and do not touch the previous step's output before that grid has completely finished.
    // Both are no-ops for a launch without the programmatic-serialization attribute.
    pdl_launch_dependents();
    if (!a.wait_prev) pdl_wait();
    unsigned cell = blockIdx.x;
    if (a.ticket != nullptr) {
        if (tid == 0) {
            // Tickets are handed out in an order of their own (a.group_blocks > 0): groups of `group_blocks` target
            // blocks, and inside a group SEGMENT major -- cells that run side by side then read the SAME source segment,
            // so at sizes whose positions exceed the L2 (N = 4 M: 67 MB) a segment is fetched from HBM once per group
            // instead of once per target block.  Any bijection is legal: a partial depends on (target, segment) only,
            // and a ticket still precedes every ticket it can wait for (the target block two groups earlier).
            unsigned t = atomicAdd(a.ticket, 1u);
            if (a.group_blocks > 0) {
                const unsigned count = (unsigned)a.segs.count, per_group = (unsigned)a.group_blocks * count;
                const unsigned g = t / per_group, r = t - g * per_group;
                const unsigned first = g * (unsigned)a.group_blocks;
                const unsigned left = (unsigned)a.n_iblocks - first;
                const unsigned nb = left < (unsigned)a.group_blocks ? left : (unsigned)a.group_blocks;   // a short last group
                const unsigned kseg = r / nb;
                t = (first + (r - kseg * nb)) * count + kseg;
            }
            s_cell = t;
        }
        __syncthreads();
        cell = s_cell;
    }
    const int ib = (int)(cell / (unsigned)a.segs.count);        // target block
    const int kseg = (int)(cell - (unsigned)ib * (unsigned)a.segs.count);   // index into this launch's segment list
    const float4 *__restrict__ src = PEER ? a.seg_src[kseg] : a.pos;   // where this cell's sources live
    if (FUSE && a.stamp_begin != nullptr && cell == 0 && tid == 0) *a.stamp_begin = global_timer_ns();
    if (PEER) {
        const unsigned long long *flag = a.seg_flag[kseg];
        if (flag != nullptr) {
            if (tid == 0) {
                const unsigned long long t0 = global_timer_ns();
                while (load_acquire_sys(flag) < a.flag_expect) {
                    __nanosleep(200);
                    if (a.error_word != nullptr && global_timer_ns() - t0 > a.wait_timeout_ns) {
                        a.error_word[1] = a.flag_expect;
                        a.error_word[0] = 1ull;     // a peer never published this step
                        __threadfence_system();
                        break;
                    }
                }
            }
            __syncthreads();
        }
    }
    {
        const int seg = a.segs.ids[kseg];
        int j0, j1;
        segment_range(a.n_sources, a.S, seg, j0, j1);
        if (a.wait_prev) {
            // dataflow instead of a grid-wide wait: the previous step must have integrated this cell's own
            // target block and the target blocks that hold the bodies of its source segment.  One thread per
            // awaited block polls (the L2 round trips overlap), the barrier below joins them.
            {
                const unsigned want = a.step_id - 1u;
                int first = j1 > j0 ? j0 / kBlockTargets : ib, last = j1 > j0 ? (j1 - 1) / kBlockTargets : ib;
                if (last >= a.n_iblocks) last = a.n_iblocks - 1;   // sources past the dispatched targets are never rewritten
                // awaited blocks: index 0 = the own target block, 1.. = first..last
                for (int w = tid; w <= last - first + 1; w += T) {
                    const int blk = w == 0 ? ib : first + w - 1;
                    if (blk >= a.n_iblocks) continue;
                    const unsigned long long t0 = global_timer_ns();
                    while ((int)(load_acquire_gpu(a.block_step + blk) - want) < 0) {
                        __nanosleep(64);
                        if (a.error_word != nullptr && global_timer_ns() - t0 > a.wait_timeout_ns) {
                            a.error_word[1] = (unsigned long long)blk;
                            a.error_word[0] = 3ull;     // a target block of the previous step never completed
                            __threadfence_system();
                            break;
                        }
                    }
                }
            }
            __syncthreads();
        }

        // targets: thread owns local bodies i_block + q*T + tid, q = 0..2P-1 (coalesced in q);
        // pair p = (q = 2p, q = 2p+1).  Out-of-range lanes are clamped and never stored.
        const int i_block = ib * kBlockTargets;
        float2 nxi[P], nyi[P], nzi[P], ax[P], ay[P], az[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            int ia = i_block + (2 * p) * T + tid;
            int ib2 = i_block + (2 * p + 1) * T + tid;
            ia = ia < a.i_cnt ? ia : a.i_cnt - 1;
            ib2 = ib2 < a.i_cnt ? ib2 : a.i_cnt - 1;
            const float4 ta = __ldcg(pos + a.i_first + ia);    // L2: may have been written by a grid still running
            const float4 tb = __ldcg(pos + a.i_first + ib2);
            nxi[p] = make_float2(-ta.x, -tb.x);
            nyi[p] = make_float2(-ta.y, -tb.y);
            nzi[p] = make_float2(-ta.z, -tb.z);
            ax[p] = ay[p] = az[p] = make_float2(0.f, 0.f);
        }

        const int n_stages = (j1 - j0 + TJ - 1) / TJ;
        int folded = 0;  // chains of this segment already folded into seg_acc
        float4 stage[kLoads];
        if (TMA) {
            if (n_stages > 0 && tid == 0) {
                const unsigned bytes = (unsigned)(((j1 - j0) < TJ ? (j1 - j0) : TJ) * sizeof(float4));
                mbar_expect_tx(&full_bar[0], bytes);
                tma_load_1d(&tile[0][0], src + j0, bytes, &full_bar[0]);
            }
        } else {
            if (n_stages > 0) {  // prologue: stage 0 -> smem buffer 0
#pragma unroll
                for (int l = 0; l < kLoads; ++l) {
                    const int j = j0 + l * T + tid;
                    stage[l] = j < j1 ? __ldcg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int l = 0; l < kLoads; ++l) tile[0][l * T + tid] = stage[l];
            }
            __syncthreads();
        }

#ifndef MAPC_HOST_EMULATION
        // Code alignment of the hot loop (profiles/r02_loop_alignment.txt): the SAME 31 instructions of the (2,128)
        // kernel run 1.3 % faster when the loop's first instruction sits in the last 16-byte slot of a 128-byte
        // instruction line than at any of the other seven positions (period 128 bytes; register names do not matter).
        // Neither CUDA C++ nor PTX can align a label, so the code in front of the loop is padded by whole
        // instructions: hot_loop_pad() holds the count per instantiation for this toolchain, written by
        // tools/align_hot_loops.py and guarded by tests/test_sass_hot_loop.py (an edit that moves the loop fails there).
        // `pmevent` is the cheapest instruction ptxas neither drops nor moves; it runs once per cell.
#pragma unroll
        for (int k = 0; k < hot_loop_pad<P, T, TJ, U, ORDER, FUSE, PEER, TMA, MASS_IN_LOOP, SHFL>(); ++k)
            asm volatile("pmevent 1;");
#endif
        for (int t = 0; t < n_stages; ++t) {
            const int buf = t & 1;
            const int jt = j0 + t * TJ;
            const bool has_next = (t + 1) < n_stages;
            if (TMA) {
                // the other buffer was released by the barrier that ended the previous iteration
                if (has_next && tid == 0) {
                    const int rest = j1 - (jt + TJ);
                    const unsigned bytes = (unsigned)((rest < TJ ? rest : TJ) * sizeof(float4));
                    mbar_expect_tx(&full_bar[buf ^ 1], bytes);
                    tma_load_1d(&tile[buf ^ 1][0], src + jt + TJ, bytes, &full_bar[buf ^ 1]);
                }
                mbar_wait(&full_bar[buf], (unsigned)((t >> 1) & 1));
            } else if (has_next) {  // prefetch the next stage into registers while this one is consumed
#pragma unroll
                for (int l = 0; l < kLoads; ++l) {
                    const int j = jt + TJ + l * T + tid;
                    stage[l] = j < j1 ? __ldcg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            const int cnt = (j1 - jt) < TJ ? (j1 - jt) : TJ;
            const int full = cnt & ~(MAPC_BLOCK_SIZE - 1);  // whole 64-body tiles of this stage
            for (int jb = 0; jb < full; jb += MAPC_BLOCK_SIZE) {
                constexpr int kU = U;
                if (SHFL) {
                    for (int h = 0; h < MAPC_BLOCK_SIZE; h += 32) {
                        const float4 mine = tile[buf][jb + h + (tid & 31)];
#pragma unroll kU
                        for (int j = 0; j < 32; ++j) {
                            float4 b;
                            b.x = __shfl_sync(0xffffffffu, mine.x, j);
                            b.y = __shfl_sync(0xffffffffu, mine.y, j);
                            b.z = __shfl_sync(0xffffffffu, mine.z, j);
                            b.w = 0.f;
                            if (ORDER == 2) {
                                group_interaction<P, MASS_IN_LOOP>(b, nxi, nyi, nzi, ax, ay, az);
                            } else {
#pragma unroll
                                for (int p = 0; p < P; ++p)
                                    pair_interaction<MASS_IN_LOOP>(b, nxi[p], nyi[p], nzi[p], ax[p], ay[p], az[p]);
                            }
                        }
                    }
                } else {
#pragma unroll kU
                    for (int j = 0; j < MAPC_BLOCK_SIZE; ++j) {
                        const float4 b = tile[buf][jb + j];
                        if (ORDER == 2) {
                            group_interaction<P, MASS_IN_LOOP>(b, nxi, nyi, nzi, ax, ay, az);
                        } else {
#pragma unroll
                            for (int p = 0; p < P; ++p)
                                pair_interaction<MASS_IN_LOOP>(b, nxi[p], nyi[p], nzi[p], ax[p], ay[p], az[p]);
                        }
                    }
                }
            }
            // ragged tail (only the last tile of all sources): bounded at the real count, no phantom bodies
#pragma unroll 1
            for (int j = full; j < cnt; ++j) {
                const float4 b = tile[buf][j];
#pragma unroll
                for (int p = 0; p < P; ++p)
                    pair_interaction<MASS_IN_LOOP>(b, nxi[p], nyi[p], nzi[p], ax[p], ay[p], az[p]);
            }
            if (!TMA && has_next) {
#pragma unroll
                for (int l = 0; l < kLoads; ++l) tile[buf ^ 1][l * T + tid] = stage[l];
            }
            __syncthreads();
            if (has_next && (t + 1) % kStagesPerChain == 0) {
                // end of a chain with more sources to come: fold the chain sum into the segment's partial
                // (thread-private shared-memory slots) and start the next chain from zero
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (!MASS_IN_LOOP) {
                        const float2 m = make_float2(MAPC_PARTICLE_MASS, MAPC_PARTICLE_MASS);
                        ax[p] = __fmul2_rn(ax[p], m);
                        ay[p] = __fmul2_rn(ay[p], m);
                        az[p] = __fmul2_rn(az[p], m);
                    }
                    if (folded == 0) {
                        seg_acc[6 * p + 0][tid] = ax[p].x;
                        seg_acc[6 * p + 1][tid] = ay[p].x;
                        seg_acc[6 * p + 2][tid] = az[p].x;
                        seg_acc[6 * p + 3][tid] = ax[p].y;
                        seg_acc[6 * p + 4][tid] = ay[p].y;
                        seg_acc[6 * p + 5][tid] = az[p].y;
                    } else {
                        seg_acc[6 * p + 0][tid] = __fadd_rn(seg_acc[6 * p + 0][tid], ax[p].x);
                        seg_acc[6 * p + 1][tid] = __fadd_rn(seg_acc[6 * p + 1][tid], ay[p].x);
                        seg_acc[6 * p + 2][tid] = __fadd_rn(seg_acc[6 * p + 2][tid], az[p].x);
                        seg_acc[6 * p + 3][tid] = __fadd_rn(seg_acc[6 * p + 3][tid], ax[p].y);
                        seg_acc[6 * p + 4][tid] = __fadd_rn(seg_acc[6 * p + 4][tid], ay[p].y);
                        seg_acc[6 * p + 5][tid] = __fadd_rn(seg_acc[6 * p + 5][tid], az[p].y);
                    }
                    ax[p] = ay[p] = az[p] = make_float2(0.f, 0.f);
                }
                ++folded;
            }
        }

        // the slot of the scratch this target block's partials go to; a ring slot must have been released
        // by the target block that used it scratch_blocks blocks earlier (its combine has read everything)
        const int slot = ib % a.scratch_blocks;
        if (a.slot_gen != nullptr) {
            if (tid == 0) {
                const unsigned want = (unsigned)(ib / a.scratch_blocks);
                const unsigned long long t0 = global_timer_ns();
                while (load_acquire_gpu(a.slot_gen + slot) != want) {
                    __nanosleep(100);
                    if (a.error_word != nullptr && global_timer_ns() - t0 > a.wait_timeout_ns) {
                        a.error_word[1] = (unsigned long long)ib;
                        a.error_word[0] = 2ull;     // a scratch-ring slot was never released
                        __threadfence_system();
                        break;
                    }
                }
            }
            __syncthreads();
        }
        float4 *out = a.partial + ((size_t)slot * a.S + seg) * kBlockTargets;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int ra = (2 * p) * T + tid, rb = (2 * p + 1) * T + tid;   // row inside the target block
            if (!MASS_IN_LOOP) {  // the uniform g_fParticleMass factor, once per chain sum
                const float2 m = make_float2(MAPC_PARTICLE_MASS, MAPC_PARTICLE_MASS);
                ax[p] = __fmul2_rn(ax[p], m);
                ay[p] = __fmul2_rn(ay[p], m);
                az[p] = __fmul2_rn(az[p], m);
            }
            if (folded > 0) {  // the last chain joins the earlier ones
                ax[p] = make_float2(__fadd_rn(seg_acc[6 * p + 0][tid], ax[p].x), __fadd_rn(seg_acc[6 * p + 3][tid], ax[p].y));
                ay[p] = make_float2(__fadd_rn(seg_acc[6 * p + 1][tid], ay[p].x), __fadd_rn(seg_acc[6 * p + 4][tid], ay[p].y));
                az[p] = make_float2(__fadd_rn(seg_acc[6 * p + 2][tid], az[p].x), __fadd_rn(seg_acc[6 * p + 5][tid], az[p].y));
            }
            if (i_block + ra < a.i_cnt) out[ra] = make_float4(ax[p].x, ay[p].x, az[p].x, 0.f);
            if (i_block + rb < a.i_cnt) out[rb] = make_float4(ax[p].y, ay[p].y, az[p].y, 0.f);
        }

        if (FUSE) {
            // last arrival for this target block combines and integrates (fixed order: any block may do it)
            __threadfence();
            __syncthreads();
            if (tid == 0) s_is_last = (atomicAdd(&a.counters[ib], 1u) + 1u == (unsigned)a.S);
            __syncthreads();
            if (s_is_last) {
                __threadfence();
                const float4 *part = a.partial + (size_t)slot * a.S * kBlockTargets;
#pragma unroll
                for (int q = 0; q < 2 * P; ++q) {
                    const int row = q * T + tid;
                    const int i = i_block + row;
                    if (i < a.i_cnt) {
                        // canonical left-to-right fold of the S partials; the loads of a batch are issued together
                        // (they are L2 round trips: one at a time they would be the longest serial piece of a
                        // small-N step), the additions stay in segment order
                        float sx = 0.f, sy = 0.f, sz = 0.f;
                        constexpr int kBatch = 8;
                        for (int s0 = 0; s0 < a.S; s0 += kBatch) {
                            float4 pp[kBatch];
#pragma unroll
                            for (int k = 0; k < kBatch; ++k)
                                if (s0 + k < a.S) pp[k] = __ldcg(part + (size_t)(s0 + k) * kBlockTargets + row);
#pragma unroll
                            for (int k = 0; k < kBatch; ++k)
                                if (s0 + k < a.S) {
                                    sx = __fadd_rn(sx, pp[k].x);
                                    sy = __fadd_rn(sy, pp[k].y);
                                    sz = __fadd_rn(sz, pp[k].z);
                                }
                        }
                        const float4 *srcpv = reinterpret_cast<const float4 *>(a.in + i);
                        float4 pos_out, vel_out;
                        integrate_body(__ldcg(srcpv), __ldcg(srcpv + 1), sx, sy, sz, a.dt, a.damping, pos_out, vel_out);
                        float4 *dst = reinterpret_cast<float4 *>(a.out + i);
                        dst[0] = pos_out;
                        dst[1] = vel_out;
                        a.pos_next[a.i_first + i] = pos_out;
                    }
                }
                // Every thread's integrate stores are ordered before the counters below by its own fence and
                // the barrier: whoever sees `done` complete (and then the fence word) sees the whole step.
                __threadfence();
                __syncthreads();
                if (tid == 0) {
                    a.counters[ib] = 0u;  // ready for the next step
                    if (a.slot_gen != nullptr) store_release_gpu(a.slot_gen + slot, (unsigned)(ib / a.scratch_blocks) + 1u);
                    const bool step_done = atomicAdd(a.done, 1u) + 1u == (unsigned)a.n_iblocks;
                    if (step_done) {
                        // the step is complete: this was its last target block.  Re-arm the step's counters.
                        __threadfence();
                        *a.done = 0u;
                        if (a.ticket != nullptr) *a.ticket = 0u;
                        if (a.slot_gen != nullptr)
                            for (int s = 0; s < a.scratch_blocks; ++s) a.slot_gen[s] = 0u;
                        __threadfence();
                    }
                    // this target block of this step is in place: published after the step's counters were
                    // re-armed (a dependent cell of the NEXT step can never touch them early) and before the
                    // system-scope stores below (they cross PCIe; the next step need not wait for them)
                    if (a.block_step != nullptr) store_release_gpu(a.block_step + ib, a.step_id);
                    if (step_done) {
                        // Host-visible completion (timer stamp, fence value).  Chained steps finish in order,
                        // but their last blocks are different threads: the fence word must never be seen going
                        // backwards, so a chained step writes it only after the previous step has (done[2]).
                        if (a.block_step != nullptr && a.wait_prev) {
                            const unsigned long long t0 = global_timer_ns();
                            while ((int)(load_acquire_gpu(a.done + 2) - (a.step_id - 1u)) < 0) {
                                __nanosleep(32);
                                if (a.error_word != nullptr && global_timer_ns() - t0 > a.wait_timeout_ns) break;
                            }
                        }
                        if (a.stamp_end != nullptr) {
                            *a.stamp_end = global_timer_ns();
                            __threadfence_system();      // the stamp is visible before the fence value that announces it
                        }
                        if (a.fence_word != nullptr)
                            *reinterpret_cast<volatile unsigned long long *>(a.fence_word) = a.fence_value;
                        __threadfence_system();
                        if (a.block_step != nullptr) store_release_gpu(a.done + 2, a.step_id);
                    }
                }
            }
        }
    }
}

// Sum the S segment partials left to right, integrate, write side b and the packed mirror.
//   in/out       local PosVelo sides (index 0 = body i_first)
//   pos_next     packed float4 positions of the side being written, global indexing
//   partial      [target block][S][block_targets] (every target block its own slot: the unfused path has no ring)
__global__ void __launch_bounds__(256)
integrate_kernel(const mapc_posvelo *__restrict__ in, mapc_posvelo *__restrict__ out,
                 float4 *__restrict__ pos_next, const float4 *__restrict__ partial,
                 int block_targets, int S, int i_first, int n_targets, float dt, float damping)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_targets) return;
    const int ib = i / block_targets, row = i - ib * block_targets;
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int s = 0; s < S; ++s) {
        const float4 p = partial[((size_t)ib * S + s) * block_targets + row];
        ax = __fadd_rn(ax, p.x);
        ay = __fadd_rn(ay, p.y);
        az = __fadd_rn(az, p.z);
    }
    const float4 *src = reinterpret_cast<const float4 *>(in + i);
    const float4 pos_in = src[0];
    const float4 vel_in = src[1];
    float4 pos_out, vel_out;
    integrate_body(pos_in, vel_in, ax, ay, az, dt, damping, pos_out, vel_out);
    float4 *dst = reinterpret_cast<float4 *>(out + i);
    dst[0] = pos_out;
    dst[1] = vel_out;
    pos_next[i_first + i] = pos_out;
}

// The step the reference actually dispatches (nBodyGravityCS.hlsl:86-109): one gravity well at
// the origin, note invDist = -1/sqrt (:97).  ~25 flop against 80 B of traffic per body (32 B PosVelo in,
// 32 B out, 16 B packed mirror): HBM-bound, so the shape is the one of a streaming copy -- B bodies per
// thread with all their loads issued before the first use (2B independent 16-byte loads in flight per
// thread), consecutive threads on consecutive bodies, streaming (evict-first) loads and stores because
// nothing is read twice inside a step.  kWellBodies bodies per thread, kWellThreads threads per block.
constexpr int kWellBodies = 2;
constexpr int kWellThreads = 256;

#ifdef MAPC_HOST_EMULATION
static inline float4 ld_stream(const float4 *p) { return *p; }
static inline void st_stream(float4 *p, float4 v) { *p = v; }
#else
__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4 *p, float4 v) { __stcs(p, v); }
#endif

__global__ void __launch_bounds__(kWellThreads)
well_step_kernel(const mapc_posvelo *__restrict__ in, mapc_posvelo *__restrict__ out,
                 float4 *__restrict__ pos_next, int i_first, int n_targets, float dt, float damping)
{
    const int base = blockIdx.x * (kWellThreads * kWellBodies) + threadIdx.x;
    float4 pos_in[kWellBodies], vel_in[kWellBodies];
#pragma unroll
    for (int k = 0; k < kWellBodies; ++k) {
        const int i = base + k * kWellThreads;
        if (i < n_targets) {
            const float4 *src = reinterpret_cast<const float4 *>(in + i);
            pos_in[k] = ld_stream(src);
            vel_in[k] = ld_stream(src + 1);
        }
    }
#pragma unroll
    for (int k = 0; k < kWellBodies; ++k) {
        const int i = base + k * kWellThreads;
        if (i >= n_targets) continue;
        const float4 p = pos_in[k];
        float d2 = __fmaf_rn(p.x, p.x, MAPC_SOFTENING_SQUARED);
        d2 = __fmaf_rn(p.y, p.y, d2);
        d2 = __fmaf_rn(p.z, p.z, d2);
        const float inv = -rsqrt_approx(d2);
        const float inv3 = __fmul_rn(__fmul_rn(inv, inv), inv);
        const float s = __fmul_rn(inv3, MAPC_PARTICLE_MASS);
        const float ax = __fmul_rn(p.x, s);
        const float ay = __fmul_rn(p.y, s);
        const float az = __fmul_rn(p.z, s);
        float4 pos_out, vel_out;
        integrate_body(p, vel_in[k], ax, ay, az, dt, damping, pos_out, vel_out);
        float4 *dst = reinterpret_cast<float4 *>(out + i);
        st_stream(dst, pos_out);
        st_stream(dst + 1, vel_out);
        st_stream(pos_next + i_first + i, pos_out);
    }
}

// packed float4 position mirror from a PosVelo array (after upload / state copy)
__global__ void __launch_bounds__(256)
pack_positions_kernel(const mapc_posvelo *__restrict__ in, float4 *__restrict__ pos, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = reinterpret_cast<const float4 *>(in + i)[0];
}

// ---- initial conditions: InitializeParticles / LoadParticles, USE_SCALAR_OPTIMIZED branch ---------------------
// Particles/Compute.cpp:596-609 fast_rand (g_seed = 214013*g_seed + 2531011; (g_seed >> 16) & 0x7FFF),
// :719-749 random walk until |delta|^2 >= 10, project on the shell of radius `spread` around `center`,
// velocity = cross(direction, perp) * speed, direction = normalize(position), perp = normalize((1,1,1) - direction);
// :831-844 two groups of N/2 around x = +-0.75 * ParticleSpread.
// The reference draws from an unseeded thread_local LCG inside a parallel_for, so its particles depend on the
// thread schedule and no run can be reproduced.  Here every particle owns an LCG stream seeded from (seed, i)
// -- one thread per particle, any order, same bits -- and the *Est normalisations (rsqrtps, ~12 bits) are
// exact; every operation is an IEEE round-to-nearest op with no contraction, so the CPU restatement
// (oracle/oracle.c mapo_init_particles) produces the same bytes.
__host__ __device__ inline unsigned ic_stream_seed(unsigned seed, unsigned i)
{
    unsigned s = seed + 0x9E3779B9u * (i + 1u);
    s ^= s >> 16;
    s *= 0x85EBCA6Bu;
    s ^= s >> 13;
    s *= 0xC2B2AE35u;
    s ^= s >> 16;
    return s;
}

__device__ __forceinline__ float ic_rand_pm1(unsigned &state)
{
    state = 214013u * state + 2531011u;                       // fast_rand(), Compute.cpp:605-609
    const int r = (int)((state >> 16) & 0x7FFFu);
    const float k_scale = (1.f / 32767.f) * 2.f;              // (1/RAND_MAX)*2, MSVC RAND_MAX (Compute.cpp:721)
    return __fadd_rn(__fmul_rn((float)r, k_scale), -1.f);
}

__device__ __forceinline__ float ic_dot3(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// Thread i generates body i of all n (both groups): its position goes to packed_a/packed_b[i] (every rank
// needs all positions as sources); bodies of the shard [i_first, i_first + n_local) also get their PosVelo
// on both ping-pong sides (Compute.cpp:881-882, :903-904 initialise both sides alike).
__global__ void __launch_bounds__(256)
init_particles_kernel(mapc_posvelo *__restrict__ side_a, mapc_posvelo *__restrict__ side_b,
                      float4 *__restrict__ packed_a, float4 *__restrict__ packed_b, unsigned n, unsigned i_first,
                      unsigned n_local, unsigned seed, float center_x, float speed, float spread)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned half = n / 2u;
    float4 pos = make_float4(0.f, 0.f, 0.f, 0.f), vel = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < 2u * half) {   // an odd N leaves its last body zero, like the value-initialised vector (Compute.cpp:825-829)
        const float cx = i < half ? center_x : -center_x;      // Compute.cpp:831-844
        unsigned state = ic_stream_seed(seed, i);
        float dx = ic_rand_pm1(state), dy = ic_rand_pm1(state), dz = ic_rand_pm1(state);
        while (ic_dot3(dx, dy, dz) < 10.f) {                    // Compute.cpp:728
            dx = __fadd_rn(dx, ic_rand_pm1(state));
            dy = __fadd_rn(dy, ic_rand_pm1(state));
            dz = __fadd_rn(dz, ic_rand_pm1(state));
        }
        const float len = __fsqrt_rn(ic_dot3(dx, dy, dz));     // XMVector3Normalize, then * spread (:738-739)
        const float px = __fadd_rn(cx, __fmul_rn(__fdiv_rn(dx, len), spread));
        const float py = __fmul_rn(__fdiv_rn(dy, len), spread);
        const float pz = __fmul_rn(__fdiv_rn(dz, len), spread);
        const float pl = __fsqrt_rn(ic_dot3(px, py, pz));
        const float ux = __fdiv_rn(px, pl), uy = __fdiv_rn(py, pl), uz = __fdiv_rn(pz, pl);   // direction (:746)
        float qx = __fadd_rn(1.f, -ux), qy = __fadd_rn(1.f, -uy), qz = __fadd_rn(1.f, -uz);   // (1,1,1) - direction (:747)
        const float ql = __fsqrt_rn(ic_dot3(qx, qy, qz));
        qx = __fdiv_rn(qx, ql); qy = __fdiv_rn(qy, ql); qz = __fdiv_rn(qz, ql);
        pos = make_float4(px, py, pz, 0.f);
        vel.x = __fmul_rn(__fadd_rn(__fmul_rn(uy, qz), -__fmul_rn(uz, qy)), speed);           // cross * speed (:748)
        vel.y = __fmul_rn(__fadd_rn(__fmul_rn(uz, qx), -__fmul_rn(ux, qz)), speed);
        vel.z = __fmul_rn(__fadd_rn(__fmul_rn(ux, qy), -__fmul_rn(uy, qx)), speed);
    }
    packed_a[i] = pos;
    packed_b[i] = pos;
    if (i >= i_first && i - i_first < n_local) {
        float4 *a = reinterpret_cast<float4 *>(side_a + (i - i_first));
        float4 *b = reinterpret_cast<float4 *>(side_b + (i - i_first));
        a[0] = pos; a[1] = vel;
        b[0] = pos; b[1] = vel;
    }
}

// FP32 roofline probe: 16 independent accumulator chains per thread, nothing but FMAs.
template <bool PACKED>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float a, float b)
{
    if (PACKED) {
        float2 acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = make_float2(threadIdx.x * 1e-3f + k, k * 0.5f);
        const float2 a2 = make_float2(a, a * 0.999f), b2 = make_float2(b, b * 1.001f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = __ffma2_rn(acc[k], a2, b2);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += acc[k].x + acc[k].y;
        if (s == 123.456f) out[0] = s;
    } else {
        float acc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = threadIdx.x * 1e-3f + k;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[k] = __fmaf_rn(acc[k], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) s += acc[k];
        if (s == 123.456f) out[0] = s;
    }
}

}  // namespace mapc
