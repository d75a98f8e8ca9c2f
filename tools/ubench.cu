// tools/ubench.cu -- FP32 pipe microbenchmarks and a sweep of force-kernel launch shapes on B200.
//
// Answers, with CUDA-event timings, the questions the force kernel's design rests on:
//   * what the FMA pipe sustains for scalar FFMA vs packed FFMA2, and how the number of distinct
//     register operands (register-file ports / reuse cache) changes it;
//   * MUFU.RSQ throughput, alone and mixed with FFMA2;
//   * which (pairs/thread P, block size T, unroll U, blocks/SM) shape of force_segments_kernel is
//     fastest at N = 262,144 (BASELINE config 3).
// Build: make -C tools   Run: tools/ubench [N]
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../multi-adapter-particles_b200/csrc/nbody_kernels.cuh"

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

constexpr int kIters = 4096;
constexpr int kChains = 8;

// Operand-fetch patterns of the packed FMA (all inputs come from memory so nothing folds):
// pattern 0: acc[k] = fma2(acc[k], s, t)      one fetched pair per op (s, t stay in the reuse cache)
// pattern 1: acc[k] = fma2(x[k], y[k], acc[k]) three distinct pairs, nothing reusable
// pattern 2: acc[k] = fma2(x[k], s, acc[k])    three pairs, s shared by neighbours
// pattern 3: acc[k] = fma2(x[k], x[k], acc[k]) two distinct pairs
// pattern 4: acc[k] = fma2(x[k], c[k].F32 broadcast, acc[k])  two pairs + one scalar
// pattern 5: acc[k] = fma2(x[k], y[k], acc[k]); y[k] = fma2(y[k], y[k], 25)   a 3-pair op followed by a 1-pair op:
//            does the operand fetch average over neighbouring instructions (2 reads/op -> ~2.06 cycles) or is it
//            paid per instruction (3.06 + 2.06)/2?  (values run off to inf: irrelevant for the timing)
// pattern 6: y[k] = fma2(y[k], y[k], 25)        one fetched pair, immediate addend (the 1-pair baseline)
// pattern 7: acc[k] = fadd2(s.x broadcast, -x[k])   the subtraction r = bj - bi: one pair + one scalar
template <int PATTERN>
__global__ void __launch_bounds__(256) mb_ffma2(float2 *out, const float2 *__restrict__ in)
{
    float2 acc[kChains], x[kChains], y[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) {
        acc[k] = in[threadIdx.x + 256 * k];
        x[k] = in[threadIdx.x + 256 * (k + kChains)];
        y[k] = in[threadIdx.x + 256 * (k + 2 * kChains)];
    }
    const float2 s = in[threadIdx.x + 256 * 3 * kChains], t = in[threadIdx.x + 256 * (3 * kChains + 1)];
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) {
            if (PATTERN == 0) acc[k] = __ffma2_rn(acc[k], s, t);
            if (PATTERN == 1) acc[k] = __ffma2_rn(x[k], y[k], acc[k]);
            if (PATTERN == 2) acc[k] = __ffma2_rn(x[k], s, acc[k]);
            if (PATTERN == 3) acc[k] = __ffma2_rn(x[k], x[k], acc[k]);
            if (PATTERN == 4) acc[k] = __ffma2_rn(x[k], make_float2(y[k].x, y[k].x), acc[k]);
            if (PATTERN == 5) {
                acc[k] = __ffma2_rn(x[k], y[k], acc[k]);
                y[k] = __ffma2_rn(y[k], y[k], make_float2(25.f, 25.f));
            }
            if (PATTERN == 6) y[k] = __ffma2_rn(y[k], y[k], make_float2(25.f, 25.f));
            if (PATTERN == 7) acc[k] = __fadd2_rn(make_float2(s.x, s.x), make_float2(-acc[k].x, -acc[k].y));
        }
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kChains; ++k) r = __fadd2_rn(r, __fadd2_rn(acc[k], PATTERN >= 5 ? y[k] : make_float2(0.f, 0.f)));
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

// scalar: pattern 0 acc = fma(acc, a, b); pattern 1 acc = fma(x[k], y[k], acc);
// pattern 2: the halves of packed pairs: acc.x = fma(x.x, y.x, acc.x), acc.y = fma(x.y, y.y, acc.y)
template <int PATTERN>
__global__ void __launch_bounds__(256) mb_ffma(float2 *out, const float2 *__restrict__ in)
{
    float2 acc[kChains], x[kChains], y[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) {
        acc[k] = in[threadIdx.x + 256 * k];
        x[k] = in[threadIdx.x + 256 * (k + kChains)];
        y[k] = in[threadIdx.x + 256 * (k + 2 * kChains)];
    }
    const float2 s = in[threadIdx.x + 256 * 3 * kChains];
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) {
            if (PATTERN == 0) {
                acc[k].x = __fmaf_rn(acc[k].x, s.x, s.y);
                acc[k].y = __fmaf_rn(acc[k].y, s.x, s.y);
            } else if (PATTERN == 3) {
                acc[k].y = __fmaf_rn(x[k].x, y[k].x, acc[k].y);
                acc[k].x = __fmaf_rn(x[k].y, y[k].y, acc[k].x);
            } else {
                acc[k].x = __fmaf_rn(x[k].x, y[k].x, acc[k].x);
                acc[k].y = __fmaf_rn(x[k].y, y[k].y, acc[k].y);
            }
        }
        if (PATTERN >= 2) {  // keep x, y alive as packed pairs so their halves sit in aligned registers
#pragma unroll
            for (int k = 0; k < kChains; ++k) x[k] = __fadd2_rn(x[k], y[k]);
        }
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kChains; ++k) r = __fadd2_rn(r, acc[k]);
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

// MUFU.RSQ alone (MIX = 0) or one MUFU per 3 two-pair FFMA2 (the n-body ratio, MIX = 1)
template <int MIX>
__global__ void __launch_bounds__(256) mb_mufu(float *out, const float2 *__restrict__ in)
{
    float v[2 * kChains];
    float2 acc[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) {
        acc[k] = in[threadIdx.x + 256 * k];
        v[2 * k] = 1.0f + fabsf(acc[k].x);
        v[2 * k + 1] = 1.0f + fabsf(acc[k].y);
    }
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) {
            v[2 * k] = mapc::rsqrt_approx(v[2 * k]);
            v[2 * k + 1] = mapc::rsqrt_approx(v[2 * k + 1]);
            if (MIX) {
                const float2 w = make_float2(v[2 * k], v[2 * k + 1]);
#pragma unroll
                for (int r = 0; r < 6 * MIX; ++r) acc[(k + r) % kChains] = __ffma2_rn(acc[(k + r) % kChains], w, w);
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kChains; ++k) r += v[k];
#pragma unroll
    for (int k = 0; k < kChains; ++k) r += acc[k].x + acc[k].y;
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
    template <class F> float best(F f, int reps = 5)
    {
        float best = 1e30f;
        for (int r = 0; r < reps; ++r) {
            CK(cudaEventRecord(a));
            f();
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (r > 0 && ms < best) best = ms;
        }
        return best;
    }
};

static int g_sms = 148;
static double g_peak_tflops = 74.45;

static int g_segments = 32;  // canonical S (argv[3]); the library's rule gives 32 up to N = 262,144
static int g_targets = 0;   // 0 = all bodies are targets; otherwise a shard of that many (multi-GPU shapes)

template <int P, int T, int TJ, int U, int MINB, int ORDER = 0, bool TMA = false, bool INLOOP = false>
void run_force(const float4 *pos, float4 *partial, int n, Timer &t)
{
    const int n_tgt = g_targets > 0 ? g_targets : n;
    const int S = g_segments;
    mapc::StepArgs args{};
    args.pos = pos;
    args.partial = partial;
    args.i_first = 0;
    args.i_cnt = n_tgt;
    args.n_sources = n;
    args.S = S;
    args.segs.count = S;
    for (int s = 0; s < S; ++s) args.segs.ids[s] = s;
    const int per_block = T * 2 * P;
    args.n_iblocks = (n_tgt + per_block - 1) / per_block;
    args.scratch_blocks = args.n_iblocks;
    auto kernel = mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, false, false, TMA, INLOOP>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, T, 0));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kernel));
    const long long cells = (long long)args.n_iblocks * S;
    const dim3 grid(args.n_iblocks * S);
    const float ms = t.best([&] { kernel<<<grid, T>>>(args); }, 4);
    const double ginter = (double)n * n_tgt / (ms * 1e-3) / 1e9;
    printf("force %s %s %s P=%d T=%3d TJ=%3d U=%d minB=%2d regs=%3d occ=%2d blk/SM (%2d warps) cells=%5lld : %8.3f ms %8.1f G int/s %5.1f %% of %.2f TF\n",
           TMA ? "tma-stage" : "ldg-stage", INLOOP ? "mass-in-loop    " : "mass-per-partial", ORDER == 0 ? "pair-major" : "op-major  ", P, T, TJ, U, MINB, fa.numRegs, occ, occ * T / 32, cells, ms, ginter,
           100.0 * ginter * 20.0 / 1e3 / g_peak_tflops, g_peak_tflops);
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 262144;
    g_targets = argc > 2 ? atoi(argv[2]) : 0;
    g_segments = argc > 3 ? atoi(argv[3]) : 32;
    if (g_segments < 1 || g_segments > MAPC_MAX_SEGMENTS) { fprintf(stderr, "S must be in 1..32\n"); return 2; }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    g_peak_tflops = g_sms * 128.0 * 2.0 * clk_khz * 1e3 / 1e12;
    printf("%s: %d SMs, max SM clock %d MHz -> FP32 peak %.2f TFLOP/s\n", prop.name, g_sms, clk_khz / 1000,
           g_peak_tflops);
    Timer t;
    const int max_blocks = g_sms * 8;
    float2 *sink, *input;
    CK(cudaMalloc(&sink, sizeof(float2) * 256 * max_blocks));
    {
        std::vector<float2> hin(256 * 32);
        unsigned st = 777u;
        for (auto &v : hin) {
            st = st * 1664525u + 1013904223u;
            v.x = 0.5f + (float)(st >> 8) / 16777216.0f;
            st = st * 1664525u + 1013904223u;
            v.y = 0.5f + (float)(st >> 8) / 16777216.0f;
        }
        CK(cudaMalloc(&input, sizeof(float2) * hin.size()));
        CK(cudaMemcpy(input, hin.data(), sizeof(float2) * hin.size(), cudaMemcpyHostToDevice));
    }
    const int threads = 256;
    for (int bps : {1, 2, 8}) {   // resident blocks per SM: 2, 4 and 16 warps per sub-partition
        const int blocks = g_sms * bps;
        printf("--- FMA-pipe operand patterns, %d block(s)/SM = %d warps per SM sub-partition ---\n", bps, bps * 2);
        auto report = [&](const char *name, float ms, double lane_fma) {
            const double tf = 2.0 * lane_fma * blocks * threads / (ms * 1e-3) / 1e12;
            printf("%-66s %8.3f ms %7.2f TFLOP/s  %5.1f %% of peak\n", name, ms, tf, 100.0 * tf / g_peak_tflops);
        };
        const double per = 2.0 * kChains * kIters;
        report("FFMA2 acc=fma2(acc,s,t)         1 fetched pair", t.best([&] { mb_ffma2<0><<<blocks, threads>>>(sink, input); }), per);
        report("FFMA2 acc=fma2(x[k],y[k],acc)   3 distinct pairs", t.best([&] { mb_ffma2<1><<<blocks, threads>>>(sink, input); }), per);
        report("FFMA2 acc=fma2(x[k],s,acc)      3 pairs, s shared (reuse)", t.best([&] { mb_ffma2<2><<<blocks, threads>>>(sink, input); }), per);
        report("FFMA2 acc=fma2(x[k],x[k],acc)   2 distinct pairs", t.best([&] { mb_ffma2<3><<<blocks, threads>>>(sink, input); }), per);
        report("FFMA2 acc=fma2(x[k],c[k].F32,acc) 2 pairs + scalar broadcast", t.best([&] { mb_ffma2<4><<<blocks, threads>>>(sink, input); }), per);
        report("FFMA2 3-pair op then 1-pair op (does operand fetch average?)", t.best([&] { mb_ffma2<5><<<blocks, threads>>>(sink, input); }), 2.0 * per);
        report("FFMA2 y=fma2(y,y,25)            1 fetched pair, immediate addend", t.best([&] { mb_ffma2<6><<<blocks, threads>>>(sink, input); }), per);
        report("FADD2 acc=fadd2(s.F32,-acc)     1 pair + scalar broadcast (r = bj - bi)", t.best([&] { mb_ffma2<7><<<blocks, threads>>>(sink, input); }), per);
        report("FFMA  acc=fma(acc,a,b)          1 fetched reg", t.best([&] { mb_ffma<0><<<blocks, threads>>>(sink, input); }), per);
        report("FFMA  acc=fma(x[k],y[k],acc)    3 distinct regs (free allocation)", t.best([&] { mb_ffma<1><<<blocks, threads>>>(sink, input); }), per);
        report("FFMA  on halves of packed pairs  3 distinct regs, same parity", t.best([&] { mb_ffma<2><<<blocks, threads>>>(sink, input); }), per);
        report("FFMA  on halves of packed pairs  acc halves crossed (e,e,o)/(o,o,e)", t.best([&] { mb_ffma<3><<<blocks, threads>>>(sink, input); }), per);
        const float ms = t.best([&] { mb_mufu<0><<<blocks, threads>>>((float *)sink, input); });
        const double per_clk_sm = 2.0 * kChains * kIters * (double)blocks * threads / (ms * 1e-3) / (clk_khz * 1e3) / g_sms;
        printf("%-66s %8.3f ms %7.2f MUFU.RSQ lanes/clk/SM\n", "MUFU.RSQ alone", ms, per_clk_sm);
        const float ms2 = t.best([&] { mb_mufu<1><<<blocks, threads>>>((float *)sink, input); });
        const double tf = 2.0 * (2.0 * 6 * kChains * kIters) * blocks * threads / (ms2 * 1e-3) / 1e12;
        printf("%-66s %8.3f ms %7.2f TFLOP/s  %5.1f %% of peak (FFMA2 part)\n", "2 MUFU.RSQ + 6 FFMA2 (XU-bound mix)", ms2, tf,
               100.0 * tf / g_peak_tflops);
        const float ms3 = t.best([&] { mb_mufu<2><<<blocks, threads>>>((float *)sink, input); });
        const double tf3 = 2.0 * (2.0 * 12 * kChains * kIters) * blocks * threads / (ms3 * 1e-3) / 1e12;
        printf("%-66s %8.3f ms %7.2f TFLOP/s  %5.1f %% of peak (FFMA2 part)\n", "2 MUFU.RSQ + 12 FFMA2 (the n-body ratio)", ms3, tf3,
               100.0 * tf3 / g_peak_tflops);
    }

    // ---- force kernel shapes at N bodies ----------------------------------------------------
    std::vector<float4> h(n);
    unsigned s = 12345u;
    auto rnd = [&] { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
    for (auto &v : h) v = make_float4(8000.f * (rnd() - 0.5f), 8000.f * (rnd() - 0.5f), 8000.f * (rnd() - 0.5f), 0.f);
    float4 *pos, *partial;
    CK(cudaMalloc(&pos, sizeof(float4) * n));
    CK(cudaMalloc(&partial, sizeof(float4) * n * 64));
    CK(cudaMemcpy(pos, h.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    printf("--- force_cells_kernel, %d sources, %d targets, S = %d ---\n", n, g_targets > 0 ? g_targets : n, g_segments);
#ifdef SWEEP_SMALL
    run_force<1, 64, 64, 8, 16, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 8, 18, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 8, 20, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 4, 20, 0>(pos, partial, n, t);
    run_force<1, 32, 64, 8, 32, 0>(pos, partial, n, t);
    run_force<2, 64, 64, 4, 8, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 4, 10, 2>(pos, partial, n, t);
    run_force<2, 32, 64, 4, 16, 2>(pos, partial, n, t);
    run_force<2, 32, 64, 4, 20, 2>(pos, partial, n, t);
    run_force<2, 32, 64, 8, 20, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 8, 8, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 8, 10, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 4, 4, 2>(pos, partial, n, t);
    return 0;
#endif
#ifdef SWEEP_BROAD
    run_force<4, 256, 256, 1, 2, 0>(pos, partial, n, t);
    run_force<4, 256, 256, 1, 3, 0>(pos, partial, n, t);
    run_force<4, 256, 256, 2, 2, 0>(pos, partial, n, t);
    run_force<4, 256, 256, 2, 3, 0>(pos, partial, n, t);
    run_force<4, 256, 256, 4, 2, 0>(pos, partial, n, t);
    run_force<4, 256, 256, 4, 3, 0>(pos, partial, n, t);
    run_force<4, 256, 256, 8, 2, 0>(pos, partial, n, t);
    run_force<4, 256, 256, 8, 3, 0>(pos, partial, n, t);
    run_force<4, 128, 256, 1, 4, 0>(pos, partial, n, t);
    run_force<4, 128, 256, 1, 6, 0>(pos, partial, n, t);
    run_force<4, 128, 256, 2, 4, 0>(pos, partial, n, t);
    run_force<4, 128, 256, 2, 6, 0>(pos, partial, n, t);
    run_force<4, 128, 256, 4, 4, 0>(pos, partial, n, t);
    run_force<4, 128, 256, 4, 6, 0>(pos, partial, n, t);
    run_force<4, 128, 256, 8, 4, 0>(pos, partial, n, t);
    run_force<4, 128, 256, 8, 6, 0>(pos, partial, n, t);
    run_force<2, 256, 256, 1, 2, 0>(pos, partial, n, t);
    run_force<2, 256, 256, 1, 4, 0>(pos, partial, n, t);
    run_force<2, 256, 256, 2, 2, 0>(pos, partial, n, t);
    run_force<2, 256, 256, 2, 4, 0>(pos, partial, n, t);
    run_force<2, 256, 256, 4, 2, 0>(pos, partial, n, t);
    run_force<2, 256, 256, 4, 4, 0>(pos, partial, n, t);
    run_force<2, 256, 256, 8, 2, 0>(pos, partial, n, t);
    run_force<2, 256, 256, 8, 4, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 1, 4, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 1, 8, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 2, 4, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 2, 8, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 4, 4, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 4, 8, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 8, 4, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 8, 8, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 1, 4, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 1, 6, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 1, 8, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 2, 4, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 2, 6, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 2, 8, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 4, 4, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 4, 6, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 4, 8, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 8, 4, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 8, 6, 0>(pos, partial, n, t);
    run_force<1, 256, 256, 8, 8, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 1, 8, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 1, 12, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 1, 16, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 2, 8, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 2, 12, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 2, 16, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 4, 8, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 4, 12, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 4, 16, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 8, 8, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 8, 12, 0>(pos, partial, n, t);
    run_force<1, 128, 256, 8, 16, 0>(pos, partial, n, t);
    run_force<2, 64, 64, 1, 8, 0>(pos, partial, n, t);
    run_force<2, 64, 64, 1, 16, 0>(pos, partial, n, t);
    run_force<2, 64, 64, 2, 8, 0>(pos, partial, n, t);
    run_force<2, 64, 64, 2, 16, 0>(pos, partial, n, t);
    run_force<2, 64, 64, 4, 8, 0>(pos, partial, n, t);
    run_force<2, 64, 64, 4, 16, 0>(pos, partial, n, t);
    run_force<2, 64, 64, 8, 8, 0>(pos, partial, n, t);
    run_force<2, 64, 64, 8, 16, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 1, 16, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 1, 24, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 2, 16, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 2, 24, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 4, 16, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 4, 24, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 8, 16, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 8, 24, 0>(pos, partial, n, t);
    run_force<4, 256, 256, 1, 2, 2>(pos, partial, n, t);
    run_force<4, 256, 256, 1, 3, 2>(pos, partial, n, t);
    run_force<4, 256, 256, 2, 2, 2>(pos, partial, n, t);
    run_force<4, 256, 256, 2, 3, 2>(pos, partial, n, t);
    run_force<4, 256, 256, 4, 2, 2>(pos, partial, n, t);
    run_force<4, 256, 256, 4, 3, 2>(pos, partial, n, t);
    run_force<4, 256, 256, 8, 2, 2>(pos, partial, n, t);
    run_force<4, 256, 256, 8, 3, 2>(pos, partial, n, t);
    run_force<4, 128, 256, 1, 4, 2>(pos, partial, n, t);
    run_force<4, 128, 256, 1, 6, 2>(pos, partial, n, t);
    run_force<4, 128, 256, 2, 4, 2>(pos, partial, n, t);
    run_force<4, 128, 256, 2, 6, 2>(pos, partial, n, t);
    run_force<4, 128, 256, 4, 4, 2>(pos, partial, n, t);
    run_force<4, 128, 256, 4, 6, 2>(pos, partial, n, t);
    run_force<4, 128, 256, 8, 4, 2>(pos, partial, n, t);
    run_force<4, 128, 256, 8, 6, 2>(pos, partial, n, t);
    run_force<2, 256, 256, 1, 2, 2>(pos, partial, n, t);
    run_force<2, 256, 256, 1, 4, 2>(pos, partial, n, t);
    run_force<2, 256, 256, 2, 2, 2>(pos, partial, n, t);
    run_force<2, 256, 256, 2, 4, 2>(pos, partial, n, t);
    run_force<2, 256, 256, 4, 2, 2>(pos, partial, n, t);
    run_force<2, 256, 256, 4, 4, 2>(pos, partial, n, t);
    run_force<2, 256, 256, 8, 2, 2>(pos, partial, n, t);
    run_force<2, 256, 256, 8, 4, 2>(pos, partial, n, t);
    run_force<2, 128, 256, 1, 4, 2>(pos, partial, n, t);
    run_force<2, 128, 256, 1, 8, 2>(pos, partial, n, t);
    run_force<2, 128, 256, 2, 4, 2>(pos, partial, n, t);
    run_force<2, 128, 256, 2, 8, 2>(pos, partial, n, t);
    run_force<2, 128, 256, 4, 4, 2>(pos, partial, n, t);
    run_force<2, 128, 256, 4, 8, 2>(pos, partial, n, t);
    run_force<2, 128, 256, 8, 4, 2>(pos, partial, n, t);
    run_force<2, 128, 256, 8, 8, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 1, 4, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 1, 6, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 1, 8, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 2, 4, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 2, 6, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 2, 8, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 4, 4, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 4, 6, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 4, 8, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 8, 4, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 8, 6, 2>(pos, partial, n, t);
    run_force<1, 256, 256, 8, 8, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 1, 8, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 1, 12, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 1, 16, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 2, 8, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 2, 12, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 2, 16, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 4, 8, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 4, 12, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 4, 16, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 8, 8, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 8, 12, 2>(pos, partial, n, t);
    run_force<1, 128, 256, 8, 16, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 1, 8, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 1, 16, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 2, 8, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 2, 16, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 4, 8, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 4, 16, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 8, 8, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 8, 16, 2>(pos, partial, n, t);
    run_force<1, 64, 64, 1, 16, 2>(pos, partial, n, t);
    run_force<1, 64, 64, 1, 24, 2>(pos, partial, n, t);
    run_force<1, 64, 64, 2, 16, 2>(pos, partial, n, t);
    run_force<1, 64, 64, 2, 24, 2>(pos, partial, n, t);
    run_force<1, 64, 64, 4, 16, 2>(pos, partial, n, t);
    run_force<1, 64, 64, 4, 24, 2>(pos, partial, n, t);
    run_force<1, 64, 64, 8, 16, 2>(pos, partial, n, t);
    run_force<1, 64, 64, 8, 24, 2>(pos, partial, n, t);
#else
    run_force<4, 256, 256, 8, 2, 0>(pos, partial, n, t);
    run_force<4, 128, 256, 8, 4, 0>(pos, partial, n, t);
    run_force<2, 256, 256, 8, 2, 0>(pos, partial, n, t);
    run_force<2, 128, 256, 4, 4, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 4, 8, 2>(pos, partial, n, t);
    run_force<2, 64, 64, 8, 8, 0>(pos, partial, n, t);
    run_force<1, 64, 64, 8, 16, 0>(pos, partial, n, t);
    run_force<1, 32, 64, 8, 32, 0>(pos, partial, n, t);
    run_force<4, 256, 256, 8, 2, 0, false, true>(pos, partial, n, t);   // the shader's per-pair mass multiply (12 lane-ops)
    run_force<4, 128, 256, 8, 4, 0, false, true>(pos, partial, n, t);
    run_force<2, 128, 256, 4, 4, 2, false, true>(pos, partial, n, t);
    run_force<4, 256, 256, 8, 2, 0, true>(pos, partial, n, t);
    run_force<4, 128, 256, 8, 4, 0, true>(pos, partial, n, t);
    run_force<2, 128, 256, 4, 4, 2, true>(pos, partial, n, t);
    run_force<1, 64, 64, 8, 16, 0, true>(pos, partial, n, t);
#endif
    {   // the TMA-staged kernel must produce the same bits as the LDG/STS-staged one
        float4 *partial2;
        CK(cudaMalloc(&partial2, sizeof(float4) * n * 8));
        const int n_tgt = g_targets > 0 ? g_targets : n;
        mapc::StepArgs args{};
        args.pos = pos; args.i_first = 0; args.i_cnt = n_tgt; args.n_sources = n; args.S = 8;
        args.segs.count = 8;
        for (int sgm = 0; sgm < 8; ++sgm) args.segs.ids[sgm] = sgm;
        args.n_iblocks = (n_tgt + 2047) / 2048;
        args.scratch_blocks = args.n_iblocks;
        CK(cudaMemset(partial, 0, sizeof(float4) * n * 8));
        CK(cudaMemset(partial2, 0, sizeof(float4) * n * 8));
        args.partial = partial;
        mapc::force_cells_kernel<4, 256, 256, 8, 2, 0, false, false, false, false><<<dim3(args.n_iblocks * 8), 256>>>(args);
        args.partial = partial2;
        mapc::force_cells_kernel<4, 256, 256, 8, 2, 0, false, false, true, false><<<dim3(args.n_iblocks * 8), 256>>>(args);
        CK(cudaDeviceSynchronize());
        std::vector<float4> a((size_t)n * 8), b((size_t)n * 8);
        CK(cudaMemcpy(a.data(), partial, sizeof(float4) * a.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), partial2, sizeof(float4) * b.size(), cudaMemcpyDeviceToHost));
        printf("tma-staged partials bit-identical to ldg-staged: %s\n",
               memcmp(a.data(), b.data(), sizeof(float4) * a.size()) == 0 ? "yes" : "NO");
    }
    return 0;
}
