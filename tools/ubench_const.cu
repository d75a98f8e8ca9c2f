// ubench_const.cu -- what would the force kernel gain if the source body were a UNIFORM-register operand?
//
// In the library's kernel every FADD2 (r = bj - bi) fetches a register pair (the two targets) plus a
// vector register holding the broadcast source coordinate (SASS `R.F32`).  When the source comes from a
// constant bank instead, ptxas loads it with LDCU straight into uniform registers and the subtraction
// becomes `FADD2 R, -R.F32x2, UR.F32` -- one vector-register-file fetch instead of two.  This benchmark
// runs the same 11-op pair math with sources in __constant__ memory (a fixed 2048-body table walked
// `reps` times, so the arithmetic volume equals a real segment) next to the shared-memory version.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../multi-adapter-particles_b200/csrc/nbody_kernels.cuh"

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int kConstBodies = 2048;
__constant__ float4 c_src[kConstBodies];

template <int P, int T, int U, int MINB, int ORDER>
__global__ void __launch_bounds__(T, MINB) force_const_kernel(const float4 *__restrict__ pos, float4 *__restrict__ partial,
                                                              int n_targets, int reps)
{
    const int tid = threadIdx.x;
    const int i_block = blockIdx.x * (T * 2 * P);
    float2 nxi[P], nyi[P], nzi[P], ax[P], ay[P], az[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        int ia = i_block + (2 * p) * T + tid, ib = i_block + (2 * p + 1) * T + tid;
        ia = ia < n_targets ? ia : n_targets - 1;
        ib = ib < n_targets ? ib : n_targets - 1;
        const float4 ta = pos[ia], tb = pos[ib];
        nxi[p] = make_float2(-ta.x, -tb.x);
        nyi[p] = make_float2(-ta.y, -tb.y);
        nzi[p] = make_float2(-ta.z, -tb.z);
        ax[p] = ay[p] = az[p] = make_float2(0.f, 0.f);
    }
    for (int r = 0; r < reps; ++r) {
#pragma unroll U
        for (int j = 0; j < kConstBodies; ++j) {
            const float4 b = c_src[j];
            if (ORDER == 2) {
                mapc::group_interaction<P, false>(b, nxi, nyi, nzi, ax, ay, az);
            } else {
#pragma unroll
                for (int p = 0; p < P; ++p) mapc::pair_interaction<false>(b, nxi[p], nyi[p], nzi[p], ax[p], ay[p], az[p]);
            }
        }
    }
    float4 *out = partial + (size_t)blockIdx.y * n_targets;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int ia = i_block + (2 * p) * T + tid, ib = i_block + (2 * p + 1) * T + tid;
        if (ia < n_targets) out[ia] = make_float4(ax[p].x, ay[p].x, az[p].x, 0.f);
        if (ib < n_targets) out[ib] = make_float4(ax[p].y, ay[p].y, az[p].y, 0.f);
    }
}

static double g_peak = 74.45;

struct Timer {
    cudaEvent_t a, b;
    Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
    template <class F> float best(F f, int reps = 4)
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

template <int P, int T, int U, int MINB, int ORDER>
void run_const(const float4 *pos, float4 *partial, int n, int S, Timer &t)
{
    auto kernel = force_const_kernel<P, T, U, MINB, ORDER>;
    const int per_block = T * 2 * P;
    const dim3 grid((n + per_block - 1) / per_block, S);
    const int reps = n / S / kConstBodies;   // sources per cell = n/S, like a canonical segment
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, T, 0));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kernel));
    const float ms = t.best([&] { kernel<<<grid, T>>>(pos, partial, n, reps); });
    const double ginter = (double)n * ((double)reps * kConstBodies * S) / (ms * 1e-3) / 1e9;
    printf("const-src %s P=%d T=%3d U=%d minB=%2d regs=%3d occ=%2d blk/SM (%2d warps) : %8.3f ms %8.1f G int/s %5.1f %%\n",
           ORDER == 0 ? "pair-major" : "op-major  ", P, T, U, MINB, fa.numRegs, occ, occ * T / 32, ms, ginter,
           100.0 * ginter * 20.0 / 1e3 / g_peak);
}

template <int P, int T, int TJ, int U, int MINB, int ORDER>
void run_smem(const float4 *pos, float4 *partial, int n, int S, Timer &t)
{
    mapc::StepArgs args{};
    args.pos = pos;
    args.partial = partial;
    args.i_cnt = n;
    args.n_sources = n;
    args.S = S;
    args.segs.count = S;
    for (int s = 0; s < S; ++s) args.segs.ids[s] = s;
    args.n_iblocks = (n + T * 2 * P - 1) / (T * 2 * P);
    args.scratch_blocks = args.n_iblocks;
    auto kernel = mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, false>;
    const dim3 grid(args.n_iblocks * S);
    const float ms = t.best([&] { kernel<<<grid, T>>>(args); });
    const double ginter = (double)n * n / (ms * 1e-3) / 1e9;
    printf("smem-src  %s P=%d T=%3d U=%d minB=%2d                              : %8.3f ms %8.1f G int/s %5.1f %%\n",
           ORDER == 0 ? "pair-major" : "op-major  ", P, T, U, MINB, ms, ginter, 100.0 * ginter * 20.0 / 1e3 / g_peak);
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 262144;
    const int S = 8;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    g_peak = prop.multiProcessorCount * 128.0 * 2.0 * clk_khz * 1e3 / 1e12;
    printf("%s: %d SMs, %d MHz -> FP32 peak %.2f TFLOP/s; N = %d, S = %d\n", prop.name, prop.multiProcessorCount,
           clk_khz / 1000, g_peak, n, S);
    std::vector<float4> h(n);
    unsigned s = 12345u;
    auto rnd = [&] { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
    for (auto &v : h) v = make_float4(8000.f * (rnd() - 0.5f), 8000.f * (rnd() - 0.5f), 8000.f * (rnd() - 0.5f), 0.f);
    float4 *pos, *partial;
    CK(cudaMalloc(&pos, sizeof(float4) * n));
    CK(cudaMalloc(&partial, sizeof(float4) * (size_t)n * S));
    CK(cudaMemcpy(pos, h.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(c_src, h.data(), sizeof(float4) * kConstBodies));
    Timer t;
    run_smem<4, 256, 256, 8, 2, 0>(pos, partial, n, S, t);
    run_const<4, 256, 8, 2, 0>(pos, partial, n, S, t);
    run_const<4, 256, 4, 2, 0>(pos, partial, n, S, t);
    run_const<4, 256, 8, 2, 2>(pos, partial, n, S, t);
    run_const<4, 256, 16, 2, 0>(pos, partial, n, S, t);
    run_const<4, 256, 8, 1, 0>(pos, partial, n, S, t);
    run_const<4, 256, 8, 3, 0>(pos, partial, n, S, t);
    run_smem<4, 128, 256, 8, 4, 0>(pos, partial, n, S, t);
    run_const<4, 128, 8, 4, 0>(pos, partial, n, S, t);
    run_const<4, 128, 8, 4, 2>(pos, partial, n, S, t);
    run_const<4, 128, 8, 5, 0>(pos, partial, n, S, t);
    run_const<4, 128, 8, 6, 0>(pos, partial, n, S, t);
    run_const<4, 64, 8, 8, 0>(pos, partial, n, S, t);
    run_smem<2, 128, 256, 4, 4, 2>(pos, partial, n, S, t);
    run_const<2, 128, 4, 4, 2>(pos, partial, n, S, t);
    run_const<2, 128, 8, 4, 0>(pos, partial, n, S, t);
    run_const<2, 128, 8, 8, 0>(pos, partial, n, S, t);
    run_const<2, 256, 8, 4, 0>(pos, partial, n, S, t);
    run_const<2, 64, 8, 16, 0>(pos, partial, n, S, t);
    run_smem<1, 64, 64, 8, 16, 0>(pos, partial, n, S, t);
    run_const<1, 64, 8, 16, 0>(pos, partial, n, S, t);
    run_const<1, 128, 8, 8, 0>(pos, partial, n, S, t);
    run_const<1, 256, 8, 4, 0>(pos, partial, n, S, t);
    run_const<1, 256, 16, 8, 0>(pos, partial, n, S, t);
    run_const<8, 128, 4, 2, 0>(pos, partial, n, S, t);
    run_const<8, 64, 4, 4, 0>(pos, partial, n, S, t);
    return 0;
}
