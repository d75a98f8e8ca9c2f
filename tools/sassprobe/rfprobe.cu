// tools/sassprobe/rfprobe.cu -- carrier kernel for SASS-level register-file probes (see rfprobe.py).
// The loop body is NOPS packed FFMA2 over a pool of live register pairs; rfprobe.py rewrites the register fields (and
// reuse flags) of those instructions in the compiled cubin to place operands in chosen registers; rfrun times each variant.
#include <cuda_runtime.h>
#ifndef NPOOL
#define NPOOL 56
#endif
#ifndef NOPS
#define NOPS 16
#endif
extern "C" __global__ void __launch_bounds__(128, 4) rfprobe(float2 *out, const float2 *__restrict__ in, int iters,
                                                             long long *cycles)
{
    float2 v[NPOOL];
#pragma unroll
    for (int k = 0; k < NPOOL; ++k) v[k] = in[threadIdx.x + 128 * k];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NOPS; ++k) v[k] = __ffma2_rn(v[NOPS + k], v[2 * NOPS + k], v[k]);
    }
    const long long t1 = clock64();
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < NPOOL; ++k) r = __fadd2_rn(r, v[k]);
    out[blockIdx.x * 128 + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
