// tools/sassprobe/rfrun.cu -- loads every cubin of a variants blob (written by rfprobe.py) with the driver API, runs the
// carrier kernel `rfprobe` and prints event time and in-kernel cycles per loop instruction.
//   rfrun variants.bin [iters] [nops]
#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); \
    printf("CUDA error %s at %s:%d\n", s_, __FILE__, __LINE__); exit(1); } } while (0)
int main(int argc, char **argv)
{
    if (argc < 2) { printf("usage: rfrun variants.bin [iters] [nops]\n"); return 2; }
    int iters = argc > 2 ? atoi(argv[2]) : 20000, nops = argc > 3 ? atoi(argv[3]) : 16;
    FILE *f = fopen(argv[1], "rb"); if (!f) { perror("open"); return 1; }
    CK(cuInit(0)); CUdevice dev; CK(cuDeviceGet(&dev, 0)); CUcontext ctx; CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
    int sms, khz; CK(cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
    CK(cuDeviceGetAttribute(&khz, CU_DEVICE_ATTRIBUTE_CLOCK_RATE, dev));
    const int blocks_per_sm = 4, threads = 128, grid = sms * blocks_per_sm;
    CUdeviceptr in, out, cyc; CK(cuMemAlloc(&in, 128 * 64 * 8)); CK(cuMemAlloc(&out, (size_t)grid * threads * 8)); CK(cuMemAlloc(&cyc, grid * 8));
    std::vector<float> h(128 * 64 * 2); for (size_t i = 0; i < h.size(); ++i) h[i] = 0.5f + 1e-3f * (i % 97);
    CK(cuMemcpyHtoD(in, h.data(), h.size() * 4));
    CUevent e0, e1; CK(cuEventCreate(&e0, 0)); CK(cuEventCreate(&e1, 0));
    printf("# %d SMs, %d kHz, grid %d x %d threads (%d warps per scheduler), %d iters x %d ops\n", sms, khz, grid, threads, blocks_per_sm, iters, nops);
    printf("# name  event_ms  cycles_per_op(event,at max clock)  cycles_per_op(clock64 median)\n");
    char name[64]; unsigned size;
    while (fread(name, 1, 64, f) == 64 && fread(&size, 4, 1, f) == 1) {
        std::vector<char> img(size); if (fread(img.data(), 1, size, f) != size) break;
        CUmodule mod; CK(cuModuleLoadData(&mod, img.data())); CUfunction fn; CK(cuModuleGetFunction(&fn, mod, "rfprobe"));
        void *args[] = {&out, &in, &iters, &cyc};
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cuEventRecord(e0, 0));
            CK(cuLaunchKernel(fn, grid, 1, 1, threads, 1, 1, 0, 0, args, nullptr));
            CK(cuEventRecord(e1, 0)); CK(cuEventSynchronize(e1));
            float ms; CK(cuEventElapsedTime(&ms, e0, e1)); if (rep) best = std::min(best, ms);
        }
        std::vector<long long> c(grid); CK(cuMemcpyDtoH(c.data(), cyc, grid * 8)); std::sort(c.begin(), c.end());
        const double ops = (double)iters * nops * blocks_per_sm;   // ops issued by one scheduler
        printf("%-40s %8.4f  %6.3f  %6.3f\n", name, best, best * 1e-3 * khz * 1e3 / ops, (double)c[grid / 2] / ops);
        CK(cuModuleUnload(mod));
    }
    return 0;
}
