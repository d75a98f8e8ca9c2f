// tests/emu/emu_asan_main.cpp -- TEST INFRASTRUCTURE ONLY: drives the emulated kernels (emu_kernels.cpp) over
// every launch shape x {fused, unfused, mass-in-loop, 256-source chains, a 2-slot scratch ring, 4 ranks with the
// peer layout, TMA staging, warp-shuffle broadcast} on
// exactly-sized heap buffers, for an AddressSanitizer + UBSan build (`make -C tests/emu asan`): any read or
// write of the kernel source outside its buffers, and any signed overflow / misaligned access, aborts.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
struct PV { float pos[4]; float velo[4]; };
extern "C" int emu_step_allpairs(const void *in, void *out, float *pos_next_out, int n, int n_active,
                      float dt, float damping, int S, int pairs, int threads, int fuse, int mass_in_loop,
                      int world, int peer, int block_order, int chain, int staging, int ring_slots,
                      unsigned long long *info);
extern "C" int emu_steps_chained(void *state, int n, int steps, float dt, float damping, int S, int pairs, int threads,
                                 int block_order);
extern "C" int emu_init_particles(void *side_a, void *side_b, float *packed_a, float *packed_b, unsigned n,
                                  unsigned i_first, unsigned n_local, unsigned seed);
extern "C" int emu_step_well(const void *in, void *out, float *pos_next_out, float *packed_out, int n,
                  int n_active, float dt, float damping, int i_first, int n_local);
int main() {
    const int shapes[7][2] = {{4,256},{4,128},{2,128},{2,64},{1,128},{1,64},{1,32}};
    int runs = 0;
    for (int n : {1, 65, 1000, 1100, 2048}) {
        std::vector<PV> in(n), out(n); std::vector<float> mirror(4*n), packed(4*n);
        srand(n);
        for (auto &b : in) { for (int k=0;k<3;++k){ b.pos[k] = (rand()%20000)/50.f-200.f; b.velo[k]=(rand()%100)/50.f; } b.pos[3]=0; b.velo[3]=0; }
        out = in;
        unsigned long long info[3];
        for (auto &sh : shapes)
            for (int S : {1, 2, 32})
                for (int variant = 0; variant < 8; ++variant) {
                    int fuse = variant != 1, inloop = variant == 2, chain = variant == 3 ? 256 : 2048, ring = variant == 4 ? 2 : 0;
                    int world = (variant == 5 && n == 2048 && S == 32) ? 4 : 1, peer = world > 1;
                    int staging = variant == 6 ? 1 : (variant == 7 ? 2 : 0);
                    if (staging == 1 && sh[1] * 2 * sh[0] < 512 && !(sh[0] == 2 && sh[1] == 128)) continue;   // TMA: 256-body-stage shapes only
                    if (staging == 2 && n > 1000) continue;                       // shuffles are slow to emulate
                    int rc = emu_step_allpairs(in.data(), out.data(), mirror.data(), n, n, 0.1f, 1.f, S, sh[0], sh[1], fuse, inloop, world, peer, variant & 1, chain, staging, ring, info);
                    if (rc != 0) { printf("rc %d n %d S %d variant %d\n", rc, n, S, variant); return 1; }
                    ++runs;
                }
        // rings of 4+ slots: tickets map to cells in segment-major groups of slots/2 target blocks (short last group)
        for (int ring : {4, 5, 8})
            for (int order : {0, 1}) {
                int rc = emu_step_allpairs(in.data(), out.data(), mirror.data(), n, n, 0.1f, 1.f, 32, 1, 32, 1, 0, 1, 0, order, 2048, 0, ring, info);
                if (rc != 0) { printf("ring rc %d n %d ring %d\n", rc, n, ring); return 1; }
                ++runs;
            }
        emu_step_well(in.data(), out.data(), mirror.data(), packed.data(), n, n, 0.1f, 1.f, 0, n);
        // chained steps (per-target-block flags) and the initial-conditions kernel, exactly-sized buffers
        for (auto &sh : shapes) {
            if (n > 1100 && sh[0] * sh[1] < 128) continue;
            std::vector<PV> state = in;
            int rc = emu_steps_chained(state.data(), n, 3, 0.1f, 1.f, 32, sh[0], sh[1], 0);
            if (rc != 0) { printf("chained rc %d n %d shape %d %d\n", rc, n, sh[0], sh[1]); return 1; }
            ++runs;
        }
        {
            const unsigned first = n >= 4 ? n / 4 : 0, count = n >= 4 ? n / 2 : n;
            std::vector<PV> a(count), b(count); std::vector<float> pa(4*n), pb(4*n);
            emu_init_particles(a.data(), b.data(), pa.data(), pb.data(), n, first, count, 99u);
            ++runs;
        }
    }
    printf("asan emulation runs: %d ok\n", runs);
    return 0;
}
