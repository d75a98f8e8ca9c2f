// tools/mapc_run.cpp -- headless driver: what Particles.{h,cpp} + Main-Particles.cpp do in the
// reference, without the window, GUI and renderer.  It owns one Compute (producer) and one
// HeadlessRender (consumer) and runs the reference's frame loop (Particles/Particles.cpp:446-456):
//
//     fence = compute.GetFenceValue();  render.Draw(nDraw, fence, nCopy);  compute.Simulate(nSim, fence);
//
// then dumps the final PosVelo state.  Plain C++17 over the C ABI (include/mapc_compute.hpp); all
// CUDA lives behind libmapc.so.
//
// Options (the reference's ArgParser tokens keep their names, Particles.cpp:248-270):
//   --numparticles N   number of bodies (default 262144 = MIN_NUM_PARTICLES)
//   --numSim n / --numCopy n / --numDraw n   bodies simulated / copied / "drawn" per frame (default N)
//   --steps K          frames to run (default 100)
//   --dt f --damping f (defaults 0.1, 1.0: Compute.cpp:545-546)
//   --mode allpairs|well     force (default allpairs; well = the kernel the reference dispatches)
//   --ic shells|sphere       initial conditions (shells = InitializeParticles, Compute.cpp:820-844)
//   --radius f --seed s      sphere radius (default 8000) and RNG seed
//   --load file / --dump file   raw little-endian PosVelo[N] (32 bytes per body) in / out
//   --device d --render-device d   CUDA devices of producer and consumer (may differ: peer copy)
//   --no-consumer      skip the consumer (pure simulation loop)
//   --async            the reference's async mode: consumer on the producer's device, no copies
//                      (Particles.cpp:202-207 chooses it when render and compute adapter are the same)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "mapc_compute.hpp"

namespace {

struct Options {
    std::uint32_t n = MAPC_MIN_NUM_PARTICLES;
    int num_sim = -1, num_copy = -1, num_draw = -1;
    int steps = 100;
    float dt = MAPC_DEFAULT_DELTA_TIME, damping = MAPC_DEFAULT_DAMPING;
    std::string mode = "allpairs", ic = "sphere", load, dump;
    float radius = 8000.f;
    std::uint32_t seed = 1;
    int device = 0, render_device = -1;
    bool consumer = true;
    bool async_mode = false;
};

[[noreturn]] void usage(const char *msg)
{
    std::fprintf(stderr, "mapc_run: %s\nsee the header of tools/mapc_run.cpp for the options\n", msg);
    std::exit(2);
}

Options parse(int argc, char **argv)
{
    Options o;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> const char * {
            if (i + 1 >= argc) usage(("missing value after " + a).c_str());
            return argv[++i];
        };
        if (a == "--numparticles" || a == "numparticles") o.n = (std::uint32_t)std::strtoul(val(), nullptr, 10);
        else if (a == "--numSim" || a == "numSim") o.num_sim = std::atoi(val());
        else if (a == "--numCopy" || a == "numCopy") o.num_copy = std::atoi(val());
        else if (a == "--numDraw" || a == "numDraw") o.num_draw = std::atoi(val());
        else if (a == "--steps") o.steps = std::atoi(val());
        else if (a == "--dt") o.dt = (float)std::atof(val());
        else if (a == "--damping") o.damping = (float)std::atof(val());
        else if (a == "--mode") o.mode = val();
        else if (a == "--ic") o.ic = val();
        else if (a == "--radius") o.radius = (float)std::atof(val());
        else if (a == "--seed") o.seed = (std::uint32_t)std::strtoul(val(), nullptr, 10);
        else if (a == "--load") o.load = val();
        else if (a == "--dump") o.dump = val();
        else if (a == "--device") o.device = std::atoi(val());
        else if (a == "--render-device") o.render_device = std::atoi(val());
        else if (a == "--no-consumer") o.consumer = false;
        else if (a == "--async") o.async_mode = true;
        else usage(("unknown option " + a).c_str());
    }
    if (o.num_sim < 0) o.num_sim = (int)o.n;
    if (o.num_copy < 0) o.num_copy = (int)o.n;
    if (o.num_draw < 0) o.num_draw = (int)o.n;
    if (o.render_device < 0) o.render_device = o.device;
    return o;
}

// uniform-in-volume sphere from a splitmix64 counter stream (same recipe as ic.py: uniform_sphere)
std::uint64_t splitmix64(std::uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    std::uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

double uniform(std::uint64_t seed, int stream, std::uint64_t k)
{
    const std::uint64_t key = splitmix64(seed * 0x100000001B3ull + (std::uint64_t)(stream + 1));
    const std::uint64_t bits = splitmix64(k ^ key);
    return ((double)(bits >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

std::vector<mapc_posvelo> uniform_sphere(std::uint32_t n, double radius, std::uint64_t seed)
{
    std::vector<mapc_posvelo> p(n);
    const double two_pi = 6.283185307179586476925286766559;
    for (std::uint32_t i = 0; i < n; ++i) {
        const double r = radius * std::cbrt(uniform(seed, 0, i));
        const double z = 2.0 * uniform(seed, 1, i) - 1.0;
        const double s = std::sqrt(std::fmax(0.0, 1.0 - z * z));
        const double phi = two_pi * uniform(seed, 2, i);
        p[i].pos[0] = (float)(s * std::cos(phi) * r);
        p[i].pos[1] = (float)(s * std::sin(phi) * r);
        p[i].pos[2] = (float)(z * r);
        p[i].pos[3] = 0.f;
        p[i].velo[0] = p[i].velo[1] = p[i].velo[2] = p[i].velo[3] = 0.f;
    }
    return p;
}

}  // namespace

int main(int argc, char **argv)
{
    const Options o = parse(argc, argv);
    try {
        mapc::Compute compute(o.n, o.device);
        if (o.mode == "well") compute.SetForceMode(MAPC_FORCE_WELL);
        else if (o.mode != "allpairs") usage("--mode must be allpairs or well");

        if (!o.load.empty()) {
            std::vector<mapc_posvelo> p(o.n);
            std::ifstream f(o.load, std::ios::binary);
            if (!f.read(reinterpret_cast<char *>(p.data()), (std::streamsize)(p.size() * sizeof(mapc_posvelo))))
                usage("cannot read --load file (need numparticles * 32 bytes)");
            compute.Upload(p.data(), o.n);
        } else if (o.ic == "shells") {
            compute.InitializeParticles(o.seed);
        } else if (o.ic == "sphere") {
            const auto p = uniform_sphere(o.n, o.radius, o.seed);
            compute.Upload(p.data(), o.n);
        } else {
            usage("--ic must be shells or sphere");
        }

        std::unique_ptr<mapc::HeadlessRender> render;
        if (o.consumer) render = std::make_unique<mapc::HeadlessRender>(compute, o.render_device, o.async_mode);

        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < o.steps; ++k) {
            std::uint64_t fence = compute.GetFenceValue();                  // Particles.cpp:446
            if (render) render->Draw(o.num_draw, fence, o.num_copy);         // Particles.cpp:447
            else fence = 0;
            compute.Simulate(o.num_sim, fence, o.dt, o.damping);             // Particles.cpp:448
        }
        compute.WaitForGpu();
        if (render) render->WaitForGpu();
        const double wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        const float gpu_ms = compute.GetGpuTimes()[0].first * 1e3f;
        std::vector<mapc_posvelo> out(o.n);
        compute.Download(out.data(), 0, o.n);
        double checksum = 0.0;
        for (const auto &b : out) checksum += std::fabs(b.pos[0]) + std::fabs(b.pos[1]) + std::fabs(b.pos[2]);
        std::uint64_t frame = 0;
        std::uint32_t count = 0;
        if (render) render->Latest(&frame, &count);

        if (!o.dump.empty()) {
            std::ofstream f(o.dump, std::ios::binary);
            f.write(reinterpret_cast<const char *>(out.data()), (std::streamsize)(out.size() * sizeof(mapc_posvelo)));
            std::ofstream j(o.dump + ".json");
            j << "{\"format\": \"PosVelo float32[N][8] little endian\", \"n\": " << o.n << ", \"steps\": " << o.steps
              << ", \"dt\": " << o.dt << ", \"damping\": " << o.damping << ", \"mode\": \"" << o.mode << "\"}\n";
        }
        const double pairs = o.mode == "well" ? (double)o.num_sim : (double)o.num_sim * (double)o.num_sim;
        std::printf("{\"n\": %u, \"steps\": %d, \"mode\": \"%s\", \"wall_ms_per_step\": %.4f, \"gpu_ms_per_step_ema\": %.4f, "
                    "\"g_interactions_per_s\": %.2f, \"consumer_frame\": %llu, \"checksum\": %.6e}\n",
                    o.n, o.steps, o.mode.c_str(), wall_s * 1e3 / o.steps, gpu_ms,
                    pairs / (wall_s / o.steps) / 1e9, (unsigned long long)frame, checksum);
    } catch (const mapc::Error &e) {
        std::fprintf(stderr, "mapc_run: error %d: %s\n", (int)e.status, e.what());
        return 1;
    }
    return 0;
}
