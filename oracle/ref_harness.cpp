// ref_harness.cpp -- C entry points around the reference shader's own functions.  TEST INFRASTRUCTURE ONLY.
//
// This file is the tail of one translation unit: oracle/Makefile concatenates hlsl_shim.hpp, the
// lowered Particles/nBodyGravityCS.hlsl (read where it lies under /root/reference, never copied into
// the repository) and this file, and compiles them into oracle/_ref/libref_shader.so.  Above this
// line the compiler has therefore seen the reference's bodyBodyInteraction (:44-57), CSMain
// (:86-109), its constants (:37-38) and its buffer / cbuffer globals.
//
// What is the reference's and what is ours:
//   ref_body_body_interaction, ref_csmain, ref_constants   pure reference code
//   ref_accel_allpairs / ref_step_allpairs                 OUR canonical loop (32 segments of 64-body
//       tiles, chains of 2,048 sources folded left to right into the segment's partial, partials
//       summed left to right: DESIGN.md section 3) around the reference's
//       bodyBodyInteraction, and the reference's integration lines (:103-108) restated because
//       CSMain fuses them with the gravity-well force.  The shipped CSMain has no all-pairs loop.
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int kTile = 64;   // Particles/defines.h:37 BLOCK_SIZE

void segment_range(int n_sources, int S, int s, int &j0, int &j1)
{
    const long long tiles = (n_sources + kTile - 1) / kTile;
    long long a = (tiles * s) / S * kTile, b = (tiles * (s + 1)) / S * kTile;
    if (a > n_sources) a = n_sources;
    if (b > n_sources) b = n_sources;
    j0 = (int)a;
    j1 = (int)b;
}

// chain: sources per sequential accumulation (0 = one chain per segment)
float3 accel_of(const float *posvelo, int n_sources, int S, int chain, int i)
{
    const float *pi = posvelo + 8 * (size_t)i;
    const float4 bi(pi[0], pi[1], pi[2], pi[3]);
    float3 total;
    for (int s = 0; s < S; ++s) {
        int j0, j1;
        segment_range(n_sources, S, s, j0, j1);
        const int step = (chain > 0 && j1 - j0 > chain) ? chain : (j1 - j0 > 0 ? j1 - j0 : 1);
        float3 partial;
        for (int c0 = j0; c0 < j1; c0 += step) {
            const int c1 = c0 + step < j1 ? c0 + step : j1;
            float3 sum;
            for (int j = c0; j < c1; ++j) {
                const float *pj = posvelo + 8 * (size_t)j;
                bodyBodyInteraction(sum, float4(pj[0], pj[1], pj[2], pj[3]), bi, g_fParticleMass, 1);
            }
            if (c0 == j0) partial = sum;
            else partial += sum;
        }
        total += partial;
    }
    return total;
}

}  // namespace

extern "C" {

void ref_constants(float *softening_squared, float *particle_mass)
{
    *softening_squared = softeningSquared;
    *particle_mass = g_fParticleMass;
}

void ref_body_body_interaction(float ai[3], const float bj[4], const float bi[4], float mass, int particles)
{
    float3 a(ai[0], ai[1], ai[2]);
    bodyBodyInteraction(a, float4(bj[0], bj[1], bj[2], bj[3]), float4(bi[0], bi[1], bi[2], bi[3]), mass, particles);
    ai[0] = a.x;
    ai[1] = a.y;
    ai[2] = a.z;
}

// Dispatch CSMain for DTid.x in [0, n_dispatch) over PosVelo arrays (8 floats per body).
void ref_csmain(const float *in_posvelo, float *out_posvelo, int n_dispatch, float dt, float damping)
{
    std::vector<Position> old_pos(n_dispatch), new_pos(n_dispatch);
    std::vector<Velocity> old_vel(n_dispatch), new_vel(n_dispatch);
    for (int i = 0; i < n_dispatch; ++i) {
        const float *p = in_posvelo + 8 * (size_t)i;
        old_pos[i].pos = float4(p[0], p[1], p[2], p[3]);
        old_vel[i].velocity = float3(p[4], p[5], p[6]);
    }
    oldPosition.data = old_pos.data();
    newPosition.data = new_pos.data();
    oldVelocity.data = old_vel.data();
    newVelocity.data = new_vel.data();
    g_paramf = float4(dt, damping, 0.f, 0.f);   // Compute.cpp:545-546
    for (int i = 0; i < n_dispatch; ++i) {
        uint3 id = {(uint)i, 0u, 0u};
        CSMain(id);
    }
    for (int i = 0; i < n_dispatch; ++i) {
        float *o = out_posvelo + 8 * (size_t)i;
        o[0] = new_pos[i].pos.x; o[1] = new_pos[i].pos.y; o[2] = new_pos[i].pos.z; o[3] = new_pos[i].pos.w;
        o[4] = new_vel[i].velocity.x; o[5] = new_vel[i].velocity.y; o[6] = new_vel[i].velocity.z;
        o[7] = 0.f;                              // Velocity is a bare float3 (:72-75)
    }
}

void ref_accel_allpairs(const float *posvelo, int n_sources, int S, int chain, const int *targets, int n_targets,
                        float *accel3, int threads)
{
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
    for (int k = 0; k < n_targets; ++k) {
        const float3 a = accel_of(posvelo, n_sources, S, chain, targets ? targets[k] : k);
        accel3[3 * (size_t)k + 0] = a.x;
        accel3[3 * (size_t)k + 1] = a.y;
        accel3[3 * (size_t)k + 2] = a.z;
    }
}

// One all-pairs step of n_targets bodies (indices `targets`, or the first n_targets when null) against
// the first n_sources; out_targets[k] receives the new state of target k.
void ref_step_allpairs_targets(const float *in_posvelo, int n_sources, const int *targets, int n_targets, int S,
                               int chain, float dt, float damping, int threads, float *out_targets)
{
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
    for (int k = 0; k < n_targets; ++k) {
        const int i = targets ? targets[k] : k;
        const float *p = in_posvelo + 8 * (size_t)i;
        float4 pos(p[0], p[1], p[2], p[3]);
        float3 vel(p[4], p[5], p[6]);
        const float3 accel = accel_of(in_posvelo, n_sources, S, chain, i);
        const float4 paramf(dt, damping, 0.f, 0.f);
        vel.xyz += accel.xyz * paramf.x;          // nBodyGravityCS.hlsl:103
        vel.xyz *= paramf.y;                      // :104
        pos.xyz += vel.xyz * paramf.x;            // :105
        const float4 out_pos = float4(pos.xyz, length(accel));   // :107
        float *o = out_targets + 8 * (size_t)k;
        o[0] = out_pos.x; o[1] = out_pos.y; o[2] = out_pos.z; o[3] = out_pos.w;
        o[4] = vel.x; o[5] = vel.y; o[6] = vel.z; o[7] = 0.f;    // :108
    }
}

}  // extern "C"
