/*
 * oracle/oracle.h -- CPU restatement of the Multi-Adapter-Particles n-body step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / reported CPU baseline.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN CODE.  The reference ships no tests, golden
 * vectors or fixtures for this path and runs its shader as HLSL
 * cs_5_0 through D3D12 on Windows (Particles/Compute.cpp:488-511), which cannot execute
 * here.  But the shader source itself compiles with g++ behind a small type shim: the
 * Makefile builds Particles/nBodyGravityCS.hlsl, unmodified in every arithmetic line, into
 * oracle/_ref/libref_shader.so (hlsl_shim.hpp, lower_hlsl.py, ref_harness.cpp), and the
 * LITERAL flavour below reproduces its bodyBodyInteraction, its CSMain and all-pairs steps
 * built from them BIT FOR BIT (tests/test_reference_shader.py; outputs committed as
 * tests/golden/ref_shader_vectors.npz so the pin travels without /root/reference).
 * What that pin cannot cover is how a D3D12 driver rounds the same HLSL (mad contraction,
 * rsq approximation): that latitude is what the 1e-5 tolerance is for.  Further pins:
 * closed-form known-answer tests (tests/test_oracle_kat.py), an independent numpy-float32
 * transcription (tests/test_oracle_numpy.py) and an fp64 direct sum.
 *
 * What it follows (all paths relative to /root/reference):
 *   Particles/nBodyGravityCS.hlsl:37-38   softeningSquared = 25, g_fParticleMass = 70000
 *   Particles/nBodyGravityCS.hlsl:44-57   bodyBodyInteraction (pair math, operation order)
 *   Particles/nBodyGravityCS.hlsl:88-89   load pos / vel
 *   Particles/nBodyGravityCS.hlsl:92-101  literal central-well acceleration (CSMain as shipped)
 *   Particles/nBodyGravityCS.hlsl:103-108 kick, damp, drift, pos.w = length(accel)
 *   Particles/Compute.cpp:542-546         N, dimx = ceil(N/64), dt = 0.1f, damping = 1.0f
 *   Particles/Compute.cpp:1041            Dispatch(ceil(nActive/64)) -> which bodies are updated
 *   Particles/defines.h:37                BLOCK_SIZE 64 (j tile)
 *
 * Canonical summation order (a build decision, frozen; the reference pins only "ascending j, tiles of
 * 64"): the sources j in [0, n_sources) are cut into S = 32 contiguous, tile-aligned segments; a segment
 * is taken as consecutive chains of MAPO_CHAIN_SOURCES = 2,048 sources counted from its first source
 * (the last one shorter); inside a chain one fp32 accumulator per axis runs over ascending j; the chain
 * sums of a segment are folded left to right into the segment's partial, (c0 + c1) + c2 ..., and the S
 * partials left to right into the acceleration, ((p0 + p1) + p2) + ... + p(S-1).
 */
#ifndef MAPO_ORACLE_H
#define MAPO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float pos[4]; float velo[4]; } mapo_posvelo; /* ParticleShared.hlsl:12-16 */

#define MAPO_SOFTENING_SQUARED 25.0f   /* nBodyGravityCS.hlsl:37 */
#define MAPO_PARTICLE_MASS     70000.0f /* nBodyGravityCS.hlsl:38 */
#define MAPO_TILE              64      /* defines.h:37 */
#define MAPO_SEGMENTS          32      /* canonical segment count, every N */
#define MAPO_CHAIN_SOURCES     2048    /* canonical chain length */

enum { MAPO_LITERAL = 0, MAPO_MIRRORED = 1 };

/* dimx of Compute.cpp:544 */
int  mapo_num_tiles(int n);
/* canonical segment count (32 for every n) and chain length (2,048 sources): see mapc_plan_segments */
int  mapo_default_segments(int n);
int  mapo_default_chain(void);
/* j range [j0, j1) of segment s out of S over n_sources sources */
void mapo_segment_range(int n_sources, int S, int s, int *j0, int *j1);
/* number of bodies one Simulate(n_active) updates: min(n, 64*ceil(n_active/64)) (Compute.cpp:1041) */
int  mapo_num_targets(int n, int n_active);

/* nBodyGravityCS.hlsl:44-57, written exactly as the shader is (separate mul/add, 1.0f/sqrtf). */
void mapo_body_body_interaction(float ai[3], const float bj[4], const float bi[4],
                                float mass, int particles);
/* same pair with the contractions the CUDA kernel uses (explicit fmaf, softening folded
 * into the first fma of the dot product); 1.0f/sqrtf stands in for MUFU.RSQ.  Like the kernel it
 * leaves the uniform mass factor OUT: the caller scales each segment partial by the mass once. */
void mapo_body_body_interaction_mirrored(float ai[3], const float bj[4], const float bi[4],
                                         float mass);

/* accel of the listed targets (NULL = targets 0..n_targets-1) from sources [0, n_sources), S segments,
 * chains of `chain` sources (0 = one chain per segment), one scalar call of the pair function per (i, j).
 * Slow; small N. */
void mapo_accel_allpairs_scalar(const mapo_posvelo *in, int n_sources, int S, int chain, int flavour,
                                const int *targets, int n_targets, float *accel3);
/* same numbers (bit-identical) with the canonical chain length, 8 targets per inner loop so gcc can
 * vectorise across i; OpenMP over target blocks; threads <= 0 means omp_get_max_threads(). */
void mapo_accel_allpairs(const mapo_posvelo *in, int n_sources, int S, int flavour,
                         const int *targets, int n_targets, float *accel3, int threads);
/* the same with chains of `chunk` sources instead of MAPO_CHAIN_SOURCES (0 = one chain per segment) */
void mapo_accel_allpairs_chunked(const mapo_posvelo *in, int n_sources, int S, int chunk, int flavour,
                                 const int *targets, int n_targets, float *accel3, int threads);
/* fp64 direct sum, ascending j, no segments -- reported alongside, never gating */
void mapo_accel_fp64(const mapo_posvelo *in, int n_sources,
                     const int *targets, int n_targets, double *accel3, int threads);

/* nBodyGravityCS.hlsl:103-108 for one body */
void mapo_integrate(const mapo_posvelo *in_i, const float accel[3], float dt, float damping,
                    int flavour, mapo_posvelo *out_i);

/* one all-pairs step: out[i] for i < mapo_num_targets(n, n_active) is written, the rest of
 * out is left untouched (the stale side of the ping-pong, as in the reference). */
void mapo_step_allpairs(const mapo_posvelo *in, mapo_posvelo *out, int n, int n_active,
                        float dt, float damping, int S, int flavour, int threads);
/* subsampled step: out_targets[k] is the new state of body targets[k] */
void mapo_step_allpairs_targets(const mapo_posvelo *in, int n_sources,
                                const int *targets, int n_targets, float dt, float damping,
                                int S, int flavour, int threads, mapo_posvelo *out_targets);
void mapo_step_allpairs_targets_chunked(const mapo_posvelo *in, int n_sources,
                                        const int *targets, int n_targets, float dt, float damping,
                                        int S, int chunk, int flavour, int threads, mapo_posvelo *out_targets);
/* the step the reference actually executes: origin gravity well, nBodyGravityCS.hlsl:86-109 */
void mapo_step_well(const mapo_posvelo *in, mapo_posvelo *out, int n, int n_active,
                    float dt, float damping, int flavour);

/* InitializeParticles / LoadParticles (Compute.cpp:596-609 fast_rand, :719-749 USE_SCALAR_OPTIMIZED branch,
 * :831-844 two groups around x = +-0.75 * ParticleSpread), with one LCG stream per particle seeded from
 * (seed, i) -- the reference's generator is unseeded and thread-schedule dependent -- and exact
 * normalisations where the reference uses the *Est forms.  Plain IEEE float ops in the written order. */
void mapo_init_particles(mapo_posvelo *out, unsigned n, unsigned seed);

int  mapo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
