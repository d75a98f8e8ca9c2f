/*
 * oracle/oracle.c -- see oracle.h.  TEST INFRASTRUCTURE ONLY.  Pinned bit for bit against the
 * reference's own shader code compiled for the CPU (oracle/_ref, tests/test_reference_shader.py),
 * plus closed-form KATs and an fp64 direct sum.
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off matters: the LITERAL flavour must keep every mul and add separate, and
 * the MIRRORED flavour asks for each fused op explicitly with fmaf().
 */
#include "oracle.h"

#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LANES 8

int mapo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Compute.cpp:544  param[1] = int(ceil(N / float(BLOCK_SIZE))) */
int mapo_num_tiles(int n) { return (n + MAPO_TILE - 1) / MAPO_TILE; }

int mapo_default_segments(int n)
{
    (void)n;
    return MAPO_SEGMENTS;
}

int mapo_default_chain(void) { return MAPO_CHAIN_SOURCES; }

void mapo_segment_range(int n_sources, int S, int s, int *j0, int *j1)
{
    const long long tiles = mapo_num_tiles(n_sources);
    long long a = (tiles * s) / S * MAPO_TILE;
    long long b = (tiles * (s + 1)) / S * MAPO_TILE;
    if (a > n_sources) a = n_sources;
    if (b > n_sources) b = n_sources;
    *j0 = (int)a;
    *j1 = (int)b;
}

/* Compute.cpp:1041 dispatches ceil(nActive/64) groups of 64 threads, one body per thread and
 * no bounds check in the shader; writes past the buffer are dropped by D3D. */
int mapo_num_targets(int n, int n_active)
{
    if (n_active <= 0) return 0;
    long long t = (long long)mapo_num_tiles(n_active) * MAPO_TILE;
    return (int)(t < n ? t : n);
}

/* nBodyGravityCS.hlsl:44-57 */
void mapo_body_body_interaction(float ai[3], const float bj[4], const float bi[4],
                                float mass, int particles)
{
    float r[3];
    r[0] = bj[0] - bi[0];                                     /* :46 */
    r[1] = bj[1] - bi[1];
    r[2] = bj[2] - bi[2];
    float distSqr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];  /* :48 dot(r, r) */
    distSqr += MAPO_SOFTENING_SQUARED;                        /* :49 */
    float invDist = 1.0f / sqrtf(distSqr);                    /* :51 */
    float invDistCube = invDist * invDist * invDist;          /* :52 left-assoc */
    float s = mass * invDistCube * (float)particles;          /* :54 */
    ai[0] += r[0] * s;                                        /* :56 */
    ai[1] += r[1] * s;
    ai[2] += r[2] * s;
}

void mapo_body_body_interaction_mirrored(float ai[3], const float bj[4], const float bi[4],
                                         float mass)
{
    float dx = bj[0] - bi[0];
    float dy = bj[1] - bi[1];
    float dz = bj[2] - bi[2];
    float d2 = fmaf(dx, dx, MAPO_SOFTENING_SQUARED);
    d2 = fmaf(dy, dy, d2);
    d2 = fmaf(dz, dz, d2);
    float inv = 1.0f / sqrtf(d2);
    float inv2 = inv * inv;
    float inv3 = inv2 * inv;
    /* the kernel leaves the uniform mass out of the pair and scales each segment partial once;
     * `mass` is applied by the caller (mapo_accel_*: partial *= mass) */
    (void)mass;
    ai[0] = fmaf(dx, inv3, ai[0]);
    ai[1] = fmaf(dy, inv3, ai[1]);
    ai[2] = fmaf(dz, inv3, ai[2]);
}

void mapo_accel_allpairs_scalar(const mapo_posvelo *in, int n_sources, int S, int chain, int flavour,
                                const int *targets, int n_targets, float *accel3)
{
    for (int k = 0; k < n_targets; ++k) {
        const int i = targets ? targets[k] : k;
        float total[3] = {0.f, 0.f, 0.f};
        for (int s = 0; s < S; ++s) {
            int j0, j1;
            mapo_segment_range(n_sources, S, s, &j0, &j1);
            float p[3] = {0.f, 0.f, 0.f};                     /* the segment's partial */
            const int step = (chain > 0 && j1 - j0 > chain) ? chain : (j1 - j0 > 0 ? j1 - j0 : 1);
            for (int c0 = j0; c0 < j1 || c0 == j0; c0 += step) {
                const int c1 = (c0 + step < j1) ? c0 + step : j1;
                float c[3] = {0.f, 0.f, 0.f};                 /* one chain */
                for (int j = c0; j < c1; ++j) {
                    if (flavour == MAPO_LITERAL)
                        mapo_body_body_interaction(c, in[j].pos, in[i].pos, MAPO_PARTICLE_MASS, 1);
                    else
                        mapo_body_body_interaction_mirrored(c, in[j].pos, in[i].pos, MAPO_PARTICLE_MASS);
                }
                if (flavour != MAPO_LITERAL) {                /* once per chain sum */
                    c[0] *= MAPO_PARTICLE_MASS; c[1] *= MAPO_PARTICLE_MASS; c[2] *= MAPO_PARTICLE_MASS;
                }
                if (c0 == j0) { p[0] = c[0]; p[1] = c[1]; p[2] = c[2]; }
                else { p[0] += c[0]; p[1] += c[1]; p[2] += c[2]; }
                if (c1 >= j1) break;
            }
            total[0] += p[0];
            total[1] += p[1];
            total[2] += p[2];
        }
        accel3[3 * k + 0] = total[0];
        accel3[3 * k + 1] = total[1];
        accel3[3 * k + 2] = total[2];
    }
}

/* One block of up to LANES targets against sources [j0, j1): each lane is an independent
 * ascending-j chain, so vectorising across lanes leaves every rounding where the scalar
 * code has it. */
static void segment_block_literal(const mapo_posvelo *in, int j0, int j1,
                                  const float *xi, const float *yi, const float *zi,
                                  float *px, float *py, float *pz)
{
    float ax[LANES], ay[LANES], az[LANES];
    for (int l = 0; l < LANES; ++l) ax[l] = ay[l] = az[l] = 0.f;
    const float mass = MAPO_PARTICLE_MASS;
    for (int j = j0; j < j1; ++j) {
        const float xj = in[j].pos[0], yj = in[j].pos[1], zj = in[j].pos[2];
#pragma omp simd
        for (int l = 0; l < LANES; ++l) {
            float rx = xj - xi[l];
            float ry = yj - yi[l];
            float rz = zj - zi[l];
            float distSqr = rx * rx + ry * ry + rz * rz;
            distSqr += MAPO_SOFTENING_SQUARED;
            float invDist = 1.0f / sqrtf(distSqr);
            float invDistCube = invDist * invDist * invDist;
            float s = mass * invDistCube * 1.0f;
            ax[l] += rx * s;
            ay[l] += ry * s;
            az[l] += rz * s;
        }
    }
    for (int l = 0; l < LANES; ++l) { px[l] = ax[l]; py[l] = ay[l]; pz[l] = az[l]; }
}

static void segment_block_mirrored(const mapo_posvelo *in, int j0, int j1,
                                   const float *xi, const float *yi, const float *zi,
                                   float *px, float *py, float *pz)
{
    float ax[LANES], ay[LANES], az[LANES];
    for (int l = 0; l < LANES; ++l) ax[l] = ay[l] = az[l] = 0.f;
    const float mass = MAPO_PARTICLE_MASS;
    for (int j = j0; j < j1; ++j) {
        const float xj = in[j].pos[0], yj = in[j].pos[1], zj = in[j].pos[2];
#pragma omp simd
        for (int l = 0; l < LANES; ++l) {
            float dx = xj - xi[l];
            float dy = yj - yi[l];
            float dz = zj - zi[l];
            float d2 = fmaf(dx, dx, MAPO_SOFTENING_SQUARED);
            d2 = fmaf(dy, dy, d2);
            d2 = fmaf(dz, dz, d2);
            float inv = 1.0f / sqrtf(d2);
            float inv2 = inv * inv;
            float inv3 = inv2 * inv;
            ax[l] = fmaf(dx, inv3, ax[l]);
            ay[l] = fmaf(dy, inv3, ay[l]);
            az[l] = fmaf(dz, inv3, az[l]);
        }
    }
    for (int l = 0; l < LANES; ++l) { px[l] = ax[l] * mass; py[l] = ay[l] * mass; pz[l] = az[l] * mass; }
}

void mapo_accel_allpairs(const mapo_posvelo *in, int n_sources, int S, int flavour,
                         const int *targets, int n_targets, float *accel3, int threads)
{
    mapo_accel_allpairs_chunked(in, n_sources, S, MAPO_CHAIN_SOURCES, flavour, targets, n_targets, accel3, threads);
}

/* The canonical order (csrc/nbody_kernels.cuh, template parameter CHAIN): a segment longer than `chunk`
 * sources is evaluated as consecutive chains of `chunk` sources counted from the segment's first source;
 * every chain is one sequential accumulation (scaled by the mass once, in the MIRRORED flavour) and the
 * chain sums are folded left to right, ((c0 + c1) + c2) + ..., into the segment's partial.  chunk == 0: one
 * chain per segment (not what the product does; kept to show what bounded chains buy). */
void mapo_accel_allpairs_chunked(const mapo_posvelo *in, int n_sources, int S, int chunk, int flavour,
                                 const int *targets, int n_targets, float *accel3, int threads)
{
    const int blocks = (n_targets + LANES - 1) / LANES;
    if (threads <= 0) threads = mapo_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (int b = 0; b < blocks; ++b) {
        float xi[LANES], yi[LANES], zi[LANES];
        float tx[LANES], ty[LANES], tz[LANES];
        int idx[LANES];
        for (int l = 0; l < LANES; ++l) {
            int k = b * LANES + l;
            if (k >= n_targets) k = n_targets - 1;            /* duplicate lane, never stored */
            idx[l] = targets ? targets[k] : k;
            xi[l] = in[idx[l]].pos[0];
            yi[l] = in[idx[l]].pos[1];
            zi[l] = in[idx[l]].pos[2];
            tx[l] = ty[l] = tz[l] = 0.f;
        }
        for (int s = 0; s < S; ++s) {
            int j0, j1;
            float px[LANES] = {0.f}, py[LANES] = {0.f}, pz[LANES] = {0.f};
            mapo_segment_range(n_sources, S, s, &j0, &j1);
            const int step = (chunk > 0 && j1 - j0 > chunk) ? chunk : (j1 - j0 > 0 ? j1 - j0 : 1);
            int c0 = j0;
            do {                                              /* one pass unless the segment is chunked */
                const int c1 = (c0 + step < j1) ? c0 + step : j1;
                float cx[LANES], cy[LANES], cz[LANES];
                if (flavour == MAPO_LITERAL)
                    segment_block_literal(in, c0, c1, xi, yi, zi, cx, cy, cz);
                else
                    segment_block_mirrored(in, c0, c1, xi, yi, zi, cx, cy, cz);
                for (int l = 0; l < LANES; ++l) {
                    if (c0 == j0) { px[l] = cx[l]; py[l] = cy[l]; pz[l] = cz[l]; }
                    else { px[l] += cx[l]; py[l] += cy[l]; pz[l] += cz[l]; }
                }
                c0 = c1;
            } while (c0 < j1);
            for (int l = 0; l < LANES; ++l) { tx[l] += px[l]; ty[l] += py[l]; tz[l] += pz[l]; }
        }
        for (int l = 0; l < LANES; ++l) {
            const int k = b * LANES + l;
            if (k < n_targets) {
                accel3[3 * (size_t)k + 0] = tx[l];
                accel3[3 * (size_t)k + 1] = ty[l];
                accel3[3 * (size_t)k + 2] = tz[l];
            }
        }
    }
}

void mapo_accel_fp64(const mapo_posvelo *in, int n_sources,
                     const int *targets, int n_targets, double *accel3, int threads)
{
    if (threads <= 0) threads = mapo_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
    for (int k = 0; k < n_targets; ++k) {
        const int i = targets ? targets[k] : k;
        const double xi = in[i].pos[0], yi = in[i].pos[1], zi = in[i].pos[2];
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int j = 0; j < n_sources; ++j) {
            const double rx = (double)in[j].pos[0] - xi;
            const double ry = (double)in[j].pos[1] - yi;
            const double rz = (double)in[j].pos[2] - zi;
            const double d2 = rx * rx + ry * ry + rz * rz + (double)MAPO_SOFTENING_SQUARED;
            const double inv = 1.0 / sqrt(d2);
            const double s = (double)MAPO_PARTICLE_MASS * inv * inv * inv;
            ax += rx * s;
            ay += ry * s;
            az += rz * s;
        }
        accel3[3 * (size_t)k + 0] = ax;
        accel3[3 * (size_t)k + 1] = ay;
        accel3[3 * (size_t)k + 2] = az;
    }
}

/* nBodyGravityCS.hlsl:103-108 */
void mapo_integrate(const mapo_posvelo *in_i, const float accel[3], float dt, float damping,
                    int flavour, mapo_posvelo *out_i)
{
    float pos[3], vel[3];
    for (int c = 0; c < 3; ++c) { pos[c] = in_i->pos[c]; vel[c] = in_i->velo[c]; }
    float len;
    if (flavour == MAPO_LITERAL) {
        for (int c = 0; c < 3; ++c) {
            vel[c] += accel[c] * dt;                          /* :103 */
            vel[c] *= damping;                                /* :104 */
            pos[c] += vel[c] * dt;                            /* :105 */
        }
        len = sqrtf(accel[0] * accel[0] + accel[1] * accel[1] + accel[2] * accel[2]);
    } else {
        for (int c = 0; c < 3; ++c) {
            vel[c] = fmaf(accel[c], dt, vel[c]);
            vel[c] *= damping;
            pos[c] = fmaf(vel[c], dt, pos[c]);
        }
        len = sqrtf(fmaf(accel[2], accel[2], fmaf(accel[1], accel[1], accel[0] * accel[0])));
    }
    out_i->pos[0] = pos[0]; out_i->pos[1] = pos[1]; out_i->pos[2] = pos[2];
    out_i->pos[3] = len;                                      /* :107 pos.w = length(accel) */
    out_i->velo[0] = vel[0]; out_i->velo[1] = vel[1]; out_i->velo[2] = vel[2];
    out_i->velo[3] = 0.0f;                                    /* :108 velocity is a bare float3 */
}

void mapo_step_allpairs_targets(const mapo_posvelo *in, int n_sources,
                                const int *targets, int n_targets, float dt, float damping,
                                int S, int flavour, int threads, mapo_posvelo *out_targets)
{
    mapo_step_allpairs_targets_chunked(in, n_sources, targets, n_targets, dt, damping, S, MAPO_CHAIN_SOURCES, flavour,
                                       threads, out_targets);
}

void mapo_step_allpairs_targets_chunked(const mapo_posvelo *in, int n_sources,
                                        const int *targets, int n_targets, float dt, float damping,
                                        int S, int chunk, int flavour, int threads, mapo_posvelo *out_targets)
{
    enum { BATCH = 4096 };   /* targets per pass: bounds the scratch, has no effect on any sum */
    float buf[3 * BATCH];
    int ids[BATCH];
    for (int base = 0; base < n_targets; base += BATCH) {
        const int cnt = (n_targets - base) < BATCH ? (n_targets - base) : BATCH;
        for (int q = 0; q < cnt; ++q) ids[q] = targets ? targets[base + q] : base + q;
        mapo_accel_allpairs_chunked(in, n_sources, S, chunk, flavour, ids, cnt, buf, threads);
        for (int q = 0; q < cnt; ++q)
            mapo_integrate(&in[ids[q]], &buf[3 * q], dt, damping, flavour, &out_targets[base + q]);
    }
}

void mapo_step_allpairs(const mapo_posvelo *in, mapo_posvelo *out, int n, int n_active,
                        float dt, float damping, int S, int flavour, int threads)
{
    const int n_targets = mapo_num_targets(n, n_active);
    const int n_sources = n_active < n ? n_active : n;
    mapo_step_allpairs_targets(in, n_sources, NULL, n_targets, dt, damping, S, flavour,
                               threads, out);
}

/* nBodyGravityCS.hlsl:86-109, the kernel the reference actually dispatches */
void mapo_step_well(const mapo_posvelo *in, mapo_posvelo *out, int n, int n_active,
                    float dt, float damping, int flavour)
{
    const int n_targets = mapo_num_targets(n, n_active);
    const float mass = MAPO_PARTICLE_MASS;                    /* :90 */
    for (int i = 0; i < n_targets; ++i) {
        const float *r = in[i].pos;                           /* :92 */
        float accel[3];
        if (flavour == MAPO_LITERAL) {
            float distSqr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2]; /* :94 */
            distSqr += MAPO_SOFTENING_SQUARED;                /* :95 */
            float invDist = -1.0f / sqrtf(distSqr);           /* :97 note the minus sign */
            float invDistCube = invDist * invDist * invDist;  /* :98 */
            float s = mass * invDistCube;                     /* :99 */
            accel[0] = r[0] * s; accel[1] = r[1] * s; accel[2] = r[2] * s; /* :101 */
        } else {
            float d2 = fmaf(r[0], r[0], MAPO_SOFTENING_SQUARED);
            d2 = fmaf(r[1], r[1], d2);
            d2 = fmaf(r[2], r[2], d2);
            float inv = -(1.0f / sqrtf(d2));
            float inv3 = inv * inv * inv;
            float s = inv3 * mass;
            accel[0] = r[0] * s; accel[1] = r[1] * s; accel[2] = r[2] * s;
        }
        mapo_integrate(&in[i], accel, dt, damping, flavour, &out[i]);
    }
}

/* ---- initial conditions (Compute.cpp:596-609, :719-749, :820-844) --------------------------------------- */
static unsigned ic_stream_seed(unsigned seed, unsigned i)
{
    unsigned s = seed + 0x9E3779B9u * (i + 1u);   /* one stream per particle (the reference's is per thread, unseeded) */
    s ^= s >> 16;
    s *= 0x85EBCA6Bu;
    s ^= s >> 13;
    s *= 0xC2B2AE35u;
    s ^= s >> 16;
    return s;
}

static float ic_rand_pm1(unsigned *state)
{
    *state = 214013u * *state + 2531011u;                     /* fast_rand(), :605-609 */
    const int r = (int)((*state >> 16) & 0x7FFFu);
    const float k_scale = (1.f / 32767.f) * 2.f;              /* (1.f / RAND_MAX) * 2.f with MSVC's RAND_MAX, :721 */
    return ((float)r * k_scale) - 1.f;                        /* :723-725 */
}

static float ic_dot3(float x, float y, float z) { return (x * x + y * y) + z * z; }

void mapo_init_particles(mapo_posvelo *out, unsigned n, unsigned seed)
{
    const float spread = 400.0f;                              /* defines.h:42 PARTICLE_SPREAD */
    const float speed = 15.0f;                                /* defines.h:39 INITIAL_PARTICLE_SPEED */
    const float center = spread * 0.750f;                     /* Compute.cpp:831 */
    const unsigned half = n / 2u;
    memset(out, 0, (size_t)n * sizeof(*out));                 /* vector::resize value-initialises, :825-829 */
    for (unsigned i = 0; i < 2u * half; ++i) {
        const float cx = i < half ? center : -center;         /* :832-844 */
        unsigned state = ic_stream_seed(seed, i);
        float dx = ic_rand_pm1(&state), dy = ic_rand_pm1(&state), dz = ic_rand_pm1(&state);
        while (ic_dot3(dx, dy, dz) < 10.f) {                  /* :728 */
            dx += ic_rand_pm1(&state);                        /* :730-735 */
            dy += ic_rand_pm1(&state);
            dz += ic_rand_pm1(&state);
        }
        const float len = sqrtf(ic_dot3(dx, dy, dz));         /* XMVector3Normalize :738 */
        const float px = cx + (dx / len) * spread;            /* :739-742 */
        const float py = (dy / len) * spread;
        const float pz = (dz / len) * spread;
        const float pl = sqrtf(ic_dot3(px, py, pz));
        const float ux = px / pl, uy = py / pl, uz = pz / pl; /* direction, :746 (exact instead of NormalizeEst) */
        float qx = 1.f - ux, qy = 1.f - uy, qz = 1.f - uz;    /* (1,1,1) - direction, :747 */
        const float ql = sqrtf(ic_dot3(qx, qy, qz));
        qx /= ql; qy /= ql; qz /= ql;
        out[i].pos[0] = px; out[i].pos[1] = py; out[i].pos[2] = pz; out[i].pos[3] = 0.f;
        out[i].velo[0] = (uy * qz - uz * qy) * speed;         /* cross(direction, perp) * initialSpeed, :748 */
        out[i].velo[1] = (uz * qx - ux * qz) * speed;
        out[i].velo[2] = (ux * qy - uy * qx) * speed;
        out[i].velo[3] = 0.f;
    }
}
