"""Second, independent restatement of the shader in numpy float32, compared BIT FOR BIT with the C oracle.

numpy's float32 add / sub / mul / div / sqrt are IEEE-754 correctly rounded, exactly like the C
operators under -ffp-contract=off, so a line-by-line transcription of
Particles/nBodyGravityCS.hlsl:44-57 and :86-109 must give the same bits as oracle/oracle.c (LITERAL
flavour).  Two restatements written separately and agreeing on every bit complement the direct pin
against the reference's shader code compiled for the CPU (tests/test_reference_shader.py).
"""
import numpy as np
import pytest

F = np.float32
SOFTENING_SQUARED = F(25)        # nBodyGravityCS.hlsl:37
PARTICLE_MASS = F(70000)         # nBodyGravityCS.hlsl:38


def body_body_interaction(ai, bj, bi, mass, particles):
    """nBodyGravityCS.hlsl:44-57, one line of numpy per line of HLSL (all float32)."""
    r = (bj[:3] - bi[:3]).astype(F)                                   # :46
    dist_sqr = F(F(F(r[0] * r[0]) + F(r[1] * r[1])) + F(r[2] * r[2]))  # :48 dot(r, r)
    dist_sqr = F(dist_sqr + SOFTENING_SQUARED)                        # :49
    inv_dist = F(F(1.0) / np.sqrt(dist_sqr, dtype=F))                 # :51
    inv_dist_cube = F(F(inv_dist * inv_dist) * inv_dist)              # :52
    s = F(F(mass * inv_dist_cube) * F(particles))                     # :54
    return (ai + (r * s).astype(F)).astype(F)                         # :56


def accel_segments(pos, i, S, segment_range):
    total = np.zeros(3, dtype=F)
    n = pos.shape[0]
    for seg in range(S):
        j0, j1 = segment_range(n, S, seg)
        p = np.zeros(3, dtype=F)
        for j in range(j0, j1):
            p = body_body_interaction(p, pos[j], pos[i], PARTICLE_MASS, 1)
        total = (total + p).astype(F)
    return total


def integrate(pos, vel, accel, dt, damping):
    """nBodyGravityCS.hlsl:103-108"""
    vel = (vel + (accel * F(dt)).astype(F)).astype(F)                 # :103
    vel = (vel * F(damping)).astype(F)                                # :104
    new = (pos[:3] + (vel * F(dt)).astype(F)).astype(F)               # :105
    length = np.sqrt(F(F(F(accel[0] * accel[0]) + F(accel[1] * accel[1])) + F(accel[2] * accel[2])), dtype=F)
    return np.array([new[0], new[1], new[2], length], dtype=F), vel   # :107-108


@pytest.mark.parametrize("n,S", [(97, 8), (200, 32)])
def test_allpairs_step_bitwise_equals_numpy_transcription(oracle, mapc, n, S):
    p = mapc.ic.uniform_sphere(n, 120.0, seed=n, speed=4.0)
    got = oracle.step_allpairs(p, dt=0.1, damping=0.97, S=S, flavour=oracle.LITERAL)
    pos, vel = p["pos"].astype(F), p["velo"][:, :3].astype(F)
    for i in range(n):
        a = accel_segments(pos, i, S, oracle.segment_range)
        new_pos, new_vel = integrate(pos[i], vel[i], a, 0.1, 0.97)
        assert got["pos"][i].tobytes() == new_pos.tobytes(), i
        assert got["velo"][i, :3].tobytes() == new_vel.tobytes(), i
        assert got["velo"][i, 3] == 0.0


def test_well_step_bitwise_equals_numpy_transcription(oracle, mapc):
    """CSMain as shipped, nBodyGravityCS.hlsl:88-108 (origin well, invDist = -1/sqrt at :97)."""
    p = mapc.ic.uniform_sphere(500, 400.0, seed=5, speed=15.0)
    got = oracle.step_well(p, dt=0.1, damping=1.0, flavour=oracle.LITERAL)
    for i in range(p.shape[0]):
        pos, vel = p["pos"][i].astype(F), p["velo"][i, :3].astype(F)
        r = pos[:3]                                                                 # :92
        dist_sqr = F(F(F(r[0] * r[0]) + F(r[1] * r[1])) + F(r[2] * r[2]))            # :94
        dist_sqr = F(dist_sqr + SOFTENING_SQUARED)                                  # :95
        inv_dist = F(F(-1.0) / np.sqrt(dist_sqr, dtype=F))                          # :97
        inv_dist_cube = F(F(inv_dist * inv_dist) * inv_dist)                        # :98
        s = F(PARTICLE_MASS * inv_dist_cube)                                        # :99
        accel = (r * s).astype(F)                                                   # :101
        new_pos, new_vel = integrate(pos, vel, accel, 0.1, 1.0)
        assert got["pos"][i].tobytes() == new_pos.tobytes(), i
        assert got["velo"][i, :3].tobytes() == new_vel.tobytes(), i


def accel_segments_chunked(pos, i, S, chunk, segment_range):
    """The canonical chains (oracle `chunk`, kernel template parameter CHAIN): a segment longer than `chunk` sources is
    taken in consecutive chunks counted from its first source, each one sequential chain; the chunk sums are
    folded left to right into the segment's partial, the partials left to right into the total."""
    total = np.zeros(3, dtype=F)
    n = pos.shape[0]
    for seg in range(S):
        j0, j1 = segment_range(n, S, seg)
        step = chunk if (j1 - j0) > chunk else max(j1 - j0, 1)
        partial = None
        for c0 in range(j0, max(j1, j0 + 1), step):
            c = np.zeros(3, dtype=F)
            for j in range(c0, min(c0 + step, j1)):
                c = body_body_interaction(c, pos[j], pos[i], PARTICLE_MASS, 1)
            partial = c if partial is None else (partial + c).astype(F)
        total = (total + partial).astype(F)
    return total


@pytest.mark.parametrize("n,S,chunk", [(200, 2, 64), (257, 1, 64), (130, 32, 64)])
def test_chunked_order_bitwise_equals_numpy_transcription(oracle, mapc, n, S, chunk):
    p = mapc.ic.uniform_sphere(n, 120.0, seed=n, speed=4.0)
    got = oracle.step_allpairs(p, dt=0.1, damping=0.97, S=S, flavour=oracle.LITERAL, chunk=chunk)
    pos, vel = p["pos"].astype(F), p["velo"][:, :3].astype(F)
    for i in range(0, n, 3):
        a = accel_segments_chunked(pos, i, S, chunk, oracle.segment_range)
        new_pos, new_vel = integrate(pos[i], vel[i], a, 0.1, 0.97)
        assert got["pos"][i].tobytes() == new_pos.tobytes(), i
        assert got["velo"][i, :3].tobytes() == new_vel.tobytes(), i
