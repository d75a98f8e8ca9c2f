"""Known-answer tests that pin the CPU oracle.

The reference ships no fixtures for this path, so each expected value here is derived by hand from
Particles/nBodyGravityCS.hlsl:44-57 (pair) and :86-109 (CSMain); tests/test_reference_shader.py pins the
oracle against the shader's own code compiled for the CPU.
"""
import numpy as np
import pytest

MASS = 70000.0


def pv(pos, vel=None):
    n = len(pos)
    a = np.zeros((n, 8), dtype=np.float32)
    a[:, :3] = np.asarray(pos, dtype=np.float32)
    if vel is not None:
        a[:, 4:7] = np.asarray(vel, dtype=np.float32)
    return a


@pytest.mark.parametrize("flavour", [0, 1])
def test_kat1_two_bodies_12_apart(oracle, flavour):
    # r = 12 on x: distSqr = 144 + 25 = 169, invDist = 1/13, a = 70000*12/2197
    acc = oracle.accel_allpairs(pv([[0, 0, 0], [12, 0, 0]]), S=1, flavour=flavour, scalar=True)
    expect = MASS * 12.0 / 2197.0
    assert acc[0, 0] == pytest.approx(expect, rel=2e-7)
    assert acc[1, 0] == pytest.approx(-expect, rel=2e-7)
    assert np.all(acc[:, 1:] == 0.0)


@pytest.mark.parametrize("flavour", [0, 1])
def test_kat2_r_8_8_4(oracle, flavour):
    # dot(r, r) = 144 -> same 1/13: a = (70000/2197) * (8, 8, 4)
    out = oracle.body_body_interaction([0, 0, 0], [8, 8, 4], [0, 0, 0], flavour=flavour)
    np.testing.assert_allclose(out, np.array([8, 8, 4]) * MASS / 2197.0, rtol=2e-7)


@pytest.mark.parametrize("flavour", [0, 1])
def test_kat3_self_pair_is_exactly_zero(oracle, flavour):
    out = oracle.body_body_interaction([0, 0, 0], [3.5, -2.25, 7.0], [3.5, -2.25, 7.0], flavour=flavour)
    assert np.all(out == 0.0)
    # and accumulating it leaves a previous sum untouched, bit for bit
    prev = np.array([1.25, -7.5, 3.0e-3], dtype=np.float32)
    out = oracle.body_body_interaction(prev, [1, 2, 3], [1, 2, 3], flavour=flavour)
    assert out.tobytes() == prev.tobytes()


def test_kat4_antisymmetry_momentum(oracle, mapc):
    p = mapc.ic.uniform_sphere(512, 300.0, seed=11)
    acc = oracle.accel_allpairs(p, S=8).astype(np.float64)
    scale = np.abs(acc).sum(axis=0)
    assert np.all(np.abs(acc.sum(axis=0)) < 1e-5 * scale)


@pytest.mark.parametrize("flavour", [0, 1])
def test_kat5_coincident_bodies(oracle, flavour):
    # r = 0: s = 70000/125 = 560 stays finite, force r*s = 0
    acc = oracle.accel_allpairs(pv([[5, 5, 5], [5, 5, 5], [5, 5, 5]]), S=1, flavour=flavour)
    assert np.all(acc == 0.0)


def test_kat5b_particles_multiplier(oracle):
    # the vestigial `particles` factor (:54): 1 is a no-op, -3 scales by -3
    one = oracle.body_body_interaction([0, 0, 0], [12, 0, 0], [0, 0, 0], particles=1)
    neg = oracle.body_body_interaction([0, 0, 0], [12, 0, 0], [0, 0, 0], particles=-3)
    assert neg[0] == pytest.approx(-3.0 * one[0], rel=1e-7)


@pytest.mark.parametrize("flavour", [0, 1])
def test_kat6_literal_well(oracle, flavour):
    # pos = (12,0,0), vel = 0: accel = (-382.33955,0,0), vel' = accel*0.1, pos' = 12 + vel'*0.1
    out = oracle.step_well(pv([[12, 0, 0]]), dt=0.1, damping=1.0, flavour=flavour)
    a = MASS * 12.0 / 2197.0
    assert out["velo"][0, 0] == pytest.approx(-a * 0.1, rel=3e-7)
    assert out["pos"][0, 0] == pytest.approx(12.0 - a * 0.01, rel=3e-7)
    assert out["pos"][0, 3] == pytest.approx(a, rel=3e-7)   # pos.w = length(accel), :107
    assert out["velo"][0, 3] == 0.0


def test_kat7_damping_after_kick_before_drift(oracle):
    # vel = (v0 + a*dt) * damping ; pos = p0 + vel*dt   (:103-105)
    p = pv([[12, 0, 0]], vel=[[100, 0, 0]])
    out = oracle.step_well(p, dt=0.1, damping=0.5)
    a = -MASS * 12.0 / 2197.0
    v = (100.0 + a * 0.1) * 0.5
    assert out["velo"][0, 0] == pytest.approx(v, rel=3e-7)
    assert out["pos"][0, 0] == pytest.approx(12.0 + v * 0.1, rel=3e-7)


def test_kat9_ragged_tail_equals_bounded_brute_force(oracle, mapc):
    # N not a multiple of 64 (nor of the segment count): no phantom bodies past N
    n = 1000
    p = mapc.ic.uniform_sphere(n, 200.0, seed=5)
    seg = oracle.accel_allpairs(p, S=32)
    one = oracle.accel_allpairs(p, S=1)
    f64 = oracle.accel_fp64(p)
    scale = np.abs(f64).max()
    assert np.abs(seg - f64).max() / scale < 2e-6
    assert np.abs(one - f64).max() / scale < 2e-6
    # a brute-force python loop over j < N for a few targets (float64) agrees too
    x = p["pos"][:, :3].astype(np.float64)
    for i in (0, 63, 64, 999):
        r = x - x[i]
        d2 = (r * r).sum(axis=1) + 25.0
        a = (r * (MASS / d2 ** 1.5)[:, None]).sum(axis=0)
        np.testing.assert_allclose(f64[i], a, rtol=1e-12, atol=1e-9)


def test_num_targets_follows_dispatch_granularity(oracle):
    # Dispatch(ceil(nActive/64)) x 64 threads, no bounds check, OOB writes dropped (Compute.cpp:1041)
    assert oracle.num_targets(10_000, 10_000) == 10_000
    assert oracle.num_targets(10_000, 100) == 128
    assert oracle.num_targets(10_000, 64) == 64
    assert oracle.num_targets(10_000, 0) == 0
    assert oracle.num_targets(100, 100) == 100


def test_segments_tile_aligned_and_cover(oracle):
    for n in (1, 63, 64, 65, 1000, 10_000, 262_144):
        for S in (1, 8, 32):
            prev = 0
            for s in range(S):
                j0, j1 = oracle.segment_range(n, S, s)
                assert j0 == prev and j1 >= j0
                assert j0 % 64 == 0 and (j1 % 64 == 0 or j1 == n)
                prev = j1
            assert prev == n
    # N = 10,000: 157 tiles, last tile has 16 valid bodies
    assert oracle.load().mapo_num_tiles(10_000) == 157
