"""BASELINE config 3 at full size, ALL 262,144 targets against the LITERAL oracle.

Collected last on purpose (file name): it is the longest test of the suite -- 6.9e10 interactions per step
on the host cores, 11 oracle steps, about a minute with the vectorised oracle -- so everything else has
reported before it starts.

Why it exists: a 2048-target subsample (test_gpu_parity.py) misses the handful of targets with a neighbour
inside a few softening lengths, and those carry the largest rounding error of the whole step -- once such a
neighbour has made a segment's fp32 accumulator large, every later term of that chain is rounded at that
magnitude.  The canonical segment rule (mapc_plan_segments: chains of at most 8,192 sources) is sized from
this: two correctly rounded CPU evaluations of the same formula (LITERAL vs MIRRORED) differ by 1.07e-5 over
all targets with 32,768-term chains and by 4.6e-6 with 8,192-term chains.
"""
import pytest

pytestmark = pytest.mark.gpu

TOL_1 = 1e-5     # BASELINE.json north_star: max relative error after one step
TOL_10 = 1e-4    # ... after ten steps


def gpu_steps(mapc, particles, steps):
    n = particles.shape[0]
    with mapc.Compute(n, 0) as c:
        c.Upload(particles)
        for _ in range(steps):
            c.Simulate(n, 0, 0.1, 1.0)
        c.WaitForGpu()
        return c.Download()


def test_full_size_262144_one_step_all_targets(mapc, oracle, gpu):
    p = mapc.ic.workload("sphere_262144")           # the bench workload (bench.py, config 3)
    assert mapc.plan_segments(p.shape[0]) == oracle.default_segments(p.shape[0]) == 32
    got = gpu_steps(mapc, p, 1)
    err = oracle.rel_errors(got, oracle.step_allpairs(p, flavour=oracle.LITERAL))
    print("N=262,144, all targets, 1 step:", err)
    assert max(err.values()) <= TOL_1, err


def test_full_size_262144_ten_steps_all_targets_lattice(mapc, oracle, gpu):
    """Ten steps at the stated 1e-4 on a close-pair-free lattice sphere of the same size (on a random sphere
    the ten-step figure measures chaotic amplification, see test_allpairs_ten_steps_10k_random_sphere)."""
    n = 262_144
    q = mapc.ic.lattice_sphere(n, 8000.0, seed=12, speed=1.0)
    got = gpu_steps(mapc, q, 10)
    ref = q
    for _ in range(10):
        ref = oracle.step_allpairs(ref, flavour=oracle.LITERAL)
    err = oracle.rel_errors(got, ref)
    print("N=262,144 lattice, all targets, 10 steps:", err)
    assert max(err.values()) <= TOL_10, err


@pytest.mark.xfail(strict=False, reason="experimental opt-in variant (MAPC_CHUNK=1): verified against the oracle by "
                   "CPU emulation only so far; informational until its first run on hardware")
def test_experimental_bounded_chain_order(mapc, oracle, gpu):
    """MAPC_CHUNK=1: chains bounded at 2,048 sources.  N = 131,072 with S = 32 gives
    segments of 4,096 sources = two chunks each, so the order really differs from the default; the result must
    match the oracle's `chunk=2048` order like the default matches the plain one."""
    import os
    n = 131_072
    p = mapc.ic.uniform_sphere(n, 6350.0, seed=5)
    default = gpu_steps(mapc, p, 1)
    try:
        os.environ["MAPC_CHUNK"] = "1"
        chunked = gpu_steps(mapc, p, 1)
    finally:
        os.environ.pop("MAPC_CHUNK", None)
    assert chunked.tobytes() != default.tobytes()
    lit = oracle.step_allpairs(p, flavour=oracle.LITERAL, chunk=2048)
    mir = oracle.step_allpairs(p, flavour=oracle.MIRRORED, chunk=2048)
    err_l, err_m = oracle.rel_errors(chunked, lit), oracle.rel_errors(chunked, mir)
    print("MAPC_CHUNK=1 vs chunked oracle: literal", err_l, "mirrored", err_m)
    # both at the 1e-5 gate: over ALL targets the distance to MIRRORED is set by the same chain noise as the
    # distance to LITERAL (MUFU.RSQ perturbs every term, after which the roundings of the chain decorrelate)
    assert max(err_l.values()) <= TOL_1 and max(err_m.values()) <= TOL_1, (err_l, err_m)
