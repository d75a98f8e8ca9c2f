"""Pins the oracle to the REFERENCE's own arithmetic.

Two sources of truth, both produced by the reference's Particles/nBodyGravityCS.hlsl compiled for the CPU
(oracle/Makefile -> oracle/_ref/libref_shader.so):
  * tests/golden/ref_shader_vectors.npz -- its committed outputs (always available);
  * the library itself, where it has been built (this container and, as a built file, the GPU box).
The oracle's LITERAL flavour must reproduce both BIT FOR BIT: same operations in the same order.
"""
import importlib
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_shader_vectors.npz")


@pytest.fixture(scope="module")
def vec():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def ref():
    mod = importlib.import_module("oracle.ref_shader")
    if not mod.available():
        pytest.skip("oracle/_ref/libref_shader.so not built (needs /root/reference at build time)")
    mod.load()
    return mod


def pv(mapc, a):
    return np.ascontiguousarray(a, dtype=np.float32).view(mapc.POSVELO_DTYPE).reshape(-1)


# ---- committed vectors ------------------------------------------------------------------------
def test_constants_match_the_shader(mapc, vec):
    assert float(vec["softening_squared"]) == mapc.SOFTENING_SQUARED == 25.0     # nBodyGravityCS.hlsl:37
    assert float(vec["particle_mass"]) == mapc.PARTICLE_MASS == 70000.0          # :38


def test_oracle_pair_equals_reference_vectors(oracle, vec):
    for k in range(vec["pair_ai"].shape[0]):
        got = oracle.body_body_interaction(vec["pair_ai"][k], vec["pair_bj"][k], vec["pair_bi"][k],
                                           float(vec["pair_mass"][k]), int(vec["pair_particles"][k]))
        assert got.tobytes() == vec["pair_out"][k].tobytes(), k


def test_oracle_well_equals_reference_vectors(mapc, oracle, vec):
    w = pv(mapc, vec["well_in"])
    assert oracle.step_well(w, dt=0.1, damping=1.0).tobytes() == vec["well_out_a"].tobytes()
    assert oracle.step_well(w, dt=0.05, damping=0.995).tobytes() == vec["well_out_b"].tobytes()


@pytest.mark.parametrize("case,dt,damping", [("a", 0.1, 1.0), ("b", 0.05, 0.995)])
def test_oracle_allpairs_equals_reference_vectors(mapc, oracle, vec, case, dt, damping):
    inp = pv(mapc, vec[f"allpairs_{case}_in"])
    S = int(vec[f"allpairs_{case}_S"])
    assert S == oracle.default_segments(inp.shape[0])
    got = oracle.step_allpairs(inp, dt=dt, damping=damping, S=S, flavour=oracle.LITERAL)
    assert got.tobytes() == vec[f"allpairs_{case}_out"].tobytes()
    scalar = oracle.accel_allpairs(inp, S=S, flavour=oracle.LITERAL, scalar=True)
    vector = oracle.accel_allpairs(inp, S=S, flavour=oracle.LITERAL)
    assert scalar.tobytes() == vector.tobytes()


def chained_case(mapc, vec):
    """the N = 98,304 case of the fixture: input regenerated from its seed and checked against the stored hash"""
    import hashlib
    n, radius, seed = int(vec["allpairs_c_n"]), float(vec["allpairs_c_radius"]), int(vec["allpairs_c_seed"])
    p = mapc.ic.uniform_sphere(n, radius, seed, speed=1.0)
    assert hashlib.sha256(p.tobytes()).digest() == vec["allpairs_c_sha256"].tobytes()
    return p, vec["allpairs_c_targets"], int(vec["allpairs_c_S"]), int(vec["allpairs_c_chain"])


def test_oracle_chained_order_equals_reference_vectors(mapc, oracle, vec):
    """Segments longer than a chain (3,072 sources = a 2,048-source chain + one of 1,024): the oracle's
    canonical order against the reference's bodyBodyInteraction driven in that order."""
    p, idx, S, chain = chained_case(mapc, vec)
    assert S == oracle.default_segments(p.shape[0]) and chain == oracle.default_chain()
    got = oracle.step_allpairs_targets(p, idx, flavour=oracle.LITERAL)
    assert got.view(np.float32).tobytes() == vec["allpairs_c_out"].tobytes()
    one_chain = oracle.step_allpairs_targets(p, idx, flavour=oracle.LITERAL, chunk=0)
    assert one_chain.tobytes() != got.tobytes()
    scalar = oracle.accel_allpairs(p, flavour=oracle.LITERAL, targets=idx[:16], scalar=True)
    vector = oracle.accel_allpairs(p, flavour=oracle.LITERAL, targets=idx[:16])
    assert scalar.tobytes() == vector.tobytes()


# ---- the live library ----------------------------------------------------------------------------
def test_vectors_are_what_the_library_produces(mapc, ref, vec):
    """the committed fixture is not stale"""
    w = pv(mapc, vec["well_in"])
    assert ref.csmain(w, dt=0.1, damping=1.0).tobytes() == vec["well_out_a"].tobytes()
    inp = pv(mapc, vec["allpairs_b_in"])
    out = ref.step_allpairs(inp, int(vec["allpairs_b_S"]), dt=0.05, damping=0.995)
    assert out.tobytes() == vec["allpairs_b_out"].tobytes()
    p, idx, S, chain = chained_case(mapc, vec)
    out = ref.step_allpairs_targets(p, idx[:64], S, chain=chain)
    assert out.view(np.float32).tobytes() == vec["allpairs_c_out"][:64].tobytes()


@pytest.mark.parametrize("n,ic", [(64, "sphere"), (100, "sphere"), (1000, "plummer"), (4097, "sphere"),
                                  (6000, "lattice")])
def test_oracle_equals_live_reference(mapc, oracle, ref, n, ic):
    if ic == "sphere":
        p = mapc.ic.uniform_sphere(n, 900.0, seed=n, speed=3.0)
    elif ic == "plummer":
        p = mapc.ic.plummer(n, 300.0, seed=n, velocity_scale=0.1)
    else:
        p = mapc.ic.lattice_sphere(n, 1500.0, seed=n, speed=1.0)
    S = oracle.default_segments(n)
    assert ref.step_allpairs(p, S).tobytes() == oracle.step_allpairs(p, flavour=oracle.LITERAL).tobytes()
    assert ref.csmain(p).tobytes() == oracle.step_well(p, flavour=oracle.LITERAL).tobytes()
    targets = np.array([0, n // 3, n - 1], dtype=np.int32)
    assert (ref.step_allpairs_targets(p, targets, S, n_sources=n - n // 4).tobytes() ==
            oracle.step_allpairs_targets(p, targets, n_sources=n - n // 4, S=S, flavour=oracle.LITERAL).tobytes())


def test_oracle_equals_live_reference_with_chains(mapc, oracle, ref):
    """N = 70,001 (ragged last tile): 32 segments of ~2,188 sources = a full 2,048-source chain + a short one."""
    n = 70_001
    p = mapc.ic.uniform_sphere(n, 5000.0, seed=77, speed=1.0)
    targets = np.sort(np.random.default_rng(3).choice(n, 96, replace=False)).astype(np.int32)
    S = oracle.default_segments(n)
    a = ref.step_allpairs_targets(p, targets, S)
    b = oracle.step_allpairs_targets(p, targets, flavour=oracle.LITERAL)
    assert a.tobytes() == b.tobytes()
    assert a.tobytes() != ref.step_allpairs_targets(p, targets, S, chain=0).tobytes()


def test_oracle_equals_live_reference_ten_steps(mapc, oracle, ref):
    p = mapc.ic.uniform_sphere(1200, 600.0, seed=5, speed=2.0)
    S = oracle.default_segments(1200)
    a = b = p
    for _ in range(10):
        a = ref.step_allpairs(a, S)
        b = oracle.step_allpairs(b, flavour=oracle.LITERAL)
    assert a.tobytes() == b.tobytes()


def test_no_reference_source_in_the_repository():
    """the shader is compiled where it lies; only its binary may exist under oracle/_ref/"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(root):
        if ".git" in dirpath.split(os.sep):
            continue
        for f in files:
            assert not f.endswith(".hlsl"), os.path.join(dirpath, f)
    ref_dir = os.path.join(root, "oracle", "_ref")
    if os.path.isdir(ref_dir):
        assert all(f.endswith(".so") for f in os.listdir(ref_dir)), os.listdir(ref_dir)
