"""Generates tests/golden/ref_shader_vectors.npz from the REFERENCE's own shader code.

oracle/_ref/libref_shader.so is Particles/nBodyGravityCS.hlsl compiled for the CPU (oracle/Makefile,
oracle/hlsl_shim.hpp); it exists only where /root/reference does, so its outputs are committed here as
fixtures: they pin the oracle (tests/test_reference_shader.py, CPU) and the CUDA path (GPU tests) to the
reference's arithmetic wherever the repository travels.
Usage (in the build container): make -C oracle && python tests/golden/make_ref_shader_vectors.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

pkg = importlib.import_module("multi-adapter-particles_b200")
orc = importlib.import_module("oracle.oracle_py")
ref = importlib.import_module("oracle.ref_shader")


def f32(a):
    return a.view(np.float32).reshape(-1, 8)


def main():
    rng = np.random.default_rng(20201)
    # bodyBodyInteraction on random pairs, incl. coincident bodies, huge separations, particles != 1
    m = 512
    ai = rng.normal(0, 50, (m, 3)).astype(np.float32)
    bj = rng.normal(0, 800, (m, 4)).astype(np.float32)
    bi = rng.normal(0, 800, (m, 4)).astype(np.float32)
    bj[:16] = bi[:16]                                   # coincident: r = 0
    bj[16:32] = bi[16:32] + np.float32(1e-3)             # inside the softening length
    bj[32:48] *= np.float32(1e4)                          # far away
    mass = np.full(m, 70000.0, np.float32)
    mass[48:64] = rng.uniform(1, 1e6, 16).astype(np.float32)
    particles = np.ones(m, np.int32)
    particles[64:80] = rng.integers(-3, 9, 16)
    pair_out = np.stack([ref.body_body_interaction(ai[k], bj[k], bi[k], float(mass[k]), int(particles[k]))
                         for k in range(m)])
    # the shipped CSMain (gravity well), two parameter sets
    w = pkg.ic.uniform_sphere(1024, 400.0, seed=31, speed=15.0)
    w["pos"][0, :3] = 0.0                               # a body sitting in the well
    well_a = ref.csmain(w, dt=0.1, damping=1.0)
    well_b = ref.csmain(w, dt=0.05, damping=0.995)
    # all-pairs steps through the reference's bodyBodyInteraction in the canonical order
    a_in = pkg.ic.uniform_sphere(1500, 700.0, seed=32, speed=2.0)
    S_a = orc.default_segments(a_in.shape[0])
    a_out = ref.step_allpairs(a_in, S_a, dt=0.1, damping=1.0)
    b_in = pkg.ic.plummer(1111, 300.0, seed=33, velocity_scale=0.1)
    S_b = orc.default_segments(b_in.shape[0])
    b_out = ref.step_allpairs(b_in, S_b, dt=0.05, damping=0.995)
    # a size at which the chains of the canonical order matter: N = 98,304 -> 32 segments of 3,072 sources = one
    # 2,048-source chain + one of 1,024, folded left to right.  The input is regenerated from its seed (its
    # sha256 is stored); outputs are kept for 384 targets only.
    import hashlib
    c_n, c_radius, c_seed = 98_304, 5800.0, 34
    c_in = pkg.ic.uniform_sphere(c_n, c_radius, c_seed, speed=1.0)
    c_idx = np.sort(rng.choice(c_n, 384, replace=False)).astype(np.int32)
    S_c = orc.default_segments(c_n)
    c_out = ref.step_allpairs_targets(c_in, c_idx, S_c, dt=0.1, damping=1.0, chain=orc.default_chain())
    assert c_out.tobytes() != ref.step_allpairs_targets(c_in, c_idx, S_c, dt=0.1, damping=1.0, chain=0).tobytes()
    soft, pmass = ref.constants()
    np.savez_compressed(os.path.join(HERE, "ref_shader_vectors.npz"),
                        pair_ai=ai, pair_bj=bj, pair_bi=bi, pair_mass=mass, pair_particles=particles,
                        pair_out=pair_out,
                        well_in=f32(w), well_out_a=f32(well_a), well_out_b=f32(well_b),
                        allpairs_a_in=f32(a_in), allpairs_a_out=f32(a_out), allpairs_a_S=S_a,
                        allpairs_b_in=f32(b_in), allpairs_b_out=f32(b_out), allpairs_b_S=S_b,
                        allpairs_c_n=c_n, allpairs_c_radius=np.float64(c_radius), allpairs_c_seed=c_seed,
                        allpairs_c_sha256=np.frombuffer(hashlib.sha256(c_in.tobytes()).digest(), dtype=np.uint8),
                        allpairs_c_targets=c_idx, allpairs_c_out=f32(c_out), allpairs_c_S=S_c,
                        allpairs_c_chain=orc.default_chain(),
                        softening_squared=np.float32(soft), particle_mass=np.float32(pmass))
    print("wrote ref_shader_vectors.npz")


if __name__ == "__main__":
    main()
