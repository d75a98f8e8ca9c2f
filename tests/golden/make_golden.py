"""Generates the golden fixtures in tests/golden/ from the CPU oracle (LITERAL flavour).

The reference has no fixtures of its own; these vectors pin the oracle against silent drift and give
the GPU parity tests a committed target.  (Vectors produced by the reference's own shader code are in
ref_shader_vectors.npz, see make_ref_shader_vectors.py.)
Usage: python tests/golden/make_golden.py
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


def allpairs(name, inp, steps, dt, damping):
    S = orc.default_segments(inp.shape[0])
    state, first = inp, None
    for k in range(steps):
        state = orc.step_allpairs(state, dt=dt, damping=damping, S=S)
        if k == 0:
            first = state.copy()
    f64 = orc.accel_fp64(inp)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), input=inp.view(np.float32).reshape(-1, 8),
                        literal_1=first.view(np.float32).reshape(-1, 8),
                        literal_last=state.view(np.float32).reshape(-1, 8), accel_fp64=f64,
                        steps=steps, dt=np.float32(dt), damping=np.float32(damping), S=S)


def main():
    allpairs("lattice_1000", pkg.ic.lattice_sphere(1000, 900.0, seed=7, speed=2.0), 10, 0.1, 1.0)
    allpairs("plummer_777", pkg.ic.plummer(777, 500.0, seed=8, velocity_scale=0.1), 10, 0.05, 0.995)
    w = pkg.ic.uniform_sphere(1000, 400.0, seed=9, speed=15.0)
    out = orc.step_well(w, dt=0.1, damping=1.0)
    np.savez_compressed(os.path.join(HERE, "well_1000.npz"), input=w.view(np.float32).reshape(-1, 8),
                        literal_1=out.view(np.float32).reshape(-1, 8), dt=np.float32(0.1),
                        damping=np.float32(1.0))


if __name__ == "__main__":
    main()
