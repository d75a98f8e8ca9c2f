#!/usr/bin/env python
"""tools/sassprobe/time_lib.py -- median in-kernel step time (ms) of the library MAPC_LIB_PATH names at one N, plus the
sha256 of the state after the steps (a renamed kernel must reproduce the bits of the original)."""
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1     # > 1: chained steps, SimulateSteps(n, batch) per call
pkg = importlib.import_module("multi-adapter-particles_b200")
pkg.load()
p = pkg.ic.uniform_sphere(n, 8000.0 * (n / 262144.0) ** (1 / 3), seed=2)
with pkg.Compute(n, 0) as c:
    c.Upload(p)
    c.SimulateSteps(n, 2)
    c.WaitForGpu()
    c.StepTimes()
    if batch > 1:
        c.SimulateSteps(n, batch)
        c.WaitForGpu()
        c.StepTimes()
        for _ in range(0, steps, batch):
            c.SimulateSteps(n, batch)
        c.WaitForGpu()
    else:
        for _ in range(steps):
            c.Simulate(n)
            c.WaitForGpu()
    t = c.StepTimes()
    plan = c.Plan()
    out = c.Download()
print("%.5f %.5f %s (%d,%d)" % (float(np.median(t)), float(t.min()), hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()[:16],
                              plan["pairs_per_thread"], plan["threads_per_block"]))
