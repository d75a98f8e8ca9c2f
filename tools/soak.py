#!/usr/bin/env python
"""tools/soak.py -- long bit-identity runs of the two in-kernel protocols that only concurrency can break:
chained small-N steps (per-target-block flags, MAPC_CHAIN) and the scratch ring (ticket + slot generations,
MAPC_RING).  Each protocol is run for many steps and compared, bitwise, with the same run with the protocol off;
any stale read, lost flag or early slot reuse shows as a different digest (or as MAPC_ERR_TIMEOUT).

    python tools/soak.py [--chain-steps 20000] [--ring-steps 100]
"""
import argparse
import hashlib
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(pkg, particles, steps, batch, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        n = particles.shape[0]
        with pkg.Compute(n, 0) as c:
            c.Upload(particles)
            c.SimulateSteps(n, 2)          # module load and first-launch costs stay out of the timing
            c.WaitForGpu()
            c.Upload(particles)
            t0 = time.perf_counter()
            for k in range(0, steps, batch):
                c.SimulateSteps(n, min(batch, steps - k))
            c.WaitForGpu()
            dt = time.perf_counter() - t0
            return hashlib.sha256(c.Download().tobytes()).hexdigest(), dt / steps
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chain-steps", type=int, default=20000)
    ap.add_argument("--ring-steps", type=int, default=100)
    args = ap.parse_args()
    pkg = importlib.import_module("multi-adapter-particles_b200")
    pkg.load()
    ok = True
    # lattice spheres: no close encounters, so 20,000 steps stay finite and every bit keeps meaning something
    for n, radius in ((10_000, 4000.0), (2_500, 2500.0)):
        p = pkg.ic.lattice_sphere(n, radius, seed=3, speed=1.0)
        a, ta = run(pkg, p, args.chain_steps, 50, {"MAPC_CHAIN": "1"})
        b, tb = run(pkg, p, args.chain_steps, 50, {"MAPC_CHAIN": "0"})
        c, tc = run(pkg, p, args.chain_steps, 1, {"MAPC_CHAIN": "1"})
        same = a == b == c
        ok = ok and same
        print(f"chained steps N={n} steps={args.chain_steps}: chained {ta * 1e6:.2f} us/step, grid-wide wait {tb * 1e6:.2f}, "
              f"single calls {tc * 1e6:.2f}; sha256 {a[:16]} bit-identical={same}", flush=True)
    p = pkg.ic.uniform_sphere(262_144, 8000.0, 2)
    a, ta = run(pkg, p, args.ring_steps, 10, {"MAPC_RING": "1"})
    b, tb = run(pkg, p, args.ring_steps, 10, {"MAPC_RING": "0"})
    same = a == b
    ok = ok and same
    print(f"scratch ring N=262144 steps={args.ring_steps}: ring {ta * 1e3:.3f} ms/step, no ring {tb * 1e3:.3f}; "
          f"sha256 {a[:16]} bit-identical={same}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
