#!/usr/bin/env python
"""tools/shape_sweep.py -- microseconds per chained step for every launch shape over a range of small and medium
N, in one process (no bench.py start-up per point): the data the launch-plan cost model (csrc/step_layout.hpp) is
checked against.  Prints one line per N: the time of every shape, the fastest, and what the library picks by itself.

    python tools/shape_sweep.py > gpurun_out/rXX_small_n_shapes.txt
"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(4, 256), (4, 128), (2, 128), (2, 64), (1, 128), (1, 64), (1, 32)]
SIZES = [1000, 2500, 4096, 6000, 8192, 10_000, 16_384, 24_576, 32_768, 65_536, 131_072]


def us_per_step(pkg, n, p, shape):
    for k in ("MAPC_PLAN_PAIRS", "MAPC_PLAN_THREADS"):
        os.environ.pop(k, None)
    if shape is not None:
        os.environ["MAPC_PLAN_PAIRS"], os.environ["MAPC_PLAN_THREADS"] = str(shape[0]), str(shape[1])
    steps = max(20, min(1000, int(2e10 / (n * n))))
    with pkg.Compute(n, 0) as c:
        c.Upload(p)
        plan = c.Plan()
        c.SimulateSteps(n, min(steps, 50))
        c.WaitForGpu()
        t0 = time.perf_counter()
        for k in range(0, steps, 50):
            c.SimulateSteps(n, min(50, steps - k))
        c.WaitForGpu()
        return (time.perf_counter() - t0) / steps * 1e6, (plan["pairs_per_thread"], plan["threads_per_block"])


def main():
    pkg = importlib.import_module("multi-adapter-particles_b200")
    pkg.load()
    print("# us per step (batches of 50, chained), one B200; columns: " + " ".join(f"({p},{t})" for p, t in SHAPES) +
          " | fastest | library's own pick and its time")
    for n in SIZES:
        p = pkg.ic.uniform_sphere(n, 2000.0 * (n / 10_000.0) ** (1 / 3), seed=1)
        times = [us_per_step(pkg, n, p, sh)[0] for sh in SHAPES]
        own, picked = us_per_step(pkg, n, p, None)
        best = min(range(len(SHAPES)), key=lambda k: times[k])
        peak = n * n / 3.722e12 * 1e6
        print(f"N={n:7d} " + " ".join(f"{t:9.2f}" for t in times) + f" | {SHAPES[best]} {times[best]:.2f} "
              f"({100 * peak / times[best]:.1f} % of peak) | {picked} {own:.2f}"
              f"{'' if picked == SHAPES[best] else f'  <- {100 * (own / times[best] - 1):.1f} % slower than the fastest'}", flush=True)


if __name__ == "__main__":
    main()
