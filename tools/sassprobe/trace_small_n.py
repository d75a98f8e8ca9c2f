#!/usr/bin/env python
"""tools/sassprobe/trace_small_n.py -- per-cell timeline of chained small-N steps (experiment build with -DMAPC_TRACE).

    MAPC_LIB_PATH=tools/sassprobe/trace_lib/libmapc.so python tools/sassprobe/trace_small_n.py [N] > gpurun_out/trace.txt
Every cell records %globaltimer at: 2 start (ticket taken), 3 inputs ready (flags of the previous step seen), 4 sources done,
5 partial written + arrival counted, 6 (last arrival only) target block integrated and published.
"""
import ctypes
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
pkg = importlib.import_module("multi-adapter-particles_b200")
lib = pkg.load()
p = pkg.ic.uniform_sphere(n, 2000.0 * (n / 10000.0) ** (1 / 3), seed=1)
with pkg.Compute(n, 0) as c:
    c.Upload(p)
    plan = c.Plan()
    for _ in range(6):
        c.SimulateSteps(n, 50)
    c.WaitForGpu()
    t = c.StepTimes()
    buf = np.zeros((64, 2048, 8), dtype=np.uint64)
    rc = lib.mapc_debug_trace_dump(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes))
    assert rc == 0, rc
print("plan", plan, "median step us", float(np.median(t)) * 1e3)
cells = plan["blocks"]
out = os.path.join(ROOT, "gpurun_out", "trace_n%d.npy" % n)
os.makedirs(os.path.dirname(out), exist_ok=True)
np.save(out, buf[:, :cells, :])
steps = {}
for s in range(64):
    r = buf[s, :cells]
    sid = int(r[0, 0] >> np.uint64(32))
    steps[sid] = r.astype(np.int64)
ids = sorted(steps)
print("steps in the buffer:", ids[0], "..", ids[-1])
prev_end = None
for sid in ids[8:28]:
    r = steps[sid]
    start, ready, done, arr = r[:, 2], r[:, 3], r[:, 4], r[:, 5]
    pub = r[:, 6][r[:, 6] > 0]
    t0 = start.min()
    end = pub.max() if pub.size else arr.max()
    ref = prev_end if prev_end is not None else t0
    q = lambda x: "%6.1f %6.1f %6.1f %6.1f" % tuple((np.percentile(x, [0, 50, 95, 100]) - ref) / 1e3)
    print(f"step {sid}: prev_end->end {(end - ref) / 1e3:6.1f} us | start[min/med/p95/max] {q(start)} | ready {q(ready)} | "
          f"sources done {q(done)} | compute us med {np.median(done - ready) / 1e3:5.1f} max {(done - ready).max() / 1e3:5.1f} | "
          f"wait-for-inputs us med {np.median(ready - start) / 1e3:5.1f} | first/last publish {(pub.min() - ref) / 1e3:6.1f} {(pub.max() - ref) / 1e3:6.1f}")
    prev_end = end
# one step in detail: busy cells over time, per-SM cell counts, which cells finish last
sid = ids[20]
r = steps[sid]
ref = steps[sid - 1][:, 6].max()
print(f"\nstep {sid} in detail (time 0 = last publish of step {sid - 1})")
ts = np.arange(-40, 50, 2.0)
for lo in ts:
    a = ((r[:, 3] - ref) / 1e3 <= lo + 2) & ((r[:, 4] - ref) / 1e3 >= lo)       # computing during [lo, lo+2)
    w = ((r[:, 2] - ref) / 1e3 <= lo + 2) & ((r[:, 3] - ref) / 1e3 >= lo)       # resident, waiting for inputs
    print(f"  t={lo:6.1f} us: computing {int(a.sum()):5d}  waiting-for-inputs {int(w.sum()):5d}")
sm = np.bincount(r[:, 1].astype(int), minlength=148)
print("cells per SM: min %d max %d; histogram %s" % (sm.min(), sm.max(), np.bincount(sm).tolist()))
order = np.argsort(r[:, 4])
segs = plan["segments"]
print("last 12 cells to finish their sources: (cell, tb, seg, start, ready, done) us")
for cidx in order[-12:]:
    print("   ", int(cidx), int(cidx) // segs, int(cidx) % segs, *["%.1f" % ((r[cidx, k] - ref) / 1e3) for k in (2, 3, 4)])
print("first 12 cells to get their inputs:")
order = np.argsort(r[:, 3])
for cidx in order[:12]:
    print("   ", int(cidx), int(cidx) // segs, int(cidx) % segs, *["%.1f" % ((r[cidx, k] - ref) / 1e3) for k in (2, 3, 4)])
