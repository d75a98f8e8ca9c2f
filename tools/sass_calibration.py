#!/usr/bin/env python
"""Does the static SASS model (tools/sass_model.py) predict the measured sweep?  No GPU needed.

Builds on: `nvcc ... -DSWEEP_BROAD -o /tmp/ubench_broad tools/ubench.cu` (compile only) and the measured table
profiles/r01_ubench_shapes_11op.txt (144 launch shapes of the unfused 11-op kernel on B200).  For every shape
it reads the hot loop's instruction mix from the SASS and regresses the measured cycles per pair-interaction
(20 / fraction of peak) on it.  Result (round 1): R^2 = 0.43-0.50; each non-FMA instruction per pair costs
0.3-0.44 cycles; the number of accumulations that fetch three distinct register pairs WITHOUT a `.reuse` hit has
no positive weight (-0.2) -- so `.reuse` adjacency is not a lever, and the model is not good enough to optimise
the source against offline.  Kept so the question does not have to be asked twice.
"""
import re, sys, subprocess, numpy as np
sys.path.insert(0, "/root/repo/tools")
import sass_model as sm
meas = {}
for line in open("/root/repo/profiles/r01_ubench_shapes_11op.txt"):
    m = re.search(r"(pair-major|op-major)\s+P=(\d+) T=\s*(\d+) TJ=\s*(\d+) U=(\d+) minB=\s*(\d+) regs=\s*(\d+) occ=\s*(\d+) blk/SM \(\s*(\d+) warps\).*?([\d.]+) % of", line)
    if m:
        order = 0 if m.group(1) == "pair-major" else 2
        key = tuple(int(m.group(i)) for i in (2, 3, 4, 5, 6)) + (order,)
        meas[key] = (float(m.group(10)), int(m.group(7)), int(m.group(9)))
print(len(meas), "measured shapes")
rows = []
for name, body in sm.functions("/tmp/ubench_broad"):
    m = re.search(r"force_cells_kernelILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELb0ELb0ELb0ELb0E", name)
    if not m:
        continue
    key = tuple(int(x) for x in m.groups())
    if key not in meas:
        continue
    r = sm.analyse(body)
    if not r:
        continue
    pct, regs, warps = meas[key]
    inter = r["mufu"]
    rows.append((key, pct, warps, r["packed_3cyc"] / (inter / 2), r["loop_instructions"] / (inter / 2), r["ceiling_pct_fp32_peak"], r["lds"] / (inter/2)))
print(len(rows), "matched")
rows.sort()
A = np.array([[r[1], r[2], r[3], r[4], r[5]] for r in rows])
print("corr(measured, model ceiling) =", np.corrcoef(A[:,0], A[:,4])[0,1])
print("corr(measured, 3-read accumulates per pair) =", np.corrcoef(A[:,0], A[:,2])[0,1])
print("corr(measured, instr per pair) =", np.corrcoef(A[:,0], A[:,3])[0,1])
print("corr(measured, warps) =", np.corrcoef(A[:,0], A[:,1])[0,1])
for r in sorted(rows, key=lambda r: -r[1])[:12]:
    print("best", r)
for r in sorted(rows, key=lambda r: r[1])[:6]:
    print("worst", r)

# regression: measured cycles per pair-interaction (20 / frac) on instruction-mix features
y = 20.0 / (A[:,0] / 100.0)
others = A[:,3] - 11.0                    # non-packed instructions per pair-interaction
three = A[:,2]                            # accumulates that fetch 3 pairs (no reuse), per pair-interaction
warps = A[:,1]
X = np.stack([np.ones_like(y), others, three], axis=1)
coef, res, rk, sv = np.linalg.lstsq(X, y, rcond=None)
pred = X @ coef
print("fit: cycles = %.2f + %.2f*others + %.2f*three_read   R2 = %.3f  rms = %.2f" % (coef[0], coef[1], coef[2], 1 - ((y-pred)**2).sum()/((y-y.mean())**2).sum(), np.sqrt(((y-pred)**2).mean())))
X2 = np.stack([np.ones_like(y), others, three, 1.0/warps], axis=1)
coef2 = np.linalg.lstsq(X2, y, rcond=None)[0]
pred2 = X2 @ coef2
print("fit2: cycles = %.2f + %.2f*others + %.2f*three + %.2f/warps  R2 = %.3f rms = %.2f" % (*coef2, 1 - ((y-pred2)**2).sum()/((y-y.mean())**2).sum(), np.sqrt(((y-pred2)**2).mean())))
for r, yy, pp in sorted(zip(rows, y, pred2), key=lambda t: -abs(t[1]-t[2]))[:8]:
    print("outlier", r[0], "measured %.2f pred %.2f" % (yy, pp), "warps", r[2])
