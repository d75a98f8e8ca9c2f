#!/usr/bin/env python
"""Static cost model of the force kernel's inner loop, read from SASS (no GPU needed).

Model (measured on B200, profiles/r01_ubench.txt): per SM sub-partition the FMA pipe retires one
packed FFMA2/FMUL2/FADD2 per 2 cycles, and the register file delivers one even- and one
odd-numbered 32-bit register per cycle, so an instruction that must fetch three distinct
register pairs takes 3 cycles unless an operand sits in the reuse cache (`.reuse` flagged on the
same slot of the instruction issued just before).  The script finds the hottest loop (most
packed ops), and reports cycles per packed op and the implied ceiling in % of FP32 peak at
20 flop/interaction.

usage: python tools/sass_model.py <lib.so|cubin> [kernel-name-substring]
"""
from __future__ import annotations

import re
import subprocess
import sys

PACKED = ("FFMA2", "FMUL2", "FADD2")
SCALAR = ("FFMA", "FMUL", "FADD")


def functions(path: str):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\*", line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, body


def parse(instr: str):
    pred = None
    if instr.startswith("@"):
        pred, instr = instr.split(None, 1)
    parts = instr.split(None, 1)
    op = parts[0]
    ops = [o.strip() for o in parts[1].split(",")] if len(parts) > 1 else []
    return pred, op, ops


def src_regs(operand: str):
    """-> (list of 32-bit register indices read, reuse flag, key)"""
    m = re.search(r"\bR(\d+)", operand)
    if not m:
        return [], False, None
    r = int(m.group(1))
    reuse = ".reuse" in operand
    if "F32x2" in operand or ".64" in operand:
        return [r, r + 1], reuse, ("pair", r)
    return [r], reuse, ("reg", r)


def analyse(body):
    # hottest backward-branch loop
    best = None
    for idx, (addr, ins) in enumerate(body):
        _, op, ops = parse(ins)
        if op.startswith("BRA") and ops:
            m = re.search(r"0x([0-9a-f]+)", ops[-1])
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt < addr:
                lo = next(i for i, (a, _) in enumerate(body) if a >= tgt)
                inner = any(parse(s)[1].startswith("BRA") and re.search(r"0x([0-9a-f]+)", s) and
                            int(re.search(r"0x([0-9a-f]+)", s).group(1), 16) < a
                            for a, s in body[lo:idx])
                if inner:
                    continue                  # only innermost loops
                n_packed = sum(1 for _, s in body[lo:idx + 1] if parse(s)[1].split(".")[0] in PACKED)
                if best is None or n_packed > best[0]:
                    best = (n_packed, lo, idx)
    if best is None or best[0] == 0:
        return None
    n_packed, lo, hi = best
    loop = body[lo:hi + 1]
    prev_reuse = {}
    fma_cycles = 0
    rf_even = rf_odd = 0
    counts = {}
    hist = {2: 0, 3: 0}
    issue = 0
    for _, ins in loop:
        _, op, ops = parse(ins)
        base = op.split(".")[0]
        counts[base] = counts.get(base, 0) + 1
        issue += 1
        srcs = ops[1:] if base not in ("STS", "STG", "BAR", "BRA") else ops
        cur_reuse = {}
        ev = od = 0
        seen = set()
        for slot, o in enumerate(srcs):
            regs, reuse, key = src_regs(o)
            if key is None:
                continue
            if reuse:
                cur_reuse[slot] = key
            if prev_reuse.get(slot) == key:
                continue                      # operand comes from the reuse cache
            for r in regs:
                if r in seen:
                    continue
                seen.add(r)
                if r % 2 == 0:
                    ev += 1
                else:
                    od += 1
        prev_reuse = cur_reuse
        rf = max(ev, od)
        rf_even += ev
        rf_odd += od
        if base in PACKED:
            c = max(2, rf)
            hist[c] = hist.get(c, 0) + 1
            fma_cycles += c
        elif base in SCALAR:
            fma_cycles += 1                   # measured: scalar FFMA with 3 distinct registers still issues every cycle
    mufu = counts.get("MUFU", 0)
    interactions = mufu                        # one MUFU.RSQ per interaction
    bound = max(fma_cycles, rf_even, rf_odd, issue)
    res = {
        "loop_instructions": len(loop), "packed": n_packed, "mufu": mufu, "lds": counts.get("LDS", 0),
        "packed_2cyc": hist.get(2, 0), "packed_3cyc": hist.get(3, 0),
        "fma_pipe_cycles": fma_cycles, "rf_even": rf_even, "rf_odd": rf_odd, "issue_slots": issue,
        "cycles_per_interaction": bound / interactions if interactions else float("nan"),
    }
    if interactions:
        # 100 % of peak = 12.8 interactions/clk/SM = 3.2 per SMSP-clk -> 10 cycles per 32-lane interaction
        res["ceiling_pct_fp32_peak"] = 100.0 * 10.0 / res["cycles_per_interaction"]
    return res


def main():
    path = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else "force"
    for name, body in functions(path):
        if pat not in name:
            continue
        r = analyse(body)
        if r:
            print(name)
            print("   ", r)


if __name__ == "__main__":
    main()
