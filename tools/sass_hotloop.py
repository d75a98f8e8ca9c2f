#!/usr/bin/env python
"""tools/sass_hotloop.py -- the hot loop of one force_cells_kernel instantiation, from the built library.

    python tools/sass_hotloop.py [P T] > profiles/r02_sass_force_default.txt

Runs `cuobjdump -sass` on lib/libmapc.so, picks the fused, non-peer, default-staging instantiation of the
launch shape (P, T) (default: 2 128, the shape the bench workload runs), finds its innermost hot loop -- the
backward branch whose body holds the most FFMA2 -- and prints an opcode histogram of that loop, how many of
its 3-operand FFMA2 carry a `.reuse` flag, and the loop's SASS.  No GPU needed.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multi-adapter-particles_b200", "lib", "libmapc.so")


def hot_loop(want, lib=LIB):
    """(function name, [(address, text)] of the innermost backward-branch loop holding the most FFMA2) of the first
    function of `lib` whose mangled name matches the regular expression `want`."""
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    want = re.compile(want)
    body = next(f for f in funcs if want.search(f.split("\n", 1)[0]))
    name = body.split("\n", 1)[0].strip()
    ins = []   # (address, text)
    for ln in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    loops = []   # (first index, last index) of every backward branch
    for k, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr_index:
            loops.append((addr_index[int(m.group(1), 16)], k))
    best = None
    for lo, hi in loops:   # innermost loops only: no other loop strictly inside
        if any((l2, h2) != (lo, hi) and l2 >= lo and h2 <= hi for l2, h2 in loops):
            continue
        loop = ins[lo:hi + 1]
        n_ffma2 = sum(1 for _, x in loop if re.search(r"\bFFMA2\b", x))
        if best is None or n_ffma2 > best[0]:
            best = (n_ffma2, loop)
    return name, best[1]


def main():
    P, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (2, 128)
    name, loop = hot_loop(rf"force_cells_kernelILi{P}ELi{T}ELi\d+ELi\d+ELi\d+ELi\d+ELb1ELb0ELb0ELb0ELb0ELi2048E")
    hist = collections.Counter()
    for _, x in loop:
        x = re.sub(r"^@!?U?P\d+\s+", "", x)
        hist[x.split()[0].split(".")[0]] += 1
    acc = [x for _, x in loop if re.search(r"\bFFMA2\b", x)]
    reuse = sum(1 for x in acc if ".reuse" in x)
    print(f"# {name}")
    print(f"# hot loop: {len(loop)} instructions, 0x{loop[0][0]:x} .. 0x{loop[-1][0]:x} (first instruction at offset "
          f"0x{loop[0][0] % 128:x} of its 128-byte instruction line)")
    print("# opcode histogram: " + ", ".join(f"{k} {v}" for k, v in hist.most_common()))
    print(f"# FFMA2 with a .reuse operand: {reuse} of {len(acc)}")
    for a, x in loop:
        print(f"/*{a:04x}*/ {x}")


if __name__ == "__main__":
    main()
