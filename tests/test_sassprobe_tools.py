"""tools/sassprobe: the cubin field patcher and the whole-kernel register renamer work on the built library (no GPU:
everything is checked through cuobjdump).  The renamer's own proof -- the patched function must disassemble to the
original text with the register names substituted -- is what makes its GPU sweeps trustworthy, so it is kept alive here."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "sassprobe"))
KERNEL = "force_cells_kernelILi2ELi128ELi256ELi1ELi8ELi2ELb1ELb0ELb0ELb0ELb0ELi2048E"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")


def test_register_renamer_proves_its_edits_and_refuses_unsafe_ones(mapc):
    from regrename import Renamer
    r = Renamer(mapc.LIB_PATH, KERNEL)
    assert len(r.dis) * 16 == r.size and r.regs[0] == 0
    quads = sorted(r.quads)
    assert quads, "the kernel has .128 accesses: their quads must be known"
    free = [p for p in range(4, (r.regs[-1] + 1) // 2) if p // 2 not in r.quads]
    # a swap of two pairs that no .128 access names, and a swap of two whole quads: both proven as text
    if len(free) >= 2:
        a, b = free[0], free[1]
        assert len(r.patches({a: b, b: a})) > 0
    q0, q1 = quads[0], quads[1]
    assert len(r.patches({2 * q0: 2 * q1, 2 * q0 + 1: 2 * q1 + 1, 2 * q1: 2 * q0, 2 * q1 + 1: 2 * q0 + 1})) > 0
    # splitting a quad, or a map that is not a permutation, is refused before anything is written
    with pytest.raises(AssertionError):
        r.patches({2 * q0: 2 * q0 + 1, 2 * q0 + 1: 2 * q0})
    with pytest.raises(AssertionError):
        r.patches({2 * q0: 2 * q1})


def test_field_patcher_round_trips_through_the_disassembler(mapc, tmp_path):
    import re
    import subprocess
    import sasspatch as sp
    from regrename import Renamer
    r = Renamer(mapc.LIB_PATH, KERNEL)
    addr, text = next((a, t) for a, t in r.dis
                      if t.startswith("FFMA2") and ".reuse" not in t and len(re.findall(r"R\d+", t)) == 4)
    img = bytearray(r.so)
    f = sp.get_fields(img, r.off + addr)
    sp.patch(img, r.off + addr, reuse=2)
    path = tmp_path / "libpatched.so"
    path.write_bytes(img)
    out = subprocess.run(["cuobjdump", "-sass", "-fun", r.mangled, str(path)], capture_output=True, text=True).stdout
    line = next(ln for ln in out.split("\n") if re.search(r"/\*%04x\*/" % addr, ln))
    regs = re.findall(r"R(\d+)(\.reuse)?", line.split("*/", 1)[1])
    assert [int(x) for x, _ in regs[:4]] == [f["d"], f["a"], f["b"], f["c"]]
    assert [bool(flag) for _, flag in regs[:4]] == [False, False, True, False]      # the flag landed on operand b only
