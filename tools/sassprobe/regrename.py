#!/usr/bin/env python
"""tools/sassprobe/regrename.py -- rename the physical registers of ONE kernel inside a built libmapc.so.

ptxas picks the registers; on B200 the time of the force kernel's hot loop depends on which banks the three operand
pairs of its accumulations fall into (profiles/r02_regfile_probe.txt).  A renaming pi of a function's registers that
is a bijection, keeps aligned pairs together (64-bit operands) and keeps every quad that a .128 access touches an
aligned quad, applied to EVERY instruction of the function (callee subroutines inside the same section included),
yields the same program on other registers.  This tool applies such a renaming to the instruction bytes of one
function inside the .so (the cubin is stored uncompressed) and proves the result as text: the patched function is
disassembled again and must equal the original disassembly with the register names substituted -- nothing else may
differ, in this function or in any other.

    from regrename import Renamer; r = Renamer(path_to_so, kernel_substr); patches = r.patches({26: 25, 25: 26})
"""
import os
import re
import struct
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import sasspatch as sp  # noqa: E402

REG = re.compile(r"(?<![A-Za-z0-9_])R(\d+)")


def disasm_function(path, mangled):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, path], capture_output=True, text=True).stdout
    res = []
    for ln in out.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", ln)
        if m:
            res.append((int(m.group(1), 16), m.group(2).strip()))
    return res


class Renamer:
    def __init__(self, so_path, kernel_substr):
        self.so_path = so_path
        self.so = open(so_path, "rb").read()
        with tempfile.TemporaryDirectory() as d:
            subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=d, capture_output=True, check=True)
            hits = []
            # both embedded ELF files carry the same name and overwrite each other: find every ELF image in the .so instead
            pos = self.so.find(b"\x7fELF", 1)
            while pos >= 0:
                if self.so[pos + 4] == 2 and struct.unpack_from("<H", self.so, pos + 0x12)[0] == 190:   # EM_CUDA
                    try:
                        off, size, name = sp.text_section(self.so[pos:], kernel_substr)
                        hits.append((pos + off, size, name))
                    except AssertionError:
                        pass
                pos = self.so.find(b"\x7fELF", pos + 1)
        assert len(hits) == 1, hits
        self.off, self.size, name = hits[0]
        self.mangled = name[len(".text."):]
        self.dis = disasm_function(so_path, self.mangled)
        assert len(self.dis) * 16 == self.size, (len(self.dis), self.size)
        # quads that a 128-bit access names: they must stay aligned quads
        self.quads = set()
        for _, t in self.dis:
            op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
            if ".128" in op and not op.startswith(("LDCU", "ULD")):
                toks = [int(x) for x in REG.findall(t)]
                data = toks[0] if op.startswith("LD") else toks[-1]
                assert data % 4 == 0, t
                self.quads.add(data // 4)
        self.regs = sorted({int(x) for _, t in self.dis for x in REG.findall(t)})
        # which of the four 8-bit fields of each instruction name a general register: flip bit 6 of one field in every
        # instruction at once, disassemble, and see whether exactly one R<n> token followed (a uniform register, an
        # immediate or an opcode bit that happens to hold the same number does not)
        self.is_reg = {a: {} for a, _ in self.dis}
        for k in "dabc":
            img = bytearray(self.so)
            for a, t in self.dis:
                v = sp.get_fields(img, self.off + a)[k]
                if v in {int(x) for x in REG.findall(t)}:       # only where the field can be a register at all
                    sp.patch(img, self.off + a, **{k: v ^ 64})
            with tempfile.NamedTemporaryFile(suffix=".so") as tmp:
                tmp.write(img); tmp.flush()
                new = disasm_function(tmp.name, self.mangled)
            assert len(new) == len(self.dis), "cuobjdump could not decode the probe image"
            new = dict(new)
            for a, t0 in self.dis:
                v = sp.get_fields(self.so, self.off + a)[k]
                t1 = new.get(a)
                ok = False
                if t1 is not None and v != 255:
                    for m in REG.finditer(t0):
                        if int(m.group(1)) == v and t0[:m.start()] + "R%d" % (v ^ 64) + t0[m.end():] == t1:
                            ok = True
                self.is_reg[a][k] = ok

    def check_map(self, pair_map):
        """pair_map: {pair index -> pair index}; must be a permutation of its keys that keeps quads whole."""
        assert sorted(pair_map) == sorted(pair_map.values()), "not a permutation"
        for q in self.quads:
            lo, hi = pair_map.get(2 * q, 2 * q), pair_map.get(2 * q + 1, 2 * q + 1)
            assert lo % 2 == 0 and hi == lo + 1, f"quad R{4*q} would be split: pairs {2*q},{2*q+1} -> {lo},{hi}"
            # the target quad must be wholly vacated by the map or be the same one: guaranteed by bijectivity

    def patches(self, pair_map):
        """[(offset in the .so, 16 new bytes)] that apply the renaming; verified through the disassembler."""
        self.check_map(pair_map)
        ren = lambda r: 2 * pair_map.get(r // 2, r // 2) + (r & 1)
        img = bytearray(self.so)
        out = []
        for addr, text in self.dis:
            toks = {int(x) for x in REG.findall(text)}
            moved = {r for r in toks if ren(r) != r}
            if not moved:
                continue
            o = self.off + addr
            f = sp.get_fields(img, o)
            kw = {k: ren(f[k]) for k in "dabc" if self.is_reg[addr][k] and f[k] in moved}
            assert kw, (hex(addr), text)
            sp.patch(img, o, **kw)
            out.append((o, bytes(img[o:o + 16])))
        with tempfile.NamedTemporaryFile(suffix=".so") as tmp:
            tmp.write(img); tmp.flush()
            new = disasm_function(tmp.name, self.mangled)
        assert len(new) == len(self.dis)
        for (a0, t0), (a1, t1) in zip(self.dis, new):
            want = REG.sub(lambda m: "R%d" % ren(int(m.group(1))), t0)
            assert a0 == a1 and t1 == want, (hex(a0), t0, "->", t1, "expected", want)
        return out

    def apply(self, patches, dst_path):
        img = bytearray(self.so)
        for o, b in patches:
            img[o:o + 16] = b
        with open(dst_path, "wb") as f:
            f.write(img)
        os.chmod(dst_path, 0o755)


if __name__ == "__main__":
    r = Renamer(sys.argv[1], sys.argv[2])
    print(r.mangled, len(r.dis), "instructions; quads:", sorted(4 * q for q in r.quads), "regs:", r.regs[0], "..", r.regs[-1])
