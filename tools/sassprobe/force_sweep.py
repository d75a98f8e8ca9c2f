#!/usr/bin/env python
"""tools/sassprobe/force_sweep.py -- time the default force kernel under register renamings (regrename.py).

    python tools/sassprobe/force_sweep.py gen [K] [seed]   # here: K random renamings -> tools/sassprobe/force_variants.json
    python tools/sassprobe/force_sweep.py run              # on the GPU box: one line per variant (ms, bits equal?)
"""
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
LIB = os.path.join(ROOT, "multi-adapter-particles_b200", "lib", "libmapc.so")
KERNEL = "force_cells_kernelILi2ELi128ELi256ELi1ELi8ELi2ELb1ELb0ELb0ELb0ELb0ELi2048E"
VARIANTS = os.path.join(HERE, "force_variants.json")
QUAD_SLOTS = list(range(2, 15))       # R8..R59 as aligned quads (pairs 4..29)
LONE_PAIRS = []                       # R60:R61 are not named by the kernel (62 registers reported, R59 the highest used)


def random_map(rng, quads):
    """quads: quad indices (reg/4) that must stay whole.  Returns {pair -> pair} over pairs 4..30."""
    free_pairs = [p for p in range(4, 30) if p // 2 not in quads]
    slots = QUAD_SLOTS[:]
    rng.shuffle(slots)
    m = {}
    for q, s in zip(sorted(quads), slots):
        m[2 * q], m[2 * q + 1] = 2 * s, 2 * s + 1
    rest = [p for s in slots[len(quads):] for p in (2 * s, 2 * s + 1)] + LONE_PAIRS
    rng.shuffle(rest)
    for p, t in zip(free_pairs, rest):
        m[p] = t
    return m


def gen(k, seed, extra=()):
    from regrename import Renamer
    r = Renamer(LIB, KERNEL)
    rng = random.Random(seed)
    out = [{"name": "identity", "map": {}, "patches": []}]
    maps = list(extra) + [("r%03d" % i, random_map(rng, r.quads)) for i in range(k)]
    for name, m in maps:
        p = r.patches(m)
        out.append({"name": name, "map": {str(a): b for a, b in m.items()}, "patches": [[o, b.hex()] for o, b in p]})
    json.dump({"lib_size": len(r.so), "variants": out}, open(VARIANTS, "w"))
    print(len(out), "variants ->", VARIANTS, file=sys.stderr)


def run():
    spec = json.load(open(VARIANTS))
    so = open(LIB, "rb").read()
    assert len(so) == spec["lib_size"], "variants were generated for another build of libmapc.so"
    base = None
    for v in spec["variants"]:
        img = bytearray(so)
        for o, b in v["patches"]:
            img[o:o + 16] = bytes.fromhex(b)
        d = "/tmp/mapc_variants/" + v["name"]
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "libmapc.so")
        open(path, "wb").write(img)
        os.chmod(path, 0o755)
        res = subprocess.run([sys.executable, os.path.join(HERE, "time_lib.py")] + sys.argv[2:], capture_output=True, text=True,
                             env=dict(os.environ, MAPC_LIB_PATH=path), timeout=300)
        line = res.stdout.strip().split("\n")[-1] if res.stdout.strip() else "ERR " + res.stderr.strip()[-200:]
        digest = line.split()[2] if len(line.split()) > 2 else None
        if base is None:
            base = digest
        print(v["name"], line, "bits_equal" if digest == base else "BITS_DIFFER", json.dumps(v["map"], separators=(",", ":")), flush=True)
        os.remove(path)


def gen_attribution():
    """Timing-only patches of the 31-instruction hot loop (results are NOT the kernel's): what would the loop cost if
    its accumulations fetched two register pairs instead of three, or never hit the reuse cache?"""
    import re
    import sasspatch as sp
    from regrename import Renamer
    r = Renamer(LIB, KERNEL)
    bra = [(k, a, t) for k, (a, t) in enumerate(r.dis) if t.startswith("BRA.U UP0")]
    loops = []
    for k, a, t in bra:
        tgt = int(re.search(r"0x([0-9a-f]+)", t).group(1), 16)
        body = [(x, y) for x, y in r.dis if tgt <= x <= a]
        if tgt < a and sum(y.startswith("FFMA2") for _, y in body) == 12 and len(body) == 31:
            loops.append(body)
    assert len(loops) == 1, len(loops)
    acc = [(a, t) for a, t in loops[0] if t.startswith("FFMA2") and len(set(re.findall(r"R(\d+)", t.split(",", 1)[1]))) >= 3]
    assert len(acc) == 6
    out = [{"name": "identity", "map": {}, "patches": []}]
    for name in ("acc_two_pairs", "acc_no_reuse", "acc_reuse_all"):
        img = bytearray(r.so)
        p = []
        for a, t in acc:
            o = r.off + a
            f = sp.get_fields(img, o)
            if name == "acc_two_pairs":
                sp.patch(img, o, a=f["b"])
            elif name == "acc_no_reuse":
                sp.patch(img, o, reuse=0)
            else:
                sp.patch(img, o, reuse=2)
            p.append([o, bytes(img[o:o + 16]).hex()])
        out.append({"name": name, "map": {}, "patches": p})
    json.dump({"lib_size": len(r.so), "variants": out}, open(VARIANTS, "w"))
    print(len(out), "attribution variants ->", VARIANTS, file=sys.stderr)


def write(name, path):
    spec = json.load(open(VARIANTS))
    img = bytearray(open(LIB, "rb").read())
    v = next(v for v in spec["variants"] if v["name"] == name)
    for o, b in v["patches"]:
        img[o:o + 16] = bytes.fromhex(b)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, "wb").write(img)
    os.chmod(path, 0o755)


if __name__ == "__main__":
    if sys.argv[1] == "attribution":
        gen_attribution()
    elif sys.argv[1] == "write":
        write(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "gen":
        minimal = [("swap_pairs_25_26", {25: 26, 26: 25}), ("swap_quads_2_3", {4: 6, 5: 7, 6: 4, 7: 5}),
                   ("pair24_to_pair30", {24: 30, 30: 24}), ("swap_quads_7_12", {14: 24, 15: 25, 24: 14, 25: 15})]
        gen(int(sys.argv[2]) if len(sys.argv) > 2 else 40, int(sys.argv[3]) if len(sys.argv) > 3 else 1, minimal)
    else:
        run()
