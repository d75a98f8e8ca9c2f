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
LONE_PAIRS = [30]                     # R60:R61


def random_map(rng, quads):
    """quads: quad indices (reg/4) that must stay whole.  Returns {pair -> pair} over pairs 4..30."""
    free_pairs = [p for p in range(4, 31) if p // 2 not in quads]
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
    maps = list(extra) + [random_map(rng, r.quads) for _ in range(k)]
    for i, m in enumerate(maps):
        m = {a: b for a, b in m.items() if a != b or True}
        p = r.patches(m)
        out.append({"name": "v%03d" % i, "map": {str(a): b for a, b in m.items()}, "patches": [[o, b.hex()] for o, b in p]})
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


if __name__ == "__main__":
    if sys.argv[1] == "gen":
        gen(int(sys.argv[2]) if len(sys.argv) > 2 else 40, int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    else:
        run()
