"""ctypes face of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The LITERAL flavour is pinned bit for bit against the reference's own shader code
compiled for the CPU (oracle/ref_shader.py, tests/test_reference_shader.py), and by closed-form KATs
(tests/test_oracle_kat.py), a numpy transcription and an fp64 direct sum.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_float, c_int, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

LITERAL = 0
MIRRORED = 1
POSVELO_DTYPE = np.dtype([("pos", np.float32, 4), ("velo", np.float32, 4)])

_lib = None


def build() -> str:
    res = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout, res.stderr)
        raise RuntimeError("building liboracle.so failed")
    return LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = ctypes.CDLL(LIB_PATH)
    fp, ip, dp = POINTER(c_float), POINTER(c_int), POINTER(c_double)
    lib.mapo_num_tiles.restype = c_int
    lib.mapo_num_tiles.argtypes = [c_int]
    lib.mapo_default_segments.restype = c_int
    lib.mapo_default_segments.argtypes = [c_int]
    lib.mapo_segment_range.restype = None
    lib.mapo_segment_range.argtypes = [c_int, c_int, c_int, ip, ip]
    lib.mapo_num_targets.restype = c_int
    lib.mapo_num_targets.argtypes = [c_int, c_int]
    lib.mapo_body_body_interaction.restype = None
    lib.mapo_body_body_interaction.argtypes = [fp, fp, fp, c_float, c_int]
    lib.mapo_body_body_interaction_mirrored.restype = None
    lib.mapo_body_body_interaction_mirrored.argtypes = [fp, fp, fp, c_float]
    lib.mapo_accel_allpairs_scalar.restype = None
    lib.mapo_accel_allpairs_scalar.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]
    lib.mapo_default_chain.restype = c_int
    lib.mapo_default_chain.argtypes = []
    lib.mapo_accel_allpairs.restype = None
    lib.mapo_accel_allpairs.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int]
    lib.mapo_accel_allpairs_chunked.restype = None
    lib.mapo_accel_allpairs_chunked.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int]
    lib.mapo_step_allpairs_targets_chunked.restype = None
    lib.mapo_step_allpairs_targets_chunked.argtypes = [c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_int,
                                                       c_int, c_int, c_int, c_void_p]
    lib.mapo_accel_fp64.restype = None
    lib.mapo_accel_fp64.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int]
    lib.mapo_integrate.restype = None
    lib.mapo_integrate.argtypes = [c_void_p, fp, c_float, c_float, c_int, c_void_p]
    lib.mapo_step_allpairs.restype = None
    lib.mapo_step_allpairs.argtypes = [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_int, c_int, c_int]
    lib.mapo_step_allpairs_targets.restype = None
    lib.mapo_step_allpairs_targets.argtypes = [c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_int,
                                               c_int, c_int, c_void_p]
    lib.mapo_step_well.restype = None
    lib.mapo_step_well.argtypes = [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_int]
    lib.mapo_init_particles.restype = None
    lib.mapo_init_particles.argtypes = [c_void_p, ctypes.c_uint, ctypes.c_uint]
    lib.mapo_max_threads.restype = c_int
    lib.mapo_max_threads.argtypes = []
    _lib = lib
    return lib


def _pv(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != POSVELO_DTYPE:
        a = np.ascontiguousarray(a, dtype=np.float32).view(POSVELO_DTYPE).reshape(-1)
    return np.ascontiguousarray(a)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


def max_threads() -> int:
    return int(load().mapo_max_threads())


def default_segments(n: int) -> int:
    return int(load().mapo_default_segments(n))


def default_chain() -> int:
    """Sources per sequential accumulation chain of the canonical order (2,048)."""
    return int(load().mapo_default_chain())


def _chain(chunk) -> int:
    """chunk=None: the canonical chain length; 0: one chain per segment (not what the product does)."""
    return default_chain() if chunk is None else int(chunk)


def num_targets(n: int, n_active: int) -> int:
    return int(load().mapo_num_targets(n, n_active))


def segment_range(n_sources: int, S: int, s: int):
    j0, j1 = c_int(0), c_int(0)
    load().mapo_segment_range(n_sources, S, s, ctypes.byref(j0), ctypes.byref(j1))
    return j0.value, j1.value


def body_body_interaction(ai, bj, bi, mass=70000.0, particles=1, flavour=LITERAL):
    """One call of bodyBodyInteraction (nBodyGravityCS.hlsl:44-57); returns the updated ai."""
    a = np.array(ai, dtype=np.float32)
    j = np.array(list(bj) + [0.0] * (4 - len(bj)), dtype=np.float32)
    i = np.array(list(bi) + [0.0] * (4 - len(bi)), dtype=np.float32)
    fp = POINTER(c_float)
    if flavour == LITERAL:
        load().mapo_body_body_interaction(a.ctypes.data_as(fp), j.ctypes.data_as(fp), i.ctypes.data_as(fp),
                                          mass, particles)
    else:
        # the mirrored pair leaves the uniform mass out (the kernel scales each segment partial once):
        # evaluate the pair from zero, scale, then add -- what a one-pair segment does
        t = np.zeros(3, dtype=np.float32)
        load().mapo_body_body_interaction_mirrored(t.ctypes.data_as(fp), j.ctypes.data_as(fp),
                                                   i.ctypes.data_as(fp), mass)
        a = (a + t * np.float32(mass)).astype(np.float32)
    return a


def accel_allpairs(particles, n_sources=None, S=None, flavour=LITERAL, targets=None, threads=0,
                   scalar=False, chunk=None) -> np.ndarray:
    p = _pv(particles)
    n_sources = p.shape[0] if n_sources is None else n_sources
    S = default_segments(n_sources) if S is None else S
    if targets is None:
        t, nt, tp = None, p.shape[0], None
    else:
        t = np.ascontiguousarray(targets, dtype=np.int32)
        nt, tp = t.shape[0], _ptr(t)
    out = np.zeros((nt, 3), dtype=np.float32)
    if scalar:
        load().mapo_accel_allpairs_scalar(_ptr(p), n_sources, S, _chain(chunk), flavour, tp, nt, _ptr(out))
    else:
        load().mapo_accel_allpairs_chunked(_ptr(p), n_sources, S, _chain(chunk), flavour, tp, nt, _ptr(out), threads)
    return out


def accel_fp64(particles, n_sources=None, targets=None, threads=0) -> np.ndarray:
    p = _pv(particles)
    n_sources = p.shape[0] if n_sources is None else n_sources
    if targets is None:
        nt, tp = p.shape[0], None
    else:
        t = np.ascontiguousarray(targets, dtype=np.int32)
        nt, tp = t.shape[0], _ptr(t)
    out = np.zeros((nt, 3), dtype=np.float64)
    load().mapo_accel_fp64(_ptr(p), n_sources, tp, nt, _ptr(out), threads)
    return out


def step_allpairs(particles, n_active=None, dt=0.1, damping=1.0, S=None, flavour=LITERAL, threads=0,
                  out=None, chunk=None) -> np.ndarray:
    """One all-pairs step in the canonical order.  `out` (the side being overwritten) defaults to a copy of
    the input.  chunk: chain length (None = canonical 2,048; 0 = one chain per segment)."""
    p = _pv(particles)
    n = p.shape[0]
    n_active = n if n_active is None else n_active
    S = default_segments(min(n_active, n)) if S is None else S
    o = p.copy() if out is None else out
    nt = num_targets(n, n_active)
    new = np.zeros(nt, dtype=POSVELO_DTYPE)
    load().mapo_step_allpairs_targets_chunked(_ptr(p), min(n_active, n), None, nt, dt, damping, S, _chain(chunk),
                                              flavour, threads, _ptr(new))
    o[:nt] = new
    return o


def step_allpairs_targets(particles, targets, n_sources=None, dt=0.1, damping=1.0, S=None, flavour=LITERAL,
                          threads=0, chunk=None) -> np.ndarray:
    p = _pv(particles)
    n_sources = p.shape[0] if n_sources is None else n_sources
    S = default_segments(n_sources) if S is None else S
    t = np.ascontiguousarray(targets, dtype=np.int32)
    o = np.zeros(t.shape[0], dtype=POSVELO_DTYPE)
    load().mapo_step_allpairs_targets_chunked(_ptr(p), n_sources, _ptr(t), t.shape[0], dt, damping, S, _chain(chunk),
                                              flavour, threads, _ptr(o))
    return o


def step_well(particles, n_active=None, dt=0.1, damping=1.0, flavour=LITERAL, out=None) -> np.ndarray:
    p = _pv(particles)
    n = p.shape[0]
    n_active = n if n_active is None else n_active
    o = p.copy() if out is None else out
    load().mapo_step_well(_ptr(p), _ptr(o), n, n_active, dt, damping, flavour)
    return o


def per_body_report(got, ref, before, floor_frac=1e-3) -> dict:
    """Per-body companion of rel_errors (which normalises by the GLOBAL maximum, so that `pos` -- of order
    the sphere radius against displacements of order 1 -- can hardly fail):
      accel_rel_l2_{p50,p99,max}  per body  |dv_got - dv_ref|_2 / max(|dv_ref|_2, floor), dv = velo - velo_before
                                  (= accel * dt * damping: the acceleration itself), floor = floor_frac *
                                  median |dv_ref|_2 so that bodies whose net force cancels do not divide by ~0
      pos_ulp_max                 largest |pos_got - pos_ref|_inf in units of the fp32 spacing at the body's
                                  largest coordinate: the step adds a displacement far smaller than pos, so
                                  anything beyond a couple of ulps means a wrong displacement, not rounding
      accel_len_rel_{p99,max}     per body |w_got - w_ref| / max(w_ref, floor) for pos.w = |accel|"""
    g, r, b = _pv(got), _pv(ref), _pv(before)
    dv_g = g["velo"][:, :3].astype(np.float64) - b["velo"][:, :3].astype(np.float64)
    dv_r = r["velo"][:, :3].astype(np.float64) - b["velo"][:, :3].astype(np.float64)
    mag = np.linalg.norm(dv_r, axis=1)
    floor = floor_frac * float(np.median(mag)) if mag.size else 0.0
    rel = np.linalg.norm(dv_g - dv_r, axis=1) / np.maximum(mag, max(floor, 1e-300))
    pr = r["pos"][:, :3]
    ulp = np.spacing(np.abs(pr).max(axis=1).astype(np.float32)).astype(np.float64)
    pos_ulps = np.abs(g["pos"][:, :3].astype(np.float64) - pr.astype(np.float64)).max(axis=1) / ulp
    w_g, w_r = g["pos"][:, 3].astype(np.float64), r["pos"][:, 3].astype(np.float64)
    w_floor = floor_frac * float(np.median(w_r)) if w_r.size else 0.0
    w_rel = np.abs(w_g - w_r) / np.maximum(w_r, max(w_floor, 1e-300))
    return {"accel_rel_l2_p50": float(np.percentile(rel, 50)), "accel_rel_l2_p99": float(np.percentile(rel, 99)),
            "accel_rel_l2_max": float(rel.max()), "pos_ulp_max": float(pos_ulps.max()),
            "accel_len_rel_p99": float(np.percentile(w_rel, 99)), "accel_len_rel_max": float(w_rel.max())}


def init_particles(n: int, seed: int) -> np.ndarray:
    """The reference's initial conditions (two shells, Compute.cpp:719-749, :820-844), seeded per particle."""
    out = np.zeros(n, dtype=POSVELO_DTYPE)
    load().mapo_init_particles(_ptr(out), n, seed & 0xFFFFFFFF)
    return out


def rel_errors(got, ref) -> dict:
    """Error metrics used by the parity tests (documented in DESIGN.md):
    per quantity q in {pos.xyz, velo.xyz, pos.w}:  max_i |got_i - ref_i|_inf / max_i |ref_i|_inf.
    The norm is global (the north star's "max relative error"); per_body_report() is the per-body view."""
    g, r = _pv(got), _pv(ref)
    out = {}
    for name, gq, rq in (("pos", g["pos"][:, :3], r["pos"][:, :3]),
                         ("velo", g["velo"][:, :3], r["velo"][:, :3]),
                         ("accel_len", g["pos"][:, 3:4], r["pos"][:, 3:4])):
        num = np.abs(gq.astype(np.float64) - rq.astype(np.float64)).max() if gq.size else 0.0
        den = np.abs(rq.astype(np.float64)).max() if rq.size else 0.0
        out[name] = float(num / den) if den > 0 else float(num)
    return out
