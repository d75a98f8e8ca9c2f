"""ctypes face of oracle/_ref/libref_shader.so: the reference's own compute shader
(Particles/nBodyGravityCS.hlsl) compiled for the CPU by oracle/Makefile.  TEST INFRASTRUCTURE ONLY.

The library exists only where it was built from /root/reference (this container; it travels to the GPU
box as a built file).  `available()` says whether it is there; nothing in the product may import this.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_float, c_int, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_shader.so")
POSVELO_DTYPE = np.dtype([("pos", np.float32, 4), ("velo", np.float32, 4)])

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(LIB_PATH + " (built by `make -C oracle` where /root/reference exists)")
        lib = ctypes.CDLL(LIB_PATH)
        fp = POINTER(c_float)
        lib.ref_constants.argtypes = [fp, fp]
        lib.ref_body_body_interaction.argtypes = [fp, fp, fp, c_float, c_int]
        lib.ref_csmain.argtypes = [c_void_p, c_void_p, c_int, c_float, c_float]
        lib.ref_accel_allpairs.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int]
        lib.ref_step_allpairs_targets.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_float,
                                                  c_int, c_void_p]
        for f in (lib.ref_constants, lib.ref_body_body_interaction, lib.ref_csmain, lib.ref_accel_allpairs,
                  lib.ref_step_allpairs_targets):
            f.restype = None
        _lib = lib
    return _lib


def _pv(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != POSVELO_DTYPE:
        a = np.ascontiguousarray(a, dtype=np.float32).view(POSVELO_DTYPE).reshape(-1)
    return np.ascontiguousarray(a)


def constants():
    """(softeningSquared, g_fParticleMass) as the shader defines them (nBodyGravityCS.hlsl:37-38)."""
    a, b = c_float(0), c_float(0)
    load().ref_constants(ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def body_body_interaction(ai, bj, bi, mass=70000.0, particles=1) -> np.ndarray:
    fp = POINTER(c_float)
    a = np.array(ai, dtype=np.float32)
    j = np.array(list(bj) + [0.0] * (4 - len(bj)), dtype=np.float32)
    i = np.array(list(bi) + [0.0] * (4 - len(bi)), dtype=np.float32)
    load().ref_body_body_interaction(a.ctypes.data_as(fp), j.ctypes.data_as(fp), i.ctypes.data_as(fp), mass, particles)
    return a


def csmain(particles, n_dispatch=None, dt=0.1, damping=1.0) -> np.ndarray:
    """The shipped CSMain (gravity well at the origin) over the first n_dispatch bodies."""
    p = _pv(particles)
    n = p.shape[0] if n_dispatch is None else n_dispatch
    out = p.copy()
    load().ref_csmain(p.ctypes.data_as(c_void_p), out.ctypes.data_as(c_void_p), n, dt, damping)
    return out


CHAIN_SOURCES = 2048   # canonical chain length (include/mapc.h MAPC_CHAIN_SOURCES)


def accel_allpairs(particles, S, n_sources=None, targets=None, threads=0, chain=CHAIN_SOURCES) -> np.ndarray:
    p = _pv(particles)
    n_sources = p.shape[0] if n_sources is None else n_sources
    if targets is None:
        nt, tp = p.shape[0], None
    else:
        t = np.ascontiguousarray(targets, dtype=np.int32)
        nt, tp = t.shape[0], t.ctypes.data_as(c_void_p)
    out = np.zeros((nt, 3), dtype=np.float32)
    load().ref_accel_allpairs(p.ctypes.data_as(c_void_p), n_sources, S, chain, tp, nt, out.ctypes.data_as(c_void_p),
                              threads)
    return out


def step_allpairs_targets(particles, targets, S, n_sources=None, dt=0.1, damping=1.0, threads=0,
                          chain=CHAIN_SOURCES) -> np.ndarray:
    """New state of the bodies `targets` (None = all) after one all-pairs step: OUR canonical loop around the
    reference's bodyBodyInteraction plus its integration lines (see ref_harness.cpp)."""
    p = _pv(particles)
    n_sources = p.shape[0] if n_sources is None else n_sources
    if targets is None:
        nt, tp = p.shape[0], None
    else:
        t = np.ascontiguousarray(targets, dtype=np.int32)
        nt, tp = t.shape[0], t.ctypes.data_as(c_void_p)
    out = np.zeros(nt, dtype=POSVELO_DTYPE)
    load().ref_step_allpairs_targets(p.ctypes.data_as(c_void_p), n_sources, tp, nt, S, chain, dt, damping, threads,
                                     out.ctypes.data_as(c_void_p))
    return out


def step_allpairs(particles, S, dt=0.1, damping=1.0, threads=0, chain=CHAIN_SOURCES) -> np.ndarray:
    return step_allpairs_targets(particles, None, S, dt=dt, damping=damping, threads=threads, chain=chain)
