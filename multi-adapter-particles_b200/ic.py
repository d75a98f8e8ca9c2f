"""Seeded synthetic initial conditions (host side, numpy) for tests, bench and the driver.

The reference's own generator (two shells, ``Particles/Compute.cpp:667-844``) is reproduced in
the C ABI (``mapc_compute_init_particles``).  The distributions here are the ones
``BASELINE.json`` names for the all-pairs workloads: uniform-in-volume sphere and Plummer.
Random numbers come from a counter-based splitmix64 stream, so the bytes depend only on
``(seed, n)`` -- not on the numpy version -- and CPU oracle and GPU are fed identical arrays.
"""
from __future__ import annotations

import numpy as np

POSVELO_DTYPE = np.dtype([("pos", np.float32, 4), ("velo", np.float32, 4)])
PARTICLE_MASS = 70000.0  # Particles/nBodyGravityCS.hlsl:38


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniforms(seed: int, n: int, streams: int) -> np.ndarray:
    """float64 uniforms in (0, 1), shape (streams, n): value k of stream s is a hash of (seed, s, k)."""
    idx = np.arange(n, dtype=np.uint64)
    out = np.empty((streams, n), dtype=np.float64)
    for s in range(streams):
        with np.errstate(over="ignore"):
            key = _splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(s + 1))
            bits = _splitmix64(idx ^ key)
        out[s] = ((bits >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / (1 << 53))
    return out


def _pack(pos: np.ndarray, vel: np.ndarray) -> np.ndarray:
    n = pos.shape[0]
    out = np.zeros(n, dtype=POSVELO_DTYPE)
    out["pos"][:, :3] = pos.astype(np.float32)
    out["velo"][:, :3] = vel.astype(np.float32)
    return out


def _directions(u_z: np.ndarray, u_phi: np.ndarray) -> np.ndarray:
    z = 2.0 * u_z - 1.0
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = 2.0 * np.pi * u_phi
    return np.stack([s * np.cos(phi), s * np.sin(phi), z], axis=1)


def uniform_sphere(n: int, radius: float, seed: int, speed: float = 0.0) -> np.ndarray:
    """Uniform-in-volume sphere of radius ``radius``; isotropic velocities of magnitude ``speed``."""
    u = uniforms(seed, n, 5)
    r = radius * np.cbrt(u[0])
    pos = _directions(u[1], u[2]) * r[:, None]
    vel = _directions(u[3], u[4]) * speed
    return _pack(pos, vel)


def plummer(n: int, scale_radius: float, seed: int, truncate: float = 10.0,
            velocity_scale: float = 1.0) -> np.ndarray:
    """Plummer sphere, r = a / sqrt(u^(-2/3) - 1) truncated at ``truncate``*a; isotropic Gaussian
    velocities with the local Plummer dispersion for total mass n * PARTICLE_MASS (G = 1)."""
    u = uniforms(seed, n, 9)
    x = truncate
    u_max = x ** 3 / (1.0 + x * x) ** 1.5
    m = u[0] * u_max
    r = scale_radius / np.sqrt(m ** (-2.0 / 3.0) - 1.0)
    pos = _directions(u[1], u[2]) * r[:, None]
    total_mass = n * PARTICLE_MASS
    sigma = np.sqrt(total_mass / (6.0 * scale_radius)) * (1.0 + (r / scale_radius) ** 2) ** -0.25
    # Box-Muller from the remaining streams
    g0 = np.sqrt(-2.0 * np.log(u[3])) * np.cos(2.0 * np.pi * u[4])
    g1 = np.sqrt(-2.0 * np.log(u[5])) * np.cos(2.0 * np.pi * u[6])
    g2 = np.sqrt(-2.0 * np.log(u[7])) * np.cos(2.0 * np.pi * u[8])
    vel = np.stack([g0, g1, g2], axis=1) * (sigma * velocity_scale)[:, None]
    return _pack(pos, vel)


def lattice_sphere(n: int, radius: float, seed: int, jitter: float = 0.25, speed: float = 0.0) -> np.ndarray:
    """``n`` bodies on a jittered cubic lattice clipped to a sphere, in seeded random order.

    Every pair is at least (1 - 2*jitter) lattice spacings apart, so there are no close
    encounters at the softening scale: multi-step runs stay well conditioned and a parity
    tolerance measures arithmetic, not chaotic amplification of the last bit."""
    spacing = radius * (4.0 / 3.0 * np.pi / (1.3 * n)) ** (1.0 / 3.0)
    while True:
        m = int(np.ceil(radius / spacing)) + 1
        g = np.arange(-m, m + 1, dtype=np.float64) * spacing
        x, y, z = np.meshgrid(g, g, g, indexing="ij")
        pts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
        pts = pts[(pts * pts).sum(axis=1) <= radius * radius]
        if pts.shape[0] >= n:
            break
        spacing *= 0.97
    u = uniforms(seed, pts.shape[0], 6)
    order = np.argsort(u[0], kind="stable")[:n]
    pts = pts[order] + (u[1:4, order].T - 0.5) * (2.0 * jitter * spacing)
    vel = _directions(u[4, order], u[5, order]) * speed
    return _pack(pts, vel)


#: the BASELINE.json / SURVEY.md section 8(d) workloads: name -> (N, generator)
def workload(name: str) -> np.ndarray:
    if name == "interactive_10k":      # config 2: N=10,000 uniform sphere R=2000 seed 1
        return uniform_sphere(10_000, 2000.0, 1)
    if name == "sphere_262144":        # config 3: N=262,144 uniform sphere R=8000 seed 2
        return uniform_sphere(262_144, 8000.0, 2)
    if name == "sphere_1048576":       # config 4: N=1,048,576 uniform sphere R=12,000 seed 3
        return uniform_sphere(1_048_576, 12_000.0, 3)
    if name == "plummer_4194304":      # config 5: N=4,194,304 Plummer a=10,000 seed 4
        return plummer(4_194_304, 10_000.0, 4)
    raise KeyError(name)
