"""The C-ABI shared library: loads, exports every symbol include/mapc.h declares, fails loudly
without a device (no CPU fallback), and its host-side plan logic agrees with the oracle's."""
import ctypes
import os
import re

import pytest

from conftest import REPO_ROOT


def declared_symbols():
    text = open(os.path.join(REPO_ROOT, "include", "mapc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"MAPC_API[^;(]*?\b(mapc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(mapc):
    lib = ctypes.CDLL(mapc.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mapc.h but not exported"
    assert set(names) == set(mapc.EXPORTED_SYMBOLS)


def test_posvelo_layout(mapc):
    # struct PosVelo {float4 pos; float4 velo;}: 32 bytes (ParticleShared.hlsl:12-16)
    assert mapc.POSVELO_DTYPE.itemsize == 32
    assert mapc.POSVELO_DTYPE.fields["velo"][1] == 16
    assert ctypes.sizeof(mapc.SharedHandlesStruct) == 72


def test_constants_match_reference(mapc):
    text = open(os.path.join(REPO_ROOT, "include", "mapc.h")).read()
    assert re.search(r"MAPC_BLOCK_SIZE\s+64\b", text)
    assert re.search(r"MAPC_SOFTENING_SQUARED\s+25\.0f", text)
    assert re.search(r"MAPC_PARTICLE_MASS\s+70000\.0f", text)
    assert re.search(r"MAPC_DEFAULT_DELTA_TIME\s+0\.1f", text)
    assert re.search(r"MAPC_DEFAULT_DAMPING\s+1\.0f", text)
    assert mapc.MIN_NUM_PARTICLES == 262144 and mapc.MAX_NUM_PARTICLES == 4194304


def test_plan_segments_matches_oracle_rule(mapc, oracle):
    for n in (1, 64, 1000, 10_000, 131_071, 131_072, 262_144, 262_145, 524_288, 524_289, 741_376, 1_048_576,
              2_097_153, 4_194_304, 8_000_000):
        assert mapc.plan_segments(n) == oracle.default_segments(n)
    # the frozen canonical order: 32 segments for every N, chains of 2,048 sources
    assert {mapc.plan_segments(n) for n in (10_000, 262_144, 370_688, 524_288, 741_376, 1_048_576, 4_194_304)} == {32}
    assert mapc.plan_chain_sources() == oracle.default_chain() == 2048


def test_no_cpu_fallback(mapc):
    """Without a CUDA device every compute entry point must raise, never compute on the host."""
    try:
        n = mapc.device_count()
    except mapc.MapcError as e:
        n = 0
        assert e.status in (2, 4)
    if n > 0:
        pytest.skip("a CUDA device is present; the no-device behaviour is checked on the CPU box")
    with pytest.raises(mapc.MapcError):
        mapc.Compute(1024, 0)
    with pytest.raises(mapc.MapcError):
        mapc.Fence(0)
    with pytest.raises(mapc.MapcError):
        mapc.fp32_peak_probe(0)


def test_invalid_arguments_are_reported(mapc):
    lib = mapc.load()
    h = ctypes.c_void_p()
    assert lib.mapc_compute_create(ctypes.byref(h), 0, 0, None) == 1
    assert b"num_particles" in lib.mapc_last_error()
    assert lib.mapc_compute_simulate(None, 1, 0.1, 1.0, 0) == 1
    assert lib.mapc_compute_wait_for_gpu(None) == 1
    assert lib.mapc_compute_destroy(None) == 0
    assert lib.mapc_fence_destroy(None) == 0


def test_ic_generators_are_deterministic(mapc):
    a = mapc.ic.uniform_sphere(1000, 100.0, seed=3)
    b = mapc.ic.uniform_sphere(1000, 100.0, seed=3)
    c = mapc.ic.uniform_sphere(1000, 100.0, seed=4)
    assert a.tobytes() == b.tobytes() and a.tobytes() != c.tobytes()
    import numpy as np
    r = np.linalg.norm(a["pos"][:, :3], axis=1)
    assert r.max() <= 100.0 * (1 + 1e-6) and np.all(a["pos"][:, 3] == 0)
    p = mapc.ic.plummer(2000, 50.0, seed=1)
    rp = np.linalg.norm(p["pos"][:, :3], axis=1)
    assert rp.max() <= 500.0 * (1 + 1e-5)
    assert 30.0 < np.median(rp) < 90.0   # Plummer half-mass radius ~ 1.3 a


def test_missing_extension_fails_loudly(mapc, monkeypatch, tmp_path):
    """No silent fallback: if libmapc.so is not there, loading the product raises."""
    monkeypatch.setattr(mapc, "_lib", None)
    monkeypatch.setattr(mapc, "LIB_PATH", str(tmp_path / "libmapc.so"))
    with pytest.raises(ImportError):
        mapc.load()
    with pytest.raises(ImportError):
        mapc.Compute(64, 0)


def test_header_is_plain_c_and_links(mapc, tmp_path):
    """include/mapc.h must be usable from C (the reference-side binding is C/C++): compile a C99 caller
    with -pedantic, link it against libmapc.so and run the calls that need no GPU."""
    import subprocess
    src = tmp_path / "caller.c"
    src.write_text(r"""
#include <stdio.h>
#include <string.h>
#include "mapc.h"
int main(void) {
    mapc_posvelo p; mapc_shared_handles h; mapc_compute *c = 0;
    memset(&p, 0, sizeof p); memset(&h, 0, sizeof h);
    if (sizeof(mapc_posvelo) != 32) return 2;
    if (mapc_plan_segments(262144u) != 32 || mapc_plan_segments(4194304u) != MAPC_MAX_SEGMENTS) return 3;
    if (mapc_plan_chain_sources() != MAPC_CHAIN_SOURCES) return 3;
    if (mapc_compute_create(&c, 0u, 0, 0) != MAPC_ERR_INVALID_ARGUMENT) return 4;
    if (strstr(mapc_last_error(), "num_particles") == 0) return 5;
    printf("%s\n", mapc_version());
    return 0;
}
""")
    exe = tmp_path / "caller"
    lib_dir = os.path.dirname(mapc.LIB_PATH)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(REPO_ROOT, "include"),
                    str(src), "-o", str(exe), "-L", lib_dir, "-lmapc", f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out
    assert "mapc" in out.stdout
