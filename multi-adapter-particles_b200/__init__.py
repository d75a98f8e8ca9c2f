"""B200-native Compute component for the Multi-Adapter-Particles n-body step.

Thin ctypes host layer over the C ABI in ``include/mapc.h`` (``lib/libmapc.so``, built from
``csrc/`` for sm_100a).  It mirrors the reference's ``Compute`` class
(``Particles/Compute.h:33-78``): same method names, argument meaning and fail-fast error
behaviour (the reference throws ``HrException`` from every failing D3D call,
``dx-samples-include/DXSampleHelper.h:22-46``; here every non-zero ``mapc_status`` raises
``MapcError``).

There is no CPU fallback and nothing here imports ``oracle/``: if ``libmapc.so`` is missing the
import of the library raises, and without a CUDA device every call that needs one raises.

The directory name is not a Python identifier; load the package with
``importlib.import_module("multi-adapter-particles_b200")``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import (POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_uint32,
                    c_uint64, c_void_p)

import numpy as np

from . import dist, ic  # noqa: F401  (re-exported)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
# MAPC_LIB_PATH: load another build of the library (A/B tools such as tools/sassprobe/force_sweep.py); default: in-tree
LIB_PATH = os.environ.get("MAPC_LIB_PATH") or os.path.join(PKG_DIR, "lib", "libmapc.so")

# constants carried over from the reference (include/mapc.h cites each)
BLOCK_SIZE = 64
SOFTENING_SQUARED = 25.0
PARTICLE_MASS = 70000.0
DEFAULT_DELTA_TIME = 0.1
DEFAULT_DAMPING = 1.0
MIN_NUM_PARTICLES = 256 * 1024
MAX_NUM_PARTICLES = 4 * 1024 * 1024
NCCL_UNIQUE_ID_BYTES = 128
IPC_BLOB_BYTES = 256

FORCE_ALLPAIRS = 0
FORCE_WELL = 1
CONSUMER_ASYNC = 1      # mapc_consumer_create_ex flag: the reference's async mode (same device, no copies)

#: numpy view of ``struct PosVelo { float4 pos; float4 velo; }`` (ParticleShared.hlsl:12-16)
POSVELO_DTYPE = np.dtype([("pos", np.float32, 4), ("velo", np.float32, 4)])

STATUS_NAMES = {
    0: "MAPC_OK", 1: "MAPC_ERR_INVALID_ARGUMENT", 2: "MAPC_ERR_CUDA", 3: "MAPC_ERR_NCCL",
    4: "MAPC_ERR_NO_DEVICE", 5: "MAPC_ERR_UNSUPPORTED", 6: "MAPC_ERR_TIMEOUT",
    7: "MAPC_ERR_OUT_OF_MEMORY",
}

#: every symbol include/mapc.h declares (tests check the built library exports each one)
EXPORTED_SYMBOLS = (
    "mapc_last_error", "mapc_version", "mapc_device_count",
    "mapc_fence_create", "mapc_fence_destroy", "mapc_fence_completed_value",
    "mapc_fence_signal_host", "mapc_fence_wait_host", "mapc_fence_signal_stream",
    "mapc_fence_wait_stream",
    "mapc_compute_create", "mapc_nccl_unique_id", "mapc_compute_create_sharded",
    "mapc_compute_destroy", "mapc_compute_upload", "mapc_compute_download", "mapc_compute_shard",
    "mapc_compute_set_force_mode", "mapc_compute_simulate", "mapc_compute_fence_value",
    "mapc_compute_wait_for_gpu", "mapc_compute_shared_handles", "mapc_compute_gpu_times",
    "mapc_compute_copy_state", "mapc_compute_init_particles", "mapc_plan_segments",
    "mapc_compute_kernel_launches", "mapc_compute_plan", "mapc_fp32_peak_probe",
    "mapc_compute_step_times", "mapc_compute_flush",
    "mapc_consumer_create", "mapc_consumer_destroy", "mapc_consumer_draw", "mapc_consumer_latest",
    "mapc_consumer_wait_for_gpu", "mapc_consumer_counters",
    "mapc_compute_ipc_export", "mapc_compute_ipc_attach", "mapc_compute_simulate_steps",
    "mapc_compute_exchange_times", "mapc_plan_chain_sources", "mapc_consumer_create_ex",
)


class MapcError(RuntimeError):
    """A C-ABI call returned a non-zero mapc_status."""

    def __init__(self, status: int, message: str):
        self.status = status
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")


class SharedHandlesStruct(Structure):
    """``mapc_shared_handles`` (``Compute::SharedHandles``, Particles/Compute.h:54-61)."""
    _fields_ = [
        ("posvelo", c_void_p * 2),
        ("packed_pos", c_void_p * 2),
        ("fence", c_void_p),
        ("compute_stream", c_void_p),
        ("aligned_data_size", c_uint64),
        ("buffer_index", c_uint32),
        ("first_particle", c_uint32),
        ("num_local", c_uint32),
        ("device", c_int32),
    ]


_lib = None


def build(verbose: bool = False) -> str:
    """Compile csrc/ into lib/libmapc.so with nvcc for sm_100a (works without a GPU)."""
    res = subprocess.run(["make", "-C", os.path.join(PKG_DIR, "csrc")], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libmapc.so failed")
    return LIB_PATH


def load() -> ctypes.CDLL:
    """dlopen lib/libmapc.so and declare the prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension is not built (run __graft_entry__.build()); "
            "there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    P = POINTER
    protos = {
        "mapc_last_error": (c_char_p, []),
        "mapc_version": (c_char_p, []),
        "mapc_device_count": (c_int, [P(c_int)]),
        "mapc_fence_create": (c_int, [P(c_void_p), c_uint64]),
        "mapc_fence_destroy": (c_int, [c_void_p]),
        "mapc_fence_completed_value": (c_uint64, [c_void_p]),
        "mapc_fence_signal_host": (c_int, [c_void_p, c_uint64]),
        "mapc_fence_wait_host": (c_int, [c_void_p, c_uint64, c_int]),
        "mapc_fence_signal_stream": (c_int, [c_void_p, c_void_p, c_uint64]),
        "mapc_fence_wait_stream": (c_int, [c_void_p, c_void_p, c_uint64]),
        "mapc_compute_create": (c_int, [P(c_void_p), c_uint32, c_int, c_void_p]),
        "mapc_nccl_unique_id": (c_int, [c_void_p]),
        "mapc_compute_create_sharded": (c_int, [P(c_void_p), c_uint32, c_int, c_int, c_int, c_void_p]),
        "mapc_compute_destroy": (c_int, [c_void_p]),
        "mapc_compute_upload": (c_int, [c_void_p, c_void_p, c_uint32]),
        "mapc_compute_download": (c_int, [c_void_p, c_void_p, c_uint32, c_uint32]),
        "mapc_compute_shard": (c_int, [c_void_p, P(c_uint32), P(c_uint32)]),
        "mapc_compute_set_force_mode": (c_int, [c_void_p, c_int]),
        "mapc_compute_simulate": (c_int, [c_void_p, c_int, c_float, c_float, c_uint64]),
        "mapc_compute_fence_value": (c_uint64, [c_void_p]),
        "mapc_compute_wait_for_gpu": (c_int, [c_void_p]),
        "mapc_compute_shared_handles": (c_int, [c_void_p, c_void_p, P(SharedHandlesStruct)]),
        "mapc_compute_gpu_times": (c_int, [c_void_p, P(c_float), P(c_float)]),
        "mapc_compute_copy_state": (c_int, [c_void_p, c_void_p]),
        "mapc_compute_init_particles": (c_int, [c_void_p, c_uint32]),
        "mapc_plan_segments": (c_int, [c_uint32]),
        "mapc_compute_kernel_launches": (c_uint64, [c_void_p]),
        "mapc_compute_plan": (c_int, [c_void_p, c_int, P(c_int), P(c_int), P(c_int), P(c_int)]),
        "mapc_fp32_peak_probe": (c_int, [c_int, c_int, P(c_float), P(c_float)]),
        "mapc_compute_step_times": (c_int, [c_void_p, P(c_float), c_int, P(c_int)]),
        "mapc_compute_flush": (c_int, [c_void_p]),
        "mapc_consumer_create": (c_int, [P(c_void_p), c_void_p, c_int]),
        "mapc_consumer_destroy": (c_int, [c_void_p]),
        "mapc_consumer_draw": (c_int, [c_void_p, c_int, P(c_uint64), c_int]),
        "mapc_consumer_latest": (c_int, [c_void_p, P(P(c_float)), P(c_uint64), P(c_uint32)]),
        "mapc_consumer_wait_for_gpu": (c_int, [c_void_p]),
        "mapc_consumer_counters": (c_int, [c_void_p, P(c_uint64)]),
        "mapc_compute_ipc_export": (c_int, [c_void_p, c_void_p]),
        "mapc_compute_ipc_attach": (c_int, [c_void_p, c_void_p, c_int]),
        "mapc_compute_simulate_steps": (c_int, [c_void_p, c_int, c_float, c_float, c_uint64, c_int]),
        "mapc_compute_exchange_times": (c_int, [c_void_p, P(c_float), P(c_float)]),
        "mapc_plan_chain_sources": (c_int, []),
        "mapc_consumer_create_ex": (c_int, [P(c_void_p), c_void_p, c_int, c_uint32]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(status: int) -> None:
    if status != 0:
        raise MapcError(status, load().mapc_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = c_int(0)
    _check(load().mapc_device_count(byref(n)))
    return n.value


def plan_segments(n_sources: int) -> int:
    """Canonical number of j segments for ``n_sources`` sources (independent of the GPU count)."""
    return int(load().mapc_plan_segments(n_sources))


def plan_chain_sources() -> int:
    """Sources per sequential accumulation chain of the canonical order (MAPC_CHAIN_SOURCES)."""
    return int(load().mapc_plan_chain_sources())


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
    _check(load().mapc_nccl_unique_id(buf))
    return buf.raw


def fp32_peak_probe(device: int = 0, packed: bool = True):
    """(TFLOP/s, ms) of the pure FFMA / FFMA2 microbenchmark kernel."""
    tf, ms = c_float(0), c_float(0)
    _check(load().mapc_fp32_peak_probe(device, 1 if packed else 0, byref(tf), byref(ms)))
    return tf.value, ms.value


def as_posvelo(arr) -> np.ndarray:
    """View/convert an (N, 8) float32 array or a structured array as contiguous PosVelo[N]."""
    a = np.asarray(arr)
    if a.dtype == POSVELO_DTYPE:
        return np.ascontiguousarray(a)
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 8:
        raise ValueError("expected PosVelo[N] or float32[N, 8]")
    return a.view(POSVELO_DTYPE).reshape(-1)


class Fence:
    """``mapc_fence``: the ID3D12Fence of the port (monotonic 64-bit value, host + any device)."""

    def __init__(self, initial_value: int = 0, _borrowed: int | None = None):
        self._owned = _borrowed is None
        if _borrowed is None:
            h = c_void_p()
            _check(load().mapc_fence_create(byref(h), initial_value))
            self._h = h
        else:
            self._h = c_void_p(_borrowed)

    @property
    def handle(self) -> c_void_p:
        return self._h

    def GetCompletedValue(self) -> int:
        return int(load().mapc_fence_completed_value(self._h))

    def Signal(self, value: int) -> None:
        _check(load().mapc_fence_signal_host(self._h, value))

    def Wait(self, value: int, timeout_ms: int = 30000) -> None:
        _check(load().mapc_fence_wait_host(self._h, value, timeout_ms))

    def SignalOnStream(self, stream: int, value: int) -> None:
        _check(load().mapc_fence_signal_stream(self._h, c_void_p(stream), value))

    def WaitOnStream(self, stream: int, value: int) -> None:
        _check(load().mapc_fence_wait_stream(self._h, c_void_p(stream), value))

    def close(self) -> None:
        if self._owned and self._h:
            load().mapc_fence_destroy(self._h)
        self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SharedHandles:
    """Python face of ``Compute::SharedHandles`` (Particles/Compute.h:54-61)."""

    def __init__(self, s: SharedHandlesStruct):
        self.posvelo = (s.posvelo[0], s.posvelo[1])
        self.packed_pos = (s.packed_pos[0], s.packed_pos[1])
        self.m_fence = Fence(_borrowed=s.fence)
        self.compute_stream = s.compute_stream
        self.m_alignedDataSize = int(s.aligned_data_size)
        self.m_bufferIndex = int(s.buffer_index)
        self.first_particle = int(s.first_particle)
        self.num_local = int(s.num_local)
        self.device = int(s.device)


class Compute:
    """Mirror of ``class Compute`` (Particles/Compute.h:33-78) over the C ABI.

    ``Compute(numParticles, device, prev=None)`` replaces
    ``Compute(UINT numParticles, IDXGIAdapter1*, bool useIntelExt, Compute* prev)``.  Pass
    ``rank/world/nccl_id`` for the i-sharded multi-GPU form (one process per GPU).
    """

    def __init__(self, in_numParticles: int, device: int = 0, in_pCompute: "Compute | None" = None,
                 *, rank: int = 0, world: int = 1, nccl_id: bytes | None = None):
        self._lib = load()
        self._h = c_void_p()
        self.m_numParticles = int(in_numParticles)
        if world > 1:
            if in_pCompute is not None:
                raise ValueError("state migration into a sharded handle is not supported")
            _check(self._lib.mapc_compute_create_sharded(byref(self._h), in_numParticles, device, rank,
                                                         world, nccl_id))
        else:
            prev = in_pCompute._h if in_pCompute is not None else None
            _check(self._lib.mapc_compute_create(byref(self._h), in_numParticles, device, prev))
        first, count = c_uint32(0), c_uint32(0)
        _check(self._lib.mapc_compute_shard(self._h, byref(first), byref(count)))
        self.first_particle, self.num_local = first.value, count.value

    # ---- reference surface -------------------------------------------------------------------
    def Simulate(self, in_numActiveParticles: int, in_sharedFenceValue: int = 0,
                 deltaTime: float = DEFAULT_DELTA_TIME, damping: float = DEFAULT_DAMPING) -> None:
        """``Compute::Simulate`` (Compute.cpp:1009-1055); asynchronous."""
        _check(self._lib.mapc_compute_simulate(self._h, in_numActiveParticles, deltaTime, damping,
                                               in_sharedFenceValue))

    def SimulateSteps(self, in_numActiveParticles: int, steps: int, in_sharedFenceValue: int = 0,
                      deltaTime: float = DEFAULT_DELTA_TIME, damping: float = DEFAULT_DAMPING) -> None:
        """`steps` Simulate calls in one submission (mapc_compute_simulate_steps)."""
        _check(self._lib.mapc_compute_simulate_steps(self._h, in_numActiveParticles, deltaTime, damping,
                                                     in_sharedFenceValue, steps))

    def GetFenceValue(self) -> int:
        return int(self._lib.mapc_compute_fence_value(self._h))

    def WaitForGpu(self) -> None:
        _check(self._lib.mapc_compute_wait_for_gpu(self._h))

    def GetSharedHandles(self, in_fence: Fence | None = None) -> SharedHandles:
        s = SharedHandlesStruct()
        _check(self._lib.mapc_compute_shared_handles(self._h, in_fence.handle if in_fence else None,
                                                     byref(s)))
        self._consumer_fence = in_fence  # keep it alive while attached
        return SharedHandles(s)

    def GetGpuTimes(self):
        """[(seconds, name)] like ``AdapterShared::GetGpuTimes`` plus the last raw sample in ms."""
        avg, last = c_float(0), c_float(0)
        _check(self._lib.mapc_compute_gpu_times(self._h, byref(avg), byref(last)))
        return [(avg.value * 1e-3, "simulate ms")], last.value

    def CopyState(self, in_pCompute: "Compute") -> None:
        _check(self._lib.mapc_compute_copy_state(self._h, in_pCompute._h))

    def InitializeParticles(self, seed: int = 0) -> None:
        _check(self._lib.mapc_compute_init_particles(self._h, seed))

    # ---- headless extras ---------------------------------------------------------------------
    def SetForceMode(self, mode: int) -> None:
        _check(self._lib.mapc_compute_set_force_mode(self._h, mode))

    def Upload(self, particles) -> None:
        a = as_posvelo(particles)
        _check(self._lib.mapc_compute_upload(self._h, a.ctypes.data_as(c_void_p), a.shape[0]))

    def Download(self, first: int | None = None, count: int | None = None, out=None) -> np.ndarray:
        first = self.first_particle if first is None else first
        count = self.num_local if count is None else count
        if out is None:
            out = np.empty(count, dtype=POSVELO_DTYPE)
        _check(self._lib.mapc_compute_download(self._h, out.ctypes.data_as(c_void_p), first, count))
        return out

    def Plan(self, n_active: int | None = None) -> dict:
        n_active = self.m_numParticles if n_active is None else n_active
        p, t, b, s = c_int(0), c_int(0), c_int(0), c_int(0)
        _check(self._lib.mapc_compute_plan(self._h, n_active, byref(p), byref(t), byref(b), byref(s)))
        return {"pairs_per_thread": p.value, "threads_per_block": t.value, "blocks": b.value,
                "segments": s.value}

    def IpcExport(self) -> bytes:
        """This rank's blob for the collective-free exchange (mapc_compute_ipc_export)."""
        buf = ctypes.create_string_buffer(IPC_BLOB_BYTES)
        _check(self._lib.mapc_compute_ipc_export(self._h, buf))
        return buf.raw

    def IpcAttach(self, blobs) -> None:
        """Attach the blobs of all ranks (rank order): steps then read peers' memory instead of all-gathering."""
        data = b"".join(blobs)
        _check(self._lib.mapc_compute_ipc_attach(self._h, data, len(blobs)))

    def StepTimes(self, capacity: int = 4096) -> np.ndarray:
        """Raw "simulate ms" samples resolved since the previous call (oldest first)."""
        buf = (c_float * capacity)()
        cnt = c_int(0)
        _check(self._lib.mapc_compute_step_times(self._h, buf, capacity, byref(cnt)))
        return np.array(buf[:cnt.value], dtype=np.float32)

    def ExchangeTimes(self):
        """(all-gather ms, ms it ran past the start of the consuming step) for the newest completed gather."""
        a, b = c_float(0), c_float(0)
        _check(self._lib.mapc_compute_exchange_times(self._h, byref(a), byref(b)))
        return a.value, b.value

    def Flush(self) -> None:
        _check(self._lib.mapc_compute_flush(self._h))

    def KernelLaunches(self) -> int:
        return int(self._lib.mapc_compute_kernel_launches(self._h))

    def close(self) -> None:
        if self._h:
            self._lib.mapc_compute_destroy(self._h)
            self._h = c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Consumer:
    """Headless stand-in for the reference's ``Render`` worker (Particles/Render.cpp:789-937): a
    copy stream pulls each step's positions to its own device and dumps them to pinned host memory,
    with the reference's copy-fence / render-fence protocol and one-frame latency."""

    def __init__(self, compute: Compute, device: int = 0, async_mode: bool = False):
        """``async_mode``: the reference's same-adapter mode (``Render::SetAsyncMode``, Render.cpp:849-852,
        :928-932): no copy stream, the producer's buffers are read in place."""
        self._lib = load()
        self._h = c_void_p()
        self._compute = compute
        self.async_mode = bool(async_mode)
        _check(self._lib.mapc_consumer_create_ex(byref(self._h), compute._h, device,
                                                 CONSUMER_ASYNC if async_mode else 0))

    def Draw(self, in_numActiveParticles: int, inout_fenceValue: int, in_numParticlesCopied: int | None = None) -> int:
        """``Render::Draw``; returns the fence value to hand to ``Compute.Simulate``."""
        f = c_uint64(inout_fenceValue)
        n_copy = in_numActiveParticles if in_numParticlesCopied is None else in_numParticlesCopied
        _check(self._lib.mapc_consumer_draw(self._h, in_numActiveParticles, byref(f), n_copy))
        return int(f.value)

    def Latest(self):
        """(frame number, float32[count, 4] copy of the newest completed position dump)."""
        ptr, frame, count = POINTER(c_float)(), c_uint64(0), c_uint32(0)
        _check(self._lib.mapc_consumer_latest(self._h, byref(ptr), byref(frame), byref(count)))
        arr = np.ctypeslib.as_array(ptr, shape=(count.value, 4)).copy() if count.value else np.zeros((0, 4), np.float32)
        return int(frame.value), arr

    def WaitForGpu(self) -> None:
        _check(self._lib.mapc_consumer_wait_for_gpu(self._h))

    def Counters(self) -> dict:
        buf = (c_uint64 * 8)()
        _check(self._lib.mapc_consumer_counters(self._h, buf))
        names = ("copy_fence_completed", "copy_fence_value", "render_fence_completed", "render_fence_value",
                 "shared_buffer_index", "current_buffer_index", "frames_drawn", "copies")
        return dict(zip(names, (int(v) for v in buf)))

    def close(self) -> None:
        if self._h:
            self._lib.mapc_consumer_destroy(self._h)
            self._h = c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
