#!/usr/bin/env python
"""bench.py -- the all-pairs gravity + integration step on N B200s (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W              one GPU, config 3 (N = 262,144)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W      one rank per GPU over NCCL
  python bench.py --impl reference ...                       the reference's shader code compiled for the
                                                             CPU (oracle/_ref; else the oracle port) timed
                                                             on the host cores, same metric

A "step" is one Simulate(N, dt, damping): force over all N^2 pairs + integration.  `value` is
whole-job G interactions/s with the state resident in HBM; `e2e` is the same metric through the
C ABI with HOST buffers (upload + simulate + download inside the timed region).  Rank 0 prints
ONE JSON line.  Only the cpu_baseline leg and --impl reference touch oracle/ (the checker).

Besides the headline (BASELINE config 3, weak-scaled with the GPU count so that the per-GPU work is fixed)
every run adds to the same line, each a short leg of its own:
  strong       BASELINE config 5, N = 4,194,304 Plummer at THIS GPU count (the 85 % target is T1 / (R * TR) of
               these lines across --gpus 1/2/4/8) and, on 8 GPUs, config 4 (N = 1,048,576) under `config4`
  determinism  sha256 of the full rank-ordered PosVelo state of a fixed problem (N = 65,536, 4 steps): the same
               string at every GPU count is the bit-identity claim, checked by comparing lines
  well         the step the reference actually dispatches (CSMain, nBodyGravityCS.hlsl:86-109) at the
               reference's default N = 4,194,304 (defines.h:45) as an HBM-bound kernel: GB/s against the
               measured copy peak (1 GPU runs only; `--mode well` makes it the headline instead)
  latency      BASELINE config 2: microseconds per step of a 1,000-step ping-pong run at N = 10,000 issued in batches
               of 50 (1 GPU runs only; `--bodies 10000 --batch 50` makes it the headline instead)
  cpu_baseline.parity   the GPU result of the headline workload against the oracle (LITERAL) and the fp64
               direct sum on a target sample -- a fast wrong kernel would show here
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "interactions_per_second"
UNIT = "G interactions/s"
FLOP_PER_INTERACTION = 20.0          # GPU-Gems convention named by BASELINE.json
BASE_N = 262_144                     # config 3, the configuration the 70 % target is quoted on
BASE_RADIUS = 8000.0
SEED = 2
DT, DAMPING = 0.1, 1.0               # Particles/Compute.cpp:545-546


STRONG_N = 4_194_304                 # config 5 / Particles/defines.h:45 MAX_NUM_PARTICLES (the reference's default)
CONFIG4_N = 1_048_576                # config 4
DETERMINISM_N, DETERMINISM_STEPS = 65_536, 4


def workload_n(pkg, world: int, scaling: str, n_override: int | None) -> int:
    """Weak scaling (default) keeps the per-GPU work fixed: N = 262,144 * sqrt(world) (dist.weak_scaled_n);
    --scaling strong runs BASELINE config 5's N = 4,194,304 at every GPU count."""
    if n_override:
        return n_override
    if scaling == "strong":
        return 4_194_304
    return pkg.dist.weak_scaled_n(world, BASE_N)


def make_particles(pkg, n: int) -> np.ndarray:
    radius = BASE_RADIUS * (n / BASE_N) ** (1.0 / 3.0)   # constant density -> same dynamics
    return pkg.ic.uniform_sphere(n, radius, SEED)


def read_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            d["_source"] = "MEASURED_PEAKS.json"
            return d
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "_source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(kernel: str, n: int, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` ("force" / "well") from the committed
    `ncu --set full` captures (profiles/ncu_traffic.json: a list, one entry per capture); null when no capture
    matches the workload (same kernel, same N, one GPU, the library's canonical order)."""
    try:
        entries = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        for t in entries if isinstance(entries, list) else [entries]:
            if world == 1 and t.get("kernel_kind") == kernel and int(t["n"]) == n and \
                    int(t.get("segments", 32)) == 32 and int(t.get("chain", 2048)) == 2048:
                return float(t["dram_bytes_per_launch"]), t.get("source")
    except Exception:
        pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.lines: list[str] = []
        self.proc = None
        self.thread = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()          # exactly the PID we started
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_leg(step_targets, what: str, particles: np.ndarray, seconds: float, threads: int, steps: int = 1,
            warmup: int = 0):
    """Times a CPU implementation on a bounded sample: `m` random targets against ALL sources, canonical
    segment order -- the same per-interaction work as the full step.  step_targets(particles, targets)."""
    n = particles.shape[0]
    rng = np.random.default_rng(1234)
    probe = np.sort(rng.choice(n, min(n, 64 * threads), replace=False)).astype(np.int32)
    t0 = time.perf_counter()
    step_targets(particles, probe)
    rate = probe.shape[0] * n / max(time.perf_counter() - t0, 1e-6)
    m = int(min(n, max(probe.shape[0], rate * seconds / n)))
    m -= m % 8 if m > 8 else 0
    targets = np.sort(rng.choice(n, m, replace=False)).astype(np.int32)
    times = []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        step_targets(particles, targets)
        if k >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = float(np.mean(times))
    ginter = m * n / per_step / 1e9
    sample = f"{m} random targets x {n} sources per step, {what}, {threads} OpenMP threads"
    return ginter, per_step, sample, m


def cpu_implementations(n: int):
    """The CPU arms, best evidence first: (kind, description, step_targets, threads).
    "reference" = the reference's own shader code (Particles/nBodyGravityCS.hlsl) compiled for the CPU into
    oracle/_ref/ (scalar per pair, OpenMP over targets); "port" = the oracle's LITERAL flavour, the same
    arithmetic bit for bit (tests/test_reference_shader.py) vectorised 8 targets wide."""
    orc = importlib.import_module("oracle.oracle_py")
    orc.load()
    # every core this process may run on, passed explicitly (num_threads clause): torchrun exports
    # OMP_NUM_THREADS=1, which would otherwise turn the CPU arm into a single-thread run
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    S = orc.default_segments(n)
    out = []
    refsh = importlib.import_module("oracle.ref_shader")
    if refsh.available():
        out.append(("reference", "reference shader compiled for the CPU (oracle/_ref, scalar per pair)",
                    lambda p, t: refsh.step_allpairs_targets(p, t, S, threads=threads), threads))
    out.append(("port", "oracle LITERAL flavour (8-wide SIMD over targets)",
                lambda p, t: orc.step_allpairs_targets(p, t, flavour=orc.LITERAL, threads=threads), threads))
    return out


def cpu_baseline(particles: np.ndarray, seconds: float, steps: int = 1, warmup: int = 0):
    """-> (cpu_baseline object, per-step seconds and sample size of the headline implementation).  The
    headline is the reference-compiled shader when it exists, and the faster vectorised port is reported
    next to it (`port_value`) so that nobody mistakes a slow scalar build for the CPU's best."""
    impls = cpu_implementations(particles.shape[0])
    kind, what, fn, threads = impls[0]
    g, per_step, sample, m = cpu_leg(fn, what, particles, seconds, threads, steps, warmup)
    obj = {"value": g, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
    if len(impls) > 1:
        _, what2, fn2, _ = impls[1]
        g2, _, sample2, _ = cpu_leg(fn2, what2, particles, max(1.0, seconds / 4), threads)
        obj["port_value"] = g2
        obj["port_sample"] = sample2
    return obj, per_step, m


def run_reference(args) -> None:
    """--impl reference: the reference executes its shader through D3D12 on Windows and cannot run here as
    shipped; its shader code compiled for the CPU (oracle/_ref, else the oracle port) is the reference arm,
    on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = importlib.import_module("multi-adapter-particles_b200")
    world = args.gpus
    n = workload_n(pkg, world, args.scaling, args.n)
    particles = make_particles(pkg, n)
    budget = max(2.0, min(20.0, 150.0 / (args.steps + args.warmup)))
    cpu, per_step, m = cpu_baseline(particles, budget, args.steps, args.warmup)
    ginter = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": ginter, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3 * (n / m),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "extrapolated": True,   # ms_per_step is the sample's time scaled to all N targets; `value` is measured
        "config": {"workload": f"allpairs uniform sphere N={n} seed={SEED} dt={DT} damping={DAMPING}",
                   "n": n, "note": "ms_per_step extrapolated from the sample to all N targets"},
        "cpu_baseline": cpu,
        "e2e": {"value": ginter, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Harness:
    """torch.distributed plumbing shared by the legs: one rank per GPU, barrier + device synchronize on both
    sides of a timed region, max over ranks of a device time."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run")
        self.pkg = importlib.import_module("multi-adapter-particles_b200")
        self.pkg.load()                              # raises if libmapc.so is missing: no fallback
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.args = args
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        return self.pkg.dist.max_over_ranks(x, self.dev)

    def compute(self, n: int, exchange: str = "nccl"):
        """A Compute for n bodies on this rank's GPU: unsharded on one GPU, else this rank's shard."""
        pkg = self.pkg
        nccl_id = None
        if self.world > 1:
            nccl_id = pkg.dist.broadcast_bytes(pkg.nccl_unique_id() if self.rank == 0 else None,
                                               pkg.NCCL_UNIQUE_ID_BYTES, 0, self.dev)
        return pkg.Compute(n, self.local_rank, rank=self.rank, world=self.world, nccl_id=nccl_id)

    def timed_steps(self, c, n: int, steps: int, warmup: int, flush: bool, batch: int = 1, sample_clocks: bool = False):
        """W untimed steps, then exactly K steps bracketed by barrier + synchronize, timed with CUDA events on
        the stream the kernels are launched on; -> (ms per step: max over ranks, in-kernel step times, launches,
        clocks summary or None)."""
        torch = self.torch
        stream = torch.cuda.ExternalStream(c.GetSharedHandles().compute_stream, device=self.dev)

        def l2_flush():
            if flush:
                with torch.cuda.stream(stream):
                    self.flush_buf.zero_()

        for _ in range(warmup):
            c.Simulate(n, 0, DT, DAMPING)
            l2_flush()
        c.WaitForGpu()
        c.StepTimes()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = c.KernelLaunches()
        sampler = ClockSampler(self.local_rank) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        try:
            e0.record(stream)
            if batch > 1:
                for k in range(0, steps, batch):
                    c.SimulateSteps(n, min(batch, steps - k), 0, DT, DAMPING)
            else:
                for k in range(steps):
                    c.Simulate(n, 0, DT, DAMPING)
                    if k + 1 < steps:
                        l2_flush()
            c.Flush()
            e1.record(stream)
            c.WaitForGpu()
            self.barrier()
        finally:
            if sampler:
                sampler.__exit__(None, None, None)
        ms = self.max_over_ranks(e0.elapsed_time(e1)) / steps
        return ms, c.StepTimes(), c.KernelLaunches() - launches0, (sampler.summary() if sampler else None)


def strong_leg(h: Harness, n: int, name: str, steps: int = 3, warmup: int = 3) -> dict:
    """A BASELINE strong-scaling configuration at this GPU count: N fixed, i-sharded over the ranks."""
    pkg = h.pkg
    particles = pkg.ic.workload(name)
    assert particles.shape[0] == n
    c = h.compute(n)
    c.Upload(particles)
    ms, _, _, _ = h.timed_steps(c, n, steps, warmup, flush=False)
    gather_ms, gather_tail_ms = c.ExchangeTimes() if h.world > 1 else (0.0, 0.0)
    mine = c.Download()
    checksum = float(np.abs(mine["pos"][:, :3].astype(np.float64)).sum())
    if h.world > 1:
        t = h.torch.tensor([checksum], dtype=h.torch.float64, device=h.dev)
        h.dist.all_reduce(t)
        checksum = float(t.item())
    c.close()
    h.barrier()
    value = float(n) * float(n) / (ms * 1e-3) / 1e9
    return {"workload": name, "n": n, "n_gpus": h.world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "value": value, "unit": UNIT, "tflops_at_20flop": value * FLOP_PER_INTERACTION / 1e3,
            "exchange": {"kind": "nccl", "allgather_ms": gather_ms, "allgather_past_step_begin_ms": gather_tail_ms}
                        if h.world > 1 else None,
            "l2": f"inputs {n * 16 / 1e6:.0f} MB > L2 for N = 4,194,304; no flush (a step is >= 0.7 s)",
            "checksum_abs_pos": checksum}


def determinism_leg(h: Harness) -> dict:
    """sha256 of the full PosVelo state, in body order, after DETERMINISM_STEPS steps of a fixed problem: the
    canonical summation order makes it the same string at every GPU count."""
    import hashlib
    pkg, torch, dist = h.pkg, h.torch, h.dist
    n = DETERMINISM_N
    particles = pkg.ic.uniform_sphere(n, 5000.0, seed=11, speed=1.0)
    c = h.compute(n)
    c.Upload(particles)
    for _ in range(DETERMINISM_STEPS):
        c.Simulate(n, 0, DT, DAMPING)
    c.WaitForGpu()
    mine = c.Download()
    c.close()
    if h.world > 1:
        t = torch.from_numpy(mine.view(np.uint8).copy()).to(h.dev)
        parts = [torch.empty_like(t) for _ in range(h.world)]
        dist.all_gather(parts, t)
        full = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
    else:
        full = mine.tobytes()
    h.barrier()
    return {"n": n, "steps": DETERMINISM_STEPS, "n_gpus": h.world, "sha256": hashlib.sha256(full).hexdigest(),
            "what": "full PosVelo state in body order (uniform sphere R=5000 seed 11 speed 1); equal strings "
                    "across --gpus 1/2/4/8 = bit-identical results"}


LATENCY_N = 10_000


def latency_leg(h: Harness, steps: int = 1000, batch: int = 50) -> dict:
    """BASELINE config 2 inside the default line: microseconds per step of a 1,000-step ping-pong run at N = 10,000
    (the same constant-density uniform sphere, issued in batches of 50: steps chained per target block, occupancy
    throttle -- DESIGN.md section 4).  Timed like the headline (CUDA events on the compute stream around exactly
    `steps` steps, no L2 flush: the state is 0.6 MB).  One GPU only: a sharded handle does not chain steps.
    Never fatal: a failure is reported in the leg instead of costing the line."""
    try:
        pkg = h.pkg
        n = LATENCY_N
        c = h.compute(n)
        c.Upload(make_particles(pkg, n))
        plan = c.Plan()
        ms, step_ms, launches, _ = h.timed_steps(c, n, steps, 50, flush=False, batch=batch)
        c.close()
        out = {"workload": f"allpairs uniform sphere N={n} seed={SEED}", "n": n, "steps": steps, "batch": batch,
               "us_per_step": ms * 1e3, "plan": plan, "gpu_launches": launches}
        if step_ms.size:
            out["kernel_us_per_step_median"] = float(np.median(step_ms)) * 1e3
        out["g_interactions_per_s"] = float(n) * float(n) / (ms * 1e-3) / 1e9
        return out
    except Exception as e:   # noqa: BLE001  (a latency leg must not take the headline down with it)
        return {"error": f"{type(e).__name__}: {e}"}


def well_leg(h: Harness, n: int, steps: int = 20, warmup: int = 5) -> dict:
    """The kernel the reference dispatches (CSMain: gravity well + integration, nBodyGravityCS.hlsl:86-109) as
    an HBM-bound kernel: 80 B per body (32 B PosVelo in, 32 B out, 16 B packed mirror).  Per-launch times are
    the library's own cudaEvent pairs around each kernel; a 256 MiB memset between steps empties the L2."""
    pkg, torch = h.pkg, h.torch
    c = pkg.Compute(n, h.local_rank)
    c.InitializeParticles(seed=1)                    # the reference's own initial conditions, on the device
    c.SetForceMode(pkg.FORCE_WELL)
    stream = torch.cuda.ExternalStream(c.GetSharedHandles().compute_stream, device=h.dev)
    for k in range(warmup + steps):
        if k == warmup:
            c.WaitForGpu()
            c.StepTimes()
        c.Simulate(n, 0, DT, DAMPING)
        with torch.cuda.stream(stream):
            h.flush_buf.zero_()
    c.WaitForGpu()
    times = c.StepTimes()
    # the same launches back to back without the flush (the state, 402 MB, is 3.2x the L2): one event pair
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        c.Simulate(n, 0, DT, DAMPING)
    e1.record(stream)
    c.WaitForGpu()
    back_to_back_ms = e0.elapsed_time(e1) / steps
    launches = c.KernelLaunches()
    c.close()
    ms = float(np.mean(times))
    peaks = read_peaks()
    bytes_per_launch = 80.0 * n   # 32 B PosVelo in + 32 B out + 16 B packed mirror, per body
    achieved = bytes_per_launch / (ms * 1e-3) / 1e9
    traffic, source = ncu_traffic("well", n, 1)
    return {"kernel": "well_step_kernel (CSMain as shipped: gravity well + integration)", "n": n, "steps": steps,
            "ms_per_step": ms, "us_min": float(times.min() * 1e3), "us_median": float(np.median(times) * 1e3),
            "back_to_back_ms_no_flush": back_to_back_ms,
            "back_to_back_gbs_no_flush": bytes_per_launch / (back_to_back_ms * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                         "frac": achieved / float(peaks["hbm_gbs"]), "traffic": traffic, "traffic_source": source,
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "peak_how": f"measured copy bandwidth ({peaks['_source']})"},
            "l2": "256 MiB memset between timed launches", "gpu_launches": int(launches)}


def parity_report(pkg, particles: np.ndarray, got: np.ndarray, targets: int = 1024) -> dict:
    """cpu_baseline leg only: the GPU's one-step result on `targets` random bodies of the headline workload
    against the oracle's LITERAL flavour (global max-norm relative error, the stated metric) and against the fp64
    direct sum (north star: "with an fp64 direct sum reported alongside")."""
    orc = importlib.import_module("oracle.oracle_py")
    n = particles.shape[0]
    idx = np.sort(np.random.default_rng(99).choice(n, min(n, targets), replace=False)).astype(np.int32)
    ref = orc.step_allpairs_targets(particles, idx, dt=DT, damping=DAMPING, flavour=orc.LITERAL)
    err = orc.rel_errors(got[idx], ref)
    a64 = orc.accel_fp64(particles, targets=idx)
    a_gpu = (got["velo"][idx, :3].astype(np.float64) / DAMPING - particles["velo"][idx, :3]) / DT
    a_lit = (ref["velo"][:, :3].astype(np.float64) / DAMPING - particles["velo"][idx, :3]) / DT
    scale = np.abs(a64).max()
    return {"targets": int(idx.size), "steps": 1, "vs_oracle_literal": err,
            "accel_vs_fp64_direct_sum": {"gpu": float(np.abs(a_gpu - a64).max() / scale),
                                         "oracle_literal": float(np.abs(a_lit - a64).max() / scale)},
            "metric": "max_i |got_i - ref_i|_inf / max_i |ref_i|_inf", "tolerance_one_step": 1e-5}


def run_mapc(args) -> None:
    h = Harness(args)
    pkg, torch, dist = h.pkg, h.torch, h.dist
    rank, local_rank, world, dev = h.rank, h.local_rank, h.world, h.dev
    peaks = read_peaks()

    if args.mode == "well":
        # the reference's shipped kernel as the headline: one GPU, HBM roofline
        n = args.n or STRONG_N
        w = well_leg(h, n, max(args.steps, 10), args.warmup)
        if rank == 0:
            value = float(n) / (w["ms_per_step"] * 1e-3) / 1e9
            print(json.dumps({"metric": "bodies_per_second", "value": value, "unit": "G bodies/s", "n_gpus": 1,
                              "steps": w["steps"], "warmup": args.warmup, "ms_per_step": w["ms_per_step"],
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic", "config": {"workload": f"well (CSMain as shipped) two shells N={n}", "l2": w["l2"]},
                              "roofline": w["roofline"], "gpu_launches": w["gpu_launches"]}), flush=True)
        return

    if args.latency_leg_only:
        if rank == 0:
            print(json.dumps({"latency": latency_leg(h)}), flush=True)
        return

    n = workload_n(pkg, world, args.scaling, args.n)
    particles = make_particles(pkg, n)
    c = h.compute(n)
    c.Upload(particles)
    if args.exchange == "peer-single":
        os.environ["MAPC_PEER_SINGLE"] = "1"
    if world > 1 and args.exchange.startswith("peer"):
        pkg.dist.enable_peer_exchange(c, dev)

    # ---- device-resident timing: W warm-up steps, then exactly K steps ------------------------
    flush = not (args.no_l2_flush or args.batch > 1)
    ms_per_step, step_ms, launches, clocks = h.timed_steps(c, n, args.steps, args.warmup, flush, args.batch,
                                                           sample_clocks=True)
    gather_ms, gather_tail_ms = c.ExchangeTimes() if world > 1 else (0.0, 0.0)
    kernel_ms = h.max_over_ranks(float(np.mean(step_ms))) if step_ms.size else float("nan")
    interactions = float(n) * float(n)
    value = interactions / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers -------------------------------------------
    # every rank holds the full host array (the caller's buffer) but the library reads only the rank's shard of
    # it: H2D per rank = n_local x 32 B, D2H per rank = n_local x 32 B; positions of the other shards arrive by
    # the device all-gather inside Upload
    host_in = torch.empty((n, 8), dtype=torch.float32, pin_memory=True)
    host_in.numpy()[:] = particles.view(np.float32).reshape(n, 8)
    host_out = torch.empty((c.num_local, 8), dtype=torch.float32, pin_memory=True)
    out_view = host_out.numpy().view(pkg.POSVELO_DTYPE).reshape(-1)
    e2e_steps = max(1, min(args.steps, 10))
    peer = world > 1 and args.exchange.startswith("peer")

    for _ in range(2):
        if peer:
            dist.barrier()                           # a peer may still be reading this rank's buffers
        c.Upload(host_in.numpy()); c.Simulate(n, 0, DT, DAMPING); c.Download(out=out_view)
    first_step = out_view.copy()                     # one step from the initial state: checked by the parity leg
    h.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if peer:
            dist.barrier()
        c.Upload(host_in.numpy())                    # H2D of this rank's shard (pinned host memory)
        c.Simulate(n, 0, DT, DAMPING)
        c.Download(out=out_view)                     # D2H of this rank's shard; blocks
    h.barrier()
    e2e_s = h.max_over_ranks(time.perf_counter() - t0) / e2e_steps
    e2e_value = interactions / e2e_s / 1e9
    checksum = float(np.abs(out_view["pos"][:, :3]).sum())

    # ---- roofline of the dominant kernel (force_cells_kernel) -------------------------------------
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    peak_tflops = sms * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    per_rank_interactions = interactions / world
    # one fused kernel IS the step: its launch duration is the CUDA-event bracket per step (which also
    # contains the L2-flush memsets, so the figure is conservative); the kernel's own %globaltimer
    # stamps (first block start -> last integrate) are reported beside it
    event_ms = ms_per_step
    achieved_tflops = per_rank_interactions * FLOP_PER_INTERACTION / (event_ms * 1e-3) / 1e12
    probe_packed, _ = pkg.fp32_peak_probe(local_rank, True)
    probe_scalar, _ = pkg.fp32_peak_probe(local_rank, False)
    # algorithmic HBM bytes per body: 64 B PosVelo read + write and the 16 B packed mirror.  The partials (one
    # 16 B value per target and canonical segment, written by the cells and read back by the combine) go through
    # a scratch ring sized to stay in L2 on an unsharded handle; `traffic` (ncu dram bytes) shows what reaches HBM.
    plan = c.Plan()
    segments = int(plan["segments"])
    hbm_bytes = 80.0 * c.num_local
    scratch_bytes = 32.0 * segments * c.num_local
    traffic, traffic_source = ncu_traffic("force", n, world)
    roofline = {
        "bound": "fp32_fma", "kernel": "force_cells_kernel (force + fused combine/integrate: the whole step is this one kernel)",
        "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
        "peak_how": f"{sms} SMs x 128 lanes x 2 flop x {sm_max_mhz:.0f} MHz ({peaks['_source']}), "
                    f"{FLOP_PER_INTERACTION:.0f} flop/interaction",
        "peak_probe_ffma2_tflops": probe_packed, "peak_probe_ffma_tflops": probe_scalar,
        "frac_of_probe": achieved_tflops / max(probe_packed, probe_scalar),
        "kernel_ms": event_ms, "kernel_ms_in_kernel_stamps": kernel_ms,
        "traffic": traffic, "traffic_source": traffic_source,
        "hbm": {"algorithmic_bytes_per_step": hbm_bytes, "achieved_gbs": hbm_bytes / (event_ms * 1e-3) / 1e9,
                "partials_scratch_bytes_per_step_l2": scratch_bytes,
                "traffic_over_algorithmic": (traffic / hbm_bytes) if traffic else None,
                "peak_gbs": peaks.get("hbm_gbs"), "note": "negligible: the step is FMA-pipe bound"},
    }
    c.close()
    h.barrier()

    # ---- the other legs (see the module docstring) ------------------------------------------------------------
    legs = {}
    if not args.headline_only:
        legs["determinism"] = determinism_leg(h)
        legs["strong"] = strong_leg(h, STRONG_N, "plummer_4194304", steps=args.strong_steps, warmup=3)
        if world == 8:
            legs["config4"] = strong_leg(h, CONFIG4_N, "sphere_1048576", steps=10, warmup=3)
        if world == 1:
            legs["well"] = well_leg(h, STRONG_N)
            legs["latency"] = latency_leg(h)

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            # cpu_baseline leg only: the one place this arm touches oracle/
            cpu, _, _ = cpu_baseline(particles, args.cpu_seconds)
            cpu["parity"] = parity_report(pkg, particles, first_step)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"allpairs uniform sphere N={n} seed={SEED} dt={DT} damping={DAMPING}",
                       "n": n, "n_per_gpu": c.num_local, "plan": plan, "parallelism": f"i-shard x{world}" + (f" ({args.exchange} exchange)" if world > 1 else ""),
                       "canonical_order": {"segments": segments, "chain_sources": pkg.plan_chain_sources()},
                       "batch": args.batch, "l2": "256 MiB memset between timed steps (inside the bracket)" if flush else
                             "no flush (latency run)"},
            "step_us": {"min": float(step_ms.min() * 1e3), "median": float(np.median(step_ms) * 1e3),
                        "p99": float(np.percentile(step_ms, 99) * 1e3), "samples": int(step_ms.size)} if step_ms.size else None,
            "tflops_at_20flop": value * FLOP_PER_INTERACTION / 1e3,
            "frac_fp32_peak": value * FLOP_PER_INTERACTION / 1e3 / (peak_tflops * world),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(c.num_local * 32 * world),
                    "d2h_bytes_per_step": int(c.num_local * 32 * world), "ms_per_step": e2e_s * 1e3,
                    "steps": e2e_steps, "checksum": checksum,
                    "note": "per rank: H2D of its shard's PosVelo, positions of the other shards by a device all-gather"
                            if world > 1 else "H2D of all N PosVelo, D2H of all N PosVelo"},
            "gpu_launches": int(launches), "clocks": clocks,
            "exchange": ({"kind": args.exchange, "allgather_ms": gather_ms,
                          "allgather_past_step_begin_ms": gather_tail_ms,
                          "note": "the gather of step k's positions runs under the local cells of step k+1"}
                         if world > 1 else None),
        }
        line.update(legs)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["mapc", "reference"], default="mapc")
    ap.add_argument("--bodies", "--n", dest="n", type=int, default=None,
                    help="override the number of bodies (use --bodies under torchrun: its parser rejects --n as ambiguous)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--exchange", choices=["nccl", "peer", "peer-single"], default="nccl",
                    help="multi-GPU position exchange: NCCL all-gather, direct peer-memory reads in the force kernel, "
                         "or the latter as one grid per step (experimental, MAPC_PEER_SINGLE=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=1,
                    help="issue the timed steps in batches of this many (mapc_compute_simulate_steps); latency runs")
    ap.add_argument("--no-l2-flush", action="store_true",
                    help="latency runs (N = 10,000, config 2): no 256 MiB memset between steps")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline leg")
    ap.add_argument("--mode", choices=["allpairs", "well"], default="allpairs",
                    help="well: the reference's shipped CSMain step (HBM-bound) at N = 4,194,304 as the headline")
    ap.add_argument("--headline-only", action="store_true",
                    help="skip the strong-scaling / determinism / well legs (profiling and latency runs)")
    ap.add_argument("--latency-leg-only", action="store_true", help="print only the config-2 latency leg (N = 10,000)")
    ap.add_argument("--strong-steps", type=int, default=3, help="timed steps of the N = 4,194,304 strong-scaling leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mapc(args)


if __name__ == "__main__":
    main()
