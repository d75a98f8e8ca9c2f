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


def ncu_traffic(n: int, world: int, segments: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the force kernel from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json); null when the workload differs from it."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        same = world == 1 and int(t["n"]) == n and int(t.get("segments", -1)) == segments
        return float(t["dram_bytes_per_launch"]) if same else None
    except Exception:
        return None


def ncu_capture_info():
    """What the committed capture was taken on (so a null `traffic` can be read against it)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return {k: t.get(k) for k in ("n", "segments", "dram_bytes_per_launch", "source")}
    except Exception:
        return None


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
        "config": {"workload": f"allpairs uniform sphere N={n} seed={SEED} dt={DT} damping={DAMPING}",
                   "n": n, "note": "ms_per_step extrapolated from the sample to all N targets"},
        "cpu_baseline": cpu,
        "e2e": {"value": ginter, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_mapc(args) -> None:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    pkg = importlib.import_module("multi-adapter-particles_b200")
    pkg.load()                                   # raises if libmapc.so is missing: no fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return pkg.dist.max_over_ranks(x, dev)

    n = workload_n(pkg, world, args.scaling, args.n)
    particles = make_particles(pkg, n)

    nccl_id = None
    if world > 1:
        nccl_id = pkg.dist.broadcast_bytes(pkg.nccl_unique_id() if rank == 0 else None,
                                           pkg.NCCL_UNIQUE_ID_BYTES, 0, dev)

    c = pkg.Compute(n, local_rank, rank=rank, world=world, nccl_id=nccl_id)
    c.Upload(particles)
    if args.exchange == "peer-single":
        os.environ["MAPC_PEER_SINGLE"] = "1"
    if world > 1 and args.exchange.startswith("peer"):
        pkg.dist.enable_peer_exchange(c, dev)
    sh = c.GetSharedHandles()
    stream = torch.cuda.ExternalStream(sh.compute_stream, device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def l2_flush():
        if args.no_l2_flush:
            return
        with torch.cuda.stream(stream):
            flush_buf.zero_()

    # ---- device-resident timing: W warm-up steps, then exactly K steps ------------------------
    for _ in range(args.warmup):
        c.Simulate(n, 0, DT, DAMPING)
        l2_flush()
    c.WaitForGpu()
    c.StepTimes()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = c.KernelLaunches()
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        if args.batch > 1:
            for k in range(0, args.steps, args.batch):
                c.SimulateSteps(n, min(args.batch, args.steps - k), 0, DT, DAMPING)
        else:
            for k in range(args.steps):
                c.Simulate(n, 0, DT, DAMPING)
                if k + 1 < args.steps:
                    l2_flush()
        c.Flush()
        e1.record(stream)
        c.WaitForGpu()
        barrier()
    launches = c.KernelLaunches() - launches0
    gather_ms, gather_tail_ms = c.ExchangeTimes() if world > 1 else (0.0, 0.0)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    step_ms = c.StepTimes()
    kernel_ms = max_over_ranks(float(np.mean(step_ms))) if step_ms.size else float("nan")
    ms_per_step = ms_total / args.steps
    interactions = float(n) * float(n)
    value = interactions / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers -------------------------------------------
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
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if peer:
            dist.barrier()
        c.Upload(host_in.numpy())                    # H2D of all N bodies (pinned host memory)
        c.Simulate(n, 0, DT, DAMPING)
        c.Download(out=out_view)                     # D2H of this rank's shard; blocks
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    e2e_value = interactions / e2e_s / 1e9
    checksum = float(np.abs(out_view["pos"][:, :3]).sum())

    # ---- roofline of the dominant kernel (force_segments_kernel) ----------------------------------
    peaks = read_peaks()
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
    # algorithmic HBM bytes per body: 64 B PosVelo read + write and the 16 B packed mirror.  The implementation
    # also writes one 16 B partial per (target, canonical segment) and reads it back in the combine; that
    # scratch is reported separately (it mostly lives in the 126 MB L2 and is noise next to the FMA time).
    segments = int(c.Plan()["segments"])
    hbm_bytes = 80.0 * c.num_local
    scratch_bytes = 32.0 * segments * c.num_local
    roofline = {
        "bound": "fp32_fma", "kernel": "force_cells_kernel (force + fused combine/integrate: the whole step is this one kernel)",
        "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
        "peak_how": f"{sms} SMs x 128 lanes x 2 flop x {sm_max_mhz:.0f} MHz ({peaks['_source']}), "
                    f"{FLOP_PER_INTERACTION:.0f} flop/interaction",
        "peak_probe_ffma2_tflops": probe_packed, "peak_probe_ffma_tflops": probe_scalar,
        "frac_of_probe": achieved_tflops / max(probe_packed, probe_scalar),
        "kernel_ms": event_ms, "kernel_ms_in_kernel_stamps": kernel_ms, "traffic": ncu_traffic(n, world, segments),
        "traffic_capture": ncu_capture_info(),
        "hbm": {"algorithmic_bytes_per_step": hbm_bytes, "achieved_gbs": hbm_bytes / (event_ms * 1e-3) / 1e9,
                "partials_scratch_bytes_per_step": scratch_bytes,
                "with_scratch_gbs": (hbm_bytes + scratch_bytes) / (event_ms * 1e-3) / 1e9,
                "peak_gbs": peaks.get("hbm_gbs"), "note": "negligible: the step is FMA-pipe bound"},
    }

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            # cpu_baseline leg only: the one place this arm touches oracle/
            cpu, _, _ = cpu_baseline(particles, args.cpu_seconds)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"allpairs uniform sphere N={n} seed={SEED} dt={DT} damping={DAMPING}",
                       "n": n, "n_per_gpu": c.num_local, "plan": c.Plan(), "parallelism": f"i-shard x{world}" + (f" ({args.exchange} exchange)" if world > 1 else ""),
                       "batch": args.batch, "l2": "no flush (latency run)" if (args.no_l2_flush or args.batch > 1) else
                             "256 MiB memset between timed steps (inside the bracket)"},
            "step_us": {"min": float(step_ms.min() * 1e3), "median": float(np.median(step_ms) * 1e3),
                        "p99": float(np.percentile(step_ms, 99) * 1e3), "samples": int(step_ms.size)} if step_ms.size else None,
            "tflops_at_20flop": value * FLOP_PER_INTERACTION / 1e3,
            "frac_fp32_peak": value * FLOP_PER_INTERACTION / 1e3 / (peak_tflops * world),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 32 * world),
                    "d2h_bytes_per_step": int(c.num_local * 32 * world), "ms_per_step": e2e_s * 1e3,
                    "steps": e2e_steps, "checksum": checksum},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "exchange": ({"kind": args.exchange, "allgather_ms": gather_ms,
                          "allgather_past_step_begin_ms": gather_tail_ms,
                          "note": "the gather of step k's positions runs under the local cells of step k+1"}
                         if world > 1 else None),
        }
    c.close()
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
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mapc(args)


if __name__ == "__main__":
    main()
