"""Multi-GPU parity worker, launched by tests/test_multi_gpu.py (or by hand) as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 \
        --master-port P tests/mgpu_worker.py [--n N ...]

Every rank drives one GPU through the C ABI (sharded handle, NCCL all-gather inside libmapc.so).
After `steps` steps each rank compares its shard, BITWISE, with an unsharded run of the same
problem on its own GPU: the canonical segment order must make results independent of the GPU count.
"""
import argparse
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bodies", dest="n", type=int, nargs="+", default=[8192, 10_000, 32_768, 262_144])
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--exchange", choices=["nccl", "peer", "peer-single", "both", "all"], default="both",
                    help="peer-single = the experimental one-grid peer exchange (MAPC_PEER_SINGLE=1); all = the three")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank, local_rank, world = (int(os.environ[k]) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("multi-adapter-particles_b200")
    pkg.load()
    ok = True
    modes = {"both": ["nccl", "peer"], "all": ["nccl", "peer", "peer-single"]}.get(args.exchange, [args.exchange])
    for n in args.n:
        if n % world:
            continue
        radius = 2000.0 * (n / 10_000.0) ** (1.0 / 3.0)
        p = pkg.ic.uniform_sphere(n, radius, seed=n, speed=1.0)
        with pkg.Compute(n, local_rank) as s:
            s.Upload(p)
            for _ in range(args.steps):
                s.Simulate(n, 0)
            s.WaitForGpu()
            full = s.Download()
        digest = hashlib.sha256(full.tobytes()).hexdigest()[:16]
        for mode in modes:
            first, count = pkg.dist.shard_range(n, rank, world)
            S = pkg.plan_segments(n)
            aligned = all((a // count) == ((b - 1) // count) for a, b in
                          (pkg.dist.segment_range(n, S, k) for k in range(S)) if b > a)
            os.environ.pop("MAPC_PEER_SINGLE", None)
            if mode == "peer-single":
                os.environ["MAPC_PEER_SINGLE"] = "1"
            if mode.startswith("peer") and not aligned:
                if rank == 0:
                    print(f"n={n} world={world} exchange={mode} skipped (segments straddle shards)", flush=True)
                continue
            nid = pkg.dist.broadcast_bytes(pkg.nccl_unique_id() if rank == 0 else None, pkg.NCCL_UNIQUE_ID_BYTES, 0, dev)
            with pkg.Compute(n, local_rank, rank=rank, world=world, nccl_id=nid) as c:
                c.Upload(p)
                if mode.startswith("peer"):
                    pkg.dist.enable_peer_exchange(c, dev)
                else:
                    dist.barrier()
                for _ in range(args.steps):
                    c.Simulate(n, 0)
                c.WaitForGpu()
                mine = c.Download()
                dist.barrier()          # nobody tears its buffers down while a peer may still read them
            same = mine.tobytes() == full[first:first + count].tobytes()
            flag = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"n={n} world={world} steps={args.steps} exchange={mode} bit-identical={bool(flag.item())} "
                      f"sha256[:16]={digest}", flush=True)
            ok = ok and bool(flag.item())
    # ---- sharded extras: device-side initial conditions, and a headless consumer per rank ------------------
    n = 16_384
    if n % world == 0:
        first, count = pkg.dist.shard_range(n, rank, world)
        with pkg.Compute(n, local_rank) as s:
            s.InitializeParticles(seed=5)
            ic_full = s.Download()
            frames_ref = [ic_full["pos"].copy()]
            for _ in range(4):
                s.Simulate(n, 0)
                s.WaitForGpu()
                frames_ref.append(s.Download()["pos"].copy())
        for async_mode in (False, True):
            nid = pkg.dist.broadcast_bytes(pkg.nccl_unique_id() if rank == 0 else None, pkg.NCCL_UNIQUE_ID_BYTES, 0, dev)
            with pkg.Compute(n, local_rank, rank=rank, world=world, nccl_id=nid) as c:
                c.InitializeParticles(seed=5)            # every rank generates its shard + all positions on its GPU
                same_ic = c.Download().tobytes() == ic_full[first:first + count].tobytes()
                dist.barrier()
                same_frames = True
                with pkg.Consumer(c, local_rank, async_mode=async_mode) as r:   # each rank dumps its own slice
                    for _ in range(4):
                        fence = c.GetFenceValue()
                        fence = r.Draw(n, fence, n)
                        c.Simulate(n, fence)
                        r.WaitForGpu()
                        frame, pos = r.Latest()
                        same_frames = same_frames and pos.shape[0] == count and \
                            pos.tobytes() == frames_ref[frame][first:first + count].tobytes()
                    c.WaitForGpu()
                dist.barrier()
            flag = torch.tensor([1 if (same_ic and same_frames) else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"n={n} world={world} sharded InitializeParticles + {'async' if async_mode else 'copying'} consumer "
                      f"per rank bit-identical={bool(flag.item())}", flush=True)
            ok = ok and bool(flag.item())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
