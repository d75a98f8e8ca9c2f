"""Host-side plumbing for the one-process-per-GPU form (torch.distributed; NCCL on GPUs, gloo on CPU).

Only rendezvous-style helpers live here: the data path's single exchange step (the all-gather of
packed positions) is issued by libmapc.so itself on its own NCCL communicator.
"""
from __future__ import annotations

import math

BASE_N = 262_144          # BASELINE.json config 3 / Particles/defines.h:44 MIN_NUM_PARTICLES
TILE = 64                 # Particles/defines.h:37


def shard_range(n: int, rank: int, world: int):
    """Targets owned by `rank`: the contiguous slice [first, first + count)."""
    if n % world:
        raise ValueError(f"N={n} is not divisible by world={world}")
    count = n // world
    return rank * count, count


def weak_scaled_n(world: int, base: int = BASE_N) -> int:
    """N that keeps the per-GPU work (N^2 / world pairs) of the base workload: base * sqrt(world),
    rounded to a multiple of 64 * 8 * world: shards are then whole tiles, and because the canonical segment
    count (32/64/128) is a multiple of every supported world size no segment straddles two shards."""
    if world == 1:
        return base
    q = TILE * 8 * world
    return int(round(base * math.sqrt(world) / q)) * q


def segment_range(n_sources: int, segments: int, s: int):
    tiles = (n_sources + TILE - 1) // TILE
    a = min(n_sources, (tiles * s) // segments * TILE)
    b = min(n_sources, (tiles * (s + 1)) // segments * TILE)
    return a, b


def local_segments(n_sources: int, segments: int, first: int, count: int):
    """Canonical segments whose sources all live in [first, first+count): a rank can evaluate them
    before the all-gather of the step lands (the overlap of SURVEY.md section 8e)."""
    out = []
    for s in range(segments):
        a, b = segment_range(n_sources, segments, s)
        if a >= first and b <= first + count:
            out.append(s)
    return out


def broadcast_bytes(data: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Broadcast a small byte string (the NCCL unique id) from `src` over the default group."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_gather_bytes(data: bytes, device=None):
    """Every rank contributes `data` (same length everywhere); returns the list in rank order."""
    import torch
    import torch.distributed as dist
    mine = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device) if device is not None else \
        torch.frombuffer(bytearray(data), dtype=torch.uint8).clone()
    outs = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, mine)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def enable_peer_exchange(compute, device=None) -> None:
    """Switch a sharded Compute to the collective-free exchange: export, all-gather the blobs, attach,
    and barrier so nobody starts stepping before every rank has uploaded and attached."""
    import torch.distributed as dist
    compute.IpcAttach(all_gather_bytes(compute.IpcExport(), device))
    dist.barrier()
