"""The N>1 host path on CPU: world_size 2 over gloo (127.0.0.1).

Covers the rendezvous helpers bench.py uses (unique-id broadcast, max-over-ranks) and the
decomposition itself: every rank advances its own target shard with the oracle, all-gathers the
packed positions, and after 3 steps the assembled state must equal the single-process oracle bit for
bit -- the property that makes 1/2/4/8-GPU runs identical (canonical segment order)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import REPO_ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, steps, out_dir):
    import importlib
    sys.path.insert(0, REPO_ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("multi-adapter-particles_b200")
    orc = importlib.import_module("oracle.oracle_py")

    ident = pkg.dist.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128, 0)
    assert ident == bytes(range(128))
    assert pkg.dist.max_over_ranks(float(rank + 1)) == float(world)

    state = pkg.ic.uniform_sphere(n, 900.0, seed=17, speed=1.0)
    first, count = pkg.dist.shard_range(n, rank, world)
    S = orc.default_segments(n)
    for _ in range(steps):
        mine = orc.step_allpairs_targets(state, np.arange(first, first + count, dtype=np.int32), S=S)
        # the exchange step: all-gather of every rank's slice (positions AND, here, velocities)
        gathered = [torch.zeros(count * 8, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine.view(np.float32).reshape(-1).copy()))
        state = np.concatenate([g.numpy() for g in gathered]).view(orc.POSVELO_DTYPE).reshape(-1)
    if rank == 0:
        np.save(os.path.join(out_dir, "sharded.npy"), state.view(np.float32).reshape(-1, 8))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_step_is_bit_identical(tmp_path, mapc, oracle):
    import torch.multiprocessing as mp
    n, steps, world = 1536, 3, 2
    mp.spawn(_worker, args=(world, _free_port(), n, steps, str(tmp_path)), nprocs=world, join=True)
    sharded = np.load(tmp_path / "sharded.npy")
    ref = mapc.ic.uniform_sphere(n, 900.0, seed=17, speed=1.0)
    for _ in range(steps):
        ref = oracle.step_allpairs(ref)
    assert sharded.tobytes() == ref.view(np.float32).tobytes()


def test_weak_scaling_sizes_and_shards(mapc):
    d = mapc.dist
    assert d.weak_scaled_n(1) == 262_144
    for world in (2, 4, 8):
        n = d.weak_scaled_n(world)
        assert n % (64 * 8 * world) == 0
        assert abs(n * n / world - 262_144 ** 2) / 262_144 ** 2 < 0.01   # same pairs per GPU
        first, count = d.shard_range(n, world - 1, world)
        assert first + count == n
    with pytest.raises(ValueError):
        d.shard_range(1001, 0, 2)


def test_local_segment_classification(mapc):
    d = mapc.dist
    # N = 1,048,576 on 8 ranks with 8 segments: segment r is exactly rank r's shard; with the canonical
    # 16 segments of that size every rank owns two
    n = 1_048_576
    for r in range(8):
        first, count = d.shard_range(n, r, 8)
        assert d.local_segments(n, 8, first, count) == [r]
        assert d.local_segments(n, 16, first, count) == [2 * r, 2 * r + 1]
    # 2 ranks: 4 local segments each
    assert d.local_segments(n, 8, 0, n // 2) == [0, 1, 2, 3]
    # ragged: N = 10,000 (157 tiles, S = 32) on 2 ranks -- a segment straddling the shard edge is remote
    loc0 = d.local_segments(10_000, 32, 0, 5000)
    loc1 = d.local_segments(10_000, 32, 5000, 5000)
    assert set(loc0).isdisjoint(loc1) and len(loc0) + len(loc1) == 31
    for s in range(32):
        a, b = d.segment_range(10_000, 32, s)
        assert (a, b) == tuple(__import__("oracle.oracle_py", fromlist=["x"]).segment_range(10_000, 32, s))


def test_canonical_segments_partition_the_sources(mapc, oracle):
    """Properties of the canonical segmentation that every size must satisfy: the S ranges tile [0, n)
    exactly, in order, on 64-body boundaries (only the last may be ragged), the three statements of the rule
    (library, oracle, host helper) agree, and S is the frozen 32 for every size (chains are bounded separately,
    at 2,048 sources inside a segment)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None, derandomize=True)
    @given(st.integers(min_value=1, max_value=5_000_000))
    def check(n):
        S = mapc.plan_segments(n)
        assert S == oracle.default_segments(n) == 32
        edge, longest = 0, 0
        for s in range(S):
            a, b = mapc.dist.segment_range(n, S, s)
            assert (a, b) == oracle.segment_range(n, S, s)
            assert a == edge and b >= a and (b % 64 == 0 or b == n)
            edge, longest = b, max(longest, b - a)
        assert edge == n
        assert longest <= (n + 63) // 64 // S * 64 + 64     # segments differ by at most one tile
    check()


def test_no_canonical_segment_straddles_a_shard(mapc):
    """The bench's multi-GPU sizes (weak-scaled and BASELINE configs 4 and 5): with S a multiple of the world
    size every segment lies inside one rank's shard, so the peer exchange never needs a gathered copy and the
    local/remote split of the NCCL path is exact."""
    d = mapc.dist
    sizes = [d.weak_scaled_n(w) for w in (2, 4, 8)] + [1_048_576, 4_194_304]
    for n in sizes:
        S = mapc.plan_segments(n)
        for world in (2, 4, 8):
            if n % (64 * world):
                continue
            count = n // world
            owned = []
            for r in range(world):
                loc = d.local_segments(n, S, r * count, count)
                assert len(loc) == S // world, (n, world, r, loc)
                owned += loc
            assert sorted(owned) == list(range(S))
