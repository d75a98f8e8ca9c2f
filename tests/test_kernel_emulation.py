"""The product's CUDA kernel SOURCE, run on the CPU (no GPU needed) and compared with the oracle bit for bit.

tests/emu/ compiles csrc/nbody_kernels.cuh -- unchanged; its few lines of PTX have host stand-ins -- with g++
over a shim in which one OS thread plays one CUDA thread and __syncthreads() is a barrier, and drives
force_cells_kernel / integrate_kernel with the StepArgs, segment lists, launch shapes (csrc/force_shapes.inc)
and two-launch local/remote split that csrc/mapc.cu uses.  With MUFU.RSQ replaced by the correctly rounded
1/sqrt, the kernel's arithmetic is exactly the oracle's MIRRORED flavour, so every output bit must agree:
this pins the kernels' indexing, staging, ragged tails, canonical segment order, arrival counters and the
fused combine + integrate -- everything except what only the hardware adds (the `-m gpu` tests cover that).
The emulation is test infrastructure: nothing in the product path includes it.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")

# (pairs per thread, threads per block) of csrc/force_shapes.inc
SHAPES = [(4, 256), (4, 128), (2, 128), (2, 64), (1, 128), (1, 64), (1, 32)]


@pytest.fixture(scope="module")
def emu():
    res = subprocess.run(["make", "-C", EMU_DIR], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lib = ctypes.CDLL(os.path.join(EMU_DIR, "libmapc_emu.so"))
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.emu_step_allpairs.restype = ci
    lib.emu_step_allpairs.argtypes = [vp, vp, vp, ci, ci, cf, cf, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp]
    lib.emu_make_plan.restype = None
    lib.emu_make_plan.argtypes = [ci, ci, ci, ci, ci, vp]
    lib.emu_local_targets.restype = ci
    lib.emu_local_targets.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ci]
    lib.emu_step_well.restype = ci
    lib.emu_step_well.argtypes = [vp, vp, vp, vp, ci, ci, cf, cf, ci, ci]
    lib.emu_steps_chained.restype = ci
    lib.emu_steps_chained.argtypes = [vp, ci, ci, cf, cf, ci, ci, ci, ci]
    lib.emu_init_particles.restype = ci
    lib.emu_init_particles.argtypes = [vp, vp, vp, vp, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint]
    return lib


CHAIN = 2048     # include/mapc.h MAPC_CHAIN_SOURCES: the product's chain length


def emu_step(lib, particles, S, shape, n_active=None, dt=0.1, damping=1.0, fuse=True, mass_in_loop=False,
             world=1, peer=False, block_order=0, stale=None, chunk=CHAIN, staging=0, ring=0):
    """-> (written side, packed mirror, info) after one emulated step.  chunk: sources per chain (2048, or 256
    so that small problems have several chains per segment); ring: slots of the scratch ring (0 = none)."""
    n = particles.shape[0]
    n_active = n if n_active is None else n_active
    inp = np.ascontiguousarray(particles)
    out = np.ascontiguousarray(stale.copy() if stale is not None else particles.copy())
    mirror = np.full((n, 4), np.nan, dtype=np.float32)
    info = np.zeros(3, dtype=np.uint64)
    rc = lib.emu_step_allpairs(inp.ctypes.data, out.ctypes.data, mirror.ctypes.data, n, n_active, dt, damping, S,
                               shape[0], shape[1], int(fuse), int(mass_in_loop), world, int(peer), block_order,
                               chunk, staging, ring, info.ctypes.data)
    assert rc == 0, f"emu_step_allpairs returned {rc}"
    return out, mirror, info


def oracle_step(oracle, particles, S, n_active=None, dt=0.1, damping=1.0, stale=None, flavour=None, chunk=CHAIN):
    out = stale.copy() if stale is not None else particles.copy()
    return oracle.step_allpairs(particles, n_active=n_active, dt=dt, damping=damping, S=S,
                                flavour=oracle.MIRRORED if flavour is None else flavour, out=out, chunk=chunk)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("n", [1, 63, 64, 65, 300, 1000])
def test_fused_kernel_equals_mirrored_oracle_bitwise(emu, oracle, mapc, shape, n):
    """Every launch shape, sizes straddling the 64-body tile and the ragged last tile (N = 1000: 15.6 tiles,
    S = 32 segments of 0 or 1 tile -- empty segments included)."""
    p = mapc.ic.uniform_sphere(n, 300.0, seed=1000 + n, speed=2.0)
    S = mapc.plan_segments(n)
    got, mirror, info = emu_step(emu, p, S, shape, dt=0.05, damping=0.995)
    ref = oracle_step(oracle, p, S, dt=0.05, damping=0.995)
    assert got.tobytes() == ref.tobytes()
    assert mirror.tobytes() == ref["pos"].tobytes()          # the packed mirror the next step's j loop reads
    assert info[0] == 1 and info[1] == 42 and info[2] == 1   # one launch; fence written; counters re-armed


@pytest.mark.parametrize("shape", [(4, 256), (2, 128), (1, 32)])
def test_longer_segments_and_stage_boundaries(emu, oracle, mapc, shape):
    """Segments longer than one 256-body stage (S = 2 over 1,100 sources: 9 and 9 tiles -> 576 + 524 bodies,
    stages of 256/64 with a partial last stage and the ragged last tile), so the prefetch / double-buffer
    path and the per-stage tile loop run several times."""
    n = 1100
    p = mapc.ic.plummer(n, 80.0, seed=5)
    for S in (1, 2, 3):
        got, mirror, _ = emu_step(emu, p, S, shape)
        ref = oracle_step(oracle, p, S)
        assert got.tobytes() == ref.tobytes(), S
        assert mirror.tobytes() == ref["pos"].tobytes()


def test_unfused_path_and_block_order_do_not_change_bits(emu, oracle, mapc):
    """MAPC_FUSE=0 (force cells + separate integrate_kernel) and the fused kernel with its blocks run in
    the opposite order (another block arrives last at every counter) give the same bits."""
    n = 700
    p = mapc.ic.uniform_sphere(n, 200.0, seed=3, speed=1.0)
    S = mapc.plan_segments(n)
    ref = oracle_step(oracle, p, S)
    for shape in ((4, 128), (1, 64)):
        fused, _, _ = emu_step(emu, p, S, shape)
        reverse, _, info_r = emu_step(emu, p, S, shape, block_order=1)
        unfused, mirror_u, info_u = emu_step(emu, p, S, shape, fuse=False)
        assert fused.tobytes() == ref.tobytes()
        assert reverse.tobytes() == ref.tobytes() and info_r[2] == 1
        assert unfused.tobytes() == ref.tobytes() and mirror_u.tobytes() == ref["pos"].tobytes()
        assert info_u[0] == 2                                # force launch + integrate launch


@pytest.mark.parametrize("n_active", [1, 64, 100, 640, 999])
def test_n_active_updates_whole_thread_groups_only(emu, oracle, mapc, n_active):
    """Simulate(n_active): sources j < n_active, targets = ceil(n_active/64) groups of 64 (Compute.cpp:1041,
    writes past N dropped); the other bodies of the written side keep their stale contents."""
    n = 1000
    p = mapc.ic.uniform_sphere(n, 250.0, seed=8, speed=1.0)
    stale = p.copy()
    stale["pos"] += 3.0
    S = mapc.plan_segments(n_active)
    got, _, _ = emu_step(emu, p, S, (2, 64), n_active=n_active, stale=stale)
    ref = oracle_step(oracle, p, S, n_active=n_active, stale=stale)
    assert got.tobytes() == ref.tobytes()
    t = oracle.num_targets(n, n_active)
    assert got[t:].tobytes() == stale[t:].tobytes()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("peer", [False, True])
def test_sharded_ranks_equal_unsharded_bitwise(emu, oracle, mapc, world, peer):
    """i-sharding as csrc/mapc.cu lays it out: every rank evaluates the cells of its own targets, local
    segments in one launch and remote segments in a second one that shares the arrival counters.  With the
    peer layout each rank's packed array is NaN outside its own shard, so a remote cell that did not go
    through seg_src[] (indexed like the launch's segment list) would poison the result."""
    n = 2048                       # 32 tiles, S = 32: one tile per segment, S/world segments per rank
    p = mapc.ic.uniform_sphere(n, 400.0, seed=11, speed=1.0)
    S = mapc.plan_segments(n)
    ref = oracle_step(oracle, p, S)
    for shape in ((2, 128), (1, 32)):
        got, mirror, info = emu_step(emu, p, S, shape, world=world, peer=peer)
        assert got.tobytes() == ref.tobytes(), (shape, world, peer)
        assert mirror.tobytes() == ref["pos"].tobytes()
        assert info[0] == 2 * world and info[2] == 1          # local + remote launch per rank


def test_mass_in_loop_variant_stays_at_rounding_level(emu, oracle, mapc):
    """MAPC_MASS_IN_LOOP=1 (the shader's per-pair mass multiply, 12 lane-ops) has no oracle flavour with its
    exact fma placement; it must sit within rounding of both flavours."""
    n = 900
    p = mapc.ic.uniform_sphere(n, 300.0, seed=21, speed=1.0)
    S = mapc.plan_segments(n)
    got, _, _ = emu_step(emu, p, S, (4, 128), mass_in_loop=True)
    for flavour in (oracle.LITERAL, oracle.MIRRORED):
        err = oracle.rel_errors(got, oracle_step(oracle, p, S, flavour=flavour))
        assert max(err.values()) < 2e-6, err


@pytest.mark.parametrize("n,n_active,shard", [(1000, 1000, (0, 1000)), (1000, 130, (0, 1000)), (1024, 1024, (512, 512))])
def test_well_and_pack_kernels(emu, oracle, mapc, n, n_active, shard):
    """well_step_kernel -- the literal CSMain the reference dispatches (nBodyGravityCS.hlsl:86-109) -- and
    pack_positions_kernel, bit for bit against the oracle's MIRRORED well step; also on the second shard of
    two, where the kernel indexes PosVelo locally and the packed mirror globally."""
    p = mapc.ic.uniform_sphere(n, 300.0, seed=31, speed=4.0)
    stale = p.copy()
    stale["velo"] += 1.0
    got = np.ascontiguousarray(stale.copy())
    mirror = np.empty((n, 4), dtype=np.float32)
    packed = np.empty((n, 4), dtype=np.float32)
    first, count = shard
    rc = emu.emu_step_well(np.ascontiguousarray(p).ctypes.data, got.ctypes.data, mirror.ctypes.data,
                           packed.ctypes.data, n, n_active, 0.05, 0.995, first, count)
    assert rc == 0
    ref = oracle.step_well(p, n_active=n_active, dt=0.05, damping=0.995, flavour=oracle.MIRRORED, out=stale.copy())
    t = oracle.num_targets(n, n_active)
    lo, hi = first, min(first + count, t)
    assert got[lo:hi].tobytes() == ref[lo:hi].tobytes()
    assert got[hi:first + count].tobytes() == stale[hi:first + count].tobytes()    # not dispatched: untouched
    assert got[:first].tobytes() == stale[:first].tobytes()                        # another rank's shard
    assert mirror[lo:hi].tobytes() == ref["pos"][lo:hi].tobytes()
    assert packed.tobytes() == p["pos"].tobytes()


@pytest.mark.parametrize("shape,n", [((1, 32), 300), ((1, 64), 1000), ((2, 64), 1100), ((4, 128), 2500)])
def test_chained_steps_equal_oracle_trajectory_bitwise(emu, oracle, mapc, shape, n):
    """Step-to-step dataflow (StepArgs::block_step / wait_prev: a cell waits for the target blocks of the previous
    step it reads instead of for the whole previous grid): four ping-pong steps whose launches 2..4 carry
    wait_prev, against the oracle's trajectory.  The emulation runs blocks one at a time, so what it pins is that
    every wait names blocks that exist and are published with the right step id (a wrong one is reported as a
    timeout), that the flags are published for every block, and that the L2 loads read the right side."""
    p = mapc.ic.uniform_sphere(n, 300.0, seed=n, speed=1.0)
    S = mapc.plan_segments(n)
    state = np.ascontiguousarray(p.copy())
    assert emu.emu_steps_chained(state.ctypes.data, n, 4, 0.1, 1.0, S, shape[0], shape[1], 0) == 0
    ref = p
    for _ in range(4):
        ref = oracle.step_allpairs(ref, S=S, flavour=oracle.MIRRORED)
    assert state.tobytes() == ref.tobytes()


@pytest.mark.parametrize("n,first,count", [(1000, 0, 1000), (1001, 0, 1001), (4096, 1024, 1024)])
def test_init_particles_kernel_equals_oracle_restatement_bitwise(emu, oracle, mapc, n, first, count):
    """init_particles_kernel (the reference's two-shell initial conditions, Compute.cpp:719-749 / :820-844,
    generated on the device) against the oracle's restatement: same bytes, on a whole handle, an odd N (the
    last body stays zero) and the second shard of four (PosVelo local, packed positions for all N)."""
    ref = oracle.init_particles(n, 1234)
    a = np.zeros(count, dtype=mapc.POSVELO_DTYPE)
    b = np.zeros(count, dtype=mapc.POSVELO_DTYPE)
    pa = np.full((n, 4), np.nan, dtype=np.float32)
    pb = np.full((n, 4), np.nan, dtype=np.float32)
    assert emu.emu_init_particles(a.ctypes.data, b.ctypes.data, pa.ctypes.data, pb.ctypes.data, n, first, count, 1234) == 0
    assert a.tobytes() == ref[first:first + count].tobytes() and b.tobytes() == a.tobytes()
    assert pa.tobytes() == ref["pos"].tobytes() and pb.tobytes() == pa.tobytes()
    # the distribution the reference describes: two shells of radius 400 around x = +-300, speed <= 15, tangential
    half = n // 2
    for sl, cx in ((slice(0, half), 300.0), (slice(half, 2 * half), -300.0)):
        d = ref["pos"][sl, :3] - np.array([cx, 0, 0], dtype=np.float32)
        np.testing.assert_allclose(np.linalg.norm(d, axis=1), 400.0, rtol=1e-5)
    speed = np.linalg.norm(ref["velo"][:, :3], axis=1)
    assert speed.max() <= 15.0 * (1 + 1e-5) and speed.mean() > 5.0
    assert oracle.init_particles(n, 1235).tobytes() != ref.tobytes()


@pytest.mark.parametrize("shape", SHAPES)
def test_chunked_order_equals_chunked_oracle_bitwise(emu, oracle, mapc, shape):
    """The chains of the canonical order (kernel template parameter CHAIN; oracle `chunk`): chain sums folded
    left to right into the segment's partial.  N = 1,100 with 256-source chains: S = 1 gives one segment of 4
    full chains + a ragged one, S = 2 segments of 576 and 524 (2 chains + remainder), S = 32 segments shorter
    than a chain (then the chain length does not matter and the result equals the 2,048-source order)."""
    n = 1100
    p = mapc.ic.plummer(n, 80.0, seed=5)
    for S in (1, 2, 32):
        got, mirror, info = emu_step(emu, p, S, shape, chunk=256)
        ref = oracle_step(oracle, p, S, chunk=256)
        assert got.tobytes() == ref.tobytes(), S
        assert mirror.tobytes() == ref["pos"].tobytes() and info[2] == 1
        plain = oracle_step(oracle, p, S)
        assert (got.tobytes() == plain.tobytes()) == (S == 32)
    # a segment that is an exact multiple of the chunk: the last chunk is folded by the final store
    q = mapc.ic.uniform_sphere(1024, 200.0, seed=6, speed=1.0)
    got, _, _ = emu_step(emu, q, 2, shape, chunk=256)
    assert got.tobytes() == oracle_step(oracle, q, 2, chunk=256).tobytes()


def test_chunked_order_product_chunk_size_and_shards(emu, oracle, mapc):
    """The chain length the library uses (2,048 sources), on segments of 2,560 sources (N = 5,120, S = 2), and
    the same through two emulated ranks (local + remote launches write the same partials)."""
    n = 5120
    p = mapc.ic.uniform_sphere(n, 600.0, seed=13, speed=1.0)
    ref = oracle_step(oracle, p, 2)
    assert ref.tobytes() != oracle_step(oracle, p, 2, chunk=0).tobytes()     # the chains really change the bits
    got, _, _ = emu_step(emu, p, 2, (4, 256))
    assert got.tobytes() == ref.tobytes()
    got2, mirror2, info2 = emu_step(emu, p, 2, (2, 128), world=2)
    assert got2.tobytes() == ref.tobytes() and mirror2.tobytes() == ref["pos"].tobytes() and info2[2] == 1


@pytest.mark.parametrize("shape,n", [((1, 32), 700), ((2, 64), 1100), ((4, 128), 5120)])
def test_scratch_ring_does_not_change_bits(emu, oracle, mapc, shape, n):
    """The L2-resident scratch ring of unsharded fused steps (csrc/mapc.cu plan_scratch): `ring` slots reused
    round-robin by the target blocks, cells handed out by the ticket counter in (target block, segment) order,
    a slot released by the combine of the block that used it before.  Whatever the ring size and whichever
    order the blocks start in, the bits are those of one slot per target block; ticket, slot generations and
    counters are back at zero for the next step."""
    p = mapc.ic.uniform_sphere(n, 300.0, seed=n, speed=1.0)
    S = mapc.plan_segments(n)
    ref = oracle_step(oracle, p, S)
    # rings of 4 and more slots hand their tickets out in groups of ring/2 target blocks, segment major inside a group
    # (the cells that run side by side read the same source segment); the last group may be short
    for ring in (1, 2, 3, 4, 5, 6, 8):
        for order in (0, 1):
            got, mirror, info = emu_step(emu, p, S, shape, ring=ring, block_order=order)
            assert got.tobytes() == ref.tobytes(), (ring, order)
            assert mirror.tobytes() == ref["pos"].tobytes() and info[2] == 1 and info[1] == 42


def test_bounded_chains_cut_the_rounding_noise(oracle, mapc):
    """What the chunked order is for: with one chain per segment the noise of the
    close-pair targets grows with the segment length, with 2,048-source chunks it does not.  N = 262,144,
    S = 8 (32,768-source segments, the rule the round started with): 1.07e-5 between the CPU flavours without
    chunks, a few 1e-6 with them."""
    from scipy.spatial import cKDTree
    p = mapc.ic.workload("sphere_262144")
    xyz = p["pos"][:, :3].astype(np.float64)
    dist, _ = cKDTree(xyz).query(xyz, k=2)
    close = np.sort(np.argsort(dist[:, 1])[:512]).astype(np.int32)

    def envelope(chunk):
        lit = oracle.step_allpairs_targets(p, close, S=8, flavour=oracle.LITERAL, chunk=chunk)
        mir = oracle.step_allpairs_targets(p, close, S=8, flavour=oracle.MIRRORED, chunk=chunk)
        return max(oracle.rel_errors(mir, lit).values())

    plain, chunked = envelope(0), envelope(2048)
    assert plain > 8e-6 and chunked < 0.4 * plain, (plain, chunked)


GOLDEN = os.path.join(HERE, "golden")


def test_emulated_kernels_against_reference_shader_vectors(emu, oracle, mapc):
    """The chain reference shader -> committed vectors -> kernel source, closed on the CPU: the emulated
    force kernel against outputs of the reference's own nBodyGravityCS.hlsl code compiled for the CPU
    (tests/golden/ref_shader_vectors.npz: all-pairs through its bodyBodyInteraction, and the shipped CSMain),
    at the tolerances the GPU tests use (the kernel's fma placement and mass-per-partial differ from the
    shader's literal order at rounding level; MUFU.RSQ is the GPU tests' business)."""
    g = np.load(os.path.join(GOLDEN, "ref_shader_vectors.npz"))
    for case, dt, damping in (("a", 0.1, 1.0), ("b", 0.05, 0.995)):
        inp = g[f"allpairs_{case}_in"].view(mapc.POSVELO_DTYPE).reshape(-1)
        S = int(g[f"allpairs_{case}_S"])
        assert S == mapc.plan_segments(inp.shape[0])
        got, _, _ = emu_step(emu, inp, S, (2, 128), dt=dt, damping=damping)
        err = oracle.rel_errors(got, g[f"allpairs_{case}_out"])
        assert max(err.values()) <= 1e-5, (case, err)
    w = np.ascontiguousarray(g["well_in"].view(mapc.POSVELO_DTYPE).reshape(-1))
    n = w.shape[0]
    for ref_key, dt, damping in (("well_out_a", 0.1, 1.0), ("well_out_b", 0.05, 0.995)):
        got = w.copy()
        mirror = np.empty((n, 4), dtype=np.float32)
        packed = np.empty((n, 4), dtype=np.float32)
        assert emu.emu_step_well(w.ctypes.data, got.ctypes.data, mirror.ctypes.data, packed.ctypes.data, n, n,
                                 dt, damping, 0, n) == 0
        err = oracle.rel_errors(got, g[ref_key])
        assert max(err.values()) <= 2e-6, (ref_key, err)


@pytest.mark.parametrize("name", ["lattice_1000", "plummer_777"])
def test_emulated_kernel_trajectory_against_golden(emu, oracle, mapc, name):
    """Ten ping-pong steps of the emulated fused kernel against the committed LITERAL trajectory (1e-5 after
    one step, 1e-4 after the last), and bit for bit against the MIRRORED oracle stepped alongside."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    state = g["input"].view(mapc.POSVELO_DTYPE).reshape(-1)
    dt, damping, steps, S = float(g["dt"]), float(g["damping"]), int(g["steps"]), int(g["S"])
    mirrored = state
    for step in range(1, steps + 1):
        state, _, _ = emu_step(emu, state, S, (1, 64) if step % 2 else (4, 128), dt=dt, damping=damping)
        mirrored = oracle.step_allpairs(mirrored, dt=dt, damping=damping, S=S, flavour=oracle.MIRRORED)
        assert state.tobytes() == mirrored.tobytes(), step
        if step == 1:
            assert max(oracle.rel_errors(state, g["literal_1"]).values()) <= 1e-5
    assert max(oracle.rel_errors(state, g["literal_last"]).values()) <= 1e-4


def test_launch_plan_and_dispatch_granularity(emu, oracle, mapc):
    """csrc/step_layout.hpp (shared by csrc/mapc.cu and the emulation): the launch shape chosen for the
    BASELINE sizes, forced shapes, and the number of targets a Simulate(n_active) updates on a shard."""
    def plan(n_targets, S, force=(0, 0), sms=148):
        out = (ctypes.c_int * 4)()
        emu.emu_make_plan(n_targets, S, sms, force[0], force[1], out)
        return tuple(out)

    # config 3: 262,144 targets, S = 32 -> (2, 128), 512 target blocks x 32 segments = 16,384 cells
    assert plan(262_144, 32) == (2, 128, 512, 32)
    # an 8-GPU shard of config 5 (4,194,304 / 8 targets): 1,024 target blocks of (2, 128)
    assert plan(524_288, 32) == (2, 128, 1024, 32)
    # config 2: 10,000 targets fill the block slots only with a small shape -> (1, 128), 40 target blocks
    assert plan(10_000, 32) == (1, 128, 40, 32)
    # a size whose cells would leave most block slots of a big shape empty never gets that shape
    assert plan(4_096, 32)[:2] in ((1, 128), (1, 64))
    # every shape can be forced, and covers all targets
    for pairs, threads in SHAPES:
        pl = plan(10_000, 32, (pairs, threads))
        assert pl[:2] == (pairs, threads) and pl[2] * pairs * 2 * threads >= 10_000 > (pl[2] - 1) * pairs * 2 * threads
    # a single target still gets a block
    assert plan(1, 32)[2] == 1
    # occupancy throttle of chained small-N steps: about three quarters of a step's cells resident, never more blocks
    # per SM than the kernel is compiled for (0 = no throttle), the measured optimum at the sizes that were swept
    # (profiles/r02_small_n_occupancy.txt)
    emu.emu_throttle_blocks_per_sm.restype = ctypes.c_int
    emu.emu_throttle_blocks_per_sm.argtypes = [ctypes.c_int] * 5
    throttle = {n: emu.emu_throttle_blocks_per_sm(n, 32, 148, 1, 128) for n in (1_000, 2_500, 4_096, 6_000, 8_192, 10_000, 12_000, 16_384)}
    assert throttle == {1_000: 1, 2_500: 2, 4_096: 3, 6_000: 4, 8_192: 5, 10_000: 6, 12_000: 0, 16_384: 0}
    # dispatch granularity (Compute.cpp:1041) agrees with the oracle's statement of it
    for n, n_active in ((1000, 0), (1000, 1), (1000, 64), (1000, 65), (1000, 999), (1000, 1000), (10_000, 9_999)):
        assert emu.emu_local_targets(n, 0, n, n_active) == oracle.num_targets(n, n_active)
    # on the second of two shards of 512: targets past the shard start only
    assert emu.emu_local_targets(1024, 512, 512, 100) == 0
    assert emu.emu_local_targets(1024, 512, 512, 600) == 128
    assert emu.emu_local_targets(1024, 512, 512, 1024) == 512


TMA, SHFL = 1, 2


@pytest.mark.parametrize("shape", [(4, 256), (4, 128), (2, 128)])
def test_tma_staging_variant_bitwise(emu, oracle, mapc, shape):
    """MAPC_TMA=1: the stages filled by 1-D bulk copies + mbarrier phases instead of LDG/STS (the shapes with
    256-body stages).  Emulated as memcpy + phase counter: checks the copy's source offsets, byte counts
    (a non-zero multiple of 16, 16-byte aligned -- what cp.async.bulk accepts) and the double-buffer parity."""
    n = 1100
    p = mapc.ic.plummer(n, 80.0, seed=5)
    for S in (1, 2, 32):
        got, mirror, info = emu_step(emu, p, S, shape, staging=TMA)
        ref = oracle_step(oracle, p, S)
        assert got.tobytes() == ref.tobytes(), S
        assert mirror.tobytes() == ref["pos"].tobytes() and info[2] == 1


@pytest.mark.parametrize("shape", [(2, 128), (1, 64), (1, 32)])
def test_shuffle_broadcast_variant_bitwise(emu, oracle, mapc, shape):
    """MAPC_SHFL=1: each lane reads one body of a 32-body group and the group is handed round with warp
    shuffles (emulated with per-warp barriers).  Same bodies in the same order: same bits."""
    n = 200
    p = mapc.ic.uniform_sphere(n, 100.0, seed=17, speed=1.0)
    for S in (1, 32):
        got, _, _ = emu_step(emu, p, S, shape, staging=SHFL)
        assert got.tobytes() == oracle_step(oracle, p, S).tobytes(), S


def test_fuzz_emulated_kernel_against_oracle(emu, oracle, mapc):
    """Property-based sweep (hypothesis): any N, any segment count, any launch shape, any n_active, fused or
    not, with or without bounded chains -- the emulated kernel equals the oracle's MIRRORED flavour bit for bit
    and leaves the bodies it does not dispatch untouched."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(n=st.integers(1, 700), S=st.integers(1, 32), shape=st.sampled_from(SHAPES),
           frac=st.floats(0.0, 1.0), fuse=st.booleans(), chunk=st.sampled_from([CHAIN, CHAIN, 256]),
           ring=st.sampled_from([0, 0, 1, 2, 4, 5, 8]), seed=st.integers(0, 1000))
    def check(n, S, shape, frac, fuse, chunk, ring, seed):
        n_active = max(1, min(n, int(round(frac * n)))) if frac < 0.9 else n
        p = mapc.ic.uniform_sphere(n, 150.0, seed=seed, speed=1.0)
        stale = p.copy()
        stale["velo"] -= 2.0
        got, _, info = emu_step(emu, p, S, shape, n_active=n_active, fuse=fuse, stale=stale, chunk=chunk,
                                ring=ring, dt=0.07, damping=0.99)
        ref = oracle_step(oracle, p, S, n_active=n_active, stale=stale, chunk=chunk, dt=0.07, damping=0.99)
        assert got.tobytes() == ref.tobytes(), (n, S, shape, n_active, fuse, chunk, ring, seed)
        assert info[2] == 1
    check()


def test_fuzz_sharded_layouts_against_oracle(emu, oracle, mapc):
    """Property-based sweep over emulated ranks: any world size, shard size, segment count and launch shape,
    NCCL layout or (where no segment straddles two shards) peer layout: the shards put together equal the
    unsharded oracle step bit for bit."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(world=st.sampled_from([2, 3, 4, 8]), k=st.integers(1, 5), S=st.integers(1, 32),
           shape=st.sampled_from(SHAPES), peer=st.booleans(), seed=st.integers(0, 1000))
    def check(world, k, S, shape, peer, seed):
        n = 64 * world * k
        p = mapc.ic.uniform_sphere(n, 150.0, seed=seed, speed=1.0)
        inp = np.ascontiguousarray(p)
        out = inp.copy()
        mirror = np.full((n, 4), np.nan, dtype=np.float32)
        info = np.zeros(3, dtype=np.uint64)
        rc = emu.emu_step_allpairs(inp.ctypes.data, out.ctypes.data, mirror.ctypes.data, n, n, 0.1, 1.0, S,
                                   shape[0], shape[1], 1, 0, world, int(peer), 0, CHAIN, 0, 0, info.ctypes.data)
        if peer and rc == -2:
            # segments straddle shards: the library refuses the peer exchange for such a step
            count = n // world
            straddles = any(a // count != (b - 1) // count for a, b in
                            (oracle.segment_range(n, S, s) for s in range(S)) if b > a)
            assert straddles
            return
        assert rc == 0
        ref = oracle_step(oracle, p, S)
        assert out.tobytes() == ref.tobytes(), (world, n, S, shape, peer, seed)
        assert mirror.tobytes() == ref["pos"].tobytes() and info[2] == 1
    check()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_single_grid_peer_exchange_bitwise(emu, oracle, mapc, world):
    """MAPC_PEER_SINGLE=1 (experimental): the peer exchange as ONE grid whose segment list holds the local
    segments first and the remote ones after them, every entry with its own source array and step flag (none
    for the local ones).  Same cells and partials, so the same bits as two launches -- and as one GPU."""
    n = 2048
    p = mapc.ic.uniform_sphere(n, 400.0, seed=11, speed=1.0)
    S = mapc.plan_segments(n)
    ref = oracle_step(oracle, p, S)
    for shape in ((4, 128), (1, 64)):
        got, mirror, info = emu_step(emu, p, S, shape, world=world, peer=2)
        assert got.tobytes() == ref.tobytes(), (shape, world)
        assert mirror.tobytes() == ref["pos"].tobytes()
        assert info[0] == world and info[2] == 1            # one launch per rank
