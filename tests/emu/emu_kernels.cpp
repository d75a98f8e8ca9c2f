// tests/emu/emu_kernels.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles csrc/nbody_kernels.cuh -- the product's
// kernel source, unchanged -- for the host over tests/emu/cuda_emu.hpp and drives it the way
// csrc/mapc.cu's enqueue_one does: same StepArgs, same segment lists (local cells first, remote cells in a
// second launch, shared arrival counters), same shape table (csrc/force_shapes.inc), same unfused
// integrate_kernel alternative, optionally `world` emulated ranks with the NCCL-style layout (every rank
// sees all N packed positions) or the peer layout (a rank's packed array is valid only inside its own
// shard, everything else is poisoned with NaN, and remote cells must go through seg_src[]).
// The Python side (tests/test_kernel_emulation.py) compares the result bit for bit with the oracle's
// MIRRORED flavour.  Built by tests/emu/Makefile with g++ -ffp-contract=off; never linked into libmapc.so.
#define MAPC_HOST_EMULATION 1
#include "../../multi-adapter-particles_b200/csrc/nbody_kernels.cuh"
#include "../../multi-adapter-particles_b200/csrc/step_layout.hpp"

#include <sched.h>

#include <cstring>
#include <limits>
#include <vector>

namespace cuda_emu {
thread_local uint3 t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local Block *t_block = nullptr;
}  // namespace cuda_emu

namespace {

using mapc::StepArgs;

template <int P, int T, int TJ, int U, int MINB, int ORDER>
bool launch_shape(bool fuse, bool peer, bool inloop, int chain, int staging, bool has_tma, const StepArgs &a, int order)
{
    // one cell per block, segment index fastest (csrc/mapc.cu launch_force)
    const dim3 grid((unsigned)a.n_iblocks * (unsigned)a.segs.count, 1, 1), block(T, 1, 1);
    if (a.segs.count == 0 || a.i_cnt <= 0) return true;
    if (chain == 256) {
        // a short chain so that small problems have several chains per segment: fused / unfused default staging
        if (peer || inloop || staging != 0) return false;
        if (fuse) cuda_emu::launch(grid, block, order, [&] { mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, true, false, false, false, false, 256>(a); });
        else cuda_emu::launch(grid, block, order, [&] { mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, false, false, false, false, false, 256>(a); });
        return true;
    }
    if (chain != MAPC_CHAIN_SOURCES) return false;
    if (staging != 0) {
        // MAPC_TMA=1 / MAPC_SHFL=1: like csrc/mapc.cu, for the fused, non-peer, mass-per-partial kernel only;
        // TMA staging exists for the 256-body-stage shapes
        if (!fuse || peer || inloop) return false;
        if (staging == 1) {
            if (!has_tma) return false;
            cuda_emu::launch(grid, block, order, [&] { mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, true, false, true>(a); });
        } else {
            cuda_emu::launch(grid, block, order, [&] { mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, true, false, false, false, true>(a); });
        }
        return true;
    }
    // the instantiations csrc/mapc.cu launches
    if (fuse && peer) cuda_emu::launch(grid, block, order, [&] { mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, true, true>(a); });
    else if (fuse && inloop) cuda_emu::launch(grid, block, order, [&] { mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, true, false, false, true>(a); });
    else if (fuse) cuda_emu::launch(grid, block, order, [&] { mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, true, false>(a); });
    else cuda_emu::launch(grid, block, order, [&] { mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, false, false>(a); });
    return true;
}

bool launch_force(int pairs, int threads, bool fuse, bool peer, bool inloop, int chain, int staging, const StepArgs &a,
                  int order)
{
#define MAPC_SHAPE(P, T, TJ, U, MINB, ORDER, HAS_TMA) \
    if (pairs == P && threads == T)                   \
        return launch_shape<P, T, TJ, U, MINB, ORDER>(fuse, peer, inloop, chain, staging, HAS_TMA, a, order);
#include "../../multi-adapter-particles_b200/csrc/force_shapes.inc"
#undef MAPC_SHAPE
    return false;
}

struct alignas(16) PV { float pos[4]; float velo[4]; };
static_assert(sizeof(PV) == sizeof(mapc_posvelo), "PosVelo layout");

}  // namespace

extern "C" {

// One all-pairs step.  in/out: n PosVelo each (out must hold the previous contents of the written side: bodies
// that are not targets keep them).  pos_next_out: n float4, the packed mirror of the written side as rank 0
// .. world-1 wrote it (own shards only).  info[0] = kernel launches, info[1] = fence word after the step,
// info[2] = 1 if every arrival counter and the `done` counter were back at zero.
// peer: 0 NCCL layout, 1 peer layout (local + remote launch), 2 peer layout as one grid (MAPC_PEER_SINGLE=1).
// chain: sources per sequential accumulation chain: MAPC_CHAIN_SOURCES (the product) or 256 (so that small
// problems have several chains per segment).
// ring_slots: 0 = every target block its own scratch slot; > 0 = the scratch ring with that many slots
// (ticket-ordered cells + slot_gen), as csrc/mapc.cu uses it for unsharded fused steps.
// staging: 0 LDG/STS + LDS broadcast (default), 1 TMA bulk copies (MAPC_TMA=1), 2 warp-shuffle broadcast (MAPC_SHFL=1).
// returns 0, or -1 for an unknown shape / variant, -2 for a layout the library would refuse (peer with
// straddling segments)
int emu_step_allpairs(const mapc_posvelo *in, mapc_posvelo *out, float *pos_next_out, int n, int n_active,
                      float dt, float damping, int S, int pairs, int threads, int fuse, int mass_in_loop,
                      int world, int peer, int block_order, int chain, int staging, int ring_slots,
                      unsigned long long *info)
{
    if (n % world) return -2;
    const int n_local = n / world;
    const int n_sources = n_active;
    const float nan = std::numeric_limits<float>::quiet_NaN();
    unsigned long long launches = 0, fence_word = 0;
    bool counters_clean = true;

    // read-side packed positions per rank
    std::vector<std::vector<float4>> packed(world, std::vector<float4>(n));
    for (int r = 0; r < world; ++r)
        for (int i = 0; i < n; ++i) {
            const bool own = i >= r * n_local && i < (r + 1) * n_local;
            if (peer && world > 1 && !own) packed[r][i] = make_float4(nan, nan, nan, nan);
            else packed[r][i] = make_float4(in[i].pos[0], in[i].pos[1], in[i].pos[2], in[i].pos[3]);
        }
    std::vector<unsigned long long> flags(world, 7);   // every owner has published step 7

    for (int r = 0; r < world; ++r) {
        const int i_first = r * n_local;
        const int n_targets = mapc::local_targets((uint32_t)n, (uint32_t)i_first, (uint32_t)n_local, n_active);
        if (n_targets <= 0) continue;
        const int per_block = threads * 2 * pairs;
        std::vector<PV> in_local(n_local), out_local(n_local);
        std::memcpy(in_local.data(), in + i_first, (size_t)n_local * sizeof(PV));
        std::memcpy(out_local.data(), out + i_first, (size_t)n_local * sizeof(PV));
        const int n_iblocks = (n_targets + per_block - 1) / per_block;
        const bool ring = ring_slots > 0 && ring_slots < n_iblocks && fuse && world == 1;
        const int slots = ring ? ring_slots : n_iblocks;
        std::vector<float4> partial((size_t)slots * S * per_block, make_float4(nan, nan, nan, nan));
        std::vector<float4> pos_next(n, make_float4(nan, nan, nan, nan));
        std::vector<unsigned> counters(n_local / 64 + 2, 0u);
        std::vector<unsigned> slot_gen(slots, 0u);
        unsigned done[3] = {0, 0, 0};   // [0] target blocks integrated, [1] cell ticket, [2] last step completed host-visibly
        unsigned long long stamps[2] = {0, 0};

        StepArgs a{};
        a.pos = packed[r].data();
        a.partial = partial.data();
        a.scratch_blocks = slots;
        a.group_blocks = (ring && slots >= 4) ? slots / 2 : 0;    // like csrc/mapc.cu
        a.ticket = ring ? &done[1] : nullptr;
        a.slot_gen = ring ? slot_gen.data() : nullptr;
        a.i_first = i_first;
        a.i_cnt = n_targets;
        a.n_sources = n_sources;
        a.S = S;
        a.n_iblocks = n_iblocks;
        a.counters = counters.data();
        a.in = reinterpret_cast<const mapc_posvelo *>(in_local.data());
        a.out = reinterpret_cast<mapc_posvelo *>(out_local.data());
        a.pos_next = pos_next.data();
        a.dt = dt;
        a.damping = damping;
        a.done = done;
        a.stamp_begin = fuse ? &stamps[0] : nullptr;
        a.stamp_end = fuse ? &stamps[1] : nullptr;
        a.fence_word = (fuse && world == 1) ? &fence_word : nullptr;
        a.fence_value = 42;

        // csrc/step_layout.hpp, the function csrc/mapc.cu uses: with the NCCL layout the emulated step is the
        // steady state (a gather of the read side is outstanding, so only the shard's own segments are local)
        const mapc::StepLayout lay = mapc::classify_segments(n_sources, S, i_first, n_local, r, world, peer != 0,
                                                             /*gather_pending=*/world > 1 && !peer);
        const mapc::SegList &local = lay.local, &remote = lay.remote;
        const int *owner = lay.owner;
        const bool peer_ok = peer && fuse && n_sources == n && !mass_in_loop && chain == MAPC_CHAIN_SOURCES && lay.aligned;
        if (peer && world > 1 && !peer_ok && remote.count > 0) return -2;

        if (peer == 2 && peer_ok && world > 1 && remote.count > 0) {
            // MAPC_PEER_SINGLE=1 as csrc/mapc.cu builds it: one grid, local segments first, then the remote ones
            mapc::SegList all{0, {}};
            for (int k = 0; k < local.count; ++k) {
                a.seg_src[all.count] = packed[r].data();
                a.seg_flag[all.count] = nullptr;
                all.ids[all.count++] = local.ids[k];
            }
            for (int k = 0; k < remote.count; ++k) {
                a.seg_src[all.count] = packed[owner[remote.ids[k]]].data();
                a.seg_flag[all.count] = &flags[owner[remote.ids[k]]];
                all.ids[all.count++] = remote.ids[k];
            }
            a.flag_expect = 7;
            a.segs = all;
            if (!launch_force(pairs, threads, fuse, true, mass_in_loop, chain, 0, a, block_order)) return -1;
            ++launches;
        } else {
        a.segs = local;
        if (!launch_force(pairs, threads, fuse, false, mass_in_loop, chain, staging, a, block_order)) return -1;
        launches += (local.count > 0);
        if (local.count > 0) a.stamp_begin = nullptr;
        if (remote.count > 0) {
            a.segs = remote;
            const bool use_peer = peer_ok && world > 1;
            if (use_peer) {
                for (int k = 0; k < remote.count; ++k) {
                    a.seg_src[k] = packed[owner[remote.ids[k]]].data();
                    a.seg_flag[k] = &flags[owner[remote.ids[k]]];
                }
                a.flag_expect = 7;
            }
            if (!launch_force(pairs, threads, fuse, use_peer, mass_in_loop, chain, use_peer ? 0 : staging, a, block_order)) return -1;
            ++launches;
        }
        }
        if (!fuse) {
            // MAPC_FUSE=0: the separate combine + integrate, as csrc/mapc.cu launches it
            cuda_emu::launch(dim3((n_targets + 255) / 256), dim3(256), 0, [&] {
                mapc::integrate_kernel(a.in, a.out, a.pos_next, a.partial, per_block, S, i_first, n_targets, dt, damping);
            });
            ++launches;
        }
        for (unsigned c : counters) counters_clean = counters_clean && c == 0u;
        counters_clean = counters_clean && done[0] == 0u && done[1] == 0u;
        for (unsigned g : slot_gen) counters_clean = counters_clean && g == 0u;
        if (fuse && !(stamps[0] != 0 && stamps[1] >= stamps[0])) counters_clean = false;
        std::memcpy(out + i_first, out_local.data(), (size_t)n_local * sizeof(PV));
        for (int i = 0; i < n_local; ++i) std::memcpy(pos_next_out + 4 * (size_t)(i_first + i), &pos_next[i_first + i], 16);
    }
    if (info) {
        info[0] = launches;
        info[1] = fence_word;
        info[2] = counters_clean ? 1 : 0;
    }
    return 0;
}

// `steps` consecutive fused steps of an unsharded handle with the step-to-step dataflow flags of csrc/mapc.cu: the
// first step waits for "the previous grid" (nothing to wait for here) and publishes block_step, every later one is
// launched with wait_prev = 1 and waits per cell for the target blocks it reads.  Blocks run one after the other in
// the emulation, so a wait can only spin forever if it names a block that is never published or a wrong step id
// -- which is what this checks (with `poison`, flags of the blocks a step must NOT need are left behind).
// state: n PosVelo in, the state after `steps` steps out.  Returns 0, -1 unknown shape, -3 a wait would not end.
int emu_steps_chained(mapc_posvelo *state, int n, int steps, float dt, float damping, int S, int pairs, int threads,
                      int block_order)
{
    const int per_block = threads * 2 * pairs;
    const int n_iblocks = (n + per_block - 1) / per_block;
    std::vector<PV> side[2] = {std::vector<PV>(n), std::vector<PV>(n)};
    std::memcpy(side[0].data(), state, (size_t)n * sizeof(PV));
    std::memcpy(side[1].data(), state, (size_t)n * sizeof(PV));
    std::vector<float4> packed[2] = {std::vector<float4>(n), std::vector<float4>(n)};
    for (int sd = 0; sd < 2; ++sd)
        for (int i = 0; i < n; ++i) packed[sd][i] = make_float4(state[i].pos[0], state[i].pos[1], state[i].pos[2], state[i].pos[3]);
    std::vector<float4> partial((size_t)n_iblocks * S * per_block);
    std::vector<unsigned> counters(n / 64 + 2, 0u), block_step(n / 64 + 2, 0u);
    unsigned done[3] = {0, 0, 0};
    unsigned long long error_word[2] = {0, 0};
    int b = 0;   // write side; both sides start alike, the first step reads side 1
    for (int k = 0; k < steps; ++k, b ^= 1) {
        StepArgs a{};
        a.pos = packed[1 - b].data();
        a.partial = partial.data();
        a.scratch_blocks = n_iblocks;
        a.i_first = 0;
        a.i_cnt = n;
        a.n_sources = n;
        a.S = S;
        a.n_iblocks = n_iblocks;
        a.counters = counters.data();
        a.in = reinterpret_cast<const mapc_posvelo *>(side[1 - b].data());
        a.out = reinterpret_cast<mapc_posvelo *>(side[b].data());
        a.pos_next = packed[b].data();
        a.dt = dt;
        a.damping = damping;
        a.done = done;
        a.block_step = block_step.data();
        a.step_id = (unsigned)(k + 1);
        a.wait_prev = k > 0 ? 1 : 0;
        a.error_word = error_word;
        a.wait_timeout_ns = 2000000000ull;   // a wait that cannot end is reported after 2 s instead of hanging the test
        a.segs.count = S;
        for (int s = 0; s < S; ++s) a.segs.ids[s] = s;
        if (!launch_force(pairs, threads, true, false, false, MAPC_CHAIN_SOURCES, 0, a, block_order)) return -1;
        if (error_word[0] != 0) return -3;
        for (int ib = 0; ib < n_iblocks; ++ib)
            if (block_step[ib] != (unsigned)(k + 1)) return -3;
    }
    std::memcpy(state, side[1 - b].data(), (size_t)n * sizeof(PV));
    return 0;
}

// The literal CSMain step (well_step_kernel) of one rank's shard [i_first, i_first + n_local): out and
// pos_next_out as above.  Also runs pack_positions_kernel over `in` into packed_out (n float4).
int emu_step_well(const mapc_posvelo *in, mapc_posvelo *out, float *pos_next_out, float *packed_out, int n,
                  int n_active, float dt, float damping, int i_first, int n_local)
{
    const int n_targets = mapc::local_targets((uint32_t)n, (uint32_t)i_first, (uint32_t)n_local, n_active);
    std::vector<PV> in_local(n_local), out_local(n_local);
    std::memcpy(in_local.data(), in + i_first, (size_t)n_local * sizeof(PV));
    std::memcpy(out_local.data(), out + i_first, (size_t)n_local * sizeof(PV));
    const float nan = std::numeric_limits<float>::quiet_NaN();
    std::vector<float4> pos_next(n, make_float4(nan, nan, nan, nan)), packed(n, make_float4(nan, nan, nan, nan));
    const mapc_posvelo *pin = reinterpret_cast<const mapc_posvelo *>(in_local.data());
    mapc_posvelo *pout = reinterpret_cast<mapc_posvelo *>(out_local.data());
    if (n_targets > 0)
        cuda_emu::launch(dim3((n_targets + mapc::kWellThreads * mapc::kWellBodies - 1) / (mapc::kWellThreads * mapc::kWellBodies)),
                         dim3(mapc::kWellThreads), 0, [&] {
            mapc::well_step_kernel(pin, pout, pos_next.data(), i_first, n_targets, dt, damping);
        });
    std::vector<PV> all(n);
    std::memcpy(all.data(), in, (size_t)n * sizeof(PV));
    cuda_emu::launch(dim3((n + 255) / 256), dim3(256), 0, [&] {
        mapc::pack_positions_kernel(reinterpret_cast<const mapc_posvelo *>(all.data()), packed.data(), n);
    });
    std::memcpy(out + i_first, out_local.data(), (size_t)n_local * sizeof(PV));
    std::memcpy(pos_next_out, pos_next.data(), (size_t)n * 16);
    std::memcpy(packed_out, packed.data(), (size_t)n * 16);
    return 0;
}

// init_particles_kernel for the shard [i_first, i_first + n_local) of n bodies: side_a / side_b n_local PosVelo
// each, packed_a / packed_b n float4 each.
int emu_init_particles(mapc_posvelo *side_a, mapc_posvelo *side_b, float *packed_a, float *packed_b, unsigned n,
                       unsigned i_first, unsigned n_local, unsigned seed)
{
    cuda_emu::launch(dim3((n + 255) / 256), dim3(256), 0, [&] {
        mapc::init_particles_kernel(side_a, side_b, reinterpret_cast<float4 *>(packed_a), reinterpret_cast<float4 *>(packed_b),
                                    n, i_first, n_local, seed, MAPC_PARTICLE_SPREAD * 0.750f, MAPC_INITIAL_PARTICLE_SPEED,
                                    MAPC_PARTICLE_SPREAD);
    });
    return 0;
}

// csrc/step_layout.hpp make_plan / local_targets, for the host-logic tests: out = {pairs, threads, blocks_x, S}
void emu_make_plan(int n_targets, int S, int sm_count, int force_pairs, int force_threads, int *out)
{
    const mapc::Plan pl = mapc::make_plan(n_targets, S, sm_count, force_pairs, force_threads);
    out[0] = pl.pairs; out[1] = pl.threads; out[2] = pl.blocks_x; out[3] = pl.segments;
}

// csrc/step_layout.hpp throttle_blocks_per_sm for the plan make_plan picks: resident blocks per SM, 0 = kernel's own
int emu_throttle_blocks_per_sm(int n_targets, int S, int sm_count, int force_pairs, int force_threads)
{
    return mapc::throttle_blocks_per_sm(mapc::make_plan(n_targets, S, sm_count, force_pairs, force_threads), sm_count);
}

int emu_local_targets(unsigned n, unsigned i_first, unsigned n_local, int n_active)
{
    return mapc::local_targets(n, i_first, n_local, n_active);
}

}  // extern "C"
