// step_layout.hpp -- the pure host-side decisions of one all-pairs step: which targets a Simulate(n_active)
// updates on a shard, which launch shape (P, T) runs them, and which canonical segments a rank evaluates at
// once (their sources are resident) or after the exchange (remote).  No CUDA calls: csrc/mapc.cu applies the
// result to streams and kernels, tests/emu/emu_kernels.cpp applies the very same functions to the emulated
// kernels, so the CPU suite exercises this logic too.
#pragma once

#include <algorithm>
#include <cmath>

#include "nbody_kernels.cuh"

namespace mapc {

struct Plan {
    int pairs;    // P: register pairs per thread (2P targets per thread)
    int threads;  // T
    int blocks_x; // target blocks of T*2P bodies
    int segments; // canonical S
    int minb;     // resident blocks per SM the shape is compiled for (csrc/force_shapes.inc)
    int block_targets() const { return threads * 2 * pairs; }
};

// Launch shapes (P, T) with the fraction of the FP32 peak each reaches at large N, measured the way the headline is
// (N = 262,144, L2 flushed between steps; profiles/r02_shape_variants.txt; the P = 1 shapes scaled from the batched
// sweep at N = 131,072, profiles/r02_small_n_shapes.txt).  At large N (2,128) wins by 1.3 %; at small N what decides is
// how many of the SMs' block slots the cells fill at all, then how evenly they fall on the 148 SMs x 4 warp
// schedulers and how many target lanes the last block leaves idle: (1,128) is the fastest shape up to N ~ 12,000.
// The shape changes neither the arithmetic nor its order, so it is free to vary with N, the shard size and the device.
struct Shape { int pairs, threads, minb; float efficiency; };
constexpr Shape kShapes[7] = {{4, 256, 2, 0.771f}, {4, 128, 4, 0.762f}, {2, 128, 8, 0.784f}, {2, 64, 16, 0.781f},
                              {1, 128, 8, 0.724f}, {1, 64, 16, 0.716f}, {1, 32, 32, 0.721f}};

// S = mapc_plan_segments(n_sources); force_pairs / force_threads != 0 pin the shape (MAPC_PLAN_PAIRS / _THREADS)
inline Plan make_plan(int n_targets, int S, int sm_count, int force_pairs, int force_threads)
{
    Plan best{1, 32, 0, S, 32};
    float best_score = -1.f;
    for (const Shape &sh : kShapes) {
        if ((force_pairs && force_pairs != sh.pairs) || (force_threads && force_threads != sh.threads)) continue;
        const int per_block = sh.threads * 2 * sh.pairs;
        const int bx = (n_targets + per_block - 1) / per_block;
        if (bx == 0) continue;
        const double used = (double)n_targets / ((double)bx * per_block);            // busy target lanes
        const double cells = (double)bx * S;
        const double blocks_per_sm = cells / sm_count;
        const double warps_per_smsp = blocks_per_sm * (sh.threads / 32) / 4.0;
        const double bal_block = blocks_per_sm / std::ceil(blocks_per_sm);
        const double bal_warp = warps_per_smsp / std::ceil(warps_per_smsp);
        // a grid that does not even fill the resident block slots runs latency-bound: the plateau figure scales with
        // the filled fraction (measured: (4,256) at N = 10,000 fills 0.54 of its slots and takes 2.2x the time of (1,128))
        const double occupancy = std::min(1.0, cells / ((double)sm_count * sh.minb));
        const float score = (float)(sh.efficiency * used * std::min(bal_block, bal_warp) * occupancy);
        if (score > best_score) {
            best_score = score;
            best = Plan{sh.pairs, sh.threads, bx, S, sh.minb};
        }
    }
    return best;
}

// Occupancy throttle of steps that chain by data flow (unsharded fused steps below the scratch-ring size, DESIGN.md
// section 4).  Such a step cannot finish a single target block before the previous step has published its LAST one
// (every target needs every source), so what bounds a run of small steps is how long a cell takes once its inputs are
// there -- and a cell that shares its SM with seven others takes four times as long as one that shares it with one.
// Fewer resident blocks per SM = the same throughput (the 8-source unroll of the P = 1 kernels keeps the FMA pipe fed
// from 2-4 warps per scheduler) at a fraction of the latency.  Measured on B200 with the (1,128) shape, us per chained
// step for 8 / best k resident blocks per SM (profiles/r02_small_n_occupancy.txt): N = 2,500 15.9 / 12.7 (k = 2),
// 4,096 19.6 / 15.9 (3), 6,000 27.3 / 22.5 (4), 8,192 33.9 / 29.9 (4-5), 10,000 41.7 / 40.5 (6), >= 12,000 no gain (8):
// the best k keeps about three quarters of a step's cells resident.  Returns 0 when the kernel's own occupancy is kept.
inline int throttle_blocks_per_sm(const Plan &pl, int sm_count)
{
    const double cells = (double)pl.blocks_x * pl.segments;
    int k = (int)std::floor(0.75 * cells / sm_count + 0.5);
    if (k < 1) k = 1;
    return k < pl.minb ? k : 0;
}

// targets of the shard [i_first, i_first + n_local) that a Simulate(n_active) updates, as a count from
// i_first: Dispatch(ceil(n/64)) groups of 64 threads (Compute.cpp:1041); writes past N are dropped
inline int local_targets(uint32_t n, uint32_t i_first, uint32_t n_local, int n_active)
{
    if (n_active <= 0) return 0;
    long long t = ((long long)n_active + MAPC_BLOCK_SIZE - 1) / MAPC_BLOCK_SIZE * MAPC_BLOCK_SIZE;
    if (t > (long long)n) t = n;
    long long loc = t - (long long)i_first;
    if (loc < 0) loc = 0;
    if (loc > (long long)n_local) loc = n_local;
    return (int)loc;
}

// Partials scratch of one step: [slot][S][block targets] float4.  Every target block gets its own slot
// (slots == blocks_x), except for an unsharded fused step whose scratch would not fit the ring budget: then
// `slots` < blocks_x slots are reused round-robin while they are still in L2 (force_cells_kernel: ticket +
// slot_gen).  The ring holds at least twice the target blocks that can be in flight at once, so the
// release wait practically never spins; correctness does not depend on that.
struct Scratch {
    int slots;
    bool ring;
    size_t bytes;
};
constexpr size_t kScratchRingBytes = 32u << 20;   // a quarter of the 126 MB L2

inline Scratch plan_scratch(const Plan &pl, int sm_count, bool allow_ring)
{
    const size_t per_block = (size_t)pl.segments * pl.block_targets() * 16u;
    Scratch sc{pl.blocks_x, false, per_block * (size_t)pl.blocks_x};
    if (!allow_ring) return sc;
    const int in_flight = (sm_count * 2 * pl.minb + pl.segments - 1) / pl.segments + 2;
    const int by_budget = (int)(kScratchRingBytes / per_block);
    const int slots = std::max(in_flight, by_budget);
    if (slots < pl.blocks_x) sc = Scratch{slots, true, per_block * (size_t)slots};
    return sc;
}

// Which canonical segments a rank can evaluate straight away and which must wait for the exchange.
//   local   every segment of an unsharded handle; on a shard the segments whose sources all lie inside it
//           (the rank wrote them itself); and, without a peer exchange, everything while no gather of the read
//           side is outstanding (right after an upload every rank holds all N positions)
//   remote  the rest: second launch, after the all-gather or through the owners' memory
//   owner[s]   rank whose shard holds segment s (a segment never counts as owned by two)
//   aligned    no segment straddles two shards (what the peer exchange needs)
struct StepLayout {
    SegList local, remote;
    int owner[MAPC_MAX_SEGMENTS];
    bool aligned;
};

inline StepLayout classify_segments(int n_sources, int S, int i_first, int n_local, int rank, int world,
                                    bool peer_mode, bool gather_pending)
{
    StepLayout out{};
    out.aligned = true;
    for (int s = 0; s < S; ++s) {
        int j0, j1;
        segment_range(n_sources, S, s, j0, j1);
        const bool inside = j0 >= i_first && j1 <= i_first + n_local;
        out.owner[s] = j1 > j0 ? j0 / n_local : rank;
        if (j1 > j0 && (j1 - 1) / n_local != out.owner[s]) out.aligned = false;  // straddles two shards
        const bool is_local = world == 1 || inside || (!peer_mode && !gather_pending);
        SegList &list = is_local ? out.local : out.remote;
        list.ids[list.count++] = s;
    }
    return out;
}

}  // namespace mapc
