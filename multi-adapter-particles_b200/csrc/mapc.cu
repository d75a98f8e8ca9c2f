// mapc.cu -- C ABI (include/mapc.h) and the Compute component behind it.
//
// Replaces Particles/Compute.{h,cpp} of the reference for the simulation path: the D3D12 compute
// queue, command lists, root signature / PSO, descriptor heap and cross-adapter heap become CUDA
// streams and device buffers; ID3D12Fence becomes mapc_fence (a 64-bit word every device and the
// host can signal / wait on); the D3D12GpuTimer becomes %globaltimer stamps written by the force
// kernel (cudaEvent pairs for the other kernels) with the same 20-sample moving average.  There is deliberately NO CPU fallback: without a CUDA device every entry
// point that needs one returns MAPC_ERR_NO_DEVICE / MAPC_ERR_CUDA.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // declarations only: libnccl is dlopen()ed when a sharded handle is created
#include <nvtx3/nvToolsExt.h>  // header-only; ranges are emitted only with MAPC_NVTX=1
#include <sched.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/mapc.h"
#include "fence.hpp"
#include "nbody_kernels.cuh"
#include "step_layout.hpp"

namespace {

thread_local std::string g_last_error;

mapc_status fail(mapc_status st, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return st;
}

#define MAPC_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            const mapc_status _st = (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) \
                                        ? MAPC_ERR_NO_DEVICE                                  \
                                        : (_e == cudaErrorMemoryAllocation ? MAPC_ERR_OUT_OF_MEMORY \
                                                                           : MAPC_ERR_CUDA);  \
            return fail(_st, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                            \
        }                                                                                     \
    } while (0)

#define MAPC_TRY(expr)                      \
    do {                                    \
        mapc_status _s = (expr);            \
        if (_s != MAPC_OK) return _s;       \
    } while (0)

// ---- driver entry points for stream memory operations (no link-time libcuda dependency) -----
typedef CUresult (*pfn_stream_value64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
pfn_stream_value64 g_write64 = nullptr, g_wait64 = nullptr;

mapc_status load_stream_memops()
{
    if (g_write64 && g_wait64) return MAPC_OK;
    void *w = nullptr, *q = nullptr;
    cudaDriverEntryPointQueryResult r1, r2;
    MAPC_CUDA(cudaGetDriverEntryPoint("cuStreamWriteValue64", &w, cudaEnableDefault, &r1));
    MAPC_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue64", &q, cudaEnableDefault, &r2));
    if (!w || !q || r1 != cudaDriverEntryPointSuccess || r2 != cudaDriverEntryPointSuccess)
        return fail(MAPC_ERR_UNSUPPORTED, "driver lacks cuStreamWriteValue64/cuStreamWaitValue64");
    g_write64 = (pfn_stream_value64)w;
    g_wait64 = (pfn_stream_value64)q;
    return MAPC_OK;
}

// ---- NCCL, loaded on demand ---------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclCommGetAsyncError) CommGetAsyncError = nullptr;   // optional
} g_nccl;

mapc_status load_nccl()
{
    if (g_nccl.lib) return MAPC_OK;
    const char *names[] = {getenv("MAPC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
    }
    if (!lib) return fail(MAPC_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define MAPC_SYM(field, name)                                                         \
    g_nccl.field = (decltype(g_nccl.field))dlsym(lib, name);                          \
    if (!g_nccl.field) return fail(MAPC_ERR_NCCL, "libnccl lacks symbol %s", name)
    MAPC_SYM(GetUniqueId, "ncclGetUniqueId");
    MAPC_SYM(CommInitRank, "ncclCommInitRank");
    MAPC_SYM(CommDestroy, "ncclCommDestroy");
    MAPC_SYM(AllGather, "ncclAllGather");
    MAPC_SYM(GetErrorString, "ncclGetErrorString");
    MAPC_SYM(GetVersion, "ncclGetVersion");
#undef MAPC_SYM
    g_nccl.CommGetAsyncError = (decltype(g_nccl.CommGetAsyncError))dlsym(lib, "ncclCommGetAsyncError");
    g_nccl.lib = lib;
    return MAPC_OK;
}

#define MAPC_NCCL(expr)                                                                  \
    do {                                                                                 \
        ncclResult_t _r = (expr);                                                        \
        if (_r != ncclSuccess)                                                           \
            return fail(MAPC_ERR_NCCL, "%s failed: %s (%s:%d)", #expr,                   \
                        g_nccl.GetErrorString(_r), __FILE__, __LINE__);                  \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Named ranges for timelines (the reference labels its D3D12 objects for PIX, Compute.cpp:428-472): off by
// default so the per-call path stays as measured; MAPC_NVTX=1 (read once) brackets the entry points below.
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char *name)
    {
        static const bool enabled = [] { const char *v = getenv("MAPC_NVTX"); return v && *v && atoi(v) != 0; }();
        on = enabled;
        if (on) nvtxRangePushA(name);
    }
    ~NvtxRange()
    {
        if (on) nvtxRangePop();
    }
};

}  // namespace

// ---- fences and gated streams (fence.hpp) -------------------------------------------------------
namespace mapc {

uint64_t fence_completed(const mapc_fence *f)
{
    return __atomic_load_n((const uint64_t *)f->word, __ATOMIC_ACQUIRE);
}

void fence_notify(mapc_fence *f)
{
    const std::vector<GatedStream *> waiters = f->waiters;  // draining may edit the list
    for (GatedStream *gs : waiters) gs_drain(gs);
}

static mapc_status fence_take_event(mapc_fence *f, int device, cudaEvent_t *out)
{
    for (size_t k = 0; k < f->spare.size(); ++k)
        if (f->spare_device[k] == device) {
            *out = f->spare[k];
            f->spare.erase(f->spare.begin() + (long)k);
            f->spare_device.erase(f->spare_device.begin() + (long)k);
            return MAPC_OK;
        }
    DeviceGuard g(device);
    MAPC_CUDA(cudaEventCreateWithFlags(out, cudaEventDisableTiming));
    return MAPC_OK;
}

mapc_status fence_submit_signal(mapc_fence *f, cudaStream_t stream, int device, uint64_t value)
{
    MAPC_TRY(load_stream_memops());
    DeviceGuard g(device);
    // drop signals that have completed (their waits are no-ops from now on)
    const uint64_t done = fence_completed(f);
    while (f->signals.size() > 4 && f->signals.front().value <= done) {
        f->spare.push_back(f->signals.front().event);   // recycled by a later signal from the same device
        f->spare_device.push_back(f->signals.front().device);
        f->signals.pop_front();
    }
    cudaEvent_t ev = nullptr;
    MAPC_TRY(fence_take_event(f, device, &ev));
    MAPC_CUDA(cudaEventRecord(ev, stream));
    const CUresult r = g_write64((CUstream)stream, (CUdeviceptr)(uintptr_t)f->word, value,
                                 CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) {
        cudaEventDestroy(ev);
        return fail(MAPC_ERR_CUDA, "cuStreamWriteValue64 failed: %d", (int)r);
    }
    f->signals.push_back(FenceSignal{value, ev, device});
    if (value > f->submitted) f->submitted = value;
    fence_notify(f);
    return MAPC_OK;
}

mapc_status fence_submit_wait(mapc_fence *f, cudaStream_t stream, uint64_t value)
{
    // A completed value needs no wait -- unless it was signalled by a kernel writing the word itself: then the
    // word says "the step is done", and the waiting stream still gets a CUDA-visible edge to the signalling
    // stream (an event, below), so a copy engine never reads the step's output on the strength of a host-side
    // poll alone.
    const bool light = f->light_stream != nullptr && value <= f->light_value && value > f->light_floor;
    if (!light && fence_completed(f) >= value) return MAPC_OK;
    for (const FenceSignal &sig : f->signals)
        if (sig.value >= value) {
            MAPC_CUDA(cudaStreamWaitEvent(stream, sig.event, 0));
            return MAPC_OK;
        }
    if (f->light_stream != nullptr && f->light_value >= value) {
        // signalled by a kernel that writes the word itself: make the dependency CUDA-visible now, with an
        // event that covers everything submitted to the signalling stream so far
        cudaEvent_t ev = nullptr;
        MAPC_TRY(fence_take_event(f, f->light_device, &ev));
        {
            DeviceGuard g(f->light_device);
            MAPC_CUDA(cudaEventRecord(ev, f->light_stream));
        }
        f->signals.push_back(FenceSignal{f->light_value, ev, f->light_device});
        MAPC_CUDA(cudaStreamWaitEvent(stream, ev, 0));
        return MAPC_OK;
    }
    // submitted >= value but no event: the value was signalled from the host and is visible already
    return MAPC_OK;
}

static mapc_status run_op(GatedStream *gs, StreamOp &op)
{
    DeviceGuard g(gs->device);
    switch (op.kind) {
    case StreamOp::kWait: return fence_submit_wait(op.fence, gs->stream, op.value);
    case StreamOp::kSignal: return fence_submit_signal(op.fence, gs->stream, gs->device, op.value);
    case StreamOp::kSignalLight: {
        // the kernel just enqueued on this stream writes the word when it finishes; no stream operation
        mapc_fence *f = op.fence;
        const uint64_t done = fence_completed(f);
        while (f->signals.size() > 4 && f->signals.front().value <= done) {
            f->spare.push_back(f->signals.front().event);
            f->spare_device.push_back(f->signals.front().device);
            f->signals.pop_front();
        }
        if (f->light_stream == nullptr) f->light_floor = op.value - 1;   // values below were signalled by stream operations
        f->light_stream = gs->stream;
        f->light_device = gs->device;
        if (op.value > f->light_value) f->light_value = op.value;
        if (op.value > f->submitted) f->submitted = op.value;
        fence_notify(f);
        return MAPC_OK;
    }
    default: return op.fn();
    }
}

mapc_status gs_drain(GatedStream *gs)
{
    if (gs->draining) return MAPC_OK;  // an outer frame of this same drain continues the loop
    gs->draining = true;
    mapc_status st = MAPC_OK;
    while (!gs->pending.empty()) {
        StreamOp &op = gs->pending.front();
        if (op.kind == StreamOp::kWait && !fence_ready(op.fence, op.value)) break;
        StreamOp local = std::move(op);
        gs->pending.pop_front();
        if (local.kind == StreamOp::kWait) {
            auto &w = local.fence->waiters;
            bool still = false;
            for (const StreamOp &o : gs->pending) still = still || (o.kind == StreamOp::kWait && o.fence == local.fence);
            if (!still) w.erase(std::remove(w.begin(), w.end(), gs), w.end());
        }
        st = run_op(gs, local);
        if (st != MAPC_OK) {
            if (gs->deferred_error == MAPC_OK) {   // sticky: see GatedStream::deferred_error
                gs->deferred_error = st;
                gs->deferred_message = g_last_error;
            }
            break;
        }
    }
    gs->draining = false;
    return st;
}

static mapc_status gs_push(GatedStream *gs, StreamOp op)
{
    if (gs->deferred_error != MAPC_OK)
        return fail(gs->deferred_error, "an earlier queued operation of this stream failed: %s",
                    gs->deferred_message.c_str());
    if (gs->pending.empty() && !(op.kind == StreamOp::kWait && !fence_ready(op.fence, op.value)))
        return run_op(gs, op);
    if (op.kind == StreamOp::kWait) {
        auto &w = op.fence->waiters;
        if (std::find(w.begin(), w.end(), gs) == w.end()) w.push_back(gs);
    }
    gs->pending.push_back(std::move(op));
    return gs_drain(gs);
}

mapc_status gs_wait(GatedStream *gs, mapc_fence *f, uint64_t value)
{
    return gs_push(gs, StreamOp{StreamOp::kWait, f, value, nullptr});
}

mapc_status gs_signal(GatedStream *gs, mapc_fence *f, uint64_t value)
{
    return gs_push(gs, StreamOp{StreamOp::kSignal, f, value, nullptr});
}

mapc_status gs_signal_light(GatedStream *gs, mapc_fence *f, uint64_t value)
{
    return gs_push(gs, StreamOp{StreamOp::kSignalLight, f, value, nullptr});
}

mapc_status gs_call(GatedStream *gs, std::function<mapc_status()> fn)
{
    return gs_push(gs, StreamOp{StreamOp::kCall, nullptr, 0, std::move(fn)});
}

void gs_detach(GatedStream *gs)
{
    for (StreamOp &op : gs->pending)
        if (op.kind == StreamOp::kWait && op.fence) {
            auto &w = op.fence->waiters;
            w.erase(std::remove(w.begin(), w.end(), gs), w.end());
        }
    gs->pending.clear();
}

// A fence is going away: whatever is still queued against it, on any stream, becomes a no-op (its waits
// count as satisfied, its signals have nobody left to see them).
static void fence_scrub(mapc_fence *f)
{
    const std::vector<GatedStream *> waiters = f->waiters;
    for (GatedStream *gs : waiters)
        for (StreamOp &op : gs->pending)
            if (op.fence == f) {
                op.kind = StreamOp::kCall;
                op.fence = nullptr;
                op.fn = []() -> mapc_status { return MAPC_OK; };
            }
    f->waiters.clear();
}

// a stream that still has host-queued work is waiting for a signal nobody has submitted
static mapc_status require_ungated(const GatedStream *gs, const char *what)
{
    if (gs->pending.empty()) return MAPC_OK;
    for (const StreamOp &op : gs->pending)
        if (op.kind == StreamOp::kWait)
            return fail(MAPC_ERR_TIMEOUT,
                        "%s: the stream is gated on fence value %llu whose signal has not been submitted "
                        "(completed %llu) -- it would never finish",
                        what, (unsigned long long)op.value, (unsigned long long)fence_completed(op.fence));
    return fail(MAPC_ERR_TIMEOUT, "%s: stream has unsubmitted work", what);
}

}  // namespace mapc

// ---- launch plan ------------------------------------------------------------------------------
namespace {

using mapc::Plan;

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// Every A/B switch of the library, read from the environment once per API call (so a queued, gated
// step runs with the switches of the call that issued it).  Defaults are the measured-best path; the
// others exist so tests and profiles can show the alternatives are the same arithmetic.
struct Switches {
    bool fuse;          // MAPC_FUSE=0: separate integrate_kernel instead of the last-arrival combine
    bool pdl;           // MAPC_PDL=0: no programmatic dependent launch between consecutive steps
    bool tma;           // MAPC_TMA=1: cp.async.bulk source staging instead of LDG/STS
    bool shfl;          // MAPC_SHFL=1: warp-shuffle broadcast of staged sources instead of LDS broadcast
    bool ring_group;    // MAPC_RING_GROUP=0: ring cells in target-block-major order instead of segment-major groups
    bool ring;          // MAPC_RING=0: every target block its own scratch slot (no L2-resident ring)
    int shape_variant;  // MAPC_SHAPE_VARIANT=v: A/B instantiations of the large-N shapes (csrc/force_shapes.inc)
    bool chain;         // MAPC_CHAIN=0: consecutive small-N steps wait for the whole previous grid, not per target block
    int wait_timeout_ms;  // MAPC_WAIT_TIMEOUT_MS: bound of every in-kernel wait (peer step flag, ring slot)
    bool mass_in_loop;  // MAPC_MASS_IN_LOOP=1: 12-op pair with the shader's per-pair mass multiply
    bool timers;        // MAPC_TIMERS=0: no "simulate ms" timer at all
    bool timer_events;  // MAPC_TIMER_EVENTS=1: cudaEvent pairs instead of in-kernel stamps
    bool kernel_fence;  // MAPC_KERNEL_FENCE=0: fence signalled by a stream operation, not by the kernel
    bool peer;          // MAPC_PEER=0: attached peer exchange not used
    bool peer_single;   // MAPC_PEER_SINGLE=1 (experimental): peer exchange as ONE grid, local cells first
    int plan_pairs, plan_threads;  // MAPC_PLAN_PAIRS / MAPC_PLAN_THREADS: force a launch shape (0 = choose)
    int blocks_per_sm;  // MAPC_BLOCKS_PER_SM: resident blocks per SM of chained small-N steps (-1 = choose, 0 = no throttle)
};

Switches read_switches()
{
    Switches w;
    w.fuse = env_int("MAPC_FUSE", 1) != 0;
    w.pdl = env_int("MAPC_PDL", 1) != 0;
    w.tma = env_int("MAPC_TMA", 0) != 0;
    w.shfl = env_int("MAPC_SHFL", 0) != 0;
    w.ring = env_int("MAPC_RING", 1) != 0;
    w.ring_group = env_int("MAPC_RING_GROUP", 1) != 0;
    w.chain = env_int("MAPC_CHAIN", 1) != 0;
    w.shape_variant = env_int("MAPC_SHAPE_VARIANT", 0);
    w.wait_timeout_ms = env_int("MAPC_WAIT_TIMEOUT_MS", 20000);
    w.mass_in_loop = env_int("MAPC_MASS_IN_LOOP", 0) != 0;
    w.timers = env_int("MAPC_TIMERS", 1) != 0;
    w.timer_events = env_int("MAPC_TIMER_EVENTS", 0) != 0;
    w.kernel_fence = env_int("MAPC_KERNEL_FENCE", 1) != 0;
    w.peer = env_int("MAPC_PEER", 1) != 0;
    w.peer_single = env_int("MAPC_PEER_SINGLE", 0) != 0;
    w.plan_pairs = env_int("MAPC_PLAN_PAIRS", 0);
    w.plan_threads = env_int("MAPC_PLAN_THREADS", 0);
    w.blocks_per_sm = env_int("MAPC_BLOCKS_PER_SM", -1);
    return w;
}

// launch shape for a step (csrc/step_layout.hpp), with the canonical S of the sources and the forced
// shape of the switches
Plan make_plan(int n_targets, int n_sources, int sm_count, const Switches &sw)
{
    return mapc::make_plan(n_targets, mapc_plan_segments((uint32_t)n_sources), sm_count, sw.plan_pairs,
                           sw.plan_threads);
}

}  // namespace

// ---- Compute ----------------------------------------------------------------------------------
struct mapc_compute {
    uint32_t n = 0;          // m_numParticles (global)
    uint32_t i_first = 0;    // shard
    uint32_t n_local = 0;
    int device = 0;
    int rank = 0, world = 1;
    int sm_count = 148;
    mapc_force_mode mode = MAPC_FORCE_ALLPAIRS;

    cudaStream_t compute = nullptr;  // m_commandQueue (compute)
    mapc::GatedStream gcompute;      // the same stream behind the fence gate (fence.hpp)
    cudaStream_t comm = nullptr;     // all-gather stream
    cudaStream_t compute2 = nullptr; // sharded runs: remote-segment cells, concurrent with the local ones
    cudaEvent_t ev_step_begin = nullptr, ev_remote_done = nullptr;
    mapc_posvelo *posvelo[2] = {nullptr, nullptr};  // ping-pong sides, local shard
    float4 *packed[2] = {nullptr, nullptr};         // packed positions, all N, per side
    float4 *partial = nullptr;                      // partials scratch [slot][segment][block targets]
    size_t partial_bytes = 0;
    unsigned *slot_gen = nullptr;                   // scratch ring: target blocks combined out of each slot
    int slot_gen_count = 0;
    unsigned *counters = nullptr;                   // per target block: segments finished this step
    int counters_key = 0;                           // block size the counters were last used with

    mapc_fence *fence = nullptr;            // m_fence
    mapc_fence *consumer_fence = nullptr;   // m_sharedRenderFence (borrowed)
    struct mapc_consumer *consumer = nullptr;  // headless consumer attached to this producer, if any
    uint64_t fence_value = 0;               // m_fenceValue
    uint32_t buffer_index = 0;              // m_bufferIndex

    ncclComm_t nccl = nullptr;
    // collective-free exchange (mapc_compute_ipc_attach)
    bool peer_mode = false;
    unsigned long long *flag = nullptr;          // own step flag (device memory, IPC-exported)
    float4 *peer_packed[16][2] = {};             // [rank][side]: own pointers for rank == c->rank
    unsigned long long *peer_flag[16] = {};
    unsigned long long step_id = 0;              // steps issued since attach (same on every rank)
    cudaEvent_t ev_integrated = nullptr;
    cudaEvent_t ev_gathered[2] = {nullptr, nullptr};     // timing-enabled: also the end of the gather
    cudaEvent_t ev_gather_begin[2] = {nullptr, nullptr};

    bool gather_pending[2] = {false, false};

    // "simulate ms" timer (Compute.cpp:445-446, D3D12GpuTimer.h:151-153)
    // steps whose timer may be unresolved at once = how far the host may run ahead of the device before Simulate
    // blocks: 32 steps keep the queue fed through host hiccups even when a step is 40 us (the reference's frame
    // loop never blocks in Simulate at all, only on the consumer's throttle)
    static constexpr int kTimerSlots = 32;
    static constexpr float kAverageOver = 20.f;
    cudaEvent_t t_begin[kTimerSlots] = {}, t_end[kTimerSlots] = {};
    bool t_pending[kTimerSlots] = {};
    int t_steps[kTimerSlots] = {};   // steps covered by the slot's event pair (batched Simulate)
    bool t_stamped[kTimerSlots] = {};          // slot timed by in-kernel %globaltimer stamps, not events
    bool t_begin_is_prev_end[kTimerSlots] = {};  // chained step: its time runs from the previous step's end stamp
    int t_cur_slot = 0;                        // slot of the enqueue in progress
    unsigned long long last_end_ns = 0;        // end stamp of the newest resolved stamped slot
    uint64_t t_fence_value[kTimerSlots] = {};  // fence value signalled after the slot's step(s)
    unsigned long long *stamps = nullptr;      // pinned host: [slot][begin, end] in ns
    unsigned *done = nullptr;                  // device: [0] target blocks integrated this step, [1] cell ticket
    unsigned long long *error_word = nullptr;  // pinned host: [0] != 0 after an in-kernel wait timed out, [1] detail
    // step-to-step dataflow (StepArgs::block_step): per target block, the id of the last step that integrated it
    unsigned *block_step = nullptr;
    unsigned step_counter = 0;                 // id of the last fused unsharded all-pairs step enqueued
    bool chain_valid = false;                  // block_step describes the newest state, written with chain_key's shape
    long long chain_key[4] = {0, 0, 0, 0};     // {pairs, threads, n_targets, n_sources} of that step
    unsigned long long *stamp_begin_next = nullptr, *stamp_end_next = nullptr;  // for the next force launch(es)
    unsigned long long fence_write_next = 0;   // != 0: the step's last block writes this value to the fence word
    uint64_t t_next = 0, t_resolved = 0;
    float ms_average = 0.f, ms_last = 0.f;
    std::vector<float> step_log;  // raw samples not yet handed out by mapc_compute_step_times

    uint64_t launches = 0;
    bool has_state = false;
    bool pdl_next = false;   // next force launch may overlap the previous one's tail (batched steps)
};

namespace {

void resolve_timers(mapc_compute *c, bool block)
{
    while (c->t_resolved < c->t_next) {
        const int slot = (int)(c->t_resolved % mapc_compute::kTimerSlots);
        if (c->t_pending[slot]) {
            float ms = 0.f;
            bool have = false;
            if (c->t_stamped[slot]) {
                // in-kernel stamps: valid once the fence value signalled after the step has completed
                if (mapc_fence_completed_value(c->fence) < c->t_fence_value[slot]) {
                    if (!block) return;
                    if (mapc_fence_wait_host(c->fence, c->t_fence_value[slot], 60000) != MAPC_OK) return;
                }
                // A chained step starts inside the drain of the previous one (its cells wait per target block, not
                // for the grid), so "first cell start -> last integrate" would count the overlap twice: its time
                // runs from the previous step's end stamp instead, like back-to-back dispatches on one queue.
                const unsigned long long t0 = c->t_begin_is_prev_end[slot] ? c->last_end_ns : c->stamps[2 * slot];
                const unsigned long long t1 = c->stamps[2 * slot + 1];
                c->last_end_ns = t1;
                have = t1 > t0 && t0 != 0;
                ms = have ? (float)((double)(t1 - t0) * 1e-6) : 0.f;
            } else {
                if (block) cudaEventSynchronize(c->t_end[slot]);
                else if (cudaEventQuery(c->t_end[slot]) != cudaSuccess) return;
                have = cudaEventElapsedTime(&ms, c->t_begin[slot], c->t_end[slot]) == cudaSuccess;
            }
            if (have) {
                if (c->t_steps[slot] > 1) ms /= (float)c->t_steps[slot];   // per-step average of a batch
                // D3D12GpuTimer.h:151-153: t = (t*(N-1) + delta)/N
                c->ms_average = (c->ms_average * (mapc_compute::kAverageOver - 1.f) + ms) /
                                mapc_compute::kAverageOver;
                c->ms_last = ms;
                if (c->step_log.size() >= 4096) c->step_log.erase(c->step_log.begin());
                c->step_log.push_back(ms);
            }
            c->t_pending[slot] = false;
        }
        ++c->t_resolved;
    }
}

// grid = target blocks x segments of this launch, one cell per thread block, segment index fastest
template <int P, int T, int TJ, int U, int MINB, int ORDER, bool FUSE, bool PEER, bool INLOOP, bool TMA,
          bool SHFL = false>
mapc_status launch_force(mapc_compute *c, const mapc::StepArgs &args, cudaStream_t stream, int blocks_per_sm = 0)
{
    auto kernel = mapc::force_cells_kernel<P, T, TJ, U, MINB, ORDER, FUSE, PEER, TMA, INLOOP, SHFL>;
    dim3 grid((unsigned)args.n_iblocks * (unsigned)args.segs.count, 1, 1);
    // Occupancy throttle (mapc::throttle_blocks_per_sm): at most `blocks_per_sm` blocks of this kernel are resident on
    // an SM when every block also asks for dynamic shared memory it never touches.  The size that yields exactly k is
    // found once per (instantiation, device, k) with the occupancy calculator.
    size_t dyn_smem = 0;
    if (blocks_per_sm > 0 && blocks_per_sm < MINB) {
        constexpr int kMaxDev = 64;
        static int dyn_for[kMaxDev][MINB] = {};     // bytes, 0 = not computed yet
        static bool optin_done[kMaxDev] = {};
        const int dev = c->device >= 0 && c->device < kMaxDev ? c->device : 0;
        if (dyn_for[dev][blocks_per_sm] == 0) {
            cudaFuncAttributes fa;
            MAPC_CUDA(cudaFuncGetAttributes(&fa, kernel));
            int per_sm = 0, reserved = 0, optin = 0;
            MAPC_CUDA(cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->device));
            MAPC_CUDA(cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, c->device));
            MAPC_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
            const int most = optin - (int)fa.sharedSizeBytes;
            if (!optin_done[dev]) {
                MAPC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
                optin_done[dev] = true;
            }
            int dyn = per_sm / blocks_per_sm - reserved - (int)fa.sharedSizeBytes;
            dyn = std::max(1024, std::min(most, dyn / 1024 * 1024));
            for (int tries = 0; tries < 64; ++tries) {   // settle on the largest size that still admits k blocks
                int nb = 0;
                MAPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, T, (size_t)dyn));
                if (nb == blocks_per_sm || (nb < blocks_per_sm && dyn <= 1024) || (nb > blocks_per_sm && dyn >= most)) break;
                dyn += nb > blocks_per_sm ? 1024 : -1024;
                dyn = std::max(1024, std::min(most, dyn));
            }
            dyn_for[dev][blocks_per_sm] = dyn;
        }
        dyn_smem = (size_t)dyn_for[dev][blocks_per_sm];
    }
    if (c->pdl_next) {
        // batched steps: the grid may be scheduled while the previous step's grid drains (the kernel
        // waits on griddepcontrol.wait before reading anything the previous step wrote)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(T, 1, 1);
        cfg.dynamicSmemBytes = dyn_smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MAPC_CUDA(cudaLaunchKernelEx(&cfg, kernel, args));
    } else {
        kernel<<<grid, T, dyn_smem, stream>>>(args);
    }
    MAPC_CUDA(cudaGetLastError());
    ++c->launches;
    return MAPC_OK;
}

// Source staging: LDG/STS with register prefetch and a uniform-address LDS.128 broadcast per source by
// default.  Two bit-identical alternatives exist for the A/B (fused, non-peer, 11-op kernel only):
//   kStageTma   MAPC_TMA=1   1-D TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) fill the stages
//                            of the 256-body-stage shapes: a wash, +0.7 % unfused in tools/ubench, -0.8 %
//                            for the fused kernel (24.29 vs 24.09 ms at N = 262,144);
//   kStageShfl  MAPC_SHFL=1  warp-shuffle broadcast of the staged bodies instead of the LDS broadcast.
enum Staging { kStageDefault = 0, kStageTma = 1, kStageShfl = 2 };

template <bool FUSE, bool PEER = false, bool INLOOP = false>
mapc_status launch_force_shape(mapc_compute *c, const Plan &pl, const mapc::StepArgs &args, cudaStream_t stream,
                               Staging staging, int variant = 0, int blocks_per_sm = 0)
{
    if (args.segs.count == 0 || args.i_cnt <= 0) return MAPC_OK;
    if (FUSE && !PEER && !INLOOP && staging == kStageDefault && variant != 0) {
#define MAPC_SHAPE(P, T, TJ, U, MINB, ORDER, HAS_TMA)
#define MAPC_SHAPE_VARIANT(V, P, T, TJ, U, MINB, ORDER)                                         \
        if (variant == V && pl.pairs == P && pl.threads == T)                                   \
            return launch_force<P, T, TJ, U, MINB, ORDER, FUSE && !PEER && !INLOOP, false, false, false>(c, args, stream, blocks_per_sm);
#include "force_shapes.inc"
#undef MAPC_SHAPE_VARIANT
#undef MAPC_SHAPE
    }
    constexpr bool kAlt = FUSE && !PEER && !INLOOP;   // the alternatives are instantiated for this path only
    const bool tma = kAlt && staging == kStageTma;
    const bool shfl = kAlt && staging == kStageShfl;
#define MAPC_SHAPE(P, T, TJ, U, MINB, ORDER, HAS_TMA)                                                          \
    if (pl.pairs == P && pl.threads == T) {                                                                   \
        if (HAS_TMA && tma) return launch_force<P, T, TJ, U, MINB, ORDER, FUSE, PEER, INLOOP, kAlt && HAS_TMA>(c, args, stream); \
        if (shfl) return launch_force<P, T, TJ, U, MINB, ORDER, FUSE, PEER, INLOOP, false, kAlt>(c, args, stream); \
        return launch_force<P, T, TJ, U, MINB, ORDER, FUSE, PEER, INLOOP, false>(c, args, stream, blocks_per_sm); \
    }
    // unroll / order / stage / blocks-per-SM per shape: the winners of sweeps of the FUSED kernel inside the library
    // (csrc/force_shapes.inc, profiles/r02_shape_variants.txt) -- an unfused tools/ubench ranking does not carry over
#include "force_shapes.inc"
#undef MAPC_SHAPE
    return fail(MAPC_ERR_INVALID_ARGUMENT, "no kernel for launch shape P=%d T=%d", pl.pairs, pl.threads);
}

// targets of this shard that a Simulate(n_active) updates, as a count from i_first
int local_targets(const mapc_compute *c, int n_active)
{
    return mapc::local_targets(c->n, c->i_first, c->n_local, n_active);
}

// partials scratch and (for the ring) its per-slot release counters, grown on demand
mapc_status ensure_partial(mapc_compute *c, const mapc::Scratch &sc)
{
    if (!c->partial || c->partial_bytes < sc.bytes) {
        if (c->partial) {
            MAPC_CUDA(cudaStreamSynchronize(c->compute));
            MAPC_CUDA(cudaStreamSynchronize(c->compute2));
            MAPC_CUDA(cudaFree(c->partial));
            c->partial = nullptr;
        }
        MAPC_CUDA(cudaMalloc(&c->partial, sc.bytes));
        c->partial_bytes = sc.bytes;
    }
    if (sc.ring && c->slot_gen_count < sc.slots) {
        if (c->slot_gen) {
            MAPC_CUDA(cudaStreamSynchronize(c->compute));
            MAPC_CUDA(cudaFree(c->slot_gen));
            c->slot_gen = nullptr;
        }
        MAPC_CUDA(cudaMalloc(&c->slot_gen, (size_t)sc.slots * sizeof(unsigned)));
        MAPC_CUDA(cudaMemsetAsync(c->slot_gen, 0, (size_t)sc.slots * sizeof(unsigned), c->compute));
        c->slot_gen_count = sc.slots;
    }
    return MAPC_OK;
}

mapc_status create_common(mapc_compute **out, uint32_t n, int device, int rank, int world,
                          const void *nccl_id)
{
    if (!out) return fail(MAPC_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (n == 0) return fail(MAPC_ERR_INVALID_ARGUMENT, "num_particles must be > 0");
    if (n > (1u << 28)) return fail(MAPC_ERR_INVALID_ARGUMENT, "num_particles too large");
    if (world < 1 || rank < 0 || rank >= world)
        return fail(MAPC_ERR_INVALID_ARGUMENT, "bad rank %d / world %d", rank, world);
    if (n % (uint32_t)world != 0)
        return fail(MAPC_ERR_INVALID_ARGUMENT, "num_particles %u not divisible by world %d", n, world);
    int count = 0;
    MAPC_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0) return fail(MAPC_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= count)
        return fail(MAPC_ERR_INVALID_ARGUMENT, "device %d out of range (0..%d)", device, count - 1);
    MAPC_CUDA(cudaSetDevice(device));
    MAPC_TRY(load_stream_memops());

    mapc_compute *c = new (std::nothrow) mapc_compute();
    if (!c) return fail(MAPC_ERR_OUT_OF_MEMORY, "host allocation failed");
    c->n = n;
    c->device = device;
    c->rank = rank;
    c->world = world;
    c->n_local = n / (uint32_t)world;
    c->i_first = c->n_local * (uint32_t)rank;
    mapc_status st = MAPC_OK;
    auto body = [&]() -> mapc_status {
        MAPC_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
        MAPC_CUDA(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
        c->gcompute.stream = c->compute;
        c->gcompute.device = device;
        MAPC_CUDA(cudaStreamCreateWithFlags(&c->comm, cudaStreamNonBlocking));
        MAPC_CUDA(cudaStreamCreateWithFlags(&c->compute2, cudaStreamNonBlocking));
        MAPC_CUDA(cudaEventCreate(&c->ev_step_begin));   // timing-enabled: mapc_compute_exchange_times
        MAPC_CUDA(cudaEventCreateWithFlags(&c->ev_remote_done, cudaEventDisableTiming));
        for (int s = 0; s < 2; ++s) {
            MAPC_CUDA(cudaMalloc(&c->posvelo[s], (size_t)c->n_local * sizeof(mapc_posvelo)));
            MAPC_CUDA(cudaMalloc(&c->packed[s], (size_t)n * sizeof(float4)));
            MAPC_CUDA(cudaEventCreate(&c->ev_gathered[s]));
            MAPC_CUDA(cudaEventCreate(&c->ev_gather_begin[s]));
        }
        MAPC_CUDA(cudaEventCreateWithFlags(&c->ev_integrated, cudaEventDisableTiming));
        for (int k = 0; k < mapc_compute::kTimerSlots; ++k) {
            MAPC_CUDA(cudaEventCreate(&c->t_begin[k]));
            MAPC_CUDA(cudaEventCreate(&c->t_end[k]));
        }
        MAPC_CUDA(cudaHostAlloc((void **)&c->stamps, 2 * mapc_compute::kTimerSlots * sizeof(unsigned long long),
                                cudaHostAllocPortable | cudaHostAllocMapped));
        memset(c->stamps, 0, 2 * mapc_compute::kTimerSlots * sizeof(unsigned long long));
        MAPC_CUDA(cudaMalloc(&c->done, 64));
        MAPC_CUDA(cudaMemset(c->done, 0, 64));
        MAPC_CUDA(cudaMalloc(&c->block_step, ((size_t)c->n_local / 64 + 2) * sizeof(unsigned)));
        MAPC_CUDA(cudaMemset(c->block_step, 0, ((size_t)c->n_local / 64 + 2) * sizeof(unsigned)));
        MAPC_CUDA(cudaHostAlloc((void **)&c->error_word, 2 * sizeof(unsigned long long),
                                cudaHostAllocPortable | cudaHostAllocMapped));
        c->error_word[0] = c->error_word[1] = 0;
        MAPC_CUDA(cudaMalloc(&c->counters, ((size_t)c->n_local / 64 + 2) * sizeof(unsigned)));
        MAPC_CUDA(cudaMemset(c->counters, 0, ((size_t)c->n_local / 64 + 2) * sizeof(unsigned)));
        // Compute.cpp:434-436: fence created with value 0, m_fenceValue++ -> 1
        MAPC_TRY(mapc_fence_create(&c->fence, c->fence_value));
        c->fence_value++;
        if (world > 1) {
            if (!nccl_id) return fail(MAPC_ERR_INVALID_ARGUMENT, "nccl_unique_id is NULL");
            MAPC_TRY(load_nccl());
            ncclUniqueId id;
            static_assert(sizeof(id) == MAPC_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
            memcpy(&id, nccl_id, sizeof(id));
            MAPC_NCCL(g_nccl.CommInitRank(&c->nccl, world, id, rank));
        }
        return MAPC_OK;
    };
    st = body();
    if (st != MAPC_OK) {
        const std::string keep = g_last_error;
        mapc_compute_destroy(c);
        g_last_error = keep;
        return st;
    }
    *out = c;
    return MAPC_OK;
}

}  // namespace

static void consumer_orphan(struct mapc_consumer *r);   // defined with the consumer, below

// =================================================================================================
extern "C" {

const char *mapc_last_error(void) { return g_last_error.c_str(); }
const char *mapc_version(void) { return "mapc 0.1 (sm_100a)"; }

mapc_status mapc_device_count(int *count)
{
    if (!count) return fail(MAPC_ERR_INVALID_ARGUMENT, "count is NULL");
    *count = 0;
    MAPC_CUDA(cudaGetDeviceCount(count));
    return MAPC_OK;
}

// Canonical summation order (frozen): 32 segments for every N -- a multiple of every supported GPU count, so
// no segment straddles two shards -- and, inside a segment, sequential fp32 chains of MAPC_CHAIN_SOURCES
// sources whose sums are folded left to right into the segment's partial.  A chain's length sets the
// rounding noise of the sum (once a close neighbour has made the accumulator large, every later term of the
// chain is rounded at that magnitude): measured over all targets at N = 262,144, two correctly rounded CPU
// evaluations of the same formula differ by 1.07e-5 with 32,768-term chains, 4.6e-6 with 8,192-term chains
// and 2.5e-6 with these 2,048-term chains -- and by no more at N = 4,194,304.
int mapc_plan_segments(uint32_t n_sources)
{
    (void)n_sources;
    return MAPC_MAX_SEGMENTS;
}

int mapc_plan_chain_sources(void) { return MAPC_CHAIN_SOURCES; }

// ---- fences ------------------------------------------------------------------------------------
mapc_status mapc_fence_create(mapc_fence **out, uint64_t initial_value)
{
    if (!out) return fail(MAPC_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    void *p = nullptr;
    MAPC_CUDA(cudaHostAlloc(&p, sizeof(uint64_t), cudaHostAllocPortable | cudaHostAllocMapped));
    mapc_fence *f = new (std::nothrow) mapc_fence();
    if (!f) {
        cudaFreeHost(p);
        return fail(MAPC_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
    f->word = (volatile uint64_t *)p;
    *f->word = initial_value;
    f->submitted = initial_value;
    *out = f;
    return MAPC_OK;
}

mapc_status mapc_fence_destroy(mapc_fence *f)
{
    if (!f) return MAPC_OK;
    // nothing may stay gated on a fence that is going away: treat its waits as satisfied, and strip what
    // is queued behind OTHER gates of its references to this fence
    f->submitted = UINT64_MAX;
    mapc::fence_notify(f);
    mapc::fence_scrub(f);
    for (mapc::FenceSignal &sig : f->signals) cudaEventDestroy(sig.event);
    for (cudaEvent_t ev : f->spare) cudaEventDestroy(ev);
    if (f->word) cudaFreeHost((void *)f->word);
    delete f;
    return MAPC_OK;
}

uint64_t mapc_fence_completed_value(const mapc_fence *f) { return f ? mapc::fence_completed(f) : 0; }

mapc_status mapc_fence_signal_host(mapc_fence *f, uint64_t value)
{
    if (!f) return fail(MAPC_ERR_INVALID_ARGUMENT, "fence is NULL");
    __atomic_store_n((uint64_t *)f->word, value, __ATOMIC_RELEASE);
    if (value > f->submitted) f->submitted = value;
    mapc::fence_notify(f);
    return MAPC_OK;
}

mapc_status mapc_fence_wait_host(const mapc_fence *f, uint64_t value, int timeout_ms)
{
    if (!f) return fail(MAPC_ERR_INVALID_ARGUMENT, "fence is NULL");
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (mapc_fence_completed_value(f) < value) {
        if (++spins > 64) sched_yield();
        if (timeout_ms >= 0 && (spins & 0xff) == 0) {
            const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(
                                std::chrono::steady_clock::now() - t0).count();
            if (ms > timeout_ms)
                return fail(MAPC_ERR_TIMEOUT, "fence wait for %llu timed out at %llu (submitted %llu)",
                            (unsigned long long)value, (unsigned long long)mapc_fence_completed_value(f),
                            (unsigned long long)f->submitted);
        }
    }
    return MAPC_OK;
}

// Raw-stream forms for callers that own their streams.  A raw stream cannot be gated, so waiting for a
// value whose signal has not been submitted is refused instead of risking the hang described in fence.hpp.
mapc_status mapc_fence_signal_stream(mapc_fence *f, void *cuda_stream, uint64_t value)
{
    if (!f) return fail(MAPC_ERR_INVALID_ARGUMENT, "fence is NULL");
    int device = 0;
    MAPC_CUDA(cudaStreamGetDevice((cudaStream_t)cuda_stream, &device));
    return mapc::fence_submit_signal(f, (cudaStream_t)cuda_stream, device, value);
}

mapc_status mapc_fence_wait_stream(const mapc_fence *f, void *cuda_stream, uint64_t value)
{
    if (!f) return fail(MAPC_ERR_INVALID_ARGUMENT, "fence is NULL");
    if (!mapc::fence_ready(f, value))
        return fail(MAPC_ERR_UNSUPPORTED,
                    "fence value %llu has not been signalled or submitted yet (submitted %llu): a raw stream "
                    "cannot wait for future values", (unsigned long long)value, (unsigned long long)f->submitted);
    int device = 0;
    MAPC_CUDA(cudaStreamGetDevice((cudaStream_t)cuda_stream, &device));
    DeviceGuard g(device);
    return mapc::fence_submit_wait(const_cast<mapc_fence *>(f), (cudaStream_t)cuda_stream, value);
}

// ---- create / destroy ------------------------------------------------------------------------
mapc_status mapc_compute_create(mapc_compute **out, uint32_t num_particles, int device,
                                mapc_compute *prev)
{
    MAPC_TRY(create_common(out, num_particles, device, 0, 1, nullptr));
    if (prev) {
        const mapc_status st = mapc_compute_copy_state(*out, prev);  // Compute.cpp:88-91
        if (st != MAPC_OK) {
            const std::string keep = g_last_error;
            mapc_compute_destroy(*out);
            *out = nullptr;
            g_last_error = keep;
            return st;
        }
    }
    return mapc_compute_wait_for_gpu(*out);  // Compute.cpp:97
}

mapc_status mapc_nccl_unique_id(void *out_id)
{
    if (!out_id) return fail(MAPC_ERR_INVALID_ARGUMENT, "out_id is NULL");
    MAPC_TRY(load_nccl());
    ncclUniqueId id;
    MAPC_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out_id, &id, sizeof(id));
    return MAPC_OK;
}

mapc_status mapc_compute_create_sharded(mapc_compute **out, uint32_t num_particles, int device,
                                        int rank, int world, const void *nccl_unique_id)
{
    MAPC_TRY(create_common(out, num_particles, device, rank, world, nccl_unique_id));
    return mapc_compute_wait_for_gpu(*out);
}

mapc_status mapc_compute_destroy(mapc_compute *c)
{
    if (!c) return MAPC_OK;
    DeviceGuard g(c->device);
    if (c->consumer) consumer_orphan(c->consumer);   // it must not touch this producer's buffers or fence again
    // Compute::~Compute drains the queue first (Compute.cpp:104)
    mapc::gs_detach(&c->gcompute);  // host-queued work gated on a signal that never came is dropped
    if (c->compute) cudaStreamSynchronize(c->compute);
    if (c->compute2) cudaStreamSynchronize(c->compute2);
    if (c->comm) cudaStreamSynchronize(c->comm);
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
    if (c->peer_mode)
        for (int p = 0; p < c->world && p < 16; ++p)
            if (p != c->rank) {
                for (int sd = 0; sd < 2; ++sd)
                    if (c->peer_packed[p][sd]) cudaIpcCloseMemHandle(c->peer_packed[p][sd]);
                if (c->peer_flag[p]) cudaIpcCloseMemHandle(c->peer_flag[p]);
            }
    if (c->flag) cudaFree(c->flag);
    for (int s = 0; s < 2; ++s) {
        if (c->posvelo[s]) cudaFree(c->posvelo[s]);
        if (c->packed[s]) cudaFree(c->packed[s]);
        if (c->ev_gathered[s]) cudaEventDestroy(c->ev_gathered[s]);
        if (c->ev_gather_begin[s]) cudaEventDestroy(c->ev_gather_begin[s]);
    }
    if (c->partial) cudaFree(c->partial);
    if (c->slot_gen) cudaFree(c->slot_gen);
    if (c->counters) cudaFree(c->counters);
    if (c->done) cudaFree(c->done);
    if (c->block_step) cudaFree(c->block_step);
    if (c->stamps) cudaFreeHost(c->stamps);
    if (c->error_word) cudaFreeHost(c->error_word);
    if (c->ev_integrated) cudaEventDestroy(c->ev_integrated);
    for (int k = 0; k < mapc_compute::kTimerSlots; ++k) {
        if (c->t_begin[k]) cudaEventDestroy(c->t_begin[k]);
        if (c->t_end[k]) cudaEventDestroy(c->t_end[k]);
    }
    if (c->ev_step_begin) cudaEventDestroy(c->ev_step_begin);
    if (c->ev_remote_done) cudaEventDestroy(c->ev_remote_done);
    if (c->compute) cudaStreamDestroy(c->compute);
    if (c->compute2) cudaStreamDestroy(c->compute2);
    if (c->comm) cudaStreamDestroy(c->comm);
    mapc_fence_destroy(c->fence);
    delete c;
    return MAPC_OK;
}

// ---- state in / out ---------------------------------------------------------------------------
mapc_status mapc_compute_upload(mapc_compute *c, const mapc_posvelo *host, uint32_t n)
{
    NvtxRange range("mapc: Upload");
    if (!c || !host) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n != c->n) return fail(MAPC_ERR_INVALID_ARGUMENT, "upload of %u bodies into a handle of %u", n, c->n);
    MAPC_TRY(mapc::require_ungated(&c->gcompute, "Upload"));
    DeviceGuard g(c->device);
    if (c->world == 1) {
        // side 0 is the landing buffer; side 1 and both packed mirrors fan out on the device
        MAPC_CUDA(cudaMemcpyAsync(c->posvelo[0], host, (size_t)n * sizeof(mapc_posvelo), cudaMemcpyHostToDevice, c->compute));
        MAPC_CUDA(cudaMemcpyAsync(c->posvelo[1], c->posvelo[0], (size_t)n * sizeof(mapc_posvelo),
                                  cudaMemcpyDeviceToDevice, c->compute));
        for (int s = 0; s < 2; ++s) {
            mapc::pack_positions_kernel<<<(n + 255) / 256, 256, 0, c->compute>>>(c->posvelo[0], c->packed[s], (int)n);
            MAPC_CUDA(cudaGetLastError());
            ++c->launches;
        }
    } else {
        // Sharded: a rank reads ONLY its own shard from `host` (n_local x 32 B of H2D instead of N x 32 B),
        // packs its positions and the ranks all-gather the N x 16 B position array on the device.  Collective:
        // every rank of the communicator must call Upload.
        MAPC_CUDA(cudaStreamSynchronize(c->comm));
        MAPC_CUDA(cudaMemcpyAsync(c->posvelo[0], host + c->i_first, (size_t)c->n_local * sizeof(mapc_posvelo),
                                  cudaMemcpyHostToDevice, c->compute));
        MAPC_CUDA(cudaMemcpyAsync(c->posvelo[1], c->posvelo[0], (size_t)c->n_local * sizeof(mapc_posvelo),
                                  cudaMemcpyDeviceToDevice, c->compute));
        mapc::pack_positions_kernel<<<(c->n_local + 255) / 256, 256, 0, c->compute>>>(
            c->posvelo[0], c->packed[0] + c->i_first, (int)c->n_local);
        MAPC_CUDA(cudaGetLastError());
        ++c->launches;
        MAPC_CUDA(cudaEventRecord(c->ev_integrated, c->compute));
        MAPC_CUDA(cudaStreamWaitEvent(c->comm, c->ev_integrated, 0));
        MAPC_NCCL(g_nccl.AllGather(c->packed[0] + c->i_first, c->packed[0], (size_t)c->n_local * 4, ncclFloat,
                                   c->nccl, c->comm));
        MAPC_CUDA(cudaEventRecord(c->ev_gathered[0], c->comm));
        MAPC_CUDA(cudaStreamWaitEvent(c->compute, c->ev_gathered[0], 0));
        MAPC_CUDA(cudaMemcpyAsync(c->packed[1], c->packed[0], (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice,
                                  c->compute));
    }
    MAPC_CUDA(cudaStreamSynchronize(c->comm));
    c->gather_pending[0] = c->gather_pending[1] = false;
    MAPC_TRY(mapc_compute_wait_for_gpu(c));  // InitializeParticles ends with WaitForGpu, Compute.cpp:922
    c->has_state = true;
    c->chain_valid = false;
    return MAPC_OK;
}

mapc_status mapc_compute_download(mapc_compute *c, mapc_posvelo *host, uint32_t first, uint32_t count)
{
    NvtxRange range("mapc: Download");
    if (!c || !host) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (first < c->i_first || (uint64_t)first + count > (uint64_t)c->i_first + c->n_local)
        return fail(MAPC_ERR_INVALID_ARGUMENT, "range [%u, %u) outside shard [%u, %u)", first,
                    first + count, c->i_first, c->i_first + c->n_local);
    MAPC_TRY(mapc::require_ungated(&c->gcompute, "Download"));
    DeviceGuard g(c->device);
    const uint32_t side = 1u - c->buffer_index;  // the side the last Simulate wrote
    MAPC_CUDA(cudaMemcpyAsync(host, c->posvelo[side] + (first - c->i_first),
                              (size_t)count * sizeof(mapc_posvelo), cudaMemcpyDeviceToHost, c->compute));
    MAPC_CUDA(cudaStreamSynchronize(c->compute));
    c->chain_valid = false;
    return MAPC_OK;
}

mapc_status mapc_compute_shard(const mapc_compute *c, uint32_t *first, uint32_t *count)
{
    if (!c) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    if (first) *first = c->i_first;
    if (count) *count = c->n_local;
    return MAPC_OK;
}

mapc_status mapc_compute_set_force_mode(mapc_compute *c, mapc_force_mode mode)
{
    if (!c) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    if (mode != MAPC_FORCE_ALLPAIRS && mode != MAPC_FORCE_WELL)
        return fail(MAPC_ERR_INVALID_ARGUMENT, "unknown force mode %d", (int)mode);
    c->mode = mode;
    return MAPC_OK;
}

// ---- the step ---------------------------------------------------------------------------------
// Enqueues one step on the compute (and comm) stream.  Runs either straight away or, when the
// compute stream is gated on a consumer-fence value nobody has submitted yet, when that gate opens.
static mapc_status enqueue_one(mapc_compute *c, uint32_t b, int n_targets, int n_sources, float delta_time,
                               float damping, mapc_force_mode mode, const Switches &sw);

// `steps` consecutive steps (ping-pong starting with write side b0) inside ONE timer pair.
static mapc_status enqueue_steps(mapc_compute *c, uint32_t b0, int n_targets, int n_sources, float delta_time,
                                 float damping, mapc_force_mode mode, int steps, uint64_t fence_value_after,
                                 bool kernel_signals, const Switches &sw)
{
    const bool timers = sw.timers;
    resolve_timers(c, false);
    const int slot = (int)(c->t_next % mapc_compute::kTimerSlots);
    if (c->t_pending[slot]) resolve_timers(c, true);
    // Two timing-enabled event records cost ~8 us of stream time per step on B200 (measured at
    // N = 10,000), so the fused all-pairs kernel stamps %globaltimer itself; the event pair remains for
    // the well kernel and the unfused path.
    const bool stamped = timers && mode == MAPC_FORCE_ALLPAIRS && n_targets > 0 && sw.fuse && !sw.timer_events;
    c->t_cur_slot = slot;
    c->t_begin_is_prev_end[slot] = false;
    if (stamped) {
        c->stamps[2 * slot] = 0;
        c->stamps[2 * slot + 1] = 0;
    } else if (timers) {
        MAPC_CUDA(cudaEventRecord(c->t_begin[slot], c->compute));  // BeginTimer, Compute.cpp:1020
    }
    // Programmatic dependent launch for every unsharded fused step, not only inside a batch: since such a
    // Simulate puts nothing but its kernel on the stream (timer stamps and fence signal are written by the
    // kernel), back-to-back Simulate calls chain kernel to kernel as well.  After any other kind of stream
    // operation the attribute is harmless (ordinary stream order applies).
    const bool pdl = c->world == 1 && mode == MAPC_FORCE_ALLPAIRS && sw.fuse && sw.pdl;
    for (int k = 0; k < steps; ++k) {
        c->pdl_next = pdl;
        c->stamp_begin_next = (stamped && k == 0) ? &c->stamps[2 * slot] : nullptr;
        c->stamp_end_next = (stamped && k == steps - 1) ? &c->stamps[2 * slot + 1] : nullptr;
        c->fence_write_next = (kernel_signals && k == steps - 1) ? fence_value_after : 0;
        const mapc_status st = enqueue_one(c, (b0 + (uint32_t)k) & 1u, n_targets, n_sources, delta_time, damping, mode, sw);
        c->pdl_next = false;
        c->stamp_begin_next = c->stamp_end_next = nullptr;
        c->fence_write_next = 0;
        if (st != MAPC_OK) return st;
    }
    if (!timers) return MAPC_OK;
    if (!stamped) MAPC_CUDA(cudaEventRecord(c->t_end[slot], c->compute));  // EndTimer, Compute.cpp:1046
    c->t_pending[slot] = true;
    c->t_stamped[slot] = stamped;
    c->t_fence_value[slot] = fence_value_after;
    c->t_steps[slot] = steps;
    ++c->t_next;
    return MAPC_OK;
}

static mapc_status enqueue_one(mapc_compute *c, uint32_t b, int n_targets, int n_sources, float delta_time,
                               float damping, mapc_force_mode mode, const Switches &sw)
{
    const uint32_t r = 1u - b;  // read side (reads 1-b, writes b)
    bool use_peer = false;
    const bool chain_was_valid = c->chain_valid;
    c->chain_valid = false;     // set again below by a step that publishes block_step for exactly this state

    if (c->peer_mode && mode == MAPC_FORCE_WELL)
        // a well step would overwrite this rank's packed positions without waiting for the peers that may
        // still be reading them over NVLink (only the fused all-pairs step carries that wait, in its cells)
        return fail(MAPC_ERR_UNSUPPORTED, "MAPC_FORCE_WELL steps are not available while the peer exchange is attached");
    if (n_targets > 0) {
        if (mode == MAPC_FORCE_WELL) {
            const int per_block = mapc::kWellThreads * mapc::kWellBodies;
            mapc::well_step_kernel<<<(n_targets + per_block - 1) / per_block, mapc::kWellThreads, 0, c->compute>>>(
                c->posvelo[r], c->posvelo[b], c->packed[b], (int)c->i_first, n_targets, delta_time, damping);
            MAPC_CUDA(cudaGetLastError());
            ++c->launches;
        } else {
            const Plan pl = make_plan(n_targets, n_sources, c->sm_count, sw);
            if (pl.blocks_x <= 0)
                return fail(MAPC_ERR_INVALID_ARGUMENT, "MAPC_PLAN_PAIRS=%d MAPC_PLAN_THREADS=%d name no launch shape "
                            "(csrc/force_shapes.inc)", sw.plan_pairs, sw.plan_threads);
            const bool fuse = sw.fuse;
            // scratch ring (stays in L2) for unsharded fused steps; everything else one slot per target block
            const mapc::Scratch sc = mapc::plan_scratch(pl, c->sm_count, c->world == 1 && fuse && sw.ring);
            MAPC_TRY(ensure_partial(c, sc));
            const int key = pl.pairs * 1024 + pl.threads;
            if (fuse && c->counters_key != key) {  // arrival counters are per target block of this shape
                MAPC_CUDA(cudaMemsetAsync(c->counters, 0, ((size_t)c->n_local / 64 + 2) * sizeof(unsigned), c->compute));
                c->counters_key = key;
            }
            // segments whose sources all live in this shard are resident already (this rank wrote
            // them in its own integrate pass); the others need the all-gather of side r.
            mapc::StepArgs args{};
            args.pos = c->packed[r];
            args.partial = c->partial;
            args.scratch_blocks = sc.slots;
            // grouped cell order (MAPC_RING_GROUP=0: target block major): half the ring per group, so that a cell only
            // ever waits for a target block two groups back
            args.group_blocks = (sc.ring && sw.ring_group && sc.slots >= 4) ? sc.slots / 2 : 0;
            args.ticket = sc.ring ? c->done + 1 : nullptr;
            args.slot_gen = sc.ring ? c->slot_gen : nullptr;
            args.i_first = (int)c->i_first;
            args.i_cnt = n_targets;
            args.n_sources = n_sources;
            args.S = pl.segments;
            args.n_iblocks = pl.blocks_x;
            args.counters = c->counters;
            args.in = c->posvelo[r];
            args.out = c->posvelo[b];
            args.pos_next = c->packed[b];
            args.dt = delta_time;
            args.damping = damping;
            args.done = c->done;
            args.error_word = c->error_word;
            args.wait_timeout_ns = (unsigned long long)sw.wait_timeout_ms * 1000000ull;
            // Step-to-step dataflow: an unsharded fused step publishes, per target block, that the block is
            // integrated; the NEXT such step with the same shape and counts, launched with programmatic
            // dependent launch and default staging, then waits per cell for the blocks it reads instead of for
            // the whole previous grid -- its cells run in the drain of the previous step (small N: the drain and
            // ramp of a ~50 us grid are a quarter of the step).  Not with the scratch ring (large N: cells of
            // milliseconds, nothing to gain).
            const long long chain_key[4] = {pl.pairs, pl.threads, n_targets, n_sources};
            const bool publishes = c->world == 1 && fuse;
            if (publishes) {
                args.block_step = c->block_step;
                args.step_id = ++c->step_counter;
                args.wait_prev = (sw.chain && c->pdl_next && chain_was_valid && !sc.ring && !sw.tma && !sw.shfl &&
                                  memcmp(chain_key, c->chain_key, sizeof(chain_key)) == 0) ? 1 : 0;
            }
            args.stamp_begin = c->stamp_begin_next;   // consumed by the first launch of the step
            if (args.wait_prev && args.stamp_begin != nullptr) {   // a chained step's time runs from the previous end stamp
                args.stamp_begin = nullptr;
                c->t_begin_is_prev_end[c->t_cur_slot] = true;
            }
            args.stamp_end = c->stamp_end_next;       // every launch: whichever finishes the step writes it
            args.fence_word = c->fence_write_next ? (unsigned long long *)c->fence->word : nullptr;
            args.fence_value = c->fence_write_next;
            const mapc::StepLayout lay = mapc::classify_segments(n_sources, pl.segments, (int)c->i_first, (int)c->n_local,
                                                                 c->rank, c->world, c->peer_mode, c->gather_pending[r]);
            const mapc::SegList &local = lay.local, &remote = lay.remote;
            const int *owner = lay.owner;
            const bool peer = c->peer_mode && fuse && n_sources == (int)c->n && sw.peer && !sw.mass_in_loop &&
                              lay.aligned;
            if (c->peer_mode && !peer && remote.count > 0 && !c->gather_pending[r]) {
                // exchange falls back to NCCL for this step, but the read side was never gathered
                return fail(MAPC_ERR_UNSUPPORTED, "peer exchange attached but this step's segments do not align "
                            "with the shards (n_active %d, N %u, S %d, world %d)", n_sources, c->n, pl.segments, c->world);
            }
            use_peer = peer && c->world > 1;
            // Local cells go on the compute stream at once; the remote cells go on a second stream that
            // waits for the all-gather, so both grids are resident together and the block scheduler
            // balances them (a lone local launch of few, long cells would leave most SMs idle).  The
            // arrival counters make the fused combine+integrate independent of which grid finishes last.
            const bool single_grid = use_peer && sw.peer_single && remote.count > 0;
            if (remote.count > 0 && !single_grid) MAPC_CUDA(cudaEventRecord(c->ev_step_begin, c->compute));
            // MAPC_MASS_IN_LOOP=1: the shader's per-pair `mass * invDistCube` (12 lane-ops) instead of the
            // default once-per-partial scale (11): A/B switch, fused non-peer path only
            const bool inloop = fuse && sw.mass_in_loop;
            const Staging staging = sw.shfl ? kStageShfl : (sw.tma ? kStageTma : kStageDefault);
            // steps that may chain by data flow run with fewer resident blocks per SM (mapc::throttle_blocks_per_sm)
            const bool may_chain = publishes && !sc.ring && sw.chain && sw.pdl && staging == kStageDefault && !inloop;
            const int throttle = !may_chain ? 0 : (sw.blocks_per_sm >= 0 ? sw.blocks_per_sm
                                                                         : mapc::throttle_blocks_per_sm(pl, c->sm_count));
            auto launch = [&](cudaStream_t st) -> mapc_status {
                if (inloop) return launch_force_shape<true, false, true>(c, pl, args, st, staging);
                return fuse ? launch_force_shape<true>(c, pl, args, st, staging, sw.shape_variant, throttle)
                            : launch_force_shape<false>(c, pl, args, st, staging);
            };
            if (single_grid) {
                // Experimental (MAPC_PEER_SINGLE=1): one grid for the whole step.  blockIdx.y walks the local
                // segments first and the remote ones after them, and blocks are dispatched in that order, so
                // the cells that wait for a peer's step flag only become resident once the local cells are
                // under way -- instead of a second grid whose blocks may sit on SMs spinning from the start.
                // Same cells, same arithmetic, same partials: the bits cannot change.
                mapc::SegList all{0, {}};
                for (int k = 0; k < local.count; ++k) {
                    args.seg_src[all.count] = c->packed[r];
                    args.seg_flag[all.count] = nullptr;
                    all.ids[all.count++] = local.ids[k];
                }
                for (int k = 0; k < remote.count; ++k) {
                    const int o = owner[remote.ids[k]];
                    args.seg_src[all.count] = c->peer_packed[o][r];
                    args.seg_flag[all.count] = c->peer_flag[o];
                    all.ids[all.count++] = remote.ids[k];
                }
                args.flag_expect = c->step_id;
                args.segs = all;
                MAPC_TRY((launch_force_shape<true, true>(c, pl, args, c->compute, kStageDefault)));
            } else {
            args.segs = local;
            MAPC_TRY(launch(c->compute));
            if (local.count > 0) args.stamp_begin = nullptr;
            if (remote.count > 0) {
                MAPC_CUDA(cudaStreamWaitEvent(c->compute2, c->ev_step_begin, 0));
                args.segs = remote;
                if (use_peer) {
                    // no collective: remote cells read the owners' memory, gated by the owners' step flags
                    for (int k = 0; k < remote.count; ++k) {
                        const int o = owner[remote.ids[k]];
                        args.seg_src[k] = c->peer_packed[o][r];
                        args.seg_flag[k] = c->peer_flag[o];
                    }
                    args.flag_expect = c->step_id;   // owners must have completed step_id steps
                    MAPC_TRY((launch_force_shape<true, true>(c, pl, args, c->compute2, kStageDefault)));
                } else {
                    MAPC_CUDA(cudaStreamWaitEvent(c->compute2, c->ev_gathered[r], 0));
                    MAPC_TRY(launch(c->compute2));
                }
                MAPC_CUDA(cudaEventRecord(c->ev_remote_done, c->compute2));
                MAPC_CUDA(cudaStreamWaitEvent(c->compute, c->ev_remote_done, 0));
            }
            }
            if (publishes && !sc.ring) {
                c->chain_valid = true;
                memcpy(c->chain_key, chain_key, sizeof(chain_key));
            }
            if (!fuse) {
                mapc::integrate_kernel<<<(n_targets + 255) / 256, 256, 0, c->compute>>>(
                    c->posvelo[r], c->posvelo[b], c->packed[b], c->partial, pl.block_targets(), pl.segments,
                    (int)c->i_first, n_targets, delta_time, damping);
                MAPC_CUDA(cudaGetLastError());
                ++c->launches;
            }
        }
    }
    if (c->world > 1 && c->peer_mode) {
        // Publish: this rank's positions of step step_id+1 are in place (and it has finished reading
        // everybody's previous ones).  A well-mode or zero-target step publishes too, so peers never wait.
        c->step_id++;
        MAPC_TRY(load_stream_memops());
        const CUresult wr = g_write64((CUstream)c->compute, (CUdeviceptr)(uintptr_t)c->flag, c->step_id,
                                      CU_STREAM_WRITE_VALUE_DEFAULT);
        if (wr != CUDA_SUCCESS) return fail(MAPC_ERR_CUDA, "cuStreamWriteValue64(step flag) failed: %d", (int)wr);
    }
    if (c->world > 1 && !use_peer && !c->peer_mode) {
        // exchange step: every rank contributes its slice of the freshly written packed positions
        if (c->gather_pending[r]) MAPC_CUDA(cudaStreamWaitEvent(c->compute, c->ev_gathered[r], 0));
        MAPC_CUDA(cudaEventRecord(c->ev_integrated, c->compute));
        MAPC_CUDA(cudaStreamWaitEvent(c->comm, c->ev_integrated, 0));
        MAPC_CUDA(cudaEventRecord(c->ev_gather_begin[b], c->comm));
        MAPC_NCCL(g_nccl.AllGather(c->packed[b] + c->i_first, c->packed[b], (size_t)c->n_local * 4,
                                   ncclFloat, c->nccl, c->comm));
        MAPC_CUDA(cudaEventRecord(c->ev_gathered[b], c->comm));
        c->gather_pending[b] = true;
    }
    return MAPC_OK;
}

mapc_status mapc_compute_simulate(mapc_compute *c, int num_active_particles, float delta_time,
                                  float damping, uint64_t consumer_fence_value)
{
    return mapc_compute_simulate_steps(c, num_active_particles, delta_time, damping, consumer_fence_value, 1);
}

mapc_status mapc_compute_simulate_steps(mapc_compute *c, int num_active_particles, float delta_time,
                                        float damping, uint64_t consumer_fence_value, int steps)
{
    NvtxRange range("mapc: Simulate");
    if (!c) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    if (steps < 1) return fail(MAPC_ERR_INVALID_ARGUMENT, "steps must be >= 1");
    if (num_active_particles < 0 || (uint32_t)num_active_particles > c->n)
        return fail(MAPC_ERR_INVALID_ARGUMENT, "num_active_particles %d outside [0, %u]",
                    num_active_particles, c->n);
    if (!c->has_state) return fail(MAPC_ERR_INVALID_ARGUMENT, "simulate before upload/init_particles");
    DeviceGuard g(c->device);

    // Compute.cpp:1012 -- "/previous/ copy must complete before overwriting the old state"
    if (c->consumer_fence && consumer_fence_value > 0)
        MAPC_TRY(mapc::gs_wait(&c->gcompute, c->consumer_fence, consumer_fence_value - 1));

    const uint32_t b = c->buffer_index;  // write side
    const int n_targets = local_targets(c, num_active_particles);
    const int n_sources = num_active_particles;
    const mapc_force_mode mode = c->mode;
    const uint64_t fence_value_after = c->fence_value + (uint64_t)(steps - 1);
    // Unsharded fused all-pairs steps signal the fence from inside the kernel (the block that finishes the
    // step writes the word): nothing but the kernel goes on the stream.  Everything else uses the
    // event + memory-operation signal.
    const Switches sw = read_switches();
    const bool kernel_signals = c->world == 1 && mode == MAPC_FORCE_ALLPAIRS && n_targets > 0 && sw.fuse &&
                                sw.timers && !sw.timer_events && sw.kernel_fence;
    MAPC_TRY(mapc::gs_call(&c->gcompute, [=]() -> mapc_status {
        return enqueue_steps(c, b, n_targets, n_sources, delta_time, damping, mode, steps, fence_value_after,
                             kernel_signals, sw);
    }));

    // MoveToNextFrame, Compute.cpp:993-1004 (a batch consumes one fence value per step and signals the last)
    c->fence_value += (uint64_t)(steps - 1);
    if (kernel_signals) MAPC_TRY(mapc::gs_signal_light(&c->gcompute, c->fence, c->fence_value));
    else MAPC_TRY(mapc::gs_signal(&c->gcompute, c->fence, c->fence_value));
    c->fence_value++;
    if (steps & 1) c->buffer_index = 1u - c->buffer_index;
    return MAPC_OK;
}

uint64_t mapc_compute_fence_value(const mapc_compute *c) { return c ? c->fence_value : 0; }

mapc_status mapc_compute_wait_for_gpu(mapc_compute *c)
{
    if (!c) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    DeviceGuard g(c->device);
    // Compute.cpp:931-938: Signal(m_fenceValue); m_fenceValue++; host wait
    MAPC_TRY(mapc::gs_call(&c->gcompute, [c]() -> mapc_status {
        for (int s = 0; s < 2; ++s)
            if (c->gather_pending[s]) MAPC_CUDA(cudaStreamWaitEvent(c->compute, c->ev_gathered[s], 0));
        return MAPC_OK;
    }));
    const uint64_t v = c->fence_value;
    MAPC_TRY(mapc::gs_signal(&c->gcompute, c->fence, v));
    c->fence_value++;
    MAPC_TRY(mapc::require_ungated(&c->gcompute, "WaitForGpu"));
    MAPC_CUDA(cudaStreamSynchronize(c->compute));
    MAPC_CUDA(cudaStreamSynchronize(c->comm));
    if (c->nccl && g_nccl.CommGetAsyncError) {
        // a collective that failed asynchronously (a peer died, a link error) must not pass as a finished step
        ncclResult_t async = ncclSuccess;
        const ncclResult_t q = g_nccl.CommGetAsyncError(c->nccl, &async);
        if (q != ncclSuccess || (async != ncclSuccess && async != ncclInProgress))
            return fail(MAPC_ERR_NCCL, "NCCL communicator reports an asynchronous error: %s",
                        g_nccl.GetErrorString(q != ncclSuccess ? q : async));
    }
    if (c->error_word && c->error_word[0] != 0) {
        const unsigned long long kind = c->error_word[0], detail = c->error_word[1];
        c->error_word[0] = c->error_word[1] = 0;
        return fail(MAPC_ERR_TIMEOUT, kind == 1 ? "a force cell gave up waiting for a peer rank to publish step %llu: "
                    "the step's results are invalid" : "a force cell gave up waiting for scratch-ring slot of target block "
                    "%llu: the step's results are invalid", detail);
    }
    if (mapc_fence_completed_value(c->fence) < v)
        return fail(MAPC_ERR_CUDA, "fence at %llu after drain, expected >= %llu",
                    (unsigned long long)mapc_fence_completed_value(c->fence), (unsigned long long)v);
    resolve_timers(c, true);
    c->chain_valid = false;   // the device is idle: the next step has nothing to chain to (and its timer starts afresh)
    return MAPC_OK;
}

mapc_status mapc_compute_shared_handles(mapc_compute *c, mapc_fence *consumer_fence,
                                        mapc_shared_handles *out)
{
    if (!c || !out) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL argument");
    c->consumer_fence = consumer_fence;  // OpenSharedHandle(in_fenceHandle), Compute.cpp:946
    for (int s = 0; s < 2; ++s) {
        out->posvelo[s] = c->posvelo[s];
        out->packed_pos[s] = c->packed[s];
    }
    out->fence = c->fence;
    out->compute_stream = c->compute;
    out->aligned_data_size = (uint64_t)c->n_local * sizeof(mapc_posvelo);
    out->buffer_index = c->buffer_index;  // Compute.cpp:948
    out->first_particle = c->i_first;
    out->num_local = c->n_local;
    out->device = c->device;
    return MAPC_OK;
}

mapc_status mapc_compute_gpu_times(mapc_compute *c, float *ms_average, float *ms_last)
{
    if (!c) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    DeviceGuard g(c->device);
    resolve_timers(c, false);
    if (ms_average) *ms_average = c->ms_average;
    if (ms_last) *ms_last = c->ms_last;
    return MAPC_OK;
}

namespace {
struct IpcBlob {
    cudaIpcMemHandle_t packed[2];
    cudaIpcMemHandle_t flag;
    uint32_t n, rank;
};
static_assert(sizeof(IpcBlob) <= MAPC_IPC_BLOB_BYTES, "IPC blob size");
}  // namespace

mapc_status mapc_compute_ipc_export(mapc_compute *c, void *out_blob)
{
    if (!c || !out_blob) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (c->world < 2) return fail(MAPC_ERR_INVALID_ARGUMENT, "peer exchange needs a sharded handle");
    DeviceGuard g(c->device);
    if (!c->flag) {
        MAPC_CUDA(cudaMalloc(&c->flag, 256));
        MAPC_CUDA(cudaMemset(c->flag, 0, 256));
    }
    IpcBlob b;
    memset(&b, 0, sizeof(b));
    for (int sd = 0; sd < 2; ++sd) MAPC_CUDA(cudaIpcGetMemHandle(&b.packed[sd], c->packed[sd]));
    MAPC_CUDA(cudaIpcGetMemHandle(&b.flag, c->flag));
    b.n = c->n;
    b.rank = (uint32_t)c->rank;
    memset(out_blob, 0, MAPC_IPC_BLOB_BYTES);
    memcpy(out_blob, &b, sizeof(b));
    return MAPC_OK;
}

mapc_status mapc_compute_ipc_attach(mapc_compute *c, const void *blobs_all_ranks, int world)
{
    if (!c || !blobs_all_ranks) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (world != c->world || world > 16) return fail(MAPC_ERR_INVALID_ARGUMENT, "world %d does not match the handle (%d, max 16)", world, c->world);
    if (!c->flag) return fail(MAPC_ERR_INVALID_ARGUMENT, "call mapc_compute_ipc_export first");
    if (c->peer_mode) return fail(MAPC_ERR_INVALID_ARGUMENT, "already attached");
    MAPC_TRY(mapc_compute_wait_for_gpu(c));
    DeviceGuard g(c->device);
    for (int p = 0; p < world; ++p) {
        IpcBlob b;
        memcpy(&b, (const char *)blobs_all_ranks + (size_t)p * MAPC_IPC_BLOB_BYTES, sizeof(b));
        if (b.n != c->n || (int)b.rank != p) return fail(MAPC_ERR_INVALID_ARGUMENT, "blob %d is from rank %u with N=%u", p, b.rank, b.n);
        if (p == c->rank) {
            c->peer_packed[p][0] = c->packed[0];
            c->peer_packed[p][1] = c->packed[1];
            c->peer_flag[p] = c->flag;
            continue;
        }
        for (int sd = 0; sd < 2; ++sd)
            MAPC_CUDA(cudaIpcOpenMemHandle((void **)&c->peer_packed[p][sd], b.packed[sd], cudaIpcMemLazyEnablePeerAccess));
        MAPC_CUDA(cudaIpcOpenMemHandle((void **)&c->peer_flag[p], b.flag, cudaIpcMemLazyEnablePeerAccess));
    }
    c->peer_mode = true;
    c->step_id = 0;
    c->gather_pending[0] = c->gather_pending[1] = false;
    return MAPC_OK;
}

mapc_status mapc_compute_exchange_times(mapc_compute *c, float *gather_ms, float *tail_past_step_begin_ms)
{
    if (!c) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    if (gather_ms) *gather_ms = 0.f;
    if (tail_past_step_begin_ms) *tail_past_step_begin_ms = 0.f;
    if (c->world < 2 || c->peer_mode) return MAPC_OK;
    DeviceGuard g(c->device);
    // the newest gather that has completed: the one into the side the last step READ, i.e. issued by the
    // step before it; the last step's timer slot gives the begin of the step that consumed it
    if (c->t_next < 2) return MAPC_OK;
    const uint32_t side = c->buffer_index;   // last step wrote 1-buffer_index... and read buffer_index
    if (cudaEventQuery(c->ev_gathered[side]) != cudaSuccess) return MAPC_OK;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_gather_begin[side], c->ev_gathered[side]) == cudaSuccess && gather_ms)
        *gather_ms = ms;
    // ev_step_begin: recorded on the compute stream when the consuming (= last) step began
    if (cudaEventQuery(c->ev_step_begin) == cudaSuccess &&
        cudaEventElapsedTime(&ms, c->ev_step_begin, c->ev_gathered[side]) == cudaSuccess && tail_past_step_begin_ms)
        *tail_past_step_begin_ms = ms;   // > 0: the gather was still running this long into the consuming step
    return MAPC_OK;
}

mapc_status mapc_compute_step_times(mapc_compute *c, float *ms_out, int capacity, int *count)
{
    if (!c || !count || (capacity > 0 && !ms_out)) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL argument");
    DeviceGuard g(c->device);
    resolve_timers(c, false);
    int k = (int)c->step_log.size();
    if (k > capacity) k = capacity < 0 ? 0 : capacity;
    for (int i = 0; i < k; ++i) ms_out[i] = c->step_log[i];
    c->step_log.erase(c->step_log.begin(), c->step_log.begin() + k);
    *count = k;
    return MAPC_OK;
}

mapc_status mapc_compute_flush(mapc_compute *c)
{
    if (!c) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    DeviceGuard g(c->device);
    return mapc::gs_call(&c->gcompute, [c]() -> mapc_status {
        for (int s = 0; s < 2; ++s)
            if (c->gather_pending[s]) MAPC_CUDA(cudaStreamWaitEvent(c->compute, c->ev_gathered[s], 0));
        return MAPC_OK;
    });
}

mapc_status mapc_compute_copy_state(mapc_compute *dst, mapc_compute *src)
{
    NvtxRange range("mapc: CopyState");
    if (!dst || !src) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    if (dst->n != src->n || dst->i_first != src->i_first || dst->n_local != src->n_local)
        return fail(MAPC_ERR_INVALID_ARGUMENT, "copy_state between different shapes (%u/%u vs %u/%u)",
                    dst->n, dst->n_local, src->n, src->n_local);
    if (!src->has_state) return fail(MAPC_ERR_INVALID_ARGUMENT, "source has no particle state");
    MAPC_TRY(mapc_compute_wait_for_gpu(src));  // drained like Particles.cpp:467-471
    MAPC_TRY(mapc_compute_wait_for_gpu(dst));
    DeviceGuard g(dst->device);
    for (int s = 0; s < 2; ++s) {
        // positions and velocities of both sides (Compute.cpp:341-344, :395-398); the reference
        // tunnels velocities through the shared position heap (:357-382), a peer copy needs no detour
        MAPC_CUDA(cudaMemcpyPeerAsync(dst->posvelo[s], dst->device, src->posvelo[s], src->device,
                                      (size_t)src->n_local * sizeof(mapc_posvelo), dst->compute));
        MAPC_CUDA(cudaMemcpyPeerAsync(dst->packed[s], dst->device, src->packed[s], src->device,
                                      (size_t)src->n * sizeof(float4), dst->compute));
    }
    // The reference restarts at m_bufferIndex 0 (Compute.cpp:80); carrying the index instead keeps
    // the trajectory bit-exact across a migration (index 0 would re-read the older side half the time).
    dst->buffer_index = src->buffer_index;
    dst->mode = src->mode;
    dst->gather_pending[0] = dst->gather_pending[1] = false;
    dst->has_state = true;
    dst->chain_valid = false;
    return mapc_compute_wait_for_gpu(dst);  // Compute.cpp:409
}

uint64_t mapc_compute_kernel_launches(const mapc_compute *c) { return c ? c->launches : 0; }

mapc_status mapc_compute_plan(const mapc_compute *c, int num_active_particles, int *pairs_per_thread,
                              int *threads_per_block, int *num_blocks, int *segments)
{
    if (!c) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    const int n_targets = local_targets(c, num_active_particles);
    const Plan pl = make_plan(n_targets, num_active_particles, c->sm_count, read_switches());
    if (pairs_per_thread) *pairs_per_thread = pl.pairs;
    if (threads_per_block) *threads_per_block = pl.threads;
    if (num_blocks) *num_blocks = pl.blocks_x * pl.segments;
    if (segments) *segments = pl.segments;
    return MAPC_OK;
}

mapc_status mapc_fp32_peak_probe(int device, int packed, float *tflops, float *ms_out)
{
    int count = 0;
    MAPC_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(MAPC_ERR_INVALID_ARGUMENT, "device %d out of range", device);
    DeviceGuard g(device);
    int sms = 0;
    MAPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const int iters = 8192, blocks = sms * 8, threads = 256;
    float best = 1e30f;
    auto body = [&]() -> mapc_status {
        MAPC_CUDA(cudaMalloc(&sink, 64));
        MAPC_CUDA(cudaEventCreate(&e0));
        MAPC_CUDA(cudaEventCreate(&e1));
        for (int rep = 0; rep < 4; ++rep) {
            MAPC_CUDA(cudaEventRecord(e0, 0));
            if (packed) mapc::fp32_peak_kernel<true><<<blocks, threads>>>(sink, iters, 0.999f, 0.001f);
            else mapc::fp32_peak_kernel<false><<<blocks, threads>>>(sink, iters, 0.999f, 0.001f);
            MAPC_CUDA(cudaEventRecord(e1, 0));
            MAPC_CUDA(cudaEventSynchronize(e1));
            MAPC_CUDA(cudaGetLastError());
            float ms = 0.f;
            MAPC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        return MAPC_OK;
    };
    const mapc_status st = body();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (sink) cudaFree(sink);
    if (st != MAPC_OK) return st;
    // 32 lane-FMAs per thread per iteration in both variants (16 x f32x2 or 32 scalar)
    const double flops = 2.0 * 32.0 * iters * (double)blocks * threads;
    if (tflops) *tflops = (float)(flops / (best * 1e-3) / 1e12);
    if (ms_out) *ms_out = best;
    return MAPC_OK;
}

// ---- headless consumer (the Render worker's role) ----------------------------------------------
}  // extern "C"

struct mapc_consumer {
    mapc_compute *producer = nullptr;
    int device = 0;
    uint32_t n = 0;                  // bodies this consumer sees: the producer's shard (all N when unsharded)
    uint32_t first = 0;              // first body of that shard
    bool async_mode = false;         // Render::m_asyncMode: same device, the producer's buffers are read in place
    cudaStream_t copy = nullptr;     // m_copyQueue
    cudaStream_t render = nullptr;   // m_commandQueue (direct queue): consumes the local buffer
    mapc::GatedStream gcopy, grender;  // the same streams behind the fence gate (fence.hpp)
    float4 *local[2] = {nullptr, nullptr};   // m_buffers: positions local to the consumer's device
    float4 *host[2] = {nullptr, nullptr};    // pinned dump targets ("the screen")
    mapc_fence *copy_fence = nullptr;        // m_copyFence, shared with the producer
    mapc_fence *render_fence = nullptr;      // m_renderFence
    mapc_fence *compute_fence = nullptr;     // m_sharedComputeFence (borrowed)
    uint64_t copy_fence_value = 0, render_fence_value = 0;
    uint64_t frame_fence_values[2] = {0, 0};
    uint32_t shared_buffer_index = 0;        // m_sharedBufferIndex
    uint32_t current_buffer_index = 0;       // m_currentBufferIndex
    uint32_t frame_index = 0;                // m_frameIndex (two frames in flight)
    uint64_t frames_drawn = 0;
    uint64_t host_frame[2] = {0, 0};         // simulation step held by host[i]
    uint64_t host_fence[2] = {0, 0};         // render fence value that marks host[i] complete
    uint32_t host_count[2] = {0, 0};
    uint64_t local_frame[2] = {0, 0};        // simulation step held by local[i]
    uint64_t copies = 0;                     // frames handed out so far = simulation step of the next frame
    uint64_t peer_copies = 0;                // cudaMemcpyPeerAsync calls issued by Draw (none in async mode)
};

// The producer is being destroyed first: finish what the consumer still has in flight against the
// producer's buffers, then cut every reference to it.  Draw fails loudly from now on.
static void consumer_orphan(mapc_consumer *r)
{
    DeviceGuard g(r->device);
    mapc::gs_detach(&r->gcopy);
    mapc::gs_detach(&r->grender);
    if (r->copy) cudaStreamSynchronize(r->copy);
    if (r->render) cudaStreamSynchronize(r->render);
    if (r->producer) {
        if (r->producer->consumer_fence == r->copy_fence || r->producer->consumer_fence == r->render_fence)
            r->producer->consumer_fence = nullptr;
        r->producer->consumer = nullptr;
    }
    r->producer = nullptr;
    r->compute_fence = nullptr;
}

extern "C" {

mapc_status mapc_consumer_destroy(mapc_consumer *r)
{
    if (!r) return MAPC_OK;
    DeviceGuard g(r->device);
    mapc::gs_detach(&r->gcopy);
    mapc::gs_detach(&r->grender);
    if (r->copy) cudaStreamSynchronize(r->copy);
    if (r->render) cudaStreamSynchronize(r->render);
    if (r->producer) {
        if (r->producer->consumer_fence == r->copy_fence || r->producer->consumer_fence == r->render_fence)
            r->producer->consumer_fence = nullptr;
        if (r->producer->consumer == r) r->producer->consumer = nullptr;
    }
    for (int i = 0; i < 2; ++i) {
        if (r->local[i]) cudaFree(r->local[i]);
        if (r->host[i]) cudaFreeHost(r->host[i]);
    }
    if (r->copy) cudaStreamDestroy(r->copy);
    if (r->render) cudaStreamDestroy(r->render);
    mapc_fence_destroy(r->copy_fence);
    mapc_fence_destroy(r->render_fence);
    delete r;
    return MAPC_OK;
}

mapc_status mapc_consumer_wait_for_gpu(mapc_consumer *r)
{
    if (!r) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL consumer");
    DeviceGuard g(r->device);
    // Render::WaitForGpu, Render.cpp:626-647: copy fence, then render fence, then host wait
    r->copy_fence_value++;
    MAPC_TRY(mapc::gs_signal(&r->gcopy, r->copy_fence, r->copy_fence_value));
    MAPC_TRY(mapc::gs_wait(&r->grender, r->copy_fence, r->copy_fence_value));
    MAPC_TRY(mapc::gs_signal(&r->grender, r->render_fence, r->render_fence_value));
    r->render_fence_value++;
    MAPC_TRY(mapc::require_ungated(&r->gcopy, "consumer WaitForGpu (copy stream)"));
    MAPC_TRY(mapc::require_ungated(&r->grender, "consumer WaitForGpu (render stream)"));
    MAPC_CUDA(cudaStreamSynchronize(r->copy));
    MAPC_CUDA(cudaStreamSynchronize(r->render));
    return MAPC_OK;
}

mapc_status mapc_consumer_create(mapc_consumer **out, mapc_compute *producer, int device)
{
    return mapc_consumer_create_ex(out, producer, device, 0u);
}

mapc_status mapc_consumer_create_ex(mapc_consumer **out, mapc_compute *producer, int device, uint32_t flags)
{
    if (!out || !producer) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = nullptr;
    if (flags & ~(uint32_t)MAPC_CONSUMER_ASYNC) return fail(MAPC_ERR_INVALID_ARGUMENT, "unknown consumer flags 0x%x", flags);
    const bool async_mode = (flags & MAPC_CONSUMER_ASYNC) != 0;
    if (async_mode && device != producer->device)
        // Particles::ShareHandles: asyncMode = (m_renderAdapterIndex == m_computeAdapterIndex), Particles.cpp:202
        return fail(MAPC_ERR_INVALID_ARGUMENT, "an async consumer lives on its producer's device (%d), not %d",
                    producer->device, device);
    if (!producer->has_state) return fail(MAPC_ERR_INVALID_ARGUMENT, "producer has no particle state");
    if (producer->consumer) return fail(MAPC_ERR_INVALID_ARGUMENT, "producer already has a consumer attached");
    int count = 0;
    MAPC_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(MAPC_ERR_INVALID_ARGUMENT, "device %d out of range", device);
    MAPC_TRY(mapc_compute_wait_for_gpu(producer));
    DeviceGuard g(device);
    mapc_consumer *r = new (std::nothrow) mapc_consumer();
    if (!r) return fail(MAPC_ERR_OUT_OF_MEMORY, "host allocation failed");
    r->producer = producer;
    r->device = device;
    r->n = producer->n_local;          // a sharded producer hands out its own slice (each rank dumps its shard)
    r->first = producer->i_first;
    r->async_mode = async_mode;
    auto body = [&]() -> mapc_status {
        MAPC_CUDA(cudaStreamCreateWithFlags(&r->copy, cudaStreamNonBlocking));
        MAPC_CUDA(cudaStreamCreateWithFlags(&r->render, cudaStreamNonBlocking));
        r->gcopy.stream = r->copy;
        r->grender.stream = r->render;
        r->gcopy.device = r->grender.device = device;
        for (int i = 0; i < 2; ++i) {
            // async mode has no local copies: the "draw" reads the producer's packed positions in place
            if (!async_mode) MAPC_CUDA(cudaMalloc(&r->local[i], (size_t)r->n * sizeof(float4)));
            MAPC_CUDA(cudaHostAlloc(&r->host[i], (size_t)r->n * sizeof(float4), cudaHostAllocPortable));
        }
        MAPC_TRY(mapc_fence_create(&r->render_fence, 0));   // Render.cpp:588-595
        r->render_fence_value = 1;
        MAPC_TRY(mapc_fence_create(&r->copy_fence, 0));     // Render.cpp:611-617 (the shared one)
        // Particles::ShareHandles, Particles.cpp:198-200; in async mode the fence the producer waits on is the
        // RENDER fence (SetAsync(m_pRender->GetFence(), ...), Particles.cpp:205, Compute.cpp:961)
        mapc_shared_handles sh;
        MAPC_TRY(mapc_compute_shared_handles(producer, async_mode ? r->render_fence : r->copy_fence, &sh));
        r->shared_buffer_index = sh.buffer_index;           // Render.cpp:224
        r->compute_fence = sh.fence;
        // "copy initial state from the other adapter" (Render.cpp:253-): both local buffers
        const uint32_t newest = 1u - sh.buffer_index;
        const float4 *src = producer->packed[newest] + r->first;
        if (!async_mode) {
            for (int i = 0; i < 2; ++i)
                MAPC_CUDA(cudaMemcpyPeerAsync(r->local[i], device, src, producer->device, (size_t)r->n * sizeof(float4), r->copy));
            MAPC_CUDA(cudaMemcpyAsync(r->host[0], r->local[0], (size_t)r->n * sizeof(float4), cudaMemcpyDeviceToHost, r->copy));
        } else {
            MAPC_CUDA(cudaMemcpyAsync(r->host[0], src, (size_t)r->n * sizeof(float4), cudaMemcpyDeviceToHost, r->copy));
        }
        r->host_count[0] = r->n;
        return mapc_consumer_wait_for_gpu(r);
    };
    const mapc_status st = body();
    if (st != MAPC_OK) {
        const std::string keep = g_last_error;
        mapc_consumer_destroy(r);
        g_last_error = keep;
        return st;
    }
    producer->consumer = r;
    *out = r;
    return MAPC_OK;
}

// The async frame (Render::Draw with m_asyncMode, Render.cpp:849-852 and :928-932): producer and consumer share
// a device, so there is no copy queue and no local buffer -- the render stream waits for the PREVIOUS Simulate
// (compute fence F-1), consumes the side that Simulate wrote where it lies, then waits for the upcoming
// Simulate (compute fence F) before it signals the render fence; the value handed back is the RENDER fence
// value, which is the fence Compute::Simulate waits on in this mode (Compute.cpp:961, :1012).
static mapc_status consumer_draw_async(mapc_consumer *r, int n_draw, uint64_t *inout_fence_value)
{
    const uint64_t compute_fence_value = *inout_fence_value;
    MAPC_TRY(mapc::gs_wait(&r->grender, r->compute_fence, compute_fence_value - 1));        // Render.cpp:851
    const uint32_t src_side = 1u - r->shared_buffer_index;   // the side the previous Simulate wrote
    r->shared_buffer_index = 1u - r->shared_buffer_index;
    const uint32_t slot = r->frame_index;
    if (n_draw > 0) {
        float4 *dst = r->host[slot];
        const float4 *src = r->producer->packed[src_side] + r->first;
        const size_t bytes = (size_t)n_draw * sizeof(float4);
        cudaStream_t render = r->render;
        MAPC_TRY(mapc::gs_call(&r->grender, [=]() -> mapc_status {
            MAPC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, render));
            return MAPC_OK;
        }));
    }
    r->host_frame[slot] = r->copies++;       // simulation step shown (0 = the initial state)
    r->host_count[slot] = (uint32_t)n_draw;
    // the next frame must not start before THIS frame's Simulate has produced its results (:930); the value
    // belongs to the Simulate submitted after this call, so the render stream gates here (fence.hpp)
    MAPC_TRY(mapc::gs_wait(&r->grender, r->compute_fence, compute_fence_value));
    *inout_fence_value = r->render_fence_value;                                              // :931
    return MAPC_OK;
}

mapc_status mapc_consumer_draw(mapc_consumer *r, int num_active_particles, uint64_t *inout_fence_value,
                               int num_particles_copied)
{
    NvtxRange range("mapc: consumer Draw");
    if (!r || !inout_fence_value) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!r->producer) return fail(MAPC_ERR_INVALID_ARGUMENT, "the producer of this consumer has been destroyed");
    const uint32_t n_global = r->producer->n;
    if (num_active_particles < 0 || (uint32_t)num_active_particles > n_global || num_particles_copied < 0 ||
        (uint32_t)num_particles_copied > n_global)
        return fail(MAPC_ERR_INVALID_ARGUMENT, "particle counts outside [0, %u]", n_global);
    // counts are global (the reference's sliders, Particles.cpp:380-394); a shard takes its part of them
    auto in_shard = [&](int count) -> int {
        long long c = (long long)count - (long long)r->first;
        if (c < 0) c = 0;
        if (c > (long long)r->n) c = r->n;
        return (int)c;
    };
    num_active_particles = in_shard(num_active_particles);
    num_particles_copied = in_shard(num_particles_copied);
    DeviceGuard g(r->device);
    if (r->async_mode) {
        MAPC_TRY(consumer_draw_async(r, num_active_particles, inout_fence_value));
    } else {
    const uint64_t compute_fence_value = *inout_fence_value;

    // ---- CopySimulationResults(in_fenceValue, in_numParticlesCopied), Render.cpp:789-831 ---------
    // wait on the previous frame's render to finish with the local buffer (:796)
    MAPC_TRY(mapc::gs_wait(&r->gcopy, r->render_fence, r->render_fence_value - 1));
    const uint32_t src_shared = 1u - r->shared_buffer_index;   // :798
    const uint32_t dst_local = 1u - r->current_buffer_index;   // :799
    r->shared_buffer_index = 1u - r->shared_buffer_index;      // :800
    if (num_particles_copied > 0) {                            // :814 copy just the particles required
        float4 *dst = r->local[dst_local];
        const float4 *src = r->producer->packed[src_shared] + r->first;
        const int dst_dev = r->device, src_dev = r->producer->device;
        const size_t bytes = (size_t)num_particles_copied * sizeof(float4);
        cudaStream_t copy = r->copy;
        MAPC_TRY(mapc::gs_call(&r->gcopy, [=]() -> mapc_status {
            MAPC_CUDA(cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, bytes, copy));
            return MAPC_OK;
        }));
        r->peer_copies++;
    }
    r->local_frame[dst_local] = r->copies++;   // results of the PREVIOUS Simulate (step number = copies so far)
    // don't start the next copy until the compute device has produced new results (:826).  The value
    // belongs to the Simulate that is submitted AFTER this call: the copy stream gates here (fence.hpp).
    MAPC_TRY(mapc::gs_wait(&r->gcopy, r->compute_fence, compute_fence_value));
    r->copy_fence_value++;                                     // :829-830
    MAPC_TRY(mapc::gs_signal(&r->gcopy, r->copy_fence, r->copy_fence_value));

    // ---- the "draw" of local[m_currentBufferIndex] (:884-891): headless = dump to pinned host ------
    const uint32_t cur = r->current_buffer_index;
    r->current_buffer_index = 1u - r->current_buffer_index;    // :885
    const uint32_t slot = r->frame_index;
    if (num_active_particles > 0) {
        float4 *dst = r->host[slot];
        const float4 *src = r->local[cur];
        const size_t bytes = (size_t)num_active_particles * sizeof(float4);
        cudaStream_t render = r->render;
        MAPC_TRY(mapc::gs_call(&r->grender, [=]() -> mapc_status {
            MAPC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, render));
            return MAPC_OK;
        }));
    }
    r->host_frame[slot] = r->local_frame[cur];
    r->host_count[slot] = (uint32_t)num_active_particles;

    // render waits for this frame's copy; hand the copy fence value to the producer (:925-926)
    MAPC_TRY(mapc::gs_wait(&r->grender, r->copy_fence, r->copy_fence_value));
    *inout_fence_value = r->copy_fence_value;
    }

    // ---- MoveToNextFrame, Render.cpp:653-677 ---------------------------------------------------------
    const uint32_t slot = r->frame_index;
    r->frame_fence_values[r->frame_index] = r->render_fence_value;
    r->host_fence[slot] = r->render_fence_value;
    MAPC_TRY(mapc::gs_signal(&r->grender, r->render_fence, r->render_fence_value));
    r->render_fence_value++;
    r->frame_index = 1u - r->frame_index;                      // two "back buffers"
    r->frames_drawn++;
    // "If the next frame is not ready to be rendered yet, wait until it is ready" (:665-674; the
    // reference returns the event and Particles::Draw waits on it, Particles.cpp:452-456)
    const uint64_t throttle = r->frame_fence_values[r->frame_index];
    if (mapc_fence_completed_value(r->render_fence) < throttle) {
        if (r->render_fence->submitted < throttle)
            return fail(MAPC_ERR_TIMEOUT,
                        "Draw called twice without the Simulate in between: render fence value %llu waits for "
                        "a compute fence value nobody has submitted", (unsigned long long)throttle);
        MAPC_TRY(mapc_fence_wait_host(r->render_fence, throttle, 60000));
    }
    return MAPC_OK;
}

mapc_status mapc_consumer_counters(const mapc_consumer *r, uint64_t out[8])
{
    if (!r || !out) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL argument");
    out[0] = mapc_fence_completed_value(r->copy_fence);
    out[1] = r->copy_fence_value;
    out[2] = mapc_fence_completed_value(r->render_fence);
    out[3] = r->render_fence_value;
    out[4] = r->shared_buffer_index;
    out[5] = r->current_buffer_index;
    out[6] = r->frames_drawn;
    out[7] = r->peer_copies;
    return MAPC_OK;
}

mapc_status mapc_consumer_latest(mapc_consumer *r, const float **host_positions, uint64_t *frame, uint32_t *count)
{
    if (!r) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL consumer");
    // newest slot whose render fence has been reached
    int best = -1;
    const uint64_t done = mapc_fence_completed_value(r->render_fence);
    for (int i = 0; i < 2; ++i)
        if (r->host_fence[i] <= done && (best < 0 || r->host_fence[i] > r->host_fence[best])) best = i;
    if (best < 0) return fail(MAPC_ERR_INVALID_ARGUMENT, "no completed frame");
    if (host_positions) *host_positions = reinterpret_cast<const float *>(r->host[best]);
    if (frame) *frame = r->host_frame[best];
    if (count) *count = r->host_count[best];
    return MAPC_OK;
}

// ---- initial conditions (InitializeParticles / LoadParticles) -----------------------------------
// Generated ON the device (init_particles_kernel, one thread per particle): the reference fills two host
// vectors with a parallel_for and uploads them (Compute.cpp:820-923); at 4 M bodies that is 134 MB of H2D
// this path does not need.  Every rank of a sharded handle generates all N positions (it needs them as
// sources) and the PosVelo of its own shard; oracle/oracle.c mapo_init_particles restates the same bytes.
mapc_status mapc_compute_init_particles(mapc_compute *c, uint32_t seed)
{
    if (!c) return fail(MAPC_ERR_INVALID_ARGUMENT, "NULL handle");
    MAPC_TRY(mapc::require_ungated(&c->gcompute, "InitializeParticles"));
    DeviceGuard g(c->device);
    MAPC_CUDA(cudaStreamSynchronize(c->comm));
    // Compute.cpp:831-844: two groups of N/2 around x = +-0.75*ParticleSpread
    const float center = MAPC_PARTICLE_SPREAD * 0.750f;
    mapc::init_particles_kernel<<<(c->n + 255) / 256, 256, 0, c->compute>>>(
        c->posvelo[0], c->posvelo[1], c->packed[0], c->packed[1], c->n, c->i_first, c->n_local, seed, center,
        MAPC_INITIAL_PARTICLE_SPEED, MAPC_PARTICLE_SPREAD);
    MAPC_CUDA(cudaGetLastError());
    ++c->launches;
    c->gather_pending[0] = c->gather_pending[1] = false;
    MAPC_TRY(mapc_compute_wait_for_gpu(c));  // InitializeParticles ends with WaitForGpu, Compute.cpp:922
    c->has_state = true;
    c->chain_valid = false;
    return MAPC_OK;
}

}  // extern "C"
