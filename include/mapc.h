/*
 * include/mapc.h -- C ABI of the B200-native Compute component.
 *
 * Drop-in boundary for ONE path of GameTechDev/Multi-Adapter-Particles: the gravity force +
 * integration step of Particles/nBodyGravityCS.hlsl behind the `Compute` surface of
 * Particles/Compute.h / Compute.cpp.  Plain C types only (no torch / CUDA types): streams and
 * device pointers cross the boundary as void*.  Every entry point names the reference
 * interface it replaces (paths relative to the reference checkout).
 *
 * Error convention: the reference wraps every D3D call in ThrowIfFailed -> HrException
 * (dx-samples-include/DXSampleHelper.h:22-46) and never catches.  Nothing may throw across a
 * C ABI, so every call returns a mapc_status (0 = OK) and records a per-thread message that
 * mapc_last_error() returns; the C++ wrapper (include/mapc_compute.hpp) rethrows.
 *
 * Threading: as in the reference (single Win32 message-loop thread, Main-Particles.cpp:76-90)
 * a handle is not thread safe; distinct handles are independent.  mapc_compute_simulate only
 * enqueues work on CUDA streams and returns; mapc_compute_wait_for_gpu / mapc_fence_wait_host
 * are the only blocking calls.
 */
#ifndef MAPC_H
#define MAPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MAPC_API __attribute__((visibility("default")))
#else
#define MAPC_API
#endif

/* ---- constants carried over from the reference ------------------------------------------- */
#define MAPC_BLOCK_SIZE             64         /* Particles/defines.h:37  BLOCK_SIZE (j tile)   */
#define MAPC_SOFTENING_SQUARED      25.0f      /* Particles/nBodyGravityCS.hlsl:37             */
#define MAPC_PARTICLE_MASS          70000.0f   /* Particles/nBodyGravityCS.hlsl:38             */
#define MAPC_DEFAULT_DELTA_TIME     0.1f       /* Particles/Compute.cpp:545  paramf[0]         */
#define MAPC_DEFAULT_DAMPING        1.0f       /* Particles/Compute.cpp:546  paramf[1]         */
#define MAPC_INITIAL_PARTICLE_SPEED 15.0f      /* Particles/defines.h:39                       */
#define MAPC_PARTICLE_SPREAD        400.0f     /* Particles/defines.h:42                       */
#define MAPC_MIN_NUM_PARTICLES      (256 * 1024)       /* Particles/defines.h:44               */
#define MAPC_MAX_NUM_PARTICLES      (4 * 1024 * 1024)  /* Particles/defines.h:45               */
/* Canonical summation order (frozen): S = 32 source segments for every N; inside a
 * segment, sequential chains of at most 2,048 sources whose sums are folded left to right. */
#define MAPC_MAX_SEGMENTS           32
#define MAPC_CHAIN_SOURCES          2048
#define MAPC_NCCL_UNIQUE_ID_BYTES   128
#define MAPC_IPC_BLOB_BYTES         256

/* struct PosVelo { float4 pos; float4 velo; }  Particles/ParticleShared.hlsl:12-16.
 * pos[3] carries length(accel) after a step (nBodyGravityCS.hlsl:107); velo[3] is 0 (the
 * reference's Velocity is a bare float3, nBodyGravityCS.hlsl:72-75, :108). */
typedef struct mapc_posvelo {
    float pos[4];
    float velo[4];
} mapc_posvelo;

typedef enum mapc_status {
    MAPC_OK = 0,
    MAPC_ERR_INVALID_ARGUMENT = 1,
    MAPC_ERR_CUDA = 2,          /* a CUDA runtime/driver call failed; see mapc_last_error()   */
    MAPC_ERR_NCCL = 3,
    MAPC_ERR_NO_DEVICE = 4,     /* no usable CUDA device / driver: there is NO CPU fallback   */
    MAPC_ERR_UNSUPPORTED = 5,
    MAPC_ERR_TIMEOUT = 6,
    MAPC_ERR_OUT_OF_MEMORY = 7
} mapc_status;

/* Which force the step applies.  ALLPAIRS is the north-star path (bodyBodyInteraction,
 * nBodyGravityCS.hlsl:44-57, summed over all j).  WELL is the kernel the reference actually
 * dispatches: a single gravity well at the origin (nBodyGravityCS.hlsl:86-109). */
typedef enum mapc_force_mode {
    MAPC_FORCE_ALLPAIRS = 0,
    MAPC_FORCE_WELL = 1
} mapc_force_mode;

typedef struct mapc_compute mapc_compute;   /* replaces class Compute, Particles/Compute.h:33 */
typedef struct mapc_fence mapc_fence;       /* replaces ID3D12Fence shared across adapters    */

MAPC_API const char *mapc_last_error(void);
MAPC_API const char *mapc_version(void);
/* number of visible CUDA devices (IDXGIFactory::EnumAdapters1 loop, Particles.cpp:101-122) */
MAPC_API mapc_status mapc_device_count(int *count);

/* ---- fences: monotonically increasing 64-bit values (ID3D12Fence) -------------------------
 * Compute.cpp:434-435 creates a cross-adapter shared fence; Render.cpp:612-617 the consumer's.
 * Here a fence is an 8-byte word in pinned, portable, device-mapped host memory: every device
 * and the host can signal and wait on it (cuStreamWriteValue64 / cuStreamWaitValue64). */
MAPC_API mapc_status mapc_fence_create(mapc_fence **out, uint64_t initial_value);
MAPC_API mapc_status mapc_fence_destroy(mapc_fence *f);
/* ID3D12Fence::GetCompletedValue */
MAPC_API uint64_t    mapc_fence_completed_value(const mapc_fence *f);
/* ID3D12Fence::Signal (CPU side) */
MAPC_API mapc_status mapc_fence_signal_host(mapc_fence *f, uint64_t value);
/* SetEventOnCompletion + WaitForSingleObject (Compute.cpp:934-938); timeout_ms < 0 = forever */
MAPC_API mapc_status mapc_fence_wait_host(const mapc_fence *f, uint64_t value, int timeout_ms);
/* ID3D12CommandQueue::Signal / ::Wait on a CUDA stream (cudaStream_t passed as void*) */
MAPC_API mapc_status mapc_fence_signal_stream(mapc_fence *f, void *cuda_stream, uint64_t value);
MAPC_API mapc_status mapc_fence_wait_stream(const mapc_fence *f, void *cuda_stream, uint64_t value);

/* ---- Compute -------------------------------------------------------------------------------
 * Compute::Compute(UINT numParticles, IDXGIAdapter1*, bool useIntelExt, Compute* prev = 0)
 * Particles/Compute.h:36-39, Compute.cpp:72-98.  `device` replaces the adapter; `prev`, when
 * not NULL, is drained and its state copied in (CopyState, Compute.cpp:303-410).  The Intel
 * command-queue extension flag has no CUDA analogue and is dropped.  Like the reference
 * constructor, returns with the device idle.  Particle state is undefined until
 * mapc_compute_upload / mapc_compute_init_particles (InitializeParticles, Compute.cpp:820). */
MAPC_API mapc_status mapc_compute_create(mapc_compute **out, uint32_t num_particles, int device,
                                         mapc_compute *prev);

/* Multi-GPU, one process per GPU: rank `rank` of `world` owns targets
 * [rank*N/world, (rank+1)*N/world) and all-gathers the N x 16-byte packed position array over
 * NCCL each step.  `nccl_unique_id` (MAPC_NCCL_UNIQUE_ID_BYTES bytes) comes from
 * mapc_nccl_unique_id() on rank 0 and is distributed by the caller (e.g. torch.distributed).
 * Requires N % world == 0.  (No reference analogue: the reference never splits the sim.) */
MAPC_API mapc_status mapc_nccl_unique_id(void *out_id);
MAPC_API mapc_status mapc_compute_create_sharded(mapc_compute **out, uint32_t num_particles,
                                                 int device, int rank, int world,
                                                 const void *nccl_unique_id);

/* Collective-free exchange (optional, after create_sharded + upload): every rank exports the IPC
 * handles of its packed position buffers and step flag (MAPC_IPC_BLOB_BYTES bytes), the caller
 * gathers the `world` blobs in rank order, and every rank attaches them.  From then on a step reads
 * the other ranks' segments straight from their memory over NVLink inside the force kernel, gated by
 * per-rank step flags, instead of all-gathering positions with NCCL.  Same canonical order, same
 * bits.  The caller must barrier the ranks between upload/attach and the first Simulate.  Once
 * attached, all-pairs steps whose segments do not align with the shards (n_active != N, or a segment
 * straddling two shards) are refused with MAPC_ERR_UNSUPPORTED.  (Closest reference idea: both
 * adapters addressing one shared heap, Compute.cpp:165-199.) */
MAPC_API mapc_status mapc_compute_ipc_export(mapc_compute *c, void *out_blob);
MAPC_API mapc_status mapc_compute_ipc_attach(mapc_compute *c, const void *blobs_all_ranks, int world);

/* Compute::~Compute, Compute.cpp:102-123: drains the device, frees everything the handle owns */
MAPC_API mapc_status mapc_compute_destroy(mapc_compute *c);

/* Upload of initial state into BOTH sides of the ping-pong (Compute.cpp:881-882, :903-904).
 * `host` points at body 0 of all N (global indexing).  An unsharded handle reads all of it; a sharded
 * handle reads ONLY its own shard, host[first .. first+count), and obtains the other ranks' positions by
 * an all-gather on the device -- so the call is collective over the ranks, and the rest of `host` need
 * not even be valid on that rank.  Blocks until the state is in place, as InitializeParticles does
 * (Compute.cpp:922). */
MAPC_API mapc_status mapc_compute_upload(mapc_compute *c, const mapc_posvelo *host, uint32_t n);
/* Current state (the side the last Simulate wrote) of bodies [first, first+count) -> host.
 * The range must lie inside the handle's shard.  Blocking. */
MAPC_API mapc_status mapc_compute_download(mapc_compute *c, mapc_posvelo *host, uint32_t first,
                                           uint32_t count);
/* shard owned by this handle (whole range for an unsharded handle) */
MAPC_API mapc_status mapc_compute_shard(const mapc_compute *c, uint32_t *first, uint32_t *count);

MAPC_API mapc_status mapc_compute_set_force_mode(mapc_compute *c, mapc_force_mode mode);

/* void Compute::Simulate(int numActiveParticles, UINT64 sharedFenceValue)
 * Particles/Compute.h:48, Compute.cpp:1009-1055.  delta_time / damping were constants uploaded
 * once (Compute.cpp:545-546); they are per-call arguments here.  Order of operations kept:
 * wait consumer fence >= consumer_fence_value-1 (:1012; skipped when no consumer fence is
 * attached or the value is 0), read side 1-b, write side b (b = buffer index), time the step,
 * signal own fence with fence_value, fence_value++, b ^= 1 (MoveToNextFrame, :993-1004).
 * Bodies updated: i < min(N, 64*ceil(n_active/64)) (Dispatch, :1041); sources: j < n_active.
 * Asynchronous: returns after enqueueing. */
MAPC_API mapc_status mapc_compute_simulate(mapc_compute *c, int num_active_particles,
                                           float delta_time, float damping,
                                           uint64_t consumer_fence_value);

/* `steps` consecutive Simulate calls in one submission (headless extension, no reference analogue):
 * only the first step waits for the consumer fence and only the last one signals -- the fence jumps
 * by `steps`, one value per step as if they had been issued one by one -- and the GPU timer records
 * the batch average.  On an unsharded handle the force kernels of a batch are chained with
 * programmatic dependent launch, so a step's grid is scheduled while the previous one drains.
 * Same arithmetic, same bits as `steps` single calls. */
MAPC_API mapc_status mapc_compute_simulate_steps(mapc_compute *c, int num_active_particles,
                                                 float delta_time, float damping,
                                                 uint64_t consumer_fence_value, int steps);

/* UINT64 Compute::GetFenceValue() const, Compute.h:64: the value the NEXT Simulate signals */
MAPC_API uint64_t    mapc_compute_fence_value(const mapc_compute *c);
/* void Compute::WaitForGpu(), Compute.cpp:928-940: signal fence_value, fence_value++, host-wait */
MAPC_API mapc_status mapc_compute_wait_for_gpu(mapc_compute *c);

/* struct Compute::SharedHandles, Compute.h:54-61 + GetSharedHandles(HANDLE consumerFence),
 * Compute.cpp:944-950.  The heap handle becomes borrowed device pointers (valid until
 * destroy); the fence handle becomes the producer's mapc_fence. */
typedef struct mapc_shared_handles {
    void       *posvelo[2];         /* device: mapc_posvelo[num_local] for side 0 / 1 (m_heap)  */
    void       *packed_pos[2];      /* device: float4[N] packed positions for side 0 / 1        */
    mapc_fence *fence;              /* producer fence (m_fence)                                 */
    void       *compute_stream;     /* cudaStream_t the steps are enqueued on                   */
    uint64_t    aligned_data_size;  /* bytes of one posvelo side (m_alignedDataSize)            */
    uint32_t    buffer_index;       /* m_bufferIndex: side the NEXT Simulate writes             */
    uint32_t    first_particle;     /* shard range                                              */
    uint32_t    num_local;
    int32_t     device;
} mapc_shared_handles;
/* consumer_fence may be NULL (headless, nobody to wait for) */
MAPC_API mapc_status mapc_compute_shared_handles(mapc_compute *c, mapc_fence *consumer_fence,
                                                 mapc_shared_handles *out);

/* AdapterShared::GetGpuTimes(), AdapterShared.h:51 -> the "simulate ms" timer
 * (Compute.cpp:445-446): exponential moving average over 20 samples as D3D12GpuTimer does
 * (include/D3D12GpuTimer.h:151-153) plus the last raw sample, both in milliseconds. */
MAPC_API mapc_status mapc_compute_gpu_times(mapc_compute *c, float *ms_average, float *ms_last);

/* Raw per-step samples of the same timer, oldest first, resolved since the previous call (at most
 * `capacity`; the log keeps the latest 4096).  Lets a benchmark average the device time of the
 * steps it issued without synchronising between them. */
MAPC_API mapc_status mapc_compute_step_times(mapc_compute *c, float *ms_out, int capacity, int *count);
/* Sharded handles, NCCL exchange: duration of the newest completed position all-gather, and how long it
 * was still running after the step that consumes it had begun (<= 0: finished before; > 0: the remote
 * cells of that step started this late -- the local cells ran meanwhile).  Evidence of the overlap. */
MAPC_API mapc_status mapc_compute_exchange_times(mapc_compute *c, float *gather_ms,
                                                 float *tail_past_step_begin_ms);
/* Make the compute stream wait for the exchange (all-gather) work issued so far, without blocking
 * the host: an event recorded on compute_stream afterwards covers the whole step. */
MAPC_API mapc_status mapc_compute_flush(mapc_compute *c);

/* void Compute::CopyState(Compute* other), Compute.cpp:303-410: drain both, copy both sides of
 * positions and velocities device-to-device (cudaMemcpyPeerAsync across devices). */
MAPC_API mapc_status mapc_compute_copy_state(mapc_compute *dst, mapc_compute *src);

/* InitializeParticles / LoadParticles, Compute.cpp:667-812, :820-923: two shells of radius
 * PARTICLE_SPREAD centred at x = +-0.75*PARTICLE_SPREAD, tangential speed
 * INITIAL_PARTICLE_SPEED, using the seeded scalar LCG of the USE_SCALAR_OPTIMIZED variant
 * (Compute.cpp:596-609, :719-749).  The reference seeds from std::random_device; here the
 * seed is an argument so runs are reproducible. */
MAPC_API mapc_status mapc_compute_init_particles(mapc_compute *c, uint32_t seed);

/* ---- headless consumer ------------------------------------------------------------------------
 * Takes the place of the Render worker as the consumer of simulation results, keeping its fence
 * protocol and one-frame latency (Particles/Render.cpp:789-831 CopySimulationResults, :839-937
 * Draw, :653-677 MoveToNextFrame): a COPY stream pulls the packed positions of the side the
 * previous Simulate wrote into a buffer local to the consumer's device (peer copy across devices,
 * the cross-adapter heap's job in the reference), and a "render" stream consumes the local buffer
 * -- headless, that is a dump to pinned host memory instead of DrawInstanced + Present.
 * Fences: copy fence (shared with the producer, Render.cpp:605-618), render fence (frame throttle).
 * Frame loop, exactly Particles::Draw (Particles.cpp:446-456):
 *     F = mapc_compute_fence_value(compute);
 *     mapc_consumer_draw(consumer, nDraw, &F, nCopy);
 *     mapc_compute_simulate(compute, nSim, dt, damping, F);
 */
typedef struct mapc_consumer mapc_consumer;
/* Render::Render + Particles::ShareHandles (Particles.cpp:191-208): attaches the consumer's copy
 * fence to `producer` (GetSharedHandles), adopts its buffer index (SetShared, Render.cpp:222-224)
 * and copies the initial positions into both local buffers. */
MAPC_API mapc_status mapc_consumer_create(mapc_consumer **out, mapc_compute *producer, int device);
/* The same with flags.  MAPC_CONSUMER_ASYNC is the reference's async mode -- renderer and simulation on ONE
 * adapter (Particles::ShareHandles, Particles.cpp:202-207; Compute::SetAsync, Compute.cpp:956-987;
 * Render::Draw, Render.cpp:849-852 and :928-932): no copy stream and no local buffers, the consumer reads
 * the producer's packed positions in place after waiting for compute fence F-1, waits for F before it signals
 * its render fence, and hands the RENDER fence value back for Simulate to wait on.  `device` must be the
 * producer's.  (The reference aliases the renderer's buffers into the compute object and copies them back
 * in ResetFromAsyncHelper, Compute.cpp:260-298; with CUDA both sides address the producer's buffers, so
 * there is nothing to alias or to restore when the consumer goes away.)  The frame a Draw dumps is then
 * the result of the previous Simulate (one step of latency instead of two).
 * A sharded producer is allowed in both modes: the consumer then sees that rank's shard only
 * (counts passed to mapc_consumer_draw stay global and are clipped to the shard). */
#define MAPC_CONSUMER_ASYNC 1u
MAPC_API mapc_status mapc_consumer_create_ex(mapc_consumer **out, mapc_compute *producer, int device,
                                             uint32_t flags);
MAPC_API mapc_status mapc_consumer_destroy(mapc_consumer *r);
/* HANDLE Render::Draw(int numActive, Particles*, UINT64& inout_fenceValue, int numCopied),
 * Render.cpp:839-938, non-async path.  in: the fence value the upcoming Simulate will signal;
 * out: the consumer fence value that Simulate must pass on.  Blocks the host only when two frames
 * are already in flight (the returned-handle wait of Particles.cpp:452-456). */
MAPC_API mapc_status mapc_consumer_draw(mapc_consumer *r, int num_active_particles,
                                        uint64_t *inout_fence_value, int num_particles_copied);
/* Positions (x, y, z, |accel|) of the newest completed frame in pinned host memory, the frame
 * number (0 = initial state) and how many bodies it holds.  Valid until the next draw call. */
MAPC_API mapc_status mapc_consumer_latest(mapc_consumer *r, const float **host_positions,
                                          uint64_t *frame, uint32_t *count);
/* Protocol counters for logs and tests: out[0] copy fence completed, out[1] copy fence value issued,
 * out[2] render fence completed, out[3] render fence value to be issued next, out[4] shared buffer
 * index, out[5] current (local) buffer index, out[6] frames drawn, out[7] device-to-device copies issued by
 * Draw (always 0 for an async consumer). */
MAPC_API mapc_status mapc_consumer_counters(const mapc_consumer *r, uint64_t out[8]);
/* Render::WaitForGpu, Render.cpp:626-647: drains copy and render streams */
MAPC_API mapc_status mapc_consumer_wait_for_gpu(mapc_consumer *r);

/* ---- plan / diagnostics -------------------------------------------------------------------- */
/* Canonical number of j segments: 32 for every n (a multiple of every supported GPU count, so no segment
 * straddles two shards).  A segment is evaluated as sequential fp32 chains of MAPC_CHAIN_SOURCES sources,
 * counted from the segment's first source; each chain sum is folded left to right into the segment's
 * partial, and the 32 partials left to right into the acceleration -- independent of the GPU count, the
 * launch shape and the order in which cells run.  Bounded chains keep the rounding noise of the sum
 * independent of N: between two correctly rounded fp32 evaluations of the formula the worst target differs
 * by ~2e-6 at N = 262,144 ... 4,194,304, against 1.07e-5 with 32,768-term chains. */
MAPC_API int mapc_plan_segments(uint32_t n_sources);
/* sources per sequential accumulation chain of the canonical order (MAPC_CHAIN_SOURCES) */
MAPC_API int mapc_plan_chain_sources(void);
/* number of this library's kernels launched by the handle so far */
MAPC_API uint64_t mapc_compute_kernel_launches(const mapc_compute *c);
/* Launch geometry the next all-pairs step would use (pairs of bodies per thread, threads per
 * block, number of blocks) -- for logs and tests. */
MAPC_API mapc_status mapc_compute_plan(const mapc_compute *c, int num_active_particles,
                                       int *pairs_per_thread, int *threads_per_block,
                                       int *num_blocks, int *segments);
/* Pure-FFMA microbenchmark on `device`: packed != 0 uses FFMA2 (fma.rn.f32x2).  Returns the
 * sustained TFLOP/s (2 flop per lane-FMA) and the kernel time. */
MAPC_API mapc_status mapc_fp32_peak_probe(int device, int packed, float *tflops, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* MAPC_H */
