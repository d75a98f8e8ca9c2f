// include/mapc_compute.hpp -- C++ host side over the C ABI (include/mapc.h).
//
// `mapc::Compute` mirrors the public surface of the reference's `class Compute`
// (Particles/Compute.h:33-78) so that the orchestration code of Particles::Draw
// (Particles/Particles.cpp:446-448) reads the same:
//
//     UINT64 fence = compute.GetFenceValue();
//     render.Draw(nDraw, fence, nCopy);                 // mapc::HeadlessRender
//     compute.Simulate(nSim, fence);
//
// Error behaviour follows the reference: every failing call throws (the reference throws
// HrException from ThrowIfFailed, dx-samples-include/DXSampleHelper.h:22-46); nothing is caught.
// Non-copyable, non-movable like the reference class (Compute.h:42-45).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mapc.h"

namespace mapc {

class Error : public std::runtime_error {
public:
    Error(mapc_status st, const std::string &what) : std::runtime_error(what), status(st) {}
    mapc_status status;
};

inline void ThrowIfFailed(mapc_status st)
{
    if (st != MAPC_OK) throw Error(st, mapc_last_error());
}

class Fence {
public:
    explicit Fence(std::uint64_t initial = 0) { ThrowIfFailed(mapc_fence_create(&m_fence, initial)); }
    ~Fence() { mapc_fence_destroy(m_fence); }
    Fence(const Fence &) = delete;
    Fence &operator=(const Fence &) = delete;
    std::uint64_t GetCompletedValue() const { return mapc_fence_completed_value(m_fence); }
    void Signal(std::uint64_t v) { ThrowIfFailed(mapc_fence_signal_host(m_fence, v)); }
    void Wait(std::uint64_t v, int timeout_ms = -1) const { ThrowIfFailed(mapc_fence_wait_host(m_fence, v, timeout_ms)); }
    mapc_fence *Get() const { return m_fence; }

private:
    mapc_fence *m_fence = nullptr;
};

class Compute {
public:
    // Compute(UINT in_numParticles, IDXGIAdapter1* in_pAdapter, bool in_useIntelCommandQueueExtension,
    //         Compute* in_pCompute = 0)   -- Compute.h:36-39.  The adapter becomes a CUDA device index.
    Compute(std::uint32_t in_numParticles, int in_device, Compute *in_pCompute = nullptr)
        : m_numParticles(in_numParticles)
    {
        ThrowIfFailed(mapc_compute_create(&m_handle, in_numParticles, in_device,
                                          in_pCompute ? in_pCompute->m_handle : nullptr));
    }
    // i-sharded form, one process per GPU (no reference analogue)
    Compute(std::uint32_t in_numParticles, int in_device, int rank, int world, const void *nccl_unique_id)
        : m_numParticles(in_numParticles)
    {
        ThrowIfFailed(mapc_compute_create_sharded(&m_handle, in_numParticles, in_device, rank, world, nccl_unique_id));
    }
    virtual ~Compute() { mapc_compute_destroy(m_handle); }

    Compute(const Compute &) = delete;
    Compute(Compute &&) = delete;
    Compute &operator=(const Compute &) = delete;
    Compute &operator=(Compute &&) = delete;

    // input is fence value of other adapter. waits to overwrite shared buffer.  (Compute.h:47-48)
    void Simulate(int in_numActiveParticles, std::uint64_t in_sharedFenceValue,
                  float in_deltaTime = MAPC_DEFAULT_DELTA_TIME, float in_damping = MAPC_DEFAULT_DAMPING)
    {
        ThrowIfFailed(mapc_compute_simulate(m_handle, in_numActiveParticles, in_deltaTime, in_damping,
                                            in_sharedFenceValue));
    }

    using SharedHandles = mapc_shared_handles;  // Compute.h:54-61
    SharedHandles GetSharedHandles(mapc_fence *in_fence)
    {
        SharedHandles h;
        ThrowIfFailed(mapc_compute_shared_handles(m_handle, in_fence, &h));
        return h;
    }

    std::uint64_t GetFenceValue() const { return mapc_compute_fence_value(m_handle); }  // Compute.h:64
    virtual void WaitForGpu() { ThrowIfFailed(mapc_compute_wait_for_gpu(m_handle)); }   // Compute.h:72

    // AdapterShared::GetGpuTimes(), AdapterShared.h:51: (seconds, name) pairs
    std::vector<std::pair<float, std::string>> GetGpuTimes()
    {
        float avg = 0.f, last = 0.f;
        ThrowIfFailed(mapc_compute_gpu_times(m_handle, &avg, &last));
        return {{avg * 1e-3f, "simulate ms"}};
    }

    void CopyState(Compute *in_pCompute) { ThrowIfFailed(mapc_compute_copy_state(m_handle, in_pCompute->m_handle)); }
    void InitializeParticles(std::uint32_t seed) { ThrowIfFailed(mapc_compute_init_particles(m_handle, seed)); }

    // headless extras
    void SetForceMode(mapc_force_mode mode) { ThrowIfFailed(mapc_compute_set_force_mode(m_handle, mode)); }
    void Upload(const mapc_posvelo *host, std::uint32_t n) { ThrowIfFailed(mapc_compute_upload(m_handle, host, n)); }
    void Download(mapc_posvelo *host, std::uint32_t first, std::uint32_t count)
    {
        ThrowIfFailed(mapc_compute_download(m_handle, host, first, count));
    }
    std::uint32_t NumParticles() const { return m_numParticles; }
    mapc_compute *Handle() const { return m_handle; }

private:
    const std::uint32_t m_numParticles;
    mapc_compute *m_handle = nullptr;
};

// The consumer role of Particles/Render.{h,cpp}, headless (see mapc_consumer_* in mapc.h)
class HeadlessRender {
public:
    // in_asyncMode: Render::SetAsyncMode(true) -- consumer and producer on one device, no copies (Particles.cpp:202-207)
    HeadlessRender(Compute &compute, int device, bool in_asyncMode = false)
    {
        ThrowIfFailed(mapc_consumer_create_ex(&m_handle, compute.Handle(), device, in_asyncMode ? MAPC_CONSUMER_ASYNC : 0u));
    }
    ~HeadlessRender() { mapc_consumer_destroy(m_handle); }
    HeadlessRender(const HeadlessRender &) = delete;
    HeadlessRender &operator=(const HeadlessRender &) = delete;

    // HANDLE Render::Draw(int numActive, Particles*, UINT64& inout_fenceValue, int numCopied) -- Render.cpp:839
    void Draw(int in_numActiveParticles, std::uint64_t &inout_fenceValue, int in_numParticlesCopied)
    {
        ThrowIfFailed(mapc_consumer_draw(m_handle, in_numActiveParticles, &inout_fenceValue, in_numParticlesCopied));
    }
    void WaitForGpu() { ThrowIfFailed(mapc_consumer_wait_for_gpu(m_handle)); }
    const float *Latest(std::uint64_t *frame, std::uint32_t *count)
    {
        const float *p = nullptr;
        ThrowIfFailed(mapc_consumer_latest(m_handle, &p, frame, count));
        return p;
    }

private:
    mapc_consumer *m_handle = nullptr;
};

}  // namespace mapc
