// fence.hpp -- ID3D12Fence semantics on CUDA streams: monotonically increasing 64-bit values that
// any stream of any device (and the host) can signal and wait on, INCLUDING waits for values whose
// signal has not been submitted yet (the reference does exactly that: Render::CopySimulationResults
// makes the copy queue wait for the value the *upcoming* Simulate will signal, Render.cpp:826).
//
// CUDA cannot express that directly.  cuStreamWaitValue64 can wait on a future value, but the
// ordering it creates is invisible to the CUDA scheduler, which may then serialise the waiting
// stream ahead of the stream that has to signal -- measured here as a hard hang once the copied
// block exceeded a few KB (tools/consumer_debug.py).  So:
//   * a signal = cudaEventRecord + cuStreamWriteValue64 (the word is what the host polls and what
//     GetCompletedValue returns; the event is what other streams wait on -- a CUDA-visible edge);
//   * a wait whose signal is already submitted = cudaStreamWaitEvent (or nothing if completed);
//   * a wait whose signal is NOT yet submitted gates the stream on the host: that wait and every
//     later operation of the same stream are queued and replayed, in order, the moment the signal
//     is submitted.  Nothing is ever enqueued on the GPU ahead of the work it depends on.
// Not thread safe (a handle belongs to one host thread, like the reference's objects).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <deque>
#include <functional>
#include <string>
#include <vector>

#include "../../include/mapc.h"

namespace mapc {

struct GatedStream;

struct FenceSignal {
    uint64_t value;
    cudaEvent_t event;
    int device;   // device the event was created on
};

}  // namespace mapc

struct mapc_fence {
    volatile uint64_t *word = nullptr;  // pinned, portable, mapped: completed value
    uint64_t submitted = 0;             // highest value whose signal has been submitted
    std::deque<mapc::FenceSignal> signals;        // submitted stream signals, ascending, not yet pruned
    std::vector<mapc::GatedStream *> waiters;     // streams gated on a value of this fence
    std::vector<cudaEvent_t> spare;               // completed signals' events, reused instead of re-created
    std::vector<int> spare_device;
    // "light" signals: the signalling kernel itself writes `word` when it finishes, and nothing is put on
    // the stream for the signal.  A later wait on another stream records an event on `light_stream` then
    // (everything submitted to it so far, hence the signalling kernel, is covered).
    cudaStream_t light_stream = nullptr;
    int light_device = -1;
    uint64_t light_value = 0;                     // highest value signalled this way
    uint64_t light_floor = 0;                     // values <= this were not (they predate the first light signal)
};

namespace mapc {

struct StreamOp {
    enum Kind { kWait, kSignal, kSignalLight, kCall } kind;
    mapc_fence *fence;
    uint64_t value;
    std::function<mapc_status()> fn;
};

struct GatedStream {
    cudaStream_t stream = nullptr;
    int device = 0;
    std::deque<StreamOp> pending;
    bool draining = false;
    // An operation that was queued behind a gate runs later, inside whichever call submits the signal
    // it waited for -- a call that may belong to another object.  Its failure is kept here and reported
    // by the next operation on THIS stream instead of being lost.
    mapc_status deferred_error = MAPC_OK;
    std::string deferred_message;
};

// implemented in mapc.cu (they need its error plumbing and driver entry points)
uint64_t fence_completed(const mapc_fence *f);
mapc_status fence_submit_signal(mapc_fence *f, cudaStream_t stream, int device, uint64_t value);
mapc_status fence_submit_wait(mapc_fence *f, cudaStream_t stream, uint64_t value);  // needs fence_ready
inline bool fence_ready(const mapc_fence *f, uint64_t value)
{
    return f->submitted >= value || fence_completed(f) >= value;
}
mapc_status gs_drain(GatedStream *gs);
mapc_status gs_wait(GatedStream *gs, mapc_fence *f, uint64_t value);
mapc_status gs_signal(GatedStream *gs, mapc_fence *f, uint64_t value);
mapc_status gs_signal_light(GatedStream *gs, mapc_fence *f, uint64_t value);   // see mapc_fence::light_stream
mapc_status gs_call(GatedStream *gs, std::function<mapc_status()> fn);
void gs_detach(GatedStream *gs);                 // before a gated stream goes away
void fence_notify(mapc_fence *f);                // a signal was submitted: replay what it unblocks

}  // namespace mapc
