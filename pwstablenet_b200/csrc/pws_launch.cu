// pws_launch.cu -- launch slots and per-device kernel attributes (see pws_launch.cuh).
#include "pws_launch.cuh"

#include <mutex>

namespace pws {

namespace {

constexpr int kMaxDevices = 64;

struct Ring {
    std::mutex m;
    cudaEvent_t ev[kLaunchSlots];
    bool have[kLaunchSlots] = {};
    unsigned next = 0;
};

Ring *ring_of(int dev, SlotRing r)
{
    static Ring rings[kMaxDevices][kNumRings];
    if (dev < 0 || dev >= kMaxDevices) return nullptr;
    return &rings[dev][r];
}

}  // namespace

SlotLease::SlotLease(SlotRing ring, cudaStream_t st) : st_(st)
{
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); return; }
    if (cap != cudaStreamCaptureStatusNone) return;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
    Ring *rg = ring_of(dev, ring);
    if (!rg) return;
    rg->m.lock();
    const int s = (int)(rg->next % kLaunchSlots);
    if (!rg->have[s]) {
        if (cudaEventCreateWithFlags(&rg->ev[s], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            rg->m.unlock();
            return;
        }
        rg->have[s] = true;
    } else if (cudaStreamWaitEvent(st, rg->ev[s], 0) != cudaSuccess) {  // no-op when the previous user has finished
        cudaGetLastError();
        rg->m.unlock();
        return;
    }
    rg->next += 1;
    ring_ = rg;
    slot_ = s;
}

SlotLease::~SlotLease()
{
    if (!ring_) return;
    Ring *rg = static_cast<Ring *>(ring_);
    if (launched_ && cudaEventRecord(rg->ev[slot_], st_) != cudaSuccess) cudaGetLastError();
    rg->m.unlock();
}

bool ensure_dynamic_smem(const void *func, int bytes, std::atomic<uint64_t> &done)
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) { cudaGetLastError(); return false; }
    const bool tracked = dev < 64;
    if (tracked && (done.load(std::memory_order_acquire) >> dev) & 1u) return true;
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (tracked) done.fetch_or((uint64_t)1 << dev, std::memory_order_release);
    return true;
}

}  // namespace pws
