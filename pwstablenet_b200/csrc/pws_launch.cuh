// pws_launch.cuh -- host-side launch plumbing shared by the persistent TMA kernels.
//
// The persistent kernels keep a few words of per-launch state in __device__ memory (the dynamic tile counter, the
// backward's zero-fill completion counters).  A launch owns one of kLaunchSlots slots of that state; the last CTA to
// leave hands the slot back clean.  SlotLease makes the reuse of a slot safe whatever the streams do: before a launch
// takes slot s, its stream is made to wait for the event the previous user of s recorded after ITS launch, so two
// launches never run on the same slot -- at worst the 65th launch in flight queues behind the first.  The lease holds
// the device's slot lock from acquire to release (a launch is a few microseconds; each device has its own lock, so
// Python threads driving different devices -- the reference's DataParallel model -- never contend).
// A stream under CUDA-graph capture gets no slot (ok() == false): a captured launch would bake the slot index into the
// graph and could be replayed next to a live launch on the same slot; the callers then take the non-persistent kernels.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

namespace pws {

constexpr int kLaunchSlots = 64;
enum SlotRing : int { kRingForward = 0, kRingBackward = 1, kNumRings = 2 };

class SlotLease {
public:
    SlotLease(SlotRing ring, cudaStream_t st);
    ~SlotLease();  // records the slot's event on the stream and unlocks
    SlotLease(const SlotLease &) = delete;
    SlotLease &operator=(const SlotLease &) = delete;
    bool ok() const { return slot_ >= 0; }
    int slot() const { return slot_; }
    void cancel() { launched_ = false; }  // nothing was launched: leave the slot's event as it was

private:
    void *ring_ = nullptr;
    cudaStream_t st_;
    int slot_ = -1;
    bool launched_ = true;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: `done` is a bit mask of the
// devices it has been set on for one kernel instantiation.
bool ensure_dynamic_smem(const void *func, int bytes, std::atomic<uint64_t> &done);

}  // namespace pws
