// Per-device caches. __constant__ / __device__ symbols and cudaFuncSetAttribute settings belong to ONE device: a flag
// that remembers "already uploaded / already set" must be kept per device, or a second GPU driven from the same process
// silently runs with zero tables. cudaGetDevice is a thread-local read.
#pragma once
#include <cuda_runtime.h>

namespace resr {

static constexpr int kMaxDevices = 64;

inline int current_device_index() {
    int d = 0;
    cudaGetDevice(&d);
    return (d < 0 || d >= kMaxDevices) ? 0 : d;
}

template <typename T>
struct PerDevice {
    T v[kMaxDevices]{};
    T& cur() { return v[current_device_index()]; }
};

}  // namespace resr
