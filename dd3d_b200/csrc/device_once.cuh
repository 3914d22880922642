// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE, while a process may hold one engine per GPU
// (include/dd3d_b200.h: "a handle is bound to the CUDA device that was current at dd3d_create").  The launchers therefore
// remember which devices they configured instead of a process-wide flag.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dd3d {

// true exactly once per (mask, current device); devices >= 64 always return true (the attribute is then set every launch)
inline bool first_use_on_device(uint64_t* mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if ((*mask >> dev) & 1ull) return false;
    *mask |= 1ull << dev;
    return true;
}

inline int current_device_or_zero() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    return dev;
}

}  // namespace dd3d
