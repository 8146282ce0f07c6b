// launch_cfg.cuh — per-device launch configuration of a kernel.
//
// The dynamic shared-memory opt-in (cudaFuncAttributeMaxDynamicSharedMemorySize) and the occupancy of a
// kernel are properties of (kernel, device), not of the process: a host that drives one handle per GPU
//  needs them on every device.  kernel_cfg<K>() caches them per device behind a mutex and
// leaves the calling thread's current device alone.
#pragma once
#include <cuda_runtime.h>
#include <mutex>

namespace nlb {

struct KernelCfg {
    int ctas_per_sm = 0;
    int num_sms = 0;
};

constexpr int NLB_MAX_DEVICES = 64;

// Configuration of `Kernel` launched with `threads` threads and `smem` bytes of dynamic shared memory on the
// CURRENT device.  Returns cudaSuccess or the failing call's error.
template <auto Kernel>
cudaError_t kernel_cfg(int threads, size_t smem, KernelCfg* out) {
    static std::mutex mu;
    static KernelCfg cache[NLB_MAX_DEVICES];
    static size_t cache_smem[NLB_MAX_DEVICES];
    static bool have[NLB_MAX_DEVICES] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    const bool cached = dev >= 0 && dev < NLB_MAX_DEVICES && have[dev] && cache_smem[dev] == smem;
    if (!cached) {
        KernelCfg c;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        e = cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.ctas_per_sm, Kernel, threads, smem);
        if (e != cudaSuccess) return e;
        if (c.ctas_per_sm < 1) c.ctas_per_sm = 1;
        if (dev < 0 || dev >= NLB_MAX_DEVICES) { *out = c; return cudaSuccess; }
        cache[dev] = c;
        cache_smem[dev] = smem;
        have[dev] = true;
    }
    *out = cache[dev];
    return cudaSuccess;
}

// RAII: make `device` current for the scope and restore the caller's device afterwards, so that an entry point
// of the C ABI never changes the calling thread's CUDA device (hosts that drive several GPUs rely on it).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != device) err = cudaSetDevice(device);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace nlb
