// ude_host_common.h — host-side helpers shared by the translation units of libude_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>

#include "../../include/ude_b200.h"

namespace ude {

inline char* cuda_err_buf() {
    static thread_local char buf[256] = "";
    return buf;
}
inline int cuda_fail(cudaError_t e, const char* where) {
    snprintf(cuda_err_buf(), 256, "%s: %s", where, cudaGetErrorString(e));
    return UDE_ERR_CUDA;
}
#define UDE_CUDA(call)                                  \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return ::ude::cuda_fail(e__, #call); \
    } while (0)

inline int round4(int x) { return (x + 3) & ~3; }
inline long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

struct DeviceInfo {
    int ok = 0;
    int n_sm = 0;
    int cc_major = 0;
    int max_smem_optin = 0;
};
// Properties of the CURRENT device, cached per device ordinal (a process may drive several GPUs).
constexpr int kMaxDevices = 64;
inline DeviceInfo& device_info() {
    static DeviceInfo info[kMaxDevices + 1];
    static std::once_flag once[kMaxDevices + 1];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        (void)cudaGetLastError();
        return info[kMaxDevices];          // stays !ok: no device visible
    }
    std::call_once(once[dev], [dev] {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return;
        info[dev].n_sm = prop.multiProcessorCount;
        info[dev].cc_major = prop.major;
        info[dev].max_smem_optin = (int)prop.sharedMemPerBlockOptin;
        info[dev].ok = 1;
    });
    return info[dev];
}

// Opt in to `bytes` of dynamic shared memory for `kernel`, once per (kernel, device, size): repeated launches -- and
// launches under CUDA-graph capture -- must not touch the function attributes again.  Keyed by the kernel's ADDRESS:
// instantiations with the same signature share one template instance of this function.
template <typename K>
inline int set_smem(K kernel, int bytes) {
    struct Have { int bytes[kMaxDevices] = {}; };
    static std::map<const void*, Have> have;
    static std::mutex mu;
    int dev = 0;
    UDE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return UDE_ERR_INVALID;
    std::lock_guard<std::mutex> lock(mu);
    int& cur = have[(const void*)kernel].bytes[dev];
    if (bytes <= cur) return UDE_OK;
    UDE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    cur = bytes;
    return UDE_OK;
}

}  // namespace ude
