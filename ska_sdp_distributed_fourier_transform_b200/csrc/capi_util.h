// SwiFTly B200 -- small helpers shared by the C-ABI translation units.
#pragma once

#include <string>

#include "plan.h"

namespace swiftly {

inline int einval(const std::string& msg) {
    set_error(msg);
    return SWIFTLY_B200_EINVAL;
}

inline int64_t floordiv(int64_t a, int64_t b) {  // python // for b > 0
    int64_t q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}

// Make the plan's device current for the duration of a call and restore the caller's
// device afterwards (also from swiftly_b200_destroy, which may run at arbitrary GC time):
// a process that drives several GPUs must not find its current device changed by a
// library call.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err != cudaSuccess) {
            prev = -1;
            return;
        }
        if (prev != device) err = cudaSetDevice(device);
        else prev = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    bool ok() const { return err == cudaSuccess; }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace swiftly

#define SW_DEVICE_GUARD(h)                                                        \
    swiftly::DeviceGuard device_guard__((h)->device);                             \
    if (!device_guard__.ok()) return swiftly::cuda_fail(device_guard__.err, "cudaSetDevice")

#define SW_CUDA(call, what)                                           \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return swiftly::cuda_fail(e__, what); \
    } while (0)

#define SW_TRY(expr)                              \
    do {                                          \
        int rc__ = (expr);                        \
        if (rc__ != SWIFTLY_B200_OK) return rc__; \
    } while (0)
