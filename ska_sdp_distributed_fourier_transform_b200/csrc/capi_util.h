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

}  // namespace swiftly

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
