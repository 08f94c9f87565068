// SwiFTly B200 -- common definitions shared by all kernels.
//
// Every kernel body in this library is written as a functor templated on an
// execution context ("Ctx") that provides the thread id, block id, shared
// memory and the CTA barrier.  The CUDA build instantiates the bodies with
// DeviceCtx inside a __global__ entry point.  A second, test-only build
// (tests/emu, -DSWIFTLY_EMU, plain g++) instantiates the very same bodies with
// a host context that runs every CUDA thread as a fibre, so that the index
// algebra of the kernels can be exercised in a container without a GPU.  The
// emulator is test tooling: the product never loads it.
#pragma once

#include <stdint.h>

#if defined(SWIFTLY_EMU)
#include "emu_runtime.h"
#else
#include <cuda_runtime.h>
#endif

#if defined(__CUDACC__)
#define SW_HD __host__ __device__ __forceinline__
#define SW_D __device__ __forceinline__
#else
#define SW_HD inline
#define SW_D inline
#endif

namespace swiftly {

typedef double2 cplx;

SW_HD cplx mk(double x, double y) {
    cplx r;
    r.x = x;
    r.y = y;
    return r;
}
SW_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
SW_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
SW_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SW_HD cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }
SW_HD cplx cconj(cplx a) { return mk(a.x, -a.y); }
// multiply by DIR * i   (DIR = -1: forward transform, +1: inverse transform)
template <int DIR>
SW_HD cplx mul_i(cplx a) {
    return DIR < 0 ? mk(a.y, -a.x) : mk(-a.y, a.x);
}

// read-only global load (LDG.E.128 through the non-coherent path on device)
SW_HD cplx ldg_c(const cplx* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
SW_HD double ldg_d(const double* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// streaming global accesses: data that is touched once per kernel (inputs and outputs are
// far larger than the 126 MB L2) should not push the few reused lines (twiddle / window
// tables, split-kernel scratch) out of L2
SW_HD cplx ld_stream(const cplx* p) {
#if defined(__CUDA_ARCH__)
    return __ldcs(p);
#else
    return *p;
#endif
}
SW_HD void st_stream(cplx* p, cplx v) {
#if defined(__CUDA_ARCH__)
    __stcs(p, v);
#else
    *p = v;
#endif
}

// software prefetch of the 32-byte sector(s) holding *p into L2 (no register is tied up):
// all warps of a CTA are in the same phase of a transform, so a plain load at the start of
// the next phase exposes the full DRAM latency; the prefetch is issued one phase ahead
SW_HD void prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// non-negative modulo for possibly negative a, n > 0
SW_HD int64_t pmod(int64_t a, int64_t n) {
    int64_t r = a % n;
    return r < 0 ? r + n : r;
}

#if defined(__CUDACC__) && !defined(SWIFTLY_EMU)
// Execution context on the device: one CTA.
struct DeviceCtx {
    int tid, bid, nblocks;
    char* smem;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    // named barrier over `count` threads (a multiple of 32; whole warps), id 1..15
    __device__ __forceinline__ void group_sync(int id, int count) const {
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
    }
};

extern __shared__ __align__(16) char swiftly_dyn_smem[];

// register budget: at least 512 resident threads per SM (<= 128 registers/thread)
template <class Body>
struct MinBlocks {
    static constexpr int V = Body::THREADS >= 512 ? 1 : 512 / Body::THREADS;
};

template <class Body>
__global__ void __launch_bounds__(Body::THREADS, MinBlocks<Body>::V) kernel_entry(const Body body) {
    DeviceCtx ctx;
    ctx.tid = threadIdx.x;
    ctx.bid = blockIdx.x;
    ctx.nblocks = gridDim.x;
    ctx.smem = swiftly_dyn_smem;
    body(ctx);
}

// launch helper: returns cudaError_t
template <class Body>
inline cudaError_t launch_body(const Body& body, int grid, size_t smem_bytes, cudaStream_t stream) {
    if (grid <= 0) return cudaSuccess;
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel_entry<Body>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem_bytes);
        if (e != cudaSuccess) return e;
    }
    kernel_entry<Body><<<grid, Body::THREADS, smem_bytes, stream>>>(body);
    return cudaGetLastError();
}
#endif

}  // namespace swiftly
