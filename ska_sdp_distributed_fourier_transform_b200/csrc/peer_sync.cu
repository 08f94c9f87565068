// SwiFTly B200 -- device-side signalling between the ranks of the sharded transform.
//
// The strips of a subgrid batch leave the axis-1 kernel straight into the owners' peer-mapped
// receive buffers (TMA bulk tensor stores over NVLink).  What is left of the "collective" is
// ordering: an owner may start its axis-0 kernel once every rank has finished writing the
// batch.  Two one-warp kernels do that with flags in peer-mapped memory:
//   signal: after my axis-1 kernel of batch k, store k into flags[my_rank] of EVERY rank
//   wait  : before my axis-0 kernel of batch k, spin until all my flags are >= k
// They are separate launches so that other work (the axis-1 kernel of batch k + 1) can sit
// between them on the stream: the wait then usually finds the flags already set.
#include "capi_util.h"

#if !defined(SWIFTLY_EMU)
namespace {

__global__ void peer_signal_kernel(long long* const* flags, int n_peers, int my_rank,
                                   long long value) {
    const int p = threadIdx.x;
    if (p < n_peers) {
        // release at system scope: everything this stream did before (the kernel boundary made
        // the strips of the batch visible) is ordered before the flag
        __threadfence_system();
        asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(flags[p] + my_rank), "l"(value)
                     : "memory");
    }
}

__global__ void peer_wait_kernel(const long long* my_flags, int n_peers, long long value,
                                 long long timeout_cycles, int* status) {
    const int p = threadIdx.x;
    if (p < n_peers) {
        const long long t0 = clock64();
        long long v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(my_flags + p)
                         : "memory");
            if (v >= value) break;
            if (clock64() - t0 > timeout_cycles) {
                atomicExch(status, 1 + p);  // a peer never arrived: report instead of hanging
                break;
            }
            __nanosleep(200);
        }
    }
}

}  // namespace
#endif

// flags: n_peers device pointers (my mapping of every rank's flag array, n_peers entries each),
// given as a HOST array of device pointers
extern "C" int swiftly_b200_peer_signal(const swiftly_b200* h, void* const* flags_dev_table,
                                        int n_peers, int my_rank, int64_t value, void* stream) {
    if (!h || !flags_dev_table) return swiftly::einval("peer_signal: NULL argument");
    if (n_peers < 1 || n_peers > 32) return swiftly::einval("peer_signal: 1..32 peers");
#if defined(SWIFTLY_EMU)
    // single process: the "peers" are plain arrays
    long long* const* flags = (long long* const*)flags_dev_table;
    for (int p = 0; p < n_peers; ++p) flags[p][my_rank] = value;
    (void)stream;
    return SWIFTLY_B200_OK;
#else
    SW_DEVICE_GUARD(h);
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((long long* const*)flags_dev_table,
                                                           n_peers, my_rank, (long long)value);
    SW_CUDA(cudaGetLastError(), "peer_signal launch");
    return SWIFTLY_B200_OK;
#endif
}

// status: device int (0 = ok; 1 + p = peer p did not arrive within the timeout)
extern "C" int swiftly_b200_peer_wait(const swiftly_b200* h, const void* my_flags, int n_peers,
                                      int64_t value, double timeout_s, void* status,
                                      void* stream) {
    if (!h || !my_flags || !status) return swiftly::einval("peer_wait: NULL argument");
    if (n_peers < 1 || n_peers > 32) return swiftly::einval("peer_wait: 1..32 peers");
#if defined(SWIFTLY_EMU)
    const long long* f = (const long long*)my_flags;
    for (int p = 0; p < n_peers; ++p)
        if (f[p] < value) *(int*)status = 1 + p;
    (void)stream;
    (void)timeout_s;
    return SWIFTLY_B200_OK;
#else
    SW_DEVICE_GUARD(h);
    const long long cycles = (long long)(timeout_s * 1.9e9);
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const long long*)my_flags, n_peers,
                                                         (long long)value, cycles, (int*)status);
    SW_CUDA(cudaGetLastError(), "peer_wait launch");
    return SWIFTLY_B200_OK;
#endif
}
