// SwiFTly B200 -- size dispatch for the line-FFT kernels (included by the
// per-primitive translation units dispatch_*.cu).
#pragma once

#include "plan.h"

namespace swiftly {

template <int N>
struct LinesPerCta {
    static constexpr int T = FftCfg<N>::T;
    static constexpr int V = T >= 256 ? 1 : ((256 / T) > 16 ? 16 : (256 / T));
};

inline int grid_for(int64_t n_lines, int lpc) {
    int64_t blocks = (n_lines + lpc - 1) / lpc;
    const int64_t cap = 148 * 32;  // grid-stride loop covers the rest
    return (int)(blocks < cap ? blocks : cap);
}

template <int N, int DIR, class Op>
int launch_lines(const swiftly_b200* h, const Op& op, bool line_fastest, cudaStream_t s) {
    const cplx* tw = twiddles(h, N);
    if (!tw) return SWIFTLY_B200_ECUDA;
    constexpr int LPC = LinesPerCta<N>::V;
    cudaError_t e;
    if (line_fastest && LPC > 1) {
        LineKernel<N, DIR, LPC, true, Op> k{op, tw};
        e = launch_body(k, grid_for(op.g.n_lines, LPC), k.SMEM, s);
    } else {
        LineKernel<N, DIR, LPC, false, Op> k{op, tw};
        e = launch_body(k, grid_for(op.g.n_lines, LPC), k.SMEM, s);
    }
    return e == cudaSuccess ? SWIFTLY_B200_OK : cuda_fail(e, "line FFT kernel launch");
}

template <int H, int DIR, class Op>
int launch_split(const swiftly_b200* h, const Op& op, cudaStream_t s) {
    const cplx* tw = twiddles(h, H);
    const cplx* tw2 = twiddles_full(h, 2 * H);
    if (!tw || !tw2) return SWIFTLY_B200_ECUDA;
    // persistent CTAs: two per SM's worth of lines in flight keeps the scratch L2 resident
    int64_t blocks = op.g.n_lines < 296 ? op.g.n_lines : 296;
    cplx* scratch = split_scratch(h, s, (size_t)blocks * H);
    if (!scratch) return SWIFTLY_B200_ECUDA;
    SplitLineKernel<H, DIR, Op> k{op, tw, tw2, scratch};
    cudaError_t e = launch_body(k, (int)blocks, k.SMEM, s);
    return e == cudaSuccess ? SWIFTLY_B200_OK : cuda_fail(e, "split line FFT kernel launch");
}

inline int unsupported(int n) {
    set_error("FFT length " + std::to_string(n) +
              " is not supported by this build (powers of two 16..16384 only)");
    return SWIFTLY_B200_EUNSUPPORTED;
}

#define SW_DIRECT_CASES(DIR, Op)                                         \
    case 16: return launch_lines<16, DIR, Op>(h, op, lf, s);             \
    case 32: return launch_lines<32, DIR, Op>(h, op, lf, s);             \
    case 64: return launch_lines<64, DIR, Op>(h, op, lf, s);             \
    case 128: return launch_lines<128, DIR, Op>(h, op, lf, s);           \
    case 256: return launch_lines<256, DIR, Op>(h, op, lf, s);           \
    case 512: return launch_lines<512, DIR, Op>(h, op, lf, s);           \
    case 1024: return launch_lines<1024, DIR, Op>(h, op, lf, s);         \
    case 2048: return launch_lines<2048, DIR, Op>(h, op, lf, s);         \
    case 4096: return launch_lines<4096, DIR, Op>(h, op, lf, s);         \
    case 8192: return launch_lines<8192, DIR, Op>(h, op, lf, s);

}  // namespace swiftly
