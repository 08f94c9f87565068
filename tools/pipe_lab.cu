// Dev tool: how do the shared-memory exchange (LSU) and the butterfly (FP64) phases of the
// FFT engine share an SM?  Times, per pass of a 4096-point line (256 threads x 16 points):
//   X  : exchange only        F : twiddle + radix-16 butterfly only      XF : both (a real pass)
// with 1 or 2 co-resident CTAs per SM (2 = the occupancy of the product kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include \
//        -I ska_sdp_distributed_fourier_transform_b200/csrc tools/pipe_lab.cu -o tools/pipe_lab
#include <cstdio>
#include <vector>
#include "fft_engine.cuh"

using namespace swiftly;

template <int MODE, bool CX>
__global__ void __launch_bounds__(256, 2) lab(const cplx* tw, cplx* out, int iters, long long* clk) {
    extern __shared__ __align__(16) char smem[];
    double* sm = (double*)smem;
    cplx* smc = (cplx*)smem;
    constexpr bool SUB = (MODE & 16) != 0;
    constexpr int N = SUB ? 1024 : 4096, T = N / 16, NS = 16, NSP = 1, RP = 16, R = 16, NB = N / R;
    const int lt = threadIdx.x % T;
    const int sub = threadIdx.x / T;
    sm += sub * (N + N / 16 + 1);
    smc += sub * (N + N / 16 + 1);
    auto SYNC = [&]() {
        if (SUB)
            asm volatile("bar.sync %0, %1;" ::"r"(1 + sub), "r"(T) : "memory");
        else
            __syncthreads();
    };
    cplx v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = mk(1.0 + 1e-9 * (lt + r), 1e-9 * (lt - r));
    __syncthreads();
    long long t0 = clock64();
    if ((MODE & 8) && (blockIdx.x / 148) % 2 == 1) {  // stagger: odd CTAs start half a pass late
        cplx w1 = mk(0.9999999, 1e-4);
        TwiddlePowers<16>::apply(v, w1);
        Radix<16, +1>::run(v);
    }
    for (int it = 0; it < iters; ++it) {
        if (MODE & 1) {
            const int j = lt;
            const int base = (j / NSP) * NS + (j & (NSP - 1));
            if (CX) {
#pragma unroll
                for (int r = 0; r < RP; ++r) smc[sm_phys(base + r * NSP)] = v[r];
                SYNC();
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = smc[sm_phys(j + r * NB)];
                SYNC();
            } else {
                double nx[16];
#pragma unroll
                for (int r = 0; r < RP; ++r) sm[sm_phys(base + r * NSP)] = v[r].x;
                SYNC();
#pragma unroll
                for (int r = 0; r < R; ++r) nx[r] = sm[sm_phys(j + r * NB)];
                SYNC();
#pragma unroll
                for (int r = 0; r < RP; ++r) sm[sm_phys(base + r * NSP)] = v[r].y;
                SYNC();
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = mk(nx[r], sm[sm_phys(j + r * NB)]);
                SYNC();
            }
        }
        if (MODE & 2) {
            cplx w1 = (MODE & 4) ? ldg_c(tw + (lt & 15)) : mk(0.9999999 + 1e-12 * it, 1e-4);
            if (MODE & 32) {  // independent twiddle "loads" (no power chain): v[r] *= const_r
#pragma unroll
                for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], mk(w1.x + 1e-9 * r, w1.y));
                Radix<16, +1>::run(v);
            } else {
            TwiddlePowers<16>::apply(v, w1);
            Radix<16, +1>::run(v);
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = cscale(v[r], 0.25);
        }
    }
    long long t1 = clock64();
    cplx s = mk(0, 0);
#pragma unroll
    for (int r = 0; r < 16; ++r) s = cadd(s, v[r]);
    out[blockIdx.x * 256 + lt] = s;
    if (lt == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
__global__ void __launch_bounds__(256, 2) peak(double* out, int iters, long long* clk) {
    double a[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) a[r] = 1.0 + 1e-9 * (threadIdx.x + r);
    const double c = 0.999999999, d = 1e-12;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (OP == 0) a[r] = fma(a[r], c, d);
                if (OP == 1) a[r] = a[r] + d;
                if (OP == 2) a[r] = a[r] * c;
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int r = 0; r < 16; ++r) s += a[r];
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run_peak(const char* name, int ctas, double* out, long long* clk) {
    const int iters = 2000;
    peak<OP><<<148 * ctas, 256>>>(out, 10, clk);
    peak<OP><<<148 * ctas, 256>>>(out, iters, clk);
    cudaDeviceSynchronize();
    std::vector<long long> h(148 * ctas);
    cudaMemcpy(h.data(), clk, sizeof(long long) * 148 * ctas, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto c : h) avg += (double)c;
    avg /= (148 * ctas);
    // per CTA: 8 warps x 128 instr per iteration
    double lanes = (double)ctas * 256.0 * 128.0 * iters / avg;
    printf("%-10s CTAs/SM %d: %.1f FP64 lane-instructions per clk per SM\n", name, ctas, lanes);
}

template <int MODE, bool CX>
void run(const char* name, int ctas_per_sm, const cplx* tw, cplx* out, long long* clk) {
    const int iters = 2000;
    // shared memory sized so that exactly ctas_per_sm CTAs fit an SM
    size_t smem = ctas_per_sm == 1 ? 120 * 1024 : (ctas_per_sm == 2 ? 100 * 1024 : 70 * 1024);
    cudaFuncSetAttribute(lab<MODE, CX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = 148 * ctas_per_sm;
    lab<MODE, CX><<<grid, 256, smem>>>(tw, out, 10, clk);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    lab<MODE, CX><<<grid, 256, smem>>>(tw, out, iters, clk);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto c : h) avg += (double)c;
    avg /= grid;
    printf("%-34s CTAs/SM %d: %8.1f clk per pass per CTA  -> %8.1f clk per line-pass per SM (%.3f ms)\n",
           name, ctas_per_sm, avg / iters, avg / iters / ctas_per_sm, ms);
}

int main() {
    cplx *tw, *out;
    long long* clk;
    cudaMalloc(&tw, 16 * 4096);
    cudaMemset(tw, 0, 16 * 4096);
    cudaMalloc(&out, 16 * 256 * 148 * 4);
    cudaMalloc(&clk, 8 * 148 * 4);
    for (int c = 1; c <= 2; ++c) {
        run_peak<0>("DFMA", c, (double*)out, clk);
        run_peak<1>("DADD", c, (double*)out, clk);
        run_peak<2>("DMUL", c, (double*)out, clk);
    }
    for (int c = 2; c <= 2; ++c) {
        run<3 | 8, false>("XF split, odd CTAs staggered", c, tw, out, clk);
        run<3 | 8, true>("XF complex, odd CTAs staggered", c, tw, out, clk);
        run<1 | 16, false>("X  split, 4 x 1024-pt groups", c, tw, out, clk);
        run<2 | 16, false>("F  4 x 1024-pt groups", c, tw, out, clk);
        run<3 | 16, false>("XF split, 4 x 1024-pt groups", c, tw, out, clk);
        run<3 | 16, true>("XF complex, 4 x 1024-pt groups", c, tw, out, clk);
        run<2 | 32, false>("F  no power chain (15 cmul)", c, tw, out, clk);
        run<3 | 32, true>("XF complex, no power chain", c, tw, out, clk);
    }
    for (int c = 1; c <= 2; ++c) {
        run<1, false>("X  exchange (re/im split)", c, tw, out, clk);
        run<1, true>("X  exchange (complex)", c, tw, out, clk);
        run<2, false>("F  twiddle powers + radix-16", c, tw, out, clk);
        run<3, false>("XF split exchange + butterfly", c, tw, out, clk);
        run<3, true>("XF complex exchange + butterfly", c, tw, out, clk);
        run<7, false>("XF split + butterfly + tw load", c, tw, out, clk);
    }
    return 0;
}
