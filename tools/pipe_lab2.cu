// Dev tool: 16 points per thread (128 registers, 512 threads per SM) against 8 points per thread
// (64 registers, 1024 threads per SM) for a complete 4096-point line through shared memory.
//   P16: radix 16,16,16  -> 3 butterfly phases, 2 exchanges, 256 threads per line
//   P8 : radix 8,8,8,8   -> 4 butterfly phases, 3 exchanges, 512 threads per line
// Two lines (CTAs) per SM in both cases; prints cycles per line per SM.
#include <cstdio>
#include <vector>
#include "fft_engine.cuh"

using namespace swiftly;

template <int P, bool CX>
__global__ void __launch_bounds__(4096 / P, 2) lab(const cplx* tw, cplx* out, int iters, long long* clk) {
    extern __shared__ __align__(16) char smem[];
    double* sm = (double*)smem;
    cplx* smc = (cplx*)smem;
    constexpr int N = 4096, T = N / P, NB = N / P;
    constexpr int PASSES = P == 16 ? 3 : 4;
    const int lt = threadIdx.x;
    cplx v[P];
#pragma unroll
    for (int r = 0; r < P; ++r) v[r] = mk(1.0 + 1e-9 * (lt + r), 1e-9 * (lt - r));
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int pass = 0; pass < PASSES; ++pass) {
            if (pass > 0) {
                const int base = lt * P;  // any conflict-free pattern of the right shape
                if (CX) {
#pragma unroll
                    for (int r = 0; r < P; ++r) smc[sm_phys(base + r)] = v[r];
                    __syncthreads();
#pragma unroll
                    for (int r = 0; r < P; ++r) v[r] = smc[sm_phys(lt + r * NB)];
                    __syncthreads();
                } else {
                    double nx[P];
#pragma unroll
                    for (int r = 0; r < P; ++r) sm[sm_phys(base + r)] = v[r].x;
                    __syncthreads();
#pragma unroll
                    for (int r = 0; r < P; ++r) nx[r] = sm[sm_phys(lt + r * NB)];
                    __syncthreads();
#pragma unroll
                    for (int r = 0; r < P; ++r) sm[sm_phys(base + r)] = v[r].y;
                    __syncthreads();
#pragma unroll
                    for (int r = 0; r < P; ++r) v[r] = mk(nx[r], sm[sm_phys(lt + r * NB)]);
                    __syncthreads();
                }
                cplx w1 = ldg_c(tw + (lt & 255));
                w1.x += 0.9999999;
                w1.y += 1e-4;
                TwiddlePowers<P>::apply(v, w1);
            }
            Radix<P, +1>::run(v);
#pragma unroll
            for (int r = 0; r < P; ++r) v[r] = cscale(v[r], 0.25);
        }
    }
    long long t1 = clock64();
    cplx s = mk(0, 0);
#pragma unroll
    for (int r = 0; r < P; ++r) s = cadd(s, v[r]);
    out[blockIdx.x * T + lt] = s;
    if (lt == 0) clk[blockIdx.x] = t1 - t0;
}

template <int P, bool CX>
void run(const char* name, const cplx* tw, cplx* out, long long* clk) {
    const int iters = 500, ctas = 2;
    size_t smem = 100 * 1024;
    cudaFuncSetAttribute(lab<P, CX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = 148 * ctas;
    lab<P, CX><<<grid, 4096 / P, smem>>>(tw, out, 10, clk);
    lab<P, CX><<<grid, 4096 / P, smem>>>(tw, out, iters, clk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto c : h) avg += (double)c;
    avg /= grid;
    printf("%-44s %8.1f clk per 4096-point line per SM (2 lines in flight)\n", name, avg / iters / ctas);
}

int main() {
    cplx *tw, *out;
    long long* clk;
    cudaMalloc(&tw, 16 * 4096);
    cudaMemset(tw, 0, 16 * 4096);
    cudaMalloc(&out, 16 * 512 * 148 * 2);
    cudaMalloc(&clk, 8 * 148 * 2);
    run<16, false>("P16 (256 thr/line, 128 regs) split exchange", tw, out, clk);
    run<16, true>("P16 (256 thr/line, 128 regs) complex exchange", tw, out, clk);
    run<8, false>("P8  (512 thr/line,  64 regs) split exchange", tw, out, clk);
    run<8, true>("P8  (512 thr/line,  64 regs) complex exchange", tw, out, clk);
    return 0;
}
