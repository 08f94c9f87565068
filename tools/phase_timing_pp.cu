// Dev tool: where do the cycles of one line of the two-group fused subgrid kernel go?
// Runs SubgridAxisKernelPP<1024,4096,false> (8 sources, cfg4 shapes, direct stores) with an
// execution context that records clock64() around every barrier for thread 0 of group 0,
// and prints the average interval between consecutive barrier events of a steady-state line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include \
//        -I ska_sdp_distributed_fourier_transform_b200/csrc tools/phase_timing_pp.cu -o tools/phase_timing_pp
#include <cstdio>
#include <vector>
#include "subgrid_pp.cuh"

using namespace swiftly;

#define MAXEV 4096
struct TimingCtx : DeviceCtx {
    long long* log;
    int* count;
    __device__ __forceinline__ void stamp(int kind) const {
        if (tid == 0) {
            int c = *count;
            if (c < MAXEV) log[(size_t)bid * MAXEV + c] = (clock64() << 8) | kind;
            *count = c + 1;
        }
    }
    __device__ __forceinline__ void group_sync(int id, int cnt) const {
        stamp(2 * id);
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(cnt) : "memory");
        stamp(2 * id + 1);
    }
};

template <class Body>
__global__ void __launch_bounds__(Body::THREADS, 1) timing_entry(const __grid_constant__ Body body,
                                                                 long long* log, int* counts) {
    extern __shared__ __align__(1024) char smem[];
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    TimingCtx ctx;
    ctx.tid = threadIdx.x;
    ctx.bid = blockIdx.x;
    ctx.nblocks = gridDim.x;
    ctx.smem = smem;
    ctx.tmaps = nullptr;
    ctx.log = log;
    ctx.count = &cnt;
    body(ctx);
    if (threadIdx.x == 0) counts[blockIdx.x] = cnt;
}

int main() {
    const int M = 1024, XM = 4096, YN = 16384, NS = 8, XA = 2048, ITER = 8;
    const int grid = 148;
    const int64_t n_lines = (int64_t)grid * 2 * ITER;
    typedef SubgridAxisKernelPP<M, XM, false> K;
    static K k;
    std::vector<cplx*> src(NS);
    for (int i = 0; i < NS; ++i) {
        cudaMalloc(&src[i], sizeof(cplx) * n_lines * YN);
        cudaMemset(src[i], 0, sizeof(cplx) * n_lines * YN);
    }
    cplx* out;
    cudaMalloc(&out, sizeof(cplx) * n_lines * XA);
    double* fn;
    cudaMalloc(&fn, 8 * M);
    cudaMemset(fn, 0, 8 * M);
    cplx *twm, *twx;
    cudaMalloc(&twm, 16 * M);
    cudaMalloc(&twx, 16 * XM);
    cudaMemset(twm, 0, 16 * M);
    cudaMemset(twx, 0, 16 * XM);
    memset(&k, 0, sizeof(k));
    for (int i = 0; i < SW_MAX_SOURCES; ++i) k.src[i].wmod = 1;
    int order[8] = {0, 2, 4, 6, 1, 3, 5, 7};
    for (int s = 0; s < NS; ++s) {
        int j = order[s];
        k.src[s].base = src[j];
        k.src[s].ls = YN;
        k.src[s].es = 1;
        k.src[s].wbase = 512 * 7;
        k.src[s].s_m = 0;
        k.src[s].wmod = YN;
        k.src[s].sf_m = (512 * j) % M;
        k.src[s].pos_base = (XM / 2 - M / 2 + 512 * j) % XM;
    }
    k.n_slots = 8;
    k.n_groups = 1;
    k.fn = fn;
    k.tw_m = twm;
    k.tw_x = twx;
    k.n_lines = n_lines;
    k.out = out;
    k.out_ls = XA;
    k.out_es = 1;
    k.sz = XA;
    k.start[0] = 1024;
    k.scale = 1.0 / XM;
    k.first_round_tiles = 1;
    k.stagger_ns = 0;
    k.cx_round0 = 0;
    long long* log;
    int* counts;
    cudaMalloc(&log, sizeof(long long) * grid * MAXEV);
    cudaMalloc(&counts, sizeof(int) * grid);
    cudaFuncSetAttribute(timing_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
    for (int rep = 0; rep < 2; ++rep) {
        timing_entry<K><<<grid, K::THREADS, K::SMEM>>>(k, log, counts);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> h((size_t)grid * MAXEV);
    std::vector<int> hc(grid);
    cudaMemcpy(h.data(), log, sizeof(long long) * grid * MAXEV, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc.data(), counts, sizeof(int) * grid, cudaMemcpyDeviceToHost);
    const int per_line = hc[10] / ITER;
    printf("events per CTA %d, per line %d\n", hc[10], per_line);
    std::vector<double> dt(per_line, 0);
    std::vector<int> kind(per_line, 0);
    int nb = 0;
    for (int b = 0; b < grid; ++b) {
        if (hc[b] != hc[10]) continue;
        for (int it = 3; it < 6; ++it) {  // steady-state lines
            ++nb;
            for (int i = 0; i < per_line; ++i) {
                size_t e = (size_t)b * MAXEV + (size_t)it * per_line + i;
                dt[i] += (double)((h[e] >> 8) - (h[e - 1] >> 8));
                kind[i] = (int)(h[e] & 255);
            }
        }
    }
    double tot = 0, inbar = 0;
    for (int i = 0; i < per_line; ++i) {
        double d = dt[i] / nb;
        tot += d;
        const int id = kind[i] / 2, after = kind[i] & 1;
        if (after) inbar += d;
        printf("%3d %s barrier %2d %8.0f\n", i, after ? "   inside" : "work  ->", id, d);
    }
    printf("total %.0f cycles per line (thread 0 of group 0): outside barriers %.0f, inside %.0f\n",
           tot, tot - inbar, inbar);
    return 0;
}
