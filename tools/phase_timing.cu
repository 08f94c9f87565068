// Dev tool: where do the cycles of one line of the fused subgrid kernel go?
// Runs SubgridAxisKernel<1024,4096,1> (8 sources) with an execution context that records
// clock64() at every barrier for one warp per CTA, prints the interval between barriers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I <csrc> tools/phase_timing.cu -o gpurun_out/phase_timing
#include <cstdio>
#include <vector>
#include <cmath>
#include "kernels.cuh"

using namespace swiftly;

#define MAXEV 512
struct TimingCtx {
    int tid, bid, nblocks;
    char* smem;
    long long* log;  // [nblocks][MAXEV]
    int* count;      // per block (only written by tid 0)
    __device__ __forceinline__ void stamp(int kind) const {
        if (tid == 0) {
            int c = *count;
            if (c < MAXEV) log[(size_t)bid * MAXEV + c] = (clock64() << 2) | kind;
            *count = c + 1;
        }
    }
    __device__ __forceinline__ void sync() const {
        stamp(0);
        __syncthreads();
        stamp(1);
    }
    __device__ __forceinline__ void group_sync(int id, int cnt) const {
        stamp(2);
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(cnt) : "memory");
        stamp(3);
    }
};

template <class Body>
__global__ void __launch_bounds__(Body::THREADS, 2) timing_entry(const Body body, long long* log, int* counts) {
    extern __shared__ __align__(16) char smem[];
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    TimingCtx ctx{(int)threadIdx.x, (int)blockIdx.x, (int)gridDim.x, smem, log, &cnt};
    body(ctx);
    if (threadIdx.x == 0) counts[blockIdx.x] = cnt;
}

int main() {
    const int M = 1024, XM = 4096, YN = 16384, NS = 8, XA = 2048;
    const int64_t n_lines = 1024;
    typedef SubgridAxisKernel<M, XM, 1> K;
    K k;
    std::vector<cplx*> src(NS);
    for (int i = 0; i < NS; ++i) {
        cudaMalloc(&src[i], sizeof(cplx) * n_lines * YN);
        cudaMemset(src[i], 0, sizeof(cplx) * n_lines * YN);
    }
    cplx* out; cudaMalloc(&out, sizeof(cplx) * n_lines * XA);
    double* fn; cudaMalloc(&fn, 8 * M); cudaMemset(fn, 0, 8 * M);
    cplx *twm, *twx; cudaMalloc(&twm, 16 * M); cudaMalloc(&twx, 16 * XM);
    cudaMemset(twm, 0, 16 * M); cudaMemset(twx, 0, 16 * XM);
    for (int i = 0; i < SW_MAX_SOURCES; ++i) { k.src[i].base = nullptr; k.src[i].ls = k.src[i].es = 0; k.src[i].wbase = k.src[i].s_m = k.src[i].sf_m = k.src[i].pos_base = 0; k.src[i].wmod = 1; }
    int order[8] = {0, 2, 4, 6, 1, 3, 5, 7};
    for (int s = 0; s < NS; ++s) {
        int j = order[s];
        k.src[s].base = src[j]; k.src[s].ls = YN; k.src[s].es = 1;
        k.src[s].wbase = 512 * 7; k.src[s].s_m = 0; k.src[s].wmod = YN;
        k.src[s].sf_m = (512 * j) % M; k.src[s].pos_base = (XM / 2 - M / 2 + 512 * j) % XM;
    }
    k.n_slots = 8; k.n_groups = 1; k.fn = fn; k.tw_m = twm; k.tw_x = twx; k.n_lines = n_lines;
    k.out = out; k.out_ls = XA; k.out_es = 1; k.out_gs = 0; k.sz = XA; k.start = 1024; k.scale = 1.0 / XM;
    k.mask = nullptr; k.first_round_tiles = 1;
    int grid = 1024;
    long long* log; int* counts;
    cudaMalloc(&log, sizeof(long long) * grid * MAXEV); cudaMalloc(&counts, sizeof(int) * grid);
    cudaFuncSetAttribute(timing_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
    for (int rep = 0; rep < 2; ++rep) {
        timing_entry<K><<<grid, K::THREADS, K::SMEM>>>(k, log, counts);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> h((size_t)grid * MAXEV); std::vector<int> hc(grid);
    cudaMemcpy(h.data(), log, sizeof(long long) * grid * MAXEV, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc.data(), counts, sizeof(int) * grid, cudaMemcpyDeviceToHost);
    // average over CTAs that are in the steady state (bid 300..700), all events of the single line
    int nev = hc[500];
    printf("events per CTA (1 line): %d\n", nev);
    std::vector<double> wait(nev, 0), work(nev, 0); std::vector<int> kind(nev, 0);
    int nb = 0;
    for (int b = 300; b < 700; ++b) {
        if (hc[b] != nev) continue;
        ++nb;
        for (int i = 0; i < nev; ++i) {
            long long t = h[(size_t)b * MAXEV + i] >> 2;
            long long tp = i ? (h[(size_t)b * MAXEV + i - 1] >> 2) : t;
            kind[i] = (int)(h[(size_t)b * MAXEV + i] & 3);
            work[i] += (double)(t - tp);
        }
    }
    double tot = 0, tw = 0, tb = 0;
    for (int i = 1; i < nev; ++i) {
        double d = work[i] / nb;
        tot += d;
        // kind 1/3 = stamp after a barrier: interval = time waiting in the barrier; 0/2 = work before it
        if (kind[i] == 1 || kind[i] == 3) tb += d; else tw += d;
        printf("%3d %s %8.0f\n", i, kind[i] == 0 ? "work->sync " : kind[i] == 1 ? "  in sync   " : kind[i] == 2 ? "work->gsync" : "  in gsync  ", d);
    }
    printf("total %.0f cycles per line (warp 0): work %.0f, inside barriers %.0f\n", tot, tw, tb);
    return 0;
}
