// Dev tool: where do the cycles of one line of K2 (extract_columns, yN = 16384) go?
// Runs the TMEM kernel (extract_tmem.cuh, MODE given on the command line: 0 DIF, 2 DIT) or the
// default 4 x 4096 kernel (mode 9) on 8 facets x 1024 lines of pre-windowed rows with an
// execution context that records clock64() around every barrier and the row wait, for thread 0
// of group 0 and of group 1, and prints the average intervals of steady-state lines.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include \
//        -I ska_sdp_distributed_fourier_transform_b200/csrc tools/phase_timing_k2.cu \
//        ska_sdp_distributed_fourier_transform_b200/csrc/tensor_map.cu -o tools/phase_timing_k2
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "extract_tma.cuh"
#include "extract_tmem.cuh"

using namespace swiftly;
namespace swiftly {
bool make_row_map(TensorMap4* tm, const cplx* base, int64_t ls, int64_t n_rows, int64_t fs,
                  int box_chunks);
}

#define MAXEV 2048
// event kinds: 2 * id (+1 after) for barrier id (0 = CTA); 40 / 41 around the row wait
struct TimingCtx : DeviceCtx {
    long long* log;
    int* count;  // [2] in shared memory
    int tg_threads;
    __device__ __forceinline__ void stamp(int kind) const {
        if (tid % tg_threads == 0) {
            const int g = tid / tg_threads;
            int c = count[g];
            if (c < MAXEV) log[((size_t)bid * 2 + g) * MAXEV + c] = (clock64() << 8) | kind;
            count[g] = c + 1;
        }
    }
    __device__ __forceinline__ void sync() const {
        stamp(0);
        __syncthreads();
        stamp(1);
    }
    __device__ __forceinline__ void group_sync(int id, int cnt) const {
        stamp(2 * id);
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(cnt) : "memory");
        stamp(2 * id + 1);
    }
    __device__ __forceinline__ void group_arrive(int id, int cnt) const {
        stamp(2 * id);
        asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(cnt) : "memory");
        stamp(2 * id + 1);
    }
    __device__ __forceinline__ void tx_wait(uint64_t* bar, uint32_t parity) const {
        stamp(40);
        DeviceCtx::tx_wait(bar, parity);
        stamp(41);
    }
    __device__ __forceinline__ uint32_t tmem_alloc(uint32_t* slot, int cols) const {
        return DeviceCtx::tmem_alloc(slot, cols);  // (its barrier is not part of a line)
    }
    __device__ __forceinline__ void tmem_free(uint32_t base, int cols) const {
        DeviceCtx::tmem_free(base, cols);
    }
};

template <class Body>
__global__ void __launch_bounds__(Body::THREADS, 1)
    timing_entry(const Body body, const __grid_constant__ typename Body::Maps maps, long long* log,
                 int* counts) {
    extern __shared__ __align__(1024) char smem[];
    __shared__ int cnt[2];
    if (threadIdx.x == 0) cnt[0] = cnt[1] = 0;
    __syncthreads();
    TimingCtx ctx;
    ctx.tid = threadIdx.x;
    ctx.bid = blockIdx.x;
    ctx.nblocks = gridDim.x;
    ctx.smem = smem;
    ctx.tmaps = &maps;
    ctx.log = log;
    ctx.count = cnt;
    ctx.tg_threads = Body::THREADS / 2;
    body(ctx);
    __syncthreads();
    if (threadIdx.x == 0) {
        counts[2 * blockIdx.x] = cnt[0];
        counts[2 * blockIdx.x + 1] = cnt[1];
    }
}

template <class K>
int run(int scratch_lines) {
    const int Q = 4096, YN = 4 * Q, FS = 8192, M = 1024, NF = 8;
    const int grid = 148;
    static K k;
    static typename K::Maps maps;
    memset(&k, 0, sizeof(k));
    std::vector<cplx*> in(NF), out(NF);
    for (int f = 0; f < NF; ++f) {
        cudaMalloc(&in[f], sizeof(cplx) * (size_t)YN * FS);
        cudaMemset(in[f], 0, sizeof(cplx) * (size_t)YN * FS);
        cudaMalloc(&out[f], sizeof(cplx) * (size_t)M * YN);
    }
    cplx *tw, *twf, *scratch = nullptr;
    cudaMalloc(&tw, 16 * Q);
    cudaMemset(tw, 0, 16 * Q);
    cudaMalloc(&twf, 16 * (YN / 2));
    cudaMemset(twf, 0, 16 * (YN / 2));
    if (scratch_lines) cudaMalloc(&scratch, sizeof(cplx) * (size_t)grid * scratch_lines * Q);
    k.op.g.n_lines = (int64_t)NF * M;
    k.op.fb = nullptr;  // pre-windowed rows
    k.op.n = YN;
    k.op.lines_per = M;
    k.op.scale = 1.0 / YN;
    k.op.rm_s_m = 0;
    k.op.rm_base = 3000;
    k.tw = tw;
    k.twf = twf;
    k.scratch = scratch;
    k.swizzled = 1;
    k.box_chunks = 256;
    k.in_cap = FS;
    for (int f = 0; f < NF; ++f) {
        k.op.fac[f].in = in[f];
        k.op.fac[f].out = out[f];
        k.op.fac[f].in_ls = FS;
        k.op.fac[f].out_ls = YN;
        k.op.fac[f].fs = FS;
        k.op.fac[f].shift_in = 4096;
        k.op.fac[f].fb_off = 0;
        if (!make_row_map(&maps.in_map[f], in[f], FS, YN, FS, k.box_chunks)) {
            printf("make_row_map failed\n");
            return 1;
        }
    }
    long long* log;
    int* counts;
    cudaMalloc(&log, sizeof(long long) * grid * 2 * MAXEV);
    cudaMalloc(&counts, sizeof(int) * grid * 2);
    const size_t smem = K::smem_bytes(k.in_cap);
    cudaFuncSetAttribute(timing_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        timing_entry<K><<<grid, K::THREADS, smem>>>(k, maps, log, counts);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("error %s\n", cudaGetErrorString(e));
            return 1;
        }
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel (instrumented) %.3f ms for %d lines\n", ms, NF * M);
    std::vector<long long> h((size_t)grid * 2 * MAXEV);
    std::vector<int> hc(grid * 2);
    cudaMemcpy(h.data(), log, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(hc.data(), counts, sizeof(int) * hc.size(), cudaMemcpyDeviceToHost);
    const int lines_cta = (NF * M) / grid;  // CTAs with the minimum number of lines
    for (int g = 0; g < 2; ++g) {
        // a CTA that ran exactly `lines_cta` lines: events per line
        int ref = -1;
        for (int b = grid - 1; b >= 0 && ref < 0; --b)
            if (hc[2 * b + g] % lines_cta == 0 && hc[2 * b + g] / lines_cta > 0) ref = b;
        if (ref < 0) {
            printf("no reference CTA (counts %d)\n", hc[g]);
            continue;
        }
        const int per_line = hc[2 * ref + g] / lines_cta;
        printf("group %d: events per CTA %d, per line %d\n", g, hc[2 * ref + g], per_line);
        std::vector<double> dt(per_line, 0);
        std::vector<int> kind(per_line, 0);
        int nb = 0;
        for (int b = 0; b < grid; ++b) {
            if (hc[2 * b + g] != hc[2 * ref + g]) continue;
            for (int it = 10; it < 40; ++it) {  // steady-state lines
                ++nb;
                for (int i = 0; i < per_line; ++i) {
                    size_t e = ((size_t)b * 2 + g) * MAXEV + (size_t)it * per_line + i;
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
            printf("%3d %s %s %2d %8.0f\n", i, after ? "   inside" : "work  ->",
                   id == 20 ? "row wait" : "barrier ", id, d);
        }
        printf("group %d total %.0f cycles per line: outside waits %.0f, inside %.0f\n", g, tot,
               tot - inbar, inbar);
    }
    return 0;
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 2;
    if (mode == 0) return run<ExtractColumnsTmemKernel<4096, 0>>(0);
    if (mode == 2) return run<ExtractColumnsTmemKernel<4096, 2>>(0);
    if (mode == 3) return run<ExtractColumnsTmemSkewKernel<4096>>(0);
    return run<ExtractColumnsTma4Kernel<4096>>(4);
}
