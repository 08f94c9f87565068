// TEST TOOLING ONLY -- host emulation of the tiny slice of the CUDA runtime and
// execution model that the SwiFTly kernels use, so that the kernel *bodies*
// (index algebra, FFT passes, barriers) can be exercised by pytest in a
// container without a GPU.  Each CUDA thread of a CTA runs as a ucontext fibre;
// __syncthreads() yields to a round-robin scheduler, which reproduces barrier
// semantics exactly (one sweep = every thread advances to its next barrier).
// The product library is never built from this header.
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <vector>

struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) {
    double2 r;
    r.x = x;
    r.y = y;
    return r;
}

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind {
    cudaMemcpyHostToHost = 0,
    cudaMemcpyHostToDevice = 1,
    cudaMemcpyDeviceToHost = 2,
    cudaMemcpyDeviceToDevice = 3,
    cudaMemcpyDefault = 4
};

static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) {
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w,
                                            size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t i = 0; i < h; ++i) memmove((char*)d + i * dp, (const char*)s + i * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) {
    *d = 0;
    return cudaSuccess;
}
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }

namespace swiftly {

struct EmuBlock;

struct HostCtx {
    int tid, bid, nblocks;
    char* smem;
    const void* tmaps;
    EmuBlock* blk;
    inline void sync() const;
    // named barriers with the hardware's counting semantics (bar.sync / bar.arrive a, b):
    // the barrier completes when `count` threads have arrived; bar.sync waits for that,
    // bar.arrive does not.  A protocol error (unbalanced arrivals) shows up as a deadlock,
    // which the scheduler detects and reports instead of hanging.
    inline void group_sync(int id, int count) const;
    inline void group_arrive(int id, int count) const;
    inline void yield() const;
    // bulk asynchronous copies: the emulated copy completes immediately (races between the
    // asynchronous copy and ordinary accesses are NOT modelled -- only the index algebra is)
    inline void tx_init(uint64_t* bar) const { *bar = 0; }
    inline void tx_expect(uint64_t*, uint32_t) const {}
    inline void tx_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) const {
        memcpy(dst, src, bytes);
        *bar += 1;
    }
    inline void tx_wait(uint64_t*, uint32_t) const {}
    inline void bulk_prefetch_l2(const void*, uint32_t) const {}
    inline void nap(unsigned) const {}
    inline void fence_async() const {}
    inline void bulk_commit() const {}
    inline void bulk_wait_read() const {}
    inline void bulk_wait_all() const {}
    inline void tensor_store(const void* map, const void* smem_src, int c1, int c2, int c3) const;
    inline void tensor_load(void* smem_dst, const void* map, int c1, int c2, uint64_t* bar) const;
    // tensor memory as parking space: 128 lanes x 512 columns x 32 bit per CTA; the emulation
    // takes (lane, column) literally (no warps: the kernel's own lane arithmetic is what is
    // tested), poisons the allocation and rejects accesses outside of it
    inline uint32_t tmem_alloc(uint32_t* smem_slot, int cols) const;
    inline void tmem_free(uint32_t base, int cols) const;
    inline void tmem_st(uint32_t base, int lane, int col, double2 v) const;
    inline double2 tmem_ld(uint32_t base, int lane, int col) const;
    struct TmemLoad {
        double2 v;
    };
    inline void tmem_ld_issue(uint32_t base, int lane, int col, TmemLoad& t) const {
        t.v = tmem_ld(base, lane, col);
    }
    inline double2 tmem_ld_wait(TmemLoad& t) const { return t.v; }
    inline void tmem_ld_wait2(TmemLoad& t, TmemLoad& u, double2& a, double2& b) const {
        a = t.v;
        b = u.v;
    }
    inline void tmem_wait_st() const {}
    inline void tmem_fence_before() const {}
    inline void tmem_fence_after() const {}
};

}  // namespace swiftly
namespace swiftly {
struct TensorMap4;
void emu_tensor_store(const TensorMap4* map, const double* src, int c1, int c2, int c3);
void emu_tensor_load(const TensorMap4* map, double* dst, int c1, int c2);
inline void HostCtx::tensor_load(void* smem_dst, const void* map, int c1, int c2,
                                 uint64_t* bar) const {
    emu_tensor_load((const TensorMap4*)map, (double*)smem_dst, c1, c2);
    *bar += 1;
}
inline void HostCtx::tensor_store(const void* map, const void* smem_src, int c1, int c2,
                                  int c3) const {
    emu_tensor_store((const TensorMap4*)map, (const double*)smem_src, c1, c2, c3);
}

struct EmuBlock {
    ucontext_t main_ctx;
    std::vector<ucontext_t> ctxs;
    std::vector<char*> stacks;
    std::vector<char> done;
    int current;
    void (*entry)(void*, HostCtx&);
    void* body;
    std::vector<HostCtx> hctx;
    int nthreads;
    int bar_arrived[16];
    unsigned bar_gen[16];
    unsigned long progress;  // barrier completions + thread exits (deadlock detection)
    std::vector<uint32_t> tmem;  // 128 lanes x 512 columns
    int tmem_cols;               // allocated columns (0: none)
};

static EmuBlock* g_emu_block = nullptr;

inline void HostCtx::yield() const {
    EmuBlock* b = blk;
    swapcontext(&b->ctxs[tid], &b->main_ctx);
}

inline void HostCtx::group_arrive(int id, int count) const {
    EmuBlock* b = blk;
    if (++b->bar_arrived[id] >= count) {
        b->bar_arrived[id] = 0;
        ++b->bar_gen[id];
        ++b->progress;
    }
}

inline void HostCtx::group_sync(int id, int count) const {
    EmuBlock* b = blk;
    const unsigned g = b->bar_gen[id];
    group_arrive(id, count);
    while (b->bar_gen[id] == g) yield();
}

inline void HostCtx::sync() const { group_sync(0, blk->nthreads); }

inline uint32_t HostCtx::tmem_alloc(uint32_t* smem_slot, int cols) const {
    if (tid == 0) {
        if (cols < 32 || cols > 512 || (cols & (cols - 1)) || blk->tmem_cols) {
            fprintf(stderr, "swiftly emulator: bad TMEM allocation (%d columns)\n", cols);
            abort();
        }
        blk->tmem.assign((size_t)128 * 512, 0xA5A5A5A5u);
        blk->tmem_cols = cols;
        *smem_slot = 0;
    }
    sync();
    return *smem_slot;
}
inline void HostCtx::tmem_free(uint32_t, int cols) const {
    sync();
    if (tid == 0) {
        if (cols != blk->tmem_cols) {
            fprintf(stderr, "swiftly emulator: TMEM free of %d columns, %d allocated\n", cols,
                    blk->tmem_cols);
            abort();
        }
        blk->tmem_cols = 0;
    }
}
static inline uint32_t* emu_tmem_cell(EmuBlock* b, uint32_t base, int lane, int col) {
    if (lane < 0 || lane >= 128 || col < 0 || (col & 3) || (int)base + col + 4 > b->tmem_cols) {
        fprintf(stderr, "swiftly emulator: TMEM access out of range (lane %d, column %d of %d)\n",
                lane, (int)base + col, b->tmem_cols);
        abort();
    }
    return &b->tmem[(size_t)lane * 512 + base + col];
}
inline void HostCtx::tmem_st(uint32_t base, int lane, int col, double2 v) const {
    memcpy(emu_tmem_cell(blk, base, lane, col), &v, sizeof v);
}
inline double2 HostCtx::tmem_ld(uint32_t base, int lane, int col) const {
    double2 v;
    memcpy(&v, emu_tmem_cell(blk, base, lane, col), sizeof v);
    return v;
}

static void emu_trampoline() {
    EmuBlock* b = g_emu_block;
    int t = b->current;
    b->entry(b->body, b->hctx[t]);
    b->done[t] = 1;
    ++b->progress;
    // returning switches to uc_link (main_ctx)
}

template <class Body>
static void emu_entry(void* body, HostCtx& ctx) {
    (*(const Body*)body)(ctx);
}

template <class Body>
inline cudaError_t launch_body_impl(const Body& body, const void* tmaps, int grid, size_t smem_bytes);

template <class Body>
inline cudaError_t launch_body(const Body& body, int grid, size_t smem_bytes, cudaStream_t) {
    return launch_body_impl(body, nullptr, grid, smem_bytes);
}

template <class Body>
inline cudaError_t launch_body_maps(const Body& body, const typename Body::Maps& maps, int grid,
                                    size_t smem_bytes, cudaStream_t) {
    return launch_body_impl(body, &maps, grid, smem_bytes);
}

template <class Body>
inline cudaError_t launch_body_impl(const Body& body, const void* tmaps, int grid, size_t smem_bytes) {
    const int T = Body::THREADS;
    const size_t STACK = 256 * 1024;
    EmuBlock blk;
    blk.ctxs.resize(T);
    blk.stacks.resize(T);
    blk.done.resize(T);
    blk.hctx.resize(T);
    blk.entry = &emu_entry<Body>;
    blk.body = (void*)&body;
    for (int t = 0; t < T; ++t) blk.stacks[t] = (char*)malloc(STACK);
    char* smem = (char*)malloc(smem_bytes + 64);
    for (int bid = 0; bid < grid; ++bid) {
        memset(smem, 0xA5, smem_bytes + 64);  // poison: uninitialised smem reads show up
        for (int t = 0; t < T; ++t) {
            getcontext(&blk.ctxs[t]);
            blk.ctxs[t].uc_stack.ss_sp = blk.stacks[t];
            blk.ctxs[t].uc_stack.ss_size = STACK;
            blk.ctxs[t].uc_link = &blk.main_ctx;
            makecontext(&blk.ctxs[t], (void (*)())emu_trampoline, 0);
            blk.done[t] = 0;
            blk.hctx[t].tid = t;
            blk.hctx[t].bid = bid;
            blk.hctx[t].nblocks = grid;
            blk.hctx[t].smem = smem;
            blk.hctx[t].tmaps = tmaps;
            blk.hctx[t].blk = &blk;
        }
        g_emu_block = &blk;
        blk.nthreads = T;
        for (int i = 0; i < 16; ++i) {
            blk.bar_arrived[i] = 0;
            blk.bar_gen[i] = 0;
        }
        blk.progress = 0;
        blk.tmem_cols = 0;
        bool any = true;
        int idle_sweeps = 0;
        while (any) {
            any = false;
            const unsigned long before = blk.progress;
            for (int t = 0; t < T; ++t) {
                if (blk.done[t]) continue;
                blk.current = t;
                swapcontext(&blk.main_ctx, &blk.ctxs[t]);
                if (!blk.done[t]) any = true;
            }
            // a sweep lets every live thread run to its next wait; two sweeps in a row without
            // a barrier completing or a thread finishing means nobody can ever proceed
            idle_sweeps = (any && blk.progress == before) ? idle_sweeps + 1 : 0;
            if (idle_sweeps >= 2) {
                fprintf(stderr, "swiftly emulator: barrier DEADLOCK in block %d (arrivals:", bid);
                for (int i = 0; i < 16; ++i) fprintf(stderr, " %d", blk.bar_arrived[i]);
                fprintf(stderr, ")\n");
                abort();
            }
        }
        if (blk.tmem_cols) {
            fprintf(stderr, "swiftly emulator: block %d exits with TMEM still allocated\n", bid);
            abort();
        }
    }
    free(smem);
    for (int t = 0; t < T; ++t) free(blk.stacks[t]);
    return cudaSuccess;
}

}  // namespace swiftly
