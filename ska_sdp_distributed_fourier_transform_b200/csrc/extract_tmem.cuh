// SwiFTly B200 -- K2 ("Fb . FFT . extract", api_helper.py:200-210 = extract_from_facet(axis 0)
// + prepare_facet(axis 1), core.py:189-253) for yN = 4 Q with the intermediate results parked
// in TENSOR MEMORY instead of an L2 scratch.
//
// Same decomposition as ExtractColumnsTmaDifKernel (extract_tma.cuh): the TMA-staged facet
// row is transformed by two independent thread groups,
//
//   group g (0: even outputs, 1: odd outputs):   y_g[j] = (z[j] + (-1)^g z[j + 2Q]) w^(g j),
//       j < 2Q, w = exp(+2 pi i / yN);  X[2 k + g] = FFT_2Q(y_g)[k]
//   inside the group:  E0 = FFT_Q(y_g[2 j]),  E1 = FFT_Q(y_g[2 j + 1]),  v = exp(+2 pi i / 2Q)
//       FFT_2Q(y_g)[k] = E0[k] + v^k E1[k],   FFT_2Q(y_g)[k + Q] = E0[k] - v^k E1[k]
//
// but nothing goes through global memory between the row and the finished line:
//   * E0 waits in TMEM.  A thread owns the same 16 output indices k in both sub-transforms, so
//     its E0[k] are thread-private: 16 tcgen05.st after the first sub-transform, 16 tcgen05.ld
//     in the last pass of the second (the DIF kernel parks them in an L2 scratch line: 128 KiB
//     of global stores and loads per line on the critical path of the one CTA an SM holds);
//   * the two groups swap halves THROUGH TMEM before storing.  Thread t of group 0 and thread t
//     of group 1 sit in warps of the same TMEM lane quarter with the same lane number and own
//     the same k: group 0 parks X[2 (k + Q)], group 1 parks X[2 k + 1]; after one CTA barrier
//     group 0 stores the pairs (X[2 k], X[2 k + 1]), group 1 the pairs (X[2 (k + Q)],
//     X[2 (k + Q) + 1]) -- 32 contiguous bytes per thread, 1 KiB per warp instruction (STG.256)
//     -- where the DIF kernel stores 16-byte samples at a 32-byte stride (the reason it lost);
//   * no scratch, no CTA-wide combine phase: one CTA barrier per line (the swap), the TMEM
//     regions are double buffered by line parity so that nobody waits for a reader.
//
// TMEM budget at Q = 4096 (256 threads per group): per thread 16 samples = 64 columns of its
// lane, x 2 threads of a group per lane (warps w and w + 4), x 2 groups, x 2 buffers = 512
// columns = all of it (256 KiB).  No tensor-core instruction is issued; tcgen05.alloc / st / ld
// only (SASS: UTCATOMSWS, STTM, LDTM).
#pragma once

#include "extract_tma.cuh"
#include "kernels.cuh"

namespace swiftly {

// MODE 0 / 1: the DIF-across form above (1: facets longer than yN/2 exist, both z[j] and
// z[j + 2Q] may be non-zero).  MODE 2: decimation in TIME across the groups as well,
//   group g:  u_g[i] = z[2 i + g], i < 2Q;  U_g = FFT_2Q(u_g) (E0 / E1 as above, from z[4 j + g]
//   and z[4 j + 2 + g]);  X[k] = U_0[k] + w^k U_1[k],  X[k + 2Q] = U_0[k] - w^k U_1[k],  k < 2Q;
// the groups swap so that group 0 combines k < Q (it gets U_1[k]) and group 1 combines Q <= k <
// 2Q (it gets U_0[k]).  Every staged sample is read by ONE group (the DIF form reads it twice),
// no twiddle at load time (group 1 of the DIF form multiplies every sample), the work of the
// two groups is equal, and the outputs are plain unit-stride 16-byte streams.
template <int Q, int MODE>
struct ExtractColumnsTmemKernel {
    static constexpr bool DIT = MODE == 2;
    static constexpr bool BOTH = MODE == 1;
    static constexpr int DIR = +1;
    static constexpr int TG = FftCfg<Q>::T;  // threads per group
    static constexpr int THREADS = 2 * TG;
    static constexpr int H = 2 * Q;
    static constexpr int N = 4 * Q;
    static constexpr int XBUF = (FftCfg<Q>::PADDED + 1) & ~1;  // doubles per exchange buffer
    static constexpr int TLANES = TG < 128 ? TG : 128;         // TMEM lanes a group covers
    static constexpr int HALVES = TG / TLANES;                 // threads of a group per lane
    static constexpr int TCOLS = 64;                           // columns per thread: 16 samples
    static constexpr int BUF_COLS = 2 * HALVES * TCOLS;        // one buffer: both groups
    static constexpr int NCOLS = 2 * BUF_COLS;                 // double buffered
    static_assert(NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM holds 512 columns");
    static_assert(N % 4 == 0, "pairs of outputs must stay adjacent after the centring rotation");
#if !defined(SWIFTLY_EMU)
    static_assert(TG % 128 == 0, "a group must cover whole TMEM lane sets (4 warps)");
#endif
    static constexpr size_t smem_bytes(int in_cap) {
        return (size_t)in_cap * sizeof(cplx) + 2 * (size_t)XBUF * sizeof(double) + 32;
    }

    ExtractColumnsOp op;
    const cplx* tw;   // compact table of the Q-point plan
    const cplx* twf;  // exp(-2 pi i t / yN), t < yN / 2
    cplx* scratch;    // unused (launch helper compatibility)
    int in_cap;
    int swizzled;
    int box_chunks;
    struct Maps {
        TensorMap4 in_map[SW_MAX_COLUMN_FACETS];
    };

    SW_HD cplx root(int t) const {  // exp(DIR 2 pi i t / N), 0 <= t < N
        const bool neg = t >= N / 2;
        cplx w = ldg_c(twf + (neg ? t - N / 2 : t));
        if (DIR > 0) w.y = -w.y;
        return neg ? mk(-w.x, -w.y) : w;
    }

    template <class Ctx>
    SW_HD void issue(const Ctx& ctx, cplx* in, uint64_t* bar, int64_t line) const {
        k2_issue_row<Maps>(ctx, op, swizzled, box_chunks, in, bar, line);
    }

    // group barrier; after the first-pass loads of the group's LAST sub-transform the staging
    // buffer is dead for this group: the second group to get there starts the next row's copy
    template <class Ctx>
    struct GroupSync {
        const Ctx& ctx;
        const ExtractColumnsTmemKernel& k;
        int grp, tg;
        cplx* in;
        uint64_t* bar;
        int* done;
        int64_t next_line;
        bool pending;
        SW_HD void operator()() {
            ctx.group_sync(1 + grp, TG);
            if (pending) {
                pending = false;
                if (tg == 0) {
#if defined(__CUDA_ARCH__)
                    const int prev = atomicAdd(done, 1);
#else
                    const int prev = (*done)++;
#endif
                    if ((prev & 1) == 1 && next_line < k.op.g.n_lines)
                        k.issue(ctx, in, bar, next_line);
                }
            }
        }
    };

    template <class Ctx>
    SW_HD void operator()(Ctx& ctx) const {
        cplx* in = (cplx*)ctx.smem;
        double* xb = (double*)(in + in_cap);
        uint64_t* bar = (uint64_t*)(xb + 2 * XBUF);
        int* done = (int*)(bar + 1);
        uint32_t* tslot = (uint32_t*)(done + 1);
        const int grp = ctx.tid / TG;
        const int tg = ctx.tid % TG;
        double* sm = xb + (size_t)grp * XBUF;
        const int n = op.n;
        // TMEM: lane of this thread, column regions of this thread and of its partner (the
        // thread with the same tg in the other group: same lane, same half)
        const int tl = tg % TLANES;
        const int half = tg / TLANES;
        const int my_col = (grp * HALVES + half) * TCOLS;
        const int partner_col = ((1 - grp) * HALVES + half) * TCOLS;
        if (ctx.tid == 0) {
            ctx.tx_init(bar);
            *done = 0;
            if ((int64_t)ctx.bid < op.g.n_lines) issue(ctx, in, bar, ctx.bid);
        }
        const uint32_t tbase = ctx.tmem_alloc(tslot, NCOLS);  // (contains the CTA barrier)
        uint32_t parity = 0;
        int buf_col = 0;
        typedef LastPass<Q> LP;
        for (int64_t line = ctx.bid; line < op.g.n_lines; line += ctx.nblocks) {
            const int f = (int)(line / op.lines_per);
            const int l = (int)(line - (int64_t)f * op.lines_per);
            const ColumnFacet& F = op.fac[f];
            const int shift_in = F.shift_in, fs = F.fs;
            const double* fb = op.fb ? op.fb + F.fb_off : nullptr;
            cplx* o = F.out + (int64_t)l * F.out_ls;
            const double scale = op.scale;
            const bool swz = swizzled != 0;
            auto sample = [&](int q) {  // natural-order sample q of the padded, rotated row
                int k = wrap_add(q, shift_in, n);
                if (k >= fs) return mk(0.0, 0.0);
                const int ks = swz ? ((k & ~7) | ((k ^ (k >> 3)) & 7)) : k;
                return fb ? cscale(in[ks], ldg_d(fb + k)) : in[ks];
            };
            // y_g[q], q < 2Q; group 1's twiddle w^q: four table loads per thread and
            // sub-transform (calls 0, 4, 8, 12), in between three multiplications by w^(2 Q/16)
            const cplx w_step = root((2 * (Q / 16)) % N);
            cplx w_ld = mk(1.0, 0.0);
            int calls = 0;
            auto y = [&](int q) {
                if constexpr (DIT) {
                    return sample(2 * q + grp);
                } else {
                    cplx v;
                    if constexpr (BOTH) {
                        cplx a = sample(q), b = sample(q + H);
                        v = grp ? csub(a, b) : cadd(a, b);
                    } else {
                        // fs <= 2Q: at most one of z[q], z[q + 2Q] is inside the facet
                        int k = wrap_add(q, shift_in, n);
                        const bool second = k >= fs;
                        if (second) k = wrap_add(k, H, n);
                        if (k >= fs) {
                            v = mk(0.0, 0.0);
                        } else {
                            const int ks = swz ? ((k & ~7) | ((k ^ (k >> 3)) & 7)) : k;
                            v = fb ? cscale(in[ks], ldg_d(fb + k)) : in[ks];
                            if (second && grp) v = mk(-v.x, -v.y);
                        }
                    }
                    if (grp) {
                        w_ld = (calls & 3) == 0 ? root(q % N) : cmul(w_ld, w_step);
                        v = cmul(v, w_ld);
                    }
                    ++calls;
                    return v;
                }
            };
            const int col0 = buf_col + my_col;
            ctx.tx_wait(bar, parity);  // this line's row has landed
            parity ^= 1;
            GroupSync<Ctx> gs{ctx, *this, grp, tg, in, bar, done, line + ctx.nblocks, false};
            {   // E0 = FFT_Q(y_g[2 j]) -> TMEM
                auto ld = [&](int j) { return y(2 * j); };
                auto st = [&](int, cplx v, int it, int r) {
                    ctx.tmem_st(tbase, tl, col0 + 4 * (it * LP::R + r), v);
                };
                line_fft<Q, DIR>(tg, sm, tw, ld, st, gs);
                ctx.tmem_wait_st();
                gs();  // the group's exchange buffer is reused by its second sub-transform
            }
            cplx keep[16];  // the half of the group's results that this thread stores itself
            {   // E1 = FFT_Q(y_g[2 j + 1]); combine with E0; park the half the partner stores
                calls = 0;
                auto ld = [&](int j) { return y(2 * j + 1); };
                // v^k = root(2 k) by recurrence over the thread's outputs k = j0 + it T + r NS
                const cplx v_it = root((2 * TG) % N), v_r = root((2 * LP::NS) % N);
                cplx w_it = mk(1.0, 0.0), w = mk(1.0, 0.0);
                auto st = [&](int k, cplx od, int it, int r) {
                    if (r == 0) {
                        w_it = it == 0 ? root((2 * k) % N) : cmul(w_it, v_it);
                        w = w_it;
                    } else {
                        w = cmul(w, v_r);
                    }
                    const int s = it * LP::R + r;
                    const cplx e = ctx.tmem_ld(tbase, tl, col0 + 4 * s);
                    const cplx wo = cmul(od, w);
                    // DIF: FFT_2Q(y_g)[k] = X[2 k + g], FFT_2Q(y_g)[k + Q] = X[2 (k + Q) + g];
                    // DIT: U_g[k], U_g[k + Q].  Group 0 keeps lo and parks hi, group 1 vice versa
                    const cplx lo = cscale(cadd(e, wo), scale);
                    const cplx hi = cscale(csub(e, wo), scale);
                    keep[s] = grp ? hi : lo;
                    ctx.tmem_st(tbase, tl, col0 + 4 * s, grp ? lo : hi);
                };
                gs.pending = true;
                line_fft<Q, DIR>(tg, sm, tw, ld, st, gs);
                ctx.tmem_wait_st();
            }
            // swap: everybody's parked half is visible to the partner after the barrier
            ctx.tmem_fence_before();
            ctx.sync();
            ctx.tmem_fence_after();
            if constexpr (DIT) {
                // group 0: k' = k (a = own U_0[k], b = U_1[k]); group 1: k' = k + Q (a = U_0[k + Q],
                // b = own U_1[k + Q]);  X[k'] = a + w^k' b, X[k' + 2Q] = a - w^k' b
                const int pcol = buf_col + partner_col;
                const cplx w_r = root(LP::NS % N);
                cplx w = mk(1.0, 0.0);
#pragma unroll
                for (int it = 0; it < LP::ITERS; ++it) {
                    const int j = tg + it * TG;
                    const int base = (j / LP::NS) * (LP::NS * LP::R) + (j & (LP::NS - 1));
#pragma unroll
                    for (int r = 0; r < LP::R; ++r) {
                        const int s = it * LP::R + r;
                        const int kk = base + r * LP::NS + (grp ? Q : 0);
                        w = r == 0 ? root(kk % N) : cmul(w, w_r);
                        const cplx p = ctx.tmem_ld(tbase, tl, pcol + 4 * s);
                        const cplx a = grp ? p : keep[s];
                        const cplx wb = cmul(grp ? keep[s] : p, w);
                        st_stream(o + wrap_add(kk, n / 2, n), cadd(a, wb));
                        st_stream(o + wrap_add(kk + H, n / 2, n), csub(a, wb));
                    }
                }
            } else {
                const int pcol = buf_col + partner_col;
#pragma unroll
                for (int it = 0; it < LP::ITERS; ++it) {
                    const int j = tg + it * TG;
                    const int base = (j / LP::NS) * (LP::NS * LP::R) + (j & (LP::NS - 1));
#pragma unroll
                    for (int r = 0; r < LP::R; ++r) {
                        const int s = it * LP::R + r;
                        const int k = base + r * LP::NS;
                        const cplx p = ctx.tmem_ld(tbase, tl, pcol + 4 * s);
                        // group 0: (X[2 k], X[2 k + 1]); group 1: (X[2 (k + Q)], X[2 (k + Q) + 1])
                        const int pc = wrap_add(2 * k + (grp ? H : 0), n / 2, n);
                        st_stream_pair(o + pc, grp ? p : keep[s], grp ? keep[s] : p);
                    }
                }
            }
            buf_col ^= BUF_COLS;
        }
        ctx.tmem_free(tbase, NCOLS);
    }
};

// ---------------------------------------------------------------------------------------
// The DIT form (MODE 2) with the two groups' STORE phases skewed by half a line.
//
// tools/phase_timing_k2.cu on the MODE 2 kernel: of 37.5 k cycles per line, 10 k are the store
// phase after the swap -- 256 KiB of 16-byte stores per line drain at about 26 bytes per clock
// and SM, and BOTH groups sit in that phase at the same time (they have just met at the swap
// barrier), so nothing else runs on the SM meanwhile.  Here group 0 stores its half of line L
// right after the swap barrier (as before, from registers), but group 1 parks its kept half in
// TMEM as well, goes straight on to line L + 1 and stores its half of line L between the two
// sub-transforms of line L + 1: each group's store burst runs under the other group's
// butterflies and exchanges; both groups still do the same amount of work between two swap
// barriers.
//
// TMEM columns per lane (512): group 0 threads 2 x 64 (E0 / parked half, double buffered by line
// parity: group 1 reads the parked half of line L while group 0 is already parking E0 of line
// L + 1), group 1 threads 64 (E0 / parked half) + 64 (kept half).  Group 1's parked half is read
// by group 0 right after the swap barrier and overwritten by group 1's E0 about a sub-transform
// later: ordered by a bar.arrive (group 0, after its loads) / bar.sync (group 1, before the
// stores of its first sub-transform's last pass) pair on named barrier 3.
template <int Q>
struct ExtractColumnsTmemSkewKernel : ExtractColumnsTmemKernel<Q, 2> {
    typedef ExtractColumnsTmemKernel<Q, 2> Base;
    static constexpr int DIR = Base::DIR, TG = Base::TG, THREADS = Base::THREADS, H = Base::H,
                         N = Base::N, XBUF = Base::XBUF, TLANES = Base::TLANES,
                         HALVES = Base::HALVES, TCOLS = Base::TCOLS;
    static constexpr int NCOLS = 4 * HALVES * TCOLS;  // 2 groups x (2 regions of 64 columns)
    static_assert(NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM holds 512 columns");
    typedef typename Base::Maps Maps;

    template <class Ctx>
    struct SkewSync {
        const Ctx& ctx;
        const ExtractColumnsTmemSkewKernel& k;
        int grp, tg;
        cplx* in;
        uint64_t* bar;
        int* done;
        int64_t next_line;
        bool pending;      // the next barrier follows the group's last first-pass load of the row
        bool wait_reader;  // group 1: the partner must have read the parked half of the last line
        SW_HD void operator()() {
            ctx.group_sync(1 + grp, TG);
            if (pending) {
                pending = false;
                if (tg == 0) {
#if defined(__CUDA_ARCH__)
                    const int prev = atomicAdd(done, 1);
#else
                    const int prev = (*done)++;
#endif
                    if ((prev & 1) == 1 && next_line < k.op.g.n_lines)
                        k.issue(ctx, in, bar, next_line);
                }
            }
        }
        SW_HD void pre_store() {
            if (wait_reader) {
                wait_reader = false;
                ctx.group_sync(3, THREADS);
                ctx.tmem_fence_after();
            }
        }
    };

    template <class Ctx>
    SW_HD void operator()(Ctx& ctx) const {
        cplx* in = (cplx*)ctx.smem;
        double* xb = (double*)(in + this->in_cap);
        uint64_t* bar = (uint64_t*)(xb + 2 * XBUF);
        int* done = (int*)(bar + 1);
        uint32_t* tslot = (uint32_t*)(done + 1);
        const int grp = ctx.tid / TG;
        const int tg = ctx.tid % TG;
        double* sm = xb + (size_t)grp * XBUF;
        const int n = this->op.n;
        const int tl = tg % TLANES;
        const int half = tg / TLANES;
        // group 0: regions [half * 128 + 64 * parity]; group 1: parked [.. + 0], kept [.. + 64]
        const int g0_col = half * 2 * TCOLS;
        const int g1_col = (HALVES + half) * 2 * TCOLS;
        if (ctx.tid == 0) {
            ctx.tx_init(bar);
            *done = 0;
            if ((int64_t)ctx.bid < this->op.g.n_lines) this->issue(ctx, in, bar, ctx.bid);
        }
        const uint32_t tbase = ctx.tmem_alloc(tslot, NCOLS);  // (contains the CTA barrier)
        uint32_t parity = 0;
        int par_col = 0;  // 0 / TCOLS: group 0's region of the current line
        typedef LastPass<Q> LP;
        cplx keep[16];         // group 0: the kept half U_0[k] of the line just finished
        cplx* o_prev = nullptr;  // output line whose stores are still due
        // X[k'] = a + w^k' b, X[k' + 2Q] = a - w^k' b for the thread's sixteen k' of the previous
        // line; group 0: k' = k, a = kept registers, b = group 1's parked half; group 1: k' = k + Q,
        // a = group 0's parked half, b = own kept half (TMEM)
        auto store_prev = [&](int a_col, int b_col) {
            const cplx w_r = this->root(LP::NS % N);
            cplx w = mk(1.0, 0.0);
            // the TMEM loads of slot s + 1 are in flight while slot s is combined and stored
            typename Ctx::TmemLoad la, lb;
            if (grp) ctx.tmem_ld_issue(tbase, tl, a_col, la);
            ctx.tmem_ld_issue(tbase, tl, b_col, lb);
#pragma unroll
            for (int it = 0; it < LP::ITERS; ++it) {
                const int j = tg + it * TG;
                const int base = (j / LP::NS) * (LP::NS * LP::R) + (j & (LP::NS - 1));
#pragma unroll
                for (int r = 0; r < LP::R; ++r) {
                    const int s = it * LP::R + r;
                    const int kk = base + r * LP::NS + (grp ? Q : 0);
                    w = r == 0 ? this->root(kk % N) : cmul(w, w_r);
                    cplx a, b;
                    if (grp) {
                        ctx.tmem_ld_wait2(la, lb, a, b);
                    } else {
                        a = keep[s];
                        b = ctx.tmem_ld_wait(lb);
                    }
                    if (s + 1 < 16) {
                        if (grp) ctx.tmem_ld_issue(tbase, tl, a_col + 4 * (s + 1), la);
                        ctx.tmem_ld_issue(tbase, tl, b_col + 4 * (s + 1), lb);
                    }
                    const cplx wb = cmul(b, w);
                    st_stream(o_prev + wrap_add(kk, n / 2, n), cadd(a, wb));
                    st_stream(o_prev + wrap_add(kk + H, n / 2, n), csub(a, wb));
                }
            }
        };
        for (int64_t line = ctx.bid; line < this->op.g.n_lines; line += ctx.nblocks) {
            const int f = (int)(line / this->op.lines_per);
            const int l = (int)(line - (int64_t)f * this->op.lines_per);
            const ColumnFacet& F = this->op.fac[f];
            const int shift_in = F.shift_in, fs = F.fs;
            const double* fb = this->op.fb ? this->op.fb + F.fb_off : nullptr;
            cplx* o = F.out + (int64_t)l * F.out_ls;
            const double scale = this->op.scale;
            const bool swz = this->swizzled != 0;
            auto sample = [&](int q) {  // natural-order sample q of the padded, rotated row
                int k = wrap_add(q, shift_in, n);
                if (k >= fs) return mk(0.0, 0.0);
                const int ks = swz ? ((k & ~7) | ((k ^ (k >> 3)) & 7)) : k;
                return fb ? cscale(in[ks], ldg_d(fb + k)) : in[ks];
            };
            if (grp == 0) {
                // group 0 stores its half of the previous line now (kept half still in registers),
                // then tells group 1 that its parked half has been read
                if (o_prev) store_prev(0, g1_col);
                ctx.tmem_fence_before();
                ctx.group_arrive(3, THREADS);
            }
            // this line's regions: E0, then the half parked for the partner
            const int col0 = grp ? g1_col : g0_col + par_col;
            ctx.tx_wait(bar, parity);  // this line's row has landed
            parity ^= 1;
            SkewSync<Ctx> gs{ctx, *this, grp, tg, in, bar, done, line + ctx.nblocks, false, grp == 1};
            {   // E0 = FFT_Q(z[4 j + g]) -> TMEM
                auto ld = [&](int j) { return sample(4 * j + grp); };
                auto st = [&](int, cplx v, int it, int r) {
                    ctx.tmem_st(tbase, tl, col0 + 4 * (it * LP::R + r), v);
                };
                line_fft<Q, DIR>(tg, sm, tw_(), ld, st, gs);
                ctx.tmem_wait_st();
                gs();  // the group's exchange buffer is reused by its second sub-transform
            }
            if (grp == 1 && o_prev) {
                // group 1 stores its half of the previous line here, under group 0's transforms:
                // a = group 0's parked half of that line (the other parity), b = own kept half
                store_prev(g0_col + (par_col ^ TCOLS), g1_col + TCOLS);
            }
            {   // E1 = FFT_Q(z[4 j + 2 + g]); U_g[k] = E0 + v^k E1, U_g[k + Q] = E0 - v^k E1
                auto ld = [&](int j) { return sample(4 * j + 2 + grp); };
                const cplx v_it = this->root((2 * TG) % N), v_r = this->root((2 * LP::NS) % N);
                cplx w_it = mk(1.0, 0.0), w = mk(1.0, 0.0);
                typename Ctx::TmemLoad le;
                auto st = [&](int k, cplx od, int it, int r) {
                    if (r == 0) {
                        w_it = it == 0 ? this->root((2 * k) % N) : cmul(w_it, v_it);
                        w = w_it;
                    } else {
                        w = cmul(w, v_r);
                    }
                    const int s = it * LP::R + r;
                    // (E0 of slot s + 1 is in flight while slot s is combined and parked)
                    if (s == 0) ctx.tmem_ld_issue(tbase, tl, col0, le);
                    const cplx e = ctx.tmem_ld_wait(le);
                    if (s + 1 < 16) ctx.tmem_ld_issue(tbase, tl, col0 + 4 * (s + 1), le);
                    const cplx wo = cmul(od, w);
                    const cplx lo = cscale(cadd(e, wo), scale);
                    const cplx hi = cscale(csub(e, wo), scale);
                    if (grp) {  // parks U_1[k] for group 0, keeps U_1[k + Q] (in TMEM)
                        ctx.tmem_st(tbase, tl, col0 + 4 * s, lo);
                        ctx.tmem_st(tbase, tl, col0 + TCOLS + 4 * s, hi);
                    } else {  // keeps U_0[k] in registers, parks U_0[k + Q] for group 1
                        keep[s] = lo;
                        ctx.tmem_st(tbase, tl, col0 + 4 * s, hi);
                    }
                };
                gs.pending = true;
                line_fft<Q, DIR>(tg, sm, tw_(), ld, st, gs);
                ctx.tmem_wait_st();
            }
            // swap: everybody's parked half is visible to the partner after the barrier
            ctx.tmem_fence_before();
            ctx.sync();
            ctx.tmem_fence_after();
            o_prev = o;
            par_col ^= TCOLS;
        }
        if (o_prev) {
            if (grp == 0)
                store_prev(0, g1_col);
            else
                store_prev(g0_col + (par_col ^ TCOLS), g1_col + TCOLS);
        }
        ctx.tmem_free(tbase, NCOLS);
    }
    SW_HD const cplx* tw_() const { return this->tw; }
};

}  // namespace swiftly
