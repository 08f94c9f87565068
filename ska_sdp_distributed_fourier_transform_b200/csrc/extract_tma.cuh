// SwiFTly B200 -- K2 ("Fb . FFT . extract", the reference's extract_column task,
// api_helper.py:200-210 = extract_from_facet(axis 0) + prepare_facet(axis 1),
// core.py:189-253) with the facet rows staged in shared memory by the TMA engine.
//
// One persistent CTA per SM walks over output lines.  A line's input is one contiguous row of
// the axis-0 prepared facet BF_F (fs samples; the row gather of extract_from_facet is just the
// choice of the row).  The row is brought into shared memory by ONE bulk asynchronous copy
// (cp.async.bulk.shared.global, completion on an mbarrier) that is issued while the PREVIOUS
// line is still being transformed: as soon as the last first-pass load of the current line has
// read the staging buffer, the copy of the next row starts and runs under the remaining
// passes, the combine and the stores.  The first-pass loads therefore hit shared memory instead
// of exposing a DRAM round trip with every warp of the SM waiting in the same phase (round 1:
// long-scoreboard stalls 4.3 warps per issue, 28 % of the HBM roofline).
//
// SPLIT = true : line length 2 H (yN = 16384 = 2 x 8192), decimation in time; E = FFT_H(even
//                samples) is parked in a per-CTA scratch line (L2), O = FFT_H(odd samples),
//                X[k] = E[k] + w^k O[k], X[k + H] = E[k] - w^k O[k] (as SplitLineKernel).
// SPLIT = false: line length H, one transform straight from the staging buffer.
#pragma once

#include "kernels.cuh"

namespace swiftly {

template <int H, bool SPLIT>
struct ExtractColumnsTmaKernel {
    static constexpr int DIR = +1;
    static constexpr int T = FftCfg<H>::T;
    static constexpr int THREADS = T;
    static constexpr int N = SPLIT ? 2 * H : H;  // line length yN
    // staging buffer: up to N - 1 facet samples (fs <= yN - 1); exchange buffer; mbarrier
    static constexpr size_t smem_bytes(int in_cap) {
        return (size_t)in_cap * sizeof(cplx) + (size_t)FftCfg<H>::PADDED * sizeof(double) + 16;
    }

    ExtractColumnsOp op;
    const cplx* tw;   // compact table of the H-point plan
    const cplx* tw2;  // exp(-2 pi i t / 2H), t < H (SPLIT only)
    cplx* scratch;    // gridDim.x * H samples (SPLIT only)
    int in_cap;       // capacity of the staging buffer in samples (>= every facet's fs, even)

    // first-pass loads must be done before the staging buffer is refilled: the refill is issued
    // by thread 0 right after the first barrier that follows them
    template <class Ctx>
    struct RefillSync {
        const Ctx& ctx;
        const ExtractColumnsTmaKernel& k;
        cplx* in;
        uint64_t* bar;
        int64_t next_line;
        bool pending;
        SW_HD void operator()() {
            ctx.sync();
            if (pending) {
                pending = false;
                if (ctx.tid == 0 && next_line < k.op.g.n_lines) k.issue(ctx, in, bar, next_line);
            }
        }
    };

    template <class Ctx>
    SW_HD void issue(const Ctx& ctx, cplx* in, uint64_t* bar, int64_t line) const {
        const int f = (int)(line / op.lines_per);
        const int l = (int)(line - (int64_t)f * op.lines_per);
        const ColumnFacet& F = op.fac[f];
        const int64_t row = wrap_add(op.rm_base, wrap_sub(l, op.rm_s_m, op.lines_per), op.n);
        const uint32_t bytes = (uint32_t)F.fs * (uint32_t)sizeof(cplx);
        ctx.tx_expect(bar, bytes);
        // one bulk copy may not exceed the engine's size field comfortably: 64 KiB pieces
        const char* src = (const char*)(F.in + row * F.in_ls);
        char* dst = (char*)in;
        for (uint32_t o = 0; o < bytes; o += 65536u) {
            const uint32_t n = bytes - o < 65536u ? bytes - o : 65536u;
            ctx.tx_copy(dst + o, src + o, n, bar);
        }
    }

    template <class Ctx>
    SW_HD void operator()(Ctx& ctx) const {
        cplx* in = (cplx*)ctx.smem;
        double* sm = (double*)(in + in_cap);
        uint64_t* bar = (uint64_t*)(sm + ((FftCfg<H>::PADDED + 1) & ~1));
        cplx* stash = SPLIT ? scratch + (size_t)ctx.bid * H : nullptr;
        const int lt = ctx.tid;
        const int n = op.n;
        if (ctx.tid == 0) {
            ctx.tx_init(bar);
            if ((int64_t)ctx.bid < op.g.n_lines) issue(ctx, in, bar, ctx.bid);
        }
        ctx.sync();
        uint32_t parity = 0;
        for (int64_t line = ctx.bid; line < op.g.n_lines; line += ctx.nblocks) {
            const int f = (int)(line / op.lines_per);
            const int l = (int)(line - (int64_t)f * op.lines_per);
            const ColumnFacet& F = op.fac[f];
            const int shift_in = F.shift_in, fs = F.fs;
            const double* fb = op.fb + F.fb_off;
            cplx* o = F.out + (int64_t)l * F.out_ls;
            const double scale = op.scale;
            // natural-order input sample q of the zero-padded, rotated, Fb-weighted row
            auto sample = [&](int q) {
                int k = wrap_add(q, shift_in, n);
                if (k >= fs) return mk(0.0, 0.0);
                return cscale(in[k], ldg_d(fb + k));
            };
            auto put = [&](int p, cplx v) {
                int pc = wrap_add(p, n / 2, n);
                st_stream(o + pc, cscale(v, scale));
            };
            ctx.tx_wait(bar, parity);  // this line's row has landed
            parity ^= 1;
            RefillSync<Ctx> refill{ctx, *this, in, bar, line + ctx.nblocks, false};
            if constexpr (SPLIT) {
                {
                    auto sync = [&]() { ctx.sync(); };
                    auto ld = [&](int q) { return sample(2 * q); };
                    auto st = [&](int k, cplx v) { stash[k] = v; };
                    line_fft<H, DIR>(lt, sm, tw, ld, st, sync);
                }
                ctx.sync();
                {
                    auto ld = [&](int q) { return sample(2 * q + 1); };
                    auto st = [&](int k, cplx od) {
                        cplx w = ldg_c(tw2 + k);
                        if (DIR > 0) w.y = -w.y;
                        cplx e = stash[k];
                        cplx wo = cmul(od, w);
                        put(k, cadd(e, wo));
                        put(k + H, csub(e, wo));
                    };
                    refill.pending = true;
                    line_fft<H, DIR>(lt, sm, tw, ld, st, refill);
                }
            } else {
                auto ld = [&](int q) { return sample(q); };
                refill.pending = true;
                line_fft<H, DIR>(lt, sm, tw, ld, put, refill);
            }
            ctx.sync();  // exchange buffer is reused by the next line
        }
    }
};

}  // namespace swiftly
