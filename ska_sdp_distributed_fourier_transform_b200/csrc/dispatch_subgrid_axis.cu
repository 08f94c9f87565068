// SwiFTly B200 -- dispatch of the fused subgrid axis kernel over (m, xM) pairs.
#include "dispatch.cuh"

namespace swiftly {

template <int M, int XM, int LINES>
static int launch_sg_axis_l(const swiftly_b200* h, const SubgridAxisArgs& a, cudaStream_t s) {
    SubgridAxisKernel<M, XM, LINES> k;
    for (int i = 0; i < SW_MAX_SOURCES; ++i) k.src[i] = a.src[i];
    k.n_slots = a.n_slots;
    k.n_groups = a.n_groups;
    k.fn = h->d_Fn;
    k.tw_m = twiddles(h, M);
    k.tw_x = twiddles(h, XM);
    if (!k.tw_m || !k.tw_x) return SWIFTLY_B200_ECUDA;
    k.n_lines = a.n_lines;
    k.out = a.out;
    k.out_ls = a.out_ls;
    k.out_es = a.out_es;
    k.out_gs = a.out_gs;
    k.sz = a.sz;
    for (int g = 0; g < SW_MAX_GROUPS; ++g) {
        k.start[g] = a.start[g];
        k.mask[g] = a.mask[g];
    }
    k.scale = 1.0 / (double)XM;
    k.first_round_tiles = a.first_round_tiles;
    cudaError_t e = launch_body(k, grid_for(((a.n_lines + LINES - 1) / LINES) * a.n_groups, 1), k.SMEM, s);
    return e == cudaSuccess ? SWIFTLY_B200_OK : cuda_fail(e, "subgrid axis kernel launch");
}

// two adjacent lines per CTA when every source and the output have unit line stride
template <int M, int XM>
static int launch_sg_axis(const swiftly_b200* h, const SubgridAxisArgs& a, cudaStream_t s) {
    bool adjacent = a.out_ls == 1 && a.n_lines > 1;
    for (int i = 0; i < SW_MAX_SOURCES && adjacent; ++i)
        if (a.src[i].base && a.src[i].ls != 1) adjacent = false;
    // 2 lines need 2 x (acc + work) of shared memory: only the pairs that fit
    if constexpr (2 * ((size_t)(XM + 4) * sizeof(cplx) + (size_t)(XM + XM / 16 + 40) * 8) <=
                  227 * 1024) {
        if (adjacent) return launch_sg_axis_l<M, XM, 2>(h, a, s);
    }
    return launch_sg_axis_l<M, XM, 1>(h, a, s);
}

#define SW_SG_PAIRS(X) \
    X(32, 64) X(32, 128) X(64, 128) X(64, 256) X(128, 256) X(128, 512) X(256, 512) X(256, 1024) \
    X(512, 1024) X(512, 2048) X(1024, 2048) X(1024, 4096) X(2048, 4096) X(2048, 8192)

int subgrid_axis_conc(int m, int xM) {
#define X(M, XM) if (m == M && xM == XM) return XM / M;
    SW_SG_PAIRS(X)
#undef X
    return 0;
}

int run_subgrid_axis(const swiftly_b200* h, const SubgridAxisArgs& a, cudaStream_t s) {
    const int m = (int)h->m, xM = (int)h->xM;
#define X(M, XM) if (m == M && xM == XM) return launch_sg_axis<M, XM>(h, a, s);
    SW_SG_PAIRS(X)
#undef X
    set_error("no fused subgrid kernel for m=" + std::to_string(m) + ", xM=" + std::to_string(xM));
    return SWIFTLY_B200_EUNSUPPORTED;
}

}  // namespace swiftly
