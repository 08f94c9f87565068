// SwiFTly B200 -- dispatch of the fused subgrid axis kernel over (m, xM) pairs.
#include "dispatch.cuh"

namespace swiftly {

template <int M, int XM>
static int launch_sg_axis(const swiftly_b200* h, const SubgridAxisArgs& a, cudaStream_t s) {
    SubgridAxisKernel<M, XM> k;
    for (int i = 0; i < SW_MAX_SOURCES; ++i) k.src[i] = a.src[i];
    k.n_slots = a.n_slots;
    k.fn = h->d_Fn;
    k.tw_m = twiddles(h, M);
    k.tw_x = twiddles(h, XM);
    if (!k.tw_m || !k.tw_x) return SWIFTLY_B200_ECUDA;
    k.n_lines = a.n_lines;
    k.out = a.out;
    k.out_ls = a.out_ls;
    k.out_es = a.out_es;
    k.sz = a.sz;
    k.start = a.start;
    k.scale = 1.0 / (double)XM;
    k.mask = a.mask;
    cudaError_t e = launch_body(k, grid_for(a.n_lines, 1), k.SMEM, s);
    return e == cudaSuccess ? SWIFTLY_B200_OK : cuda_fail(e, "subgrid axis kernel launch");
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
