// SwiFTly B200 -- size dispatch of subgrid_to_facets (one translation unit per primitive keeps
// the heavy FP64 template instantiations compiling in parallel).
#include "dispatch.cuh"

namespace swiftly {

int run_subgrid_to_facets(const swiftly_b200* h, const SubgridToFacetsOp& op, bool lf, cudaStream_t s) {
    const int n = op.m;
    switch (n) {
        SW_DIRECT_CASES(+1, SubgridToFacetsOp)
        default: break;
    }
    {
        int M = 0, F = 0;
        if (split_f_plan(n, &M, &F)) {
            SW_SPLIT_F_CASES(+1, SubgridToFacetsOp, M, F)
        }
    }
    return unsupported(n);
}

}  // namespace swiftly
