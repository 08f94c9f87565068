// SwiFTly B200 -- size dispatch of extract_from_subgrid (one translation unit per primitive keeps
// the heavy FP64 template instantiations compiling in parallel).
#include "dispatch.cuh"

namespace swiftly {

int run_extract_from_subgrid(const swiftly_b200* h, const ExtractFromSubgridOp& op, bool lf, cudaStream_t s) {
    const int n = op.m;
    switch (n) {
        SW_DIRECT_CASES(+1, ExtractFromSubgridOp)
        default: break;
    }
    {
        int M = 0, F = 0;
        if (split_f_plan(n, &M, &F)) {
            SW_SPLIT_F_CASES(+1, ExtractFromSubgridOp, M, F)
        }
    }
    return unsupported(n);
}

}  // namespace swiftly
