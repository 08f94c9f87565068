// SwiFTly B200 -- size dispatch of add_to_subgrid (one translation unit per primitive keeps
// the heavy FP64 template instantiations compiling in parallel).
#include "dispatch.cuh"

namespace swiftly {

int run_add_to_subgrid(const swiftly_b200* h, const AddToSubgridOp& op, bool lf, cudaStream_t s) {
    const int n = op.m;
    switch (n) {
        SW_DIRECT_CASES(-1, AddToSubgridOp)
        default: break;
    }
    {
        int M = 0, F = 0;
        if (split_f_plan(n, &M, &F)) {
            SW_SPLIT_F_CASES(-1, AddToSubgridOp, M, F)
        }
    }
    return unsupported(n);
}

}  // namespace swiftly
