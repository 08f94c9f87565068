// SwiFTly B200 -- size dispatch of finish_subgrid (one translation unit per primitive keeps
// the heavy FP64 template instantiations compiling in parallel).
#include "dispatch.cuh"

namespace swiftly {

int run_finish_subgrid(const swiftly_b200* h, const FinishSubgridOp& op, bool lf, cudaStream_t s) {
    const int n = op.xM;
    switch (n) {
        SW_DIRECT_CASES(+1, FinishSubgridOp)
        default: break;
    }
    {
        int M = 0, F = 0;
        if (split_f_plan(n, &M, &F)) {
            SW_SPLIT_F_CASES(+1, FinishSubgridOp, M, F)
        }
    }
    return unsupported(n);
}

}  // namespace swiftly
