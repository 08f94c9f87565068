// SwiFTly B200 -- size dispatch of finish_subgrid (one translation unit per primitive keeps
// the heavy FP64 template instantiations compiling in parallel).
#include "dispatch.cuh"

namespace swiftly {

int run_finish_subgrid(const swiftly_b200* h, const FinishSubgridOp& op, bool lf, cudaStream_t s) {
    const int n = op.xM;
    switch (n) {
        SW_DIRECT_CASES(+1, FinishSubgridOp)
        default: return unsupported(n);
    }
}

}  // namespace swiftly
