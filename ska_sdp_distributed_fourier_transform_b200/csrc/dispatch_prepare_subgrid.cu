// SwiFTly B200 -- size dispatch of prepare_subgrid (one translation unit per primitive keeps
// the heavy FP64 template instantiations compiling in parallel).
#include "dispatch.cuh"

namespace swiftly {

int run_prepare_subgrid(const swiftly_b200* h, const PrepareSubgridOp& op, bool lf, cudaStream_t s) {
    const int n = op.xM;
    switch (n) {
        SW_DIRECT_CASES(-1, PrepareSubgridOp)
        default: return unsupported(n);
    }
}

}  // namespace swiftly
