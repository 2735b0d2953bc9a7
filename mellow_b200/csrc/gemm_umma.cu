// tcgen05 / TMA GEMM engine (sm_100a).  Placeholder until the engine lands: reports "not handled" so the caller
// falls through to the mma.sync engine.
#include "gemm.cuh"

namespace mb {

cudaError_t launch_gemm_umma(const GemmArgs& g, int epi, cudaStream_t st, bool* handled) {
    (void)g; (void)epi; (void)st;
    *handled = false;
    return cudaSuccess;
}

}  // namespace mb
