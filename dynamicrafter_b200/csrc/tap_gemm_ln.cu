// Instantiations of the tcgen05 implicit-GEMM kernel with LayerNorm folding (see tap_gemm_kernel.cuh / dcb200.h): the
// fp32-out variants also write a bf16 operand copy and per-row (sum, sum of squares); the bf16-out variants apply the
// per-row normalisation of a gamma-folded weight matrix in their epilogue.  Own translation unit: the plain kernels
// carry no code or register cost for it.
#include "tap_gemm_kernel.cuh"

namespace dcb {

int tap_gemm_launch_ln(int bn, bool out_bf16, bool pair, const CUtensorMap& tmA, const CUtensorMap* tmB,
                       const CUtensorMap& tmOut, const CUtensorMap& tmRes, TapGemmParams& p, cudaStream_t s) {
    return launch_all<2>(bn, out_bf16, pair, tmA, tmB, tmOut, tmRes, p, s);
}

}  // namespace dcb
