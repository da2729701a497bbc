// Instantiations of the tcgen05 implicit-GEMM kernel that also emit per-column (sum, sum of squares) partials of the
// output from the epilogue (fused GroupNorm statistics, see tap_gemm_kernel.cuh / dcb200.h).
#include "tap_gemm_kernel.cuh"

namespace dcb {

int tap_gemm_launch_cs(int bn, bool out_bf16, bool pair, const CUtensorMap& tmA, const CUtensorMap* tmB,
                       const CUtensorMap& tmOut, const CUtensorMap& tmRes, TapGemmParams& p, cudaStream_t s) {
    return launch_all<1>(bn, out_bf16, pair, tmA, tmB, tmOut, tmRes, p, s);
}

}  // namespace dcb
