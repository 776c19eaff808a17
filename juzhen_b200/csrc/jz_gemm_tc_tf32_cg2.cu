// jz_gemm_tc_tf32_cg2.cu -- instantiates gemm_tcgen05_kernel<CG = 2, MODE_TF32, TN, AMN, BMN> (jz_gemm_tc.cuh) for every
// tile width and operand-major combination of this mode / CTA-group size.  One translation unit per (mode, CG) so
// the instantiations compile in parallel.
#define JZ_GEMM_TC_IMPL
#include "jz_gemm_tc.cuh"

namespace jz {
namespace tc {

template <>
int launch_tc_cg<MODE_TF32, 2>(int tn, const Operand& a, const Operand& b, const GemmArgs& args, unsigned batch, cudaStream_t s) {
    if (tn == 256) return launch_tc_major<2, MODE_TF32, 256>(a, b, args, batch, s);
    if (tn == 128) return launch_tc_major<2, MODE_TF32, 128>(a, b, args, batch, s);
    return fail(JZ_ERR_ARG, "gemm: no tensor-core kernel for CG=2 TN=%d", tn);
}

}  // namespace tc
}  // namespace jz
