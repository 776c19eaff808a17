// jz_gemm_tc_xform_cg1.cu -- instantiates gemm_tcgen05_kernel<CG = 1, MODE_XFORM, TN, AMN, BMN> (jz_gemm_tc.cuh) for every
// tile width and operand-major combination of this mode / CTA-group size.  One translation unit per (mode, CG) so
// the instantiations compile in parallel.
#define JZ_GEMM_TC_IMPL
#include "jz_gemm_tc.cuh"

namespace jz {
namespace tc {

template <>
int launch_tc_cg<MODE_XFORM, 1>(int tn, const Operand& a, const Operand& b, const GemmArgs& args, unsigned batch, cudaStream_t s) {
    if (tn == 256) return launch_tc_major<1, MODE_XFORM, 256, false>(a, b, args, batch, s);
    if (tn == 128) return launch_tc_major<1, MODE_XFORM, 128, false>(a, b, args, batch, s);
    if (tn == 64) return launch_tc_major<1, MODE_XFORM, 64, false>(a, b, args, batch, s);
    return fail(JZ_ERR_ARG, "gemm: no tensor-core kernel for CG=1 TN=%d", tn);
}

}  // namespace tc
}  // namespace jz
