// jz_gemm_tc_xform_ts.cu -- instantiates gemm_tcgen05_kernel<CG = 1, MODE_XFORM_TS, TN, AMN, BMN> (jz_gemm_tc.cuh): 3xTF32 with
// the A operand's hi / lo parts staged in tensor memory (tcgen05.st by the transform warps, tcgen05.mma with a TMEM A
// operand), for products with a narrow or short output (n <= 128 or m <= 128) that the shared-memory form of the
// transform bounds by shared-memory bandwidth.
#define JZ_GEMM_TC_IMPL
#include "jz_gemm_tc.cuh"

namespace jz {
namespace tc {

int launch_tc_ts(int tn, const Operand& a, const Operand& b, const GemmArgs& args, unsigned batch, cudaStream_t s) {
    if (tn == 128) return launch_tc_major<1, MODE_XFORM_TS, 128>(a, b, args, batch, s);
    if (tn == 64) return launch_tc_major<1, MODE_XFORM_TS, 64>(a, b, args, batch, s);
    return fail(JZ_ERR_ARG, "gemm: no TMEM-A kernel for TN=%d", tn);
}

}  // namespace tc
}  // namespace jz
