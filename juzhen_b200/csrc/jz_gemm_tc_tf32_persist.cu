// jz_gemm_tc_tf32_persist.cu -- instantiates gemm_tf32_persistent_kernel<AMN, BMN> (jz_gemm_tc.cuh): the persistent
// single-pass TF32 GEMM whose epilogue runs under the next tile's mainloop.
#define JZ_GEMM_TC_IMPL
#define JZ_GEMM_TC_PERSIST_IMPL
#include "jz_gemm_tc.cuh"

namespace jz {
namespace tc {

int launch_tc_tf32_persistent(const Operand& a, const Operand& b, const GemmArgs& args, cudaStream_t s) {
    return launch_tf32_persistent_major(a, b, args, s);
}

}  // namespace tc
}  // namespace jz
