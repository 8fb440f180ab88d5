// plip_b200 — C-ABI entry points (see include/plip_b200.h).
#include "plip_b200.h"

#include "common.cuh"
#include "gemm.cuh"

namespace plip {
const char* get_last_error();
}

extern "C" {

PLIP_API const char* plip_last_error(void) { return plip::get_last_error(); }
PLIP_API int plip_abi_version(void) { return PLIP_B200_ABI_VERSION; }
PLIP_API uint64_t plip_launch_count(void) { return plip::g_launch_count; }

PLIP_API int plip_dbg_gemm(const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K,
                           const float* bias, void* out, int ldo, const float* pos, int epilogue, int cta_group,
                           int block_n, void* stream) {
  plip::GemmArgs g;
  g.A = static_cast<const __nv_bfloat16*>(A_bf16);
  g.lda = lda;
  g.W = static_cast<const __nv_bfloat16*>(W_bf16);
  g.ldw = ldw;
  g.M = M; g.N = N; g.K = K;
  g.bias = bias;
  g.out = out;
  g.ldo = ldo;
  g.pos = pos;
  g.epi = epilogue;
  g.force_cg = cta_group;
  g.force_bn = block_n;
  return plip::launch_gemm(g, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
