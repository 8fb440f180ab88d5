// plip_b200 — tcgen05 GEMM interface (host side).
#pragma once
#include "common.cuh"

namespace plip {

// Fused epilogues.  All GEMMs compute acc[M,N] = A[M,K] (bf16) x W[N,K]^T (bf16), fp32 accumulate.
enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,       // out_bf16 = acc + bias                     (QKV projection, TF:modeling_clip.py:310-312)
  EPI_BIAS_GELU_BF16 = 1,  // out_bf16 = quick_gelu(acc + bias)         (fc1, TF:modeling_clip.py:348-349)
  EPI_BIAS_RESID_F32 = 2,  // x_f32   += acc + bias  (in place)         (out_proj / fc2 + residual, :334,:377,:350,:382)
  EPI_PATCH_F32 = 3,       // x_f32[b*50+1+p] = acc + pos[1+p]          (patch conv + position embedding, :209-217)
  EPI_F32 = 4,             // out_f32 = acc                             (visual/text projection, :861,:823)
  EPI_COUNT = 5
};

struct GemmArgs {
  const __nv_bfloat16* A = nullptr;  // [M, K] row-major, row stride lda elements
  int lda = 0;
  const __nv_bfloat16* W = nullptr;  // [N, K] row-major (nn.Linear weight layout), row stride ldw
  int ldw = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;       // [N]
  void* out = nullptr;               // bf16 or fp32 depending on epilogue; row stride ldo elements
  int ldo = 0;
  const float* pos = nullptr;        // EPI_PATCH_F32: vision position embedding [50, N]
  int epi = EPI_F32;
  int force_cg = 0;                  // 0 = auto; 1 / 2 = CTA-group size (test hook)
  int force_bn = 0;                  // 0 = auto; 128 / 256 = N tile (test hook)
};

// Enqueue the GEMM on `stream`.  Returns 0 on success (see last_error otherwise).
int launch_gemm(const GemmArgs& g, cudaStream_t stream);

// Number of kernel launches issued by this translation unit since load (bench accounting).
extern unsigned long long g_launch_count;

}  // namespace plip
