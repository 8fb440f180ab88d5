// plip_b200 — host-side launchers of the non-GEMM kernels.
#pragma once
#include "plip_b200.h"

#include "common.cuh"
#include "gemm.cuh"

namespace plip {

// elementwise.cu  (f16 = 1: 16-bit outputs are IEEE half instead of bfloat16; the pointer type stays a 16-bit tag)
int launch_im2col(const void* pixels, int fmt, int64_t n, __nv_bfloat16* out, int f16, cudaStream_t st);
int launch_layernorm(const float* x, const int32_t* row_index, int64_t in_row_stride, int64_t rows, int dim,
                     const float* gamma, const float* beta, float* out_f32, __nv_bfloat16* out_bf16, int f16,
                     cudaStream_t st);
int launch_rowstats_cast(const float* x, int64_t rows, int dim, __nv_bfloat16* xb, float2* stats, int f16, cudaStream_t st);
int launch_text_embed(const void* ids, int ids_dtype, int64_t n, int seq_len, int ids_stride, const float* tok,
                      const float* pos, float* x, int32_t* eos_rows, int eos_id, int no_eos_argmax, cudaStream_t st);
int launch_mask_to_i32(const void* mask, int dtype, int64_t count, int seq_len, int stride, int32_t* out,
                       cudaStream_t st);
int launch_cls_rows(const float* cls, const float* pos, int64_t n, float* x, cudaStream_t st);
int launch_l2_normalize(float* x, int64_t rows, int dim, cudaStream_t st);
// rows idx(i) (= row_index[i], or i * row_stride) of a 16-bit [*, dim] and an fp32 [*, dim] matrix -> compact [n, dim]
int launch_gather_rows(const __nv_bfloat16* a16, const float* x32, const int32_t* row_index, int64_t row_stride,
                       int64_t n, int dim, __nv_bfloat16* a16_out, float* x32_out, cudaStream_t st);

// resize.cu: Pillow-exact bicubic resize + crop of packed RGB uint8 images into [n,224,224,3] tiles.
int launch_resize_crop(const uint8_t* src, size_t src_bytes, const plip_resize_desc_t* descs_host, int64_t n,
                       uint8_t* tiles, cudaStream_t st);

int resize_filter_host(int in_size, int out_size, int xx, int32_t* k, int k_cap, int* xmin, int* count);

// attention.cu: softmax(q k^T [+causal/padding mask]) v per (sequence, head); q pre-scaled by dh^-0.5.
// qkv: bf16 [n_seq*seq_len, 3*heads*64]; key_mask: optional int32 [n_seq, seq_len] (0 = masked key).
// f16 = 1: q, k, v, P and the output are IEEE half instead of bfloat16 (the engine's operand format).
int launch_attention(const __nv_bfloat16* qkv, int64_t n_seq, int seq_len, int heads, bool causal,
                     const int32_t* key_mask, __nv_bfloat16* out, int f16, cudaStream_t st);

// similarity.cu
int launch_similarity(const float* a, int64_t n, const float* b, int64_t m, float scale, bool norm_a, bool norm_b,
                      float* out, int64_t ldo, cudaStream_t st);
int launch_similarity_topk(const float* q, int64_t n, const float* s, int64_t m, float scale, bool norm_q,
                           bool norm_s, int k, int32_t* idx, float* val, cudaStream_t st);

}  // namespace plip
