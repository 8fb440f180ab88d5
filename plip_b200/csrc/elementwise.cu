// plip_b200 — HBM-bound helper kernels around the GEMMs: patch im2col (+ uint8 preprocessing),
// LayerNorm, token/position embedding gather, class-token rows, EOS search, L2 normalisation.
//
// All are pure streaming kernels (roofline: HBM); accesses are 16-byte vectorised and coalesced.
#include "kernels.cuh"

namespace plip {

namespace {

constexpr int kEwThreads = 256;

__device__ __forceinline__ uint4 pack8(const float (&f)[8], int f16) {
  uint4 u;
  u.x = pack_op2_rt(f[0], f[1], f16);
  u.y = pack_op2_rt(f[2], f[3], f16);
  u.z = pack_op2_rt(f[4], f[5], f16);
  u.w = pack_op2_rt(f[6], f[7], f16);
  return u;
}

// ------------------------------------------------------------------------------------------------
// im2col of non-overlapping 32x32 patches: pixels -> A0[b*49 + py*7 + px][c*1024 + ky*32 + kx] (bf16).
// Restates nn.Conv2d(3,768,32,32,bias=False) input gathering (TF:modeling_clip.py:148-154,209-210);
// the uint8 path fuses CLIPImageProcessor's rescale + normalise (TF:image_processing_clip.py:50-62).
// One thread moves 8 consecutive x of one image row.
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(kEwThreads) im2col_kernel(const void* __restrict__ pixels,
                                                            __nv_bfloat16* __restrict__ out, int64_t n, int f16) {
  constexpr int kX8 = kImage / 8;  // 28 groups of 8 pixels per image row
  const int64_t total = (FMT == PLIP_PIX_U8_NHWC) ? n * kImage * kX8 : n * 3 * kImage * kX8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int x8 = (int)(i % kX8);
    int64_t r = i / kX8;
    const int y = (int)(r % kImage);
    r /= kImage;
    const int x = x8 * 8;
    const int py = y >> 5, ky = y & 31, px = x >> 5, kx = x & 31;
    if constexpr (FMT == PLIP_PIX_U8_NHWC) {
      const int64_t b = r;
      const uint8_t* src = static_cast<const uint8_t*>(pixels) + ((b * kImage + y) * kImage + x) * 3;
      const uint2* s2 = reinterpret_cast<const uint2*>(src);  // 24 bytes, 8-byte aligned
      uint2 w0 = __ldg(s2), w1 = __ldg(s2 + 1), w2 = __ldg(s2 + 2);
      uint8_t bytes[24];
      *reinterpret_cast<uint2*>(bytes) = w0;
      *reinterpret_cast<uint2*>(bytes + 8) = w1;
      *reinterpret_cast<uint2*>(bytes + 16) = w2;
      const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
      const float istd[3] = {1.0f / 0.26862954f, 1.0f / 0.26130258f, 1.0f / 0.27577711f};
      __nv_bfloat16* dst = out + (b * kPatches + py * kGrid + px) * (int64_t)kPatchK + ky * 32 + kx;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (bytes[j * 3 + c] * (1.0f / 255.0f) - mean[c]) * istd[c];
        *reinterpret_cast<uint4*>(dst + c * 1024) = pack8(f, f16);
      }
    } else {
      const int c = (int)(r % 3);
      const int64_t b = r / 3;
      const int64_t src_off = ((b * 3 + c) * kImage + y) * kImage + x;
      __nv_bfloat16* dst =
          out + (b * kPatches + py * kGrid + px) * (int64_t)kPatchK + c * 1024 + ky * 32 + kx;
      if constexpr (FMT == PLIP_PIX_F32_NCHW) {
        const float4* s4 = reinterpret_cast<const float4*>(static_cast<const float*>(pixels) + src_off);
        const float4 a = __ldg(s4), bq = __ldg(s4 + 1);
        const float f[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
        *reinterpret_cast<uint4*>(dst) = pack8(f, f16);
      } else {  // bf16 NCHW: straight 16-byte copy (re-rounded to half for the fp16 operand format)
        uint4 w = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(pixels) + src_off));
        if (f16) {
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&w);
          float f[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 t = __bfloat1622float2(h2[j]);
            f[2 * j] = t.x;
            f[2 * j + 1] = t.y;
          }
          w = pack8(f, 1);
        }
        *reinterpret_cast<uint4*>(dst) = w;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (eps 1e-5, affine), one warp per row, fp32 statistics (two-pass in
// registers).  TF:modeling_clip.py:371,380 (layer_norm1/2), :677 (pre_layrnorm), :686 (post_layernorm
// on the CLS row), :562 (final_layer_norm; only the pooled EOS row is needed downstream).
// Rows are addressed as x + row_index[r] * in_row_stride (row_index == nullptr -> r).
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kEwThreads) layernorm_kernel(const float* __restrict__ x,
                                                               const int32_t* __restrict__ row_index,
                                                               int64_t in_row_stride, int64_t rows,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               float* __restrict__ out_f32,
                                                               __nv_bfloat16* __restrict__ out_bf16, int f16) {
  constexpr int V = D / 128;  // float4 per lane
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const int64_t src_row = row_index ? (int64_t)row_index[r] : r;
    const float4* xr = reinterpret_cast<const float4*>(x + src_row * in_row_stride);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = xr[lane + 32 * j];
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + kLnEps);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
      float4 y;
      y.x = (v[j].x - mean) * rstd * g.x + b.x;
      y.y = (v[j].y - mean) * rstd * g.y + b.y;
      y.z = (v[j].z - mean) * rstd * g.z + b.z;
      y.w = (v[j].w - mean) * rstd * g.w + b.w;
      if (out_f32) reinterpret_cast<float4*>(out_f32 + r * D)[lane + 32 * j] = y;
      if (out_bf16) {
        uint2 u;
        u.x = pack_op2_rt(y.x, y.y, f16);
        u.y = pack_op2_rt(y.z, y.w, f16);
        reinterpret_cast<uint2*>(out_bf16 + r * D)[lane + 32 * j] = u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bf16 copy + per-row (sum, sum of squares) of the fp32 residual stream: the A operand and the
// statistics a LayerNorm-folded GEMM needs (EPI_LN_*), for rows that were not produced by a residual
// GEMM epilogue (start of a tower).  One warp per row; slot 0 of the statistics row is written.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kEwThreads) rowstats_cast_kernel(const float* __restrict__ x, int64_t rows,
                                                                   __nv_bfloat16* __restrict__ xb,
                                                                   float2* __restrict__ stats, int f16) {
  constexpr int V = D / 128;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * D);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 v = xr[lane + 32 * j];
      s1 += (v.x + v.y) + (v.z + v.w);
      s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      uint2 u;
      u.x = pack_op2_rt(v.x, v.y, f16);
      u.y = pack_op2_rt(v.z, v.w, f16);
      reinterpret_cast<uint2*>(xb + r * D)[lane + 32 * j] = u;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) stats[r * kStatSlots] = make_float2(s1, s2);
  }
}

// ------------------------------------------------------------------------------------------------
// Text embeddings: x[b*S + t] = token_embedding[ids[b,t]] + position_embedding[t]   (TF:253-256).
// One warp per token row (512 fp32 = 4 float4 per lane).  Ids are clamped into the vocabulary for
// memory safety (the reference would raise an IndexError on out-of-range ids).
// ------------------------------------------------------------------------------------------------
template <typename IdT>
__global__ void __launch_bounds__(kEwThreads) text_embed_kernel(const IdT* __restrict__ ids, int64_t n,
                                                                int seq_len, int ids_stride,
                                                                const float* __restrict__ tok,
                                                                const float* __restrict__ pos,
                                                                float* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = n * seq_len;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const int t = (int)(r % seq_len);
    long long id = (long long)ids[(r / seq_len) * ids_stride + t];  // rows may be a prefix of longer id rows
    id = id < 0 ? 0 : (id >= kVocab ? kVocab - 1 : id);
    const float4* e = reinterpret_cast<const float4*>(tok + id * kTxtDim);
    const float4* p = reinterpret_cast<const float4*>(pos + (int64_t)t * kTxtDim);
    float4* o = reinterpret_cast<float4*>(x + r * kTxtDim);
#pragma unroll
    for (int j = 0; j < kTxtDim / 128; ++j) {
      const float4 a = __ldg(e + lane + 32 * j), b = __ldg(p + lane + 32 * j);
      o[lane + 32 * j] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
  }
}

// Pooling row of each caption: b*S + (first t with ids[b,t] == eos, else 0) — the semantics of
// (input_ids == eos_token_id).int().argmax(-1) (TF:571-584).  One warp per caption.
template <typename IdT>
__global__ void __launch_bounds__(kEwThreads) eos_row_kernel(const IdT* __restrict__ ids, int64_t n,
                                                             int seq_len, int ids_stride, int eos_id,
                                                             int no_eos_argmax, int32_t* __restrict__ row_index) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < n; b += nwarps) {
    int first = seq_len;
    for (int t0 = 0; t0 < seq_len && first == seq_len; t0 += 32) {
      const int t = t0 + lane;
      const bool hit = t < seq_len && (long long)ids[b * ids_stride + t] == (long long)eos_id;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) first = t0 + __ffs(m) - 1;
    }
    if (first == seq_len) {
      // no eos in the row.  HF with eos_token_id == 49407 pools position 0 here ((ids == eos).argmax() of all zeros,
      // TF:571-584); legacy configs (eos_token_id == 2, what openai/clip-vit-base-patch32 ships) and OpenAI clip pool
      // the first position of the largest id (TF:564-570) — selected by plip_set_text_pooling.
      first = 0;
      if (no_eos_argmax) {
        long long best = -(1ll << 62);
        int best_t = 0;
        for (int t = lane; t < seq_len; t += 32) {
          const long long v = (long long)ids[b * ids_stride + t];
          if (v > best) { best = v; best_t = t; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const long long ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int ot = __shfl_xor_sync(0xffffffffu, best_t, o);
          if (ov > best || (ov == best && ot < best_t)) { best = ov; best_t = ot; }
        }
        first = best_t;
      }
    }
    if (lane == 0) row_index[b] = (int32_t)(b * seq_len + first);
  }
}

// Key-padding mask -> int32 (1 = attend, 0 = padded key).
template <typename IdT>
__global__ void __launch_bounds__(kEwThreads) mask_to_i32_kernel(const IdT* __restrict__ m, int64_t count,
                                                                 int seq_len, int stride,
                                                                 int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = m[(i / seq_len) * stride + (i % seq_len)] != 0 ? 1 : 0;
}

// Class-token rows: x[b*50] = class_embedding + position_embedding[0]   (TF:212-217).
__global__ void __launch_bounds__(kEwThreads) cls_rows_kernel(const float* __restrict__ cls,
                                                              const float* __restrict__ pos, int64_t n,
                                                              float* __restrict__ x) {
  constexpr int V = kVisDim / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * V;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / V;
    const int j = (int)(i % V);
    const float4 c = __ldg(reinterpret_cast<const float4*>(cls) + j);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + j);
    reinterpret_cast<float4*>(x + b * kVisSeq * kVisDim)[j] =
        make_float4(c.x + p.x, c.y + p.y, c.z + p.z, c.w + p.w);
  }
}

// Pooled-row gather for the pruned last layer (engine.cu run_layers): copies row idx(i) of the 16-bit attention
// output and of the fp32 residual stream into compact [n, dim] buffers.  idx(i) = row_index[i], or i * row_stride
// when row_index is null (the vision CLS rows).  16-byte pieces, one per thread.
__global__ void __launch_bounds__(kEwThreads) gather_rows_kernel(const uint4* __restrict__ a16, const uint4* __restrict__ x32,
                                                                 const int32_t* __restrict__ row_index, int64_t row_stride,
                                                                 int64_t n, int dim, uint4* __restrict__ a16_out,
                                                                 uint4* __restrict__ x32_out) {
  const int v16 = dim / 8, v32 = dim / 4, per_row = v16 + v32;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * per_row;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / per_row;
    const int j = (int)(i % per_row);
    const int64_t src = row_index ? (int64_t)row_index[r] : r * row_stride;
    if (j < v16) a16_out[r * v16 + j] = a16[src * v16 + j];
    else x32_out[r * v32 + (j - v16)] = x32[src * v32 + (j - v16)];
  }
}

// x[r] /= sqrt(sum x[r]^2): _get_vector_norm, no epsilon (TF:57-65,923-924). One warp per row.
__global__ void __launch_bounds__(kEwThreads) l2_normalize_kernel(float* __restrict__ x, int64_t rows,
                                                                  int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    float* xr = x + r * dim;
    float s = 0.f;
    for (int j = lane; j < dim; j += 32) s += xr[j] * xr[j];
    const float inv = 1.0f / sqrtf(warp_sum(s));
    for (int j = lane; j < dim; j += 32) xr[j] *= inv;
  }
}

inline int grid_for(int64_t work_items, int per_block) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  const int64_t cap = 148LL * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

int launch_im2col(const void* pixels, int fmt, int64_t n, __nv_bfloat16* out, int f16, cudaStream_t st) {
  PLIP_REQUIRE(n > 0, "im2col: n must be positive");
  PLIP_REQUIRE((reinterpret_cast<uintptr_t>(pixels) & 15) == 0, "im2col: pixels must be 16-byte aligned");
  const int64_t items = (fmt == PLIP_PIX_U8_NHWC ? 1 : 3) * n * kImage * (kImage / 8);
  const int grid = grid_for(items, kEwThreads);
  switch (fmt) {
    case PLIP_PIX_F32_NCHW: PLIP_CUDA_CHECK(launch_kernel(im2col_kernel<PLIP_PIX_F32_NCHW>, dim3(grid), dim3(kEwThreads), 0, st, 1, pixels, out, n, f16)); break;
    case PLIP_PIX_BF16_NCHW: PLIP_CUDA_CHECK(launch_kernel(im2col_kernel<PLIP_PIX_BF16_NCHW>, dim3(grid), dim3(kEwThreads), 0, st, 1, pixels, out, n, f16)); break;
    case PLIP_PIX_U8_NHWC: PLIP_CUDA_CHECK(launch_kernel(im2col_kernel<PLIP_PIX_U8_NHWC>, dim3(grid), dim3(kEwThreads), 0, st, 1, pixels, out, n, f16)); break;
    default: set_last_error("im2col: unknown pixel format %d", fmt); return -2;
  }
  PLIP_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int launch_layernorm(const float* x, const int32_t* row_index, int64_t in_row_stride, int64_t rows, int dim,
                     const float* gamma, const float* beta, float* out_f32, __nv_bfloat16* out_bf16, int f16,
                     cudaStream_t st) {
  PLIP_REQUIRE(rows > 0, "layernorm: rows must be positive");
  PLIP_REQUIRE(in_row_stride % 4 == 0, "layernorm: row stride must be a multiple of 4 floats");
  const int grid = grid_for(rows, kEwThreads / 32);
  if (dim == kVisDim)
    PLIP_CUDA_CHECK(launch_kernel(layernorm_kernel<kVisDim>, dim3(grid), dim3(kEwThreads), 0, st, 1, x, row_index, in_row_stride, rows, gamma, beta, out_f32, out_bf16, f16));
  else if (dim == kTxtDim)
    PLIP_CUDA_CHECK(launch_kernel(layernorm_kernel<kTxtDim>, dim3(grid), dim3(kEwThreads), 0, st, 1, x, row_index, in_row_stride, rows, gamma, beta, out_f32, out_bf16, f16));
  else {
    set_last_error("layernorm: unsupported dim %d (768 or 512)", dim);
    return -2;
  }
  PLIP_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int launch_rowstats_cast(const float* x, int64_t rows, int dim, __nv_bfloat16* xb, float2* stats, int f16, cudaStream_t st) {
  PLIP_REQUIRE(rows > 0, "rowstats_cast: rows must be positive");
  const int grid = grid_for(rows, kEwThreads / 32);
  if (dim == kVisDim)
    PLIP_CUDA_CHECK(launch_kernel(rowstats_cast_kernel<kVisDim>, dim3(grid), dim3(kEwThreads), 0, st, 1, x, rows, xb, stats, f16));
  else if (dim == kTxtDim)
    PLIP_CUDA_CHECK(launch_kernel(rowstats_cast_kernel<kTxtDim>, dim3(grid), dim3(kEwThreads), 0, st, 1, x, rows, xb, stats, f16));
  else {
    set_last_error("rowstats_cast: unsupported dim %d", dim);
    return -2;
  }
  ++g_launch_count;
  return 0;
}

int launch_text_embed(const void* ids, int ids_dtype, int64_t n, int seq_len, int ids_stride, const float* tok,
                      const float* pos, float* x, int32_t* eos_rows, int eos_id, int no_eos_argmax, cudaStream_t st) {
  PLIP_REQUIRE(ids_stride >= seq_len, "text_embed: ids row stride %d < seq_len %d", ids_stride, seq_len);
  PLIP_REQUIRE(n > 0 && seq_len > 0 && seq_len <= kTxtSeq, "text_embed: bad shape n=%lld seq_len=%d",
               (long long)n, seq_len);
  const int grid = grid_for(n * seq_len, kEwThreads / 32);
  const int grid2 = grid_for(n, kEwThreads / 32);
  if (ids_dtype == PLIP_IDS_I64) {
    PLIP_CUDA_CHECK(launch_kernel(text_embed_kernel<long long>, dim3(grid), dim3(kEwThreads), 0, st, 1, static_cast<const long long*>(ids), n, seq_len, ids_stride, tok, pos, x));
    PLIP_CUDA_CHECK(launch_kernel(eos_row_kernel<long long>, dim3(grid2), dim3(kEwThreads), 0, st, 1, static_cast<const long long*>(ids), n, seq_len, ids_stride, eos_id, no_eos_argmax, eos_rows));
  } else if (ids_dtype == PLIP_IDS_I32) {
    PLIP_CUDA_CHECK(launch_kernel(text_embed_kernel<int>, dim3(grid), dim3(kEwThreads), 0, st, 1, static_cast<const int*>(ids), n, seq_len, ids_stride, tok, pos, x));
    PLIP_CUDA_CHECK(launch_kernel(eos_row_kernel<int>, dim3(grid2), dim3(kEwThreads), 0, st, 1, static_cast<const int*>(ids), n, seq_len, ids_stride, eos_id, no_eos_argmax, eos_rows));
  } else {
    set_last_error("text_embed: unknown ids dtype %d", ids_dtype);
    return -2;
  }
  PLIP_CUDA_CHECK(cudaGetLastError());
  g_launch_count += 2;
  return 0;
}

int launch_mask_to_i32(const void* mask, int dtype, int64_t count, int seq_len, int stride, int32_t* out,
                       cudaStream_t st) {
  const int grid = grid_for(count, kEwThreads);
  if (dtype == PLIP_IDS_I64)
    PLIP_CUDA_CHECK(launch_kernel(mask_to_i32_kernel<long long>, dim3(grid), dim3(kEwThreads), 0, st, 1, static_cast<const long long*>(mask), count, seq_len, stride, out));
  else
    PLIP_CUDA_CHECK(launch_kernel(mask_to_i32_kernel<int>, dim3(grid), dim3(kEwThreads), 0, st, 1, static_cast<const int*>(mask), count, seq_len, stride, out));
  PLIP_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int launch_cls_rows(const float* cls, const float* pos, int64_t n, float* x, cudaStream_t st) {
  PLIP_CUDA_CHECK(launch_kernel(cls_rows_kernel, dim3(grid_for(n * (kVisDim / 4), kEwThreads)), dim3(kEwThreads), 0, st, 1, cls, pos, n, x));
  PLIP_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int launch_gather_rows(const __nv_bfloat16* a16, const float* x32, const int32_t* row_index, int64_t row_stride,
                       int64_t n, int dim, __nv_bfloat16* a16_out, float* x32_out, cudaStream_t st) {
  PLIP_REQUIRE(n > 0 && dim > 0 && dim % 8 == 0, "gather_rows: bad shape");
  const int64_t items = n * (dim / 8 + dim / 4);
  PLIP_CUDA_CHECK(launch_kernel(gather_rows_kernel, dim3(grid_for(items, kEwThreads)), dim3(kEwThreads), 0, st, 1,
                                reinterpret_cast<const uint4*>(a16), reinterpret_cast<const uint4*>(x32), row_index,
                                row_stride, n, dim, reinterpret_cast<uint4*>(a16_out), reinterpret_cast<uint4*>(x32_out)));
  PLIP_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int launch_l2_normalize(float* x, int64_t rows, int dim, cudaStream_t st) {
  PLIP_REQUIRE(rows > 0 && dim > 0, "l2_normalize: bad shape");
  PLIP_CUDA_CHECK(launch_kernel(l2_normalize_kernel, dim3(grid_for(rows, kEwThreads / 32)), dim3(kEwThreads), 0, st, 1, x, rows, dim));
  PLIP_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

}  // namespace plip
