// plip_b200 — similarity head: L2-normalise + scale * A . B^T in one fp32 kernel, and a fused top-k.
//
// Replaces   x / _get_vector_norm(x);  text @ image.T * exp(logit_scale);  .t()   (TF:modeling_clip.py:57-65,923-930)
// and the numpy heads  key.dot(space.T) + argmax / argsort top-k
//   (plip.py:73-87,99-102; evaluation/zero_shot/zero_shot.py:12-13; evaluation/retrieval/retrieval.py:13-16).
//
// The [n,512] x [m,512]^T product is kept in fp32 FMA arithmetic (K = 512 only): the |dlogits| <= 1e-3
// bar at logit scales up to 100 rules out 16-bit operands.  The row norms are
// accumulated from the same operand tiles that feed the product, so each input is read once per tile.
#include "kernels.cuh"

namespace plip {

namespace {

constexpr int kSimThreads = 256;
constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(kSimThreads)
similarity_kernel(const float* __restrict__ A, int64_t n, const float* __restrict__ B, int64_t m, int K,
                  float scale, int norm_a, int norm_b, float* __restrict__ C, int64_t ldc) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ float inv_a[TM], inv_b[TN];

  pdl_wait();
  pdl_launch_dependents();
  const int t = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.y * TM;
  const int64_t col0 = (int64_t)blockIdx.x * TN;
  const int lr = t >> 2;         // tile row loaded by this thread (0..63)
  const int lk = (t & 3) * 4;    // k offset inside the TK slab
  const int ty = t >> 4, tx = t & 15;

  float acc[4][4] = {};
  float ssa = 0.f, ssb = 0.f;
  const bool a_ok = row0 + lr < n, b_ok = col0 + lr < m;
  const float* ap = A + (row0 + lr) * K + lk;
  const float* bp = B + (col0 + lr) * K + lk;

  for (int k0 = 0; k0 < K; k0 += TK) {
    float4 av = a_ok ? __ldg(reinterpret_cast<const float4*>(ap + k0)) : make_float4(0, 0, 0, 0);
    float4 bv = b_ok ? __ldg(reinterpret_cast<const float4*>(bp + k0)) : make_float4(0, 0, 0, 0);
    ssa += av.x * av.x + av.y * av.y + av.z * av.z + av.w * av.w;
    ssb += bv.x * bv.x + bv.y * bv.y + bv.z * bv.z + bv.w * bv.w;
    __syncthreads();
    As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Bs[lk + 0][lr] = bv.x; Bs[lk + 1][lr] = bv.y; Bs[lk + 2][lr] = bv.z; Bs[lk + 3][lr] = bv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w};
      const float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
  // the 4 threads that loaded one row are adjacent lanes
  ssa += __shfl_xor_sync(0xffffffffu, ssa, 1);
  ssa += __shfl_xor_sync(0xffffffffu, ssa, 2);
  ssb += __shfl_xor_sync(0xffffffffu, ssb, 1);
  ssb += __shfl_xor_sync(0xffffffffu, ssb, 2);
  if ((t & 3) == 0) {
    inv_a[lr] = norm_a ? 1.0f / sqrtf(ssa) : 1.0f;
    inv_b[lr] = norm_b ? 1.0f / sqrtf(ssb) : 1.0f;
  }
  __syncthreads();

  const bool vec_ok = (ldc % 4 == 0) && (col0 + tx * 4 + 3 < m);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    if (r >= n) continue;
    const float sa = scale * inv_a[ty * 4 + i];
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = acc[i][j] * sa * inv_b[tx * 4 + j];
    float* cp = C + r * ldc + col0 + tx * 4;
    if (vec_ok) {
      *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col0 + tx * 4 + j < m) cp[j] = o[j];
    }
  }
}

// ---- fused similarity + top-k (round-1 version: one CTA per query, streaming the space) ---------
constexpr int kTopkThreads = 128;
constexpr int kTopkMax = 64;

__device__ __forceinline__ bool better(float s, int i, float s2, int i2) {
  return s > s2 || (s == s2 && i < i2);
}

__global__ void __launch_bounds__(kTopkThreads)
similarity_topk_kernel(const float* __restrict__ Q, int64_t n, const float* __restrict__ Sp, int64_t m, int K,
                       float scale, int norm_q, int norm_s, int k, int32_t* __restrict__ idx,
                       float* __restrict__ val) {
  __shared__ float q[kProj];
  __shared__ float ls[4][kTopkMax];
  __shared__ int li[4][kTopkMax];
  __shared__ float red[4];
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x;

  float ss = 0.f;
  for (int j = threadIdx.x; j < K; j += kTopkThreads) {
    const float v = Q[row * K + j];
    q[j] = v;
    ss += v * v;
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  for (int j = lane; j < k; j += 32) {
    ls[warp][j] = -INFINITY;
    li[warp][j] = 0x7fffffff;
  }
  __syncthreads();
  const float qs = scale * (norm_q ? 1.0f / sqrtf(red[0] + red[1] + red[2] + red[3]) : 1.0f);

  for (int64_t c = warp; c < m; c += 4) {
    const float* sp = Sp + c * K;
    float dot = 0.f, s2 = 0.f;
    for (int j = lane * 4; j < K; j += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(sp + j));
      const float4 w = *reinterpret_cast<const float4*>(&q[j]);
      dot = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, dot))));
      s2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    dot = warp_sum(dot);
    s2 = warp_sum(s2);
    const float score = dot * qs * (norm_s ? 1.0f / sqrtf(s2) : 1.0f);
    if (lane == 0 && better(score, (int)c, ls[warp][k - 1], li[warp][k - 1])) {
      int pos = k - 1;  // insertion into the warp's descending list
      while (pos > 0 && better(score, (int)c, ls[warp][pos - 1], li[warp][pos - 1])) {
        ls[warp][pos] = ls[warp][pos - 1];
        li[warp][pos] = li[warp][pos - 1];
        --pos;
      }
      ls[warp][pos] = score;
      li[warp][pos] = (int)c;
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int head[4] = {0, 0, 0, 0};
    for (int o = 0; o < k; ++o) {
      int bw = -1;
      for (int w = 0; w < 4; ++w) {
        if (head[w] >= k) continue;
        if (bw < 0 || better(ls[w][head[w]], li[w][head[w]], ls[bw][head[bw]], li[bw][head[bw]])) bw = w;
      }
      const int id = li[bw][head[bw]];
      idx[row * k + o] = (id == 0x7fffffff) ? -1 : id;
      if (val) val[row * k + o] = ls[bw][head[bw]];
      ++head[bw];
    }
  }
}

}  // namespace

int launch_similarity(const float* a, int64_t n, const float* b, int64_t m, float scale, bool norm_a, bool norm_b,
                      float* out, int64_t ldo, cudaStream_t st) {
  PLIP_REQUIRE(n > 0 && m > 0, "similarity: empty operand n=%lld m=%lld", (long long)n, (long long)m);
  PLIP_REQUIRE(ldo >= m, "similarity: ld_logits %lld < m %lld", (long long)ldo, (long long)m);
  PLIP_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(out) & 15) == 0, "similarity: operands must be 16-byte aligned");
  const int64_t gy = (n + TM - 1) / TM, gx = (m + TN - 1) / TN;
  PLIP_REQUIRE(gy <= 65535, "similarity: n=%lld too large for one launch (chunk rows)", (long long)n);
  dim3 grid((unsigned)gx, (unsigned)gy);
  PLIP_CUDA_CHECK(launch_pdl(similarity_kernel, grid, dim3(kSimThreads), 0, st, 1, a, n, b, m, (int)kProj, scale,
                             norm_a ? 1 : 0, norm_b ? 1 : 0, out, ldo));
  ++g_launch_count;
  return 0;
}

int launch_similarity_topk(const float* q, int64_t n, const float* s, int64_t m, float scale, bool norm_q,
                           bool norm_s, int k, int32_t* idx, float* val, cudaStream_t st) {
  PLIP_REQUIRE(n > 0 && m > 0, "similarity_topk: empty operand");
  PLIP_REQUIRE(k >= 1 && k <= kTopkMax, "similarity_topk: k=%d out of range [1,%d]", k, kTopkMax);
  PLIP_REQUIRE(n <= 0x7fffffff && m <= 0x7fffffff, "similarity_topk: operand too large");
  PLIP_CUDA_CHECK(launch_pdl(similarity_topk_kernel, dim3((unsigned)n), dim3(kTopkThreads), 0, st, 1, q, n, s, m,
                             (int)kProj, scale, norm_q ? 1 : 0, norm_s ? 1 : 0, k, idx, val));
  ++g_launch_count;
  return 0;
}

}  // namespace plip
