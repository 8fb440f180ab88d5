// plip_b200 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   acc[M,N] = A[M,K] (bf16, K-major) x W[N,K]^T (bf16, K-major), fp32 accumulation in TMEM,
//   fused epilogue (bias / QuickGELU / fp32 residual add / patch-embedding scatter).
//
// Replaces, for the PLIP (CLIP ViT-B/32) hot path, the cuBLAS/cuDNN calls behind
//   nn.Conv2d patch embedding   TF:modeling_clip.py:148-154,209
//   q/k/v_proj, out_proj        TF:modeling_clip.py:310-312,334
//   fc1 + QuickGELU, fc2        TF:modeling_clip.py:347-351, TF:activations.py:117-123
//   visual/text_projection      TF:modeling_clip.py:861,823
//
// Structure (one CTA per SM, 256 threads, persistent over output tiles):
//   warp 0   TMA producer: A/W k-blocks (64 bf16 = one 128B-swizzle atom wide) -> smem ring
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma (UMMA 128xBNx16, or 256xBNx16 for a
//            CTA pair), releasing smem stages and publishing accumulators through tcgen05.commit
//   warp 2   TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4-7 epilogue: tcgen05.ld (thread == accumulator row) -> fused math -> global stores,
//            overlapped with the MMAs of the next tile through the second accumulator stage
// CG == 2 pairs two SMs (cta_group::2, cluster (2,1,1)): each CTA stages its 128 rows of A and its
// half of the W tile, the leader CTA issues the 256-row MMA, halving per-SM L2->smem operand traffic.
#include "gemm.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace plip {

unsigned long long g_launch_count = 0;

namespace {

constexpr int BM = 128;  // accumulator rows per CTA (TMEM lanes)
constexpr int BK = 64;   // k-block: 64 bf16 = 128 B = one swizzle atom
constexpr int kThreads = 256;
constexpr uint32_t A_STAGE = BM * BK * 2;

template <int CG, int BN>
struct Cfg {
  static constexpr int LOAD_N = BN / CG;  // W rows staged by each CTA
  static constexpr uint32_t B_STAGE = LOAD_N * BK * 2;
  static constexpr uint32_t STAGE = A_STAGE + B_STAGE;
  static constexpr int kMaxStages = (227 * 1024 - 1024 - 512) / STAGE;
  static constexpr int STAGES = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE + 1024 + 512;
};

struct GemmDev {
  int M, N, K;
  const float* bias;
  void* out;
  int ldo;
  const float* pos;
};

template <int EPI>
__device__ __forceinline__ void epilogue_store(const GemmDev& p, int row, int col0,
                                               const uint32_t (&v)[32]) {
  if (row >= p.M) return;
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);

  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RESID_F32) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(b4 + i);
      f[4 * i + 0] += b.x;
      f[4 * i + 1] += b.y;
      f[4 * i + 2] += b.z;
      f[4 * i + 3] += b.w;
    }
  }
  if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = quick_gelu(f[i]);
  }

  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
    uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]);
      u.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
      u.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
      u.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
      o4[i] = u;
    }
  } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
    float4* x4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                           static_cast<size_t>(row) * p.ldo + col0);
    float4 r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = x4[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      r[i].x += f[4 * i + 0];
      r[i].y += f[4 * i + 1];
      r[i].z += f[4 * i + 2];
      r[i].w += f[4 * i + 3];
      x4[i] = r[i];
    }
  } else if constexpr (EPI == EPI_PATCH_F32) {
    const int b = row / kPatches;
    const int pp = row - b * kPatches;
    const float4* pos4 =
        reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(1 + pp) * p.N + col0);
    float4* x4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                           static_cast<size_t>(b * kVisSeq + 1 + pp) * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 q = __ldg(pos4 + i);
      x4[i] = make_float4(f[4 * i + 0] + q.x, f[4 * i + 1] + q.y, f[4 * i + 2] + q.z,
                          f[4 * i + 3] + q.w);
    }
  } else {  // EPI_F32
    float4* x4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                           static_cast<size_t>(row) * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      x4[i] = make_float4(f[4 * i + 0], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  }
}

template <int CG, int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmDev p) {
  using C = Cfg<CG, BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (smem_raw_u32 + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * C::STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), CG);  // leader's arrive.expect_tx (+ peer's remote arrive)
      mbar_init(empty_bar(s), 1);  // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);        // tcgen05.commit
      mbar_init(tempty_bar(a), 4 * CG);  // one arrive per epilogue warp (both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base =
      *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_raw_u32));

  const int num_n_blk = p.N / BN;
  const int num_m_blk = (p.M + BM * CG - 1) / (BM * CG);
  const int num_tiles = num_m_blk * num_n_blk;
  const int num_kb = p.K / BK;
  const int tile0 = blockIdx.x / CG;
  const int tile_step = gridDim.x / CG;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = tile0; t < num_tiles; t += tile_step) {
        const int m_blk = t / num_n_blk, n_blk = t - m_blk * num_n_blk;
        const int m0 = m_blk * BM * CG + cta_rank * BM;
        const int n0 = n_blk * BN + cta_rank * C::LOAD_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t sa = smem_base + s * C::STAGE;
          const uint32_t sb = sa + A_STAGE;
          if constexpr (CG == 1) {
            mbar_arrive_expect_tx(full_bar(s), C::STAGE);
            tma_load_2d(sa, &tmA, full_bar(s), kb * BK, m0);
            tma_load_2d(sb, &tmB, full_bar(s), kb * BK, n0);
          } else {
            const uint32_t lfull = mapa_shared(full_bar(s), 0);
            if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * C::STAGE);
            else mbar_arrive_cluster(lfull);
            tma_load_2d_cg2(sa, &tmA, lfull, kb * BK, m0);
            tma_load_2d_cg2(sb, &tmB, lfull, kb * BK, n0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN, 0, 0);
      int s = 0, a = 0;
      uint32_t ph = 0, aph = 0;
      for (int t = tile0; t < num_tiles; t += tile_step) {
        mbar_wait(tempty_bar(a), aph ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + a * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * C::STAGE;
          const uint64_t adesc = make_smem_desc_sw128(sa, 1024, 16);
          const uint64_t bdesc = make_smem_desc_sw128(sa + A_STAGE, 1024, 16);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_ss<CG>(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit<CG>(empty_bar(s));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        umma_commit<CG>(tfull_bar(a));
        a ^= 1;
        if (a == 0) aph ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may access
    int a = 0;
    uint32_t aph = 0;
    for (int t = tile0; t < num_tiles; t += tile_step) {
      const int m_blk = t / num_n_blk, n_blk = t - m_blk * num_n_blk;
      mbar_wait(tfull_bar(a), aph);
      tc_fence_after();
      const int row = m_blk * BM * CG + cta_rank * BM + q * 32 + lane;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tbase + c * 32, v);
        tmem_ld_wait();
        epilogue_store<EPI>(p, row, n_blk * BN + c * 32, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(tempty_bar(a));
        else mbar_arrive_cluster(mapa_shared(tempty_bar(a), 0));
      }
      a ^= 1;
      if (a == 0) aph ^= 1u;
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int CG, int BN, int EPI>
int launch_inst(const GemmArgs& g, cudaStream_t stream) {
  using C = Cfg<CG, BN>;
  auto kern = gemm_kernel<CG, BN, EPI>;
  static bool configured = false;
  if (!configured) {
    PLIP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM_BYTES));
    configured = true;
  }
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_bf16_2d(&tmA, g.A, g.M, g.K, (uint64_t)g.lda * 2, BM, BK)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmB, g.W, g.N, g.K, (uint64_t)g.ldw * 2, C::LOAD_N, BK)) return rc;

  GemmDev p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.bias = g.bias; p.out = g.out; p.ldo = g.ldo; p.pos = g.pos;

  const int num_tiles = ((g.M + BM * CG - 1) / (BM * CG)) * (g.N / BN);
  int groups = num_sms() / CG;
  if (groups > num_tiles) groups = num_tiles;

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CG);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PLIP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
  ++g_launch_count;
  return 0;
}

template <int CG, int BN>
int launch_epi(const GemmArgs& g, cudaStream_t stream) {
  switch (g.epi) {
    case EPI_BIAS_BF16: return launch_inst<CG, BN, EPI_BIAS_BF16>(g, stream);
    case EPI_BIAS_GELU_BF16: return launch_inst<CG, BN, EPI_BIAS_GELU_BF16>(g, stream);
    case EPI_BIAS_RESID_F32: return launch_inst<CG, BN, EPI_BIAS_RESID_F32>(g, stream);
    case EPI_PATCH_F32: return launch_inst<CG, BN, EPI_PATCH_F32>(g, stream);
    case EPI_F32: return launch_inst<CG, BN, EPI_F32>(g, stream);
    default: set_last_error("launch_gemm: bad epilogue %d", g.epi); return -2;
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  PLIP_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "launch_gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  PLIP_REQUIRE(g.K % BK == 0, "launch_gemm: K=%d must be a multiple of %d", g.K, BK);
  PLIP_REQUIRE(g.N % 128 == 0, "launch_gemm: N=%d must be a multiple of 128", g.N);
  PLIP_REQUIRE((g.lda % 8) == 0 && (g.ldw % 8) == 0 && (g.ldo % 8) == 0,
               "launch_gemm: leading dimensions must be multiples of 8 elements");
  PLIP_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.W) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(g.out) & 15) == 0,
               "launch_gemm: operands must be 16-byte aligned");
  static const int env_cg = env_int("PLIP_GEMM_CG", 0);
  static const int env_bn = env_int("PLIP_GEMM_BN", 0);
  int cg = g.force_cg ? g.force_cg : (env_cg ? env_cg : 2);
  int bn = g.force_bn ? g.force_bn : (env_bn ? env_bn : 256);
  if (g.N % bn != 0) bn = 128;
  PLIP_REQUIRE((cg == 1 || cg == 2) && (bn == 128 || bn == 256), "launch_gemm: bad config cg=%d bn=%d", cg, bn);
  if (cg == 1) return bn == 256 ? launch_epi<1, 256>(g, stream) : launch_epi<1, 128>(g, stream);
  return bn == 256 ? launch_epi<2, 256>(g, stream) : launch_epi<2, 128>(g, stream);
}

}  // namespace plip
