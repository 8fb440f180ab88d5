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
// Structure (one CTA per SM, 384 threads, persistent over output tiles):
//   warps 0-7 epilogue: tcgen05.ld (thread == accumulator row) -> fused math -> global stores,
//            overlapped with the MMAs of the next tile through the second accumulator stage
//   warp 8   TMA producer: A/W k-blocks (64 bf16 = one 128B-swizzle atom wide) -> smem ring
//   warp 9   MMA issuer: one elected thread issues tcgen05.mma (UMMA 128xBNx16, or 256xBNx16 for a
//            CTA pair), releasing smem stages and publishing accumulators through tcgen05.commit
//   warp 10  TMEM allocator (2 accumulator stages of BN fp32 columns)
// (the control warps carry the highest warp ids on purpose: the sub-partition arbiter prefers them)
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
constexpr int kThreads = 384;      // 8 epilogue warps (0-7) + 4 control warps (8-11)
constexpr int kWarpTma = 8, kWarpMma = 9, kWarpTmem = 10;  // highest warp ids: the SM sub-partition arbiter
                                                           // favours them over the instruction-heavy epilogue warps
constexpr int kEpiWarps = 8;       // two per TMEM lane quarter, splitting the tile's column blocks
constexpr uint32_t A_STAGE = BM * BK * 2;

template <int CG, int BN, int EPI = EPI_LN_BIAS_BF16>
struct Cfg {
  static constexpr int LOAD_N = BN / CG;  // W rows staged by each CTA
  static constexpr uint32_t B_STAGE = LOAD_N * BK * 2;
  static constexpr uint32_t STAGE = A_STAGE + B_STAGE;
  static constexpr bool LN_FOLD = (EPI == EPI_LN_BIAS_BF16 || EPI == EPI_LN_BIAS_GELU_BF16);
  // per-warp staging blocks + per-warp bias slice (+ colsum slice for the LN fold): a warp owns BN / 2 columns
  // (a warp's share is the larger half of the tile's 64-column blocks: BN = 192 splits 2 + 1)
  static constexpr int VEC_FLOATS = ((BN / 64 + 1) / 2) * 64;
  static constexpr uint32_t VEC_BYTES = (LN_FOLD ? 2 : 1) * VEC_FLOATS * 4;  // per warp
  // Short-K residual GEMMs (out_proj: N tile 192 / 128) are paced by the fp32 read-modify-write of the residual tile,
  // not by the MMAs: they double-buffer the residual blocks in shared memory with cp.async, one 32-column block ahead
  // (and the first block of a tile while its MMAs still run), at the price of operand-ring stages they do not need.
#if defined(PLIP_NO_RPF)   // A/B build switches (tools/r2_call7.sh, r2_call8.sh)
  static constexpr bool RPF = false;
#elif defined(PLIP_RPF_ALL)
  static constexpr bool RPF = (EPI == EPI_BIAS_RESID_F32);
#else
  static constexpr bool RPF = (EPI == EPI_BIAS_RESID_F32) && (BN < 256);
#endif
  static constexpr uint32_t RPF_BYTES = RPF ? kEpiWarps * 2 * 4096 : 0;
  static constexpr uint32_t EPI_BYTES = kEpiWarps * 32 * 128 + kEpiWarps * VEC_BYTES + RPF_BYTES;
  static constexpr uint32_t BAR_BYTES = 256;
  // The dynamic smem window starts 1024-aligned (checked at kernel entry), so no alignment slack is
  // reserved: that is what lets the residual epilogues run a 6-deep 32 KB operand ring.
  static constexpr int kMaxStages = (227 * 1024 - BAR_BYTES - EPI_BYTES) / STAGE;
  static constexpr int STAGES = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE + EPI_BYTES + BAR_BYTES;
};

struct GemmDev {
  int M, N, K;
  int tma_store;  // bf16 epilogues: write the staged blocks with TMA bulk stores instead of read-back + STG
  int f32_serial;  // A/B switch (PLIP_GEMM_F32_SERIAL=1): fp32-output epilogues load each block from TMEM right before its math
  int dbg;  // diagnostic (PLIP_GEMM_DBG): 1 = no global stores, 2 = no staging and no stores, 3 = no math either
  const float* bias;
  const float* rowscale;
  void* out;
  int ldo;
  const float* pos;
  const float* colsum;
  const float2* stats_in;
  int n_partials;
  __nv_bfloat16* xb_out;
  float2* stats_out;
};

// ---- epilogue ---------------------------------------------------------------------------------
// A thread owns one accumulator row (TMEM lane).  Writing rows straight to global memory makes every
// warp-level store touch 32 different cache lines; instead each warp stages a [32 rows x 128 B] block
// in shared memory (16-byte chunks XOR-swizzled by row, the SWIZZLE_128B pattern) and reads it back
// with 8 lanes per row, so each global access covers 4 rows x 128 contiguous bytes.
constexpr uint32_t kEpiStageBytes = 32 * 128;  // per epilogue warp

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One [32 rows x 32 fp32] block of the residual stream -> this warp's prefetch buffer: lane (rb_row, rb_chunk) copies
// the eight 16-byte pieces it will read back itself (rows past M: zero fill), as one cp.async group.
__device__ __forceinline__ void rpf_issue(const float* x, int ldo, int M, uint32_t buf, int row_base, int col0, int lane) {
  const int rb_row = lane >> 3, rb_chunk = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rb_row;
    const int grow = row_base + r;
    const float* src = x + static_cast<size_t>(grow < M ? grow : 0) * ldo + col0 + rb_chunk * 4;
    cp_async_16(buf + r * 128 + rb_chunk * 16, src, grow < M ? 16 : 0);
  }
  cp_async_commit();
}

// acc (+ bias from the smem bias tile) for 32 consecutive columns of this thread's row, as 16 float2 (packed fp32 math).
// LN_FOLD: rstd * acc + (bias' - (rstd * mean) * colsum)  ==  rstd * (acc - mean * colsum) + bias'; the warp's colsum
// slice is stored Cfg::VEC_FLOATS floats after its bias slice.  `nrm` = -rstd * mean.
template <bool HAS_BIAS, bool LN_FOLD, int BN>
__device__ __forceinline__ void acc_math32(const uint32_t (&v)[32], uint32_t bias_smem, float nrm, float rstd,
                                           float2 (&f)[16]) {
  const float2 nrm2 = make_float2(nrm, nrm), rstd2 = make_float2(rstd, rstd);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (HAS_BIAS) b = ld_shared_f4(bias_smem + 16 * i);  // broadcast read
    const float2 a0 = make_float2(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1]));
    const float2 a1 = make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
    if constexpr (LN_FOLD) {
      const float4 cs = ld_shared_f4(bias_smem + (((BN / 64 + 1) / 2) * 64) * 4 + 16 * i);
      f[2 * i + 0] = __ffma2_rn(rstd2, a0, __ffma2_rn(nrm2, make_float2(cs.x, cs.y), make_float2(b.x, b.y)));
      f[2 * i + 1] = __ffma2_rn(rstd2, a1, __ffma2_rn(nrm2, make_float2(cs.z, cs.w), make_float2(b.z, b.w)));
    } else if constexpr (HAS_BIAS) {
      f[2 * i + 0] = __fadd2_rn(a0, make_float2(b.x, b.y));
      f[2 * i + 1] = __fadd2_rn(a1, make_float2(b.z, b.w));
    } else {
      f[2 * i + 0] = a0;
      f[2 * i + 1] = a1;
    }
  }
}

// One accumulator tile (this warp's 32 rows x BN columns) -> global memory.
// `release_bar`: the accumulator's "empty" barrier (cluster address of the pair leader's for CG == 2).  The 16-bit-output
// path arrives on it itself, as soon as its last TMEM load has landed (before the remaining math and stores), and
// returns true; the other paths return false and the caller releases the accumulator after the call.
template <int CG, int BN, int EPI, bool F16>
__device__ __forceinline__ bool epilogue_tile(const GemmDev& p, const CUtensorMap* tmC, uint32_t tmem_row_base,
                                              uint32_t stage_smem,
                                              uint32_t bias_smem, uint32_t rpf_smem, int row_base, int col_base, int n_blk,
                                              int half, int lane, float ln_mean, float ln_rstd, uint32_t release_bar) {
  constexpr bool LN_FOLD = (EPI == EPI_LN_BIAS_BF16 || EPI == EPI_LN_BIAS_GELU_BF16);
  constexpr bool GELU = (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_LN_BIAS_GELU_BF16);
  constexpr bool HAS_BIAS = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RESID_F32 || LN_FOLD);
  constexpr bool OUT_BF16 = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || LN_FOLD);
  float mean = ln_mean, rstd = ln_rstd;
  if constexpr (EPI == EPI_NULL) {
    uint32_t acc = 0;
#pragma unroll 1
    for (int blk = half; blk < BN / 32; blk += 2) {
      uint32_t v[32];
      tmem_ld32(tmem_row_base + blk * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
    if (acc == 0x12345678u && p.M < 0) reinterpret_cast<uint32_t*>(p.out)[0] = acc;  // keep the loads alive
    return false;
  }
  const uint32_t my_row = stage_smem + lane * 128;
  const int sw = lane & 7;
  const int rb_row = lane >> 3;  // read-back: row within a group of 4
  const int rb_chunk = lane & 7;  // read-back: 16-byte chunk of the row

  if constexpr (OUT_BF16) {
    // Software pipeline over 32-column pieces (r2 probe, profiles/r2_notes.md §9: the epilogue, not the MMAs, paced the
    // short-K GEMMs — all eight warps loaded, then all computed, then all stored, so TMEM-read bandwidth (64 B/clk),
    // MUFU / FMA issue and the store wait never overlapped).  A warp always has the NEXT piece's tcgen05.ld in flight
    // while it runs the math of the current one; the accumulator is handed back to the MMA warp as soon as the last
    // piece has landed in registers, i.e. before that piece's math and the last store.
    constexpr int kMaxBlk = (BN / 64 + 1) / 2;            // 64-column blocks a warp may own (half, half + 2, ...)
    const int nblk = (BN / 64 - half + 1) / 2;
    uint32_t va[32], vb[32];
    tmem_ld32(tmem_row_base + half * 64, va);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < kMaxBlk; ++j) {
      if (j >= nblk) break;
      const int blk = half + 2 * j;
      const bool more = j + 1 < nblk;
      uint32_t pk[32];
      float dbg_acc = 0.f;
      tmem_ld32(tmem_row_base + blk * 64 + 32, vb);       // in flight during the first half's math
      {
        float2 f[16];
        acc_math32<HAS_BIAS, LN_FOLD, BN>(va, bias_smem + (j * 64) * 4, mean, rstd, f);
        if (p.dbg >= 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) dbg_acc += f[i].x + f[i].y;
        } else {
          if constexpr (GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = quick_gelu2(f[i]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_op2<F16>(f[i].x, f[i].y);
        }
      }
      tmem_ld_wait();
      if (more) {
        tmem_ld32(tmem_row_base + (blk + 2) * 64, va);    // next block's first half: in flight during this math + store
      } else {
        // every accumulator column this warp owns is in registers: release the TMEM buffer now
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) mbar_arrive(release_bar);
          else mbar_arrive_cluster(release_bar);
        }
      }
      {
        float2 f[16];
        acc_math32<HAS_BIAS, LN_FOLD, BN>(vb, bias_smem + (j * 64 + 32) * 4, mean, rstd, f);
        if (p.dbg >= 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) dbg_acc += f[i].x + f[i].y;
        } else {
          if constexpr (GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = quick_gelu2(f[i]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[16 + i] = pack_op2<F16>(f[i].x, f[i].y);
        }
      }
      if (p.dbg >= 2) {
        if (dbg_acc == 1.2345e30f) reinterpret_cast<float*>(p.out)[0] = dbg_acc;
      } else {
        if (p.tma_store) {  // the previous bulk store must have finished reading this staging block
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        }
#pragma unroll
        for (int chunk = 0; chunk < 8; ++chunk)
          st_shared_v4(my_row + ((chunk ^ sw) << 4), pk[4 * chunk + 0], pk[4 * chunk + 1], pk[4 * chunk + 2], pk[4 * chunk + 3]);
        if (p.tma_store) {
          // the staging block is laid out exactly as a SWIZZLE_128B [32 rows x 64 bf16] TMA box
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && p.dbg != 1) {
            tma_store_2d(tmC, stage_smem, col_base + blk * 64, row_base);
            tma_store_commit();
          }
        } else {
          __syncwarp();
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + rb_row;
            const uint4 v = ld_shared_v4(stage_smem + r * 128 + ((rb_chunk ^ (r & 7)) << 4));
            const int grow = row_base + r;
            if (grow < p.M && p.dbg != 1)
              *reinterpret_cast<uint4*>(out + static_cast<size_t>(grow) * p.ldo + col_base + blk * 64 + rb_chunk * 8) = v;
          }
          __syncwarp();
        }
      }
      if (more) tmem_ld_wait();
    }
    // (no wait for the last bulk store here: the next tile's first block waits before it touches the staging block,
    //  and the kernel ends with tma_store_wait_all)
    return true;
  } else {
    // per-row (sum, sum of squares) of the updated residual rows this lane writes (rows i*4 + rb_row)
    float st1[8], st2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) st1[i] = st2[i] = 0.f;
    const bool emit = (EPI == EPI_BIAS_RESID_F32) && p.xb_out != nullptr;
#if defined(PLIP_NO_RPF)
    constexpr bool RPF = false;
#elif defined(PLIP_RPF_ALL)
    constexpr bool RPF = (EPI == EPI_BIAS_RESID_F32);
#else
    constexpr bool RPF = (EPI == EPI_BIAS_RESID_F32) && (BN < 256);  // == Cfg::RPF
#endif
    int jblk = 0;
    // same software pipeline as the 16-bit path: the next block's accumulator columns are on their way from TMEM while
    // this block is staged, read back and stored; the accumulator is released once the last block is in registers
    uint32_t vacc[32];
    const bool pipe = p.f32_serial == 0;
    if (pipe) {
      tmem_ld32(tmem_row_base + half * 32, vacc);
      tmem_ld_wait();
    }
#pragma unroll 1
    for (int blk = half; blk < BN / 32; blk += 2, ++jblk) {
      // 32 columns -> 128 B of fp32 per row
      const int col = col_base + blk * 32 + rb_chunk * 4;
      float* out = reinterpret_cast<float*>(p.out);
      const uint32_t rpf_cur = rpf_smem + static_cast<uint32_t>(jblk & 1) * 4096u;
      const bool has_next = blk + 2 < BN / 32;
      if constexpr (RPF) {
        // block j sits (or is landing) in buffer j & 1 — the first one of the tile was issued before the accumulator
        // wait; fetch block j + 1 into the other buffer (its last reader, block j - 1, ended with a __syncwarp)
        if (has_next) rpf_issue(out, p.ldo, p.M, rpf_smem + static_cast<uint32_t>((jblk + 1) & 1) * 4096u, row_base, col_base + (blk + 2) * 32, lane);
      }
      // Residual rows are fetched before the TMEM load / staging so their DRAM latency overlaps it
      // (issued back to back: a load->add->store chain per row serialises 8 round trips per block).
      float4 xr[8];
      if constexpr (EPI == EPI_BIAS_RESID_F32 && !RPF) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int grow = row_base + i * 4 + rb_row;
          xr[i] = (grow < p.M) ? *reinterpret_cast<const float4*>(out + static_cast<size_t>(grow) * p.ldo + col)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      {
        float2 f[16];
        if (!pipe) {
          tmem_ld32(tmem_row_base + blk * 32, vacc);
          tmem_ld_wait();
        }
        acc_math32<HAS_BIAS, false, BN>(vacc, bias_smem + (blk >> 1) * 32 * 4, 0.f, 1.f, f);
        if (has_next) {
          if (pipe) tmem_ld32(tmem_row_base + (blk + 2) * 32, vacc);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 1) mbar_arrive(release_bar);
            else mbar_arrive_cluster(release_bar);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_shared_v4(my_row + ((c ^ sw) << 4), __float_as_uint(f[2 * c + 0].x), __float_as_uint(f[2 * c + 0].y),
                       __float_as_uint(f[2 * c + 1].x), __float_as_uint(f[2 * c + 1].y));
      }
      if constexpr (RPF) {
        if (has_next) cp_async_wait<1>(); else cp_async_wait<0>();   // this block's residual rows have landed
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + rb_row;
        float4 v = ld_shared_f4(stage_smem + r * 128 + ((rb_chunk ^ (r & 7)) << 4));
        const int grow = row_base + r;
        if constexpr (RPF) xr[i] = ld_shared_f4(rpf_cur + r * 128 + rb_chunk * 16);  // this lane's own copies
        if (grow < p.M) {
          if constexpr (EPI == EPI_BIAS_RESID_F32) {
            const float2 y01 = __fadd2_rn(make_float2(xr[i].x, xr[i].y), make_float2(v.x, v.y));
            const float2 y23 = __fadd2_rn(make_float2(xr[i].z, xr[i].w), make_float2(v.z, v.w));
            const float4 y = make_float4(y01.x, y01.y, y23.x, y23.y);
            *reinterpret_cast<float4*>(out + static_cast<size_t>(grow) * p.ldo + col) = y;
            if (emit) {
              // bf16 copy (A operand of the next, LayerNorm-folded GEMM) + statistics of the fp32 row
              uint2 u;
              u.x = pack_op2<F16>(y.x, y.y);
              u.y = pack_op2<F16>(y.z, y.w);
              *reinterpret_cast<uint2*>(p.xb_out + static_cast<size_t>(grow) * p.ldo + col) = u;
              const float2 s = __fadd2_rn(y01, y23);
              const float2 q = __ffma2_rn(y01, y01, __fmul2_rn(y23, y23));
              st1[i] += s.x + s.y;
              st2[i] += q.x + q.y;
            }
          } else if constexpr (EPI == EPI_SIM_F32) {
            const float rs = __ldg(p.rowscale + grow);
            const float4 cs = ld_shared_f4(bias_smem + ((blk >> 1) * 32 + rb_chunk * 4) * 4);
            *reinterpret_cast<float4*>(out + static_cast<size_t>(grow) * p.ldo + col) =
                make_float4(v.x * rs * cs.x, v.y * rs * cs.y, v.z * rs * cs.z, v.w * rs * cs.w);
          } else if constexpr (EPI == EPI_PATCH_F32) {
            const int b = grow / kPatches;
            const int pp = grow - b * kPatches;
            const float4 q = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(1 + pp) * p.N + col));
            *reinterpret_cast<float4*>(out + static_cast<size_t>(b * kVisSeq + 1 + pp) * p.ldo + col) =
                make_float4(v.x + q.x, v.y + q.y, v.z + q.z, v.w + q.w);
          } else {
            *reinterpret_cast<float4*>(out + static_cast<size_t>(grow) * p.ldo + col) = v;
          }
        }
      }
      if (has_next && pipe) tmem_ld_wait();
      __syncwarp();
    }
    if constexpr (EPI == EPI_BIAS_RESID_F32) {
      if (emit) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = st1[i], b = st2[i];
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {  // the 8 lanes that share a row
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          const int grow = row_base + i * 4 + rb_row;
          if (rb_chunk == 0 && grow < p.M)
            p.stats_out[static_cast<size_t>(grow) * kStatSlots + 2 * n_blk + half] = make_float2(a, b);
        }
      }
    }
  }
  return true;
}

// QUAD (experimental, PLIP_GEMM_QUAD=1): a cluster of TWO CTA pairs works on neighbouring M blocks of the same N block
// and takes the W tile with ONE multicast TMA load per half (pair 0 loads, both pairs receive) — 25 % fewer operand
// bytes through L2, the resource every layer GEMM is bound by (profiles/r2_notes.md §6).  Pair 0's stage slots are
// released by the MMA commits of BOTH pairs; everything else (accumulator barriers, TMEM, epilogue) stays per pair.
template <int CG, int BN, int EPI, bool F16, bool QUAD = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const GemmDev p) {
  static_assert(!QUAD || CG == 2, "the two-pair cluster is built from CTA pairs");
  constexpr int CL = QUAD ? 4 : CG;  // CTAs per cluster
  using C = Cfg<CG, BN, EPI>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = smem_raw_u32;
  if ((smem_raw_u32 & 1023u) != 0) {  // SWIZZLE_128B operand tiles need 1024-byte alignment
    if (threadIdx.x == 0) printf("plip_b200: dynamic shared memory base 0x%x is not 1024-byte aligned\n", smem_raw_u32);
    __trap();
  }
  const uint32_t epi_base = smem_base + STAGES * C::STAGE;   // 1024-aligned: 4 x 4 KB staging blocks
  const uint32_t bias_base = epi_base + kEpiWarps * kEpiStageBytes;  // per warp: BN / 2 bias (+ BN / 2 colsum) floats
  const uint32_t rpf_base = bias_base + kEpiWarps * C::VEC_BYTES;    // per warp: 2 x 4 KB residual prefetch buffers (C::RPF)
  const uint32_t bar_base = smem_base + STAGES * C::STAGE + C::EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;  // rank inside the cluster
  const uint32_t cta_rank = crank & (CG - 1);                  // rank inside the CTA pair
  const uint32_t pair = QUAD ? (crank >> 1) : 0u;              // which pair of the cluster
  const uint32_t lead_rank = crank & ~1u;                       // cluster rank of this pair's leader
  const bool leader = (cta_rank == 0);

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), CG);  // leader's arrive.expect_tx (+ peer's remote arrive)
      mbar_init(empty_bar(s), (QUAD && pair == 0) ? 2 : 1);  // tcgen05.commit (QUAD: of both pairs for the W multicaster)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);        // tcgen05.commit
      mbar_init(tempty_bar(a), kEpiWarps * CG);  // one arrive per epilogue warp (both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if (warp == kWarpTmem) tmem_alloc<CG>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base =
      *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_raw_u32));

  const int num_n_blk = p.N / BN;
  const int num_m_blk = (p.M + BM * CG - 1) / (BM * CG);
  // QUAD: a work item is a pair of neighbouring M blocks x one N block; an odd M-block count leaves the second pair of the
  // last item with rows past M only (zero-filled loads, nothing stored)
  const int num_tiles = (QUAD ? (num_m_blk + 1) / 2 : num_m_blk) * num_n_blk;
  const int num_kb = p.K / BK;
  const int tile0 = blockIdx.x / CL;
  const int tile_step = gridDim.x / CL;
  auto m_block_of = [&](int t) { const int mb = t / num_n_blk; return QUAD ? 2 * mb + (int)pair : mb; };

  if (warp == kWarpTma) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = tile0; t < num_tiles; t += tile_step) {
        const int m_blk = m_block_of(t), n_blk = t % num_n_blk;
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
            const uint32_t lfull = mapa_shared(full_bar(s), lead_rank);
            if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * C::STAGE);
            else mbar_arrive_cluster(lfull);
            tma_load_2d_cg2(sa, &tmA, lfull, kb * BK, m0);
            if constexpr (!QUAD) {
              tma_load_2d_cg2(sb, &tmB, lfull, kb * BK, n0);
            } else if (pair == 0) {
              // this CTA's half of the W tile, delivered to the same slot of CTA cta_rank of BOTH pairs
              // (barrier operand: this CTA's own offset with the pair's peer bit cleared = "the even CTA of the pair",
              // which the hardware resolves per destination CTA — the addressing CUTLASS' 2-SM multicast atoms use)
              tma_load_2d_cg2_mc(sb, &tmB, full_bar(s) & 0xFEFFFFFFu, kb * BK, n0,
                                 static_cast<uint16_t>((1u << cta_rank) | (1u << (cta_rank + 2))));
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_op(BM * CG, BN, 0, 0, F16);
      int s = 0, a = 0;
      uint32_t ph = 0, aph = 0;
      // commit targets: the stage barriers of this pair — plus, for the second pair of a QUAD cluster, those of the first
      // pair, whose producers refill the shared W slot — and the accumulator barrier of this pair
      const uint16_t empty_mask = QUAD ? (pair == 0 ? 0x3 : 0xF) : 0x3;
      const uint16_t tfull_mask = QUAD ? static_cast<uint16_t>(0x3u << (2 * pair)) : 0x3;
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
          if constexpr (CG == 2) umma_commit_mask(empty_bar(s), empty_mask); else umma_commit<CG>(empty_bar(s));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if constexpr (CG == 2) umma_commit_mask(tfull_bar(a), tfull_mask); else umma_commit<CG>(tfull_bar(a));
        a ^= 1;
        if (a == 0) aph ^= 1u;
      }
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue =====================
    constexpr bool LN_FOLD = (EPI == EPI_LN_BIAS_BF16 || EPI == EPI_LN_BIAS_GELU_BF16);
    constexpr bool HAS_BIAS = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RESID_F32 || LN_FOLD ||
                               EPI == EPI_SIM_F32);  // the smem "bias" tile carries the column scales for EPI_SIM_F32
    const int q = warp & 3;          // the TMEM lane quarter this warp may access (warp % 4)
    const int half = warp >> 2;      // which interleaved half of the tile's column blocks it handles
    int a = 0;
    uint32_t aph = 0;
    for (int t = tile0; t < num_tiles; t += tile_step) {
      const int m_blk = m_block_of(t), n_blk = t % num_n_blk;
      if constexpr (HAS_BIAS) {
        // this warp's slice of the bias (and colsum) vector: the (about BN / 2) columns it will touch, in its own order
        // (column blocks half, half + 2, ... of W columns).  Private to the warp: __syncwarp instead of a 256-thread
        // barrier per tile (ncu r2b: "barrier" was the second stall reason of every epilogue).
        constexpr bool kOut16 = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || LN_FOLD);
        constexpr int W = kOut16 ? 64 : 32;
        float* bs = reinterpret_cast<float*>(smem_raw + (bias_base - smem_raw_u32) + warp * C::VEC_BYTES);
        __syncwarp();
        for (int i = lane; i < C::VEC_FLOATS; i += 32) {
          const int cb = (half + 2 * (i / W)) * W + (i % W);       // column inside the tile
          if (cb < BN) {
            bs[i] = __ldg(p.bias + n_blk * BN + cb);
            if constexpr (LN_FOLD) bs[C::VEC_FLOATS + i] = __ldg(p.colsum + n_blk * BN + cb);
          }
        }
        __syncwarp();
      }
      const int row_base = m_blk * BM * CG + cta_rank * BM + q * 32;
      if constexpr (C::RPF)  // first residual block of the tile -> prefetch buffer 0, while the MMAs of the tile run
        rpf_issue(reinterpret_cast<const float*>(p.out), p.ldo, p.M, rpf_base + warp * 8192u, row_base, n_blk * BN + half * 32, lane);
      if constexpr (EPI == EPI_BIAS_RESID_F32) {
        // pull this warp's [32 rows x BN] slice of the residual stream into L2 while the MMAs run
        const float* xin = reinterpret_cast<const float*>(p.out);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int grow = row_base + j * 4 + (lane >> 3);
          if (grow < p.M) {
#pragma unroll
            for (int c = 0; c < BN / 256 + (BN % 256 != 0); ++c) {
              const int cofs = c * 256 + (lane & 7) * 32;  // one 128-byte line = one 32-column block
              const float* ptr = xin + static_cast<size_t>(grow) * p.ldo + n_blk * BN + cofs;
              if (cofs < BN && (((lane & 7) & 1) == half)) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
            }
          }
        }
      }
      float ln_mean = 0.f, ln_rstd = 1.f;
      if constexpr (LN_FOLD) {
        // LayerNorm statistics of this thread's row from the partials the producing GEMM left, summed in a
        // fixed order (bitwise reproducible); var = E[x^2] - mean^2 in fp32, eps = 1e-5 (TF:371,380).
        // All slots are fetched at once, before the accumulator wait, so the loads overlap the MMAs.
        const int grow = row_base + lane;
        float2 t[kStatSlots];
#pragma unroll
        for (int j = 0; j < kStatSlots; ++j)
          t[j] = (grow < p.M && j < p.n_partials) ? __ldg(p.stats_in + static_cast<size_t>(grow) * kStatSlots + j)
                                                  : make_float2(0.f, 0.f);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < kStatSlots; ++j) {
          s1 += t[j].x;
          s2 += t[j].y;
        }
        const float inv_k = 1.0f / static_cast<float>(p.K);
        ln_mean = s1 * inv_k;
        ln_rstd = rsqrtf(fmaxf(s2 * inv_k - ln_mean * ln_mean, 0.f) + kLnEps);
        ln_mean = -ln_rstd * ln_mean;  // the epilogue wants rstd * acc + (bias' + (-rstd * mean) * colsum)
      }
      mbar_wait(tfull_bar(a), aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN;
      uint32_t release_bar = tempty_bar(a);
      if constexpr (CG == 2) release_bar = mapa_shared(release_bar, lead_rank);
      const bool released =
          epilogue_tile<CG, BN, EPI, F16>(p, &tmC, trow, epi_base + warp * kEpiStageBytes, bias_base + warp * C::VEC_BYTES,
                                          rpf_base + warp * 8192u, row_base, n_blk * BN, n_blk, half, lane, ln_mean, ln_rstd,
                                          release_bar);
      if (!released) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) mbar_arrive(release_bar);
          else mbar_arrive_cluster(release_bar);
        }
      }
      a ^= 1;
      if (a == 0) aph ^= 1u;
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == kWarpTmem) tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
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

template <int CG, int BN, int EPI, bool F16, bool QUAD = false>
int launch_inst(const GemmArgs& g, cudaStream_t stream) {
  using C = Cfg<CG, BN, EPI>;
  constexpr int CL = QUAD ? 4 : CG;  // CTAs per cluster
  auto kern = gemm_kernel<CG, BN, EPI, F16, QUAD>;
  static unsigned long long configured = 0;
  static int max_groups = 0;  // co-resident clusters (CTAs for CG == 1, CTA pairs for CG == 2, two pairs for QUAD)
  if (first_use_on_device(configured)) {
    PLIP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM_BYTES));
    max_groups = num_sms() / CL;
    if (CL >= 2) {
      // A persistent grid must be fully co-resident: ask how many CTA pairs fit at once (SM pairs
      // must share a TPC, so this can be below num_sms / 2).
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(num_sms());
      q.blockDim = dim3(kThreads);
      q.dynamicSmemBytes = C::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int n_clusters = 0;
      PLIP_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n_clusters, kern, &q));
      if (n_clusters > 0 && n_clusters < max_groups) max_groups = n_clusters;
    }
    const int env_groups = env_int("PLIP_GEMM_GROUPS", 0);
    if (env_groups > 0) max_groups = env_groups;
    if (env_int("PLIP_DEBUG", 0))
      fprintf(stderr, "plip_b200: gemm<cg=%d,bn=%d,epi=%d,cluster=%d> stages=%d smem=%u max_groups=%d\n", CG, BN, EPI, CL,
              C::STAGES, C::SMEM_BYTES, max_groups);
  }
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_bf16_2d(&tmA, g.A, g.M, g.K, (uint64_t)g.lda * 2, BM, BK)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmB, g.W, g.N, g.K, (uint64_t)g.ldw * 2, C::LOAD_N, BK)) return rc;
  constexpr bool kOutBf16 = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_LN_BIAS_BF16 ||
                             EPI == EPI_LN_BIAS_GELU_BF16);
  static const int env_tma_store = env_int("PLIP_GEMM_TMA_STORE", 1);  // 0 = read-back + STG.128 path
  const bool tma_store = kOutBf16 && env_tma_store != 0;
  CUtensorMap tmC = tmA;  // placeholder when unused
  if (tma_store)
    if (int rc = make_tmap_bf16_2d(&tmC, g.out, g.M, g.N, (uint64_t)g.ldo * 2, 32, 64)) return rc;

  GemmDev p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  static const int env_dbg = env_int("PLIP_GEMM_DBG", 0);
  p.dbg = env_dbg;
  static const int env_f32_serial = env_int("PLIP_GEMM_F32_SERIAL", 0);
  p.f32_serial = env_f32_serial;
  p.tma_store = tma_store ? 1 : 0;
  p.bias = g.bias; p.rowscale = g.rowscale; p.out = g.out; p.ldo = g.ldo; p.pos = g.pos;
  p.colsum = g.colsum; p.stats_in = g.stats_in; p.n_partials = g.n_partials;
  p.xb_out = g.xb_out; p.stats_out = g.stats_out;
  if (g.n_tiles_used) *g.n_tiles_used = 2 * (g.N / BN);

  const int num_m_blk = (g.M + BM * CG - 1) / (BM * CG);
  const int num_tiles = (QUAD ? (num_m_blk + 1) / 2 : num_m_blk) * (g.N / BN);
  int groups = max_groups;
  if (groups > num_tiles) groups = num_tiles;

  PLIP_CUDA_CHECK(launch_kernel(kern, dim3(groups * CL), dim3(kThreads), C::SMEM_BYTES, stream, CL, tmA, tmB, tmC, p));
  ++g_launch_count;
  return 0;
}

template <int CG, int BN, bool F16>
int launch_epi_fmt(const GemmArgs& g, cudaStream_t stream) {
  if constexpr (CG == 2) {
    // experimental two-pair clusters with a multicast W tile: the three layer epilogues only, opt-in
    static const int env_quad = env_int("PLIP_GEMM_QUAD", 0);
    if (env_quad && !g.force_cg) {
      switch (g.epi) {
        case EPI_BIAS_RESID_F32: return launch_inst<2, BN, EPI_BIAS_RESID_F32, F16, true>(g, stream);
        case EPI_LN_BIAS_BF16: return launch_inst<2, BN, EPI_LN_BIAS_BF16, F16, true>(g, stream);
        case EPI_LN_BIAS_GELU_BF16: return launch_inst<2, BN, EPI_LN_BIAS_GELU_BF16, F16, true>(g, stream);
        default: break;
      }
    }
  }
  switch (g.epi) {
    case EPI_BIAS_BF16: return launch_inst<CG, BN, EPI_BIAS_BF16, F16>(g, stream);
    case EPI_BIAS_GELU_BF16: return launch_inst<CG, BN, EPI_BIAS_GELU_BF16, F16>(g, stream);
    case EPI_BIAS_RESID_F32: return launch_inst<CG, BN, EPI_BIAS_RESID_F32, F16>(g, stream);
    case EPI_PATCH_F32: return launch_inst<CG, BN, EPI_PATCH_F32, F16>(g, stream);
    case EPI_F32: return launch_inst<CG, BN, EPI_F32, F16>(g, stream);
    case EPI_LN_BIAS_BF16: return launch_inst<CG, BN, EPI_LN_BIAS_BF16, F16>(g, stream);
    case EPI_LN_BIAS_GELU_BF16: return launch_inst<CG, BN, EPI_LN_BIAS_GELU_BF16, F16>(g, stream);
    case EPI_NULL: return launch_inst<CG, BN, EPI_NULL, F16>(g, stream);
    case EPI_SIM_F32: return launch_inst<CG, BN, EPI_SIM_F32, F16>(g, stream);
    default: set_last_error("launch_gemm: bad epilogue %d", g.epi); return -2;
  }
}
template <int CG, int BN>
int launch_epi(const GemmArgs& g, cudaStream_t stream) {
  return g.f16 ? launch_epi_fmt<CG, BN, true>(g, stream) : launch_epi_fmt<CG, BN, false>(g, stream);
}

}  // namespace

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  PLIP_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "launch_gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  PLIP_REQUIRE(g.K % BK == 0, "launch_gemm: K=%d must be a multiple of %d", g.K, BK);
  PLIP_REQUIRE(g.N % 128 == 0 || g.N % 192 == 0, "launch_gemm: N=%d must be a multiple of 128 or 192", g.N);
  PLIP_REQUIRE((g.lda % 8) == 0 && (g.ldw % 8) == 0 && (g.ldo % 8) == 0,
               "launch_gemm: leading dimensions must be multiples of 8 elements");
  PLIP_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.W) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(g.out) & 15) == 0,
               "launch_gemm: operands must be 16-byte aligned");
  if (g.epi == EPI_LN_BIAS_BF16 || g.epi == EPI_LN_BIAS_GELU_BF16)
    PLIP_REQUIRE(g.colsum && g.stats_in && g.bias && g.n_partials >= 1 && g.n_partials <= kStatSlots,
                 "launch_gemm: LayerNorm-folded epilogue needs colsum, stats and 1..%d partials", kStatSlots);
  if (g.epi == EPI_SIM_F32)
    PLIP_REQUIRE(g.bias && g.rowscale, "launch_gemm: the similarity epilogue needs row and column scales");
  if (g.xb_out || g.stats_out)
    PLIP_REQUIRE(g.epi == EPI_BIAS_RESID_F32 && g.xb_out && g.stats_out,
                 "launch_gemm: xb/stats outputs belong to the residual epilogue");
  static const int env_cg = env_int("PLIP_GEMM_CG", 0);
  static const int env_bn = env_int("PLIP_GEMM_BN", 0);
  int cg = g.force_cg ? g.force_cg : (env_cg ? env_cg : 2);
  int bn = g.force_bn ? g.force_bn : (env_bn ? env_bn : 256);
  // Memory-bound residual GEMM with a short K (out_proj): 192-wide tiles quantise better on 74 CTA
  // pairs (800 tiles = 10.8 waves instead of 600 = 8.1 -> 9) and were measured 7 % faster (exp8).
  if (!g.force_bn && !env_bn && cg == 2 && g.epi == EPI_BIAS_RESID_F32 && g.K <= 1024) {
    if (g.N % 192 == 0) bn = 192;
    else if (g.N == 512) bn = 128;  // text out_proj: 86.5 vs 91.8 us (exp14)
  }
  // (192-wide tiles for the long-K residual GEMM, fc2, were measured slower in round 2: 252.9 vs 238.0 us in step)
  if (g.N % bn != 0) bn = 128;
  PLIP_REQUIRE((cg == 1 || cg == 2) && (bn == 128 || bn == 256 || (bn == 192 && cg == 2)),
               "launch_gemm: bad config cg=%d bn=%d", cg, bn);
  PLIP_REQUIRE(!g.stats_out || 2 * (g.N / bn) <= kStatSlots, "launch_gemm: N=%d / BN=%d exceeds %d statistics slots",
               g.N, bn, kStatSlots);
  if (cg == 1) return bn == 256 ? launch_epi<1, 256>(g, stream) : launch_epi<1, 128>(g, stream);
  if (bn == 192) return launch_epi<2, 192>(g, stream);
  return bn == 256 ? launch_epi<2, 256>(g, stream) : launch_epi<2, 128>(g, stream);
}

}  // namespace plip
