// plip_b200 — image resize + crop on the device: variable-size RGB uint8 images -> 224x224 uint8 tiles.
//
// Replaces the PIL pass the reference runs on the host before every forward: CLIPProcessor's shortest-edge
// bicubic resize + centre crop (/root/reference/plip.py:35; TF:models/clip/image_processing_clip.py:50-62) and
// torchvision's Resize(n_px, BICUBIC) + CenterCrop (reproducibility/embedders/transform.py:45-52), both of which
// end in Pillow's ImagingResample: separable antialiased bicubic (a = -0.5), window half-width 2*max(scale,1),
// double-precision weights normalised per output pixel and quantised to 22-bit fixed point, int32 accumulation
// from 1<<21, >>22 and clamp, uint8 image between the horizontal and the vertical pass.  The result here is
// bit-identical to PIL.Image.resize(...).crop(...) (tests/test_resize.py): the weights are rebuilt on the device
// with the same sequence of IEEE double operations (explicit _rn intrinsics: no FMA contraction).
//
// One CTA produces 8 output rows of one tile.  It builds the 224 horizontal filter rows it needs (only the
// cropped columns) and its 8 vertical ones in shared memory, then for `rows_per_pass` output rows at a time
// runs the horizontal pass over the source rows those outputs touch into a uint8 shared-memory strip and the
// vertical pass out of that strip.  Source pixels are read from HBM once per CTA strip (strips of adjacent CTAs
// overlap by the filter support); the roofline is HBM: source bytes + 150,528 tile bytes per image.
#include "kernels.cuh"

#include <math.h>

namespace plip {

namespace {

constexpr int kRsThreads = 256;
constexpr int kRsRowsPerCta = 8;
constexpr int kRsBatch = 64;                 // images per launch (descriptors travel as kernel parameters)
constexpr int kTileRowBytes = kImage * 3;    // 672
constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow's PRECISION_BITS for 8-bit channels

struct ResizeImg {
  long long src_off;   // byte offset of the image in the packed source buffer
  int w, h;            // source size
  int new_w, new_h;    // size after the resize
  int left, top;       // crop origin in the resized image
  int rows_per_pass;   // output rows per strip (8, 4, 2 or 1)
  int strip_rows;      // capacity of the uint8 strip, in source rows
};

struct ResizeBatch {
  ResizeImg img[kRsBatch];
};

struct AxisFilter {
  double scale, support, ss;
  int ksize;
};

// IEEE double operations without FMA contraction: intrinsics on the device, plain operators on the host (x86-64
// gcc does not contract without -mfma).  The host instantiation only serves plip_dbg_resize_filter (CPU tests).
#ifdef __CUDA_ARCH__
#define RN_ADD(a, b) __dadd_rn((a), (b))
#define RN_SUB(a, b) __dsub_rn((a), (b))
#define RN_MUL(a, b) __dmul_rn((a), (b))
#define RN_DIV(a, b) __ddiv_rn((a), (b))
#else
#define RN_ADD(a, b) ((a) + (b))
#define RN_SUB(a, b) ((a) - (b))
#define RN_MUL(a, b) ((a) * (b))
#define RN_DIV(a, b) ((a) / (b))
#endif

__host__ __device__ inline int axis_ksize(int in_size, int out_size) {
  double fs = (double)in_size / (double)out_size;
  if (fs < 1.0) fs = 1.0;
  return (int)ceil(2.0 * fs) * 2 + 1;
}

__host__ __device__ __forceinline__ AxisFilter make_axis(int in_size, int out_size) {
  AxisFilter f;
  f.scale = RN_DIV((double)in_size, (double)out_size);
  const double fs = f.scale < 1.0 ? 1.0 : f.scale;
  f.support = RN_MUL(2.0, fs);
  f.ksize = (int)ceil(f.support) * 2 + 1;
  f.ss = RN_DIV(1.0, fs);
  return f;
}

__host__ __device__ __forceinline__ double bicubic_rn(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return RN_ADD(RN_MUL(RN_MUL(RN_SUB(RN_MUL(a + 2.0, x), a + 3.0), x), x), 1.0);
  if (x < 2.0) return RN_MUL(RN_SUB(RN_MUL(RN_ADD(RN_MUL(RN_SUB(x, 5.0), x), 8.0), x), 4.0), a);
  return 0.0;
}

// Filter row of output index `xx`: window [xmin, xmin+count) and fixed-point weights k[0..count).
__host__ __device__ __forceinline__ void filter_row(const AxisFilter& f, int in_size, int xx, int* k,
                                                    int& xmin_out, int& count_out) {
  const double center = RN_MUL((double)xx + 0.5, f.scale);
  int xmin = (int)RN_ADD(RN_SUB(center, f.support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)RN_ADD(RN_ADD(center, f.support), 0.5);
  if (xmax > in_size) xmax = in_size;
  const int count = xmax - xmin;
  double ww = 0.0;
  for (int x = 0; x < count; ++x)
    ww = RN_ADD(ww, bicubic_rn(RN_MUL(RN_ADD(RN_SUB((double)(x + xmin), center), 0.5), f.ss)));
  const double one = (double)(1 << kPrecisionBits);
  for (int x = 0; x < count; ++x) {
    double w = bicubic_rn(RN_MUL(RN_ADD(RN_SUB((double)(x + xmin), center), 0.5), f.ss));
    if (ww != 0.0) w = RN_DIV(w, ww);
    k[x] = w < 0.0 ? (int)RN_ADD(-0.5, RN_MUL(w, one)) : (int)RN_ADD(0.5, RN_MUL(w, one));
  }
  xmin_out = xmin;
  count_out = count;
}

__device__ __forceinline__ uint8_t clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__global__ void __launch_bounds__(kRsThreads) resize_crop_kernel(const uint8_t* __restrict__ src,
                                                                 uint8_t* __restrict__ tiles, const ResizeBatch batch,
                                                                 int64_t first_image) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  pdl_wait();
  pdl_launch_dependents();
  const ResizeImg& im = batch.img[blockIdx.y];
  const int row0 = blockIdx.x * kRsRowsPerCta;  // first output row of this CTA
  const AxisFilter fh = make_axis(im.w, im.new_w), fv = make_axis(im.h, im.new_h);

  int* kh = reinterpret_cast<int*>(rs_smem);                 // [224][ksize_h]
  int* kv = kh + kImage * fh.ksize;                          // [8][ksize_v]
  int* bh = kv + kRsRowsPerCta * fv.ksize;                   // [224][2] (xmin, count)
  int* bv = bh + kImage * 2;                                 // [8][2]
  uint8_t* strip = reinterpret_cast<uint8_t*>(bv + kRsRowsPerCta * 2);  // [strip_rows][224*3]

  const int t = threadIdx.x;
  if (t < kImage) {
    int xmin, cnt;
    filter_row(fh, im.w, im.left + t, kh + t * fh.ksize, xmin, cnt);
    bh[2 * t] = xmin;
    bh[2 * t + 1] = cnt;
  } else if (t < kImage + kRsRowsPerCta) {
    const int j = t - kImage;
    int ymin, cnt;
    filter_row(fv, im.h, im.top + row0 + j, kv + j * fv.ksize, ymin, cnt);
    bv[2 * j] = ymin;
    bv[2 * j + 1] = cnt;
  }
  __syncthreads();

  const uint8_t* img = src + im.src_off;
  const int64_t src_row_bytes = (int64_t)im.w * 3;
  uint8_t* out = tiles + ((first_image + blockIdx.y) * kImage + row0) * (int64_t)kTileRowBytes;
  const int rp = im.rows_per_pass;
  for (int sub = 0; sub < kRsRowsPerCta; sub += rp) {
    const int s0 = bv[2 * sub];
    const int s1 = bv[2 * (sub + rp - 1)] + bv[2 * (sub + rp - 1) + 1];
    const int nrows = s1 - s0;
    if (nrows > im.strip_rows) __trap();  // host sizing bug: never expected
    // horizontal pass: strip[r][xx][c] for the source rows [s0, s1)
    for (int idx = t; idx < nrows * kTileRowBytes; idx += kRsThreads) {
      const int r = idx / kTileRowBytes, e = idx - r * kTileRowBytes;
      const int xx = e / 3, c = e - xx * 3;
      const int xmin = bh[2 * xx], cnt = bh[2 * xx + 1];
      const uint8_t* p = img + (int64_t)(s0 + r) * src_row_bytes + (int64_t)xmin * 3 + c;
      const int* k = kh + xx * fh.ksize;
      int acc = 1 << (kPrecisionBits - 1);
      for (int x = 0; x < cnt; ++x) acc += (int)__ldg(p + x * 3) * k[x];
      strip[idx] = clip8(acc);
    }
    __syncthreads();
    // vertical pass: rp output rows out of the strip
    for (int idx = t; idx < rp * kTileRowBytes; idx += kRsThreads) {
      const int j = idx / kTileRowBytes, e = idx - j * kTileRowBytes;
      const int ymin = bv[2 * (sub + j)], cnt = bv[2 * (sub + j) + 1];
      const uint8_t* p = strip + (ymin - s0) * kTileRowBytes + e;
      const int* k = kv + (sub + j) * fv.ksize;
      int acc = 1 << (kPrecisionBits - 1);
      for (int y = 0; y < cnt; ++y) acc += (int)p[y * kTileRowBytes] * k[y];
      out[(sub + j) * kTileRowBytes + e] = clip8(acc);
    }
    __syncthreads();
  }
}

constexpr size_t kRsSmemSoft = 96 * 1024;    // preferred ceiling: two CTAs per SM
constexpr size_t kRsSmemHard = 200 * 1024;

size_t table_bytes(int ksh, int ksv) {
  return ((size_t)kImage * ksh + (size_t)kRsRowsPerCta * ksv + (kImage + kRsRowsPerCta) * 2) * sizeof(int);
}

// Source rows one strip of `rp` output rows can touch: (rp-1)*scale + 2*support + rounding slack.
int strip_rows_for(int in_h, int out_h, int rp) {
  const double scale = (double)in_h / (double)out_h;
  const double support = 2.0 * (scale < 1.0 ? 1.0 : scale);
  int rows = (int)ceil((rp - 1) * scale + 2.0 * support) + 3;
  return rows > in_h ? in_h : rows;
}

}  // namespace

int resize_filter_host(int in_size, int out_size, int xx, int32_t* k, int k_cap, int* xmin, int* count) {
  const AxisFilter f = make_axis(in_size, out_size);
  if (k_cap < f.ksize) return -f.ksize;
  filter_row(f, in_size, xx, k, *xmin, *count);
  return f.ksize;
}

int launch_resize_crop(const uint8_t* src, size_t src_bytes, const plip_resize_desc_t* d, int64_t n, uint8_t* tiles,
                       cudaStream_t st) {
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    PLIP_CUDA_CHECK(cudaFuncSetAttribute(resize_crop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kRsSmemHard));
  for (int64_t base = 0; base < n; base += kRsBatch) {
    const int cnt = (int)((n - base) < kRsBatch ? (n - base) : kRsBatch);
    ResizeBatch b = {};
    size_t smem = 0;
    for (int i = 0; i < cnt; ++i) {
      const plip_resize_desc_t& s = d[base + i];
      const long long idx = (long long)(base + i);
      PLIP_REQUIRE(s.width > 0 && s.height > 0 && s.width <= 65536 && s.height <= 65536,
                   "plip_resize_crop_u8: image %lld has invalid size %dx%d", idx, s.width, s.height);
      PLIP_REQUIRE(s.offset >= 0 && (uint64_t)s.offset + (uint64_t)s.width * s.height * 3 <= src_bytes,
                   "plip_resize_crop_u8: image %lld (%dx%d at byte %lld) exceeds the %llu-byte source buffer", idx,
                   s.width, s.height, (long long)s.offset, (unsigned long long)src_bytes);
      PLIP_REQUIRE(s.new_width >= kImage && s.new_height >= kImage && s.new_width <= 65536 && s.new_height <= 65536,
                   "plip_resize_crop_u8: image %lld: resized size %dx%d is smaller than the %dx%d tile", idx,
                   s.new_width, s.new_height, kImage, kImage);
      PLIP_REQUIRE(s.left >= 0 && s.top >= 0 && s.left + kImage <= s.new_width && s.top + kImage <= s.new_height,
                   "plip_resize_crop_u8: image %lld: crop origin (%d,%d) leaves the %dx%d resized image", idx, s.left,
                   s.top, s.new_width, s.new_height);
      const int ksh = axis_ksize(s.width, s.new_width), ksv = axis_ksize(s.height, s.new_height);
      const size_t tb = table_bytes(ksh, ksv);
      int rp = 0, rows = 0;
      for (size_t limit : {kRsSmemSoft, kRsSmemHard}) {
        for (int cand = kRsRowsPerCta; cand >= 1 && !rp; cand >>= 1) {
          const int r = strip_rows_for(s.height, s.new_height, cand);
          if (tb + (size_t)r * kTileRowBytes <= limit) rp = cand, rows = r;
        }
        if (rp) break;
      }
      PLIP_REQUIRE(rp, "plip_resize_crop_u8: image %lld (%dx%d -> %dx%d) shrinks too much for the on-device "
                   "resize (filter tables need %zu bytes of shared memory); reduce it on the host first",
                   idx, s.width, s.height, s.new_width, s.new_height, tb);
      ResizeImg& o = b.img[i];
      o.src_off = s.offset;
      o.w = s.width, o.h = s.height, o.new_w = s.new_width, o.new_h = s.new_height;
      o.left = s.left, o.top = s.top, o.rows_per_pass = rp, o.strip_rows = rows;
      const size_t need = tb + (size_t)rows * kTileRowBytes;
      smem = need > smem ? need : smem;
    }
    PLIP_CUDA_CHECK(launch_pdl(resize_crop_kernel, dim3(kImage / kRsRowsPerCta, cnt), dim3(kRsThreads), smem, st, 1,
                               src, tiles, b, base));
    ++g_launch_count;
  }
  return 0;
}

}  // namespace plip
