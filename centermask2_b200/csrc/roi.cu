// ROI-stage and output-side kernels: FPN level assignment fused into ROIAlign, SAG spatial attention,
// class-gathered mask predictor + sigmoid, MaskIoU input/score glue, box rescale and mask paste-back.
// All bandwidth-bound: coalesced over channels (NHWC) / over x (paste), 8 channels per thread.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>
#include <type_traits>

namespace cm2 {

// ---------------------------------------------------------------------------------------------
// ROIAlign over an FPN pyramid with the level chosen per box.
//   centermask/pooler.py:80-118 (ratio) / :121-152 (area) + torchvision roi_align(aligned=True).
// One thread = (roi slot, output bin, 8-channel vector); neighbouring threads walk channels, so every
// bilinear tap is a coalesced row segment of the NHWC feature map.
// ---------------------------------------------------------------------------------------------
constexpr int ROI_MAX_LEVELS = 4;

template <typename T>
struct RoiAlignParams {
  View<const T> feat[ROI_MAX_LEVELS];
  float scale[ROI_MAX_LEVELS];
  int num_levels, min_level, max_level;
  const float* boxes;
  const int* det_count;
  const float* image_area;
  int n, r_cap, crit, sampling_ratio, res;
  View<T> out;
  int* level_out;
  int* order;                     // column kernel: ROI slots in launch order (largest first), [n * r_cap]
};

__device__ __forceinline__ int assign_level(float x0, float y0, float x1, float y1, float img_area, int crit,
                                            int min_level, int max_level) {
  const float eps = 2.220446049250313e-16f;
  float area = (x1 - x0) * (y1 - y0);
  float lv;
  if (crit == 0) {
    lv = ceilf((float)max_level - log2f(img_area / area + eps));              // pooler.py:110-112
  } else {
    lv = floorf(4.0f + log2f(sqrtf(area) / 224.0f + eps));                     // pooler.py:145-147
  }
  if (!(lv >= (float)min_level)) lv = (float)min_level;                       // also catches NaN
  if (lv > (float)max_level) lv = (float)max_level;
  return (int)lv - min_level;
}

template <typename T>
__global__ void roialign_fpn_kernel(const RoiAlignParams<T> p) {
  const int c8 = p.out.c >> 3;
  const int bins = p.res * p.res;
  const int64_t total = (int64_t)p.n * p.r_cap * bins * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % c8);
    int64_t t = i / c8;
    int bin = (int)(t % bins);
    int slot = (int)(t / bins);
    int img = slot / p.r_cap;
    int r = slot - img * p.r_cap;
    int ph = bin / p.res, pw = bin - ph * p.res;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < p.det_count[img]) {
      const float4 bx = reinterpret_cast<const float4*>(p.boxes)[slot];
      int lvl = assign_level(bx.x, bx.y, bx.z, bx.w, p.image_area[img], p.crit, p.min_level, p.max_level);
      if (lvl >= p.num_levels) lvl = p.num_levels - 1;
      if (p.level_out && bin == 0 && cv == 0) p.level_out[slot] = lvl;
      // pick the level's view without dynamic indexing of the param array (keeps it in constant bank)
      View<const T> f = p.feat[0];
      float scale = p.scale[0];
#pragma unroll
      for (int l = 1; l < ROI_MAX_LEVELS; ++l)
        if (l == lvl) { f = p.feat[l]; scale = p.scale[l]; }
      const int height = f.h, width = f.w;
      float roi_start_w = bx.x * scale - 0.5f, roi_start_h = bx.y * scale - 0.5f;
      float roi_end_w = bx.z * scale - 0.5f, roi_end_h = bx.w * scale - 0.5f;
      float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
      float bin_h = roi_height / (float)p.res, bin_w = roi_width / (float)p.res;
      int grid_h = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_height / (float)p.res);
      int grid_w = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_width / (float)p.res);
      float count = fmaxf((float)(grid_h * grid_w), 1.f);
      for (int iy = 0; iy < grid_h; ++iy) {
        float y = roi_start_h + ph * bin_h + ((float)iy + 0.5f) * bin_h / (float)grid_h;
        for (int ix = 0; ix < grid_w; ++ix) {
          float x = roi_start_w + pw * bin_w + ((float)ix + 0.5f) * bin_w / (float)grid_w;
          if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) continue;
          float yy = y <= 0.f ? 0.f : y, xx = x <= 0.f ? 0.f : x;
          int y_low = (int)yy, x_low = (int)xx, y_high, x_high;
          if (y_low >= height - 1) { y_high = y_low = height - 1; yy = (float)y_low; } else y_high = y_low + 1;
          if (x_low >= width - 1) { x_high = x_low = width - 1; xx = (float)x_low; } else x_high = x_low + 1;
          float ly = yy - (float)y_low, lx = xx - (float)x_low, hy = 1.f - ly, hx = 1.f - lx;
          float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
          float v1[8], v2[8], v3[8], v4[8];
          Vec8<T>::load(f.at(img, y_low, x_low) + cv * 8, v1);
          Vec8<T>::load(f.at(img, y_low, x_high) + cv * 8, v2);
          Vec8<T>::load(f.at(img, y_high, x_low) + cv * 8, v3);
          Vec8<T>::load(f.at(img, y_high, x_high) + cv * 8, v4);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += w1 * v1[k] + w2 * v2[k] + w3 * v3[k] + w4 * v4[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] /= count;
    }
    Vec8<T>::store(p.out.at(slot, ph, pw) + cv * 8, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// ROIAlign, one CTA per ROI slot, MERGED TAPS.  The average over the adaptive sampling grid is separable:
//   out[ph][pw] = 1/count * sum_Y sum_X Wy[ph][Y] * Wx[pw][X] * f[Y][X],
// where Wy[ph][Y] collects the bilinear row weights of the bin's grid_h samples (hy on y_low, ly on y_high; samples
// outside [-1, H] dropped, coordinates clamped -- torchvision's rules) and Wx likewise.  A bin therefore reads
// (grid_h + 1) x (grid_w + 1) feature vectors instead of 4 x grid_h x grid_w, every one of them a contiguous 16 bytes
// per lane / 512 bytes per warp (NHWC).  The two weight tables (res x ROI_MAXT each) are built once per CTA in shared
// memory; a box whose grid does not fit the tables falls back to the sample loop.  Sums are reordered with respect to
// torchvision (fp32 rounding only).
// ---------------------------------------------------------------------------------------------
// 8 channels as loaded (the conversion happens when the value is used, so several loads can be in flight cheaply)
__device__ __forceinline__ void ffma2_roi(float& d0, float& d1, float a0, float a1, float b) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %4};\nmov.b64 rc, {%0, %1};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b));
}
template <typename T> struct Raw8;
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void fma(float w, float (&acc)[8]) const {
    ffma2_roi(acc[0], acc[1], a.x, a.y, w); ffma2_roi(acc[2], acc[3], a.z, a.w, w);
    ffma2_roi(acc[4], acc[5], b.x, b.y, w); ffma2_roi(acc[6], acc[7], b.z, b.w, w);
  }
};
template <> struct Raw8<__nv_bfloat16> {
  uint4 q;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { q = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void fma(float w, float (&acc)[8]) const {
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      ffma2_roi(acc[2 * k], acc[2 * k + 1], __uint_as_float(u[k] << 16), __uint_as_float(u[k] & 0xffff0000u), w);
  }
};

constexpr int ROI_MAXT = 32;
constexpr int ROI_MAX_RES = 32;

template <typename T>
__device__ __forceinline__ void roialign_sample_loop(const View<const T>& f, int img, int cv, int ph, int pw, float roi_start_h,
                                                     float roi_start_w, float bin_h, float bin_w, int grid_h, int grid_w,
                                                     float (&acc)[8]) {
  const int height = f.h, width = f.w;
  for (int iy = 0; iy < grid_h; ++iy) {
    float y = roi_start_h + ph * bin_h + ((float)iy + 0.5f) * bin_h / (float)grid_h;
    for (int ix = 0; ix < grid_w; ++ix) {
      float x = roi_start_w + pw * bin_w + ((float)ix + 0.5f) * bin_w / (float)grid_w;
      if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) continue;
      float yy = y <= 0.f ? 0.f : y, xx = x <= 0.f ? 0.f : x;
      int y_low = (int)yy, x_low = (int)xx, y_high, x_high;
      if (y_low >= height - 1) { y_high = y_low = height - 1; yy = (float)y_low; } else y_high = y_low + 1;
      if (x_low >= width - 1) { x_high = x_low = width - 1; xx = (float)x_low; } else x_high = x_low + 1;
      float ly = yy - (float)y_low, lx = xx - (float)x_low, hy = 1.f - ly, hx = 1.f - lx;
      float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
      // rare path (boxes wider than 14 * 30 feature pixels): one tap at a time keeps the register footprint of the
      // kernel that of the merged-tap loop
      const float wt[4] = {w1, w2, w3, w4};
      const int ty[4] = {y_low, y_low, y_high, y_high}, tx[4] = {x_low, x_high, x_low, x_high};
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        Raw8<T> rv;
        rv.load(f.at(img, ty[t], tx[t]) + cv * 8);
        rv.fma(wt[t], acc);
      }
    }
  }
}

// weight tables of one ROI (all threads of the CTA call this; ends with a barrier)
__device__ __forceinline__ void roi_build_tables(float (*s_w)[ROI_MAX_RES][ROI_MAXT], int (*s_start)[ROI_MAX_RES],
                                                 int (*s_num)[ROI_MAX_RES], int res, float roi_start_h, float roi_start_w,
                                                 float bin_h, float bin_w, int grid_h, int grid_w, int fh, int fw) {
  for (int i = threadIdx.x; i < 2 * res * ROI_MAXT; i += blockDim.x) {
    const int axis = i / (res * ROI_MAXT), rem = i - axis * res * ROI_MAXT;
    s_w[axis][rem / ROI_MAXT][rem % ROI_MAXT] = 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * res; t += blockDim.x) {
    const int axis = t / res, pb = t - axis * res;
    const float start = axis ? roi_start_w : roi_start_h, bin = axis ? bin_w : bin_h;
    const int grid = axis ? grid_w : grid_h, extent = axis ? fw : fh;
    int s0 = -1, num = 0;
    for (int i = 0; i < grid; ++i) {
      const float v = start + pb * bin + ((float)i + 0.5f) * bin / (float)grid;
      if (v < -1.0f || v > (float)extent) continue;
      float vv = v <= 0.f ? 0.f : v;
      int low = (int)vv, high;
      if (low >= extent - 1) { high = low = extent - 1; vv = (float)low; } else high = low + 1;
      const float l = vv - (float)low, h = 1.f - l;
      if (s0 < 0) s0 = low;
      s_w[axis][pb][low - s0] += h;
      s_w[axis][pb][high - s0] += l;
      num = high - s0 + 1;
    }
    s_start[axis][pb] = s0 < 0 ? 0 : s0;
    s_num[axis][pb] = num;
  }
  __syncthreads();
}

// merged-tap evaluation of one bin for 8 channels (weights and tap ranges from the tables)
template <typename T>
__device__ __forceinline__ void roi_bin_merged(const T* __restrict__ img_base, int sh32, int sw32,
                                               const float (*s_w)[ROI_MAX_RES][ROI_MAXT], const int (*s_start)[ROI_MAX_RES],
                                               const int (*s_num)[ROI_MAX_RES], int ph, int pw, int cv, float (&acc)[8]) {
  const int ny = s_num[0][ph], nx = s_num[1][pw];
  const float* wy = s_w[0][ph];
  const float* wx = s_w[1][pw];
  int off0 = s_start[0][ph] * sh32 + s_start[1][pw] * sw32 + cv * 8;
  for (int jy = 0; jy < ny; ++jy, off0 += sh32) {
    const float wyv = wy[jy];
    int off = off0, jx = 0;
    for (; jx + 4 <= nx; jx += 4, off += 4 * sw32) {
      Raw8<T> r0, r1, r2, r3;
      r0.load(img_base + off); r1.load(img_base + off + sw32);
      r2.load(img_base + off + 2 * sw32); r3.load(img_base + off + 3 * sw32);
      r0.fma(wyv * wx[jx], acc); r1.fma(wyv * wx[jx + 1], acc);
      r2.fma(wyv * wx[jx + 2], acc); r3.fma(wyv * wx[jx + 3], acc);
    }
    if (nx - jx >= 2) {
      Raw8<T> r0, r1;
      r0.load(img_base + off); r1.load(img_base + off + sw32);
      r0.fma(wyv * wx[jx], acc); r1.fma(wyv * wx[jx + 1], acc);
      jx += 2; off += 2 * sw32;
    }
    if (nx - jx >= 1) {
      Raw8<T> r0;
      r0.load(img_base + off);
      r0.fma(wyv * wx[jx], acc);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 3) roialign_roi_kernel(const RoiAlignParams<T> p) {
  __shared__ float s_w[2][ROI_MAX_RES][ROI_MAXT];       // [axis][bin][tap] merged weights (axis 0 = y)
  __shared__ int s_start[2][ROI_MAX_RES], s_num[2][ROI_MAX_RES];
  const int slot = blockIdx.x;
  const int img = slot / p.r_cap, r = slot - img * p.r_cap;
  const int res = p.res, c8 = p.out.c >> 3, bins = res * res;
  const int items = bins * c8;
  if (r >= p.det_count[img]) {                          // empty slot: zeros
    const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
      const int bin = i / c8, cv = i - bin * c8;
      const int ph = bin / res, pw = bin - ph * res;
      Vec8<T>::store(p.out.at(slot, ph, pw) + cv * 8, z);
    }
    return;
  }
  const float4 bx = reinterpret_cast<const float4*>(p.boxes)[slot];
  int lvl = assign_level(bx.x, bx.y, bx.z, bx.w, p.image_area[img], p.crit, p.min_level, p.max_level);
  if (lvl >= p.num_levels) lvl = p.num_levels - 1;
  if (p.level_out && threadIdx.x == 0) p.level_out[slot] = lvl;
  View<const T> f = p.feat[0];
  float scale = p.scale[0];
#pragma unroll
  for (int l = 1; l < ROI_MAX_LEVELS; ++l)
    if (l == lvl) { f = p.feat[l]; scale = p.scale[l]; }
  const float roi_start_w = bx.x * scale - 0.5f, roi_start_h = bx.y * scale - 0.5f;
  const float roi_end_w = bx.z * scale - 0.5f, roi_end_h = bx.w * scale - 0.5f;
  const float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
  const float bin_h = roi_height / (float)res, bin_w = roi_width / (float)res;
  const int grid_h = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_height / (float)res);
  const int grid_w = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_width / (float)res);
  const float count = fmaxf((float)(grid_h * grid_w), 1.f);
  const bool merged = grid_h + 2 <= ROI_MAXT && grid_w + 2 <= ROI_MAXT;      // CTA-uniform
  if (merged)
    roi_build_tables(s_w, s_start, s_num, res, roi_start_h, roi_start_w, bin_h, bin_w, grid_h, grid_w, f.h, f.w);
  // exact division of small indices by the run-time constants c8 / res: q = umulhi(i, ceil(2^32 / d)) for i * d < 2^32
  const uint32_t magic_c8 = (uint32_t)((0x100000000ull + c8 - 1) / (uint32_t)c8);
  const uint32_t magic_res = (uint32_t)((0x100000000ull + res - 1) / (uint32_t)res);
  const float inv_count = 1.0f / count;
  // 32-bit element offsets inside the image (the host checks sn < 2^31): one IMAD.WIDE per load address
  const T* img_base = f.p + (size_t)img * f.sn;
  const int sh32 = (int)f.sh, sw32 = (int)f.sw;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int bin = c8 == 32 ? (i >> 5) : (c8 == 1 ? i : (int)__umulhi((uint32_t)i, magic_c8)), cv = i - bin * c8;
    const int ph = res == 1 ? bin : (int)__umulhi((uint32_t)bin, magic_res), pw = bin - ph * res;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (merged) {
      roi_bin_merged<T>(img_base, sh32, sw32, s_w, s_start, s_num, ph, pw, cv, acc);
    } else {
      roialign_sample_loop<T>(f, img, cv, ph, pw, roi_start_h, roi_start_w, bin_h, bin_w, grid_h, grid_w, acc);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv_count;
    Vec8<T>::store(p.out.at(slot, ph, pw) + cv * 8, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// ROIAlign, COLUMN WALK (separable evaluation), one WARP-SIZED CTA per (ROI slot, bin column).
// ncu of the merged-tap kernel above: 234 M warp instructions for 7.0 M warp-taps (33 per tap: 12 are the bf16 unpack +
// FMA floor, the rest loop / address / weight overhead), issue-bound at 28 % of the HBM peak.  Here a thread owns
// (bin column pw, 8 channels) and walks the bin rows ph = 0..res-1 downwards:
//   out[ph][pw] = 1/count * sum_Y Wy[ph][Y] * v[Y],      v[Y] = sum_X Wx[pw][X] * f[Y][X]      (x pass of ONE feature row)
// * The x weights of the column live in registers for the whole walk (compile-time tap count NX: immediate-offset
//   accesses, no weight fetch, no inner loop), and v[Y] is kept for the following bins: with an adaptive sampling grid
//   neighbouring samples are <= 1 pixel apart, so the rows of bin ph + 1 start at most two rows before the last row of
//   bin ph -- the last two x-passed rows are carried in registers and a feature row is loaded and unpacked ONCE per
//   column (h + 2 rows instead of sum_ph ny[ph] ~ h + 21).
// * Raw feature vectors of the rows ahead are staged through shared memory with cp.async: every thread copies the 16
//   (bf16) / 32 (fp32) bytes of ITS OWN channels of tap k of row t + 2 into its private cell [slot = t & 1][k][lane]
//   while rows t and t + 1 are processed, and reads them back with one LDS.128 -- no barrier (a thread only ever reads
//   what it copied itself), a full row of lead time for the loads, no registers held across the latency.
// * Work granularity is one column (44 800 CTAs at cfg 5): with one CTA per ROI the box-size spread left the SMs idle
//   for 17 % of the kernel (tail) and every warp waited at the table-building barriers; a warp now builds the small
//   tables it needs itself (lane = bin row for the y weights, shuffles for the carry bookkeeping) and nothing waits.
// Columns with more than RoiColNx taps run the same walk with a run-time tap loop; ROIs whose rows do not follow each
// other without a gap (fixed sampling ratios with bins > 2 pixels) or whose tables do not fit use the sample loop.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void roi_fma8(float (&acc)[8], float w, const float (&v)[8]) {
  ffma2_roi(acc[0], acc[1], v[0], v[1], w); ffma2_roi(acc[2], acc[3], v[2], v[3], w);
  ffma2_roi(acc[4], acc[5], v[4], v[5], w); ffma2_roi(acc[6], acc[7], v[6], v[7], w);
}
// .cg: the staged lines bypass L1 (a column's taps are read once; with .ca the spill / table traffic lost its L1 hits)
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
template <typename T> struct RoiStage;
template <> struct RoiStage<__nv_bfloat16> {
  static constexpr int BYTES = 16;
  static __device__ __forceinline__ void copy(uint32_t saddr, const __nv_bfloat16* g) { cp_async16(saddr, g); }
  static __device__ __forceinline__ void read(uint32_t saddr, Raw8<__nv_bfloat16>& r) { r.q = lds128(saddr); }
};
template <> struct RoiStage<float> {
  static constexpr int BYTES = 32;
  static __device__ __forceinline__ void copy(uint32_t saddr, const float* g) { cp_async16(saddr, g); cp_async16(saddr + 16, g + 4); }
  static __device__ __forceinline__ void read(uint32_t saddr, Raw8<float>& r) {
    const uint4 a = lds128(saddr), b = lds128(saddr + 16);
    r.a = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
    r.b = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w));
  }
};
constexpr int ROI_COL_CLASSES = 6;                       // cost classes of the largest-first launch order
constexpr int ROI_WYT = 16;                              // row taps per bin the column kernel keeps (grid_h <= 15)
constexpr int ROI_COL_CELL = 96;                         // staging bytes per lane and slot (6 bf16 taps / 3 fp32 taps)
template <typename T> struct RoiColNx { static constexpr int value = ROI_COL_CELL / RoiStage<T>::BYTES; };

// shared-memory layout of the column kernel (one warp): [res][ROI_WYT] y weights, [ROI_MAXT] x weights, [res] rows per
// bin, [res] carried rows per bin, then 2 staging slots of ROI_COL_CELL bytes per lane
__host__ __device__ inline int roi_col_tables_bytes(int res) { return (res * ROI_WYT + ROI_MAXT + 2 * res) * 4; }
__host__ __device__ inline int roi_col_smem_bytes(int res) { return (roi_col_tables_bytes(res) + 15) / 16 * 16 + 2 * ROI_COL_CELL * 32; }

// NX > 0: compile-time tap count, weights in registers, rows staged by cp.async.  NX == 0: run-time tap count (wide
// columns), weights from shared memory, direct loads.  SW: compile-time pixel stride in elements (0 = run-time).
template <typename T, int NX, int SW>
__device__ __forceinline__ void roi_col_walk(const T* __restrict__ colp, int sh32, int sw_rt, int nx_rt,
                                             const float* __restrict__ wxs, const float* __restrict__ wys,
                                             const int* __restrict__ ny_s, const int* __restrict__ carry_s, int res,
                                             int y_first, int y_end, T* __restrict__ outp, long long out_sh,
                                             uint32_t stage) {
  constexpr int B = RoiStage<T>::BYTES;
  constexpr uint32_t kstride = 32 * B;                    // [slot][tap][lane] cells
  constexpr int NXR = NX > 0 ? NX : 1;
  const int sw = SW ? SW : sw_rt;
  const uint32_t cell0 = stage + (threadIdx.x & 31) * B;
  float wx[NXR];
  if (NX > 0) {
#pragma unroll
    for (int k = 0; k < NXR; ++k) wx[k] = wxs[k];
  }
  float cl[8], cp[8];                                     // x-passed rows y_done (cl) and y_done - 1 (cp)
#pragma unroll
  for (int k = 0; k < 8; ++k) cl[k] = cp[k] = 0.f;
  // rows y_first .. y_end - 1 are consumed in order, each exactly once
  const T* nextp = colp + (long long)y_first * sh32;      // next row to copy (NX > 0) / to read (NX == 0)
  int t_next = y_first, t_cur = y_first;
  if (NX > 0) {
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      if (t_next < y_end) {
#pragma unroll
        for (int k = 0; k < NXR; ++k) RoiStage<T>::copy(cell0 + (d * NXR + k) * kstride, nextp + k * sw);
        ++t_next;
        nextp += sh32;
      }
      cp_async_commit();
    }
  }
  for (int ph = 0; ph < res; ++ph, outp += out_sh) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int ny = ny_s[ph];
    const float* wy = wys + ph * ROI_WYT;                 // already divided by the sample count
    const int c = carry_s[ph];                            // leading rows of this bin that are x-passed already
    int j = 0;
    if (c == 2) {
      roi_fma8(acc, wy[0], cp);
      if (ny > 1) roi_fma8(acc, wy[1], cl);
      j = 2;
    } else if (c == 1) {
      roi_fma8(acc, wy[0], cl);
      j = 1;
    }
    for (; j < ny; ++j, ++t_cur) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { cp[k] = cl[k]; cl[k] = 0.f; }
      if (NX > 0) {
        const uint32_t cell = cell0 + ((t_cur - y_first) & 1) * (NXR * kstride);
        cp_async_wait1();                                 // row t_cur has landed (row t_cur + 1 may be in flight)
        Raw8<T> raw[NXR];
#pragma unroll
        for (int k = 0; k < NXR; ++k) RoiStage<T>::read(cell + k * kstride, raw[k]);
#pragma unroll
        for (int k = 0; k < NXR; ++k) raw[k].fma(wx[k], cl);
        if (t_next < y_end) {                             // rows below y_end are in bounds by construction
#pragma unroll
          for (int k = 0; k < NXR; ++k) RoiStage<T>::copy(cell + k * kstride, nextp + k * sw);
          ++t_next;
          nextp += sh32;
        }
        cp_async_commit();
      } else {
        const T* tp = nextp;
        int k = 0;
        for (; k + 2 <= nx_rt; k += 2, tp += 2 * sw) {
          Raw8<T> r0, r1;
          r0.load(tp); r1.load(tp + sw);
          r0.fma(wxs[k], cl); r1.fma(wxs[k + 1], cl);
        }
        if (k < nx_rt) {
          Raw8<T> r0;
          r0.load(tp);
          r0.fma(wxs[k], cl);
        }
        nextp += sh32;
      }
      roi_fma8(acc, wy[j], cl);
    }
    Vec8<T>::store(outp, acc);
  }
}

template <typename T, int SW>
__device__ __forceinline__ void roi_col_dispatch(int nx, const T* __restrict__ colp, int sh32, int sw_rt,
                                                 const float* __restrict__ wxs, const float* __restrict__ wys,
                                                 const int* __restrict__ ny_s, const int* __restrict__ carry_s, int res,
                                                 int y_first, int y_end, T* __restrict__ outp, long long out_sh,
                                                 uint32_t stage) {
#define CM2_ROI_COL_CASE(NXV)                                                                                          \
  case NXV:                                                                                                            \
    if constexpr (NXV <= RoiColNx<T>::value) {                                                                         \
      roi_col_walk<T, NXV, SW>(colp, sh32, sw_rt, nx, wxs, wys, ny_s, carry_s, res, y_first, y_end, outp, out_sh, stage); \
      return;                                                                                                          \
    }                                                                                                                  \
    break;
  switch (nx) {
    CM2_ROI_COL_CASE(1) CM2_ROI_COL_CASE(2) CM2_ROI_COL_CASE(3) CM2_ROI_COL_CASE(4) CM2_ROI_COL_CASE(5) CM2_ROI_COL_CASE(6)
    default: break;
  }
#undef CM2_ROI_COL_CASE
  roi_col_walk<T, 0, SW>(colp, sh32, sw_rt, nx, wxs, wys, ny_s, carry_s, res, y_first, y_end, outp, out_sh, stage);
}

// Launch order of the column kernel: ROI slots bucketed by decreasing cost of one bin column (~ feature rows x taps per
// row; octave classes, empty slots last), slots ascending inside a class (STABLE counting sort: the CTAs in flight at
// any time then belong to few images, whose feature maps stay in L2).  One CTA; warp w owns a contiguous range of slots,
// ranks inside a 32-slot chunk come from __match_any_sync.  The order only affects scheduling, never a result.
template <typename T>
__global__ void __launch_bounds__(1024) roi_order_kernel(const RoiAlignParams<T> p) {
  constexpr int NC = ROI_COL_CLASSES + 1;
  __shared__ int s_cnt[32][NC];                         // per warp and class: count, then start position
  __shared__ int s_base[NC];
  const int nslots = p.n * p.r_cap, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = ((nslots + 31) / 32 + 31) / 32 * 32;  // slots per warp, a multiple of the chunk size
  const int lo = warp * per, hi = min(lo + per, nslots);
  const unsigned full = 0xffffffffu;
  auto cost_class = [&](int slot) {
    const int img = slot / p.r_cap, r = slot - img * p.r_cap;
    if (r >= p.det_count[img]) return ROI_COL_CLASSES;
    const float4 bx = reinterpret_cast<const float4*>(p.boxes)[slot];
    int lvl = assign_level(bx.x, bx.y, bx.z, bx.w, p.image_area[img], p.crit, p.min_level, p.max_level);
    if (lvl >= p.num_levels) lvl = p.num_levels - 1;
    float scale = p.scale[0];
#pragma unroll
    for (int l = 1; l < ROI_MAX_LEVELS; ++l)
      if (l == lvl) scale = p.scale[l];
    const float rows = (bx.w - bx.y) * scale + 2.f, taps = (bx.z - bx.x) * scale / (float)p.res + 2.f;
    const float cost = fmaxf(rows * taps, 1.f);                        // NaN -> 1
    // cost < 32 -> last class, doubling per class (2, 3, 4 and 12 classes were measured too: 6 is as good as any)
    const int cls = ROI_COL_CLASSES - 1 - ((int)log2f(cost) - 4);
    return min(max(cls, 0), ROI_COL_CLASSES - 1);
  };
  for (int i = threadIdx.x; i < 32 * NC; i += blockDim.x) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  for (int c0 = lo; c0 < hi; c0 += 32) {
    const int slot = c0 + lane;
    const int cls = slot < hi ? cost_class(slot) : NC;                 // NC: no slot
    const unsigned m = __match_any_sync(full, cls);
    if (cls < NC && lane == __ffs(m) - 1) s_cnt[warp][cls] += __popc(m);
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < NC) {                               // per class: exclusive prefix over the warps
    int run = 0;
    for (int w = 0; w < 32; ++w) { const int c = s_cnt[w][threadIdx.x]; s_cnt[w][threadIdx.x] = run; run += c; }
    s_base[threadIdx.x] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 0; k < NC; ++k) { const int c = s_base[k]; s_base[k] = run; run += c; }
  }
  __syncthreads();
  for (int c0 = lo; c0 < hi; c0 += 32) {
    const int slot = c0 + lane;
    const int cls = slot < hi ? cost_class(slot) : NC;
    const unsigned m = __match_any_sync(full, cls);
    if (cls < NC) p.order[s_base[cls] + s_cnt[warp][cls] + __popc(m & ((1u << lane) - 1u))] = slot;
    __syncwarp();
    if (cls < NC && lane == __ffs(m) - 1) s_cnt[warp][cls] += __popc(m);
    __syncwarp();
  }
}

template <typename T>
__global__ void __launch_bounds__(32, 24) roialign_col_kernel(const RoiAlignParams<T> p) {
  extern __shared__ __align__(16) unsigned char s_col[];
  const int res = p.res, c8 = p.out.c >> 3;             // c8 <= 32 lanes walk, all 32 build the tables
  float* s_wy = reinterpret_cast<float*>(s_col);        // [res][ROI_WYT]
  float* s_wx = s_wy + res * ROI_WYT;                   // [ROI_MAXT]
  int* s_ny = reinterpret_cast<int*>(s_wx + ROI_MAXT);  // [res]
  int* s_carry = s_ny + res;                            // [res]
  const uint32_t stage = (uint32_t)__cvta_generic_to_shared(s_col) + (roi_col_tables_bytes(res) + 15) / 16 * 16;
  const int lane = threadIdx.x;
  // CTAs start in blockIdx order: p.order lists the ROI slots by decreasing column cost (roi_order_kernel), so the long
  // columns (up to ~90 rows x 6 taps, tens of microseconds for one warp) run first and the short ones fill the tail.
  // (CTAs of 7 or 14 warps = neighbouring columns sharing taps through L1 were measured too: same time.)
  const int oi = blockIdx.x / res, pw = blockIdx.x - oi * res;
  const int slot = p.order[oi];
  const int img = slot / p.r_cap, r = slot - img * p.r_cap;
  T* outp = p.out.at(slot, 0, pw) + lane * 8;
  if (r >= p.det_count[img]) {                          // empty slot: zeros
    if (lane >= c8) return;
    const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int ph = 0; ph < res; ++ph, outp += p.out.sh) Vec8<T>::store(outp, z);
    return;
  }
  const float4 bx = reinterpret_cast<const float4*>(p.boxes)[slot];
  int lvl = assign_level(bx.x, bx.y, bx.z, bx.w, p.image_area[img], p.crit, p.min_level, p.max_level);
  if (lvl >= p.num_levels) lvl = p.num_levels - 1;
  if (p.level_out && lane == 0 && pw == 0) p.level_out[slot] = lvl;
  View<const T> f = p.feat[0];
  float scale = p.scale[0];
#pragma unroll
  for (int l = 1; l < ROI_MAX_LEVELS; ++l)
    if (l == lvl) { f = p.feat[l]; scale = p.scale[l]; }
  const float roi_start_w = bx.x * scale - 0.5f, roi_start_h = bx.y * scale - 0.5f;
  const float roi_end_w = bx.z * scale - 0.5f, roi_end_h = bx.w * scale - 0.5f;
  const float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
  const float bin_h = roi_height / (float)res, bin_w = roi_width / (float)res;
  const int grid_h = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_height / (float)res);
  const int grid_w = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_width / (float)res);
  const float inv_count = 1.0f / fmaxf((float)(grid_h * grid_w), 1.f);
  const T* img_base = f.p + (size_t)img * f.sn;
  const int sh32 = (int)f.sh, sw32 = (int)f.sw;

  // ---- tables: lane b < res builds the y weights of bin row b, lane 31 the x weights of this column
  for (int i = lane; i < res * ROI_WYT + ROI_MAXT; i += 32) s_wy[i] = 0.f;      // s_wx follows s_wy
  __syncwarp();
  int t0 = 0, tn = 0;                                   // first feature row / column of "my" bin, rows / columns it spans
  bool bad = false;
  {
    const bool xlane = lane == 31;
    const bool active = xlane || lane < res;
    const float start = xlane ? roi_start_w : roi_start_h, bin = xlane ? bin_w : bin_h;
    const int pb = xlane ? pw : lane, grid = active ? (xlane ? grid_w : grid_h) : 0;
    const int extent = xlane ? f.w : f.h, cap = xlane ? ROI_MAXT : ROI_WYT;
    const float wscale = xlane ? 1.f : inv_count;       // 1 / count folded into the y weights
    float* wrow = xlane ? s_wx : s_wy + lane * ROI_WYT;
    int s0 = -1;
    for (int i = 0; i < grid; ++i) {
      const float v = start + pb * bin + ((float)i + 0.5f) * bin / (float)grid;
      if (v < -1.0f || v > (float)extent) continue;
      float vv = v <= 0.f ? 0.f : v;
      int low = (int)vv, high;
      if (low >= extent - 1) { high = low = extent - 1; vv = (float)low; } else high = low + 1;
      const float l = vv - (float)low, h = 1.f - l;
      if (s0 < 0) s0 = low;
      if (high - s0 >= cap) { bad = true; break; }
      wrow[low - s0] += h * wscale;
      wrow[high - s0] += l * wscale;
      tn = high - s0 + 1;
    }
    t0 = s0 < 0 ? 0 : s0;
  }
  const unsigned full = 0xffffffffu;
  const int sx = __shfl_sync(full, t0, 31), nx = __shfl_sync(full, tn, 31);
  // carry bookkeeping with shuffles: rows of the bins above end at y_done; rows y_done - 1, y_done are in (cp, cl)
  const bool ybin = lane < res && tn > 0;
  int pm = ybin ? t0 + tn - 1 : -1;                     // inclusive prefix max of the last row of a bin
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(full, pm, o);
    if (lane >= o) pm = max(pm, t);
  }
  int y_done = __shfl_up_sync(full, pm, 1);
  if (lane == 0) y_done = -1;
  int y_first = ybin ? t0 : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) y_first = min(y_first, __shfl_xor_sync(full, y_first, o));
  int y_end = __shfl_sync(full, pm, 31) + 1;
  if (y_first == 0x7fffffff) { y_first = 0; y_end = 0; }
  int c = 0;
  if (ybin && y_done >= 0) {
    c = y_done - t0 + 1;
    if (c < 0 || c > min(2, y_done - y_first + 1)) bad = true;       // gap, or a row that is not carried any more
  }
  const bool walk = !__any_sync(full, bad);
  if (lane < res) { s_ny[lane] = tn; s_carry[lane] = c; }
  __syncwarp();
  if (lane >= c8) return;

  if (walk) {
    if (nx == 0) {                                       // column entirely outside the feature map
      const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
      for (int ph = 0; ph < res; ++ph, outp += p.out.sh) Vec8<T>::store(outp, z);
      return;
    }
    const T* colp = img_base + sx * sw32 + lane * 8;
    if (sw32 == 256)
      roi_col_dispatch<T, 256>(nx, colp, sh32, sw32, s_wx, s_wy, s_ny, s_carry, res, y_first, y_end, outp, p.out.sh, stage);
    else
      roi_col_dispatch<T, 0>(nx, colp, sh32, sw32, s_wx, s_wy, s_ny, s_carry, res, y_first, y_end, outp, p.out.sh, stage);
    return;
  }
  // sample loop (torchvision's own order of evaluation) for the bins of this column
#pragma unroll 1
  for (int ph = 0; ph < res; ++ph, outp += p.out.sh) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    roialign_sample_loop<T>(f, img, lane, ph, pw, roi_start_h, roi_start_w, bin_h, bin_w, grid_h, grid_w, acc);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv_count;
    Vec8<T>::store(outp, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// Spatial attention (sam.py:23-28).  One CTA per ROI: phase 1 channel mean / max per pixel (warp per
// pixel), phase 2 3x3 conv over the [2, s, s] map + sigmoid in shared memory, phase 3 scale.
// ---------------------------------------------------------------------------------------------
constexpr int SAM_MAX_S = 32;

template <typename T>
__global__ void __launch_bounds__(256) spatial_attention_kernel(View<const T> x, View<T> out,
                                                                const float* __restrict__ w18) {
  __shared__ float s_avg[SAM_MAX_S * SAM_MAX_S], s_max[SAM_MAX_S * SAM_MAX_S], s_att[SAM_MAX_S * SAM_MAX_S];
  __shared__ float s_w[18];
  const int r = blockIdx.x;
  const int S = x.h, C = x.c, c8 = C >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x < 18) s_w[threadIdx.x] = w18[threadIdx.x];
  for (int pix = warp; pix < S * S; pix += nwarps) {
    int y = pix / S, xx = pix - y * S;
    const T* row = x.at(r, y, xx);
    float sum = 0.f, mx = -INFINITY;
    for (int cv = lane; cv < c8; cv += 32) {
      float v[8];
      Vec8<T>::load(row + cv * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) { sum += v[k]; mx = fmaxf(mx, v[k]); }
    }
    sum = warp_sum(sum);
    mx = warp_max(mx);
    if (lane == 0) { s_avg[pix] = sum / (float)C; s_max[pix] = mx; }
  }
  __syncthreads();
  for (int pix = threadIdx.x; pix < S * S; pix += blockDim.x) {
    int y = pix / S, xx = pix - y * S;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        int iy = y + ky - 1, ix = xx + kx - 1;
        if (iy >= 0 && iy < S && ix >= 0 && ix < S) {
          acc = fmaf(s_w[ky * 3 + kx], s_avg[iy * S + ix], acc);
          acc = fmaf(s_w[9 + ky * 3 + kx], s_max[iy * S + ix], acc);
        }
      }
    s_att[pix] = sigmoid_f32(acc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * S * c8; i += blockDim.x) {
    int cv = i % c8, pix = i / c8;
    int y = pix / S, xx = pix - y * S;
    float v[8];
    Vec8<T>::load(x.at(r, y, xx) + cv * 8, v);
    float a = s_att[pix];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= a;
    Vec8<T>::store(out.at(r, y, xx) + cv * 8, v);
  }
}

// bf16 variant that keeps the ROI's [S, S, C] tile in shared memory between the pooling and the scaling phase
// (100 KB for 14x14x256): the tensor is read once and written once instead of read twice.
__global__ void __launch_bounds__(512) spatial_attention_smem_kernel(View<const __nv_bfloat16> x, View<__nv_bfloat16> out,
                                                                     const float* __restrict__ w18) {
  extern __shared__ uint4 s_x[];                        // [S*S][c8]
  __shared__ float s_avg[SAM_MAX_S * SAM_MAX_S], s_max[SAM_MAX_S * SAM_MAX_S], s_att[SAM_MAX_S * SAM_MAX_S];
  __shared__ float s_w[18];
  const int r = blockIdx.x;
  const int S = x.h, C = x.c, c8 = C >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x < 18) s_w[threadIdx.x] = w18[threadIdx.x];
  // phase 1: a warp owns pixels warp, warp + nwarps, ...; SAM_BATCH pixel rows are requested before the first one is
  // reduced (a load -> reduce loop pays one memory latency per pixel: 13 per warp for a 14x14 ROI)
  constexpr int SAM_BATCH = 7;
  for (int pix0 = warp; pix0 < S * S; pix0 += nwarps * SAM_BATCH) {
    for (int cv = lane; cv < c8; cv += 32) {
      uint4 q[SAM_BATCH];
#pragma unroll
      for (int b = 0; b < SAM_BATCH; ++b) {
        const int pix = pix0 + b * nwarps;
        if (pix < S * S) {
          const int y = pix / S, xx = pix - y * S;
          q[b] = __ldg(reinterpret_cast<const uint4*>(x.at(r, y, xx)) + cv);
        }
      }
#pragma unroll
      for (int b = 0; b < SAM_BATCH; ++b) {
        const int pix = pix0 + b * nwarps;
        if (pix < S * S) s_x[pix * c8 + cv] = q[b];
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int b = 0; b < SAM_BATCH; ++b) {
      const int pix = pix0 + b * nwarps;
      if (pix >= S * S) break;
      float sum = 0.f, mx = -INFINITY;
      for (int cv = lane; cv < c8; cv += 32) {
        const uint4 q = s_x[pix * c8 + cv];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = __uint_as_float(w[k] << 16), hi = __uint_as_float(w[k] & 0xffff0000u);
          sum += lo; sum += hi;
          mx = fmaxf(mx, fmaxf(lo, hi));
        }
      }
      sum = warp_sum(sum);
      mx = warp_max(mx);
      if (lane == 0) { s_avg[pix] = sum / (float)C; s_max[pix] = mx; }
    }
  }
  __syncthreads();
  for (int pix = threadIdx.x; pix < S * S; pix += blockDim.x) {
    const int y = pix / S, xx = pix - y * S;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = y + ky - 1, ix = xx + kx - 1;
        if (iy >= 0 && iy < S && ix >= 0 && ix < S) {
          acc = fmaf(s_w[ky * 3 + kx], s_avg[iy * S + ix], acc);
          acc = fmaf(s_w[9 + ky * 3 + kx], s_max[iy * S + ix], acc);
        }
      }
    s_att[pix] = sigmoid_f32(acc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * S * c8; i += blockDim.x) {
    const int pix = i / c8, cv = i - pix * c8;
    const int y = pix / S, xx = pix - y * S;
    const uint4 q = s_x[i];
    const float a = s_att[pix];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(w[k] << 16) * a, __uint_as_float(w[k] & 0xffff0000u) * a);
      o[k] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(out.at(r, y, xx) + cv * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Persistent, double-buffered variant: one CTA per SM walks over the ROIs; the [S, S, C] tile of the NEXT ROI is
// fetched by the copy engine (cp.async.bulk, one contiguous S * C run per tile row, completion on an mbarrier) into the
// other shared-memory buffer while the current tile is pooled, convolved, scaled and stored -- the load latency that
// the CTA-per-ROI kernel pays three times per SM residency slot disappears behind the store phase.
__device__ __forceinline__ void sam_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 26)) { printf("spatial_attention_pipe: mbarrier wait timed out (block %d)\n", blockIdx.x); __trap(); }
  }
}

__global__ void __launch_bounds__(1024, 1) spatial_attention_pipe_kernel(View<const __nv_bfloat16> x, View<__nv_bfloat16> out,
                                                                         const float* __restrict__ w18) {
  extern __shared__ uint4 s_x[];                        // [2][S*S*c8] tiles, then avg / max / att maps
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ float s_w[18];
  const int S = x.h, C = x.c, c8 = C >> 3, npix = S * S;
  const int tile_vecs = npix * c8;
  float* s_avg = reinterpret_cast<float*>(s_x + 2 * tile_vecs);
  float* s_max = s_avg + npix;
  float* s_att = s_max + npix;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t row_bytes = (uint32_t)S * C * 2;
  if (threadIdx.x < 18) s_w[threadIdx.x] = w18[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar[b])), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int r, int buf) {                    // one thread
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[buf]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic reads of the buffer precede the async writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (uint32_t)S) : "memory");
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(s_x + (size_t)buf * tile_vecs);
    for (int y = 0; y < S; ++y)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(dst0 + (uint32_t)y * row_bytes), "l"(x.at(r, y, 0)), "r"(row_bytes), "r"(bar) : "memory");
  };
  const int nroi = x.n;
  if (threadIdx.x == 0 && (int)blockIdx.x < nroi) issue(blockIdx.x, 0);
  int k = 0;
  for (int r = blockIdx.x; r < nroi; r += gridDim.x, ++k) {
    const int buf = k & 1;
    if (threadIdx.x == 0 && r + (int)gridDim.x < nroi) issue(r + gridDim.x, buf ^ 1);
    sam_mbar_wait((uint32_t)__cvta_generic_to_shared(&s_bar[buf]), (uint32_t)(k >> 1) & 1u);
    const uint4* tile = s_x + (size_t)buf * tile_vecs;
    // phase 1: channel mean / max per pixel (warp per pixel)
    for (int pix = warp; pix < npix; pix += nwarps) {
      float sum = 0.f, mx = -INFINITY;
      for (int cv = lane; cv < c8; cv += 32) {
        const uint4 q = tile[pix * c8 + cv];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
          sum += lo; sum += hi;
          mx = fmaxf(mx, fmaxf(lo, hi));
        }
      }
      sum = warp_sum(sum);
      mx = warp_max(mx);
      if (lane == 0) { s_avg[pix] = sum / (float)C; s_max[pix] = mx; }
    }
    __syncthreads();
    // phase 2: 3x3 conv over (avg, max) + sigmoid
    for (int pix = threadIdx.x; pix < npix; pix += blockDim.x) {
      const int y = pix / S, xx = pix - y * S;
      float acc = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int iy = y + ky - 1, ix = xx + kx - 1;
          if (iy >= 0 && iy < S && ix >= 0 && ix < S) {
            acc = fmaf(s_w[ky * 3 + kx], s_avg[iy * S + ix], acc);
            acc = fmaf(s_w[9 + ky * 3 + kx], s_max[iy * S + ix], acc);
          }
        }
      s_att[pix] = sigmoid_f32(acc);
    }
    __syncthreads();
    // phase 3: scale and store
    for (int i = threadIdx.x; i < tile_vecs; i += blockDim.x) {
      const int pix = i / c8, cv = i - pix * c8;
      const int y = pix / S, xx = pix - y * S;
      const uint4 q = tile[i];
      const float a = s_att[pix];
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(w[j] << 16) * a, __uint_as_float(w[j] & 0xffff0000u) * a);
        o[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(out.at(r, y, xx) + cv * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();                                    // the tile and the maps are free again
  }
}

// ---------------------------------------------------------------------------------------------
// ROIAlign, column walk over a SHARED ROW RING (variant 3, CM2_ROIALIGN_VARIANT=3; NOT the default: it moves a third
// of the L2 -> SM bytes of variant 2 and runs in the same time, 0.263 vs 0.258 ms at cfg 5 -- with the launch order
// mixing images both read the feature maps ~2x from DRAM, ~1 GB of DRAM traffic in 0.26 ms).  ncu of the
// warp-per-column kernel: the L2 -> SM path is the
// busiest unit (143 M sectors = 4.6 GB at cfg 5, 71 % of the L2 peak) because every column fetches its own copy of the
// taps it shares with its neighbours and with the bins above.  Here one CTA owns one ROI: a producer warp streams the
// feature rows of the ROI (one contiguous run of (x_hi - x_lo) pixels per row, NHWC) into a ring of shared-memory slots
// with cp.async.bulk + mbarriers (full / empty per slot, up to 8 rows ahead), and the 14 column warps run the same walk
// as roi_col_walk with their taps read from the ring by LDS.128 -- every feature byte crosses L2 -> SM once per ROI.
// Rows wider than half the ring, pixel strides != c, and irregular ROIs use the direct-load walk / the sample loop.
// ---------------------------------------------------------------------------------------------
constexpr int ROI_RING_BYTES = 96 * 1024;
constexpr int ROI_RING_MAXD = 8;
constexpr int ROI_RING_COLS = 14;                        // consumer warps (bin columns) per CTA; one more warp produces
constexpr int ROI_RING_NX = 4;                           // taps per row with the weights in registers (64-register budget)

__device__ __forceinline__ void roi_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// NX > 0: compile-time tap count; NX == 0: run-time count, weights from shared memory.  PB: pixel bytes (0 = run-time).
template <typename T, int NX, int PB>
__device__ __forceinline__ void roi_ring_walk(uint32_t ring, uint32_t slot_bytes, int depth, uint32_t bars, uint32_t col_off,
                                              uint32_t pb_rt, int nx_rt, bool active, const float* __restrict__ wxs,
                                              const float* __restrict__ wys, const int* __restrict__ ny_s,
                                              const int* __restrict__ carry_s, int res, T* __restrict__ outp,
                                              long long out_sh) {
  constexpr int NXR = NX > 0 ? NX : 1;
  const uint32_t pb = PB ? PB : pb_rt;
  const int lane = threadIdx.x & 31;
  float wx[NXR];
  if (NX > 0) {
#pragma unroll
    for (int k = 0; k < NXR; ++k) wx[k] = wxs[k];
  }
  float cl[8], cp[8];                                     // x-passed rows y_done (cl) and y_done - 1 (cp)
#pragma unroll
  for (int k = 0; k < 8; ++k) cl[k] = cp[k] = 0.f;
  int slot = 0;
  uint32_t phase = 0;
  for (int ph = 0; ph < res; ++ph, outp += out_sh) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int ny = ny_s[ph];
    const float* wy = wys + ph * ROI_WYT;                 // already divided by the sample count
    const int c = carry_s[ph];                            // leading rows of this bin that are x-passed already
    int j = 0;
    if (c == 2) {
      roi_fma8(acc, wy[0], cp);
      if (ny > 1) roi_fma8(acc, wy[1], cl);
      j = 2;
    } else if (c == 1) {
      roi_fma8(acc, wy[0], cl);
      j = 1;
    }
    for (; j < ny; ++j) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { cp[k] = cl[k]; cl[k] = 0.f; }
      sam_mbar_wait(bars + slot * 8, phase);              // the row has landed in the slot
      const uint32_t base = ring + slot * slot_bytes + col_off;
      if (active) {
        if (NX > 0) {
          Raw8<T> raw[NXR];
#pragma unroll
          for (int k = 0; k < NXR; ++k) RoiStage<T>::read(base + k * pb, raw[k]);
#pragma unroll
          for (int k = 0; k < NXR; ++k) raw[k].fma(wx[k], cl);
        } else {
          uint32_t a = base;
          for (int k = 0; k < nx_rt; ++k, a += pb) {
            Raw8<T> r0;
            RoiStage<T>::read(a, r0);
            r0.fma(wxs[k], cl);
          }
        }
      }
      __syncwarp();                                       // every lane has its values in registers
      if (lane == 0) roi_mbar_arrive(bars + (ROI_RING_MAXD + slot) * 8);
      if (++slot == depth) { slot = 0; phase ^= 1u; }
      roi_fma8(acc, wy[j], cl);
    }
    if (active) Vec8<T>::store(outp, acc);
  }
}

template <typename T, int PB>
__device__ __forceinline__ void roi_ring_dispatch(int nx, uint32_t ring, uint32_t slot_bytes, int depth, uint32_t bars,
                                                  uint32_t col_off, uint32_t pb_rt, bool active, const float* __restrict__ wxs,
                                                  const float* __restrict__ wys, const int* __restrict__ ny_s,
                                                  const int* __restrict__ carry_s, int res, T* __restrict__ outp,
                                                  long long out_sh) {
  constexpr int NXMAX = sizeof(T) == 2 ? ROI_RING_NX : 2;
#define CM2_ROI_RING_CASE(NXV)                                                                                          \
  case NXV:                                                                                                             \
    if constexpr (NXV <= NXMAX) {                                                                                       \
      roi_ring_walk<T, NXV, PB>(ring, slot_bytes, depth, bars, col_off, pb_rt, nx, active, wxs, wys, ny_s, carry_s, res, \
                                outp, out_sh);                                                                          \
      return;                                                                                                           \
    }                                                                                                                   \
    break;
  switch (nx) {
    CM2_ROI_RING_CASE(1) CM2_ROI_RING_CASE(2) CM2_ROI_RING_CASE(3) CM2_ROI_RING_CASE(4)
    default: break;
  }
#undef CM2_ROI_RING_CASE
  roi_ring_walk<T, 0, PB>(ring, slot_bytes, depth, bars, col_off, pb_rt, nx, active, wxs, wys, ny_s, carry_s, res, outp, out_sh);
}

template <typename T>
__global__ void __launch_bounds__(32 * (ROI_RING_COLS + 1), 2) roialign_ring_kernel(const RoiAlignParams<T> p) {
  extern __shared__ __align__(128) unsigned char s_ring[];
  __shared__ __align__(8) unsigned long long s_bar[2 * ROI_RING_MAXD];       // full[0..7], empty[8..15]
  __shared__ float s_wy[ROI_MAX_RES * ROI_WYT];
  __shared__ float s_wx[ROI_RING_COLS][ROI_MAXT];
  __shared__ int s_ny[ROI_MAX_RES], s_carry[ROI_MAX_RES], s_sx[ROI_RING_COLS], s_nx[ROI_RING_COLS];
  __shared__ int s_info[3];                             // first needed feature row, one past the last, walk usable
  const int res = p.res, c8 = p.out.c >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = (res + ROI_RING_COLS - 1) / ROI_RING_COLS;
  const int oi = blockIdx.x / groups, pw0 = (blockIdx.x - oi * groups) * ROI_RING_COLS;
  const int ncons = min(ROI_RING_COLS, res - pw0);      // column warps of this CTA
  const int pw = pw0 + warp;
  const bool consumer = warp < ncons, active = lane < c8;
  const int slot = p.order[oi];
  const int img = slot / p.r_cap, r = slot - img * p.r_cap;
  T* outp = p.out.at(slot, 0, consumer ? pw : 0) + lane * 8;
  if (r >= p.det_count[img]) {                          // empty slot: zeros (CTA-uniform)
    if (!consumer || !active) return;
    const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int ph = 0; ph < res; ++ph, outp += p.out.sh) Vec8<T>::store(outp, z);
    return;
  }
  const float4 bx = reinterpret_cast<const float4*>(p.boxes)[slot];
  int lvl = assign_level(bx.x, bx.y, bx.z, bx.w, p.image_area[img], p.crit, p.min_level, p.max_level);
  if (lvl >= p.num_levels) lvl = p.num_levels - 1;
  if (p.level_out && threadIdx.x == 0 && pw0 == 0) p.level_out[slot] = lvl;
  View<const T> f = p.feat[0];
  float scale = p.scale[0];
#pragma unroll
  for (int l = 1; l < ROI_MAX_LEVELS; ++l)
    if (l == lvl) { f = p.feat[l]; scale = p.scale[l]; }
  const float roi_start_w = bx.x * scale - 0.5f, roi_start_h = bx.y * scale - 0.5f;
  const float roi_end_w = bx.z * scale - 0.5f, roi_end_h = bx.w * scale - 0.5f;
  const float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
  const float bin_h = roi_height / (float)res, bin_w = roi_width / (float)res;
  const int grid_h = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_height / (float)res);
  const int grid_w = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_width / (float)res);
  const float inv_count = 1.0f / fmaxf((float)(grid_h * grid_w), 1.f);
  const T* img_base = f.p + (size_t)img * f.sn;
  const int sh32 = (int)f.sh, sw32 = (int)f.sw;
  const unsigned full = 0xffffffffu;

  // ---- tables: warp 0 builds the y weights (lane = bin row) and the carry bookkeeping, lane 31 of every column warp
  //      the x weights of its column
  if (warp == 0)
    for (int i = lane; i < res * ROI_WYT; i += 32) s_wy[i] = 0.f;
  if (consumer) s_wx[warp][lane] = 0.f;                 // ROI_MAXT == 32
  __syncwarp();
  if (consumer) {
    int t0 = 0, tn = 0;                                 // first feature row / column of "my" bin, rows / columns it spans
    bool bad = false;
    const bool xlane = lane == 31;
    const bool build = xlane || (warp == 0 && lane < res);
    const float start = xlane ? roi_start_w : roi_start_h, bin = xlane ? bin_w : bin_h;
    const int pb = xlane ? pw : lane, grid = build ? (xlane ? grid_w : grid_h) : 0;
    const int extent = xlane ? f.w : f.h, cap = xlane ? ROI_MAXT : ROI_WYT;
    const float wscale = xlane ? 1.f : inv_count;       // 1 / count folded into the y weights
    float* wrow = xlane ? s_wx[warp] : s_wy + lane * ROI_WYT;
    int s0 = -1;
    for (int i = 0; i < grid; ++i) {
      const float v = start + pb * bin + ((float)i + 0.5f) * bin / (float)grid;
      if (v < -1.0f || v > (float)extent) continue;
      float vv = v <= 0.f ? 0.f : v;
      int low = (int)vv, high;
      if (low >= extent - 1) { high = low = extent - 1; vv = (float)low; } else high = low + 1;
      const float l = vv - (float)low, h = 1.f - l;
      if (s0 < 0) s0 = low;
      if (high - s0 >= cap) { bad = true; break; }
      wrow[low - s0] += h * wscale;
      wrow[high - s0] += l * wscale;
      tn = high - s0 + 1;
    }
    t0 = s0 < 0 ? 0 : s0;
    if (xlane) { s_sx[warp] = t0; s_nx[warp] = bad ? -1 : tn; }
    if (warp == 0) {
      // carry bookkeeping with shuffles: rows of the bins above end at y_done; rows y_done - 1, y_done are in (cp, cl)
      if (xlane) { bad = false; tn = 0; }
      const bool ybin = lane < res && tn > 0;
      int pm = ybin ? t0 + tn - 1 : -1;                 // inclusive prefix max of the last row of a bin
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(full, pm, o);
        if (lane >= o) pm = max(pm, t);
      }
      int y_done = __shfl_up_sync(full, pm, 1);
      if (lane == 0) y_done = -1;
      int y_first = ybin ? t0 : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) y_first = min(y_first, __shfl_xor_sync(full, y_first, o));
      int y_end = __shfl_sync(full, pm, 31) + 1;
      if (y_first == 0x7fffffff) { y_first = 0; y_end = 0; }
      int c = 0;
      if (ybin && y_done >= 0) {
        c = y_done - t0 + 1;
        if (c < 0 || c > min(2, y_done - y_first + 1)) bad = true;     // gap, or a row that is not carried any more
      }
      const bool walk = !__any_sync(full, bad);
      if (lane < res) { s_ny[lane] = tn; s_carry[lane] = c; }
      if (lane == 0) { s_info[0] = y_first; s_info[1] = y_end; s_info[2] = walk ? 1 : 0; }
    }
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < ROI_RING_MAXD; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar[b])), "r"(1u) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar[ROI_RING_MAXD + b])),
                   "r"((uint32_t)ncons) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int y_first = s_info[0], y_end = s_info[1];
  int x_lo = 0x7fffffff, x_hi = 0;
  bool xbad = false;
  for (int w = 0; w < ncons; ++w) {
    const int n = s_nx[w], x0 = s_sx[w];
    if (n < 0) xbad = true;
    if (n > 0) { x_lo = min(x_lo, x0); x_hi = max(x_hi, x0 + n); }
  }
  const int nrows = y_end - y_first;
  const uint32_t pix_bytes = (uint32_t)f.c * sizeof(T);
  const bool walk = s_info[2] != 0 && !xbad;
  const bool none = nrows <= 0 || x_hi <= x_lo;         // nothing inside the feature map: zeros
  const uint32_t row_bytes = none ? 0u : (uint32_t)(x_hi - x_lo) * pix_bytes;
  const uint32_t slot_bytes = (row_bytes + 127u) & ~127u;
  const bool ring_ok = walk && !none && sw32 == f.c && 2u * slot_bytes <= (uint32_t)ROI_RING_BYTES;
  if (walk && (none || ring_ok)) {
    const int depth = none ? 1 : min(ROI_RING_MAXD, (int)((uint32_t)ROI_RING_BYTES / slot_bytes));
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(s_ring), bars = (uint32_t)__cvta_generic_to_shared(s_bar);
    if (warp == ROI_RING_COLS) {                        // producer
      if (lane == 0 && !none) {
        const T* src = img_base + (long long)y_first * sh32 + (long long)x_lo * sw32;
        int sl = 0;
        uint32_t phase = 0;
        for (int t = 0; t < nrows; ++t, src += sh32) {
          if (t >= depth) sam_mbar_wait(bars + (ROI_RING_MAXD + sl) * 8, phase ^ 1u);      // consumers released the slot
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + sl * 8), "r"(row_bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       :: "r"(ring + sl * slot_bytes), "l"(src), "r"(row_bytes), "r"(bars + sl * 8) : "memory");
          if (++sl == depth) { sl = 0; phase ^= 1u; }
        }
      }
      return;
    }
    if (!consumer) return;
    const int nx = none ? 0 : s_nx[warp];
    const uint32_t col_off = none ? 0u : (uint32_t)(s_sx[warp] - x_lo) * pix_bytes + lane * RoiStage<T>::BYTES;
    if (none) {
      if (!active) return;
      const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
      for (int ph = 0; ph < res; ++ph, outp += p.out.sh) Vec8<T>::store(outp, z);
      return;
    }
    // a column outside the feature map (nx == 0) still consumes every row: the run-time walk with zero taps
    if (pix_bytes == 512)
      roi_ring_dispatch<T, 512>(nx, ring, slot_bytes, depth, bars, nx > 0 ? col_off : 0u, pix_bytes, active, s_wx[warp], s_wy, s_ny,
                                s_carry, res, outp, p.out.sh);
    else
      roi_ring_dispatch<T, 0>(nx, ring, slot_bytes, depth, bars, nx > 0 ? col_off : 0u, pix_bytes, active, s_wx[warp], s_wy, s_ny,
                              s_carry, res, outp, p.out.sh);
    return;
  }
  if (!consumer || !active) return;
  if (walk) {                                           // rows too wide for the ring / pitched pixels: direct-load walk
    const int nx = s_nx[warp];
    const T* colp = img_base + s_sx[warp] * sw32 + lane * 8;
    roi_col_walk<T, 0, 0>(colp, sh32, sw32, nx, s_wx[warp], s_wy, s_ny, s_carry, res, y_first, y_end, outp, p.out.sh, 0u);
    return;
  }
  // sample loop (torchvision's own order of evaluation) for the bins of this column
#pragma unroll 1
  for (int ph = 0; ph < res; ++ph, outp += p.out.sh) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    roialign_sample_loop<T>(f, img, lane, ph, pw, roi_start_h, roi_start_w, bin_h, bin_w, grid_h, grid_w, acc);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv_count;
    Vec8<T>::store(outp, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// ROIAlign with the X PASS ON THE TENSOR CORES (variant 4, CM2_ROIALIGN_VARIANT=4; bf16 maps, res <= 16, c % 32 == 0).
// ncu of the column walk (variant 2): ~140 M warp instructions at cfg 5, two thirds of them the bf16 unpack + FMA of the
// x pass.  The x pass of one feature row of an ROI is a small GEMM,
//   T[pw, ch] = sum_X Wx[pw, X] * f[Y, X, ch]          (M = 16 padded bin columns, K = the ROI's pixel columns, N = channels),
// so it runs on the warp-level tensor-core path here (mma.sync.m16n8k16, bf16 x bf16 -> fp32); the SIMT pipes keep the y pass
// (out[ph] += Wy[ph][Y] * T on the accumulator fragments, with the same carried-row bookkeeping as roi_col_walk) and the
// stores.  The fp32 x weights are split into bf16 hi + lo (two MMAs into the same accumulator; the feature values are bf16
// already, so every product is exact in fp32 and the result differs from the SIMT walk by summation order only;
// SPLIT = false keeps hi alone: 2^-9 relative weight error, measured as an alternative, not used).
// One CTA per ROI (c / 32 warps), launched largest-first like the column kernel.  The tables (y weights / carry counts by
// warp 0, dense x weights [16][K] by the last warp) are built once per ROI; after that the warps never meet again: warp w
// owns channels [32 w, 32 w + 32) and stages ITS OWN 64-byte slice of every pixel of a row through a private ring of
// STAGES slots (cp.async 16 B, zero-filled past the last tap column; 80-byte pixel pitch = conflict-free ldmatrix.trans),
// so the loop needs cp.async.wait_group + __syncwarp only.  A ring item is (feature row, chunk of <= 32 pixel columns);
// the MMA accumulators run over the chunks of a row.  ROIs whose tap range exceeds ROI_MMA_KMAX columns or whose rows do not
// follow each other without a gap use the sample loop.  Non-finite feature values spread over the bin columns of their
// row chunk (0 * inf in the dense weight matrix) -- the SIMT variants confine them to the bins that tap them.
// MEASURED (B200, cfg 5: 32 images x 100 ROIs; profiles/r2_roialign_mma.txt, profiles/r2_ncu_roialign_mma_*.txt): 0.285 ms
// against 0.259 ms of the column walk, which stays the default -- and is the faster one in every box-size class (sides 32-120
// px: 0.132 vs 0.120 ms, 256-520 px: 0.61 vs 0.51 ms).  Why: the MMAs do 16 / ~2 times the useful multiply-adds (every bin
// column against every pixel column of a k tile), the staging adds 24 shared-memory wavefronts per k tile (8 copy + 8
// features + 8 weights; weights in registers when the tap range is one chunk) = 62 % of the LSU data path, and 48 fp32
// accumulator / carried-row registers per lane allow 16 warps per SM (the column walk: 24), each issuing an instruction
// every ~8 clocks (tensor pipe 38 % active, issue slots 49 %, no unit saturated: dependent-latency-bound).  Dropping the lo
// MMAs (SPLIT = false) buys 5 %, a fourth ring stage 2.5 %, register-resident parity buffers instead of the 16 moves per row
// and both k tiles' fragments loaded up front (124 registers) were 3 % SLOWER.  Kept as a tested alternative.
// ---------------------------------------------------------------------------------------------
constexpr int ROI_MMA_M = 16;                            // bin columns of the weight operand (res <= 16)
constexpr int ROI_MMA_KC = 32;                           // pixel columns per ring item (two k tiles)
constexpr int ROI_MMA_KMAX = 128;                        // widest tap range (pixel columns) the tensor-core walk handles
constexpr int ROI_MMA_APITCH = ROI_MMA_KMAX + 8;         // bf16 per weight row: 272 B, 8 rows hit 8 different bank groups
constexpr int ROI_MMA_PIX = 80;                          // bytes per staged pixel of one warp: 64 B of channels + 16 B pad
constexpr int ROI_MMA_SLOT = ROI_MMA_KC * ROI_MMA_PIX;   // bytes per ring slot of one warp

inline int roi_mma_smem_bytes(int c, int stages) {
  return std::max((c / 32) * stages * ROI_MMA_SLOT, ROI_MMA_M * ROI_MMA_KMAX * 4);     // the fp32 x table aliases the rings
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t saddr, const void* g, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(src_bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_n() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
// d += a (16 x 16, row) * b (16 x 8, col), fp32 accumulation
__device__ __forceinline__ void mma_bf16_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// d = a * b (the accumulator registers need no zeroing before the first k tile of a row)
__device__ __forceinline__ void mma_bf16_m16n8k16_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void roi_fma16(float (&acc)[4][4], float w, const float (&v)[4][4]) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    ffma2_roi(acc[nt][0], acc[nt][1], v[nt][0], v[nt][1], w);
    ffma2_roi(acc[nt][2], acc[nt][3], v[nt][2], v[nt][3], w);
  }
}

// SW: compile-time pixel stride of the feature maps in elements (0 = run-time); MINB: CTAs per SM the register budget allows
template <int STAGES, bool SPLIT, int SW, int MINB>
__global__ void __launch_bounds__(256, MINB) roialign_mma_kernel(const RoiAlignParams<__nv_bfloat16> p) {
  using T = __nv_bfloat16;
  extern __shared__ __align__(128) unsigned char s_mma[];
  __shared__ __align__(16) T s_a[2][ROI_MMA_M][ROI_MMA_APITCH];              // x weights, bf16 hi / lo, [bin column][pixel column]
  __shared__ float s_wy[ROI_MMA_M * ROI_WYT];
  __shared__ int s_ny[ROI_MMA_M], s_carry[ROI_MMA_M];
  __shared__ int s_info[6];                             // y_first, y_end, rows walkable, first tap column, tap columns, x bad
  const int res = p.res, c8 = p.out.c >> 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int slot = p.order[blockIdx.x];
  const int img = slot / p.r_cap, r = slot - img * p.r_cap;
  auto fill_zero = [&]() {
    const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < res * res * c8; i += blockDim.x) {
      const int bin = i / c8, cv = i - bin * c8;
      const int ph = bin / res, pw = bin - ph * res;
      Vec8<T>::store(p.out.at(slot, ph, pw) + cv * 8, z);
    }
  };
  if (r >= p.det_count[img]) {                          // empty slot: zeros (CTA-uniform)
    fill_zero();
    return;
  }
  const float4 bx = reinterpret_cast<const float4*>(p.boxes)[slot];
  int lvl = assign_level(bx.x, bx.y, bx.z, bx.w, p.image_area[img], p.crit, p.min_level, p.max_level);
  if (lvl >= p.num_levels) lvl = p.num_levels - 1;
  if (p.level_out && tid == 0) p.level_out[slot] = lvl;
  View<const T> f = p.feat[0];
  float scale = p.scale[0];
#pragma unroll
  for (int l = 1; l < ROI_MAX_LEVELS; ++l)
    if (l == lvl) { f = p.feat[l]; scale = p.scale[l]; }
  const float roi_start_w = bx.x * scale - 0.5f, roi_start_h = bx.y * scale - 0.5f;
  const float roi_end_w = bx.z * scale - 0.5f, roi_end_h = bx.w * scale - 0.5f;
  const float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
  const float bin_h = roi_height / (float)res, bin_w = roi_width / (float)res;
  const int grid_h = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_height / (float)res);
  const int grid_w = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(roi_width / (float)res);
  const float inv_count = 1.0f / fmaxf((float)(grid_h * grid_w), 1.f);
  const T* img_base = f.p + (size_t)img * f.sn;
  const unsigned full = 0xffffffffu;

  // ---- tables
  float* s_wxf = reinterpret_cast<float*>(s_mma);       // [ROI_MMA_M][ROI_MMA_KMAX] fp32 x weights (the rings come later)
  for (int i = tid; i < ROI_MMA_M * ROI_MMA_KMAX; i += blockDim.x) s_wxf[i] = 0.f;
  for (int i = tid; i < ROI_MMA_M * ROI_WYT; i += blockDim.x) s_wy[i] = 0.f;
  __syncthreads();
  if (warp == 0) {                                      // y weights (lane = bin row) and the carry bookkeeping of the walk
    int t0 = 0, tn = 0;
    bool bad = false;
    if (lane < res) {
      float* wrow = s_wy + lane * ROI_WYT;
      int s0 = -1;
      for (int i = 0; i < grid_h; ++i) {
        const float v = roi_start_h + lane * bin_h + ((float)i + 0.5f) * bin_h / (float)grid_h;
        if (v < -1.0f || v > (float)f.h) continue;
        float vv = v <= 0.f ? 0.f : v;
        int low = (int)vv, high;
        if (low >= f.h - 1) { high = low = f.h - 1; vv = (float)low; } else high = low + 1;
        const float l = vv - (float)low, h = 1.f - l;
        if (s0 < 0) s0 = low;
        if (high - s0 >= ROI_WYT) { bad = true; break; }
        wrow[low - s0] += h * inv_count;                // 1 / count folded into the y weights
        wrow[high - s0] += l * inv_count;
        tn = high - s0 + 1;
      }
      t0 = s0 < 0 ? 0 : s0;
    }
    // rows of the bins above end at y_done; rows y_done - 1, y_done are carried in (cp, cl)
    const bool ybin = lane < res && tn > 0;
    int pm = ybin ? t0 + tn - 1 : -1;                   // inclusive prefix max of the last row of a bin
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(full, pm, o);
      if (lane >= o) pm = max(pm, t);
    }
    int y_done = __shfl_up_sync(full, pm, 1);
    if (lane == 0) y_done = -1;
    int y_first = ybin ? t0 : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) y_first = min(y_first, __shfl_xor_sync(full, y_first, o));
    int y_end = __shfl_sync(full, pm, 31) + 1;
    if (y_first == 0x7fffffff) { y_first = 0; y_end = 0; }
    int cy = 0;
    if (ybin && y_done >= 0) {
      cy = y_done - t0 + 1;
      if (cy < 0 || cy > min(2, y_done - y_first + 1)) bad = true;      // gap, or a row that is not carried any more
    }
    const bool walk_y = !__any_sync(full, bad);
    if (lane < ROI_MMA_M) { s_ny[lane] = lane < res ? tn : 0; s_carry[lane] = cy; }
    if (lane == 0) { s_info[0] = y_first; s_info[1] = y_end; s_info[2] = walk_y ? 1 : 0; }
  }
  if (warp == nwarps - 1) {                             // x weights: lane = bin column, dense rows over the ROI's tap columns
    int xlo = 0x7fffffff, xhi = -1;
    if (lane < res) {
      for (int i = 0; i < grid_w; ++i) {                // pass 1: the tap range (taps are monotone in i)
        const float v = roi_start_w + lane * bin_w + ((float)i + 0.5f) * bin_w / (float)grid_w;
        if (v < -1.0f || v > (float)f.w) continue;
        const float vv = v <= 0.f ? 0.f : v;
        int low = (int)vv, high;
        if (low >= f.w - 1) high = low = f.w - 1; else high = low + 1;
        xlo = min(xlo, low);
        xhi = max(xhi, high);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      xlo = min(xlo, __shfl_xor_sync(full, xlo, o));
      xhi = max(xhi, __shfl_xor_sync(full, xhi, o));
    }
    const int kx = xhi >= 0 ? xhi - xlo + 1 : 0;
    const bool xbad = kx > ROI_MMA_KMAX;
    if (!xbad && kx > 0 && lane < res) {
      float* wrow = s_wxf + lane * ROI_MMA_KMAX;
      for (int i = 0; i < grid_w; ++i) {                // pass 2: the weights
        const float v = roi_start_w + lane * bin_w + ((float)i + 0.5f) * bin_w / (float)grid_w;
        if (v < -1.0f || v > (float)f.w) continue;
        float vv = v <= 0.f ? 0.f : v;
        int low = (int)vv, high;
        if (low >= f.w - 1) { high = low = f.w - 1; vv = (float)low; } else high = low + 1;
        const float l = vv - (float)low, h = 1.f - l;
        wrow[low - xlo] += h;
        wrow[high - xlo] += l;
      }
    }
    if (lane == 0) { s_info[3] = kx > 0 ? xlo : 0; s_info[4] = kx; s_info[5] = xbad ? 1 : 0; }
  }
  __syncthreads();
  const int y_first = s_info[0], nrows = s_info[1] - s_info[0], x_first = s_info[3], kx = s_info[4];
  const bool walk = s_info[2] != 0 && s_info[5] == 0;
  if (walk && (nrows <= 0 || kx <= 0)) {                // nothing inside the feature map: zeros
    fill_zero();
    return;
  }
  if (!walk) {                                          // sample loop (torchvision's own order of evaluation), a warp per bin
#pragma unroll 1
    for (int bin = warp; bin < res * res; bin += nwarps) {
      const int ph = bin / res, pw = bin - ph * res;
#pragma unroll 1
      for (int cv = lane; cv < c8; cv += 32) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        roialign_sample_loop<T>(f, img, cv, ph, pw, roi_start_h, roi_start_w, bin_h, bin_w, grid_h, grid_w, acc);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= inv_count;
        Vec8<T>::store(p.out.at(slot, ph, pw) + cv * 8, acc);
      }
    }
    return;
  }
  const int nchunk = (kx + ROI_MMA_KC - 1) / ROI_MMA_KC, kpad = nchunk * ROI_MMA_KC;
  for (int i = tid; i < ROI_MMA_M * kpad; i += blockDim.x) {                  // fp32 -> bf16 hi + lo
    const int row = i / kpad, k = i - row * kpad;
    const float w = s_wxf[row * ROI_MMA_KMAX + k];
    const T hi = __float2bfloat16_rn(w);
    s_a[0][row][k] = hi;
    s_a[1][row][k] = __float2bfloat16_rn(w - __bfloat162float(hi));
  }
  __syncthreads();                                      // the fp32 table is dead: its memory becomes the rings

  // ---- the walk of this warp's 32 channels
  const uint32_t wring = (uint32_t)__cvta_generic_to_shared(s_mma) + warp * (STAGES * ROI_MMA_SLOT);
  const int mat = lane >> 3, r8 = lane & 7;
  // ldmatrix row addresses: A (weights) matrices = (rows 0-7 | 8-15) x (k 0-7 | 8-15); B (pixels, transposed on load)
  // matrices = (pixels 0-7 | 8-15) x (channels 0-7 | 8-15) of a 16-channel pair of n tiles
  const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(&s_a[0][0][0]) + ((r8 + (mat & 1) * 8) * ROI_MMA_APITCH + (mat >> 1) * 8) * 2;
  constexpr uint32_t A_LO = ROI_MMA_M * ROI_MMA_APITCH * 2;
  const uint32_t b_lane = wring + (r8 + (mat & 1) * 8) * ROI_MMA_PIX + (mat >> 1) * 16;
  // copy role of a lane: pixel cpx (+ 8, 16, 24) of the chunk, 16-byte quarter cq of the warp's 64 bytes per pixel -- the 8
  // lanes of a quarter warp write 8 different pixels (80-byte pitch: 8 different bank groups, like the ldmatrix rows).
  // Element offsets inside an image fit 32 bits (the host checks sn < 2^31).
  const int cpx = lane & 7, cq = lane >> 3;
  const int sw32 = SW ? SW : (int)f.sw, sh32 = (int)f.sh;
  const T* iss_src = img_base + (y_first * sh32 + (x_first + cpx) * sw32 + warp * 32 + cq * 8);   // this lane's pixel of the next item
  const uint32_t lane_dst = wring + cpx * ROI_MMA_PIX + cq * 16;
  int iss_left = nrows * nchunk;
  uint32_t iss_slot = 0, cons_slot = 0;
  const int g = lane >> 2, t4 = lane & 3;
  T* outp = p.out.at(slot, 0, 0) + warp * 32 + t4 * 8;
  const int out_sw = (int)p.out.sw;

  float cl[4][4], cp[4][4];                             // x-passed rows y_done (cl) and y_done - 1 (cp): [n tile][fragment]
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int k = 0; k < 4; ++k) cl[nt][k] = cp[nt][k] = 0.f;
  // cl (+)= Wx[:, 16 columns] * f[row, 16 columns, 32 channels]; FRESH: the first k tile of a row overwrites cl
  auto ktile = [&](const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t baddr, auto fresh) {
    constexpr bool FRESH = decltype(fresh)::value;
    uint32_t b0[4], b1[4];
    ldsm_x4_trans(baddr, b0);
    ldsm_x4_trans(baddr + 32, b1);
    if (FRESH) {
      mma_bf16_m16n8k16_zero(cl[0], ahi, b0[0], b0[1]);
      mma_bf16_m16n8k16_zero(cl[1], ahi, b0[2], b0[3]);
      mma_bf16_m16n8k16_zero(cl[2], ahi, b1[0], b1[1]);
      mma_bf16_m16n8k16_zero(cl[3], ahi, b1[2], b1[3]);
    } else {
      mma_bf16_m16n8k16(cl[0], ahi, b0[0], b0[1]);
      mma_bf16_m16n8k16(cl[1], ahi, b0[2], b0[3]);
      mma_bf16_m16n8k16(cl[2], ahi, b1[0], b1[1]);
      mma_bf16_m16n8k16(cl[3], ahi, b1[2], b1[3]);
    }
    if (SPLIT) {
      mma_bf16_m16n8k16(cl[0], alo, b0[0], b0[1]);
      mma_bf16_m16n8k16(cl[1], alo, b0[2], b0[3]);
      mma_bf16_m16n8k16(cl[2], alo, b1[0], b1[1]);
      mma_bf16_m16n8k16(cl[3], alo, b1[2], b1[3]);
    }
  };
  auto ktile_smem = [&](uint32_t aaddr, uint32_t baddr, auto fresh) {       // weights fetched per k tile
    uint32_t ahi[4], alo[4] = {0u, 0u, 0u, 0u};
    ldsm_x4(aaddr, ahi);
    if (SPLIT) ldsm_x4(aaddr + A_LO, alo);
    ktile(ahi, alo, baddr, fresh);
  };
  // 4 x 4 transpose of one 32-bit word per (lane of a quad, n tile): afterwards lane t4 holds the 8 consecutive channels
  // 8 t4 .. 8 t4 + 7 of a bin column as four bf16 pairs -> one 16-byte store per lane instead of four 4-byte stores
  auto quad_transpose = [&](uint32_t (&v)[4]) {
    const bool b1 = (t4 & 2) != 0, b0 = (t4 & 1) != 0;
    uint32_t x = b1 ? v[0] : v[2], y = b1 ? v[1] : v[3];
    x = __shfl_xor_sync(full, x, 2);
    y = __shfl_xor_sync(full, y, 2);
    if (b1) { v[0] = x; v[1] = y; } else { v[2] = x; v[3] = y; }
    x = b0 ? v[0] : v[1];
    y = b0 ? v[2] : v[3];
    x = __shfl_xor_sync(full, x, 1);
    y = __shfl_xor_sync(full, y, 1);
    if (b0) { v[0] = x; v[2] = y; } else { v[1] = x; v[3] = y; }
  };
  auto pack2 = [](float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  };

  // SINGLE: the tap range fits one chunk (<= 32 columns, the common case): an item is a whole feature row, the weight
  // fragments of its one or two k tiles stay in registers for the whole ROI and the copy predicates are ROI constants
  auto run = [&](auto single) {
    constexpr bool SINGLE = decltype(single)::value;
    // ---- SINGLE state
    const int rem0 = kx - cpx;                          // pixel cpx + 8 i is a tap column iff rem0 > 8 i
    const bool two = kx > 16;                           // second k tile
    uint32_t w0hi[4] = {0u, 0u, 0u, 0u}, w0lo[4] = {0u, 0u, 0u, 0u}, w1hi[4] = {0u, 0u, 0u, 0u}, w1lo[4] = {0u, 0u, 0u, 0u};
    if (SINGLE) {
      ldsm_x4(a_hi, w0hi);
      if (SPLIT) ldsm_x4(a_hi + A_LO, w0lo);
      if (two) {
        ldsm_x4(a_hi + 32, w1hi);
        if (SPLIT) ldsm_x4(a_hi + 32 + A_LO, w1lo);
      }
    }
    // ---- chunked state
    const int row_wrap = sh32 - (nchunk - 1) * ROI_MMA_KC * sw32;    // from the last chunk of a row to the first of the next
    int iss_rem = rem0;
    // whole k tiles are written; columns past the last tap are zero-filled (src-size 0: nothing is read)
    auto issue_next = [&]() {
      if (iss_left > 0) {
        const uint32_t dst = lane_dst + iss_slot;
        if (SINGLE) {
          cp_async16_zfill(dst, iss_src, rem0 > 0 ? 16u : 0u);
          cp_async16_zfill(dst + 8 * ROI_MMA_PIX, iss_src + 8 * sw32, rem0 > 8 ? 16u : 0u);
          if (two) {
            cp_async16_zfill(dst + 16 * ROI_MMA_PIX, iss_src + 16 * sw32, rem0 > 16 ? 16u : 0u);
            cp_async16_zfill(dst + 24 * ROI_MMA_PIX, iss_src + 24 * sw32, rem0 > 24 ? 16u : 0u);
          }
          iss_src += sh32;
        } else {
          cp_async16_zfill(dst, iss_src, iss_rem > 0 ? 16u : 0u);
          cp_async16_zfill(dst + 8 * ROI_MMA_PIX, iss_src + 8 * sw32, iss_rem > 8 ? 16u : 0u);
          if (iss_rem + cpx > 16) {
            cp_async16_zfill(dst + 16 * ROI_MMA_PIX, iss_src + 16 * sw32, iss_rem > 16 ? 16u : 0u);
            cp_async16_zfill(dst + 24 * ROI_MMA_PIX, iss_src + 24 * sw32, iss_rem > 24 ? 16u : 0u);
          }
          iss_rem -= ROI_MMA_KC;
          if (iss_rem + cpx <= 0) { iss_rem = rem0; iss_src += row_wrap; } else iss_src += ROI_MMA_KC * sw32;
        }
        --iss_left;
        iss_slot += ROI_MMA_SLOT;
        if (iss_slot == STAGES * ROI_MMA_SLOT) iss_slot = 0;
      }
      cp_async_commit();                                // one group per step, empty or not: the wait count stays fixed
    };
#pragma unroll 1
    for (int d = 0; d < STAGES - 1; ++d) issue_next();

#pragma unroll 1
    for (int ph = 0; ph < res; ++ph, outp += p.out.sh) {
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[nt][k] = 0.f;
      const int ny = s_ny[ph];
      const float* wy = s_wy + ph * ROI_WYT;
      const int cy = s_carry[ph];                       // leading rows of this bin that are x-passed already
      int j = 0;
      if (cy == 2) {
        roi_fma16(acc, wy[0], cp);
        if (ny > 1) roi_fma16(acc, wy[1], cl);
        j = 2;
      } else if (cy == 1) {
        roi_fma16(acc, wy[0], cl);
        j = 1;
      }
#pragma unroll 1
      for (; j < ny; ++j) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int k = 0; k < 4; ++k) cp[nt][k] = cl[nt][k];
        if (SINGLE) {
          cp_async_wait_n<STAGES - 2>();                // this lane's copies of the row have landed ...
          __syncwarp();                                 // ... and so have the other lanes'; the slot of the previous row is free
          issue_next();
          const uint32_t baddr = b_lane + cons_slot;
          ktile(w0hi, w0lo, baddr, std::true_type());
          if (two) ktile(w1hi, w1lo, baddr + 16 * ROI_MMA_PIX, std::false_type());
          cons_slot += ROI_MMA_SLOT;
          if (cons_slot == STAGES * ROI_MMA_SLOT) cons_slot = 0;
        } else {
          int kb = 0;
#pragma unroll 1
          do {                                          // the chunks of one feature row
            cp_async_wait_n<STAGES - 2>();
            __syncwarp();
            issue_next();
            const uint32_t aaddr = a_hi + kb * 2, baddr = b_lane + cons_slot;
            if (kb == 0) ktile_smem(aaddr, baddr, std::true_type()); else ktile_smem(aaddr, baddr, std::false_type());
            if (kx - kb > 16) ktile_smem(aaddr + 32, baddr + 16 * ROI_MMA_PIX, std::false_type());
            cons_slot += ROI_MMA_SLOT;
            if (cons_slot == STAGES * ROI_MMA_SLOT) cons_slot = 0;
            kb += ROI_MMA_KC;
          } while (kb < kx);
        }
        roi_fma16(acc, wy[j], cl);
      }
      // fragment rows = bin columns g and g + 8, fragment columns = channels 2 t4, 2 t4 + 1 of each n tile
      uint32_t v[4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) v[nt] = pack2(acc[nt][0], acc[nt][1]);
      quad_transpose(v);
      if (g < res) *reinterpret_cast<uint4*>(outp + g * out_sw) = make_uint4(v[0], v[1], v[2], v[3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) v[nt] = pack2(acc[nt][2], acc[nt][3]);
      quad_transpose(v);
      if (g + 8 < res) *reinterpret_cast<uint4*>(outp + (g + 8) * out_sw) = make_uint4(v[0], v[1], v[2], v[3]);
    }
  };
  if (nchunk == 1) run(std::true_type()); else run(std::false_type());
  cp_async_wait_n<0>();
}

// ---------------------------------------------------------------------------------------------
// predictor restricted to the predicted class + sigmoid (sam.py:97, mask_head.py:196-216).
// warp per output pixel; lanes stride over channels.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void mask_predict_kernel(View<const T> x, const float* __restrict__ wp, const float* __restrict__ bp,
                                    const int64_t* __restrict__ classes, int ncls, float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int px = x.h * x.w, c8 = x.c >> 3;
  const int64_t total = (int64_t)x.n * px;
  for (int64_t i = warp; i < total; i += nwarps) {
    int r = (int)(i / px), pix = (int)(i - (int64_t)r * px);
    int y = pix / x.w, xx = pix - y * x.w;
    int cls = ncls == 1 ? 0 : (int)classes[r];
    cls = min(max(cls, 0), ncls - 1);
    const float* wrow = wp + (size_t)cls * x.c;
    const T* row = x.at(r, y, xx);
    float acc = 0.f;
    for (int cv = lane; cv < c8; cv += 32) {
      float v[8];
      Vec8<T>::load(row + cv * 8, v);
      float4 wa = __ldg(reinterpret_cast<const float4*>(wrow + cv * 8));
      float4 wb = __ldg(reinterpret_cast<const float4*>(wrow + cv * 8) + 1);
      acc = fmaf(v[0], wa.x, acc); acc = fmaf(v[1], wa.y, acc); acc = fmaf(v[2], wa.z, acc); acc = fmaf(v[3], wa.w, acc);
      acc = fmaf(v[4], wb.x, acc); acc = fmaf(v[5], wb.y, acc); acc = fmaf(v[6], wb.z, acc); acc = fmaf(v[7], wb.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) probs[i] = sigmoid_f32(acc + bp[cls]);
  }
}

// 2x2 max pool of probs [r, 2s, 2s] into channel 0 of out [r, s, s, cpad] (other channels zero)
template <typename T>
__global__ void maskiou_prep_kernel(const float* __restrict__ probs, View<T> out) {
  const int S = out.h;
  int64_t total = (int64_t)out.n * S * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / (S * S)), pix = (int)(i - (int64_t)r * S * S);
    int y = pix / S, x = pix - y * S;
    const float* src = probs + ((size_t)r * 2 * S + 2 * y) * 2 * S + 2 * x;
    float m = fmaxf(fmaxf(src[0], src[1]), fmaxf(src[2 * S], src[2 * S + 1]));
    T* q = out.at(r, y, x);
    q[0] = from_f32<T>(m);
    for (int c = 1; c < out.c; ++c) q[c] = from_f32<T>(0.f);
  }
}

template <typename T>
__global__ void maskiou_score_kernel(const T* __restrict__ iou, int r, int ncls, const int64_t* __restrict__ classes,
                                     const float* __restrict__ scores, float* __restrict__ mask_scores) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r) return;
  int cls = min(max((int)classes[i], 0), ncls - 1);
  mask_scores[i] = scores[i] * to_f32<T>(iou[(size_t)i * ncls + cls]);
}

// ---------------------------------------------------------------------------------------------
// detector_postprocess: box rescale/clip and mask paste-back
// ---------------------------------------------------------------------------------------------
__global__ void scale_clip_boxes_kernel(const float* __restrict__ in, float* __restrict__ out, uint8_t* __restrict__ valid,
                                        int r, float sx, float sy, float out_w, float out_h) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r) return;
  float4 b = reinterpret_cast<const float4*>(in)[i];
  b.x = fminf(fmaxf(b.x * sx, 0.f), out_w);
  b.z = fminf(fmaxf(b.z * sx, 0.f), out_w);
  b.y = fminf(fmaxf(b.y * sy, 0.f), out_h);
  b.w = fminf(fmaxf(b.w * sy, 0.f), out_h);
  reinterpret_cast<float4*>(out)[i] = b;
  valid[i] = ((b.z - b.x) > 0.f && (b.w - b.y) > 0.f) ? 1 : 0;
}

// Batched variant: params[img] = (sx, sy, out_w, out_h) in device memory, r_cap slots per image.
__global__ void scale_clip_boxes_batch_kernel(const float* __restrict__ in, float* __restrict__ out, uint8_t* __restrict__ valid,
                                              int total, int r_cap, const float* __restrict__ params,
                                              const int* __restrict__ det_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (det_count && i % r_cap >= det_count[i / r_cap]) {       // empty slot: no box, no mask
    reinterpret_cast<float4*>(out)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    valid[i] = 0;
    return;
  }
  const float4 q = __ldg(reinterpret_cast<const float4*>(params) + i / r_cap);
  float4 b = reinterpret_cast<const float4*>(in)[i];
  b.x = fminf(fmaxf(b.x * q.x, 0.f), q.z);
  b.z = fminf(fmaxf(b.z * q.x, 0.f), q.z);
  b.y = fminf(fmaxf(b.y * q.y, 0.f), q.w);
  b.w = fminf(fmaxf(b.w * q.y, 0.f), q.w);
  reinterpret_cast<float4*>(out)[i] = b;
  valid[i] = ((b.z - b.x) > 0.f && (b.w - b.y) > 0.f) ? 1 : 0;
}

// Fixed-size result record of every detection slot (centermask2_b200/parallel.py: RECORD_FIELDS = 11):
// box[4], score, class, mask score, location[2], valid, detections of the image.  One thread per slot.
__global__ void pack_records_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const long long* __restrict__ classes,
                                    const float* __restrict__ mask_scores, const float* __restrict__ locations,
                                    const uint8_t* __restrict__ valid, const int* __restrict__ count, int total, int r_cap,
                                    float* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int img = i / r_cap, k = i - img * r_cap, n = count[img];
  float* o = rec + (size_t)i * 11;
  const bool live = k < n;
  const float4 b = live ? reinterpret_cast<const float4*>(boxes)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w;
  o[4] = live ? scores[i] : 0.f;
  o[5] = live ? (float)classes[i] : 0.f;
  o[6] = live && mask_scores ? mask_scores[i] : 0.f;
  o[7] = live && locations ? locations[2 * i] : 0.f;
  o[8] = live && locations ? locations[2 * i + 1] : 0.f;
  o[9] = live && (!valid || valid[i]) ? 1.f : 0.f;
  o[10] = (float)n;
}

// Mask paste-back in two steps: (1) the whole [r, out_h, out_w] buffer is cleared with one memset at store bandwidth
// (most bytes lie outside their box); (2) paste_window_kernel visits only the dilated box window
// [floor(x0)-1, ceil(x1)+1) x [floor(y0)-1, ceil(y1)+1) of every valid ROI.  grid (PASTE_CTAS_PER_ROI, r): a CTA keeps
// the 28x28 probabilities in shared memory with a zero border (index clamping replaces the four bounds tests of
// grid_sample's zero padding) and walks the window in tiles of 8 rows x 256 columns; a thread owns one column of the
// tile, so the x interpolation coefficients are computed once per 8 pixels.  Rows of an 800x1333 mask are not even
// 2-byte aligned, so the window is written with byte stores (32 consecutive bytes per warp instruction).
// grid_sample(bilinear, zeros, align_corners=False): ix = ((gx + 1) * m - 1) / 2.
constexpr int PASTE_CTAS_PER_ROI = 8;
constexpr int PASTE_MAX_W = 4096;                      // widest output the column table supports (else byte kernel)

// Word variant (out_h * out_w % 4 == 0 and a 4-byte aligned buffer, so every ROI plane starts on a word): a warp takes
// one window row at a time and its lanes own consecutive aligned 32-bit words of the flat plane, i.e. 128 pixels per
// warp instruction and full-sector stores (byte stores into lines that are not L2-resident cost a DRAM fill each).
// Bytes of a boundary word that fall outside the window are zeros by definition, so whole words can be written.  The
// x interpolation (tap columns + weights) is tabulated once per CTA in shared memory.
struct __align__(16) PasteCol { float wx0, wx1; int cl, ch; };

__global__ void __launch_bounds__(256) paste_window_words_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                                 const uint8_t* __restrict__ valid, uint8_t* __restrict__ out,
                                                                 int m, int out_h, int out_w, float threshold) {
  extern __shared__ float s_dyn[];                     // (m + 2)^2 mask with zero border, then the column table
  const int r = blockIdx.y;
  if (valid[r] == 0) return;
  const float4 b = reinterpret_cast<const float4*>(boxes)[r];
  const int xa = max((int)floorf(b.x) - 1, 0), ya = max((int)floorf(b.y) - 1, 0);
  const int xb = min((int)ceilf(b.z) + 1, out_w), yb = min((int)ceilf(b.w) + 1, out_h);
  if (xa >= xb || ya >= yb) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if ((int)blockIdx.x * nwarps >= yb - ya) return;
  const int mp = m + 2;
  float* s_mask = s_dyn;
  PasteCol* s_col = reinterpret_cast<PasteCol*>(s_dyn + ((mp * mp + 3) & ~3));
  for (int i = threadIdx.x; i < mp * mp; i += blockDim.x) {
    const int yy = i / mp - 1, xx = i - (yy + 1) * mp - 1;
    s_mask[i] = (yy >= 0 && yy < m && xx >= 0 && xx < m) ? probs[(size_t)r * m * m + yy * m + xx] : 0.f;
  }
  const float bw = b.z - b.x, bh = b.w - b.y, fm = (float)m;
  for (int x = xa + threadIdx.x; x < xb; x += blockDim.x) {
    const float gx = ((float)x + 0.5f - b.x) / bw * 2.f - 1.f;
    const float ix = ((gx + 1.f) * fm - 1.f) * 0.5f;
    const float fx = floorf(ix);
    // float -> int of an out-of-range value is clamped first; both taps of a far-outside column land on the zero border
    const int xl = (int)fminf(fmaxf(fx, -2.f), fm + 1.f);
    PasteCol c;
    c.wx1 = ix - fx; c.wx0 = (fx + 1.f) - ix;
    c.cl = min(max(xl + 1, 0), m + 1); c.ch = min(max(xl + 2, 0), m + 1);
    s_col[x - xa] = c;
  }
  __syncthreads();
  uint32_t* obase = reinterpret_cast<uint32_t*>(out + (size_t)r * out_h * out_w);
  for (int y = ya + blockIdx.x * nwarps + warp; y < yb; y += gridDim.x * nwarps) {
    const float gy = ((float)y + 0.5f - b.y) / bh * 2.f - 1.f;
    const float iy = ((gy + 1.f) * fm - 1.f) * 0.5f;
    const float fy = floorf(iy);
    const float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;        // ATen: (iy_se - iy), (iy - iy_nw)
    const int yl = (int)fminf(fmaxf(fy, -2.f), fm + 1.f);
    const float* r0 = s_mask + min(max(yl + 1, 0), m + 1) * mp;
    const float* r1 = s_mask + min(max(yl + 2, 0), m + 1) * mp;
    const long long row0 = (long long)y * out_w;
    const long long w_first = (row0 + xa) >> 2, w_last = (row0 + xb - 1) >> 2;
    // 32 words = 128 consecutive pixels per warp pass: in step k lane l evaluates pixel 32k + l (consecutive lanes read
    // consecutive 16-byte table entries: conflict-free), a ballot collects the 32 decisions, and lane l finally stores
    // word l assembled from ballot l / 8 -- one coalesced 128-byte store per pass.
    for (long long w0 = w_first; w0 <= w_last; w0 += 32) {
      uint32_t bits[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int x = (int)(w0 * 4 + 32 * k + lane - row0), yy = y;
        // a boundary word can reach into the neighbouring row (out_w is not a multiple of 4); its bytes there are
        // evaluated with that row's coefficients so that both rows write identical words
        if (x < 0) { x += out_w; yy = y - 1; } else if (x >= out_w) { x -= out_w; yy = y + 1; }
        bool on = false;
        if (x >= xa && x < xb && yy >= ya && yy < yb) {
          const PasteCol c = s_col[x - xa];
          float a0 = wy0, a1 = wy1;
          const float* q0 = r0;
          const float* q1 = r1;
          if (yy != y) {                                    // rare
            const float gy2 = ((float)yy + 0.5f - b.y) / bh * 2.f - 1.f;
            const float iy2 = ((gy2 + 1.f) * fm - 1.f) * 0.5f;
            const float fy2 = floorf(iy2);
            a1 = iy2 - fy2; a0 = (fy2 + 1.f) - iy2;
            const int yl2 = (int)fminf(fmaxf(fy2, -2.f), fm + 1.f);
            q0 = s_mask + min(max(yl2 + 1, 0), m + 1) * mp;
            q1 = s_mask + min(max(yl2 + 2, 0), m + 1) * mp;
          }
          // same accumulation order as ATen's grid_sampler_2d (nw, ne, sw, se); border cells contribute exact zeros
          float v = 0.f;
          v += q0[c.cl] * (c.wx0 * a0);
          v += q0[c.ch] * (c.wx1 * a0);
          v += q1[c.cl] * (c.wx0 * a1);
          v += q1[c.ch] * (c.wx1 * a1);
          on = v >= threshold;
        }
        bits[k] = __ballot_sync(0xffffffffu, on);
      }
      const uint32_t sel = lane < 8 ? bits[0] : (lane < 16 ? bits[1] : (lane < 24 ? bits[2] : bits[3]));
      const uint32_t nib = (sel >> (4 * (lane & 7))) & 0xfu;
      const uint32_t word = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
      if (w0 + lane <= w_last) obase[w0 + lane] = word;
    }
  }
}

// Byte-store variant for buffers whose ROI planes are not word aligned: a thread owns one column of an 8-row tile.
constexpr int PASTE_TILE_ROWS = 8;
__global__ void __launch_bounds__(256) paste_window_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                           const uint8_t* __restrict__ valid, uint8_t* __restrict__ out,
                                                           int m, int out_h, int out_w, float threshold) {
  extern __shared__ float s_mask[];                    // (m + 2) x (m + 2), zero border
  const int r = blockIdx.y;
  if (valid[r] == 0) return;
  const float4 b = reinterpret_cast<const float4*>(boxes)[r];
  const int xa = max((int)floorf(b.x) - 1, 0), ya = max((int)floorf(b.y) - 1, 0);
  const int xb = min((int)ceilf(b.z) + 1, out_w), yb = min((int)ceilf(b.w) + 1, out_h);
  if (xa >= xb || ya >= yb) return;
  const int ntx = (xb - xa + 255) >> 8, nty = (yb - ya + PASTE_TILE_ROWS - 1) / PASTE_TILE_ROWS;
  if ((int)blockIdx.x >= ntx * nty) return;
  const int mp = m + 2;
  for (int i = threadIdx.x; i < mp * mp; i += blockDim.x) {
    const int yy = i / mp - 1, xx = i - (yy + 1) * mp - 1;
    s_mask[i] = (yy >= 0 && yy < m && xx >= 0 && xx < m) ? probs[(size_t)r * m * m + yy * m + xx] : 0.f;
  }
  __syncthreads();
  const float bw = b.z - b.x, bh = b.w - b.y, fm = (float)m;
  uint8_t* obase = out + (size_t)r * out_h * out_w;
  for (int t = blockIdx.x; t < ntx * nty; t += gridDim.x) {
    const int ty = t / ntx, tx = t - ty * ntx;
    const int x = xa + tx * 256 + threadIdx.x;
    if (x >= xb) continue;
    const float gx = ((float)x + 0.5f - b.x) / bw * 2.f - 1.f;
    const float ix = ((gx + 1.f) * fm - 1.f) * 0.5f;
    const float fx = floorf(ix);
    const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
    const int xl = (int)fminf(fmaxf(fx, -2.f), fm + 1.f);
    const int cl = min(max(xl + 1, 0), m + 1), ch = min(max(xl + 2, 0), m + 1);
    const int y_end = min(ya + (ty + 1) * PASTE_TILE_ROWS, yb);
    for (int y = ya + ty * PASTE_TILE_ROWS; y < y_end; ++y) {
      const float gy = ((float)y + 0.5f - b.y) / bh * 2.f - 1.f;
      const float iy = ((gy + 1.f) * fm - 1.f) * 0.5f;
      const float fy = floorf(iy);
      const float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;      // ATen: (iy_se - iy), (iy - iy_nw)
      const int yl = (int)fminf(fmaxf(fy, -2.f), fm + 1.f);
      const float* r0 = s_mask + min(max(yl + 1, 0), m + 1) * mp;
      const float* r1 = s_mask + min(max(yl + 2, 0), m + 1) * mp;
      float v = 0.f;
      v += r0[cl] * (wx0 * wy0);
      v += r0[ch] * (wx1 * wy0);
      v += r1[cl] * (wx0 * wy1);
      v += r1[ch] * (wx1 * wy1);
      obase[(size_t)y * out_w + x] = v >= threshold ? 1 : 0;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Fused paste (out_h * out_w % 16 == 0, 16-byte aligned buffer): every byte of the [r, out_h, out_w] buffer is written
// exactly once with 16-byte stores, no memset.  grid (slabs, r): a CTA owns a contiguous run of 16-byte chunks of one
// ROI plane (about 48 image rows).
//   phase Z  chunks that do not touch the dilated box window are stored as zeros (a slab that misses the window rows
//            does nothing else);
//   phase C  the window part of the slab is evaluated with FIXED COLUMNS per lane: a warp owns a strip of 32 columns,
//            the x interpolation (tap column + two weights) lives in registers for the whole strip, the y
//            interpolation comes from a per-row table and the four taps of a pixel are ONE 16-byte shared-memory load
//            from a pre-gathered table M4[yl][xl] = (m[yl][xl], m[yl][xl+1], m[yl+1][xl], m[yl+1][xl+1]) with a zero
//            border of two cells (grid_sample's zero padding without bounds tests).  A ballot per (row, strip) packs
//            the 32 decisions into one word of a bit tile in shared memory;
//   phase W  the chunks that touch the window gather their 16 bits from the bit tile (funnel shift over two words,
//            rows of an 800x1333 plane are not aligned to anything), expand them to 16 bytes and store.
// The arithmetic (operand order included) is that of paste_window_words_kernel, i.e. ATen's grid_sampler_2d.
// ---------------------------------------------------------------------------------------------
struct __align__(16) PasteRow { float wy0, wy1; int base, pad; };

__device__ __forceinline__ uint32_t spread4(uint32_t nib) {          // bits 0..3 -> bytes 0..3 (0 / 1)
  return (nib * 0x00204081u) & 0x01010101u;
}

__global__ void __launch_bounds__(256) paste_fused_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                          const uint8_t* __restrict__ valid, uint8_t* __restrict__ out,
                                                          int m, int out_h, int out_w, float threshold, int rows_cap,
                                                          int pitch, int chunks_per_slab) {
  extern __shared__ uint4 s_dyn4[];
  const int r = blockIdx.y, tid = threadIdx.x;
  const int nchunks = (out_h * out_w) >> 4;
  const int cA = blockIdx.x * chunks_per_slab;
  const int cB = min(cA + chunks_per_slab, nchunks);
  if (cA >= cB) return;
  uint4* __restrict__ oplane = reinterpret_cast<uint4*>(out + (size_t)r * out_h * out_w);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  int xa = 0, xb = 0, ya = 0, yb = 0;
  float4 b = make_float4(0.f, 0.f, 1.f, 1.f);
  if (valid[r] != 0) {
    b = reinterpret_cast<const float4*>(boxes)[r];
    xa = max((int)floorf(b.x) - 1, 0); ya = max((int)floorf(b.y) - 1, 0);
    xb = min((int)ceilf(b.z) + 1, out_w); yb = min((int)ceilf(b.w) + 1, out_h);
    if (xa >= xb || ya >= yb) xa = xb = ya = yb = 0;
  }
  const int yA = (int)((uint32_t)(cA * 16) / (uint32_t)out_w), yB = (int)((uint32_t)(cB * 16 - 1) / (uint32_t)out_w);   // image rows this slab touches
  const int r0 = max(ya, yA), r1 = min(yb, yB + 1);                   // window rows inside the slab: [r0, r1)
  if (r0 >= r1) {                                                     // common: nothing but zeros
    for (int c = cA + tid; c < cB; c += 256) oplane[c] = zero4;
    return;
  }
  // The slab is walked by image rows: row y owns the chunks that START in it, [ceil(y W / 16), ceil((y + 1) W / 16)),
  // clipped to the slab; only the last of them can reach into row y + 1.  A warp takes a row at a time, its lanes
  // consecutive chunks (512 contiguous bytes per store instruction).  `window` selects the rows whose chunks can touch
  // the window (the row itself or the next one is a window row); the other rows are plain zeros.
  const int warp = tid >> 5, lane = tid & 31;
  const int nstr = (xb - xa + 31) >> 5;                               // 32-column strips; words 1 .. nstr of a tile row
  const uint32_t* s_bits_c = nullptr;
  auto row_pass = [&](bool window) {
    for (int y = yA + warp; y <= yB; y += 8) {
      const bool in_y = y >= r0 && y < r1, in_next = y + 1 >= r0 && y + 1 < r1;
      if ((in_y || in_next) != window) continue;
      const int row0 = y * out_w;
      const int cs = max((row0 + 15) >> 4, cA), ce = min((row0 + out_w + 15) >> 4, cB);
      if (!window) {
        for (int c = cs + lane; c < ce; c += 32) oplane[c] = zero4;
        continue;
      }
      const uint32_t* trow = s_bits_c + (y - r0) * pitch;             // only dereferenced when the row is a window row
      for (int c = cs + lane; c < ce; c += 32) {
        const int x = c * 16 - row0, xe = x + 15;
        uint32_t bits = 0u;
        if (in_y && xe >= xa && x < xb) {                             // x - xa + 32 lies in [17, 32 nstr + 32)
          const int rel = x - xa + 32;
          bits = __funnelshift_r(trow[rel >> 5], trow[(rel >> 5) + 1], rel & 31) & 0xffffu;
        }
        if (in_next && xe >= out_w && xa <= xe - out_w) {             // the chunk's tail lies in columns 0 .. of row y + 1
          const int n1 = out_w - x, rel = 32 - xa;                    // xa <= 14 here
          const uint32_t nb = __funnelshift_r(trow[pitch + (rel >> 5)], trow[pitch + (rel >> 5) + 1], rel & 31);
          bits = (bits & ((1u << n1) - 1u)) | ((nb << n1) & 0xffffu);
        }
        oplane[c] = bits ? make_uint4(spread4(bits & 15u), spread4((bits >> 4) & 15u), spread4((bits >> 8) & 15u), spread4(bits >> 12))
                         : zero4;
      }
    }
  };
  // ---- phase Z
  row_pass(false);
  // ---- tables
  const int mq = m + 3;
  float4* s_m4 = reinterpret_cast<float4*>(s_dyn4);
  PasteRow* s_row = reinterpret_cast<PasteRow*>(s_dyn4 + mq * mq);
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_dyn4 + mq * mq + rows_cap);
  // the mask with a zero border of two cells goes to shared memory first (it lives in the not yet used bit tile), so
  // that the gathered table is built without bounds tests
  {
    const float* pm = probs + (size_t)r * m * m;
    float* s_pad = reinterpret_cast<float*>(s_bits);
    const int mpd = m + 4;
    for (int i = tid; i < mpd * mpd; i += 256) s_pad[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < m * m; i += 256) {
      const int yy = (int)((uint32_t)i / (uint32_t)m), xx = i - yy * m;
      s_pad[(yy + 2) * mpd + xx + 2] = __ldg(pm + i);
    }
    __syncthreads();
    for (int i = tid; i < mq * mq; i += 256) {
      const int yi = (int)((uint32_t)i / (uint32_t)mq), xi = i - yi * mq;        // (yl + 2, xl + 2)
      const float* q = s_pad + yi * mpd + xi;
      s_m4[i] = make_float4(q[0], q[1], q[mpd], q[mpd + 1]);
    }
    __syncthreads();
  }
  const float bw = b.z - b.x, bh = b.w - b.y, fm = (float)m;
  const int nrows = r1 - r0;
  for (int i = tid; i < nrows; i += 256) {
    const int y = r0 + i;
    const float gy = ((float)y + 0.5f - b.y) / bh * 2.f - 1.f;
    const float iy = ((gy + 1.f) * fm - 1.f) * 0.5f;
    const float fy = floorf(iy);
    PasteRow pr;
    pr.wy1 = iy - fy; pr.wy0 = (fy + 1.f) - iy;                      // ATen: (iy_se - iy), (iy - iy_nw)
    const int yl = (int)fminf(fmaxf(fy, -2.f), fm);
    pr.base = (yl + 2) * mq * 16; pr.pad = 0;
    s_row[i] = pr;
  }
  for (int i = tid; i < nrows * pitch; i += 256) s_bits[i] = 0u;
  __syncthreads();
  // ---- phase C
  {
    // work items = strips x row groups, at least ~24 of them so that the eight warps finish together
    int groups = (24 + nstr - 1) / nstr;
    groups = max(1, min(groups, nrows >> 2));
    const int rows_per_group = (nrows + groups - 1) / groups;
    for (int item = warp; item < nstr * groups; item += 8) {
      const int g = item / nstr, j = item - g * nstr;
      const int x = xa + 32 * j + lane;
      const float gx = ((float)x + 0.5f - b.x) / bw * 2.f - 1.f;
      const float ix = ((gx + 1.f) * fm - 1.f) * 0.5f;
      const float fx = floorf(ix);
      const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
      const int xl = (int)fminf(fmaxf(fx, -2.f), fm);
      const uint32_t tap0 = (uint32_t)__cvta_generic_to_shared(s_m4) + (uint32_t)(xl + 2) * 16u;
      const bool in_x = x < xb;
      const int i_begin = g * rows_per_group, i_end = min((g + 1) * rows_per_group, nrows);
      uint32_t row_addr = (uint32_t)__cvta_generic_to_shared(s_row + i_begin);
      uint32_t bit_addr = (uint32_t)__cvta_generic_to_shared(s_bits + i_begin * pitch + j + 1);
      const uint32_t bit_step = (uint32_t)pitch * 4u;
      for (int i = i_begin; i < i_end; ++i, row_addr += 16u, bit_addr += bit_step) {
        float wy0, wy1, t0, t1, t2, t3;
        uint32_t base;                                    // the fourth word of a row record is padding
        asm volatile("{\n.reg .b32 pad;\nld.shared.v4.b32 {%0, %1, %2, pad}, [%3];\n}" : "=f"(wy0), "=f"(wy1), "=r"(base) : "r"(row_addr) : "memory");
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t0), "=f"(t1), "=f"(t2), "=f"(t3) : "r"(tap0 + base) : "memory");
        // same accumulation order as ATen's grid_sampler_2d (nw, ne, sw, se); border cells contribute exact zeros
        float v = 0.f;
        v += t0 * (wx0 * wy0);
        v += t1 * (wx1 * wy0);
        v += t2 * (wx0 * wy1);
        v += t3 * (wx1 * wy1);
        const uint32_t bits = __ballot_sync(0xffffffffu, in_x && v >= threshold);
        if (lane == 0) asm volatile("st.shared.b32 [%0], %1;" :: "r"(bit_addr), "r"(bits) : "memory");
      }
    }
  }
  __syncthreads();
  // ---- phase W
  s_bits_c = s_bits;
  row_pass(true);
}

int grid_for(int64_t work, int block);

}  // namespace cm2

using namespace cm2;

template <typename T>
static int roialign_launch(const cm2_act* feats, const int32_t* feat_stride, int num_levels, const float* boxes,
                           const int32_t* det_count, int n, int r_cap, const float* image_area, int crit,
                           int sampling_ratio, const cm2_act* out, int32_t* level_out, void* workspace, cudaStream_t s) {
  RoiAlignParams<T> p;
  for (int l = 0; l < ROI_MAX_LEVELS; ++l) {
    int k = l < num_levels ? l : num_levels - 1;
    p.feat[l] = make_view<const T>(feats[k]);
    p.scale[l] = 1.0f / (float)feat_stride[k];
  }
  auto ilog2 = [](int v) { int l = 0; while ((1 << (l + 1)) <= v) ++l; return l; };
  p.num_levels = num_levels;
  p.min_level = ilog2(feat_stride[0]);
  p.max_level = ilog2(feat_stride[num_levels - 1]);
  p.boxes = boxes; p.det_count = det_count; p.image_area = image_area;
  p.n = n; p.r_cap = r_cap; p.crit = crit; p.sampling_ratio = sampling_ratio; p.res = out->h;
  p.out = make_view<T>(*out);
  p.level_out = level_out;
  int64_t total = (int64_t)n * r_cap * out->h * out->w * (out->c / 8);
  const int variant = getenv("CM2_ROIALIGN_VARIANT") ? atoi(getenv("CM2_ROIALIGN_VARIANT")) : 2;
  bool small = true;
  for (int l = 0; l < num_levels; ++l) small = small && feats[l].sn < (1ll << 31);
  p.order = reinterpret_cast<int*>(workspace);
  if (variant == 3 && workspace && p.res < 32 && small && out->c <= 256) {
    CM2_ENSURE_DYN_SMEM(roialign_ring_kernel<T>, ROI_RING_BYTES, "roialign_ring");      // per instantiation (T) and device
    roi_order_kernel<T><<<1, 1024, 0, s>>>(p);
    roialign_ring_kernel<T><<<n * r_cap * ((p.res + ROI_RING_COLS - 1) / ROI_RING_COLS), 32 * (ROI_RING_COLS + 1), ROI_RING_BYTES, s>>>(p);
    return 0;
  }
  if constexpr (sizeof(T) == 2) {
    if (variant == 4 && workspace && p.res <= ROI_MMA_M && small && out->c % 32 == 0 && out->c <= 256) {
      const int stages = getenv("CM2_ROIALIGN_STAGES") ? atoi(getenv("CM2_ROIALIGN_STAGES")) : 4;
      const bool split = !(getenv("CM2_ROIALIGN_SPLIT") && atoi(getenv("CM2_ROIALIGN_SPLIT")) == 0);
      CM2_CHECK_ARG(stages >= 3 && stages <= 4, "roialign: CM2_ROIALIGN_STAGES %d not in [3,4]", stages);
      bool sw256 = true;
      for (int l = 0; l < num_levels; ++l) sw256 = sw256 && feats[l].sw == 256;
      const int smem = roi_mma_smem_bytes(out->c, stages);
      roi_order_kernel<T><<<1, 1024, 0, s>>>(p);
#define CM2_ROI_MMA_LAUNCH(ST, SP, SWV, MB)                                                                   \
  do {                                                                                                        \
    CM2_ENSURE_DYN_SMEM((roialign_mma_kernel<ST, SP, SWV, MB>), smem, "roialign_mma");                        \
    roialign_mma_kernel<ST, SP, SWV, MB><<<n * r_cap, out->c, smem, s>>>(p);                                  \
  } while (0)
#define CM2_ROI_MMA_LAUNCH_SW(ST, SP, MB)                                                                     \
  do {                                                                                                        \
    if (sw256) CM2_ROI_MMA_LAUNCH(ST, SP, 256, MB); else CM2_ROI_MMA_LAUNCH(ST, SP, 0, MB);                   \
  } while (0)
#define CM2_ROI_MMA_LAUNCH_SP(ST, MB)                                                                         \
  do {                                                                                                        \
    if (split) CM2_ROI_MMA_LAUNCH_SW(ST, true, MB); else CM2_ROI_MMA_LAUNCH_SW(ST, false, MB);                \
  } while (0)
      if (stages == 3) CM2_ROI_MMA_LAUNCH_SP(3, 2); else CM2_ROI_MMA_LAUNCH_SP(4, 2);
#undef CM2_ROI_MMA_LAUNCH_SP
#undef CM2_ROI_MMA_LAUNCH_SW
#undef CM2_ROI_MMA_LAUNCH
      return 0;
    }
  }
  // variant 4 on fp32 maps / other shapes: the column walk
  if ((variant == 2 || variant == 4) && workspace && p.res < 32 && small && out->c <= 256) {   // lane 31 builds the x table, lanes < res the y table
    roi_order_kernel<T><<<1, 1024, 0, s>>>(p);
    roialign_col_kernel<T><<<n * r_cap * p.res, 32, roi_col_smem_bytes(p.res), s>>>(p);
    return 0;
  }
  if (variant >= 1 && p.res <= ROI_MAX_RES && small) {
    const int items = p.res * p.res * (out->c / 8);
    const int threads = std::min(256, std::max(64, (items + 31) / 32 * 32));
    roialign_roi_kernel<T><<<n * r_cap, threads, 0, s>>>(p);
    return 0;
  }
  roialign_fpn_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(p);
  return 0;
}

extern "C" int cm2_roialign_fpn(const cm2_act* feats, const int32_t* feat_stride, int32_t num_levels, int32_t dtype,
                                const float* boxes, const int32_t* det_count, int32_t n, int32_t r_cap,
                                const float* image_area, int32_t crit, int32_t sampling_ratio, const cm2_act* out,
                                int32_t* level_out, void* workspace, void* stream) {
  CM2_CHECK_ARG(feats && feat_stride && boxes && det_count && image_area && out && out->data, "roialign: null pointer");
  CM2_CHECK_ARG(num_levels >= 1 && num_levels <= ROI_MAX_LEVELS, "roialign: num_levels %d not in [1,%d]", num_levels,
                ROI_MAX_LEVELS);
  CM2_CHECK_ARG(dtype == CM2_F32 || dtype == CM2_BF16, "roialign: dtype %d not supported", dtype);
  CM2_CHECK_ARG(crit == 0 || crit == 1, "roialign: crit %d", crit);
  CM2_CHECK_ARG(out->n == n * r_cap && out->h == out->w && out->h > 0, "roialign: out view [%d,%d,%d,%d] vs n=%d r_cap=%d",
                out->n, out->h, out->w, out->c, n, r_cap);
  int eb = elem_bytes(dtype);
  CM2_CHECK_ARG(vec8_ok(*out, eb), "roialign: out channels/strides must be multiples of 8");
  for (int l = 0; l < num_levels; ++l) {
    CM2_CHECK_ARG(feats[l].data && feats[l].c == out->c && feats[l].n == n && vec8_ok(feats[l], eb),
                  "roialign: feature level %d invalid", l);
    CM2_CHECK_ARG(feat_stride[l] > 0 && (feat_stride[l] & (feat_stride[l] - 1)) == 0, "roialign: stride %d not a power of 2",
                  feat_stride[l]);
  }
  if (n * r_cap == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if (dtype == CM2_F32)
    rc = roialign_launch<float>(feats, feat_stride, num_levels, boxes, det_count, n, r_cap, image_area, crit, sampling_ratio,
                                out, level_out, workspace, s);
  else
    rc = roialign_launch<__nv_bfloat16>(feats, feat_stride, num_levels, boxes, det_count, n, r_cap, image_area, crit,
                                        sampling_ratio, out, level_out, workspace, s);
  if (rc != CM2_OK) return rc;
  CM2_CHECK_LAUNCH("roialign_fpn");
  return CM2_OK;
}

extern "C" int cm2_spatial_attention(const cm2_act* x, const cm2_act* out, int32_t dtype, const float* w18,
                                     void* stream) {
  CM2_CHECK_ARG(x && out && x->data && out->data && w18, "spatial_attention: null pointer");
  CM2_CHECK_ARG(dtype == CM2_F32 || dtype == CM2_BF16, "spatial_attention: dtype %d not supported", dtype);
  int eb = elem_bytes(dtype);
  CM2_CHECK_ARG(same_extent(*x, *out) && vec8_ok(*x, eb) && vec8_ok(*out, eb) && x->h == x->w && x->h <= SAM_MAX_S,
                "spatial_attention: bad views [%d,%d,%d,%d]", x->n, x->h, x->w, x->c);
  if (x->n == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t tile_bytes = (size_t)x->h * x->w * x->c * 2;
  const int sam_variant = getenv("CM2_SAM_VARIANT") ? atoi(getenv("CM2_SAM_VARIANT")) : 1;
  const size_t pipe_smem = 2 * tile_bytes + (size_t)3 * x->h * x->w * sizeof(float);
  if (sam_variant == 1 && dtype == CM2_BF16 && pipe_smem <= 220 * 1024 && x->sw == x->c && x->n >= 2 * 148 &&
      ((size_t)x->w * x->c * 2) % 16 == 0) {
    // persistent double-buffered kernel: worth it once every SM gets at least two ROIs
    CM2_ENSURE_DYN_SMEM(spatial_attention_pipe_kernel, 220 * 1024, "spatial_attention_pipe");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    spatial_attention_pipe_kernel<<<std::min(sms, x->n), 1024, pipe_smem, s>>>(make_view<const __nv_bfloat16>(*x),
                                                                               make_view<__nv_bfloat16>(*out), w18);
    CM2_CHECK_LAUNCH("spatial_attention_pipe");
    return CM2_OK;
  }
  if (dtype == CM2_BF16 && tile_bytes <= 110 * 1024) {     // two CTAs per SM
    CM2_ENSURE_DYN_SMEM(spatial_attention_smem_kernel, 110 * 1024, "spatial_attention_smem");
    spatial_attention_smem_kernel<<<x->n, 512, tile_bytes, s>>>(make_view<const __nv_bfloat16>(*x), make_view<__nv_bfloat16>(*out), w18);
    CM2_CHECK_LAUNCH("spatial_attention_smem");
    return CM2_OK;
  }
  if (dtype == CM2_F32)
    spatial_attention_kernel<float><<<x->n, 256, 0, s>>>(make_view<const float>(*x), make_view<float>(*out), w18);
  else
    spatial_attention_kernel<__nv_bfloat16><<<x->n, 256, 0, s>>>(make_view<const __nv_bfloat16>(*x),
                                                               make_view<__nv_bfloat16>(*out), w18);
  CM2_CHECK_LAUNCH("spatial_attention");
  return CM2_OK;
}

extern "C" int cm2_mask_predict(const cm2_act* x, int32_t dtype, const float* wp, const float* bp,
                                const int64_t* classes, int32_t ncls, float* probs, void* stream) {
  CM2_CHECK_ARG(x && x->data && wp && bp && probs && (classes || ncls == 1), "mask_predict: null pointer");
  CM2_CHECK_ARG(dtype == CM2_F32 || dtype == CM2_BF16, "mask_predict: dtype %d not supported", dtype);
  CM2_CHECK_ARG(vec8_ok(*x, elem_bytes(dtype)) && ncls >= 1, "mask_predict: bad view (c=%d)", x->c);
  int64_t total = (int64_t)x->n * x->h * x->w;
  if (total == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int grid = grid_for(total * 32, 256);
  if (dtype == CM2_F32)
    mask_predict_kernel<float><<<grid, 256, 0, s>>>(make_view<const float>(*x), wp, bp, classes, ncls, probs);
  else
    mask_predict_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(make_view<const __nv_bfloat16>(*x), wp, bp, classes, ncls,
                                                          probs);
  CM2_CHECK_LAUNCH("mask_predict");
  return CM2_OK;
}

extern "C" int cm2_maskiou_prep(const float* probs, const cm2_act* out, int32_t dtype, void* stream) {
  CM2_CHECK_ARG(probs && out && out->data, "maskiou_prep: null pointer");
  CM2_CHECK_ARG(dtype == CM2_F32 || dtype == CM2_BF16, "maskiou_prep: dtype %d not supported", dtype);
  CM2_CHECK_ARG(out->h == out->w && out->c >= 1, "maskiou_prep: bad out view");
  int64_t total = (int64_t)out->n * out->h * out->w;
  if (total == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CM2_F32)
    maskiou_prep_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(probs, make_view<float>(*out));
  else
    maskiou_prep_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>(probs, make_view<__nv_bfloat16>(*out));
  CM2_CHECK_LAUNCH("maskiou_prep");
  return CM2_OK;
}

extern "C" int cm2_maskiou_score(const void* iou, int32_t dtype, int32_t r, int32_t ncls, const int64_t* classes,
                                 const float* scores, float* mask_scores, void* stream) {
  CM2_CHECK_ARG(iou && classes && scores && mask_scores, "maskiou_score: null pointer");
  CM2_CHECK_ARG(dtype == CM2_F32 || dtype == CM2_BF16, "maskiou_score: dtype %d not supported", dtype);
  if (r == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CM2_F32)
    maskiou_score_kernel<float><<<ceil_div(r, 128), 128, 0, s>>>((const float*)iou, r, ncls, classes, scores, mask_scores);
  else
    maskiou_score_kernel<__nv_bfloat16><<<ceil_div(r, 128), 128, 0, s>>>((const __nv_bfloat16*)iou, r, ncls, classes,
                                                                       scores, mask_scores);
  CM2_CHECK_LAUNCH("maskiou_score");
  return CM2_OK;
}

extern "C" int cm2_scale_clip_boxes(const float* boxes_in, float* boxes_out, uint8_t* valid, int32_t r, float sx,
                                    float sy, float out_w, float out_h, void* stream) {
  CM2_CHECK_ARG(boxes_in && boxes_out && valid, "scale_clip_boxes: null pointer");
  if (r == 0) return CM2_OK;
  scale_clip_boxes_kernel<<<ceil_div(r, 128), 128, 0, (cudaStream_t)stream>>>(boxes_in, boxes_out, valid, r, sx, sy,
                                                                             out_w, out_h);
  CM2_CHECK_LAUNCH("scale_clip_boxes");
  return CM2_OK;
}

extern "C" int cm2_scale_clip_boxes_batch(const float* boxes_in, float* boxes_out, uint8_t* valid, int32_t n, int32_t r_cap,
                                          const float* params, const int32_t* det_count, void* stream) {
  CM2_CHECK_ARG(boxes_in && boxes_out && valid && params, "scale_clip_boxes_batch: null pointer");
  CM2_CHECK_ARG(n >= 0 && r_cap > 0, "scale_clip_boxes_batch: bad extents n=%d r_cap=%d", n, r_cap);
  if (n == 0) return CM2_OK;
  scale_clip_boxes_batch_kernel<<<ceil_div(n * r_cap, 128), 128, 0, (cudaStream_t)stream>>>(boxes_in, boxes_out, valid, n * r_cap,
                                                                                           r_cap, params, det_count);
  CM2_CHECK_LAUNCH("scale_clip_boxes_batch");
  return CM2_OK;
}

extern "C" int cm2_pack_records(const float* boxes, const float* scores, const int64_t* classes, const float* mask_scores,
                                const float* locations, const uint8_t* valid, const int32_t* det_count, int32_t n, int32_t r_cap,
                                float* records, void* stream) {
  CM2_CHECK_ARG(boxes && scores && classes && det_count && records, "pack_records: null pointer");
  CM2_CHECK_ARG(n >= 0 && r_cap > 0 && (reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "pack_records: bad extents n=%d r_cap=%d", n, r_cap);
  if (n == 0) return CM2_OK;
  pack_records_kernel<<<ceil_div(n * r_cap, 128), 128, 0, (cudaStream_t)stream>>>(
      boxes, scores, reinterpret_cast<const long long*>(classes), mask_scores, locations, valid, det_count, n * r_cap, r_cap, records);
  CM2_CHECK_LAUNCH("pack_records");
  return CM2_OK;
}

extern "C" int cm2_paste_masks(const float* probs, const float* boxes, const uint8_t* valid, uint8_t* out, int32_t r,
                               int32_t m, int32_t out_h, int32_t out_w, float threshold, void* stream) {
  CM2_CHECK_ARG(probs && boxes && valid && out, "paste_masks: null pointer");
  CM2_CHECK_ARG(m > 0 && m <= 64 && out_h > 0 && out_w > 0, "paste_masks: bad extents m=%d out=%dx%d", m, out_h, out_w);
  if (r == 0) return CM2_OK;
  CM2_CHECK_ARG(r <= 65535, "paste_masks: too many ROIs in one call (%d)", r);
  cudaStream_t s = (cudaStream_t)stream;
  const int variant = getenv("CM2_PASTE_VARIANT") ? atoi(getenv("CM2_PASTE_VARIANT")) : 1;
  const long long plane = (long long)out_h * out_w;
  if (variant == 1 && plane % 16 == 0 && plane >= 16 && plane < (1ll << 27) && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    // slabs of about 48 image rows; shared memory: M4 table, row table, bit tile
    const int nchunks = (int)(plane >> 4);
    int slabs = std::max(1, std::min(ceil_div(out_h, 48), nchunks));
    const int chunks_per_slab = ceil_div(nchunks, slabs);
    slabs = ceil_div(nchunks, chunks_per_slab);
    const int rows_cap = (chunks_per_slab * 16) / out_w + 3;
    const int pitch = ceil_div(out_w, 32) + 4;
    const size_t tile_bytes = std::max((size_t)rows_cap * pitch * 4, (size_t)(m + 4) * (m + 4) * 4);
    const size_t smem = ((size_t)(m + 3) * (m + 3) + rows_cap) * 16 + tile_bytes;
    if (smem <= 200 * 1024) {
      CM2_ENSURE_DYN_SMEM(paste_fused_kernel, smem, "paste_masks_fused");
      paste_fused_kernel<<<dim3(slabs, r), 256, smem, s>>>(probs, boxes, valid, out, m, out_h, out_w, threshold, rows_cap, pitch,
                                                           chunks_per_slab);
      CM2_CHECK_LAUNCH("paste_masks_fused");
      return CM2_OK;
    }
  }
  if (cudaMemsetAsync(out, 0, (size_t)r * out_h * out_w, s) != cudaSuccess) {
    set_error("paste_masks: cudaMemsetAsync failed");
    return CM2_ERR_CUDA;
  }
  dim3 grid(PASTE_CTAS_PER_ROI, r);
  const size_t mask_floats = (size_t)(((m + 2) * (m + 2) + 3) & ~3);
  if (((long long)out_h * out_w) % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0 && out_w <= PASTE_MAX_W) {
    const size_t smem = mask_floats * sizeof(float) + (size_t)out_w * sizeof(PasteCol);
    CM2_ENSURE_DYN_SMEM(paste_window_words_kernel, ((64 + 2) * (64 + 2) + 4) * sizeof(float) + PASTE_MAX_W * sizeof(PasteCol),
                        "paste_masks_words");
    paste_window_words_kernel<<<grid, 256, smem, s>>>(probs, boxes, valid, out, m, out_h, out_w, threshold);
    CM2_CHECK_LAUNCH("paste_masks_words");
    return CM2_OK;
  }
  paste_window_kernel<<<grid, 256, (size_t)(m + 2) * (m + 2) * sizeof(float), s>>>(probs, boxes, valid, out, m, out_h, out_w, threshold);
  CM2_CHECK_LAUNCH("paste_masks");
  return CM2_OK;
}
