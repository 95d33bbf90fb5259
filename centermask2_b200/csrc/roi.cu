// ROI-stage and output-side kernels: FPN level assignment fused into ROIAlign, SAG spatial attention,
// class-gathered mask predictor + sigmoid, MaskIoU input/score glue, box rescale and mask paste-back.
// All bandwidth-bound: coalesced over channels (NHWC) / over x (paste), 8 channels per thread.
#include "common.cuh"

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
                                              int total, int r_cap, const float* __restrict__ params) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float4 q = __ldg(reinterpret_cast<const float4*>(params) + i / r_cap);
  float4 b = reinterpret_cast<const float4*>(in)[i];
  b.x = fminf(fmaxf(b.x * q.x, 0.f), q.z);
  b.z = fminf(fmaxf(b.z * q.x, 0.f), q.z);
  b.y = fminf(fmaxf(b.y * q.y, 0.f), q.w);
  b.w = fminf(fmaxf(b.w * q.y, 0.f), q.w);
  reinterpret_cast<float4*>(out)[i] = b;
  valid[i] = ((b.z - b.x) > 0.f && (b.w - b.y) > 0.f) ? 1 : 0;
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

int grid_for(int64_t work, int block);

}  // namespace cm2

using namespace cm2;

template <typename T>
static int roialign_launch(const cm2_act* feats, const int32_t* feat_stride, int num_levels, const float* boxes,
                           const int32_t* det_count, int n, int r_cap, const float* image_area, int crit,
                           int sampling_ratio, const cm2_act* out, int32_t* level_out, cudaStream_t s) {
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
  roialign_fpn_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(p);
  return 0;
}

extern "C" int cm2_roialign_fpn(const cm2_act* feats, const int32_t* feat_stride, int32_t num_levels, int32_t dtype,
                                const float* boxes, const int32_t* det_count, int32_t n, int32_t r_cap,
                                const float* image_area, int32_t crit, int32_t sampling_ratio, const cm2_act* out,
                                int32_t* level_out, void* stream) {
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
  if (dtype == CM2_F32)
    roialign_launch<float>(feats, feat_stride, num_levels, boxes, det_count, n, r_cap, image_area, crit, sampling_ratio,
                           out, level_out, s);
  else
    roialign_launch<__nv_bfloat16>(feats, feat_stride, num_levels, boxes, det_count, n, r_cap, image_area, crit,
                                   sampling_ratio, out, level_out, s);
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
  if (dtype == CM2_BF16 && tile_bytes <= 110 * 1024) {     // two CTAs per SM
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(spatial_attention_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
      attr_done = true;
    }
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
                                          const float* params, void* stream) {
  CM2_CHECK_ARG(boxes_in && boxes_out && valid && params, "scale_clip_boxes_batch: null pointer");
  CM2_CHECK_ARG(n >= 0 && r_cap > 0, "scale_clip_boxes_batch: bad extents n=%d r_cap=%d", n, r_cap);
  if (n == 0) return CM2_OK;
  scale_clip_boxes_batch_kernel<<<ceil_div(n * r_cap, 128), 128, 0, (cudaStream_t)stream>>>(boxes_in, boxes_out, valid, n * r_cap,
                                                                                           r_cap, params);
  CM2_CHECK_LAUNCH("scale_clip_boxes_batch");
  return CM2_OK;
}

extern "C" int cm2_paste_masks(const float* probs, const float* boxes, const uint8_t* valid, uint8_t* out, int32_t r,
                               int32_t m, int32_t out_h, int32_t out_w, float threshold, void* stream) {
  CM2_CHECK_ARG(probs && boxes && valid && out, "paste_masks: null pointer");
  CM2_CHECK_ARG(m > 0 && m <= 64 && out_h > 0 && out_w > 0, "paste_masks: bad extents m=%d out=%dx%d", m, out_h, out_w);
  if (r == 0) return CM2_OK;
  CM2_CHECK_ARG(r <= 65535, "paste_masks: too many ROIs in one call (%d)", r);
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(out, 0, (size_t)r * out_h * out_w, s) != cudaSuccess) {
    set_error("paste_masks: cudaMemsetAsync failed");
    return CM2_ERR_CUDA;
  }
  dim3 grid(PASTE_CTAS_PER_ROI, r);
  const size_t mask_floats = (size_t)(((m + 2) * (m + 2) + 3) & ~3);
  if (((long long)out_h * out_w) % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0 && out_w <= PASTE_MAX_W) {
    const size_t smem = mask_floats * sizeof(float) + (size_t)out_w * sizeof(PasteCol);
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(paste_window_words_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(((64 + 2) * (64 + 2) + 4) * sizeof(float) + PASTE_MAX_W * sizeof(PasteCol)));
      attr_done = true;
    }
    paste_window_words_kernel<<<grid, 256, smem, s>>>(probs, boxes, valid, out, m, out_h, out_w, threshold);
    CM2_CHECK_LAUNCH("paste_masks_words");
    return CM2_OK;
  }
  paste_window_kernel<<<grid, 256, (size_t)(m + 2) * (m + 2) * sizeof(float), s>>>(probs, boxes, valid, out, m, out_h, out_w, threshold);
  CM2_CHECK_LAUNCH("paste_masks");
  return CM2_OK;
}
