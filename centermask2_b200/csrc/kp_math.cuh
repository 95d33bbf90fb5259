// Arithmetic of the keypoint decode (bilinear x2 of the low-resolution logits, bicubic resize to the ROI's own
// pixels), shared by the kernel (keypoint.cu) and by the host harness of tests/test_keypoint_math.py, which compiles
// this header with g++ to check the index / coefficient logic against torch's CPU interpolate without a GPU.
//
// The formulas follow ATen's upsample kernels (area_pixel_compute_source_index, guard_index_and_lambda,
// get_cubic_upsample_coefficients with A = -0.75), the leaf ops behind the reference's
//   keypoint_head.py:221  interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
//   detectron2 heatmaps_to_keypoints [d2]: F.interpolate(maps[[i]], size=(ceil(h), ceil(w)), mode="bicubic",
//                                          align_corners=False)      (call site keypoint_head.py:113)
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define KP_HD __host__ __device__ __forceinline__
#else
#define KP_HD inline
#endif

namespace cm2 {

// low-resolution logits as the conv engine writes them: [roi][res][res][4][k] with phase = (y & 1) * 2 + (x & 1)
// of the 2*res x 2*res map that ConvTranspose2d(k 4, s 2, p 1) produces (keypoint_head.py:205-208).
KP_HD size_t kp_lowres_offset(int roi, int y, int x, int kp, int res, int k) {
  return ((((size_t)roi * res + (y >> 1)) * res + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) * k + kp;
}

// bilinear, align_corners=False, scale 1/2: source index and weight of output index o (ATen clamps src at 0).
KP_HD void kp_bilinear_src(int o, int in_size, int& i0, int& i1, float& l1) {
  float src = 0.5f * ((float)o + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

// one pixel of the x2 bilinear map from the low-res map `low` [in][in]; accumulation order of ATen's separable
// kernel: x inside each row, then y.
KP_HD float kp_bilinear_at(const float* low, int in_size, int oy, int ox) {
  int y0, y1, x0, x1;
  float ly, lx;
  kp_bilinear_src(oy, in_size, y0, y1, ly);
  kp_bilinear_src(ox, in_size, x0, x1, lx);
  const float wy0 = 1.f - ly, wx0 = 1.f - lx;
  float t0 = low[y0 * in_size + x0] * wx0;
  t0 = fmaf(low[y0 * in_size + x1], lx, t0);
  float t1 = low[y1 * in_size + x0] * wx0;
  t1 = fmaf(low[y1 * in_size + x1], lx, t1);
  float out = t0 * wy0;
  return fmaf(t1, ly, out);
}

struct KpCubic {
  int idx[4];      // clamped source indices
  float w[4];      // cubic convolution coefficients
};

// bicubic, align_corners=False, scale = in / out (float): taps of output index o.
KP_HD KpCubic kp_cubic_taps(float scale, int o, int in_size) {
  const float A = -0.75f;
  const float real = scale * ((float)o + 0.5f) - 0.5f;
  float fl = floorf(real);
  int i = (int)fl;
  if (i > in_size - 1) i = in_size - 1;                      // guard_index_and_lambda
  float t = real - (float)i;
  t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
  KpCubic c;
  float x = t + 1.f;
  c.w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;
  c.w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;
  c.w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = x + 1.f;
  c.w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int q = i - 1 + j;
    c.idx[j] = q < 0 ? 0 : (q > in_size - 1 ? in_size - 1 : q);
  }
  return c;
}

// one pixel of the bicubic-resized map from `hi` [in][in]; x inside each of the four rows, then y.
KP_HD float kp_bicubic_at(const float* hi, int in_size, const KpCubic& cy, const KpCubic& cx) {
  float out = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float* row = hi + cy.idx[a] * in_size;
    float t = row[cx.idx[0]] * cx.w[0];
    t = fmaf(row[cx.idx[1]], cx.w[1], t);
    t = fmaf(row[cx.idx[2]], cx.w[2], t);
    t = fmaf(row[cx.idx[3]], cx.w[3], t);
    out = a == 0 ? t * cy.w[0] : fmaf(t, cy.w[a], out);
  }
  return out;
}

// ROI geometry of heatmaps_to_keypoints [d2]: widths / heights clamped to >= 1, resized map = ceil of them.
struct KpRoi {
  float x0, y0, w, h;
  int wc, hc;
};
KP_HD KpRoi kp_roi(float bx0, float by0, float bx1, float by1) {
  KpRoi r;
  r.x0 = bx0;
  r.y0 = by0;
  r.w = bx1 - bx0;
  r.h = by1 - by0;
  r.w = r.w < 1.f ? 1.f : r.w;
  r.h = r.h < 1.f ? 1.f : r.h;
  r.wc = (int)ceilf(r.w);
  r.hc = (int)ceilf(r.h);
  return r;
}

}  // namespace cm2
