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

// The same taps without float index arithmetic: for scale 1/2 the source coordinate of o = 2i is i - 1/4 (clamped to 0
// for o = 0) and of o = 2i + 1 is i + 1/4, so the weights are the exact constants 3/4, 1/4 (or 0) and both
// formulations give bit-identical results (checked exhaustively per size in tests/test_keypoint_math.py).
KP_HD void kp_bilinear2_src(int o, int in_size, int& i0, int& i1, float& l1) {
  const int i = o >> 1;
  if (o & 1) {
    i0 = i;
    l1 = 0.25f;
  } else {
    i0 = i > 0 ? i - 1 : 0;
    l1 = i > 0 ? 0.75f : 0.f;
  }
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
}
KP_HD float kp_bilinear2_at(const float* low, int in_size, int oy, int ox) {
  int y0, y1, x0, x1;
  float ly, lx;
  kp_bilinear2_src(oy, in_size, y0, y1, ly);
  kp_bilinear2_src(ox, in_size, x0, x1, lx);
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
  int base;        // guarded floor of the source coordinate: idx[j] = clamp(base - 1 + j)
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
  c.base = i;
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

// y taps of one resized row as the column walk keeps them (a table in shared memory, or computed on the fly)
struct alignas(16) KpW4 {
  float w0, w1, w2, w3;
};
struct KpRowTaps {
  KpW4 w;
  int base;
};
KP_HD KpRowTaps kp_row_taps(float scale, int o, int in_size) {
  const KpCubic c = kp_cubic_taps(scale, o, in_size);
  KpRowTaps r;
  r.w.w0 = c.w[0]; r.w.w1 = c.w[1]; r.w.w2 = c.w[2]; r.w.w3 = c.w[3];
  r.base = c.base;
  return r;
}

// How the column walk reads the 4res x 4res map and the y-tap table.  This one goes through ordinary pointers (host
// harness, generic addressing); keypoint.cu has the shared-space twin (32-bit addresses, ld.shared) with the same
// three members, because generic addressing of dynamic shared memory re-derives the window base inside the loop.
struct KpMemPtr {
  const float* hi;
  const KpW4* wtab;
  const int* btab;
  KP_HD float hi_at(int elem) const { return hi[elem]; }
  KP_HD KpW4 w_at(int oy) const { return wtab[oy]; }
  KP_HD int base_at(int oy) const { return btab[oy]; }
};

// x pass of one source row: the inner sum of kp_bicubic_at.
template <typename Mem>
KP_HD float kp_row_interp(const Mem& m, int in_size, int row, const KpCubic& cx) {
  row = row < 0 ? 0 : (row > in_size - 1 ? in_size - 1 : row);
  const int o = row * in_size;
  float t = m.hi_at(o + cx.idx[0]) * cx.w[0];
  t = fmaf(m.hi_at(o + cx.idx[1]), cx.w[1], t);
  t = fmaf(m.hi_at(o + cx.idx[2]), cx.w[2], t);
  return fmaf(m.hi_at(o + cx.idx[3]), cx.w[3], t);
}

struct KpBest {
  float v;
  long long p;     // flat index oy * wc + ox; the smaller index wins among equal values (torch.argmax on CPU)
};
KP_HD void kp_best_merge(KpBest& a, float v, long long p) {
  if (v > a.v || (v == a.v && p < a.p)) { a.v = v; a.p = p; }
}

// Split of the hc x wc resized pixels of one (ROI, keypoint) into work items (column, row segment): about four
// items per thread for balance, segments of at least KP_MIN_SEG rows so that the window restarts stay cheap.
constexpr int KP_MIN_SEG = 8;
struct KpSplit {
  int seg_len, nseg;
  long long items;
};
KP_HD KpSplit kp_split(int hc, int wc, int nthreads) {
  long long want = (4LL * nthreads + wc - 1) / wc;
  const long long cap = (hc + KP_MIN_SEG - 1) / KP_MIN_SEG;
  want = want < 1 ? 1 : (want > cap ? cap : want);
  KpSplit s;
  s.seg_len = (int)((hc + want - 1) / want);
  s.nseg = (hc + s.seg_len - 1) / s.seg_len;
  s.items = (long long)s.nseg * wc;
  return s;
}

// Which ROIs take the column walk: those whose rows fit the y-tap table and whose item count fits 31 bits (wc up to
// 2^20 with at most tab_rows / KP_MIN_SEG segments); the others take the flat per-pixel loop.
KP_HD bool kp_walk_applies(int hc, int wc, int tab_rows) {
  return hc <= tab_rows && tab_rows <= 2048 && wc <= (1 << 20);
}

// The column walk of thread `tid`: for each of its (column, segment) items the x taps are computed once, the x pass of
// the four source rows under the current resized row is kept in registers and advanced when the source row changes
// (once per hc / in_size rows), and a resized pixel costs the four y FMAs + a compare.  Same expression tree per pixel
// as kp_bicubic_at, hence bit-identical values.  The y taps of every resized row come from the table behind `m`
// (weights as one 16-byte load + the source row base); ROIs taller than the table take the flat per-pixel loop.
template <typename Mem>
KP_HD KpBest kp_column_walk(const Mem& m, int in_size, int hc, int wc, float scale_x, int tid, int nthreads) {
  KpBest best;
  best.v = -INFINITY;
  best.p = 0x7fffffffffffffffLL;
  const KpSplit sp = kp_split(hc, wc, nthreads);
  // precondition (kp_walk_applies): the item count fits 31 bits, so the item -> (segment, column) split is one
  // 32-bit division per item (the 64-bit one was ~190 instructions per item in the first version)
  const unsigned items = (unsigned)sp.items, uwc = (unsigned)wc;
  for (unsigned item = (unsigned)tid; item < items; item += (unsigned)nthreads) {
    const int seg = (int)(item / uwc);
    const int ox = (int)(item - (unsigned)seg * uwc);
    const int row0 = seg * sp.seg_len;
    const int row1 = row0 + sp.seg_len < hc ? row0 + sp.seg_len : hc;
    const KpCubic cx = kp_cubic_taps(scale_x, ox, in_size);
    KpRowTaps ty;
    ty.w = m.w_at(row0);
    ty.base = m.base_at(row0);
    int base = ty.base;
    float t0 = kp_row_interp(m, in_size, base - 1, cx);
    float t1 = kp_row_interp(m, in_size, base, cx);
    float t2 = kp_row_interp(m, in_size, base + 1, cx);
    float t3 = kp_row_interp(m, in_size, base + 2, cx);
    float col_v = -INFINITY;
    int col_y = row0;
    for (int oy = row0;;) {
      float v = t0 * ty.w.w0;
      v = fmaf(t1, ty.w.w1, v);
      v = fmaf(t2, ty.w.w2, v);
      v = fmaf(t3, ty.w.w3, v);
      if (v > col_v) { col_v = v; col_y = oy; }
      if (++oy >= row1) break;
      ty.w = m.w_at(oy);
      ty.base = m.base_at(oy);
      int shift = ty.base - base;              // the source coordinate is monotonic in oy: shift >= 0
      if (shift != 0) {
        if (shift >= 4) {
          base = ty.base;
          t0 = kp_row_interp(m, in_size, base - 1, cx);
          t1 = kp_row_interp(m, in_size, base, cx);
          t2 = kp_row_interp(m, in_size, base + 1, cx);
          t3 = kp_row_interp(m, in_size, base + 2, cx);
        } else {
          do {
            ++base;
            t0 = t1; t1 = t2; t2 = t3;
            t3 = kp_row_interp(m, in_size, base + 2, cx);
          } while (--shift);
        }
      }
    }
    kp_best_merge(best, col_v, (long long)col_y * wc + ox);
  }
  return best;
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
