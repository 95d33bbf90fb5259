// TEST INFRASTRUCTURE ONLY: serial host evaluation of csrc/kp_math.cuh (the arithmetic of keypoints_decode_kernel),
// built by tests/test_keypoint_math.py with g++ so that the index / coefficient logic of the kernel is checked against
// torch's CPU interpolate in the CPU test tier.  Never loaded by the package.
#include <stddef.h>
#include <vector>
#include "../../centermask2_b200/csrc/kp_math.cuh"

// walk_threads > 0: evaluate through kp_column_walk with that many threads (the kernel uses 256) when the ROI fits the
// y-tap table of tab_rows rows, as the kernel does; otherwise, and with walk_threads == 0: the flat per-pixel loop.
extern "C" void kp_host_decode(const float* lowres, const float* boxes, int r, int res, int k, float* out, float* hi_out,
                               int walk_threads, int tab_rows) {
  using namespace cm2;
  const int s_low = 2 * res, s_hi = 4 * res;
  std::vector<float> low(s_low * s_low), hi(s_hi * s_hi);
  for (int roi = 0; roi < r; ++roi)
    for (int kp = 0; kp < k; ++kp) {
      for (int y = 0; y < s_low; ++y)
        for (int x = 0; x < s_low; ++x) low[y * s_low + x] = lowres[kp_lowres_offset(roi, y, x, kp, res, k)];
      for (int y = 0; y < s_hi; ++y)
        for (int x = 0; x < s_hi; ++x) hi[y * s_hi + x] = kp_bilinear2_at(low.data(), s_low, y, x);
      if (hi_out)
        for (int i = 0; i < s_hi * s_hi; ++i) hi_out[((size_t)roi * k + kp) * s_hi * s_hi + i] = hi[i];
      const float* b = boxes + 4 * roi;
      const KpRoi g = kp_roi(b[0], b[1], b[2], b[3]);
      const float sy = (float)s_hi / (float)g.hc, sx = (float)s_hi / (float)g.wc;
      float best = -INFINITY;
      long long best_p = 0;
      if (walk_threads > 0 && kp_walk_applies(g.hc, g.wc, tab_rows)) {
        // the kernel's default decomposition: every thread's column walk, merged like the block reduction
        std::vector<KpW4> wtab;
        std::vector<int> btab;
        for (int oy = 0; oy < g.hc; ++oy) {
          const KpRowTaps t = kp_row_taps(sy, oy, s_hi);
          wtab.push_back(t.w);
          btab.push_back(t.base);
        }
        KpMemPtr m;
        m.hi = hi.data();
        m.wtab = wtab.data();
        m.btab = btab.data();
        KpBest b;
        b.v = -INFINITY;
        b.p = 0x7fffffffffffffffLL;
        for (int tid = 0; tid < walk_threads; ++tid) {
          const KpBest t = kp_column_walk(m, s_hi, g.hc, g.wc, sx, tid, walk_threads);
          kp_best_merge(b, t.v, t.p);
        }
        best = b.v;
        best_p = b.p;
      } else {
        for (int oy = 0; oy < g.hc; ++oy) {
          const KpCubic cy = kp_cubic_taps(sy, oy, s_hi);
          for (int ox = 0; ox < g.wc; ++ox) {
            const KpCubic cx = kp_cubic_taps(sx, ox, s_hi);
            const float v = kp_bicubic_at(hi.data(), s_hi, cy, cx);
            if (v > best) { best = v; best_p = (long long)oy * g.wc + ox; }
          }
        }
      }
      float pool = 0.f;
      for (int i = 0; i < s_hi * s_hi; ++i) pool += expf(hi[i] - best);
      const long long yi = best_p / g.wc;
      const int xi = (int)(best_p - yi * g.wc);
      float* o = out + ((size_t)roi * k + kp) * 4;
      volatile float fx = ((float)xi + 0.5f) * (g.w / (float)g.wc);
      volatile float fy = ((float)yi + 0.5f) * (g.h / (float)g.hc);
      o[0] = fx + g.x0;
      o[1] = fy + g.y0;
      o[2] = best;
      o[3] = 1.0f / pool;
    }
}

// number of pixels of the x2 map where the constant-weight formulation differs (bitwise) from the float-index one
extern "C" int kp_host_bilinear_mismatches(const float* low, int in_size) {
  using namespace cm2;
  int bad = 0;
  for (int y = 0; y < 2 * in_size; ++y)
    for (int x = 0; x < 2 * in_size; ++x) {
      const float a = kp_bilinear_at(low, in_size, y, x), b = kp_bilinear2_at(low, in_size, y, x);
      if (!(a == b)) ++bad;
    }
  return bad;
}
