// Keypoint decode: the tail of KRCNNConvDeconvUpsampleHead.layers (bilinear x2, keypoint_head.py:221) fused with
// keypoint_rcnn_inference -> detectron2 heatmaps_to_keypoints [d2] (keypoint_head.py:95-120).
//
// One CTA per (keypoint, ROI slot).  The 2res x 2res low-resolution logits of that keypoint are gathered into shared
// memory, expanded x2 (bilinear) into a 4res x 4res map that also stays in shared memory, and the bicubic resize to
// the ROI's own ceil(h) x ceil(w) pixels is evaluated on the fly -- the resized maps (up to image size per ROI and
// keypoint) are never stored.  What leaves the SM is 16 bytes per (ROI, keypoint).
//
// Bound: fp32 issue, not HBM (algorithmic bytes are 4 * (2res)^2 in + 16 out per CTA).  WALK (default): the resized
// pixels are split into (column, row segment) items; a thread computes the x taps of its column once, keeps the x pass
// of the four source rows under the current resized row in registers, and spends the four y FMAs + a compare per
// pixel, with the y taps of every row in a shared table (kp_column_walk in kp_math.cuh).  The flat variant
// (CM2_KP_VARIANT=0) evaluates all 16 taps and both coefficient sets per pixel (~100 instructions); both build the same
// expression tree per pixel, so their results are bit-identical.
#include "common.cuh"
#include "kp_math.cuh"
#include <stdlib.h>

namespace cm2 {

constexpr int KP_THREADS = 256;

// Shared-space twin of KpMemPtr (kp_math.cuh): 32-bit shared addresses and ld.shared, so that the inner loop of the
// column walk does not rebuild generic addresses (ncu: ~10 of ~45 instructions per resized pixel were that).
struct KpMemShared {
  uint32_t hi, wtab, btab;
  __device__ __forceinline__ float hi_at(int elem) const {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(hi + 4u * (uint32_t)elem));
    return v;
  }
  __device__ __forceinline__ KpW4 w_at(int oy) const {
    KpW4 w;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(w.w0), "=f"(w.w1), "=f"(w.w2), "=f"(w.w3)
                 : "r"(wtab + 16u * (uint32_t)oy));
    return w;
  }
  __device__ __forceinline__ int base_at(int oy) const {
    int b;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(b) : "r"(btab + 4u * (uint32_t)oy));
    return b;
  }
};

template <bool WALK>
__global__ void __launch_bounds__(KP_THREADS, 4)
keypoints_decode_kernel(const float* __restrict__ lowres, const float* __restrict__ boxes, const int32_t* __restrict__ count,
                        int r_cap, int res, int k, int tab_rows, float* __restrict__ out) {
  extern __shared__ __align__(16) float kp_smem[];
  const int s_low = 2 * res, s_hi = 4 * res;
  float* low = kp_smem;                       // [s_low][s_low]
  float* hi = kp_smem + s_low * s_low;        // [s_hi][s_hi]
  __shared__ float red_v[KP_THREADS / 32];
  __shared__ long long red_p[KP_THREADS / 32];
  __shared__ float s_max;

  const int kp = blockIdx.x, slot = blockIdx.y;
  const int img = slot / r_cap;
  float* o = out + ((size_t)slot * k + kp) * 4;
  if (slot - img * r_cap >= count[img]) {     // empty ROI slot: defined output, no work
    if (threadIdx.x < 4) o[threadIdx.x] = 0.f;
    return;
  }
  // gather + bilinear x2, rows over warps and columns over lanes (no divisions by run-time extents)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* src = lowres + (size_t)slot * (res * res * 4) * k + kp;       // this ROI's [res][res][4][k] block
  for (int cy = warp; cy < res; cy += KP_THREADS / 32)
    for (int t = lane; t < res * 4; t += 32) {                               // t = (cell column, phase)
      const int cx = t >> 2, ph = t & 3;
      low[(2 * cy + (ph >> 1)) * s_low + 2 * cx + (ph & 1)] = __ldg(src + (size_t)(cy * res * 4 + t) * k);
    }
  __syncthreads();
  for (int y = warp; y < s_hi; y += KP_THREADS / 32)
    for (int x = lane; x < s_hi; x += 32) hi[y * s_hi + x] = kp_bilinear2_at(low, s_low, y, x);
  __syncthreads();

  const float4 b = __ldg(reinterpret_cast<const float4*>(boxes) + slot);
  const KpRoi roi = kp_roi(b.x, b.y, b.z, b.w);
  const float scale_y = (float)s_hi / (float)roi.hc, scale_x = (float)s_hi / (float)roi.wc;
  const long long total = (long long)roi.hc * roi.wc;

  // arg-max of the resized map; first index wins among equal values (torch.argmax on CPU)
  float best = -INFINITY;
  long long best_p = 0x7fffffffffffffffLL;
  if (WALK && kp_walk_applies(roi.hc, roi.wc, tab_rows)) {      // taller ROIs (> 1024 px at res 14) take the flat loop below
    KpW4* wtab = reinterpret_cast<KpW4*>(hi + s_hi * s_hi);               // [tab_rows]; 20 res^2 floats precede: 16-byte aligned
    int* btab = reinterpret_cast<int*>(wtab + tab_rows);                  // [tab_rows]
    for (int oy = threadIdx.x; oy < roi.hc; oy += KP_THREADS) {
      const KpRowTaps t = kp_row_taps(scale_y, oy, s_hi);
      wtab[oy] = t.w;
      btab[oy] = t.base;
    }
    __syncthreads();
    KpMemShared m;
    m.hi = (uint32_t)__cvta_generic_to_shared(hi);
    m.wtab = (uint32_t)__cvta_generic_to_shared(wtab);
    m.btab = (uint32_t)__cvta_generic_to_shared(btab);
    const KpBest b = kp_column_walk(m, s_hi, roi.hc, roi.wc, scale_x, (int)threadIdx.x, KP_THREADS);
    best = b.v;
    best_p = b.p;
  } else if (total <= 0x7fffffffLL) {
    const unsigned wc = (unsigned)roi.wc;
    for (unsigned p = threadIdx.x; p < (unsigned)total; p += KP_THREADS) {
      const unsigned oy = p / wc, ox = p - oy * wc;
      const KpCubic cy = kp_cubic_taps(scale_y, (int)oy, s_hi);
      const KpCubic cx = kp_cubic_taps(scale_x, (int)ox, s_hi);
      const float v = kp_bicubic_at(hi, s_hi, cy, cx);
      if (v > best) { best = v; best_p = p; }
    }
  } else {
    for (long long p = threadIdx.x; p < total; p += KP_THREADS) {
      const long long oy = p / roi.wc;
      const int ox = (int)(p - oy * roi.wc);
      const KpCubic cy = kp_cubic_taps(scale_y, (int)oy, s_hi);
      const KpCubic cx = kp_cubic_taps(scale_x, ox, s_hi);
      const float v = kp_bicubic_at(hi, s_hi, cy, cx);
      if (v > best) { best = v; best_p = p; }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, best, d);
    const long long p2 = __shfl_xor_sync(0xffffffffu, best_p, d);
    if (v2 > best || (v2 == best && p2 < best_p)) { best = v2; best_p = p2; }
  }
  if (lane == 0) { red_v[warp] = best; red_p[warp] = best_p; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < KP_THREADS / 32; ++w)
      if (red_v[w] > best || (red_v[w] == best && red_p[w] < best_p)) { best = red_v[w]; best_p = red_p[w]; }
    red_p[0] = best_p;
    s_max = best;
  }
  __syncthreads();
  const float mx = s_max;
  best_p = red_p[0];

  // score = exp(logit - max) / sum over the pool-resolution map of exp(map - max), logit == max
  float acc = 0.f;
  for (int i = threadIdx.x; i < s_hi * s_hi; i += KP_THREADS) acc += expf(hi[i] - mx);
  acc = warp_sum(acc);
  __syncthreads();                             // red_v is reused
  if (lane == 0) red_v[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float pool = 0.f;
    for (int w = 0; w < KP_THREADS / 32; ++w) pool += red_v[w];
    if (best_p == 0x7fffffffffffffffLL) best_p = 0;          // all-NaN map: torch would also return some index
    const long long y_int = best_p / roi.wc;
    const int x_int = (int)(best_p - y_int * roi.wc);
    // separate multiply and add, as the reference's tensor expressions do (no contraction)
    o[0] = __fadd_rn(__fmul_rn((float)x_int + 0.5f, __fdiv_rn(roi.w, (float)roi.wc)), roi.x0);
    o[1] = __fadd_rn(__fmul_rn((float)y_int + 0.5f, __fdiv_rn(roi.h, (float)roi.hc)), roi.y0);
    o[2] = mx;
    o[3] = __fdiv_rn(1.0f, pool);
  }
}

}  // namespace cm2

using namespace cm2;

extern "C" int cm2_keypoints_decode(const float* lowres, const float* boxes, const int32_t* det_count, int32_t n,
                                    int32_t r_cap, int32_t res, int32_t num_keypoints, float* out, void* stream) {
  CM2_CHECK_ARG(lowres && boxes && det_count && out, "keypoints_decode: null pointer");
  CM2_CHECK_ARG(n >= 0 && r_cap > 0 && num_keypoints > 0, "keypoints_decode: bad extents n=%d r_cap=%d k=%d", n, r_cap,
                num_keypoints);
  CM2_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "keypoints_decode: boxes must be 16-byte aligned");
  CM2_CHECK_ARG(res > 0 && res <= 24, "keypoints_decode: pooler resolution %d not in [1, 24]", res);
  CM2_CHECK_ARG((long long)n * r_cap <= 65535, "keypoints_decode: %lld ROI slots exceed the grid limit", (long long)n * r_cap);
  if (n == 0) return CM2_OK;
  const size_t maps = (size_t)20 * res * res * sizeof(float);           // (2res)^2 + (4res)^2 floats
  dim3 grid(num_keypoints, n * r_cap);
  const int variant = getenv("CM2_KP_VARIANT") ? atoi(getenv("CM2_KP_VARIANT")) : 1;
  if (variant == 1) {
    // the y-tap table takes what is left of the default 48 KB (static shared memory: ~112 bytes); taller ROIs compute
    // their y taps on the fly
    int tab_rows = (int)((48 * 1024 - 256 - maps) / 20);               // 16 bytes of weights + the source row base
    tab_rows = tab_rows > 1024 ? 1024 : tab_rows;           // 20 KB: six CTAs per SM at res 14
    keypoints_decode_kernel<true><<<grid, KP_THREADS, maps + (size_t)tab_rows * 20, (cudaStream_t)stream>>>(
        lowres, boxes, det_count, r_cap, res, num_keypoints, tab_rows, out);
  } else {
    keypoints_decode_kernel<false><<<grid, KP_THREADS, maps, (cudaStream_t)stream>>>(lowres, boxes, det_count, r_cap, res,
                                                                                    num_keypoints, 0, out);
  }
  CM2_CHECK_LAUNCH("keypoints_decode");
  return CM2_OK;
}
