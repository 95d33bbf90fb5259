// Bilinear resize of uint8 HWC images with Pillow's arithmetic (src/libImaging/Resample.c, 8 bits per channel):
// two separable passes on uint8 with 22-bit fixed-point coefficient tables (built on the host by
// centermask2_b200/transforms.py::pil_bilinear_coeffs, a restatement of precompute_coeffs + normalize_coeffs_8bpc).
// This is what detectron2's ResizeTransform / ResizeShortestEdge run on the CPU before the reference model
// (/root/reference/deploy_utils.py:60-73); bit-identical output, so everything downstream sees the same pixels.
#include "common.cuh"

namespace cm2 {

constexpr int RESIZE_PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ uint8_t resize_clip8(int v) {
  v >>= RESIZE_PRECISION_BITS;
  return (uint8_t)min(max(v, 0), 255);
}

// horizontal pass: in [h][w][c] -> tmp [h][ow][c]; one thread per (row, output column)
template <int C>
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ tmp, int h, int w,
                                                       int ow, const int* __restrict__ bounds, const int* __restrict__ kk,
                                                       int ksize) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (xx >= ow) return;
  const int x0 = __ldg(bounds + 2 * xx), n = __ldg(bounds + 2 * xx + 1);
  int acc[C];
#pragma unroll
  for (int ch = 0; ch < C; ++ch) acc[ch] = 1 << (RESIZE_PRECISION_BITS - 1);
  const uint8_t* src = in + ((size_t)y * w + x0) * C;
  for (int t = 0; t < n; ++t) {
    const int k = __ldg(kk + (size_t)xx * ksize + t);
#pragma unroll
    for (int ch = 0; ch < C; ++ch) acc[ch] += (int)src[t * C + ch] * k;
  }
  uint8_t* dst = tmp + ((size_t)y * ow + xx) * C;
#pragma unroll
  for (int ch = 0; ch < C; ++ch) dst[ch] = resize_clip8(acc[ch]);
}

// vertical pass: tmp [h][ow][c] -> out [oh][ow][c] (chw == 0) or [c][oh][ow] (chw == 1)
template <int C>
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ tmp, uint8_t* __restrict__ out, int ow, int oh,
                                                       const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                       int chw) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int yy = blockIdx.y;
  if (xx >= ow) return;
  const int y0 = __ldg(bounds + 2 * yy), n = __ldg(bounds + 2 * yy + 1);
  int acc[C];
#pragma unroll
  for (int ch = 0; ch < C; ++ch) acc[ch] = 1 << (RESIZE_PRECISION_BITS - 1);
  for (int t = 0; t < n; ++t) {
    const int k = __ldg(kk + (size_t)yy * ksize + t);
    const uint8_t* src = tmp + ((size_t)(y0 + t) * ow + xx) * C;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) acc[ch] += (int)src[ch] * k;
  }
#pragma unroll
  for (int ch = 0; ch < C; ++ch) {
    if (chw) out[((size_t)ch * oh + yy) * ow + xx] = resize_clip8(acc[ch]);
    else out[((size_t)yy * ow + xx) * C + ch] = resize_clip8(acc[ch]);
  }
}

}  // namespace cm2

using namespace cm2;

extern "C" int cm2_resize_pil_u8(const uint8_t* src, uint8_t* tmp, uint8_t* dst, int32_t h, int32_t w, int32_t c, int32_t oh,
                                 int32_t ow, const int32_t* bounds_x, const int32_t* kk_x, int32_t ksize_x,
                                 const int32_t* bounds_y, const int32_t* kk_y, int32_t ksize_y, int32_t chw, void* stream) {
  CM2_CHECK_ARG(src && tmp && dst && bounds_x && kk_x && bounds_y && kk_y, "resize_pil_u8: null pointer");
  CM2_CHECK_ARG(h > 0 && w > 0 && oh > 0 && ow > 0 && (c == 1 || c == 3 || c == 4) && ksize_x > 0 && ksize_y > 0 && h <= 65535 &&
                oh <= 65535, "resize_pil_u8: bad extents %dx%dx%d -> %dx%d", h, w, c, oh, ow);
  cudaStream_t s = (cudaStream_t)stream;
  dim3 gh(ceil_div(ow, 256), h), gv(ceil_div(ow, 256), oh);
#define CM2_RESIZE(C)                                                                                   \
  do {                                                                                                  \
    resize_h_kernel<C><<<gh, 256, 0, s>>>(src, tmp, h, w, ow, bounds_x, kk_x, ksize_x);                 \
    CM2_CHECK_LAUNCH("resize_h");                                                                       \
    resize_v_kernel<C><<<gv, 256, 0, s>>>(tmp, dst, ow, oh, bounds_y, kk_y, ksize_y, chw);              \
    CM2_CHECK_LAUNCH("resize_v");                                                                       \
  } while (0)
  if (c == 1) CM2_RESIZE(1);
  else if (c == 3) CM2_RESIZE(3);
  else CM2_RESIZE(4);
#undef CM2_RESIZE
  return CM2_OK;
}
