// Bandwidth-bound kernels on pitched NHWC views: input normalisation, max-pool, eSE, GroupNorm, ReLU.
// All are coalesced over the channel dimension (innermost) and vectorised 8 channels/thread.
// Reductions are deterministic (fixed two-stage trees, no atomics).
#include "common.cuh"
#include <cuda_fp16.h>
#include <algorithm>
#include <string.h>

namespace cm2 {

// ---------------------------------------------------------------------------------------------
// preprocess: CHW (u8 | f32) -> normalised, zero-padded HWC
// ---------------------------------------------------------------------------------------------
template <typename InT, typename OutT>
__global__ void preprocess_kernel(const InT* __restrict__ img, int h, int w, float m0, float m1, float m2,
                                  float s0, float s1, float s2, View<OutT> out, int b) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= out.w) return;
  OutT* q = out.at(b, y, x);
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  if (y < h && x < w) {
    size_t plane = (size_t)h * w;
    size_t o = (size_t)y * w + x;
    v0 = ((float)img[o] - m0) / s0;
    v1 = ((float)img[plane + o] - m1) / s1;
    v2 = ((float)img[2 * plane + o] - m2) / s2;
  }
  q[0] = from_f32<OutT>(v0);
  q[1] = from_f32<OutT>(v1);
  q[2] = from_f32<OutT>(v2);
  for (int c = 3; c < out.c; ++c) q[c] = from_f32<OutT>(0.f);
}

// ---------------------------------------------------------------------------------------------
// MaxPool 3x3 stride 2, ceil_mode, no padding (windows clipped at the bottom/right edge)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool3s2_kernel(View<const T> in, View<T> out) {
  const int c8 = in.c >> 3;
  int64_t total = (int64_t)out.n * out.h * out.w * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % c8);
    int64_t pix = i / c8;
    int ox = (int)(pix % out.w);
    int64_t t = pix / out.w;
    int oy = (int)(t % out.h);
    int b = (int)(t / out.h);
    float best[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) best[k] = -INFINITY;
    int y0 = oy * 2, x0 = ox * 2;
    for (int dy = 0; dy < 3; ++dy) {
      int y = y0 + dy;
      if (y >= in.h) break;
      for (int dx = 0; dx < 3; ++dx) {
        int x = x0 + dx;
        if (x >= in.w) break;
        float v[8];
        Vec8<T>::load(in.at(b, y, x) + cv * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) best[k] = fmaxf(best[k], v[k]);
      }
    }
    Vec8<T>::store(out.at(b, oy, ox) + cv * 8, best);
  }
}

// ---------------------------------------------------------------------------------------------
// Depthwise 3x3 convolution, padding 1, stride 1 or 2, no bias (vovnet.py:110-130 "dw_conv3x3": Conv2d(c, c, 3,
// groups=c); the norm / ReLU follow the pointwise 1x1 that comes next).  HBM-bound: one thread = 8 channels of TWO
// neighbouring output pixels (the 3x3 windows of a stride-1 pair share 6 of their 12 input vectors), weights
// [9][c] fp32 read through L1.  fp32 accumulation in tap order (ky, kx), one rounding on store.
// ---------------------------------------------------------------------------------------------
template <typename T, int STRIDE>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(View<const T> in, View<T> out, const float* __restrict__ w) {
  const int c8 = in.c >> 3, c = in.c;
  const int wpairs = (out.w + 1) >> 1;
  const int64_t total = (int64_t)out.n * out.h * wpairs * c8;
  constexpr int COLS = STRIDE == 1 ? 4 : 5;             // input columns under two neighbouring windows
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % c8);
    int64_t t = i / c8;
    const int op = (int)(t % wpairs);
    t /= wpairs;
    const int oy = (int)(t % out.h), b = (int)(t / out.h);
    const int ox = op * 2;
    float acc0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, acc1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = oy * STRIDE + ky - 1;
      if (y < 0 || y >= in.h) continue;
      float wk[3][8];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * c + cv * 8));
        const float4 d = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * c + cv * 8) + 1);
        wk[kx][0] = a.x; wk[kx][1] = a.y; wk[kx][2] = a.z; wk[kx][3] = a.w;
        wk[kx][4] = d.x; wk[kx][5] = d.y; wk[kx][6] = d.z; wk[kx][7] = d.w;
      }
#pragma unroll
      for (int j = 0; j < COLS; ++j) {
        const int x = ox * STRIDE + j - 1;
        if (x < 0 || x >= in.w) continue;
        float v[8];
        Vec8<T>::load(in.at(b, y, x) + cv * 8, v);
        if (j < 3) {                                      // window of output pixel ox: columns 0..2
#pragma unroll
          for (int k = 0; k < 8; ++k) acc0[k] = fmaf(v[k], wk[j][k], acc0[k]);
        }
        if (j >= STRIDE) {                                // window of output pixel ox + 1: columns STRIDE..STRIDE+2
#pragma unroll
          for (int k = 0; k < 8; ++k) acc1[k] = fmaf(v[k], wk[j - STRIDE][k], acc1[k]);
        }
      }
    }
    Vec8<T>::store(out.at(b, oy, ox) + cv * 8, acc0);
    if (ox + 1 < out.w) Vec8<T>::store(out.at(b, oy, ox + 1) + cv * 8, acc1);
  }
}

// ---------------------------------------------------------------------------------------------
// eSE
// ---------------------------------------------------------------------------------------------
constexpr int ESE_PIX_PER_CHUNK = 256;

// stage 1: grid (chunks, n), 256 threads.  Thread t owns the 8-channel column cv = t % c8 on pixel lane
// pl = t / c8 (lanes = 256 / c8); the lanes are combined through shared memory in a fixed order, so the
// result is deterministic.  (The first version walked 256 pixels serially per thread and ran at ~1/6 of HBM.)
template <typename T>
__global__ void __launch_bounds__(256) ese_pool_partial_kernel(View<const T> x, int chunks, float* __restrict__ ws) {
  __shared__ float sm[8 * 1024];                      // [lanes][c], lanes * c <= 8192 (c <= 1024)
  int chunk = blockIdx.x, b = blockIdx.y;
  int c = x.c, c8 = c >> 3;
  int hw = x.h * x.w;
  int lanes = blockDim.x / c8;
  if (lanes > 8) lanes = 8;
  int cv = threadIdx.x % c8, pl = threadIdx.x / c8;
  int p0 = chunk * ESE_PIX_PER_CHUNK;
  int p1 = min(p0 + ESE_PIX_PER_CHUNK, hw);
  if (pl < lanes) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = p0 + pl; p < p1; p += lanes) {
      int y = p / x.w, xx = p - y * x.w;
      float v[8];
      Vec8<T>::load(x.at(b, y, xx) + cv * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[pl * c + cv * 8 + k] = acc[k];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += sm[l * c + ch];
    ws[((size_t)b * chunks + chunk) * c + ch] = s;
  }
}

__global__ void ese_pool_final_kernel(const float* __restrict__ ws, int chunks, int c, float inv_hw,
                                      float* __restrict__ pooled) {
  int b = blockIdx.y;
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float acc = 0.f;
  for (int k = 0; k < chunks; ++k) acc += ws[((size_t)b * chunks + k) * c + ch];
  pooled[(size_t)b * c + ch] = acc * inv_hw;
}

// gate[b, o] = relu6(sum_i W[o,i]*pooled[b,i]*inv + bias[o] + 3) / 6 ; one warp per output channel.
__global__ void ese_gate_kernel(const float* __restrict__ pooled, float inv_count, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ gate, int c) {
  int b = blockIdx.y;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= c) return;
  const float* wr = w + (size_t)warp * c;
  const float* pv = pooled + (size_t)b * c;
  float acc = 0.f;
  for (int i = lane; i < c; i += 32) acc = fmaf(__ldg(wr + i), pv[i] * inv_count, acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + bias[warp] + 3.0f;
    v = fminf(fmaxf(v, 0.f), 6.f) / 6.0f;
    gate[(size_t)b * c + warp] = v;
  }
}

template <typename T>
__global__ void ese_apply_kernel(View<const T> x, const float* __restrict__ gate, View<const T> idn, View<T> out) {
  int c8 = x.c >> 3;
  int64_t total8 = (int64_t)x.n * x.h * x.w * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % c8);
    int64_t pix = i / c8;
    int xx = (int)(pix % x.w);
    int64_t t = pix / x.w;
    int y = (int)(t % x.h);
    int b = (int)(t / x.h);
    float v[8];
    Vec8<T>::load(x.at(b, y, xx) + cv * 8, v);
    const float* g = gate + (size_t)b * x.c + cv * 8;
    if (idn.p) {
      float r[8];
      Vec8<T>::load(idn.at(b, y, xx) + cv * 8, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = v[k] * g[k] + r[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = v[k] * g[k];
    }
    Vec8<T>::store(out.at(b, y, xx) + cv * 8, v);
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm (+ReLU), in place.  stage 1: per (chunk, sample) per-channel partial sum / sum of squares,
// stage 2: double-precision combine per (group, sample) -> mean, rstd; stage 3: normalise.
// ---------------------------------------------------------------------------------------------
constexpr int GN_PIX_PER_CHUNK = 128;

template <typename T>
__global__ void gn_partial_kernel(View<const T> x, int chunks, float* __restrict__ ws) {
  // grid (chunks, n), 256 threads.  Thread t owns the 8-channel column cv = t % c8 on pixel lane
  // pl = t / c8; lanes are combined through shared memory in a fixed order.
  __shared__ float sm[2048 * 2];          // [lanes][c][2], lanes * c == 2048 at most
  int chunk = blockIdx.x, b = blockIdx.y;
  int c = x.c, c8 = c >> 3;
  int hw = x.h * x.w;
  int lanes = blockDim.x / c8;
  int cv = threadIdx.x % c8, pl = threadIdx.x / c8;
  int p0 = chunk * GN_PIX_PER_CHUNK, p1 = min(p0 + GN_PIX_PER_CHUNK, hw);
  if (pl < lanes) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = p0 + pl; p < p1; p += lanes) {
      int y = p / x.w, xx = p - y * x.w;
      float v[8];
      Vec8<T>::load(x.at(b, y, xx) + cv * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sm[((size_t)pl * c + cv * 8 + k) * 2 + 0] = s[k];
      sm[((size_t)pl * c + cv * 8 + k) * 2 + 1] = q[k];
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int l = 0; l < lanes; ++l) { s += sm[((size_t)l * c + ch) * 2]; q += sm[((size_t)l * c + ch) * 2 + 1]; }
    ws[(((size_t)b * chunks + chunk) * c + ch) * 2 + 0] = s;
    ws[(((size_t)b * chunks + chunk) * c + ch) * 2 + 1] = q;
  }
}

// one warp per (group, sample): combines the per-chunk per-channel partials in double precision.
__global__ void gn_final_kernel(const float* __restrict__ ws, int hw, int c, int groups, int chunks, float eps,
                                float* __restrict__ stats) {
  int b = blockIdx.y, g = blockIdx.x;
  int cpg = c / groups;
  int lane = threadIdx.x;
  double sum = 0.0, sq = 0.0;
  int total = chunks * cpg;
  for (int i = lane; i < total; i += 32) {
    int chunk = i / cpg, ch = g * cpg + (i - chunk * cpg);
    const float* q = ws + (((size_t)b * chunks + chunk) * c + ch) * 2;
    sum += (double)q[0];
    sq += (double)q[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if (lane == 0) {
    double cnt = (double)hw * cpg;
    double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0) var = 0;
    stats[((size_t)b * groups + g) * 2 + 0] = (float)mean;
    stats[((size_t)b * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

template <typename T>
__global__ void gn_apply_kernel(View<T> x, int groups, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int relu) {
  int c8 = x.c >> 3, cpg = x.c / groups;
  int64_t total8 = (int64_t)x.n * x.h * x.w * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % c8);
    int64_t pix = i / c8;
    int xx = (int)(pix % x.w);
    int64_t t = pix / x.w;
    int y = (int)(t % x.h);
    int b = (int)(t / x.h);
    float v[8];
    T* ptr = x.at(b, y, xx) + cv * 8;
    Vec8<T>::load(ptr, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int g = (cv * 8 + k) / cpg;
      float mean = stats[((size_t)b * groups + g) * 2], rstd = stats[((size_t)b * groups + g) * 2 + 1];
      float yv = (v[k] - mean) * rstd * __ldg(gamma + cv * 8 + k) + __ldg(beta + cv * 8 + k);
      v[k] = relu ? fmaxf(yv, 0.f) : yv;
    }
    Vec8<T>::store(ptr, v);
  }
}

template <typename T>
__global__ void relu_kernel(View<const T> in, View<T> out) {
  int c8 = in.c >> 3;
  int64_t total8 = (int64_t)in.n * in.h * in.w * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % c8);
    int64_t pix = i / c8;
    int xx = (int)(pix % in.w);
    int64_t t = pix / in.w;
    int y = (int)(t % in.h);
    int b = (int)(t / in.h);
    float v[8];
    Vec8<T>::load(in.at(b, y, xx) + cv * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
    Vec8<T>::store(out.at(b, y, xx) + cv * 8, v);
  }
}

// phase-split copy: pixel (y, x) -> plane (y&1)*2 + (x&1) at (y>>1, x>>1), optional ReLU
template <typename T>
__global__ void phase_split_kernel(View<const T> in, View<T> out, long long plane_stride, int relu) {
  int c8 = in.c >> 3;
  int64_t total8 = (int64_t)in.n * in.h * in.w * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % c8);
    int64_t pix = i / c8;
    int xx = (int)(pix % in.w);
    int64_t t = pix / in.w;
    int y = (int)(t % in.h);
    int b = (int)(t / in.h);
    float v[8];
    Vec8<T>::load(in.at(b, y, xx) + cv * 8, v);
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
    }
    Vec8<T>::store(out.at(b, y >> 1, xx >> 1) + ((y & 1) * 2 + (xx & 1)) * plane_stride + cv * 8, v);
  }
}

// fused normalise + pad + im2col (3x3, stride 2, pad 1, Cin 3) -> 32 bf16 channels per output pixel
template <typename InT>
__global__ void preprocess_im2col_kernel(const InT* __restrict__ img, int h, int w, int ho, int wo, float m0, float m1,
                                         float m2, float r0, float r1, float r2, View<__nv_bfloat16> out, int b) {
  int ox = blockIdx.x * blockDim.x + threadIdx.x;
  int oy = blockIdx.y;
  if (ox >= wo) return;
  const size_t plane = (size_t)h * w;
  const float mean[3] = {m0, m1, m2}, rstd[3] = {r0, r1, r2};
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    int iy = 2 * oy + ky - 1;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      int ix = 2 * ox + kx - 1;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
        size_t o = (size_t)iy * w + ix;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = ((float)img[c * plane + o] - mean[c]) / rstd[c];
      }
    }
  }
  __nv_bfloat16* q = out.at(b, oy, ox);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float t8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t8[k] = v[j * 8 + k];
    Vec8<__nv_bfloat16>::store(q + j * 8, t8);
  }
}

// The same for a whole batch in one launch: grid (pixel-chunks, 1, image).  One thread per (output pixel, 8-channel
// chunk): 8 byte loads and ONE 16-byte store, consecutive threads write consecutive 16-byte pieces (a warp stores
// 512 contiguous bytes).  UNIT_STD skips the division when std == 1 (x / 1.0f is exact, so the result is unchanged).
struct Im2colBatch {
  const void* img[CM2_MAX_BATCH_PTRS];
  int h[CM2_MAX_BATCH_PTRS], w[CM2_MAX_BATCH_PTRS];
};

constexpr int IM2COL_PIX_PER_BLOCK = 1024;

template <typename InT, bool UNIT_STD, int CHUNK>
__device__ __forceinline__ void im2col_chunk(const InT* __restrict__ img, int h, int w, size_t plane, int oy, int ox, float m0,
                                             float m1, float m2, float r0, float r1, float r2, __nv_bfloat16* dst) {
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    constexpr int dummy = 0;
    (void)dummy;
    const int e = CHUNK * 8 + k;                       // compile-time: element (ky, kx, c) of the 27 (+5 zero) channels
    const int tap = e / 3, c = e - tap * 3;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
    float val = 0.f;
    if (e < 27 && iy >= 0 && iy < h && ix >= 0 && ix < w) {
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
      val = (float)img[c * plane + (size_t)iy * w + ix] - mean;
      if (!UNIT_STD) val = val / (c == 0 ? r0 : (c == 1 ? r1 : r2));
    }
    v[k] = val;
  }
  Vec8<__nv_bfloat16>::store(dst + CHUNK * 8, v);
}

template <typename InT, bool UNIT_STD>
__global__ void __launch_bounds__(256) preprocess_im2col_batch_kernel(Im2colBatch bt, int ho, int wo, float m0, float m1, float m2,
                                                                      float r0, float r1, float r2, View<__nv_bfloat16> out, int b0) {
  const int b = blockIdx.z;
  const InT* __restrict__ img = reinterpret_cast<const InT*>(bt.img[b]);
  const int h = bt.h[b], w = bt.w[b];
  const size_t plane = (size_t)h * w;
  // a block covers IM2COL_PIX_PER_BLOCK consecutive output pixels, 64 per pass (67 200 tiny blocks per image batch were
  // block-scheduling bound)
#pragma unroll 1
  for (int k = 0; k < IM2COL_PIX_PER_BLOCK / 64; ++k) {
    const unsigned pix = blockIdx.x * IM2COL_PIX_PER_BLOCK + k * 64 + (threadIdx.x >> 2);
    if (pix >= (unsigned)(ho * wo)) return;
    const int oy = pix / (unsigned)wo, ox = pix - oy * wo;
    __nv_bfloat16* dst = out.at(b0 + b, oy, ox);
    switch (threadIdx.x & 3) {                          // the chunk index is a template argument: all tap / channel arithmetic folds
      case 0: im2col_chunk<InT, UNIT_STD, 0>(img, h, w, plane, oy, ox, m0, m1, m2, r0, r1, r2, dst); break;
      case 1: im2col_chunk<InT, UNIT_STD, 1>(img, h, w, plane, oy, ox, m0, m1, m2, r0, r1, r2, dst); break;
      case 2: im2col_chunk<InT, UNIT_STD, 2>(img, h, w, plane, oy, ox, m0, m1, m2, r0, r1, r2, dst); break;
      default: im2col_chunk<InT, UNIT_STD, 3>(img, h, w, plane, oy, ox, m0, m1, m2, r0, r1, r2, dst); break;
    }
  }
}

// uint8 fast path: grid (x-tiles, output rows, images).  The block first stages the 3 planes x 3 input rows x 513 columns
// its 256 output pixels read into shared memory with coalesced byte loads (the direct version issued 8 scattered byte
// loads per thread and was L1-transaction bound at ~1 TB/s), then every thread assembles 16-byte pieces from it.
constexpr int IM2COL_TILE = 256;
template <bool UNIT_STD, int CHUNK>
__device__ __forceinline__ void im2col_chunk_smem(const uint8_t (*tile)[2 * IM2COL_TILE + 8], int h, int w, int oy, int ox, int lx,
                                                  float m0, float m1, float m2, float r0, float r1, float r2, __nv_bfloat16* dst) {
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = CHUNK * 8 + k;
    const int tap = e / 3, c = e - tap * 3;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
    float val = 0.f;
    if (e < 27 && iy >= 0 && iy < h && ix >= 0 && ix < w) {
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
      val = (float)tile[(c < 3 ? c : 0) * 3 + (ky < 3 ? ky : 0)][2 * lx + kx] - mean;
      if (!UNIT_STD) val = val / (c == 0 ? r0 : (c == 1 ? r1 : r2));
    }
    v[k] = val;
  }
  Vec8<__nv_bfloat16>::store(dst + CHUNK * 8, v);
}

template <bool UNIT_STD>
__global__ void __launch_bounds__(256) preprocess_im2col_u8_tiled_kernel(Im2colBatch bt, int ho, int wo, float m0, float m1, float m2,
                                                                         float r0, float r1, float r2, View<__nv_bfloat16> out, int b0) {
  __shared__ uint8_t tile[9][2 * IM2COL_TILE + 8];
  const int b = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * IM2COL_TILE;
  const uint8_t* __restrict__ img = reinterpret_cast<const uint8_t*>(bt.img[b]);
  const int h = bt.h[b], w = bt.w[b];
  const size_t plane = (size_t)h * w;
  constexpr int COLS = 2 * IM2COL_TILE + 1;
  constexpr int ITERS = (9 * COLS + 255) / 256;
  // all loads of a thread are issued before the first shared-memory store (a load -> store loop pays one L2 latency per
  // iteration: 19 x ~700 cycles per block)
  uint8_t stage[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int i = it * 256 + threadIdx.x;
    const int pr = i / COLS, col = i - pr * COLS;
    const int c = pr / 3, ky = pr - c * 3;
    const int iy = 2 * oy + ky - 1, ix = 2 * ox0 - 1 + col;
    stage[it] = 0;
    if (i < 9 * COLS && iy >= 0 && iy < h && ix >= 0 && ix < w) stage[it] = __ldg(img + c * plane + (size_t)iy * w + ix);
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int i = it * 256 + threadIdx.x;
    const int pr = i / COLS, col = i - pr * COLS;
    if (i < 9 * COLS) tile[pr][col] = stage[it];
  }
  __syncthreads();
#pragma unroll 1
  for (int k = 0; k < IM2COL_TILE / 64; ++k) {
    const int lx = k * 64 + (threadIdx.x >> 2);
    const int ox = ox0 + lx;
    if (ox >= wo) break;
    __nv_bfloat16* dst = out.at(b0 + b, oy, ox);
    switch (threadIdx.x & 3) {
      case 0: im2col_chunk_smem<UNIT_STD, 0>(tile, h, w, oy, ox, lx, m0, m1, m2, r0, r1, r2, dst); break;
      case 1: im2col_chunk_smem<UNIT_STD, 1>(tile, h, w, oy, ox, lx, m0, m1, m2, r0, r1, r2, dst); break;
      case 2: im2col_chunk_smem<UNIT_STD, 2>(tile, h, w, oy, ox, lx, m0, m1, m2, r0, r1, r2, dst); break;
      default: im2col_chunk_smem<UNIT_STD, 3>(tile, h, w, oy, ox, lx, m0, m1, m2, r0, r1, r2, dst); break;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm on a segmented halo tensor (all FPN levels of an FCOS tower in one launch set).
// ---------------------------------------------------------------------------------------------
struct GnSegs {
  int num;
  long long row0[CM2_MAX_SEG];
  int n[CM2_MAX_SEG], h[CM2_MAX_SEG], w[CM2_MAX_SEG];
  int chunk_prefix[CM2_MAX_SEG + 1];      // chunks (GN_PIX_PER_CHUNK interior pixels of one image) before segment s
  int img_prefix[CM2_MAX_SEG + 1];        // images before segment s
  long long vec_prefix[CM2_MAX_SEG + 1];  // 8-channel vectors of interior pixels before segment s
};

template <typename T>
__device__ __forceinline__ T* gn_seg_pixel(T* base, const GnSegs& g, int s, int b, int y, int x, int c) {
  long long row = g.row0[s] + (long long)b * (g.h[s] + 2) * (g.w[s] + 2) + (long long)(y + 1) * (g.w[s] + 2) + (x + 1);
  return base + row * c;
}

template <typename T>
__global__ void gn_seg_partial_kernel(const T* __restrict__ x, int c, GnSegs g, float* __restrict__ ws) {
  __shared__ float sm[2048 * 2];
  int s = 0;
  for (int i = 1; i < g.num; ++i)
    if ((int)blockIdx.x >= g.chunk_prefix[i]) s = i;
  const int hw = g.h[s] * g.w[s], w = g.w[s];
  const int cpi = (hw + GN_PIX_PER_CHUNK - 1) / GN_PIX_PER_CHUNK;
  const int local = blockIdx.x - g.chunk_prefix[s];
  const int b = local / cpi, chunk = local - b * cpi;
  const int c8 = c >> 3;
  const int lanes = blockDim.x / c8;
  const int cv = threadIdx.x % c8, pl = threadIdx.x / c8;
  const int p0 = chunk * GN_PIX_PER_CHUNK, p1 = min(p0 + GN_PIX_PER_CHUNK, hw);
  if (pl < lanes) {
    float sa[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = p0 + pl; p < p1; p += lanes) {
      int y = p / w, xx = p - y * w;
      float v[8];
      Vec8<T>::load(gn_seg_pixel(x, g, s, b, y, xx, c) + cv * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) { sa[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sm[((size_t)pl * c + cv * 8 + k) * 2 + 0] = sa[k];
      sm[((size_t)pl * c + cv * 8 + k) * 2 + 1] = q[k];
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int l = 0; l < lanes; ++l) { a += sm[((size_t)l * c + ch) * 2]; q += sm[((size_t)l * c + ch) * 2 + 1]; }
    ws[((size_t)blockIdx.x * c + ch) * 2 + 0] = a;
    ws[((size_t)blockIdx.x * c + ch) * 2 + 1] = q;
  }
}

// grid (groups, total images): one warp combines the chunk partials of one (segment, image, group)
__global__ void gn_seg_final_kernel(const float* __restrict__ ws, int c, int groups, GnSegs g, float eps,
                                    float* __restrict__ stats) {
  const int gi = blockIdx.y, grp = blockIdx.x;
  int s = 0;
  for (int i = 1; i < g.num; ++i)
    if (gi >= g.img_prefix[i]) s = i;
  const int b = gi - g.img_prefix[s];
  const int hw = g.h[s] * g.w[s];
  const int cpi = (hw + GN_PIX_PER_CHUNK - 1) / GN_PIX_PER_CHUNK;
  const int chunk0 = g.chunk_prefix[s] + b * cpi;
  const int cpg = c / groups;
  const int lane = threadIdx.x;
  double sum = 0.0, sq = 0.0;
  const int total = cpi * cpg;
  for (int i = lane; i < total; i += 32) {
    int chunk = i / cpg, ch = grp * cpg + (i - chunk * cpg);
    const float* q = ws + ((size_t)(chunk0 + chunk) * c + ch) * 2;
    sum += (double)q[0];
    sq += (double)q[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if (lane == 0) {
    double cnt = (double)hw * cpg;
    double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0) var = 0;
    stats[((size_t)gi * groups + grp) * 2 + 0] = (float)mean;
    stats[((size_t)gi * groups + grp) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

template <typename T>
__global__ void gn_seg_apply_kernel(T* __restrict__ x, int c, int groups, GnSegs g, const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, int relu) {
  const int c8 = c >> 3, cpg = c / groups;
  const long long total8 = g.vec_prefix[g.num];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    int s = 0;
    for (int k = 1; k < g.num; ++k)
      if (i >= g.vec_prefix[k]) s = k;
    long long li = i - g.vec_prefix[s];
    int cv = (int)(li % c8);
    long long pix = li / c8;
    int xx = (int)(pix % g.w[s]);
    long long t = pix / g.w[s];
    int y = (int)(t % g.h[s]);
    int b = (int)(t / g.h[s]);
    const int gi = g.img_prefix[s] + b;
    float v[8];
    T* ptr = gn_seg_pixel(x, g, s, b, y, xx, c) + cv * 8;
    Vec8<T>::load(ptr, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int grp = (cv * 8 + k) / cpg;
      float mean = stats[((size_t)gi * groups + grp) * 2], rstd = stats[((size_t)gi * groups + grp) * 2 + 1];
      float yv = (v[k] - mean) * rstd * __ldg(gamma + cv * 8 + k) + __ldg(beta + cv * 8 + k);
      v[k] = relu ? fmaxf(yv, 0.f) : yv;
    }
    Vec8<T>::store(ptr, v);
  }
}

static bool gn_make_segs(int32_t num_seg, const cm2_seg* seg, int c, GnSegs* g) {
  if (num_seg < 1 || num_seg > CM2_MAX_SEG) return false;
  g->num = num_seg;
  g->chunk_prefix[0] = 0; g->img_prefix[0] = 0; g->vec_prefix[0] = 0;
  for (int i = 0; i < num_seg; ++i) {
    if (seg[i].n <= 0 || seg[i].h <= 0 || seg[i].w <= 0 || seg[i].halo != 0) return false;    // (own-frame segments only)
    g->row0[i] = seg[i].row0; g->n[i] = seg[i].n; g->h[i] = seg[i].h; g->w[i] = seg[i].w;
    int cpi = ceil_div(seg[i].h * seg[i].w, GN_PIX_PER_CHUNK);
    g->chunk_prefix[i + 1] = g->chunk_prefix[i] + cpi * seg[i].n;
    g->img_prefix[i + 1] = g->img_prefix[i] + seg[i].n;
    g->vec_prefix[i + 1] = g->vec_prefix[i] + (long long)seg[i].n * seg[i].h * seg[i].w * (c / 8);
  }
  return true;
}

// Split-precision operand preparation (cm2.h: cm2_split_f16x2): one thread = 8 channels of one pixel; reads 32 bytes,
// writes the 16-byte hi piece to [pixel][ch] and the 16-byte lo piece to [pixel][c + ch].  hi = half(x) (round to nearest
// even), lo = half(x - float(hi)): x - hi is exact in fp32 (Sterbenz / the difference has at most 13 significant bits left
// of x's 24), so hi + lo carries 22 significant bits of x; below 2^-14 the lo part is a half subnormal with 2^-24
// absolute resolution.
__global__ void __launch_bounds__(256) split_f16x2_kernel(const float* __restrict__ x, __half* __restrict__ out, long long total8,
                                                           int c8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / c8;
    const int ch = (int)(i - pix * c8) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + i * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + i * 8) + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    __half* o = out + pix * (2ll * c8 * 8) + ch;
    *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(o + c8 * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

int grid_for(int64_t work, int block) {
  int64_t g = ceil_div64(work, block);
  int64_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace cm2

using namespace cm2;

#define CM2_CHECK_DTYPE(dtype, name) \
  CM2_CHECK_ARG((dtype) == CM2_F32 || (dtype) == CM2_BF16, name ": dtype %d not supported", (int)(dtype))

extern "C" int cm2_preprocess_image(const void* img, int32_t in_dtype, int32_t h, int32_t w, const float* mean3,
                                    const float* std3, const cm2_act* out, int32_t out_dtype, int32_t out_index,
                                    void* stream) {
  CM2_CHECK_ARG(img && out && out->data && mean3 && std3, "preprocess: null pointer");
  CM2_CHECK_ARG(h > 0 && w > 0 && out->h >= h && out->w >= w && out->c >= 3 && out_index >= 0 && out_index < out->n,
                "preprocess: bad extents %dx%d -> %dx%d (image %d of %d)", h, w, out->h, out->w, out_index, out->n);
  dim3 grid(ceil_div(out->w, 256), out->h);
  cudaStream_t s = (cudaStream_t)stream;
#define CM2_PRE(IN, OUT)                                                                                   \
  preprocess_kernel<IN, OUT><<<grid, 256, 0, s>>>((const IN*)img, h, w, mean3[0], mean3[1], mean3[2], std3[0], \
                                                  std3[1], std3[2], make_view<OUT>(*out), out_index)
  if (in_dtype == CM2_F32 && out_dtype == CM2_F32) CM2_PRE(float, float);
  else if (in_dtype == CM2_F32 && out_dtype == CM2_BF16) CM2_PRE(float, __nv_bfloat16);
  else if (in_dtype == CM2_U8 && out_dtype == CM2_F32) CM2_PRE(uint8_t, float);
  else if (in_dtype == CM2_U8 && out_dtype == CM2_BF16) CM2_PRE(uint8_t, __nv_bfloat16);
  else { set_error("preprocess: unsupported dtypes %d -> %d", in_dtype, out_dtype); return CM2_ERR_UNSUPPORTED; }
#undef CM2_PRE
  CM2_CHECK_LAUNCH("preprocess");
  return CM2_OK;
}

extern "C" int cm2_preprocess_im2col(const void* img, int32_t in_dtype, int32_t h, int32_t w, int32_t hp, int32_t wp,
                                     const float* mean3, const float* std3, const cm2_act* out, int32_t out_index,
                                     void* stream) {
  CM2_CHECK_ARG(img && out && out->data && mean3 && std3, "preprocess_im2col: null pointer");
  CM2_CHECK_ARG(h > 0 && w > 0 && hp >= h && wp >= w && hp % 2 == 0 && wp % 2 == 0, "preprocess_im2col: bad extents");
  CM2_CHECK_ARG(out->h == hp / 2 && out->w == wp / 2 && out->c == 32 && out_index >= 0 && out_index < out->n &&
                vec8_ok(*out, 2), "preprocess_im2col: out view [%d,%d,%d,%d] != [n,%d,%d,32]", out->n, out->h, out->w,
                out->c, hp / 2, wp / 2);
  dim3 grid(ceil_div(out->w, 128), out->h);
  cudaStream_t s = (cudaStream_t)stream;
  // (x - mean) / std is computed with a true division, exactly as cm2_preprocess_image does
  if (in_dtype == CM2_F32)
    preprocess_im2col_kernel<float><<<grid, 128, 0, s>>>((const float*)img, h, w, out->h, out->w, mean3[0], mean3[1],
                                                         mean3[2], std3[0], std3[1], std3[2],
                                                         make_view<__nv_bfloat16>(*out), out_index);
  else if (in_dtype == CM2_U8)
    preprocess_im2col_kernel<uint8_t><<<grid, 128, 0, s>>>((const uint8_t*)img, h, w, out->h, out->w, mean3[0], mean3[1],
                                                           mean3[2], std3[0], std3[1], std3[2],
                                                           make_view<__nv_bfloat16>(*out), out_index);
  else { set_error("preprocess_im2col: unsupported input dtype %d", in_dtype); return CM2_ERR_UNSUPPORTED; }
  CM2_CHECK_LAUNCH("preprocess_im2col");
  return CM2_OK;
}

extern "C" int cm2_preprocess_im2col_batch(const void* const* imgs, const int32_t* hs, const int32_t* ws, int32_t n, int32_t in_dtype,
                                           int32_t hp, int32_t wp, const float* mean3, const float* std3, const cm2_act* out,
                                           int32_t out_index0, void* stream) {
  CM2_CHECK_ARG(imgs && hs && ws && out && out->data && mean3 && std3, "preprocess_im2col_batch: null pointer");
  CM2_CHECK_ARG(n >= 0 && hp > 0 && wp > 0 && hp % 2 == 0 && wp % 2 == 0, "preprocess_im2col_batch: bad extents");
  CM2_CHECK_ARG(out->h == hp / 2 && out->w == wp / 2 && out->c == 32 && out_index0 >= 0 && out_index0 + n <= out->n &&
                vec8_ok(*out, 2), "preprocess_im2col_batch: out view [%d,%d,%d,%d] != [>=%d,%d,%d,32]", out->n, out->h, out->w,
                out->c, out_index0 + n, hp / 2, wp / 2);
  CM2_CHECK_ARG(in_dtype == CM2_F32 || in_dtype == CM2_U8, "preprocess_im2col_batch: unsupported input dtype %d", in_dtype);
  cudaStream_t s = (cudaStream_t)stream;
  const bool unit = std3[0] == 1.f && std3[1] == 1.f && std3[2] == 1.f;
  for (int i0 = 0; i0 < n; i0 += CM2_MAX_BATCH_PTRS) {
    const int nb = std::min(n - i0, (int)CM2_MAX_BATCH_PTRS);
    Im2colBatch bt;
    memset(&bt, 0, sizeof(bt));
    for (int i = 0; i < nb; ++i) {
      CM2_CHECK_ARG(imgs[i0 + i] && hs[i0 + i] > 0 && ws[i0 + i] > 0 && hs[i0 + i] <= hp && ws[i0 + i] <= wp,
                    "preprocess_im2col_batch: image %d is %dx%d, padded extent %dx%d", i0 + i, hs[i0 + i], ws[i0 + i], hp, wp);
      bt.img[i] = imgs[i0 + i]; bt.h[i] = hs[i0 + i]; bt.w[i] = ws[i0 + i];
    }
    dim3 grid(ceil_div(out->h * out->w, IM2COL_PIX_PER_BLOCK), 1, nb);
    View<__nv_bfloat16> ov = make_view<__nv_bfloat16>(*out);
#define CM2_IM2COL(T, U) preprocess_im2col_batch_kernel<T, U><<<grid, 256, 0, s>>>(bt, out->h, out->w, mean3[0], mean3[1], mean3[2], \
                                                                                   std3[0], std3[1], std3[2], ov, out_index0 + i0)
    if (in_dtype == CM2_F32) { if (unit) CM2_IM2COL(float, true); else CM2_IM2COL(float, false); }
    else {
      dim3 tgrid(ceil_div(out->w, IM2COL_TILE), out->h, nb);
      if (unit) preprocess_im2col_u8_tiled_kernel<true><<<tgrid, 256, 0, s>>>(bt, out->h, out->w, mean3[0], mean3[1], mean3[2], std3[0],
                                                                              std3[1], std3[2], ov, out_index0 + i0);
      else preprocess_im2col_u8_tiled_kernel<false><<<tgrid, 256, 0, s>>>(bt, out->h, out->w, mean3[0], mean3[1], mean3[2], std3[0],
                                                                           std3[1], std3[2], ov, out_index0 + i0);
    }
#undef CM2_IM2COL
    CM2_CHECK_LAUNCH("preprocess_im2col_batch");
  }
  return CM2_OK;
}

extern "C" int cm2_phase_split(const cm2_act* in, const cm2_act* out_plane0, int32_t dtype, int32_t relu, void* stream) {
  CM2_CHECK_ARG(in && out_plane0 && in->data && out_plane0->data, "phase_split: null pointer");
  CM2_CHECK_DTYPE(dtype, "phase_split");
  int eb = elem_bytes(dtype);
  CM2_CHECK_ARG(vec8_ok(*in, eb) && vec8_ok(*out_plane0, eb), "phase_split: channels/strides must be multiples of 8");
  CM2_CHECK_ARG(out_plane0->n == in->n && out_plane0->c == in->c && out_plane0->h == (in->h + 1) / 2 &&
                out_plane0->w == (in->w + 1) / 2, "phase_split: plane view [%d,%d,%d,%d] does not match input [%d,%d,%d,%d]",
                out_plane0->n, out_plane0->h, out_plane0->w, out_plane0->c, in->n, in->h, in->w, in->c);
  int64_t total8 = (int64_t)in->n * in->h * in->w * (in->c / 8);
  if (total8 == 0) return CM2_OK;
  long long ps = (long long)out_plane0->n * out_plane0->sn;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CM2_F32)
    phase_split_kernel<float><<<grid_for(total8, 256), 256, 0, s>>>(make_view<const float>(*in),
                                                                   make_view<float>(*out_plane0), ps, relu);
  else
    phase_split_kernel<__nv_bfloat16><<<grid_for(total8, 256), 256, 0, s>>>(make_view<const __nv_bfloat16>(*in),
                                                                          make_view<__nv_bfloat16>(*out_plane0), ps, relu);
  CM2_CHECK_LAUNCH("phase_split");
  return CM2_OK;
}

extern "C" int cm2_maxpool3x3s2_ceil(const cm2_act* in, const cm2_act* out, int32_t dtype, void* stream) {
  CM2_CHECK_ARG(in && out && in->data && out->data, "maxpool: null pointer");
  CM2_CHECK_DTYPE(dtype, "maxpool");
  CM2_CHECK_ARG(vec8_ok(*in, elem_bytes(dtype)) && vec8_ok(*out, elem_bytes(dtype)),
                "maxpool: channels/strides must be multiples of 8 (c=%d)", in->c);
  int h = in->h, w = in->w;
  int eh = (h - 3 + 1) / 2 + 1, ew = (w - 3 + 1) / 2 + 1;   // ceil((h-3)/2)+1
  if ((eh - 1) * 2 >= h) --eh;                               // last window must start inside the input
  if ((ew - 1) * 2 >= w) --ew;
  CM2_CHECK_ARG(out->h == eh && out->w == ew && out->n == in->n && out->c == in->c,
                "maxpool: out [%d,%d,%d,%d] != expected [%d,%d,%d,%d]", out->n, out->h, out->w, out->c, in->n, eh, ew,
                in->c);
  int64_t total = (int64_t)out->n * out->h * out->w * (in->c / 8);
  if (total == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CM2_F32)
    maxpool3s2_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(make_view<const float>(*in), make_view<float>(*out));
  else
    maxpool3s2_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>(make_view<const __nv_bfloat16>(*in),
                                                                        make_view<__nv_bfloat16>(*out));
  CM2_CHECK_LAUNCH("maxpool3x3s2");
  return CM2_OK;
}

extern "C" int cm2_dwconv3x3(const cm2_act* in, const cm2_act* out, int32_t dtype, const float* w, int32_t stride,
                             void* stream) {
  CM2_CHECK_ARG(in && out && in->data && out->data && w, "dwconv3x3: null pointer");
  CM2_CHECK_DTYPE(dtype, "dwconv3x3");
  CM2_CHECK_ARG(stride == 1 || stride == 2, "dwconv3x3: stride %d", stride);
  CM2_CHECK_ARG(in->n == out->n && in->c == out->c && out->h == (in->h - 1) / stride + 1 && out->w == (in->w - 1) / stride + 1,
                "dwconv3x3: in [%d,%d,%d,%d] / out [%d,%d,%d,%d] do not match stride %d", in->n, in->h, in->w, in->c, out->n,
                out->h, out->w, out->c, stride);
  CM2_CHECK_ARG(vec8_ok(*in, elem_bytes(dtype)) && vec8_ok(*out, elem_bytes(dtype)) && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                "dwconv3x3: channels / strides must be multiples of 8, weights 16-byte aligned");
  const int64_t total = (int64_t)out->n * out->h * ((out->w + 1) / 2) * (out->c / 8);
  if (total == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
#define CM2_DW(T, S) dwconv3x3_kernel<T, S><<<grid_for(total, 256), 256, 0, s>>>(make_view<const T>(*in), make_view<T>(*out), w)
  if (dtype == CM2_F32) { if (stride == 1) CM2_DW(float, 1); else CM2_DW(float, 2); }
  else { if (stride == 1) CM2_DW(__nv_bfloat16, 1); else CM2_DW(__nv_bfloat16, 2); }
#undef CM2_DW
  CM2_CHECK_LAUNCH("dwconv3x3");
  return CM2_OK;
}

extern "C" int32_t cm2_ese_pool_chunks(int32_t hw) { return ceil_div(hw, ESE_PIX_PER_CHUNK); }

extern "C" int cm2_ese_pool(const cm2_act* x, int32_t dtype, float* workspace, float* pooled, void* stream) {
  CM2_CHECK_ARG(x && x->data && workspace && pooled, "ese_pool: null pointer");
  CM2_CHECK_DTYPE(dtype, "ese_pool");
  CM2_CHECK_ARG(vec8_ok(*x, elem_bytes(dtype)) && x->h > 0 && x->w > 0 && x->c <= 1024 && x->c / 8 <= 256,
                "ese_pool: bad shape %dx%d c=%d (need c %% 8 == 0, c <= 1024)", x->h, x->w, x->c);
  if (x->n == 0) return CM2_OK;
  int hw = x->h * x->w;
  int chunks = cm2_ese_pool_chunks(hw);
  cudaStream_t s = (cudaStream_t)stream;
  dim3 g1(chunks, x->n);
  if (dtype == CM2_F32)
    ese_pool_partial_kernel<float><<<g1, 256, 0, s>>>(make_view<const float>(*x), chunks, workspace);
  else
    ese_pool_partial_kernel<__nv_bfloat16><<<g1, 256, 0, s>>>(make_view<const __nv_bfloat16>(*x), chunks, workspace);
  CM2_CHECK_LAUNCH("ese_pool_partial");
  dim3 g2(ceil_div(x->c, 128), x->n);
  ese_pool_final_kernel<<<g2, 128, 0, s>>>(workspace, chunks, x->c, 1.0f / (float)hw, pooled);
  CM2_CHECK_LAUNCH("ese_pool_final");
  return CM2_OK;
}

extern "C" int cm2_ese_gate(const float* pooled, float inv_count, const float* fc_w, const float* fc_b, float* gate,
                            int32_t n, int32_t c, void* stream) {
  CM2_CHECK_ARG(pooled && fc_w && fc_b && gate, "ese_gate: null pointer");
  if (n == 0) return CM2_OK;
  dim3 grid(ceil_div(c * 32, 256), n);
  ese_gate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pooled, inv_count, fc_w, fc_b, gate, c);
  CM2_CHECK_LAUNCH("ese_gate");
  return CM2_OK;
}

extern "C" int cm2_ese_apply(const cm2_act* x, const float* gate, const cm2_act* identity, const cm2_act* out,
                             int32_t dtype, void* stream) {
  CM2_CHECK_ARG(x && x->data && gate && out && out->data, "ese_apply: null pointer");
  CM2_CHECK_DTYPE(dtype, "ese_apply");
  int eb = elem_bytes(dtype);
  CM2_CHECK_ARG(vec8_ok(*x, eb) && vec8_ok(*out, eb) && same_extent(*x, *out), "ese_apply: bad views (c=%d)", x->c);
  cm2_act idn;
  memset(&idn, 0, sizeof(idn));
  if (identity && identity->data) {
    CM2_CHECK_ARG(vec8_ok(*identity, eb) && same_extent(*x, *identity), "ese_apply: identity view mismatch");
    idn = *identity;
  }
  int64_t total8 = (int64_t)x->n * x->h * x->w * (x->c / 8);
  if (total8 == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CM2_F32)
    ese_apply_kernel<float><<<grid_for(total8, 256), 256, 0, s>>>(make_view<const float>(*x), gate,
                                                                 make_view<const float>(idn), make_view<float>(*out));
  else
    ese_apply_kernel<__nv_bfloat16><<<grid_for(total8, 256), 256, 0, s>>>(
        make_view<const __nv_bfloat16>(*x), gate, make_view<const __nv_bfloat16>(idn), make_view<__nv_bfloat16>(*out));
  CM2_CHECK_LAUNCH("ese_apply");
  return CM2_OK;
}

extern "C" int64_t cm2_gn_workspace_floats(int32_t n, int32_t hw, int32_t c, int32_t groups) {
  int chunks = ceil_div(hw, GN_PIX_PER_CHUNK);
  return (int64_t)2 * n * chunks * c + (int64_t)2 * n * groups;
}

extern "C" int cm2_groupnorm_relu(const cm2_act* x, int32_t dtype, int32_t groups, const float* gamma,
                                  const float* beta, float eps, int32_t relu, float* workspace, void* stream) {
  CM2_CHECK_ARG(x && x->data && gamma && beta && workspace, "groupnorm: null pointer");
  CM2_CHECK_DTYPE(dtype, "groupnorm");
  int c = x->c, n = x->n, hw = x->h * x->w;
  CM2_CHECK_ARG(groups > 0 && c % groups == 0 && vec8_ok(*x, elem_bytes(dtype)) && c / 8 <= 256 && 256 % (c / 8) == 0,
                "groupnorm: unsupported c=%d groups=%d (need c %% 8 == 0, c/8 a divisor of 256)", c, groups);
  if (n == 0 || hw == 0) return CM2_OK;
  const int nthreads = 256;
  int chunks = ceil_div(hw, GN_PIX_PER_CHUNK);
  cudaStream_t s = (cudaStream_t)stream;
  float* partial = workspace;
  float* stats = workspace + (size_t)2 * n * chunks * c;
  dim3 g1(chunks, n);
  if (dtype == CM2_F32)
    gn_partial_kernel<float><<<g1, nthreads, 0, s>>>(make_view<const float>(*x), chunks, partial);
  else
    gn_partial_kernel<__nv_bfloat16><<<g1, nthreads, 0, s>>>(make_view<const __nv_bfloat16>(*x), chunks, partial);
  CM2_CHECK_LAUNCH("gn_partial");
  dim3 g2(groups, n);
  gn_final_kernel<<<g2, 32, 0, s>>>(partial, hw, c, groups, chunks, eps, stats);
  CM2_CHECK_LAUNCH("gn_final");
  int64_t total8 = (int64_t)n * hw * (c / 8);
  if (dtype == CM2_F32)
    gn_apply_kernel<float><<<grid_for(total8, 256), 256, 0, s>>>(make_view<float>(*x), groups, stats, gamma, beta, relu);
  else
    gn_apply_kernel<__nv_bfloat16><<<grid_for(total8, 256), 256, 0, s>>>(make_view<__nv_bfloat16>(*x), groups, stats,
                                                                        gamma, beta, relu);
  CM2_CHECK_LAUNCH("gn_apply");
  return CM2_OK;
}

extern "C" int64_t cm2_gn_seg_workspace_floats(int32_t num_seg, const cm2_seg* seg, int32_t c, int32_t groups) {
  GnSegs g;
  if (!seg || !gn_make_segs(num_seg, seg, c, &g)) return 0;
  return (int64_t)2 * g.chunk_prefix[num_seg] * c + (int64_t)2 * g.img_prefix[num_seg] * groups;
}

extern "C" int cm2_groupnorm_relu_seg(void* x, int32_t dtype, int32_t c, int32_t num_seg, const cm2_seg* seg,
                                      int32_t groups, const float* gamma, const float* beta, float eps, int32_t relu,
                                      float* workspace, void* stream) {
  CM2_CHECK_ARG(x && seg && gamma && beta && workspace, "groupnorm_seg: null pointer");
  CM2_CHECK_DTYPE(dtype, "groupnorm_seg");
  CM2_CHECK_ARG(groups > 0 && c % groups == 0 && c % 8 == 0 && c / 8 <= 256 && 256 % (c / 8) == 0 &&
                (reinterpret_cast<uintptr_t>(x) % (size_t)(8 * elem_bytes(dtype))) == 0,
                "groupnorm_seg: unsupported c=%d groups=%d", c, groups);
  GnSegs g;
  CM2_CHECK_ARG(gn_make_segs(num_seg, seg, c, &g), "groupnorm_seg: bad segment table");
  cudaStream_t s = (cudaStream_t)stream;
  float* partial = workspace;
  float* stats = workspace + (size_t)2 * g.chunk_prefix[num_seg] * c;
  const int chunks = g.chunk_prefix[num_seg], imgs = g.img_prefix[num_seg];
  if (dtype == CM2_F32)
    gn_seg_partial_kernel<float><<<chunks, 256, 0, s>>>((const float*)x, c, g, partial);
  else
    gn_seg_partial_kernel<__nv_bfloat16><<<chunks, 256, 0, s>>>((const __nv_bfloat16*)x, c, g, partial);
  CM2_CHECK_LAUNCH("gn_seg_partial");
  gn_seg_final_kernel<<<dim3(groups, imgs), 32, 0, s>>>(partial, c, groups, g, eps, stats);
  CM2_CHECK_LAUNCH("gn_seg_final");
  const long long total8 = g.vec_prefix[num_seg];
  if (dtype == CM2_F32)
    gn_seg_apply_kernel<float><<<grid_for(total8, 256), 256, 0, s>>>((float*)x, c, groups, g, stats, gamma, beta, relu);
  else
    gn_seg_apply_kernel<__nv_bfloat16><<<grid_for(total8, 256), 256, 0, s>>>((__nv_bfloat16*)x, c, groups, g, stats, gamma,
                                                                          beta, relu);
  CM2_CHECK_LAUNCH("gn_seg_apply");
  return CM2_OK;
}

extern "C" int cm2_split_f16x2(const float* x, void* out, int64_t pixels, int32_t c, void* stream) {
  CM2_CHECK_ARG(x && out, "split_f16x2: null pointer");
  CM2_CHECK_ARG(pixels >= 0 && c > 0 && c % 8 == 0, "split_f16x2: c=%d must be a positive multiple of 8", c);
  CM2_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "split_f16x2: pointers must be 16-byte aligned");
  const long long total8 = (long long)pixels * (c / 8);
  if (total8 == 0) return CM2_OK;
  split_f16x2_kernel<<<grid_for(total8, 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__half*>(out), total8, c / 8);
  CM2_CHECK_LAUNCH("split_f16x2");
  return CM2_OK;
}

extern "C" int cm2_relu(const cm2_act* in, const cm2_act* out, int32_t dtype, void* stream) {
  CM2_CHECK_ARG(in && out && in->data && out->data, "relu: null pointer");
  CM2_CHECK_DTYPE(dtype, "relu");
  int eb = elem_bytes(dtype);
  CM2_CHECK_ARG(vec8_ok(*in, eb) && vec8_ok(*out, eb) && same_extent(*in, *out), "relu: bad views");
  int64_t total8 = (int64_t)in->n * in->h * in->w * (in->c / 8);
  if (total8 == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CM2_F32)
    relu_kernel<float><<<grid_for(total8, 256), 256, 0, s>>>(make_view<const float>(*in), make_view<float>(*out));
  else
    relu_kernel<__nv_bfloat16><<<grid_for(total8, 256), 256, 0, s>>>(make_view<const __nv_bfloat16>(*in),
                                                                    make_view<__nv_bfloat16>(*out));
  CM2_CHECK_LAUNCH("relu");
  return CM2_OK;
}
