// stem_1 fused with the input side: normalise + pad + 3x3 / stride 2 / pad 1 convolution (3 -> 64) + FrozenBN + ReLU in ONE pass
// (reference: deploy_utils.py:76-98 preprocess_image, vovnet.py:205-236 conv3x3 unit, vovnet.py:392-400 stem_1).
//
// Why its own kernel: Cin = 3 makes the layer HBM-bound (51 MB of uint8 in, 550 MB of bf16 out at batch 16, 14.9 GFLOP).  The
// tcgen05 engine needs K-major operand tiles in shared memory, so the previous path first wrote a 32-channel im2col tensor
// (221 MB) and then read it back as a K = 32 GEMM: 0.19 + 0.20 ms.  Here the normalised bf16 image tile lives in shared
// memory only, the A fragments of the warp-level tensor-core instruction (mma.sync.m16n8k16, bf16 -> fp32; the K = 30 band is far
// too thin for a tcgen05 tile) are gathered from it with 4-byte loads, and the only HBM traffic is the image and the output.
//
// K order: k = ky * 10 + kx * 3 + c (ky rows padded from 9 to 10 so that a (k, k + 1) operand pair never straddles a filter row
// and is one aligned 32-bit shared-memory load: the tile is stored [row][pixel][channel], so the 9 values of a filter row are
// contiguous).  Columns k = 9, 19, 29, 30, 31 carry zero weights; their A values are whatever finite bf16 the tile holds.
// Arithmetic identical to the two-pass path: (float(u8) - mean) / std rounded to bf16, fp32 accumulation, fma(acc, scale,
// shift), ReLU, round to bf16.
#include "common.cuh"
#include <cuda_fp16.h>
#include <string.h>
#include <algorithm>

namespace cm2 {

constexpr int ST_TR = 4, ST_TC = 128;                   // output tile: rows x columns
constexpr int ST_IR = 2 * ST_TR + 1;                    // input rows of a tile
constexpr int ST_IC = 2 * ST_TC + 2;                    // input pixels per row (one more than needed: the k = 9 over-read)
constexpr int ST_PITCH = ST_IC * 3 + 2;                 // bf16 elements per tile row (even: rows stay 4-byte aligned)
constexpr int ST_OPITCH = 144;                          // staged output bytes per pixel (128 + 16: conflict-free 4-byte writes)
constexpr int ST_THREADS = 256;
static_assert(ST_IC == ST_THREADS + 2, "the tile loader maps thread -> pixel");

struct StemBatch {
  const void* img[CM2_MAX_BATCH_PTRS];
  int h[CM2_MAX_BATCH_PTRS], w[CM2_MAX_BATCH_PTRS];
};

template <bool F16>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// shared-memory carve-up (dynamic): NT tiles of 16-bit elements, the per-warp output staging, scale / shift, NT sets of B fragments
constexpr int ST_TILE_BYTES = ST_IR * ST_PITCH * 2;
constexpr int ST_OST_BYTES = (ST_THREADS / 32) * 16 * ST_OPITCH;
constexpr int ST_B_BYTES = 8 * 2 * 32 * 8;
__host__ __device__ constexpr int stem_smem_bytes(bool split) {
  return (split ? 2 : 1) * (ST_TILE_BYTES + ST_B_BYTES) + ST_OST_BYTES + 2 * 64 * 4;
}
static_assert(ST_TILE_BYTES % 16 == 0 && ST_OST_BYTES % 16 == 0, "16-byte aligned regions");

// SPLIT = false: bf16 engine (one bf16 tile, bf16 weights w30 [64][32], bf16 output of 64 channels).
// SPLIT = true : fp32 engine (include/cm2.h "Split precision"): the normalised fp32 value is kept as an f16 pair hi = half(v),
//   lo = half(v - hi) in two tiles, the weights come as W_hi / W_lo (w30 [2][64][32], pre-scaled per output channel by a power of
//   two that `scale` undoes), acc = x_hi W_hi + x_lo W_hi + x_hi W_lo in fp32, and the result is stored as the [hi | lo] f16 operand
//   pair of stem_2 (128 channels per pixel: hi at co, lo at 64 + co) -- no fp32 copy, no split pass.
template <typename InT, bool UNIT_STD, bool SPLIT>
__global__ void __launch_bounds__(ST_THREADS, SPLIT ? 2 : 3) stem1_fused_kernel(StemBatch bt, int ho, int wo, float m0, float m1, float m2,
                                                                            float r0, float r1, float r2, const uint16_t* __restrict__ w30,
                                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                                            int relu, View<uint16_t> out, int b0) {
  constexpr int NT = SPLIT ? 2 : 1;
  extern __shared__ __align__(16) unsigned char st_smem[];
  uint16_t* tile = reinterpret_cast<uint16_t*>(st_smem);                               // [NT][ST_IR * ST_PITCH]
  unsigned char* ostage = st_smem + NT * ST_TILE_BYTES;
  float* s_sc = reinterpret_cast<float*>(ostage + ST_OST_BYTES);
  float* s_sh = s_sc + 64;
  uint2* s_b = reinterpret_cast<uint2*>(s_sh + 64);                                    // [NT][8 * 2 * 32]
  const int b = blockIdx.z, oy0 = blockIdx.y * ST_TR, ox0 = blockIdx.x * ST_TC;
  const InT* __restrict__ img = reinterpret_cast<const InT*>(bt.img[b]);
  const int h = bt.h[b], w = bt.w[b];
  const size_t plane = (size_t)h * w;
  // ---- stage the normalised tile.  One (plane, row) per unrolled step, thread = pixel: no index arithmetic, all global loads of
  // a thread are issued before its first shared-memory store
  InT stage[3 * ST_IR];
  const int ix_t = 2 * ox0 - 1 + (int)threadIdx.x;
  const bool x_ok = ix_t >= 0 && ix_t < w;
#pragma unroll
  for (int cr = 0; cr < 3 * ST_IR; ++cr) {
    const int c = cr / ST_IR, r = cr - c * ST_IR;        // compile-time
    const int iy = 2 * oy0 - 1 + r;
    stage[cr] = (InT)0;
    if (x_ok && iy >= 0 && iy < h) stage[cr] = __ldg(img + c * plane + (size_t)iy * w + ix_t);
  }
  // the ST_IC - ST_THREADS = 2 pixels per (plane, row) left over: thread i < 2 * 27 takes pixel 256 + (i & 1) of row i >> 1
  InT tail = (InT)0;
  const int tcr = threadIdx.x >> 1, tt = ST_THREADS + (threadIdx.x & 1);
  const int tc_ = tcr / ST_IR, tr_ = tcr - tc_ * ST_IR;
  const int tiy = 2 * oy0 - 1 + tr_, tix = 2 * ox0 - 1 + tt;
  const bool tail_on = threadIdx.x < 2 * 3 * ST_IR;
  const bool tail_in = tail_on && tiy >= 0 && tiy < h && tix >= 0 && tix < w;
  if (tail_in) tail = __ldg(img + tc_ * plane + (size_t)tiy * w + tix);
  auto put = [&](int idx, InT raw, int c, bool inside) {
    float val = 0.f;                                     // padding (conv pad and the /32 image pad) is zero AFTER normalisation
    if (inside) {
      val = (float)raw - (c == 0 ? m0 : (c == 1 ? m1 : m2));
      if (!UNIT_STD) val = val / (c == 0 ? r0 : (c == 1 ? r1 : r2));
    }
    if (SPLIT) {
      const __half hi = __float2half_rn(val);
      const __half lo = __float2half_rn(val - __half2float(hi));
      tile[idx] = __half_as_ushort(hi);
      tile[ST_IR * ST_PITCH + idx] = __half_as_ushort(lo);
    } else {
      tile[idx] = __bfloat16_as_ushort(__float2bfloat16_rn(val));
    }
  };
#pragma unroll
  for (int cr = 0; cr < 3 * ST_IR; ++cr) {
    const int c = cr / ST_IR, r = cr - c * ST_IR;
    const int iy = 2 * oy0 - 1 + r;
    put(r * ST_PITCH + (int)threadIdx.x * 3 + c, stage[cr], c, x_ok && iy >= 0 && iy < h);
  }
  if (tail_on) put(tr_ * ST_PITCH + tt * 3 + tc_, tail, tc_, tail_in);
  if (threadIdx.x < ST_IR) {                             // the two spare elements of every row (read by the last pixel's k = 9)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      tile[t * ST_IR * ST_PITCH + threadIdx.x * ST_PITCH + ST_IC * 3] = 0;
      tile[t * ST_IR * ST_PITCH + threadIdx.x * ST_PITCH + ST_IC * 3 + 1] = 0;
    }
  }
  // ---- per-lane constants: scale / shift, B fragments, A offsets
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  if (threadIdx.x < 64) {
    s_sc[threadIdx.x] = scale ? __ldg(scale + threadIdx.x) : 1.f;
    s_sh[threadIdx.x] = shift ? __ldg(shift + threadIdx.x) : 0.f;
  }
  // B fragments [set][j][s][lane] as uint2 = {(k = 16 s + 2 q, + 1), (k = 16 s + 2 q + 8, + 9)} of output channel 8 j + g: a warp
  // reads 256 contiguous bytes per (j, s)
  for (int i = threadIdx.x; i < NT * 8 * 2 * 32; i += ST_THREADS) {
    const int l = i & 31, s2 = (i >> 5) & 1, j = (i >> 6) & 7, set = i >> 9;
    const uint32_t* wr = reinterpret_cast<const uint32_t*>(w30 + set * 64 * 32 + (8 * j + (l >> 2)) * 32);
    s_b[i] = make_uint2(__ldg(wr + 8 * s2 + (l & 3)), __ldg(wr + 8 * s2 + (l & 3) + 4));
  }
  int aoff[4];                                           // element offsets of the operand pairs p = q, q + 4, q + 8, q + 12
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int p = q + 4 * i;
    if (p > 14) p = 14;                                  // k = 30, 31: zero weights, any in-tile address
    aoff[i] = (p / 5) * ST_PITCH + 2 * (p % 5);
  }
  const float floor_v = relu ? 0.f : -INFINITY;
  const uint32_t ost = (uint32_t)__cvta_generic_to_shared(ostage) + (uint32_t)warp * 16u * ST_OPITCH;
  __syncthreads();
  // ---- 16-pixel groups: TR rows x TC / 16 groups per row, round-robin over the warps
#pragma unroll 1
  for (int gi = warp; gi < ST_TR * (ST_TC / 16); gi += ST_THREADS / 32) {
    const int r = gi / (ST_TC / 16), cg = gi - r * (ST_TC / 16);
    const int oy = oy0 + r, oxg = ox0 + cg * 16;
    if (oy >= ho || oxg >= wo) continue;                 // warp-uniform
    uint32_t a[NT][2][4];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const uint16_t* arow = tile + t * ST_IR * ST_PITCH + (2 * r) * ST_PITCH + (cg * 16 + g) * 6;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        a[t][s][0] = *reinterpret_cast<const uint32_t*>(arow + aoff[2 * s]);            // row g,     k = 16 s + 2 q
        a[t][s][1] = *reinterpret_cast<const uint32_t*>(arow + 48 + aoff[2 * s]);       // row g + 8 (8 pixels = 48 elements on)
        a[t][s][2] = *reinterpret_cast<const uint32_t*>(arow + aoff[2 * s + 1]);        // row g,     k = 16 s + 2 q + 8
        a[t][s][3] = *reinterpret_cast<const uint32_t*>(arow + 48 + aoff[2 * s + 1]);   // row g + 8
      }
    }
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      const uint2 b0 = s_b[(2 * j) * 32 + lane], b1 = s_b[(2 * j + 1) * 32 + lane];
      if (SPLIT) {
        // cross terms first (2^-11 of the main term), then the main term: x_lo W_hi + x_hi W_lo + x_hi W_hi
        const uint2 l0 = s_b[512 + (2 * j) * 32 + lane], l1 = s_b[512 + (2 * j + 1) * 32 + lane];
        mma_16816<true>(acc[j], a[NT - 1][0], b0.x, b0.y);
        mma_16816<true>(acc[j], a[NT - 1][1], b1.x, b1.y);
        mma_16816<true>(acc[j], a[0][0], l0.x, l0.y);
        mma_16816<true>(acc[j], a[0][1], l1.x, l1.y);
      }
      mma_16816<SPLIT>(acc[j], a[0][0], b0.x, b0.y);
      mma_16816<SPLIT>(acc[j], a[0][1], b1.x, b1.y);
    }
    // epilogue: scale / shift / ReLU, staged so that the warp writes 16 pixels x 128 bytes as 16-byte pieces (two rounds for the
    // [hi | lo] pair)
    uint16_t* orow = out.at(b0 + b, oy, oxg);
#pragma unroll
    for (int round = 0; round < NT; ++round) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 sc = *reinterpret_cast<const float2*>(s_sc + 8 * j + 2 * q), sh = *reinterpret_cast<const float2*>(s_sh + 8 * j + 2 * q);
        const float v0 = fmaxf(fmaf(acc[j][0], sc.x, sh.x), floor_v), v1 = fmaxf(fmaf(acc[j][1], sc.y, sh.y), floor_v);
        const float v2 = fmaxf(fmaf(acc[j][2], sc.x, sh.x), floor_v), v3 = fmaxf(fmaf(acc[j][3], sc.y, sh.y), floor_v);
        uint32_t w01, w23;
        if (SPLIT) {
          const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
          if (round == 0) {
            w01 = *reinterpret_cast<const uint32_t*>(&h01); w23 = *reinterpret_cast<const uint32_t*>(&h23);
          } else {
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y), l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
            w01 = *reinterpret_cast<const uint32_t*>(&l01); w23 = *reinterpret_cast<const uint32_t*>(&l23);
          }
        } else {
          const __nv_bfloat162 p01 = __floats2bfloat162_rn(v0, v1), p23 = __floats2bfloat162_rn(v2, v3);
          w01 = *reinterpret_cast<const uint32_t*>(&p01); w23 = *reinterpret_cast<const uint32_t*>(&p23);
        }
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ost + (uint32_t)g * ST_OPITCH + 16u * j + 4u * q), "r"(w01) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ost + (uint32_t)(g + 8) * ST_OPITCH + 16u * j + 4u * q), "r"(w23) : "memory");
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int px = it * 4 + (lane >> 3), chunk = lane & 7;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(ost + (uint32_t)px * ST_OPITCH + 16u * chunk) : "memory");
        if (oxg + px < wo) *reinterpret_cast<uint4*>(orow + (long long)px * out.sw + round * 64 + chunk * 8) = v;
      }
    }
  }
}

template <bool SPLIT>
static int stem_launch(const void* const* imgs, const int32_t* hs, const int32_t* ws, int32_t n, int32_t in_dtype, int32_t hp, int32_t wp,
                       const float* mean3, const float* std3, const void* w30, const float* scale, const float* shift, int32_t relu,
                       const cm2_act* out, int32_t out_index0, void* stream, const char* who) {
  CM2_CHECK_ARG(imgs && hs && ws && out && out->data && mean3 && std3 && w30, "%s: null pointer", who);
  CM2_CHECK_ARG(n >= 0 && hp > 0 && wp > 0 && hp % 2 == 0 && wp % 2 == 0, "%s: bad extents", who);
  constexpr int OC = SPLIT ? 128 : 64;
  CM2_CHECK_ARG(out->h == hp / 2 && out->w == wp / 2 && out->c == OC && out_index0 >= 0 && out_index0 + n <= out->n &&
                vec8_ok(*out, 2), "%s: out view [%d,%d,%d,%d] != [>=%d,%d,%d,%d]", who, out->n, out->h, out->w,
                out->c, out_index0 + n, hp / 2, wp / 2, OC);
  CM2_CHECK_ARG(in_dtype == CM2_F32 || in_dtype == CM2_U8, "%s: unsupported input dtype %d", who, in_dtype);
  CM2_CHECK_ARG((reinterpret_cast<uintptr_t>(w30) & 3) == 0, "%s: weights not 4-byte aligned", who);
  cudaStream_t s = (cudaStream_t)stream;
  const bool unit = std3[0] == 1.f && std3[1] == 1.f && std3[2] == 1.f;
  const uint16_t* wq = reinterpret_cast<const uint16_t*>(w30);
  constexpr int smem = stem_smem_bytes(SPLIT);
  CM2_ENSURE_DYN_SMEM((stem1_fused_kernel<float, true, SPLIT>), smem, who);
  CM2_ENSURE_DYN_SMEM((stem1_fused_kernel<float, false, SPLIT>), smem, who);
  CM2_ENSURE_DYN_SMEM((stem1_fused_kernel<uint8_t, true, SPLIT>), smem, who);
  CM2_ENSURE_DYN_SMEM((stem1_fused_kernel<uint8_t, false, SPLIT>), smem, who);
  for (int i0 = 0; i0 < n; i0 += CM2_MAX_BATCH_PTRS) {
    const int nb = std::min(n - i0, (int)CM2_MAX_BATCH_PTRS);
    StemBatch bt;
    memset(&bt, 0, sizeof(bt));
    for (int i = 0; i < nb; ++i) {
      CM2_CHECK_ARG(imgs[i0 + i] && hs[i0 + i] > 0 && ws[i0 + i] > 0 && hs[i0 + i] <= hp && ws[i0 + i] <= wp,
                    "%s: image %d is %dx%d, padded extent %dx%d", who, i0 + i, hs[i0 + i], ws[i0 + i], hp, wp);
      bt.img[i] = imgs[i0 + i]; bt.h[i] = hs[i0 + i]; bt.w[i] = ws[i0 + i];
    }
    dim3 grid(ceil_div(out->w, ST_TC), ceil_div(out->h, ST_TR), nb);
    View<uint16_t> ov = make_view<uint16_t>(*out);
#define CM2_STEM(T, U) stem1_fused_kernel<T, U, SPLIT><<<grid, ST_THREADS, smem, s>>>(bt, out->h, out->w, mean3[0], mean3[1], mean3[2], std3[0], \
                                                                                      std3[1], std3[2], wq, scale, shift, relu, ov, out_index0 + i0)
    if (in_dtype == CM2_F32) { if (unit) CM2_STEM(float, true); else CM2_STEM(float, false); }
    else { if (unit) CM2_STEM(uint8_t, true); else CM2_STEM(uint8_t, false); }
#undef CM2_STEM
    CM2_CHECK_LAUNCH(who);
  }
  return CM2_OK;
}

}  // namespace cm2

extern "C" int cm2_stem1_fused_batch(const void* const* imgs, const int32_t* hs, const int32_t* ws, int32_t n, int32_t in_dtype,
                                     int32_t hp, int32_t wp, const float* mean3, const float* std3, const void* w30,
                                     const float* scale, const float* shift, int32_t relu, const cm2_act* out,
                                     int32_t out_index0, void* stream) {
  return cm2::stem_launch<false>(imgs, hs, ws, n, in_dtype, hp, wp, mean3, std3, w30, scale, shift, relu, out, out_index0, stream,
                                 "stem1_fused_batch");
}

extern "C" int cm2_stem1_fused_split_batch(const void* const* imgs, const int32_t* hs, const int32_t* ws, int32_t n, int32_t in_dtype,
                                           int32_t hp, int32_t wp, const float* mean3, const float* std3, const void* w30_hi_lo,
                                           const float* scale, const float* shift, int32_t relu, const cm2_act* out_split,
                                           int32_t out_index0, void* stream) {
  return cm2::stem_launch<true>(imgs, hs, ws, n, in_dtype, hp, wp, mean3, std3, w30_hi_lo, scale, shift, relu, out_split, out_index0,
                                stream, "stem1_fused_split_batch");
}
