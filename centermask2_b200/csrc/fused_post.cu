// Consumers of the statistics the tensor-core convolution accumulates in its epilogue (cm2_conv_desc.stats):
//   * GroupNorm(+ReLU) of the FCOS towers (fcos.py:182-186) from per-(image, 8-channel chunk) fp64 sums: ONE
//     in-place pass over the segmented tower tensor (the separate partial / final reduction passes are gone);
//   * eSE (vovnet.py:247-260): gate from the fp64 channel sums, then x * gate (+ identity) fused with the
//     MaxPool2d(3, 2, ceil_mode=True) that opens the next stage (vovnet.py:349-350) -- the stage output is read once.
// All HBM-bound: 16-byte channel vectors, 32-bit index arithmetic, a few rows per thread.
#include "common.cuh"
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>

namespace cm2 {

int grid_for(int64_t work, int block);

// ---------------------------------------------------------------------------------------------
// GroupNorm apply on a segmented halo tensor
// ---------------------------------------------------------------------------------------------
struct GnRowSegs {
  int num;
  int row0[CM2_MAX_SEG], rows[CM2_MAX_SEG], pitch[CM2_MAX_SEG], plane[CM2_MAX_SEG];
  int h[CM2_MAX_SEG], w[CM2_MAX_SEG], img0[CM2_MAX_SEG];
};

constexpr int GN_ROWS_PER_THREAD = 8;
constexpr int GN_BATCH = 4;

// Thread (cv, row block): GN_ROWS_PER_THREAD consecutive flat rows of one 8-channel vector column.  The row ->
// (segment, image, y, x) decomposition is done once per thread and advanced incrementally (32-bit arithmetic only);
// the per-(image, group) mean / rstd are recomputed only when the image changes.
template <typename T>
__global__ void __launch_bounds__(256) gn_seg_apply_stats_kernel(T* __restrict__ x, int c, int cpg, GnRowSegs g, int total_rows,
                                                                 const double* __restrict__ stats,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps, int relu) {
  const unsigned c8 = (unsigned)c >> 3;
  const int chunks_per_group = cpg >> 3;
  const unsigned nvec = (unsigned)((total_rows + GN_ROWS_PER_THREAD - 1) / GN_ROWS_PER_THREAD) * c8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
    const int cv = (int)(i % c8);
    const int rb = (int)(i / c8) * GN_ROWS_PER_THREAD;
    int cur_gi = -1;
    float a_mul[8], a_add[8];                            // y = v * a_mul + a_add
    int s = -1, r = 0, img = 0, yy = 0, xx = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < GN_ROWS_PER_THREAD; k0 += GN_BATCH) {
      // ---- pass 1: decode GN_BATCH rows and issue their loads together (memory-level parallelism)
      uint4 raw[GN_BATCH];
      int gis[GN_BATCH];
#pragma unroll
      for (int k = 0; k < GN_BATCH; ++k) {
        const int row = rb + k0 + k;
        gis[k] = -1;
        if (row >= total_rows) continue;
        if (s < 0 || r >= g.rows[s]) {                   // (re)locate: first row, or the previous segment ended
          s = 0;
#pragma unroll
          for (int j = 1; j < CM2_MAX_SEG; ++j)
            if (j < g.num && row >= g.row0[j]) s = j;
          r = row - g.row0[s];
          img = r / g.plane[s];
          const int rr = r - img * g.plane[s];
          yy = rr / g.pitch[s];
          xx = rr - yy * g.pitch[s];
        }
        if (r < g.rows[s] && yy >= 1 && yy <= g.h[s] && xx >= 1 && xx <= g.w[s]) {      // halo stays zero
          gis[k] = (s << 24) | (g.img0[s] + img);
          raw[k] = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(x) + ((size_t)row * c + cv * 8) * sizeof(T));
        }
        ++r;
        if (++xx == g.pitch[s]) {
          xx = 0;
          if (++yy == g.h[s] + 2) { yy = 0; ++img; }
        }
      }
      // ---- pass 2: normalise and store
#pragma unroll
      for (int k = 0; k < GN_BATCH; ++k) {
        if (gis[k] < 0) continue;
        if (gis[k] != cur_gi) {
          cur_gi = gis[k];
          const int sg = cur_gi >> 24, gi = cur_gi & 0xffffff;
          const int grp = (cv * 8) / cpg;
          const double* q = stats + ((size_t)gi * c8 + (size_t)grp * chunks_per_group) * 2;
          double sum = 0.0, sq = 0.0;
          for (int j = 0; j < chunks_per_group; ++j) { sum += q[2 * j]; sq += q[2 * j + 1]; }
          const double cnt = (double)g.h[sg] * (double)g.w[sg] * (double)cpg;
          const double m = sum / cnt;
          double var = sq / cnt - m * m;
          if (var < 0.0) var = 0.0;
          const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)eps));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float gm = __ldg(gamma + cv * 8 + j);
            a_mul[j] = rstd * gm;
            a_add[j] = __ldg(beta + cv * 8 + j) - mean * rstd * gm;
          }
        }
        T* ptr = x + (size_t)(rb + k0 + k) * c + cv * 8;
        float v[8];
        if (sizeof(T) == 2) {
          const uint32_t w[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[2 * j] = __uint_as_float(w[j] << 16); v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
        } else {
          const float4 lo = *reinterpret_cast<const float4*>(&raw[k]);
          const float4 hi = *(reinterpret_cast<const float4*>(ptr) + 1);
          v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float yv = fmaf(v[j], a_mul[j], a_add[j]);
          v[j] = relu ? fmaxf(yv, 0.f) : yv;
        }
        Vec8<T>::store(ptr, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// packed helpers (bf16 vectors of 8 channels = one uint4)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmul.rn.f32x2 rd, ra, rb;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void ffma2p(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {          // round-to-nearest-even
  uint32_t w;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
  return w;
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {     // max(x, 0) then round
  uint32_t w;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
  return w;
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ---------------------------------------------------------------------------------------------
// GroupNorm apply, bf16, one block per interior image line (no per-pixel index arithmetic): thread = (pixel lane,
// 8-channel vector); the per-(image, group) coefficients are computed once per thread.  ~20 instructions per
// 16-byte vector (the generic kernel above needed ~130 and was issue-bound at half of HBM speed).
// ---------------------------------------------------------------------------------------------
struct GnLineSegs {
  int num;
  int row0[CM2_MAX_SEG], pitch[CM2_MAX_SEG], plane[CM2_MAX_SEG], h[CM2_MAX_SEG], w[CM2_MAX_SEG], img0[CM2_MAX_SEG];
  int line_prefix[CM2_MAX_SEG + 1];
};

__global__ void __launch_bounds__(256) gn_seg_apply_lines_kernel(__nv_bfloat16* __restrict__ x, int c, int cpg, GnLineSegs g,
                                                                 const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps, int relu) {
  // CTAs walk the tensor BACKWARDS: the convolution that produced it wrote it front to back, so its tail is what the 126 MB L2
  // still holds; read back to front, that part never comes from DRAM, and the front -- written last here -- is in L2 when the
  // next convolution starts reading front to back (measured: see DESIGN.md section 3, round 2)
  const int line = (int)(gridDim.x - 1 - blockIdx.x);
  int s = 0;
#pragma unroll
  for (int j = 1; j < CM2_MAX_SEG; ++j)
    if (j < g.num && line >= g.line_prefix[j]) s = j;
  const int l = line - g.line_prefix[s];
  const int img = l / g.h[s], y = l - img * g.h[s];
  const int c8 = c >> 3;
  const int cv = threadIdx.x % c8, p0 = threadIdx.x / c8, pstep = blockDim.x / c8;
  const int w = g.w[s];
  // coefficients of this thread's 8 channels: y = v * a + b
  float a[8], b[8];
  {
    const int gi = g.img0[s] + img;
    const int chunks_per_group = cpg >> 3;
    const int grp = (cv * 8) / cpg;
    const double* q = stats + ((size_t)gi * c8 + (size_t)grp * chunks_per_group) * 2;
    double sum = 0.0, sq = 0.0;
    for (int j = 0; j < chunks_per_group; ++j) { sum += q[2 * j]; sq += q[2 * j + 1]; }
    const double cnt = (double)g.h[s] * (double)w * (double)cpg;
    const double m = sum / cnt;
    double var = sq / cnt - m * m;
    if (var < 0.0) var = 0.0;
    const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)eps));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gm = __ldg(gamma + cv * 8 + j);
      a[j] = rstd * gm;
      b[j] = __ldg(beta + cv * 8 + j) - mean * rstd * gm;
    }
  }
  uint4* base = reinterpret_cast<uint4*>(x + ((size_t)g.row0[s] + (size_t)img * g.plane[s] + (size_t)(y + 1) * g.pitch[s] + 1) * c) + cv;
  constexpr int U = 4;                                  // independent 16-byte loads in flight per thread
  for (int px = p0; px < w; px += U * pstep) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (px + u * pstep < w) raw[u] = base[(size_t)(px + u * pstep) * c8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (px + u * pstep >= w) break;
      const uint32_t in[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v0, v1;
        ffma2p(v0, v1, bf_lo(in[j]), bf_hi(in[j]), a[2 * j], a[2 * j + 1], b[2 * j], b[2 * j + 1]);
        o[j] = relu ? pack_bf16x2_relu(v0, v1) : pack_bf16x2(v0, v1);
      }
      base[(size_t)(px + u * pstep) * c8] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// fp32 engine: GroupNorm normalise + ReLU of an fp32 segmented tensor written straight into the [hi | lo] f16 operand form
// of the next convolution (include/cm2.h "Split precision"): the separate cm2_split_f16x2 pass over the tower tensor
// disappears.  One CTA per image line; a thread owns 8 channels of a pixel: 32 bytes in, 16 + 16 bytes out.  Only
// interior pixels are written: the halo of the (zero-initialised) output stays zero.
__global__ void __launch_bounds__(256) gn_seg_apply_split_lines_kernel(const float* __restrict__ x, __half* __restrict__ out, int c, int cpg,
                                                                       GnLineSegs g, const double* __restrict__ stats,
                                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                       float eps, int relu) {
  // CTAs walk the tensor BACKWARDS: the convolution that produced it wrote it front to back, so its tail is what the 126 MB L2
  // still holds; read back to front, that part never comes from DRAM, and the front -- written last here -- is in L2 when the
  // next convolution starts reading front to back (measured: see DESIGN.md section 3, round 2)
  const int line = (int)(gridDim.x - 1 - blockIdx.x);
  int s = 0;
#pragma unroll
  for (int j = 1; j < CM2_MAX_SEG; ++j)
    if (j < g.num && line >= g.line_prefix[j]) s = j;
  const int l = line - g.line_prefix[s];
  const int img = l / g.h[s], y = l - img * g.h[s];
  const int c8 = c >> 3;
  const int w = g.w[s];
  for (int cv = threadIdx.x % min(c8, (int)blockDim.x); cv < c8; cv += blockDim.x) {      // c8 <= 256: one pass for c <= 2048
    float a[8], b[8];
    {
      const int gi = g.img0[s] + img;
      const int chunks_per_group = cpg >> 3;
      const int grp = (cv * 8) / cpg;
      const double* q = stats + ((size_t)gi * c8 + (size_t)grp * chunks_per_group) * 2;
      double sum = 0.0, sq = 0.0;
      for (int j = 0; j < chunks_per_group; ++j) { sum += q[2 * j]; sq += q[2 * j + 1]; }
      const double cnt = (double)g.h[s] * (double)w * (double)cpg;
      const double m = sum / cnt;
      double var = sq / cnt - m * m;
      if (var < 0.0) var = 0.0;
      const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)eps));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gm = __ldg(gamma + cv * 8 + j);
        a[j] = rstd * gm;
        b[j] = __ldg(beta + cv * 8 + j) - mean * rstd * gm;
      }
    }
    const size_t row_first = (size_t)g.row0[s] + (size_t)img * g.plane[s] + (size_t)(y + 1) * g.pitch[s] + 1;
    const int pstep = max(1, (int)blockDim.x / c8), p0 = threadIdx.x / c8;
    for (int px = p0; px < w; px += pstep) {
      const float4* src = reinterpret_cast<const float4*>(x + (row_first + px) * c + cv * 8);
      const float4 u0 = __ldg(src), u1 = __ldg(src + 1);
      float v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float r0 = fmaf(v[2 * j], a[2 * j], b[2 * j]), r1 = fmaf(v[2 * j + 1], a[2 * j + 1], b[2 * j + 1]);
        if (relu) { r0 = fmaxf(r0, 0.f); r1 = fmaxf(r1, 0.f); }
        const __half2 h = __floats2half2_rn(r0, r1);
        const float2 hf = __half22float2(h);
        const __half2 lw = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
        lo[j] = *reinterpret_cast<const uint32_t*>(&lw);
      }
      __half* o = out + (row_first + px) * (size_t)(2 * c) + cv * 8;
      *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(o + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// eSE gate from fp64 channel sums: gate[b, o] = relu6(sum_i W[o,i] * mean[b,i] + bias[o] + 3) / 6
// ---------------------------------------------------------------------------------------------
__global__ void ese_gate_f64_kernel(const double* __restrict__ sums, double inv_count, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ gate, int c) {
  const int b = blockIdx.y;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= c) return;
  const float* wr = w + (size_t)warp * c;
  const double* pv = sums + (size_t)b * c;
  float acc = 0.f;
  for (int i = lane; i < c; i += 32) acc = fmaf(__ldg(wr + i), (float)(pv[i] * inv_count), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + bias[warp] + 3.0f;
    v = fminf(fmaxf(v, 0.f), 6.f) / 6.0f;
    gate[(size_t)b * c + warp] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// eSE apply fused with the following 3x3 / stride-2 / ceil-mode max-pool.  One thread per (pooled pixel,
// 8-channel vector): it evaluates y = round_T(x * gate (+ identity)) on its 3x3 window, writes the max to `pool`,
// and -- when `full` is given -- stores y for the window pixels it owns (rows/cols 2o, 2o+1, plus the trailing
// row/col for the last pooled index), so every full-resolution pixel is written exactly once.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ese_apply_pool_kernel(View<const T> x, const float* __restrict__ gate, View<const T> idn,
                                                             View<T> full, View<T> pool) {
  const unsigned c8 = (unsigned)x.c >> 3;
  const unsigned total = (unsigned)pool.n * pool.h * pool.w * c8;          // < 2^32 (checked by the host wrapper)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cv = (int)(i % c8);
    const unsigned pix = i / c8;
    const int ox = (int)(pix % (unsigned)pool.w);
    const unsigned t = pix / (unsigned)pool.w;
    const int oy = (int)(t % (unsigned)pool.h), b = (int)(t / (unsigned)pool.h);
    float gt[8];
    {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * x.c + cv * 8));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * x.c + cv * 8) + 1);
      gt[0] = a0.x; gt[1] = a0.y; gt[2] = a0.z; gt[3] = a0.w; gt[4] = a1.x; gt[5] = a1.y; gt[6] = a1.z; gt[7] = a1.w;
    }
    float best[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) best[k] = -INFINITY;
    const int own_y = (oy == pool.h - 1) ? 3 : 2, own_x = (ox == pool.w - 1) ? 3 : 2;
    // issue every load of the window first (9 or 18 independent 16-byte loads in flight per thread)
    uint4 rx[9], ri[9];
    constexpr int VB = 16 / (8 * sizeof(T) / 8) / 2;      // uint4 loads per 8-channel vector: 1 (bf16) or 2 (f32)
    static_assert(sizeof(T) == 2 || sizeof(T) == 4, "T");
    if (sizeof(T) == 2) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int y = 2 * oy + dy, xx = 2 * ox + dx;
          if (y < x.h && xx < x.w) {
            rx[dy * 3 + dx] = __ldg(reinterpret_cast<const uint4*>(x.at(b, y, xx) + cv * 8));
            if (idn.p) ri[dy * 3 + dx] = __ldg(reinterpret_cast<const uint4*>(idn.at(b, y, xx) + cv * 8));
          }
        }
    }
    (void)VB;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int y = 2 * oy + dy;
      if (y >= x.h) break;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = 2 * ox + dx;
        if (xx >= x.w) break;
        float v[8], r[8];
        if (sizeof(T) == 2) {
          const uint32_t w[4] = {rx[dy * 3 + dx].x, rx[dy * 3 + dx].y, rx[dy * 3 + dx].z, rx[dy * 3 + dx].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[2 * j] = __uint_as_float(w[j] << 16); v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
          if (idn.p) {
            const uint32_t u[4] = {ri[dy * 3 + dx].x, ri[dy * 3 + dx].y, ri[dy * 3 + dx].z, ri[dy * 3 + dx].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { r[2 * j] = __uint_as_float(u[j] << 16); r[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u); }
          }
        } else {
          Vec8<T>::load(x.at(b, y, xx) + cv * 8, v);
          if (idn.p) Vec8<T>::load(idn.at(b, y, xx) + cv * 8, r);
        }
        if (idn.p) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __fadd_rn(__fmul_rn(v[k], gt[k]), r[k]);      // two roundings, as ATen (no FMA contraction)
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = v[k] * gt[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = to_f32<T>(from_f32<T>(v[k]));           // the value as stored
          best[k] = fmaxf(best[k], v[k]);
        }
        if (full.p && dy < own_y && dx < own_x) Vec8<T>::store(full.at(b, y, xx) + cv * 8, v);
      }
    }
    Vec8<T>::store(pool.at(b, oy, ox) + cv * 8, best);
  }
}

// bf16 fast path of the above: one block per (image, pooled line), thread = (pooled-pixel lane, 8-channel vector),
// packed bf16x2 / f32x2 arithmetic.  Without identity the max is taken on the raw bf16 inputs and multiplied once
// (gate >= 0 and rounding is monotone, so max_k round(x_k * g) == round(max_k(x_k) * g)): ~50 instructions per
// pooled vector instead of ~400.
template <bool IDN, bool FULL>
__global__ void __launch_bounds__(256) ese_apply_pool_lines_kernel(View<const __nv_bfloat16> x, const float* __restrict__ gate,
                                                                   View<const __nv_bfloat16> idn, View<__nv_bfloat16> full,
                                                                   View<__nv_bfloat16> pool) {
  const int b = (int)(gridDim.y - 1 - blockIdx.y), oy = (int)(gridDim.x - 1 - blockIdx.x);     // back to front (L2, see gn_seg_apply_lines_kernel)
  const int c8 = x.c >> 3;
  const int cv = threadIdx.x % c8, o0 = threadIdx.x / c8, ostep = blockDim.x / c8;
  float gt[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * x.c + cv * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * x.c + cv * 8) + 1);
    gt[0] = a0.x; gt[1] = a0.y; gt[2] = a0.z; gt[3] = a0.w; gt[4] = a1.x; gt[5] = a1.y; gt[6] = a1.z; gt[7] = a1.w;
  }
  const int own_y = (oy == pool.h - 1) ? 3 : 2;
  const int ny = min(3, x.h - 2 * oy);                  // valid window rows / cols (>= 1)
  for (int ox = o0; ox < pool.w; ox += ostep) {
    const int own_x = (ox == pool.w - 1) ? 3 : 2;
    const int nx = min(3, x.w - 2 * ox);
    uint4 rx[9], ri[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
        if (dy < ny && dx < nx) {
          rx[dy * 3 + dx] = __ldg(reinterpret_cast<const uint4*>(x.at(b, 2 * oy + dy, 2 * ox + dx) + cv * 8));
          if (IDN) ri[dy * 3 + dx] = __ldg(reinterpret_cast<const uint4*>(idn.at(b, 2 * oy + dy, 2 * ox + dx) + cv * 8));
        }
    uint32_t m[4];
    if (!IDN) {
      m[0] = rx[0].x; m[1] = rx[0].y; m[2] = rx[0].z; m[3] = rx[0].w;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
          if (dy < ny && dx < nx) {
            const uint4 r = rx[dy * 3 + dx];
            m[0] = max_bf16x2(m[0], r.x); m[1] = max_bf16x2(m[1], r.y); m[2] = max_bf16x2(m[2], r.z); m[3] = max_bf16x2(m[3], r.w);
            if (FULL && dy < own_y && dx < own_x) {
              const uint32_t in[4] = {r.x, r.y, r.z, r.w};
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float v0, v1;
                fmul2(v0, v1, bf_lo(in[j]), bf_hi(in[j]), gt[2 * j], gt[2 * j + 1]);
                o[j] = pack_bf16x2(v0, v1);
              }
              *reinterpret_cast<uint4*>(full.at(b, 2 * oy + dy, 2 * ox + dx) + cv * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v0, v1;
        fmul2(v0, v1, bf_lo(m[j]), bf_hi(m[j]), gt[2 * j], gt[2 * j + 1]);
        m[j] = pack_bf16x2(v0, v1);
      }
    } else {
      m[0] = m[1] = m[2] = m[3] = 0xff80ff80u;             // (-inf, -inf)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
          if (dy < ny && dx < nx) {
            const uint4 r = rx[dy * 3 + dx], q = ri[dy * 3 + dx];
            const uint32_t in[4] = {r.x, r.y, r.z, r.w}, id[4] = {q.x, q.y, q.z, q.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // two roundings, as ATen (x * g, then + identity): scalar intrinsics are never contracted
              const float v0 = __fadd_rn(__fmul_rn(bf_lo(in[j]), gt[2 * j]), bf_lo(id[j]));
              const float v1 = __fadd_rn(__fmul_rn(bf_hi(in[j]), gt[2 * j + 1]), bf_hi(id[j]));
              o[j] = pack_bf16x2(v0, v1);
              m[j] = max_bf16x2(m[j], o[j]);
            }
            if (FULL && dy < own_y && dx < own_x)
              *reinterpret_cast<uint4*>(full.at(b, 2 * oy + dy, 2 * ox + dx) + cv * 8) = make_uint4(o[0], o[1], o[2], o[3]);
          }
    }
    *reinterpret_cast<uint4*>(pool.at(b, oy, ox) + cv * 8) = make_uint4(m[0], m[1], m[2], m[3]);
  }
}

// eSE apply over whole halo buffers (x, identity and out are interior views of identically shaped, contiguous
// one-pixel-halo buffers): a flat streaming pass, halo included (0 * g + 0 = 0 keeps it zero).
template <typename T>
__global__ void __launch_bounds__(256) ese_apply_flat_kernel(const T* __restrict__ x, const float* __restrict__ gate,
                                                             const T* __restrict__ idn, T* __restrict__ out, int c, int plane,
                                                             long long total8) {
  const unsigned c8 = (unsigned)c >> 3;
  for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < (unsigned)total8; j += gridDim.x * blockDim.x) {
    const unsigned i = (unsigned)total8 - 1u - j;                      // back to front (L2, see gn_seg_apply_lines_kernel)
    const int cv = (int)(i % c8);
    const unsigned row = i / c8;
    const int b = (int)(row / (unsigned)plane);
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * c + cv * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * c + cv * 8) + 1);
    const float gt[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float v[8];
    Vec8<T>::load(x + (size_t)i * 8, v);
    if (idn) {
      float r[8];
      Vec8<T>::load(idn + (size_t)i * 8, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __fadd_rn(__fmul_rn(v[k], gt[k]), r[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = v[k] * gt[k];
    }
    Vec8<T>::store(out + (size_t)i * 8, v);
  }
}

static bool halo_view(const cm2_act& a) {
  return a.sw == a.c && a.sh == (long long)(a.w + 2) * a.sw && a.sn == (long long)(a.h + 2) * a.sh;
}

}  // namespace cm2

using namespace cm2;

#define CM2_CHECK_DTYPE(dtype, name) \
  CM2_CHECK_ARG((dtype) == CM2_F32 || (dtype) == CM2_BF16, name ": dtype %d not supported", (int)(dtype))

extern "C" int cm2_groupnorm_apply_seg(void* x, int32_t dtype, int32_t c, int32_t num_seg, const cm2_seg* seg, int32_t groups,
                                       const float* gamma, const float* beta, float eps, int32_t relu, const double* stats,
                                       void* stream) {
  CM2_CHECK_ARG(x && seg && gamma && beta && stats, "groupnorm_apply_seg: null pointer");
  CM2_CHECK_DTYPE(dtype, "groupnorm_apply_seg");
  CM2_CHECK_ARG(groups > 0 && c % groups == 0 && c % 8 == 0 && (c / groups) % 8 == 0 &&
                (reinterpret_cast<uintptr_t>(x) % (size_t)(8 * elem_bytes(dtype))) == 0,
                "groupnorm_apply_seg: unsupported c=%d groups=%d (channels per group must be a multiple of 8)", c, groups);
  CM2_CHECK_ARG(num_seg >= 1 && num_seg <= CM2_MAX_SEG, "groupnorm_apply_seg: bad segment count %d", num_seg);
  GnRowSegs g;
  memset(&g, 0, sizeof(g));
  g.num = num_seg;
  long long end = 0;
  int img0 = 0;
  for (int i = 0; i < num_seg; ++i) {
    CM2_CHECK_ARG(seg[i].n > 0 && seg[i].h > 0 && seg[i].w > 0 && seg[i].row0 >= end, "groupnorm_apply_seg: bad segment %d", i);
    CM2_CHECK_ARG(seg[i].halo == 0 || (seg[i].halo == 1 && dtype == CM2_BF16 && 256 % (c / 8) == 0),
                  "groupnorm_apply_seg: segment %d: halo kind %d not supported here", i, seg[i].halo);
    const int fr = seg[i].halo == 1 ? 1 : 2;                      // shared zero frame: line pitch w + 1, image pitch (h + 1)(w + 1)
    const long long rows = (long long)seg[i].n * (seg[i].h + fr) * (seg[i].w + fr);
    CM2_CHECK_ARG(seg[i].row0 + rows < (1ll << 31) - 4096, "groupnorm_apply_seg: segment %d out of range", i);
    g.row0[i] = (int)seg[i].row0; g.rows[i] = (int)rows; g.pitch[i] = seg[i].w + fr; g.plane[i] = (seg[i].h + fr) * (seg[i].w + fr);
    g.h[i] = seg[i].h; g.w[i] = seg[i].w; g.img0[i] = img0;
    img0 += seg[i].n;
    end = seg[i].row0 + rows;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CM2_BF16 && 256 % (c / 8) == 0) {
    GnLineSegs gl;
    memset(&gl, 0, sizeof(gl));
    gl.num = num_seg;
    for (int i = 0; i < num_seg; ++i) {
      gl.row0[i] = g.row0[i]; gl.pitch[i] = g.pitch[i]; gl.plane[i] = g.plane[i]; gl.h[i] = g.h[i]; gl.w[i] = g.w[i];
      gl.img0[i] = g.img0[i];
      gl.line_prefix[i + 1] = gl.line_prefix[i] + seg[i].n * seg[i].h;
    }
    // 128-thread CTAs: measured 0.48 against 0.50 ms per step for 256 (more CTAs in flight per SM).  (They would also fit beside
    // a resident convolution CTA -- 58 registers x 128 against the 11.7 K the convolution leaves free -- but running the pass under
    // the other tower's convolution was measured slower: Engine.tower_overlap, off.)
    static const int apply_threads = getenv("CM2_GN_APPLY_THREADS") ? atoi(getenv("CM2_GN_APPLY_THREADS")) : 128;
    gn_seg_apply_lines_kernel<<<gl.line_prefix[num_seg], apply_threads, 0, s>>>((__nv_bfloat16*)x, c, c / groups, gl, stats, gamma, beta, eps, relu);
    CM2_CHECK_LAUNCH("gn_seg_apply_lines");
    return CM2_OK;
  }
  const int total_rows = (int)end;
  const long long nvec = (long long)ceil_div(total_rows, GN_ROWS_PER_THREAD) * (c / 8);
  CM2_CHECK_ARG(nvec < (1ll << 32) - (1ll << 24), "groupnorm_apply_seg: too many elements for 32-bit indexing");
  const int grid = (int)std::min<long long>(ceil_div64(nvec, 256), 148 * 32);
  if (dtype == CM2_F32)
    gn_seg_apply_stats_kernel<float><<<grid, 256, 0, s>>>((float*)x, c, c / groups, g, total_rows, stats, gamma, beta, eps, relu);
  else
    gn_seg_apply_stats_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)x, c, c / groups, g, total_rows, stats, gamma,
                                                                  beta, eps, relu);
  CM2_CHECK_LAUNCH("gn_seg_apply_stats");
  return CM2_OK;
}

extern "C" int cm2_groupnorm_apply_seg_split(const float* x, void* out_split, int32_t c, int32_t num_seg, const cm2_seg* seg,
                                             int32_t groups, const float* gamma, const float* beta, float eps, int32_t relu,
                                             const double* stats, void* stream) {
  CM2_CHECK_ARG(x && out_split && seg && gamma && beta && stats, "groupnorm_apply_seg_split: null pointer");
  CM2_CHECK_ARG(groups > 0 && c % groups == 0 && c % 8 == 0 && (c / groups) % 8 == 0 && c / 8 <= 256 && 256 % (c / 8) == 0 &&
                (reinterpret_cast<uintptr_t>(x) & 31) == 0 && (reinterpret_cast<uintptr_t>(out_split) & 15) == 0,
                "groupnorm_apply_seg_split: unsupported c=%d groups=%d", c, groups);
  CM2_CHECK_ARG(num_seg >= 1 && num_seg <= CM2_MAX_SEG, "groupnorm_apply_seg_split: bad segment count %d", num_seg);
  GnLineSegs gl;
  memset(&gl, 0, sizeof(gl));
  gl.num = num_seg;
  long long end = 0;
  int img0 = 0;
  for (int i = 0; i < num_seg; ++i) {
    CM2_CHECK_ARG(seg[i].n > 0 && seg[i].h > 0 && seg[i].w > 0 && seg[i].row0 >= end, "groupnorm_apply_seg_split: bad segment %d", i);
    CM2_CHECK_ARG(seg[i].halo == 0 || seg[i].halo == 1, "groupnorm_apply_seg_split: segment %d: halo kind %d not supported", i, seg[i].halo);
    const int fr = seg[i].halo == 1 ? 1 : 2;
    const long long rows = (long long)seg[i].n * (seg[i].h + fr) * (seg[i].w + fr);
    CM2_CHECK_ARG(seg[i].row0 + rows < (1ll << 31) - 4096, "groupnorm_apply_seg_split: segment %d out of range", i);
    gl.row0[i] = (int)seg[i].row0; gl.pitch[i] = seg[i].w + fr; gl.plane[i] = (seg[i].h + fr) * (seg[i].w + fr);
    gl.h[i] = seg[i].h; gl.w[i] = seg[i].w; gl.img0[i] = img0;
    gl.line_prefix[i + 1] = gl.line_prefix[i] + seg[i].n * seg[i].h;
    img0 += seg[i].n;
    end = seg[i].row0 + rows;
  }
  gn_seg_apply_split_lines_kernel<<<gl.line_prefix[num_seg], 256, 0, (cudaStream_t)stream>>>(
      x, reinterpret_cast<__half*>(out_split), c, c / groups, gl, stats, gamma, beta, eps, relu);
  CM2_CHECK_LAUNCH("gn_seg_apply_split_lines");
  return CM2_OK;
}

extern "C" int cm2_ese_gate_f64(const double* sums, double inv_count, const float* fc_w, const float* fc_b, float* gate,
                                int32_t n, int32_t c, void* stream) {
  CM2_CHECK_ARG(sums && fc_w && fc_b && gate, "ese_gate_f64: null pointer");
  if (n == 0) return CM2_OK;
  dim3 grid(ceil_div(c * 32, 256), n);
  ese_gate_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sums, inv_count, fc_w, fc_b, gate, c);
  CM2_CHECK_LAUNCH("ese_gate_f64");
  return CM2_OK;
}

extern "C" int cm2_ese_apply_pool(const cm2_act* x, const float* gate, const cm2_act* identity, const cm2_act* full,
                                  const cm2_act* pool, int32_t dtype, void* stream) {
  CM2_CHECK_ARG(x && x->data && gate, "ese_apply_pool: null pointer");
  CM2_CHECK_DTYPE(dtype, "ese_apply_pool");
  const int eb = elem_bytes(dtype);
  const bool has_full = full && full->data, has_pool = pool && pool->data;
  CM2_CHECK_ARG(has_full || has_pool, "ese_apply_pool: no output");
  CM2_CHECK_ARG(vec8_ok(*x, eb) && x->c % 8 == 0 && (reinterpret_cast<uintptr_t>(gate) & 15) == 0, "ese_apply_pool: bad x view (c=%d)", x->c);
  cm2_act idn, fu, po;
  memset(&idn, 0, sizeof(idn)); memset(&fu, 0, sizeof(fu)); memset(&po, 0, sizeof(po));
  if (identity && identity->data) {
    CM2_CHECK_ARG(vec8_ok(*identity, eb) && same_extent(*x, *identity), "ese_apply_pool: identity view mismatch");
    idn = *identity;
  }
  if (has_full) {
    CM2_CHECK_ARG(vec8_ok(*full, eb) && same_extent(*x, *full), "ese_apply_pool: full-resolution output view mismatch");
    fu = *full;
  }
  if (x->n == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (has_pool) {
    int ho = ceil_div(x->h - 3, 2) + 1, wo = ceil_div(x->w - 3, 2) + 1;      // MaxPool2d(3, 2, ceil_mode=True)
    if ((ho - 1) * 2 >= x->h) --ho;
    if ((wo - 1) * 2 >= x->w) --wo;
    CM2_CHECK_ARG(x->h >= 3 && x->w >= 3 && vec8_ok(*pool, eb) && pool->n == x->n && pool->c == x->c && pool->h == ho && pool->w == wo,
                  "ese_apply_pool: pooled output must be [%d,%d,%d,%d]", x->n, ho, wo, x->c);
    po = *pool;
    const long long total = (long long)po.n * po.h * po.w * (x->c / 8);
    CM2_CHECK_ARG(total < (1ll << 32) - (1ll << 24), "ese_apply_pool: too many elements for 32-bit indexing");
    const int c8 = x->c / 8;
    if (dtype == CM2_BF16 && c8 <= 256 && x->n <= 65535) {
      static const int max_threads = getenv("CM2_ESE_APPLY_THREADS") ? atoi(getenv("CM2_ESE_APPLY_THREADS")) : 256;
      const int threads = c8 * std::max(1, max_threads / c8);
      dim3 grid(po.h, x->n);
      View<const __nv_bfloat16> vx = make_view<const __nv_bfloat16>(*x), vi = make_view<const __nv_bfloat16>(idn);
      View<__nv_bfloat16> vf = make_view<__nv_bfloat16>(fu), vp = make_view<__nv_bfloat16>(po);
      if (idn.data && has_full) ese_apply_pool_lines_kernel<true, true><<<grid, threads, 0, s>>>(vx, gate, vi, vf, vp);
      else if (idn.data) ese_apply_pool_lines_kernel<true, false><<<grid, threads, 0, s>>>(vx, gate, vi, vf, vp);
      else if (has_full) ese_apply_pool_lines_kernel<false, true><<<grid, threads, 0, s>>>(vx, gate, vi, vf, vp);
      else ese_apply_pool_lines_kernel<false, false><<<grid, threads, 0, s>>>(vx, gate, vi, vf, vp);
      CM2_CHECK_LAUNCH("ese_apply_pool_lines");
      return CM2_OK;
    }
    const int grid = (int)std::min<long long>(ceil_div64(total, 256), 148 * 32);
    if (dtype == CM2_F32)
      ese_apply_pool_kernel<float><<<grid, 256, 0, s>>>(make_view<const float>(*x), gate, make_view<const float>(idn),
                                                        make_view<float>(fu), make_view<float>(po));
    else
      ese_apply_pool_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(make_view<const __nv_bfloat16>(*x), gate,
                                                                make_view<const __nv_bfloat16>(idn), make_view<__nv_bfloat16>(fu),
                                                                make_view<__nv_bfloat16>(po));
    CM2_CHECK_LAUNCH("ese_apply_pool");
    return CM2_OK;
  }
  // no pooling: flat streaming pass when everything is a whole halo buffer of the same geometry
  const bool flat = halo_view(*x) && halo_view(fu) && (!idn.data || halo_view(idn));
  CM2_CHECK_ARG(flat, "ese_apply_pool: without a pooled output x / identity / full must be interior views of halo-1 buffers "
                      "(use cm2_ese_apply for arbitrary views)");
  const long long plane = (long long)(x->h + 2) * (x->w + 2);
  const long long total8 = (long long)x->n * plane * (x->c / 8);
  CM2_CHECK_ARG(total8 < (1ll << 32) - (1ll << 24), "ese_apply_pool: too many elements for 32-bit indexing");
  const long long back = (x->sh + x->sw);                                    // elements from padded (0,0) to interior (0,0)
  const int grid = (int)std::min<long long>(ceil_div64(total8, 256), 148 * 32);
  if (dtype == CM2_F32)
    ese_apply_flat_kernel<float><<<grid, 256, 0, s>>>((const float*)x->data - back, gate, idn.data ? (const float*)idn.data - back : nullptr,
                                                      (float*)fu.data - back, x->c, (int)plane, total8);
  else
    ese_apply_flat_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x->data - back, gate,
                                                              idn.data ? (const __nv_bfloat16*)idn.data - back : nullptr,
                                                              (__nv_bfloat16*)fu.data - back, x->c, (int)plane, total8);
  CM2_CHECK_LAUNCH("ese_apply_flat");
  return CM2_OK;
}
