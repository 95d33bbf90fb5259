// Generic implicit-GEMM convolution on CUDA cores (fp32 accumulate).
//
// This is the engine of the strict-fp32 variant (parity gate: kept-detection index sets identical to
// the oracle) and the fallback for the few layers the tcgen05 engine does not take (Cin = 3 stem
// conv, stride-2 convs until the strided-TMA path is validated).  Any kernel size / stride / padding,
// any channel counts, virtual channel-concat over up to CM2_MAX_SRC sources, fused scale/shift
// (folded FrozenBN or bias), residual / nearest-2x-upsample add, ReLU, and the 2x2 transposed-conv
// scatter store.
//
// GEMM view: M = n*ho*wo output pixels, N = cout, K = kh*kw*cin_total.  64x64 tile per 256-thread
// CTA, BK = 16, 4x4 register micro-tile, register-staged double buffering.
#include "common.cuh"

namespace cm2 {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct SimtConvParams {
  const void* src[CM2_MAX_SRC];
  int src_c[CM2_MAX_SRC];
  int src_off[CM2_MAX_SRC + 1];
  long long src_sn[CM2_MAX_SRC], src_sh[CM2_MAX_SRC], src_sw[CM2_MAX_SRC];
  int num_src;
  int n, h, w, cin, cout, kh, kw, stride, pad, ho, wo;
  int M, K;
  const void* weight;
  const float* scale;
  const float* shift;
  int relu, in_relu;
  const void* residual;
  long long res_sn, res_sh, res_sw;
  int res_mode, out_mode;
  void* out;
  long long out_sn, out_sh, out_sw;
};

template <typename T, typename OutT>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const SimtConvParams p) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  __shared__ int pix_n[BM], pix_y[BM], pix_x[BM];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  if (tid < BM) {
    int m = m0 + tid;
    if (m < p.M) {
      int hw = p.ho * p.wo;
      int n = m / hw;
      int r = m - n * hw;
      int oy = r / p.wo;
      int ox = r - oy * p.wo;
      pix_n[tid] = n;
      pix_y[tid] = oy * p.stride - p.pad;
      pix_x[tid] = ox * p.stride - p.pad;
    } else {
      pix_n[tid] = -1;
      pix_y[tid] = 0;
      pix_x[tid] = 0;
    }
  }
  __syncthreads();

  const int a_kk = tid & 15;   // k within chunk handled by this thread for A
  const int a_r0 = tid >> 4;   // first pixel row (then +16, +32, +48)
  const int b_kk = tid >> 4;   // k row for B
  const int b_c0 = (tid & 15) * 4;

  float a_reg[4], b_reg[4];

  auto load_chunk = [&](int k0) {
    // ---- A: gather from the (virtually concatenated) sources
    int kg = k0 + a_kk;
    bool kvalid = kg < p.K;
    int tap = 0, cc = 0, s = 0, ky = 0, kx = 0;
    const T* base = nullptr;
    long long sn = 0, sh = 0, sw = 0;
    if (kvalid) {
      tap = kg / p.cin;
      cc = kg - tap * p.cin;
      ky = tap / p.kw;
      kx = tap - ky * p.kw;
#pragma unroll
      for (int i = 1; i < CM2_MAX_SRC; ++i)
        if (i < p.num_src && cc >= p.src_off[i]) s = i;
      cc -= p.src_off[s];
      sn = p.src_sn[s]; sh = p.src_sh[s]; sw = p.src_sw[s];
      base = reinterpret_cast<const T*>(p.src[s]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = a_r0 + 16 * j;
      float v = 0.f;
      int n = pix_n[r];
      if (kvalid && n >= 0) {
        int iy = pix_y[r] + ky;
        int ix = pix_x[r] + kx;
        if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
          v = to_f32<T>(base[n * sn + iy * sh + ix * sw + cc]);
          if (p.in_relu) v = fmaxf(v, 0.f);
        }
      }
      a_reg[j] = v;
    }
    // ---- B: weights [K][cout]
    int kb = k0 + b_kk;
    const T* wptr = reinterpret_cast<const T*>(p.weight);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int co = n0 + b_c0 + i;
      b_reg[i] = (kb < p.K && co < p.cout) ? to_f32<T>(wptr[(size_t)kb * p.cout + co]) : 0.f;
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) As[buf][a_kk][a_r0 + 16 * j] = a_reg[j];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[buf][b_kk][b_c0 + i] = b_reg[i];
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nchunks = (p.K + BK - 1) / BK;
  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    int buf = c & 1;
    if (c + 1 < nchunks) load_chunk((c + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (c + 1 < nchunks) store_chunk(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  const T* res = reinterpret_cast<const T*>(p.residual);
  OutT* out = reinterpret_cast<OutT*>(p.out);
  const int cq_n = p.cout >> 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = ty * 4 + i;
    int n = pix_n[r];
    if (n < 0) continue;
    int m = m0 + r;
    int hw = p.ho * p.wo;
    int rem = m - n * hw;
    int oy = rem / p.wo;
    int ox = rem - oy * p.wo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co >= p.cout) continue;
      float v = acc[i][j];
      if (p.scale) v *= __ldg(p.scale + co);
      if (p.shift) v += __ldg(p.shift + co);
      if (p.res_mode == 1) {
        v += to_f32<T>(res[n * p.res_sn + oy * p.res_sh + ox * p.res_sw + co]);
      } else if (p.res_mode == 2) {
        v += to_f32<T>(res[n * p.res_sn + (oy >> 1) * p.res_sh + (ox >> 1) * p.res_sw + co]);
      }
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.out_mode == 0) {
        out[n * p.out_sn + oy * p.out_sh + ox * p.out_sw + co] = from_f32<OutT>(v);
      } else {
        int q = co / cq_n;
        int cq = co - q * cq_n;
        int dy = q >> 1, dx = q & 1;
        out[n * p.out_sn + (2 * oy + dy) * p.out_sh + (2 * ox + dx) * p.out_sw + cq] = from_f32<OutT>(v);
      }
    }
  }
}

void conv_out_extent(const cm2_conv_desc* d, int* ho, int* wo) {
  if (d->src_phase) { *ho = d->src[0].h; *wo = d->src[0].w; return; }   // plane extent == output extent
  *ho = (d->src[0].h + 2 * d->pad - d->kh) / d->stride + 1;
  *wo = (d->src[0].w + 2 * d->pad - d->kw) / d->stride + 1;
}

int conv_simt_launch(const cm2_conv_desc* d, cudaStream_t stream) {
  SimtConvParams p;
  int off = 0;
  for (int i = 0; i < CM2_MAX_SRC; ++i) {
    bool on = i < d->num_src;
    p.src[i] = on ? d->src[i].data : nullptr;
    p.src_c[i] = on ? d->src[i].c : 0;
    p.src_sn[i] = on ? d->src[i].sn : 0;
    p.src_sh[i] = on ? d->src[i].sh : 0;
    p.src_sw[i] = on ? d->src[i].sw : 0;
    p.src_off[i] = off;
    off += p.src_c[i];
  }
  p.src_off[CM2_MAX_SRC] = off;
  p.num_src = d->num_src;
  int ho, wo;
  conv_out_extent(d, &ho, &wo);
  p.n = d->src[0].n; p.h = d->src[0].h; p.w = d->src[0].w; p.cin = off; p.cout = d->cout;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.ho = ho; p.wo = wo;
  int64_t M = (int64_t)p.n * ho * wo;
  CM2_CHECK_ARG(M < (1ll << 31) && (int64_t)M * d->cout < (1ll << 40), "conv_simt: problem too large");
  p.M = (int)M;
  p.K = d->kh * d->kw * off;
  p.weight = d->weight; p.scale = d->scale; p.shift = d->shift;
  p.relu = d->relu; p.in_relu = d->in_relu;
  p.residual = d->residual.data; p.res_mode = d->residual.data ? d->res_mode : 0;
  p.res_sn = d->residual.sn; p.res_sh = d->residual.sh; p.res_sw = d->residual.sw;
  p.out_mode = d->out_mode; p.out = d->out.data;
  p.out_sn = d->out.sn; p.out_sh = d->out.sh; p.out_sw = d->out.sw;
  if (p.M == 0) return CM2_OK;
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.cout, BN));
  if (d->dtype == CM2_F32 && d->out_dtype == CM2_F32)
    conv_simt_kernel<float, float><<<grid, NT, 0, stream>>>(p);
  else if (d->dtype == CM2_BF16 && d->out_dtype == CM2_BF16)
    conv_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, NT, 0, stream>>>(p);
  else if (d->dtype == CM2_BF16 && d->out_dtype == CM2_F32)
    conv_simt_kernel<__nv_bfloat16, float><<<grid, NT, 0, stream>>>(p);
  else {
    set_error("conv_simt: unsupported dtype combination %d -> %d", d->dtype, d->out_dtype);
    return CM2_ERR_UNSUPPORTED;
  }
  CM2_CHECK_LAUNCH("conv_simt");
  return CM2_OK;
}

int conv_tc_launch(const cm2_conv_desc* d, cudaStream_t stream);  // conv_tc.cu

}  // namespace cm2

extern "C" int cm2_conv2d(const cm2_conv_desc* d, void* stream) {
  using namespace cm2;
  CM2_CHECK_ARG(d != nullptr, "conv2d: null descriptor");
  CM2_CHECK_ARG(d->num_src >= 1 && d->num_src <= CM2_MAX_SRC, "conv2d: num_src %d out of range", d->num_src);
  CM2_CHECK_ARG(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->pad >= 0 && d->cout > 0, "conv2d: bad kernel geometry");
  if (d->num_seg > 0) {                      // segmented tensors: validated and run by the TC engine only
    CM2_CHECK_ARG(d->num_seg <= CM2_MAX_SEG && d->engine == CM2_ENGINE_TC && d->out_mode == 0 && !d->residual.data &&
                  !d->src_phase && d->weight && d->out.data, "conv2d: bad segmented descriptor");
    return conv_tc_launch(d, (cudaStream_t)stream);
  }
  const cm2_act& s0 = d->src[0];
  CM2_CHECK_ARG(s0.n >= 0 && s0.h > 0 && s0.w > 0, "conv2d: bad extents");
  for (int i = 0; i < d->num_src; ++i) {
    const cm2_act& s = d->src[i];
    CM2_CHECK_ARG(s.data != nullptr && s.c > 0, "conv2d: source %d invalid", i);
    CM2_CHECK_ARG(s.n == s0.n && s.h == s0.h && s.w == s0.w, "conv2d: source %d extent differs from source 0", i);
  }
  int ho, wo;
  conv_out_extent(d, &ho, &wo);
  CM2_CHECK_ARG(ho > 0 && wo > 0, "conv2d: empty output");
  CM2_CHECK_ARG(d->out_mode == 0 || ((d->out_mode == 1 || d->out_mode == 3) && d->cout % 4 == 0) || d->out_mode == 2,
                "conv2d: bad out_mode");
  CM2_CHECK_ARG(d->weight != nullptr && d->out.data != nullptr, "conv2d: null weight/out");
  // split-precision output (CM2_F16 out of CM2_F16 sources): [hi | lo] pair, 2 * cout channels
  const int out_c = (d->dtype == CM2_F16 && d->out_dtype == CM2_F16) ? 2 * d->cout : d->cout;
  if (d->out_mode == 0)
    CM2_CHECK_ARG(d->out.n == s0.n && d->out.h == ho && d->out.w == wo && d->out.c == out_c,
                  "conv2d: out view [%d,%d,%d,%d] != expected [%d,%d,%d,%d]", d->out.n, d->out.h, d->out.w, d->out.c,
                  s0.n, ho, wo, d->cout);
  else if (d->out_mode == 2)
    CM2_CHECK_ARG(d->out.n == s0.n && d->out.h == (ho + 1) / 2 && d->out.w == (wo + 1) / 2 && d->out.c == out_c,
                  "conv2d: phase-split out view [%d,%d,%d,%d] != expected [%d,%d,%d,%d]", d->out.n, d->out.h, d->out.w,
                  d->out.c, s0.n, (ho + 1) / 2, (wo + 1) / 2, d->cout);
  else if (d->out_mode == 3)
    CM2_CHECK_ARG(d->out.n == s0.n && d->out.h == 2 * ho && d->out.w == 2 * wo && d->out.c == 1 && d->engine == CM2_ENGINE_TC,
                  "conv2d: fused deconv + predictor (TC engine) writes [%d,%d,%d,1]", s0.n, 2 * ho, 2 * wo);
  else
    CM2_CHECK_ARG(d->out.n == s0.n && d->out.h == 2 * ho && d->out.w == 2 * wo && d->out.c == d->cout / 4,
                  "conv2d: deconv out view [%d,%d,%d,%d] != expected [%d,%d,%d,%d]", d->out.n, d->out.h, d->out.w,
                  d->out.c, s0.n, 2 * ho, 2 * wo, d->cout / 4);
  if (d->residual.data != nullptr) {
    const cm2_act& r = d->residual;
    CM2_CHECK_ARG(d->out_mode == 0 && r.c == d->cout && r.n == s0.n, "conv2d: residual channel/batch mismatch");
    CM2_CHECK_ARG((d->res_mode == 1 && r.h == ho && r.w == wo) ||
                  (d->res_mode == 2 && ho % 2 == 0 && wo % 2 == 0 && r.h == ho / 2 && r.w == wo / 2),
                  "conv2d: residual extent %dx%d incompatible with res_mode %d and output %dx%d", r.h, r.w, d->res_mode,
                  ho, wo);
  }
  if (d->engine == CM2_ENGINE_SIMT) {
    CM2_CHECK_ARG(d->stats == nullptr, "conv2d: fused statistics need the TC engine");
    if (d->src_phase || d->out_mode == 2) {
      set_error("conv2d: phase-split layouts need the TC engine");
      return CM2_ERR_UNSUPPORTED;
    }
    return conv_simt_launch(d, (cudaStream_t)stream);
  }
  if (d->engine == CM2_ENGINE_TC) return conv_tc_launch(d, (cudaStream_t)stream);
  set_error("conv2d: unknown engine %d", d->engine);
  return CM2_ERR_UNSUPPORTED;
}
