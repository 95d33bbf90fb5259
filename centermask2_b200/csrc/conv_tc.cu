// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA into 128B-swizzled shared memory) -- sm_100a only.
//
// GEMM view.  Activations live in NHWC buffers with a one-pixel zero halo, [n, h+2, w+2, c], which this
// kernel treats as a *flat* 2-D matrix  A[rows = n*(h+2)*(w+2)][c].  For a stride-1 3x3 convolution the
// operand of tap (ky, kx) is the same matrix shifted by (ky-1)*(w+2) + (kx-1) rows, so every K-block is
// ONE plain 2-D TMA box load [128 rows x 64 channels] at row coordinate m0 + shift (out-of-range rows
// are zero-filled by TMA).  Outputs are produced for every padded position; the epilogue writes zeros
// to the halo rows, which keeps the invariant "halo == 0" for the next layer.  1x1 convolutions and
// linear layers use the same path with a single tap; a channel concat (OSA aggregation, vovnet.py:324)
// is a K loop over several source matrices (one tensor map each) and is never materialised.
//
//   D[128 x BN] (fp32, TMEM)  +=  A_tile[128 x 64] (bf16, smem, K-major SW128)  x  W_tile[BN x 64]^T
//
// Warp roles (192 threads, one CTA per SM, persistent over output tiles):
//   warp 0     TMA producer (one elected lane): A + W boxes per K-block into a STAGES-deep ring
//   warp 1     TMEM allocation + MMA issue (one elected lane): 4 x tcgen05.mma (K=16) per K-block,
//              tcgen05.commit releases the smem slot / publishes the accumulator
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 16 columns), scale/shift (folded FrozenBN or bias),
//              residual / nearest-2x upsample add, ReLU, halo masking, bf16/f32 store, deconv scatter.
// Two accumulator stages (2 x 256 TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <algorithm>

namespace cm2 {

constexpr int TC_BM = 128;          // rows per tile (UMMA M)
constexpr int TC_BK = 64;           // bf16 channels per K-block = 128 bytes = one swizzle row
constexpr int TC_ACC_COLS = 256;    // TMEM columns per accumulator stage
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;

struct alignas(64) TcParams {
  CUtensorMap a_map[CM2_MAX_SRC];
  CUtensorMap b_map;
  int num_src;
  int src_c[CM2_MAX_SRC];
  int taps;                 // 1 or 9
  int tap_shift[9];         // row shift of every tap in the flat source matrix
  int pitch;                // w + 2 (rows per padded image line); 0 in dense mode
  int plane;                // (h + 2) * (w + 2); h*w in dense mode
  int h, w, halo;           // halo: 1 = padded geometry, 0 = dense rows
  int rows;                 // total GEMM rows
  int m_tiles, n_tiles, bn, cout;
  int stages;
  const float* scale;
  const float* shift;
  int relu;
  // output
  void* out;
  long long out_sn, out_sh, out_sw;
  long long out_plane;      // out_mode 2: element stride between the four phase planes
  int out_f32, out_halo, out_mode, out_vec;
  // residual
  const __nv_bfloat16* res;
  long long res_sn, res_sh, res_sw;
  int res_mode;
  // ---- v2 kernel (256-row tiles, separate A / B rings, optional kx-merged A slabs)
  int kx_merge;             // 1: one A slab per (ky, source, k-block) serves the three kx taps
  int a_box_rows;           // rows per A TMA box: 128, or 136 with kx_merge (two boxes per slab)
  int sa_stages, sb_stages; // ring depths
  int spin;                 // control warps poll their barriers with test_wait instead of the suspending try_wait
  int b_resident;           // v2: all weight tiles of the (single) N tile stay in shared memory for the whole kernel
  int acc_stages;           // 2 when bn <= 128 (2 x 2 x 128 TMEM columns), else 1
  int nblk_total;           // sum over sources of ceil(c / 64)
  int desc_mode;            // 0: base_offset field 0;  1: base_offset = (start >> 7) & 7 for unaligned starts
  int phase;                // 1: sources are stride-2 phase planes (tap_shift holds plane + line offsets)
  int num_seg;              // > 0: segmented halo tensor (several maps of different extent in one flat buffer)
  int seg_row0[CM2_MAX_SEG], seg_rows[CM2_MAX_SEG], seg_pitch[CM2_MAX_SEG], seg_plane[CM2_MAX_SEG];
  int seg_h[CM2_MAX_SEG], seg_w[CM2_MAX_SEG];
  double* stats;            // fused output statistics (see cm2_conv_desc.stats_mode); nullptr: off
  int stats_mode;           // 1: per (image, channel) sum;  2: per (image, 8-channel chunk) sum and sum of squares
  int stats_stride;         // doubles per image
  int seg_img0[CM2_MAX_SEG];// global image index of the first image of every segment
  double rcp_plane, rcp_pitch;                       // 1 / plane, 1 / pitch (1 / w in dense mode): exact fast division
  double seg_rcp_plane[CM2_MAX_SEG], seg_rcp_pitch[CM2_MAX_SEG];
  // out_mode 3: class-gathered mask predictor fused behind the 2x2 transposed conv
  const float* pred_w; const float* pred_b; const long long* pred_cls; int pred_ncls;
  int epi_kind;             // staged-epilogue variant (see tc_epilogue_dispatch)
  int epi_sets;             // column sets of epilogue warps per 128-row accumulator (1, 2 or 4)
  int fast_store;           // 1: epilogue transposes through shared memory and writes 64-byte row segments
  int dbg;                  // tuning experiments (CM2_TC_DEBUG): 1 no epilogue stores, 2 no TMA loads, 4 no MMAs, 8 no epilogue body
  int row_begin;            // v1 only: first GEMM row covered by a tile (leading halo rows trimmed to save a wave of tiles)
  int pair;                 // v2 only: 1 = launched as CTA pairs, cta_group::2 MMAs (see conv_tc2_kernel<.., true>)
  int variant;              // host only: 1 = conv_tc_kernel (128-row tiles), 2 = conv_tc2_kernel
  unsigned smem_bytes;      // host only: dynamic shared memory of the launch
  uint32_t idesc_ab;        // operand format bits of the instruction descriptor: bf16 (1 << 7 | 1 << 10) or f16 (0)
  int res_f32;              // the residual tensor is fp32 (split-precision convolutions), else bf16
  // ---- v3 kernel (split precision): src_c[] holds the LOGICAL channel counts, a source matrix is [rows][hi(c) | lo(c)]
  int split;                // 1: conv_tc3_kernel
  int chunk;                // main-accumulator K-blocks between two drains into the fp32 register sums
  int cout_pad;             // W_lo tiles start cout_pad rows below the W_hi tiles in the weight matrix
  int split_out;            // 1: the output is the [hi | lo] f16 pair of 2 * cout channels (epilogue kind 11)
  // ---- split-K (v1 kernel): tile t = (k-slice, m tile, n tile); slice ks accumulates K-blocks [ks * kb_per_split, ..) and stores
  // its fp32 partial sums split_stride elements behind the previous slice's (cm2_conv_desc.splitk; splitk_finish_kernel adds them)
  // trimmed tile range (row_begin > 0): the halo rows of the output outside it are zeroed by the first / last CTA's epilogue warps
  char* trim_base;          // address of flat output row 0; nullptr: nothing to zero
  long long trim_lo_bytes, trim_hi_off, trim_hi_bytes;
  int ksplit;               // 0 / 1: off
  int kb_per_split;
  long long split_stride;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 1023u) == 0) {          // a protocol bug must fail loudly, not hang the GPU box
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) {  // ~2 s
        printf("conv_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
// One lane of a fully converged warp.  The control warps run their loops warp-uniformly and predicate only the
// TMA / MMA / commit instructions with this: their operands then live in uniform registers and each UTCHMMA /
// UTMALDG is a single instruction.  (Wrapping the loops in `if (lane == 0)` made ptxas emit an ELECT / BRA.U.ANY
// uniformisation loop plus R2UR moves around every one of them -- ~170 cycles per MMA in the ncu source view.)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}
// Polling wait (mbarrier.test_wait never suspends the thread): used by the two control warps, whose hand-off latency
// bounds short pipeline stages.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 0xfffffu) == 0) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) {
        printf("conv_tc: mbarrier spin wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait_ctl(int spin, uint32_t bar, uint32_t parity) {
  if (spin) mbar_wait_spin(bar, parity);
  else mbar_wait(bar, parity);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- cta_group::2 (CTA pair) forms.  Shared-memory addresses of a cluster launch carry the CTA rank in bit 24;
// clearing it addresses the same offset in the even (leader) CTA of the pair.
constexpr uint32_t TC_PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA box into THIS CTA's shared memory, transaction bytes credited to the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & TC_PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the same barrier of both CTAs of the pair when the MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & TC_PEER_MASK) : "memory");
}
// D[256 x N] += A[256 x 16] * B[N x 16]^T over the CTA pair: rows 0-127 / A and B[0, N/2) from the leader's shared
// memory, rows 128-255 / B[N/2, N) from the peer's (same offsets); each CTA's TMEM holds its 128 rows
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool PAIR> __device__ __forceinline__ void tc_mma_x(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR) tc_mma_bf16_pair(d, a, b, idesc, acc); else tc_mma_bf16(d, a, b, idesc, acc);
}
template <bool PAIR> __device__ __forceinline__ void tc_commit_x(uint32_t bar) {
  if (PAIR) tc_commit_pair(bar); else tc_commit(bar);
}
template <bool PAIR> __device__ __forceinline__ void tma_load_2d_x(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  if (PAIR) tma_load_2d_pair(dst, map, bar, c0, c1); else tma_load_2d(dst, map, bar, c0, c1);
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile (rows of 128 bytes, 8-row groups 1024 bytes apart):
//   start address >> 4 | LBO (ignored for swizzled K-major) = 1 | SBO = 1024 >> 4 | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// Geometry of the map a tile belongs to (segments start at multiples of 256 rows, so a tile never straddles two).
struct TileGeom { int row0, rows, pitch, plane, h, w, img0; double rcp_plane, rcp_pitch; long long out_extra; };
__device__ __forceinline__ TileGeom tc_geom(const TcParams& p, int m0) {
  TileGeom g;
  g.out_extra = 0;
  if (p.num_seg == 0) {
    g.row0 = 0; g.rows = p.rows; g.pitch = p.pitch; g.plane = p.plane; g.h = p.h; g.w = p.w; g.img0 = 0;
    g.rcp_plane = p.rcp_plane; g.rcp_pitch = p.rcp_pitch;
  } else {
    int s = 0;
#pragma unroll
    for (int i = 1; i < CM2_MAX_SEG; ++i)
      if (i < p.num_seg && m0 >= p.seg_row0[i]) s = i;
    g.row0 = p.seg_row0[s]; g.rows = p.seg_rows[s]; g.pitch = p.seg_pitch[s]; g.plane = p.seg_plane[s];
    g.h = p.seg_h[s]; g.w = p.seg_w[s]; g.img0 = p.seg_img0[s];
    g.rcp_plane = p.seg_rcp_plane[s]; g.rcp_pitch = p.seg_rcp_pitch[s];
  }
  return g;
}
__device__ __forceinline__ int tc_tap_shift(const TcParams& p, int tap, int pitch) {
  if (p.taps == 1) return 0;
  if (p.phase) return p.tap_shift[tap];
  return (tap / 3 - 1) * pitch + (tap % 3 - 1);
}

// Per-tile epilogue vectors in shared memory: the epilogue warps cooperatively copy scale/shift of the tile's
// bn columns (1 / 0 where absent or beyond cout) and meet on a named barrier; per-element __ldg in the
// epilogue loop showed up as the top stall (long scoreboard on every FMUL) in ncu.
__device__ __forceinline__ void tc_stage_scale_shift(const TcParams& p, uint32_t ss_smem, int n0, int tid_e, int n_epi, int m0) {
  const float* pred_row = nullptr;
  if (p.out_mode == 3) {                             // rows of a tile belong to one ROI (plane rows per ROI, plane % 128 == 0)
    const int roi = m0 / p.plane;
    int cls = p.pred_ncls == 1 ? 0 : (int)p.pred_cls[roi];
    cls = min(max(cls, 0), p.pred_ncls - 1);
    pred_row = p.pred_w + (size_t)cls * (p.cout >> 2);
  }
  for (int i = tid_e; i < p.bn; i += n_epi) {
    const int co = n0 + i;
    float sc = 1.f, sh = 0.f;
    if (co < p.cout) {
      if (p.scale) sc = __ldg(p.scale + co);
      if (p.shift) sh = __ldg(p.shift + co);
      if (pred_row) sc = __ldg(pred_row + (co % (p.cout >> 2)));
    }
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(ss_smem + 4u * i), "f"(sc) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(ss_smem + 1024u + 4u * i), "f"(sh) : "memory");
  }
  asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
constexpr uint32_t EPI_SS_BYTES = 2 * 2048;         // two tile parities x (scale[256] + shift[256])

// ------------------------------------------------------------------------------------------------
// epilogue of one accumulator (128 rows x bn columns): the calling warp owns TMEM lanes [32q, 32q+32),
// i.e. GEMM rows m = tile_row0 + 32q + lane.  `taddr` = TMEM address of (lane 32q, column 0 of the tile).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_epilogue_rows(const TcParams& p, const TileGeom& g, uint32_t taddr, int m, int n0,
                                                 uint32_t ss_smem) {
      // decode the GEMM row into (image, y, x) of the *unpadded* feature map
  const int mr = m - g.row0;
  bool in_range = mr >= 0 && mr < g.rows, interior = false;
  int img = 0, y = 0, x = 0;
  if (in_range) {
    img = mr / g.plane;
    int r = mr - img * g.plane;
    if (p.halo) {
      int yy = r / g.pitch, xx = r - yy * g.pitch;
      y = yy - 1; x = xx - 1;
      interior = y >= 0 && y < g.h && x >= 0 && x < g.w;
    } else {
      y = r / g.w; x = r - y * g.w;
      interior = true;
    }
  }
  // rows that get written: interior rows always; halo rows (as zeros) only when the output keeps the halo
  const bool do_store = in_range && (interior || (p.out_halo && p.out_mode == 0)) && !(p.dbg & 1);
  long long out_off;
  if (p.num_seg)
    out_off = (long long)m * p.out_sw;                 // segmented output: same flat row, pitch = cout
  else if (p.out_mode == 0)
    out_off = (long long)img * p.out_sn + (long long)y * p.out_sh + (long long)x * p.out_sw;
  else if (p.out_mode == 1)       // 2x2 transposed-conv scatter: quadrant offset added per column chunk
    out_off = (long long)img * p.out_sn + (long long)(2 * y) * p.out_sh + (long long)(2 * x) * p.out_sw;
  else                            // phase-split store for a following stride-2 convolution
    out_off = (long long)((y & 1) * 2 + (x & 1)) * p.out_plane + (long long)img * p.out_sn +
              (long long)(y >> 1) * p.out_sh + (long long)(x >> 1) * p.out_sw;
  const __nv_bfloat16* res_row = nullptr;
  const float* res_row_f = nullptr;
  if (p.res_mode && interior) {
    const long long roff = (long long)img * p.res_sn + (long long)(p.res_mode == 2 ? (y >> 1) : y) * p.res_sh +
                           (long long)(p.res_mode == 2 ? (x >> 1) : x) * p.res_sw;
    if (p.res_f32) res_row_f = reinterpret_cast<const float*>(p.res) + roff;
    else res_row = p.res + roff;
  }

  for (int c0 = 0; c0 < p.bn; c0 += 16) {
    uint32_t raw[16];
    __syncwarp();                                       // tcgen05.ld is warp-collective (.sync.aligned)
    tc_ld16(taddr + (uint32_t)c0, raw);
    tc_ld_wait();
    const int co0 = n0 + c0;
    if (!do_store || co0 >= p.cout) continue;
    float v[16];
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const float4 sc = lds_f4(ss_smem + 4u * (uint32_t)(c0 + 4 * j4)), sh = lds_f4(ss_smem + 1024u + 4u * (uint32_t)(c0 + 4 * j4));
      v[4 * j4 + 0] = fmaf(__uint_as_float(raw[4 * j4 + 0]), sc.x, sh.x);
      v[4 * j4 + 1] = fmaf(__uint_as_float(raw[4 * j4 + 1]), sc.y, sh.y);
      v[4 * j4 + 2] = fmaf(__uint_as_float(raw[4 * j4 + 2]), sc.z, sh.z);
      v[4 * j4 + 3] = fmaf(__uint_as_float(raw[4 * j4 + 3]), sc.w, sh.w);
    }
    if (res_row_f) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (co0 + j < p.cout) v[j] += __ldg(res_row_f + co0 + j);
    }
    if (res_row) {
      if (p.out_vec) {
        float r0[8], r1[8];
        Vec8<__nv_bfloat16>::load(res_row + co0, r0);
        Vec8<__nv_bfloat16>::load(res_row + co0 + 8, r1);
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] += r0[j]; v[8 + j] += r1[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (co0 + j < p.cout) v[j] += __bfloat162float(res_row[co0 + j]);
      }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (!interior) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    long long off = out_off;
    int cc = co0;
    if (p.out_mode == 1) {                              // 2x2 transposed-conv scatter (sam.py:74-80)
      const int cq = p.cout >> 2;
      const int quad = co0 / cq;
      cc = co0 - quad * cq;
      off += (long long)(quad >> 1) * p.out_sh + (long long)(quad & 1) * p.out_sw;
    }
    if (p.out_f32) {
      float* o = reinterpret_cast<float*>(p.out) + off + cc;
      if (p.out_vec) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (co0 + j < p.cout) o[j] = v[j];
      }
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off + cc;
      if (p.out_vec) {
        float lo[8], hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
        Vec8<__nv_bfloat16>::store(o, lo);
        Vec8<__nv_bfloat16>::store(o + 8, hi);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (co0 + j < p.cout) o[j] = __float2bfloat16_rn(v[j]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused output statistics (staged epilogue only).  v[0..31] = this row's 32 stored values of the pass (zeros for
// halo / out-of-range rows).  Rows of a warp usually belong to one image; the loop handles tiles that straddle
// images.  Values are reduced across the warp with a halving butterfly (each exchange halves the number of live
// values), then added to the fp64 accumulators -- fp64 keeps the result independent of the atomic order to ~1e-16,
// i.e. bit-identical after the consumer's conversion to fp32.
//   mode 1 (eSE pool, vovnet.py:254): stats[img][c]         += sum over pixels
//   mode 2 (GroupNorm, fcos.py:182):  stats[img][c/8][2]    += (sum, sum of squares) over pixels x 8 channels
// ------------------------------------------------------------------------------------------------
template <int STATS>
__device__ __forceinline__ void tc_epilogue_stats(const TcParams& p, const float (&v)[32], unsigned rows, bool interior,
                                                  int img_g, int co0, int ncol, int lane) {
  while (rows) {                                     // warp-uniform
    const int leader = __ffs(rows) - 1;
    const int li = __shfl_sync(0xffffffffu, img_g, leader);
    const bool mine = interior && img_g == li;
    rows &= ~__ballot_sync(0xffffffffu, mine);
    if (STATS == 2) {
      float r[8];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float sa = 0.f, sq = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float x = mine ? v[8 * ch + k] : 0.f; sa += x; sq = fmaf(x, x, sq); }
        r[2 * ch] = sa; r[2 * ch + 1] = sq;
      }
      // 8 -> 4 -> 2 -> 1 live values over lane bits 4, 3, 2; then a plain sum over bits 1, 0
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool up = lane & 16;
        const float keep = up ? r[i + 4] : r[i], send = up ? r[i] : r[i + 4];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool up = lane & 8;
        const float keep = up ? r[i + 2] : r[i], send = up ? r[i] : r[i + 2];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      {
        const bool up = lane & 4;
        const float keep = up ? r[1] : r[0], send = up ? r[0] : r[1];
        r[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      r[0] += __shfl_xor_sync(0xffffffffu, r[0], 2);
      r[0] += __shfl_xor_sync(0xffffffffu, r[0], 1);
      const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);    // value index = chunk * 2 + {sum, sq}
      if ((lane & 3) == 0 && (k >> 1) * 8 < ncol)
        atomicAdd(p.stats + (size_t)li * p.stats_stride + (size_t)((co0 >> 3) + (k >> 1)) * 2 + (k & 1), (double)r[0]);
    } else {
      float r[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) r[k] = mine ? v[k] : 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const bool up = lane & 16;
        const float keep = up ? r[i + 16] : r[i], send = up ? r[i] : r[i + 16];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool up = lane & 8;
        const float keep = up ? r[i + 8] : r[i], send = up ? r[i] : r[i + 8];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool up = lane & 4;
        const float keep = up ? r[i + 4] : r[i], send = up ? r[i] : r[i + 4];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool up = lane & 2;
        const float keep = up ? r[i + 2] : r[i], send = up ? r[i] : r[i + 2];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      {
        const bool up = lane & 1;
        const float keep = up ? r[1] : r[0], send = up ? r[0] : r[1];
        r[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
      // lane now holds the column sum of channel index (bit4,bit3,bit2,bit1,bit0) = lane
      if (lane < ncol) atomicAdd(p.stats + (size_t)li * p.stats_stride + (size_t)(co0 + lane), (double)r[0]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Coalescing epilogue.  In the TMEM register layout a thread owns one output ROW, so direct stores make
// every warp instruction touch 32 different rows 16 bytes at a time (measured: the stores alone cost as
// much as all MMAs of a 3x3 256->256 layer).  Here each pass moves a [32 rows x 64 bytes] block through
// a per-warp staging buffer (row pitch 80 B: conflict-free 16-byte st.shared) and writes it back as
// 8 rows x 64 contiguous bytes per instruction.  Used when the output is vectorisable and there is no
// residual; everything else takes tc_epilogue_rows.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t EPI_PITCH = 80;                  // bytes per staged row (64 payload + 16 pad)
constexpr uint32_t EPI_WARP_BYTES = 32 * EPI_PITCH;

// two fp32 FMAs in one instruction (FFMA2)
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

// The epilogue warps of one 128-row accumulator are organised in `ncset` column sets (one warp per TMEM lane quarter
// and set); set `cset` handles every ncset-th 64-byte column pass.  Several warps per scheduler hide the
// tcgen05.ld -> FMA -> st.shared -> ld.shared -> st.global dependency chain, which is what bounds short-K layers.
// Compile-time variants (runtime flags in the pass loop cost ~150 of its 250 SASS instructions -- measured with ncu on
// the K = 32 stem layer, whose epilogue is the whole kernel): F32 output, residual add, fused statistics (0 / 1 / 2),
// deconv scatter.
__device__ __forceinline__ int tc_fast_div(int n, int d, double rcp) {          // n >= 0, d > 0; exact
  int q = __double2int_rz((double)n * rcp);
  const int r = n - q * d;
  if (r < 0) --q;
  else if (r >= d) ++q;
  return q;
}

// SPLIT (fp32 engine): the fp32 result is stored as the [hi | lo] f16 operand pair of the next convolution -- f16 output
// tensor of 2 * cout channels, hi = half(v) at channel co, lo = half(v - hi) at cout + co (include/cm2.h "Split precision") --
// so no fp32 copy is written and no separate split pass reads it back.  Two staged store rounds per 32-channel pass.
template <bool F32, bool RES, int STATS, bool DECONV, bool SPLIT = false>
__device__ __forceinline__ void tc_epilogue_rows_staged(const TcParams& p, const TileGeom& g, uint32_t taddr, int m, int n0,
                                                        uint32_t stage_smem, int lane, uint32_t ss_smem, int cset, int ncset) {
  constexpr int esize_res = F32 ? 4 : 2;             // the residual has the output's element type
  const int mr = m - g.row0;
  bool in_range = mr >= 0 && mr < g.rows, interior = false;
  int img = 0, y = 0, x = 0;
  if (in_range) {
    img = tc_fast_div(mr, g.plane, g.rcp_plane);
    int r = mr - img * g.plane;
    if (p.halo) {
      int yy = tc_fast_div(r, g.pitch, g.rcp_pitch), xx = r - yy * g.pitch;
      y = yy - 1; x = xx - 1;
      interior = y >= 0 && y < g.h && x >= 0 && x < g.w;
    } else {
      y = tc_fast_div(r, g.w, g.rcp_pitch); x = r - y * g.w;
      interior = true;
    }
  }
  const bool do_store = in_range && (interior || (p.out_halo && p.out_mode == 0)) && !(p.dbg & 1);
  long long out_off;
  if (p.num_seg)
    out_off = (long long)m * p.out_sw;                 // segmented output: same flat row, pitch = cout
  else if (DECONV)
    out_off = (long long)img * p.out_sn + (long long)(2 * y) * p.out_sh + (long long)(2 * x) * p.out_sw;
  else if (p.out_mode == 0)
    out_off = (long long)img * p.out_sn + (long long)y * p.out_sh + (long long)x * p.out_sw + g.out_extra;
  else
    out_off = (long long)((y & 1) * 2 + (x & 1)) * p.out_plane + (long long)img * p.out_sn +
              (long long)(y >> 1) * p.out_sh + (long long)(x >> 1) * p.out_sw;
  // residual row: bf16, or fp32 when the output is fp32 (split-precision convolutions)
  const char* res_row = nullptr;
  if (RES && interior)
    res_row = reinterpret_cast<const char*>(p.res) +
              ((long long)img * p.res_sn + (long long)(p.res_mode == 2 ? (y >> 1) : y) * p.res_sh +
               (long long)(p.res_mode == 2 ? (x >> 1) : x) * p.res_sw) * esize_res;
  const unsigned store_mask = __ballot_sync(0xffffffffu, do_store);      // rows of this warp that get written
  const unsigned int_mask = __ballot_sync(0xffffffffu, interior);
  const bool all_int = int_mask == 0xffffffffu;
  const unsigned stat_rows = STATS ? int_mask : 0u;
  const int img_g = img + g.img0;                    // image index across segments (statistics slot)
  constexpr int esize = F32 ? 4 : 2;
  constexpr int cpp = F32 ? 16 : 32;                 // channels per pass = 64 bytes per row
  const float relu_floor = p.relu ? 0.f : -INFINITY;
  uint64_t keep_policy = 0;
  if (STATS) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
  const uint32_t my_row = stage_smem + (uint32_t)lane * EPI_PITCH;
  // write-back role of this lane: rows it*8 + lane/4, 16-byte piece lane%4 -- row addresses are fixed for the tile
  char* wptr[4];
  bool wok[4];
  uint32_t wsm[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2);
    const long long roff = __shfl_sync(0xffffffffu, out_off, row);
    wptr[it] = reinterpret_cast<char*>(p.out) + roff * esize + 16 * (lane & 3);
    wok[it] = (store_mask >> row) & 1u;
    wsm[it] = stage_smem + (uint32_t)row * EPI_PITCH + 16u * (uint32_t)(lane & 3);
  }
  for (int c0 = cset * cpp; c0 < p.bn; c0 += ncset * cpp) {
    const int co0 = n0 + c0;
    float v[32];
    uint4 rres[4];
    const bool has_res = RES && res_row != nullptr && co0 < p.cout;
    if (RES && has_res) {                                     // issue the residual loads before waiting on TMEM
#pragma unroll
      for (int j = 0; j < 4; ++j) {                             // 16 bytes each: 8 bf16 or 4 fp32 channels
        constexpr int per = F32 ? 4 : 8;
        if (per * j < cpp && co0 + per * j < p.cout)
          rres[j] = __ldg(reinterpret_cast<const uint4*>(res_row + (size_t)(co0 + per * j) * esize_res));
        else rres[j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    {
      uint32_t raw[16];
      __syncwarp();
      tc_ld16(taddr + (uint32_t)c0, raw);
      if (!F32 && c0 + 16 < p.bn) {
        uint32_t raw2[16];
        tc_ld16(taddr + (uint32_t)(c0 + 16), raw2);
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 + j] = __uint_as_float(raw2[j]);
      } else {
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 + j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
    }
    if (co0 >= p.cout) continue;                     // warp-uniform
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      if (4 * j4 < cpp && c0 + 4 * j4 < p.bn) {
        const float4 sc = lds_f4(ss_smem + 4u * (uint32_t)(c0 + 4 * j4)), sh = lds_f4(ss_smem + 1024u + 4u * (uint32_t)(c0 + 4 * j4));
        ffma2(v[4 * j4 + 0], v[4 * j4 + 1], v[4 * j4 + 0], v[4 * j4 + 1], sc.x, sc.y, sh.x, sh.y);
        ffma2(v[4 * j4 + 2], v[4 * j4 + 3], v[4 * j4 + 2], v[4 * j4 + 3], sc.z, sc.w, sh.z, sh.w);
      }
    }
    if (RES && has_res) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {rres[j].x, rres[j].y, rres[j].z, rres[j].w};
        if (F32) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[4 * j + i] += __uint_as_float(w[i]);
        } else if (8 * j < cpp) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[8 * j + 2 * i] += __uint_as_float(w[i] << 16);
            v[8 * j + 2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
          }
        }
      }
    }
    // ---- stage this thread's 64 bytes
    if (F32) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], relu_floor);
      if (!all_int) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = interior ? v[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + 16u * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                     "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
    } else if (SPLIT) {
      uint32_t whi[16], wlo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a0 = fmaxf(v[2 * i], relu_floor), a1 = fmaxf(v[2 * i + 1], relu_floor);
        const __half2 h = __floats2half2_rn(a0, a1);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
        whi[i] = interior ? *reinterpret_cast<const uint32_t*>(&h) : 0u;
        wlo[i] = interior ? *reinterpret_cast<const uint32_t*>(&l) : 0u;
      }
      const int valid_pieces_s = min(4, (min(p.bn - c0, p.cout - co0) + 7) / 8);
      const bool piece_ok_s = (lane & 3) < valid_pieces_s;
#pragma unroll
      for (int part = 0; part < 2; ++part) {      // hi, then lo
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t* w = part ? wlo : whi;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + 16u * j), "r"(w[4 * j]), "r"(w[4 * j + 1]),
                       "r"(w[4 * j + 2]), "r"(w[4 * j + 3]) : "memory");
        }
        __syncwarp();
        const long long col_bytes_s = (long long)(co0 + (part ? p.cout : 0)) * 2;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          uint4 val;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                       : "r"(wsm[it]) : "memory");
          if (wok[it] && piece_ok_s) *reinterpret_cast<uint4*>(wptr[it] + col_bytes_s) = val;
        }
        __syncwarp();
      }
      continue;
    } else {
      const __nv_bfloat162 floor2 = __floats2bfloat162_rn(relu_floor, relu_floor);
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        h2 = __hmax2(h2, floor2);                    // max(round(x), 0) == round(max(x, 0))
        w[i] = *reinterpret_cast<uint32_t*>(&h2);
      }
      if (!all_int) {
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = interior ? w[i] : 0u;
      }
      if (STATS) {                                   // statistics of the values as stored (bf16-rounded)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          v[2 * i] = __uint_as_float(w[i] << 16);
          v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + 16u * j), "r"(w[4 * j]), "r"(w[4 * j + 1]),
                     "r"(w[4 * j + 2]), "r"(w[4 * j + 3]) : "memory");
    }
    if (STATS) tc_epilogue_stats<STATS>(p, v, stat_rows, interior, img_g, co0, min(cpp, min(p.bn - c0, p.cout - co0)), lane);
    __syncwarp();
    // ---- write back: 4 instructions x (8 rows x 64 bytes)
    long long extra = 0;
    int cc = co0;
    if (DECONV) {
      const int cq = p.cout >> 2;
      const int quad = co0 / cq;
      cc = co0 - quad * cq;
      extra = (long long)(quad >> 1) * p.out_sh + (long long)(quad & 1) * p.out_sw;
    }
    const long long col_bytes = (extra + cc) * esize;
    constexpr int elems16 = F32 ? 4 : 8;             // elements per 16-byte piece
    // 16-byte pieces of this pass that belong to this tile (bn need not be a multiple of the pass width) and to cout
    const int valid_pieces = min(4, (min(p.bn - c0, p.cout - co0) + elems16 - 1) / elems16);
    const bool piece_ok = (lane & 3) < valid_pieces;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      uint4 val;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                   : "r"(wsm[it]) : "memory");
      if (wok[it] && piece_ok) {
        if (STATS) {
          // outputs whose statistics are taken here are read next by a bandwidth-bound pass (GroupNorm apply / eSE apply) that
          // walks the tensor back to front: ask the L2 to keep these lines (evict_last) rather than the operand tiles streaming by
          asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(wptr[it] + col_bytes), "r"(val.x), "r"(val.y),
                       "r"(val.z), "r"(val.w), "l"(keep_policy) : "memory");
        } else {
          *reinterpret_cast<uint4*>(wptr[it] + col_bytes) = val;
        }
      }
    }
  }
  __syncwarp();
}

// out_mode 3 (sam.py:74-83, :96-97 + mask_head.py:196-216): the GEMM is the 2x2 transposed conv (column = quadrant *
// cq + channel, one N tile per quadrant); each thread owns an input pixel, applies bias + ReLU + bf16 rounding (the
// value the unfused pipeline stores), dots it with the predictor row of the ROI's class (staged in the "scale" slot)
// and writes sigmoid(dot + b[cls]) to probs[roi, 2y + dy, 2x + dx].  The [R, 28, 28, 256] tensor is never materialised.
__device__ __forceinline__ void tc_epilogue_deconv_predict(const TcParams& p, const TileGeom& g, uint32_t taddr, int m, int n0,
                                                           uint32_t ss_smem) {
  const int mr = m - g.row0;
  const bool in_range = mr >= 0 && mr < g.rows;
  int img = 0, y = 0, x = 0;
  bool interior = false;
  if (in_range) {
    img = tc_fast_div(mr, g.plane, g.rcp_plane);
    const int r = mr - img * g.plane;
    const int yy = tc_fast_div(r, g.pitch, g.rcp_pitch), xx = r - yy * g.pitch;
    y = yy - 1; x = xx - 1;
    interior = y >= 0 && y < g.h && x >= 0 && x < g.w;
  }
  float dot0 = 0.f, dot1 = 0.f;
  for (int c0 = 0; c0 < p.bn; c0 += 32) {
    float v[32];
    {
      uint32_t raw[16], raw2[16];
      __syncwarp();
      tc_ld16(taddr + (uint32_t)c0, raw);
      tc_ld16(taddr + (uint32_t)(c0 + 16), raw2);
      tc_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) { v[j] = __uint_as_float(raw[j]); v[16 + j] = __uint_as_float(raw2[j]); }
    }
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 wv = lds_f4(ss_smem + 4u * (uint32_t)(c0 + 4 * j4)), sh = lds_f4(ss_smem + 1024u + 4u * (uint32_t)(c0 + 4 * j4));
      // relu(acc + bias) rounded to bf16, as stored by the unfused deconv
      __nv_bfloat162 a = __floats2bfloat162_rn(fmaxf(v[4 * j4 + 0] + sh.x, 0.f), fmaxf(v[4 * j4 + 1] + sh.y, 0.f));
      __nv_bfloat162 b = __floats2bfloat162_rn(fmaxf(v[4 * j4 + 2] + sh.z, 0.f), fmaxf(v[4 * j4 + 3] + sh.w, 0.f));
      const uint32_t ua = *reinterpret_cast<uint32_t*>(&a), ub = *reinterpret_cast<uint32_t*>(&b);
      ffma2(dot0, dot1, __uint_as_float(ua << 16), __uint_as_float(ua & 0xffff0000u), wv.x, wv.y, dot0, dot1);
      ffma2(dot0, dot1, __uint_as_float(ub << 16), __uint_as_float(ub & 0xffff0000u), wv.z, wv.w, dot0, dot1);
    }
  }
  if (interior && !(p.dbg & 1)) {
    int cls = p.pred_ncls == 1 ? 0 : (int)p.pred_cls[img];
    cls = min(max(cls, 0), p.pred_ncls - 1);
    const int quad = n0 / (p.cout >> 2);
    const float logit = (dot0 + dot1) + __ldg(p.pred_b + cls);
    float* o = reinterpret_cast<float*>(p.out) + (long long)img * p.out_sn + (long long)(2 * y + (quad >> 1)) * p.out_sh +
               (long long)(2 * x + (quad & 1)) * p.out_sw;
    *o = 1.0f / (1.0f + expf(-logit));
  }
}

// kind: 0 plain bf16, 1 bf16 + channel sums, 2 bf16 + GroupNorm sums, 3 bf16 + residual, 4 f32, 5 bf16 deconv scatter,
//       6 fused deconv + predictor, 7 f32 + channel sums, 8 f32 + GroupNorm sums, 9 f32 + f32 residual, 10 f32 deconv scatter,
//       11 split f16 [hi | lo] output
__device__ __forceinline__ void tc_epilogue_dispatch(const TcParams& p, const TileGeom& g, uint32_t taddr, int m, int n0,
                                                     uint32_t stage_smem, int lane, uint32_t ss_smem, int cset, int ncset) {
  if (p.epi_kind == 6) {
    if (cset == 0) tc_epilogue_deconv_predict(p, g, taddr, m, n0, ss_smem);
    return;
  }
  switch (p.epi_kind) {
    case 0: tc_epilogue_rows_staged<false, false, 0, false>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    case 1: tc_epilogue_rows_staged<false, false, 1, false>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    case 2: tc_epilogue_rows_staged<false, false, 2, false>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    case 3: tc_epilogue_rows_staged<false, true, 0, false>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    case 4: tc_epilogue_rows_staged<true, false, 0, false>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    case 7: tc_epilogue_rows_staged<true, false, 1, false>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    case 8: tc_epilogue_rows_staged<true, false, 2, false>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    case 9: tc_epilogue_rows_staged<true, true, 0, false>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    case 10: tc_epilogue_rows_staged<true, false, 0, true>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    case 11: tc_epilogue_rows_staged<false, false, 0, false, true>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
    default: tc_epilogue_rows_staged<false, false, 0, true>(p, g, taddr, m, n0, stage_smem, lane, ss_smem, cset, ncset); break;
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.bn * TC_BK * 2;
  const uint32_t stage_bytes = TC_A_BYTES + b_bytes;
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;
  // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2]; then the TMEM base address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
  const int n_epi_warps = 4 * p.epi_sets;
  const uint32_t epi_base = (tmem_slot + 16u + 15u) & ~15u;       // n_epi_warps x EPI_WARP_BYTES of staging
  const uint32_t ss_base = epi_base + (uint32_t)n_epi_warps * EPI_WARP_BYTES;         // EPI_SS_BYTES of scale/shift

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mn_tiles = p.m_tiles * p.n_tiles;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int total_tiles = mn_tiles * ksplit;        // split-K: slice-major, so the slices of one output tile run side by side

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), (uint32_t)n_epi_warps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  // programmatic dependent launch: the setup above overlapped the previous kernel's tail; let the next kernel start its own
  // setup, then wait until everything this one reads (and overwrites) is final
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ks = t / mn_tiles, tt = t - ks * mn_tiles;
        const int m0 = p.row_begin + (tt / p.n_tiles) * TC_BM, n0 = (tt % p.n_tiles) * p.bn;
        const int kb_lo = ksplit > 1 ? ks * p.kb_per_split : 0, kb_hi = ksplit > 1 ? kb_lo + p.kb_per_split : 0x7fffffff;
        int kb = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int shift = tc_tap_shift(p, tap, p.num_seg ? tc_geom(p, m0).pitch : p.pitch);
          for (int s = 0; s < p.num_src; ++s) {
            const int nblk = (p.src_c[s] + TC_BK - 1) / TC_BK;
            for (int cb = 0; cb < nblk; ++cb, ++kb) {
              if (kb < kb_lo || kb >= kb_hi) continue;               // another K slice's block
              mbar_wait_ctl(p.spin, empty_bar(stage), phase ^ 1u);
              const uint32_t sa = base + (uint32_t)stage * stage_bytes;
              if (elect_one_sync()) {
                if (p.dbg & 2) {
                  mbar_arrive(full_bar(stage));
                } else {
                  mbar_expect_tx(full_bar(stage), TC_A_BYTES + b_bytes);
                  tma_load_2d(sa, &p.a_map[s], full_bar(stage), cb * TC_BK, m0 + shift);
                  tma_load_2d(sa + TC_A_BYTES, &p.b_map, full_bar(stage), kb * TC_BK, n0);
                }
              }
              __syncwarp();
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    {
      // instruction descriptor: D=f32, A=B=bf16 (or f16: p.idesc_ab), both K-major, N = bn, M = 128
      const uint32_t idesc = (1u << 4) | p.idesc_ab | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const int taps = p.taps, num_src = p.num_src, n_stages = p.stages, spin = p.spin;
      const bool no_mma = (p.dbg & 4) != 0;
      const uint64_t adesc_base = umma_desc_sw128(base), bdesc_base = umma_desc_sw128(base + TC_A_BYTES);
      const uint64_t desc_step = (uint64_t)(stage_bytes >> 4);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait_ctl(spin, tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_ACC_COLS);
        uint32_t accumulate = 0;
        const int ks = t / mn_tiles;
        const int kb_lo = ksplit > 1 ? ks * p.kb_per_split : 0, kb_hi = ksplit > 1 ? kb_lo + p.kb_per_split : 0x7fffffff;
        int kb = 0;
        for (int tap = 0; tap < taps; ++tap) {
          for (int s = 0; s < num_src; ++s) {
            const int c = p.src_c[s];
            const int nblk = (c + TC_BK - 1) / TC_BK;
            for (int cb = 0; cb < nblk; ++cb, ++kb) {
              if (kb < kb_lo || kb >= kb_hi) continue;
              const int nk = (min(TC_BK, c - cb * TC_BK) + 15) >> 4;        // 16-channel MMAs in this block
              mbar_wait_ctl(spin, full_bar(stage), phase);
              tc_fence_after();
              const uint64_t adesc = adesc_base + (uint64_t)stage * desc_step, bdesc = bdesc_base + (uint64_t)stage * desc_step;
              if (elect_one_sync()) {
                if (!no_mma) {
                  // +32 bytes per K=16 step inside the 128B swizzle row (start-address field is in 16B units)
                  if (nk == 4) {
                    tc_mma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
                    tc_mma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                    tc_mma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                    tc_mma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                  } else {
                    for (int k = 0; k < nk; ++k) tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accumulate | (uint32_t)k);
                  }
                }
                tc_commit(empty_bar(stage));          // frees the smem slot when these MMAs have read it
              }
              __syncwarp();
              accumulate = 1;
              if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
        if (elect_one_sync()) tc_commit(tfull_bar(acc));                  // accumulator complete -> epilogue
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    if (p.trim_base && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
      // output rows outside the trimmed tile range are halo pixels: keep them zero (was two memset nodes per launch)
      const long long t0 = ((long long)threadIdx.x - 64) * 16, step = 32ll * n_epi_warps * 16;
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      if (blockIdx.x == 0)
        for (long long i = t0; i < p.trim_lo_bytes; i += step) *reinterpret_cast<uint4*>(p.trim_base + i) = z;
      if (blockIdx.x == gridDim.x - 1)
        for (long long i = t0; i < p.trim_hi_bytes; i += step) *reinterpret_cast<uint4*>(p.trim_base + p.trim_hi_off + i) = z;
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t parity = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, parity ^= 1u) {
      const int ks = t / mn_tiles, tt = t - ks * mn_tiles;
      const int m0 = p.row_begin + (tt / p.n_tiles) * TC_BM, n0 = (tt % p.n_tiles) * p.bn;
      // scale / shift of the tile's columns: staged once when there is a single N tile, else per tile (double buffered)
      const uint32_t ss = ss_base + (p.n_tiles == 1 ? 0u : parity * 2048u);
      if (p.n_tiles > 1 || t == (int)blockIdx.x)
        tc_stage_scale_shift(p, ss, n0, (int)threadIdx.x - 64, 32 * n_epi_warps, m0);      // overlaps the MMAs of this tile
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * TC_ACC_COLS) + ((uint32_t)(q * 32) << 16);
      TileGeom g = tc_geom(p, m0);
      g.out_extra = (long long)ks * p.split_stride;
      const int cset = (warp - 2) >> 2;
      if (p.dbg & 8) {
      } else if (p.fast_store)
        tc_epilogue_dispatch(p, g, taddr, m0 + q * 32 + lane, n0, epi_base + (uint32_t)(warp - 2) * EPI_WARP_BYTES, lane, ss,
                             cset, p.epi_sets);
      else if (cset == 0)
        tc_epilogue_rows(p, g, taddr, m0 + q * 32 + lane, n0, ss);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// v2: 256 GEMM rows per tile (two 128-row accumulators that share every weight tile), separate smem
// rings for A slabs and W tiles, 8 epilogue warps.  The conv kernels are bound by L2 -> smem traffic
// (ncu: tensor pipe 30-55 %, L2 ~7 TB/s), so this variant cuts bytes per MAC:
//   * W tile traffic per row halves (one W box feeds 256 rows);
//   * kx-merge (stride-1 3x3): the three taps (ky, 0..2) of a k-block read the SAME rows shifted by
//     -1/0/+1, so ONE slab of 272 rows (two TMA boxes of 136) is loaded per (ky, source, k-block) and the
//     MMA A-descriptor simply starts 0, 1 or 2 rows (128 B each) into it -- A traffic drops 3x.
// Warps: 0 = TMA producer, 1 = TMEM alloc + MMA issue, 2..9 = epilogue (warps 2-5: rows 0-127, 6-9: 128-255).
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t umma_desc_sw128_at(uint32_t smem_addr, int desc_mode) {
  uint64_t d = umma_desc_sw128(smem_addr);
  if (desc_mode == 1) d |= (uint64_t)((smem_addr >> 7) & 7u) << 49;      // matrix base offset
  return d;
}

// PAIR = true: the kernel is launched in clusters of two CTAs (one TPC) and every MMA is a cta_group::2 instruction of
// M = 256: the pair works on 512 consecutive GEMM rows (256 per CTA, as above), each CTA loads its own A slab and HALF
// of every weight tile (N/2 rows), and the leader CTA's MMA warp issues for both.  Per MMA a CTA then reads 4 KB of A
// and N/2 x 32 B of W from its shared memory instead of 4 KB + N x 32 B -- for N = 128 that is 96 instead of 128 bytes
// per clock, the shared-memory bandwidth that bounds the single-CTA kernel on N <= 128 layers -- and the L2 -> shared
// weight traffic halves as well.  Barrier protocol: `full` barriers live in the leader (both CTAs' TMA loads credit
// their bytes there), `empty` / `tmem full` barriers are signalled in both CTAs by multicast commits, the epilogue
// warps of both CTAs arrive on the leader's `tmem empty` barrier.
template <int MAXT, bool PAIR>
__global__ void __launch_bounds__(MAXT, 1) conv_tc2_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_half_bytes = (uint32_t)p.a_box_rows * 128u;
  const uint32_t a_slab_bytes = 2u * a_half_bytes;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  constexpr int NCTA = PAIR ? 2 : 1;
  const uint32_t b_bytes = (uint32_t)(p.bn / NCTA) * TC_BK * 2;      // weight rows held by THIS CTA
  const uint32_t b_base = base + (uint32_t)p.sa_stages * a_slab_bytes;
  const uint32_t bar_base = b_base + (uint32_t)p.sb_stages * b_bytes;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (p.sa_stages + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * p.sa_stages + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (2 * p.sa_stages + p.sb_stages + s); };
  const uint32_t tbar = bar_base + 8u * (2 * p.sa_stages + 2 * p.sb_stages);
  auto tfull_bar = [&](int a) { return tbar + 8u * a; };
  auto tempty_bar = [&](int a) { return tbar + 8u * (2 + a); };
  const uint32_t tmem_slot = tbar + 8u * 4;
  const int n_epi_warps = 8 * p.epi_sets;
  const uint32_t epi_base = (tmem_slot + 16u + 15u) & ~15u;       // n_epi_warps x EPI_WARP_BYTES of staging
  const uint32_t ss_base = epi_base + (uint32_t)n_epi_warps * EPI_WARP_BYTES;         // EPI_SS_BYTES of scale/shift

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tiles are walked per CTA (256 rows) or per pair (512 rows, 256 per CTA)
  const int total_tiles = PAIR ? ((p.m_tiles + 1) >> 1) * p.n_tiles : p.m_tiles * p.n_tiles;
  const int tile_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tile_rows = PAIR ? 512 : 256, row_off = rank * 256;
  const int groups_per_tap_row = p.kx_merge ? 3 : 1;          // W tiles consumed per A slab
  const int ngroup_outer = p.kx_merge ? 3 : p.taps;            // ky (merged) or tap

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
    for (int s = 0; s < p.sa_stages; ++s) { mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), 1); }
    for (int s = 0; s < p.sb_stages; ++s) { mbar_init(bfull_bar(s), 1); mbar_init(bempty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), (uint32_t)(n_epi_warps * NCTA)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  // programmatic dependent launch: the setup above overlapped the previous kernel's tail; let the next kernel start its own
  // setup, then wait until everything this one reads (and overwrites) is final
  pdl_launch_dependents();
  pdl_wait();
  const int half_cols = p.acc_stages == 2 ? 128 : 256;        // TMEM columns per 128-row accumulator

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      if (p.b_resident) {
        // weights of small-N layers (stem, bbox/centerness head) are loaded once per CTA: slot i = (tap, k-block) i
        for (int i = 0; i < p.sb_stages; ++i) {
          if (elect_one_sync()) {
            if (p.dbg & 2) {
              mbar_arrive(bfull_bar(i));
            } else {
              mbar_expect_tx(bfull_bar(i), b_bytes);
              tma_load_2d(b_base + (uint32_t)i * b_bytes, &p.b_map, bfull_bar(i), i * TC_BK, 0);
            }
          }
          __syncwarp();
        }
      }
      for (int t = tile_first; t < total_tiles; t += tile_step) {
        const int m0 = (t / p.n_tiles) * tile_rows + row_off, n0 = (t % p.n_tiles) * p.bn;
        for (int g = 0; g < ngroup_outer; ++g) {
          // first row of the slab in the flat source matrix
          const int pitch = p.num_seg ? tc_geom(p, m0).pitch : p.pitch;
          const int row0 = m0 + (p.kx_merge ? tc_tap_shift(p, g * 3 + 1, pitch) - 1 : tc_tap_shift(p, g, pitch));
          int blk = 0;
          for (int s = 0; s < p.num_src; ++s) {
            const int nblk = (p.src_c[s] + TC_BK - 1) / TC_BK;
            for (int cb = 0; cb < nblk; ++cb, ++blk) {
              mbar_wait_ctl(p.spin, aempty_bar(sa), pa ^ 1u);
              const uint32_t slab = base + (uint32_t)sa * a_slab_bytes;
              if (elect_one_sync()) {
                if (p.dbg & 2) {
                  if (rank == 0) mbar_arrive(afull_bar(sa));
                } else {
                  if (rank == 0) mbar_expect_tx(afull_bar(sa), a_slab_bytes * NCTA);
                  tma_load_2d_x<PAIR>(slab, &p.a_map[s], afull_bar(sa), cb * TC_BK, row0);
                  tma_load_2d_x<PAIR>(slab + a_half_bytes, &p.a_map[s], afull_bar(sa), cb * TC_BK, row0 + p.a_box_rows);
                }
              }
              __syncwarp();
              if (++sa == p.sa_stages) { sa = 0; pa ^= 1u; }
              for (int j = 0; j < groups_per_tap_row && !p.b_resident; ++j) {
                const int tap = p.kx_merge ? g * 3 + j : g;
                mbar_wait_ctl(p.spin, bempty_bar(sb), pb ^ 1u);
                if (elect_one_sync()) {
                  if (p.dbg & 2) {
                    if (rank == 0) mbar_arrive(bfull_bar(sb));
                  } else {
                    if (rank == 0) mbar_expect_tx(bfull_bar(sb), b_bytes * NCTA);
                    tma_load_2d_x<PAIR>(b_base + (uint32_t)sb * b_bytes, &p.b_map, bfull_bar(sb), (tap * p.nblk_total + blk) * TC_BK,
                                        n0 + rank * (p.bn / NCTA));
                  }
                }
                __syncwarp();
                if (++sb == p.sb_stages) { sb = 0; pb ^= 1u; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (rank == 0) {
      // The issue loop is the critical path of short-N layers (ncu: the producer waits for THIS warp, which never waits
      // itself): every parameter it needs is copied to a local first, descriptors are advanced incrementally, a full
      // 64-channel block takes the branch-free 8-MMA path, and a whole slab is issued under a single elect.
      const uint32_t idesc = (1u << 4) | p.idesc_ab | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)((TC_BM * NCTA) >> 4) << 24);
      const int kx_merge = p.kx_merge, b_res = p.b_resident, nblk_total = p.nblk_total, n_acc = p.acc_stages;
      const int n_sa = p.sa_stages, n_sb = p.sb_stages, num_src = p.num_src, spin = p.spin;
      const bool no_mma = (p.dbg & 4) != 0;
      const uint64_t bdesc_step = (uint64_t)(b_bytes >> 4);          // descriptor start-address units (16 B)
      const uint64_t bdesc_base = umma_desc_sw128(b_base);
      const uint64_t adesc_half = (uint64_t)((128u * 128u) >> 4);    // rows 128.. of the slab feed the second accumulator
      int sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, acc_phase = 0;
      if (b_res) {
        for (int i = 0; i < n_sb; ++i) mbar_wait_ctl(spin, bfull_bar(i), 0u);
        tc_fence_after();
      }
      for (int t = tile_first; t < total_tiles; t += tile_step) {
        mbar_wait_ctl(spin, tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * 2 * half_cols);
        const uint32_t d1 = d0 + (uint32_t)half_cols;
        uint32_t accumulate = 0;
        for (int g = 0; g < ngroup_outer; ++g) {
          int blk = 0;
          for (int s = 0; s < num_src; ++s) {
            const int c = p.src_c[s];
            const int nblk = (c + TC_BK - 1) / TC_BK;
            for (int cb = 0; cb < nblk; ++cb, ++blk) {
              const int nk = (min(TC_BK, c - cb * TC_BK) + 15) >> 4;
              mbar_wait_ctl(spin, afull_bar(sa), pa);
              tc_fence_after();
              const uint64_t adesc_slab = umma_desc_sw128(base + (uint32_t)sa * a_slab_bytes);
              if (b_res) {
                // weights resident: tile (tap, k-block) lives in slot tap * nblk_total + blk; one elect per slab
                if (elect_one_sync()) {
                  for (int j = 0; j < groups_per_tap_row; ++j) {
                    const int tap = kx_merge ? g * 3 + j : g;
                    const uint64_t bdesc = bdesc_base + (uint64_t)(tap * nblk_total + blk) * bdesc_step;
                    // tap kx = j reads the slab j rows (128 B = 8 descriptor units each) further down
                    const uint64_t adesc0 = adesc_slab + (uint64_t)(kx_merge ? 8 * j : 0), adesc1 = adesc0 + adesc_half;
                    if (!no_mma) {
                      if (nk == 4) {
                        tc_mma_x<PAIR>(d0, adesc0, bdesc, idesc, accumulate);
                        tc_mma_x<PAIR>(d0, adesc0 + 2, bdesc + 2, idesc, 1u);
                        tc_mma_x<PAIR>(d0, adesc0 + 4, bdesc + 4, idesc, 1u);
                        tc_mma_x<PAIR>(d0, adesc0 + 6, bdesc + 6, idesc, 1u);
                        tc_mma_x<PAIR>(d1, adesc1, bdesc, idesc, accumulate);
                        tc_mma_x<PAIR>(d1, adesc1 + 2, bdesc + 2, idesc, 1u);
                        tc_mma_x<PAIR>(d1, adesc1 + 4, bdesc + 4, idesc, 1u);
                        tc_mma_x<PAIR>(d1, adesc1 + 6, bdesc + 6, idesc, 1u);
                      } else {
                        for (int k = 0; k < nk; ++k) tc_mma_x<PAIR>(d0, adesc0 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accumulate | (uint32_t)k);
                        for (int k = 0; k < nk; ++k) tc_mma_x<PAIR>(d1, adesc1 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accumulate | (uint32_t)k);
                      }
                    }
                    accumulate = 1;
                  }
                  tc_commit_x<PAIR>(aempty_bar(sa));
                }
                __syncwarp();
                accumulate = 1;
              } else {
                // weights streamed through their own ring (measured: letting the elected thread wait for the weight tiles
                // itself, to issue a whole slab from one elect region, is 5-10 % slower than these warp-level waits)
                for (int j = 0; j < groups_per_tap_row; ++j) {
                  mbar_wait_ctl(spin, bfull_bar(sb), pb);
                  tc_fence_after();
                  const uint64_t bdesc = bdesc_base + (uint64_t)sb * bdesc_step;
                  const uint64_t adesc0 = adesc_slab + (uint64_t)(kx_merge ? 8 * j : 0), adesc1 = adesc0 + adesc_half;
                  if (elect_one_sync()) {
                    if (!no_mma) {
                      if (nk == 4) {
                        tc_mma_x<PAIR>(d0, adesc0, bdesc, idesc, accumulate);
                        tc_mma_x<PAIR>(d0, adesc0 + 2, bdesc + 2, idesc, 1u);
                        tc_mma_x<PAIR>(d0, adesc0 + 4, bdesc + 4, idesc, 1u);
                        tc_mma_x<PAIR>(d0, adesc0 + 6, bdesc + 6, idesc, 1u);
                        tc_mma_x<PAIR>(d1, adesc1, bdesc, idesc, accumulate);
                        tc_mma_x<PAIR>(d1, adesc1 + 2, bdesc + 2, idesc, 1u);
                        tc_mma_x<PAIR>(d1, adesc1 + 4, bdesc + 4, idesc, 1u);
                        tc_mma_x<PAIR>(d1, adesc1 + 6, bdesc + 6, idesc, 1u);
                      } else {
                        for (int k = 0; k < nk; ++k) tc_mma_x<PAIR>(d0, adesc0 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accumulate | (uint32_t)k);
                        for (int k = 0; k < nk; ++k) tc_mma_x<PAIR>(d1, adesc1 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accumulate | (uint32_t)k);
                      }
                    }
                    tc_commit_x<PAIR>(bempty_bar(sb));
                    if (j + 1 == groups_per_tap_row) tc_commit_x<PAIR>(aempty_bar(sa));
                  }
                  __syncwarp();
                  accumulate = 1;
                  if (++sb == n_sb) { sb = 0; pb ^= 1u; }
                }
              }
              if (++sa == n_sa) { sa = 0; pa ^= 1u; }
            }
          }
        }
        if (elect_one_sync()) tc_commit_x<PAIR>(tfull_bar(acc));
        __syncwarp();
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;            // warps 2-5 / 10-13: rows 0-127, warps 6-9 / 14-17: rows 128-255
    const int cset = (warp - 2) >> 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t parity = 0;
    for (int t = tile_first; t < total_tiles; t += tile_step, parity ^= 1u) {
      const int m0 = (t / p.n_tiles) * tile_rows + row_off, n0 = (t % p.n_tiles) * p.bn;
      const uint32_t ss = ss_base + (p.n_tiles == 1 ? 0u : parity * 2048u);
      if (p.n_tiles > 1 || t == tile_first)
        tc_stage_scale_shift(p, ss, n0, (int)threadIdx.x - 64, 32 * n_epi_warps, m0);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * 2 * half_cols + half * half_cols) + ((uint32_t)(q * 32) << 16);
      const TileGeom tg = tc_geom(p, m0);
      if (p.dbg & 8) {
      } else if (p.fast_store)
        tc_epilogue_dispatch(p, tg, taddr, m0 + half * 128 + q * 32 + lane, n0, epi_base + (uint32_t)(warp - 2) * EPI_WARP_BYTES, lane, ss,
                             cset, p.epi_sets);
      else if (cset == 0)
        tc_epilogue_rows(p, tg, taddr, m0 + half * 128 + q * 32 + lane, n0, ss);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();    // the peer may still read this CTA's shared memory / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------
// v3: split-precision convolution (include/cm2.h "Split precision"): fp32-grade results from f16 tensor-core MMAs.
//
//   x = x_hi + x_lo, W = W_hi + W_lo (f16 pairs, 22 significant bits each):  x*W ~= x_hi*W_hi + x_lo*W_hi + x_hi*W_lo
//
// Measured on B200 (tests/test_gpu_conv_split.py): tcgen05.mma adds into its fp32 TMEM accumulator with truncation -- a
// same-sign sum of 2304 terms accumulated in ONE accumulator came out 1.3e-5 low (3e-8 per MMA), ten times the error of
// the CUDA-core fp32 engine and systematic.  So (Ootomo & Yokota's scheme, re-cut for TMEM):
//   * the dominant x_hi*W_hi products go to a MAIN accumulator that is drained every `chunk` K-blocks (<= 4 * chunk
//     accumulations) into fp32 registers of the epilogue warps, where the partial sums are added with round-to-nearest;
//     two main accumulators alternate so that the drain of one overlaps the MMAs into the other;
//   * the two cross terms are 2^-11 of the main term, so is their truncation error: they accumulate over the whole
//     K loop in a CROSS accumulator that is added once per tile (two of them alternate between tiles).
// One pipeline stage holds A_hi, A_lo (two boxes of the source's [hi | lo] matrix), W_hi and W_lo of one K-block: x_hi
// and W_hi are loaded once and used by two MMA groups.  128-row tiles, N <= 128: 4 x 128 TMEM columns.
// At the end of a tile the fp32 sums are written back to the cross accumulator's columns (tcgen05.st) and the ordinary
// epilogue (scale / shift / residual / statistics / stores) runs on them unchanged.
// Warps: 0 = TMA producer, 1 = TMEM alloc + MMA issue, 2..9 = epilogue (lane quarter q = warp & 3, column set = (warp - 2) >> 2).
// ------------------------------------------------------------------------------------------------
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) conv_tc3_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // A ring: slots of {x_hi slab, x_lo slab} (a_box_rows rows of 128 B each); W ring: slots of {W_hi, W_lo} (bn rows each).
  // kx-merge (stride-1 3x3): one slab of 136 rows per (ky, source, k-block) serves the three kx taps -- the MMA A
  // descriptor starts 0 / 1 / 2 rows into it, as in conv_tc2_kernel -- so x traffic drops 3x; this kernel is bound by
  // L2 -> shared-memory bytes (64 KB per 12 MMAs unmerged), so that is where its time goes.
  const uint32_t a_half = (uint32_t)p.a_box_rows * 128u, a_slot = 2u * a_half;
  const uint32_t b_bytes = (uint32_t)p.bn * TC_BK * 2, w_slot = 2u * b_bytes;
  const uint32_t w_base = base + (uint32_t)p.sa_stages * a_slot;
  const uint32_t bar_base = w_base + (uint32_t)p.sb_stages * w_slot;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (p.sa_stages + s); };
  auto wfull_bar = [&](int s) { return bar_base + 8u * (2 * p.sa_stages + s); };
  auto wempty_bar = [&](int s) { return bar_base + 8u * (2 * p.sa_stages + p.sb_stages + s); };
  const uint32_t tb = bar_base + 8u * (2 * p.sa_stages + 2 * p.sb_stages);
  auto mfull_bar = [&](int a) { return tb + 8u * a; };
  auto mempty_bar = [&](int a) { return tb + 8u * (2 + a); };
  auto xfull_bar = [&](int a) { return tb + 8u * (4 + a); };
  auto xempty_bar = [&](int a) { return tb + 8u * (6 + a); };
  const uint32_t tmem_slot = tb + 8u * 8;
  constexpr int n_epi_warps = 8;
  const uint32_t epi_base = (tmem_slot + 16u + 15u) & ~15u;
  const uint32_t ss_base = epi_base + (uint32_t)n_epi_warps * EPI_WARP_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int nkb = p.taps * p.nblk_total;                  // K-blocks per tile
  const int nchunks = (nkb + p.chunk - 1) / p.chunk;
  const int per_slab = p.kx_merge ? 3 : 1;                // W tiles consumed per A slab
  const int ngroup_outer = p.kx_merge ? 3 : p.taps;       // ky (merged) or tap

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.a_map[s]);
    tma_prefetch_desc(&p.b_map);
    for (int s = 0; s < p.sa_stages; ++s) { mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), 1); }
    for (int s = 0; s < p.sb_stages; ++s) { mbar_init(wfull_bar(s), 1); mbar_init(wempty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(mfull_bar(a), 1); mbar_init(mempty_bar(a), (uint32_t)n_epi_warps);
      mbar_init(xfull_bar(a), 1); mbar_init(xempty_bar(a), (uint32_t)n_epi_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  // programmatic dependent launch: the setup above overlapped the previous kernel's tail; let the next kernel start its own
  // setup, then wait until everything this one reads (and overwrites) is final
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m0 = p.row_begin + (t / p.n_tiles) * TC_BM, n0 = (t % p.n_tiles) * p.bn;
      const int pitch = p.num_seg ? tc_geom(p, m0).pitch : p.pitch;
      for (int g = 0; g < ngroup_outer; ++g) {
        const int row0 = m0 + (p.kx_merge ? tc_tap_shift(p, g * 3 + 1, pitch) - 1 : tc_tap_shift(p, g, pitch));
        int blk = 0;
        for (int s = 0; s < p.num_src; ++s) {
          const int c = p.src_c[s];
          const int nblk = (c + TC_BK - 1) / TC_BK;
          for (int cb = 0; cb < nblk; ++cb, ++blk) {
            mbar_wait_ctl(p.spin, aempty_bar(sa), pa ^ 1u);
            const uint32_t slab = base + (uint32_t)sa * a_slot;
            if (elect_one_sync()) {
              mbar_expect_tx(afull_bar(sa), a_slot);
              tma_load_2d(slab, &p.a_map[s], afull_bar(sa), cb * TC_BK, row0);                    // x_hi
              tma_load_2d(slab + a_half, &p.a_map[s], afull_bar(sa), c + cb * TC_BK, row0);       // x_lo
            }
            __syncwarp();
            if (++sa == p.sa_stages) { sa = 0; pa ^= 1u; }
            for (int j = 0; j < per_slab; ++j) {
              const int tap = p.kx_merge ? g * 3 + j : g;
              const int kcol = (tap * p.nblk_total + blk) * TC_BK;
              mbar_wait_ctl(p.spin, wempty_bar(sb), pb ^ 1u);
              const uint32_t wt = w_base + (uint32_t)sb * w_slot;
              if (elect_one_sync()) {
                mbar_expect_tx(wfull_bar(sb), w_slot);
                tma_load_2d(wt, &p.b_map, wfull_bar(sb), kcol, n0);                               // W_hi
                tma_load_2d(wt + b_bytes, &p.b_map, wfull_bar(sb), kcol, p.cout_pad + n0);        // W_lo
              }
              __syncwarp();
              if (++sb == p.sb_stages) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    const uint32_t idesc = (1u << 4) | p.idesc_ab | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const int num_src = p.num_src, n_sa = p.sa_stages, n_sb = p.sb_stages, spin = p.spin, chunk = p.chunk, kx_merge = p.kx_merge;
    const uint64_t a_step = (uint64_t)(a_slot >> 4), w_step = (uint64_t)(w_slot >> 4);
    const uint64_t ahi_base = umma_desc_sw128(base), alo_base = umma_desc_sw128(base + a_half);
    const uint64_t whi_base = umma_desc_sw128(w_base), wlo_base = umma_desc_sw128(w_base + b_bytes);
    int sa = 0, sb = 0, mbuf = 0, xbuf = 0;
    uint32_t pa = 0, pb = 0, mphase = 0, xphase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait_ctl(spin, xempty_bar(xbuf), xphase ^ 1u);
      tc_fence_after();
      const uint32_t dx = tmem_base + 256u + (uint32_t)(xbuf * 128);
      uint32_t acc_m = 0, acc_x = 0;
      int in_chunk = 0, kb = 0;
      for (int g = 0; g < ngroup_outer; ++g) {
        for (int s = 0; s < num_src; ++s) {
          const int c = p.src_c[s];
          const int nblk = (c + TC_BK - 1) / TC_BK;
          for (int cb = 0; cb < nblk; ++cb) {
            const int nk = (min(TC_BK, c - cb * TC_BK) + 15) >> 4;
            mbar_wait_ctl(spin, afull_bar(sa), pa);
            tc_fence_after();
            const uint64_t aoff = (uint64_t)sa * a_step;
            for (int j = 0; j < per_slab; ++j, ++kb) {
              if (in_chunk == 0) {                            // a fresh main accumulator
                mbar_wait_ctl(spin, mempty_bar(mbuf), mphase ^ 1u);
                tc_fence_after();
                acc_m = 0;
              }
              const uint32_t dm = tmem_base + (uint32_t)(mbuf * 128);
              mbar_wait_ctl(spin, wfull_bar(sb), pb);
              tc_fence_after();
              // tap kx = j reads the slab j rows (128 B = 8 descriptor units each) further down
              const uint64_t ashift = aoff + (uint64_t)(kx_merge ? 8 * j : 0);
              const uint64_t ahi = ahi_base + ashift, alo = alo_base + ashift;
              const uint64_t woff = (uint64_t)sb * w_step;
              const uint64_t whi = whi_base + woff, wlo = wlo_base + woff;
              const bool close_chunk = (in_chunk + 1 == chunk) || (kb + 1 == nkb);
              if (elect_one_sync()) {
                if (nk == 4) {                                // full 64-channel block: branch-free issue
                  tc_mma_bf16(dm, ahi, whi, idesc, acc_m);
                  tc_mma_bf16(dm, ahi + 2, whi + 2, idesc, 1u);
                  tc_mma_bf16(dm, ahi + 4, whi + 4, idesc, 1u);
                  tc_mma_bf16(dm, ahi + 6, whi + 6, idesc, 1u);
                  tc_mma_bf16(dx, alo, whi, idesc, acc_x);
                  tc_mma_bf16(dx, alo + 2, whi + 2, idesc, 1u);
                  tc_mma_bf16(dx, alo + 4, whi + 4, idesc, 1u);
                  tc_mma_bf16(dx, alo + 6, whi + 6, idesc, 1u);
                  tc_mma_bf16(dx, ahi, wlo, idesc, 1u);
                  tc_mma_bf16(dx, ahi + 2, wlo + 2, idesc, 1u);
                  tc_mma_bf16(dx, ahi + 4, wlo + 4, idesc, 1u);
                  tc_mma_bf16(dx, ahi + 6, wlo + 6, idesc, 1u);
                } else {
                  for (int k = 0; k < nk; ++k) tc_mma_bf16(dm, ahi + (uint64_t)(2 * k), whi + (uint64_t)(2 * k), idesc, acc_m | (uint32_t)k);
                  for (int k = 0; k < nk; ++k) tc_mma_bf16(dx, alo + (uint64_t)(2 * k), whi + (uint64_t)(2 * k), idesc, acc_x | (uint32_t)k);
                  for (int k = 0; k < nk; ++k) tc_mma_bf16(dx, ahi + (uint64_t)(2 * k), wlo + (uint64_t)(2 * k), idesc, 1u);
                }
                tc_commit(wempty_bar(sb));
                if (j + 1 == per_slab) tc_commit(aempty_bar(sa));
                if (close_chunk) tc_commit(mfull_bar(mbuf));
              }
              __syncwarp();
              acc_m = 1; acc_x = 1;
              if (++sb == n_sb) { sb = 0; pb ^= 1u; }
              if (close_chunk) {
                in_chunk = 0;
                if (++mbuf == 2) { mbuf = 0; mphase ^= 1u; }
              } else {
                ++in_chunk;
              }
            }
            if (++sa == n_sa) { sa = 0; pa ^= 1u; }
          }
        }
      }
      if (elect_one_sync()) tc_commit(xfull_bar(xbuf));
      __syncwarp();
      if (++xbuf == 2) { xbuf = 0; xphase ^= 1u; }
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3, cset = (warp - 2) >> 2;
    int mbuf = 0, xbuf = 0;
    uint32_t mphase = 0, xphase = 0, parity = 0;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, parity ^= 1u) {
      const int m0 = p.row_begin + (t / p.n_tiles) * TC_BM, n0 = (t % p.n_tiles) * p.bn;
      const uint32_t ss = ss_base + (p.n_tiles == 1 ? 0u : parity * 2048u);
      if (p.n_tiles > 1 || t == (int)blockIdx.x)
        tc_stage_scale_shift(p, ss, n0, (int)threadIdx.x - 64, 32 * n_epi_warps, m0);
      // this warp sums four 16-column groups of its 32 rows -- the groups its own staged epilogue passes read back (16
      // columns per pass for an fp32 output: g = cset, cset + 2, ..; 32 per pass for the split f16 output: g = 2 cset,
      // 2 cset + 1, 2 cset + 4, ..), so it later reads only what it wrote itself
      const int so = p.split_out;
      auto group_col = [&](int j) { return 16 * (so ? ((j >> 1) * 4 + cset * 2 + (j & 1)) : (cset + 2 * j)); };
      float run[4][16];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 16; ++i) run[j][i] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(mfull_bar(mbuf), mphase);
        tc_fence_after();
        const uint32_t ta = tmem_base + (uint32_t)(mbuf * 128) + lane_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c0 = group_col(j);
          if (c0 < p.bn) {                                  // warp-uniform
            uint32_t raw[16];
            __syncwarp();
            tc_ld16(ta + (uint32_t)c0, raw);
            tc_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) run[j][i] += __uint_as_float(raw[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(mempty_bar(mbuf));
        if (++mbuf == 2) { mbuf = 0; mphase ^= 1u; }
      }
      mbar_wait(xfull_bar(xbuf), xphase);
      tc_fence_after();
      const uint32_t tx = tmem_base + 256u + (uint32_t)(xbuf * 128) + lane_off;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c0 = group_col(j);
        if (c0 < p.bn) {
          uint32_t raw[16];
          __syncwarp();
          tc_ld16(tx + (uint32_t)c0, raw);
          tc_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) run[j][i] += __uint_as_float(raw[i]);
          tc_st16(tx + (uint32_t)c0, run[j]);
        }
      }
      tc_st_wait();
      const TileGeom g = tc_geom(p, m0);
      if (p.fast_store) {
        tc_epilogue_dispatch(p, g, tx, m0 + q * 32 + lane, n0, epi_base + (uint32_t)(warp - 2) * EPI_WARP_BYTES, lane, ss, cset, 2);
      } else {
        // the row-per-thread epilogue reads all columns from the cset-0 warps: make the other set's sums visible first
        tc_fence_before();
        asm volatile("bar.sync 2, %0;" ::"r"(32 * n_epi_warps) : "memory");
        tc_fence_after();
        if (cset == 0) tc_epilogue_rows(p, g, tx, m0 + q * 32 + lane, n0, ss);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(xempty_bar(xbuf));
      if (++xbuf == 2) { xbuf = 0; xphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    else
      cudaGetLastError();
  }
  return fn;
}

// 16-bit matrix [rows][cols] (cols contiguous, row pitch `pitch` >= cols elements: a channel slice of a wider tensor
// is a matrix too); box = [box_rows][64 cols], 128B swizzle; columns beyond `cols` read as zero
static bool encode_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows, bool f16) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Split-K finish: out[n, y, x, c] = act(scale[c] * sum_s part[s][n, y, x, c] + shift[c]) over the interior pixels of the output
// view; the partial buffers have the output's strides (fp32).  Fixed summation order: results do not depend on scheduling.
template <typename OutT>
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* __restrict__ part, long long split_stride, int ksplit,
                                                            View<OutT> out, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, int relu) {
  const int c8 = out.c >> 3;
  const long long total = (long long)out.n * out.h * out.w * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % c8);
    long long t = i / c8;
    const int x = (int)(t % out.w); t /= out.w;
    const int y = (int)(t % out.h);
    const int b = (int)(t / out.h);
    const long long off = (long long)b * out.sn + (long long)y * out.sh + (long long)x * out.sw + cv * 8;
    float acc[8];
    Vec8<float>::load(part + off, acc);
    for (int s = 1; s < ksplit; ++s) {
      float v[8];
      Vec8<float>::load(part + (long long)s * split_stride + off, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float sc = scale ? __ldg(scale + cv * 8 + k) : 1.f, sh = shift ? __ldg(shift + cv * 8 + k) : 0.f;
      acc[k] = fmaf(acc[k], sc, sh);
      if (relu) acc[k] = fmaxf(acc[k], 0.f);
    }
    Vec8<OutT>::store(out.p + off, acc);
  }
}

// Halo kinds of an interior view: 1 = every image carries its own one-pixel zero frame ([n, h+2, w+2, c] buffer); 2 = SHARED
// halo: line pitch w + 1 and image pitch (h + 1)(w + 1) pixels -- the zero pixel right of a line is the zero pixel left of the
// next line, the zero line under an image the zero line above the next image (the buffer ends with one more zero line + pixel;
// reads past it are TMA out-of-bounds zero fill).  A 3x3 tap is the same flat row shift (ky - 1) * pitch + (kx - 1) in both; for
// the 14x14 ROI maps the shared form has 225 GEMM rows per ROI instead of 256 (196 of them interior).
static int halo_kind(const cm2_act& a, bool slice_ok = false) {
  if (!(slice_ok ? a.sw >= a.c : a.sw == a.c)) return 0;
  if (a.sh == (long long)(a.w + 2) * a.sw && a.sn == (long long)(a.h + 2) * a.sh) return 1;
  if (a.sh == (long long)(a.w + 1) * a.sw && a.sn == (long long)(a.h + 1) * a.sh) return 2;
  return 0;
}
static bool is_dense_view(const cm2_act& a, bool slice_ok = false) {
  return (slice_ok ? a.sw >= a.c : a.sw == a.c) && a.sh == (long long)a.w * a.sw && a.sn == (long long)a.h * a.sh;
}

static int device_is_sm100() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    cached = major == 10 ? 1 : 0;
  }
  return cached;
}

static thread_local bool g_plan_splitk = false;         // set by conv_tc_launch_splitk around its tc_plan call

static int pick_bn(int cout_pad, int m_tiles, int sms) {
  int bn = cout_pad;
  if (bn > 256) {
    bn = 256;
    while (cout_pad % bn) bn -= 16;
  }
  // small problems: narrower tiles give more CTAs (weights are streamed once per N tile either way)
  while (bn >= 64 && bn % 32 == 0 && m_tiles * (cout_pad / bn) < sms) bn >>= 1;
  return bn;
}

// Validates the descriptor for the TC engine; on success fills `p` (without tensor maps when maps == false).
static int tc_plan(const cm2_conv_desc* d, TcParams* p, bool maps) {
  const cm2_act& s0 = d->src[0];
  if (!device_is_sm100()) { set_error("conv_tc: device is not sm_100"); return CM2_ERR_UNSUPPORTED; }
#define TC_REQUIRE(cond, ...) do { if (!(cond)) { set_error(__VA_ARGS__); return CM2_ERR_UNSUPPORTED; } } while (0)
  const bool f16 = d->dtype == CM2_F16;
  TC_REQUIRE((d->dtype == CM2_BF16 && (d->out_dtype == CM2_BF16 || d->out_dtype == CM2_F32)) ||
                 (f16 && (d->out_dtype == CM2_F32 || d->out_dtype == CM2_F16)),
             "conv_tc: needs bf16 sources (bf16 / f32 output) or f16 split sources (f32 or split f16 output)");
  const bool split_out = f16 && d->out_dtype == CM2_F16;
  if (split_out)
    TC_REQUIRE(!d->residual.data && !d->stats && (d->out_mode == 0 || d->out_mode == 2) && d->cout % 16 == 0,
               "conv_tc: the split f16 output takes out_mode 0 / 2 without residual or statistics, cout %% 16 == 0");
  TC_REQUIRE(!(f16 && d->out_mode == 3), "conv_tc: the fused deconv + predictor epilogue is bf16 only");
  const bool phase = d->src_phase != 0;
  if (phase)
    TC_REQUIRE(d->stride == 2 && d->kh == 3 && d->kw == 3 && d->pad == 1, "conv_tc: phase-split sources need a 3x3/s2/p1 conv");
  else
    TC_REQUIRE(d->stride == 1 && d->kh == d->kw && ((d->kh == 1 && d->pad == 0) || (d->kh == 3 && d->pad == 1)),
               "conv_tc: only stride-1 1x1/p0 and 3x3/p1, or 3x3/s2/p1 on phase-split sources (got k%d s%d p%d)", d->kh,
               d->stride, d->pad);
  TC_REQUIRE(!d->in_relu, "conv_tc: in_relu not supported");
  TC_REQUIRE(!d->stats || d->stats_mode == 1 || d->stats_mode == 2, "conv_tc: stats_mode must be 1 or 2");
  const bool seg = d->num_seg > 0;
  if (seg)
    TC_REQUIRE(d->num_seg <= CM2_MAX_SEG && !phase && d->out_mode == 0 && !d->residual.data,
               "conv_tc: segmented tensors need stride 1, out_mode 0, no residual");
  const int hk = seg ? 1 : halo_kind(s0, true);
  const bool halo = hk != 0;
  TC_REQUIRE(halo || (d->kh == 1 && is_dense_view(s0, true)), "conv_tc: source 0 is neither a halo-1 view nor (for 1x1) dense");
  TC_REQUIRE(hk != 2 || (!phase && d->out_mode != 3), "conv_tc: shared-halo views take stride-1 convolutions without the fused predictor");
  for (int i = 0; i < d->num_src; ++i) {
    const cm2_act& s = d->src[i];
    TC_REQUIRE(s.c % 16 == 0 && (reinterpret_cast<uintptr_t>(s.data) & 15) == 0 && s.sw % 8 == 0 && s.sw >= s.c,
               "conv_tc: source %d channels %d / pitch %lld / alignment", i, s.c, (long long)s.sw);
    if (!seg) TC_REQUIRE(halo ? halo_kind(s, true) == hk : is_dense_view(s, true), "conv_tc: source %d geometry differs from source 0", i);
  }
  memset(p, 0, sizeof(*p));
  p->idesc_ab = f16 ? 0u : ((1u << 7) | (1u << 10));
  p->num_src = d->num_src;
  for (int i = 0; i < d->num_src; ++i) p->src_c[i] = d->src[i].c;
  if (f16) {
    // split precision: a source is the [hi | lo] tensor of 2c channels; the K loop runs over the c logical channels
    for (int i = 0; i < d->num_src; ++i) {
      TC_REQUIRE(d->src[i].c % 32 == 0, "conv_tc: split-precision source %d needs 2c %% 32 == 0 (got %d)", i, d->src[i].c);
      p->src_c[i] = d->src[i].c / 2;
    }
    p->split = 1;
  }
  p->taps = d->kh * d->kw;
  p->halo = halo ? 1 : 0;
  p->phase = phase ? 1 : 0;
  p->h = s0.h; p->w = s0.w;
  p->pitch = halo ? s0.w + (hk == 2 ? 1 : 2) : 0;
  p->plane = halo ? (s0.h + (hk == 2 ? 1 : 2)) * p->pitch : s0.h * s0.w;
  long long rows = (long long)s0.n * p->plane;
  if (seg) {
    p->num_seg = d->num_seg;
    rows = 0;
    for (int i = 0; i < d->num_seg; ++i) {
      const cm2_seg& g = d->seg[i];
      TC_REQUIRE(g.row0 % 256 == 0 && g.row0 >= rows && g.n > 0 && g.h > 0 && g.w > 0, "conv_tc: segment %d badly placed", i);
      // g.halo: 0 = every image with its own zero frame, 1 = shared frame (line pitch w + 1, image pitch (h + 1)(w + 1); see halo_kind)
      TC_REQUIRE(g.halo == 0 || g.halo == 1, "conv_tc: segment %d: halo kind %d not supported", i, g.halo);
      const int fr = g.halo == 1 ? 1 : 2;
      const long long srows = (long long)g.n * (g.h + fr) * (g.w + fr);
      TC_REQUIRE(g.row0 + srows < (1ll << 31) - 4096, "conv_tc: segment %d out of range", i);
      p->seg_row0[i] = (int)g.row0; p->seg_rows[i] = (int)srows; p->seg_pitch[i] = g.w + fr;
      p->seg_plane[i] = (g.h + fr) * (g.w + fr); p->seg_h[i] = g.h; p->seg_w[i] = g.w;
      p->seg_img0[i] = i ? p->seg_img0[i - 1] + d->seg[i - 1].n : 0;
      rows = g.row0 + srows;
    }
    p->h = p->w = p->pitch = p->plane = 0;
  }
  TC_REQUIRE(rows > 0 && rows < (1ll << 31) - 4096, "conv_tc: %lld rows out of range", rows);
  if (seg) {
    for (int i = 0; i < d->num_seg; ++i) {
      p->seg_rcp_plane[i] = 1.0 / (double)p->seg_plane[i];
      p->seg_rcp_pitch[i] = 1.0 / (double)p->seg_pitch[i];
    }
  } else {
    p->rcp_plane = 1.0 / (double)p->plane;
    p->rcp_pitch = 1.0 / (double)(halo ? p->pitch : s0.w);
  }
  TC_REQUIRE(!phase || halo, "conv_tc: phase-split sources must be halo views");
  TC_REQUIRE(!phase || rows * 4 < (1ll << 31) - 4096, "conv_tc: phase-split source too large");
  for (int tap = 0; tap < p->taps; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    if (p->taps == 1) p->tap_shift[tap] = 0;
    else if (!phase) p->tap_shift[tap] = (ky - 1) * p->pitch + (kx - 1);
    else {
      // input row 2*oy + ky - 1: ky = 0 -> odd plane, one line up; ky = 1 -> even plane; ky = 2 -> odd plane, same line
      const int py = ky == 1 ? 0 : 1, px = kx == 1 ? 0 : 1, dy = ky == 0 ? -1 : 0, dx = kx == 0 ? -1 : 0;
      p->tap_shift[tap] = (int)((py * 2 + px) * rows) + dy * p->pitch + dx;
    }
  }
  p->rows = (int)rows;
  p->cout = d->cout;
  const int cout_pad = (d->cout + 15) / 16 * 16;
  const int sms = 148;
  if (d->out_mode == 1) {
    TC_REQUIRE((d->cout / 4) % 16 == 0, "conv_tc: deconv scatter needs cout/4 %% 16 == 0");
  }
  const bool pred = d->out_mode == 3;
  if (pred) {
    TC_REQUIRE(halo && !seg && !phase && p->taps == 1 && d->cout % 4 == 0 && (d->cout / 4) % 32 == 0 && d->cout / 4 <= 256,
               "conv_tc: fused deconv + predictor needs a 1x1 GEMM over halo buffers with cout/4 a multiple of 32, <= 256");
    TC_REQUIRE(p->plane % TC_BM == 0, "conv_tc: fused deconv + predictor needs (h+2)*(w+2) %% 128 == 0 (one ROI per tile)");
    TC_REQUIRE(d->pred_w && d->pred_b && d->pred_cls && d->pred_ncls > 0 && d->out_dtype == CM2_F32 && !d->residual.data &&
               !d->scale && !d->stats && d->relu, "conv_tc: fused deconv + predictor: missing predictor / unsupported epilogue flags");
    TC_REQUIRE(d->out.c == 1 && d->out.h == 2 * s0.h && d->out.w == 2 * s0.w && d->out.n == s0.n,
               "conv_tc: fused deconv + predictor writes probabilities [n, 2h, 2w, 1]");
  }
  p->nblk_total = 0;
  for (int i = 0; i < d->num_src; ++i) p->nblk_total += (p->src_c[i] + TC_BK - 1) / TC_BK;
  p->cout_pad = cout_pad;
  // ---- kernel variant: 256-row tiles (v2) whenever that still fills the machine, else 128-row tiles (v1)
  static const int env_variant = getenv("CM2_TC_VARIANT") ? atoi(getenv("CM2_TC_VARIANT")) : 0;
  static const int env_desc = getenv("CM2_TC_DESC_MODE") ? atoi(getenv("CM2_TC_DESC_MODE")) : 0;
  static const int env_dbg = getenv("CM2_TC_DEBUG") ? atoi(getenv("CM2_TC_DEBUG")) : 0;
  p->dbg = env_dbg;
  static const int env_spin = getenv("CM2_TC_SPIN") ? atoi(getenv("CM2_TC_SPIN")) : 0;
  p->spin = env_spin;
  const int m_tiles256 = (int)((rows + 255) / 256);
  int bn2 = pick_bn(cout_pad, m_tiles256, sms);
  // N = 16 (mod 32) layers (FCOS logits, 80 classes): one N tile padded to the next multiple of 32 makes the layer eligible for
  // CTA pairs; the weight rows beyond cout_pad are TMA out-of-bounds zero fill, the epilogue stores co < cout only
  static const int env_padn = getenv("CM2_TC_PAD_N") ? atoi(getenv("CM2_TC_PAD_N")) : 1;   // measured: logits 0.189 -> 0.164 ms
  const bool padded_n = env_padn && bn2 == cout_pad && cout_pad % 32 == 16 && cout_pad >= 48 && cout_pad <= 208 && p->taps == 9 && !phase && !pred;
  if (padded_n) bn2 += 16;
  // Measured (tools/conv_bench.py, B200, batch 16; profiles/r1_convbench_variants_b16.txt): 256-row tiles win
  //   * with the kx-merged slab while both accumulators stay double-buffered (bn <= 128) and the problem has at least
  //     two waves of tiles (small maps, e.g. OSA5 25x42, are faster on 128-row tiles);
  //   * for stride-2 convolutions on phase planes (stem_3: 760 vs 650 TFLOP/s);
  //   * for very short K 1x1 layers whose epilogue is the whole kernel (stem_1, K = 32).
  // Everything else (N = 256 towers, 1x1 aggregations / laterals with K >= 512) is as fast or faster on 128-row tiles.
  const bool merge_ok = p->taps == 9 && !phase;
  int k_total = 0;
  for (int i = 0; i < d->num_src; ++i) k_total += d->src[i].c;
  const int tiles256 = m_tiles256 * ((cout_pad + bn2 - 1) / bn2);
  bool use_v2 = !pred && (merge_ok && bn2 <= 128 && tiles256 >= 2 * sms) || (phase && bn2 <= 128 && tiles256 >= sms) ||
                (p->taps == 1 && k_total <= 128 && tiles256 >= sms);
  // CTA pairs (cta_group::2 MMAs, conv_tc2_kernel<.., true>): measured on the same tool (profiles/r1b_convbench_pair_b16.txt)
  // +17 % / +14 % / +10 % on the stride-1 3x3 layers with N = 128 / 160 / 192 (their single-CTA MMAs are bound by
  // shared-memory operand bandwidth); N = 256 layers and maps with fewer than ~1.5 pair tiles per cluster stay on v1.
  static const int env_pair = getenv("CM2_TC_PAIR") ? atoi(getenv("CM2_TC_PAIR")) : 1;
  const int pair_tiles = ((m_tiles256 + 1) / 2) * ((cout_pad + bn2 - 1) / bn2);
  const bool pair_ok = env_pair >= 1 && !pred && merge_ok && bn2 % 32 == 0 && bn2 <= 224 && (cout_pad == bn2 || padded_n) &&
                       2 * pair_tiles >= 3 * (sms / 2);
  if (pair_ok) use_v2 = true;
  if (env_variant == 1) use_v2 = false;
  if (env_variant >= 2 && !pred) use_v2 = true;
  if (f16 || g_plan_splitk) use_v2 = false;
  static const int env_sets1 = getenv("CM2_TC_EPI_SETS_V1") ? atoi(getenv("CM2_TC_EPI_SETS_V1")) : 2;
  static const int env_sets2 = getenv("CM2_TC_EPI_SETS_V2") ? atoi(getenv("CM2_TC_EPI_SETS_V2")) : 1;
  // 16 epilogue warps (576 threads, <= 96 registers per thread) were measured 3-20 % slower on every layer
  // (profiles/r1_convbench_epilogue_sets_b16.txt), so only 4 or 8 epilogue warps are built
  const int sets1 = env_sets1 == 1 ? 1 : 2;
  const int sets2 = 1;
  (void)env_sets2;
  const size_t tail_v1 = 8 * (2 * 8 + 4) + 48 + 4 * sets1 * EPI_WARP_BYTES + EPI_SS_BYTES;
  const size_t tail_v2 = 8 * (2 * 6 + 2 * 9 + 4) + 48 + 8 * sets2 * EPI_WARP_BYTES + EPI_SS_BYTES;
  const size_t smem_max = 227u * 1024u - 1024u;          // minus the 1 KB alignment slack
  if (use_v2) {
    p->variant = 2;
    p->epi_sets = sets2;
    p->bn = bn2;
    p->m_tiles = m_tiles256;
    p->n_tiles = (cout_pad + p->bn - 1) / p->bn;
    p->kx_merge = (p->taps == 9 && !phase && env_variant != 2) ? 1 : 0;
    p->a_box_rows = p->kx_merge ? 136 : 128;
    p->acc_stages = p->bn <= 128 ? 2 : 1;
    p->desc_mode = env_desc;
    const size_t a_slab = 2u * (size_t)p->a_box_rows * 128u, b_bytes = (size_t)p->bn * TC_BK * 2;
    const int b_tiles = p->taps * p->nblk_total;
    static const int env_bres = getenv("CM2_TC_B_RESIDENT") ? atoi(getenv("CM2_TC_B_RESIDENT")) : 1;
    size_t tail2 = tail_v2;
    if (env_bres && p->n_tiles == 1 && b_tiles <= 64 && (size_t)b_tiles * b_bytes <= 80 * 1024) {
      // small-N layers: weights resident, only A slabs flow through the ring (3 instead of 12 stage hand-offs per tile
      // for stem_2, and the per-tile weight re-read from L2 disappears)
      p->b_resident = 1;
      p->sb_stages = b_tiles;
      tail2 += 16 * (size_t)b_tiles;
      size_t sa = (smem_max - tail2 - (size_t)b_tiles * b_bytes) / a_slab;
      p->sa_stages = (int)(sa > 6 ? 6 : sa);
    } else if (p->kx_merge) {
      p->sa_stages = 3;
      size_t sb = (smem_max - tail_v2 - p->sa_stages * a_slab) / b_bytes;
      p->sb_stages = (int)(sb > 9 ? 9 : sb);
    } else {
      size_t st = (smem_max - tail_v2) / (a_slab + b_bytes);
      p->sa_stages = p->sb_stages = (int)(st > 6 ? 6 : st);
    }
    // CTA pairs (cta_group::2): streamed weights only, N a multiple of 32 (each CTA holds N/2 rows, N/2 % 16 == 0)
    // streamed weights only, N a multiple of 32 (each CTA holds N/2 rows, N/2 % 16 == 0); CM2_TC_PAIR=2 forces pairs on
    // every v2 layer (tests), 0 disables them
    p->pair = ((pair_ok || env_pair == 2) && !p->b_resident && p->bn % 32 == 0 && p->bn >= 32) ? 1 : 0;
    p->smem_bytes = (unsigned)(1024 + p->sa_stages * a_slab + p->sb_stages * b_bytes + tail2);
  } else {
    p->variant = f16 ? 3 : 1;
    p->epi_sets = f16 ? 2 : sets1;
    p->m_tiles = (int)((rows + TC_BM - 1) / TC_BM);
    p->bn = pred ? d->cout / 4 : pick_bn(cout_pad, p->m_tiles, g_plan_splitk ? 1 : sms);   // fused predictor: one N tile per quadrant;
                                                                                           // split-K: widest tile (K slices fill the SMs)
    if (f16) {
      // four accumulators (2 main + 2 cross) share the 512 TMEM columns: N <= 128
      while (p->bn > 128 || cout_pad % p->bn) p->bn -= 16;
      // K-blocks (4 accumulations each) of the main term between two drains.  Measured on the full-size workload
      // (profiles/r2_split_chunk_sweep.txt): chunk 1 / 2 / 4 / 6 -> box error 0.0026 / 0.0051 / 0.0058 / 0.0078 px against
      // the fp32 oracle (the truncation bias grows with the chain length), 316 / 340 / 357 / 354 img/s
      // Final round-2 tree, kx-merged slabs (below).  A chunk = `chunk` weight tiles = 4 * chunk accumulations into the main accumulator.
      // Box error against the fp32 oracle at 800x1333 (tolerance 1e-2 px), three weight / image samples (profiles/
      // r2_parity_seeds.json):   chunk 1: 2.4e-3 / 4.3e-3 / 6.0e-3 px     chunk 2: 2.3e-3 / 5.4e-3 / 1.57e-2 px (the unmerged
      // chunk 2 of the first version: 1.65e-2 on the third sample; the CUDA-core fp32 engine: 3.5e-3 / 6.1e-3 / 4.0e-3).
      // The truncation bias of a long same-sign chain depends on the weights: chunk 2 holds the tolerance on two samples with
      // a wide margin and misses it on the third, so the default is chunk 1 (352 img/s; chunk 2: 385, chunk 4: 404).
      static const int env_chunk = getenv("CM2_TC_CHUNK") ? atoi(getenv("CM2_TC_CHUNK")) : 1;
      p->chunk = env_chunk < 1 ? 1 : env_chunk;
    }
    p->n_tiles = (cout_pad + p->bn - 1) / p->bn;
    // Wave quantisation: the first pitch + 1 and the last pitch + 1 rows of a halo matrix are halo pixels of the first /
    // last image.  When leaving them out saves a whole wave of tiles (16 images of 25x42: 149 tiles on 148 SMs -> 148),
    // the tiles start at the first interior pixel and the launch clears those output rows with two small memsets.
    static const int env_trim = getenv("CM2_TC_TRIM") ? atoi(getenv("CM2_TC_TRIM")) : 1;
    if (env_trim && halo && !seg && !phase && !pred && rows > 4 * (long long)(p->pitch + 1)) {
      const int lead = p->pitch + 1;
      // (shared-halo views end with an interior pixel: only the leading rows can be left out)
      const int t_trim = (int)((rows - (hk == 2 ? 1 : 2) * lead + TC_BM - 1) / TC_BM);
      const int waves_full = (p->m_tiles * p->n_tiles + sms - 1) / sms, waves_trim = (t_trim * p->n_tiles + sms - 1) / sms;
      if (waves_trim < waves_full) { p->row_begin = lead; p->m_tiles = t_trim; }
    }
    p->a_box_rows = TC_BM;
    if (f16) {
      // separate rings for the {x_hi, x_lo} slabs and the {W_hi, W_lo} tiles.  kx-merge (one 136-row slab per filter row of a
      // stride-1 3x3) cuts the x traffic 3x but was measured NOT faster (341 vs 350 img/s end to end, profiles/
      // r2_split_kx_merge.txt): at N = 128 the MMAs themselves read 128 B / clock of shared memory, which is the bound, not
      // the L2 -> shared-memory fill.  Kept selectable (CM2_TC3_MERGE=1; tests/test_gpu_conv_split.py runs both).
      // (Re-measured on the final round-2 tree, chunk 4: 367.6 -> 395.3 img/s with the merged slabs -- with half as many
      // accumulator drains the x traffic is the bound again -- so they are ON; CM2_TC3_MERGE=0 switches back.)
      const int env_merge = getenv("CM2_TC3_MERGE") ? atoi(getenv("CM2_TC3_MERGE")) : 1;
      p->kx_merge = (env_merge && p->taps == 9 && !phase) ? 1 : 0;
      p->a_box_rows = p->kx_merge ? 136 : 128;
      const size_t a_slot = 2u * (size_t)p->a_box_rows * 128u, w_slot = 2u * (size_t)p->bn * TC_BK * 2;
      const size_t tail = (size_t)(8 * (2 * 4 + 2 * 8 + 8) + 48 + 8 * EPI_WARP_BYTES + EPI_SS_BYTES);
      p->sa_stages = p->kx_merge ? 2 : 3;
      size_t sb = (smem_max - tail - p->sa_stages * a_slot) / w_slot;
      p->sb_stages = (int)(sb > 8 ? 8 : sb);
      TC_REQUIRE(p->sb_stages >= 2, "conv_tc: tile does not fit in shared memory");
      p->smem_bytes = (unsigned)(1024 + p->sa_stages * a_slot + p->sb_stages * w_slot + tail);
    } else {
      const uint32_t stage_bytes = TC_A_BYTES + (uint32_t)p->bn * TC_BK * 2;
      int stages = (int)((smem_max - tail_v1) / stage_bytes);
      p->stages = stages > 8 ? 8 : stages;
      p->smem_bytes = (unsigned)(1024 + (size_t)p->stages * stage_bytes + tail_v1);
    }
  }
  p->scale = d->scale; p->shift = d->shift; p->relu = d->relu;
  p->out = d->out.data;
  p->out_sn = d->out.sn; p->out_sh = d->out.sh; p->out_sw = d->out.sw;
  p->out_f32 = d->out_dtype == CM2_F32;
  p->out_mode = d->out_mode;
  p->out_plane = (long long)d->out.n * d->out.sn;
  p->out_halo = (halo && d->out_mode == 0 && halo_kind(d->out) == hk && d->out.h == s0.h && d->out.w == s0.w && !g_plan_splitk) ? 1 : 0;
  const int oc = d->out_mode == 1 ? d->cout / 4 : d->cout;
  const int oeb = p->out_f32 ? 4 : 2;
  p->out_vec = (oc % 16 == 0 && d->out.sn % 8 == 0 && d->out.sh % 8 == 0 && d->out.sw % 8 == 0 &&
                (reinterpret_cast<uintptr_t>(d->out.data) % (size_t)(8 * oeb)) == 0) ? 1 : 0;
  if (seg) {
    TC_REQUIRE(d->out.c == d->cout && d->cout % 16 == 0 && (reinterpret_cast<uintptr_t>(d->out.data) % (size_t)(8 * oeb)) == 0,
               "conv_tc: segmented output needs out.c == cout, cout %% 16 == 0");
    p->out_sw = d->cout; p->out_sn = p->out_sh = 0;
    p->out_halo = 1;
    p->out_vec = 1;
  }
  if (d->residual.data) {
    p->res = reinterpret_cast<const __nv_bfloat16*>(d->residual.data);
    p->res_sn = d->residual.sn; p->res_sh = d->residual.sh; p->res_sw = d->residual.sw;
    p->res_mode = d->res_mode;
    p->res_f32 = f16 ? 1 : 0;                          // the residual has dtype `dtype`; f32 for split-precision convolutions
    if (p->out_vec && !(d->residual.sn % 8 == 0 && d->residual.sh % 8 == 0 && d->residual.sw % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(d->residual.data) & 15) == 0))
      p->out_vec = 0;
  }
  static const int env_store = getenv("CM2_TC_FAST_STORE") ? atoi(getenv("CM2_TC_FAST_STORE")) : 1;
  p->fast_store = (env_store && p->out_vec && (d->out_mode != 1 || (d->cout / 4) % 32 == 0)) ? 1 : 0;
  // the staged epilogue adds a residual of the OUTPUT's element type: bf16 + bf16 residual, or f32 + f32 residual
  if (p->res_mode && (p->out_f32 != 0) != (p->res_f32 != 0)) p->fast_store = 0;
  if (p->out_f32) p->epi_kind = p->res_mode ? 9 : (d->out_mode == 1 ? 10 : 4);
  else p->epi_kind = p->res_mode ? 3 : (d->out_mode == 1 ? 5 : 0);
  if (split_out) {
    TC_REQUIRE(p->fast_store, "conv_tc: the split f16 output needs vectorisable output strides");
    p->split_out = 1;
    p->epi_kind = 11;
  }
  if (pred) {
    p->fast_store = 1;                                 // dispatcher path; nothing is staged
    p->epi_kind = 6;
    p->pred_w = d->pred_w; p->pred_b = d->pred_b; p->pred_cls = reinterpret_cast<const long long*>(d->pred_cls);
    p->pred_ncls = d->pred_ncls;
  }
  if (p->fast_store) TC_REQUIRE(!(p->res_mode && d->out_mode == 1), "conv_tc: unsupported epilogue combination (residual + deconv)");
  if (d->stats) {
    TC_REQUIRE(p->fast_store && !p->res_mode && d->out_mode == 0 && d->cout % 8 == 0,
               "conv_tc: fused statistics need an out_mode-0 output without residual, cout %% 8 == 0");
    p->stats = reinterpret_cast<double*>(d->stats);
    p->stats_mode = d->stats_mode;
    p->epi_kind = p->out_f32 ? 6 + d->stats_mode : d->stats_mode;
    p->stats_stride = d->stats_mode == 1 ? d->cout : (d->cout / 8) * 2;
  }
#undef TC_REQUIRE
  if (!maps) return CM2_OK;
  for (int i = 0; i < d->num_src; ++i) {
    // base of the flat matrix = address of padded pixel (0,0) of image 0
    const char* basep = reinterpret_cast<const char*>(d->src[i].data);
    if (halo && !seg) basep -= (size_t)(d->src[i].sh + d->src[i].sw) * 2;
    if (!encode_2d(&p->a_map[i], basep, (uint64_t)rows * (phase ? 4 : 1), (uint64_t)d->src[i].c, (uint64_t)d->src[i].sw,
                   (uint32_t)p->a_box_rows, f16)) {
      set_error("conv_tc: cuTensorMapEncodeTiled failed for source %d", i);
      return CM2_ERR_CUDA;
    }
  }
  const int64_t ktc = cm2_conv_tc_klen(d->kh, d->kw, d->num_src, p->src_c);
  // split precision: rows [0, cout_pad) hold W_hi, rows [cout_pad, 2 cout_pad) W_lo
  if (!encode_2d(&p->b_map, d->weight, (uint64_t)cout_pad * (f16 ? 2 : 1), (uint64_t)ktc, (uint64_t)ktc,
                 (uint32_t)(p->pair ? p->bn / 2 : p->bn), f16)) {
    set_error("conv_tc: cuTensorMapEncodeTiled failed for the weights");
    return CM2_ERR_CUDA;
  }
  return CM2_OK;
}

// Split-K launch (cm2_conv_desc.splitk >= 2): the K loop of every output tile is cut into `splitk` slices that run as separate
// tiles of the 128-row kernel (small-M layers: the MaskIoU linear layers, P6 / P7, whole stages at small batch -- a handful of
// output tiles cannot fill 148 SMs, their K loops can); fp32 partial sums go to the workspace, splitk_finish_kernel reduces them
// in a fixed order and applies scale / shift / ReLU.  Returns CM2_ERR_UNSUPPORTED (nothing launched) when the layer does not qualify.
static bool tc_pdl_enabled() { return pdl_enabled(); }
// Launch with the programmatic-stream-serialisation attribute (CM2_PDL=0: plain launch; the kernels' griddepcontrol
// instructions are no-ops then).
template <typename K>
static void tc_launch_pdl(K kernel, int grid, int block, unsigned smem, cudaStream_t stream, const TcParams& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, p);
}

static int splitk_plan(const cm2_conv_desc* d, TcParams* pp, bool maps) {
#define SK_REQUIRE(cond, ...) do { if (!(cond)) { set_error(__VA_ARGS__); return CM2_ERR_UNSUPPORTED; } } while (0)
  SK_REQUIRE(d->dtype == CM2_BF16 && (d->out_dtype == CM2_BF16 || d->out_dtype == CM2_F32) && d->out_mode == 0 && !d->residual.data &&
             !d->stats && d->num_seg == 0 && d->cout % 8 == 0, "conv_tc: split-K takes plain bf16 convolutions (out_mode 0, no residual / statistics / segments)");
  SK_REQUIRE(d->splitk_ws && (reinterpret_cast<uintptr_t>(d->splitk_ws) & 31) == 0, "conv_tc: split-K workspace missing or misaligned");
  const long long split_stride = (long long)d->out.n * d->out.sn;
  SK_REQUIRE(d->out.sn % 8 == 0 && d->out.sh % 8 == 0 && d->out.sw % 8 == 0 && d->out.c == d->cout,
             "conv_tc: split-K needs 8-element aligned output strides");
  cm2_conv_desc d2 = *d;
  d2.splitk = 0;
  d2.out.data = d->splitk_ws;
  d2.out_dtype = CM2_F32;
  d2.scale = nullptr; d2.shift = nullptr; d2.relu = 0;
  TcParams& p = *pp;
  g_plan_splitk = true;                                  // tc_plan: 128-row kernel, widest N tile, no halo stores
  int rc = tc_plan(&d2, &p, maps);
  g_plan_splitk = false;
  if (rc != CM2_OK) return rc;
  SK_REQUIRE(p.variant == 1 && p.fast_store && !p.out_halo, "conv_tc: split-K plan fell off the staged 128-row path");
  int ksplit = d->splitk;
  const int kb_total = p.taps * p.nblk_total;
  if (ksplit > kb_total) ksplit = kb_total;
  const int per = (kb_total + ksplit - 1) / ksplit;
  ksplit = (kb_total + per - 1) / per;                   // no empty slice
  SK_REQUIRE(ksplit >= 2, "conv_tc: split-K with fewer than two slices");
  SK_REQUIRE((long long)ksplit * split_stride * 4 <= d->splitk_ws_bytes, "conv_tc: split-K workspace too small (%lld bytes needed)",
             (long long)ksplit * split_stride * 4);
  p.ksplit = ksplit; p.kb_per_split = per; p.split_stride = split_stride;
#undef SK_REQUIRE
  return CM2_OK;
}

static int conv_tc_launch_splitk(const cm2_conv_desc* d, cudaStream_t stream, int sms) {
  TcParams p;
  int rc = splitk_plan(d, &p, true);
  if (rc != CM2_OK) return rc;
  const int ksplit = p.ksplit;
  const long long split_stride = p.split_stride;
  CM2_ENSURE_DYN_SMEM(conv_tc_kernel<320>, 227 * 1024, "conv_tc");
  const int tiles = p.m_tiles * p.n_tiles * ksplit;
  tc_launch_pdl(conv_tc_kernel<320>, tiles < sms ? tiles : sms, 64 + 128 * p.epi_sets, p.smem_bytes, stream, p);
  CM2_CHECK_LAUNCH("conv_tc (split-K)");
  const long long vecs = (long long)d->out.n * d->out.h * d->out.w * (d->cout / 8);
  const int blocks = (int)std::min<long long>((vecs + 255) / 256, 148 * 8);
  const float* part = reinterpret_cast<const float*>(d->splitk_ws);
  if (d->out_dtype == CM2_F32)
    splitk_finish_kernel<float><<<blocks, 256, 0, stream>>>(part, split_stride, ksplit, make_view<float>(d->out), d->scale, d->shift, d->relu);
  else
    splitk_finish_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(part, split_stride, ksplit, make_view<__nv_bfloat16>(d->out), d->scale,
                                                                   d->shift, d->relu);
  CM2_CHECK_LAUNCH("splitk_finish");
  return CM2_OK;
}

int conv_tc_launch(const cm2_conv_desc* d, cudaStream_t stream) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (d->splitk >= 2) return conv_tc_launch_splitk(d, stream, sms);
  TcParams p;
  int rc = tc_plan(d, &p, true);
  if (rc != CM2_OK) return rc;
  CM2_ENSURE_DYN_SMEM(conv_tc_kernel<320>, 227 * 1024, "conv_tc");
  CM2_ENSURE_DYN_SMEM((conv_tc2_kernel<320, false>), 227 * 1024, "conv_tc2");
  CM2_ENSURE_DYN_SMEM((conv_tc2_kernel<320, true>), 227 * 1024, "conv_tc2 (pair)");
  CM2_ENSURE_DYN_SMEM(conv_tc3_kernel<320>, 227 * 1024, "conv_tc3");
  if (p.stats) {
    long long imgs = d->src[0].n;
    if (d->num_seg > 0) {
      imgs = 0;
      for (int i = 0; i < d->num_seg; ++i) imgs += d->seg[i].n;
    }
    if (cudaMemsetAsync(p.stats, 0, (size_t)imgs * p.stats_stride * sizeof(double), stream) != cudaSuccess) {
      set_error("conv_tc: cudaMemsetAsync(stats) failed");
      return CM2_ERR_CUDA;
    }
  }
  if (p.row_begin > 0 && p.out_halo && d->out_mode == 0) {
    // output rows outside the trimmed tile range are halo pixels: keep them zero
    const size_t eb = p.out_f32 ? 4 : 2, row_bytes = (size_t)d->out.sw * eb;
    char* obase = reinterpret_cast<char*>(d->out.data) - (size_t)(d->out.sh + d->out.sw) * eb;
    const long long covered_end = (long long)p.row_begin + (long long)p.m_tiles * TC_BM;
    if (p.variant == 1 && row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(obase) & 15) == 0) {
      p.trim_base = obase;                              // zeroed inside the kernel
      p.trim_lo_bytes = (long long)p.row_begin * (long long)row_bytes;
      p.trim_hi_off = covered_end * (long long)row_bytes;
      p.trim_hi_bytes = covered_end < p.rows ? (long long)(p.rows - covered_end) * (long long)row_bytes : 0;
    } else {
      bool ok = cudaMemsetAsync(obase, 0, (size_t)p.row_begin * row_bytes, stream) == cudaSuccess;
      if (ok && covered_end < p.rows)
        ok = cudaMemsetAsync(obase + (size_t)covered_end * row_bytes, 0, (size_t)(p.rows - covered_end) * row_bytes, stream) == cudaSuccess;
      if (!ok) { set_error("conv_tc: cudaMemsetAsync(halo rows) failed"); return CM2_ERR_CUDA; }
    }
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < sms ? tiles : sms;
  if (p.variant == 2 && p.pair) {
    const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
    int clusters = sms / 2;
    if (pair_tiles < clusters) clusters = pair_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(64 + 256 * p.epi_sets);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tc_pdl_enabled() ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, conv_tc2_kernel<320, true>, p) != cudaSuccess) {
      set_error("conv_tc: cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      return CM2_ERR_CUDA;
    }
  } else if (p.variant == 2)
    tc_launch_pdl(conv_tc2_kernel<320, false>, grid, 64 + 256 * p.epi_sets, p.smem_bytes, stream, p);
  else if (p.variant == 3)
    tc_launch_pdl(conv_tc3_kernel<320>, grid, 320, p.smem_bytes, stream, p);
  else
    tc_launch_pdl(conv_tc_kernel<320>, grid, 64 + 128 * p.epi_sets, p.smem_bytes, stream, p);
  CM2_CHECK_LAUNCH("conv_tc");
  return CM2_OK;
}

}  // namespace cm2

extern "C" int64_t cm2_conv_tc_klen(int32_t kh, int32_t kw, int32_t num_src, const int32_t* src_c) {
  int64_t k = 0;
  for (int i = 0; i < num_src; ++i) k += (src_c[i] + 63) / 64 * 64;
  return k * kh * kw;
}

extern "C" int cm2_conv_tc_supported(const cm2_conv_desc* d) {
  if (!d || d->num_src < 1 || d->num_src > CM2_MAX_SRC) return 0;
  cm2::TcParams p;
  if (d->splitk >= 2) return cm2::splitk_plan(d, &p, false) == CM2_OK ? 1 : 0;
  return cm2::tc_plan(d, &p, false) == CM2_OK ? 1 : 0;
}
