// FCOS post-process: decode + threshold + compaction, per-level top-k, class-aware NMS, post top-k.
//
// Reference: modeling/fcos/fcos_outputs.py:372-495 and layers/ml_nms.py:93-96.  The reference loops
// over images and levels in Python and calls nonzero / index / torchvision NMS; here the whole
// post-process of a batch is three launches per batch (plus one decode launch per level).
//
// Determinism: candidates are appended with atomics (unordered), then every (image, level) segment
// is sorted by (raw score desc, flat index asc), which makes all later stages order-independent.
#include "common.cuh"
#include <algorithm>
#include <string.h>
#include <stdlib.h>

namespace cm2 {

// ---------------------------------------------------------------------------------------------
// decode: one warp per location; lanes stride over classes so the logits row is read coalesced.
// ---------------------------------------------------------------------------------------------
// `logit_floor` = logit(thresh) - 1e-3: logits below it cannot pass `sigmoid(x) > thresh` (nor `sigmoid(x) * ctr > thresh`,
// ctr <= 1), so the exact ATen-arithmetic sigmoid is evaluated only for the ~0.1% of logits near or above the threshold.
__device__ __forceinline__ void fcos_emit_candidate(const View<const float>& regctr, int img, int py_, int px_, int c, int ncls,
                                                    float xl, int stride, float reg_scale, float thresh, int thresh_with_ctr,
                                                    int level, int num_levels, int cap, const cm2_cand_buffers& cand) {
  const float* rrow = regctr.at(img, py_, px_);
  const float ctr = sigmoid_f32(__ldg(rrow + 4));
  const float p = sigmoid_f32(xl);
  const float s = p * ctr;
  if (!(thresh_with_ctr ? (s > thresh) : (p > thresh))) return;
  const int seg = img * num_levels + level;
  const int slot = atomicAdd(cand.count + seg, 1);
  if (slot >= cap) return;
  float l = fmaxf(__ldg(rrow + 0) * reg_scale, 0.f) * (float)stride;
  float t = fmaxf(__ldg(rrow + 1) * reg_scale, 0.f) * (float)stride;
  float r = fmaxf(__ldg(rrow + 2) * reg_scale, 0.f) * (float)stride;
  float b = fmaxf(__ldg(rrow + 3) * reg_scale, 0.f) * (float)stride;
  float px = (float)(px_ * stride + stride / 2);
  float py = (float)(py_ * stride + stride / 2);
  size_t o = (size_t)seg * cap + slot;
  reinterpret_cast<float4*>(cand.boxes)[o] = make_float4(px - l, py - t, px + r, py + b);
  cand.score[o] = s;
  cand.cls[o] = c;
  cand.flat[o] = (py_ * regctr.w + px_) * ncls + c;
}

// Streaming variant (needs ncls % 4 == 0, pixel stride == ncls and 16-byte aligned rows): grid (row-chunks, h, n); an
// image row of the level is w * ncls contiguous floats, read as float4 with four loads in flight per thread -- the
// common case costs one 128-bit load and four compares per four logits.
__global__ void __launch_bounds__(256) fcos_decode_rows_kernel(View<const float> logits, View<const float> regctr, int stride,
                                                               float reg_scale, float thresh, float logit_floor, int thresh_with_ctr,
                                                               int level, int num_levels, int cap, cm2_cand_buffers cand) {
  const int img = blockIdx.z, py_ = blockIdx.y;
  const int ncls = logits.c;
  const int n4 = (logits.w * ncls) >> 2;
  const float4* row = reinterpret_cast<const float4*>(logits.at(img, py_, 0));
  constexpr int U = 4;
  for (int i0 = blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n4; i0 += gridDim.x * blockDim.x * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * (int)blockDim.x < n4) v[u] = __ldg(row + i0 + u * blockDim.x);
      else v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float m4 = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
      if (i0 + u * (int)blockDim.x < n4 && m4 >= logit_floor) {      // rare
        const int e = (i0 + u * (int)blockDim.x) * 4;
        const int px_ = e / ncls, c = e - px_ * ncls;
        const float vals[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (vals[k] >= logit_floor)
            fcos_emit_candidate(regctr, img, py_, px_, c + k, ncls, vals[k], stride, reg_scale, thresh, thresh_with_ctr, level,
                                num_levels, cap, cand);
      }
    }
  }
}

// All pyramid levels in one launch (the four small levels are a few microseconds of work each): grid
// (row-chunks, sum of level heights, n); the block looks its level up from the row prefix.
constexpr int FCOS_MAX_LEVELS = 8;
struct DecodeLevels {
  View<const float> logits[FCOS_MAX_LEVELS], regctr[FCOS_MAX_LEVELS];
  int stride[FCOS_MAX_LEVELS], row_prefix[FCOS_MAX_LEVELS + 1];
  float reg_scale[FCOS_MAX_LEVELS];
  int num;
};

__global__ void __launch_bounds__(256) fcos_decode_levels_kernel(const __grid_constant__ DecodeLevels lv, float thresh, float logit_floor,
                                                                 int thresh_with_ctr, int cap, cm2_cand_buffers cand) {
  int level = 0;
#pragma unroll
  for (int l = 1; l < FCOS_MAX_LEVELS; ++l)
    if (l < lv.num && (int)blockIdx.y >= lv.row_prefix[l]) level = l;
  const View<const float>& logits = lv.logits[level];
  const View<const float>& regctr = lv.regctr[level];
  const int img = blockIdx.z, py_ = blockIdx.y - lv.row_prefix[level];
  const int ncls = logits.c;
  const int n4 = (logits.w * ncls) >> 2;
  const float4* row = reinterpret_cast<const float4*>(logits.at(img, py_, 0));
  constexpr int U = 4;
  for (int i0 = blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n4; i0 += gridDim.x * blockDim.x * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * (int)blockDim.x < n4) v[u] = __ldg(row + i0 + u * blockDim.x);
      else v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float m4 = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
      if (i0 + u * (int)blockDim.x < n4 && m4 >= logit_floor) {      // rare
        const int e = (i0 + u * (int)blockDim.x) * 4;
        const int px_ = e / ncls, c = e - px_ * ncls;
        const float vals[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (vals[k] >= logit_floor)
            fcos_emit_candidate(regctr, img, py_, px_, c + k, ncls, vals[k], lv.stride[level], lv.reg_scale[level], thresh,
                                thresh_with_ctr, level, lv.num, cap, cand);
      }
    }
  }
}

// Flat-tile variant (every level's image is one dense run of h * w * ncls floats): the batch is cut into tiles of
// DECODE_TILE float4 (32 KB); a CTA takes one tile and every thread has its DECODE_U 16-byte loads in flight before the
// first compare -- 7 104 full CTAs for the 800x1344 pyramid at batch 32 instead of 24 960 mostly idle row CTAs.
constexpr int DECODE_U = 8;
constexpr int DECODE_TILE = 256 * DECODE_U;
struct DecodeTiles {
  DecodeLevels lv;
  int tile_prefix[FCOS_MAX_LEVELS + 1];                 // tiles per image, prefix over levels
};

__global__ void __launch_bounds__(256) fcos_decode_tiles_kernel(const __grid_constant__ DecodeTiles dt, float thresh, float logit_floor,
                                                                int thresh_with_ctr, int cap, cm2_cand_buffers cand) {
  const DecodeLevels& lv = dt.lv;
  // grid (n, tiles): the image index varies fastest, so the CTAs in flight append to the counters of many
  // (image, level) segments instead of queueing on a handful of addresses
  const int img = blockIdx.x, tile = blockIdx.y;
  int level = 0;
#pragma unroll
  for (int l = 1; l < FCOS_MAX_LEVELS; ++l)
    if (l < lv.num && tile >= dt.tile_prefix[l]) level = l;
  const View<const float>& logits = lv.logits[level];
  const int ncls = logits.c;
  const int n4 = (logits.h * logits.w * ncls) >> 2;
  const float* img_base = logits.p + img * logits.sn;
  const float4* base = reinterpret_cast<const float4*>(img_base);
  const int i0 = (tile - dt.tile_prefix[level]) * DECODE_TILE + threadIdx.x;
  float4 v[DECODE_U];
#pragma unroll
  for (int u = 0; u < DECODE_U; ++u)
    if (i0 + u * 256 < n4) v[u] = __ldg(base + i0 + u * 256);
    else v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  // one bit per logit that reaches the floor; the (rare) candidates are then handled by ONE copy of the emit code in a
  // loop over the set bits -- unrolling the emit path DECODE_U x 4 times made the kernel instruction-fetch bound
  uint32_t hits = 0u;
#pragma unroll
  for (int u = 0; u < DECODE_U; ++u) {
    hits |= (v[u].x >= logit_floor ? 1u : 0u) << (4 * u);
    hits |= (v[u].y >= logit_floor ? 2u : 0u) << (4 * u);
    hits |= (v[u].z >= logit_floor ? 4u : 0u) << (4 * u);
    hits |= (v[u].w >= logit_floor ? 8u : 0u) << (4 * u);
  }
  while (hits) {
    const int bit = __ffs(hits) - 1;
    hits &= hits - 1u;
    const int e = (i0 + (bit >> 2) * 256) * 4 + (bit & 3);
    const int pos = e / ncls, c = e - pos * ncls;
    const int py_ = pos / logits.w, px_ = pos - py_ * logits.w;
    fcos_emit_candidate(lv.regctr[level], img, py_, px_, c, ncls, __ldg(img_base + e), lv.stride[level], lv.reg_scale[level], thresh,
                        thresh_with_ctr, level, lv.num, cap, cand);
  }
}

// Generic variant: one warp per location; lanes stride over classes.
__global__ void fcos_decode_kernel(View<const float> logits, View<const float> regctr, int stride, float reg_scale, float thresh,
                                   float logit_floor, int thresh_with_ctr, int level, int num_levels, int cap,
                                   cm2_cand_buffers cand) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int h = logits.h, w = logits.w, ncls = logits.c;
  const int hw = h * w;
  const int64_t total = (int64_t)logits.n * hw;
  for (int64_t loc = warp; loc < total; loc += nwarps) {
    const int img = (int)(loc / hw);
    const int pos = (int)(loc - (int64_t)img * hw);
    const int py_ = pos / w, px_ = pos - py_ * w;
    const float* lrow = logits.at(img, py_, px_);
    for (int c0 = 0; c0 < ncls; c0 += 32) {
      const int c = c0 + lane;
      const float xl = c < ncls ? __ldg(lrow + c) : 0.f;
      if (c < ncls && xl >= logit_floor)
        fcos_emit_candidate(regctr, img, py_, px_, c, ncls, xl, stride, reg_scale, thresh, thresh_with_ctr, level, num_levels, cap,
                            cand);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// in-CTA bitonic sort of 64-bit keys, descending.  `n` must be a power of two.
// ---------------------------------------------------------------------------------------------
__device__ void bitonic_sort_desc(unsigned long long* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          bool up = (i & k) == 0;          // descending in the "up" half
          if (up ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

struct SelectWs {
  float* boxes;        // [n][levels][pre][4]   per-level sorted + truncated candidates
  float* score;        // [n][levels][pre]
  int* cls;            // [n][levels][pre]
  int* flat;           // [n][levels][pre]
  int* kept;           // [n][levels]
};

// grid (levels, n).  Sort one (image, level) segment by (raw score desc, flat asc), keep pre_topk.
__global__ void fcos_select_level_kernel(cm2_cand_buffers cand, int num_levels, int cap, int pre_topk, int sort_n,
                                         SelectWs ws) {
  extern __shared__ unsigned long long keys[];
  const int level = blockIdx.x, img = blockIdx.y;
  const int seg = img * num_levels + level;
  int cnt = min(cand.count[seg], cap);
  // sort only as many keys as there are candidates (power of two, <= the launch-time bound)
  {
    int sn = 32;
    while (sn < cnt) sn <<= 1;
    sort_n = min(sort_n, sn);
  }
  for (int i = threadIdx.x; i < sort_n; i += blockDim.x) {
    unsigned long long k = 0ull;
    if (i < cnt) {
      size_t o = (size_t)seg * cap + i;
      // primary: raw score bits (positive floats order as unsigned ints); secondary: lower flat first.
      // The slot index rides in a side array because 64 bits are used up.
      k = ((unsigned long long)__float_as_uint(cand.score[o]) << 32) | (unsigned)(0xffffffffu - (unsigned)cand.flat[o]);
    }
    keys[i] = k;
  }
  __syncthreads();
  bitonic_sort_desc(keys, sort_n);
  int keep = min(cnt, pre_topk);
  if (threadIdx.x == 0) ws.kept[seg] = keep;
  // map sorted keys back to slots: flat is unique within a segment, so search it.  To stay O(n log n)
  // we sort a second time on (flat) would be wasteful; instead build an index: every thread scans for
  // its key's slot through a shared flat->slot pass below.
  // Pass: each candidate slot finds its rank by binary search over the sorted keys.
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    size_t o = (size_t)seg * cap + i;
    unsigned long long k =
        ((unsigned long long)__float_as_uint(cand.score[o]) << 32) | (unsigned)(0xffffffffu - (unsigned)cand.flat[o]);
    int lo = 0, hi = cnt - 1;             // keys[0..cnt) descending, all distinct
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (keys[mid] > k) lo = mid + 1; else hi = mid;
    }
    int rank = lo;
    if (rank < keep) {
      size_t d = (size_t)seg * pre_topk + rank;
      reinterpret_cast<float4*>(ws.boxes)[d] = reinterpret_cast<const float4*>(cand.boxes)[o];
      ws.score[d] = cand.score[o];
      ws.cls[d] = cand.cls[o];
      ws.flat[d] = cand.flat[o];
    }
  }
}

__device__ __forceinline__ bool iou_gt(const float4& a, float area_a, const float4& b, float area_b, float thr) {
  // torchvision nms_kernel (CPU): inter / (area_a + area_b - inter) > thr
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
  float inter = w * h;
  float ovr = inter / (area_a + area_b - inter);
  return ovr > thr;
}

// grid (n); 1024 threads order the candidates of all levels, warp 0 runs the greedy sweep with early exit at post_topk
// survivors.  MERGE: the per-level lists arrive sorted (fcos_select_level_kernel), so the global order is a multi-way
// merge: rank = own rank + the number of larger keys in every other level (binary searches), written straight to its
// place -- ncu of the bitonic version (8192 keys, 91 stages) showed the sort as ~90 % of the kernel's 570 k warp
// instructions per image.  Same total order (keys are distinct), hence identical results.
constexpr int NMS_MAX_LEVELS = 8;
template <bool MERGE>
__global__ void fcos_nms_image_kernel(SelectWs ws, int num_levels, int pre_topk, int sort_n, const int* level_w,
                                      const int* level_stride, int ncls, float nms_thresh, int post_topk,
                                      cm2_det_buffers det) {
  extern __shared__ unsigned long long keys_all[];      // MERGE: [2][levels * pre_topk], else [sort_n]
  __shared__ float4 kept_box[256];
  __shared__ float kept_area[256];
  __shared__ int kept_cls[256];
  __shared__ int kept_idx[256];
  __shared__ int s_off[NMS_MAX_LEVELS + 1];
  const int img = blockIdx.x;
  const int total_slots = num_levels * pre_topk;
  unsigned long long* keys = keys_all;
  auto make_key = [&](int level, int r) {
    const size_t o = (size_t)(img * num_levels + level) * pre_topk + r;
    const float sc = sqrtf(ws.score[o]);                   // fcos_outputs.py:460
    // secondary key: earlier (level, rank) first -- deterministic because segments are sorted
    unsigned long long k = ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned)(0xffffffffu - (unsigned)(level * pre_topk + r));
    return k == 0ull ? 1ull : k;
  };
  if (MERGE) {
    if (threadIdx.x == 0) {
      int run = 0;
      for (int l = 0; l < num_levels; ++l) { s_off[l] = run; run += ws.kept[img * num_levels + l]; }
      s_off[num_levels] = run;
    }
    __syncthreads();
    unsigned long long* lists = keys_all + total_slots;   // level-sorted keys, packed back to back
    for (int i = threadIdx.x; i < total_slots; i += blockDim.x) {
      const int level = i / pre_topk, r = i - level * pre_topk;
      if (r < s_off[level + 1] - s_off[level]) lists[s_off[level] + r] = make_key(level, r);
      keys[i] = 0ull;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total_slots; i += blockDim.x) {
      const int level = i / pre_topk, r = i - level * pre_topk;
      if (r >= s_off[level + 1] - s_off[level]) continue;
      const unsigned long long k = lists[s_off[level] + r];
      int rank = r;
      for (int l = 0; l < num_levels; ++l) {
        if (l == level) continue;
        int lo = s_off[l], hi = s_off[l + 1];              // descending list: count the keys > k
        const int base = lo;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (lists[mid] > k) lo = mid + 1; else hi = mid;
        }
        rank += lo - base;
      }
      keys[rank] = k;
    }
    __syncthreads();
  } else {
    for (int i = threadIdx.x; i < sort_n; i += blockDim.x) {
      unsigned long long k = 0ull;
      if (i < total_slots) {
        int level = i / pre_topk, r = i - level * pre_topk;
        if (r < ws.kept[img * num_levels + level]) k = make_key(level, r);
      }
      keys[i] = k;
    }
    __syncthreads();
    bitonic_sort_desc(keys, sort_n);
  }
  sort_n = MERGE ? total_slots : sort_n;

  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int nkept = 0;
    for (int base = 0; base < total_slots && nkept < post_topk; base += 32) {
      unsigned long long k = (base + lane < sort_n) ? keys[base + lane] : 0ull;
      bool alive = k != 0ull;
      float4 bx = make_float4(0, 0, 0, 0);
      float area = 0.f;
      int cls = -1, slot = 0;
      if (alive) {
        slot = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
        int level = slot / pre_topk, r = slot - level * pre_topk;
        size_t o = (size_t)(img * num_levels + level) * pre_topk + r;
        bx = reinterpret_cast<const float4*>(ws.boxes)[o];
        area = (bx.z - bx.x) * (bx.w - bx.y);
        cls = ws.cls[o];
        for (int j = 0; j < nkept; ++j)
          if (kept_cls[j] == cls && iou_gt(kept_box[j], kept_area[j], bx, area, nms_thresh)) { alive = false; break; }
      }
      unsigned m = __ballot_sync(0xffffffffu, alive);
      if (__ballot_sync(0xffffffffu, k != 0ull) == 0u) break;      // ran out of candidates
      while (m != 0u && nkept < post_topk) {
        int leader = __ffs(m) - 1;
        float4 lb;
        lb.x = __shfl_sync(0xffffffffu, bx.x, leader);
        lb.y = __shfl_sync(0xffffffffu, bx.y, leader);
        lb.z = __shfl_sync(0xffffffffu, bx.z, leader);
        lb.w = __shfl_sync(0xffffffffu, bx.w, leader);
        float la = __shfl_sync(0xffffffffu, area, leader);
        int lc = __shfl_sync(0xffffffffu, cls, leader);
        int ls = __shfl_sync(0xffffffffu, slot, leader);
        if (lane == 0) { kept_box[nkept] = lb; kept_area[nkept] = la; kept_cls[nkept] = lc; kept_idx[nkept] = ls; }
        ++nkept;
        if (alive && lane > leader && cls == lc && iou_gt(lb, la, bx, area, nms_thresh)) alive = false;
        if (lane == leader) alive = false;
        m = __ballot_sync(0xffffffffu, alive);
      }
      __syncwarp();
    }
    __syncwarp();
    // write-out: survivors are already in descending score order
    for (int j = lane; j < post_topk; j += 32) {
      size_t d = (size_t)img * post_topk + j;
      if (j < nkept) {
        int slot = kept_idx[j];
        int level = slot / pre_topk, r = slot - level * pre_topk;
        size_t o = (size_t)(img * num_levels + level) * pre_topk + r;
        reinterpret_cast<float4*>(det.boxes)[d] = kept_box[j];
        det.scores[d] = sqrtf(ws.score[o]);
        det.classes[d] = (int64_t)kept_cls[j];
        int pos = ws.flat[o] / ncls;
        int w = level_w[level], st = level_stride[level];
        int y = pos / w, x = pos - y * w;
        det.locations[2 * d] = (float)(x * st + st / 2);
        det.locations[2 * d + 1] = (float)(y * st + st / 2);
      } else {
        reinterpret_cast<float4*>(det.boxes)[d] = make_float4(0, 0, 0, 0);
        det.scores[d] = 0.f;
        det.classes[d] = 0;
        det.locations[2 * d] = 0.f;
        det.locations[2 * d + 1] = 0.f;
      }
    }
    if (lane == 0) det.count[img] = nkept;
  }
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static SelectWs carve_ws(void* workspace, int n, int num_levels, int pre) {
  SelectWs ws;
  char* p = reinterpret_cast<char*>(workspace);
  size_t slots = (size_t)n * num_levels * pre;
  ws.boxes = reinterpret_cast<float*>(p); p += slots * 16;
  ws.score = reinterpret_cast<float*>(p); p += slots * 4;
  ws.cls = reinterpret_cast<int*>(p); p += slots * 4;
  ws.flat = reinterpret_cast<int*>(p); p += slots * 4;
  ws.kept = reinterpret_cast<int*>(p);
  return ws;
}

}  // namespace cm2

using namespace cm2;

extern "C" int cm2_fcos_decode(const cm2_act* logits, const cm2_act* regctr, int32_t stride, float reg_scale, float thresh,
                               int32_t thresh_with_ctr, int32_t level, int32_t num_levels, int32_t cap,
                               const cm2_cand_buffers* cand, void* stream) {
  CM2_CHECK_ARG(logits && regctr && logits->data && regctr->data && cand && cand->boxes && cand->score && cand->cls &&
                cand->flat && cand->count, "fcos_decode: null pointer");
  CM2_CHECK_ARG(regctr->c >= 5 && logits->c > 0 && cap > 0 && level >= 0 && level < num_levels,
                "fcos_decode: bad arguments");
  CM2_CHECK_ARG(regctr->n == logits->n && regctr->h == logits->h && regctr->w == logits->w,
                "fcos_decode: logits / regctr extents differ");
  CM2_CHECK_ARG((int64_t)logits->h * logits->w * logits->c < (1ll << 31), "fcos_decode: level too large");
  int64_t total = (int64_t)logits->n * logits->h * logits->w;
  if (total == 0) return CM2_OK;
  int64_t blocks = ceil_div64(total * 32, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  // thresh in (0, 1): prefilter in logit space with a 1e-3 safety margin; otherwise evaluate everything
  const float logit_floor = (thresh > 0.f && thresh < 1.f) ? (float)(log((double)thresh / (1.0 - (double)thresh)) - 1e-3) : -INFINITY;
  if (logits->c % 4 == 0 && logits->sw == logits->c && logits->sh % 4 == 0 && logits->sn % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(logits->data) & 15) == 0 && logits->h <= 65535 && logits->n <= 65535) {
    const int n4 = logits->w * logits->c / 4;
    dim3 grid(std::max(1, std::min(8, ceil_div(n4, 256 * 4))), logits->h, logits->n);
    fcos_decode_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(make_view<const float>(*logits), make_view<const float>(*regctr),
                                                                    stride, reg_scale, thresh, logit_floor, thresh_with_ctr, level,
                                                                    num_levels, cap, *cand);
    CM2_CHECK_LAUNCH("fcos_decode_rows");
    return CM2_OK;
  }
  fcos_decode_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(make_view<const float>(*logits),
                                                                   make_view<const float>(*regctr), stride, reg_scale, thresh,
                                                                   logit_floor, thresh_with_ctr, level, num_levels, cap, *cand);
  CM2_CHECK_LAUNCH("fcos_decode");
  return CM2_OK;
}

extern "C" int cm2_fcos_decode_levels(const cm2_act* logits, const cm2_act* regctr, const int32_t* strides, const float* reg_scales,
                                      int32_t num_levels, float thresh, int32_t thresh_with_ctr, int32_t cap,
                                      const cm2_cand_buffers* cand, void* stream) {
  CM2_CHECK_ARG(logits && regctr && strides && reg_scales && cand, "fcos_decode_levels: null pointer");
  CM2_CHECK_ARG(num_levels >= 1 && num_levels <= FCOS_MAX_LEVELS && cap > 0, "fcos_decode_levels: bad level count %d / cap %d", num_levels, cap);
  bool fast = true;
  for (int l = 0; l < num_levels; ++l) {
    const cm2_act& a = logits[l];
    CM2_CHECK_ARG(a.data && regctr[l].data && a.n == logits[0].n && regctr[l].n == a.n && regctr[l].h == a.h && regctr[l].w == a.w &&
                  regctr[l].c >= 5 && a.c == logits[0].c, "fcos_decode_levels: level %d views do not match", l);
    CM2_CHECK_ARG((int64_t)a.h * a.w * a.c < (1ll << 31), "fcos_decode_levels: level too large");
    fast = fast && a.c % 4 == 0 && a.sw == a.c && a.sh % 4 == 0 && a.sn % 4 == 0 && (reinterpret_cast<uintptr_t>(a.data) & 15) == 0;
  }
  if (logits[0].n == 0) return CM2_OK;
  if (!fast || logits[0].n > 65535) {                      // generic per-level path
    for (int l = 0; l < num_levels; ++l) {
      int rc = cm2_fcos_decode(&logits[l], &regctr[l], strides[l], reg_scales[l], thresh, thresh_with_ctr, l, num_levels, cap, cand, stream);
      if (rc != CM2_OK) return rc;
    }
    return CM2_OK;
  }
  DecodeLevels lv;
  memset(&lv, 0, sizeof(lv));
  lv.num = num_levels;
  int max_n4 = 0;
  for (int l = 0; l < num_levels; ++l) {
    lv.logits[l] = make_view<const float>(logits[l]);
    lv.regctr[l] = make_view<const float>(regctr[l]);
    lv.stride[l] = strides[l];
    lv.reg_scale[l] = reg_scales[l];
    lv.row_prefix[l + 1] = lv.row_prefix[l] + logits[l].h;
    max_n4 = std::max(max_n4, logits[l].w * logits[l].c / 4);
  }
  CM2_CHECK_ARG(lv.row_prefix[num_levels] <= 65535, "fcos_decode_levels: too many rows");
  const float logit_floor = (thresh > 0.f && thresh < 1.f) ? (float)(log((double)thresh / (1.0 - (double)thresh)) - 1e-3) : -INFINITY;
  const int variant = getenv("CM2_DECODE_VARIANT") ? atoi(getenv("CM2_DECODE_VARIANT")) : 1;
  bool dense = variant == 1;
  for (int l = 0; l < num_levels; ++l) dense = dense && logits[l].sh == (int64_t)logits[l].w * logits[l].c;
  int tiles_total = 0;
  for (int l = 0; l < num_levels; ++l) tiles_total += ceil_div(logits[l].h * logits[l].w * logits[l].c / 4, DECODE_TILE);
  dense = dense && tiles_total <= 65535;
  if (dense) {
    DecodeTiles dt;
    dt.lv = lv;
    dt.tile_prefix[0] = 0;
    for (int l = 0; l < FCOS_MAX_LEVELS; ++l) {
      const int n4 = l < num_levels ? logits[l].h * logits[l].w * logits[l].c / 4 : 0;
      dt.tile_prefix[l + 1] = dt.tile_prefix[l] + ceil_div(n4, DECODE_TILE);
    }
    dim3 grid(logits[0].n, dt.tile_prefix[num_levels]);
    fcos_decode_tiles_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dt, thresh, logit_floor, thresh_with_ctr, cap, *cand);
    CM2_CHECK_LAUNCH("fcos_decode_tiles");
    return CM2_OK;
  }
  dim3 grid(std::max(1, std::min(4, ceil_div(max_n4, 256 * 4))), lv.row_prefix[num_levels], logits[0].n);
  fcos_decode_levels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(lv, thresh, logit_floor, thresh_with_ctr, cap, *cand);
  CM2_CHECK_LAUNCH("fcos_decode_levels");
  return CM2_OK;
}

extern "C" int64_t cm2_fcos_select_workspace(int32_t n, int32_t num_levels, int32_t cap) {
  // sized for pre_topk <= cap
  size_t slots = (size_t)n * num_levels * cap;
  return (int64_t)(slots * 28 + (size_t)n * num_levels * 4 + 256);
}

extern "C" int cm2_fcos_select(const cm2_cand_buffers* cand, int32_t n, int32_t num_levels, int32_t cap,
                               const int32_t* level_w, const int32_t* level_stride, int32_t ncls, int32_t pre_topk,
                               float nms_thresh, int32_t post_topk, const cm2_det_buffers* det, void* workspace,
                               void* stream) {
  CM2_CHECK_ARG(cand && det && workspace && level_w && level_stride, "fcos_select: null pointer");
  CM2_CHECK_ARG(pre_topk > 0 && pre_topk <= cap, "fcos_select: pre_topk %d must be in (0, cap=%d]", pre_topk, cap);
  CM2_CHECK_ARG(post_topk > 0 && post_topk <= 256, "fcos_select: post_topk %d must be in (0, 256]", post_topk);
  int sort_a = next_pow2(cap);
  int sort_b = next_pow2(num_levels * pre_topk);
  CM2_CHECK_ARG(sort_a <= 16384 && sort_b <= 16384, "fcos_select: cap %d / levels*pre_topk %d exceed the in-CTA sort (16384)",
                cap, num_levels * pre_topk);
  if (n == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  SelectWs ws = carve_ws(workspace, n, num_levels, pre_topk);
  CM2_ENSURE_DYN_SMEM(fcos_select_level_kernel, 16384 * 8, "fcos_select_level");
  CM2_ENSURE_DYN_SMEM(fcos_nms_image_kernel<false>, 16384 * 8, "fcos_nms_image");
  CM2_ENSURE_DYN_SMEM(fcos_nms_image_kernel<true>, 16384 * 8, "fcos_nms_image");
  dim3 g1(num_levels, n);
  fcos_select_level_kernel<<<g1, 1024, (size_t)sort_a * 8, s>>>(*cand, num_levels, cap, pre_topk, sort_a, ws);
  CM2_CHECK_LAUNCH("fcos_select_level");
  const int nms_variant = getenv("CM2_NMS_VARIANT") ? atoi(getenv("CM2_NMS_VARIANT")) : 1;
  if (nms_variant == 1 && num_levels <= NMS_MAX_LEVELS && 2 * num_levels * pre_topk <= 16384)
    fcos_nms_image_kernel<true><<<n, 1024, (size_t)2 * num_levels * pre_topk * 8, s>>>(ws, num_levels, pre_topk, sort_b, level_w,
                                                                                     level_stride, ncls, nms_thresh, post_topk, *det);
  else
    fcos_nms_image_kernel<false><<<n, 1024, (size_t)sort_b * 8, s>>>(ws, num_levels, pre_topk, sort_b, level_w, level_stride, ncls,
                                                                   nms_thresh, post_topk, *det);
  CM2_CHECK_LAUNCH("fcos_nms_image");
  return CM2_OK;
}
