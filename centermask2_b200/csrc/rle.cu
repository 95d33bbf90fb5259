// COCO run-length encoding of pasted masks on the device.  Replaces, for the masks that
// paste-back leaves in HBM, the per-mask `mask_util.encode(np.array(mask[:, :, None], order="F"))` loop of
// /root/reference/centermask2/centermask/evaluation/coco_evaluation.py:388-391 (pycocotools' rleEncode): instead of
// copying R x H x W bytes to the host (1 MB per mask), only the run lengths travel.
//
// A mask is scanned in COLUMN-major order (the COCO convention); a "transition" is a pixel whose value differs from
// its predecessor in that order (the predecessor of pixel 0 is 0).  Runs = differences of transition positions.
//   1. rle_count_kernel      transitions per (mask, column)                       -> col_count [r][w]
//   2. rle_scan_kernel       exclusive scan over the columns of each mask         -> col_offset [r][w], total [r]
//   3. rle_write_kernel      run lengths, written at mask_offset[m] + col_offset  -> runs (uint32)
// One thread walks one column (adjacent threads read adjacent bytes of a row: coalesced); the caller sizes `runs`
// from `total` (one small D2H copy) between steps 2 and 3.
#include "common.cuh"

namespace cm2 {

// Optional window hint: a pasted mask is zero outside the dilated box window of its ROI (paste_masks_in_image [d2] writes
// only [floor(x0) - 1, ceil(x1) + 1) x [floor(y0) - 1, ceil(y1) + 1)), so the column walk can skip everything else -- the
// boxes cover ~10 % of an image on average, which is what the scan then reads.  The window is widened by one more pixel on
// every side; an invalid slot has an empty window.
struct RleWindow { int xa, xb, ya, yb; };
constexpr int RLE_ROWS = 8;                                // rows fetched per step of a column walk
__device__ __forceinline__ RleWindow rle_window(const float* __restrict__ boxes, const uint8_t* __restrict__ valid, int m, int h, int w) {
  RleWindow q;
  if (!boxes) { q.xa = 0; q.xb = w; q.ya = 0; q.yb = h; return q; }
  if (valid && !valid[m]) { q.xa = q.xb = q.ya = q.yb = 0; return q; }
  const float4 b = __ldg(reinterpret_cast<const float4*>(boxes) + m);
  q.xa = max((int)floorf(b.x) - 2, 0); q.xb = min((int)ceilf(b.z) + 2, w);
  q.ya = max((int)floorf(b.y) - 2, 0); q.yb = min((int)ceilf(b.w) + 2, h);
  if (q.xb <= q.xa || q.yb <= q.ya) q.xa = q.xb = q.ya = q.yb = 0;
  return q;
}
// value of the pixel that precedes (x, ya) in column-major order, given that everything outside the window is zero
__device__ __forceinline__ uint8_t rle_prev(const uint8_t* __restrict__ p, const RleWindow& q, int x, int h, int w) {
  if (q.ya > 0) return 0;                                  // the pixel above lies outside the window
  if (x == 0 || x - 1 < q.xa || q.yb < h) return 0;        // last pixel of the previous column lies outside the window
  return p[(size_t)(h - 1) * w + x - 1] != 0;
}

__global__ void __launch_bounds__(256) rle_count_kernel(const uint8_t* __restrict__ masks, int h, int w, int* __restrict__ col_count,
                                                        const float* __restrict__ boxes, const uint8_t* __restrict__ valid) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (x >= w) return;
  const RleWindow q = rle_window(boxes, valid, m, h, w);
  int cnt = 0;
  if (x >= q.xa && x < q.xb) {
    const uint8_t* p = masks + (size_t)m * h * w;
    uint8_t prev = rle_prev(p, q, x, h, w);
    const int ye = min(q.yb + 1, h);                       // one row past the window: the closing transition of the column
    // the walk is a chain of dependent compares but independent loads: eight rows are fetched before the first compare, so that
    // a tall window costs one L2 round trip per eight rows instead of one per row (the kernel is bound by its tallest columns)
    int y = q.ya;
    const uint8_t* col = p + (size_t)y * w + x;
    for (; y + RLE_ROWS <= ye; y += RLE_ROWS, col += (size_t)RLE_ROWS * w) {
      uint8_t v[RLE_ROWS];
#pragma unroll
      for (int i = 0; i < RLE_ROWS; ++i) v[i] = col[(size_t)i * w];
#pragma unroll
      for (int i = 0; i < RLE_ROWS; ++i) {
        const uint8_t b = v[i] != 0;
        cnt += b != prev;
        prev = b;
      }
    }
    for (; y < ye; ++y, col += w) {
      const uint8_t v = *col != 0;
      cnt += v != prev;
      prev = v;
    }
  }
  col_count[(size_t)m * w + x] = cnt;
}

// one CTA per mask: exclusive scan of w column counts (w <= 1024 * RLE_SCAN_ITEMS)
constexpr int RLE_SCAN_ITEMS = 8;
__global__ void __launch_bounds__(1024) rle_scan_kernel(const int* __restrict__ col_count, int w, int* __restrict__ col_offset,
                                                        int* __restrict__ total) {
  __shared__ int s_warp[32];
  const int m = blockIdx.x;
  const int* c = col_count + (size_t)m * w;
  int* o = col_offset + (size_t)m * w;
  int v[RLE_SCAN_ITEMS];
  int sum = 0;
  const int base = threadIdx.x * RLE_SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < RLE_SCAN_ITEMS; ++i) {
    v[i] = base + i < w ? c[base + i] : 0;
    sum += v[i];
  }
  // inclusive scan of the per-thread sums: warp shuffle + one pass over the warp totals
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int ws = s_warp[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ws, d);
      if (lane >= d) ws += t;
    }
    s_warp[lane] = ws;
  }
  __syncthreads();
  int run = incl - sum + (warp ? s_warp[warp - 1] : 0);
#pragma unroll
  for (int i = 0; i < RLE_SCAN_ITEMS; ++i) {
    if (base + i < w) o[base + i] = run;
    run += v[i];
  }
  if (threadIdx.x == blockDim.x - 1) total[m] = s_warp[31];
}

// mask_offset[m] = sum_{i < m} (total[i] + 1), mask_offset[r] = number of runs of all masks: the device-side form of the
// host scan between cm2_rle_count and cm2_rle_write, so that a pipelined caller never waits for `total`.
__global__ void __launch_bounds__(1024) rle_offsets_kernel(const int* __restrict__ total, int r, long long* __restrict__ mask_offset) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < r; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = i < r ? (long long)total[i] + 1 : 0;
    long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      long long ws = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, ws, d);
        if (lane >= d) ws += t;
      }
      s_warp[lane] = ws;
    }
    __syncthreads();
    const long long carry = s_carry;
    const long long excl = carry + incl - v + (warp ? s_warp[warp - 1] : 0);
    if (i < r) mask_offset[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) mask_offset[r] = s_carry;
}

// runs[mask_offset[m] + k] for k = 0 .. total[m]: run k ends at transition k; the last run ends at h * w.
// Each column thread needs the position of the last transition before its column: the previous column's last
// transition is not known locally, so positions are written first and differenced in place by rle_diff_kernel.
__global__ void __launch_bounds__(256) rle_write_kernel(const uint8_t* __restrict__ masks, int h, int w, const int* __restrict__ col_offset,
                                                        const long long* __restrict__ mask_offset, unsigned* __restrict__ runs,
                                                        const int* __restrict__ total, long long capacity,
                                                        const float* __restrict__ boxes, const uint8_t* __restrict__ valid) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (x >= w) return;
  if (capacity >= 0 && mask_offset[m] + total[m] + 1 > capacity) return;      // does not fit: the caller sees mask_offset[r] > capacity
  const RleWindow q = rle_window(boxes, valid, m, h, w);
  if (x < q.xa || x >= q.xb) return;
  const uint8_t* p = masks + (size_t)m * h * w;
  unsigned* out = runs + mask_offset[m] + col_offset[(size_t)m * w + x];
  uint8_t prev = rle_prev(p, q, x, h, w);
  const int ye = min(q.yb + 1, h);
  int k = 0;
  int y = q.ya;
  const uint8_t* col = p + (size_t)y * w + x;
  for (; y + RLE_ROWS <= ye; y += RLE_ROWS, col += (size_t)RLE_ROWS * w) {
    uint8_t v[RLE_ROWS];
#pragma unroll
    for (int i = 0; i < RLE_ROWS; ++i) v[i] = col[(size_t)i * w];
#pragma unroll
    for (int i = 0; i < RLE_ROWS; ++i) {
      const uint8_t b = v[i] != 0;
      if (b != prev) out[k++] = (unsigned)(x * h + y + i);      // column-major position of the transition
      prev = b;
    }
  }
  for (; y < ye; ++y, col += w) {
    const uint8_t v = *col != 0;
    if (v != prev) out[k++] = (unsigned)(x * h + y);
    prev = v;
  }
}

// positions -> run lengths, one thread per entry (entry total[m] of every mask is the closing run)
__global__ void __launch_bounds__(256) rle_diff_kernel(const unsigned* __restrict__ pos, const long long* __restrict__ mask_offset,
                                                       const int* __restrict__ total, int h, int w, unsigned* __restrict__ runs_out,
                                                       long long capacity) {
  const int m = blockIdx.y;
  const int n = total[m];
  const long long base = mask_offset[m];
  if (capacity >= 0 && base + n + 1 > capacity) return;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= n; k += gridDim.x * blockDim.x) {
    const unsigned hi = k < n ? pos[base + k] : (unsigned)(h * w);
    const unsigned lo = k > 0 ? pos[base + k - 1] : 0u;
    runs_out[base + k] = hi - lo;
  }
}

}  // namespace cm2

using namespace cm2;

extern "C" int cm2_rle_count(const uint8_t* masks, int32_t r, int32_t h, int32_t w, int32_t* col_count, int32_t* col_offset,
                             int32_t* total, void* stream) {
  CM2_CHECK_ARG(masks && col_count && col_offset && total, "rle_count: null pointer");
  CM2_CHECK_ARG(r >= 0 && h > 0 && w > 0 && w <= 1024 * RLE_SCAN_ITEMS && r <= 65535 && (long long)h * w < (1ll << 31),
                "rle_count: bad extents r=%d %dx%d", r, h, w);
  if (r == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  rle_count_kernel<<<dim3(ceil_div(w, 256), r), 256, 0, s>>>(masks, h, w, col_count, nullptr, nullptr);
  CM2_CHECK_LAUNCH("rle_count");
  rle_scan_kernel<<<r, 1024, 0, s>>>(col_count, w, col_offset, total);
  CM2_CHECK_LAUNCH("rle_scan");
  return CM2_OK;
}

extern "C" int cm2_rle_write(const uint8_t* masks, int32_t r, int32_t h, int32_t w, const int32_t* col_offset, const int32_t* total,
                             const int64_t* mask_offset, uint32_t* positions, uint32_t* runs, void* stream) {
  CM2_CHECK_ARG(masks && col_offset && total && mask_offset && positions && runs && positions != runs, "rle_write: null / aliased pointer");
  CM2_CHECK_ARG(r >= 0 && h > 0 && w > 0 && r <= 65535 && (long long)h * w < (1ll << 31), "rle_write: bad extents r=%d %dx%d", r, h, w);
  if (r == 0) return CM2_OK;
  cudaStream_t s = (cudaStream_t)stream;
  rle_write_kernel<<<dim3(ceil_div(w, 256), r), 256, 0, s>>>(masks, h, w, col_offset, reinterpret_cast<const long long*>(mask_offset), positions,
                                                             total, -1, nullptr, nullptr);
  CM2_CHECK_LAUNCH("rle_write");
  rle_diff_kernel<<<dim3(64, r), 256, 0, s>>>(positions, reinterpret_cast<const long long*>(mask_offset), total, h, w, runs, -1);
  CM2_CHECK_LAUNCH("rle_diff");
  return CM2_OK;
}

extern "C" int cm2_rle_encode(const uint8_t* masks, int32_t r, int32_t h, int32_t w, int32_t* col_count, int32_t* col_offset,
                              int32_t* total, int64_t* mask_offset, uint32_t* positions, uint32_t* runs, int64_t capacity,
                              const float* boxes, const uint8_t* valid, void* stream) {
  CM2_CHECK_ARG(masks && col_count && col_offset && total && mask_offset && positions && runs && positions != runs,
                "rle_encode: null / aliased pointer");
  CM2_CHECK_ARG(r >= 0 && h > 0 && w > 0 && w <= 1024 * RLE_SCAN_ITEMS && r <= 65535 && (long long)h * w < (1ll << 31) && capacity >= 0,
                "rle_encode: bad extents r=%d %dx%d capacity=%lld", r, h, w, (long long)capacity);
  cudaStream_t s = (cudaStream_t)stream;
  if (r == 0) {
    if (cudaMemsetAsync(mask_offset, 0, sizeof(int64_t), s) != cudaSuccess) { set_error("rle_encode: memset failed"); return CM2_ERR_CUDA; }
    return CM2_OK;
  }
  CM2_CHECK_ARG(!boxes || (reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "rle_encode: boxes must be 16-byte aligned");
  rle_count_kernel<<<dim3(ceil_div(w, 256), r), 256, 0, s>>>(masks, h, w, col_count, boxes, valid);
  CM2_CHECK_LAUNCH("rle_count");
  rle_scan_kernel<<<r, 1024, 0, s>>>(col_count, w, col_offset, total);
  CM2_CHECK_LAUNCH("rle_scan");
  rle_offsets_kernel<<<1, 1024, 0, s>>>(total, r, reinterpret_cast<long long*>(mask_offset));
  CM2_CHECK_LAUNCH("rle_offsets");
  rle_write_kernel<<<dim3(ceil_div(w, 256), r), 256, 0, s>>>(masks, h, w, col_offset, reinterpret_cast<const long long*>(mask_offset), positions,
                                                             total, capacity, boxes, valid);
  CM2_CHECK_LAUNCH("rle_write");
  rle_diff_kernel<<<dim3(64, r), 256, 0, s>>>(positions, reinterpret_cast<const long long*>(mask_offset), total, h, w, runs, capacity);
  CM2_CHECK_LAUNCH("rle_diff");
  return CM2_OK;
}
