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

__global__ void __launch_bounds__(256) rle_count_kernel(const uint8_t* __restrict__ masks, int h, int w, int* __restrict__ col_count) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (x >= w) return;
  const uint8_t* p = masks + (size_t)m * h * w;
  uint8_t prev = x == 0 ? 0 : (p[(size_t)(h - 1) * w + x - 1] != 0);
  int cnt = 0;
  for (int y = 0; y < h; ++y) {
    const uint8_t v = p[(size_t)y * w + x] != 0;
    cnt += v != prev;
    prev = v;
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

// runs[mask_offset[m] + k] for k = 0 .. total[m]: run k ends at transition k; the last run ends at h * w.
// Each column thread needs the position of the last transition before its column: the previous column's last
// transition is not known locally, so positions are written first and differenced in place by rle_diff_kernel.
__global__ void __launch_bounds__(256) rle_write_kernel(const uint8_t* __restrict__ masks, int h, int w, const int* __restrict__ col_offset,
                                                        const long long* __restrict__ mask_offset, unsigned* __restrict__ runs) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (x >= w) return;
  const uint8_t* p = masks + (size_t)m * h * w;
  unsigned* out = runs + mask_offset[m] + col_offset[(size_t)m * w + x];
  uint8_t prev = x == 0 ? 0 : (p[(size_t)(h - 1) * w + x - 1] != 0);
  int k = 0;
  for (int y = 0; y < h; ++y) {
    const uint8_t v = p[(size_t)y * w + x] != 0;
    if (v != prev) out[k++] = (unsigned)(x * h + y);      // column-major position of the transition
    prev = v;
  }
}

// positions -> run lengths, one thread per entry (entry total[m] of every mask is the closing run)
__global__ void __launch_bounds__(256) rle_diff_kernel(const unsigned* __restrict__ pos, const long long* __restrict__ mask_offset,
                                                       const int* __restrict__ total, int h, int w, unsigned* __restrict__ runs_out) {
  const int m = blockIdx.y;
  const int n = total[m];
  const long long base = mask_offset[m];
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
  rle_count_kernel<<<dim3(ceil_div(w, 256), r), 256, 0, s>>>(masks, h, w, col_count);
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
  rle_write_kernel<<<dim3(ceil_div(w, 256), r), 256, 0, s>>>(masks, h, w, col_offset, reinterpret_cast<const long long*>(mask_offset), positions);
  CM2_CHECK_LAUNCH("rle_write");
  rle_diff_kernel<<<dim3(64, r), 256, 0, s>>>(positions, reinterpret_cast<const long long*>(mask_offset), total, h, w, runs);
  CM2_CHECK_LAUNCH("rle_diff");
  return CM2_OK;
}
