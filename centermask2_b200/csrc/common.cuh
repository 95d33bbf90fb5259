// Shared helpers for libcm2 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/cm2.h"

namespace cm2 {

void set_error(const char* fmt, ...);

#define CM2_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      cm2::set_error(__VA_ARGS__);          \
      return CM2_ERR_BAD_SHAPE;             \
    }                                       \
  } while (0)

#define CM2_CHECK_LAUNCH(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      cm2::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return CM2_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute (MODEL.DEVICE=cuda:N selects the device per cfg
// within one process): remember the largest value set on every device, and fail loudly when the driver refuses it.
constexpr int CM2_MAX_DEVICES = 64;
#define CM2_ENSURE_DYN_SMEM(kernel, nbytes, name)                                                                  \
  do {                                                                                                             \
    static int set__[cm2::CM2_MAX_DEVICES];                                                                        \
    int dev__ = 0;                                                                                                 \
    if (cudaGetDevice(&dev__) != cudaSuccess || dev__ < 0 || dev__ >= cm2::CM2_MAX_DEVICES) {                      \
      cm2::set_error("%s: cudaGetDevice failed", name);                                                            \
      return CM2_ERR_CUDA;                                                                                         \
    }                                                                                                              \
    if (set__[dev__] < (int)(nbytes)) {                                                                            \
      cudaError_t e__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(nbytes));  \
      if (e__ != cudaSuccess) {                                                                                    \
        cm2::set_error("%s: cudaFuncSetAttribute(%d bytes of dynamic shared memory) failed: %s", name,             \
                       (int)(nbytes), cudaGetErrorString(e__));                                                    \
        return CM2_ERR_CUDA;                                                                                       \
      }                                                                                                            \
      set__[dev__] = (int)(nbytes);                                                                                \
    }                                                                                                              \
  } while (0)

// Programmatic dependent launch (convolution kernels only, csrc/conv_tc.cu): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may be scheduled while its predecessor in the stream is still draining;
// after its setup it lets the NEXT kernel be scheduled (launch_dependents) and blocks until the predecessor has completed and
// its writes are visible (wait) -- nothing global is read or written before that.  Measured (same box, batch 16): 11.99 ->
// 11.73 ms per step with the convolutions alone; giving the bandwidth-bound kernels between them the same treatment was 4 %
// SLOWER than no PDL at all (12.06 vs 11.60 ms: convolution CTAs that become resident early take registers and shared memory
// from the streaming kernel that is still running), so those are launched plainly.  CM2_PDL=0 switches it off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();                                      // api.cu

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive channels as fp32, from either dtype (pointer must be 16B (bf16) / 32B (f32) aligned).
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// sigmoid with the same arithmetic as ATen's CPU/CUDA kernel: 1 / (1 + exp(-x)) in fp32.
__device__ __forceinline__ float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }

// device-side copy of a cm2_act
template <typename T>
struct View {
  T* p;
  int n, h, w, c;
  long long sn, sh, sw;
  __device__ __forceinline__ T* at(int b, int y, int x) const { return p + b * sn + y * sh + x * sw; }
};
template <typename T>
inline View<T> make_view(const cm2_act& a) {
  View<T> v;
  v.p = reinterpret_cast<T*>(a.data);
  v.n = a.n; v.h = a.h; v.w = a.w; v.c = a.c;
  v.sn = a.sn; v.sh = a.sh; v.sw = a.sw;
  return v;
}
inline bool same_extent(const cm2_act& a, const cm2_act& b) {
  return a.n == b.n && a.h == b.h && a.w == b.w && a.c == b.c;
}
// vector (8-channel) access needs every stride and the base address to be a multiple of 8 elements
inline bool vec8_ok(const cm2_act& a, int elem_bytes) {
  return a.c % 8 == 0 && a.sn % 8 == 0 && a.sh % 8 == 0 && a.sw % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a.data) % (8 * elem_bytes)) == 0;
}
inline int elem_bytes(int dtype) { return dtype == CM2_F32 ? 4 : (dtype == CM2_BF16 ? 2 : 1); }

void conv_out_extent(const cm2_conv_desc* d, int* ho, int* wo);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace cm2
