// Throughput of the legacy warp-level tensor-core path (mma.sync.m16n8k16 bf16 -> fp32) and of ldmatrix.x4.trans on
// sm_100a, per SM -- input for the planned tensor-core x-pass of ROIAlign (DESIGN.md "ROIAlign, next step").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mma_rate tools/micro/mma_rate.cu && tools/micro/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void mma_kernel(float* out, int iters, int ilp_dummy) {
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
  float c[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)                           // 8 independent accumulators
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  if (s == 12345.f && ilp_dummy) out[0] = s;
}

__global__ void ldsm_kernel(float* out, int iters, int pitch_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // matrix row r (0..31 over the four 8x8 matrices) = lane: rows pitch_bytes apart (512: unpadded pixels, 528: padded)
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem) + (lane & 15) * pitch_bytes + (lane >> 4) * 16 + warp * 64;
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    uint32_t r0, r1, r2, r3;
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
    acc += r0 ^ r1 ^ r2 ^ r3;
    addr ^= 32;
  }
  if (acc == 0x12345678u) out[0] = 1.f;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      mma_kernel<<<p.multiProcessorCount, 32 * warps>>>(out, iters, 0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)p.multiProcessorCount * warps * iters * 8;
    printf("mma.sync m16n8k16 bf16: %2d warps/SM  %.3f ms  %.1f TFLOP/s  %.2f clk/MMA/SM at %.0f MHz nominal\n", warps, ms,
           mmas * 4096 / ms / 1e9, ms * 1e-3 * p.clockRate * 1e3 / ((double)warps * iters * 8), p.clockRate / 1e3);
  }
  cudaFuncSetAttribute(ldsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int pitch : {512, 528}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      ldsm_kernel<<<p.multiProcessorCount, 256, 65536>>>(out, iters, pitch);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("ldmatrix.x4.trans, row pitch %d B: %.3f ms  %.2f clk per instruction per SM (8 warps)\n", pitch, ms,
           ms * 1e-3 * p.clockRate * 1e3 / ((double)8 * iters));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
