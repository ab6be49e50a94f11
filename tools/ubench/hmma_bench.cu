// Micro-benchmark: legacy mma.sync (HMMA) bf16 m16n8k16 issue rate per SM on sm_100a, alone and fed by ldmatrix.x4
// from 128B-swizzled shared memory (the access pattern of the grouped-conv kernel).  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t* a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

template <int MODE>
__global__ void __launch_bounds__(416, 2) bench(float* out, int iters) {
  extern __shared__ __align__(1024) uint8_t sm[];
  for (int i = threadIdx.x; i < 144 * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  float acc[4][4] = {};
  uint32_t b[3][2] = {{0x3c003c00u, 0x3c003c01u}, {0x3c013c00u, 0x3c003c02u}, {0x3c003c00u, 0x3c003c03u}};
  uint32_t a[4] = {0x3c003c00u + lane, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
  const int chunk = warp % 6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int row = (lane & 7) + 8 * ((lane >> 3) & 1) + ((lane >> 4) + 2 * i);
      const uint32_t ad = base + row * 128 + ((chunk ^ (row & 7)) << 4);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (MODE == 1) ldsm4(ad + mt * 2048 + ((warp / 6) * 8192), a);
        mma16816(acc[mt], a, b[i]);
      }
    }
  }
  float s = 0;
  for (int mt = 0; mt < 4; ++mt) for (int i = 0; i < 4; ++i) s += acc[mt][i];
  if (s == 123.456f) out[0] = s;
}

int main() {
  float* out; cudaMalloc(&out, 4);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) for (int warps : {4, 8, 12}) for (int cps : {1, 2}) {
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) bench<0><<<pr.multiProcessorCount * cps, warps * 32, 20480>>>(out, iters);
      else bench<1><<<pr.multiProcessorCount * cps, warps * 32, 20480>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas_per_sm = (double)iters * 12 * warps * cps;
    double cyc = ms * 1e-3 * clk_khz * 1e3;
    printf("mode %s warps/CTA %2d CTAs/SM %d: %.3f ms, %.2f cyc/mma/SM (at %d MHz nominal), %.1f dense TFLOP/s chip\n",
           mode ? "ldmatrix+mma" : "mma only", warps, cps, ms, cyc / mmas_per_sm, clk_khz / 1000,
           mmas_per_sm * pr.multiProcessorCount * 4096.0 / (ms * 1e-3) / 1e12);
  }
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
