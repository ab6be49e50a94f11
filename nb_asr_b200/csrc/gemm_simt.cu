// Generic strided SIMT GEMM (fp32 accumulate) with the fused epilogue.
// This is the fp32-parity path (activations + weights in fp32, 1e-4 tolerance of north_star)
// and the path for small / oddly shaped products (K or N not TMA friendly).
// The bf16 hot path is gemm_sm100.cu (tcgen05 / TMEM / TMA).
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename TA, typename TB>
__global__ void __launch_bounds__(NT) simt_gemm_kernel(SimtGemmArgs p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float Cs[BM][BN + 1];
  const TA* __restrict__ A = reinterpret_cast<const TA*>(p.a);
  const TB* __restrict__ Bm = reinterpret_cast<const TB*>(p.b);
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * BN;
  const int i0 = blockIdx.y * BM;
  const int ib = blockIdx.z;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kfast = (p.a_kr == 1);
  const bool b_kfast = (p.b_kr == 1);
  for (int kb = 0; kb < p.nkb; ++kb) {
    for (int k0 = 0; k0 < p.nkr; k0 += BK) {
#pragma unroll
      for (int q = 0; q < (BM * BK) / NT; ++q) {
        int idx = tid + q * NT;
        int k = a_kfast ? (idx % BK) : (idx / BM);
        int i = a_kfast ? (idx / BK) : (idx % BM);
        float v = 0.f;
        if (i0 + i < p.nir && k0 + k < p.nkr)
          v = static_cast<float>(A[ib * p.a_ib + (int64_t)(i0 + i) * p.a_ir + kb * p.a_kb + (int64_t)(k0 + k) * p.a_kr]);
        As[k][i] = v;
      }
#pragma unroll
      for (int q = 0; q < (BN * BK) / NT; ++q) {
        int idx = tid + q * NT;
        int k = b_kfast ? (idx % BK) : (idx / BN);
        int j = b_kfast ? (idx / BK) : (idx % BN);
        float v = 0.f;
        if (j0 + j < p.N && k0 + k < p.nkr)
          v = static_cast<float>(Bm[(int64_t)(j0 + j) * p.b_j + kb * p.b_kb + (int64_t)(k0 + k) * p.b_kr]);
        Bs[k][j] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Cs[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();
  if (tid < BM * (BN / 32)) {
    int r = tid / (BN / 32), ch = tid % (BN / 32);
    int i = i0 + r, c0 = j0 + ch * 32;
    if (i < p.nir && c0 < p.N) {
      float v[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) v[x] = Cs[r][ch * 32 + x];
      int64_t rho = p.o_r0 + (int64_t)ib * p.o_bs + (int64_t)i * p.o_rs;
      epilogue_chunk(p.epi, rho, c0, p.N, v);
    }
  }
}

}  // namespace

int simt_gemm_launch(const SimtGemmArgs& a, cudaStream_t st) {
  dim3 grid((a.N + BN - 1) / BN, (a.nir + BM - 1) / BM, a.nib);
  if (a.a_dtype == NBASR_F32 && a.b_dtype == NBASR_F32) simt_gemm_kernel<float, float><<<grid, NT, 0, st>>>(a);
  else if (a.a_dtype == NBASR_BF16 && a.b_dtype == NBASR_F32) simt_gemm_kernel<bf16, float><<<grid, NT, 0, st>>>(a);
  else if (a.a_dtype == NBASR_BF16 && a.b_dtype == NBASR_BF16) simt_gemm_kernel<bf16, bf16><<<grid, NT, 0, st>>>(a);
  else simt_gemm_kernel<float, bf16><<<grid, NT, 0, st>>>(a);
  NBASR_CHECK_LAUNCH();
  return 0;
}
