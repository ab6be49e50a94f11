// HBM-bound row kernels: LayerNorm fwd/bwd, element-wise epilogue pass, column sums, input
// transpose, weight packing.  One warp per frame row for the norm kernels (a row of C <= 1280
// channels lives in registers), 16-byte vector accesses everywhere.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int LN_MAXG = 5;  // groups of 8 channels per lane -> C <= 1280

template <typename T>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int Tt,
                                                            int Tp, int C, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            float* __restrict__ mean_o, float* __restrict__ rstd_o) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int ngroups = C >> 3;
  const int64_t nrows = (int64_t)B * Tt;
  for (int64_t r = warp; r < nrows; r += nwarps) {
    int b = (int)(r / Tt), t = (int)(r % Tt);
    int64_t rho = (int64_t)b * Tp + NBASR_PAD_L + t;
    const T* xr = x + rho * C;
    float v[LN_MAXG][8];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < LN_MAXG; ++q) {
      int g = lane + 32 * q;
      if (g < ngroups) {
        load8(xr + g * 8, v[q]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[q][i];
      }
    }
    float mean = warp_sum(s) / C;
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < LN_MAXG; ++q) {
      int g = lane + 32 * q;
      if (g < ngroups) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float d = v[q][i] - mean;
          ss += d * d;
        }
      }
    }
    float rstd = rsqrtf(warp_sum(ss) / C + eps);
    if (lane == 0 && mean_o) {
      mean_o[rho] = mean;
      rstd_o[rho] = rstd;
    }
    T* yr = y + rho * C;
#pragma unroll
    for (int q = 0; q < LN_MAXG; ++q) {
      int g = lane + 32 * q;
      if (g < ngroups) {
        float ga[8], be[8], o[8];
        load8(gamma + g * 8, ga);
        load8(beta + g * 8, be);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (v[q][i] - mean) * rstd * ga[i] + be[i];
        store8(yr + g * 8, o);
      }
    }
  }
}

// Backward.  Each warp owns a private multi-stage ring in shared memory that 1-D bulk copies (TMA,
// cp.async.bulk + mbarrier) fill with whole (x, dy) rows several rows ahead: memory-level parallelism comes
// from the copy engine instead of from warp count, rows are read from smem twice (statistics, then output)
// and the registers hold the dgamma / dbeta partials of the lane's channels.
__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ln_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void ln_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LN_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra LN_DONE;\n"
      "bra LN_WAIT;\n"
      "LN_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

template <typename T, int NST>
__global__ void __launch_bounds__(256, 1) layernorm_bwd_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean_i,
    const float* __restrict__ rstd_i, const float* __restrict__ gamma, int B, int Tt, int Tp, int C,
    T* __restrict__ dx, T* __restrict__ dx2, const uint32_t* __restrict__ mask2, float scale2, int64_t mask_rows, int mask2_w,
    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ __align__(128) uint8_t lnsm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int ngroups = C >> 3;
  const int64_t nrows = (int64_t)B * Tt;
  const uint32_t row_bytes = (uint32_t)C * sizeof(T);
  const uint32_t stage_bytes = 2 * row_bytes;                      // x row then dy row
  uint8_t* ring = lnsm + (size_t)wib * NST * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lnsm + (size_t)8 * NST * stage_bytes) + wib * NST;
  float* red = reinterpret_cast<float*>(lnsm + (size_t)8 * NST * stage_bytes + 8 * NST * 8);   // 2*C floats, block reduction
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  if (lane == 0) {
    for (int s = 0; s < NST; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ln_smem_u32(bars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto row_of = [&](int64_t r) { return (int64_t)(r / Tt) * Tp + NBASR_PAD_L + (r % Tt); };
  auto issue = [&](int64_t r, int s) {          // lane 0 only
    const uint32_t bar = ln_smem_u32(bars + s);
    const int64_t rho = row_of(r);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
    ln_bulk_load(ln_smem_u32(ring + (size_t)s * stage_bytes), x + rho * C, row_bytes, bar);
    ln_bulk_load(ln_smem_u32(ring + (size_t)s * stage_bytes + row_bytes), dy + rho * C, row_bytes, bar);
  };
  float dg[LN_MAXG][8], db[LN_MAXG][8], ga[LN_MAXG][8];
#pragma unroll
  for (int q = 0; q < LN_MAXG; ++q) {
    int g = lane + 32 * q;
#pragma unroll
    for (int i = 0; i < 8; ++i) dg[q][i] = db[q][i] = ga[q][i] = 0.f;
    if (g < ngroups) load8(gamma + g * 8, ga[q]);
  }
  if (lane == 0)
    for (int s = 0; s < NST; ++s) {
      int64_t r = warp + (int64_t)s * nwarps;
      if (r < nrows) issue(r, s);
    }
  int st = 0;
  uint32_t ph = 0;
  for (int64_t r = warp; r < nrows; r += nwarps) {
    const int64_t rho = row_of(r);
    const float mean = mean_i[rho], rstd = rstd_i[rho];
    ln_mbar_wait(ln_smem_u32(bars + st), ph);
    const T* xs = reinterpret_cast<const T*>(ring + (size_t)st * stage_bytes);
    const T* ds = reinterpret_cast<const T*>(ring + (size_t)st * stage_bytes + row_bytes);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < LN_MAXG; ++q) {
      int g = lane + 32 * q;
      if (g < ngroups) {
        float xv[8], dv[8];
        load8(xs + g * 8, xv);
        load8(ds + g * 8, dv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float xh = (xv[i] - mean) * rstd;
          dg[q][i] += dv[i] * xh;
          db[q][i] += dv[i];
          float gy = dv[i] * ga[q][i];
          s1 += gy;
          s2 += gy * xh;
        }
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int q = 0; q < LN_MAXG; ++q) {
      int g = lane + 32 * q;
      if (g < ngroups) {
        float xv[8], dv[8], o[8];
        load8(xs + g * 8, xv);
        load8(ds + g * 8, dv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float xh = (xv[i] - mean) * rstd;
          o[i] = rstd * (dv[i] * ga[q][i] - s1 - xh * s2);
        }
        if (dx) store8(dx + rho * C + g * 8, o);
        if (dx2) {
          uint32_t w = mask2 ? reinterpret_cast<const uint8_t*>(mask2)[mask_byte_addr(rho, g * 8, mask2_w, mask_rows)] : 0xffu;
          float o2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o2[i] = ((w >> i) & 1u) ? o[i] * scale2 : 0.f;
          store8(dx2 + rho * C + g * 8, o2);
        }
      }
    }
    // the stage is consumed: refill it with the row NST iterations ahead (generic reads -> async write)
    __syncwarp();
    const int64_t rn = r + (int64_t)NST * nwarps;
    if (lane == 0 && rn < nrows) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(rn, st);
    }
    if (++st == NST) { st = 0; ph ^= 1; }
  }
#pragma unroll
  for (int q = 0; q < LN_MAXG; ++q) {
    int g = lane + 32 * q;
    if (g < ngroups) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(&red[g * 8 + i], dg[q][i]);
        atomicAdd(&red[C + g * 8 + i], db[q][i]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
  }
}

__global__ void eltwise_kernel(int src_dtype, const void* __restrict__ src, int64_t ld_src, int B, int Tt, int Tp, int C,
                               nbasr_epilogue e) {
  const int nch = (C + 31) >> 5;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * Tt * nch;
  if (idx >= total) return;
  int ch = (int)(idx % nch);
  int64_t r = idx / nch;
  int b = (int)(r / Tt), t = (int)(r % Tt);
  int64_t rho = (int64_t)b * Tp + NBASR_PAD_L + t;
  int c0 = ch * 32;
  float v[32];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    int nrem = C - c0 - g * 8;
    if (src && nrem > 0) load8_dt_n(src, src_dtype, rho * ld_src + c0 + g * 8, v + g * 8, nrem);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[g * 8 + i] = 0.f;
    }
  }
  epilogue_chunk(e, rho, c0, C, v);
}

constexpr int CS_MAXG = 8;  // C <= 2048

// column sums: warp per row (16-byte vector loads, fully coalesced), register partials, block reduce, atomics
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int B, int Tt, int Tp, int C, float* __restrict__ out) {
  extern __shared__ float red[];  // C floats
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int ngroups = C >> 3;
  const int64_t nrows = (int64_t)B * Tt;
  for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[CS_MAXG][8];
#pragma unroll
  for (int q = 0; q < CS_MAXG; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
  for (int64_t r = warp; r < nrows; r += nwarps) {
    int b = (int)(r / Tt), t = (int)(r % Tt);
    const T* xr = x + ((int64_t)b * Tp + NBASR_PAD_L + t) * C;
#pragma unroll
    for (int q = 0; q < CS_MAXG; ++q) {
      int g = lane + 32 * q;
      if (g < ngroups) {
        float v[8];
        load8(xr + g * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[q][i] += v[i];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < CS_MAXG; ++q) {
    int g = lane + 32 * q;
    if (g < ngroups) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&red[g * 8 + i], acc[q][i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, red[i]);
}

template <typename T>
__global__ void transpose_in_kernel(const float* __restrict__ a, T* __restrict__ out, int B, int F, int Tt, int Tp) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int f = f0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (f < F && t < Tt) ? a[((int64_t)b * F + f) * Tt + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, f = f0 + threadIdx.x;
    if (t < Tt && f < F) out[((int64_t)b * Tp + NBASR_PAD_L + t) * F + f] = static_cast<T>(tile[threadIdx.x][i]);
  }
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int M, int N, int nq, int t0,
                                   int tstep, int64_t ws_m, int64_t ws_n, int64_t ws_t) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)N * nq * M;
  if (idx >= total) return;
  int m = (int)(idx % M);
  int q = (int)((idx / M) % nq);
  int n = (int)(idx / ((int64_t)M * nq));
  out[idx] = static_cast<T>(w[m * ws_m + n * ws_n + (int64_t)(t0 + q * tstep) * ws_t]);
}

template <typename T>
__global__ void convert_kernel(const float* __restrict__ s, T* __restrict__ d, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = static_cast<T>(s[i]);
}

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

extern "C" {

int nbasr_layernorm_fwd(int dtype, const void* x, void* y, int B, int T, int Tp, int C, const float* gamma,
                        const float* beta, float eps, float* mean, float* rstd, void* stream) {
  NBASR_REQUIRE(C % 8 == 0 && C <= 8 * 32 * LN_MAXG, "C");
  int64_t rows = (int64_t)B * T;
  int blocks = (int)std::min<int64_t>((rows + 7) / 8, 148 * 8);
  if (blocks < 1) return 0;
  if (dtype == NBASR_BF16)
    layernorm_fwd_kernel<bf16><<<blocks, 256, 0, as_stream(stream)>>>((const bf16*)x, (bf16*)y, B, T, Tp, C, gamma, beta, eps, mean, rstd);
  else
    layernorm_fwd_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>((const float*)x, (float*)y, B, T, Tp, C, gamma, beta, eps, mean, rstd);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_layernorm_bwd(int dtype, const void* dy, const void* x, const float* mean, const float* rstd,
                        const float* gamma, int B, int T, int Tp, int C, void* dx, void* dx2, const uint32_t* mask2,
                        float scale2, int64_t mask_rows, int mask2_w, float* dgamma, float* dbeta, void* stream) {
  NBASR_REQUIRE(C % 8 == 0 && C <= 8 * 32 * LN_MAXG, "C");
  int64_t rows = (int64_t)B * T;
  int blocks = (int)std::min<int64_t>((rows + 7) / 8, nbasr_sm_count());
  if (blocks < 1) return 0;
  const int nst_bf16 = 3, nst_f32 = 2;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(layernorm_bwd_kernel<bf16, nst_bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(layernorm_bwd_kernel<float, nst_f32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr = true;
  }
  if (dtype == NBASR_BF16) {
    size_t sm = (size_t)8 * nst_bf16 * 2 * C * 2 + 8 * nst_bf16 * 8 + (size_t)2 * C * 4;
    layernorm_bwd_kernel<bf16, nst_bf16><<<blocks, 256, sm, as_stream(stream)>>>((const bf16*)dy, (const bf16*)x, mean, rstd, gamma, B, T, Tp, C, (bf16*)dx, (bf16*)dx2, mask2, scale2, mask_rows, mask2_w, dgamma, dbeta);
  } else {
    size_t sm = (size_t)8 * nst_f32 * 2 * C * 4 + 8 * nst_f32 * 8 + (size_t)2 * C * 4;
    layernorm_bwd_kernel<float, nst_f32><<<blocks, 256, sm, as_stream(stream)>>>((const float*)dy, (const float*)x, mean, rstd, gamma, B, T, Tp, C, (float*)dx, (float*)dx2, mask2, scale2, mask_rows, mask2_w, dgamma, dbeta);
  }
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_eltwise(int src_dtype, const void* src, int64_t ld_src, int B, int T, int Tp, int C,
                  const nbasr_epilogue* epi, void* stream) {
  int64_t total = (int64_t)B * T * ((C + 31) / 32);
  if (total == 0) return 0;
  eltwise_kernel<<<(unsigned)((total + 127) / 128), 128, 0, as_stream(stream)>>>(src_dtype, src, ld_src, B, T, Tp, C, *epi);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_colsum(int dtype, const void* x, int B, int T, int Tp, int C, float* out, void* stream) {
  NBASR_REQUIRE(C % 8 == 0 && C <= 8 * 32 * CS_MAXG, "C");
  int64_t rows = (int64_t)B * T;
  if (rows < 1) return 0;
  int grid = (int)std::min<int64_t>((rows + 7) / 8, 148 * 4);
  size_t sm = C * sizeof(float);
  if (dtype == NBASR_BF16) colsum_kernel<bf16><<<grid, 256, sm, as_stream(stream)>>>((const bf16*)x, B, T, Tp, C, out);
  else colsum_kernel<float><<<grid, 256, sm, as_stream(stream)>>>((const float*)x, B, T, Tp, C, out);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_transpose_in(const float* audio, void* out, int dtype, int B, int F, int T, int Tp, void* stream) {
  dim3 grid((T + 31) / 32, (F + 31) / 32, B), block(32, 8);
  if (dtype == NBASR_BF16) transpose_in_kernel<bf16><<<grid, block, 0, as_stream(stream)>>>(audio, (bf16*)out, B, F, T, Tp);
  else transpose_in_kernel<float><<<grid, block, 0, as_stream(stream)>>>(audio, (float*)out, B, F, T, Tp);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_pack_weight(const float* w, void* out, int out_dtype, int M, int N, int nq, int t0, int tstep, int64_t ws_m,
                      int64_t ws_n, int64_t ws_t, void* stream) {
  int64_t total = (int64_t)N * nq * M;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (out_dtype == NBASR_BF16)
    pack_weight_kernel<bf16><<<blocks, 256, 0, as_stream(stream)>>>(w, (bf16*)out, M, N, nq, t0, tstep, ws_m, ws_n, ws_t);
  else
    pack_weight_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(w, (float*)out, M, N, nq, t0, tstep, ws_m, ws_n, ws_t);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_convert(const float* src, void* dst, int dst_dtype, int64_t n, void* stream) {
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (dst_dtype == NBASR_BF16) convert_kernel<bf16><<<blocks, 256, 0, as_stream(stream)>>>(src, (bf16*)dst, n);
  else convert_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(src, (float*)dst, n);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_fill_u32(uint32_t* p, uint32_t val, int64_t n, void* stream) {
  fill_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(p, val, n);
  NBASR_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
