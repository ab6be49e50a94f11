// Shared device helpers for libnbasr (sm_100a). See include/nbasr.h for the ABI.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "../../include/nbasr.h"

typedef __nv_bfloat16 bf16;
typedef __half f16;

// 16-bit pair conversions (round to nearest; fp16 saturates to +-65504 instead of overflowing to inf)
__device__ __forceinline__ uint32_t f2_to_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t f2_to_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 f16x2_to_f2(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

extern thread_local char g_nbasr_err[512];
int nbasr_fail(const char* fmt, ...);

#define NBASR_CHECK_LAUNCH()                                                     \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) return nbasr_fail("%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)
#define NBASR_REQUIRE(cond, msg)                                                 \
  do {                                                                           \
    if (!(cond)) return nbasr_fail("%s:%d requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------------
// 8-wide vector load/store with dtype conversion (all row pitches / chunk offsets are 8-aligned)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8(const float* p, float* v) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float* v) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const f16* p, float* v) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t* h = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = f16x2_to_f2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(f16* p, const float* v) {
  uint4 r;
  uint32_t* h = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = f2_to_f16x2(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ void store8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float* v) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ void load8_dt(const void* base, int dtype, int64_t idx, float* v) {
  if (dtype == NBASR_BF16) load8(reinterpret_cast<const bf16*>(base) + idx, v);
  else if (dtype == NBASR_F16) load8(reinterpret_cast<const f16*>(base) + idx, v);
  else load8(reinterpret_cast<const float*>(base) + idx, v);
}
__device__ __forceinline__ void store8_dt(void* base, int dtype, int64_t idx, const float* v) {
  if (dtype == NBASR_BF16) store8(reinterpret_cast<bf16*>(base) + idx, v);
  else if (dtype == NBASR_F16) store8(reinterpret_cast<f16*>(base) + idx, v);
  else store8(reinterpret_cast<float*>(base) + idx, v);
}
// 16 bytes of a 16-bit tensor (dtype BF16 or F16) -> 8 values
__device__ __forceinline__ void load8_h(const void* p, int dtype, float* v) {
  if (dtype == NBASR_F16) load8(reinterpret_cast<const f16*>(p), v);
  else load8(reinterpret_cast<const bf16*>(p), v);
}
// 8 values -> 16 bytes of a 16-bit tensor (dtype BF16 or F16) at a shared / global address
__device__ __forceinline__ void store8_h(void* p, int dtype, const float* v) {
  if (dtype == NBASR_F16) store8(reinterpret_cast<f16*>(p), v);
  else store8(reinterpret_cast<bf16*>(p), v);
}
__device__ __forceinline__ float ld1_dt(const void* base, int dtype, int64_t idx) {
  return dtype == NBASR_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(base)[idx])
       : dtype == NBASR_F16  ? __half2float(reinterpret_cast<const f16*>(base)[idx])
                             : reinterpret_cast<const float*>(base)[idx];
}
__device__ __forceinline__ void st1_dt(void* base, int dtype, int64_t idx, float v) {
  if (dtype == NBASR_BF16) reinterpret_cast<bf16*>(base)[idx] = __float2bfloat16(v);
  else if (dtype == NBASR_F16) reinterpret_cast<f16*>(base)[idx] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  else reinterpret_cast<float*>(base)[idx] = v;
}
// tail-safe variants: nrem = number of valid elements in this group of 8 (may be < 8)
__device__ __forceinline__ void load8_dt_n(const void* base, int dtype, int64_t idx, float* v, int nrem) {
  if (nrem >= 8) { load8_dt(base, dtype, idx, v); return; }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (i < nrem) ? ld1_dt(base, dtype, idx + i) : 0.f;
}
__device__ __forceinline__ void store8_dt_n(void* base, int dtype, int64_t idx, const float* v, int nrem) {
  if (nrem >= 8) { store8_dt(base, dtype, idx, v); return; }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nrem) st1_dt(base, dtype, idx + i, v[i]);
}
__device__ __forceinline__ float ld_dt(const void* base, int dtype, int64_t idx) { return ld1_dt(base, dtype, idx); }

// epilogue scale fields: 0 means "unset" (see nbasr.h)
__device__ __forceinline__ float epi_acc_scale(const nbasr_epilogue& e) { return e.acc_scale != 0.f ? e.acc_scale : 1.f; }
__device__ __forceinline__ float epi_bias_scale(const nbasr_epilogue& e) { return e.bias_scale != 0.f ? e.bias_scale : 1.f; }
__device__ __forceinline__ float epi_relu_hi(const nbasr_epilogue& e) { return e.relu_hi != 0.f ? e.relu_hi : 20.f; }

// plane-major gate-bit masks (see nbasr.h): byte address of (row, column group starting at col, col % 8 == 0)
__device__ __forceinline__ int64_t mask_byte_addr(int64_t rho, int col, int w, int64_t rows) {
  const int plane = col / w;
  const int eb = (w == 32) ? 4 : 8;
  return ((int64_t)plane * rows + rho) * eb + ((col - plane * w) >> 3);
}

// Counter-based RNG for dropout: one 32-bit hash per element (seed, element index).
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
  uint64_t z = idx * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<uint32_t>(z >> 32);
}

// ---------------------------------------------------------------------------------------------
// Fused output stage on one row and NV (multiple of 8) consecutive columns starting at c0 (multiple
// of 8).  v[NV] holds the accumulator values; nvalid = number of valid columns (<= NV).  Gate-bit masks
// are plane-major bit arrays (mask_byte_addr): any 8-aligned column range is whole bytes of one entry.
// ---------------------------------------------------------------------------------------------
// compute half: bias, ReLU20 (+ gate bits), dropout, skip-sum.  m[g] = gate bits of column group g.
// NOADD: the caller sums the skip tensors itself (the GEMM kernel stages them through shared memory with TMA).
template <int NV, bool FULL = false, bool NOBIAS = false, bool NOADD = false>
__device__ __forceinline__ void epilogue_compute(const nbasr_epilogue& e, int64_t rho, int c0, int nvalid_, float* v, uint32_t* m) {
  constexpr int NG = NV / 8;
  const int nvalid = FULL ? NV : nvalid_;   // FULL: every column valid -> all guards fold at compile time
#pragma unroll
  for (int g = 0; g < NG; ++g) m[g] = 0xffu;
  if (!NOBIAS) {      // (NOBIAS: the caller has already applied acc_scale and the bias)
    const float as = epi_acc_scale(e);
    if (as != 1.f) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] *= as;
    }
    if (e.bias) {
      const float bs = epi_bias_scale(e);
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (i < nvalid) v[i] = fmaf(__ldg(e.bias + c0 + i), bs, v[i]);
    }
  }
  if (e.relu20) {
    const float hi = epi_relu_hi(e);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      uint32_t mm = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float z = v[g * 8 + i];
        if (z > 0.f && z <= hi) mm |= (1u << i);
        v[g * 8 + i] = fminf(fmaxf(z, 0.f), hi);
      }
      m[g] = mm;
    }
  }
  if (e.drop_p > 0.f) {
    const float scale = 1.f / (1.f - e.drop_p);
    const uint32_t thr = static_cast<uint32_t>(e.drop_p * 4294967296.0);
    const uint64_t seed = e.drop_seed + (e.drop_step ? __ldg(e.drop_step) * 0xD1B54A32D192ED03ull : 0ull);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint32_t h = hash_u32(seed, static_cast<uint64_t>(rho) * 4096ull + c0 + i);
      bool keep = h >= thr;
      if (!keep) m[i >> 3] &= ~(1u << (i & 7));
      v[i] = keep ? v[i] * scale : 0.f;
    }
  }
  for (int a = 0; a < (NOADD ? 0 : e.n_add); ++a) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (g * 8 < nvalid) {
        float t[8];
        load8_dt_n(e.add[a], e.add_dtype, rho * e.ld_out + c0 + g * 8, t, nvalid - g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[g * 8 + i] += t[i];
      }
    }
  }
}

// store half: out, gate-bit bytes, out2 = v * bit(mask2) * scale2 (per-thread global accesses)
template <int NV, bool FULL = false>
__device__ __forceinline__ void epilogue_store(const nbasr_epilogue& e, int64_t rho, int c0, int nvalid_, float* v, const uint32_t* m) {
  constexpr int NG = NV / 8;
  const int nvalid = FULL ? NV : nvalid_;
  if (e.out) {
    if (e.accumulate) {
      float* o = reinterpret_cast<float*>(e.out) + rho * e.ld_out + c0;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (i < nvalid) o[i] += v[i];
    } else {
#pragma unroll
      for (int g = 0; g < NG; ++g)
        if (g * 8 < nvalid) store8_dt_n(e.out, e.out_dtype, rho * e.ld_out + c0 + g * 8, v + g * 8, nvalid - g * 8);
    }
  }
  if (e.mask_out) {
    uint8_t* mo = reinterpret_cast<uint8_t*>(e.mask_out);
    if (NV == 32 && e.mask_w == 32 && (c0 & 31) == 0) {   // one aligned 32-bit entry per (row, plane): coalesced across rows
      uint32_t word = 0;
#pragma unroll
      for (int g = 0; g < NG; ++g) word |= (g * 8 < nvalid ? m[g] : 0u) << (8 * g);
      *reinterpret_cast<uint32_t*>(mo + mask_byte_addr(rho, c0, 32, e.mask_rows)) = word;
    } else {
#pragma unroll
      for (int g = 0; g < NG; ++g)
        if (g * 8 < nvalid) mo[mask_byte_addr(rho, c0 + g * 8, e.mask_w, e.mask_rows)] = static_cast<uint8_t>(m[g]);
    }
  }
  if (e.out2) {
    const uint8_t* mi = reinterpret_cast<const uint8_t*>(e.mask2);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (g * 8 < nvalid) {
        uint32_t w = mi ? mi[mask_byte_addr(rho, c0 + g * 8, e.mask2_w, e.mask_rows)] : 0xffu;
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = ((w >> i) & 1u) ? v[g * 8 + i] * e.scale2 : 0.f;
        store8_dt_n(e.out2, e.out2_dtype, rho * e.ld_out + c0 + g * 8, t, nvalid - g * 8);
      }
    }
  }
}

template <int NV, bool FULL = false>
__device__ __forceinline__ void epilogue_cols(const nbasr_epilogue& e, int64_t rho, int c0, int nvalid, float* v) {
  uint32_t m[NV / 8];
  epilogue_compute<NV, FULL>(e, rho, c0, nvalid, v, m);
  epilogue_store<NV, FULL>(e, rho, c0, nvalid, v, m);
}

// one (row, aligned 32-column chunk); ncol = number of valid columns of the tensor
__device__ __forceinline__ void epilogue_chunk(const nbasr_epilogue& e, int64_t rho, int c0, int ncol, float* v) {
  if (ncol - c0 >= 32) epilogue_cols<32, true>(e, rho, c0, 32, v);
  else epilogue_cols<32, false>(e, rho, c0, ncol - c0, v);
}

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still draining; it must call pdl_wait() before its first access to global memory (the wait returns when the
// predecessor has completed and flushed), and calls pdl_launch_dependents() at its start so that ITS successor's
// prologue (barrier init, TMEM allocation, descriptor prefetch) and launch latency overlap its own tail.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
