// Classifier + log-softmax, CTC loss / gradient (log-space alpha-beta), greedy decode + fold +
// Levenshtein PER.  Small, latency-bound kernels; all arithmetic in fp32 (fp64 for the PER mean).
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int HEAD_MAXK32 = 40;  // K <= 1280

// ---------------------------------------------------------------- head forward: warp per frame row
__global__ void __launch_bounds__(256) head_fwd_kernel(int h_dtype, const void* __restrict__ h, int64_t h_bs, int64_t h_rs,
                                                       int B, int T, int K, int V, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ logits,
                                                       float* __restrict__ logp) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nk = (K + 31) / 32;
  for (int64_t r = warp; r < (int64_t)B * T; r += nwarps) {
    int b = (int)(r / T), t = (int)(r % T);
    int64_t base = (int64_t)b * h_bs + (int64_t)t * h_rs;
    float hv[HEAD_MAXK32];
#pragma unroll
    for (int q = 0; q < HEAD_MAXK32; ++q) {
      int k = lane + 32 * q;
      hv[q] = (q < nk && k < K) ? ld_dt(h, h_dtype, base + k) : 0.f;
    }
    float l0 = -CUDART_INF_F, l1 = -CUDART_INF_F;  // lane v holds logit v and v+32
    for (int v = 0; v < V; ++v) {
      const float* wr = w + (int64_t)v * K;
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < HEAD_MAXK32; ++q) {
        int k = lane + 32 * q;
        if (q < nk && k < K) s = fmaf(hv[q], __ldg(wr + k), s);
      }
      s = warp_sum(s) + bias[v];
      if ((v & 31) == lane) {
        if (v < 32) l0 = s; else l1 = s;
      }
    }
    float m = fmaxf(l0, l1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float e = (lane < V ? __expf(l0 - m) : 0.f) + (lane + 32 < V ? __expf(l1 - m) : 0.f);
    float lse = m + __logf(warp_sum(e));
    if (lane < V) {
      if (logits) logits[r * V + lane] = l0;
      if (logp) logp[r * V + lane] = l0 - lse;
    }
    if (lane + 32 < V) {
      if (logits) logits[r * V + lane + 32] = l1;
      if (logp) logp[r * V + lane + 32] = l1 - lse;
    }
  }
}

// head backward, fused: dh = dl W, db += sum dl, dW += dl^T h.  W and the block's dW partial live in shared
// memory; a block walks its rows 4 at a time (each W element read from smem feeds 4 FMAs); the 8 warps split the
// classes for dW (warp-private rows -> no atomics) and the hidden range for dh.
constexpr int HB_R = 4;

__global__ void __launch_bounds__(256) head_bwd_kernel(int h_dtype, const void* __restrict__ h, int64_t h_bs, int64_t h_rs,
                                                       int B, int T, int K, int V, const float* __restrict__ w,
                                                       const float* __restrict__ dl, float* __restrict__ dh,
                                                       int64_t dh_bs, int64_t dh_rs, float* __restrict__ dw,
                                                       float* __restrict__ db) {
  extern __shared__ float sm[];
  float* ws = sm;                           // V x K
  float* sdw = ws + V * K;                  // V x K (only if dw)
  float* hrow = sdw + (dw ? V * K : 0);     // HB_R x K
  float* drow = hrow + HB_R * K;            // HB_R x 64
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < V * K; i += blockDim.x) {
    ws[i] = w[i];
    if (dw) sdw[i] = 0.f;
  }
  float dbacc = 0.f;                        // thread v < V accumulates db[v]
  const int64_t nrows = (int64_t)B * T;
  const int64_t ngroups = (nrows + HB_R - 1) / HB_R;
  for (int64_t gidx = blockIdx.x; gidx < ngroups; gidx += gridDim.x) {
    __syncthreads();
    const int64_t r0 = gidx * HB_R;
    for (int i = tid; i < HB_R * K; i += blockDim.x) {
      int rr = i / K, k = i - rr * K;
      int64_t r = r0 + rr;
      float v = 0.f;
      if (r < nrows) v = ld_dt(h, h_dtype, (r / T) * h_bs + (r % T) * h_rs + k);
      hrow[i] = v;
    }
    if (tid < HB_R * 64) {
      int rr = tid >> 6, v = tid & 63;
      int64_t r = r0 + rr;
      float d = (v < V && r < nrows) ? dl[r * V + v] : 0.f;
      drow[tid] = d;
    }
    __syncthreads();
    if (tid < V) {
#pragma unroll
      for (int rr = 0; rr < HB_R; ++rr) dbacc += drow[rr * 64 + tid];
    }
    // dh[r][k] = sum_v d[r][v] W[v][k]
    for (int k = tid; k < K; k += blockDim.x) {
      float acc[HB_R];
#pragma unroll
      for (int rr = 0; rr < HB_R; ++rr) acc[rr] = 0.f;
      for (int v = 0; v < V; ++v) {
        float wv = ws[v * K + k];
#pragma unroll
        for (int rr = 0; rr < HB_R; ++rr) acc[rr] = fmaf(drow[rr * 64 + v], wv, acc[rr]);
      }
#pragma unroll
      for (int rr = 0; rr < HB_R; ++rr) {
        int64_t r = r0 + rr;
        if (r < nrows) dh[(r / T) * dh_bs + (r % T) * dh_rs + k] = acc[rr];
      }
    }
    if (dw) {
      for (int v = warp; v < V; v += 8) {
        float d[HB_R];
#pragma unroll
        for (int rr = 0; rr < HB_R; ++rr) d[rr] = drow[rr * 64 + v];
        float* row = sdw + v * K;
        for (int k = lane; k < K; k += 32) {
          float a = row[k];
#pragma unroll
          for (int rr = 0; rr < HB_R; ++rr) a = fmaf(d[rr], hrow[rr * K + k], a);
          row[k] = a;
        }
      }
    }
  }
  __syncthreads();
  if (db && tid < V) atomicAdd(db + tid, dbacc);
  if (dw) for (int i = tid; i < V * K; i += blockDim.x) atomicAdd(dw + i, sdw[i]);
}


// global fp32 array -> shared memory with 8 independent 16-byte loads in flight per thread (a load-then-store loop pays one
// global-memory latency per iteration)
__device__ __forceinline__ void fill_smem_f32(float* dst, const float* __restrict__ src, int n) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i0 = tid; i0 < n4; i0 += 8 * nt) {
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * nt;
        t[u] = i < n4 ? __ldg(reinterpret_cast<const float4*>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * nt;
        if (i < n4) reinterpret_cast<float4*>(dst)[i] = t[u];
      }
    }
  } else {
    for (int i = tid; i < n; i += nt) dst[i] = src[i];
  }
}

// ---------------------------------------------------------------- head forward / dh, register-tiled (K <= 600)
// logits[r][v] = sum_k h[r][k] W[v][k] + b[v] with W resident in shared memory (fp32) and h streamed in 64-wide K
// chunks; a thread owns 4 rows x up to 8 classes (v = c8 + 8 i), so every shared-memory operand feeds 4..8 FMAs and
// no dot product needs a warp reduction.  The 8 threads of a row group finish log-softmax with three shuffles.
constexpr int HT_ROWS = 128;
constexpr int HT_KC = 64;
constexpr int HT_HP = 68;                    // staged h row pitch in floats (16-byte aligned rows)
__global__ void __launch_bounds__(256, 1) head_fwd_tiled_kernel(int h_dtype, const void* __restrict__ h, int64_t h_bs, int64_t h_rs,
                                                                int B, int T, int K, int V, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ logits,
                                                                float* __restrict__ logp) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                       // V x K
  float* hs = ws + V * K;               // HT_ROWS x HT_HP (V * K is a multiple of 4 on the float4 path)
  __shared__ int64_t rbase[HT_ROWS];    // element offset of each row of the tile (-1: past the end)
  const int tid = threadIdx.x, rg = tid >> 3, c8 = tid & 7, lane = tid & 31, wrp = tid >> 5;
  fill_smem_f32(ws, w, V * K);            // classifier weights -> shared memory
  const int64_t nrows = (int64_t)B * T;
  const int ntiles = (int)((nrows + HT_ROWS - 1) / HT_ROWS);
  int vcls[8];
  float bv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int v = c8 + 8 * i;
    vcls[i] = min(v, V - 1);            // clamped duplicates are computed and dropped
    bv[i] = bias[vcls[i]];
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = (int64_t)tile * HT_ROWS;
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
    __syncthreads();
    if (tid < HT_ROWS) {
      const int64_t r = r0 + tid;
      rbase[tid] = r < nrows ? (r / T) * h_bs + (r % T) * h_rs : -1;
    }
    for (int k0 = 0; k0 < K; k0 += HT_KC) {
      __syncthreads();
      // warp w stages rows w, w + 8, ...: one coalesced 128-byte (bf16) / 256-byte (fp32) row segment per pass.  All 16 row
      // loads of a warp are issued before the first shared-memory store (one after the other they cost 16 global-memory
      // latencies per K chunk, 128 per tile: that, not the FMAs, was the time of this kernel)
      {
        float sv0[HT_ROWS / 8], sv1[HT_ROWS / 8];
        const int k = k0 + 2 * lane;
        if (h_dtype == NBASR_BF16 && (K & 1) == 0) {
          // branch-free: every lane loads (a clamped address when its row / column pair is out of range), so the 16 loads are
          // independent instructions the compiler can issue back to back.  The dispatcher guarantees even strides and a
          // 4-byte aligned base, so (base + k) is a bf16x2 boundary.
          __nv_bfloat162 raw[HT_ROWS / 8];
          bool okv[HT_ROWS / 8];
#pragma unroll
          for (int i = 0; i < HT_ROWS / 8; ++i) {
            const int64_t base = rbase[wrp + 8 * i];
            okv[i] = base >= 0 && k < K;
            raw[i] = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const bf16*>(h) + (okv[i] ? base + k : 0));
          }
#pragma unroll
          for (int i = 0; i < HT_ROWS / 8; ++i) {
            const float2 f = __bfloat1622float2(raw[i]);
            sv0[i] = okv[i] ? f.x : 0.f;
            sv1[i] = okv[i] ? f.y : 0.f;
          }
        } else {
#pragma unroll
        for (int i = 0; i < HT_ROWS / 8; ++i) {
          const int64_t base = rbase[wrp + 8 * i];
          float v0 = 0.f, v1 = 0.f;
          if (base >= 0) {
            if (h_dtype == NBASR_BF16) {
              // the dispatcher guarantees even strides and a 4-byte aligned base, so (base + k) is a bf16x2 boundary
              if (k + 1 < K) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const bf16*>(h) + base + k));
                v0 = f.x; v1 = f.y;
              } else if (k < K) {
                v0 = __bfloat162float(reinterpret_cast<const bf16*>(h)[base + k]);
              }
            } else {
              if (k < K) v0 = reinterpret_cast<const float*>(h)[base + k];
              if (k + 1 < K) v1 = reinterpret_cast<const float*>(h)[base + k + 1];
            }
          }
          sv0[i] = v0; sv1[i] = v1;
        }
        }
#pragma unroll
        for (int i = 0; i < HT_ROWS / 8; ++i)
          *reinterpret_cast<float2*>(hs + (wrp + 8 * i) * HT_HP + 2 * lane) = make_float2(sv0[i], sv1[i]);
      }
      __syncthreads();
      const int kn = min(HT_KC, K - k0);
      const float* hp = hs + (4 * rg) * HT_HP;
      const float* wp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) wp[i] = ws + vcls[i] * K + k0;
      auto step = [&](int kk) {
        float hv[4], wv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) hv[j] = hp[j * HT_HP + kk];
#pragma unroll
        for (int i = 0; i < 8; ++i) wv[i] = wp[i][kk];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(hv[j], wv[i], acc[j][i]);
      };
      // four k at a time with 128-bit shared loads (W rows and staged h rows are 16-byte aligned when K % 4 == 0): the
      // scalar version issued 12 shared-memory wavefronts per 32 FMAs and was LSU-bound
      auto step4 = [&](int kk) {
        float4 hv[4], wv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) hv[j] = *reinterpret_cast<const float4*>(hp + j * HT_HP + kk);
#pragma unroll
        for (int i = 0; i < 8; ++i) wv[i] = *reinterpret_cast<const float4*>(wp[i] + kk);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[j][i] = fmaf(hv[j].x, wv[i].x, acc[j][i]);
            acc[j][i] = fmaf(hv[j].y, wv[i].y, acc[j][i]);
            acc[j][i] = fmaf(hv[j].z, wv[i].z, acc[j][i]);
            acc[j][i] = fmaf(hv[j].w, wv[i].w, acc[j][i]);
          }
      };
      if ((K & 3) == 0) {
        const int k4 = kn & ~3;
#pragma unroll 4
        for (int kk = 0; kk < k4; kk += 4) step4(kk);
        for (int kk = k4; kk < kn; ++kk) step(kk);
      } else if (kn == HT_KC) {
#pragma unroll 8
        for (int kk = 0; kk < HT_KC; ++kk) step(kk);
      } else {
        for (int kk = 0; kk < kn; ++kk) step(kk);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = r0 + 4 * rg + j;
      float m = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[j][i] += bv[i];
        if (c8 + 8 * i < V) m = fmaxf(m, acc[j][i]);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float e = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c8 + 8 * i < V) e += __expf(acc[j][i] - m);
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      const float lse = m + __logf(e);
      if (r < nrows) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int v = c8 + 8 * i;
          if (v < V) {
            if (logits) logits[r * V + v] = acc[j][i];
            if (logp) logp[r * V + v] = acc[j][i] - lse;
          }
        }
      }
    }
  }
}

// dh[r][k] = sum_v dl[r][v] W[v][k] (fp32), plus a bf16 copy of dl padded to 64 columns: the operand of the
// tensor-core weight-gradient GEMM that computes dW and db (nbasr_gemm_wgrad) instead of this kernel.
constexpr int HD_ROWS = 64;
__global__ void __launch_bounds__(256, 1) head_bwd_dh_kernel(int B, int T, int K, int V, const float* __restrict__ w,
                                                             const float* __restrict__ dl, float* __restrict__ dh, int64_t dh_bs,
                                                             int64_t dh_rs, bf16* __restrict__ dl16) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                       // V x K
  float* ds = ws + V * K;               // HD_ROWS x 65
  const int tid = threadIdx.x;
  fill_smem_f32(ws, w, V * K);
  const int64_t nrows = (int64_t)B * T;
  const int ntiles = (int)((nrows + HD_ROWS - 1) / HD_ROWS);
  const int k0 = tid, k1 = tid + 256, k2 = tid + 512;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = (int64_t)tile * HD_ROWS;
    __syncthreads();
    // 8 independent loads per thread before the first store (same reason as fill_smem_f32)
    for (int i0 = tid; i0 < HD_ROWS * 64; i0 += 8 * blockDim.x) {
      float dv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        const int rr = i >> 6, v = i & 63;
        const int64_t r = r0 + rr;
        const bool ok = i < HD_ROWS * 64 && v < V && r < nrows;
        dv[u] = __ldg(dl + (ok ? r * V + v : 0));
        if (!ok) dv[u] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < HD_ROWS * 64) {
          const int rr = i >> 6, v = i & 63;
          const int64_t r = r0 + rr;
          ds[rr * 65 + v] = dv[u];
          if (dl16 && r < nrows) dl16[r * 64 + v] = __float2bfloat16(dv[u]);
        }
      }
    }
    __syncthreads();
    for (int r8 = 0; r8 < HD_ROWS; r8 += 8) {
      float acc[8][3];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;
      for (int v = 0; v < V; ++v) {
        const float w0 = k0 < K ? ws[v * K + k0] : 0.f, w1 = k1 < K ? ws[v * K + k1] : 0.f, w2 = k2 < K ? ws[v * K + k2] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = ds[(r8 + j) * 65 + v];
          acc[j][0] = fmaf(d, w0, acc[j][0]);
          acc[j][1] = fmaf(d, w1, acc[j][1]);
          acc[j][2] = fmaf(d, w2, acc[j][2]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t r = r0 + r8 + j;
        if (r < nrows) {
          float* o = dh + (r / T) * dh_bs + (r % T) * dh_rs;
          if (k0 < K) o[k0] = acc[j][0];
          if (k1 < K) o[k1] = acc[j][1];
          if (k2 < K) o[k2] = acc[j][2];
        }
      }
    }
  }
}

// ---------------------------------------------------------------- CTC
__device__ __forceinline__ float logadd(float a, float b) {
  if (a == -CUDART_INF_F) return b;
  if (b == -CUDART_INF_F) return a;
  float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}

// one CTA per utterance; threads over the extended label sequence (2S+1 states, blank = 0)
__global__ void ctc_kernel(const float* __restrict__ logp, int B, int T, int V, const int32_t* __restrict__ targets, int S,
                           const int64_t* __restrict__ audio_len, int len_div, const int64_t* __restrict__ targets_len,
                           float* __restrict__ nll_out, float* __restrict__ loss_out, float* __restrict__ dlogits,
                           float* __restrict__ work) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int Lmax = 2 * S + 1;
  int Tb = (int)(audio_len[b] / len_div);
  if (Tb > T) Tb = T;
  int Sb = (int)targets_len[b];
  if (Sb > S) Sb = S;
  const int L = 2 * Sb + 1;
  // The alpha and beta recursions are independent chains of Tb - 1 latency-bound steps (shared-memory neighbours, two
  // logadds, one emission, a barrier): the first half of the CTA walks alpha forward while the second half walks beta
  // backward, sharing the per-step barrier; the emission of the NEXT step is requested before this step's arithmetic.
  const int half = blockDim.x >> 1;
  const bool do_beta = dlogits != nullptr;
  const bool is_b = threadIdx.x >= half;
  const int tid = threadIdx.x - (is_b ? half : 0);
  float* prev = sm + (is_b ? 2 * Lmax : 0);      // [alpha prev | alpha cur | beta prev | beta cur]
  float* cur = prev + Lmax;
  int* ext = reinterpret_cast<int*>(sm + 4 * Lmax);
  float* alpha = work + (int64_t)b * T * Lmax;
  float* beta = work + (int64_t)B * T * Lmax + (int64_t)b * T * Lmax;
  const float* lp = logp + (int64_t)b * T * V;
  const float NEG = -CUDART_INF_F;
  for (int s = threadIdx.x; s < Lmax; s += blockDim.x) ext[s] = (s < L && (s & 1)) ? targets[(int64_t)b * S + (s >> 1)] : 0;
  __syncthreads();
  __shared__ float s_nll;
  if (!is_b) {
    for (int s = tid; s < L; s += half) {
      float a = NEG;
      if (Tb > 0) {
        if (s == 0) a = lp[0];
        else if (s == 1) a = lp[ext[1]];
      }
      prev[s] = a;
      if (Tb > 0) alpha[s] = a;
    }
  } else if (do_beta && Tb > 0) {
    for (int s = tid; s < L; s += half) {
      float bv = NEG;
      if (s == L - 1) bv = lp[(int64_t)(Tb - 1) * V];
      else if (s == L - 2) bv = lp[(int64_t)(Tb - 1) * V + ext[L - 2]];
      prev[s] = bv;
      beta[(int64_t)(Tb - 1) * Lmax + s] = bv;
    }
  }
  __syncthreads();
  if (L <= half) {
    // common case: one state per thread, emission software-pipelined
    const int s = tid;
    const bool on = s < L && (!is_b || do_beta);
    const int es = on ? ext[s] : 0;
    const bool skip_ok = on && (is_b ? (s + 2 < L && ext[s + 2] != 0 && ext[s + 2] != es) : (s >= 2 && es != 0 && es != ext[s - 2]));
    float e_next = 0.f;
    if (on && Tb > 1) e_next = lp[(int64_t)(is_b ? Tb - 2 : 1) * V + es];
    for (int k = 1; k < Tb; ++k) {
      const int t = is_b ? Tb - 1 - k : k;
      const float e = e_next;
      if (on && k + 1 < Tb) e_next = lp[(int64_t)(is_b ? t - 1 : t + 1) * V + es];
      if (on) {
        float a = prev[s];
        if (!is_b) {
          if (s >= 1) a = logadd(a, prev[s - 1]);
          if (skip_ok) a = logadd(a, prev[s - 2]);
        } else {
          if (s + 1 < L) a = logadd(a, prev[s + 1]);
          if (skip_ok) a = logadd(a, prev[s + 2]);
        }
        if (a != NEG) a += e;
        cur[s] = a;
        (is_b ? beta : alpha)[(int64_t)t * Lmax + s] = a;
      }
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    }
  } else {
    for (int k = 1; k < Tb; ++k) {
      const int t = is_b ? Tb - 1 - k : k;
      if (!is_b) {
        for (int s = tid; s < L; s += half) {
          float a = prev[s];
          if (s >= 1) a = logadd(a, prev[s - 1]);
          if (s >= 2 && ext[s] != 0 && ext[s] != ext[s - 2]) a = logadd(a, prev[s - 2]);
          if (a != NEG) a += lp[(int64_t)t * V + ext[s]];
          cur[s] = a;
          alpha[(int64_t)t * Lmax + s] = a;
        }
      } else if (do_beta) {
        for (int s = tid; s < L; s += half) {
          float bv = prev[s];
          if (s + 1 < L) bv = logadd(bv, prev[s + 1]);
          if (s + 2 < L && ext[s + 2] != 0 && ext[s + 2] != ext[s]) bv = logadd(bv, prev[s + 2]);
          if (bv != NEG) bv += lp[(int64_t)t * V + ext[s]];
          cur[s] = bv;
          beta[(int64_t)t * Lmax + s] = bv;
        }
      }
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    }
  }
  if (threadIdx.x == 0) {          // an alpha thread: its `prev` is the last alpha row
    float ll = NEG;
    if (Tb > 0) {
      ll = prev[L - 1];
      if (L > 1) ll = logadd(ll, prev[L - 2]);
    }
    float nll = -ll;
    if (nll == CUDART_INF_F || nll != nll) nll = CUDART_INF_F;
    s_nll = nll;
    float z = (nll == CUDART_INF_F) ? 0.f : nll;   // zero_infinity=True
    nll_out[b] = z;
    if (loss_out && Tb > 0) atomicAdd(loss_out, z / (float)Tb / (float)B);
  }
  __syncthreads();
  if (!dlogits) return;
  const float nll = s_nll;
  float* dl = dlogits + (int64_t)b * T * V;
  if (nll == CUDART_INF_F || Tb == 0) {
    for (int i = threadIdx.x; i < T * V; i += blockDim.x) dl[i] = 0.f;
    return;
  }
  // ---- gradient wrt logits (log-softmax backward folded in); thread per frame
  const float scale = 1.f / ((float)Tb * (float)B);
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float* row = dl + (int64_t)t * V;
    if (t >= Tb) {
      for (int c = 0; c < V; ++c) row[c] = 0.f;
      continue;
    }
    for (int c = 0; c < V; ++c) row[c] = NEG;
    for (int s = 0; s < L; ++s) {
      float ab = alpha[(int64_t)t * Lmax + s] + beta[(int64_t)t * Lmax + s];
      int c = ext[s];
      row[c] = logadd(row[c], ab);
    }
    for (int c = 0; c < V; ++c) {
      float l = lp[(int64_t)t * V + c];
      float occ = (row[c] == NEG) ? 0.f : expf(row[c] + nll - l);
      row[c] = (expf(l) - occ) * scale;
    }
  }
}

// ---------------------------------------------------------------- CTC prefix beam search (width <= 16)
// What Trainer.decode actually calls (trainer.py:71,236): ctcdecode.CTCBeamDecoder(labels, beam_width=12,
// log_probs_input=True) with its defaults -- no language model, cutoff_top_n = 40, cutoff_prob = 1, blank 0.  The
// library (parlance/ctcdecode @9a20e00f, not under /root/reference) implements the prefix beam search of the
// PaddlePaddle DeepSpeech decoder; restated here (and in oracle/decode_np.py:beam_search) as:
//   per frame: keep the 40 most probable classes; for every beam prefix l (p_b, p_nb = log-prob of l ending in blank /
//   non-blank, score = logaddexp):  blank -> p_b'(l) += lp[0] + score(l);  c = last(l) -> p_nb'(l) += lp[c] + p_nb(l);
//   extension l+c -> p_nb'(l+c) += lp[c] + (c == last(l) ? p_b(l) : score(l)), merged with l+c if that prefix is in the
//   beam already; keep the `width` best prefixes by score'.  Scores are fp64 here and in the oracle (ctcdecode uses
//   fp32); ties break on (last label, slot).  Output: best prefix per utterance.  One CTA per utterance.
constexpr int BS_MAXW = 16;
constexpr int BS_MAXV = 64;
constexpr int BS_MAXC = BS_MAXW * BS_MAXV + BS_MAXW;       // candidate slots per frame

__device__ __forceinline__ double bs_logadd(double a, double b) {
  if (a == -CUDART_INF) return b;
  if (b == -CUDART_INF) return a;
  const double m = fmax(a, b);
  return m + log(exp(a - m) + exp(b - m));
}

__global__ void __launch_bounds__(256) beam_search_kernel(const float* __restrict__ logp, int B, int T, int V,
                                                          const int64_t* __restrict__ audio_len, int len_div, int width, int top_n,
                                                          int32_t* __restrict__ out, int32_t* __restrict__ out_len) {
  extern __shared__ int16_t bs_seq[];               // [2][width][T]
  __shared__ double lp[BS_MAXV];
  __shared__ int keep[BS_MAXV];
  __shared__ double pb[BS_MAXW], pnb[BS_MAXW], sc[BS_MAXW];      // current beam (previous frame's p_b, p_nb, score)
  __shared__ int blen[BS_MAXW], blast[BS_MAXW];
  __shared__ unsigned long long bhash[BS_MAXW];
  __shared__ double c_pb[BS_MAXC], c_pnb[BS_MAXC], c_sc[BS_MAXC];
  __shared__ int c_par[BS_MAXC], c_chr[BS_MAXC], c_ok[BS_MAXC];
  __shared__ int win[BS_MAXW];
  __shared__ int s_n;
  const int b = blockIdx.x, tid = threadIdx.x;
  int Tb = (int)(audio_len[b] / len_div);
  if (Tb > T) Tb = T;
  const float* lpg = logp + (int64_t)b * T * V;
  int cur = 0;
  if (tid == 0) {           // root prefix: empty, p_b = 0 (log 1), p_nb = -inf
    s_n = 1; pb[0] = 0.0; pnb[0] = -CUDART_INF; sc[0] = 0.0; blen[0] = 0; blast[0] = -1; bhash[0] = 1469598103934665603ull;
  }
  __syncthreads();
  for (int t = 0; t < Tb; ++t) {
    const int n = s_n;
    if (tid < V) lp[tid] = (double)lpg[(int64_t)t * V + tid];
    __syncthreads();
    if (tid < V) {          // rank among the frame's classes (larger first, ties by index): keep the top_n
      int r = 0;
      const double me = lp[tid];
      for (int c = 0; c < V; ++c) r += (lp[c] > me || (lp[c] == me && c < tid)) ? 1 : 0;
      keep[tid] = r < top_n;
    }
    const int nslots = n + n * V;
    for (int s = tid; s < nslots; s += blockDim.x) c_ok[s] = 0;
    __syncthreads();
    // existing prefixes: slot i
    if (tid < n) {
      const int i = tid;
      c_pb[i] = keep[0] ? sc[i] + lp[0] : -CUDART_INF;
      c_pnb[i] = (blast[i] > 0 && keep[blast[i]]) ? lp[blast[i]] + pnb[i] : -CUDART_INF;
      c_par[i] = i; c_chr[i] = -1; c_ok[i] = 1;
    }
    __syncthreads();
    // extensions: slot n + i*V + c
    for (int e = tid; e < n * V; e += blockDim.x) {
      const int i = e / V, c = e - i * V;
      if (c == 0 || !keep[c]) continue;
      const double l = (c == blast[i]) ? (pb[i] == -CUDART_INF ? -CUDART_INF : lp[c] + pb[i]) : lp[c] + sc[i];
      if (l == -CUDART_INF) continue;
      // is prefix_i + c already in the beam?  (len, last, hash, then the labels themselves)
      const unsigned long long h = bhash[i] * 1099511628211ull + (unsigned long long)c;
      int j_same = -1;
      for (int j = 0; j < n; ++j) {
        if (blen[j] == blen[i] + 1 && blast[j] == c && bhash[j] == h) {
          const int16_t* a = bs_seq + ((size_t)cur * width + i) * T;
          const int16_t* q = bs_seq + ((size_t)cur * width + j) * T;
          bool eq = true;
          for (int k = 0; k < blen[i]; ++k) eq = eq && (a[k] == q[k]);
          if (eq) { j_same = j; break; }
        }
      }
      const int s = n + e;
      c_pb[s] = -CUDART_INF; c_pnb[s] = l; c_par[s] = i; c_chr[s] = c;
      c_ok[s] = j_same < 0 ? 1 : -(j_same + 1);          // negative: merge into existing slot j_same
    }
    __syncthreads();
    // merges: at most one parent per existing prefix, so slot j is updated by exactly one thread
    for (int s = n + tid; s < nslots; s += blockDim.x)
      if (c_ok[s] < 0) { const int j = -c_ok[s] - 1; c_pnb[j] = bs_logadd(c_pnb[j], c_pnb[s]); c_ok[s] = 0; }
    __syncthreads();
    for (int s = tid; s < nslots; s += blockDim.x)
      if (c_ok[s]) {
        c_sc[s] = bs_logadd(c_pb[s], c_pnb[s]);
        if (c_sc[s] == -CUDART_INF) c_ok[s] = 0;         // dead prefix (possible when blank / last label were cut off)
      }
    __syncthreads();
    // top `width` by (score desc, last label asc, slot asc): rank by counting
    if (tid < BS_MAXW) win[tid] = -1;
    __syncthreads();
    for (int s = tid; s < nslots; s += blockDim.x) {
      if (!c_ok[s]) continue;
      const double me = c_sc[s];
      const int mc = c_chr[s] >= 0 ? c_chr[s] : blast[c_par[s]];
      int r = 0;
      for (int u = 0; u < nslots; ++u) {
        if (!c_ok[u] || u == s) continue;
        const double ot = c_sc[u];
        const int oc = c_chr[u] >= 0 ? c_chr[u] : blast[c_par[u]];
        if (ot > me || (ot == me && (oc < mc || (oc == mc && u < s)))) ++r;
      }
      if (r < width) win[r] = s;
    }
    __syncthreads();
    int nn = 0;
    for (int r = 0; r < width; ++r) nn += win[r] >= 0 ? 1 : 0;      // winners fill ranks 0..nn-1 contiguously
    // new beam into the other half of the label store
    for (int r = 0; r < nn; ++r) {
      const int s = win[r], i = c_par[s];
      const int16_t* a = bs_seq + ((size_t)cur * width + i) * T;
      int16_t* q = bs_seq + ((size_t)(cur ^ 1) * width + r) * T;
      for (int k = tid; k < blen[i]; k += blockDim.x) q[k] = a[k];
      if (tid == 0 && c_chr[s] >= 0) q[blen[i]] = (int16_t)c_chr[s];
    }
    __syncthreads();
    double npb = 0, npnb = 0, nsc = 0; int nlen = 0, nlast = 0; unsigned long long nh = 0;
    if (tid < nn) {
      const int s = win[tid], i = c_par[s];
      npb = c_pb[s]; npnb = c_pnb[s]; nsc = c_sc[s];
      nlen = blen[i] + (c_chr[s] >= 0 ? 1 : 0);
      nlast = c_chr[s] >= 0 ? c_chr[s] : blast[i];
      nh = c_chr[s] >= 0 ? bhash[i] * 1099511628211ull + (unsigned long long)c_chr[s] : bhash[i];
    }
    __syncthreads();
    if (tid < nn) { pb[tid] = npb; pnb[tid] = npnb; sc[tid] = nsc; blen[tid] = nlen; blast[tid] = nlast; bhash[tid] = nh; }
    if (tid == 0) s_n = nn;
    cur ^= 1;
    __syncthreads();
  }
  // best prefix = rank 0 of the last selection (the root if there were no frames)
  const int L = s_n > 0 ? blen[0] : 0;
  const int16_t* a = bs_seq + ((size_t)cur * width + 0) * T;
  for (int k = tid; k < T; k += blockDim.x) out[(int64_t)b * T + k] = k < L ? (int32_t)a[k] : 0;
  if (tid == 0) out_len[b] = L;
}

// ---------------------------------------------------------------- greedy decode (or a given hypothesis) + PER
// pre_hyp / pre_len (optional): unfolded label sequences from beam_search_kernel, used instead of the argmax path.
__global__ void greedy_per_kernel(const float* __restrict__ logp, int B, int T, int V, const int64_t* __restrict__ audio_len,
                                  int len_div, const int32_t* __restrict__ targets, int S,
                                  const int64_t* __restrict__ targets_len, const int32_t* __restrict__ lut,
                                  int32_t* __restrict__ hyp, int32_t* __restrict__ hyp_len, int32_t* __restrict__ dist,
                                  double* __restrict__ per, int32_t* __restrict__ work, const int32_t* __restrict__ pre_hyp,
                                  const int32_t* __restrict__ pre_len) {
  extern __shared__ int smi[];
  int* best = smi;              // T
  int* hy = best + T;           // T
  int* ref = hy + T;            // S
  int* d0 = ref + S;            // S+1  (three anti-diagonals)
  int* d1 = d0 + S + 1;
  int* d2 = d1 + S + 1;
  __shared__ int s_n, s_m;
  const int b = blockIdx.x;
  int Tb = (int)(audio_len[b] / len_div);
  if (Tb > T) Tb = T;
  const float* lp = logp + (int64_t)b * T * V;
  for (int t = threadIdx.x; t < Tb && !pre_hyp; t += blockDim.x) {
    const float* row = lp + (int64_t)t * V;
    float bv = row[0];
    int bi = 0;
    for (int c = 1; c < V; ++c) {
      float v = row[c];
      if (v > bv) { bv = v; bi = c; }     // strict > keeps the first maximum, like torch.argmax
    }
    best[t] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0, prev = -1;
    if (pre_hyp) {
      const int L = min(pre_len[b], T);
      for (int k = 0; k < L; ++k) {
        int c = pre_hyp[(int64_t)b * T + k];
        int f = lut ? lut[c] : c;
        if (f != 0) hy[n++] = f;
      }
    } else {
      for (int t = 0; t < Tb; ++t) {
        int c = best[t];
        if (c != prev && c != 0) {
          int f = lut ? lut[c] : c;
          if (f != 0) hy[n++] = f;
        }
        prev = c;
      }
    }
    s_n = n;
    int m = 0;
    int tl = (int)targets_len[b];
    if (tl > S) tl = S;
    for (int j = 0; j < tl; ++j) {
      int c = targets[(int64_t)b * S + j];
      int f = lut ? lut[c] : c;
      if (f != 0) ref[m++] = f;
    }
    s_m = m;
    hyp_len[b] = n;
  }
  __syncthreads();
  const int n = s_n, m = s_m;
  for (int t = threadIdx.x; t < T; t += blockDim.x) hyp[(int64_t)b * T + t] = t < n ? hy[t] : 0;
  // Levenshtein over anti-diagonals k = i + j; cell (i, j) lives at index j of diagonal k.
  // dk[j] = min(d(k-1)[j] + 1 [delete hyp i], d(k-1)[j-1] + 1 [insert ref j], d(k-2)[j-1] + (hy[i-1] != ref[j-1]))
  int* pm2 = d0; int* pm1 = d1; int* cur = d2;
  for (int k = 0; k <= n + m; ++k) {
    int jlo = max(0, k - n), jhi = min(m, k);
    for (int j = jlo + threadIdx.x; j <= jhi; j += blockDim.x) {
      int i = k - j;
      int v;
      if (i == 0) v = j;
      else if (j == 0) v = i;
      else {
        int sub = pm2[j - 1] + (hy[i - 1] != ref[j - 1] ? 1 : 0);
        int del = pm1[j] + 1;
        int ins = pm1[j - 1] + 1;
        v = min(sub, min(del, ins));
      }
      cur[j] = v;
    }
    __syncthreads();
    int* tmp = pm2; pm2 = pm1; pm1 = cur; cur = tmp;
  }
  if (threadIdx.x == 0) {
    dist[b] = pm1[m];       // after the last swap pm1 holds diagonal n+m
    __threadfence();
    int ticket = atomicAdd(work, 1);
    if (ticket == B - 1) {   // last utterance finished: sequential fp64 mean, deterministic order
      __threadfence();
      double acc = 0.0;
      for (int i = 0; i < B; ++i) {
        int d = *((volatile int32_t*)dist + i);
        acc += (double)d / (double)targets_len[i];
      }
      per[0] = acc / (double)B;
      reinterpret_cast<float*>(per + 1)[0] = (float)(acc / (double)B);
      *work = 0;
    }
  }
}

}  // namespace

extern "C" {

int nbasr_head_fwd(int h_dtype, const void* h, int64_t h_bs, int64_t h_rs, int B, int T, int K, int V, const float* w,
                   const float* bias, float* logits, float* logp, void* stream) {
  NBASR_REQUIRE(V <= 64 && K <= 32 * HEAD_MAXK32, "head shape");
  int64_t rows = (int64_t)B * T;
  if (rows == 0) return 0;
  const size_t smt = sizeof(float) * ((size_t)V * K + HT_ROWS * HT_HP);
  const bool aligned = h_dtype != NBASR_BF16 || (((h_bs | h_rs) & 1) == 0 && ((uintptr_t)h & 3) == 0);
  if (smt <= 200 * 1024 && rows >= 4 * HT_ROWS && aligned) {
    static DevOnce attr;
    if (!attr) { cudaFuncSetAttribute(head_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
    int grid = (int)std::min<int64_t>((rows + HT_ROWS - 1) / HT_ROWS, nbasr_sm_count());
    head_fwd_tiled_kernel<<<grid, 256, smt, as_stream(stream)>>>(h_dtype, h, h_bs, h_rs, B, T, K, V, w, bias, logits, logp);
    NBASR_CHECK_LAUNCH();
    return 0;
  }
  int blocks = (int)std::min<int64_t>((rows + 7) / 8, 148 * 8);
  head_fwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(h_dtype, h, h_bs, h_rs, B, T, K, V, w, bias, logits, logp);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_head_bwd(int h_dtype, const void* h, int64_t h_bs, int64_t h_rs, int B, int T, int K, int V, const float* w,
                   const float* dlogits, float* dh, int64_t dh_bs, int64_t dh_rs, float* dw, float* db, void* stream) {
  NBASR_REQUIRE(V <= 64 && K <= 32 * HEAD_MAXK32, "head shape");
  int64_t rows = (int64_t)B * T;
  if (rows == 0) return 0;
  const bool fuse_dw = dw && (size_t)(2 * V * K + HB_R * K + HB_R * 64) * sizeof(float) <= 220 * 1024;
  size_t sm = sizeof(float) * ((size_t)V * K + (fuse_dw ? (size_t)V * K : 0) + HB_R * K + HB_R * 64);
  NBASR_REQUIRE(sm <= 220 * 1024, "head too wide for the fused backward kernel");
  static DevOnce attr;
  if (!attr) { cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); attr = true; }
  int blocks = (int)std::min<int64_t>((rows + HB_R - 1) / HB_R, nbasr_sm_count());
  head_bwd_kernel<<<blocks, 256, sm, as_stream(stream)>>>(h_dtype, h, h_bs, h_rs, B, T, K, V, w, dlogits, dh, dh_bs, dh_rs,
                                                          fuse_dw ? dw : nullptr, db);
  NBASR_CHECK_LAUNCH();
  if (dw && !fuse_dw) {
    // dW[v, k] += sum_{b,t} dl[b,t,v] * h[b,t,k]
    SimtGemmArgs a{};
    a.a = dlogits; a.a_dtype = NBASR_F32; a.a_ib = 0; a.a_ir = 1; a.a_kb = (int64_t)T * V; a.a_kr = V;
    a.nib = 1; a.nir = V;
    a.b = h; a.b_dtype = h_dtype; a.b_j = 1; a.b_kb = h_bs; a.b_kr = h_rs;
    a.nkb = B; a.nkr = T; a.N = K;
    a.o_r0 = 0; a.o_bs = 0; a.o_rs = 1;
    a.epi.out = dw; a.epi.out_dtype = NBASR_F32; a.epi.ld_out = K; a.epi.accumulate = 1;
    return simt_gemm_launch(a, as_stream(stream));
  }
  return 0;
}

int nbasr_head_bwd_dh(int B, int T, int K, int V, const float* w, const float* dlogits, float* dh, int64_t dh_bs, int64_t dh_rs,
                      void* dl16, void* stream) {
  NBASR_REQUIRE(V <= 64 && K <= 768, "head shape");
  int64_t rows = (int64_t)B * T;
  if (rows == 0) return 0;
  const size_t smt = sizeof(float) * ((size_t)V * K + HD_ROWS * 65);
  NBASR_REQUIRE(smt <= 200 * 1024, "head too wide for head_bwd_dh");
  static DevOnce attr;
  if (!attr) { cudaFuncSetAttribute(head_bwd_dh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  int grid = (int)std::min<int64_t>((rows + HD_ROWS - 1) / HD_ROWS, nbasr_sm_count());
  head_bwd_dh_kernel<<<grid, 256, smt, as_stream(stream)>>>(B, T, K, V, w, dlogits, dh, dh_bs, dh_rs, (bf16*)dl16);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_ctc(const float* logp, int B, int T, int V, const int32_t* targets, int S, const int64_t* audio_len, int len_div,
              const int64_t* targets_len, float* nll, float* loss, float* dlogits, float* work, void* stream) {
  if (B == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (loss) cudaMemsetAsync(loss, 0, sizeof(float), st);
  int Lmax = 2 * S + 1;
  // two halves: alpha threads | beta threads (one state per thread when 2S+1 <= 512)
  int threads = 2 * std::min(512, std::max(32, ((Lmax + 31) / 32) * 32));
  size_t sm = sizeof(float) * 5 * Lmax;
  ctc_kernel<<<B, threads, sm, st>>>(logp, B, T, V, targets, S, audio_len, len_div, targets_len, nll, loss, dlogits, work);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_greedy_per(const float* logp, int B, int T, int V, const int64_t* audio_len, int len_div, const int32_t* targets,
                     int S, const int64_t* targets_len, const int32_t* lut, int32_t* hyp, int32_t* hyp_len, int32_t* dist,
                     double* per, int32_t* work, void* stream) {
  if (B == 0) return 0;
  size_t sm = sizeof(int) * ((size_t)2 * T + S + 3 * (S + 1));
  NBASR_REQUIRE(sm <= 200 * 1024, "sequence too long for the decode kernel");
  static DevOnce attr;
  if (!attr) { cudaFuncSetAttribute(greedy_per_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  greedy_per_kernel<<<B, 256, sm, as_stream(stream)>>>(logp, B, T, V, audio_len, len_div, targets, S, targets_len, lut, hyp,
                                                       hyp_len, dist, per, work, nullptr, nullptr);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int nbasr_beam_per(const float* logp, int B, int T, int V, const int64_t* audio_len, int len_div, int beam_width, int cutoff_top_n,
                   const int32_t* targets, int S, const int64_t* targets_len, const int32_t* lut, int32_t* raw, int32_t* raw_len,
                   int32_t* hyp, int32_t* hyp_len, int32_t* dist, double* per, int32_t* work, void* stream) {
  if (B == 0) return 0;
  NBASR_REQUIRE(beam_width >= 1 && beam_width <= BS_MAXW && V <= BS_MAXV && T < 32768, "beam search shape");
  const size_t smb = (size_t)2 * beam_width * T * sizeof(int16_t);
  NBASR_REQUIRE(smb <= 150 * 1024, "sequence too long for the beam-search kernel");
  static DevOnce attr;
  if (!attr) {
    cudaFuncSetAttribute(beam_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
    cudaFuncSetAttribute(greedy_per_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  beam_search_kernel<<<B, 256, smb, as_stream(stream)>>>(logp, B, T, V, audio_len, len_div, beam_width,
                                                         cutoff_top_n > 0 ? cutoff_top_n : V, raw, raw_len);
  NBASR_CHECK_LAUNCH();
  size_t sm = sizeof(int) * ((size_t)2 * T + S + 3 * (S + 1));
  NBASR_REQUIRE(sm <= 200 * 1024, "sequence too long for the decode kernel");
  greedy_per_kernel<<<B, 256, sm, as_stream(stream)>>>(logp, B, T, V, audio_len, len_div, targets, S, targets_len, lut, hyp,
                                                       hyp_len, dist, per, work, raw, raw_len);
  NBASR_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
