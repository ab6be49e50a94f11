// bf16 LayerNorm forward / backward (model.py:47,92: nn.LayerNorm(C, eps=1e-3), biased variance, affine), second
// generation.  Same structure as the kernels in elementwise.cu -- one persistent CTA of 16 warps per SM, a warp per frame
// row, every lane streaming its own 16-byte slices of the next rows into a private shared-memory ring with cp.async -- but
// on an instruction diet, because ncu showed those kernels issue-limited rather than HBM-limited (478 / 996 issued warp
// instructions per 800-channel row, profiles/r1_layernorm_v2.txt):
//   * all fp32 arithmetic is done on PAIRS with the Blackwell packed instructions (fma/add/mul.rn.f32x2 -> FFMA2 / FADD2 /
//     FMUL2): a bf16x2 word unpacks into exactly one such pair (lo = w << 16, hi = w & 0xffff0000), so the conversion cost
//     is unchanged and every arithmetic instruction handles two channels;
//   * row indices are 32-bit and advance incrementally (the old kernels did two 64-bit divisions per row);
//   * only the LAST 32-group slice of a row can be partial, so only that slice is predicated (no per-slice branches).
// Per-element arithmetic is bit-identical to the first generation (same fma nesting); only the order of the row sums differs.
#include "common.cuh"
#include "kernels.h"

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float hsum2(u64 v) {
  float lo, hi;
  upk2(v, lo, hi);
  return lo + hi;
}
// bf16x2 word -> fp32 pair {element 0, element 1}
__device__ __forceinline__ u64 bf2_to_f2(uint32_t w) { return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ uint32_t f2_to_bf2(u64 v) {
  float lo, hi;
  upk2(v, lo, hi);
  return f2_to_bf16x2(lo, hi);
}
// 16-bit pair of either format (F16 = the scaled fp16 forward activations of the 16-bit mode) <-> fp32 pair
template <bool F16>
__device__ __forceinline__ u64 h2_to_f2(uint32_t w) {
  if (F16) {
    const float2 f = f16x2_to_f2(w);
    return pk2(f.x, f.y);
  }
  return bf2_to_f2(w);
}
template <bool F16>
__device__ __forceinline__ uint32_t f2_to_h2(u64 v) {
  float lo, hi;
  upk2(v, lo, hi);
  return F16 ? f2_to_f16x2(lo, hi) : f2_to_bf16x2(lo, hi);
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
// 8 consecutive fp32 from shared memory as 4 pairs
__device__ __forceinline__ void lds8p(uint32_t a, u64* p) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(p[0]), "=l"(p[1]) : "r"(a));
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+16];" : "=l"(p[2]), "=l"(p[3]) : "r"(a));
}
__device__ __forceinline__ void cp16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void stg128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}

// rows are dealt round-robin to the warps of the grid; (b, t) advance incrementally
struct RowCur {
  int r, b, t;
};

constexpr int LN2_D = 4;   // forward: rows in flight per warp

// F16: x and y are fp16 (y multiplied by out_scale, folded into gamma / beta); y2 (optional) = the unscaled result as bf16.
template <int QN, bool F16>
__global__ void __launch_bounds__(512, 1)
ln2_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int Tt, int Tp, int C, const float* __restrict__ gamma,
               const float* __restrict__ beta, float eps, float* __restrict__ mean_o, float* __restrict__ rstd_o, float out_scale,
               bf16* __restrict__ y2) {
  extern __shared__ __align__(16) uint8_t lnsm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * 16 + wib, nW = gridDim.x * 16;
  const int ngroups = C >> 3;
  const int rowb = C * 2;
  const int nrows = B * Tt;
  float* gs = reinterpret_cast<float*>(lnsm);          // gamma | beta
  const uint32_t gsa = (uint32_t)__cvta_generic_to_shared(gs);
  const uint32_t wbuf = (uint32_t)__cvta_generic_to_shared(lnsm) + C * 8 + wib * LN2_D * rowb;
  const bool tail_ok = lane + 32 * (QN - 1) < ngroups;  // only the last slice of a row can be partial
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < C; i += blockDim.x) { gs[i] = gamma[i] * out_scale; gs[C + i] = beta[i] * out_scale; }
  __syncthreads();
  const float inv_os = 1.f / out_scale;            // out_scale is a power of two: y2 = y / out_scale is exact
  const u64 inv_os2 = pk2(inv_os, inv_os);
  const int dq = nW / Tt, dr = nW - dq * Tt;
  auto adv = [&](RowCur& c) {
    c.r += nW; c.b += dq; c.t += dr;
    if (c.t >= Tt) { c.t -= Tt; ++c.b; }
  };
  auto issue = [&](const RowCur& c, int buf) {
    if (c.r < nrows) {
      const bf16* src = x + (size_t)(c.b * Tp + NBASR_PAD_L + c.t) * C + lane * 8;
      const uint32_t dst = wbuf + buf * rowb + lane * 16;
#pragma unroll
      for (int q = 0; q < QN; ++q)
        if (q < QN - 1 || tail_ok) cp16(dst + q * 512, src + q * 256);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  RowCur ci{gw, gw / Tt, gw % Tt};
  RowCur cc = ci;
#pragma unroll
  for (int d = 0; d < LN2_D; ++d) { issue(ci, d); adv(ci); }
  int buf = 0;
  const float invC = 1.f / C;
  for (; cc.r < nrows; adv(cc), buf = (buf + 1 == LN2_D) ? 0 : buf + 1) {
    const int rho = cc.b * Tp + NBASR_PAD_L + cc.t;
    asm volatile("cp.async.wait_group %0;" ::"n"(LN2_D - 1) : "memory");
    const uint32_t xb = wbuf + buf * rowb + lane * 16;
    u64 v[QN][4];
    u64 sa = 0ull, sb = 0ull;                       // {+0.f, +0.f}
#pragma unroll
    for (int q = 0; q < QN; ++q) {
      if (q < QN - 1 || tail_ok) {
        const uint4 w = lds128(xb + q * 512);
        v[q][0] = h2_to_f2<F16>(w.x); v[q][1] = h2_to_f2<F16>(w.y); v[q][2] = h2_to_f2<F16>(w.z); v[q][3] = h2_to_f2<F16>(w.w);
        sa = add2(sa, add2(v[q][0], v[q][1]));
        sb = add2(sb, add2(v[q][2], v[q][3]));
      } else {
        v[q][0] = v[q][1] = v[q][2] = v[q][3] = 0ull;
      }
    }
    const float mean = warp_sum(hsum2(add2(sa, sb))) * invC;
    const u64 nmean2 = pk2(-mean, -mean);
    u64 qa = 0ull, qb = 0ull;
#pragma unroll
    for (int q = 0; q < QN; ++q) {
      if (q < QN - 1 || tail_ok) {
        const u64 d0 = add2(v[q][0], nmean2), d1 = add2(v[q][1], nmean2), d2 = add2(v[q][2], nmean2), d3 = add2(v[q][3], nmean2);
        qa = fma2(d0, d0, qa); qb = fma2(d1, d1, qb);
        qa = fma2(d2, d2, qa); qb = fma2(d3, d3, qb);
      }
    }
    const float rstd = rsqrtf(warp_sum(hsum2(add2(qa, qb))) * invC + eps);
    if (lane == 0 && mean_o) {
      mean_o[rho] = mean;
      rstd_o[rho] = rstd;
    }
    const float nm = -mean * rstd;
    const u64 rstd2 = pk2(rstd, rstd), nm2 = pk2(nm, nm);
    bf16* dst = y + (size_t)rho * C + lane * 8;
#pragma unroll
    for (int q = 0; q < QN; ++q) {
      if (q < QN - 1 || tail_ok) {
        u64 ga[4], be[4];
        lds8p(gsa + (lane * 8 + q * 256) * 4, ga);
        lds8p(gsa + (C + lane * 8 + q * 256) * 4, be);
        u64 r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = fma2(fma2(v[q][i], rstd2, nm2), ga[i], be[i]);
        stg128(dst + q * 256, f2_to_h2<F16>(r[0]), f2_to_h2<F16>(r[1]), f2_to_h2<F16>(r[2]), f2_to_h2<F16>(r[3]));
        if (y2)
          stg128(y2 + (size_t)rho * C + lane * 8 + q * 256, f2_to_bf2(mul2(r[0], inv_os2)), f2_to_bf2(mul2(r[1], inv_os2)),
                 f2_to_bf2(mul2(r[2], inv_os2)), f2_to_bf2(mul2(r[3], inv_os2)));
      }
    }
    issue(ci, buf);       // refill the slot this lane has just finished reading
    adv(ci);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------- backward
// dx = rstd (gy - mean(gy) - xh mean(gy xh)), gy = dy gamma, xh = (x - mean) rstd; dgamma += dy xh, dbeta += dy.
// Optional dx2 = dx * bit(mask2) * scale2 (dZ of the node that produced the pre-norm tensor).  Ring of 2 (x, dy) row pairs.
// The dgamma / dbeta partial sums are 16 QN registers per lane.  For QN >= 4 that spills (measured: slower than the first
// generation), so there the dbeta sums live in a private, conflict-free shared-memory slot per lane (DBS) and QN = 5 runs 12
// warps per CTA instead of 16.
// XF16: x is the scaled fp16 activation x_scale * x_true (mean / rstd are those of the stored tensor, so xhat is exact);
// the gradient wrt x_true is x_scale times the gradient wrt the stored tensor.  dy / dx / dx2 are bf16.
template <int QN, int NWARP, bool DBS, bool XF16>
__global__ void __launch_bounds__(NWARP * 32, 1)
ln2_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
               const float* __restrict__ gamma, int B, int Tt, int Tp, int C, bf16* __restrict__ dx, bf16* __restrict__ dx2,
               const uint32_t* __restrict__ mask2, float scale2, int64_t mask_rows, int mask2_w, float* __restrict__ dgamma,
               float* __restrict__ dbeta, float x_scale) {
  extern __shared__ __align__(16) uint8_t lnsm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * NWARP + wib, nW = gridDim.x * NWARP;
  const int ngroups = C >> 3;
  const int rowb = C * 2;
  const int nrows = B * Tt;
  float* gs = reinterpret_cast<float*>(lnsm);
  const uint32_t gsa = (uint32_t)__cvta_generic_to_shared(gs);
  const uint32_t wbuf = (uint32_t)__cvta_generic_to_shared(lnsm) + C * 4 + wib * 4 * rowb;     // [2 buffers][x row | dy row]
  const bool tail_ok = lane + 32 * (QN - 1) < ngroups;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < C; i += blockDim.x) gs[i] = gamma[i];
  __syncthreads();

  u64 dg[QN][4], db[DBS ? 1 : QN][4];
#pragma unroll
  for (int q = 0; q < QN; ++q)
#pragma unroll
    for (int i = 0; i < 4; ++i) dg[q][i] = 0ull;
  // DBS: [slice q][16-byte half][lane] after the row ring
  const uint32_t dbsa = (uint32_t)__cvta_generic_to_shared(lnsm) + C * 4 + max(NWARP * 4 * rowb, NWARP * 2 * QN * 1024) + wib * QN * 1024 + lane * 16;
#pragma unroll
  for (int q = 0; q < (DBS ? 1 : QN); ++q)
#pragma unroll
    for (int i = 0; i < 4; ++i) db[q][i] = 0ull;
  if (DBS) {
#pragma unroll
    for (int q = 0; q < QN; ++q) {
      asm volatile("st.shared.v2.b64 [%0], {%1, %1};" ::"r"(dbsa + q * 1024), "l"(0ull) : "memory");
      asm volatile("st.shared.v2.b64 [%0], {%1, %1};" ::"r"(dbsa + q * 1024 + 512), "l"(0ull) : "memory");
    }
  }
  const bool has_m2 = dx2 && mask2;
  const int meb = (mask2_w == 32) ? 4 : 8;
  // byte offset of (row 0, this lane's group of slice q) inside the mask planes: < 2^31 (checked by the caller)
  const uint8_t* m2b = reinterpret_cast<const uint8_t*>(mask2);
  int moff[QN];
#pragma unroll
  for (int q = 0; q < QN; ++q) moff[q] = has_m2 ? (int)mask_byte_addr(0, (lane + 32 * q) * 8, mask2_w, mask_rows) : 0;

  const int dq = nW / Tt, dr = nW - dq * Tt;
  auto adv = [&](RowCur& c) {
    c.r += nW; c.b += dq; c.t += dr;
    if (c.t >= Tt) { c.t -= Tt; ++c.b; }
  };
  auto issue = [&](const RowCur& c, int buf) {
    if (c.r < nrows) {
      const size_t e = (size_t)(c.b * Tp + NBASR_PAD_L + c.t) * C + lane * 8;
      const uint32_t dst = wbuf + buf * 2 * rowb + lane * 16;
#pragma unroll
      for (int q = 0; q < QN; ++q)
        if (q < QN - 1 || tail_ok) {
          cp16(dst + q * 512, x + e + q * 256);
          cp16(dst + rowb + q * 512, dy + e + q * 256);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  RowCur cc{gw, gw / Tt, gw % Tt};
  RowCur cn = cc;
  issue(cn, 0);
  float mean_n = 0.f, rstd_n = 0.f;
  if (cc.r < nrows) { const int rho = cc.b * Tp + NBASR_PAD_L + cc.t; mean_n = mean_i[rho]; rstd_n = rstd_i[rho]; }
  int buf = 0;
  const float invC = 1.f / C;
  const u64 sc2 = pk2(scale2, scale2);
  for (; cc.r < nrows; cc = cn, buf ^= 1) {
    const int rho = cc.b * Tp + NBASR_PAD_L + cc.t;
    const float mean = mean_n, rstd = rstd_n;
    adv(cn);
    issue(cn, buf ^ 1);
    if (cn.r < nrows) { const int rhon = cn.b * Tp + NBASR_PAD_L + cn.t; mean_n = mean_i[rhon]; rstd_n = rstd_i[rhon]; }
    // gate bits of this row: issued before any arithmetic so their latency hides behind pass 1
    uint32_t mw[QN];
#pragma unroll
    for (int q = 0; q < QN; ++q) mw[q] = (has_m2 && (q < QN - 1 || tail_ok)) ? (uint32_t)m2b[moff[q] + rho * meb] : 0xffu;
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    const uint32_t xb = wbuf + buf * 2 * rowb + lane * 16;
    const float nm = -mean * rstd;
    const u64 rstd2 = pk2(rstd, rstd), nm2 = pk2(nm, nm);
    u64 s1 = 0ull, s2 = 0ull;
#pragma unroll
    for (int q = 0; q < QN; ++q) {
      if (q < QN - 1 || tail_ok) {
        const uint4 xw = lds128(xb + q * 512), dw = lds128(xb + rowb + q * 512);
        u64 ga[4];
        lds8p(gsa + (lane * 8 + q * 256) * 4, ga);
        const uint32_t xs[4] = {xw.x, xw.y, xw.z, xw.w}, ds[4] = {dw.x, dw.y, dw.z, dw.w};
        u64 dbq[4];
        if (DBS) {
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(dbq[0]), "=l"(dbq[1]) : "r"(dbsa + q * 1024));
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(dbq[2]), "=l"(dbq[3]) : "r"(dbsa + q * 1024 + 512));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const u64 dv = bf2_to_f2(ds[i]);
          const u64 xh = fma2(h2_to_f2<XF16>(xs[i]), rstd2, nm2);
          const u64 gy = mul2(dv, ga[i]);
          s1 = add2(s1, gy);
          s2 = fma2(gy, xh, s2);
          dg[q][i] = fma2(dv, xh, dg[q][i]);
          if (DBS) dbq[i] = add2(dbq[i], dv);
          else db[q][i] = add2(db[q][i], dv);
        }
        if (DBS) {
          asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(dbsa + q * 1024), "l"(dbq[0]), "l"(dbq[1]) : "memory");
          asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(dbsa + q * 1024 + 512), "l"(dbq[2]), "l"(dbq[3]) : "memory");
        }
      }
    }
    const float m1 = warp_sum(hsum2(s1)) * invC;
    const float m2 = warp_sum(hsum2(s2)) * invC;
    // dx = rstd_t (gy - m1 - xh m2) = gy rstd_t + x k1 + k0   (xh = x rstd + nm; rstd_t = x_scale rstd: true-domain rstd)
    const float rstd_t = rstd * x_scale;
    const float k1 = -rstd_t * rstd * m2, k0 = -rstd_t * (m1 + nm * m2);
    const u64 k1p = pk2(k1, k1), k0p = pk2(k0, k0), rstd_t2 = pk2(rstd_t, rstd_t);
    const size_t eo = (size_t)rho * C + lane * 8;
#pragma unroll
    for (int q = 0; q < QN; ++q) {
      if (q < QN - 1 || tail_ok) {
        const uint4 xw = lds128(xb + q * 512), dw = lds128(xb + rowb + q * 512);
        u64 ga[4];
        lds8p(gsa + (lane * 8 + q * 256) * 4, ga);
        const uint32_t xs[4] = {xw.x, xw.y, xw.z, xw.w}, ds[4] = {dw.x, dw.y, dw.z, dw.w};
        u64 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fma2(h2_to_f2<XF16>(xs[i]), k1p, fma2(mul2(bf2_to_f2(ds[i]), ga[i]), rstd_t2, k0p));
        if (dx) stg128(dx + eo + q * 256, f2_to_bf2(o[0]), f2_to_bf2(o[1]), f2_to_bf2(o[2]), f2_to_bf2(o[3]));
        if (dx2) {
          const uint32_t w = mw[q];
          uint32_t o2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float lo, hi;
            upk2(mul2(o[i], sc2), lo, hi);
            o2[i] = f2_to_bf2(pk2(((w >> (2 * i)) & 1u) ? lo : 0.f, ((w >> (2 * i + 1)) & 1u) ? hi : 0.f));
          }
          stg128(dx2 + eo + q * 256, o2[0], o2[1], o2[2], o2[3]);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // CTA reduction over the warps (the row buffers are dead now), then one atomic per channel
  float* red = reinterpret_cast<float*>(lnsm + C * 4);
  const int CP = QN * 256;
#pragma unroll
  for (int q = 0; q < QN; ++q) {
    u64* d0 = reinterpret_cast<u64*>(red + ((size_t)wib * 2) * CP + (lane + 32 * q) * 8);
    u64* d1 = reinterpret_cast<u64*>(red + ((size_t)wib * 2 + 1) * CP + (lane + 32 * q) * 8);
    u64 dbq[4];
    if (DBS) {
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(dbq[0]), "=l"(dbq[1]) : "r"(dbsa + q * 1024));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(dbq[2]), "=l"(dbq[3]) : "r"(dbsa + q * 1024 + 512));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { d0[i] = dg[q][i]; d1[i] = DBS ? dbq[i] : db[q][i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, bsum = 0.f;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
      a += red[((size_t)w * 2) * CP + c];
      bsum += red[((size_t)w * 2 + 1) * CP + c];
    }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, bsum);
  }
}

template <int QN, bool F16>
int ln2_fwd_launch(const bf16* x, bf16* y, int B, int T, int Tp, int C, const float* gamma, const float* beta, float eps, float* mean,
                   float* rstd, float out_scale, bf16* y2, cudaStream_t st) {
  const int64_t rows = (int64_t)B * T;
  const int grid = (int)std::min<int64_t>((rows + 15) / 16, nbasr_sm_count());
  const size_t smb = (size_t)C * 8 + (size_t)16 * LN2_D * C * 2;
  static DevOnce attr;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(ln2_fwd_kernel<QN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return nbasr_fail("ln2_fwd smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  cudaError_t le = launch_pdl(ln2_fwd_kernel<QN, F16>, dim3(grid), dim3(512), smb, st, 1, x, y, B, T, Tp, C, gamma, beta, eps, mean, rstd,
                              out_scale, y2);
  if (le != cudaSuccess) return nbasr_fail("ln2_fwd launch: %s", cudaGetErrorString(le));
  return 0;
}

template <int QN, bool XF16>
int ln2_bwd_launch(const bf16* dy, const bf16* x, const float* mean, const float* rstd, const float* gamma, int B, int T, int Tp, int C,
                   bf16* dx, bf16* dx2, const uint32_t* mask2, float scale2, int64_t mask_rows, int mask2_w, float* dgamma, float* dbeta,
                   float x_scale, cudaStream_t st) {
  constexpr int NWARP = QN <= 4 ? 16 : 12;
  constexpr bool DBS = QN >= 4;
  const int64_t rows = (int64_t)B * T;
  const int grid = (int)std::min<int64_t>((rows + NWARP - 1) / NWARP, nbasr_sm_count());
  // gamma | row ring (re-used by the final reduction) | dbeta slots
  const size_t ring = (size_t)NWARP * 4 * C * 2;
  const size_t smb = (size_t)C * 4 + std::max(ring, (size_t)NWARP * 2 * QN * 256 * 4) + (DBS ? (size_t)NWARP * QN * 1024 : 0);
  static DevOnce attr;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(ln2_bwd_kernel<QN, NWARP, DBS, XF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return nbasr_fail("ln2_bwd smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  cudaError_t le = launch_pdl(ln2_bwd_kernel<QN, NWARP, DBS, XF16>, dim3(grid), dim3(NWARP * 32), smb, st, 1, dy, x, mean, rstd, gamma, B, T, Tp, C,
                              dx, dx2, mask2, scale2, mask_rows, mask2_w, dgamma, dbeta, x_scale);
  if (le != cudaSuccess) return nbasr_fail("ln2_bwd launch: %s", cudaGetErrorString(le));
  return 0;
}

}  // namespace

// rows = B*T must fit 32 bits and (B*Tp + pad) * C must fit size_t arithmetic from 32-bit row indices: checked by the callers
int ln2_fwd(const void* x, void* y, int f16, int B, int T, int Tp, int C, const float* gamma, const float* beta, float eps, float* mean,
            float* rstd, float out_scale, void* y2, cudaStream_t st) {
  const bf16* xx = (const bf16*)x;
  bf16 *yy = (bf16*)y, *y2b = (bf16*)y2;
#define NBASR_LN2F(Q)                                                                                                       \
  case Q:                                                                                                                   \
    return f16 ? ln2_fwd_launch<Q, true>(xx, yy, B, T, Tp, C, gamma, beta, eps, mean, rstd, out_scale, y2b, st)             \
               : ln2_fwd_launch<Q, false>(xx, yy, B, T, Tp, C, gamma, beta, eps, mean, rstd, out_scale, y2b, st);
  switch ((C / 8 + 31) / 32) {
    NBASR_LN2F(1) NBASR_LN2F(2) NBASR_LN2F(3) NBASR_LN2F(4) NBASR_LN2F(5)
  }
#undef NBASR_LN2F
  return nbasr_fail("ln2_fwd: C = %d out of range", C);
}

int ln2_bwd(const void* dy, const void* x, int x_f16, float x_scale, const float* mean, const float* rstd, const float* gamma, int B, int T,
            int Tp, int C, void* dx, void* dx2, const uint32_t* mask2, float scale2, int64_t mask_rows, int mask2_w, float* dgamma,
            float* dbeta, cudaStream_t st) {
  const bf16 *a = (const bf16*)dy, *b = (const bf16*)x;
  bf16 *o = (bf16*)dx, *o2 = (bf16*)dx2;
#define NBASR_LN2B(Q)                                                                                                                   \
  case Q:                                                                                                                               \
    return x_f16 ? ln2_bwd_launch<Q, true>(a, b, mean, rstd, gamma, B, T, Tp, C, o, o2, mask2, scale2, mask_rows, mask2_w, dgamma, dbeta, \
                                           x_scale, st)                                                                                 \
                 : ln2_bwd_launch<Q, false>(a, b, mean, rstd, gamma, B, T, Tp, C, o, o2, mask2, scale2, mask_rows, mask2_w, dgamma,       \
                                            dbeta, x_scale, st);
  switch ((C / 8 + 31) / 32) {
    NBASR_LN2B(1) NBASR_LN2B(2) NBASR_LN2B(3) NBASR_LN2B(4) NBASR_LN2B(5)
  }
#undef NBASR_LN2B
  return nbasr_fail("ln2_bwd: C = %d out of range", C);
}
