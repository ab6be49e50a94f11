// LSTM recurrence (model.py:100,118-121: nn.LSTM(1200,500,batch_first), h0=c0=0, gates i,f,g,o).
//
// Persistent kernels, one launch per sequence pass.  The hidden units are split into slices of
// 16 (64 gate rows) and the batch into independent groups of BG utterances; CTA (slice, group)
// keeps its W_hh slice in REGISTERS for the whole sequence (64 floats / thread), reads h_{t-1}
// of its batch group through L2 each step and synchronises only with the 32 CTAs of its own
// batch group via a global arrive/wait counter (cooperative launch guarantees co-residency).
// Gate math and the cell state stay in fp32; c_t lives in a register of the owning thread.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int L_U = 16;        // hidden units per CTA
constexpr int L_ROWS = 64;     // 4 gates x 16 units
constexpr int L_KS = 8;        // k slices (forward)
constexpr int L_KW = 64;       // k per slice
constexpr int L_KP = 512;      // padded hidden size
constexpr int L_NT = 512;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ void group_barrier(unsigned int* cnt, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(cnt, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt));
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

struct LstmFwdArgs {
  const float* gx;     // (B, T, 4H)
  const float* w_hh;   // (4H, H)
  int T, B, H, BG;
  void* h_seq;
  int h_dtype;
  int64_t h_bs, h_rs;
  float* gates;        // (B, T, 4H)
  float* cstate;       // (B, T, H)
  float* hbuf;         // (2, B, H) ping-pong
  unsigned int* cnt;   // per batch group
};

__global__ void __launch_bounds__(L_NT, 1) lstm_fwd_kernel(LstmFwdArgs p) {
  extern __shared__ float smem[];
  const int BG = p.BG, BGP = BG + 1;
  float* h_s = smem;                       // BG x L_KP
  float* part = h_s + BG * L_KP;           // L_KS x L_ROWS x BGP
  const int tid = threadIdx.x;
  const int u0 = blockIdx.x * L_U;
  const int b0 = blockIdx.y * BG;
  const int nslices = gridDim.x;
  const int H = p.H, H4 = 4 * p.H;
  const int ks = tid / L_ROWS, rr = tid % L_ROWS;
  const int gate = rr / L_U, ul = rr % L_U;
  const bool unit_ok = (u0 + ul) < H;
  float w[L_KW];
#pragma unroll
  for (int i = 0; i < L_KW; ++i) {
    int k = ks * L_KW + i;
    w[i] = (unit_ok && k < H) ? p.w_hh[(int64_t)(gate * H + u0 + ul) * H + k] : 0.f;
  }
  // pair role: (unit, batch) -> cell state register
  const int pul = tid % L_U, pbl = tid / L_U;
  const bool pair_ok = (pbl < BG) && (u0 + pul < H) && (b0 + pbl < p.B);
  float c_reg = 0.f;
  unsigned int* cnt = p.cnt + blockIdx.y;
  // input projection of the NEXT step is fetched before the barrier wait (it does not depend on h)
  float gx_next[4] = {0.f, 0.f, 0.f, 0.f};
  if (pair_ok && p.T > 0) {
    const float* gxr = p.gx + ((int64_t)(b0 + pbl) * p.T) * H4;
#pragma unroll
    for (int g = 0; g < 4; ++g) gx_next[g] = __ldg(gxr + g * H + u0 + pul);
  }

  for (int t = 0; t < p.T; ++t) {
    // stage h_{t-1} of this batch group (zero for t = 0 and for the K padding)
    const float* hprev = p.hbuf + (int64_t)((t + 1) & 1) * p.B * H;
    for (int idx = tid; idx < BG * (L_KP / 4); idx += L_NT) {
      int bl = idx / (L_KP / 4), k = (idx % (L_KP / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t > 0 && k < H && b0 + bl < p.B) v = __ldcg(reinterpret_cast<const float4*>(hprev + (int64_t)(b0 + bl) * H + k));   // H % 4 == 0
      *reinterpret_cast<float4*>(h_s + bl * L_KP + k) = v;
    }
    __syncthreads();
    for (int bl = 0; bl < BG; bl += 4) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float4* h0 = reinterpret_cast<const float4*>(h_s + (bl + 0) * L_KP + ks * L_KW);
      const float4* h1 = reinterpret_cast<const float4*>(h_s + (bl + 1) * L_KP + ks * L_KW);
      const float4* h2 = reinterpret_cast<const float4*>(h_s + (bl + 2) * L_KP + ks * L_KW);
      const float4* h3 = reinterpret_cast<const float4*>(h_s + (bl + 3) * L_KP + ks * L_KW);
#pragma unroll
      for (int i = 0; i < L_KW / 4; ++i) {
        float4 x0 = h0[i], x1 = h1[i], x2 = h2[i], x3 = h3[i];
        a0 = fmaf(w[4 * i], x0.x, a0); a0 = fmaf(w[4 * i + 1], x0.y, a0); a0 = fmaf(w[4 * i + 2], x0.z, a0); a0 = fmaf(w[4 * i + 3], x0.w, a0);
        a1 = fmaf(w[4 * i], x1.x, a1); a1 = fmaf(w[4 * i + 1], x1.y, a1); a1 = fmaf(w[4 * i + 2], x1.z, a1); a1 = fmaf(w[4 * i + 3], x1.w, a1);
        a2 = fmaf(w[4 * i], x2.x, a2); a2 = fmaf(w[4 * i + 1], x2.y, a2); a2 = fmaf(w[4 * i + 2], x2.z, a2); a2 = fmaf(w[4 * i + 3], x2.w, a2);
        a3 = fmaf(w[4 * i], x3.x, a3); a3 = fmaf(w[4 * i + 1], x3.y, a3); a3 = fmaf(w[4 * i + 2], x3.z, a3); a3 = fmaf(w[4 * i + 3], x3.w, a3);
      }
      float* pp = part + (ks * L_ROWS + rr) * BGP + bl;
      pp[0] = a0; pp[1] = a1; pp[2] = a2; pp[3] = a3;
    }
    __syncthreads();
    if (pair_ok) {
      const int b = b0 + pbl, u = u0 + pul;
      float pre[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float s = gx_next[g];
#pragma unroll
        for (int q = 0; q < L_KS; ++q) s += part[(q * L_ROWS + g * L_U + pul) * BGP + pbl];
        pre[g] = s;
      }
      float ig = sigmoidf_(pre[0]), fg = sigmoidf_(pre[1]), gg = tanhf(pre[2]), og = sigmoidf_(pre[3]);
      c_reg = fg * c_reg + ig * gg;
      float h = og * tanhf(c_reg);
      __stcg(p.hbuf + (int64_t)(t & 1) * p.B * H + (int64_t)b * H + u, h);
      int64_t ho = (int64_t)b * p.h_bs + (int64_t)t * p.h_rs + u;
      if (p.h_dtype == NBASR_BF16) reinterpret_cast<bf16*>(p.h_seq)[ho] = __float2bfloat16(h);
      else reinterpret_cast<float*>(p.h_seq)[ho] = h;
      if (p.gates) {
        float* gr = p.gates + ((int64_t)b * p.T + t) * H4;
        gr[u] = ig; gr[H + u] = fg; gr[2 * H + u] = gg; gr[3 * H + u] = og;
        p.cstate[((int64_t)b * p.T + t) * H + u] = c_reg;
      }
      if (t + 1 < p.T) {
        const float* gxr = p.gx + ((int64_t)b * p.T + t + 1) * H4;
#pragma unroll
        for (int g = 0; g < 4; ++g) gx_next[g] = __ldg(gxr + g * H + u);
      }
    }
    if (t + 1 < p.T) group_barrier(cnt, (unsigned)(t + 1) * nslices);
  }
}

struct LstmBwdArgs {
  const float* dh_seq;
  int64_t dh_bs, dh_rs;
  const float* w_hh;
  const float* gates;
  const float* cstate;
  int T, B, H, BG;
  float* dgx;          // (B, T, 4H)
  unsigned int* cnt;
};

constexpr int LB_RS = 32;    // row slices
constexpr int LB_RW = 64;    // rows per slice (4H padded to 2048)
constexpr int LB_RP = 2048;

__global__ void __launch_bounds__(L_NT, 1) lstm_bwd_kernel(LstmBwdArgs p) {
  extern __shared__ float smem[];
  const int BG = p.BG, BGP = BG + 1;
  float* dg_s = smem;                      // BG x LB_RP
  float* part = dg_s + BG * LB_RP;         // LB_RS x L_U x BGP
  const int tid = threadIdx.x;
  const int u0 = blockIdx.x * L_U;
  const int b0 = blockIdx.y * BG;
  const int nslices = gridDim.x;
  const int H = p.H, H4 = 4 * p.H;
  const int rs = tid / L_U, kl = tid % L_U;
  float w[LB_RW];
#pragma unroll
  for (int i = 0; i < LB_RW; ++i) {
    int row = rs * LB_RW + i;
    w[i] = (row < H4 && u0 + kl < H) ? p.w_hh[(int64_t)row * H + u0 + kl] : 0.f;
  }
  const int pul = tid % L_U, pbl = tid / L_U;
  const bool pair_ok = (pbl < BG) && (u0 + pul < H) && (b0 + pbl < p.B);
  float dc_next = 0.f, dh_rec = 0.f;
  unsigned int* cnt = p.cnt + blockIdx.y;

  // saved activations of the step are fetched one step ahead (they do not depend on the recurrence)
  float nx[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // dh_seq, i, f, g, o, c, c_prev
  auto prefetch = [&](int t) {
    const int b = b0 + pbl, u = u0 + pul;
    nx[0] = __ldg(p.dh_seq + (int64_t)b * p.dh_bs + (int64_t)t * p.dh_rs + u);
    const float* gr = p.gates + ((int64_t)b * p.T + t) * H4;
    nx[1] = __ldg(gr + u); nx[2] = __ldg(gr + H + u); nx[3] = __ldg(gr + 2 * H + u); nx[4] = __ldg(gr + 3 * H + u);
    nx[5] = __ldg(p.cstate + ((int64_t)b * p.T + t) * H + u);
    nx[6] = t > 0 ? __ldg(p.cstate + ((int64_t)b * p.T + t - 1) * H + u) : 0.f;
  };
  if (pair_ok && p.T > 0) prefetch(p.T - 1);

  for (int step = 0; step < p.T; ++step) {
    const int t = p.T - 1 - step;
    if (pair_ok) {
      const int b = b0 + pbl, u = u0 + pul;
      float dh = nx[0] + dh_rec;
      float ig = nx[1], fg = nx[2], gg = nx[3], og = nx[4];
      float c = nx[5];
      float cprev = nx[6];
      float tc = tanhf(c);
      float dov = dh * tc;
      float dc = dc_next + dh * og * (1.f - tc * tc);
      float di = dc * gg, dgv = dc * ig, df = dc * cprev;
      dc_next = dc * fg;
      float* dr = p.dgx + ((int64_t)b * p.T + t) * H4;
      __stcg(dr + u, di * ig * (1.f - ig));
      __stcg(dr + H + u, df * fg * (1.f - fg));
      __stcg(dr + 2 * H + u, dgv * (1.f - gg * gg));
      __stcg(dr + 3 * H + u, dov * og * (1.f - og));
      if (t > 0) prefetch(t - 1);
    }
    if (t == 0) break;
    group_barrier(cnt, (unsigned)(step + 1) * nslices);
    // stage dgates_t of the batch group, then dh_rec[b][k] = sum_row dg[b][row] * W_hh[row][k]
    for (int idx = tid; idx < BG * (LB_RP / 4); idx += L_NT) {
      int bl = idx / (LB_RP / 4), row = (idx % (LB_RP / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < H4 && b0 + bl < p.B) v = __ldcg(reinterpret_cast<const float4*>(p.dgx + ((int64_t)(b0 + bl) * p.T + t) * H4 + row));
      *reinterpret_cast<float4*>(dg_s + bl * LB_RP + row) = v;
    }
    __syncthreads();
    for (int bl = 0; bl < BG; bl += 4) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float4* d0 = reinterpret_cast<const float4*>(dg_s + (bl + 0) * LB_RP + rs * LB_RW);
      const float4* d1 = reinterpret_cast<const float4*>(dg_s + (bl + 1) * LB_RP + rs * LB_RW);
      const float4* d2 = reinterpret_cast<const float4*>(dg_s + (bl + 2) * LB_RP + rs * LB_RW);
      const float4* d3 = reinterpret_cast<const float4*>(dg_s + (bl + 3) * LB_RP + rs * LB_RW);
#pragma unroll
      for (int i = 0; i < LB_RW / 4; ++i) {
        float4 x0 = d0[i], x1 = d1[i], x2 = d2[i], x3 = d3[i];
        a0 = fmaf(w[4 * i], x0.x, a0); a0 = fmaf(w[4 * i + 1], x0.y, a0); a0 = fmaf(w[4 * i + 2], x0.z, a0); a0 = fmaf(w[4 * i + 3], x0.w, a0);
        a1 = fmaf(w[4 * i], x1.x, a1); a1 = fmaf(w[4 * i + 1], x1.y, a1); a1 = fmaf(w[4 * i + 2], x1.z, a1); a1 = fmaf(w[4 * i + 3], x1.w, a1);
        a2 = fmaf(w[4 * i], x2.x, a2); a2 = fmaf(w[4 * i + 1], x2.y, a2); a2 = fmaf(w[4 * i + 2], x2.z, a2); a2 = fmaf(w[4 * i + 3], x2.w, a2);
        a3 = fmaf(w[4 * i], x3.x, a3); a3 = fmaf(w[4 * i + 1], x3.y, a3); a3 = fmaf(w[4 * i + 2], x3.z, a3); a3 = fmaf(w[4 * i + 3], x3.w, a3);
      }
      float* pp = part + (rs * L_U + kl) * BGP + bl;
      pp[0] = a0; pp[1] = a1; pp[2] = a2; pp[3] = a3;
    }
    __syncthreads();
    if (pair_ok) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < LB_RS; ++q) s += part[(q * L_U + pul) * BGP + pbl];
      dh_rec = s;
    }
    // part / dg_s are rewritten only after the next group_barrier's __syncthreads
  }
}

int pick_bg(int B, int sms, int nslices, int max_bg) {
  for (int bg = 16; bg <= max_bg; bg *= 2)
    if (nslices * ((B + bg - 1) / bg) <= sms) return bg;
  return -1;
}

}  // namespace

extern "C" {

int nbasr_lstm_fwd(const float* gx, const float* w_hh, int T, int B, int H, void* h_seq, int h_dtype, int64_t h_bs,
                   int64_t h_rs, int64_t ld_h, float* gates, float* cstate, const void* hstate, float* work, void* stream) {
  (void)ld_h;
  if (hstate && h_dtype == NBASR_BF16 && !nbasr_env_flag(NBASR_ENV_FORCE_SIMT))   // bf16 mode: tensor-core cluster kernel
    return sm100_lstm_fwd(gx, hstate, T, B, H, h_seq, h_dtype, h_bs, h_rs, gates, cstate, as_stream(stream));
  NBASR_REQUIRE(H <= L_KP && H % 4 == 0, "hidden size");
  int sms = nbasr_sm_count();
  int nslices = (H + L_U - 1) / L_U;
  int bg = pick_bg(B, sms, nslices, 32);
  NBASR_REQUIRE(bg > 0, "batch too large for one cooperative launch (split the batch)");
  int ngroups = (B + bg - 1) / bg;
  // work layout: [0, 2*B*H) h ping-pong, then ngroups counters
  unsigned int* cnt = reinterpret_cast<unsigned int*>(work + (size_t)2 * B * H);
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * ngroups, st);
  LstmFwdArgs a{gx, w_hh, T, B, H, bg, h_seq, h_dtype, h_bs, h_rs, gates, cstate, work, cnt};
  size_t sm = sizeof(float) * ((size_t)bg * L_KP + (size_t)L_KS * L_ROWS * (bg + 1));
  static DevOnce attr;
  if (!attr) { cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  void* args[] = {&a};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)lstm_fwd_kernel, dim3(nslices, ngroups), dim3(L_NT), args, sm, st);
  if (e != cudaSuccess) return nbasr_fail("lstm_fwd launch: %s", cudaGetErrorString(e));
  return 0;
}

int nbasr_lstm_bwd(const float* dh_seq, int64_t dh_bs, int64_t dh_rs, int64_t ld_dh, const float* w_hh,
                   const float* gates, const float* cstate, int T, int B, int H, float* dgx, float* work,
                   const void* w_hh_packed, void* dgx_bf16, void* stream) {
  (void)ld_dh;
  if (w_hh_packed && !nbasr_env_flag(NBASR_ENV_FORCE_SIMT))
    return sm100_lstm_bwd(dh_seq, dh_bs, dh_rs, w_hh_packed, gates, cstate, T, B, H, dgx, dgx_bf16, as_stream(stream));
  if (dgx_bf16) return nbasr_fail("the fp32 SIMT recurrence does not write a bf16 copy");
  NBASR_REQUIRE(4 * H <= LB_RP && H <= L_KP && H % 4 == 0, "hidden size");
  int sms = nbasr_sm_count();
  int nslices = (H + L_U - 1) / L_U;
  int bg = pick_bg(B, sms, nslices, 16);
  NBASR_REQUIRE(bg > 0, "batch too large for one cooperative launch (split the batch)");
  int ngroups = (B + bg - 1) / bg;
  unsigned int* cnt = reinterpret_cast<unsigned int*>(work + (size_t)2 * B * H) + 64;
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * ngroups, st);
  LstmBwdArgs a{dh_seq, dh_bs, dh_rs, w_hh, gates, cstate, T, B, H, bg, dgx, cnt};
  size_t sm = sizeof(float) * ((size_t)bg * LB_RP + (size_t)LB_RS * L_U * (bg + 1));
  static DevOnce attr;
  if (!attr) { cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  void* args[] = {&a};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)lstm_bwd_kernel, dim3(nslices, ngroups), dim3(L_NT), args, sm, st);
  if (e != cudaSuccess) return nbasr_fail("lstm_bwd launch: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
