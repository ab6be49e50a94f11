// LSTM recurrence on tcgen05 tensor cores with thread-block clusters (bf16 mode).
//
// One cluster of 16 CTAs per group of 16 utterances.  CTA j owns hidden units [32j, 32j+32): its 128 gate
// rows (i,f,g,o x 32 units) of W_hh stay in shared memory for the whole sequence (128 x 512 bf16 = 128 KB,
// K-major, 128B swizzle, TMA-loaded once).  Per time step
//     D[128 gate rows x 16 utterances] = W_slice[128 x 512] * h_{t-1}^T[512 x 16]
// is 32 tcgen05.mma (M=128, N=16, K=16) into TMEM; 4 epilogue warps read the accumulator, add the input
// projection, apply the gates (fp32, cell state in registers), and the CTA's new h slice (32 units x 16
// utterances, bf16) is pushed into the h-operand buffer of ALL 16 CTAs of the cluster with one 1-KB
// cp.async.bulk (smem -> distributed smem) per destination.  The h operand uses the un-swizzled K-major
// canonical layout [k-chunk][row group][8 rows][8 elems] in which a CTA's slice is one contiguous KB; each
// destination's mbarrier counts the 16 KB it expects, so data arrival IS the inter-CTA synchronisation --
// there is no cluster barrier and no global-memory round trip inside the time loop.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int LC_NCTA = 16;
constexpr int LC_U = 32;
constexpr int LC_BG = 16;
constexpr int LC_KP = 512;
constexpr int LC_W_BYTES = 128 * LC_KP * 2;      // 131072
constexpr int LC_H_BYTES = LC_BG * LC_KP * 2;    // 16384
constexpr int LC_SLICE = LC_U * LC_BG * 2;       // 1024: one CTA's h slice
constexpr int LC_PRE_LD = 17;
constexpr int LC_FWD_EPI = 512;                  // both kernels: 16 epilogue warps, one cell per thread, + the control warp (TMA, MMA) last
constexpr int LC_FWD_THREADS = 32 + LC_FWD_EPI;
constexpr int LC_SMEM = LC_W_BYTES + 2 * LC_H_BYTES + 2 * LC_SLICE + 128 * LC_PRE_LD * 4 + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + __expf(-x)); }

struct LstmCArgs {
  const float* gx;     // (B, T, 4H)
  int T, B, H;
  void* h_seq;
  int h_dtype;
  int64_t h_bs, h_rs;
  float* gates;        // (B, T, 4H)
  float* cstate;       // (B, T, H)
};

// A operand from tensor memory: D[tmem] (+)= A[tmem] * B[smem].  Lane r of the A region holds row r, K packed two bf16 per
// 32-bit column (16 K elements = 8 columns), i.e. exactly what the row's owner thread writes with tcgen05.st.32x32b.x8.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
constexpr int LC_A_COL = 32;             // first TMEM column of the resident W slice (TS variant): 256 columns

// TS = true: the W slice is copied ONCE from shared memory into tensor memory and every step's 32 MMAs read their A
// operand from there.  With both operands in shared memory an MMA costs >= 88 cycles whatever its N (the 128 x 16 A tile
// is streamed through the ~46 B/clk operand port: 2 816 cycles = 1.43 us of the ~3.7 us step); from tensor memory the
// N = 16 MMA is bounded by its tiny B operand only.
template <bool TS>
__global__ void __launch_bounds__(LC_FWD_THREADS, 1) lstm_cluster_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const LstmCArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t wsm = base;
  const uint32_t hsm = base + LC_W_BYTES;                         // 2 buffers
  const uint32_t stg = hsm + 2 * LC_H_BYTES;                      // 2 x 1 KB staging of the own slice
  uint8_t* stg_p = al + LC_W_BYTES + 2 * LC_H_BYTES;
  float* pre_s = reinterpret_cast<float*>(al + LC_W_BYTES + 2 * LC_H_BYTES + 2 * LC_SLICE);
  const uint32_t bar0 = stg + 2 * LC_SLICE + 128 * LC_PRE_LD * 4;
  const uint32_t wbar = bar0, mma_bar = bar0 + 8;
  auto hfull = [&](int b) { return bar0 + 16 + 8u * b; };
  const uint32_t aready = bar0 + 32;           // TS: the W slice is resident in tensor memory
  uint32_t* tptr = reinterpret_cast<uint32_t*>(al + LC_W_BYTES + 2 * LC_H_BYTES + 2 * LC_SLICE + 128 * LC_PRE_LD * 4 + 64);
  // warp roles: 0..15 gate arithmetic (0..3 also read the accumulator), 16 MMA issue -- LAST, because the issue arbiter prefers
  // the highest warp id and the epilogue warps poll their barrier while the MMAs of a step are being issued
  constexpr int W_MMA = LC_FWD_EPI / 32;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t j = cluster_ctarank();
  const int cluster_id = blockIdx.x / LC_NCTA;
  const int b0 = cluster_id * LC_BG;
  const int H = p.H, H4 = 4 * p.H;

  // zero both h operand buffers (h_{-1} = 0; K padding stays 0 because invalid units always send 0)
  for (int i = threadIdx.x; i < 2 * LC_H_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(al + LC_W_BYTES)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmW);
    mbar_init(wbar, 1);
    mbar_init(mma_bar, 1);
    mbar_init(hfull(0), 1);
    mbar_init(hfull(1), 1);
    mbar_init(aready, 128);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(tptr), TS ? 512 : 32);
  fence_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();                     // every CTA's barriers / buffers are ready before any remote traffic
  const uint32_t tm = *tptr;

  if (warp == W_MMA) {
    // Warp-uniform loop, one elected lane issues: inside `if (lane == 0)` ptxas wraps every tcgen05.mma in an ELECT /
    // R2UR.BROADCAST sequence (~20 instructions); the 32 MMAs of a step sit in the recurrence's serial latency chain.
    if (elect_one()) {
      mbar_expect_tx(wbar, LC_W_BYTES);
      for (int kb = 0; kb < 8; ++kb) tma_load_2d(wsm + kb * 16384, &tmW, wbar, kb * 64, (int)j * 128);
    }
    __syncwarp();
    const uint32_t idesc = make_idesc(128, LC_BG, 0, 0);
    const uint64_t ad0 = make_smem_desc(wsm, 16, 1024);
    mbar_wait(TS ? aready : wbar, 0);
    for (int t = 0; t < p.T; ++t) {
      const int pb = t & 1;
      if (t + 1 < p.T && elect_one()) mbar_expect_tx(hfull(pb ^ 1), LC_H_BYTES);        // arm the buffer that receives h_t
      __syncwarp();
      if (t > 0) mbar_wait(hfull(pb), ((t - 1) >> 1) & 1);                // h_{t-1} from all 16 CTAs has landed
      tcgen05_fence_after();
      // un-swizzled K-major operand: LBO = stride between K-adjacent core matrices (256 B), SBO = row groups (128 B)
      const uint64_t bd0 = make_smem_desc(hsm + pb * LC_H_BYTES, 256, 128) & ~((uint64_t)7 << 61);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < LC_KP / 16; ++k) {
          const uint64_t bd = bd0 + (uint64_t)(k * 32);                    // + k * 512 bytes
          if (TS) umma_bf16_ts(tm, tm + LC_A_COL + 8 * k, bd, idesc, k != 0);
          else umma_bf16(tm, ad0 + (uint64_t)((k >> 2) * 1024 + (k & 3) * 2), bd, idesc, k != 0);
        }
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
  } else {
    // 16 epilogue warps: warps 0-3 additionally read the accumulator (their TMEM lane quadrant = gate index, rows are
    // gate-major); every thread owns ONE cell (unit ul, utterance bl), so the gate arithmetic of a step -- three sigmoids
    // and two tanh per cell, part of the step's serial latency chain -- is 4x shorter than with 4 cells per thread.
    const int q = warp & 3;
    const bool reader = warp < 4;
    const int etid = threadIdx.x;                 // 0..511
    const int ul = etid & 31, bl = etid >> 5;     // unit ul of this CTA, utterance bl of the cluster's 16
    const int u = (int)j * LC_U + ul;
    const int b = b0 + bl;
    const bool ok = u < H && b < p.B;
    float c_reg = 0.f;
    float gxr[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) gxr[g] = (ok && p.T > 0) ? __ldg(p.gx + ((int64_t)b * p.T) * H4 + g * H + u) : 0.f;
    if (TS && reader) {
      // row r = q*32 + lane of the W slice: shared memory (128B-swizzled K-major, 8 K blocks of 64) -> registers -> the
      // thread's own TMEM lane, 16 K elements (8 columns) at a time
      mbar_wait(wbar, 0);
      const int r = q * 32 + lane;
      const uint8_t* wrow = al + r * 128;
#pragma unroll 4
      for (int k = 0; k < LC_KP / 16; ++k) {
        const uint8_t* kb = wrow + (k >> 2) * 16384;
        const int c0 = (k & 3) * 2;
        const uint4 lo = *reinterpret_cast<const uint4*>(kb + (((c0) ^ (r & 7)) << 4));
        const uint4 hi = *reinterpret_cast<const uint4*>(kb + (((c0 + 1) ^ (r & 7)) << 4));
        const uint32_t regs[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        tmem_st8(tm + ((uint32_t)(q * 32) << 16) + LC_A_COL + 8 * k, regs);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(aready);
    }
    for (int t = 0; t < p.T; ++t) {
      if (reader) {
        mbar_wait(mma_bar, t & 1);
        tcgen05_fence_after();
        float v[16];
        tmem_ld16_nowait(tm + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
        tcgen05_fence_before();
#pragma unroll
        for (int bb = 0; bb < 16; ++bb) pre_s[(q * 32 + lane) * LC_PRE_LD + bb] = v[bb];
      }
      named_bar_sync(1, LC_FWD_EPI);
      uint8_t* hs = stg_p + (t & 1) * LC_SLICE;
      {
        float pre[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) pre[g] = pre_s[(g * 32 + ul) * LC_PRE_LD + bl] + gxr[g];
        const float ig = sigmoid_(pre[0]), fg = sigmoid_(pre[1]), gg = tanhf(pre[2]), og = sigmoid_(pre[3]);
        c_reg = fg * c_reg + ig * gg;
        const float h = ok ? og * tanhf(c_reg) : 0.f;
        // own slice in the destination layout: [chunk = ul/8][row group = bl/8][row = bl%8][elem = ul%8]
        reinterpret_cast<bf16*>(hs)[(((ul >> 3) * 2 + (bl >> 3)) * 8 + (bl & 7)) * 8 + (ul & 7)] = __float2bfloat16(h);
        if (ok) {
          const int64_t ho = (int64_t)b * p.h_bs + (int64_t)t * p.h_rs + u;
          if (p.h_dtype == NBASR_BF16) reinterpret_cast<bf16*>(p.h_seq)[ho] = __float2bfloat16(h);
          else reinterpret_cast<float*>(p.h_seq)[ho] = h;
          if (p.gates) {
            float* gr = p.gates + ((int64_t)b * p.T + t) * H4;
            gr[u] = ig; gr[H + u] = fg; gr[2 * H + u] = gg; gr[3 * H + u] = og;
            p.cstate[((int64_t)b * p.T + t) * H + u] = c_reg;
          }
          if (t + 1 < p.T) {
#pragma unroll
            for (int g = 0; g < 4; ++g) gxr[g] = __ldg(p.gx + ((int64_t)b * p.T + t + 1) * H4 + g * H + u);
          }
        }
      }
      fence_async_smem();
      named_bar_sync(1, LC_FWD_EPI);
      if (etid < LC_NCTA && t + 1 < p.T) {
        // push the 1-KB slice into CTA `etid`'s buffer for step t+1; its mbarrier counts the bytes
        const uint32_t dst = mapa(hsm + ((t & 1) ^ 1) * LC_H_BYTES + j * LC_SLICE, etid);
        const uint32_t bar = mapa(hfull((t & 1) ^ 1), etid);
        dsmem_bulk_copy(dst, stg + (t & 1) * LC_SLICE, LC_SLICE, bar);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                     // nobody leaves while a peer may still write into its shared memory
  if (warp == W_MMA) {
    tcgen05_fence_after();
    tmem_dealloc(tm, TS ? 512 : 32);
  }
}

// ---------------------------------------------------------------- backward (BPTT)
// Same ownership (CTA j: units [32j, 32j+32), its 128 gate rows of W_hh in shared memory).  Per step the CTA turns
// dh_t of its units into the 128 x 16 gate gradients dg (fp32 to global for the weight-gradient GEMMs, bf16 into the
// un-swizzled MMA operand buffer) and computes its PARTIAL contribution to dh_{t-1} of ALL 512 units:
//     partial[k][b] = sum_{own rows r} W_hh[r][k] dg[r][b]   -- 4 M-tiles x 8 tcgen05.mma (M=128 k's, N=16, K=16 rows),
// reading the SAME shared-memory W slice as an MN-major A operand.  The 32 x 16 block that belongs to CTA i's units is
// sent to CTA i (reduce-scatter over distributed shared memory, 1 KB bulk copies, bf16 partials); the receiver sums the
// 16 partials in fp32.  Again the mbarrier byte count is the only inter-CTA synchronisation.
constexpr int LB_DG_BYTES = 128 * LC_BG * 2;          // 4096: dg operand (N=16 x K=128)
constexpr int LB_SEND_BYTES = LC_NCTA * LC_SLICE;     // 16384: one 1-KB block per destination
constexpr int LB_A_COL = 64;             // backward: first TMEM column of the resident W^T tiles (4 m-tiles x 64 columns)
constexpr int LB_SMEM = LC_W_BYTES + LB_DG_BYTES + 2 * LB_SEND_BYTES + 2 * LB_SEND_BYTES + 1024 + 256;

struct LstmCBwdArgs {
  const float* dh_seq;
  int64_t dh_bs, dh_rs;
  const float* gates;
  const float* cstate;
  int T, B, H;
  float* dgx;          // (B, T, 4H) fp32
  bf16* dgx_bf16;      // optional bf16 copy for the tensor-core GEMMs that follow
};

__global__ void __launch_bounds__(LC_FWD_THREADS, 1) lstm_cluster_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const LstmCBwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t wsm = base;
  const uint32_t dsm = base + LC_W_BYTES;                       // dg operand
  const uint32_t snd = dsm + LB_DG_BYTES;                       // 2 x 16 KB send staging
  const uint32_t rcv = snd + 2 * LB_SEND_BYTES;                 // 2 x 16 KB receive buffers [src][l][b]
  uint8_t* dsm_p = al + LC_W_BYTES;
  uint8_t* snd_p = dsm_p + LB_DG_BYTES;
  uint8_t* rcv_p = snd_p + 2 * LB_SEND_BYTES;
  const uint32_t bar0 = rcv + 2 * LB_SEND_BYTES;
  const uint32_t wbar = bar0, mma_bar = bar0 + 8, dg_bar = bar0 + 16;
  auto rfull = [&](int b) { return bar0 + 24 + 8u * b; };
  const uint32_t aready = bar0 + 40;           // the transposed W slice is resident in tensor memory
  uint32_t* tptr = reinterpret_cast<uint32_t*>(rcv_p + 2 * LB_SEND_BYTES + 64);
  constexpr int W_MMA = LC_FWD_EPI / 32;       // warp roles as in the forward kernel: 0..15 cells / drain, 16 MMA issue
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t j = cluster_ctarank();
  const int b0 = (blockIdx.x / LC_NCTA) * LC_BG;
  const int H = p.H, H4 = 4 * p.H;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmW);
    mbar_init(wbar, 1);
    mbar_init(mma_bar, 1);
    mbar_init(dg_bar, LC_FWD_EPI);
    mbar_init(rfull(0), 1);
    mbar_init(rfull(1), 1);
    mbar_init(aready, LC_FWD_EPI);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(tptr), 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();
  const uint32_t tm = *tptr;

  if (warp == W_MMA) {
    if (elect_one()) {
      mbar_expect_tx(wbar, LC_W_BYTES);
      for (int kb = 0; kb < 8; ++kb) tma_load_2d(wsm + kb * 16384, &tmW, wbar, kb * 64, (int)j * 128);
    }
    __syncwarp();
    // A operand = W^T tiles resident in TENSOR MEMORY (copied once by the epilogue warps below): with both operands in shared
    // memory each of the 32 MMAs of a step costs >= 88 cycles (1.4 us of the ~3.6 us step); from tensor memory the N = 16 MMA
    // is bounded by its small B operand only, as in the forward kernel.
    const uint32_t idesc = make_idesc(128, LC_BG, 0, 0);
    const uint64_t bd0 = make_smem_desc(dsm, 256, 128) & ~((uint64_t)7 << 61);
    mbar_wait(aready, 0);
    for (int s = 0; s + 1 < p.T; ++s) {                          // step s handles t = T-1-s; the last step needs no matmul
      if (elect_one()) mbar_expect_tx(rfull(s & 1), LB_SEND_BYTES);             // arm the buffer that receives this step's partials
      __syncwarp();
      mbar_wait(dg_bar, s & 1);                                  // dg operand of this step is in shared memory
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
          for (int k = 0; k < 8; ++k)       // A: 8 columns (16 gate rows) per K step, B: + k * 512 bytes
            umma_bf16_ts(tm + mt * 16, tm + LB_A_COL + mt * 64 + 8 * k, bd0 + (uint64_t)(k * 32), idesc, k != 0);
        }
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
  } else {
    // 16 epilogue warps, one cell (unit ul, utterance bl) per thread; warp w drains ONE of the four accumulator m-tiles
    // (its TMEM lane quadrant is w % 4, its m-tile w / 4) -- both used to be 4 sequential items per thread inside the
    // step's latency chain.
    const int q = warp & 3;
    const int mt_own = warp >> 2;
    const int etid = threadIdx.x;                 // 0..511
    const int ul = etid & 31, bl = etid >> 5;
    const int u = (int)j * LC_U + ul;
    const int b = b0 + bl;
    const bool ok = u < H && b < p.B;
    float dc_next = 0.f, dh_rec = 0.f;
    float nx[7];
    auto prefetch = [&](int t) {
      if (ok) {
        nx[0] = __ldg(p.dh_seq + (int64_t)b * p.dh_bs + (int64_t)t * p.dh_rs + u);
        const float* gr = p.gates + ((int64_t)b * p.T + t) * H4;
        nx[1] = __ldg(gr + u); nx[2] = __ldg(gr + H + u); nx[3] = __ldg(gr + 2 * H + u); nx[4] = __ldg(gr + 3 * H + u);
        nx[5] = __ldg(p.cstate + ((int64_t)b * p.T + t) * H + u);
        nx[6] = t > 0 ? __ldg(p.cstate + ((int64_t)b * p.T + t - 1) * H + u) : 0.f;
      } else {
#pragma unroll
        for (int z = 0; z < 7; ++z) nx[z] = 0.f;
      }
    };
    if (p.T > 0) prefetch(p.T - 1);
    {
      // W^T into tensor memory, once: lane (q*32 + lane) of m-tile mt_own is hidden unit kcol; its 128 K elements are the own
      // gate rows r of W_hh[r][kcol], read column-wise from the 128B-swizzled K-major slice and packed two per 32-bit column
      mbar_wait(wbar, 0);
      const int kcol = mt_own * 128 + q * 32 + lane;
      const uint8_t* wcol = al + (kcol >> 6) * 16384 + (kcol & 7) * 2;
      const int chunk = (kcol & 63) >> 3;
#pragma unroll 2
      for (int k = 0; k < 8; ++k) {
        uint32_t regs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r0 = 16 * k + 2 * i, r1 = r0 + 1;
          const uint32_t lo = *reinterpret_cast<const uint16_t*>(wcol + r0 * 128 + ((chunk ^ (r0 & 7)) << 4));
          const uint32_t hi = *reinterpret_cast<const uint16_t*>(wcol + r1 * 128 + ((chunk ^ (r1 & 7)) << 4));
          regs[i] = lo | (hi << 16);
        }
        tmem_st8(tm + ((uint32_t)(q * 32) << 16) + LB_A_COL + mt_own * 64 + 8 * k, regs);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(aready);
    }
    for (int s = 0; s < p.T; ++s) {
      const int t = p.T - 1 - s;
      if (s > 0) {
        // partial sums of dh_t for the own units from all 16 CTAs (sent during step s-1)
        mbar_wait(rfull((s - 1) & 1), ((s - 1) >> 1) & 1);
        const bf16* rb = reinterpret_cast<const bf16*>(rcv_p + ((s - 1) & 1) * LB_SEND_BYTES);
        float acc = 0.f;
#pragma unroll
        for (int src = 0; src < LC_NCTA; ++src) acc += __bfloat162float(rb[(src * 32 + ul) * LC_BG + bl]);
        dh_rec = acc;
      }
      bf16* dgo = reinterpret_cast<bf16*>(dsm_p);
      {
        const float dh = nx[0] + dh_rec;
        const float ig = nx[1], fg = nx[2], gg = nx[3], og = nx[4], c = nx[5], cprev = nx[6];
        const float tc = tanhf(c);
        const float dov = dh * tc;
        const float dc = dc_next + dh * og * (1.f - tc * tc);
        dc_next = dc * fg;
        float d[4];
        d[0] = dc * gg * ig * (1.f - ig);
        d[1] = dc * cprev * fg * (1.f - fg);
        d[2] = dc * ig * (1.f - gg * gg);
        d[3] = dov * og * (1.f - og);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float v = ok ? d[g] : 0.f;
          const int r = g * 32 + ul;                         // own gate row = K index of the MMA
          dgo[(((r >> 3) * 2 + (bl >> 3)) * 8 + (bl & 7)) * 8 + (r & 7)] = __float2bfloat16(v);
          if (ok) {
            const int64_t o = ((int64_t)b * p.T + t) * H4 + g * H + u;
            p.dgx[o] = v;
            if (p.dgx_bf16) p.dgx_bf16[o] = __float2bfloat16(v);
          }
        }
      }
      if (t == 0) break;
      prefetch(t - 1);
      fence_async_smem();
      mbar_arrive(dg_bar);                                   // 512 arrivals -> the MMA thread may read the dg operand
      mbar_wait(mma_bar, s & 1);
      tcgen05_fence_after();
      uint8_t* sb = snd_p + (s & 1) * LB_SEND_BYTES;
      {
        float v[16];
        tmem_ld16_nowait(tm + ((uint32_t)(q * 32) << 16) + mt_own * 16, v);
        tmem_ld_wait();
        // lane l of quadrant q holds k = mt*128 + q*32 + l -> unit l of CTA (4*mt + q)
        bf16* dst = reinterpret_cast<bf16*>(sb + (4 * mt_own + q) * LC_SLICE) + lane * LC_BG;
        uint4 pk[2];
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
        for (int x = 0; x < 8; ++x) h2[x] = __floats2bfloat162_rn(v[2 * x], v[2 * x + 1]);
        reinterpret_cast<uint4*>(dst)[0] = pk[0];
        reinterpret_cast<uint4*>(dst)[1] = pk[1];
      }
      tcgen05_fence_before();
      fence_async_smem();
      named_bar_sync(1, LC_FWD_EPI);
      if (etid < LC_NCTA) {
        const uint32_t dst = mapa(rcv + (s & 1) * LB_SEND_BYTES + j * LC_SLICE, etid);
        const uint32_t bar = mapa(rfull(s & 1), etid);
        dsmem_bulk_copy(dst, snd + (s & 1) * LB_SEND_BYTES + etid * LC_SLICE, LC_SLICE, bar);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == W_MMA) {
    tcgen05_fence_after();
    tmem_dealloc(tm, 512);
  }
}

}  // namespace

// w_packed: [16][128 rows = gate*32 + unit][512] bf16, made by nbasr_pack_batch kind 4
int sm100_lstm_fwd(const float* gx, const void* w_packed, int T, int B, int H, void* h_seq, int h_dtype, int64_t h_bs, int64_t h_rs,
                   float* gates, float* cstate, cudaStream_t st) {
  NBASR_REQUIRE(H <= LC_KP && H <= LC_NCTA * LC_U, "hidden size");
  LstmCArgs a{gx, T, B, H, h_seq, h_dtype, h_bs, h_rs, gates, cstate};
  CUtensorMap tmW;
  uint64_t dw[2] = {LC_KP, (uint64_t)LC_NCTA * 128};
  int64_t sw[2] = {1, LC_KP};
  uint32_t bw[2] = {64, 128};
  if (sm100_get_map(w_packed, 2, dw, sw, bw, &tmW)) return 1;
  const bool ts = !nbasr_env_flag(NBASR_ENV_LSTM_SS);      // default: W slice resident in tensor memory
  static DevOnce attr;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lstm_cluster_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LC_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_cluster_fwd_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_cluster_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LC_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_cluster_fwd_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return nbasr_fail("lstm_cluster attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int nclusters = (B + LC_BG - 1) / LC_BG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * LC_NCTA);
  cfg.blockDim = dim3(LC_FWD_THREADS);
  cfg.dynamicSmemBytes = LC_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = LC_NCTA;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = ts ? cudaLaunchKernelEx(&cfg, lstm_cluster_fwd_kernel<true>, tmW, a)
                     : cudaLaunchKernelEx(&cfg, lstm_cluster_fwd_kernel<false>, tmW, a);
  if (e != cudaSuccess) return nbasr_fail("lstm_cluster_fwd launch: %s", cudaGetErrorString(e));
  return 0;
}

int sm100_lstm_bwd(const float* dh_seq, int64_t dh_bs, int64_t dh_rs, const void* w_packed, const float* gates, const float* cstate,
                   int T, int B, int H, float* dgx, void* dgx_bf16, cudaStream_t st) {
  NBASR_REQUIRE(H <= LC_KP && H <= LC_NCTA * LC_U, "hidden size");
  LstmCBwdArgs a{dh_seq, dh_bs, dh_rs, gates, cstate, T, B, H, dgx, reinterpret_cast<bf16*>(dgx_bf16)};
  CUtensorMap tmW;
  uint64_t dw[2] = {LC_KP, (uint64_t)LC_NCTA * 128};
  int64_t sw[2] = {1, LC_KP};
  uint32_t bw[2] = {64, 128};
  if (sm100_get_map(w_packed, 2, dw, sw, bw, &tmW)) return 1;
  static DevOnce attr;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lstm_cluster_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_cluster_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return nbasr_fail("lstm_cluster_bwd attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int nclusters = (B + LC_BG - 1) / LC_BG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * LC_NCTA);
  cfg.blockDim = dim3(LC_FWD_THREADS);
  cfg.dynamicSmemBytes = LB_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = LC_NCTA;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_cluster_bwd_kernel, tmW, a);
  if (e != cudaSuccess) return nbasr_fail("lstm_cluster_bwd launch: %s", cudaGetErrorString(e));
  return 0;
}
