// Grouped conv edges, forward and input-gradient (ops.py:73-76, groups=100) on tcgen05 -- tap-merged version.
//
// Measured on B200 (tools/dbg_bench.py): one CTA thread can retire at most one tcgen05.mma per ~88 cycles per SM,
// whatever N <= 176 is (N=192: 96, N=240: 120, N=256: 128 cycles).  The per-tap block-diagonal kernel
// (gconv_sm100.cu: 128 x 48 x 16 MMAs, 3 per tap and tile) therefore ran exactly at that issue floor
// (15 x 88 cycles per 128 x 48 output tile) and far below both HBM and tensor peak.  This kernel makes every MMA
// instruction carry m taps:
//
//   * B operand = the packed block-diagonal weights of m consecutive taps stacked along N (the pack layout
//     [slab][tap][48][64] is already contiguous in taps), so one 128 x (48 m) x 16 MMA reads the activation tile
//     ONCE and produces m partial products P_b[t][co] = sum_ci X[t + off0 + (g m) d][ci] W_{g m + b}[co][ci]
//     in m column blocks of the accumulator; tap groups g accumulate into the same blocks with the A start
//     address advanced by whole 128-byte rows (g m d frames).
//   * the remaining shift is done by the epilogue: Y[t] = sum_b P_b[t + b d] -- a warp shuffle by b d lanes
//     (TMEM lane = frame), with the first (m-1) d rows of the next warp's quadrant exchanged through shared
//     memory (one named barrier per tile).  A tile therefore yields 128 - (m-1) d valid frames and tiles step by that amount.
//   * one persistent CTA per SM owns a PAIR of adjacent slabs (both weight packs resident) and walks frame tiles
//     lane, lane + L, ...; CTAs of different slab pairs work on the same frames at the same time, so the 128-byte
//     column slices they fetch are neighbours in DRAM / L2.  warp 0 = TMA producer, warp 1 = MMA issuer, two
//     epilogue groups of 8 warps each own whole tiles (own staging buffer, own named barrier): a group pulls all
//     m x 48 partial columns into registers at once and releases the TMEM stage before it does any arithmetic.
//   * epilogue (bias, ReLU20, dropout, skip-sum, gate bits, masked second output) as in the per-tap kernel.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int AROWS = 144;                 // 128 + tap reach, multiple of 8
constexpr int A_BYTES = AROWS * 128;       // 18432
constexpr int NW = 48;                     // accumulator block width / weight rows per tap
constexpr int WTAP_BYTES = NW * 128;       // 6144
constexpr int OSTAGE_BYTES = 128 * NW * 2; // 12288
constexpr int XB_BYTES = 6144;             // [4 quadrants][2 halves][2 blocks][4 rows][24 floats]
constexpr int MST_BYTES = 1024;            // 128 x 8-byte gate-bit entries
constexpr int MAXM_ALL = 3;
constexpr int MAXG = 2;
constexpr int THREADS = 64 + 256 * MAXG;   // 576
constexpr int SMEM_LIMIT = 225 * 1024;

struct Args {
  int B, T, C, OUT, ktaps, dstep, off0;
  int m, ngroups, GTV;
  int nslabs, npairs, nlanes, ntiles, tiles_per_utt;
  int nstage, nacc, accstride, G, has_out2;
  nbasr_epilogue epi;
  int64_t Tp;
  unsigned long long* dbg;   // optional timeline dump (tools/trace_gconv.py): [cta < 8][item < 64][8] globaltimer ns
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define GC_STAMP(slot)                                                                         \
  do {                                                                                         \
    if (p.dbg && blockIdx.x < 8 && li < 64) p.dbg[((size_t)blockIdx.x * 64 + li) * 8 + (slot)] = gtime(); \
  } while (0)

template <int MAXM>
__global__ void __launch_bounds__(THREADS, 1)
gconv_fwd_merged_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                        const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2, const Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
  const int NS = p.nstage, NACC = p.nacc;
  const uint32_t wsz = p.ktaps * WTAP_BYTES;
  const uint32_t gbytes = OSTAGE_BYTES * (1 + p.has_out2) + XB_BYTES + MST_BYTES;
  const uint32_t w0 = base;
  const uint32_t a0 = w0 + 2 * wsz;
  const uint32_t g0 = a0 + NS * A_BYTES;
  const uint32_t bar0 = g0 + p.G * gbytes;
  const uint32_t wfull = bar0;
  auto full_bar = [&](int s) { return bar0 + 8u * (4 + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (12 + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (20 + s); };
  uint32_t* tptr = reinterpret_cast<uint32_t*>(al + (bar0 - base) + 8 * 28);
  float* sbias = reinterpret_cast<float*>(al + (bar0 - base) + 256);      // [2 slabs][48]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work list of this CTA: tiles lane, lane + nlanes, ... x the (1 or 2) slabs of its pair; item li -> (li / nsl, li % nsl)
  const int pair = blockIdx.x % p.npairs, lane_id = blockIdx.x / p.npairs;
  const int slab_first = 2 * pair;
  const int nsl = min(2, p.nslabs - slab_first);
  const int ntl = lane_id < p.ntiles ? (p.ntiles - lane_id + p.nlanes - 1) / p.nlanes : 0;
  const int nli = ntl * nsl;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    mbar_init(wfull, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < NACC; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 256); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tptr), 512);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * NW) {      // bias of the CTA's (up to) two slabs
    const int i = threadIdx.x - 64, c = (slab_first + i / NW) * p.OUT + i % NW;
    sbias[i] = (p.epi.bias && i / NW < nsl && i % NW < p.OUT && c < p.C) ? __ldg(p.epi.bias + c) : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = *tptr;

  if (warp == 0) {
    if (lane == 0 && nli > 0) {
      mbar_expect_tx(wfull, nsl * wsz);
      for (int so = 0; so < nsl; ++so)
        for (int j = 0; j < p.ktaps; ++j)
          tma_load_2d(w0 + so * wsz + j * WTAP_BYTES, &tmW, wfull, 0, ((slab_first + so) * p.ktaps + j) * NW);
      int stage = 0;
      uint32_t phase = 0;
      for (int li = 0; li < nli; ++li) {
        const int tile = lane_id + (li / nsl) * p.nlanes, slab = slab_first + li % nsl;
        const int b = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * p.GTV;
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(full_bar(stage), A_BYTES);
        tma_load_3d(a0 + stage * A_BYTES, &tmX, full_bar(stage), slab * p.OUT, NBASR_PAD_L + t0 + p.off0, b);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nli > 0) {
      const int mlast = p.ktaps - (p.ngroups - 1) * p.m;
      const uint32_t idesc_full = make_idesc(128, NW * p.m, 0, 0);
      const uint32_t idesc_last = make_idesc(128, NW * mlast, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(wfull, 0);
      for (int li = 0; li < nli; ++li) {
        const int so = li % nsl;
        const int as = li % NACC;
        const uint32_t aphase = (li / NACC) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        mbar_wait(full_bar(stage), phase);
        GC_STAMP(1);
        tcgen05_fence_after();
        const uint32_t sa = a0 + stage * A_BYTES;
        const uint32_t sw = w0 + so * wsz;
        for (int g = 0; g < p.ngroups; ++g) {
          const uint32_t idesc = (g == p.ngroups - 1) ? idesc_last : idesc_full;
#pragma unroll
          for (int k = 0; k < NW / 16; ++k) {
            uint64_t ad = make_smem_desc(sa + (g * p.m * p.dstep) * 128 + k * 32, 16, 1024);
            uint64_t bd = make_smem_desc(sw + (g * p.m) * WTAP_BYTES + k * 32, 16, 1024);
            umma_bf16(tm + as * p.accstride, ad, bd, idesc, (g | k) != 0);
          }
        }
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(as));
        GC_STAMP(2);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 2 + 8 * p.G) {
    // Epilogue group = 8 warps: the pair (w, w + 4) shares a TMEM lane quadrant and splits the 48 columns in halves of 24.
    const int ew = warp - 2;
    const int grp = ew >> 3;
    const int q = warp & 3;                   // TMEM lane quadrant this warp may read
    const int hh = (ew & 7) >> 2;             // column half
    const int gtid = threadIdx.x - 64 - grp * 256;
    const int row = q * 32 + lane;
    uint8_t* gsm = al + (g0 - base) + grp * gbytes;
    uint8_t* ost = gsm;
    uint8_t* ost2 = gsm + OSTAGE_BYTES;
    float* xb = reinterpret_cast<float*>(gsm + OSTAGE_BYTES * (1 + p.has_out2));
    uint8_t* mst = gsm + OSTAGE_BYTES * (1 + p.has_out2) + XB_BYTES;      // 128 x 8-byte gate-bit entries
    const uint32_t osm = g0 + grp * gbytes;
    const nbasr_epilogue& epi = p.epi;
    const int OUTB = p.OUT * 2;
    const int m = p.m, d = p.dstep;
    const bool lean = epi.drop_p == 0.f && epi.n_add == 0;
    for (int li = grp; li < nli; li += p.G) {
      const int tile = lane_id + (li / nsl) * p.nlanes, so = li % nsl, slab = slab_first + so;
      const int as = li % NACC;
      const uint32_t aphase = (li / NACC) & 1;
      const int b = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * p.GTV;
      const int t = t0 + row;
      const bool rowok = row < p.GTV && t < p.T;
      const int64_t rho = (int64_t)b * p.Tp + NBASR_PAD_L + t;
      const int c0 = slab * p.OUT;
      const int cbeg = c0 + 24 * hh;
      const int nvalid = max(0, min(24, min(p.C, c0 + p.OUT) - cbeg));     // multiple of 8
      const bool m2slab = epi.mask2_w == p.OUT;      // mask2 written by a grouped-conv node: one 8-byte entry per (slab, row)
      uint64_t m2bits = ~0ull;
      if (epi.out2 && epi.mask2 && rowok && m2slab)
        m2bits = reinterpret_cast<const uint64_t*>(epi.mask2)[(int64_t)slab * epi.mask_rows + rho] >> (24 * hh);
      mbar_wait(tfull_bar(as), aphase);
      if (gtid == 0) GC_STAMP(3);
      tcgen05_fence_after();
      const uint32_t ta = tm + ((uint32_t)(q * 32) << 16) + as * p.accstride + 24 * hh;
      float P[MAXM][24];
#pragma unroll
      for (int b2 = 0; b2 < MAXM; ++b2) {
        if (b2 < m) {
          tmem_ld16_nowait(ta + b2 * NW, P[b2]);
          tmem_ld8_nowait(ta + b2 * NW + 16, P[b2] + 16);
        }
      }
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(as));             // all partial products are in registers: release the TMEM stage
      if (gtid == 0) GC_STAMP(4);
      // publish the first b2*d rows of this quadrant for the warps that own the previous 32 rows
#pragma unroll
      for (int b2 = 1; b2 < MAXM; ++b2) {
        if (b2 < m && lane < b2 * d) {
          float4* dst = reinterpret_cast<float4*>(xb + ((((q * 2 + hh) * 2 + (b2 - 1)) * 4 + lane) * 24));
#pragma unroll
          for (int i = 0; i < 6; ++i) dst[i] = make_float4(P[b2][4 * i], P[b2][4 * i + 1], P[b2][4 * i + 2], P[b2][4 * i + 3]);
        }
      }
      if (gtid == 0) bulk_wait_read0();        // staging buffers free: previous tile's TMA stores have read them
      named_bar_sync(1 + grp, 256);
      if (gtid == 0) GC_STAMP(5);
      // Y[t] = sum_b P_b[t + b d]: lanes < b d take over the rows published by the next quadrant, then one rotation
#pragma unroll
      for (int b2 = 1; b2 < MAXM; ++b2) {
        if (b2 < m) {
          const int s = b2 * d;
          if (lane < s) {
            const float4* src = reinterpret_cast<const float4*>(xb + ((((((q + 1) & 3) * 2 + hh) * 2 + (b2 - 1)) * 4 + lane) * 24));
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              const float4 f = src[i];
              P[b2][4 * i] = f.x; P[b2][4 * i + 1] = f.y; P[b2][4 * i + 2] = f.z; P[b2][4 * i + 3] = f.w;
            }
          }
          const int srcl = (lane + s) & 31;
#pragma unroll
          for (int i = 0; i < 24; ++i) P[0][i] += __shfl_sync(0xffffffffu, P[b2][i], srcl);
        }
      }
      float* v = P[0];
      uint32_t mg[3] = {0, 0, 0};
      if (rowok && nvalid > 0) {
        if (epi.bias) {
          const float4* bs = reinterpret_cast<const float4*>(sbias + so * NW + 24 * hh);
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float4 bb = bs[i];
            v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
          }
        }
        if (lean && nvalid == 24) {
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            uint32_t mm = 0xffu;
            if (epi.relu20) {
              mm = 0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float z = v[g * 8 + i];
                // 0 < z <= 20  <=>  bits(z) - 1 < bits(20.0f) as unsigned (negative z and +0 wrap to huge values)
                mm |= ((__float_as_uint(z) - 1u) < 0x41A00000u) ? (1u << i) : 0u;
                v[g * 8 + i] = fminf(fmaxf(z, 0.f), 20.f);
              }
            }
            mg[g] = mm;
          }
        } else if (nvalid == 24) {
          epilogue_compute<24, true, true>(epi, rho, cbeg, 24, v, mg);
        } else {
          epilogue_compute<24, false, true>(epi, rho, cbeg, nvalid, v, mg);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 24; ++i) v[i] = 0.f;      // rows past the utterance land on zero pad rows / are clipped
      }
      uint8_t* orow = ost + row * OUTB + 48 * hh;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        if (g * 8 < nvalid) {
          if (epi.out) store8(reinterpret_cast<bf16*>(orow + g * 16), v + g * 8);
          if (epi.out2) {
            uint32_t w = (uint32_t)(m2bits >> (8 * g)) & 0xffu;
            if (epi.mask2 && !m2slab && rowok)
              w = reinterpret_cast<const uint8_t*>(epi.mask2)[mask_byte_addr(rho, cbeg + g * 8, epi.mask2_w, epi.mask_rows)];
            float t2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t2[i] = ((w >> i) & 1u) ? v[g * 8 + i] * epi.scale2 : 0.f;
            store8(reinterpret_cast<bf16*>(ost2 + row * OUTB + 48 * hh + g * 16), t2);
          }
        }
        if (epi.mask_out) mst[row * 8 + 3 * hh + g] = (g * 8 < nvalid && rowok) ? (uint8_t)mg[g] : (uint8_t)0;
      }
      if (gtid == 0) GC_STAMP(6);
      fence_async_smem();
      named_bar_sync(1 + grp, 256);
      if (gtid == 0) GC_STAMP(0);
      if (epi.mask_out && gtid < p.GTV && t0 + gtid < p.T) {
        // 8-byte entries of consecutive rows of this slab's mask plane: one fully coalesced store per warp
        const int64_t r2 = (int64_t)b * p.Tp + NBASR_PAD_L + t0 + gtid;
        reinterpret_cast<uint64_t*>(epi.mask_out)[(int64_t)slab * epi.mask_rows + r2] = reinterpret_cast<const uint64_t*>(mst)[gtid];
      }
      if (gtid == 0) {
        if (epi.out) tma_store_3d(&tmO, osm, c0, NBASR_PAD_L + t0, b);
        if (epi.out2) tma_store_3d(&tmO2, osm + OSTAGE_BYTES, c0, NBASR_PAD_L + t0, b);
        bulk_commit();
        GC_STAMP(7);
      }
    }
    if (gtid == 0) bulk_wait0();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tm, 512);
  }
}

}  // namespace

int sm100_gconv_fwd_v1(const nbasr_gconv* g, cudaStream_t st);   // per-tap kernel (gconv_sm100.cu)
int sm100_gconv_fwd_frag(const nbasr_gconv* g, cudaStream_t st); // mma.sync kernel (gconv_frag_sm100.cu)
extern unsigned long long* g_gconv_dbg;
int g_gconv_impl = 0;   // 0 = tcgen05 per-tap kernel (default), 1 = warp-level MMA kernel
extern "C" void nbasr_dbg_gconv_impl(int impl) { g_gconv_impl = impl; }

int sm100_gconv_fwd(const nbasr_gconv* g, cudaStream_t st) {
  // Default = the per-tap kernel.  The tap-merged kernel below is numerically identical (same tests) and cuts the MMA
  // count 1.7-2.5x, but its epilogue (shuffle-sum + ~600-750 issued instructions per warp and tile, ncu: 2.45 IPC,
  // issue-bound) ends up slower on B200: C=1000 k=5 cold-L2 72.6 us (m=2) / 108 us (m=3) vs 56 us.  NBASR_GCONV_MERGED=1
  // selects it for experiments (tools/bench_gconv.py, tools/trace_gconv.py).
  static const char* merged = getenv("NBASR_GCONV_MERGED");
  static const char* frag = getenv("NBASR_GCONV_FRAG");
  // Warp-level MMA kernel (gconv_frag_sm100.cu): correct (same tests) but instruction-issue bound on B200 -- ~6 600 issued
  // warp instructions per 128 x 48 tile against 4 400 issue slots at HBM speed (ncu, profiles/r1_gconv_frag_experiment.txt):
  // 75-99 us vs 55 us for C=1000 k=5.  Experimental: NBASR_GCONV_FRAG=1 or nbasr_dbg_gconv_impl(1).
  if ((frag || g_gconv_impl == 1) && (g->ktaps == 5 || g->ktaps == 7) && g->cpg <= 16 && (g->cpg == 10 ? 40 : 48) % g->cpg == 0 && g->C % 8 == 0)
    return sm100_gconv_fwd_frag(g, st);
  if (!merged) return sm100_gconv_fwd_v1(g, st);
  Args a{};
  a.B = g->B; a.T = g->T; a.Tp = g->Tp; a.C = g->C; a.OUT = g->cpg == 10 ? 40 : 48;
  a.ktaps = g->ktaps; a.dstep = g->dstep; a.off0 = g->off0;
  static const char* env_m = getenv("NBASR_GCONV_M");
  a.m = env_m ? atoi(env_m) : (g->dstep == 1 ? 3 : 2);
  a.m = std::max(1, std::min(a.m, std::min(MAXM_ALL, a.ktaps)));
  while (a.m > 1 && (a.m - 1) * a.dstep > 4) --a.m;
  a.ngroups = (a.ktaps + a.m - 1) / a.m;
  a.GTV = 128 - (a.m - 1) * a.dstep;
  NBASR_REQUIRE(a.ktaps <= 7 && a.off0 >= -NBASR_PAD_L && (a.ktaps - 1) * a.dstep <= AROWS - 128, "tap reach");
  NBASR_REQUIRE(g->epi.ld_out == g->C, "grouped conv writes dense (B,Tp,C) tensors");
  NBASR_REQUIRE((!g->epi.out || g->epi.out_dtype == NBASR_BF16) && (!g->epi.out2 || g->epi.out2_dtype == NBASR_BF16) &&
                    !g->epi.accumulate, "tcgen05 grouped conv stores bf16");
  NBASR_REQUIRE(!g->epi.mask_out || g->epi.mask_w == a.OUT, "grouped-conv mask planes are slab wide");
  a.nslabs = (g->C + a.OUT - 1) / a.OUT;
  a.tiles_per_utt = (g->T + a.GTV - 1) / a.GTV;
  a.ntiles = a.tiles_per_utt * g->B;
  a.npairs = (a.nslabs + 1) / 2;
  a.nlanes = std::max(1, std::min(a.ntiles, nbasr_sm_count() / a.npairs));
  a.has_out2 = g->epi.out2 ? 1 : 0;
  a.accstride = a.m == 1 ? 64 : a.m == 2 ? 128 : a.m == 3 ? 160 : 256;
  a.nacc = std::min(8, 512 / a.accstride);
  static const char* env_g = getenv("NBASR_GCONV_G");
  a.G = env_g ? std::max(1, std::min(MAXG, atoi(env_g))) : MAXG;
  const int fixed = 2 * a.ktaps * WTAP_BYTES + 1024 /*align*/ + 1024 /*barriers, bias*/;
  auto gb = [&](int G) { return G * (OSTAGE_BYTES * (1 + a.has_out2) + XB_BYTES + MST_BYTES); };
  a.nstage = std::min(4, (SMEM_LIMIT - fixed - gb(a.G)) / A_BYTES);
  NBASR_REQUIRE(a.nstage >= 2, "shared memory budget");
  a.epi = g->epi;
  a.dbg = g_gconv_dbg;
  CUtensorMap tmX, tmW, tmO, tmO2;
  uint64_t dx[3] = {(uint64_t)g->C, (uint64_t)g->Tp, (uint64_t)g->B};
  int64_t sx[3] = {1, g->C, (int64_t)g->Tp * g->C};
  uint32_t bx[3] = {64, AROWS, 1};
  if (sm100_get_map(g->x, 3, dx, sx, bx, &tmX)) return 1;
  uint64_t dw[2] = {64, (uint64_t)a.nslabs * a.ktaps * NW};
  int64_t sw[2] = {1, 64};
  uint32_t bw[2] = {64, NW};
  if (sm100_get_map(g->w, 2, dw, sw, bw, &tmW)) return 1;
  uint32_t bo[3] = {(uint32_t)a.OUT, (uint32_t)a.GTV, 1};
  const void* o1 = g->epi.out ? g->epi.out : g->x;       // unused maps still need a valid descriptor
  const void* o2 = g->epi.out2 ? g->epi.out2 : g->x;
  if (sm100_get_map(o1, 3, dx, sx, bo, &tmO, 0)) return 1;
  if (sm100_get_map(o2, 3, dx, sx, bo, &tmO2, 0)) return 1;
  size_t smem = (size_t)fixed + gb(a.G) + (size_t)a.nstage * A_BYTES;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gconv_fwd_merged_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gconv_fwd_merged_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return nbasr_fail("gconv_fwd_merged smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  int grid = a.npairs * a.nlanes;
  if (a.m <= 2) gconv_fwd_merged_kernel<2><<<grid, THREADS, smem, st>>>(tmX, tmW, tmO, tmO2, a);
  else gconv_fwd_merged_kernel<3><<<grid, THREADS, smem, st>>>(tmX, tmW, tmO, tmO2, a);
  NBASR_CHECK_LAUNCH();
  return 0;
}
