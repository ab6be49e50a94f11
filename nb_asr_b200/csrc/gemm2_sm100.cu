// Dense GEMM  C[(b,r), n] = sum_k A[(b,r), k] W[n, k]  on CTA PAIRS (tcgen05 cta_group::2): the time-reduction
// convolutions, `linear` edges, the LSTM input projection and their input-gradients (see include/nbasr.h).
//
// Why pairs: measured on B200 (tools/bench_gemm.py) the 1-CTA kernel (gemm_sm100.cu, 128 x 256 tile) needs ~750 cycles
// per 64-deep K block although its four MMAs take 512 -- an SM ingests ~64 B/clk from L2 and the tile needs
// 16 KB (A) + 32 KB (B) per K block.  With cta_group::2 two SMs of a TPC share one 256 x BN MMA: each CTA loads ITS
// 128 rows of A and only HALF of the B tile (BN/2 rows), so ingest drops to 16 + 16 KB per K block (= 512 cycles at
// 64 B/clk, the MMA time), and a stage is 32 KB, so the TMA ring is 6 deep.
//
// Structure per CTA (same warp roles as the 1-CTA kernel): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA
// only), 8 epilogue warps.  Synchronisation:
//   full[s]   lives in the LEADER: both producers arrive.expect_tx on it (count 2) and both CTAs' TMA loads
//             complete_tx on it (.cta_group::2 lets a load signal the peer's barrier);
//   empty[s]  in each CTA, arrived by the leader's tcgen05.commit multicast to both CTAs;
//   tfull[a]  in each CTA, same multicast commit; each CTA's epilogue drains ITS half (TMEM lanes = its 128 rows);
//   tempty[a] in the leader: one arrive per epilogue warp of both CTAs (count 16, the peer arrives remotely).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int BM = 128;                 // rows per CTA (256 per pair)
constexpr int BK = 64;
constexpr int STAGES = 6;
constexpr int A_STAGE_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_STAGE_BYTES = 128 * BK * 2;         // up to BN/2 = 128 rows: 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int THREADS = 320;

struct Args {
  int nb, nr, K, N, BN;
  int mt_per_utt, m_tiles, n_tiles, pair_tiles;
  int64_t o_r0, o_bs, o_rs;
  nbasr_epilogue epi;
};

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
// arrive (+ expect_tx) on a barrier that may live in the peer CTA (shared::cluster address)
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster, uint32_t bytes) {
  // .relaxed: a cluster-scope release here compiles to MEMBAR + ERRBAR, which waits for the producer's outstanding TMA
  // loads and serialises the whole ring (ncu: ~1 us per K block); the data itself is published by complete_tx.
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// TMA loads whose completion is signalled on a barrier of either CTA of the pair
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs -> one arrive on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_al + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_kb = (p.K + BK - 1) / BK;
  const int half_bn = p.BN >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);       // one arrive.expect_tx per CTA of the pair (used in the leader only)
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 16);    // 8 epilogue warps x 2 CTAs (used in the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(smem_u32(tmem_ptr_smem), 512);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                  // both CTAs' barriers are initialised before any remote arrive / multicast
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = A_STAGE_BYTES + half_bn * BK * 2;
      for (int pt = pair_id; pt < p.pair_tiles; pt += npairs) {
        const int n_idx = pt % p.n_tiles;
        const int m_idx = min(2 * (pt / p.n_tiles) + (int)rank, p.m_tiles - 1);   // odd tile count: the peer repeats the last tile
        const int b = m_idx / p.mt_per_utt, r0 = (m_idx % p.mt_per_utt) * BM;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          const uint32_t fb = mapa(full_bar(stage), 0);
          mbar_expect_tx_cluster(fb, stage_tx);
          tma2_load_3d(sa, &tmA, fb, kb * BK, r0, b);
          tma2_load_2d(sb, &tmB, fb, kb * BK, n_idx * p.BN + (int)rank * half_bn);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(2 * BM, p.BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int pt = pair_id; pt < p.pair_tiles; pt += npairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t ad = make_smem_desc(sa + k * 32, 16, 1024);
            uint64_t bd = make_smem_desc(sb + k * 32, 16, 1024);
            umma2_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma2_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma2_commit(tfull_bar(as));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;           // TMEM lane quadrant this warp may access
    const int hh = (warp - 2) >> 2;   // the two warps of a quadrant take alternate 32-column chunks
    int it = 0;
    for (int pt = pair_id; pt < p.pair_tiles; pt += npairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int n_idx = pt % p.n_tiles;
      const int m_raw = 2 * (pt / p.n_tiles) + (int)rank;
      const int m_idx = min(m_raw, p.m_tiles - 1);
      const int b = m_idx / p.mt_per_utt, r0 = (m_idx % p.mt_per_utt) * BM;
      const int r = r0 + q * 32 + lane;
      const int n0 = n_idx * p.BN;
      const int ncol = min(p.N, n0 + p.BN);
      mbar_wait(tfull_bar(as), aphase);
      tcgen05_fence_after();
      const int64_t rho = p.o_r0 + (int64_t)b * p.o_bs + (int64_t)r * p.o_rs;
      for (int c = 32 * hh; c < p.BN; c += 64) {
        if (n0 + c >= ncol) break;   // warp-uniform
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * 256 + c, v);
        if (r < p.nr && m_raw < p.m_tiles) epilogue_chunk(p.epi, rho, n0 + c, ncol, v);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(tempty_bar(as), 0));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                  // nobody leaves (or frees TMEM) while the pair's MMAs / multicasts may still touch it
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// waves x cycles per K block (see gemm_sm100.cu pick_bn), for the paired tile 256 x BN: four MMAs of max(88, BN/2)
// cycles against 16 KB of A + 64 BN bytes of B ingest per CTA
int pick_bn_pair(int N, int m_tiles, int pairs) {
  int best = 256;
  long best_cost = -1;
  for (int bn = 128; bn <= 256; bn += 32) {
    const long tiles = (long)((m_tiles + 1) / 2) * ((N + bn - 1) / bn);
    const long waves = (tiles + pairs - 1) / pairs;
    const long mma = 4L * std::max(88, bn / 2), ingest = (16384 + 64L * bn) / 64;
    const long cost = waves * std::max(mma, ingest) * 16 + bn / 32;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace

int sm100_gemm_tn_pair(const nbasr_gemm* g, cudaStream_t st) {
  NBASR_REQUIRE(g->K % 8 == 0, "K must keep 16-byte row alignment");
  Args a{};
  a.nb = g->nb; a.nr = g->nr; a.K = g->K; a.N = g->N;
  a.mt_per_utt = (g->nr + BM - 1) / BM;
  a.m_tiles = a.mt_per_utt * g->nb;
  const int sms = nbasr_sm_count();
  a.BN = pick_bn_pair(g->N, a.m_tiles, sms / 2);
  static const char* env_bn = getenv("NBASR_GEMM_BN");      // tuning override (tools/bench_gemm.py)
  if (env_bn) a.BN = std::max(64, std::min(256, atoi(env_bn) / 32 * 32));
  a.n_tiles = (g->N + a.BN - 1) / a.BN;
  a.pair_tiles = ((a.m_tiles + 1) / 2) * a.n_tiles;
  a.o_r0 = g->o_r0; a.o_bs = g->o_bs; a.o_rs = g->o_rs;
  a.epi = g->epi;
  CUtensorMap tmA, tmB;
  uint64_t da[3] = {(uint64_t)g->K, (uint64_t)g->nr, (uint64_t)g->nb};
  int64_t sa[3] = {1, g->a_rs, g->a_bs};
  uint32_t ba[3] = {BK, BM, 1};
  if (sm100_get_map(g->a, 3, da, sa, ba, &tmA)) return 1;
  uint64_t db[2] = {(uint64_t)g->K, (uint64_t)g->N};
  int64_t sb[2] = {1, g->ldw};
  uint32_t bb[2] = {BK, (uint32_t)(a.BN / 2)};
  if (sm100_get_map(g->w, 2, db, sb, bb, &tmB)) return 1;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return nbasr_fail("gemm_tn_pair smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  // persistent pairs: never launch more clusters than can be co-resident (a second wave would double the time)
  static int max_pairs = 0;
  if (!max_pairs) {
    cfg.gridDim = dim3(sms);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tn_pair_kernel, &cfg) != cudaSuccess || n < 1) n = sms / 2;
    max_pairs = std::min(n, sms / 2);
    if (getenv("NBASR_DEBUG")) fprintf(stderr, "[nbasr] gemm_tn_pair: %d co-resident CTA pairs on %d SMs\n", max_pairs, sms);
  }
  const int npairs = std::max(1, std::min(a.pair_tiles, max_pairs));
  cfg.gridDim = dim3(2 * npairs);
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tn_pair_kernel, tmA, tmB, a);
  if (e != cudaSuccess) return nbasr_fail("gemm_tn_pair launch: %s", cudaGetErrorString(e));
  return 0;
}
