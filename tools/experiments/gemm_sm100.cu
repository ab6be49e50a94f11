// tcgen05 / TMEM / TMA GEMM kernels for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
//  gemm_tn   : C[(b,r), n] = sum_k A[(b,r), k] W[n, k]  -- both operands K-major, 128B swizzle.
//              Persistent CTAs (one per SM), static tile scheduler, 4-stage TMA->smem ring,
//              double-buffered TMEM accumulator so the fused epilogue (bias, ReLU20, dropout,
//              skip-sum, gradient mask) of tile i overlaps the MMAs of tile i+1.
//              Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..5 = epilogue.
//              The A operand is described by a 3-D tensor map (K, rows-per-utterance, batch)
//              whose row stride may be smaller than K: every tap of the k=8 time-reduction
//              convolutions is then just a column range of the same map (implicit GEMM with no
//              im2col copy); zero padding = zero pad rows + TMA out-of-bounds fill.
//  gemm_wgrad: dW[m, n] += sum_{b,r} dY[(b,r), m] X[(b,r), n] -- both operands MN-major
//              (the reduction index is the slow index in memory), split-K over (utterance,
//              64-frame chunk) units, fp32 red.global.add epilogue.
#include <cuda.h>

#include <cstdlib>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;             // 64 bf16 = 128 bytes = one swizzle span
constexpr int BN_MAX = 256;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KB
constexpr int B_STAGE_BYTES = BN_MAX * BK * 2;    // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;        // weight-gradient kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int TN_THREADS = 320;         // gemm_tn: TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant)

using namespace sm100;

struct TnArgs {
  int nb, nr, K, N, BN;
  int mt_per_utt, n_tiles, total_tiles;
  int64_t o_r0, o_bs, o_rs;
  nbasr_epilogue epi;
};

__global__ void __launch_bounds__(TN_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  // barrier slots (8 B each): full[0..S), empty[S..2S), tmem_full[2S..2S+2), tmem_empty[2S+2..2S+4)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_al + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_idx = tile % p.n_tiles, m_idx = tile / p.n_tiles;
        const int b = m_idx / p.mt_per_utt, r0 = (m_idx % p.mt_per_utt) * BM;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), A_STAGE_BYTES + p.BN * BK * 2);
          tma_load_3d(sa, &tmA, full_bar(stage), kb * BK, r0, b);
          tma_load_2d(sb, &tmB, full_bar(stage), kb * BK, n_idx * p.BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, p.BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN_MAX;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t ad = make_smem_desc(sa + k * 32, 16, 1024);
            uint64_t bd = make_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tfull_bar(as));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int hh = (warp - 2) >> 2;   // the two warps of a quadrant take alternate 32-column chunks
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int n_idx = tile % p.n_tiles, m_idx = tile / p.n_tiles;
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
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN_MAX + c, v);
        if (r < p.nr) epilogue_chunk(p.epi, rho, n0 + c, ncol, v);
      }
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(as));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- weight gradient (MN-major operands)
struct WgArgs {
  int nb, nr, M, N, BN;
  int m_tiles, n_tiles, chunks_per_utt, total_units, units_per_split;
  float* dw;
  int64_t ldw;
  float* dbias;   // optional bias gradient: column sums of dY via one extra N=16 MMA against a tile of ones
};

constexpr int WG_ONES_BYTES = 8192;   // 64 K-rows x 128 B, every element 1.0 (swizzle invariant)
constexpr int WG_SMEM_BYTES = STAGES * STAGE_BYTES + WG_ONES_BYTES + 1024 + 256;

__device__ __forceinline__ void tmem_ld16_(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ones_sm = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = ones_sm + WG_ONES_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_al + STAGES * STAGE_BYTES + WG_ONES_BYTES + 8 * (2 * STAGES + 4));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int m0 = (tile % p.m_tiles) * BM, n0 = (tile / p.m_tiles) * p.BN;
  const int u_begin = blockIdx.y * p.units_per_split;
  const int u_end = min(p.total_units, u_begin + p.units_per_split);
  const int nbox_b = (p.BN + 63) / 64;
  const bool do_bias = p.dbias != nullptr && n0 == 0;
  if (do_bias) {
    uint32_t* o = reinterpret_cast<uint32_t*>(smem_al + STAGES * STAGE_BYTES);
    for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += blockDim.x) o[i] = 0x3F803F80u;   // bf16 (1.0, 1.0)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (u_begin >= u_end) {  // nothing to do for this split (uniform per CTA)
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
    return;
  }

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = u_begin; u < u_end; ++u) {
        const int b = u / p.chunks_per_utt, r0 = (u % p.chunks_per_utt) * BK;
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
        mbar_expect_tx(full_bar(stage), (2 + nbox_b) * 64 * 64 * 2);
        tma_load_3d(sa, &tmDY, full_bar(stage), m0, r0, b);
        tma_load_3d(sa + 8192, &tmDY, full_bar(stage), m0 + 64, r0, b);
        for (int h = 0; h < nbox_b; ++h) tma_load_3d(sb + h * 8192, &tmX, full_bar(stage), n0 + h * 64, r0, b);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, p.BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int u = u_begin; u < u_end; ++u) {
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // MN-major SW128: 64-element (128 B) MN spans, LBO = distance between spans (one 64x64
          // box = 8 KB), SBO = distance between 8-row K groups (1 KB); 16 K rows = 2 KB per MMA.
          uint64_t ad = make_smem_desc(sa + k * 2048, 8192, 1024);
          uint64_t bd = make_smem_desc(sb + k * 2048, 8192, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (u > u_begin || k > 0) ? 1u : 0u);
          if (do_bias)
            umma_bf16(tmem_base + 256, ad, make_smem_desc(ones_sm + k * 2048, 8192, 1024), make_idesc(BM, 16, 1, 1),
                      (u > u_begin || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (u == u_end - 1) umma_commit(tfull_bar);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tcgen05_fence_after();
    const int ncol = min(p.N, n0 + p.BN);
    if (do_bias) {
      float one[16];
      tmem_ld16_(tmem_base + ((uint32_t)(q * 32) << 16) + 256, one);
      if (m < p.M) atomicAdd(p.dbias + m, one[0]);
    }
    for (int c = 0; c < p.BN; c += 32) {
      if (n0 + c >= ncol) break;
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
      if (m < p.M) {
        float* dst = p.dw + (int64_t)m * p.ldw + n0 + c;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (n0 + c + g * 4 < ncol)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4), "f"(v[g * 4]), "f"(v[g * 4 + 1]),
                         "f"(v[g * 4 + 2]), "f"(v[g * 4 + 3])
                         : "memory");
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

struct MapKey {
  uint64_t v[10];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 10; ++i) { h ^= k.v[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

}  // namespace

// bf16 tensor map of rank 2 or 3: dims[0] is the contiguous one; strides in ELEMENTS for dims 1,2.
int sm100_get_map(const void* base, int rank, const uint64_t* dims, const int64_t* strides_el, const uint32_t* box, CUtensorMap* out,
                  int swizzle128) {
  MapKey k{};
  k.v[0] = (uint64_t)base; k.v[1] = rank | (swizzle128 ? 0 : 16);
  for (int i = 0; i < rank; ++i) { k.v[2 + i] = dims[i]; k.v[7 + i] = box[i]; }
  for (int i = 1; i < rank; ++i) k.v[4 + i] = (uint64_t)strides_el[i];
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(k);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  EncodeTiledFn enc = get_encode();
  if (!enc) return nbasr_fail("cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 1; i < rank; ++i) gstr[i - 1] = (cuuint64_t)strides_el[i] * 2;
  if ((uint64_t)base % 16 != 0) return nbasr_fail("TMA base %p not 16-byte aligned", base);
  for (int i = 1; i < rank; ++i)
    if (gstr[i - 1] % 16 != 0) return nbasr_fail("TMA stride %llu not a multiple of 16 bytes", (unsigned long long)gstr[i - 1]);
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return nbasr_fail("cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu %llu)", (int)r, rank,
                                           (unsigned long long)dims[0], (unsigned long long)dims[1],
                                           (unsigned long long)(rank > 2 ? dims[2] : 0));
  g_maps[k] = m;
  *out = m;
  return 0;
}

namespace {

// Tile width: minimise  waves x cycles-per-K-block.  A K block (64) costs max(MMA issue, operand ingest): four MMAs of
// max(88, BN/2) cycles (tools/dbg_bench.py) against (16 KB of A + 128 BN bytes of B) at ~64 B/clk/SM from L2 (measured:
// the 128 x 256 tile runs ~750 cycles per K block, not 512).  Narrow tiles (the old "least padding" rule picked BN = 64
// for N = 1200) multiply the A re-reads and the wave count.
int pick_bn(int N, int m_tiles, int sms) {
  int best = 256;
  long best_cost = -1;
  for (int bn = 96; bn <= 256; bn += 32) {
    const long tiles = (long)m_tiles * ((N + bn - 1) / bn);
    const long waves = (tiles + sms - 1) / sms;
    const long mma = 4L * std::max(88, bn / 2), ingest = (16384 + 128L * bn) / 64;
    const long cost = waves * std::max(mma, ingest) * 16 + bn / 32;   // ties -> narrower tile (less padding work)
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace

int sm100_gemm_tn(const nbasr_gemm* g, cudaStream_t st) {
  NBASR_REQUIRE(g->K % 8 == 0, "K must keep 16-byte row alignment");
  TnArgs a{};
  a.nb = g->nb; a.nr = g->nr; a.K = g->K; a.N = g->N;
  a.mt_per_utt = (g->nr + BM - 1) / BM;
  a.BN = pick_bn(g->N, a.mt_per_utt * g->nb, nbasr_sm_count());
  static const char* env_bn = getenv("NBASR_GEMM_BN");      // tuning override (tools/bench_gemm.py)
  if (env_bn) a.BN = std::max(32, std::min(256, atoi(env_bn) / 32 * 32));
  a.mt_per_utt = (g->nr + BM - 1) / BM;
  a.n_tiles = (g->N + a.BN - 1) / a.BN;
  a.total_tiles = a.mt_per_utt * g->nb * a.n_tiles;
  a.o_r0 = g->o_r0; a.o_bs = g->o_bs; a.o_rs = g->o_rs;
  a.epi = g->epi;
  CUtensorMap tmA, tmB;
  uint64_t da[3] = {(uint64_t)g->K, (uint64_t)g->nr, (uint64_t)g->nb};
  int64_t sa[3] = {1, g->a_rs, g->a_bs};
  uint32_t ba[3] = {BK, BM, 1};
  if (sm100_get_map(g->a, 3, da, sa, ba, &tmA)) return 1;
  uint64_t db[2] = {(uint64_t)g->K, (uint64_t)g->N};
  int64_t sb[2] = {1, g->ldw};
  uint32_t bb[2] = {BK, (uint32_t)a.BN};
  if (sm100_get_map(g->w, 2, db, sb, bb, &tmB)) return 1;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return nbasr_fail("gemm_tn smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  int grid = std::min(a.total_tiles, nbasr_sm_count());
  gemm_tn_kernel<<<grid, TN_THREADS, SMEM_BYTES, st>>>(tmA, tmB, a);
  NBASR_CHECK_LAUNCH();
  return 0;
}

int sm100_gemm_wgrad(const nbasr_wgrad* g, cudaStream_t st) {
  NBASR_REQUIRE(g->N % 4 == 0 && g->ldw % 4 == 0, "wgrad N / ldw must be multiples of 4");
  WgArgs a{};
  a.nb = g->nb; a.nr = g->nr; a.M = g->M; a.N = g->N;
  a.BN = g->N >= 256 ? 256 : ((g->N + 63) / 64) * 64;
  a.m_tiles = (g->M + BM - 1) / BM;
  a.n_tiles = (g->N + a.BN - 1) / a.BN;
  a.chunks_per_utt = (g->nr + BK - 1) / BK;
  a.total_units = a.chunks_per_utt * g->nb;
  int tiles = a.m_tiles * a.n_tiles;
  int sms = nbasr_sm_count();
  int splits = std::max(1, std::min(a.total_units, (2 * sms + tiles - 1) / tiles));
  a.units_per_split = (a.total_units + splits - 1) / splits;
  splits = (a.total_units + a.units_per_split - 1) / a.units_per_split;
  a.dw = g->dw; a.ldw = g->ldw; a.dbias = g->dbias;
  CUtensorMap tmDY, tmX;
  uint64_t dd[3] = {(uint64_t)g->M, (uint64_t)g->nr, (uint64_t)g->nb};
  int64_t sd[3] = {1, g->dy_rs, g->dy_bs};
  uint32_t bx[3] = {64, BK, 1};
  if (sm100_get_map(g->dy, 3, dd, sd, bx, &tmDY)) return 1;
  uint64_t dx[3] = {(uint64_t)g->N, (uint64_t)g->nr, (uint64_t)g->nb};
  int64_t sx[3] = {1, g->x_rs, g->x_bs};
  if (sm100_get_map(g->x, 3, dx, sx, bx, &tmX)) return 1;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e != cudaSuccess) return nbasr_fail("gemm_wgrad smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  gemm_wgrad_kernel<<<dim3(tiles, splits), NUM_THREADS, WG_SMEM_BYTES, st>>>(tmDY, tmX, a);
  NBASR_CHECK_LAUNCH();
  return 0;
}
