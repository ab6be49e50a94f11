// Grouped conv edges (ops.py:73-76, groups=100, cpg = C/100 in {6,8,10,12}): forward and input-gradient on the
// warp-level tensor path (mma.sync m16n8k16 / m16n8k8, bf16 x bf16 -> fp32) fed by ldmatrix from a TMA-loaded tile.
//
// EXPERIMENT (NBASR_GCONV_FRAG=1 / nbasr_dbg_gconv_impl(1)); the product path is the tcgen05 kernel in gconv_sm100.cu -- measured
// result at the end of this comment.  Why try it: a grouped conv is HBM-bound (15-42 FLOP/B), but a
// tcgen05.mma with both operands in shared memory costs >= 88 cycles per instruction whatever its N (tools/dbg_bench.py),
// and the block-diagonal formulation needs 3*k of them per 128 x 48 tile = 1 320 cycles against ~1 100 cycles of HBM
// time: that kernel runs at its MMA issue floor.  The warp-level MMA has no such floor (measured 2.15 cycles per
// m16n8k16 per SM, 2.24 with the ldmatrix.x4 that feeds it: tools/ubench/hmma_bench.cu) and its K index can be laid
// out freely, so the TAPS go into K and the block-diagonal zero padding mostly disappears:
//
//   * an N-tile is 8 consecutive output channels; it needs the 1..3 aligned 8-channel input chunks that cover its
//     group(s).  One K=16 step = (chunk c, taps 2i and 2i+1): the A fragment is ONE ldmatrix.x4 whose 8x8 matrices are
//     the chunk's 16-byte row segments at frames m..m+15 shifted by tap 2i / 2i+1 (any dilation = a row offset); the odd
//     last tap is a K=8 step (ldmatrix.x2).  cpg = 8, k = 5: 2.5 MMAs per 16 frames x 8 channels, no zero padding at all.
//   * the B fragments come from the same block-diagonal bf16 pack the tcgen05 kernel uses (nbasr_pack_gconv_mma:
//     forward / group-transposed + tap-flipped for the input gradient); they are re-ordered once per CTA into shared
//     memory in fragment order (one conflict-free 8-byte load per K step and warp).
//   * persistent CTAs (2 per SM): one producer warp runs a TMA ring of (128 + halo) x 64-channel tiles in the
//     128B-swizzled layout (conflict-free ldmatrix); 12 compute warps (3 per scheduler) = 6 N-tiles x 2 frame halves of
//     a 48-channel slab, 4 independent accumulator tiles each.
//   * the fused epilogue (bias = accumulator init, ReLU20 + gate bits, dropout, skip-sum, second masked output) runs on
//     the accumulator fragments in registers; results are staged as a dense bf16 tile (double-buffered) and leave by TMA
//     store, gate bits as one coalesced 8-byte entry per row.  Same mask-plane format / slab width as the tcgen05
//     kernel: drop-in.
//
// Measured on B200 (profiles/r1_gconv_frag_experiment.txt): correct on all 16 shapes, but ~6 600 issued warp instructions per
// 128 x 48 tile against 4 400 issue slots at HBM speed -- the fragment-layout epilogue (~9 instructions per output element) and
// the ldmatrix/MMA stream share the same warps -- so it is 1.35-1.8x SLOWER than the tcgen05 kernel and stays an experiment.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int GT = 128;                    // frames per tile
constexpr int AROWS = 144;                 // 128 + max halo 12, multiple of 8
constexpr int A_BYTES = AROWS * 128;       // 18432
constexpr int NW = 48;                     // rows of one tap of the weight pack
constexpr int CW = 12;                     // compute warps: (N-tile, frame half)
constexpr int NCOMP = CW * 32;
constexpr int THREADS = 32 + NCOMP;        // + producer warp
constexpr int OST_BYTES = GT * NW * 2;     // 12288: one staged bf16 output tile
constexpr int MST_BYTES = GT * 8;          // gate-bit entries of one tile
constexpr int SMEM_LIMIT = 113 * 1024;     // 2 CTAs / SM

struct FragArgs {
  int B, T, C, OUT, cpg, dstep, off0;
  int nslabs, ntiles, tiles_per_utt, nlanes, dbg, ns, has_out2;
  const bf16* w;
  long long* trace;    // optional timeline dump (tools/trace_gconv.py): [cta < 8][tile < 64][8] clock64 stamps
  nbasr_epilogue epi;
  int64_t Tp;
};

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t* a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t addr, uint32_t* a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(a[0]), "=r"(a[1]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float* c, const uint32_t* a, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
}
// {lo, hi} -> bf16x2, round to nearest even; the .relu form clamps negatives (and NaN) to +0
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t min_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

__device__ __forceinline__ float2 ld2_dt(const void* base, int dtype, int64_t idx) {
  if (dtype == NBASR_BF16) return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const bf16*>(base) + idx));
  return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + idx);
}

// Fused output stage on the accumulator fragments of one warp: 4 m-tiles x rows (gid, gid + 8) x channels (col, col + 1);
// the accumulators already contain the bias.
struct FragEpi {
  uint8_t* o;          // staging address of (row r0, col); out2 tile at + OST_BYTES
  uint8_t* mrow;       // gate-bit byte of (row r0, this N-tile)
  const uint8_t* m2;   // mask2 byte of (tile row 0, this N-tile), rows eb2 bytes apart; nullptr = all ones
  int64_t rho0;        // padded row index of tile row 0
  int ostep, r0, nvr, tig, col, eb2;
};

// Lean variants (every launch of a skip-free, dropout-free model):
//   FWD : ReLU20 + gate bits -> out, mask_out;   !FWD: plain store (input gradient)
//   OUT2: second output = v * bit(mask2) * scale2 (dZ of the previous node)
//   PART: the last tile of an utterance: rows >= nvr are stored as zeros (they land on zero pad rows / are clipped)
template <bool FWD, bool OUT2, bool PART>
__device__ __forceinline__ void frag_epilogue_lean(const nbasr_epilogue& epi, const float (&acc)[4][4], const FragEpi& e) {
  uint32_t w2[8];
  if (OUT2) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = e.r0 + 8 * q;
      w2[q] = (e.m2 && (!PART || r < e.nvr)) ? (uint32_t)__ldg(e.m2 + r * e.eb2) >> (2 * e.tig) : 3u;
    }
  }
  const uint32_t cap = 0x41A041A0u;          // bf16x2 {20, 20}
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    uint32_t mword = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = 2 * mt + h;
      const float v0 = acc[mt][2 * h], v1 = acc[mt][2 * h + 1];
      const bool ok = !PART || (e.r0 + 8 * q < e.nvr);
      uint32_t y;
      if (FWD) {
        // 0 < z <= 20  <=>  bits(z) - 1 < bits(20.0f) as unsigned (negative z and +0 wrap to huge values)
        if ((__float_as_uint(v0) - 1u) < 0x41A00000u && ok) mword |= 1u << (8 * h);
        if ((__float_as_uint(v1) - 1u) < 0x41A00000u && ok) mword |= 2u << (8 * h);
        // clamp after rounding = rounding after clamp (rounding is monotonic; 0 and 20 are bf16 numbers)
        y = min_bf16x2(pack_relu_bf16x2(v0, v1), cap);
      } else {
        y = pack_bf16x2(v0, v1);
      }
      if (PART && !ok) y = 0u;
      *reinterpret_cast<uint32_t*>(e.o + q * e.ostep) = y;
      if (OUT2) {
        // the reference order: value (fp32) * gate * scale, rounded once
        float f0 = v0, f1 = v1;
        if (FWD) { f0 = fminf(fmaxf(v0, 0.f), 20.f); f1 = fminf(fmaxf(v1, 0.f), 20.f); }
        uint32_t y2 = pack_bf16x2((w2[q] & 1u) ? f0 * epi.scale2 : 0.f, (w2[q] & 2u) ? f1 * epi.scale2 : 0.f);
        if (PART && !ok) y2 = 0u;
        *reinterpret_cast<uint32_t*>(e.o + q * e.ostep + OST_BYTES) = y2;
      }
    }
    if (FWD) {
      mword <<= 2 * e.tig;
      mword |= __shfl_xor_sync(0xffffffffu, mword, 1);
      mword |= __shfl_xor_sync(0xffffffffu, mword, 2);
      // lane tig = 0 stores the byte of row gid, tig = 1 that of row gid + 8 (one predicated store, no divergence)
      if (e.tig < 2) e.mrow[mt * 128 + e.tig * 64] = (uint8_t)(mword >> (8 * e.tig));
    }
  }
}

// General variant: dropout, skip-sum operands, optional stores.  Not unrolled over the skip operands / rows more than
// needed: it is the path of the skip-connected architectures and of training with dropout.
__device__ __noinline__ void frag_epilogue_gen(const nbasr_epilogue& epi, const float (&acc)[4][4], const FragEpi& e) {
  float dscale = 1.f;
  uint32_t thr = 0;
  uint64_t seed = 0;
  if (epi.drop_p > 0.f) {
    dscale = 1.f / (1.f - epi.drop_p);
    thr = static_cast<uint32_t>(epi.drop_p * 4294967296.0);
    seed = epi.drop_seed + (epi.drop_step ? __ldg(epi.drop_step) * 0xD1B54A32D192ED03ull : 0ull);
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    uint32_t mword = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = 2 * mt + h;
      const int r = e.r0 + 8 * q;
      float v0 = acc[mt][2 * h], v1 = acc[mt][2 * h + 1];
      uint32_t bits = 3u;
      if (epi.relu20) {
        bits = (((__float_as_uint(v0) - 1u) < 0x41A00000u) ? 1u : 0u) | (((__float_as_uint(v1) - 1u) < 0x41A00000u) ? 2u : 0u);
        v0 = fminf(fmaxf(v0, 0.f), 20.f);
        v1 = fminf(fmaxf(v1, 0.f), 20.f);
      }
      uint32_t w = 3u;
      if (r < e.nvr) {
        const int64_t rho = e.rho0 + r;
        if (epi.drop_p > 0.f) {
          const bool k0 = hash_u32(seed, static_cast<uint64_t>(rho) * 4096ull + e.col) >= thr;
          const bool k1 = hash_u32(seed, static_cast<uint64_t>(rho) * 4096ull + e.col + 1) >= thr;
          if (!k0) bits &= ~1u;
          if (!k1) bits &= ~2u;
          v0 = k0 ? v0 * dscale : 0.f;
          v1 = k1 ? v1 * dscale : 0.f;
        }
#pragma unroll 1
        for (int a = 0; a < epi.n_add; ++a) {
          const float2 s = ld2_dt(epi.add[a], epi.add_dtype, rho * epi.ld_out + e.col);
          v0 += s.x;
          v1 += s.y;
        }
        if (e.m2) w = (uint32_t)__ldg(e.m2 + r * e.eb2) >> (2 * e.tig);
      } else {
        v0 = 0.f; v1 = 0.f; bits = 0u;     // rows past the utterance land on zero pad rows / are clipped
      }
      mword |= bits << (8 * h);
      if (epi.out) *reinterpret_cast<uint32_t*>(e.o + q * e.ostep) = pack_bf16x2(v0, v1);
      if (epi.out2)
        *reinterpret_cast<uint32_t*>(e.o + q * e.ostep + OST_BYTES) =
            pack_bf16x2((w & 1u) ? v0 * epi.scale2 : 0.f, (w & 2u) ? v1 * epi.scale2 : 0.f);
    }
    if (epi.mask_out) {
      mword <<= 2 * e.tig;
      mword |= __shfl_xor_sync(0xffffffffu, mword, 1);
      mword |= __shfl_xor_sync(0xffffffffu, mword, 2);
      if (e.tig < 2) e.mrow[mt * 128 + e.tig * 64] = (uint8_t)(mword >> (8 * e.tig));
    }
  }
}

#define FR_STAMP(slot)                                                                          \
  do {                                                                                          \
    if (p.trace && blockIdx.x < 8 && it < 64) p.trace[((size_t)blockIdx.x * 64 + it) * 8 + (slot)] = clock64(); \
  } while (0)

template <int KT>
__global__ void __launch_bounds__(THREADS, 2)
gconv_frag_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                  const __grid_constant__ CUtensorMap tmO2, const FragArgs p) {
  constexpr int NP = KT / 2;               // K=16 steps per chunk (tap pairs); the odd last tap is a K=8 step
  constexpr int NKS = NP + 1;
  constexpr int BSM_BYTES = 6 * 3 * NKS * 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
  const int NS = p.ns;
  const int obuf = p.has_out2 ? 2 * OST_BYTES : OST_BYTES;   // bytes of one staging buffer (out [+ out2])
  const uint32_t osm = base + NS * A_BYTES;               // [buf][out | out2] staging
  uint8_t* ost = al + NS * A_BYTES;
  uint8_t* mst = ost + 2 * obuf;                          // [buf][128] 8-byte gate-bit entries
  uint8_t* bsm = mst + 2 * MST_BYTES;                     // B fragments [nt][c][ks][lane] x 8 bytes
  const uint32_t bar0 = osm + 2 * obuf + 2 * MST_BYTES + BSM_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (4 + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.x % p.nslabs;
  const int lane_id = blockIdx.x / p.nslabs;
  const int c0 = slab * p.OUT;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmO);
    prefetch_tmap(&tmO2);
    for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CW); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * MST_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(mst)[i] = 0u;
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();                                  // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = lane_id; tile < p.ntiles; tile += p.nlanes, ++it) {
        const int b = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * GT;
        mbar_wait(empty_bar(stage), phase ^ 1);
        FR_STAMP(0);
        mbar_expect_tx(full_bar(stage), A_BYTES);
        tma_load_3d(base + stage * A_BYTES, &tmX, full_bar(stage), c0, NBASR_PAD_L + t0 + p.off0, b);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
    return;
  }

  // ---------------------------------------------------------------- compute warps
  const int cw = warp - 1;                   // 0..11
  const int nt = cw % 6, mh = cw / 6;        // N-tile of the slab, frame half of the tile
  const int ctid = threadIdx.x - 32;
  const int gid = lane >> 2, tig = lane & 3;
  const int colg = c0 + 8 * nt;              // first channel of the N-tile
  const bool active = 8 * nt < p.OUT && colg < p.C;
  const int col = colg + 2 * tig;            // this thread's two output channels: col, col + 1
  const nbasr_epilogue& epi = p.epi;         // stays in the kernel-parameter bank

  // input chunks covering the group(s) of this N-tile (slab-local 8-channel chunks clo .. clo + nch - 1, nch <= 3)
  const int g_lo = (8 * nt) / p.cpg, g_hi = (8 * nt + 7) / p.cpg;
  const int clo = (g_lo * p.cpg) >> 3;
  const int nch = active ? min(3, ((min((g_hi + 1) * p.cpg, p.OUT) + 7) >> 3) - clo) : 0;

  // B fragments -> shared memory.  K index = (tap parity, channel in chunk), n = output channel; pack[slab][tap][n (48)][kk (64)]
  uint2* bfr = reinterpret_cast<uint2*>(bsm) + (nt * 3 * NKS) * 32 + lane;
  if (mh == 0) {
    for (int c = 0; c < nch; ++c) {
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(p.w + ((int64_t)slab * KT * NW + 8 * nt + gid) * 64 + 8 * (clo + c) + 2 * tig);
#pragma unroll
      for (int i = 0; i < NP; ++i)
        bfr[(c * NKS + i) * 32] = make_uint2(__ldg(wp + (2 * i) * (NW * 32)), __ldg(wp + (2 * i + 1) * (NW * 32)));
      bfr[(c * NKS + NP) * 32] = make_uint2(__ldg(wp + (KT - 1) * (NW * 32)), 0u);
    }
  }
  float bias0 = 0.f, bias1 = 0.f;
  if (epi.bias && active) { bias0 = __ldg(epi.bias + col); bias1 = __ldg(epi.bias + col + 1); }
  named_bar_sync(1, NCOMP);

  // per-lane ldmatrix row inside a 16-frame m-tile: matrices 0/1 = frames 0-7 / 8-15 at tap 2i, 2/3 at tap 2i+1
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8 + mh * 64;
  const int ltap = lane >> 4;
  const int OUTB = p.OUT * 2;
  FragEpi e;
  e.ostep = 8 * OUTB;
  e.r0 = mh * 64 + gid;
  e.tig = tig;
  e.col = col;
  e.eb2 = 0;
  int plane2 = 0, byte2 = 0;
  if (epi.out2 && epi.mask2) {
    plane2 = colg / epi.mask2_w;
    byte2 = (colg - plane2 * epi.mask2_w) >> 3;
    e.eb2 = epi.mask2_w == 32 ? 4 : 8;
  }
  // general path: dropout / skip operands and the odd store combinations
  const bool gen = epi.drop_p != 0.f || epi.n_add != 0 || !epi.out || (epi.relu20 != 0) != (epi.mask_out != nullptr);

  int stage = 0, it = 0;
  uint32_t phase = 0;
  for (int tile = lane_id; tile < p.ntiles; tile += p.nlanes, ++it) {
    const int b = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * GT;
    float acc[4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) { acc[mt][0] = bias0; acc[mt][1] = bias1; acc[mt][2] = bias0; acc[mt][3] = bias1; }

    if (ctid == 0) FR_STAMP(1);
    mbar_wait(full_bar(stage), phase);
    if (ctid == 0) FR_STAMP(2);
    const uint32_t sa = base + stage * A_BYTES;
    if (!(p.dbg & 1))
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      const int chunk = clo + c;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const uint2 bq = bfr[(c * NKS + i) * 32];
        const int row = lrow + (2 * i + ltap) * p.dstep;
        const uint32_t ad = sa + row * 128 + ((chunk ^ (row & 7)) << 4);
        uint32_t a[4][4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) ldsm4(ad + mt * 2048, a[mt]);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) mma16816(acc[mt], a[mt], bq.x, bq.y);
      }
      {
        const uint2 bq = bfr[(c * NKS + NP) * 32];
        const int row = lrow + (KT - 1) * p.dstep;       // lanes 16-31: addresses unused by .x2 but valid
        const uint32_t ad = sa + row * 128 + ((chunk ^ (row & 7)) << 4);
        uint32_t a[4][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) ldsm2(ad + mt * 2048, a[mt]);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) mma1688(acc[mt], a[mt], bq.x);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(stage));          // this warp is done reading the input stage
    if (ctid == 0) FR_STAMP(3);
    if (++stage == NS) { stage = 0; phase ^= 1; }

    // ---- epilogue on the fragments: thread holds rows (gid, gid + 8) x channels (col, col + 1) of each m-tile
    const int buf = it & 1;
    if (active && !(p.dbg & 2)) {
      e.o = ost + buf * obuf + (mh * 64 + gid) * OUTB + (8 * nt + 2 * tig) * 2;
      e.mrow = mst + buf * MST_BYTES + (mh * 64 + gid) * 8 + nt;
      e.nvr = p.T - t0;
      e.rho0 = (int64_t)b * p.Tp + NBASR_PAD_L + t0;
      e.m2 = e.eb2 ? reinterpret_cast<const uint8_t*>(epi.mask2) + ((int64_t)plane2 * epi.mask_rows + e.rho0) * e.eb2 + byte2 : nullptr;
      const bool part = e.nvr < GT;
      if (gen) {
        frag_epilogue_gen(epi, acc, e);
      } else if (epi.relu20) {
        if (epi.out2) { if (part) frag_epilogue_lean<true, true, true>(epi, acc, e); else frag_epilogue_lean<true, true, false>(epi, acc, e); }
        else { if (part) frag_epilogue_lean<true, false, true>(epi, acc, e); else frag_epilogue_lean<true, false, false>(epi, acc, e); }
      } else {
        if (epi.out2) { if (part) frag_epilogue_lean<false, true, true>(epi, acc, e); else frag_epilogue_lean<false, true, false>(epi, acc, e); }
        else { if (part) frag_epilogue_lean<false, false, true>(epi, acc, e); else frag_epilogue_lean<false, false, false>(epi, acc, e); }
      }
    }
    uint8_t* mb = mst + buf * MST_BYTES;
    if (ctid == 0) FR_STAMP(4);
    // this barrier also tells everybody that the TMA stores of the previous tile have finished READING their staging
    // buffer, i.e. that the OTHER buffer may be overwritten by the next tile
    if (ctid == 0) bulk_wait_read0();
    fence_async_smem();
    named_bar_sync(1, NCOMP);
    if (ctid == 0) FR_STAMP(5);
    if (epi.mask_out && ctid < GT && t0 + ctid < p.T) {
      // 128 consecutive 8-byte entries of this slab's mask plane: one fully coalesced store per warp
      const int64_t r2 = (int64_t)b * p.Tp + NBASR_PAD_L + t0 + ctid;
      reinterpret_cast<uint64_t*>(epi.mask_out)[(int64_t)slab * epi.mask_rows + r2] = reinterpret_cast<const uint64_t*>(mb)[ctid];
    }
    if (ctid == 0) {
      const uint32_t so = osm + buf * obuf;
      if (epi.out) tma_store_3d(&tmO, so, c0, NBASR_PAD_L + t0, b);
      if (epi.out2) tma_store_3d(&tmO2, so + OST_BYTES, c0, NBASR_PAD_L + t0, b);
      bulk_commit();
      FR_STAMP(6);
    }
  }
  if (ctid == 0) bulk_wait0();
}

template <int KT>
int frag_launch(FragArgs& a, const CUtensorMap& tmX, const CUtensorMap& tmO, const CUtensorMap& tmO2, cudaStream_t st) {
  constexpr int BSM_BYTES = 6 * 3 * (KT / 2 + 1) * 256;
  const int obuf = a.has_out2 ? 2 * OST_BYTES : OST_BYTES;
  const int fixed = 2 * obuf + 2 * MST_BYTES + BSM_BYTES + 64 + 1024;
  a.ns = std::min(4, (SMEM_LIMIT - fixed) / A_BYTES);
  const size_t smem = (size_t)fixed + (size_t)a.ns * A_BYTES;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gconv_frag_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return nbasr_fail("gconv_frag smem attr: %s", cudaGetErrorString(e));
    attr = true;
    if (a.dbg) {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gconv_frag_kernel<KT>, THREADS, smem);
      fprintf(stderr, "gconv_frag<%d>: %d CTAs/SM, smem %zu, %d stages\n", KT, nb, smem, a.ns);
    }
  }
  cudaError_t le = launch_pdl(gconv_frag_kernel<KT>, dim3(a.nslabs * a.nlanes), dim3(THREADS), smem, st, 1, tmX, tmO, tmO2, a);
  if (le != cudaSuccess) return nbasr_fail("gconv_frag launch: %s", cudaGetErrorString(le));
  return 0;
}

}  // namespace

extern unsigned long long* g_gconv_dbg;   // gconv_sm100.cu (nbasr_dbg_gconv_trace)

int sm100_gconv_fwd_frag(const nbasr_gconv* g, cudaStream_t st) {
  FragArgs a{};
  a.B = g->B; a.T = g->T; a.Tp = g->Tp; a.C = g->C; a.cpg = g->cpg; a.OUT = g->cpg == 10 ? 40 : 48;
  a.dstep = g->dstep; a.off0 = g->off0;
  a.w = reinterpret_cast<const bf16*>(g->w);
  NBASR_REQUIRE(g->ktaps == 5 || g->ktaps == 7, "fragment grouped conv: 5 or 7 taps");
  NBASR_REQUIRE(a.off0 >= -NBASR_PAD_L && (g->ktaps - 1) * a.dstep <= AROWS - GT, "tap reach");
  NBASR_REQUIRE(a.OUT % a.cpg == 0 && a.cpg <= 16, "slab = whole groups, <= 3 chunks per N-tile");
  NBASR_REQUIRE(g->C % 8 == 0, "channel count multiple of 8");
  NBASR_REQUIRE(g->epi.ld_out == g->C, "grouped conv writes dense (B,Tp,C) tensors");
  NBASR_REQUIRE((!g->epi.out || g->epi.out_dtype == NBASR_BF16) && (!g->epi.out2 || g->epi.out2_dtype == NBASR_BF16) &&
                    !g->epi.accumulate, "grouped conv stores bf16");
  NBASR_REQUIRE(!g->epi.mask_out || g->epi.mask_w == a.OUT, "grouped-conv mask planes are slab wide");
  a.nslabs = (g->C + a.OUT - 1) / a.OUT;
  a.tiles_per_utt = (g->T + GT - 1) / GT;
  a.ntiles = a.tiles_per_utt * g->B;
  const int slots = 2 * nbasr_sm_count();
  a.nlanes = std::max(1, std::min(a.ntiles, slots / a.nslabs));
  a.epi = g->epi;
  a.has_out2 = g->epi.out2 ? 1 : 0;
  a.trace = reinterpret_cast<long long*>(g_gconv_dbg);
  static const char* dbg = getenv("NBASR_FRAG_DBG");
  a.dbg = dbg ? atoi(dbg) : 0;
  CUtensorMap tmX, tmO, tmO2;
  uint64_t dx[3] = {(uint64_t)g->C, (uint64_t)g->Tp, (uint64_t)g->B};
  int64_t sx[3] = {1, g->C, (int64_t)g->Tp * g->C};
  uint32_t bx[3] = {64, AROWS, 1};
  if (sm100_get_map(g->x, 3, dx, sx, bx, &tmX)) return 1;
  uint32_t bo[3] = {(uint32_t)a.OUT, GT, 1};
  const void* o1 = g->epi.out ? g->epi.out : g->x;       // unused maps still need a valid descriptor
  const void* o2 = g->epi.out2 ? g->epi.out2 : g->x;
  if (sm100_get_map(o1, 3, dx, sx, bo, &tmO, 0)) return 1;
  if (sm100_get_map(o2, 3, dx, sx, bo, &tmO2, 0)) return 1;
  return g->ktaps == 5 ? frag_launch<5>(a, tmX, tmO, tmO2, st) : frag_launch<7>(a, tmX, tmO, tmO2, st);
}
