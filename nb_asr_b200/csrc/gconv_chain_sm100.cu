// A chain of grouped-conv edges in ONE launch (model.py:13-22,49-59: node n+1's op reads node n's output).
//
// Groups never mix channels, so a CTA that owns (slab of 48/40 channels, contiguous range of 128-frame tiles) can run
// node 0, 1, 2 back to back: the only cross-CTA dependency is the halo of (k-1)*dstep frames that belongs to the
// neighbouring CTA of the same slab.  One launch therefore replaces up to three of gconv_mma_fwd_kernel (same tile
// engine: TMA-loaded input tile, block-diagonal tcgen05 MMAs per tap, fused epilogue, TMA store):
//   * the per-launch set-up / first-tile latency / tail (10-15 us of a ~30 us launch, profiles/r2_ncu_launches_summary.txt)
//     is paid once, the TMA ring and the TMEM accumulator stages keep running across the node boundary, and the
//     intermediate tensors are re-read from L2 while still resident;
//   * the epilogue thread that issued a tile's TMA stores waits for their completion (lazily, one tile late, with
//     cp.async.bulk.wait_group 1; immediately for the last tile of a node) and then advances a progress word in SHARED
//     memory; the producer thread polls it for the tiles of its own range, and only when its requirement exceeds the
//     progress it already knows -- in steady state no poll and no fence at all.  Only the FIRST and LAST tile of a range
//     are visible to other CTAs: they have a flag word in a caller-provided work buffer, published as epoch + 1 with
//     st.release.gpu and acquired by the neighbour's producer.  After any acquire: fence.proxy.async, then the TMA load.
//     (ld.acquire.gpu / st.release.gpu compile to MEMBAR.GPU + CCTL.IVALL: per tile they cost 30 % of the kernel.)
//   * no per-launch memset: `epoch` lives in the work buffer and the last CTA to leave (atomic ticket) advances it, so
//     flags of earlier launches never match.  The state is device memory, so CUDA-graph replay and eager launches mix;
//   * skip-sum operands produced earlier in the same chain are read with ld.global.cg (L1 could hold sectors that the
//     other CTA of the SM fetched before they were written);
//   * a flag wait gives up after ~1 s and sets work[2] (the tests check it): a protocol error can not hang the GPU.
// Used for the forward node chain of a cell and for its input-gradient chain (dZ_n -> dZ_{n-1}); x of node i + 1 must be
// `out` or `out2` of node i, and every tensor written inside the chain must be a distinct buffer.
#include <cuda.h>

#include "common.cuh"
#include "gconv_sm100.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int MAXCHAIN = 3;
constexpr int CH_THREADS = 320;             // 8 epilogue warps (two column halves x four TMEM lane quadrants), producer, MMA
constexpr int CH_NEPI = 256;
constexpr int CH_WORK_HDR = 16;            // u32 words in front of the flags: [0] epoch, [1] exit ticket, [2] error

struct GcChainMaps {
  CUtensorMap x[MAXCHAIN], w[MAXCHAIN], o[MAXCHAIN], o2[MAXCHAIN];
};
struct GcChainNode {
  int ktaps, dstep, off0, pad_;
  nbasr_epilogue epi;
};
struct GcChainArgs {
  int B, T, C, OUT, n_nodes, nslabs, ntiles, tiles_per_utt, nlanes, nstage, maxtaps, no_prefetch, w_stable, f16, dbg;
  int64_t Tp;
  uint32_t* work;
  GcChainNode node[MAXCHAIN];
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t saddr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const uint32_t* f, uint32_t done, uint32_t* err) {
  uint32_t spins = 0;
  while (ld_acquire_gpu(f) != done) {
    __nanosleep(64);
    if (++spins > (1u << 24)) {
      atomicExch(err, 1u);
      break;
    }
  }
}
// v[0..8) += 16 bytes of a 16-bit tensor.  cg: read through L2 only.  Needed when a slab is not a whole number of 32-byte
// sectors (40-channel slabs): the neighbouring slab's CTA may then have pulled a sector into this SM's L1 before the bytes
// of THIS slab -- produced earlier in the same launch -- were written.  48-channel slabs own their sectors: cached loads.
__device__ __forceinline__ void add8_cg(const void* base, int dtype, int64_t idx, float* v, bool cg = true) {
  const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + idx);
  const uint4 r = cg ? __ldcg(q) : *q;
  const uint32_t* h = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = dtype == NBASR_F16 ? f16x2_to_f2(h[i]) : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h[i]));
    v[2 * i] += f.x;
    v[2 * i + 1] += f.y;
  }
}

// ---- packed fp32 pairs (FMUL2)
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// Specialised epilogue of one thread's 24 accumulator columns, for the configurations the train / eval step actually uses
// (full slab, no dropout, no skip operand): straight-line code, packed fp32 arithmetic, outputs left PACKED in registers.
// The runtime-configurable path below spends ~630 issued instructions per warp and tile, most of them branches and flag
// tests (ncu r2: the kernel ran at IPC 2.0 = the issue limit of the fma / alu pipes, i.e. it was bound by its own epilogue,
// at twice the MMA issue floor).  Arithmetic is identical to the general path (same fma, same roundings).
//   RELU : z = acc * acc_s + bias, gate bit = (0 < z <= hi), z = clamp(z, 0, hi)      (forward edge)
//          else z = acc                                                              (input gradient: no bias, acc_s = 1)
//   F16OUT: `out` is fp16 (else bf16);  O2K: 0 no second output, 1 bf16(z * scale2) (the weight gradient's unscaled copy),
//          2 bf16(z * scale2) where the gate bit of w2 is set, else 0 (dZ of the previous node)
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float r;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// mm |= bit  if  t != 0   (one compare + one predicated OR on the alu pipe)
__device__ __forceinline__ void or_if_nonzero(uint32_t& mm, float t, uint32_t bit) {
  asm("{\n .reg .pred q;\n setp.ne.b32 q, %1, 0;\n @q or.b32 %0, %0, %2;\n}" : "+r"(mm) : "r"(__float_as_uint(t)), "r"(bit));
}
// RELU (forward edge).  ncu r2: the epilogue, not the MMAs, bounds this kernel, and within it the ALU pipe (one instruction per
// 2 cycles and scheduler): 11.6 alu instructions per element in the runtime-configurable path (ReLU20 = 2 FMNMX, gate bit =
// VIADD + ISETP + SEL + LOP3, bit tests and selects of the second output, two conversions).  Here the clamp moves to the FMA
// pipe and the gate bit costs two alu instructions:
//   u = fma.sat(acc, acc_s / hi, bias / hi) = clamp(z / hi, 0, 1)        (the caller passes bias / hi and acc_s / hi)
//   gate bit = (u - u*u != 0)  <=>  0 < u < 1        (one FFMA, one compare, one predicated OR)
//   out = u * hi, second output = u * (hi * scale2)  (packed FMUL2, then one conversion per pair)
// Differences from the clamp-based form: out differs by <= 2 ulp of fp32 (invisible after the 16-bit rounding) and z == hi
// EXACTLY (or within half an ulp below it) counts as clamped (gate 0; the reference's clamp_max_ passes the gradient at
// z == 20) -- a set of measure zero; the gate-bit tests allow 1e-3 of mismatches against the fp32 formula.
// else (input gradient: no bias, acc_s = 1): z = acc.
template <bool RELU, bool F16OUT, int O2K>
__device__ __forceinline__ void tile_fast(const float* v, const float4* bias4, float acc_s, float relu_hi, uint32_t hi_bits,
                                          const uint32_t* w2, float scale2, uint32_t* po, uint32_t* po2, uint32_t* m,
                                          const nbasr_epilogue& epi, int64_t eidx, bool cg) {
  const u64 sc2 = pk2(scale2, scale2);
  const float a_hi = acc_s / relu_hi;
  const u64 hi2 = pk2(relu_hi, relu_hi), hs2 = pk2(relu_hi * scale2, relu_hi * scale2);
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    float z[8];
    uint32_t mm = 0xffu;
    if (RELU) {
      const float4 b0 = bias4[2 * g], b1 = bias4[2 * g + 1];
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      mm = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        z[i] = fma_sat(v[g * 8 + i], a_hi, bb[i]);
        or_if_nonzero(mm, fmaf(-z[i], z[i], z[i]), 1u << i);
      }
      m[g] = mm;
      if (epi.n_add == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a, b;
          upk2(mul2(pk2(z[2 * k], z[2 * k + 1]), hi2), a, b);
          po[g * 4 + k] = F16OUT ? f2_to_f16x2(a, b) : f2_to_bf16x2(a, b);
          if (O2K == 1) {
            upk2(mul2(pk2(z[2 * k], z[2 * k + 1]), hs2), a, b);
            po2[g * 4 + k] = f2_to_bf16x2(a, b);
          }
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] *= relu_hi;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = v[g * 8 + i];
      m[g] = mm;
    }
    for (int a = 0; a < epi.n_add; ++a) add8_cg(epi.add[a], epi.add_dtype, eidx + g * 8, z, cg);      // skip-connection sum
#pragma unroll
    for (int k = 0; k < 4; ++k) po[g * 4 + k] = F16OUT ? f2_to_f16x2(z[2 * k], z[2 * k + 1]) : f2_to_bf16x2(z[2 * k], z[2 * k + 1]);
    if (O2K == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a, b;
        upk2(mul2(pk2(z[2 * k], z[2 * k + 1]), sc2), a, b);
        po2[g * 4 + k] = f2_to_bf16x2(a, b);
      }
    } else if (O2K == 2) {
      const uint32_t w = w2[g];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = ((w >> (2 * k)) & 1u) ? z[2 * k] * scale2 : 0.f;
        const float b = ((w >> (2 * k + 1)) & 1u) ? z[2 * k + 1] * scale2 : 0.f;
        po2[g * 4 + k] = f2_to_bf16x2(a, b);
      }
    }
  }
}

__global__ void __launch_bounds__(CH_THREADS, 2)
gconv_chain_kernel(const __grid_constant__ GcChainMaps maps, const __grid_constant__ GcChainArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
  const int NS = p.nstage;
  const uint32_t wsm = base;
  const uint32_t asm0 = base + p.maxtaps * WTAP_BYTES;
  const uint32_t osm = asm0 + NS * A_BYTES;          // out staging, then out2 staging
  const uint32_t bar0 = osm + 2 * OSTAGE_BYTES;
  uint8_t* ost = al + p.maxtaps * WTAP_BYTES + NS * A_BYTES;
  uint8_t* mst = ost + 2 * OSTAGE_BYTES + 256;    // 128 x 8-byte gate-bit entries
  const uint32_t wbar = bar0;
  auto full_bar = [&](int s) { return bar0 + 8u * (1 + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (1 + 4 + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (1 + 8 + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (1 + 8 + NACC + s); };
  const uint32_t wfree = bar0 + 8u * (1 + 8 + 2 * NACC);       // "every MMA of this node has read the weight tile"
  uint32_t* tptr = reinterpret_cast<uint32_t*>(ost + 2 * OSTAGE_BYTES + 8 * (2 + 8 + 2 * NACC));
  // number of this CTA's tiles (node-major order) whose stores have landed: the producer polls THIS word for the tiles of
  // its own range (a global acquire per dependency, ~3 L2 round trips per tile, made the producer the bottleneck: r2)
  const uint32_t own_prog = bar0 + 8u * (2 + 8 + 2 * NACC) + 8u;
  // warp roles: 0..7 epilogue, 8 producer, 9 MMA issue (highest id = issue priority; warp-uniform MMA loop, see gconv_sm100.cu)
  constexpr int W_PROD = CH_NEPI / 32, W_MMA = CH_NEPI / 32 + 1;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int slab = blockIdx.x % p.nslabs;
  const int lane_id = blockIdx.x / p.nslabs;
  const int c0 = slab * p.OUT;
  // contiguous range of frame tiles: only its two ends depend on other CTAs
  const bool strided = (p.dbg & 4) != 0;       // (timing experiment only: the strided assignment of gconv_mma_fwd_kernel)
  const int tb = strided ? lane_id : (int)((int64_t)lane_id * p.ntiles / p.nlanes);
  const int te = strided ? p.ntiles : (int)((int64_t)(lane_id + 1) * p.ntiles / p.nlanes);
  const int tstep = strided ? p.nlanes : 1;

  if (warp == W_PROD && lane == 0) {
    prefetch_tmap(&maps.x[0]);
    prefetch_tmap(&maps.w[0]);
    mbar_init(wbar, 1);
    mbar_init(wfree, 1);
    st_release_cta_shared(own_prog, 0u);
    for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < NACC; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), CH_NEPI); }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(tptr), 256);
  pdl_launch_dependents();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = *tptr;
  auto load_weights = [&](int nd) {
    const int kt = p.node[nd].ktaps;
    mbar_expect_tx(wbar, kt * WTAP_BYTES);
    for (int j = 0; j < kt; ++j) tma_load_2d(wsm + j * WTAP_BYTES, &maps.w[nd], wbar, 0, (slab * kt + j) * NW);
  };
  if (p.w_stable && warp == W_PROD && lane == 0) load_weights(0);
  pdl_wait();                                  // everything above overlapped the previous kernel's tail
  if (!p.w_stable && warp == W_PROD && lane == 0) load_weights(0);
  // flag value of THIS launch (the epoch only advances when every CTA of a launch has left)
  const uint32_t done = *reinterpret_cast<volatile uint32_t*>(p.work) + 1u;
  uint32_t* flags = p.work + CH_WORK_HDR + (int64_t)slab * p.ntiles;
  const int64_t node_flags = (int64_t)p.nslabs * p.ntiles;

  if (warp == W_PROD) {
    int stage = 0;
    uint32_t phase = 0;
    int known_prog = 0;        // (lane 0) own tiles known to have landed AND already ordered by a proxy fence
    for (int nd = 0; nd < p.n_nodes; ++nd) {
      const GcChainNode& N = p.node[nd];
      const int esz = 2;
      const bool pf_mask = N.epi.out2 && N.epi.mask2;
      const int pl2 = pf_mask ? c0 / N.epi.mask2_w : 0;
      const int eb2 = N.epi.mask2_w == 32 ? 4 : 8;
      for (int tile = tb; tile < te; tile += tstep) {
        const int tu = tile % p.tiles_per_utt;
        const int b = tile / p.tiles_per_utt, t0 = tu * GT;
        if (lane == 0) {
          if (nd > 0 && !(p.dbg & 3)) {
            // tiles of the previous node's output that the input window [t0 + off0, t0 + off0 + 128 + (k-1) dstep) touches
            const int r0 = t0 + N.off0, r1 = r0 + GT - 1 + (N.ktaps - 1) * N.dstep;
            const int tlo = max(0, r0) / GT, thi = min(p.tiles_per_utt - 1, r1 / GT);
            const uint32_t* f = flags + (nd - 1) * node_flags;
            int own_need = 0;                        // tiles of my own range: one shared-memory poll for the latest of them
            bool fence = false;
            for (int tt = tlo; tt <= thi; ++tt) {
              const int gt = tile - tu + tt;
              if (gt >= tb && gt < te) {
                own_need = max(own_need, (nd - 1) * (te - tb) + (gt - tb) + 1);
              } else {                               // a neighbour's first / last tile
                wait_flag(f + gt, done, p.work + 2);
                fence = true;
              }
            }
            if (own_need > known_prog) {
              // the poll returns the CURRENT progress, usually far ahead of the requirement: once the previous node has
              // drained, the following tiles of this node need neither a poll nor a fence
              uint32_t spins = 0;
              while ((known_prog = (int)ld_acquire_cta_shared(own_prog)) < own_need)
                if (++spins > (1u << 28)) { atomicExch(p.work + 2, 2u); break; }
              fence = true;
            }
            if (fence) fence_proxy_async_global();
          }
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), A_BYTES);
          tma_load_3d(asm0 + stage * A_BYTES, &maps.x[nd], full_bar(stage), c0, NBASR_PAD_L + t0 + N.off0, b);
          if (nd > 0 && tile == tb) {
            mbar_wait(wfree, (nd - 1) & 1);       // the previous node's MMAs are done with the weight tile
            load_weights(nd);
          }
        }
        __syncwarp();
        if (!p.no_prefetch) {
          const int64_t rho0 = (int64_t)b * p.Tp + NBASR_PAD_L + t0;
          const int nr = min(GT, p.T - t0);
          const int ncols = min(p.OUT, p.C - c0);
          for (int a = 0; a < N.epi.n_add; ++a) {
            const char* ab = reinterpret_cast<const char*>(N.epi.add[a]) + (rho0 * N.epi.ld_out + c0) * esz;
            for (int r = lane; r < nr; r += 32) {
              const char* q = ab + (int64_t)r * N.epi.ld_out * esz;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(q + ncols * esz - 1));
            }
          }
          if (pf_mask) {
            const char* m0 = reinterpret_cast<const char*>(N.epi.mask2) + ((int64_t)pl2 * N.epi.mask_rows + rho0) * eb2;
            const char* m1 = reinterpret_cast<const char*>(N.epi.mask2) +
                             ((int64_t)((c0 + ncols - 1) / N.epi.mask2_w) * N.epi.mask_rows + rho0) * eb2;
            const int nb = nr * eb2;
            for (int o = lane * 128; o < nb + 127; o += 32 * 128) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(m0 + min(o, nb - 1)));
              if (m1 != m0) asm volatile("prefetch.global.L2 [%0];" ::"l"(m1 + min(o, nb - 1)));
            }
          }
        }
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == W_MMA) {
    const uint32_t idesc = make_idesc(128, NW, 0, 0, p.f16);
    const uint64_t bd0 = make_smem_desc(wsm, 16, 1024);
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int nd = 0; nd < p.n_nodes; ++nd) {
      const int kt = p.node[nd].ktaps, ds = p.node[nd].dstep;
      mbar_wait(wbar, nd & 1);
      for (int tile = tb; tile < te; tile += tstep, ++it) {
        const int as = it % NACC;
        const uint32_t aphase = (it / NACC) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
        const uint64_t ad0 = make_smem_desc(asm0 + stage * A_BYTES, 16, 1024);
        if (elect_one()) {
          for (int j = 0; j < kt; ++j) {
#pragma unroll
            for (int k = 0; k < NW / 16; ++k)
              umma_bf16(tm + as * 64, ad0 + (uint64_t)(j * ds * 8 + k * 2), bd0 + (uint64_t)(j * (WTAP_BYTES >> 4) + k * 2), idesc, (j | k) != 0);
          }
          umma_commit(empty_bar(stage));
          umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(wfree);
      __syncwarp();
    }
  } else {
    const int ew = warp;                     // 0..7
    const int q = warp & 3;                  // TMEM lane quadrant of this warp
    const int hh = ew >> 2;                  // column half: cols [24*hh, 24*hh + 24)
    const int etid = threadIdx.x;
    const int row = q * 32 + lane;
    const int cbeg = c0 + 24 * hh;
    const int nvalid = max(0, min(24, min(p.C, c0 + p.OUT) - cbeg));     // multiple of 8
    const int OUTB = p.OUT * 2;              // staged row pitch in bytes
    const bool cg_adds = (OUTB & 31) != 0;   // slabs that share 32-byte sectors with their neighbours (see add8_cg)
    int it = 0;
    // (thread etid 0) Landing of the TMA stores is observed in BATCHES: `cp.async.bulk.wait_group` (the full-completion
    // form) compiles to DEPBAR + CCTL.IVALL -- it invalidates the SM's whole L1 -- so it runs once per PUBLISH_EVERY tiles
    // (wait_group 1: every group but the newest has landed) and once at the end of a node.  The consumer of node n's tiles
    // is node n + 1, a whole pass later: batches never stall it (ncu r2: per-tile waits cost 16 % more instructions issued
    // and every register-spill reload missed L1).
    constexpr uint32_t PUBLISH_EVERY = 4;
    uint32_t n_issued = 0, n_landed = 0;     // tiles (node-major) whose stores were committed / are known to have landed
    uint32_t* first_flag = nullptr;          // global flag of this node's first tile while unpublished
    for (int nd = 0; nd < p.n_nodes; ++nd) {
      const nbasr_epilogue& epi = p.node[nd].epi;     // stays in the kernel-parameter bank
      const bool more = nd + 1 < p.n_nodes;
      const float acc_s = epi_acc_scale(epi), bias_s = epi_bias_scale(epi), relu_hi = epi_relu_hi(epi);
      const uint32_t hi_bits = __float_as_uint(relu_hi);
      // specialised epilogue (tile_fast) when this node's configuration is one of the six the step uses; -1: general path
      int mode_cta = -1;                  // CTA-uniform: decides the scaling of the shared bias table
      if (epi.drop_p == 0.f && (epi.n_add == 0 || epi.add_dtype != NBASR_F32) && epi.out && !(p.dbg & 16) && (!epi.out2 || epi.out2_dtype == NBASR_BF16)) {
        const int o2k = !epi.out2 ? 0 : (epi.mask2 ? 2 : 1);
        if (epi.relu20 && o2k != 2) mode_cta = (epi.out_dtype == NBASR_F16 ? 0 : 2) + o2k;              // 0..3
        else if (!epi.relu20 && !epi.bias && acc_s == 1.f && epi.out_dtype == NBASR_BF16 && o2k != 1) mode_cta = 4 + (o2k >> 1);   // 4, 5
      }
      const int mode = nvalid == 24 ? mode_cta : -1;       // threads whose 24 columns are not all valid take the general path
      // slab bias (scaled) in shared memory, double-buffered by node parity: 24 registers less than a per-thread copy
      // (the register copy spilled, and the reloads sat behind the TMEM load of every tile).  The ReLU fast modes work on
      // z / hi (tile_fast): their table holds bias / hi.
      float* bsm = reinterpret_cast<float*>(mst + 1024) + (nd & 1) * 64;
      const bool bias_by_hi = mode_cta >= 0 && mode_cta < 4;
      const float bias_f = bias_by_hi ? bias_s / relu_hi : bias_s;
      const float bias_back = bias_by_hi ? relu_hi : 1.f;      // general-path threads of such a CTA undo the table's scaling
      if (etid < NW) bsm[etid] = (epi.bias && c0 + etid < min(p.C, c0 + p.OUT)) ? __ldg(epi.bias + c0 + etid) * bias_f : 0.f;
      named_bar_sync(1, CH_NEPI);
      const float4* bias4 = reinterpret_cast<const float4*>(bsm + 24 * hh);
      for (int tile = tb; tile < te; tile += tstep, ++it) {
        const int as = it % NACC;
        const uint32_t aphase = (it / NACC) & 1;
        const int b = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * GT;
        const int t = t0 + row;
        // gate bits of the second output (input gradient: dZ of the previous node), requested before the accumulator wait
        uint32_t w2[3] = {0xffu, 0xffu, 0xffu};
        if (epi.out2 && epi.mask2 && t < p.T) {
          const int64_t rho2 = (int64_t)b * p.Tp + NBASR_PAD_L + t;
#pragma unroll
          for (int g = 0; g < 3; ++g)
            if (g * 8 < nvalid) w2[g] = reinterpret_cast<const uint8_t*>(epi.mask2)[mask_byte_addr(rho2, cbeg + g * 8, epi.mask2_w, epi.mask_rows)];
        }
        // the previous tile's stores have had a whole tile time to land: publish its flag (off the critical path)
        if (etid == 0 && n_issued - n_landed > PUBLISH_EVERY && !(p.dbg & 1)) {
          asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");      // every group but the newest has landed
          n_landed = n_issued - 1;
          if (first_flag) {                  // first tile of the range: the neighbour's last tile reads its halo
            fence_proxy_async_global();
            st_release_gpu(first_flag, done);
            first_flag = nullptr;
          }
          st_release_cta_shared(own_prog, n_landed);
        }
        mbar_wait(tfull_bar(as), aphase);
        tcgen05_fence_after();
        float v[24];
        const uint32_t ta = tm + ((uint32_t)(q * 32) << 16) + as * 64 + 24 * hh;
        tmem_ld16_nowait(ta, v);
        tmem_ld8_nowait(ta + 16, v + 16);
        tmem_ld_wait();
        tcgen05_fence_before();
        mbar_arrive(tempty_bar(as));           // accumulator is in registers: release the TMEM stage early
        const bool rowok = t < p.T;
        const int64_t rho = (int64_t)b * p.Tp + NBASR_PAD_L + t;
        if (p.dbg & 512) continue;          // (timing experiment: the epilogue only drains the accumulator)
        uint32_t m[3] = {0, 0, 0};
        uint32_t po[12], po2[12];            // this thread's 24 columns of `out` / `out2`, packed 16-bit pairs
        if (p.dbg & 32) {                    // (timing experiment: no epilogue arithmetic)
#pragma unroll
          for (int i = 0; i < 12; ++i) { po[i] = __float_as_uint(v[2 * i]); po2[i] = __float_as_uint(v[2 * i + 1]); }
        } else if (rowok && mode >= 0) {
          switch (mode) {
            case 0: tile_fast<true, true, 0>(v, bias4, acc_s, relu_hi, hi_bits, w2, epi.scale2, po, po2, m, epi, rho * epi.ld_out + cbeg, cg_adds); break;
            case 1: tile_fast<true, true, 1>(v, bias4, acc_s, relu_hi, hi_bits, w2, epi.scale2, po, po2, m, epi, rho * epi.ld_out + cbeg, cg_adds); break;
            case 2: tile_fast<true, false, 0>(v, bias4, acc_s, relu_hi, hi_bits, w2, epi.scale2, po, po2, m, epi, rho * epi.ld_out + cbeg, cg_adds); break;
            case 3: tile_fast<true, false, 1>(v, bias4, acc_s, relu_hi, hi_bits, w2, epi.scale2, po, po2, m, epi, rho * epi.ld_out + cbeg, cg_adds); break;
            case 4: tile_fast<false, false, 0>(v, bias4, acc_s, relu_hi, hi_bits, w2, epi.scale2, po, po2, m, epi, rho * epi.ld_out + cbeg, cg_adds); break;
            default: tile_fast<false, false, 2>(v, bias4, acc_s, relu_hi, hi_bits, w2, epi.scale2, po, po2, m, epi, rho * epi.ld_out + cbeg, cg_adds); break;
          }
        } else {
          if (rowok) {
#pragma unroll
            for (int g = 0; g < 6; ++g) {
              const float4 bq = bias4[g];
              v[4 * g] = fmaf(v[4 * g], acc_s, bq.x * bias_back);
              v[4 * g + 1] = fmaf(v[4 * g + 1], acc_s, bq.y * bias_back);
              v[4 * g + 2] = fmaf(v[4 * g + 2], acc_s, bq.z * bias_back);
              v[4 * g + 3] = fmaf(v[4 * g + 3], acc_s, bq.w * bias_back);
            }
            if (nvalid == 24) epilogue_compute<24, true, true, true>(epi, rho, cbeg, 24, v, m);
            else epilogue_compute<24, false, true, true>(epi, rho, cbeg, nvalid, v, m);
            for (int a = 0; a < epi.n_add; ++a) {
#pragma unroll
              for (int g = 0; g < 3; ++g)
                if (g * 8 < nvalid) add8_cg(epi.add[a], epi.add_dtype, rho * epi.ld_out + cbeg + g * 8, v + g * 8, cg_adds);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 24; ++i) v[i] = 0.f;      // rows past the utterance land on zero pad rows / are clipped
          }
          const bool o16 = epi.out_dtype == NBASR_F16, o216 = epi.out2_dtype == NBASR_F16;
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const uint32_t w = w2[g];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float a = v[g * 8 + 2 * k], b2 = v[g * 8 + 2 * k + 1];
              po[g * 4 + k] = o16 ? f2_to_f16x2(a, b2) : f2_to_bf16x2(a, b2);
              const float a2 = ((w >> (2 * k)) & 1u) ? a * epi.scale2 : 0.f, c2 = ((w >> (2 * k + 1)) & 1u) ? b2 * epi.scale2 : 0.f;
              po2[g * 4 + k] = o216 ? f2_to_f16x2(a2, c2) : f2_to_bf16x2(a2, c2);
            }
          }
        }
        // staging buffers are free once the previous tile's TMA stores have finished READING shared memory
        if (etid == 0) bulk_wait_read0();
        named_bar_sync(1, CH_NEPI);
        uint8_t* orow = ost + row * OUTB + 48 * hh;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          if (g * 8 < nvalid) {
            *reinterpret_cast<uint4*>(orow + g * 16) = make_uint4(po[g * 4], po[g * 4 + 1], po[g * 4 + 2], po[g * 4 + 3]);
            if (epi.out2)
              *reinterpret_cast<uint4*>(orow + OSTAGE_BYTES + g * 16) = make_uint4(po2[g * 4], po2[g * 4 + 1], po2[g * 4 + 2], po2[g * 4 + 3]);
          }
          if (epi.mask_out) mst[row * 8 + 3 * hh + g] = (g * 8 < nvalid) ? (uint8_t)m[g] : (uint8_t)0;   // 8-byte entry per row
        }
        fence_async_smem();
        named_bar_sync(1, CH_NEPI);
        if (epi.mask_out && etid < GT && t0 + etid < p.T && !(p.dbg & 256)) {
          // 128 consecutive 8-byte entries of this slab's mask plane: one fully coalesced store per warp
          const int64_t r2 = (int64_t)b * p.Tp + NBASR_PAD_L + t0 + etid;
          reinterpret_cast<uint64_t*>(epi.mask_out)[(int64_t)slab * epi.mask_rows + r2] = reinterpret_cast<const uint64_t*>(mst)[etid];
        }
        if (etid == 0) {
          if (epi.out && !(p.dbg & 64)) tma_store_3d(&maps.o[nd], osm, c0, NBASR_PAD_L + t0, b);
          if (epi.out2 && !(p.dbg & (64 | 128))) tma_store_3d(&maps.o2[nd], osm + OSTAGE_BYTES, c0, NBASR_PAD_L + t0, b);
          bulk_commit();
          if (more && !(p.dbg & 1)) {
            ++n_issued;
            uint32_t* f = flags + nd * node_flags + tile;
            if (tile == tb) first_flag = f;
            if (tile + tstep >= te) {          // last tile of the node: everything of this node lands and is published now
              bulk_wait0();
              n_landed = n_issued;
              fence_proxy_async_global();
              if (first_flag) st_release_gpu(first_flag, done);
              first_flag = nullptr;
              if (tile != tb) st_release_gpu(f, done);      // (the neighbour's first tile reads the halo behind this one)
              st_release_cta_shared(own_prog, n_landed);
            }
          }
        }
      }
    }
    if (etid == 0) bulk_wait0();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tcgen05_fence_after();
    tmem_dealloc(tm, 256);
  }
  if (warp == W_PROD && lane == 0) {
    // exit ticket: the last CTA advances the epoch (every flag of this launch becomes stale) and re-arms the ticket
    __threadfence();
    const uint32_t tk = atomicAdd(p.work + 1, 1u);
    if (tk == gridDim.x - 1) {
      p.work[1] = 0;
      __threadfence();
      atomicExch(p.work, done);
    }
  }
}

}  // namespace

int64_t sm100_gconv_chain_work_bytes(int B, int T, int C, int cpg, int n) {
  const int OUT = slab_out(cpg);
  const int64_t nslabs = (C + OUT - 1) / OUT, ntiles = (int64_t)((T + GT - 1) / GT) * B;
  return (CH_WORK_HDR + (int64_t)n * nslabs * ntiles) * 4;
}

int sm100_gconv_chain(const nbasr_gconv* g, int n, int fused, void* work, int64_t work_bytes, cudaStream_t st) {
  NBASR_REQUIRE(n >= 1 && n <= MAXCHAIN, "chain length");
  if (!fused || n == 1) {       // one launch of gconv_mma_fwd_kernel per node (programmatic dependent launches)
    for (int i = 0; i < n; ++i)
      if (sm100_gconv_fwd(g + i, st)) return 1;
    return 0;
  }
  GcChainArgs a{};
  GcChainMaps maps;
  const nbasr_gconv& g0 = g[0];
  a.B = g0.B; a.T = g0.T; a.Tp = g0.Tp; a.C = g0.C; a.OUT = slab_out(g0.cpg);
  a.n_nodes = n;
  a.f16 = g0.dtype == NBASR_F16 ? 1 : 0;
  a.nslabs = (g0.C + a.OUT - 1) / a.OUT;
  a.tiles_per_utt = (g0.T + GT - 1) / GT;
  a.ntiles = a.tiles_per_utt * g0.B;
  const int slots = GCONV_SLOTS * nbasr_sm_count();
  a.nlanes = std::max(1, std::min(a.ntiles, slots / a.nslabs));
  a.no_prefetch = nbasr_env_flag(NBASR_ENV_GCONV_NO_PREFETCH) ? 1 : 0;
  a.w_stable = (g0.w_packed & 2) ? 1 : 0;
  a.work = reinterpret_cast<uint32_t*>(work);
  a.dbg = nbasr_env_chain_dbg();      // timing experiments only (results are wrong when set)
  NBASR_REQUIRE(work != nullptr && work_bytes >= sm100_gconv_chain_work_bytes(g0.B, g0.T, g0.C, g0.cpg, n),
                "chain work buffer (nbasr_gconv_chain_work_bytes, zero-initialised once)");
  uint64_t dx[3] = {(uint64_t)g0.C, (uint64_t)g0.Tp, (uint64_t)g0.B};
  int64_t sx[3] = {1, g0.C, (int64_t)g0.Tp * g0.C};
  uint32_t bx[3] = {64, AROWS, 1};
  uint32_t bo[3] = {(uint32_t)a.OUT, GT, 1};
  for (int i = 0; i < n; ++i) {
    const nbasr_gconv& gi = g[i];
    NBASR_REQUIRE((gi.w_packed & 1) && gi.dtype == g0.dtype && gi.B == g0.B && gi.T == g0.T && gi.Tp == g0.Tp && gi.C == g0.C &&
                      gi.cpg == g0.cpg, "chain nodes share dtype and geometry");
    NBASR_REQUIRE(gi.off0 >= -NBASR_PAD_L && (gi.ktaps - 1) * gi.dstep <= AROWS - GT, "tap reach");
    NBASR_REQUIRE(gi.epi.ld_out == gi.C, "grouped conv writes dense (B,Tp,C) tensors");
    NBASR_REQUIRE((!gi.epi.out || gi.epi.out_dtype != NBASR_F32) && (!gi.epi.out2 || gi.epi.out2_dtype != NBASR_F32) &&
                      !gi.epi.accumulate, "tcgen05 grouped conv stores 16-bit tensors");
    NBASR_REQUIRE(gi.epi.n_add == 0 || gi.epi.add_dtype != NBASR_F32, "tcgen05 grouped conv adds 16-bit skip tensors");
    NBASR_REQUIRE(!gi.epi.mask_out || gi.epi.mask_w == a.OUT, "grouped-conv mask planes are slab wide");
    if (i > 0)
      NBASR_REQUIRE(gi.x != nullptr && (gi.x == g[i - 1].epi.out || gi.x == g[i - 1].epi.out2),
                    "node i+1 of a chain reads an output of node i");
    a.maxtaps = std::max(a.maxtaps, gi.ktaps);
    a.node[i].ktaps = gi.ktaps; a.node[i].dstep = gi.dstep; a.node[i].off0 = gi.off0;
    a.node[i].epi = gi.epi;
    if (sm100_get_map(gi.x, 3, dx, sx, bx, &maps.x[i])) return 1;
    uint64_t dw[2] = {64, (uint64_t)a.nslabs * gi.ktaps * NW};
    int64_t sw[2] = {1, 64};
    uint32_t bw[2] = {64, NW};
    if (sm100_get_map(gi.w, 2, dw, sw, bw, &maps.w[i])) return 1;
    const void* o1 = gi.epi.out ? gi.epi.out : gi.x;       // unused maps still need a valid descriptor
    const void* o2 = gi.epi.out2 ? gi.epi.out2 : gi.x;
    if (sm100_get_map(o1, 3, dx, sx, bo, &maps.o[i], 0)) return 1;
    if (sm100_get_map(o2, 3, dx, sx, bo, &maps.o2[i], 0)) return 1;
  }
  for (int i = n; i < MAXCHAIN; ++i) { maps.x[i] = maps.x[0]; maps.w[i] = maps.w[0]; maps.o[i] = maps.o[0]; maps.o2[i] = maps.o2[0]; }
  a.nstage = fwd_nstage(a.maxtaps);
  // ring + staging + barriers (256) + gate-bit staging (1024) + two slab-bias buffers (512) + alignment slack (1024)
  const size_t smem = (size_t)a.maxtaps * WTAP_BYTES + (size_t)a.nstage * A_BYTES + 2 * OSTAGE_BYTES + 256 + 1024 + 512 + 1024;
  static DevOnce attr;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gconv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BUDGET);
    if (e != cudaSuccess) return nbasr_fail("gconv_chain smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  cudaError_t le = launch_pdl(gconv_chain_kernel, dim3(a.nslabs * a.nlanes), dim3(CH_THREADS), smem, st, 1, maps, a);
  if (le != cudaSuccess) return nbasr_fail("gconv_chain launch: %s", cudaGetErrorString(le));
  return 0;
}
