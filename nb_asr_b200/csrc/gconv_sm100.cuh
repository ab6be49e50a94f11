// Geometry shared by the tcgen05 grouped-conv kernels (gconv_sm100.cu, gconv_chain_sm100.cu).
#pragma once
#include "sm100_ptx.cuh"

namespace {
using namespace sm100;

constexpr int GT = 128;                    // frames per tile
constexpr int AROWS = 144;                 // 128 + max halo 12, multiple of 8
constexpr int A_BYTES = AROWS * 128;       // 18432
constexpr int NW = 48;                     // MMA N (and K window) of a slab
constexpr int WTAP_BYTES = NW * 128;       // 6144
constexpr int NSTAGE = 3;                  // (weight-gradient kernel)
constexpr int NACC = 4;
constexpr int GC_THREADS = 192;            // weight-gradient kernel: 4 drain warps, producer, MMA
constexpr int FWD_THREADS = 448;           // forward kernel: 12 epilogue warps, producer, MMA
constexpr int NEPI = 384;                  // (three column thirds x four TMEM lane quadrants)
constexpr int OSTAGE_BYTES = GT * NW * 2;  // 12288: bf16 output tile staged for the TMA store
constexpr int FWD_SMEM_BUDGET = 112 * 1024;

__host__ __device__ inline int slab_out(int cpg) { return cpg == 10 ? 40 : 48; }
// weight-gradient slabs: as many whole groups as fit in 56 channels: 54 / 56 / 50 / 48 for cpg 6 / 8 / 10 / 12 -> fewer,
// fuller 128-byte loads.  TMA boxes must start on a 16-byte (8-channel) boundary, so the box starts at c0 & ~7 and the
// slab sits at offset c0 & 7 inside the 64-channel window; channel 63 of the window is never part of a slab and carries
// the ones column of the fused bias gradient.
__host__ __device__ inline int wg_slab_out(int cpg) { return (56 / cpg) * cpg; }
__host__ __device__ inline int fwd_nstage(int ktaps) {
  int n = (FWD_SMEM_BUDGET - 3072 - ktaps * WTAP_BYTES - 2 * OSTAGE_BYTES) / A_BYTES;
  return n > 4 ? 4 : n;
}


constexpr int GCONV_SLOTS = 2;   // resident CTAs per SM the grouped-conv kernels size their persistent grids for
}  // namespace
