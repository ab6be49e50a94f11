// Micro-experiment kernel (test-only entry point): does a UMMA shared-memory descriptor whose start
// address is shifted by whole 128-byte rows inside a 128B-swizzled tile address the right data?
// mode 0: K-major A tile of (128+pad) rows x 64 bf16; D[m][n] = sum_k A[m+shift][k] * B[n][k]
// mode 1: MN-major: D[m][n] = sum_{k<64} Y[k][m] * X[k+shift][n]   (Y: 64 x 128, X: (64+pad) x 64)
// bo_mode 0: base_offset field = 0; 1: base_offset = (start_addr >> 7) & 7.
// mode 2: B operand in the NO-SWIZZLE K-major canonical layout [k-chunk of 8][row group][8 rows][8 elems], N = 16,
//         K = 64, filled by plain st.shared (the LSTM kernel's h operand): D[m][n] = sum_k A[m][k] * B[n][k]
#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int PADROWS = 16;

__global__ void __launch_bounds__(128, 1)
dbg_shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, int mode_in, int shift,
                 int bo_mode) {
  // modes 3/4/5: mode-0 layout with operand formats (A,B) = (f16,bf16) / (bf16,f16) / (f16,f16) in the instruction descriptor
  const int mode = mode_in >= 3 ? 0 : mode_in;
  const uint32_t fmt_clear = (mode_in == 3 || mode_in == 5 ? (7u << 7) : 0u) | (mode_in == 4 || mode_in == 5 ? (7u << 10) : 0u);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sa = base;                    // up to 144 rows x 128 B = 18 KB  (mode 1: Y, 2 boxes of 64x64)
  const uint32_t sb = base + 32768;            // B / X tile
  const uint32_t bar = base + 65536;
  const uint32_t tbar = bar + 8;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(al + 65536 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(tbar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tptr), 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = *tptr;
  if (mode == 2) {
    // B (16 x 64 bf16, row-major in global, passed via D's neighbour pointer trick: tmB unused) -> canonical layout
    const bf16* Bg = reinterpret_cast<const bf16*>(D + 128 * 64);
    bf16* bs = reinterpret_cast<bf16*>(al + 32768);
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
      int n = i / 64, k = i % 64;
      bs[((k >> 3) * 2 + (n >> 3)) * 64 + (n & 7) * 8 + (k & 7)] = Bg[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (mode == 2) {
      mbar_expect_tx(bar, 128 * 128);
      tma_load_2d(sa, &tmA, bar, 0, 0);
    } else if (mode == 0) {
      mbar_expect_tx(bar, (128 + PADROWS) * 128 + 64 * 128);
      tma_load_2d(sa, &tmA, bar, 0, 0);
      tma_load_2d(sb, &tmB, bar, 0, 0);
    } else {
      mbar_expect_tx(bar, 2 * 64 * 128 + (64 + PADROWS) * 128);
      tma_load_2d(sa, &tmA, bar, 0, 0);
      tma_load_2d(sa + 8192, &tmA, bar, 64, 0);
      tma_load_2d(sb, &tmB, bar, 0, 0);
    }
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    for (int k = 0; k < 4; ++k) {
      uint64_t ad, bd;
      if (mode == 2) {
        ad = make_smem_desc(sa + k * 32, 16, 1024);
        // no swizzle: layout type 0; LBO = distance between K-adjacent 8x16B core matrices, SBO = between row groups
        bd = make_smem_desc(sb + k * 512, shift ? 128 : 256, shift ? 256 : 128) & ~((uint64_t)7 << 61);
        umma_bf16(tm, ad, bd, make_idesc(128, 16, 0, 0), k != 0);
        continue;
      }
      if (mode == 0) {
        uint32_t a0 = sa + shift * 128 + k * 32;
        ad = make_smem_desc(a0, 16, 1024, bo_mode ? (a0 >> 7) & 7 : 0);
        bd = make_smem_desc(sb + k * 32, 16, 1024);
      } else {
        uint32_t b0 = sb + shift * 128 + k * 2048;
        ad = make_smem_desc(sa + k * 2048, 8192, 1024);
        bd = make_smem_desc(b0, 8192, 1024, bo_mode ? (b0 >> 7) & 7 : 0);
      }
      umma_bf16(tm, ad, bd, make_idesc(128, 64, mode, mode) & ~fmt_clear, k != 0);
    }
    umma_commit(tbar);
  }
  mbar_wait(tbar, 0);
  tcgen05_fence_after();
  for (int c = 0; c < 64; c += 32) {
    float v[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 64 + c + i] = v[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

}  // namespace

extern "C" int nbasr_dbg_shift(const void* A, const void* B, float* D, int mode, int shift, int bo_mode, void* stream) {
  CUtensorMap tmA, tmB;
  if (mode == 2) {
    uint64_t da[2] = {64, 128};
    int64_t sa[2] = {1, 64};
    uint32_t ba[2] = {64, 128};
    if (sm100_get_map(A, 2, da, sa, ba, &tmA)) return 1;
    tmB = tmA;
  } else if (mode == 0 || mode >= 3) {
    uint64_t da[2] = {64, 128 + PADROWS};
    int64_t sa[2] = {1, 64};
    uint32_t ba[2] = {64, 128 + PADROWS};
    if (sm100_get_map(A, 2, da, sa, ba, &tmA)) return 1;
    uint64_t db[2] = {64, 64};
    uint32_t bb[2] = {64, 64};
    if (sm100_get_map(B, 2, db, sa, bb, &tmB)) return 1;
  } else {
    uint64_t da[2] = {128, 64};
    int64_t sa[2] = {1, 128};
    uint32_t ba[2] = {64, 64};
    if (sm100_get_map(A, 2, da, sa, ba, &tmA)) return 1;
    uint64_t db[2] = {64, 64 + PADROWS};
    int64_t sb[2] = {1, 64};
    uint32_t bb[2] = {64, 64 + PADROWS};
    if (sm100_get_map(B, 2, db, sb, bb, &tmB)) return 1;
  }
  cudaFuncSetAttribute(dbg_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  dbg_shift_kernel<<<1, 128, 70 * 1024, as_stream(stream)>>>(tmA, tmB, D, mode, shift, bo_mode);
  NBASR_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Micro-benchmarks (test-only): what=0: back-to-back tcgen05.mma 128 x N x 16 (SS, bf16) issue+execute cycles;
// what=1: tcgen05.ld 32x32b.x32 throughput over `nw` warps; what=2: warp-shuffle throughput over `nw` warps.
// out[cta] = elapsed clock64 cycles for `iters` operations (per warp for what=1/2).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(512, 1)
dbg_bench_kernel(unsigned long long* out, int what, int N, int iters, int variant) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sa = base;                 // 160 rows x 128 B
  const uint32_t sb = base + 24576;         // 256 rows x 128 B
  const uint32_t bar = base + 24576 + 32768;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(al + 24576 + 32768 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (24576 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(al)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tptr), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = *tptr;
  if (what == 0) {
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc(128, N, 0, 0);
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        // variant bit0: alternate accumulators; bit1: walk taps (row shift) and k-chunks like the grouped-conv kernel
        const int j = (variant & 2) ? (i / 3) % 5 : 0, k = (variant & 2) ? i % 3 : 0;
        uint64_t ad = make_smem_desc(sa + j * 128 + k * 32, 16, 1024);
        uint64_t bd = make_smem_desc(sb + k * 32, 16, 1024);
        umma_bf16(tm + ((variant & 1) ? (i & 1) * 256 : 0), ad, bd, idesc, i != 0);
      }
      long long t1 = clock64();
      umma_commit(bar);
      mbar_wait(bar, 0);
      long long t2 = clock64();
      out[blockIdx.x * 2] = t2 - t0;
      out[blockIdx.x * 2 + 1] = t1 - t0;
    }
  } else if (what == 1) {
    if (warp < N) {   // N = number of warps taking part
      float acc = 0.f;
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        float v[32];
        tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16) + ((i * 32) & 255) + (warp >> 2) * 256 % 512, v);
#pragma unroll
        for (int q = 0; q < 32; ++q) acc += v[q];
      }
      long long t1 = clock64();
      if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
      if (acc == 123.456f) out[0] = 0;
    }
  } else {
    if (warp < N) {
      float v[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = lane * 0.5f + q;
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], 1);
      }
      long long t1 = clock64();
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 32; ++q) acc += v[q];
      if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
      if (acc == 123.456f) out[0] = 0;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tm, 512);
  }
}
}  // namespace

extern "C" int nbasr_dbg_bench(unsigned long long* out, int what, int N, int iters, int variant, int ctas, void* stream) {
  cudaFuncSetAttribute(dbg_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  dbg_bench_kernel<<<ctas, 512, 60 * 1024, as_stream(stream)>>>(out, what, N, iters, variant);
  NBASR_CHECK_LAUNCH();
  return 0;
}
