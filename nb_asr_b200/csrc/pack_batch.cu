// One-launch refresh of every derived weight operand (bf16 copies, transposes, tap-flipped and block-diagonal
// packs): a device-resident job table, each block handles one 4096-element chunk of one job.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int PB_CHUNK = 4096;

// store one value as out_dtype (F32 / BF16 / F16)
__device__ __forceinline__ void put_dt(void* dst, int dtype, int64_t i, float v) { st1_dt(dst, dtype, i, v); }

__global__ void __launch_bounds__(256) pack_batch_kernel(const nbasr_pack_job* __restrict__ jobs, int n,
                                                         const int2* __restrict__ blockmap) {
  // blockmap[b] = (job, chunk) -- precomputed by the host that built the table (no per-block search)
  const int2 bm = blockmap[blockIdx.x];
  const int j = bm.x;
  const int64_t blk = bm.y;
  if (j >= n) return;
  const nbasr_pack_job J = jobs[j];
  if (J.kind == 1) {
    // transpose job: chunk = one 64 (m) x 64 (n) tile of tap q, staged through shared memory so that both the fp32 reads
    // (along n, ws_n == 1) and the writes (along m) are coalesced.  chunk -> (q, tile_m, tile_n), tile_n fastest.
    __shared__ float tile[64][65];
    const int M = J.a[0], N = J.a[1], nq = J.a[2], t0 = J.a[3], ts = J.a[4];
    const int tn = (N + 63) >> 6, tm = (M + 63) >> 6;
    const int q = (int)(blk / ((int64_t)tm * tn));
    const int m0 = (int)((blk / tn) % tm) * 64, n0 = (int)(blk % tn) * 64;
    const float* src = J.src + (int64_t)(t0 + q * ts) * J.s[2];
    const int c = threadIdx.x & 63, r4 = threadIdx.x >> 6;
    for (int r = r4; r < 64; r += 4)
      tile[r][c] = (m0 + r < M && n0 + c < N) ? src[(int64_t)(m0 + r) * J.s[0] + (int64_t)(n0 + c) * J.s[1]] : 0.f;
    __syncthreads();
    for (int r = r4; r < 64; r += 4) {          // r = n within the tile, c = m within the tile
      if (n0 + r < N && m0 + c < M) {
        const int64_t o = (int64_t)(n0 + r) * nq * M + (int64_t)q * M + m0 + c;
        put_dt(J.dst, J.out_dtype, o, tile[c][r]);
      }
    }
    return;
  }
  if (J.kind == 2) {
    // block-diagonal grouped-conv operand: only the C*cpg*ktaps weights are (re)written -- the off-diagonal zeros of the
    // [slab][tap][48][64] pack are written once when the buffer is allocated.  idx walks the SOURCE tensor (co, i, tap).
    const int Cc = J.a[0], cpg = J.a[1], ktaps = J.a[2], tr = J.a[3];
    const int OUT = cpg == 10 ? 40 : 48;
    const int64_t nsrc = (int64_t)Cc * cpg * ktaps;
    const int64_t s0 = blk * PB_CHUNK, s1 = min(nsrc, s0 + PB_CHUNK);
    for (int64_t idx = s0 + threadIdx.x; idx < s1; idx += blockDim.x) {
      const int jt = (int)(idx % ktaps);
      const int i = (int)((idx / ktaps) % cpg);
      const int co = (int)(idx / ((int64_t)ktaps * cpg));
      const int ci = (co / cpg) * cpg + i;
      const int sl = co / OUT;
      const int row = (tr ? ci : co) - sl * OUT, col = (tr ? co : ci) - sl * OUT, tap = tr ? ktaps - 1 - jt : jt;
      put_dt(J.dst, J.out_dtype, (((int64_t)sl * ktaps + tap) * 48 + row) * 64 + col, J.src[idx]);
    }
    return;
  }
  if (J.kind == 0 && J.out_dtype != NBASR_F32 && (J.n_out & 3) == 0) {      // fp32 -> 16-bit copy, 4 elements per thread
    const int64_t s0 = blk * (PB_CHUNK / 4), s1 = min(J.n_out >> 2, s0 + PB_CHUNK / 4);
    const float4* src = reinterpret_cast<const float4*>(J.src);
    uint2* dst = reinterpret_cast<uint2*>(J.dst);
    for (int64_t i = s0 + threadIdx.x; i < s1; i += blockDim.x) {
      const float4 f = src[i];
      uint2 o;
      if (J.out_dtype == NBASR_F16) { o.x = f2_to_f16x2(f.x, f.y); o.y = f2_to_f16x2(f.z, f.w); }
      else { o.x = f2_to_bf16x2(f.x, f.y); o.y = f2_to_bf16x2(f.z, f.w); }
      dst[i] = o;
    }
    return;
  }
  const int64_t start = blk * PB_CHUNK;
  const int64_t end = min(J.n_out, start + PB_CHUNK);
  for (int64_t idx = start + threadIdx.x; idx < end; idx += blockDim.x) {
    float v = 0.f;
    if (J.kind == 0) {
      v = J.src[idx];
    } else if (J.kind == 4) {           // LSTM W_hh for the cluster kernel: [cta 16][row = gate*32 + unit][k 512] (lstm_sm100.cu)
      const int H = J.a[0];
      int k = (int)(idx % 512);
      int r = (int)((idx / 512) % 128);
      int cj = (int)(idx / (512 * 128));
      int g = r >> 5, u = cj * 32 + (r & 31);
      if (u < H && k < H) v = J.src[((int64_t)g * H + u) * H + k];
    } else {                            // group-transposed, tap-flipped fp32 (SIMT input-gradient operand)
      const int cpg = J.a[1], ktaps = J.a[2];
      int jt = (int)(idx % ktaps), o = (int)((idx / ktaps) % cpg), ci = (int)(idx / (ktaps * cpg));
      int g = ci / cpg, i = ci % cpg;
      v = J.src[((int64_t)(g * cpg + o) * cpg + i) * ktaps + (ktaps - 1 - jt)];
    }
    put_dt(J.dst, J.out_dtype, idx, v);
  }
}

}  // namespace

extern "C" int nbasr_pack_batch(const nbasr_pack_job* jobs, int n, const int32_t* blockmap, int64_t blocks, void* stream) {
  if (n <= 0 || blocks <= 0) return 0;
  pack_batch_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(jobs, n, reinterpret_cast<const int2*>(blockmap));
  NBASR_CHECK_LAUNCH();
  return 0;
}
