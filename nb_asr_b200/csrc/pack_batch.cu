// One-launch refresh of every derived weight operand (bf16 copies, transposes, tap-flipped and block-diagonal
// packs): a device-resident job table, each block handles one 4096-element chunk of one job.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int PB_CHUNK = 4096;

template <typename T>
__device__ __forceinline__ void put(void* dst, int64_t i, float v) { reinterpret_cast<T*>(dst)[i] = static_cast<T>(v); }

__global__ void __launch_bounds__(256) pack_batch_kernel(const nbasr_pack_job* __restrict__ jobs, int n,
                                                         const int2* __restrict__ blockmap) {
  // blockmap[b] = (job, chunk) -- precomputed by the host that built the table (no per-block search)
  const int2 bm = blockmap[blockIdx.x];
  const int j = bm.x;
  const int64_t blk = bm.y;
  if (j >= n) return;
  const nbasr_pack_job J = jobs[j];
  const int64_t start = blk * PB_CHUNK;
  const int64_t end = min(J.n_out, start + PB_CHUNK);
  for (int64_t idx = start + threadIdx.x; idx < end; idx += blockDim.x) {
    float v = 0.f;
    if (J.kind == 0) {
      v = J.src[idx];
    } else if (J.kind == 1) {           // out[n][q*M + m] = w[m*ws_m + n*ws_n + (t0 + q*tstep)*ws_t]
      const int M = J.a[0], nq = J.a[2], t0 = J.a[3], ts = J.a[4];
      int m = (int)(idx % M);
      int q = (int)((idx / M) % nq);
      int nn = (int)(idx / ((int64_t)M * nq));
      v = J.src[m * J.s[0] + nn * J.s[1] + (int64_t)(t0 + q * ts) * J.s[2]];
    } else if (J.kind == 2) {           // block-diagonal [slab][tap][48][64] (gconv_sm100.cu)
      const int Cc = J.a[0], cpg = J.a[1], ktaps = J.a[2], tr = J.a[3];
      const int OUT = cpg == 10 ? 40 : 48;
      int kk = (int)(idx % 64);
      int nn = (int)((idx / 64) % 48);
      int jt = (int)((idx / (64 * 48)) % ktaps);
      int s = (int)(idx / ((int64_t)64 * 48 * ktaps));
      int c0 = s * OUT, cn = c0 + nn, ck = c0 + kk;
      if (nn < OUT && cn < Cc && ck < Cc && kk < 48 && (cn / cpg) == (ck / cpg))
        v = tr ? J.src[((int64_t)ck * cpg + (cn % cpg)) * ktaps + (ktaps - 1 - jt)] : J.src[((int64_t)cn * cpg + (ck % cpg)) * ktaps + jt];
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
    if (J.out_dtype == NBASR_BF16) put<bf16>(J.dst, idx, v);
    else put<float>(J.dst, idx, v);
  }
}

}  // namespace

extern "C" int nbasr_pack_batch(const nbasr_pack_job* jobs, int n, const int32_t* blockmap, int64_t blocks, void* stream) {
  if (n <= 0 || blocks <= 0) return 0;
  pack_batch_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(jobs, n, reinterpret_cast<const int2*>(blockmap));
  NBASR_CHECK_LAUNCH();
  return 0;
}
