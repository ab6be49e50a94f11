// Log-mel front end of the reference data pipeline on the GPU (training/torch/timit.py:90-106):
//   torchaudio MelSpectrogram(sample_rate=16000, n_fft=win_length=400, hop_length=160, n_mels=80; periodic Hann window,
//   center=True with reflect padding of 200 samples, power 2, HTK mel scale, no filter normalisation) -> log ->
//   (x - mean) / (variance + eps)  [sic, no sqrt: timit.py:83] -> zero padding to the longest utterance (timit.py:104).
// The windowed 400-point DFT and the mel projection are two fp32 GEMMs on the library's own GEMM entry point
// (nbasr_gemm_tn with overlapping rows: frame r is samples [160 r, 160 r + 400) of the reflect-padded waveform, the same
// trick as the k=8 convolutions); the three kernels here are the glue: reflect padding, power, log + normalise + layout.
#include "common.cuh"
#include "kernels.h"

namespace {

// wav (B, L) fp32 zero padded, len[b] samples -> padded (B, Lp): [200 reflected | len | 200 reflected | zeros]
__global__ void fe_reflect_pad_kernel(const float* __restrict__ wav, const int64_t* __restrict__ len, int B, int64_t L, int64_t Lp,
                                      int pad, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t n = len[b];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Lp; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t j = i - pad;                       // position in the unpadded signal
    float v = 0.f;
    if (j < n + pad && n > 0) {
      if (j < 0) j = -j;                       // reflect (no edge repeat), as torch.nn.functional.pad(mode='reflect')
      if (j >= n) j = 2 * (n - 1) - j;
      if (j >= 0 && j < n) v = wav[b * L + j];
    }
    out[b * Lp + i] = v;
  }
}

// spec (rows, ld_s): [re(0..nf) | im(0..nf)] -> pw (rows, ld_p): re^2 + im^2, zero beyond nf
__global__ void fe_power_kernel(const float* __restrict__ spec, int64_t rows, int nf, int ld_s, int ld_p, float* __restrict__ pw) {
  const int64_t total = rows * ld_p;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_p;
    const int n = (int)(i - r * ld_p);
    float v = 0.f;
    if (n < nf) {
      const float re = spec[r * ld_s + n], im = spec[r * ld_s + nf + n];
      v = re * re + im * im;
    }
    pw[i] = v;
  }
}

// mel (B*F, ld_m) -> out (B, n_mels, T): (log(mel) - mean) / (var + eps) for t < frames[b], 0 beyond (collate padding)
__global__ void fe_lognorm_kernel(const float* __restrict__ mel, const int64_t* __restrict__ len, int hop, int B, int F, int n_mels,
                                  int ld_m, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                  float* __restrict__ out, int T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int64_t nfr = len[b] > 0 ? 1 + len[b] / hop : 0;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, m = m0 + threadIdx.x;
    float v = 0.f;
    if (t < F && m < n_mels && t < nfr) v = (logf(mel[((int64_t)b * F + t) * ld_m + m]) - mean[m]) / (var[m] + eps);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int m = m0 + i, t = t0 + threadIdx.x;
    if (m < n_mels && t < T) out[((int64_t)b * n_mels + m) * T + t] = t < F ? tile[threadIdx.x][i] : 0.f;
  }
}

}  // namespace

extern "C" int nbasr_logmel(const float* wav, const int64_t* len, int B, int64_t L, const float* dft, const float* melfb,
                            const float* mean, const float* var, float eps, float* out, int T, float* work, int64_t work_floats,
                            void* stream) {
  constexpr int NFFT = 400, HOP = 160, NF = NFFT / 2 + 1, NMEL = 80, PAD = NFFT / 2;
  constexpr int LD_S = 408, LD_P = 208;       // 2*NF = 402 and NF = 201 rounded up to multiples of 8
  if (B <= 0 || L <= 0) return 0;
  const int F = (int)(1 + L / HOP);           // frames of the longest possible utterance
  NBASR_REQUIRE(T >= F, "output frame count");
  const int64_t Lp = ((L + 2 * PAD + NFFT + 7) / 8) * 8;
  const int64_t need = (int64_t)B * Lp + (int64_t)B * F * (LD_S + LD_P + NMEL);
  NBASR_REQUIRE(work_floats >= need, "nbasr_logmel workspace too small (see nbasr_logmel_work_floats)");
  cudaStream_t st = as_stream(stream);
  float* padded = work;
  float* spec = padded + (int64_t)B * Lp;
  float* pw = spec + (int64_t)B * F * LD_S;
  float* mel = pw + (int64_t)B * F * LD_P;
  fe_reflect_pad_kernel<<<dim3((unsigned)std::min<int64_t>((Lp + 255) / 256, 1024), B), 256, 0, st>>>(wav, len, B, L, Lp, PAD, padded);
  NBASR_CHECK_LAUNCH();
  nbasr_gemm g{};
  g.dtype = NBASR_F32; g.a = padded; g.a_bs = Lp; g.a_rs = HOP; g.nb = B; g.nr = F; g.K = NFFT; g.N = 2 * NF;
  g.w = dft; g.ldw = NFFT; g.o_r0 = 0; g.o_bs = F; g.o_rs = 1;
  g.epi.out = spec; g.epi.out_dtype = NBASR_F32; g.epi.ld_out = LD_S; g.epi.mask_w = 32; g.epi.mask2_w = 32;
  if (nbasr_gemm_tn(&g, stream)) return 1;
  const int64_t rows = (int64_t)B * F;
  fe_power_kernel<<<(unsigned)std::min<int64_t>((rows * LD_P + 255) / 256, 148 * 16), 256, 0, st>>>(spec, rows, NF, LD_S, LD_P, pw);
  NBASR_CHECK_LAUNCH();
  nbasr_gemm m{};
  m.dtype = NBASR_F32; m.a = pw; m.a_bs = (int64_t)F * LD_P; m.a_rs = LD_P; m.nb = B; m.nr = F; m.K = LD_P; m.N = NMEL;
  m.w = melfb; m.ldw = LD_P; m.o_r0 = 0; m.o_bs = F; m.o_rs = 1;
  m.epi.out = mel; m.epi.out_dtype = NBASR_F32; m.epi.ld_out = NMEL; m.epi.mask_w = 32; m.epi.mask2_w = 32;
  if (nbasr_gemm_tn(&m, stream)) return 1;
  fe_lognorm_kernel<<<dim3((T + 31) / 32, (NMEL + 31) / 32, B), dim3(32, 8), 0, st>>>(mel, len, HOP, B, F, NMEL, NMEL, mean, var, eps,
                                                                                      out, T);
  NBASR_CHECK_LAUNCH();
  return 0;
}

extern "C" int64_t nbasr_logmel_work_floats(int B, int64_t L) {
  const int F = (int)(1 + L / 160);
  const int64_t Lp = ((L + 400 + 400 + 7) / 8) * 8;
  return (int64_t)B * Lp + (int64_t)B * F * (408 + 208 + 80);
}
