// Optimiser tail of Trainer.step (trainer.py:221-225) over flat fp32 buffers:
//   grad += 0.01 * W / ||W||_F for every PadConvRelu weight (gradient of the norm regulariser),
//   coef = min(1, max_norm / (||grad||_2 + 1e-6))   (torch.nn.utils.clip_grad_norm_),
//   Adam(betas, eps) with bias correction (torch.optim.Adam, amsgrad=False, weight_decay=0).
// Step count, lr and every reduction result live in a small device `state` array, so the whole
// step replays inside a CUDA graph.  Every reduction is DETERMINISTIC (per-block partials summed in a fixed order, no
// float atomics): data-parallel replicas that receive the same all-reduced gradient must compute bit-identical clip
// coefficients and updates, or they drift apart.
#include "common.cuh"
#include "kernels.h"

namespace {

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  return v;
}

// Regularised segments are cut into chunks of SEG_CHUNK elements; block b finds its (segment, chunk) by walking
// the (<= 64) segment lengths -- big tensors get proportionally many blocks.
constexpr int SEG_CHUNK = 16384;

__device__ __forceinline__ bool seg_locate(const int64_t* __restrict__ len, int nseg, int64_t blk, int& s, int64_t& start) {
  for (s = 0; s < nseg; ++s) {
    int64_t nch = (len[s] + SEG_CHUNK - 1) / SEG_CHUNK;
    if (blk < nch) { start = blk * SEG_CHUNK; return true; }
    blk -= nch;
  }
  return false;
}

// (the segment bases are 16-byte aligned in the engine's flat buffer; the scalar loops handle any other caller and the tails)
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr int SUMSQ_BLOCKS = 592;

__global__ void seg_sumsq_kernel(const float* __restrict__ p, const int64_t* __restrict__ off, const int64_t* __restrict__ len,
                                 int nseg, float* __restrict__ part) {
  int s; int64_t start;
  if (!seg_locate(len, nseg, blockIdx.x, s, start)) return;
  const float* x = p + off[s] + start;
  const int n = (int)(min(len[s], start + SEG_CHUNK) - start);
  float acc = 0.f;
  int i0 = 0;
  if (al16(x)) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + threadIdx.x; i < n; i += blockDim.x) acc += x[i] * x[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// block s: state[4 + s] = sum of the chunk partials of segment s, in chunk order
__global__ void seg_reduce_kernel(const float* __restrict__ part, const int64_t* __restrict__ len, int nseg, float* __restrict__ state) {
  const int s = blockIdx.x;
  int64_t first = 0;
  for (int i = 0; i < s; ++i) first += (len[i] + SEG_CHUNK - 1) / SEG_CHUNK;
  const int nch = (int)((len[s] + SEG_CHUNK - 1) / SEG_CHUNK);
  float acc = 0.f;
  for (int i = threadIdx.x; i < nch; i += blockDim.x) acc += part[first + i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) state[4 + s] = acc;
}

// state[2] = sum of n per-block partials, fixed order
__global__ void total_reduce_kernel(const float* __restrict__ part, int n, float* __restrict__ state) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) state[2] = acc;
}

// g += k W for one chunk of a regularised segment
__global__ void reg_grad_kernel(const float* __restrict__ p, float* __restrict__ g, const int64_t* __restrict__ off,
                                const int64_t* __restrict__ len, int nseg, const float* __restrict__ state, float reg_coef) {
  int s; int64_t start;
  if (!seg_locate(len, nseg, blockIdx.x, s, start)) return;
  const float nrm = sqrtf(state[4 + s]);
  const float k = nrm > 0.f ? reg_coef / nrm : 0.f;
  const float* x = p + off[s] + start;
  float* gg = g + off[s] + start;
  const int n = (int)(min(len[s], start + SEG_CHUNK) - start);
  int i0 = 0;
  if (al16(x) && al16(gg)) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      float4 w = reinterpret_cast<float4*>(gg)[i];
      w.x += k * v.x; w.y += k * v.y; w.z += k * v.z; w.w += k * v.w;
      reinterpret_cast<float4*>(gg)[i] = w;
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + threadIdx.x; i < n; i += blockDim.x) gg[i] += k * x[i];
}

__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  float acc = 0.f;
  int64_t i0 = 0;
  if (al16(g)) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(g)[i];
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    i0 = n4 << 2;
  }
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += g[i] * g[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float max_norm, float b1, float b2, float eps, float* __restrict__ state) {
  const float step = state[0] + 1.f;
  const float lr = state[1];
  const float total = sqrtf(state[2]);
  const float coef = fminf(1.f, max_norm / (total + 1e-6f));
  const float bc1 = 1.f - powf(b1, step);
  const float bc2s = sqrtf(1.f - powf(b2, step));
  const float step_size = lr / bc1;
  const int64_t n4 = n >> 2;   // flat buffers are padded to multiples of 64 elements
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    float* gp = &g4.x; float* mp = &m4.x; float* vp = &v4.x; float* pp = &p4.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gi = gp[j] * coef;
      mp[j] = b1 * mp[j] + (1.f - b1) * gi;
      vp[j] = b2 * vp[j] + (1.f - b2) * gi * gi;
      pp[j] -= step_size * mp[j] / (sqrtf(vp[j]) / bc2s + eps);
    }
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    reinterpret_cast<float4*>(p)[i] = p4;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * coef;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) / bc2s + eps);
  }
}

__global__ void optim_finish_kernel(float* state, int nseg, float max_norm) {
  if (threadIdx.x == 0) {
    state[3] = fminf(1.f, max_norm / (sqrtf(state[2]) + 1e-6f));
    state[0] += 1.f;
  }
}

}  // namespace

extern "C" int nbasr_optim_step(float* param, float* grad, float* m, float* v, int64_t n, const int64_t* seg_off,
                                const int64_t* seg_len, int nseg, int64_t seg_chunks, float reg_coef, float max_norm, float beta1,
                                float beta2, float eps, float* state, void* stream) {
  cudaStream_t st = as_stream(stream);
  float* part_total = state + 8 + nseg;              // SUMSQ_BLOCKS partials of ||grad||^2
  float* part_seg = part_total + SUMSQ_BLOCKS;       // seg_chunks partials of the per-segment ||W||^2
  if (nseg > 0 && reg_coef != 0.f) {
    // seg_chunks = sum_s ceil(seg_len[s] / 16384), computed once by the host that owns the segment table
    seg_sumsq_kernel<<<(unsigned)seg_chunks, 256, 0, st>>>(param, seg_off, seg_len, nseg, part_seg);
    NBASR_CHECK_LAUNCH();
    seg_reduce_kernel<<<nseg, 256, 0, st>>>(part_seg, seg_len, nseg, state);
    NBASR_CHECK_LAUNCH();
    reg_grad_kernel<<<(unsigned)seg_chunks, 256, 0, st>>>(param, grad, seg_off, seg_len, nseg, state, reg_coef);
    NBASR_CHECK_LAUNCH();
  }
  sumsq_kernel<<<SUMSQ_BLOCKS, 256, 0, st>>>(grad, n, part_total);
  NBASR_CHECK_LAUNCH();
  total_reduce_kernel<<<1, 256, 0, st>>>(part_total, SUMSQ_BLOCKS, state);
  NBASR_CHECK_LAUNCH();
  adam_kernel<<<1184, 256, 0, st>>>(param, grad, m, v, n, max_norm, beta1, beta2, eps, state);
  NBASR_CHECK_LAUNCH();
  optim_finish_kernel<<<1, 32, 0, st>>>(state, nseg, max_norm);
  NBASR_CHECK_LAUNCH();
  return 0;
}
