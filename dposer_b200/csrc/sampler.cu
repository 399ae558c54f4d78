// Elementwise pieces of the sampler / prior loss for the fp32 engine, Langevin corrector
// kernels, Philox fill.  The tcgen05 engine fuses the same update into its last-layer
// epilogue (score_tc.cu) using the identical formulas and the identical Philox addressing.
//
// Reference: EulerMaruyamaPredictor.update_fn sampling.py:182-188, imputation :413-422,
// LangevinCorrector.update_fn :282-302, prior loss run/completion.py:131-149.
#include "score.h"

namespace dpb {

// one thread per (row, quad of 4 columns): 16 quads per row
__global__ void em_update_kernel(float* __restrict__ x_io, const float* __restrict__ raw,
                                 const float* __restrict__ coef, const float* __restrict__ obs,
                                 const float* __restrict__ mask, const float* __restrict__ noise, int noise_k,
                                 uint64_t seed, uint32_t step, float* __restrict__ traj,
                                 float* __restrict__ x_mean_out, int64_t B, int impute) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 16) return;
  const int64_t row = i >> 4;
  const int quad = (int)(i & 15);
  const float a = coef[0], b = coef[1], c = coef[2], alpha = coef[3], sd = coef[4];
  float zp[4], zi[4] = {0, 0, 0, 0};
  const size_t plane = (size_t)B * D;
  if (noise) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = quad * 4 + j;
      size_t e = (size_t)row * D + col;
      bool ok = col < D;
      if (noise_k == 3) {  // planes: [draw after corrector | predictor draw | draw after predictor]
        zp[j] = ok ? noise[plane + e] : 0.f;
        zi[j] = ok ? noise[2 * plane + e] : 0.f;
      } else {
        zp[j] = ok ? noise[e] : 0.f;
      }
    }
  } else {
    normal4(seed, (uint64_t)row, step, 1, quad, zp);
    if (impute) normal4(seed, (uint64_t)row, step, 2, quad, zi);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int col = quad * 4 + j;
    if (col >= D) continue;
    size_t e = (size_t)row * D + col;
    // x already carries the imputation of the corrector slot (impute_kernel ran before the score eval)
    float x = x_io[e];
    float xm = a * x + b * raw[(size_t)row * DP + col];
    float xn = xm + c * zp[j];
    if (impute) {
      float m = mask[e];
      xn = xn * (1.0f - m) + (alpha * obs[e] + zi[j] * sd) * m;   // sampling.py:413-422 after the predictor
    }
    x_io[e] = xn;
    if (traj) traj[e] = xn;
    if (x_mean_out) x_mean_out[e] = xm;
  }
}

// imputation that precedes the score evaluation of a step (corrector slot, sampling.py:459)
__global__ void impute_kernel(float* __restrict__ x_io, const float* __restrict__ coef,
                              const float* __restrict__ obs, const float* __restrict__ mask,
                              const float* __restrict__ noise, uint64_t seed, uint32_t step, int64_t B) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 16) return;
  const int64_t row = i >> 4;
  const int quad = (int)(i & 15);
  const float alpha = coef[3], sd = coef[4];
  float z[4];
  if (noise) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = quad * 4 + j;
      z[j] = col < D ? noise[(size_t)row * D + col] : 0.f;
    }
  } else {
    normal4(seed, (uint64_t)row, step, 0, quad, z);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int col = quad * 4 + j;
    if (col >= D) continue;
    size_t e = (size_t)row * D + col;
    float m = mask[e];
    x_io[e] = x_io[e] * (1.0f - m) + (alpha * obs[e] + z[j] * sd) * m;
  }
}

__global__ void scale_out_kernel(const float* __restrict__ raw, const float* __restrict__ row_scale, float scale,
                                 float* __restrict__ out, int64_t B) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  int64_t row = i / D;
  int col = (int)(i % D);
  float s = row_scale ? row_scale[row] : scale;
  out[i] = raw[row * DP + col] * s;
}

// ------------------------------------------------------------------ Langevin
__global__ void __launch_bounds__(256) langevin_norms_kernel(const float* __restrict__ grad,
                                                             const float* __restrict__ noise,
                                                             float* __restrict__ sums, int64_t B) {
  // one warp per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * 8 + warp;
  float g = 0.f, z = 0.f;
  if (row < B) {
    for (int c = lane; c < D; c += 32) {
      float a = grad[row * D + c], b = noise[row * D + c];
      g = fmaf(a, a, g);
      z = fmaf(b, b, z);
    }
  }
  for (int s = 16; s > 0; s >>= 1) {
    g += __shfl_xor_sync(0xffffffffu, g, s);
    z += __shfl_xor_sync(0xffffffffu, z, s);
  }
  __shared__ float sg[8], sz[8];
  if (lane == 0) { sg[warp] = row < B ? sqrtf(g) : 0.f; sz[warp] = row < B ? sqrtf(z) : 0.f; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += sg[w]; b += sz[w]; }
    atomicAdd(&sums[0], a);
    atomicAdd(&sums[1], b);
  }
}

__global__ void langevin_update_kernel(float* __restrict__ x_io, float* __restrict__ x_mean,
                                       const float* __restrict__ grad, const float* __restrict__ noise,
                                       const float* __restrict__ sums, float snr, float alpha, int64_t B) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  // both means share the divisor, so the ratio of sums equals the ratio of means (sampling.py:296-298)
  float r = snr * sums[1] / sums[0];
  float step = r * r * 2.0f * alpha;
  float xm = x_io[i] + step * grad[i];
  if (x_mean) x_mean[i] = xm;
  x_io[i] = xm + sqrtf(step * 2.0f) * noise[i];
}

__global__ void normal_fill_kernel(float* __restrict__ out, int64_t B, uint64_t seed, uint32_t step, uint32_t slot) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 16) return;
  const int64_t row = i >> 4;
  const int quad = (int)(i & 15);
  float z[4];
  normal4(seed, (uint64_t)row, step, slot, quad, z);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int col = quad * 4 + j;
    if (col < D) out[(size_t)row * D + col] = z[j];
  }
}

// ------------------------------------------------------------------ prior loss (fp32 engine)
__global__ void perturb_kernel(const float* __restrict__ x0, const float* __restrict__ z, uint64_t seed,
                               uint32_t step, float alpha, float sd, float* __restrict__ xt, int64_t B) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 16) return;
  const int64_t row = i >> 4;
  const int quad = (int)(i & 15);
  float zz[4];
  if (z) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = quad * 4 + j;
      zz[j] = col < D ? z[(size_t)row * D + col] : 0.f;
    }
  } else {
    normal4(seed, (uint64_t)row, step, 3, quad, zz);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int col = quad * 4 + j;
    if (col < D) xt[(size_t)row * D + col] = alpha * x0[(size_t)row * D + col] + sd * zz[j];
  }
}

// one warp per row; loss_out accumulated with one atomic per CTA
__global__ void __launch_bounds__(256) prior_loss_kernel(const float* __restrict__ x0, const float* __restrict__ xt,
                                                         const float* __restrict__ raw, float alpha, float sd,
                                                         float inv_sigma_std, float w, float inv_div,
                                                         float* __restrict__ loss_out, float* __restrict__ grad_out,
                                                         float* __restrict__ row_loss, int64_t B) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * 8 + warp;
  float acc = 0.f;
  if (row < B) {
    for (int c = lane; c < D; c += 32) {
      size_t e = (size_t)row * D + c;
      float score = -raw[(size_t)row * DP + c] * inv_sigma_std;       // utils.py:162
      float x0h = (xt[e] + (sd * sd) * score) / alpha;                  // completion.py:107
      float d = x0[e] - x0h;
      acc = fmaf(w * d, d, acc);
      if (grad_out) grad_out[e] = 2.0f * w * d * inv_div;
    }
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ float sacc[8];
  if (lane == 0) {
    sacc[warp] = acc;
    if (row_loss && row < B) row_loss[row] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int k = 0; k < 8; ++k) a += sacc[k];
    atomicAdd(loss_out, a * inv_div);
  }
}

// ------------------------------------------------------------------ host-side launchers used by api.cu
static inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

int launch_em_update(float* x_io, const float* raw, const float* coef, const float* obs, const float* mask,
                     const float* noise, int noise_k, uint64_t seed, uint32_t step, float* traj, float* x_mean,
                     int64_t B, int impute, cudaStream_t st) {
  em_update_kernel<<<blocks_for(B * 16, 256), 256, 0, st>>>(x_io, raw, coef, obs, mask, noise, noise_k, seed, step,
                                                            traj, x_mean, B, impute);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

int launch_impute(float* x_io, const float* coef, const float* obs, const float* mask, const float* noise,
                  uint64_t seed, uint32_t step, int64_t B, cudaStream_t st) {
  impute_kernel<<<blocks_for(B * 16, 256), 256, 0, st>>>(x_io, coef, obs, mask, noise, seed, step, B);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

int launch_scale_out(const float* raw, const float* row_scale, float scale, float* out, int64_t B, cudaStream_t st) {
  scale_out_kernel<<<blocks_for(B * D, 256), 256, 0, st>>>(raw, row_scale, scale, out, B);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

int launch_perturb(const float* x0, const float* z, uint64_t seed, uint32_t step, float alpha, float sd, float* xt,
                   int64_t B, cudaStream_t st) {
  perturb_kernel<<<blocks_for(B * 16, 256), 256, 0, st>>>(x0, z, seed, step, alpha, sd, xt, B);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

int launch_prior_loss(const float* x0, const float* xt, const float* raw, float alpha, float sd, float inv_sigma_std,
                      float w, float inv_div, float* loss_out, float* grad_out, float* row_loss, int64_t B,
                      cudaStream_t st) {
  DPB_CUDA_CHECK(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  prior_loss_kernel<<<blocks_for(B, 8), 256, 0, st>>>(x0, xt, raw, alpha, sd, inv_sigma_std, w, inv_div, loss_out,
                                                      grad_out, row_loss, B);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

}  // namespace dpb

// ------------------------------------------------------------------ C ABI (stateless helpers)
extern "C" int dpb_langevin_norms(const float* grad, const float* noise, float* sums, int64_t B, void* stream) {
  if (!grad || !noise || !sums || B <= 0) return dpb::fail(DPB_EINVAL, "dpb_langevin_norms: bad argument");
  dpb::PtrDeviceGuard guard(grad);
  dpb::langevin_norms_kernel<<<dpb::blocks_for(B, 8), 256, 0, (cudaStream_t)stream>>>(grad, noise, sums, B);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

extern "C" int dpb_langevin_update(float* x_io, float* x_mean, const float* grad, const float* noise,
                                   const float* sums, float snr, float alpha, int64_t B, void* stream) {
  if (!x_io || !grad || !noise || !sums || B <= 0) return dpb::fail(DPB_EINVAL, "dpb_langevin_update: bad argument");
  dpb::PtrDeviceGuard guard(x_io);
  dpb::langevin_update_kernel<<<dpb::blocks_for(B * dpb::D, 256), 256, 0, (cudaStream_t)stream>>>(
      x_io, x_mean, grad, noise, sums, snr, alpha, B);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

extern "C" int dpb_normal_fill(float* out, int64_t B, uint64_t seed, uint64_t step, int slot, void* stream) {
  if (!out || B <= 0) return dpb::fail(DPB_EINVAL, "dpb_normal_fill: bad argument");
  dpb::PtrDeviceGuard guard(out);
  dpb::normal_fill_kernel<<<dpb::blocks_for(B * 16, 256), 256, 0, (cudaStream_t)stream>>>(out, B, seed,
                                                                                         (uint32_t)step, (uint32_t)slot);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}
