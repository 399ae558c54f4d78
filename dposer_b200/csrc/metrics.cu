// Metric reductions: APD (lib/utils/metric.py:8-37) and per-sample mean point error
// (Evaler.eval_bodys, lib/dataset/AMASS.py:286-296), as device reductions.
#include "common.cuh"

namespace dpb {

// One block per row i: threads stride over all columns j, block-reduce in a fixed order, one store per row.
// No atomics: row_sums[i - row0] is deterministic, and the caller adds the rows in double precision (a single fp32
// accumulator over B^2 pair distances loses digits for B well above 10^4).
__global__ void __launch_bounds__(256) apd_kernel(const float* __restrict__ joints, int64_t B, int nj, int64_t row0,
                                                  float* __restrict__ row_sums) {
  extern __shared__ float ji[];  // [nj*3] joints of row i
  const int64_t i = row0 + blockIdx.x;
  for (int k = threadIdx.x; k < nj * 3; k += 256) ji[k] = joints[i * nj * 3 + k];
  __syncthreads();
  float acc = 0.f;
  for (int64_t j = threadIdx.x; j < B; j += 256) {
    if (j == i) continue;
    const float* jj = joints + j * nj * 3;
    float d = 0.f;
    for (int k = 0; k < nj; ++k) {
      float dx = ji[k * 3] - jj[k * 3], dy = ji[k * 3 + 1] - jj[k * 3 + 1], dz = ji[k * 3 + 2] - jj[k * 3 + 2];
      d += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    acc += d / (float)nj;
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0)
    row_sums[blockIdx.x] = ((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]));
}

// one warp per sample
__global__ void __launch_bounds__(256) point_error_kernel(const float* __restrict__ a, const float* __restrict__ c,
                                                          int64_t B, int n_points, const int32_t* __restrict__ idx,
                                                          int n_idx, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  const int n = idx ? n_idx : n_points;
  float s = 0.f;
  for (int k = lane; k < n; k += 32) {
    const int p = idx ? idx[k] : k;
    const float* pa = a + ((size_t)b * n_points + p) * 3;
    const float* pc = c + ((size_t)b * n_points + p) * 3;
    float dx = pa[0] - pc[0], dy = pa[1] - pc[1], dz = pa[2] - pc[2];
    s += sqrtf(dx * dx + dy * dy + dz * dz);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[b] = s / (float)n * 1000.0f;
}

}  // namespace dpb

extern "C" int dpb_apd_partial(const float* joints, int64_t B, int n_joints, int64_t row0, int64_t nrows,
                               float* row_sums, void* stream) {
  if (!joints || !row_sums || B <= 0 || n_joints <= 0 || row0 < 0 || nrows < 0 || row0 + nrows > B)
    return dpb::fail(DPB_EINVAL, "dpb_apd_partial: bad argument");
  if (nrows == 0) return DPB_OK;
  dpb::PtrDeviceGuard guard(joints);
  dpb::apd_kernel<<<(unsigned)nrows, 256, n_joints * 3 * sizeof(float), (cudaStream_t)stream>>>(joints, B, n_joints,
                                                                                              row0, row_sums);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

extern "C" int dpb_mean_point_error(const float* a, const float* c, int64_t B, int n_points, const int32_t* idx,
                                    int n_idx, float* out, void* stream) {
  if (!a || !c || !out || B <= 0 || n_points <= 0 || (idx && n_idx <= 0))
    return dpb::fail(DPB_EINVAL, "dpb_mean_point_error: bad argument");
  dpb::PtrDeviceGuard guard(a);
  dpb::point_error_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(a, c, B, n_points, idx, n_idx,
                                                                                    out);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}
