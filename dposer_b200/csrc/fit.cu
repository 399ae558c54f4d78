// SMPLify body-fitting loss, fused forward + closed-form backward (lib/body_model/fitting_losses.py:59-103):
//   reprojection  conf^2 * sum_xy GMoF(f * X/Z + c - kp2d, sigma)      (perspective_projection :6-38 with R = I,
//                                                                       `translation` unused as in the reference :30)
//   angle prior   w_a^2 * sum_i exp(s_i * pose[idx_i])^2, idx = {52, 55, 9, 12}, s = {+1, -1, -1, -1}   (:50-56)
//   shape prior   w_s^2 * sum betas^2
// One warp per sample: lanes stride over the joints, the per-sample loss is a shuffle reduction, and the cotangents
// of joints / pose / betas are written in the same pass (the caller scales them by dL/dloss[b]).
#include "common.cuh"

namespace dpb {

__global__ void __launch_bounds__(256) fit_loss_kernel(
    const float* __restrict__ joints, const float* __restrict__ kp2d, const float* __restrict__ conf,
    const float* __restrict__ center, const float* __restrict__ pose, int pose_dim, const float* __restrict__ betas,
    int n_betas, int K, const float* __restrict__ focal_b, float focal, float sigma, float w_angle, float w_shape, float* __restrict__ loss,
    float* __restrict__ reproj, float* __restrict__ g_joints, float* __restrict__ g_pose,
    float* __restrict__ g_betas, int64_t B) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  const float cx = center[b * 2], cy = center[b * 2 + 1];
  if (focal_b) focal = focal_b[b];   // K[:,0,0] = K[:,1,1] = focal_length accepts a per-image tensor (fitting_losses.py:24-26)
  const float s2 = sigma * sigma;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float* X = joints + (b * K + k) * 3;
    const float x = X[0], y = X[1], z = X[2];
    const float iz = 1.0f / z;
    const float dx = focal * (x * iz) + cx - kp2d[(b * K + k) * 2];
    const float dy = focal * (y * iz) + cy - kp2d[(b * K + k) * 2 + 1];
    const float c2 = conf[b * K + k] * conf[b * K + k];
    const float qx = s2 + dx * dx, qy = s2 + dy * dy;
    const float r = c2 * (s2 * dx * dx / qx + s2 * dy * dy / qy);
    acc += r;
    if (reproj) reproj[b * K + k] = r;
    if (g_joints) {
      // d GMoF / d d = 2 s^4 d / (s^2 + d^2)^2
      const float gx = c2 * 2.0f * s2 * s2 * dx / (qx * qx), gy = c2 * 2.0f * s2 * s2 * dy / (qy * qy);
      float* G = g_joints + (b * K + k) * 3;
      G[0] = gx * focal * iz;
      G[1] = gy * focal * iz;
      G[2] = -(gx * x + gy * y) * focal * iz * iz;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  // angle prior: one lane per term
  float ang = 0.f;
  if (pose && g_pose)
    for (int i = lane; i < pose_dim; i += 32) g_pose[b * pose_dim + i] = 0.f;
  __syncwarp();
  if (pose && lane < 4) {
    const int idx = lane == 0 ? 52 : lane == 1 ? 55 : lane == 2 ? 9 : 12;
    const float sgn = lane == 0 ? 1.f : -1.f;
    if (idx < pose_dim) {
      const float e = expf(sgn * pose[b * pose_dim + idx]);
      ang = w_angle * w_angle * e * e;
      if (g_pose) g_pose[b * pose_dim + idx] = 2.0f * sgn * ang;
    }
  }
  float shp = 0.f;
  if (betas)
    for (int i = lane; i < n_betas; i += 32) {
      const float v = betas[b * n_betas + i];
      shp += v * v;
      if (g_betas) g_betas[b * n_betas + i] = 2.0f * w_shape * w_shape * v;
    }
  float extra = ang + w_shape * w_shape * shp;
  for (int o = 16; o > 0; o >>= 1) extra += __shfl_xor_sync(0xffffffffu, extra, o);
  if (lane == 0) loss[b] = acc + extra;
}

}  // namespace dpb

using namespace dpb;

extern "C" int dpb_fit_loss(const float* joints, const float* joints_2d, const float* conf, const float* center,
                            const float* body_pose, int pose_dim, const float* betas, int n_betas, int n_joints,
                            const float* focal_b, float focal, float sigma, float w_angle, float w_shape, float* loss,
                            float* reproj,
                            float* g_joints, float* g_pose, float* g_betas, int64_t B, void* stream) {
  DPB_REQUIRE(joints && joints_2d && conf && center && loss, "dpb_fit_loss: joints, joints_2d, conf, center, loss are required");
  DPB_REQUIRE(n_joints > 0 && pose_dim >= 0 && n_betas >= 0, "dpb_fit_loss: bad sizes");
  DPB_REQUIRE(!body_pose || pose_dim > 55, "dpb_fit_loss: the angle prior reads body_pose[:, 52] and [:, 55]");
  if (B <= 0) return DPB_OK;
  PtrDeviceGuard guard(joints);
  fit_loss_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      joints, joints_2d, conf, center, body_pose, pose_dim, betas, n_betas, n_joints, focal_b, focal, sigma, w_angle, w_shape,
      loss, reproj, g_joints, g_pose, g_betas, B);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}
