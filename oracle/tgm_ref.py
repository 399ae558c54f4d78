"""TEST INFRASTRUCTURE (oracle): CPU restatement of the two third-party functions the reference's rot6d helpers call.

The reference's `lib/utils/transforms.py:197-255` builds `rot6d_to_axis_angle` / `axis_angle_to_rot6d` /
`axis_angle_to_mat3x3` on `torchgeometry` (reference `requirements.txt:13`, version not pinned there; the published
release is 0.1.2), which is neither vendored under /root/reference nor installable offline.  This module restates the
PUBLISHED algorithms of that release from their documented definitions (the same code is widely reproduced, e.g. in
SPIN's `utils/geometry.py`), so that `dposer_b200/transforms.py` is checked against the dependency's own arithmetic
(including its quirks) rather than only by properties:

* `angle_axis_to_rotation_matrix`  -- Rodrigues with the axis taken as `angle_axis / (theta + 1e-6)` (the epsilon is
  INSIDE the normalisation, so the axis is very slightly shorter than unit: relative 1e-6 / theta) and a first-order
  Taylor matrix where `theta^2 <= 1e-6`; returns 4x4 homogeneous matrices.
* `rotation_matrix_to_angle_axis`  -- `rotation_matrix_to_quaternion` (four-branch selection on the TRANSPOSED matrix,
  normalised by `0.5 / sqrt(t_k)`) followed by `quaternion_to_angle_axis` (`2 atan2(+-sin, +-cos)` so that the angle
  lies in (-pi, pi]).

PARITY UNPINNED for this file: the real torchgeometry cannot be imported here, and the reference holds no golden vector
for these calls.  Only tests/ may import this module."""
import torch


def angle_axis_to_rotation_matrix(angle_axis):
    """[N,3] -> [N,4,4] (torchgeometry 0.1.2 `angle_axis_to_rotation_matrix`; call site transforms.py:249,255)."""
    aa = angle_axis.reshape(-1, 3)
    theta2 = (aa * aa).sum(dim=1, keepdim=True)
    theta = torch.sqrt(theta2)
    w = aa / (theta + 1e-6)
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    c, s = torch.cos(theta), torch.sin(theta)
    normal = torch.cat([c + wx * wx * (1 - c), wx * wy * (1 - c) - wz * s, wy * s + wx * wz * (1 - c),
                        wz * s + wx * wy * (1 - c), c + wy * wy * (1 - c), -wx * s + wy * wz * (1 - c),
                        -wy * s + wx * wz * (1 - c), wx * s + wy * wz * (1 - c), c + wz * wz * (1 - c)], dim=1)
    rx, ry, rz = aa[:, 0:1], aa[:, 1:2], aa[:, 2:3]
    one = torch.ones_like(rx)
    taylor = torch.cat([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1)
    big = (theta2 > 1e-6).to(aa.dtype)
    R = torch.eye(4, dtype=aa.dtype).repeat(aa.shape[0], 1, 1)
    R[:, :3, :3] = (big * normal + (1 - big) * taylor).reshape(-1, 3, 3)
    return R


def rotation_matrix_to_quaternion(rotation_matrix, eps=1e-6):
    """[N,3,4] -> [N,4] (w, x, y, z) (torchgeometry 0.1.2 `rotation_matrix_to_quaternion`)."""
    m = rotation_matrix[:, :3, :3].transpose(1, 2)            # the published code indexes the transposed matrix
    d2 = m[:, 2, 2] < eps
    d0_d1 = m[:, 0, 0] > m[:, 1, 1]
    d0_nd1 = m[:, 0, 0] < -m[:, 1, 1]
    t0 = 1 + m[:, 0, 0] - m[:, 1, 1] - m[:, 2, 2]
    q0 = torch.stack([m[:, 1, 2] - m[:, 2, 1], t0, m[:, 0, 1] + m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2]], -1)
    t1 = 1 - m[:, 0, 0] + m[:, 1, 1] - m[:, 2, 2]
    q1 = torch.stack([m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] + m[:, 1, 0], t1, m[:, 1, 2] + m[:, 2, 1]], -1)
    t2 = 1 - m[:, 0, 0] - m[:, 1, 1] + m[:, 2, 2]
    q2 = torch.stack([m[:, 0, 1] - m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2], m[:, 1, 2] + m[:, 2, 1], t2], -1)
    t3 = 1 + m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    q3 = torch.stack([t3, m[:, 1, 2] - m[:, 2, 1], m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] - m[:, 1, 0]], -1)
    c0 = (d2 & d0_d1).to(m.dtype)[:, None]
    c1 = (d2 & ~d0_d1).to(m.dtype)[:, None]
    c2 = (~d2 & d0_nd1).to(m.dtype)[:, None]
    c3 = (~d2 & ~d0_nd1).to(m.dtype)[:, None]
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    q = q / torch.sqrt(t0[:, None] * c0 + t1[:, None] * c1 + t2[:, None] * c2 + t3[:, None] * c3)
    return q * 0.5


def quaternion_to_angle_axis(q):
    """[N,4] (w, x, y, z) -> [N,3] (torchgeometry 0.1.2 `quaternion_to_angle_axis`)."""
    q1, q2, q3 = q[:, 1], q[:, 2], q[:, 3]
    sin2 = q1 * q1 + q2 * q2 + q3 * q3
    sin_t = torch.sqrt(sin2)
    cos_t = q[:, 0]
    two_theta = 2.0 * torch.where(cos_t < 0.0, torch.atan2(-sin_t, -cos_t), torch.atan2(sin_t, cos_t))
    k = torch.where(sin2 > 0.0, two_theta / sin_t, 2.0 * torch.ones_like(sin_t))
    return torch.stack([q1 * k, q2 * k, q3 * k], dim=1)


def rotation_matrix_to_angle_axis(rotation_matrix):
    """[N,3,4] -> [N,3] (torchgeometry 0.1.2; call site transforms.py:220)."""
    return quaternion_to_angle_axis(rotation_matrix_to_quaternion(rotation_matrix))


# ---- the reference's three helpers on top of them (lib/utils/transforms.py:197-222, 237-258)
def rot6d_to_axis_angle(rot6d):
    import torch.nn.functional as F
    x = rot6d.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - torch.einsum('bi,bi->b', b1, a2).unsqueeze(-1) * b1)
    b3 = torch.cross(b1, b2, dim=1)
    R = torch.stack((b1, b2, b3), dim=-1)
    R = torch.cat([R, torch.zeros(R.shape[0], 3, 1, dtype=R.dtype)], 2)
    aa = rotation_matrix_to_angle_axis(R).reshape(-1, 3)
    aa[torch.isnan(aa)] = 0.0
    return aa


def axis_angle_to_rot6d(angle_axis):
    return angle_axis_to_rotation_matrix(angle_axis)[:, :3, :2].reshape(-1, 6)


def axis_angle_to_mat3x3(angle_axis):
    return angle_axis_to_rotation_matrix(angle_axis)[:, :3, :3]
