"""Rotation-representation helpers with the reference's names (lib/utils/transforms.py:197-255).

Data-preparation utilities (not on the hot path): plain torch ops on whatever device the input lives on.  The
reference builds them on third-party ``torchgeometry`` (``angle_axis_to_rotation_matrix`` /
``rotation_matrix_to_angle_axis``), which is not installable offline -- PARITY UNPINNED: they are checked by
properties (round trips, orthonormality, agreement with the LBS oracle's Rodrigues), not against the reference.
``misc.Posenormalizer(rot_rep='rot6d')`` and ``misc.create_mask(observation_type='mean')`` call them at the data boundary;
the hot-path kernels take the shipped ``rot_rep='axis'`` (63-D) configuration only (a 126-D score net would need its own
first / last layer geometry)."""
import torch
import torch.nn.functional as F


def axis_angle_to_mat3x3(angle_axis):
    """[N,3] -> [N,3,3] (Rodrigues; first-order branch below 1e-6 rad like torchgeometry's Taylor branch)."""
    aa = angle_axis.reshape(-1, 3)
    theta = aa.norm(dim=1, keepdim=True)
    small = theta < 1e-6
    k = aa / theta.clamp_min(1e-12)
    kx, ky, kz = k[:, 0], k[:, 1], k[:, 2]
    z = torch.zeros_like(kx)
    K = torch.stack([z, -kz, ky, kz, z, -kx, -ky, kx, z], dim=1).reshape(-1, 3, 3)
    s, c = torch.sin(theta)[:, :, None], torch.cos(theta)[:, :, None]
    eye = torch.eye(3, dtype=aa.dtype, device=aa.device)[None]
    R = eye + s * K + (1 - c) * (K @ K)
    rx, ry, rz = aa[:, 0], aa[:, 1], aa[:, 2]
    one = torch.ones_like(rx)
    R1 = torch.stack([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1).reshape(-1, 3, 3)
    return torch.where(small[:, :, None], R1, R)


def axis_angle_to_rot6d(angle_axis):
    """[N,3] -> [N,6]: the first two columns of the rotation matrix, row-major (transforms.py:237-252)."""
    return axis_angle_to_mat3x3(angle_axis)[:, :3, :2].reshape(-1, 6)


def rot6d_to_mat3x3(rot6d):
    """[N,6] -> [N,3,3] by Gram-Schmidt (transforms.py:225-234)."""
    x = rot6d.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1, dim=1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def rot6d_to_axis_angle(rot6d):
    """[N,6] -> [N,3] (transforms.py:197-222); NaNs (degenerate input) become 0 like the reference."""
    R = rot6d_to_mat3x3(rot6d)
    # through the unit quaternion: stable for angles near 0 and pi
    t = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    qw = torch.sqrt((1 + t).clamp_min(0)) / 2
    qx = torch.sqrt((1 + R[:, 0, 0] - R[:, 1, 1] - R[:, 2, 2]).clamp_min(0)) / 2
    qy = torch.sqrt((1 - R[:, 0, 0] + R[:, 1, 1] - R[:, 2, 2]).clamp_min(0)) / 2
    qz = torch.sqrt((1 - R[:, 0, 0] - R[:, 1, 1] + R[:, 2, 2]).clamp_min(0)) / 2
    qx = torch.copysign(qx, R[:, 2, 1] - R[:, 1, 2])
    qy = torch.copysign(qy, R[:, 0, 2] - R[:, 2, 0])
    qz = torch.copysign(qz, R[:, 1, 0] - R[:, 0, 1])
    v = torch.stack([qx, qy, qz], dim=1)
    n = v.norm(dim=1)
    angle = 2 * torch.atan2(n, qw)
    aa = v * (angle / n.clamp_min(1e-12))[:, None]
    aa = torch.where((n < 1e-12)[:, None], 2 * v, aa)
    aa[torch.isnan(aa)] = 0.0
    return aa
