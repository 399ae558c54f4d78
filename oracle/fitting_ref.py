"""Oracle: the small host-side pieces around the score net and LBS.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against the real reference
modules by tests/golden/make_golden.py.

Reference anchors:
  lib/dataset/AMASS.py:187-259      Posenormalizer (z-score branch: min_max=False, normalize=True)
  lib/utils/misc.py:27-55,84-95     create_mask, gaussian_smoothing
  lib/body_model/utils.py:39-47     BodyPartIndices
  lib/utils/metric.py:8-37          average_pairwise_distance
  lib/dataset/AMASS.py:275-316      Evaler (MPVPE / MPJPE, min over hypotheses)
  lib/body_model/fitting_losses.py  projection, GMoF, angle prior, body/camera losses
  lib/dataset/EvaSampler.py:77-106  contiguous shard rule
"""
import numpy as np
import torch
import torch.nn.functional as F

# joint ids EXCLUDING pelvis (lib/body_model/utils.py:36-47)
BODY_PARTS = {
    'left_leg': [0, 3, 6, 9], 'right_leg': [1, 4, 7, 10],
    'left_arm': [12, 15, 17, 19], 'right_arm': [13, 16, 18, 20],
    'trunk': [2, 5, 8, 15, 16], 'hands': [19, 20],
}
BODY_PARTS['legs'] = sorted(BODY_PARTS['left_leg'] + BODY_PARTS['right_leg'])
BODY_PARTS['arms'] = sorted(BODY_PARTS['left_arm'] + BODY_PARTS['right_arm'])


def normalize(poses, mean, std):
    """AMASS.py:218-229 z-score."""
    return (poses - mean.view(1, -1)) / std.view(1, -1)


def denormalize(poses, mean, std):
    """AMASS.py:247-256."""
    return poses * std.view(1, -1) + mean.view(1, -1)


def mask_indices(part, rot_n=3):
    """misc.py:33-36 -- flat pose dims zeroed in the mask for a body part."""
    j = torch.tensor(BODY_PARTS[part]).view(-1, 1) * rot_n + torch.arange(rot_n).view(1, -1)
    return j.flatten()


def create_mask(body_poses, part='legs', noise=None):
    """misc.py:27-41 (observation_type='noise'); the Gaussian fill is injected."""
    idx = mask_indices(part, body_poses.shape[1] // 21)
    mask = body_poses.new_ones(body_poses.shape)
    mask[:, idx] = 0
    obs = body_poses.clone()
    obs[:, idx] = torch.randn_like(obs[:, idx]) if noise is None else noise
    return mask, obs


def gaussian_smoothing(data, window_size, sigma):
    """misc.py:84-95 -- zero-padded conv1d along dim 0."""
    k = torch.arange(window_size).float() - window_size // 2
    k = torch.exp(-0.5 * (k / sigma) ** 2)
    k = (k / k.sum()).view(1, 1, -1)
    d = data.transpose(0, 1).unsqueeze(1)
    return F.conv1d(d, k, padding=window_size // 2).squeeze(1).transpose(0, 1)


def apd(joints3d):
    """metric.py:8-37 -- mean over ordered pairs (i != j) of mean-over-joints L2 distance."""
    B = joints3d.shape[0]
    diff = joints3d[:, None] - joints3d[None]                 # [B,B,J,3]
    d = diff.norm(dim=-1).mean(dim=-1)                        # [B,B]
    d.fill_diagonal_(0)
    return d.sum() / (B * (B - 1))


def eval_bodies(v_out, v_gt, j_out, j_gt, vert_idx=None, joint_idx=None):
    """AMASS.py:275-298 -- per-sample MPVPE / MPJPE in mm (numpy fp32 like the reference)."""
    v_out, v_gt, j_out, j_gt = [np.asarray(a) for a in (v_out, v_gt, j_out, j_gt)]
    vi = slice(None) if vert_idx is None else np.asarray(vert_idx)
    ji = slice(None) if joint_idx is None else np.asarray(joint_idx)
    mpvpe = np.sqrt(((v_out[:, vi] - v_gt[:, vi]) ** 2).sum(-1)).mean(-1) * 1000
    mpjpe = np.sqrt(((j_out[:, ji] - j_gt[:, ji]) ** 2).sum(-1)).mean(-1) * 1000
    return mpvpe, mpjpe


def shard_range(total, world, rank):
    """EvaSampler.py:77-106 -- contiguous chunk; the first ``total % world`` ranks get one extra."""
    base, mod = divmod(total, world)
    if rank <= mod:
        start = rank * (base + 1)
    else:
        start = mod * (base + 1) + (rank - mod) * base
    n = base + 1 if rank < mod else base
    return start, n


# ------------------------------------------------------------------ fitting losses
def perspective_projection(points, focal, center):
    """fitting_losses.py:6-38 with rotation = I (the translation argument is unused there, :30)."""
    p = points / points[:, :, -1].unsqueeze(-1)
    return torch.stack([focal * p[..., 0] + center[:, None, 0] * p[..., 2],
                        focal * p[..., 1] + center[:, None, 1] * p[..., 2]], dim=-1)


def gmof(x, sigma):
    """fitting_losses.py:41-47."""
    return (sigma ** 2 * x ** 2) / (sigma ** 2 + x ** 2)


def angle_prior(pose):
    """fitting_losses.py:50-56."""
    return torch.exp(pose[:, [52, 55, 9, 12]] * torch.tensor([1., -1., -1., -1.])) ** 2


def body_fitting_loss(body_pose, betas, joints, center, kp2d, conf, prior_scalar,
                      focal=5000., sigma=100., w_pose=4.78, w_shape=5., w_angle=15.2):
    """fitting_losses.py:59-103, output='mean'; ``prior_scalar`` is pose_prior(...) (already sum/B)."""
    proj = perspective_projection(joints, focal, center)
    reproj = (conf ** 2) * gmof(proj - kp2d, sigma).sum(-1)
    total = reproj.sum(-1) + (w_pose ** 2) * prior_scalar + (w_angle ** 2) * angle_prior(body_pose).sum(-1) \
        + (w_shape ** 2) * (betas ** 2).sum(-1)
    return total.mean()


def camera_fitting_loss(joints, cam_t, cam_t_est, center, kp2d, conf, focal=5000., w_depth=100.,
                        op_ind=(9, 12, 2, 5), gt_ind=(27, 28, 33, 34)):
    """fitting_losses.py:106-136; op/gt indices are constants.JOINT_IDS of the 4 torso joints."""
    proj = perspective_projection(joints, focal, center)
    op_ind, gt_ind = list(op_ind), list(gt_ind)
    e_op = (kp2d[:, op_ind] - proj[:, op_ind]) ** 2
    e_gt = (kp2d[:, gt_ind] - proj[:, gt_ind]) ** 2
    valid = (conf[:, op_ind].min(dim=-1)[0][:, None, None] > 0).float()
    reproj = (valid * e_op + (1 - valid) * e_gt).sum(dim=(1, 2))
    depth = (w_depth ** 2) * (cam_t[:, 2] - cam_t_est[:, 2]) ** 2
    return (reproj + depth).sum()
