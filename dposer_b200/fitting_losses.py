"""SMPLify fitting losses with the reference's names (lib/body_model/fitting_losses.py:6-136).

On CUDA tensors ``body_fitting_loss`` runs as one fused kernel (``dpb_fit_loss``: projection, GMoF, angle and shape
priors, loss and cotangents in a single pass) wrapped in an autograd.Function, so autograd still reaches the LBS
kernel's backward (``dpb_lbs_backward``) and the prior kernel's closed-form gradient.  The torch-op versions of the
reference's helpers serve the camera loss (8 torso joints) and ``output='reprojection'`` on device tensors; host
tensors are refused like everywhere else in the package (no CPU fallback).  ``per_problem=True``
reproduces the reference's B=1 normalisation for a batch of independent images (SURVEY App. B-6,B-10)."""
import torch

from . import _lib as L

# constants.JOINT_IDS of the four torso joints used by the camera loss (lib/body_model/constants.py:89)
OP_TORSO = [9, 12, 2, 5]        # OP RHip, OP LHip, OP RShoulder, OP LShoulder
GT_TORSO = [27, 28, 33, 34]     # Right Hip, Left Hip, Right Shoulder, Left Shoulder


def perspective_projection(points, rotation, translation, focal_length, camera_center):
    """fitting_losses.py:6-38.  ``translation`` is unused by the reference (:30) and therefore here too."""
    batch_size = points.shape[0]
    K = torch.zeros([batch_size, 3, 3], device=points.device)
    K[:, 0, 0] = focal_length
    K[:, 1, 1] = focal_length
    K[:, 2, 2] = 1.
    K[:, :-1, -1] = camera_center
    points = torch.einsum('bij,bkj->bki', rotation, points)
    projected = points / points[:, :, -1].unsqueeze(-1)
    return torch.einsum('bij,bkj->bki', K, projected)[:, :, :-1]


def gmof(x, sigma):
    """Geman-McClure, fitting_losses.py:41-47."""
    x2, s2 = x ** 2, sigma ** 2
    return (s2 * x2) / (s2 + x2)


def angle_prior(pose):
    """fitting_losses.py:50-56."""
    return torch.exp(pose[:, [55 - 3, 58 - 3, 12 - 3, 15 - 3]] *
                     torch.tensor([1., -1., -1, -1.], device=pose.device)) ** 2


class _FitLossFn(torch.autograd.Function):
    """loss[b] = reprojection + angle prior + shape prior of sample b (fitting_losses.py:72-92 without the pose
    prior, which the caller adds).  The kernel returns the cotangents with the loss; backward only scales them."""

    @staticmethod
    def forward(ctx, joints, body_pose, betas, joints_2d, conf, center, focal, sigma, w_angle, w_shape):
        dev = joints.device
        B, K = joints.shape[0], joints.shape[1]
        f = lambda t: t.detach().to(torch.float32).contiguous()
        j, p, b_, kp, c, ce = f(joints), f(body_pose), f(betas), f(joints_2d), f(conf), f(center)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        gj, gp, gb = torch.empty_like(j), torch.empty_like(p), torch.empty_like(b_)
        # focal: python number, or a [B] tensor of per-image focal lengths (run/fitting.py passes batch['focal_length'])
        focal_b, focal_s = None, 0.0
        if torch.is_tensor(focal) and focal.numel() > 1:
            focal_b = focal.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
            if focal_b.numel() != B:
                raise ValueError(f'focal_length has {focal_b.numel()} entries for a batch of {B}')
        else:
            focal_s = float(focal)        # a 1-element tensor costs one device read here: pass a float in loops
        L.check(L.load().dpb_fit_loss(L.ptr(j), L.ptr(kp), L.ptr(c), L.ptr(ce), L.ptr(p), p.shape[1], L.ptr(b_),
                                      b_.shape[1], K, L.ptr(focal_b), focal_s, float(sigma), float(w_angle),
                                      float(w_shape),
                                      L.ptr(loss), None, L.ptr(gj), L.ptr(gp), L.ptr(gb), B, L.current_stream(dev)))
        ctx.save_for_backward(gj, gp, gb)
        return loss

    @staticmethod
    def backward(ctx, g):
        gj, gp, gb = ctx.saved_tensors
        return gj * g[:, None, None], gp * g[:, None], gb * g[:, None], None, None, None, None, None, None, None


def _fused_ok(model_joints, camera_center, output, verbose):
    return (model_joints.is_cuda and output != 'reprojection' and not verbose and torch.is_tensor(camera_center)
            and camera_center.dim() == 2)


def body_fitting_loss(body_pose, betas, model_joints, camera_t, camera_center, joints_2d, joints_conf, pose_prior,
                      quan_t, focal_length=5000, sigma=100, pose_prior_weight=4.78, shape_prior_weight=5,
                      angle_prior_weight=15.2, output='mean', verbose=False, per_problem=False):
    """fitting_losses.py:59-103.  pose_prior(body_pose, betas, quan_t) returns a scalar that is broadcast-added
    to the per-sample vector (:79,90).  With per_problem=True the result is the SUM of per-image losses (each
    image normalised as a batch of one), i.e. B independent reference problems solved at once."""
    batch_size = body_pose.shape[0]
    L.require_cuda(model_joints, 'model_joints')   # no CPU fallback: the joints come from the LBS kernels
    if _fused_ok(model_joints, camera_center, output, verbose):
        per_sample = _FitLossFn.apply(model_joints, body_pose, betas, joints_2d, joints_conf, camera_center,
                                      focal_length, sigma, angle_prior_weight, shape_prior_weight)
        prior = (pose_prior_weight ** 2) * pose_prior(body_pose, betas, quan_t) if pose_prior is not None else 0.0
        if per_problem:
            return per_sample.sum() + prior
        total = per_sample + prior
        return total.sum() if output == 'sum' else total.mean()
    rotation = torch.eye(3, device=body_pose.device).unsqueeze(0).expand(batch_size, -1, -1)
    projected = perspective_projection(model_joints, rotation, camera_t, focal_length, camera_center)
    reprojection_loss = (joints_conf ** 2) * gmof(projected - joints_2d, sigma).sum(dim=-1)
    pose_prior_loss = (pose_prior_weight ** 2) * pose_prior(body_pose, betas, quan_t) if pose_prior is not None else 0.0
    angle_prior_loss = (angle_prior_weight ** 2) * angle_prior(body_pose).sum(dim=-1)
    shape_prior_loss = (shape_prior_weight ** 2) * (betas ** 2).sum(dim=-1)
    if per_problem:
        return (reprojection_loss.sum(dim=-1) + angle_prior_loss + shape_prior_loss).sum() + pose_prior_loss
    total = reprojection_loss.sum(dim=-1) + pose_prior_loss + angle_prior_loss + shape_prior_loss
    if output == 'sum':
        return total.sum()
    if output == 'reprojection':
        return reprojection_loss
    return total.mean()


def camera_fitting_loss(model_joints, camera_t, camera_t_est, camera_center, joints_2d, joints_conf,
                        focal_length=5000, depth_loss_weight=100):
    """fitting_losses.py:106-136."""
    batch_size = model_joints.shape[0]
    L.require_cuda(model_joints, 'model_joints')
    rotation = torch.eye(3, device=model_joints.device).unsqueeze(0).expand(batch_size, -1, -1)
    projected = perspective_projection(model_joints, rotation, camera_t, focal_length, camera_center)
    err_op = (joints_2d[:, OP_TORSO] - projected[:, OP_TORSO]) ** 2
    err_gt = (joints_2d[:, GT_TORSO] - projected[:, GT_TORSO]) ** 2
    is_valid = (joints_conf[:, OP_TORSO].min(dim=-1)[0][:, None, None] > 0).float()
    reprojection_loss = (is_valid * err_op + (1 - is_valid) * err_gt).sum(dim=(1, 2))
    depth_loss = (depth_loss_weight ** 2) * (camera_t[:, 2] - camera_t_est[:, 2]) ** 2
    return (reprojection_loss + depth_loss).sum()
