"""Fitting loops of the hot path with the reference's class surface:

* ``MotionDenoise``  (run/motion_denoising.py:63-300)  Adam over the body pose with DPoser prior +
  temporal vertex smoothness + joint data term; LBS forward/backward on the native kernels.
* ``SMPLify``        (run/smplify.py:118-281)          camera stage + 5 weighted body stages.

Both accept a *batch of independent problems* (many 60-frame sequences / many images): every loss term is
normalised per problem exactly as the reference does for its single problem and
the per-problem losses are summed, so each problem sees the gradients (and Adam updates) it would see alone.
"""
import math

import numpy as np
import torch

from . import utils as mutils
from .fitting_losses import body_fitting_loss, camera_fitting_loss
from .misc import gaussian_smoothing
from .prior import MotionPrior

# constants.JOINT_IDS of ['OP Neck', 'OP RHip', 'OP LHip', 'Right Hip', 'Left Hip'] (run/smplify.py:136-137)
IGN_JOINTS = [1, 9, 12, 27, 28]


class MotionDenoise(MotionPrior):
    def __init__(self, config, args, diffusion_model, body_model, sde, normalizer, sde_N=1000, dposer_weight=1.0,
                 out_path=None, debug=False, batch_size=1, seq_len=None):
        """batch_size = total rows (frames).  seq_len = frames per independent sequence (default: one sequence)."""
        sde.N = sde_N                                            # run/motion_denoising.py:91
        super().__init__(diffusion_model, sde, config.training.continuous, batch_size)
        self.args, self.debug, self.device = args, debug, args.device
        self.body_model = body_model
        self.dposer_weight = dposer_weight
        self.out_path = out_path
        self.seq_len = seq_len or batch_size
        assert batch_size % self.seq_len == 0
        self.n_seq = batch_size // self.seq_len
        self.betas = torch.zeros((batch_size, 10), device=self.device)
        self.poses = torch.randn((batch_size, 63), device=self.device) * 0.01
        self.Normalizer = normalizer
        # per-sequence normalisation of the prior: sum / seq_len  (reference: sum / batch_size with one sequence)
        self.batch_size_divisor = self.seq_len

    def DPoser_loss(self, x_0, vec_t, quan_t=None, weighted=False, multi_denoise=False, z=None):
        if multi_denoise:
            return self._ddim_loss(x_0, vec_t, weighted, self.batch_size_divisor, 10, z)
        return self._fused_loss(x_0, self._host_t(vec_t), bool(weighted), float(self.batch_size_divisor), z)

    def get_loss_weights(self):
        """run/motion_denoising.py:156-162."""
        return {'temp': lambda cst, it: 10. ** 1 * cst * (1 + it),
                'data': lambda cst, it: 10. ** 2 * cst / (1 + it * it),
                'dposer': lambda cst, it: 10. ** -1 * cst * (1 + it) * self.dposer_weight}

    @staticmethod
    def backward_step(loss_dict, weight_dict, it):
        return torch.stack([weight_dict[k](loss_dict[k], it) for k in loss_dict]).sum()

    def _temporal_term(self, v):
        """mean ||v[t]-v[t+1]|| over the (seq_len-1) x V pairs of each sequence, summed over sequences
        (run/motion_denoising.py:256-257; adjacent rows of DIFFERENT sequences are excluded, App. B-7)."""
        vs = v.view(self.n_seq, self.seq_len, -1, 3)
        d = vs[:, :-1] - vs[:, 1:]
        return torch.sqrt(torch.sum(d * d, dim=3)).mean(dim=(1, 2)).sum()

    def _data_term(self, joints, target):
        d = (joints[:, :22] - target).view(self.n_seq, self.seq_len, 22, 3)
        return torch.sqrt(torch.sum(d * d, dim=3)).mean(dim=(1, 2))           # [n_seq]

    def optimize(self, joints3d, gt_poses=None, time_strategy='1', sample_trun=2.0, sample_time=990, iterations=5,
                 steps_per_iter=50, verbose=False, vis=False):
        """run/motion_denoising.py:199-300 (visualisation dropped)."""
        bm = self.body_model
        with torch.no_grad():
            smpl_gt = bm(betas=self.betas, pose_body=gt_poses) if gt_poses is not None else None
        init_joints = joints3d.detach()
        init_MPJPE = None
        if smpl_gt is not None:
            e = joints3d - smpl_gt.Jtr[:, :22]
            init_MPJPE = torch.mean(torch.sqrt(torch.sum(e * e, dim=2)), dim=1) * 100.
        pose_body = self.poses.clone().detach().requires_grad_(True)
        optimizer = torch.optim.Adam([pose_body], 0.03, betas=(0.9, 0.999))
        weight_dict = self.get_loss_weights()
        timesteps = mutils.timestep_grid(self.sde, 1e-3)
        total_steps = iterations * steps_per_iter
        for it in range(iterations):
            for i in range(steps_per_iter):
                step = it * steps_per_iter + i
                optimizer.zero_grad()
                loss_dict = dict()
                poses = self.Normalizer.offline_normalize(pose_body, from_axis=True)
                if time_strategy == '1':
                    quan_t = int(torch.randint(self.sde.N, [1]))
                elif time_strategy == '2':
                    quan_t = int(sample_time)
                elif time_strategy == '3':
                    quan_t = self.sde.N - math.floor(
                        torch.tensor(total_steps - step - 1) * (self.sde.N / (sample_trun * total_steps))) - 2
                else:
                    raise NotImplementedError('unsupported time sampling strategy')
                loss_dict['dposer'] = self.DPoser_loss(poses, float(timesteps[quan_t]), quan_t)
                out = bm(betas=self.betas, pose_body=pose_body)
                loss_dict['temp'] = self._temporal_term(out.v)
                data = self._data_term(out.Jtr, init_joints)
                # reference: `if data_term > 0` drops NaN / zero data terms with a host sync (:262);
                # here the predicate stays on the device, per sequence
                loss_dict['data'] = torch.where(data > 0, data, torch.zeros_like(data)).sum()
                self.backward_step(loss_dict, weight_dict, it).backward()
                optimizer.step()
        with torch.no_grad():
            pose_final = pose_body.detach()
            ps = pose_final.view(self.n_seq, self.seq_len, -1)
            smooth = torch.stack([gaussian_smoothing(s, window_size=3, sigma=2) for s in ps])
            smooth[:, 0], smooth[:, -1] = ps[:, 0], ps[:, -1]                    # :283-285
            smooth = smooth.reshape(-1, ps.shape[-1])
            final = bm(betas=self.betas, pose_body=smooth)
            results = {'pose_body': smooth}
            if smpl_gt is not None:
                je = final.Jtr[:, :22] - smpl_gt.Jtr[:, :22]
                ve = final.v - smpl_gt.v
                results.update(init_MPJPE=init_MPJPE.cpu().numpy(),
                               MPJPE=(torch.mean(torch.sqrt(torch.sum(je * je, dim=2)), dim=1) * 100.).cpu().numpy(),
                               MPVPE=(torch.mean(torch.sqrt(torch.sum(ve * ve, dim=2)), dim=1) * 100.).cpu().numpy())
        return results


class SMPLify:
    """Single-stage SMPLify (run/smplify.py:118-281) on a batch of independent images."""

    def __init__(self, body_model, step_size=1e-2, batch_size=32, num_iters=100, focal_length=5000, args=None,
                 pose_prior=None, per_problem=True):
        self.smpl = body_model
        self.device = getattr(args, 'device', None)
        self.focal_length = focal_length
        self.step_size = step_size
        self.ign_joints = IGN_JOINTS
        self.num_iters = num_iters
        self.pose_prior = pose_prior                       # a dposer_b200.prior.DPoser
        self.sde_N = args.sde_N
        self.time_strategy = args.time_strategy
        self.sample_time = round(args.sde_N * 0.9)
        self.sample_trun = 20.0
        self.loss_weights = {'pose_prior_weight': [50, 20, 10, 5, 2],
                             'shape_prior_weight': [50, 20, 10, 5, 2],
                             'angle_prior_weight': [150, 50, 30, 15, 5]}
        self.stages = len(self.loss_weights['pose_prior_weight'])
        self.per_problem = per_problem
        if per_problem and pose_prior is not None:
            pose_prior.batch_size = 1                       # reference runs B=1: prior = sum / 1 per image

    def sample_discrete_time(self, iteration):
        total_steps = self.stages * self.num_iters
        if self.time_strategy == '1':
            return int(torch.randint(self.sde_N, [1]))
        if self.time_strategy == '2':
            return int(self.sample_time)
        if self.time_strategy == '3':
            return self.sde_N - math.floor(
                torch.tensor(total_steps - iteration - 1) * (self.sde_N / (self.sample_trun * total_steps))) - 5
        raise NotImplementedError

    def __call__(self, init_pose, init_betas, init_cam_t, camera_center, keypoints_2d):
        camera_translation = init_cam_t.clone()
        joints_2d = keypoints_2d[:, :, :2]
        joints_conf = keypoints_2d[:, :, -1]
        body_pose = init_pose[:, 3:].detach().clone()
        global_orient = init_pose[:, :3].detach().clone()
        betas = init_betas.detach().clone()
        # ---- stage 1: camera translation + global orientation (:208-221)
        global_orient.requires_grad = True
        camera_translation.requires_grad = True
        cam_opt = torch.optim.Adam([global_orient, camera_translation], lr=self.step_size, betas=(0.9, 0.999))
        for _ in range(self.num_iters):
            out = self.smpl(betas=betas, body_pose=body_pose, global_orient=global_orient, pose2rot=True,
                            transl=camera_translation)
            loss = camera_fitting_loss(out.joints, camera_translation, init_cam_t, camera_center, joints_2d,
                                       joints_conf, focal_length=self.focal_length)
            cam_opt.zero_grad()
            loss.backward()
            cam_opt.step()
        camera_translation = camera_translation.detach()
        # ---- stage 2: body pose, shape, orientation (:224-260)
        body_pose.requires_grad = True
        betas.requires_grad = True
        joints_conf[:, self.ign_joints] = 0.               # mutates the caller's keypoints view, as the reference
        body_opt = torch.optim.Adam([body_pose, betas, global_orient], lr=self.step_size, betas=(0.9, 0.999))
        stage_weights = [dict(zip(self.loss_weights.keys(), vals)) for vals in zip(*self.loss_weights.values())]
        for stage, w in enumerate(stage_weights):
            for i in range(self.num_iters):
                out = self.smpl(betas=betas, body_pose=body_pose, global_orient=global_orient, pose2rot=True,
                                transl=camera_translation)
                quan_t = self.sample_discrete_time(iteration=stage * self.num_iters + i)
                loss = body_fitting_loss(body_pose, betas, out.joints, camera_translation, camera_center, joints_2d,
                                         joints_conf, self.pose_prior, quan_t=quan_t, focal_length=self.focal_length,
                                         per_problem=self.per_problem, **w)
                body_opt.zero_grad()
                loss.backward()
                body_opt.step()
        with torch.no_grad():
            out = self.smpl(betas=betas, body_pose=body_pose, global_orient=global_orient, pose2rot=True,
                            transl=camera_translation)
            quan_t = self.sample_discrete_time(iteration=self.num_iters - 1)
            reprojection_loss = body_fitting_loss(body_pose, betas, out.joints, camera_translation, camera_center,
                                                  joints_2d, joints_conf, self.pose_prior, quan_t=quan_t,
                                                  focal_length=self.focal_length, output='reprojection')
        pose = torch.cat([global_orient, body_pose], dim=-1).detach()
        return pose, betas.detach(), camera_translation, reprojection_loss
