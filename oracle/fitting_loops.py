"""Oracle: CPU restatement of the three task loops (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

  completion_optimize  run/completion.py:167-207       (DPoserComp.optimize)
  motion_denoise       run/motion_denoising.py:199-300 (one or several independent sequences)
  smplify              run/smplify.py:168-281          (per-image normalisation, B independent images)

PINNED: tests/golden/make_golden_loops.py runs the REAL reference classes (run.completion.DPoserComp.optimize,
run.motion_denoising.MotionDenoise.optimize, run.smplify.SMPLify.__call__, imported from /root/reference with
the SURVEY Appendix C shims; body model = a BodyModel-compatible object over oracle/lbs_ref.py because smplx is
absent) with the Gaussian draws replayed, asserts these functions reproduce them, and stores the fixtures in
tests/golden/loops_golden.npz; tests/test_oracle_golden.py re-checks that on every CPU run.

torch autograd differentiates the oracle LBS (oracle/lbs_ref.py); the prior term detaches x0_hat exactly as the
reference does.  Every Gaussian draw is injected so the GPU path can be compared step for step.
"""
import math

import torch

from . import fitting_ref as Fr
from . import lbs_ref, score_ref


def _prior_term(sd, sde, x0, t, z, weighted, divisor):
    vt = torch.ones(x0.shape[0]) * t
    mean, std = sde.marginal(x0, vt)
    with torch.no_grad():
        x0_hat, snr = score_ref.one_step_denoise(sd, sde, (mean + std[:, None] * z).detach(), vt)
    w = 0.5 * torch.sqrt(1 + snr) if weighted else 0.5
    return (w * (x0 - x0_hat) ** 2).sum() / divisor


def _smplx_pose(body_pose, global_orient=None, hand=None):
    B = body_pose.shape[0]
    go = torch.zeros(B, 3) if global_orient is None else global_orient
    hd = torch.zeros(B, 90) if hand is None else hand
    return torch.cat([go, body_pose, torch.zeros(B, 9), hd], dim=1)


def completion_optimize(sd, observation, mask, z_list, sde_N=1000, lr=0.1, sample_trun=5.0, iterations=2,
                        steps_per_iter=100):
    """DPoserComp.optimize, time strategy '3' (run/completion.py:167-207).  The reference passes quan_t into the
    `weighted` slot (:196, SURVEY B-3): the loss is weighted whenever quan_t != 0.  mean over B*63 (:147)."""
    sde = score_ref.SubVP(0.1, 20., sde_N)
    ts = torch.linspace(1.0, 1e-3, sde_N)
    total = iterations * steps_per_iter
    x = observation.clone().requires_grad_(True)
    opt = torch.optim.Adam([x], lr, betas=(0.9, 0.999))
    for it in range(iterations):
        for i in range(steps_per_iter):
            step = it * steps_per_iter + i
            opt.zero_grad()
            quan_t = sde_N - math.floor(torch.tensor(total - step - 1) * (sde_N / (sample_trun * total))) - 2
            l_prior = _prior_term(sd, sde, x, ts[quan_t], z_list[step], bool(quan_t), float(x.numel()))
            l_data = torch.mean((x * mask - observation * mask) ** 2)
            tot = 100. * l_data / (1 + it) + 0.1 * l_prior * (it + 1)
            tot.backward()
            opt.step()
    return (observation * mask + x * (1.0 - mask)).detach()


def motion_denoise(sd, model, joints3d, init_pose, mean, std, z_list, seq_len, sde_N=500, iterations=1,
                   steps_per_iter=3, sample_trun=4.0, dposer_weight=1.0):
    """Returns the optimised [rows,63] body pose after iterations*steps_per_iter Adam steps (lr 0.03)."""
    sde = score_ref.SubVP(0.1, 20., sde_N)
    rows = init_pose.shape[0]
    n_seq = rows // seq_len
    pose = init_pose.clone().requires_grad_(True)
    opt = torch.optim.Adam([pose], 0.03, betas=(0.9, 0.999))
    ts = torch.linspace(1.0, 1e-3, sde_N)
    total = iterations * steps_per_iter
    betas = torch.zeros(rows, model['shapedirs'].shape[2])
    for it in range(iterations):
        for i in range(steps_per_iter):
            step = it * steps_per_iter + i
            opt.zero_grad()
            quan_t = sde_N - math.floor(torch.tensor(total - step - 1) * (sde_N / (sample_trun * total))) - 2
            x0 = Fr.normalize(pose, mean, std)
            l_prior = _prior_term(sd, sde, x0, ts[quan_t], z_list[step], False, seq_len)
            v, j = lbs_ref.body_forward(model, betas, _smplx_pose(pose))
            vs = v.view(n_seq, seq_len, -1, 3)
            d = vs[:, :-1] - vs[:, 1:]
            l_temp = torch.sqrt((d * d).sum(3)).mean(dim=(1, 2)).sum()
            e = (j[:, :22] - joints3d).view(n_seq, seq_len, 22, 3)
            l_data = torch.sqrt((e * e).sum(3)).mean(dim=(1, 2)).sum()
            tot = 10. * l_temp * (1 + it) + 100. * l_data / (1 + it * it) + 0.1 * l_prior * (1 + it) * dposer_weight
            tot.backward()
            opt.step()
    return pose.detach()


def smplify(sd, model, joint_map, init_pose, init_betas, init_cam_t, center, kp2d, mean, std, z_list, num_iters=2,
            sde_N=500, focal=5000., step_size=1e-2, ign_joints=(1, 9, 12, 27, 28), hand_mean=None):
    """Returns (pose[B,66], betas, cam_t) after num_iters camera steps + 5*num_iters body steps.
    hand_mean [90]: the constant hand pose of lib/body_model/smpl.py's SMPLX (smplx defaults: flat_hand_mean=False)."""
    hand = None if hand_mean is None else hand_mean[None].expand(init_pose.shape[0], -1)
    sde = score_ref.SubVP(0.1, 20., sde_N)
    ts = torch.linspace(1.0, 1e-3, sde_N)
    B = init_pose.shape[0]
    cam_t = init_cam_t.clone().requires_grad_(True)
    j2d, conf = kp2d[:, :, :2], kp2d[:, :, -1].clone()
    body = init_pose[:, 3:].clone()
    glob = init_pose[:, :3].clone().requires_grad_(True)
    betas = init_betas.clone()

    def joints(bt, bp, go, tr):
        shape = torch.cat([bt, torch.zeros(B, model['shapedirs'].shape[2] - bt.shape[1])], 1)
        _, j = lbs_ref.body_forward(model, shape, _smplx_pose(bp, go, hand), tr)
        return j[:, joint_map]
    opt = torch.optim.Adam([glob, cam_t], lr=step_size, betas=(0.9, 0.999))
    for _ in range(num_iters):
        loss = Fr.camera_fitting_loss(joints(betas, body, glob, cam_t), cam_t, init_cam_t, center, j2d, conf, focal)
        opt.zero_grad()
        loss.backward()
        opt.step()
    cam_t = cam_t.detach()
    body.requires_grad_(True)
    betas.requires_grad_(True)
    conf[:, list(ign_joints)] = 0.
    opt = torch.optim.Adam([body, betas, glob], lr=step_size, betas=(0.9, 0.999))
    weights = list(zip([50, 20, 10, 5, 2], [50, 20, 10, 5, 2], [150, 50, 30, 15, 5]))
    total = 5 * num_iters
    k = 0
    for stage, (wp, ws, wa) in enumerate(weights):
        for i in range(num_iters):
            it = stage * num_iters + i
            quan_t = sde_N - math.floor(torch.tensor(total - it - 1) * (sde_N / (20.0 * total))) - 5
            j = joints(betas, body, glob, cam_t)
            x0 = Fr.normalize(body[:, :63], mean, std)
            prior = _prior_term(sd, sde, x0, ts[quan_t], z_list[k], True, 1.0)       # B=1 per image: sum / 1
            proj = Fr.perspective_projection(j, focal, center)
            reproj = (conf ** 2) * Fr.gmof(proj - j2d, 100.).sum(-1)
            per_img = reproj.sum(-1) + (wa ** 2) * Fr.angle_prior(body).sum(-1) + (ws ** 2) * (betas ** 2).sum(-1)
            loss = per_img.sum() + (wp ** 2) * prior
            opt.zero_grad()
            loss.backward()
            opt.step()
            k += 1
    return torch.cat([glob, body], 1).detach(), betas.detach(), cam_t
