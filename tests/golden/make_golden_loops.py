#!/usr/bin/env python
"""Pin the three task LOOPS to the REAL reference (run in the build container only; needs /root/reference):

    python tests/golden/make_golden_loops.py

Runs, unmodified and on CPU,
  * run.completion.DPoserComp.optimize            (run/completion.py:167-207)
  * run.motion_denoising.MotionDenoise.optimize   (run/motion_denoising.py:199-300)
  * run.smplify.SMPLify.__call__                  (run/smplify.py:168-281)
for a handful of Adam steps with the reference's own RNG draws replayed (torch.manual_seed before the call, the
same draws regenerated afterwards), asserts that oracle/fitting_loops.py reproduces the results, and writes
tests/golden/loops_golden.npz.  The body model handed to the reference classes is a BodyModel / SMPLX-compatible
object over oracle/lbs_ref.py (third-party smplx is absent: SURVEY 8c), on the synthetic SMPL-X tensors of
dposer_b200.synthetic.make_body_tensors('smplx') -- so the loops (loss assembly, normalisation quirks B-3/B-6/B-13,
schedules, Adam) are pinned to the reference, the LBS inside them to the restatement.
The reference processes ONE problem per call (one 60-frame sequence / one image); the batched oracle is checked
against per-problem reference runs.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the Appendix-C shims and imports the reference modules)

from lib.algorithms.advanced import sde_lib  # noqa: E402
from run.completion import DPoserComp  # noqa: E402
from run.motion_denoising import MotionDenoise  # noqa: E402
import run.smplify as RS  # noqa: E402

from dposer_b200 import synthetic  # noqa: E402
from dposer_b200.body_model import JOINT_MAP_49  # noqa: E402
from oracle import fitting_loops, lbs_ref  # noqa: E402
from oracle import fitting_ref as Fr  # noqa: E402
from oracle import score_ref as S  # noqa: E402

REF = MG.REF
DATA = os.path.join(REF, 'data/AMASS/amass_processed')


def close(a, b, tol, what):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = float((a - b).norm() / b.norm().clamp_min(1e-30))
    assert err <= tol, f'{what}: oracle != reference, rel err {err:.3e} > {tol:.1e}'
    print(f'  ok {what}: rel err {err:.2e}')


class Struct:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class RefBodyModel:
    """lib/body_model/body_model.py BodyModel surface (model_type='smplx', flat hands) over the LBS restatement."""

    def __init__(self, m):
        self.m = m

    def __call__(self, betas=None, pose_body=None, **kw):
        B = pose_body.shape[0]
        S_ = self.m['shapedirs'].shape[2]
        shape = torch.cat([betas, torch.zeros(B, S_ - betas.shape[1])], 1)
        full = torch.cat([torch.zeros(B, 3), pose_body, torch.zeros(B, 9 + 90)], 1)
        v, j = lbs_ref.body_forward(self.m, shape, full)
        return Struct(v=v, f=self.m['faces'], betas=betas, Jtr=j, pose_body=pose_body)


class RefSMPLX:
    """lib/body_model/smpl.py SMPLX surface: smplx defaults => constant non-zero mean hand pose, 49-joint map."""

    def __init__(self, m):
        self.m = m
        self.joint_map = torch.tensor(JOINT_MAP_49)

    def __call__(self, betas=None, body_pose=None, global_orient=None, pose2rot=True, transl=None, **kw):
        B = body_pose.shape[0]
        S_ = self.m['shapedirs'].shape[2]
        shape = torch.cat([betas, torch.zeros(B, S_ - betas.shape[1])], 1)
        full = torch.cat([global_orient, body_pose, torch.zeros(B, 9), self.m['hands_mean'][None].expand(B, -1)], 1)
        v, j = lbs_ref.body_forward(self.m, shape, full, transl)
        return Struct(vertices=v, joints=j[:, self.joint_map], global_orient=global_orient, body_pose=body_pose,
                      betas=betas, full_pose=full)


def replay(seed, n, shape):
    torch.manual_seed(seed)
    return [torch.randn(*shape) for _ in range(n)]


def main():
    cfg = MG.get_config()
    cfg.device = torch.device('cpu')
    model = MG.build_reference_model(cfg)
    sd = S.make_state_dict(42)
    out = {}
    stats = torch.load(os.path.join(DATA, 'version1/train/axis_normalize2.pt'))
    mean, std = stats['mean_poses'], stats['std_poses']
    toy = torch.tensor(np.load(os.path.join(REF, 'examples/toy_data.npz'))['pose_samples'])

    # ------------------------------------------------------------------ completion (config 3A)
    B, iters, spi = 6, 2, 4
    sde = sde_lib.subVPSDE(beta_min=0.1, beta_max=20., N=1000)
    comp = DPoserComp(model, sde, True, batch_size=B)
    x_gt = Fr.normalize(toy[:B], mean, std)
    torch.manual_seed(3)
    mask, obs = MG.create_mask(x_gt, part='legs')
    torch.manual_seed(101)
    ref = comp.optimize(obs, mask, time_strategy='3', lr=0.1, sample_trun=5.0, iterations=iters, steps_per_iter=spi)
    z_list = replay(101, iters * spi, (B, 63))
    got = fitting_loops.completion_optimize(sd, obs, mask, z_list, sde_N=1000, lr=0.1, sample_trun=5.0,
                                            iterations=iters, steps_per_iter=spi)
    close(got - obs, ref.detach() - obs, 1e-5, 'DPoserComp.optimize update')
    out.update(comp_obs=obs.numpy(), comp_mask=mask.numpy(), comp_z=torch.stack(z_list).numpy(),
               comp_out=ref.detach().numpy(), comp_iters=np.array([iters, spi]))

    # ------------------------------------------------------------------ motion denoising (config 4)
    m = synthetic.make_body_tensors('smplx')
    bm = RefBodyModel(m)
    seq_len, n_seq, iters, spi = 8, 2, 2, 3
    ges = torch.tensor(np.load(os.path.join(REF, 'examples/Gestures_3_poses_batch005.npz'))['pose_body']).float()
    args = types.SimpleNamespace(device='cpu', dataset_folder=DATA, version='version1')
    finals, smooths, zs, inits, noisies, gts, res_all = [], [], [], [], [], [], []
    for s in range(n_seq):
        gt = ges[s * 60:s * 60 + seq_len].clone()
        g = torch.Generator().manual_seed(31 + s)
        with torch.no_grad():
            j_gt = bm(betas=torch.zeros(seq_len, 10), pose_body=gt).Jtr[:, :22]
        noisy = j_gt + 0.04 * torch.randn(seq_len, 22, 3, generator=g)
        torch.manual_seed(32 + s)
        md = MotionDenoise(cfg, args, model, bm, sde_N=500, dposer_weight=1.0, batch_size=seq_len)
        init = md.poses.clone()
        seen = []
        orig = bm.__call__

        def rec(betas=None, pose_body=None, _seen=seen, **kw):
            _seen.append(pose_body.detach().clone())
            return RefBodyModel.__call__(bm, betas=betas, pose_body=pose_body, **kw)
        md.body_model = rec
        torch.manual_seed(200 + s)
        res = md.optimize(noisy, gt_poses=gt, time_strategy='3', sample_trun=4.0, iterations=iters,
                          steps_per_iter=spi)
        zs.append(torch.stack(replay(200 + s, iters * spi, (seq_len, 63))))
        finals.append(md.poses.detach().clone())        # the Adam leaf IS md.poses (motion_denoising.py:202,217)
        smooths.append(seen[-1])                        # last body-model call = smoothed pose (:283-286)
        inits.append(init), noisies.append(noisy), gts.append(gt), res_all.append(res)
    init, noisy, gt = torch.cat(inits), torch.cat(noisies), torch.cat(gts)
    z_list = [torch.cat([z[k] for z in zs]) for k in range(iters * spi)]
    got = fitting_loops.motion_denoise(sd, m, noisy, init, mean, std, z_list, seq_len, sde_N=500, iterations=iters,
                                       steps_per_iter=spi, sample_trun=4.0)
    ref_final = torch.cat(finals)
    close(got - init, ref_final - init, 2e-4, 'MotionDenoise.optimize update (2 sequences, per-sequence runs)')
    ps = got.view(n_seq, seq_len, -1)
    sm = torch.stack([Fr.gaussian_smoothing(s_, 3, 2) for s_ in ps])
    sm[:, 0], sm[:, -1] = ps[:, 0], ps[:, -1]
    close(sm.reshape(-1, 63), torch.cat(smooths), 2e-4, 'MotionDenoise smoothed pose')
    out.update(md_init=init.numpy(), md_noisy=noisy.numpy(), md_gt=gt.numpy(), md_z=torch.stack(z_list).numpy(),
               md_final=ref_final.numpy(), md_smooth=torch.cat(smooths).numpy(),
               md_MPJPE=np.concatenate([r['MPJPE'] for r in res_all]),
               md_MPVPE=np.concatenate([r['MPVPE'] for r in res_all]),
               md_init_MPJPE=np.concatenate([r['init_MPJPE'] for r in res_all]),
               md_geom=np.array([seq_len, n_seq, iters, spi]))

    # ------------------------------------------------------------------ SMPLify (config 5)
    B, iters = 3, 2
    smpl = RefSMPLX(m)
    g = torch.Generator().manual_seed(41)
    mean_params = np.load(os.path.join(REF, 'lib/body_model/smpl_mean_params.npz'))
    from dposer_b200.body_model import rot6d_to_axis_angle
    mean_pose = rot6d_to_axis_angle(torch.tensor(mean_params['pose'], dtype=torch.float32)).reshape(-1)
    mean_shape = torch.tensor(mean_params['shape'], dtype=torch.float32)
    gt_body = toy[:B]
    gt_glob = torch.tensor([3.14159, 0., 0.]) + 0.2 * torch.randn(B, 3, generator=g)
    cam = torch.stack([0.2 * torch.randn(B, generator=g), 0.2 * torch.randn(B, generator=g),
                       20 + 20 * torch.rand(B, generator=g)], 1)
    betas_gt = torch.randn(B, 10, generator=g)
    with torch.no_grad():
        j = smpl(betas=betas_gt, body_pose=gt_body, global_orient=gt_glob, transl=cam).joints
    center = torch.full((B, 2), 512.)
    kp = Fr.perspective_projection(j, 5000., center) + 2.0 * torch.randn(B, 49, 2, generator=g)
    conf = 0.3 + 0.7 * torch.rand(B, 49, generator=g)
    conf[:, 25:] = 0.
    kp2d = torch.cat([kp, conf[..., None]], -1)
    init_pose = torch.cat([gt_glob + 0.1, mean_pose[3:66][None].repeat(B, 1)], 1)
    init_betas = mean_shape[None].repeat(B, 1)
    init_cam = cam + torch.tensor([0.1, -0.1, 2.0])
    RS.DPoser.load_model = lambda self, config, args: model          # no checkpoint offline: random-init weights
    RS.tqdm = lambda it, **kw: it
    sargs = types.SimpleNamespace(device='cpu', dataset_folder=DATA, version='version1', ckpt_path=None,
                                  config_path='configs.subvp.amass_scorefc_continuous.get_config', sde_N=500,
                                  time_strategy='3')
    poses, betas_o, cams, reprojs, zs = [], [], [], [], []
    for b in range(B):                                               # the reference runs B=1 (run/fitting.py:74)
        fit = RS.SMPLify(body_model=smpl, step_size=1e-2, batch_size=1, num_iters=iters, focal_length=5000.,
                         args=sargs)
        torch.manual_seed(300 + b)
        kp_b = kp2d[b:b + 1].clone()
        pose, bet, cam_t, reproj = fit(init_pose[b:b + 1], init_betas[b:b + 1], init_cam[b:b + 1], center[b:b + 1],
                                       kp_b)
        assert float(kp_b[0, 9, 2]) == 0.                            # B-13: ign_joints zeroed in the caller's tensor
        zs.append(torch.stack(replay(300 + b, 5 * iters + 1, (1, 63))))
        poses.append(pose), betas_o.append(bet), cams.append(cam_t.detach()), reprojs.append(reproj)
    z_list = [torch.cat([z[k] for z in zs]) for k in range(5 * iters + 1)]
    got_pose, got_betas, got_cam = fitting_loops.smplify(sd, m, smpl.joint_map, init_pose, init_betas, init_cam,
                                                         center, kp2d.clone(), mean, std, z_list, num_iters=iters,
                                                         sde_N=500, hand_mean=m['hands_mean'])
    ref_pose, ref_betas, ref_cam = torch.cat(poses), torch.cat(betas_o), torch.cat(cams)
    close(got_pose - init_pose, ref_pose - init_pose, 2e-4, 'SMPLify pose update (3 images, B=1 runs)')
    close(got_betas - init_betas, ref_betas - init_betas, 2e-4, 'SMPLify betas update')
    close(got_cam - init_cam, ref_cam - init_cam, 2e-4, 'SMPLify camera update')
    out.update(sf_init_pose=init_pose.numpy(), sf_init_betas=init_betas.numpy(), sf_init_cam=init_cam.numpy(),
               sf_center=center.numpy(), sf_kp2d=kp2d.numpy(), sf_z=torch.stack(z_list).numpy(),
               sf_pose=ref_pose.numpy(), sf_betas=ref_betas.numpy(), sf_cam=ref_cam.numpy(),
               sf_reproj=torch.cat(reprojs).numpy(), sf_iters=np.array([iters]))
    np.savez_compressed(os.path.join(HERE, 'loops_golden.npz'), **out)
    print('loops_golden.npz written')


if __name__ == '__main__':
    main()
