"""GPU parity: the two fitting loops (configs 4 and 5) against the oracle loops, a few Adam steps with the
Gaussian draws injected; LBS forward+backward, prior loss+grad and the fitting losses all take part."""
import types

import pytest
import torch

from conftest import golden, rel_err
from dposer_b200 import _lib as L
from dposer_b200 import fitting, prior, sde_lib, synthetic
from dposer_b200.body_model import BodyModel, SMPLX
from dposer_b200.misc import Posenormalizer
from oracle import fitting_loops, lbs_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('engine,tol', [(L.ENGINE_FP32, 2e-3), (L.ENGINE_TC, 5e-3)])
def test_motion_denoise_steps_vs_oracle(gpu_model, oracle_sd, engine, tol):
    m = synthetic.make_body_tensors('smplx')
    seq_len, n_seq, steps = 6, 2, 3
    rows = seq_len * n_seq
    g = torch.Generator().manual_seed(31)
    gt = synthetic.toy_poses()[:rows] + 0.02 * torch.randn(rows, 63, generator=g)
    _, j_gt = lbs_ref.body_forward(m, torch.zeros(rows, 20), torch.cat([torch.zeros(rows, 3), gt, torch.zeros(rows, 99)], 1))
    noisy = j_gt[:, :22] + 0.04 * torch.randn(rows, 22, 3, generator=g)
    init = 0.01 * torch.randn(rows, 63, generator=g)
    z_list = [torch.randn(rows, 63, generator=g) for _ in range(steps)]
    norm = Posenormalizer(None, device='cuda', normalize=True, min_max=False, rot_rep='axis')
    ref = fitting_loops.motion_denoise(oracle_sd, m, noisy, init, norm.mean_poses.cpu(), norm.std_poses.cpu(), z_list,
                                       seq_len, sde_N=500, iterations=1, steps_per_iter=steps, sample_trun=4.0)
    cfg = synthetic.default_config()
    args = types.SimpleNamespace(device='cuda')
    bm = BodyModel(m, num_betas=10, batch_size=rows, model_type='smplx').cuda()
    gpu_model.engine = engine
    try:
        md = fitting.MotionDenoise(cfg, args, gpu_model, bm, sde_lib.subVPSDE(0.1, 20., 1000), norm, sde_N=500,
                                   batch_size=rows, seq_len=seq_len)
        md.poses = init.cuda()
        # smoothing off for the comparison: take the raw optimised pose through a 1-frame window trick
        res = md.optimize(noisy.cuda(), gt_poses=gt.cuda(), time_strategy='3', sample_trun=4.0, iterations=1,
                          steps_per_iter=steps, z_list=z_list)
    finally:
        gpu_model.engine = L.ENGINE_AUTO
    from oracle import fitting_ref as Fr
    ps = ref.view(n_seq, seq_len, -1)
    sm = torch.stack([Fr.gaussian_smoothing(s, 3, 2) for s in ps])
    sm[:, 0], sm[:, -1] = ps[:, 0], ps[:, -1]
    ref_sm = sm.reshape(rows, -1)
    # Adam's first steps move every coordinate by ~lr regardless of gradient scale: compare the update itself
    upd_ref, upd = ref_sm - init, res['pose_body'].cpu() - init
    assert rel_err(upd, upd_ref) < tol
    assert res['MPJPE'].shape == (rows,)


def test_smplify_steps_vs_oracle(gpu_model, oracle_sd):
    m = synthetic.make_body_tensors('smplx')
    B, iters = 3, 2
    g = torch.Generator().manual_seed(41)
    norm = Posenormalizer(None, device='cuda', normalize=True, min_max=False, rot_rep='axis')
    smpl = SMPLX(m, batch_size=B).cuda()
    jm = smpl.joint_map
    gt_body = synthetic.toy_poses()[:B]
    gt_glob = torch.tensor([3.14159, 0., 0.]) + 0.2 * torch.randn(B, 3, generator=g)
    cam = torch.stack([0.2 * torch.randn(B, generator=g), 0.2 * torch.randn(B, generator=g),
                       20 + 20 * torch.rand(B, generator=g)], 1)
    betas_gt = torch.randn(B, 10, generator=g)
    hm = m['hands_mean'][None].expand(B, -1)        # SMPLX runs with the model's constant mean hand pose
    _, j = lbs_ref.body_forward(m, torch.cat([betas_gt, torch.zeros(B, 10)], 1),
                                torch.cat([gt_glob, gt_body, torch.zeros(B, 9), hm], 1), cam)
    j = j[:, jm]
    center = torch.full((B, 2), 512.)
    from oracle import fitting_ref as Fr
    kp = Fr.perspective_projection(j, 5000., center) + 2.0 * torch.randn(B, 49, 2, generator=g)
    conf = 0.3 + 0.7 * torch.rand(B, 49, generator=g)
    conf[:, 25:] = 0.
    kp2d = torch.cat([kp, conf[..., None]], -1)
    init_pose = torch.cat([gt_glob + 0.1, smpl.mean_poses[3:66].cpu()[None].repeat(B, 1)], 1)
    init_betas = smpl.mean_shape.cpu()[None].repeat(B, 1)
    init_cam = cam + torch.tensor([0.1, -0.1, 2.0])
    z_list = [torch.randn(B, 63, generator=g) for _ in range(5 * iters + 1)]
    ref_pose, ref_betas, ref_cam = fitting_loops.smplify(oracle_sd, m, jm, init_pose, init_betas, init_cam, center,
                                                         kp2d.clone(), norm.mean_poses.cpu(), norm.std_poses.cpu(),
                                                         z_list, num_iters=iters, sde_N=500, hand_mean=m['hands_mean'])
    args = types.SimpleNamespace(device='cuda', sde_N=500, time_strategy='3')
    gpu_model.engine = L.ENGINE_FP32
    try:
        pp = prior.DPoser(batch_size=B, args=args, model=gpu_model, sde=sde_lib.subVPSDE(0.1, 20., 1000),
                          normalizer=norm)
        fit = fitting.SMPLify(smpl, step_size=1e-2, batch_size=B, num_iters=iters, focal_length=5000., args=args,
                              pose_prior=pp)
        pose, betas, cam_t, reproj = fit(init_pose.cuda(), init_betas.cuda(), init_cam.cuda(), center.cuda(),
                                         kp2d.clone().cuda(), z_list=z_list)
    finally:
        gpu_model.engine = L.ENGINE_AUTO
    assert rel_err(pose.cpu() - init_pose, ref_pose - init_pose) < 5e-3
    assert rel_err(betas.cpu() - init_betas, ref_betas - init_betas) < 5e-3
    assert rel_err(cam_t.cpu() - init_cam, ref_cam - init_cam) < 5e-3
    assert reproj.shape == (B, 49)


@pytest.mark.parametrize('engine,tol', [(L.ENGINE_FP32, 2e-3), (L.ENGINE_TC, 5e-3)])
def test_motion_denoise_vs_reference_golden(gpu_model, engine, tol):
    """MotionDenoise.optimize against the REAL run/motion_denoising.py:199-300 (loops_golden.npz: two sequences,
    2 x 3 Adam steps, the reference's own draws replayed; LBS = the restatement on both sides)."""
    g = golden('loops_golden.npz')
    t = lambda k: torch.tensor(g[k])         # noqa: E731
    seq_len, n_seq, iters, spi = g['md_geom'].tolist()
    rows = seq_len * n_seq
    m = synthetic.make_body_tensors('smplx')
    norm = Posenormalizer(None, device='cuda', normalize=True, min_max=False, rot_rep='axis')
    bm = BodyModel(m, num_betas=10, batch_size=rows, model_type='smplx').cuda()
    gpu_model.engine = engine
    try:
        md = fitting.MotionDenoise(synthetic.default_config(), types.SimpleNamespace(device='cuda'), gpu_model, bm,
                                   sde_lib.subVPSDE(0.1, 20., 1000), norm, sde_N=500, batch_size=rows,
                                   seq_len=seq_len)
        md.poses = t('md_init').cuda()
        res = md.optimize(t('md_noisy').cuda(), gt_poses=t('md_gt').cuda(), time_strategy='3', sample_trun=4.0,
                          iterations=iters, steps_per_iter=spi, z_list=list(t('md_z')))
    finally:
        gpu_model.engine = L.ENGINE_AUTO
    init = t('md_init')
    assert rel_err(res['pose_body'].cpu() - init, t('md_smooth') - init) < tol
    assert rel_err(res['MPJPE'], g['md_MPJPE']) < tol and rel_err(res['MPVPE'], g['md_MPVPE']) < tol
    assert rel_err(res['init_MPJPE'], g['md_init_MPJPE']) < 1e-4


@pytest.mark.parametrize('engine,tol', [(L.ENGINE_FP32, 2e-3), (L.ENGINE_TC, 5e-3)])
def test_smplify_vs_reference_golden(gpu_model, engine, tol):
    """SMPLify.__call__ against the REAL run/smplify.py:168-281 run image by image (B=1, as the reference requires):
    2 camera + 10 body Adam steps, hands at the model's non-zero mean pose (smplx defaults)."""
    g = golden('loops_golden.npz')
    t = lambda k: torch.tensor(g[k])         # noqa: E731
    (iters,) = g['sf_iters'].tolist()
    m = synthetic.make_body_tensors('smplx')
    B = g['sf_init_pose'].shape[0]
    norm = Posenormalizer(None, device='cuda', normalize=True, min_max=False, rot_rep='axis')
    smpl = SMPLX(m, batch_size=B).cuda()
    args = types.SimpleNamespace(device='cuda', sde_N=500, time_strategy='3')
    gpu_model.engine = engine
    try:
        pp = prior.DPoser(batch_size=B, args=args, model=gpu_model, sde=sde_lib.subVPSDE(0.1, 20., 1000),
                          normalizer=norm)
        fit = fitting.SMPLify(smpl, step_size=1e-2, batch_size=B, num_iters=iters, focal_length=5000., args=args,
                              pose_prior=pp)
        kp2d = t('sf_kp2d').cuda()
        pose, betas, cam_t, reproj = fit(t('sf_init_pose').cuda(), t('sf_init_betas').cuda(), t('sf_init_cam').cuda(),
                                         t('sf_center').cuda(), kp2d, z_list=list(t('sf_z')))
    finally:
        gpu_model.engine = L.ENGINE_AUTO
    assert float(kp2d[:, 9, 2].abs().max()) == 0.            # B-13: ignored joints zeroed in the caller's tensor
    assert rel_err(pose.cpu() - t('sf_init_pose'), t('sf_pose') - t('sf_init_pose')) < tol
    assert rel_err(betas.cpu() - t('sf_init_betas'), t('sf_betas') - t('sf_init_betas')) < tol
    assert rel_err(cam_t.cpu() - t('sf_init_cam'), t('sf_cam') - t('sf_init_cam')) < tol
    assert rel_err(reproj.cpu(), t('sf_reproj')) < tol


@pytest.mark.parametrize('per_problem', [False, True])
def test_fused_body_fitting_loss_vs_oracle(per_problem):
    """dpb_fit_loss (projection + GMoF + angle / shape priors, loss and cotangents in one kernel) against the
    oracle's restatement of fitting_losses.py:59-103, value and gradients w.r.t. joints, pose and betas."""
    from dposer_b200 import fitting_losses as F
    from oracle import fitting_ref as Fr
    B, K = 37, 49
    g = torch.Generator().manual_seed(3)
    joints = torch.randn(B, K, 3, generator=g) * 0.4 + torch.tensor([0., 0., 25.])
    kp = torch.randn(B, K, 2, generator=g) * 60 + 512
    conf = torch.rand(B, K, generator=g)
    conf[:, 30:] = 0
    center = torch.full((B, 2), 512.) + torch.randn(B, 2, generator=g)
    pose = torch.randn(B, 69, generator=g) * 0.3
    betas = torch.randn(B, 10, generator=g)
    prior_scalar = torch.tensor(0.731)

    def run(fn, dev):
        leaves = [t.clone().to(dev).requires_grad_(True) for t in (joints, pose, betas)]
        return fn(*leaves), leaves

    def oracle(j, p, b):
        if not per_problem:
            return Fr.body_fitting_loss(p, b, j, center, kp, conf, prior_scalar)
        # B independent single-image problems (each normalised as a batch of one) + the shared prior term
        zero = torch.tensor(0.)
        per = [Fr.body_fitting_loss(p[i:i + 1], b[i:i + 1], j[i:i + 1], center[i:i + 1], kp[i:i + 1],
                                    conf[i:i + 1], zero) for i in range(B)]
        return torch.stack(per).sum() + (4.78 ** 2) * prior_scalar

    ref, lr = run(oracle, 'cpu')
    ref.backward()
    got, lg = run(lambda j, p, b: F.body_fitting_loss(p, b, j, None, center.cuda(), kp.cuda(), conf.cuda(),
                                                      lambda *_: prior_scalar.cuda(), 0, per_problem=per_problem),
                  'cuda')
    got.backward()
    assert abs(float(got.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    for a, b_ in zip(lg, lr):
        scale = b_.grad.abs().max().clamp_min(1e-12)
        assert float((a.grad.cpu() - b_.grad).abs().max() / scale) < 1e-5


def test_motion_denoise_step_graphs_replay(gpu_model):
    """graphs=True captures every Adam step once and replays it for further batches of sequences: the first (capturing)
    call, a replaying call and the plain launch sequence must agree for the same seed."""
    m = synthetic.make_body_tensors('smplx')
    seq_len, n_seq = 6, 2
    rows = seq_len * n_seq
    g = torch.Generator().manual_seed(77)
    gt = synthetic.toy_poses()[:rows] + 0.02 * torch.randn(rows, 63, generator=g)
    norm = Posenormalizer(None, device='cuda', normalize=True, min_max=False, rot_rep='axis')
    bm = BodyModel(m, num_betas=10, batch_size=rows, model_type='smplx').cuda()
    with torch.no_grad():
        noisy = bm(pose_body=gt.cuda()).Jtr[:, :22] + 0.04 * torch.randn(rows, 22, 3, generator=g).cuda()
    init = (0.01 * torch.randn(rows, 63, generator=g)).cuda()
    kw = dict(time_strategy='3', sample_trun=4.0, iterations=1, steps_per_iter=3)

    def run(md, graphs):
        md.poses = init.clone()
        return md.optimize(noisy, graphs=graphs, **kw)['pose_body_raw']

    def make():
        return fitting.MotionDenoise(synthetic.default_config(), types.SimpleNamespace(device='cuda'), gpu_model, bm,
                                     sde_lib.subVPSDE(0.1, 20., 1000), norm, sde_N=500, batch_size=rows, seq_len=seq_len)
    torch.manual_seed(5)
    plain = run(make(), False)
    torch.manual_seed(5)
    md = make()
    captured = run(md, True)
    replayed = run(md, True)          # same buffers, same schedule: graph replays only
    assert torch.isfinite(plain).all()
    assert rel_err(captured.cpu() - init.cpu(), plain.cpu() - init.cpu()) < 1e-5
    assert rel_err(replayed.cpu() - init.cpu(), plain.cpu() - init.cpu()) < 1e-5
