"""One Adam step of configs 4 and 5 for an ncu launch list (run under gpurun):
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python scripts/profile_steps.py"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dposer_b200 import fitting, prior, sde_lib, synthetic
from dposer_b200.body_model import BodyModel, SMPLX
from dposer_b200.misc import Posenormalizer

dev = 'cuda'
model = synthetic.make_score_model(42).cuda()
norm = Posenormalizer(None, device=dev, normalize=True, min_max=False, rot_rep='axis')
cfg = synthetic.default_config()
mx = synthetic.make_body_tensors('smplx')
which = os.environ.get('PROF_TASK', 'c4c5')
if 'c4' in which:
    n_seq, L = int(os.environ.get('C4_SEQ', 256)), 60
    rows = n_seq * L
    bm = BodyModel(mx, num_betas=10, batch_size=rows, model_type='smplx').cuda()
    ges, _ = synthetic.gesture_sequences()
    gt = ges[:L].repeat(n_seq, 1).cuda()
    with torch.no_grad():
        jn = bm(pose_body=gt, need_verts=False).Jtr[:, :22] + 0.04 * torch.randn(rows, 22, 3, device=dev)
    md = fitting.MotionDenoise(cfg, types.SimpleNamespace(device=dev), model, bm, sde_lib.subVPSDE(0.1, 20., 1000), norm,
                               sde_N=500, batch_size=rows, seq_len=L)
    torch.cuda.synchronize()
    print('== c4 steps', flush=True)
    md.optimize(jn, time_strategy='3', sample_trun=4.0, iterations=1, steps_per_iter=2)
    torch.cuda.synchronize()
    del md, bm
if 'c5' in which:
    B = int(os.environ.get('C5_B', 131072))
    smpl = SMPLX(mx, batch_size=B).cuda()
    g = torch.Generator().manual_seed(41)
    body = synthetic.toy_poses().repeat((B + 499) // 500, 1)[:B]
    glob = torch.tensor([3.14159, 0., 0.]) + 0.2 * torch.randn(B, 3, generator=g)
    cam = torch.stack([0.2 * torch.randn(B, generator=g), 0.2 * torch.randn(B, generator=g), 20 + 20 * torch.rand(B, generator=g)], 1)
    with torch.no_grad():
        j = smpl(betas=torch.randn(B, 10, generator=g).cuda(), body_pose=body.cuda(), global_orient=glob.cuda(), transl=cam.cuda()).joints
    center = torch.full((B, 2), 512., device=dev)
    kp = torch.stack([5000 * j[..., 0] / j[..., 2] + 512, 5000 * j[..., 1] / j[..., 2] + 512], -1)
    conf = torch.ones(B, 49, device=dev); conf[:, 25:] = 0
    kp2d = torch.cat([kp, conf[..., None]], -1)
    args = types.SimpleNamespace(device=dev, sde_N=500, time_strategy='3')
    pp = prior.DPoser(batch_size=B, args=args, model=model, sde=sde_lib.subVPSDE(0.1, 20., 1000), normalizer=norm)
    fit = fitting.SMPLify(smpl, step_size=1e-2, batch_size=B, num_iters=1, focal_length=5000., args=args, pose_prior=pp)
    init_pose = torch.cat([glob + 0.1, smpl.mean_poses[3:66].cpu()[None].repeat(B, 1)], 1).cuda()
    torch.cuda.synchronize()
    print('== c5 steps', flush=True)
    fit(init_pose, smpl.mean_shape[None].repeat(B, 1).cuda(), (cam + torch.tensor([0.1, -0.1, 2.0])).cuda(), center, kp2d)
    torch.cuda.synchronize()
