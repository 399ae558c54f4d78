"""torch.profiler view of the task-level loops (configs 4 and 5): which kernels the time goes to (run under gpurun)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dposer_b200 import fitting, prior, sde_lib, synthetic
from dposer_b200.body_model import BodyModel, SMPLX
from dposer_b200.misc import Posenormalizer

dev = 'cuda'
model = synthetic.make_score_model(42).cuda()
norm = Posenormalizer(None, device=dev, normalize=True, min_max=False, rot_rep='axis')
mx = synthetic.make_body_tensors('smplx')


def report(name, fn):
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    print('=====', name)
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))


n_seq, L = 32, 60
rows = n_seq * L
bm = BodyModel(mx, num_betas=10, batch_size=rows, model_type='smplx').cuda()
ges, _ = synthetic.gesture_sequences()
gt = ges[:L].repeat(n_seq, 1).cuda()
with torch.no_grad():
    jn = bm(pose_body=gt).Jtr[:, :22] + 0.04 * torch.randn(rows, 22, 3, device=dev)
cfg = synthetic.default_config()
md = fitting.MotionDenoise(cfg, types.SimpleNamespace(device=dev), model, bm, sde_lib.subVPSDE(0.1, 20., 1000), norm,
                           sde_N=500, batch_size=rows, seq_len=L)
report('motion denoising, 1920 frames, 20 Adam steps',
       lambda: md.optimize(jn, gt_poses=gt, time_strategy='3', sample_trun=4.0, iterations=1, steps_per_iter=20))

B, iters = 2048, 4
smpl = SMPLX(mx, batch_size=B).cuda()
g = torch.Generator().manual_seed(41)
body = synthetic.toy_poses().repeat((B + 499) // 500, 1)[:B]
glob = torch.tensor([3.14159, 0., 0.]) + 0.2 * torch.randn(B, 3, generator=g)
cam = torch.stack([0.2 * torch.randn(B, generator=g), 0.2 * torch.randn(B, generator=g), 20 + 20 * torch.rand(B, generator=g)], 1)
betas = torch.randn(B, 10, generator=g)
with torch.no_grad():
    j = smpl(betas=betas.cuda(), body_pose=body.cuda(), global_orient=glob.cuda(), transl=cam.cuda()).joints
center = torch.full((B, 2), 512., device=dev)
kp = torch.stack([5000 * j[..., 0] / j[..., 2] + 512, 5000 * j[..., 1] / j[..., 2] + 512], -1) + 2 * torch.randn(B, 49, 2, device=dev)
conf = 0.3 + 0.7 * torch.rand(B, 49, device=dev); conf[:, 25:] = 0
kp2d = torch.cat([kp, conf[..., None]], -1)
args = types.SimpleNamespace(device=dev, sde_N=500, time_strategy='3')
pp = prior.DPoser(batch_size=B, args=args, model=model, sde=sde_lib.subVPSDE(0.1, 20., 1000), normalizer=norm)
fit = fitting.SMPLify(smpl, step_size=1e-2, batch_size=B, num_iters=iters, focal_length=5000., args=args, pose_prior=pp)
init_pose = torch.cat([glob + 0.1, smpl.mean_poses[3:66].cpu()[None].repeat(B, 1)], 1).cuda()
report('SMPLify, 2048 poses, 24 Adam steps',
       lambda: fit(init_pose, smpl.mean_shape[None].repeat(B, 1), (cam + torch.tensor([0.1, -0.1, 2.0])).cuda(), center, kp2d))
