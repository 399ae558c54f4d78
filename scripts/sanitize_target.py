"""Small pass over every tcgen05 kernel and the LBS backward for `compute-sanitizer --tool memcheck` (run under gpurun)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from dposer_b200 import _lib as L, sampling, sde_lib, synthetic
from dposer_b200.body_model import BodyModel
model = synthetic.make_score_model(42).cuda(); model.engine = L.ENGINE_TC
cfg = synthetic.default_config()
for B, N in [(300, 3), (19200, 2)]:
    fn = sampling.get_sampling_fn(cfg, sde_lib.subVPSDE(0.1, 20., N), (B, 63), lambda x: x, 1e-3, device='cuda')
    traj, x = fn(model, z=torch.randn(B, 63))
    torch.cuda.synchronize(); print('sampler ok', B, float(x.abs().mean()))
x = torch.randn(500, 63).cuda()
out = model(x, torch.full((500,), 300., device='cuda')); torch.cuda.synchronize(); print('fwd ok')
for mt, B in [('smpl', 200), ('smplx', 130)]:
    bm = BodyModel(synthetic.make_body_tensors(mt), batch_size=B, model_type=mt).cuda()
    inp = {k: v.cuda().requires_grad_(True) for k, v in synthetic.lbs_inputs(B, mt).items()}
    o = bm(**inp); (o.v.sum() + o.Jtr.sum()).backward(); torch.cuda.synchronize(); print('lbs ok', mt)
