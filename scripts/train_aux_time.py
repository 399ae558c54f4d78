"""Time the auxiliary training step (reference config: batch 1280, denoise_steps 10, SMPL-X body model).  Run under gpurun."""
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from dposer_b200 import losses, sde_lib, synthetic
from dposer_b200.body_model import BodyModel
from dposer_b200.ema import ExponentialMovingAverage
from dposer_b200.misc import Posenormalizer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = synthetic.default_config()
model = synthetic.make_score_model(42).cuda(); model.train()
bm = BodyModel(synthetic.make_body_tensors('smplx'), num_betas=10, batch_size=B, model_type='smplx').cuda()
norm = Posenormalizer(None, device='cuda', normalize=True, min_max=False, rot_rep='axis')
state = dict(optimizer=losses.get_optimizer(cfg, model.parameters()), model=model,
             ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
fn = losses.get_step_fn(sde_lib.subVPSDE(0.1, 20., 1000), True, losses.optimization_manager(cfg), reduce_mean=True,
                        auxiliary_loss=True, denormalize=norm.offline_denormalize, body_model=bm, rot_rep='axis', denoise_steps=N)
toy = synthetic.toy_poses()
data = norm.offline_normalize(toy[torch.randint(0, toy.shape[0], (B,))].cuda())
for _ in range(3): ld = fn(state, data)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
t0 = time.perf_counter(); e0.record()
for _ in range(n): ld = fn(state, data)
e1.record(); torch.cuda.synchronize()
print(f'aux step B={B} N={N}: device {e0.elapsed_time(e1)/n:.2f} ms/step, wall {(time.perf_counter()-t0)*1e3/n:.2f} ms/step', {k: round(float(v), 4) for k, v in ld.items()})
