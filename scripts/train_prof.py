"""Two eager training steps at the reference batch size, for ncu (run under gpurun)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from dposer_b200 import losses, sde_lib, synthetic
from dposer_b200.ema import ExponentialMovingAverage
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
cfg = synthetic.default_config()
model = synthetic.make_score_model(42).cuda(); model.train()
state = dict(optimizer=losses.get_optimizer(cfg, model.parameters()), model=model,
             ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=100)
step_fn = losses.get_step_fn(sde_lib.subVPSDE(0.1, 20., 1000), True, losses.optimization_manager(cfg), reduce_mean=True)
data = synthetic.toy_poses()
data = data[torch.randint(0, data.shape[0], (B,))].cuda()
for _ in range(2):
    ld = step_fn(state, data)
torch.cuda.synchronize(); print('ok', float(ld['step_loss']))
