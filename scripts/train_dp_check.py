"""2-rank data-parallel training step == one-rank step on the concatenated batch (run: gpurun --gpus 2 -- torchrun ...)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from dposer_b200 import dist as D, losses, sde_lib, synthetic
from dposer_b200.ema import ExponentialMovingAverage
rank, local, world = D.init_from_env()
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
cfg = synthetic.default_config(); cfg.optim.warmup = 0; cfg.model.dropout = 0.0
sde = sde_lib.subVPSDE(0.1, 20., 1000)
Bl = 96
g = torch.Generator().manual_seed(3)
data = synthetic.toy_poses()[:Bl * world]
t = torch.rand(Bl * world, generator=g) * (1 - 1e-5) + 1e-5
z = torch.randn(Bl * world, 63, generator=g)
def run(dp, sl):
    model = synthetic.make_score_model(42).to(dev); model.train()
    state = dict(optimizer=losses.get_optimizer(cfg, model.parameters()), model=model,
                 ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
    fn = losses.get_step_fn(sde, True, losses.optimization_manager(cfg), reduce_mean=True, data_parallel=dp)
    for _ in range(2):
        ld = fn(state, data[sl].to(dev), t=t[sl], z=z[sl], drop_mask=torch.ones(5, sl.stop - sl.start, 1024, dtype=torch.uint8))
    return state['optimizer'].flat_p.clone(), float(ld['step_loss'])
p_dp, l_dp = run(True, slice(rank * Bl, (rank + 1) * Bl))
p_one, l_one = run(False, slice(0, Bl * world))
d = float((p_dp - p_one).abs().max())
ref = torch.cat([p.detach().reshape(-1) for p in synthetic.make_score_model(42).parameters()]).to(dev)
step = float((p_one - ref).abs().max())
# all ranks must hold identical parameters
chk = p_dp.double().sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN); torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
print(f'rank {rank}: max |p_dp - p_one| = {d:.3e} (largest update {step:.3e}), ranks identical: {bool(lo == hi)}', flush=True)
assert d < 2e-2 * step and bool(lo == hi)
torch.distributed.destroy_process_group()
