"""Time the native training step (run under gpurun): python scripts/train_time.py [B]"""
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from dposer_b200 import losses, sde_lib, synthetic
from dposer_b200.ema import ExponentialMovingAverage
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
GRAPH = len(sys.argv) > 2 and sys.argv[2] == 'graph'
cfg = synthetic.default_config()
model = synthetic.make_score_model(42).cuda(); model.train()
opt = losses.get_optimizer(cfg, model.parameters())
ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
state = dict(optimizer=opt, model=model, ema=ema, step=0)
sde = sde_lib.subVPSDE(0.1, 20., 1000)
step_fn = losses.get_step_fn(sde, True, losses.optimization_manager(cfg), reduce_mean=True, graph=GRAPH)
data = synthetic.toy_poses()
data = data[torch.randint(0, data.shape[0], (B,))].cuda()
for _ in range(5): step_fn(state, data)
torch.cuda.synchronize()
for n in (20, 100):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): ld = step_fn(state, data)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f'B={B} graph={GRAPH} n={n}: device {e0.elapsed_time(e1)/n:.3f} ms/step, wall {(t1-t0)*1e3/n:.3f} ms/step, loss {float(ld["step_loss"]):.3f}')
# device-only: loss+grad call alone
import ctypes as C
from dposer_b200 import _lib as L
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): step_fn(state, data)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))
