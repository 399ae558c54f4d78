import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dposer_b200 import sampling, sde_lib, synthetic
model = synthetic.make_score_model(42).cuda()
cfg = synthetic.default_config(); cfg.sampling.corrector = 'langevin'
fn = sampling.get_sampling_fn(cfg, sde_lib.subVPSDE(0.1, 20., 1000), (500, 63), lambda x: x, 1e-3, device='cuda', return_trajs=False)
z = torch.randn(500, 63).cuda()
for dbg in sys.argv[1].split():
    os.environ['DPB_TC_DEBUG'] = dbg
    fn(model, z=z); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(model, z=z); e1.record(); torch.cuda.synchronize()
    print('debug', dbg, e0.elapsed_time(e1), 'ms', flush=True)
# EM-only reference at N and 2N steps
for N in (1000, 2000):
    cfg2 = synthetic.default_config()
    fn2 = sampling.get_sampling_fn(cfg2, sde_lib.subVPSDE(0.1, 20., N), (500, 63), lambda x: x, 1e-3, device='cuda', return_trajs=False)
    os.environ['DPB_TC_DEBUG'] = '0'
    fn2(model, z=z); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn2(model, z=z); e1.record(); torch.cuda.synchronize()
    print('EM steps', N, e0.elapsed_time(e1), 'ms', flush=True)
