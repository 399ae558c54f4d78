import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dposer_b200 import sampling, sde_lib, synthetic
model = synthetic.make_score_model(42).cuda()
cfg = synthetic.default_config(); cfg.sampling.corrector = 'langevin'
fn = sampling.get_sampling_fn(cfg, sde_lib.subVPSDE(0.1, 20., 6), (500, 63), lambda x: x, 1e-3, device='cuda', return_trajs=False)
z = torch.randn(500, 63).cuda()
fn(model, z=z); torch.cuda.synchronize()
