"""Small fixed workload for ncu captures (run under gpurun):  sampler 148 tiles x N steps + SMPL LBS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dposer_b200 import _lib as L, sampling, sde_lib, synthetic
from dposer_b200.body_model import BodyModel

B = int(os.environ.get('PROF_B', 18944))
N = int(os.environ.get('PROF_N', 4))
BL = int(os.environ.get('PROF_BL', 16384))
engine = {'tc': L.ENGINE_TC, 'fp32': L.ENGINE_FP32}[os.environ.get('PROF_ENGINE', 'tc')]
model = synthetic.make_score_model(42).cuda()
model.engine = engine
cfg = synthetic.default_config()
fn = sampling.get_sampling_fn(cfg, sde_lib.subVPSDE(0.1, 20., N), (B, 63), lambda x: x, 1e-3, device='cuda',
                              return_trajs=False)
z = torch.randn(B, 63)
bm = BodyModel(synthetic.make_body_tensors('smpl'), batch_size=BL, model_type='smpl').cuda()
inp = {k: v.cuda() for k, v in synthetic.lbs_inputs(BL, 'smpl').items()}
with torch.no_grad():
    for _ in range(2):
        fn(model, z=z)
        bm(**inp)
    torch.cuda.synchronize()
