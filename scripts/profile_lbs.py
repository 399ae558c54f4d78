"""LBS-only workload for ncu captures (run under gpurun): SMPL (and optionally SMPL-X) forward at PROF_BL poses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dposer_b200 import synthetic
from dposer_b200.body_model import BodyModel

BL = int(os.environ.get('PROF_BL', 65536))
mt = os.environ.get('PROF_MODEL', 'smpl')
bm = BodyModel(synthetic.make_body_tensors(mt), batch_size=BL, model_type=mt).cuda()
inp = {k: v.cuda() for k, v in synthetic.lbs_inputs(BL, mt).items()}
with torch.no_grad():
    for _ in range(3):
        bm(**inp)
    torch.cuda.synchronize()
