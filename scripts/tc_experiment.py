"""Timing experiment for the fused sampler kernel (run under gpurun): DPB_TC_DEBUG variants.

The knobs exist only in an instrumented build of the library:
    DPB_BUILD_DEFINES=DPB_TC_KNOBS python -c "from dposer_b200 import build; build.build(force=True)"   # or DPB_TC_PROFILE
    DPB_LIBRARY=$PWD/dposer_b200/lib/libdposer_b200_prof.so PROF_DBG=0,1,64 python scripts/tc_experiment.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dposer_b200 import _lib as L, sampling, sde_lib, synthetic
B = int(os.environ.get('PROF_B', 18944)); N = int(os.environ.get('PROF_N', 20))
model = synthetic.make_score_model(42).cuda(); model.engine = L.ENGINE_TC
cfg = synthetic.default_config()
fn = sampling.get_sampling_fn(cfg, sde_lib.subVPSDE(0.1, 20., N), (B, 63), lambda x: x, 1e-3, device='cuda', return_trajs=False)
z = torch.randn(B, 63)
for dbg in os.environ.get('PROF_DBG', '0,1,2,3').split(','):
    os.environ['DPB_TC_DEBUG'] = dbg
    for _ in range(2): fn(model, z=z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn(model, z=z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f'debug={dbg}: {ms:.3f} ms  -> {8646656*B*N/(ms*1e-3)/1e12:.1f} TFLOP/s-equivalent', flush=True)
