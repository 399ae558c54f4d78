"""GPU diagnostic: error of both engines vs the oracle per t / batch (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import torch
from dposer_b200 import _lib as L, synthetic, sde_lib, utils as mutils
from oracle import score_ref as S

model = synthetic.make_score_model(42).cuda()
sd = S.make_state_dict(42)
sde, osde = sde_lib.subVPSDE(0.1, 20., 1000), S.SubVP()
score_fn = mutils.get_score_fn(sde, model, train=False, continuous=True)
for engine, name in [(L.ENGINE_FP32, 'fp32'), (L.ENGINE_TC, 'tc')]:
    model.engine = engine
    for B in [7, 128, 500]:
        gen = torch.Generator().manual_seed(B)
        x = torch.randn(B, 63, generator=gen) * 1.5
        for tv in [1.0, 0.5, 0.1, 0.01, 1e-3]:
            ref = S.score_fn(sd, osde, x, torch.ones(B) * tv)
            try:
                out = score_fn(x.cuda(), torch.ones(B, device='cuda') * tv, None, None).cpu()
            except Exception as e:
                print(name, B, tv, 'ERR', repr(e)[:100]); continue
            rel = float((out - ref).norm() / ref.norm())
            row = float(((out - ref).norm(dim=1) / ref.norm(dim=1)).max())
            mx = float((out - ref).abs().max() / ref.abs().max())
            print(f'{name} B={B} t={tv}: rel {rel:.2e} rowmax {row:.2e} maxrel {mx:.2e} nan={bool(torch.isnan(out).any())}')
