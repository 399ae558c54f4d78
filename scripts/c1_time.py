"""configs[0] literal (500 poses, N = 1000): EM and EM + Langevin, small-batch engine on / off (DPB_TC_SMALL)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dposer_b200 import sampling, sde_lib, synthetic

model = synthetic.make_score_model(42).cuda()
sde = sde_lib.subVPSDE(0.1, 20., 1000)
for B in (500, 1024, 128):
    z = torch.randn(B, 63, generator=torch.Generator().manual_seed(5)).cuda()
    for corr in ('none', 'langevin'):
        cfg = synthetic.default_config()
        cfg.sampling.corrector = corr
        fn = sampling.get_sampling_fn(cfg, sde, (B, 63), lambda x: x, 1e-3, device='cuda', return_trajs=False)
        outs = {}
        for small in ('1', '0'):
            os.environ['DPB_TC_SMALL'] = small
            torch.manual_seed(0)
            fn(model, z=z)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(2):
                torch.manual_seed(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); _, x = fn(model, z=z); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            outs[small] = x
            print(f'B={B} corrector={corr} small={small}: {best:.1f} ms ({best:.3f} us/step... {B / best * 1e3:.0f} poses/s)', flush=True)
        d = (outs['1'] - outs['0']).norm() / outs['0'].norm()
        print(f'   rel diff small vs whole-tile engine: {float(d):.2e}', flush=True)
