#!/usr/bin/env python
"""Golden fixtures for the VPSDE / VESDE code paths of the score function and the Euler-Maruyama sampler, from the
REAL reference (build container only):   python tests/golden/make_golden_sde_variants.py
Reuses the import shims and the weight recipe of make_golden.py."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (sets up the shims, imports the reference modules)


def main():
    cfg = G.get_config()
    cfg.device = torch.device('cpu')
    model = G.build_reference_model(cfg)
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(7, 63, generator=gen) * 1.2
    out = {'x': x.numpy()}
    sdes = {'vp': G.sde_lib.VPSDE(0.1, 20., N=1000), 've': G.sde_lib.VESDE(0.01, 50., N=1000)}
    with torch.no_grad():
        for name, sde in sdes.items():
            fn = G.mutils.get_score_fn(sde, model, train=False, continuous=True)
            for tv in [1.0, 0.5, 0.01]:
                out[f'{name}_score_{tv}'] = fn(x, torch.ones(7) * tv, None, None).numpy()
        # Euler-Maruyama sampler, VPSDE, 8 steps, replayed draws (sampling.py:182-188)
        N, B = 8, 5
        cfg.sampling.corrector = 'none'
        cfg.sampling.probability_flow = False
        vp = G.sde_lib.VPSDE(0.1, 20., N=N)
        sfn = G.sampling.get_sampling_fn(cfg, vp, (B, 63), lambda v: v, 1e-3, device='cpu')
        z0 = torch.randn(B, 63, generator=gen)
        torch.manual_seed(4321)
        traj, xm = sfn(model, z=z0.clone())
        torch.manual_seed(4321)
        noise = torch.stack([torch.randn(B, 63) for _ in range(N)])
        out.update(vp_em_z0=z0.numpy(), vp_em_noise=noise.numpy(), vp_em_out=xm.numpy(), vp_em_last=traj[-1].numpy())
    np.savez(os.path.join(HERE, 'sde_variants_golden.npz'), **out)
    print('wrote sde_variants_golden.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
