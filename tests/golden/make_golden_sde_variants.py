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
        # ReverseDiffusion / AncestralSampling predictors (sampling.py:210-259): the reference classes called directly, step
        # by step over pc_sampler's time grid (its own pc_sampler cannot drive them: update_fn arity), draws replayed
        # (N = 1000 schedules: VPSDE's discrete betas are only valid for beta_max / N < 1; the LAST 8 steps of the grid,
        # entered like pc_sampler's 'denoise' task does with start_step = N - 8)
        vp1k, sub1k = G.sde_lib.VPSDE(0.1, 20., N=1000), G.sde_lib.subVPSDE(0.1, 20., N=1000)
        for tag, sde, cls in [('vp_rd', vp1k, G.sampling.ReverseDiffusionPredictor),
                              ('vp_anc', vp1k, G.sampling.AncestralSamplingPredictor),
                              ('sub_rd', sub1k, G.sampling.ReverseDiffusionPredictor)]:
            score_fn = G.mutils.get_score_fn(sde, model, train=False, continuous=True)
            sf = lambda xx, tt, _f=score_fn: _f(xx, tt, None, None)   # noqa: E731  (predictors call score_fn(x, t))
            pred = cls(sde, sf, False)
            if hasattr(pred, 'rsde'):
                pred.rsde = sde.reverse(lambda xx, tt, c=None, m=None, _f=score_fn: _f(xx, tt, c, m), False)
            xx = z0.clone()
            torch.manual_seed(999)
            for tt in torch.linspace(sde.T, 1e-3, sde.N)[sde.N - N:]:
                xx, xm = pred.update_fn(xx, torch.ones(B) * tt)
            torch.manual_seed(999)
            noise = torch.stack([torch.randn(B, 63) for _ in range(N)])
            out.update({f'{tag}_noise': noise.numpy(), f'{tag}_out': xm.numpy(), f'{tag}_last': xx.numpy()})
    # VESDE through the same three predictors (sde_lib.py:234-295; score_fn utils.py:164-180: labels = sigma(t), no std
    # division): Euler-Maruyama through the reference's own pc_sampler, the other two called directly, draws replayed
    with torch.no_grad():
        ve8 = G.sde_lib.VESDE(0.01, 50., N=N)
        sfn = G.sampling.get_sampling_fn(cfg, ve8, (B, 63), lambda v: v, 1e-5, device='cpu')
        zv = z0 * 50.
        torch.manual_seed(2468)
        traj, xm = sfn(model, z=zv.clone())
        torch.manual_seed(2468)
        noise = torch.stack([torch.randn(B, 63) for _ in range(N)])
        out.update(ve_em_z0=zv.numpy(), ve_em_noise=noise.numpy(), ve_em_out=xm.numpy(), ve_em_last=traj[-1].numpy())
        for tag, cls in [('ve_rd', G.sampling.ReverseDiffusionPredictor), ('ve_anc', G.sampling.AncestralSamplingPredictor)]:
            score_fn = G.mutils.get_score_fn(ve8, model, train=False, continuous=True)
            sf = lambda xx, tt, _f=score_fn: _f(xx, tt, None, None)   # noqa: E731
            pred = cls(ve8, sf, False)
            if hasattr(pred, 'rsde'):
                pred.rsde = ve8.reverse(lambda xx, tt, c=None, m=None, _f=score_fn: _f(xx, tt, c, m), False)
            xx = zv.clone()
            torch.manual_seed(1357)
            for tt in torch.linspace(ve8.T, 1e-5, ve8.N):
                xx, xm = pred.update_fn(xx, torch.ones(B) * tt)
            torch.manual_seed(1357)
            noise = torch.stack([torch.randn(B, 63) for _ in range(N)])
            out.update({f'{tag}_noise': noise.numpy(), f'{tag}_out': xm.numpy(), f'{tag}_last': xx.numpy()})
    # likelihood / latent code under the probability-flow ODE (likelihood.py:40-113), Rademacher probe replayed
    from lib.algorithms.advanced import likelihood as ref_lik
    sub1000 = G.sde_lib.subVPSDE(0.1, 20., N=1000)
    data = torch.tensor(np.load(os.path.join(G.REF, 'examples/toy_data.npz'))['pose_samples'][:4]).float()
    norm = torch.load(os.path.join(G.REF, 'data/AMASS/amass_processed/version1/train/axis_normalize2.pt'))
    data = (data - norm['mean_poses']) / norm['std_poses']
    lfn = ref_lik.get_likelihood_fn(sub1000, lambda v: v, rtol=1e-5, atol=1e-5, eps=1e-5)
    torch.manual_seed(77)
    bpd, z, nfe = lfn(model, data.clone())
    torch.manual_seed(77)
    epsilon = torch.randint_like(data, low=0, high=2).float() * 2 - 1.
    # one function evaluation too (drift and divergence at t = 0.3), the unit the kernel replaces
    x_t = data * 0.7 + 0.2
    vec_t = torch.ones(4) * 0.3
    score_fn = G.mutils.get_score_fn(sub1000, model, train=False, continuous=True)
    rs = sub1000.reverse(score_fn, probability_flow=True)
    drift = rs.sde(x_t, vec_t, condition=None, mask=None)[0]
    div = ref_lik.get_div_fn(lambda xx, tt: rs.sde(xx, tt, condition=None, mask=None)[0])(x_t.clone(), vec_t, epsilon)
    out.update(lik_data=data.numpy(), lik_eps=epsilon.numpy(), lik_bpd=bpd.numpy(), lik_z=z.numpy(), lik_nfe=np.array(nfe),
               lik_xt=x_t.numpy(), lik_drift=drift.detach().numpy(), lik_div=div.detach().numpy())
    np.savez(os.path.join(HERE, 'sde_variants_golden.npz'), **out)
    print('wrote sde_variants_golden.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
