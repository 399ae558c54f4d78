#!/usr/bin/env python
"""Golden fixture for the auxiliary training loss from the REAL reference (build container only):
    python tests/golden/make_golden_train_aux.py
Runs lib/algorithms/advanced/losses.get_step_fn(train=True, auxiliary_loss=True) -- multi_step_denoise under autograd, the
estimate through denormalize + a BodyModel-compatible object over oracle/lbs_ref.py (third-party smplx is absent) on the
synthetic SMPL-X tensors -- for one step on 6 normalised AMASS poses with 3 denoise steps; records t, z and the 15 dropout
masks, the four loss values, the gradient norm and sampled gradients.  Asserts that oracle/train_ref.aux_loss reproduces it."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
import make_golden_train as GT  # noqa: E402

from lib.algorithms.advanced import losses as ref_losses  # noqa: E402
from lib.algorithms.ema import ExponentialMovingAverage  # noqa: E402
from oracle import lbs_ref  # noqa: E402
from oracle import score_ref as S  # noqa: E402
from oracle import train_ref as T  # noqa: E402
from dposer_b200 import synthetic  # noqa: E402

B, NSTEPS = 6, 3


class Struct:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class RefBodyModel:
    """lib/body_model/body_model.py BodyModel surface (model_type='smplx', zero betas / hands) over the LBS restatement."""

    def __init__(self, m):
        self.m = m

    def __call__(self, pose_body=None, **kw):
        Bn = pose_body.shape[0]
        shape = torch.zeros(Bn, self.m['shapedirs'].shape[2])
        full = torch.cat([torch.zeros(Bn, 3), pose_body, torch.zeros(Bn, 9 + 90)], 1)
        v, j = lbs_ref.body_forward(self.m, shape, full)
        return Struct(v=v, Jtr=j)


def main():
    cfg = G.get_config()
    cfg.device = torch.device('cpu')
    model = G.build_reference_model(cfg)
    model.train()
    names = [n for n, _ in model.named_parameters()]
    toy = np.load(os.path.join(G.REF, 'examples/toy_data.npz'))['pose_samples']
    norm = torch.load(os.path.join(G.REF, 'data/AMASS/amass_processed/version1/train/axis_normalize2.pt'))
    mean, std = norm['mean_poses'], norm['std_poses']
    data = (torch.tensor(toy[:B]).float() - mean) / std
    denorm = lambda v: v * std + mean                     # noqa: E731  (Posenormalizer.offline_denormalize, z-score)
    m = synthetic.make_body_tensors('smplx')
    bm = RefBodyModel(m)
    sde = G.sde_lib.subVPSDE(0.1, 20., N=1000)
    ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    optimizer = ref_losses.get_optimizer(cfg, model.parameters())
    state = dict(optimizer=optimizer, model=model, ema=ema, step=4000)
    grads_seen = []

    def recording_optimize_fn(optimizer, params, step, **kw):
        grads_seen.append({n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in zip(names, list(params))})
    step_fn = ref_losses.get_step_fn(sde, train=True, optimize_fn=recording_optimize_fn, reduce_mean=True, continuous=True,
                                     likelihood_weighting=False, auxiliary_loss=True, denormalize=denorm, body_model=bm,
                                     rot_rep='axis', denoise_steps=NSTEPS)
    masks_store, hook = GT.capture_masks(model)
    torch.manual_seed(900)
    rec, restore = GT.patched_draws()
    ld = step_fn(state, batch=data.clone(), condition=None, mask=None)
    restore()
    hook.remove()
    t = rec['u'] * (sde.T - 1e-5) + 1e-5
    masks = torch.stack(masks_store).reshape(NSTEPS, 5, B, 1024)
    g = grads_seen[-1]
    total = torch.sqrt(sum((v.double() ** 2).sum() for v in g.values() if v is not None)).float()
    out = {'data': data.numpy(), 't': t.numpy(), 'z': rec['z'].numpy(), 'masks': np.packbits(masks.numpy().astype(bool), axis=-1),
           'gnorm': total.numpy(), 'nsteps': np.array(NSTEPS)}
    for k in ('step_loss', 'score_loss', 'v2v_loss', 'j2j_loss'):
        out[k] = ld[k].detach().numpy()
    for n in names:
        if g[n] is None:
            continue
        flat = g[n].reshape(-1)
        out[f'g_{n}'] = flat[GT.sample_idx(flat.numel())].numpy()
        out[f'gmax_{n}'] = np.array(float(flat.abs().max()))
    # the oracle must reproduce it
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pn = T.param_names(sd)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in pn}
    full = dict(sd)
    full.update(leaves)
    body_fn = lambda pose: (lambda o: (o.v, o.Jtr))(bm(pose_body=pose))     # noqa: E731
    ol, osc, ov, oj = T.aux_loss(full, S.SubVP(0.1, 20., 1000), data, t, rec['z'], masks, cfg.model.dropout, denorm, body_fn,
                                 NSTEPS, reduce_mean=True)
    used = [k for k in pn if not k.startswith('pre_dense_cond')]
    og = dict(zip(used, torch.autograd.grad(ol, [leaves[k] for k in used])))
    for a, b, w in [(ol, ld['step_loss'], 'loss'), (osc, ld['score_loss'], 'score'), (ov, ld['v2v_loss'], 'v2v'),
                    (oj, ld['j2j_loss'], 'j2j')]:
        assert abs(float(a) - float(b)) <= 2e-5 * abs(float(b)) + 1e-12, (w, float(a), float(b))
    for n in used:
        assert (og[n] - g[n]).abs().max() <= 2e-4 * g[n].abs().max() + 1e-10, n
    np.savez_compressed(os.path.join(HERE, 'train_aux_golden.npz'), **out)
    print('wrote train_aux_golden.npz', {k: float(out[k]) for k in ('step_loss', 'score_loss', 'v2v_loss', 'j2j_loss', 'gnorm')})


if __name__ == '__main__':
    main()
