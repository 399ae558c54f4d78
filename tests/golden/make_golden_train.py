#!/usr/bin/env python
"""Golden fixtures for the training step from the REAL reference (build container only):
    python tests/golden/make_golden_train.py
Runs lib/algorithms/advanced/losses.get_step_fn(train=True) with the reference's own ScoreModelFC (train mode, dropout
0.1), torch.optim.Adam, optimization_manager (warm-up + clip) and ExponentialMovingAverage for three steps on 96
normalised AMASS poses, records every random draw (t, z, the five dropout masks per step -- captured by a forward hook on
the model's Dropout module) and stores losses, gradient statistics and parameter / EMA / Adam-moment deltas; plus the
evaluation step, the likelihood-weighted and reduce_mean=False losses and the legacy DDPM / SMLD losses.  While
generating it asserts that oracle/train_ref.py reproduces the reference."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402

from lib.algorithms.advanced import losses as ref_losses  # noqa: E402
from lib.algorithms.ema import ExponentialMovingAverage  # noqa: E402
from oracle import score_ref as S  # noqa: E402
from oracle import train_ref as T  # noqa: E402

B, STEPS, STEP0 = 96, 3, 4000
SAMPLE = 48          # entries sampled per tensor


def sample_idx(n):
    return torch.linspace(0, n - 1, min(SAMPLE, n)).long()


def capture_masks(model):
    """forward hook on the Dropout module: keep-mask = (output != 0) | (input == 0)."""
    store = []
    h = model.dropout.register_forward_hook(lambda mod, inp, out: store.append(((out != 0) | (inp[0] == 0)).to(torch.uint8)))
    return store, h


def patched_draws():
    """record torch.rand / torch.randn_like draws made by loss_fn"""
    rec = {}
    orig_rand, orig_randn_like, orig_randint = torch.rand, torch.randn_like, torch.randint

    def rand(*a, **k):
        v = orig_rand(*a, **k)
        rec['u'] = v.clone()
        return v

    def randn_like(x, **k):
        v = orig_randn_like(x, **k)
        rec['z'] = v.clone()
        return v

    def randint(*a, **k):
        v = orig_randint(*a, **k)
        rec['labels'] = v.clone()
        return v
    torch.rand, torch.randn_like, torch.randint = rand, randn_like, randint
    return rec, lambda: (setattr(torch, 'rand', orig_rand), setattr(torch, 'randn_like', orig_randn_like),
                         setattr(torch, 'randint', orig_randint))


def main():
    cfg = G.get_config()
    cfg.device = torch.device('cpu')
    model = G.build_reference_model(cfg)
    model.train()
    names = [n for n, _ in model.named_parameters()]
    toy = np.load(os.path.join(G.REF, 'examples/toy_data.npz'))['pose_samples']
    norm = torch.load(os.path.join(G.REF, 'data/AMASS/amass_processed/version1/train/axis_normalize2.pt'))
    data = (torch.tensor(toy[:B * STEPS]).float() - norm['mean_poses']) / norm['std_poses']
    sde = G.sde_lib.subVPSDE(0.1, 20., N=1000)
    osde = S.SubVP(0.1, 20., 1000)
    out = {'data': data.numpy(), 'step0': np.array(STEP0)}

    # ---- oracle state (copies of the initial weights)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    osd = {k: v.clone() for k, v in sd0.items()}
    oopt = {k: (torch.zeros_like(osd[k]), torch.zeros_like(osd[k])) for k in T.param_names(osd)}
    oopt['step'] = 0
    oema = {k: osd[k].clone() for k in T.param_names(osd)}
    oema['num_updates'] = 0
    assert T.param_names(osd) == names, (T.param_names(osd), names)

    ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    optimizer = ref_losses.get_optimizer(cfg, model.parameters())
    state = dict(optimizer=optimizer, model=model, ema=ema, step=STEP0)
    optimize_fn = ref_losses.optimization_manager(cfg)
    grads_seen = []

    def recording_optimize_fn(optimizer, params, step, **kw):
        params = list(params)
        grads_seen.append({n: (p.grad.detach().clone() if p.grad is not None else None)
                           for n, p in zip(names, params)})
        optimize_fn(optimizer, params, step, **kw)
    step_fn = ref_losses.get_step_fn(sde, train=True, optimize_fn=recording_optimize_fn, reduce_mean=cfg.training.reduce_mean,
                                     continuous=True, likelihood_weighting=False)
    masks_store, hook = capture_masks(model)
    for s in range(STEPS):
        batch = data[s * B:(s + 1) * B]
        torch.manual_seed(100 + s)
        rec, restore = patched_draws()
        del masks_store[:]
        ld = step_fn(state, batch=batch, condition=None, mask=None)
        restore()
        t = rec['u'] * (sde.T - 1e-5) + 1e-5
        masks = torch.stack(masks_store)                     # [5, B, 1024]
        g = grads_seen[-1]
        assert g['pre_dense_cond.weight'] is None
        total = torch.sqrt(sum((v.double() ** 2).sum() for v in g.values() if v is not None)).float()
        out[f's{s}_t'] = t.numpy()
        out[f's{s}_z'] = rec['z'].numpy()
        out[f's{s}_masks'] = np.packbits(masks.numpy().astype(bool), axis=-1)
        out[f's{s}_loss'] = ld['step_loss'].detach().numpy()
        out[f's{s}_gnorm'] = total.numpy()
        for n in names:
            if g[n] is None:
                continue
            flat = g[n].reshape(-1)
            out[f's{s}_g_{n}'] = flat[sample_idx(flat.numel())].numpy()
            out[f's{s}_gsum_{n}'] = np.array([flat.double().sum().item(), flat.double().abs().sum().item()])
        # the oracle must reproduce this step
        ol, og, ot = T.train_step(osd, oopt, oema, osde, batch, t, rec['z'], masks, STEP0 + s, p=cfg.model.dropout,
                                  lr=cfg.optim.lr, warmup=cfg.optim.warmup, grad_clip=cfg.optim.grad_clip,
                                  reduce_mean=cfg.training.reduce_mean, ema_rate=cfg.model.ema_rate)
        assert abs(float(ol) - float(ld['step_loss'])) <= 1e-5 * abs(float(ol)), (float(ol), float(ld['step_loss']))
        assert abs(float(ot) - float(total)) <= 1e-4 * float(total)
        for n in names:
            if g[n] is not None:
                assert (og[n] - g[n]).abs().max() <= 1e-4 * g[n].abs().max() + 1e-9, n
    hook.remove()
    sd1 = model.state_dict()
    for i, n in enumerate(names):
        d = (sd1[n] - sd0[n]).reshape(-1)
        idx = sample_idx(d.numel())
        out[f'dp_{n}'] = d[idx].numpy()
        out[f'dpnorm_{n}'] = np.array(d.double().norm().item())
        e = (ema.shadow_params[i] - sd0[n]).reshape(-1)
        out[f'dema_{n}'] = e[idx].numpy()
        st = optimizer.state.get(list(model.parameters())[i])
        if st:
            out[f'm_{n}'] = st['exp_avg'].reshape(-1)[idx].numpy()
            out[f'v_{n}'] = st['exp_avg_sq'].reshape(-1)[idx].numpy()
        od = (osd[n] - sd0[n]).reshape(-1)
        assert (od - d).abs().max() <= 2e-2 * d.abs().max() + 1e-12, (n, float((od - d).abs().max()), float(d.abs().max()))
        assert (oema[n] - ema.shadow_params[i]).abs().max() <= 2e-2 * d.abs().max() + 1e-12
        # (Adam turns an element whose gradients nearly cancel into a +-lr step: compare against the largest delta)
    out['ema_num_updates'] = np.array(ema.num_updates)

    # ---- evaluation step (EMA weights, eval mode) and loss variants on the updated model (no dropout: eval mode)
    model.eval()
    batch = data[:B]
    eval_fn = ref_losses.get_step_fn(sde, train=False, reduce_mean=True, continuous=True, likelihood_weighting=False)
    torch.manual_seed(200)
    rec, restore = patched_draws()
    ld = eval_fn(state, batch=batch)
    restore()
    out.update(eval_t=(rec['u'] * (sde.T - 1e-5) + 1e-5).numpy(), eval_z=rec['z'].numpy(), eval_loss=ld['step_loss'].numpy())
    for tag, kw in [('lw', dict(reduce_mean=True, likelihood_weighting=True)),
                    ('sum', dict(reduce_mean=False, likelihood_weighting=False))]:
        fn = ref_losses.get_sde_loss_fn(sde, train=False, continuous=True, **kw)
        torch.manual_seed(300)
        rec, restore = patched_draws()
        model.zero_grad()
        loss = fn(model, batch, None, None)
        loss.backward()
        restore()
        gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None)).float()
        out.update({f'{tag}_t': (rec['u'] * (sde.T - 1e-5) + 1e-5).numpy(), f'{tag}_z': rec['z'].numpy(),
                    f'{tag}_loss': loss.detach().numpy(), f'{tag}_gnorm': gn.numpy(),
                    f'{tag}_g_post': model.post_dense.bias.grad.numpy().copy()})
    vp, ve = G.sde_lib.VPSDE(0.1, 20., N=1000), G.sde_lib.VESDE(0.01, 50., N=1000)
    for tag, fn in [('ddpm', ref_losses.get_ddpm_loss_fn(vp, train=False, reduce_mean=True)),
                    ('smld', ref_losses.get_smld_loss_fn(ve, train=False, reduce_mean=False))]:
        torch.manual_seed(400)
        rec, restore = patched_draws()
        model.zero_grad()
        loss = fn(model, batch, None, None)
        loss.backward()
        restore()
        gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None)).float()
        out.update({f'{tag}_labels': rec['labels'].numpy(), f'{tag}_z': rec['z'].numpy(), f'{tag}_loss': loss.detach().numpy(),
                    f'{tag}_gnorm': gn.numpy(), f'{tag}_g_post': model.post_dense.bias.grad.numpy().copy()})
    np.savez_compressed(os.path.join(HERE, 'train_golden.npz'), **out)
    print('wrote train_golden.npz', len(out), 'arrays;', 'losses', [float(out[f's{s}_loss']) for s in range(STEPS)],
          'gnorm', [float(out[f's{s}_gnorm']) for s in range(STEPS)], 'eval', float(out['eval_loss']),
          'lw/sum/ddpm/smld', float(out['lw_loss']), float(out['sum_loss']), float(out['ddpm_loss']), float(out['smld_loss']))


if __name__ == '__main__':
    main()
