"""TEST INFRASTRUCTURE ONLY -- never imported by dposer_b200/.  CPU restatement (torch autograd) of one training step of the
reference: lib/algorithms/advanced/losses.py:31-57 (Adam, warm-up, clip), :61-137 (SDE loss), :140-184 (SMLD / DDPM
losses), :187-275 (step_fn), model.py:141-196 in train mode (dropout masks passed in), lib/algorithms/ema.py:35-50.
Pinned against the REAL reference by tests/golden/make_golden_train.py -> train_golden.npz (tests/test_oracle_golden.py).
"""
import math

import torch
import torch.nn.functional as F

from . import score_ref as S

PARAM_ORDER = None  # filled by param_names()


def param_names(sd):
    """model.parameters() order of the reference's ScoreModelFC (registration order, model.py:100-139)."""
    names = ['pre_dense', 'pre_dense_t', 'pre_dense_cond', 'pre_gnorm', 'shared_time_embed.0']
    for b in (1, 2):
        names += [f'b{b}_dense1', f'b{b}_dense1_t', f'b{b}_gnorm1', f'b{b}_dense2', f'b{b}_dense2_t', f'b{b}_gnorm2']
    names += ['post_dense']
    out = []
    for n in names:
        out += [n + '.weight', n + '.bias']
    return [k for k in out if k in sd]


def forward_train(sd, x, labels, masks=None, p=0.0, scale_by_sigma=True):
    """model.py:141-196 with dropout(h) = h * mask / (1 - p) after each of the five activations (masks [5,B,1024] 0/1)."""
    def drop(h, i):
        return h if masks is None or p <= 0 else h * masks[i].to(h.dtype) / (1.0 - p)
    temb = F.silu(S._lin(sd, 'shared_time_embed.0', S.timestep_embedding(labels)))
    h = S._lin(sd, 'pre_dense', x) + S._lin(sd, 'pre_dense_t', temb)
    h = drop(F.silu(S._gn(sd, 'pre_gnorm', h)), 0)
    i = 1
    for b in (1, 2):
        h1 = S._lin(sd, f'b{b}_dense1', h) + S._lin(sd, f'b{b}_dense1_t', temb)
        h1 = drop(F.silu(S._gn(sd, f'b{b}_gnorm1', h1)), i)
        h2 = S._lin(sd, f'b{b}_dense2', h1) + S._lin(sd, f'b{b}_dense2_t', temb)
        h2 = drop(F.silu(S._gn(sd, f'b{b}_gnorm2', h2)), i + 1)
        h = h + h2
        i += 2
    res = S._lin(sd, 'post_dense', h)
    if scale_by_sigma:
        res = res / sd['sigmas'][labels.long()].reshape(-1, 1)
    return res


def sde_loss(sd, sde, batch, t, z, masks=None, p=0.0, reduce_mean=True, likelihood_weighting=False):
    """losses.py:108-131 for the sub-VP SDE (sde: oracle SubVP): score = -model(x_t, 999 t) / std."""
    mean, std = sde.marginal(batch, t)
    xt = mean + std[:, None] * z
    score = -forward_train(sd, xt, t * 999, masks, p) / std[:, None]
    red = (lambda v: v.mean(dim=-1)) if reduce_mean else (lambda v: 0.5 * v.sum(dim=-1))
    if not likelihood_weighting:
        losses = red(torch.square(score * std[:, None] + z))
    else:
        g2 = sde.sde(torch.zeros_like(batch), t)[1] ** 2
        losses = red(torch.square(score + z / std[:, None])) * g2
    return losses.mean()


def aux_loss(sd, sde, batch, t, z, masks, p, denormalize, body_fn, n_steps, reduce_mean=True):
    """losses.py:91-121,244-258: score loss on the first evaluation + log(1 + SNR)-weighted v2v / j2j terms on the estimate
    of an ``n_steps`` DDIM chain (t -> t / (2 n_steps)); masks [n_steps, 5, B, 1024]; body_fn(pose) -> (verts, joints)."""
    mean, std = sde.marginal(batch, t)
    x = mean + std[:, None] * z
    alpha0, sigma0 = sde.alpha_sigma(t)
    snr = alpha0 / sigma0[:, None]
    lin = torch.linspace(0, 1, n_steps + 1)[:, None]
    traj = (1 - lin) * t + lin * (t / (2 * n_steps))
    score0 = None
    for i in range(n_steps):
        tc, tb = traj[i], traj[i + 1]
        a_c, s_c = sde.alpha_sigma(tc)
        a_b, s_b = sde.alpha_sigma(tb)
        score = -forward_train(sd, x, tc * 999, None if masks is None else masks[i], p) / sde.marginal(x, tc)[1][:, None]
        if i == 0:
            score0 = score
        noise = -score * s_c[:, None]
        x = a_b / a_c * (x - s_c[:, None] * noise) + s_b[:, None] * noise
    red = (lambda v: v.mean(dim=-1)) if reduce_mean else (lambda v: 0.5 * v.sum(dim=-1))
    score_loss = red(torch.square(score0 * std[:, None] + z)).mean()
    weight = torch.log(1.0 + snr)
    gv, gj = body_fn(denormalize(batch))
    pv, pj = body_fn(denormalize(x))
    v2v = torch.mean(weight * ((gv - pv) ** 2).sum(dim=-1))
    j2j = torch.mean(weight * ((gj - pj) ** 2).sum(dim=-1))
    return score_loss + v2v + j2j, score_loss, v2v, j2j


def clip_coef(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_: max_norm / (||g|| + 1e-6), clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    return torch.clamp(max_norm / (total + 1e-6), max=1.0), total


def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (amsgrad off, weight decay 0); in place."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def ema_decay(decay, num_updates):
    return min(decay, (1 + num_updates) / (10 + num_updates))


def train_step(sd, opt, ema, sde, batch, t, z, masks, step, p=0.1, lr=2e-4, warmup=5000, grad_clip=1.0, reduce_mean=True,
               likelihood_weighting=False, ema_rate=0.9999):
    """One losses.py:234-262 step.  sd: dict of leaf tensors (updated in place); opt: dict name -> (m, v) + 'step';
    ema: dict name -> shadow + 'num_updates'.  Returns (loss, grads by name, total grad norm)."""
    names = param_names(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    loss = sde_loss(full, sde, batch, t, z, masks, p, reduce_mean, likelihood_weighting)
    used = [k for k in names if not k.startswith('pre_dense_cond')]
    grads = dict(zip(used, torch.autograd.grad(loss, [leaves[k] for k in used])))
    lr_t = lr * min(step / warmup, 1.0) if warmup > 0 else lr
    coef, total = clip_coef(list(grads.values()), grad_clip) if grad_clip >= 0 else (1.0, None)
    opt['step'] += 1
    with torch.no_grad():
        for k in used:
            m, v = opt[k]
            adam_update(sd[k], grads[k] * coef, m, v, opt['step'], lr_t)
        ema['num_updates'] += 1
        omd = 1.0 - ema_decay(ema_rate, ema['num_updates'])
        for k in names:
            ema[k].sub_(omd * (ema[k] - sd[k]))
    return loss.detach(), grads, total
