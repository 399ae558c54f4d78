"""Oracle: score network, subVP-SDE scalars, score_fn, PC sampler, prior loss.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain torch-CPU fp32, written
functionally over a ``state_dict`` so it does not depend on the product classes.
Pinned against the real reference by tests/golden/make_golden.py.

Reference anchors (relative to the reference checkout):
  lib/algorithms/advanced/model.py:24-51,141-196   sigmas, embedding, ScoreModelFC.forward
  lib/algorithms/advanced/sde_lib.py:184-231,75-119 subVPSDE, reverse SDE
  lib/algorithms/advanced/utils.py:127-163          get_score_fn
  lib/algorithms/advanced/sampling.py:177-188,273-302,375-468   EM, Langevin, pc_sampler
  run/completion.py:105-149, run/motion_denoising.py:99-143, run/smplify.py:69-107  prior loss
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

HIDDEN, EMBED, POSE_D, N_BLOCKS, GROUPS = 1024, 512, 63, 2, 32


# ----------------------------------------------------------------------------- net
def sigma_table(sigma_min=0.01, sigma_max=50.0, num_scales=1000):
    """model.py:24-34 -- geometric table from sigma_max DOWN to sigma_min (fp64 -> fp32)."""
    tab = np.exp(np.linspace(np.log(sigma_max), np.log(sigma_min), num_scales))
    return torch.tensor(tab, dtype=torch.float)


def timestep_embedding(labels, dim=EMBED, max_positions=10000):
    """model.py:37-51 -- [sin(l*f_k) | cos(l*f_k)], f_k = exp(-k ln(1e4)/(dim/2-1))."""
    half = dim // 2
    scale = math.log(max_positions) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -scale)
    arg = labels.float()[:, None] * freqs[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def _lin(sd, name, x):
    return F.linear(x, sd[name + '.weight'], sd[name + '.bias'])


def _gn(sd, name, x):
    return F.group_norm(x, GROUPS, sd[name + '.weight'], sd[name + '.bias'], eps=1e-5)


def score_model_forward(sd, x, labels, scale_by_sigma=True, n_blocks=N_BLOCKS):
    """model.py:141-196 (positional embedding, swish, eval-mode dropout = identity)."""
    temb = F.silu(_lin(sd, 'shared_time_embed.0', timestep_embedding(labels)))
    h = _lin(sd, 'pre_dense', x) + _lin(sd, 'pre_dense_t', temb)
    h = F.silu(_gn(sd, 'pre_gnorm', h))
    for b in range(1, n_blocks + 1):
        h1 = _lin(sd, f'b{b}_dense1', h) + _lin(sd, f'b{b}_dense1_t', temb)
        h1 = F.silu(_gn(sd, f'b{b}_gnorm1', h1))
        h2 = _lin(sd, f'b{b}_dense2', h1) + _lin(sd, f'b{b}_dense2_t', temb)
        h2 = F.silu(_gn(sd, f'b{b}_gnorm2', h2))
        h = h + h2
    res = _lin(sd, 'post_dense', h)
    if scale_by_sigma:
        res = res / sd['sigmas'][labels.long()].reshape(-1, 1)   # model.py:159,192-194
    return res


# ----------------------------------------------------------------------------- SDE
class SubVP:
    """sde_lib.py:184-231 scalar schedules (all fp32 torch ops, as the reference)."""

    def __init__(self, beta_min=0.1, beta_max=20.0, N=1000, T=1.0):
        self.b0, self.b1, self.N, self.T = beta_min, beta_max, N, T
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas

    def beta(self, t):
        return self.b0 + t * (self.b1 - self.b0)

    def lmc(self, t):                                    # log mean coefficient
        return -0.25 * t ** 2 * (self.b1 - self.b0) - 0.5 * t * self.b0

    def sde(self, x, t):                                 # :206-211
        bt = self.beta(t)
        drift = -0.5 * bt[:, None] * x
        discount = 1.0 - torch.exp(-2 * self.b0 * t - (self.b1 - self.b0) * t ** 2)
        return drift, torch.sqrt(bt * discount)

    def marginal(self, x, t):                            # :213-217  (std is NOT a sqrt)
        c = self.lmc(t)
        return torch.exp(c)[:, None] * x, 1 - torch.exp(2.0 * c)

    def alpha_sigma(self, t):                            # :227-231
        c = self.lmc(t)
        return torch.exp(c[:, None]), 1.0 - torch.exp(2.0 * c)


def score_fn(sd, sde, x, t):
    """utils.py:141-163 -- labels = t*999, score = -net/std(t)."""
    labels = t * 999
    out = score_model_forward(sd, x, labels)
    std = sde.marginal(torch.zeros_like(x), t)[1]
    return -out / std[:, None]


def reverse_drift(sd, sde, x, t, probability_flow=False):
    """sde_lib.py:98-106."""
    drift, g = sde.sde(x, t)
    s = score_fn(sd, sde, x, t)
    drift = drift - g[:, None] ** 2 * s * (0.5 if probability_flow else 1.0)
    if probability_flow:
        g = torch.zeros(1)
    return drift, g, s


# ----------------------------------------------------------------------------- sampler
def em_step(sd, sde, x, t, z, probability_flow=False):
    """sampling.py:182-188 with the Gaussian draw ``z`` supplied by the caller."""
    dt = -1.0 / sde.N
    drift, g, _ = reverse_drift(sd, sde, x, t, probability_flow)
    x_mean = x + drift * dt
    x = x_mean + g[:, None] * np.sqrt(-dt) * z
    return x, x_mean


def langevin_step(sd, sde, x, t, noise, snr=0.16):
    """sampling.py:282-302 (n_steps_each = 1); batch-global norm means (:296-297)."""
    timestep = (t * (sde.N - 1) / sde.T).long()
    alpha = sde.alphas[timestep]
    grad = score_fn(sd, sde, x, t)
    grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
    noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
    step = (snr * noise_norm / grad_norm) ** 2 * 2 * alpha
    x_mean = x + step[:, None] * grad
    x = x_mean + torch.sqrt(step * 2)[:, None] * noise
    return x, x_mean


def pc_sample(sd, sde, x_init, eps=1e-3, noise=None, corrector='none', snr=0.16,
              observation=None, mask=None, task=None, start_step=0,
              probability_flow=False, denoise=True, n_run=None, keep_traj=False):
    """sampling.py:429-466 with every Gaussian draw injected.

    noise[i] is a dict of [B,D] tensors for step i:
      'pred' (EM draw), 'corr' (Langevin draw), 'imp_c' / 'imp_p' (imputation draws
      after corrector / predictor, sampling.py:413-422).  Missing keys -> zeros.
    n_run limits the number of executed steps (for short parity runs on the same grid).
    """
    B, D = x_init.shape
    x = x_init.clone()
    timesteps = torch.linspace(sde.T, eps, sde.N)
    start = start_step if task == 'denoise' else 0
    stop = sde.N if n_run is None else min(sde.N, start + n_run)
    traj, x_mean = [], x
    zero = torch.zeros(B, D)

    def impute(xx, vec_t, z):
        if task != 'completion':
            return xx
        mean, std = sde.marginal(observation, vec_t)
        return xx * (1 - mask) + (mean + z * std[:, None]) * mask

    for i in range(start, stop):
        nz = noise[i] if noise is not None else {}
        vec_t = torch.ones(B) * timesteps[i]
        if corrector == 'langevin':
            x, x_mean = langevin_step(sd, sde, x, vec_t, nz.get('corr', zero), snr)
        x = impute(x, vec_t, nz.get('imp_c', zero))
        x, x_mean = em_step(sd, sde, x, vec_t, nz.get('pred', zero), probability_flow)
        x = impute(x, vec_t, nz.get('imp_p', zero))
        if keep_traj:
            traj.append(x.clone())
    out = x_mean if denoise else x
    return (torch.stack(traj) if keep_traj else None), out


# ----------------------------------------------------------------------------- prior loss
def one_step_denoise(sd, sde, x_t, t):
    """run/completion.py:105-110 -- x0_hat = (x_t + std^2 * score)/alpha, SNR = alpha/std."""
    alpha, sigma = sde.alpha_sigma(t)
    s = score_fn(sd, sde, x_t, t)
    x0_hat = (x_t + (sigma ** 2)[:, None] * s) / alpha
    snr = alpha / torch.sqrt(sigma ** 2)[:, None]
    return x0_hat.detach(), snr


def multi_step_denoise(sd, sde, x_t, t, t_end, n=10):
    """run/completion.py:112-129 -- DDIM from t to t_end in n uniform sub-steps."""
    lam = torch.linspace(0, 1, n + 1)[:, None]
    traj = (1 - lam) * t + lam * t_end                  # lib/utils/misc.py:58-61
    cur = x_t
    for i in range(n):
        a_c, s_c = sde.alpha_sigma(traj[i])
        a_b, s_b = sde.alpha_sigma(traj[i + 1])
        eps_hat = -score_fn(sd, sde, cur, traj[i]) * s_c[:, None]
        cur = a_b / a_c * (cur - s_c[:, None] * eps_hat) + s_b[:, None] * eps_hat
    alpha, sigma = sde.alpha_sigma(traj[0])
    return cur.detach(), alpha / sigma[:, None]


def prior_loss(sd, sde, x0, t, z, weighted, reduce, divisor=None, multi_denoise=False, ddim_n=10):
    """Appendix A.5 of SURVEY.md.

    reduce='mean'  -> torch.mean(w*(x0-x0_hat)^2)            (run/completion.py:147)
    reduce='sum'   -> torch.sum(w*(x0-x0_hat)^2)/divisor     (run/motion_denoising.py:141, run/smplify.py:105)
    Returns (loss, d loss / d x0) -- x0_hat is detached, so grad = 2 w (x0-x0_hat)/div.
    """
    x0 = x0.detach().clone().requires_grad_(True)
    mean, std = sde.marginal(x0, t)
    x_t = mean + std[:, None] * z
    with torch.no_grad():
        if multi_denoise:
            x0_hat, snr = multi_step_denoise(sd, sde, x_t.detach(), t, t / (2 * ddim_n), ddim_n)
        else:
            x0_hat, snr = one_step_denoise(sd, sde, x_t.detach(), t)
    w = 0.5 * torch.sqrt(1 + snr) if weighted else 0.5
    sq = w * (x0 - x0_hat) ** 2
    loss = sq.mean() if reduce == 'mean' else sq.sum() / divisor
    loss.backward()
    return loss.detach(), x0.grad.detach()


def quan_t_schedule(N, total_steps, trun, offset):
    """Strategy '3' (run/completion.py:189-190, run/motion_denoising.py:245, run/smplify.py:153-166):
    python math.floor of a torch scalar product -- reproduced verbatim for bit-exact ints."""
    out = []
    for step in range(total_steps):
        out.append(N - math.floor(torch.tensor(total_steps - step - 1) * (N / (trun * total_steps))) - offset)
    return out


# ----------------------------------------------------------------------------- synthetic weights
def make_state_dict(seed=42):
    """SURVEY 8(d) "Score weights": default nn.Linear/GroupNorm init in the reference's module
    registration order under torch.manual_seed(seed), then GroupNorm affine randomised
    (weight~U(0.5,1.5), bias~N(0,0.1^2)) in module order with the same generator."""
    import torch.nn as nn
    torch.manual_seed(seed)
    mods = {}
    mods['pre_dense'] = nn.Linear(POSE_D, HIDDEN)
    mods['pre_dense_t'] = nn.Linear(EMBED, HIDDEN)
    mods['pre_dense_cond'] = nn.Linear(HIDDEN, HIDDEN)
    mods['pre_gnorm'] = nn.GroupNorm(GROUPS, HIDDEN)
    mods['shared_time_embed.0'] = nn.Linear(EMBED, EMBED)
    for b in range(1, N_BLOCKS + 1):
        mods[f'b{b}_dense1'] = nn.Linear(HIDDEN, HIDDEN)
        mods[f'b{b}_dense1_t'] = nn.Linear(EMBED, HIDDEN)
        mods[f'b{b}_gnorm1'] = nn.GroupNorm(GROUPS, HIDDEN)
        mods[f'b{b}_dense2'] = nn.Linear(HIDDEN, HIDDEN)
        mods[f'b{b}_dense2_t'] = nn.Linear(EMBED, HIDDEN)
        mods[f'b{b}_gnorm2'] = nn.GroupNorm(GROUPS, HIDDEN)
    mods['post_dense'] = nn.Linear(HIDDEN, POSE_D)
    for name, m in mods.items():
        if isinstance(m, nn.GroupNorm):
            with torch.no_grad():
                m.weight.copy_(torch.rand(HIDDEN) + 0.5)
                m.bias.copy_(torch.randn(HIDDEN) * 0.1)
    sd = {}
    for name, m in mods.items():
        sd[name + '.weight'] = m.weight.detach().clone()
        sd[name + '.bias'] = m.bias.detach().clone()
    sd['sigmas'] = sigma_table()
    return sd
