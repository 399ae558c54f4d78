"""SDE objects with the reference's call surface (lib/algorithms/advanced/sde_lib.py).

``subVPSDE`` is the shipped configuration; ``VPSDE`` / ``VESDE`` are kept as thin host-side
variants (same kernels, different scalar schedules).  Everything here is scalar-per-row
schedule arithmetic in fp32 torch ops -- the heavy work is the score evaluation, which the
``score_fn`` handed to ``reverse()`` runs on the GPU kernels.
"""
import abc

import numpy as np
import torch


class SDE(abc.ABC):
    """Forward SDE  dx = f(x,t) dt + g(t) dw  (sde_lib.py:7-73)."""

    def __init__(self, N):
        super().__init__()
        self.N = N

    @property
    @abc.abstractmethod
    def T(self):
        ...

    @abc.abstractmethod
    def sde(self, x, t):
        ...

    @abc.abstractmethod
    def marginal_prob(self, x, t):
        ...

    @abc.abstractmethod
    def prior_sampling(self, shape):
        ...

    @abc.abstractmethod
    def prior_logp(self, z):
        ...

    @abc.abstractmethod
    def return_alpha_sigma(self, t):
        ...

    def discretize(self, x, t):
        """Euler-Maruyama discretisation f*dt, g*sqrt(dt) (sde_lib.py:51-69)."""
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t)
        return drift * dt, diffusion * torch.sqrt(torch.tensor(dt, device=t.device))

    def reverse(self, score_fn, probability_flow=False):
        """Reverse-time SDE / probability-flow ODE object (sde_lib.py:75-119)."""
        return _Reverse(self, score_fn, probability_flow)


class _Reverse:
    """What the reference builds as the local class ``RSDE`` on every ``reverse()`` call."""

    def __init__(self, fwd, score_fn, probability_flow):
        self._fwd = fwd
        self._score_fn = score_fn
        self.N = fwd.N
        self.probability_flow = probability_flow

    @property
    def T(self):
        return self._fwd.T

    def sde(self, x, t, condition=None, mask=None, guide=False):
        drift, diffusion = self._fwd.sde(x, t)
        score = self._score_fn(x, t, condition, mask)
        drift = drift - diffusion[:, None] ** 2 * score * (0.5 if self.probability_flow else 1.)
        if self.probability_flow:
            diffusion = torch.zeros(1, device=drift.device)
        if not guide:
            return drift, diffusion
        alpha, sigma = self._fwd.return_alpha_sigma(t)
        return drift, diffusion, alpha, sigma ** 2, score

    def discretize(self, x, t, condition=None, mask=None):
        f, G = self._fwd.discretize(x, t)
        rev_f = f - G[:, None] ** 2 * self._score_fn(x, t, condition, mask)
        rev_G = torch.zeros_like(G) if self.probability_flow else G
        return rev_f, rev_G

    def __getattr__(self, name):          # marginal_prob etc. fall through to the forward SDE
        return getattr(self._fwd, name)


class VPSDE(SDE):
    """sde_lib.py:122-181."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000, T=1):
        super().__init__(N)
        self.beta_0, self.beta_1, self.N, self._T = beta_min, beta_max, N, T
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)

    @property
    def T(self):
        return self._T

    def _lmc(self, t):
        return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        return -0.5 * beta_t[:, None] * x, torch.sqrt(beta_t)

    def marginal_prob(self, x, t):
        c = self._lmc(t)
        return torch.exp(c[:, None]) * x, torch.sqrt(1. - torch.exp(2. * c))

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2. * np.log(2 * np.pi) - torch.sum(z ** 2, dim=1) / 2.

    def return_alpha_sigma(self, t):
        c = self._lmc(t)
        return torch.exp(c[:, None]), torch.sqrt(1. - torch.exp(2. * c))

    def discretize(self, x, t):
        timestep = (t * (self.N - 1) / self.T).long()
        beta = self.discrete_betas.to(x.device)[timestep]
        alpha = self.alphas.to(x.device)[timestep]
        return torch.sqrt(alpha)[:, None] * x - x, torch.sqrt(beta)


class subVPSDE(SDE):
    """sde_lib.py:184-231.  NB ``std = 1 - exp(2 lmc)`` -- no square root in the sub-VP marginal."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000, T=1):
        super().__init__(N)
        self.beta_0, self.beta_1, self.N, self._T = beta_min, beta_max, N, T
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas

    @property
    def T(self):
        return self._T

    def _lmc(self, t):
        return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        discount = 1. - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
        return -0.5 * beta_t[:, None] * x, torch.sqrt(beta_t * discount)

    def marginal_prob(self, x, t):
        c = self._lmc(t)
        return torch.exp(c)[:, None] * x, 1 - torch.exp(2. * c)

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2. * np.log(2 * np.pi) - torch.sum(z ** 2, dim=1) / 2.

    def return_alpha_sigma(self, t):
        c = self._lmc(t)
        return torch.exp(c[:, None]), 1. - torch.exp(2. * c)


class VESDE(SDE):
    """sde_lib.py:234-291."""

    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000, T=1):
        super().__init__(N)
        self.sigma_min, self.sigma_max, self.N, self._T = sigma_min, sigma_max, N, T
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

    @property
    def T(self):
        return self._T

    def sde(self, x, t):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        diffusion = sigma * torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)),
                                                    device=t.device))
        return torch.zeros_like(x), diffusion

    def marginal_prob(self, x, t):
        return x, self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def prior_sampling(self, shape):
        return torch.randn(*shape) * self.sigma_max

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2. * np.log(2 * np.pi * self.sigma_max ** 2) - torch.sum(z ** 2, dim=1) / (2 * self.sigma_max ** 2)

    def return_alpha_sigma(self, t):
        return torch.ones_like(t)[:, None], self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def discretize(self, x, t):
        timestep = (t * (self.N - 1) / self.T).long()
        sigmas = self.discrete_sigmas.to(t.device)
        sigma = sigmas[timestep]
        adjacent = torch.where(timestep == 0, torch.zeros_like(t), sigmas[timestep - 1])
        return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adjacent ** 2)
