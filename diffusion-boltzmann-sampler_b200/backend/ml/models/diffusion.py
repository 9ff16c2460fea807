"""VP-SDE schedule — host-side scalar code behind the sampler's per-step sigma table.

Replacement for the parts of models/diffusion.py the sampling path uses (`noise_level` :68-97,
`beta` :52-66, `get_config` :40-46), with the training helpers (`forward`, `score_target`,
`reverse_step`, `compute_loss`) kept so the module still satisfies its callers.  This is O(num_steps) scalar
work per sampling call; the engine consumes the resulting fp32 coefficients.
"""
from typing import Optional, Tuple

import torch

EPS = 1e-8
MIN_SIGMA = 1e-4
MAX_BETA = 500.0


class DiffusionProcess:
    """dx = -0.5 beta(t) x dt + sqrt(beta(t)) dW, beta linear in t unless a schedule object
    (anything with .beta(t) and .noise_level(t)) is supplied."""

    def __init__(self, beta_min: float = 0.1, beta_max: float = 20.0, schedule=None):
        self.beta_min = beta_min
        self.beta_max = beta_max
        self._schedule = schedule

    def get_config(self) -> dict:
        return {"beta_min": self.beta_min, "beta_max": self.beta_max}

    def _clamp_time(self, t: torch.Tensor) -> torch.Tensor:
        return torch.clamp(t, min=0.0, max=1.0)

    def beta(self, t: torch.Tensor) -> torch.Tensor:
        t = self._clamp_time(t)
        b = self._schedule.beta(t) if self._schedule is not None else self.beta_min + t * (self.beta_max - self.beta_min)
        return torch.clamp(b, min=0.0, max=MAX_BETA)

    def noise_level(self, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(alpha_t, sigma_t): alpha = exp(-0.5 * int_0^t beta), sigma = sqrt(max(1 - alpha^2, EPS)),
        then sigma >= MIN_SIGMA — the same fp32 operation order as models/diffusion.py:80-97 so the
        sampler's coefficients are bit-identical to the reference's."""
        t = self._clamp_time(t)
        if self._schedule is not None:
            alpha_t, sigma_t = self._schedule.noise_level(t)
        else:
            integral = self.beta_min * t + 0.5 * (self.beta_max - self.beta_min) * t ** 2
            alpha_t = torch.exp(-0.5 * integral)
            sigma_t = torch.sqrt(torch.clamp(1 - alpha_t ** 2, min=EPS))
        return alpha_t, torch.clamp(sigma_t, min=MIN_SIGMA)

    def sigma_table(self, num_steps: int):
        """Python-float sigma for i = num_steps .. 1 at t = i/num_steps, each evaluated on a
        1-element fp32 tensor exactly as samplers/diffusion.py:223-227 does."""
        out = []
        for i in range(num_steps, 0, -1):
            _, s = self.noise_level(torch.full((1,), i / num_steps))
            out.append(s[0].item())
        return out

    # ---- training-side helpers (not on the sampling path) --------------------------------------
    def forward(self, x_0: torch.Tensor, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """x_t = alpha_t x_0 + sigma_t eps (models/diffusion.py:99-123)."""
        a, s = self.noise_level(t)
        shape = (-1,) + (1,) * (x_0.dim() - 1)
        eps = torch.randn_like(x_0)
        return a.view(shape) * x_0 + s.view(shape) * eps, eps

    def score_target(self, noise: torch.Tensor, sigma_t: torch.Tensor) -> torch.Tensor:
        return -noise / (sigma_t + EPS)

    def reverse_step(self, x_t, score, t, dt: float, stochastic: bool = True):
        """One reverse-SDE step in beta form (models/diffusion.py:139-179)."""
        b = self.beta(t)
        if b.dim() == 1:
            b = b.view((-1,) + (1,) * (x_t.dim() - 1))
        x = x_t + (-0.5 * b * x_t - b * score) * dt
        if stochastic:
            x = x + torch.sqrt(b * abs(dt)) * torch.randn_like(x_t)
        return x

    def compute_loss(self, score_pred: torch.Tensor, noise: torch.Tensor, t: torch.Tensor, loss_type: str = "l2",
                     weighting: str = "uniform") -> torch.Tensor:
        """Denoising score-matching loss against the target -eps / (sigma_t + EPS): squared ("l2") or absolute
        ("l1") error, weighted by 1 ("uniform") or sigma_t^2 ("snr"), mean over every element
        (models/diffusion.py:181-227).  Differentiable torch code for the trainer; not on the sampling path."""
        sigma = self.noise_level(t)[1].view((-1,) + (1,) * (score_pred.dim() - 1))
        err = score_pred - self.score_target(noise, sigma)
        if loss_type == "l2":
            err = err.pow(2)
        elif loss_type == "l1":
            err = err.abs()
        else:
            raise ValueError(f"Unknown loss type: {loss_type}")
        if weighting == "snr":
            err = err * sigma.pow(2)
        elif weighting != "uniform":
            raise ValueError(f"Unknown weighting: {weighting}")
        return err.mean()
