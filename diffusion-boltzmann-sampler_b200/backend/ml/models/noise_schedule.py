"""beta(t) schedules of the VP-SDE — host-side scalar code that feeds the sampler's per-step sigma table.

Same public surface as the reference's models/noise_schedule.py (`NoiseSchedule` :20-74, `LinearNoiseSchedule`
:77-103, `CosineNoiseSchedule` :106-151, `SigmoidNoiseSchedule` :154-187, `get_schedule` :190-227).  A schedule
plugs into `DiffusionProcess(schedule=...)`; `DiffusionProcess.sigma_table` evaluates it once per step on a
1-element fp32 tensor exactly as the reference sampler does (samplers/diffusion.py:223-227), so the fp32
coefficients handed to the Euler-Maruyama kernel are the reference's.  Nothing here runs on the GPU: it is
O(num_steps) scalar work per sampling call.
"""
import math
from typing import Tuple

import torch

_TRAPEZOID_PANELS = 100      # the reference's fixed-panel trapezoid rule for schedules without a closed form (:58-74)


class NoiseSchedule:
    """t in [0, 1] -> beta(t); alpha_t = exp(-0.5 * int_0^t beta), sigma_t = sqrt(max(1 - alpha_t^2, 1e-8))."""

    def beta(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def _integrate_beta(self, t: torch.Tensor) -> torch.Tensor:
        # 100 trapezoid panels on [0, t], summed in panel order (fp32, like the reference)
        h = t / _TRAPEZOID_PANELS
        total = torch.zeros_like(t)
        for k in range(_TRAPEZOID_PANELS):
            total = total + 0.5 * (self.beta(k * h) + self.beta((k + 1) * h)) * h
        return total

    def noise_level(self, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        alpha = torch.exp(-0.5 * self._integrate_beta(t))
        return alpha, torch.sqrt(torch.clamp(1 - alpha ** 2, min=1e-8))


class LinearNoiseSchedule(NoiseSchedule):
    """beta(t) = beta_min + t * (beta_max - beta_min), integral in closed form."""

    def __init__(self, beta_min: float = 0.1, beta_max: float = 20.0):
        self.beta_min, self.beta_max = beta_min, beta_max

    def beta(self, t: torch.Tensor) -> torch.Tensor:
        return self.beta_min + t * (self.beta_max - self.beta_min)

    def _integrate_beta(self, t: torch.Tensor) -> torch.Tensor:
        return self.beta_min * t + 0.5 * (self.beta_max - self.beta_min) * t ** 2


class CosineNoiseSchedule(NoiseSchedule):
    """alpha_t^2 = f(t) / f(0), f(t) = cos^2(pi/2 * (t + s) / (1 + s)); beta by a forward difference of log alpha."""

    def __init__(self, s: float = 0.008):
        self.s = s
        self._alpha_0 = self._f(torch.tensor(0.0)).item()

    def _f(self, t: torch.Tensor) -> torch.Tensor:
        return torch.cos(math.pi / 2 * (t + self.s) / (1 + self.s)) ** 2

    def noise_level(self, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        a2 = self._f(t) / self._alpha_0
        return torch.sqrt(torch.clamp(a2, min=1e-8, max=1.0)), torch.sqrt(torch.clamp(1 - a2, min=1e-8))

    def beta(self, t: torch.Tensor) -> torch.Tensor:
        step = 1e-4
        a0, _ = self.noise_level(t)
        a1, _ = self.noise_level(t + step)
        return torch.clamp(-2 * (torch.log(a1 + 1e-8) - torch.log(a0 + 1e-8)) / step, min=0.0, max=100.0)


class SigmoidNoiseSchedule(NoiseSchedule):
    """beta(t) = beta_min + (beta_max - beta_min) * sigmoid(start + (end - start) * t); integral by the trapezoid rule."""

    def __init__(self, beta_min: float = 0.1, beta_max: float = 20.0, start: float = -6.0, end: float = 6.0):
        self.beta_min, self.beta_max, self.start, self.end = beta_min, beta_max, start, end

    def beta(self, t: torch.Tensor) -> torch.Tensor:
        return self.beta_min + (self.beta_max - self.beta_min) * torch.sigmoid(self.start + (self.end - self.start) * t)


def get_schedule(schedule_type: str = "linear", beta_min: float = 0.1, beta_max: float = 20.0, **kwargs) -> NoiseSchedule:
    kind = schedule_type.lower()
    if kind == "linear":
        return LinearNoiseSchedule(beta_min=beta_min, beta_max=beta_max)
    if kind == "cosine":
        return CosineNoiseSchedule(s=kwargs.get("s", 0.008))
    if kind == "sigmoid":
        return SigmoidNoiseSchedule(beta_min=beta_min, beta_max=beta_max, start=kwargs.get("start", -6.0),
                                    end=kwargs.get("end", 6.0))
    raise ValueError(f"Unknown schedule type: {schedule_type}. Choose from: linear, cosine, sigmoid")
