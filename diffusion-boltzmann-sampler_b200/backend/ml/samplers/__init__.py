"""Samplers of the hot path (MCMC, diffusion, Langevin); names as the reference's samplers package, plus the pretrained wrapper."""
from . import diffusion as _diffusion, langevin as _langevin, mcmc as _mcmc

MetropolisHastings = _mcmc.MetropolisHastings
DiffusionSampler = _diffusion.DiffusionSampler
PretrainedDiffusionSampler = _diffusion.PretrainedDiffusionSampler
LangevinDynamics = _langevin.LangevinDynamics
UnderdampedLangevin = _langevin.UnderdampedLangevin

__all__ = ("MetropolisHastings", "DiffusionSampler", "PretrainedDiffusionSampler", "LangevinDynamics", "UnderdampedLangevin")
