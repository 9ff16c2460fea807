from .mcmc import MetropolisHastings
from .diffusion import DiffusionSampler, PretrainedDiffusionSampler
from .langevin import LangevinDynamics, UnderdampedLangevin

__all__ = ["MetropolisHastings", "DiffusionSampler", "PretrainedDiffusionSampler", "LangevinDynamics",
           "UnderdampedLangevin"]
