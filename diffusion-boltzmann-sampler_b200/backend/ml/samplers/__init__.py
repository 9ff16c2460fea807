from .mcmc import MetropolisHastings
from .diffusion import DiffusionSampler, PretrainedDiffusionSampler

__all__ = ["MetropolisHastings", "DiffusionSampler", "PretrainedDiffusionSampler"]
