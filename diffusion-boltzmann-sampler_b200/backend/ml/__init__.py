"""Hot-path ML package: same import paths and class/method signatures as the reference's
backend/ml (ml/__init__.py:16-17 re-exports), backed by libdbs_b200.so."""
from .samplers.mcmc import MetropolisHastings
from .samplers.diffusion import DiffusionSampler, PretrainedDiffusionSampler
from .analysis import pair_correlation

__all__ = ["MetropolisHastings", "DiffusionSampler", "PretrainedDiffusionSampler", "pair_correlation"]
