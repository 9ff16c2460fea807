"""Hot-path ML package: same import paths and class/method signatures as the reference's
backend/ml (ml/__init__.py:16-36 re-exports samplers and the Ising analysis), backed by libdbs_b200.so."""
from . import analysis as _analysis, samplers as _samplers
from .analysis.correlation import *  # noqa: F401,F403  (the fourteen Ising analysis functions)

MetropolisHastings = _samplers.MetropolisHastings
DiffusionSampler = _samplers.DiffusionSampler
PretrainedDiffusionSampler = _samplers.PretrainedDiffusionSampler

__all__ = ("MetropolisHastings", "DiffusionSampler", "PretrainedDiffusionSampler") + tuple(_analysis.correlation.__all__)
