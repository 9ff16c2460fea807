"""Hot-path ML package: same import paths and class/method signatures as the reference's
backend/ml (ml/__init__.py:16-36 re-exports), backed by libdbs_b200.so."""
from .samplers.mcmc import MetropolisHastings
from .samplers.diffusion import DiffusionSampler, PretrainedDiffusionSampler
from .analysis import (pair_correlation, magnetization_distribution, autocorrelation_time, energy_histogram,
                       compare_distributions, kl_divergence, symmetric_kl_divergence, magnetization_kl_divergence,
                       energy_kl_divergence, wasserstein_distance_1d, magnetization_wasserstein, energy_wasserstein,
                       correlation_function_comparison, comprehensive_comparison)

__all__ = ["MetropolisHastings", "DiffusionSampler", "PretrainedDiffusionSampler", "pair_correlation",
           "magnetization_distribution", "autocorrelation_time", "energy_histogram", "compare_distributions",
           "kl_divergence", "symmetric_kl_divergence", "magnetization_kl_divergence", "energy_kl_divergence",
           "wasserstein_distance_1d", "magnetization_wasserstein", "energy_wasserstein",
           "correlation_function_comparison", "comprehensive_comparison"]
