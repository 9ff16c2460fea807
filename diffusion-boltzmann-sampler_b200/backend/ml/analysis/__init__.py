"""Same export list as the reference's analysis/__init__.py."""
from .correlation import (pair_correlation, magnetization_distribution, autocorrelation_time, energy_histogram,
                          compare_distributions, kl_divergence, symmetric_kl_divergence,
                          magnetization_kl_divergence, energy_kl_divergence, wasserstein_distance_1d,
                          magnetization_wasserstein, energy_wasserstein, correlation_function_comparison,
                          comprehensive_comparison)
from .rdf import compute_rdf, compute_rdf_from_system, rdf_statistics, compare_rdfs

__all__ = [
    "pair_correlation", "magnetization_distribution", "autocorrelation_time", "energy_histogram",
    "compare_distributions", "kl_divergence", "symmetric_kl_divergence", "magnetization_kl_divergence",
    "energy_kl_divergence", "wasserstein_distance_1d", "magnetization_wasserstein", "energy_wasserstein",
    "correlation_function_comparison", "comprehensive_comparison", "compute_rdf", "compute_rdf_from_system",
    "rdf_statistics", "compare_rdfs",
]
