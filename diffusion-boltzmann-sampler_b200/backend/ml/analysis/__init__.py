from .correlation import pair_correlation
from .rdf import compute_rdf, compute_rdf_from_system

__all__ = ["pair_correlation", "compute_rdf", "compute_rdf_from_system"]
