from .correlation import pair_correlation

__all__ = ["pair_correlation"]
