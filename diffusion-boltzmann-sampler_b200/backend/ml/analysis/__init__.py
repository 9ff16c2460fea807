"""Ensemble analysis; exports exactly the names of the reference's analysis package (each submodule lists its own)."""
from . import correlation as _correlation, rdf as _rdf
from .correlation import *  # noqa: F401,F403
from .rdf import *  # noqa: F401,F403

__all__ = tuple(_correlation.__all__) + tuple(_rdf.__all__)
