"""Physical systems on the hot path; the names the reference's systems package exports."""
from . import base as _base, ising as _ising, lennard_jones as _lj

PhysicalSystem = _base.PhysicalSystem
IsingModel = _ising.IsingModel
LennardJonesSystem = _lj.LennardJonesSystem

__all__ = ("PhysicalSystem", "IsingModel", "LennardJonesSystem")
