from .ising import IsingModel
from .lennard_jones import LennardJonesSystem

__all__ = ["IsingModel", "LennardJonesSystem"]
