from .base import PhysicalSystem
from .ising import IsingModel
from .lennard_jones import LennardJonesSystem

__all__ = ["PhysicalSystem", "IsingModel", "LennardJonesSystem"]
