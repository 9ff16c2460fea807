"""Interface a sampled physical system exposes — the contract of the reference's systems/base.py:8-87
(energy, random_configuration, energy_per_particle, to/from_continuous, num_particles,
configuration_shape).  The engine's IsingModel and LennardJonesSystem are, like the reference's,
plain classes that do not derive from it; it is kept for third-party systems that do."""
import abc
from typing import Tuple

import torch


class PhysicalSystem(abc.ABC):
    @abc.abstractmethod
    def energy(self, configuration: torch.Tensor) -> torch.Tensor:
        """Energy of each configuration."""

    @abc.abstractmethod
    def random_configuration(self, batch_size: int = 1) -> torch.Tensor:
        """`batch_size` random configurations."""

    @abc.abstractmethod
    def energy_per_particle(self, configuration: torch.Tensor) -> torch.Tensor:
        """Energy divided by the number of particles / spins."""

    @abc.abstractmethod
    def to_continuous(self, configuration: torch.Tensor) -> torch.Tensor:
        """Configuration -> the continuous tensor the diffusion process noises."""

    @abc.abstractmethod
    def from_continuous(self, x: torch.Tensor) -> torch.Tensor:
        """Inverse of to_continuous."""

    @property
    @abc.abstractmethod
    def num_particles(self) -> int:
        """Number of particles / spins."""

    @property
    @abc.abstractmethod
    def configuration_shape(self) -> Tuple[int, ...]:
        """Shape of one configuration."""
