"""2-D periodic Lennard-Jones fluid — engine-backed replacement for systems/lennard_jones.py.

Public surface of the reference kept (systems/lennard_jones.py:7-268).  `energy`, `forces` and
`score` run the fused pair kernel of libdbs_b200.so (one CTA per configuration, energy and forces
in one pass, no (B,N,N) temporaries); the coordinate helpers are O(N) glue on whatever device the
caller's tensor lives on.
"""
from typing import Tuple

import numpy as np
import torch

from dbs_b200 import abi, runtime


class LennardJonesSystem:
    REDUCED_DENSITY_LIQUID = 0.8
    REDUCED_DENSITY_GAS = 0.3
    TRIPLE_POINT_TEMP = 0.694
    CRITICAL_TEMP = 1.326

    def __init__(self, n_particles: int, box_size: float, epsilon: float = 1.0, sigma: float = 1.0,
                 cutoff: float = 2.5):
        self.n_particles = n_particles
        self.box_size = box_size
        self.epsilon = epsilon
        self.sigma = sigma
        self.cutoff = cutoff * sigma
        self._cutoff_squared = self.cutoff ** 2
        sr6 = (self.sigma / self.cutoff) ** 6
        self._energy_shift = 4 * self.epsilon * (sr6 ** 2 - sr6)  # U(rc), systems/lennard_jones.py:53-56

    # ---------------------------------------------------------------- fused pair kernel
    def energy_and_forces(self, positions: torch.Tensor, want_energy=True, want_forces=True):
        """(E, F) from ONE pass over the pairs; either may be skipped (None)."""
        N = self.n_particles
        if positions.shape[-2] != N or positions.shape[-1] != 2:
            raise ValueError(f"expected (..., {N}, 2) positions, got {tuple(positions.shape)}")
        x = runtime.to_device(positions)
        lead = x.shape[:-2]
        B = int(np.prod(lead)) if len(lead) else 1
        e = torch.empty(B, dtype=torch.float32, device=x.device) if want_energy else None
        f = torch.empty((B, N, 2), dtype=torch.float32, device=x.device) if want_forces else None
        abi.check(abi.lib.dbs_lj_energy_forces(x.data_ptr(), B, N, float(self.box_size), float(self.epsilon),
                                               float(self.sigma), float(self.cutoff), float(self._energy_shift),
                                               runtime.ptr(e), runtime.ptr(f), runtime.current_stream()),
                  "lj_energy_forces")
        e = None if e is None else runtime.like_input(e.reshape(lead), positions)
        f = None if f is None else runtime.like_input(f.reshape(*lead, N, 2), positions)
        return e, f

    def energy(self, positions: torch.Tensor) -> torch.Tensor:
        """Shifted-truncated LJ energy, 0.5 * sum over ordered pairs (systems/lennard_jones.py:91-118)."""
        return self.energy_and_forces(positions, True, False)[0]

    def forces(self, positions: torch.Tensor) -> torch.Tensor:
        """F_i = sum_j f(r_ij) * unit(r_ij) (systems/lennard_jones.py:120-155)."""
        return self.energy_and_forces(positions, False, True)[1]

    def score(self, positions: torch.Tensor, temperature: float) -> torch.Tensor:
        """grad log p = F / T (systems/lennard_jones.py:157-170)."""
        return self.forces(positions) / temperature

    def energy_per_particle(self, positions: torch.Tensor) -> torch.Tensor:
        return self.energy(positions) / self.n_particles

    # ---------------------------------------------------------------- O(N) glue (not on the hot path)
    def minimum_image(self, dr: torch.Tensor) -> torch.Tensor:
        """Wrap displacements into [-L/2, L/2] (systems/lennard_jones.py:58-67)."""
        return dr - self.box_size * torch.round(dr / self.box_size)

    def pairwise_distances(self, positions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(dist (...,N,N), dr (...,N,N,2)) as systems/lennard_jones.py:69-89.  Kept for API
        completeness (analysis code reads it); the engine never materialises these tensors."""
        dr = self.minimum_image(positions[..., :, None, :] - positions[..., None, :, :])
        return torch.sqrt((dr ** 2).sum(dim=-1) + 1e-10), dr

    def random_configuration(self, batch_size: int = 1) -> torch.Tensor:
        return torch.rand(batch_size, self.n_particles, 2) * self.box_size

    def lattice_configuration(self, batch_size: int = 1) -> torch.Tensor:
        """Square-lattice start (systems/lennard_jones.py:197-220)."""
        side = int(np.ceil(np.sqrt(self.n_particles)))
        a = self.box_size / side
        k = torch.arange(self.n_particles, dtype=torch.float64)
        pos = torch.stack([(k % side + 0.5) * a, (torch.floor(k / side) + 0.5) * a], dim=-1)
        return pos.to(torch.float32).unsqueeze(0).repeat(batch_size, 1, 1)

    def to_continuous(self, positions: torch.Tensor) -> torch.Tensor:
        return 2 * positions / self.box_size - 1

    def from_continuous(self, x: torch.Tensor) -> torch.Tensor:
        return ((x + 1) * self.box_size / 2) % self.box_size

    @property
    def num_particles(self) -> int:
        return self.n_particles

    @property
    def configuration_shape(self) -> Tuple[int, int]:
        return (self.n_particles, 2)

    @property
    def density(self) -> float:
        return self.n_particles / (self.box_size ** 2)

    @property
    def reduced_density(self) -> float:
        return self.density * self.sigma ** 2
