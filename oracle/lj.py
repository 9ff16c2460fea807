"""Lennard-Jones pair energy / forces restated in numpy float32.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates systems/lennard_jones.py:29-155 of /root/reference/backend/ml op by op (fp32, the same
masks and epsilons); only pow(x, 6) is written as a product and the pair sum is numpy's.
"""
import numpy as np

F = np.float32


class LJParams:
    def __init__(self, n_particles, box_size, epsilon=1.0, sigma=1.0, cutoff=2.5):
        # systems/lennard_jones.py:46-56
        self.n_particles = n_particles
        self.box_size = box_size
        self.epsilon = epsilon
        self.sigma = sigma
        self.cutoff = cutoff * sigma
        r6 = (self.sigma / self.cutoff) ** 6
        self.energy_shift = 4 * self.epsilon * (r6 ** 2 - r6)


def _pairs(p, pos):
    pos = np.asarray(pos, dtype=F)
    dr = pos[..., :, None, :] - pos[..., None, :, :]
    dr = dr - F(p.box_size) * np.round(dr / F(p.box_size))          # :67  (round half to even)
    dist = np.sqrt((dr * dr).sum(axis=-1, dtype=F) + F(1e-10))        # :88
    mask = (dist > F(0.1)) & (dist < F(p.cutoff))                     # :104
    ds = np.where(mask, dist, F(1.0))
    q = F(p.sigma) / ds
    q2 = q * q
    r6 = q2 * q2 * q2
    r12 = r6 * r6
    return dr, ds, mask, r6, r12


def energy(p, pos):
    """systems/lennard_jones.py:91-118."""
    _, _, mask, r6, r12 = _pairs(p, pos)
    pot = F(4 * p.epsilon) * (r12 - r6) - F(p.energy_shift)
    pot = np.where(mask, pot, F(0))
    return (F(0.5) * pot.astype(np.float64).sum(axis=(-1, -2))).astype(F)


def forces(p, pos):
    """systems/lennard_jones.py:120-155."""
    dr, ds, mask, r6, r12 = _pairs(p, pos)
    fm = F(24 * p.epsilon) / ds * (F(2) * r12 - r6)
    fm = np.where(mask, fm, F(0))
    unit = dr / (ds[..., None] + F(1e-10))
    return (fm[..., None] * unit).astype(np.float64).sum(axis=-2).astype(F)
