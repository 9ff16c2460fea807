"""Langevin samplers and the RDF histogram restated in numpy float32.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).

Restates, citing /root/reference/backend/ml:
  LangevinDynamics.step      samplers/langevin.py:46-89    (noise and the accept uniform injected)
  UnderdampedLangevin.step   samplers/langevin.py:192-225  (noise injected)
  compute_rdf                analysis/rdf.py:9-83
"""
import numpy as np

from . import lj as o_lj

F = np.float32


def _wrap(x, box):
    return np.remainder(x, F(box)).astype(F)   # torch's % : sign of the divisor


def mala_step(p, pos, temperature, dt, noise, u, use_mh=True):
    """One step on a single configuration (N,2); returns (new positions, accepted flag)."""
    pos = np.asarray(pos, dtype=F)
    forces = o_lj.forces(p, pos[None])[0]
    nz = np.asarray(noise, dtype=F) * F(np.sqrt(2 * dt))
    new = pos + F(dt) * forces / F(temperature) + nz
    new = _wrap(new, p.box_size)
    if not use_mh:
        return new, True
    e_old = float(o_lj.energy(p, pos[None])[0])
    e_new = float(o_lj.energy(p, new[None])[0])
    f_new = o_lj.forces(p, new[None])[0]
    drift_f = pos + F(dt) * forces / F(temperature)
    lq_f = F(-0.5) * ((new - drift_f) ** 2).sum(dtype=F) / F(2 * dt)
    drift_b = new + F(dt) * f_new / F(temperature)
    lq_b = F(-0.5) * ((pos - drift_b) ** 2).sum(dtype=F) / F(2 * dt)
    log_alpha = F(-(1.0 / temperature) * (e_new - e_old)) + lq_b - lq_f
    if np.log(u) < float(log_alpha):
        return new, True
    return pos, False


def baoab_step(p, pos, vel, temperature, dt, mass, friction, noise):
    pos, vel = np.asarray(pos, dtype=F), np.asarray(vel, dtype=F)
    c1 = np.exp(-friction * dt)
    c2 = np.sqrt(1 - c1 ** 2) * np.sqrt(temperature / mass)
    f = o_lj.forces(p, pos[None])[0]
    vel = vel + F(0.5 * dt) * f / F(mass)
    pos = pos + F(0.5 * dt) * vel
    vel = F(c1) * vel + F(c2) * np.asarray(noise, dtype=F)
    pos = pos + F(0.5 * dt) * vel
    pos = _wrap(pos, p.box_size)
    f = o_lj.forces(p, pos[None])[0]
    vel = vel + F(0.5 * dt) * f / F(mass)
    return pos.astype(F), vel.astype(F)


def rdf_histogram(positions, box_size, n_bins=100, r_max=None):
    """Integer pair histogram + (centres, g_r) as analysis/rdf.py:40-83."""
    pos = np.asarray(positions, dtype=F)
    if pos.ndim == 2:
        pos = pos[None]
    n_samples, n, _ = pos.shape
    if r_max is None:
        r_max = box_size / 2
    edges = np.linspace(0, r_max, n_bins + 1)
    centers = (edges[:-1] + edges[1:]) / 2
    dr = edges[1] - edges[0]
    iu = np.triu_indices(n, k=1)
    hist = np.zeros(n_bins, dtype=np.int64)
    for s in range(n_samples):
        d = pos[s][:, None, :] - pos[s][None, :, :]
        d = d - F(box_size) * np.round(d / F(box_size))
        dist = np.sqrt((d * d).sum(axis=-1, dtype=F))
        hist += np.histogram(dist[iu], bins=edges)[0]
    total_pairs = n * (n - 1) / 2 * n_samples
    rho = n / box_size ** 2
    g_r = hist / (total_pairs * (2 * np.pi * centers * dr) * rho + 1e-10)
    return hist, centers, g_r
