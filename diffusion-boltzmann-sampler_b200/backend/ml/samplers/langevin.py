"""Langevin samplers for the Lennard-Jones fluid — engine-backed replacement for samplers/langevin.py.

Public surface of the reference kept (samplers/langevin.py:9-267): `LangevinDynamics(system,
temperature, dt, use_mh_correction)` with `step / sample / sample_with_trajectory`, and
`UnderdampedLangevin(system, temperature, dt, mass, friction)` with `step / sample`.  Both accept a
single configuration (N, 2) as the reference does, or a batch (B, N, 2) — an additive extension:
B independent chains advance in lock-step on the GPU, one fused pair pass per step (the reference
spends four O(N^2) passes per MALA step on one configuration).
"""
from typing import Generator, Optional

import numpy as np
import torch

from dbs_b200 import abi, runtime
from ..systems.lennard_jones import LennardJonesSystem


def _seed_from_numpy() -> int:
    # the reference's accept test draws from NumPy's global RNG (samplers/langevin.py:84)
    return int(np.random.randint(0, 2 ** 31 - 1))


class _State:
    """Device-resident chain state: positions with their cached forces and energies."""

    def __init__(self, system: LennardJonesSystem, positions: torch.Tensor):
        self.system = system
        self.single = positions.dim() == 2
        self.pos = runtime.to_device(positions).reshape(-1, system.n_particles, 2).clone()
        self.B = self.pos.shape[0]
        self.force = torch.empty_like(self.pos)
        self.energy = torch.empty(self.B, dtype=torch.float32, device=self.pos.device)
        self.refresh()

    def evaluate(self, pos, energy, force):
        s = self.system
        abi.check(abi.lib.dbs_lj_energy_forces(pos.data_ptr(), self.B, s.n_particles, float(s.box_size), float(s.epsilon),
                                               float(s.sigma), float(s.cutoff), float(s._energy_shift),
                                               runtime.ptr(energy), force.data_ptr(), runtime.current_stream()),
                  "lj_energy_forces")

    def refresh(self):
        self.evaluate(self.pos, self.energy, self.force)

    def positions(self, like: torch.Tensor) -> torch.Tensor:
        out = self.pos[0] if self.single else self.pos
        return runtime.like_input(out.clone(), like)


class LangevinDynamics:
    """Overdamped Langevin proposal + Metropolis-Hastings correction (MALA),
    x' = x + dt F(x)/T + sqrt(2 dt) eta  (samplers/langevin.py:9-44)."""

    def __init__(self, system: LennardJonesSystem, temperature: float, dt: float = 0.001,
                 use_mh_correction: bool = True, seed: Optional[int] = None):
        self.system = system
        self.temperature = temperature
        self.dt = dt
        self.use_mh_correction = use_mh_correction
        self.beta = 1.0 / temperature
        self.noise_scale = np.sqrt(2 * dt)
        self.seed = _seed_from_numpy() if seed is None else int(seed)
        self._steps_done = 0
        self.last_accepted = None

    # ---------------------------------------------------------------- engine
    def _advance(self, st: _State, n_steps: int, noise=None, uniforms=None, chain_offset: int = 0):
        s = self.system
        N = s.n_particles
        pos_new = torch.empty_like(st.pos)
        force_new = torch.empty_like(st.pos)
        energy_new = torch.empty_like(st.energy)
        acc = torch.empty(st.B, dtype=torch.int32, device=st.pos.device)
        stream = runtime.current_stream()
        for k in range(n_steps):
            step = self._steps_done + k
            nz = None if noise is None else runtime.to_device(noise[k]).reshape(st.B, N, 2)
            abi.check(abi.lib.dbs_mala_propose(st.pos.data_ptr(), st.force.data_ptr(), runtime.ptr(nz), pos_new.data_ptr(),
                                               st.B, N, float(self.dt), float(self.temperature), float(self.noise_scale),
                                               float(s.box_size), self.seed, chain_offset, step, stream), "mala_propose")
            st.evaluate(pos_new, energy_new, force_new)
            if self.use_mh_correction:
                un = None if uniforms is None else runtime.to_device(uniforms[k]).reshape(st.B)
                abi.check(abi.lib.dbs_mala_accept(st.pos.data_ptr(), st.force.data_ptr(), st.energy.data_ptr(),
                                                  pos_new.data_ptr(), force_new.data_ptr(), energy_new.data_ptr(),
                                                  runtime.ptr(un), st.B, N, float(self.dt), float(self.temperature),
                                                  self.seed, chain_offset, step, acc.data_ptr(), stream), "mala_accept")
            else:
                st.pos, pos_new = pos_new, st.pos
                st.force, force_new = force_new, st.force
                st.energy, energy_new = energy_new, st.energy
        self._steps_done += n_steps
        if self.use_mh_correction and n_steps:
            self.last_accepted = acc

    # ---------------------------------------------------------------- reference API
    def step(self, positions: torch.Tensor, noise=None, uniform=None) -> torch.Tensor:
        """One step; returns the updated positions, the argument is left untouched
        (samplers/langevin.py:46-89).  `noise` / `uniform` inject the Gaussian and the accept draw."""
        st = _State(self.system, positions)
        self._advance(st, 1, None if noise is None else [noise], None if uniform is None else [uniform])
        return st.positions(positions)

    def _initial(self, initial, n_chains):
        if initial is not None:
            return initial
        if n_chains:
            return self.system.lattice_configuration(batch_size=n_chains)
        return self.system.lattice_configuration(batch_size=1).squeeze(0)

    def sample(self, n_samples: int, n_steps: int = 100, initial: Optional[torch.Tensor] = None, burn_in: int = 1000,
               n_chains: Optional[int] = None, return_device: bool = False) -> torch.Tensor:
        """(n_samples, N, 2) — or (n_chains, n_samples, N, 2) for a batch — after burn-in, one
        snapshot every n_steps steps (samplers/langevin.py:91-126)."""
        init = self._initial(initial, n_chains)
        st = _State(self.system, init)
        self._advance(st, burn_in)
        out = torch.empty((n_samples, st.B, self.system.n_particles, 2), dtype=torch.float32, device=st.pos.device)
        for k in range(n_samples):
            self._advance(st, n_steps)
            out[k].copy_(st.pos)
        out = out[:, 0] if st.single else out.permute(1, 0, 2, 3).contiguous()
        return out if return_device else out.cpu()

    def sample_with_trajectory(self, n_steps: int, initial: Optional[torch.Tensor] = None
                               ) -> Generator[torch.Tensor, None, None]:
        """Yield the configuration after every step, first the initial one (samplers/langevin.py:128-153)."""
        init = self._initial(initial, None)
        st = _State(self.system, init)
        yield st.positions(init)
        for _ in range(n_steps):
            self._advance(st, 1)
            yield st.positions(init)


class UnderdampedLangevin:
    """BAOAB integrator of underdamped Langevin dynamics (samplers/langevin.py:156-267)."""

    def __init__(self, system: LennardJonesSystem, temperature: float, dt: float = 0.001, mass: float = 1.0,
                 friction: float = 1.0, seed: Optional[int] = None):
        self.system = system
        self.temperature = temperature
        self.dt = dt
        self.mass = mass
        self.friction = friction
        self.beta = 1.0 / temperature
        self.c1 = np.exp(-friction * dt)
        self.c2 = np.sqrt(1 - self.c1 ** 2) * np.sqrt(temperature / mass)
        self.seed = _seed_from_numpy() if seed is None else int(seed)
        self._steps_done = 0

    def _advance(self, st: _State, vel: torch.Tensor, n_steps: int, noise=None, chain_offset: int = 0):
        s = self.system
        N = s.n_particles
        stream = runtime.current_stream()
        for k in range(n_steps):
            nz = None if noise is None else runtime.to_device(noise[k]).reshape(st.B, N, 2)
            abi.check(abi.lib.dbs_baoab_first(st.pos.data_ptr(), vel.data_ptr(), st.force.data_ptr(), runtime.ptr(nz), st.B,
                                              N, float(self.dt), float(self.mass), float(self.c1), float(self.c2),
                                              float(s.box_size), self.seed, chain_offset, self._steps_done + k, stream),
                      "baoab_first")
            st.evaluate(st.pos, None, st.force)     # the force of the new positions is cached for the next step too
            abi.check(abi.lib.dbs_baoab_second(vel.data_ptr(), st.force.data_ptr(), st.B, N, float(self.dt),
                                               float(self.mass), stream), "baoab_second")
        self._steps_done += n_steps

    def step(self, positions: torch.Tensor, velocities: torch.Tensor, noise=None):
        """(positions, velocities) after one BAOAB step (samplers/langevin.py:192-225)."""
        st = _State(self.system, positions)
        vel = runtime.to_device(velocities).reshape(st.B, self.system.n_particles, 2).clone()
        self._advance(st, vel, 1, None if noise is None else [noise])
        v = vel[0] if st.single else vel
        return st.positions(positions), runtime.like_input(v.clone(), velocities)

    def sample(self, n_samples: int, n_steps: int = 100, initial: Optional[torch.Tensor] = None, burn_in: int = 1000,
               n_chains: Optional[int] = None, return_device: bool = False) -> torch.Tensor:
        """Maxwell-Boltzmann initial velocities, burn-in, snapshots (samplers/langevin.py:227-267)."""
        if initial is None:
            initial = self.system.lattice_configuration(batch_size=n_chains) if n_chains else \
                self.system.lattice_configuration(batch_size=1).squeeze(0)
        st = _State(self.system, initial)
        vel = torch.randn(st.pos.shape, device=st.pos.device) * float(np.sqrt(self.temperature / self.mass))
        self._advance(st, vel, burn_in)
        out = torch.empty((n_samples, st.B, self.system.n_particles, 2), dtype=torch.float32, device=st.pos.device)
        for k in range(n_samples):
            self._advance(st, vel, n_steps)
            out[k].copy_(st.pos)
        out = out[:, 0] if st.single else out.permute(1, 0, 2, 3).contiguous()
        return out if return_device else out.cpu()
