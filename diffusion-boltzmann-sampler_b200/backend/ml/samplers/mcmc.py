"""Metropolis sampler for the 2-D Ising model — engine-backed replacement for samplers/mcmc.py.

Public surface of the reference kept verbatim (samplers/mcmc.py:15-124): `MetropolisHastings(model,
temperature)` with `.model/.temperature/.beta`, `step`, `sweep`, `sample`, `sample_with_trajectory`.

What changes underneath: `sweep`/`sample`/`sample_with_trajectory` run the
engine's bit-packed *checkerboard* sweep (same stationary distribution, same acceptance rule
`dE <= 0 or u < exp(-beta*dE)`), with counter-based Philox draws; `step` stays one random-site
attempt that mutates and returns its argument.  Additive extensions: `n_chains=`, per-chain
`temperature` tensors, `seed=`, and packed-state entry points for bulk runs that never leave the GPU.
"""
from typing import Generator, Optional

import numpy as np
import torch

from dbs_b200 import abi, runtime
from ..systems.ising import IsingModel, acceptance_table


class MetropolisHastings:
    def __init__(self, model: IsingModel, temperature, seed: Optional[int] = None):
        self.model = model
        self.temperature = temperature
        temps = np.atleast_1d(np.asarray(temperature, dtype=np.float64))
        self._per_chain = np.ndim(temperature) > 0
        if not self._per_chain:
            t = float(temperature)
            self.beta = 1.0 / t if t > 0 else float("inf")  # samplers/mcmc.py:24
        else:
            with np.errstate(divide="ignore"):
                self.beta = np.where(temps > 0, 1.0 / temps, np.inf)
        # one table per distinct temperature, chains index into it
        uniq, inv = np.unique(temps, return_inverse=True)
        self._tables_host = np.stack([acceptance_table(model.J, model.h, float(t)) for t in uniq])
        self._table_index_host = inv.astype(np.int32)
        self._tables_dev = None
        self._index_dev = None
        # the reference draws from NumPy's global RNG (samplers/mcmc.py:36-46); take the engine seed
        # from it so np.random.seed() keeps runs reproducible
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self._sweeps_done = 0
        self._starts_drawn = 0
        # h == 0: the kernel's multi-spin path (agreement counts for 32 sites at once, two thresholds)
        self.zero_field = 0 if (model.h != 0 or model.J == 0) else (1 if model.J > 0 else -1)

    # ---------------------------------------------------------------- device-side state
    def _tables(self):
        if self._tables_dev is None:
            dev = runtime.require_cuda()
            self._tables_dev = torch.from_numpy(self._tables_host.view(np.int32).copy()).to(dev)
            if self._per_chain:
                self._index_dev = torch.from_numpy(self._table_index_host).to(dev)
        return self._tables_dev, self._index_dev

    def sweeps_packed(self, bits: torch.Tensor, n_sweeps: int, chain_offset: int = 0, uniforms=None,
                      sweep_offset: Optional[int] = None, chain_stride: int = 1) -> torch.Tensor:
        """n_sweeps checkerboard sweeps in place on packed GPU state (n_chains, L, W): ONE launch,
        lattices resident in shared memory throughout.  `uniforms` (uint32-as-int32
        [n_sweeps][n_chains][L][L]) is the test-mode injection port.  Chain c draws from the Philox
        stream of global id chain_offset + c * chain_stride (striped multi-GPU sharding)."""
        tables, index = self._tables()
        n = bits.shape[0]
        if self._per_chain and index.shape[0] != n:
            raise ValueError(f"per-chain temperature has {index.shape[0]} entries for {n} chains")
        off = self._sweeps_done if sweep_offset is None else sweep_offset
        abi.check(abi.lib.dbs_ising_sweeps_strided(bits.data_ptr(), self.model.size, n, int(n_sweeps), tables.data_ptr(),
                                                   tables.shape[0], runtime.ptr(index), self.seed, chain_offset,
                                                   chain_stride, off, runtime.ptr(uniforms), self.zero_field,
                                                   runtime.current_stream()), "ising_sweeps")
        if sweep_offset is None:
            self._sweeps_done += int(n_sweeps)
        return bits

    # ---------------------------------------------------------------- reference API
    def step(self, spins: torch.Tensor) -> torch.Tensor:
        """One random-site attempt (samplers/mcmc.py:26-50): mutates and returns `spins`, flips at
        most one spin.  Site and uniform come from NumPy's global RNG like the reference's."""
        if self._per_chain:
            raise NotImplementedError("MetropolisHastings.step: one random-site attempt uses one acceptance table; with a "
                                      "per-chain temperature tensor use sweep() / sample() (tables are selected per chain there)")
        L = self.model.size
        i = int(np.random.randint(L))
        j = int(np.random.randint(L))
        # (the reference draws its uniform only when dE > 0, which it learns from a host read of dE; here the draw is
        # made unconditionally so that no device->host synchronisation is needed: same acceptance rule, a different
        # position in NumPy's stream afterwards)
        r = int(np.random.randint(0, 2 ** 32, dtype=np.uint64))
        tables, _ = self._tables()
        x = runtime.to_device(spins)
        lat = x.reshape(-1, L, L)
        abi.check(abi.lib.dbs_ising_site_update(lat.data_ptr(), L, lat.shape[0], i, j, tables.data_ptr(), r,
                                                runtime.current_stream()), "ising_site_update")
        if x.data_ptr() != spins.data_ptr():
            spins.copy_(x.reshape(spins.shape))
        return spins

    def sweep(self, spins: torch.Tensor) -> torch.Tensor:
        """One sweep = L*L attempts (samplers/mcmc.py:52-60), done as one checkerboard sweep.
        Mutates and returns `spins` ((L, L) or (n, L, L))."""
        bits = self.model.pack(spins)
        self.sweeps_packed(bits, 1)
        out = self.model.unpack(bits)
        spins.copy_(out.reshape(spins.shape))
        return spins

    def _initial_bits(self, initial, n_chains):
        L = self.model.size
        if initial is None:
            # hot start (samplers/mcmc.py:82) from the engine's Philox INIT stream; every call draws a fresh one, as the
            # reference does (chain ids continue where the previous call's stopped)
            off = self._starts_drawn
            self._starts_drawn += n_chains
            return self.model.random_packed(n_chains, self.seed, chain_offset=off)
        init = initial if isinstance(initial, torch.Tensor) else torch.as_tensor(initial)
        if init.dim() == 2:
            init = init.unsqueeze(0).expand(n_chains, L, L)
        elif init.shape[0] != n_chains:
            raise ValueError(f"initial has {init.shape[0]} lattices for n_chains={n_chains}")
        return self.model.pack(init.contiguous())

    def sample(self, n_samples: int, n_sweeps: int = 10, initial: Optional[torch.Tensor] = None, burn_in: int = 100,
               n_chains: Optional[int] = None, return_device: bool = False) -> torch.Tensor:
        """Burn in, then record a snapshot every `n_sweeps` sweeps (samplers/mcmc.py:62-97).
        Returns (n_samples, L, L) — or (n_chains, n_samples, L, L) when `n_chains` is given — on the
        host (reference convention) unless return_device=True."""
        L = self.model.size
        batched = n_chains is not None
        if not batched:
            n_chains = initial.shape[0] if (initial is not None and initial.dim() == 3) else 1
            batched = initial is not None and initial.dim() == 3
        bits = self._initial_bits(initial, n_chains)
        if burn_in > 0:
            self.sweeps_packed(bits, burn_in)
        out = torch.empty((n_samples, n_chains, L, L), dtype=torch.float32, device=bits.device)
        for k in range(n_samples):
            if n_sweeps > 0:
                self.sweeps_packed(bits, n_sweeps)
            self.model.unpack(bits, out[k])
        out = out.permute(1, 0, 2, 3) if batched else out[:, 0]
        out = out.contiguous()
        return out if return_device else out.cpu()

    def sample_with_trajectory(self, n_steps: int, initial: Optional[torch.Tensor] = None
                               ) -> Generator[torch.Tensor, None, None]:
        """Yield the lattice after every sweep, first frame = the initial state
        (samplers/mcmc.py:99-124)."""
        bits = self._initial_bits(initial if initial is None or initial.dim() == 2 else initial[0], 1)
        yield self.model.unpack(bits)[0].cpu()
        for _ in range(n_steps):
            self.sweeps_packed(bits, 1)
            yield self.model.unpack(bits)[0].cpu()
