"""2-D periodic Ising model — engine-backed replacement for the reference's systems/ising.py.

Same public surface as the reference class (systems/ising.py:7-120): `IsingModel(size, J, h)`,
`T_CRITICAL`, `random_configuration`, `energy`, `local_energy_diff`, `magnetization`,
`energy_per_spin`, `to_continuous`, `from_continuous`.  Tensors may live on the host (the
reference's convention; results come back on the host) or on the GPU (results stay there).
All lattice arithmetic runs in libdbs_b200.so; nothing is computed on the CPU.
"""
import math

import numpy as np
import torch

from dbs_b200 import abi, runtime


def acceptance_table(J, h, temperature):
    """Host-side integer acceptance table for the sweep kernel: uint32 [16][2] {lo, hi} per site
    class k = 8*s_bit + (#up neighbours).  dE_k is evaluated in fp32 in the reference's order
    (systems/ising.py:91: 2*s*(J*nb + h)), widened to float64 for exp(-beta*dE) as
    samplers/mcmc.py:46 does, and turned into the exact integer threshold ceil(p*2^32)
    (u = r*2^-32 < p  <=>  r < ceil(p*2^32)).  dE <= 0 (samplers/mcmc.py:43) sets hi = 1."""
    beta = 1.0 / temperature if temperature > 0 else float("inf")
    tbl = np.zeros((16, 2), dtype=np.uint32)
    two, Jf, hf = np.float32(2.0), np.float32(J), np.float32(h)
    for k in range(16):
        u = k & 7
        if u > 4:
            continue
        s = np.float32(1.0 if k >= 8 else -1.0)
        dE = float((two * s) * (Jf * np.float32(2 * u - 4) + hf))
        if dE <= 0.0:
            tbl[k, 1] = 1
            continue
        p = math.exp(-beta * dE) if beta * dE < 745.2 else 0.0
        thr = math.ceil(p * 4294967296.0)
        if thr >= 4294967296:
            tbl[k, 1] = 1
        else:
            tbl[k, 0] = thr
    return tbl


class IsingModel:
    """H = -J sum_<ij> s_i s_j - h sum_i s_i on an L x L torus (systems/ising.py:7-34)."""

    T_CRITICAL = 2.0 / math.log(1.0 + math.sqrt(2.0))  # 2.269185...

    def __init__(self, size: int, J: float = 1.0, h: float = 0.0):
        self.size = size
        self.J = J
        self.h = h

    # ---------------------------------------------------------------- configurations
    def random_configuration(self, batch_size: int = 1, seed=None, device=None) -> torch.Tensor:
        """(batch, L, L) of +/-1 (systems/ising.py:36-45).  The bits come from the engine's Philox
        INIT stream; `seed` defaults to a draw from torch's global generator so that
        torch.manual_seed() keeps the call reproducible.  Host tensor unless device='cuda'."""
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        bits = self.random_packed(batch_size, seed)
        out = self.unpack(bits)
        return out if (device is not None and str(device).startswith("cuda")) else out.cpu()

    def random_packed(self, n_chains: int, seed: int, chain_offset: int = 0, chain_stride: int = 1) -> torch.Tensor:
        """Bit-packed hot start on the GPU: int32 tensor (n_chains, L, W); chain c is the chain of global id
        chain_offset + c * chain_stride."""
        dev = runtime.require_cuda()
        L = self.size
        wpc = abi.lib.dbs_ising_words_per_chain(L)
        if wpc < 0:
            raise ValueError(f"IsingModel: lattice size {L} is not supported by the packed engine (4..128, not 3)")
        bits = torch.empty((n_chains, L, wpc // L), dtype=torch.int32, device=dev)
        abi.check(abi.lib.dbs_ising_init_packed_strided(bits.data_ptr(), L, n_chains, seed, chain_offset, chain_stride,
                                                        runtime.current_stream()), "ising_init_packed")
        return bits

    def pack(self, spins: torch.Tensor) -> torch.Tensor:
        """fp32 (..., L, L) -> packed int32 (n, L, W) on the GPU (x > 0 -> bit 1)."""
        L = self.size
        x = runtime.to_device(spins).reshape(-1, L, L)
        wpc = abi.lib.dbs_ising_words_per_chain(L)
        if wpc < 0:
            raise ValueError(f"IsingModel: lattice size {L} is not supported by the packed engine (4..128, not 3)")
        bits = torch.empty((x.shape[0], L, wpc // L), dtype=torch.int32, device=x.device)
        abi.check(abi.lib.dbs_ising_pack(x.data_ptr(), bits.data_ptr(), L, x.shape[0], runtime.current_stream()),
                  "ising_pack")
        return bits

    def unpack(self, bits: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """packed (n, L, W) -> fp32 (n, L, L) of +/-1 on the GPU."""
        L = self.size
        n = bits.shape[0]
        if out is None:
            out = torch.empty((n, L, L), dtype=torch.float32, device=bits.device)
        abi.check(abi.lib.dbs_ising_unpack(bits.data_ptr(), out.data_ptr(), L, n, runtime.current_stream()),
                  "ising_unpack")
        return out

    # ---------------------------------------------------------------- observables
    def _observables(self, spins, want_e, want_m):
        L = self.size
        if spins.shape[-1] != L or spins.shape[-2] != L:
            raise ValueError(f"IsingModel({L}): expected (..., {L}, {L}) spins, got {tuple(spins.shape)}")
        x = runtime.to_device(spins)
        lead = x.shape[:-2]
        n = int(np.prod(lead)) if len(lead) else 1
        e = torch.empty(n, dtype=torch.float32, device=x.device) if want_e else None
        m = torch.empty(n, dtype=torch.float32, device=x.device) if want_m else None
        abi.check(abi.lib.dbs_ising_observables_f32(x.data_ptr(), L, n, float(self.J), float(self.h), runtime.ptr(e),
                                                    runtime.ptr(m), runtime.current_stream()), "ising_observables")
        res = []
        for t in (e, m):
            res.append(None if t is None else runtime.like_input(t.reshape(lead), spins))
        return res

    def energy(self, spins: torch.Tensor) -> torch.Tensor:
        """Total energy per configuration, shape (...,) (systems/ising.py:47-66)."""
        return self._observables(spins, True, False)[0]

    def magnetization(self, spins: torch.Tensor) -> torch.Tensor:
        """Mean spin per configuration (systems/ising.py:93-102)."""
        return self._observables(spins, False, True)[1]

    def energy_per_spin(self, spins: torch.Tensor) -> torch.Tensor:
        """systems/ising.py:104-106."""
        return self.energy(spins) / (self.size * self.size)

    def packed_observables(self, bits: torch.Tensor):
        """(E, M) fp32 tensors of shape (n,) straight from packed GPU state (no unpack)."""
        n = bits.shape[0]
        e = torch.empty(n, dtype=torch.float32, device=bits.device)
        m = torch.empty(n, dtype=torch.float32, device=bits.device)
        abi.check(abi.lib.dbs_ising_observables(bits.data_ptr(), self.size, n, float(self.J), float(self.h),
                                                e.data_ptr(), m.data_ptr(), runtime.current_stream()),
                  "ising_observables")
        return e, m

    def local_energy_diff(self, spins: torch.Tensor, i: int, j: int) -> torch.Tensor:
        """dE of flipping (i, j): 2*s*(J*sum_nbrs + h) (systems/ising.py:68-91).  O(1) indexing on
        whatever device holds `spins`; the sweep kernel folds this into its class table instead."""
        L = self.size
        s = spins[..., i, j]
        nb = (spins[..., (i + 1) % L, j] + spins[..., (i - 1) % L, j]
              + spins[..., i, (j + 1) % L] + spins[..., i, (j - 1) % L])
        return 2 * s * (self.J * nb + self.h)

    # ---------------------------------------------------------------- diffusion adapters
    def to_continuous(self, spins: torch.Tensor) -> torch.Tensor:
        """(B, L, L) -> (B, 1, L, L) (systems/ising.py:108-113)."""
        return spins.unsqueeze(1)

    def from_continuous(self, x: torch.Tensor) -> torch.Tensor:
        """(B, 1, L, L) -> sign (B, L, L); sign(0) = 0 as in the reference (systems/ising.py:115-120)."""
        return torch.sign(x.squeeze(1))
