#!/usr/bin/env python
"""Generate tests/golden/ising_odd.npz: odd lattice sizes, by running the UNMODIFIED reference.

    python oracle/make_golden_odd.py        (build container only; /root/reference is not on the GPU box)

An odd torus is not bipartite, so the engine sweeps three colour classes (oracle.ising.site_colours).  Sites of
one class do not interact, hence driving the reference's own MetropolisHastings.step() over class 0, 1, 2 with
the uniform assigned to each site reproduces the parallel class update bit for bit — the same construction as
make_golden.py's checkerboard cases.  Also stored: the reference's energy / magnetisation and pair_correlation
on odd lattices.  Only inputs and outputs are stored.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DBS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.dont_write_bytecode = True

import torch  # noqa: E402

import backend.ml.samplers.mcmc as ref_mcmc  # noqa: E402
from backend.ml.systems.ising import IsingModel  # noqa: E402
from backend.ml.analysis.correlation import pair_correlation  # noqa: E402

from oracle import ising as o_ising  # noqa: E402  (Philox draw layout, hot start, colour classes)


class _Shim:
    """Stands in for np.random inside the reference's mcmc.py only."""

    def __init__(self):
        self.q = []
        self.u = 0.0

    def randint(self, n):
        return self.q.pop(0)

    def random(self):
        return self.u


def main():
    g = {}
    specs = [
        # L, J, h, T, n_sweeps, seed
        (9, 1.0, 0.0, 2.269, 4, 21),
        (9, 0.7, -0.45, 1.1, 3, 22),
        (15, 1.0, 0.0, 1.5, 3, 23),
        (15, -1.0, 0.2, 2.0, 2, 24),
        (33, 1.0, 0.0, 2.269, 2, 25),     # one full word + one bit per row
        (5, 1.0, 0.0, 0.0, 3, 26),        # smallest odd lattice, T = 0
    ]
    shim = _Shim()
    saved = ref_mcmc.np
    ref_mcmc.np = types.SimpleNamespace(random=shim, exp=np.exp)
    try:
        for n_, (L, J, h, T, n_sweeps, seed) in enumerate(specs):
            n_colours, colour_of = o_ising.site_colours(L)
            assert n_colours == 3
            sampler = ref_mcmc.MetropolisHastings(IsingModel(L, J=J, h=h), T)
            n_chains = 2
            start = o_ising.unpack(o_ising.init_bits_words(seed, np.arange(n_chains), L), L)
            r = np.stack([o_ising.sweep_draws(seed, np.arange(n_chains), s, L) for s in range(n_sweeps)])
            final = np.zeros_like(start)
            for c in range(n_chains):
                x = torch.tensor(start[c].astype(np.float32))
                for s in range(n_sweeps):
                    for colour in range(3):
                        for i in range(L):
                            for j in range(L):
                                if colour_of[i, j] != colour:
                                    continue
                                shim.q = [i, j]
                                shim.u = float(r[s, c, i, j]) * 2.0 ** -32
                                x = sampler.step(x)
                final[c] = x.numpy().astype(np.int8)
            for k, v in dict(L=L, J=J, h=h, T=T, n_sweeps=n_sweeps, seed=seed, start=start, r=r, final=final).items():
                g[f"c{n_}_{k}"] = np.asarray(v)
        g["n_cases"] = np.asarray(len(specs))
    finally:
        ref_mcmc.np = saved

    rng = np.random.RandomState(4)
    for L in (9, 33, 127):
        lat = torch.tensor(rng.choice([-1.0, 1.0], size=(4, L, L)).astype(np.float32))
        m = IsingModel(L, J=0.7, h=-0.45)
        g[f"obs_L{L}_spins"] = lat.numpy().astype(np.int8)
        g[f"obs_L{L}_E"] = m.energy(lat).numpy()
        g[f"obs_L{L}_M"] = m.magnetization(lat).numpy()
    for L in (9, 15, 33):
        samples = torch.tensor(rng.choice([-1.0, 1.0], size=(6, L, L), p=[0.35, 0.65]).astype(np.float32))
        np.random.seed(200 + L)
        res = pair_correlation(samples)
        np.random.seed(200 + L)
        g[f"corr_L{L}_ref_indices"] = np.random.choice(L * L, min(100, L * L), replace=False)
        g[f"corr_L{L}_samples"] = samples.numpy().astype(np.int8)
        g[f"corr_L{L}_C"] = np.array(res["C_r"])

    path = os.path.join(ROOT, "tests", "golden", "ising_odd.npz")
    np.savez_compressed(path, **g)
    print(f"wrote {path}: {len(g)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
