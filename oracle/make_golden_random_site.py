#!/usr/bin/env python
"""Golden for the reference's own random-site Metropolis chain (samplers/mcmc.py:26-60), run UNMODIFIED from
/root/reference under np.random.seed -> tests/golden/ising_random_site.npz.  It pins oracle.ising.random_site_sweeps,
the CPU baseline bench.py times for BASELINE configs 1 and 3.  Build container only:  python oracle/make_golden_random_site.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DBS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import torch  # noqa: E402

from backend.ml.samplers.mcmc import MetropolisHastings  # noqa: E402
from backend.ml.systems.ising import IsingModel  # noqa: E402

# (L, J, h, T, sweeps, numpy seed)
CASES = [(16, 1.0, 0.0, 2.269, 3, 123), (8, 1.0, 0.3, 1.5, 4, 7), (12, -1.0, 0.0, 3.5, 2, 99), (16, 1.0, 0.0, 0.0, 2, 5)]


def main():
    g = {"cases": np.array(CASES, dtype=np.float64)}
    for n, (L, J, h, T, sweeps, seed) in enumerate(CASES):
        torch.manual_seed(seed)
        model = IsingModel(L, J=J, h=h)
        start = model.random_configuration(1).squeeze(0)
        mh = MetropolisHastings(model, T)
        x = start.clone()
        np.random.seed(seed)
        for _ in range(sweeps):
            x = mh.sweep(x)
        g[f"start_{n}"], g[f"end_{n}"] = start.numpy(), x.numpy()
        g[f"next_uniform_{n}"] = np.array(np.random.random())     # the generator's position after the run
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ising_random_site.npz"), **g)
    print("wrote ising_random_site.npz")


if __name__ == "__main__":
    main()
