#!/usr/bin/env python
"""Ensemble observables of the UNMODIFIED reference Metropolis chain (random-site, samplers/mcmc.py:26-60)
-> tests/golden/ising_ensemble.npz.

BASELINE config 1 as-is: IsingModel(16), T = 2.269, 64 chains, hot start, 1000 sweeps (the first 200 discarded),
plus two shorter runs at T = 1.5 (cold start) and T = 3.5 (hot start).  For every chain the time averages of
E/N, |M|, M^2, M^4 over the measured sweeps are stored; chains are independent, so their spread gives the
standard error of the ensemble mean whatever the autocorrelation time is.  The engine's checkerboard chains
sample the same Boltzmann distribution with different dynamics: tests/test_ising_gpu.py compares the two
ensembles within 3 sigma (north_star: "ensemble observables agree within statistical error").

Run in the build container only (the reference does not exist on the GPU box), ~3 min on 8 cores:
    python oracle/make_golden_ensemble.py            (and with --odd: a 9x9 lattice -> ising_ensemble_odd.npz)
Nothing from the reference is copied: only its outputs are stored.
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DBS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

# (T, n_chains, start, burn_in sweeps, measured sweeps)
if "--odd" in sys.argv:
    # an odd lattice (the engine sweeps three colour classes there): python oracle/make_golden_ensemble.py --odd
    L, OUT = 9, "ising_ensemble_odd.npz"
    RUNS = [(2.269, 48, "hot", 150, 450), (1.5, 16, "cold", 100, 300), (3.5, 16, "hot", 100, 300)]
else:
    L, OUT = 16, "ising_ensemble.npz"
    RUNS = [(2.269, 64, "hot", 200, 800), (1.5, 16, "cold", 100, 300), (3.5, 16, "hot", 100, 300)]


def one_chain(job):
    T, chain, start, burn, meas = job
    import torch
    torch.set_num_threads(1)
    from backend.ml.samplers.mcmc import MetropolisHastings
    from backend.ml.systems.ising import IsingModel
    seed = 1234 + 1000 * int(round(T * 1000)) + chain
    np.random.seed(seed)            # the reference draws sites and uniforms from the global NumPy state
    torch.manual_seed(seed)         # ... and the hot start from torch.randint
    model = IsingModel(L)
    mh = MetropolisHastings(model, T)
    x = model.random_configuration(1).squeeze(0) if start == "hot" else torch.ones(L, L)
    acc = np.zeros(4)
    for s in range(burn + meas):
        x = mh.sweep(x)
        if s >= burn:
            e = float(model.energy_per_spin(x))
            m = float(model.magnetization(x))
            acc += (e, abs(m), m * m, m ** 4)
    return acc / meas


def main():
    jobs = [(T, c, start, burn, meas) for T, n, start, burn, meas in RUNS for c in range(n)]
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(one_chain, jobs, chunksize=1)
    out = {"L": np.array(L), "runs": np.array([(T, n, 1.0 if st == "hot" else 0.0, b, m) for T, n, st, b, m in RUNS])}
    k = 0
    for i, (T, n, st, b, m) in enumerate(RUNS):
        out[f"chains_{i}"] = np.array(res[k:k + n])      # [n_chains][E/N, |M|, M^2, M^4] time averages
        k += n
        mean = out[f"chains_{i}"].mean(0)
        sem = out[f"chains_{i}"].std(0, ddof=1) / np.sqrt(n)
        print(f"T={T}: E/N={mean[0]:.4f}+-{sem[0]:.4f} |M|={mean[1]:.4f}+-{sem[1]:.4f} "
              f"M2={mean[2]:.4f}+-{sem[2]:.4f} M4={mean[3]:.4f}+-{sem[3]:.4f}")
    path = os.path.join(ROOT, "tests", "golden", OUT)
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
