#!/usr/bin/env python
"""Generate tests/golden/analysis.npz by running the UNMODIFIED reference's analysis functions
(analysis/correlation.py:50-491, analysis/rdf.py:111-226) on fixed sample sets.

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/make_golden_analysis.py
Only inputs and outputs are stored.  Sample sets: two 16x16 Ising ensembles produced by the
reference's own MetropolisHastings at T = 2.0 and T = 3.0 (time-ordered, so the autocorrelation
estimate sees a real chain), one continuous (B, 1, L, L) set, and two synthetic g(r) curves; plus DiffusionProcess.compute_loss values.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DBS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import torch  # noqa: E402

from backend.ml.systems.ising import IsingModel  # noqa: E402
from backend.ml.samplers.mcmc import MetropolisHastings  # noqa: E402
from backend.ml.analysis import correlation as rc  # noqa: E402
from backend.ml.analysis import rdf as rr  # noqa: E402


def flat(prefix, d, out):
    for k, v in d.items():
        if isinstance(v, dict):
            flat(f"{prefix}{k}.", v, out)
        else:
            out[prefix + k] = np.asarray(v, dtype=np.float64)


def main():
    out = {}
    L = 16
    model = IsingModel(L)
    sets = {}
    for tag, T, seed in (("a", 2.0, 7), ("b", 3.0, 8)):
        np.random.seed(seed)
        torch.manual_seed(seed)
        sets[tag] = MetropolisHastings(model, T).sample(n_samples=120, n_sweeps=1, burn_in=20).float()
        out[f"spins_{tag}"] = sets[tag].numpy().astype(np.int8)
    a, b = sets["a"], sets["b"]
    torch.manual_seed(3)
    cont = (0.8 * a[:40].unsqueeze(1) + 0.3 * torch.randn(40, 1, L, L)).float()
    out["cont"] = cont.numpy()

    for tag, s in (("a", a), ("b", b), ("a4", a.unsqueeze(1)), ("cont", cont)):
        flat(f"magdist_{tag}.", rc.magnetization_distribution(s, 20), out)
        out[f"tau_{tag}"] = np.float64(rc.autocorrelation_time(s))
        flat(f"ehist_{tag}.", rc.energy_histogram(s, model, 15), out)
    out["tau_short"] = np.float64(rc.autocorrelation_time(a[:5]))
    out["tau_const"] = np.float64(rc.autocorrelation_time(torch.ones(20, L, L)))

    p = np.array([0.1, 0.4, 0.0, 0.5, 2.0])
    q = np.array([0.3, 0.3, 0.2, 0.0, 1.0])
    out["kl_p"], out["kl_q"] = p, q
    out["kl_pq"] = np.float64(rc.kl_divergence(p, q))
    out["kl_qp"] = np.float64(rc.kl_divergence(q, p))
    out["kl_sym"] = np.float64(rc.symmetric_kl_divergence(p, q))
    out["kl_eps"] = np.float64(rc.kl_divergence(p, q, epsilon=1e-4))
    rng = np.random.default_rng(0)
    w1, w2 = rng.normal(0, 1, 200).astype(np.float32), rng.normal(0.3, 1.5, 150).astype(np.float32)
    out["w1"], out["w2"] = w1, w2
    out["w_12"] = np.float64(rc.wasserstein_distance_1d(w1, w2))
    out["w_11"] = np.float64(rc.wasserstein_distance_1d(w1, w1))

    flat("magkl.", rc.magnetization_kl_divergence(a, b, 20), out)
    flat("ekl.", rc.energy_kl_divergence(a, b, model, 15), out)
    out["magw"] = np.float64(rc.magnetization_wasserstein(a, b))
    out["ew"] = np.float64(rc.energy_wasserstein(a, b, model))
    out["magw_cont"] = np.float64(rc.magnetization_wasserstein(cont, a[:40].unsqueeze(1)))
    flat("cmpdist.", rc.compare_distributions(a, b, model), out)
    # pair_correlation draws its reference sites from np.random: fix the seed right before each call chain
    np.random.seed(42)
    flat("corrcmp.", rc.correlation_function_comparison(a, b), out)
    np.random.seed(43)
    flat("comp.", rc.comprehensive_comparison(a, b, model, 20), out)

    # RDF curve statistics on two synthetic liquid-like curves (different grids)
    r1 = (np.arange(60) + 0.5) * 0.05
    g1 = np.where(r1 < 0.95, 0.0, 1 + 1.8 * np.exp(-(r1 - 1.1) ** 2 / 0.02) * 1.0 + 0.3 * np.cos(6 * (r1 - 1.1)) * np.exp(-(r1 - 1.1)))
    r2 = (np.arange(45) + 0.5) * 0.06
    g2 = np.where(r2 < 0.9, 0.0, 1 + 1.2 * np.exp(-(r2 - 1.2) ** 2 / 0.03) + 0.2 * np.cos(5 * (r2 - 1.2)) * np.exp(-(r2 - 1.2)))
    out["rdf_r1"], out["rdf_g1"], out["rdf_r2"], out["rdf_g2"] = r1, g1, r2, g2
    flat("rdfstat1.", rr.rdf_statistics(r1, g1), out)
    flat("rdfstat2.", rr.rdf_statistics(r2, g2, sigma=1.1), out)
    flat("rdfstat_none.", rr.rdf_statistics(r1[:10], g1[:10]), out)
    mono = np.linspace(2.0, 1.0, 60)
    out["rdf_mono"] = mono
    flat("rdfstat_mono.", rr.rdf_statistics(r1, mono), out)
    flat("rdfcmp_same.", rr.compare_rdfs(r1, g1, r1, g1 * 0.9 + 0.05), out)
    flat("rdfcmp_diff.", rr.compare_rdfs(r1, g1, r2, g2), out)

    # DiffusionProcess.compute_loss (models/diffusion.py:181-227): a trainer-side helper the replacement module keeps
    from backend.ml.models.diffusion import DiffusionProcess
    torch.manual_seed(5)
    sp, nz, t = torch.randn(6, 1, 8, 8), torch.randn(6, 1, 8, 8), torch.rand(6)
    out["closs_pred"], out["closs_noise"], out["closs_t"] = sp.numpy(), nz.numpy(), t.numpy()
    for lt in ("l2", "l1"):
        for w in ("uniform", "snr"):
            out[f"closs_{lt}_{w}"] = np.float64(DiffusionProcess().compute_loss(sp, nz, t, lt, w).item())

    path = os.path.join(ROOT, "tests", "golden", "analysis.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
