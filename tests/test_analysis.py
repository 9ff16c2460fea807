"""Ensemble analysis: the replacement analysis/correlation.py + analysis/rdf.py against the outputs of
the unmodified reference (tests/golden/analysis.npz, oracle/make_golden_analysis.py).

CPU tests: the host statistics on series computed by the oracle.  GPU tests: the public functions end to end, their
per-sample series coming from the engine's observable kernels through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ising as o_ising

G = load_golden("analysis.npz")
TOL = dict(rtol=1e-6, atol=1e-9)


def _sets():
    return torch.tensor(G["spins_a"].astype(np.float32)), torch.tensor(G["spins_b"].astype(np.float32))


def _sub(prefix):
    return {k[len(prefix):]: G[k] for k in G.files if k.startswith(prefix)}


def _check_dict(got, prefix, **tol):
    want = _sub(prefix)
    tol = tol or TOL

    def walk(d, pre):
        for k, v in d.items():
            if isinstance(v, dict):
                walk(v, f"{pre}{k}.")
            else:
                np.testing.assert_allclose(np.asarray(v, dtype=np.float64), want[pre + k], err_msg=pre + k, **tol)
    walk(got, "")
    flat_keys = set()

    def keys(d, pre):
        for k, v in d.items():
            keys(v, f"{pre}{k}.") if isinstance(v, dict) else flat_keys.add(pre + k)
    keys(got, "")
    assert flat_keys == set(want), (flat_keys ^ set(want))


# ------------------------------------------------------------------------------------------ host statistics (CPU)
def test_export_list_equals_reference():
    import backend.ml.analysis as an
    import backend.ml as ml
    ref = ["pair_correlation", "magnetization_distribution", "autocorrelation_time", "energy_histogram",
           "compare_distributions", "kl_divergence", "symmetric_kl_divergence", "magnetization_kl_divergence",
           "energy_kl_divergence", "wasserstein_distance_1d", "magnetization_wasserstein", "energy_wasserstein",
           "correlation_function_comparison", "comprehensive_comparison", "compute_rdf", "compute_rdf_from_system",
           "rdf_statistics", "compare_rdfs"]          # analysis/__init__.py:23-42 of the reference
    assert sorted(an.__all__) == sorted(ref)
    for name in ref:
        assert callable(getattr(an, name))
    for name in ref[:14]:                             # ml/__init__.py:20-35 re-exports
        assert getattr(ml, name) is getattr(an, name)


def test_kl_and_wasserstein_match_reference():
    from backend.ml.analysis import kl_divergence, symmetric_kl_divergence, wasserstein_distance_1d
    p, q = G["kl_p"], G["kl_q"]
    np.testing.assert_allclose(kl_divergence(p, q), G["kl_pq"], rtol=1e-12)
    np.testing.assert_allclose(kl_divergence(q, p), G["kl_qp"], rtol=1e-12)
    np.testing.assert_allclose(symmetric_kl_divergence(p, q), G["kl_sym"], rtol=1e-12)
    np.testing.assert_allclose(kl_divergence(p, q, epsilon=1e-4), G["kl_eps"], rtol=1e-12)
    assert kl_divergence(p, p) == 0.0
    np.testing.assert_allclose(wasserstein_distance_1d(G["w1"], G["w2"]), G["w_12"], rtol=1e-12)
    assert wasserstein_distance_1d(G["w1"], G["w1"]) == 0.0 == float(G["w_11"])
    # shifting a sample set by c moves it a distance c (the reference weights each interval with the CDF gap at its
    # RIGHT end, a Riemann sum that is off by O(1/n) — kept as is)
    np.testing.assert_allclose(wasserstein_distance_1d(G["w1"], G["w1"] + 0.5), 0.5, rtol=2e-3)


def test_series_statistics_match_reference_on_oracle_series():
    """Histogram, autocorrelation and distance arithmetic, with the magnetisation / energy series taken from the
    pinned oracle (bit-identical to the reference for +/-1 lattices)."""
    from backend.ml.analysis import correlation as c
    for tag in "ab":
        s = G[f"spins_{tag}"].astype(np.float32)
        m = o_ising.magnetization(s)
        e = o_ising.energy(s) / np.float32(s.shape[-1] ** 2)
        centres, h = c._density_histogram(m, 20, (-1, 1))
        np.testing.assert_allclose(centres, G[f"magdist_{tag}.M"], **TOL)
        np.testing.assert_allclose(h, G[f"magdist_{tag}.P_M"], **TOL)
        centres, h = c._density_histogram(e, 15)
        np.testing.assert_allclose(centres, G[f"ehist_{tag}.E"], **TOL)
        np.testing.assert_allclose(h, G[f"ehist_{tag}.P_E"], **TOL)
        np.testing.assert_allclose(c.integrated_autocorrelation(m), G[f"tau_{tag}"], rtol=1e-5)
    assert c.integrated_autocorrelation(np.zeros(5, np.float32)) == 1.0 == float(G["tau_short"])
    assert c.integrated_autocorrelation(np.ones(20, np.float32)) == 1.0 == float(G["tau_const"])
    ma, mb = (o_ising.magnetization(G[f"spins_{t}"].astype(np.float32)) for t in "ab")
    np.testing.assert_allclose(c.wasserstein_distance_1d(ma, mb), G["magw"], rtol=1e-6)
    assert c._correlation_length(np.array([0.0, 0.9, 0.5, 0.2])) == 0
    assert c._correlation_length(np.array([1.0, 0.9, 0.5, 0.2])) == 3
    assert c._correlation_length(np.array([1.0, 0.9])) == 2


def test_rdf_curve_statistics_match_reference():
    from backend.ml.analysis import rdf_statistics, compare_rdfs
    r1, g1, r2, g2 = G["rdf_r1"], G["rdf_g1"], G["rdf_r2"], G["rdf_g2"]
    _check_dict(rdf_statistics(r1, g1), "rdfstat1.", rtol=1e-12)
    _check_dict(rdf_statistics(r2, g2, sigma=1.1), "rdfstat2.", rtol=1e-12)
    _check_dict(rdf_statistics(r1[:10], g1[:10]), "rdfstat_none.", rtol=1e-12)
    _check_dict(rdf_statistics(r1, G["rdf_mono"]), "rdfstat_mono.", rtol=1e-12)
    _check_dict(compare_rdfs(r1, g1, r1, g1 * 0.9 + 0.05), "rdfcmp_same.", rtol=1e-12)
    _check_dict(compare_rdfs(r1, g1, r2, g2), "rdfcmp_diff.", rtol=1e-12)


def test_lattice_functions_raise_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from backend.ml.analysis import magnetization_distribution, comprehensive_comparison
    from backend.ml.systems.ising import IsingModel
    a, b = _sets()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        magnetization_distribution(a)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        comprehensive_comparison(a, b, IsingModel(16))


# ------------------------------------------------------------------------------------------ end to end (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b", "a4", "cont"])
def test_distributions_and_autocorrelation_equal_reference(tag):
    from backend.ml.analysis import magnetization_distribution, energy_histogram, autocorrelation_time
    from backend.ml.systems.ising import IsingModel
    a, b = _sets()
    s = {"a": a, "b": b, "a4": a.unsqueeze(1), "cont": torch.tensor(G["cont"])}[tag]
    tol = dict(rtol=1e-6, atol=1e-9) if tag != "cont" else dict(rtol=2e-4, atol=1e-6)
    _check_dict(magnetization_distribution(s, 20), f"magdist_{tag}.", **tol)
    _check_dict(energy_histogram(s, IsingModel(16), 15), f"ehist_{tag}.", **tol)
    np.testing.assert_allclose(autocorrelation_time(s), G[f"tau_{tag}"], rtol=1e-4)
    with pytest.raises(ValueError, match="Unknown observable"):
        autocorrelation_time(s, observable="energy")


@pytest.mark.gpu
def test_series_are_bit_exact_for_spin_lattices_and_stay_on_device_path():
    from backend.ml.analysis.correlation import magnetization_series, energy_series
    from backend.ml.systems.ising import IsingModel
    a, _ = _sets()
    assert np.array_equal(magnetization_series(a), o_ising.magnetization(a.numpy()))
    assert np.array_equal(magnetization_series(a.cuda()), o_ising.magnetization(a.numpy()))
    e = o_ising.energy(a.numpy()) / np.float32(256)
    assert np.array_equal(energy_series(a.unsqueeze(1), IsingModel(16)), e)
    assert autocorr_short() == 1.0


def autocorr_short():
    from backend.ml.analysis import autocorrelation_time
    return autocorrelation_time(torch.ones(5, 16, 16))


@pytest.mark.gpu
def test_two_sample_comparisons_equal_reference():
    from backend.ml import analysis as an
    from backend.ml.systems.ising import IsingModel
    a, b = _sets()
    model = IsingModel(16)
    _check_dict(an.magnetization_kl_divergence(a, b, 20), "magkl.")
    _check_dict(an.energy_kl_divergence(a, b, model, 15), "ekl.")
    np.testing.assert_allclose(an.magnetization_wasserstein(a, b), G["magw"], rtol=1e-6)
    np.testing.assert_allclose(an.energy_wasserstein(a, b, model), G["ew"], rtol=1e-6)
    np.testing.assert_allclose(an.magnetization_wasserstein(torch.tensor(G["cont"]), a[:40].unsqueeze(1)),
                               G["magw_cont"], rtol=1e-4)
    _check_dict(an.compare_distributions(a, b, model), "cmpdist.", rtol=1e-5, atol=1e-9)
    np.random.seed(42)        # the reference draws pair_correlation's sites from np.random (analysis/correlation.py:25-26)
    _check_dict(an.correlation_function_comparison(a, b), "corrcmp.", rtol=1e-12, atol=1e-15)
    np.random.seed(43)
    _check_dict(an.comprehensive_comparison(a, b, model, 20), "comp.", rtol=1e-5, atol=1e-9)


# ------------------------------------------------------------------------------------------ module surface kept for callers
def test_compute_loss_and_physical_system_kept_for_the_trainer():
    """training/trainer.py:11-14 imports DiffusionProcess and calls compute_loss; systems/__init__.py:1 exports
    PhysicalSystem.  Both are host-side torch code outside the sampling path; values pinned to the reference."""
    from backend.ml.models.diffusion import DiffusionProcess
    from backend.ml.systems import PhysicalSystem
    sp, nz, t = (torch.tensor(G[k]) for k in ("closs_pred", "closs_noise", "closs_t"))
    sp.requires_grad_(True)
    for lt in ("l2", "l1"):
        for w in ("uniform", "snr"):
            loss = DiffusionProcess().compute_loss(sp, nz, t, lt, w)
            np.testing.assert_allclose(loss.item(), G[f"closs_{lt}_{w}"], rtol=1e-6)
    loss.backward()
    assert torch.isfinite(sp.grad).all()
    with pytest.raises(ValueError, match="Unknown loss type"):
        DiffusionProcess().compute_loss(sp, nz, t, "huber")
    with pytest.raises(ValueError, match="Unknown weighting"):
        DiffusionProcess().compute_loss(sp, nz, t, "l2", "min-snr")
    with pytest.raises(TypeError):
        PhysicalSystem()
    assert {"energy", "random_configuration", "energy_per_particle", "to_continuous", "from_continuous",
            "num_particles", "configuration_shape"} == set(PhysicalSystem.__abstractmethods__)
