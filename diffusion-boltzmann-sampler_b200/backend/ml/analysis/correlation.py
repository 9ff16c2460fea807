"""Ising ensemble analysis — engine-backed replacement for the whole of the reference's
analysis/correlation.py (file:line citations per function).

Everything that touches a lattice runs in libdbs_b200.so: C(r) in `dbs_ising_correlation`, the
per-sample magnetisation and energy series in `dbs_ising_observables_f32` (through `IsingModel`).
What remains on the host is what the reference also leaves there: histograms, KL / Wasserstein
distances and the autocorrelation sum over a series of at most n_samples scalars.  Those helpers
take plain arrays, so they are importable (and tested) without a GPU; every function that is
given lattices raises without one."""
from typing import Any, Dict, List

import numpy as np
import torch

from dbs_b200 import abi, runtime
from ..systems.ising import IsingModel

__all__ = ("pair_correlation", "magnetization_distribution", "autocorrelation_time", "energy_histogram",
           "compare_distributions", "kl_divergence", "symmetric_kl_divergence", "magnetization_kl_divergence",
           "energy_kl_divergence", "wasserstein_distance_1d", "magnetization_wasserstein", "energy_wasserstein",
           "correlation_function_comparison", "comprehensive_comparison")



def _displacement_counts(L: int) -> np.ndarray:
    """Number of displacements (di, dj) in [-R, R]^2 that fall in bin r = int(sqrt(di^2+dj^2)),
    0 < r < R (analysis/correlation.py:34-37)."""
    R = L // 2
    d = np.arange(-R, R + 1)
    rr = np.sqrt(d[:, None] ** 2 + d[None, :] ** 2).astype(np.int64)
    cnt = np.zeros(R, dtype=np.int64)
    for r in range(1, R):
        cnt[r] = int((rr == r).sum())
    return cnt


def pair_correlation(samples, ref_indices=None) -> Dict[str, List[float]]:
    """C(r) = <s_0 s_r> binned by r = floor(|d|), r < L/2; C[0] = 0.

    ref_indices: None reproduces the reference — min(100, L*L) reference sites drawn once per
    call with np.random.choice (analysis/correlation.py:25-26); "all" uses every site (the same
    estimator with less noise); or an explicit list of flattened site indices.
    +/-1 lattices (4 <= L <= 128) go through the bit-packed XOR / popcount kernel (integer-exact sums); lattices with any
    other value are correlated as real numbers (_pair_correlation_real), as the reference's product does.
    """
    x = samples if isinstance(samples, torch.Tensor) else torch.as_tensor(np.asarray(samples))
    n, L, _ = x.shape
    R = L // 2
    if ref_indices is None:
        ref_indices = np.random.choice(L * L, min(100, L * L), replace=False)
    xd = runtime.to_device(x)
    if abi.lib.dbs_ising_words_per_chain(L) < 0 or not bool(((xd == 1) | (xd == -1)).all()):
        # real-valued samples (undiscretised diffusion output, the (B, 1, L, L) sets of comprehensive_comparison) and
        # lattice sizes the packed engine does not take: the
        # reference multiplies the values themselves (analysis/correlation.py:42); the bit-packed kernel would
        # binarise them.  Same sums on the GPU from the circular cross-correlation of the masked with the full lattice.
        return _pair_correlation_real(xd, ref_indices)
    model = IsingModel(L)
    bits = model.pack(xd)
    W = bits.shape[2]
    mask_dev = None
    if isinstance(ref_indices, str):
        if ref_indices != "all":
            raise ValueError("ref_indices must be None, 'all' or a sequence of site indices")
        n_ref = L * L
    else:
        idx = np.asarray(ref_indices, dtype=np.int64)
        mask = np.zeros((L, W), dtype=np.uint32)
        for k in idx:
            i, j = divmod(int(k), L)
            mask[i, j // 32] |= np.uint32(1 << (j % 32))
        n_ref = int(sum(bin(int(v)).count("1") for v in mask.ravel()))
        mask_dev = torch.from_numpy(mask.view(np.int32).copy()).to(bits.device)
    acc = torch.zeros(R, dtype=torch.int64, device=bits.device)
    abi.check(abi.lib.dbs_ising_correlation(bits.data_ptr(), L, n, runtime.ptr(mask_dev), 0, acc.data_ptr(),
                                            runtime.current_stream()), "ising_correlation")
    sums = acc.cpu().numpy().astype(np.float64)
    counts = (_displacement_counts(L) * n_ref * n).astype(np.float64)
    c = np.divide(sums, counts, where=counts > 0, out=np.zeros_like(sums))
    return {"r": list(range(R)), "C_r": c.tolist()}


def _pair_correlation_real(xd: torch.Tensor, ref_indices) -> Dict[str, List[float]]:
    """pair_correlation for lattices of arbitrary real values, on the GPU in fp64:
    sum over reference sites x and samples of s(x) s(x + d) = [corr(mask * s, s)](d), evaluated for every displacement at
    once through the FFT of the periodic lattice, then binned by r = int(|d|) over the window [-R, R]^2 exactly as the
    reference's loops do (a displacement and its negative are distinct entries of the window; at even L the offsets
    -R and +R reach the same site and are both counted, as in the reference)."""
    n, L, _ = xd.shape
    R = L // 2
    s = xd.to(torch.float64)
    mask = torch.zeros(L * L, dtype=torch.float64, device=xd.device)
    if isinstance(ref_indices, str):
        if ref_indices != "all":
            raise ValueError("ref_indices must be None, 'all' or a sequence of site indices")
        mask[:] = 1.0
    else:
        mask[torch.as_tensor(np.unique(np.asarray(ref_indices, dtype=np.int64)), device=xd.device)] = 1.0
    n_ref = int(mask.sum().item())
    ms = s * mask.reshape(1, L, L)
    # cc[d] = sum_x ms(x) s(x + d)  (circular), summed over the samples
    cc = torch.fft.irfft2(torch.conj(torch.fft.rfft2(ms)) * torch.fft.rfft2(s), s=(L, L)).sum(0)
    d = torch.arange(-R, R + 1, device=xd.device)
    rr = torch.sqrt((d[:, None] ** 2 + d[None, :] ** 2).to(torch.float64)).to(torch.int64)
    win = cc[(d % L)[:, None], (d % L)[None, :]]
    sums = torch.zeros(R, dtype=torch.float64, device=xd.device)
    keep = (rr > 0) & (rr < R)
    sums.index_add_(0, rr[keep], win[keep])
    counts = (_displacement_counts(L) * n_ref * n).astype(np.float64)
    sm = sums.cpu().numpy()
    c = np.divide(sm, counts, where=counts > 0, out=np.zeros_like(sm))
    return {"r": list(range(R)), "C_r": c.tolist()}


# ------------------------------------------------------------------------------------------------
# per-sample series (GPU) and the scalar statistics over them (host)
def _lattices(samples) -> torch.Tensor:
    """(n, L, L) or (n, C, L, L) tensor as given; array-likes are wrapped."""
    x = samples if isinstance(samples, torch.Tensor) else torch.as_tensor(np.asarray(samples))
    if x.dim() not in (3, 4):
        raise ValueError(f"expected (n, L, L) or (n, C, L, L) samples, got {tuple(x.shape)}")
    return x


def magnetization_series(samples) -> np.ndarray:
    """Mean spin of every sample as an fp32 numpy vector — the `samples.mean(dim=...)` that opens
    magnetization_distribution / autocorrelation_time / magnetization_wasserstein / compare_distributions
    (analysis/correlation.py:62-66, 88-92, 301-307, 421-426), computed by the engine's observable kernel."""
    x = _lattices(samples)
    m = IsingModel(x.shape[-1]).magnetization(x)
    if x.dim() == 4:                      # mean over (C, H, W) = mean over C of the per-channel means
        m = m.mean(dim=1)
    return m.detach().cpu().numpy()


def energy_series(samples, ising_model) -> np.ndarray:
    """Energy per spin of every sample (analysis/correlation.py:136-141, 329-339)."""
    x = _lattices(samples)
    if x.dim() == 4:
        x = x.squeeze(1)
    return ising_model.energy_per_spin(x).detach().cpu().numpy()


def _density_histogram(values: np.ndarray, n_bins: int, value_range=None):
    h, edges = np.histogram(values, bins=n_bins, density=True, range=value_range)
    return 0.5 * (edges[1:] + edges[:-1]), h


def magnetization_distribution(samples, n_bins: int = 50) -> Dict[str, List[float]]:
    """P(M) on [-1, 1] as a density histogram (analysis/correlation.py:50-72)."""
    centres, h = _density_histogram(magnetization_series(samples), n_bins, (-1, 1))
    return {"M": centres.tolist(), "P_M": h.tolist()}


def energy_histogram(samples, ising_model, n_bins: int = 50) -> Dict[str, List[float]]:
    """P(E/N) over the sampled range (analysis/correlation.py:122-146)."""
    centres, h = _density_histogram(energy_series(samples, ising_model), n_bins)
    return {"E": centres.tolist(), "P_E": h.tolist()}


def integrated_autocorrelation(series: np.ndarray) -> float:
    """tau_int = 1/2 + sum of the normalised autocorrelation over lags 1..99 up to its first negative
    value; 1.0 for fewer than 10 points or a constant series (analysis/correlation.py:94-119)."""
    obs = np.asarray(series)
    n = obs.shape[0]
    if n < 10:
        return 1.0
    var = obs.var()
    if var < 1e-10:
        return 1.0
    d = obs - obs.mean()
    rho = np.correlate(d, d, mode="full")[n - 1:] / (var * np.arange(n, 0, -1))
    tau = 0.5
    for lag in range(1, min(n, 100)):
        if rho[lag] < 0:
            break
        tau += rho[lag]
    return float(tau)


def autocorrelation_time(samples, observable: str = "magnetization") -> float:
    """Integrated autocorrelation time of a time-ordered sample sequence (analysis/correlation.py:75-119)."""
    if observable != "magnetization":
        raise ValueError(f"Unknown observable: {observable}")
    return integrated_autocorrelation(magnetization_series(samples))


# ------------------------------------------------------------------------------------------------
# distances between histograms / sample sets (host, arrays only)
def kl_divergence(p, q, epsilon: float = 1e-10) -> float:
    """D_KL(P || Q) of two histograms with the reference's double normalisation
    (analysis/correlation.py:149-180)."""
    def prob(v):
        v = np.asarray(v, dtype=np.float64)
        v = v / (v.sum() + epsilon) + epsilon
        return v / v.sum()
    p, q = prob(p), prob(q)
    return float(np.sum(p * np.log(p / q)))


def symmetric_kl_divergence(p, q, epsilon: float = 1e-10) -> float:
    """(D_KL(P||Q) + D_KL(Q||P)) / 2 (analysis/correlation.py:183-198)."""
    return 0.5 * (kl_divergence(p, q, epsilon) + kl_divergence(q, p, epsilon))


def _kl_pair(p, q) -> Dict[str, float]:
    p, q = np.array(p), np.array(q)
    return {"kl_divergence": kl_divergence(p, q), "symmetric_kl_divergence": symmetric_kl_divergence(p, q)}


def magnetization_kl_divergence(samples1, samples2, n_bins: int = 50) -> Dict[str, float]:
    """analysis/correlation.py:201-225."""
    return _kl_pair(magnetization_distribution(samples1, n_bins)["P_M"],
                    magnetization_distribution(samples2, n_bins)["P_M"])


def energy_kl_divergence(samples1, samples2, ising_model, n_bins: int = 50) -> Dict[str, float]:
    """analysis/correlation.py:228-254."""
    return _kl_pair(energy_histogram(samples1, ising_model, n_bins)["P_E"],
                    energy_histogram(samples2, ising_model, n_bins)["P_E"])


def wasserstein_distance_1d(samples1, samples2) -> float:
    """W_1 = integral |F - G| between two empirical CDFs, evaluated on the union of the sample values
    (analysis/correlation.py:257-285)."""
    a = np.sort(np.asarray(samples1).ravel())
    b = np.sort(np.asarray(samples2).ravel())
    grid = np.unique(np.concatenate([a, b]))
    fa = np.searchsorted(a, grid, side="right") / len(a)
    fb = np.searchsorted(b, grid, side="right") / len(b)
    widths = np.diff(grid, prepend=grid[0])
    return float(np.sum(np.abs(fa - fb) * widths))


def magnetization_wasserstein(samples1, samples2) -> float:
    """analysis/correlation.py:288-309."""
    return wasserstein_distance_1d(magnetization_series(samples1), magnetization_series(samples2))


def energy_wasserstein(samples1, samples2, ising_model) -> float:
    """analysis/correlation.py:312-338."""
    return wasserstein_distance_1d(energy_series(samples1, ising_model), energy_series(samples2, ising_model))


def _correlation_length(c: np.ndarray) -> int:
    """First r with C(r) < 1/e, len(c) if none (analysis/correlation.py:371-377)."""
    below = np.nonzero(c < 1.0 / np.e)[0]
    return int(below[0]) if below.size else len(c)


def correlation_function_comparison(samples1, samples2) -> Dict[str, Any]:
    """C(r) of both sets, their RMSE / max difference and the ratio of the 1/e lengths
    (analysis/correlation.py:341-392).  4-D inputs are read as (n, 1, L, L)."""
    def lattices3(s):
        x = _lattices(s)
        return x.squeeze(1) if x.dim() == 4 else x
    corr1, corr2 = pair_correlation(lattices3(samples1)), pair_correlation(lattices3(samples2))
    c1, c2 = np.array(corr1["C_r"]), np.array(corr2["C_r"])
    xi1, xi2 = _correlation_length(c1), _correlation_length(c2)
    return {
        "r": corr1["r"], "C1_r": corr1["C_r"], "C2_r": corr2["C_r"],
        "rmse": float(np.sqrt(np.mean((c1 - c2) ** 2))),
        "max_diff": float(np.max(np.abs(c1 - c2))),
        "correlation_length_1": xi1, "correlation_length_2": xi2,
        "correlation_length_ratio": xi2 / xi1 if xi1 > 0 else float("inf"),
    }


def compare_distributions(samples1, samples2, ising_model) -> Dict[str, float]:
    """Mean / unbiased variance of the two magnetisation series (analysis/correlation.py:395-433; the
    histograms the reference builds there are unused and not rebuilt)."""
    m1 = torch.from_numpy(magnetization_series(samples1))
    m2 = torch.from_numpy(magnetization_series(samples2))
    mean1, mean2, var1, var2 = m1.mean().item(), m2.mean().item(), m1.var().item(), m2.var().item()
    return {
        "mag_mean_diff": abs(mean1 - mean2), "mag_var_diff": abs(var1 - var2),
        "samples1_mean_mag": mean1, "samples2_mean_mag": mean2,
        "samples1_var_mag": var1, "samples2_var_mag": var2,
    }


def comprehensive_comparison(samples1, samples2, ising_model, n_bins: int = 50) -> Dict[str, Any]:
    """The report `/sample/compare` returns (analysis/correlation.py:436-491; routes/sampling.py:271-300).
    Each per-sample series is computed once on the GPU and shared by the histogram, KL and Wasserstein
    entries (the reference recomputes it per entry)."""
    mags = [magnetization_series(s) for s in (samples1, samples2)]
    ens = [energy_series(s, ising_model) for s in (samples1, samples2)]
    t1, t2 = torch.from_numpy(mags[0]), torch.from_numpy(mags[1])
    basic = {
        "mag_mean_diff": abs(t1.mean().item() - t2.mean().item()),
        "mag_var_diff": abs(t1.var().item() - t2.var().item()),
        "samples1_mean_mag": t1.mean().item(), "samples2_mean_mag": t2.mean().item(),
        "samples1_var_mag": t1.var().item(), "samples2_var_mag": t2.var().item(),
    }
    mag_kl = _kl_pair(*[_density_histogram(m, n_bins, (-1, 1))[1] for m in mags])
    energy_kl = _kl_pair(*[_density_histogram(e, n_bins)[1] for e in ens])
    mag_w = wasserstein_distance_1d(*mags)
    energy_w = wasserstein_distance_1d(*ens)
    corr = correlation_function_comparison(samples1, samples2)
    return {
        "basic_statistics": basic,
        "kl_divergence": {"magnetization": mag_kl, "energy": energy_kl},
        "wasserstein": {"magnetization": mag_w, "energy": energy_w},
        "correlation": {k: corr[k] for k in ("rmse", "max_diff", "correlation_length_ratio")},
        "summary": {
            "mag_kl": mag_kl["symmetric_kl_divergence"], "energy_kl": energy_kl["symmetric_kl_divergence"],
            "mag_wasserstein": mag_w, "energy_wasserstein": energy_w, "correlation_rmse": corr["rmse"],
        },
    }
