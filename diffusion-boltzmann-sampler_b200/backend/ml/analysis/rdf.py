"""Radial distribution function — engine-backed replacement for compute_rdf / compute_rdf_from_system
(analysis/rdf.py:9-109 of the reference) plus that file's O(n_bins) curve statistics (rdf_statistics,
compare_rdfs: host arithmetic on the finished g(r), no lattice or particle data involved)."""
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from dbs_b200 import abi, runtime

__all__ = ("compute_rdf", "compute_rdf_from_system", "rdf_statistics", "compare_rdfs")


def compute_rdf(positions: torch.Tensor, box_size: float, n_bins: int = 100, r_max: Optional[float] = None
                ) -> Tuple[np.ndarray, np.ndarray]:
    """(bin centres, g(r)) with the reference's normalisation (analysis/rdf.py:40-83); the pair
    histogram is counted on the GPU with np.histogram's exact bin rule."""
    if positions.dim() == 2:
        positions = positions.unsqueeze(0)
    n_samples, n_particles, _ = positions.shape
    if r_max is None:
        r_max = box_size / 2
    edges = np.linspace(0, r_max, n_bins + 1)
    centers = (edges[:-1] + edges[1:]) / 2
    dr = edges[1] - edges[0]
    x = runtime.to_device(positions)
    e_dev = torch.from_numpy(edges).to(x.device)
    hist = torch.zeros(n_bins, dtype=torch.int64, device=x.device)
    abi.check(abi.lib.dbs_rdf_histogram(x.data_ptr(), n_samples, n_particles, float(box_size), e_dev.data_ptr(), n_bins,
                                        hist.data_ptr(), runtime.current_stream()), "rdf_histogram")
    h = hist.cpu().numpy()
    total_pairs = n_particles * (n_particles - 1) / 2 * n_samples
    rho = n_particles / (box_size ** 2)
    g_r = h / (total_pairs * (2 * np.pi * centers * dr) * rho + 1e-10)
    return centers, g_r


def compute_rdf_from_system(positions: torch.Tensor, system, n_bins: int = 100, r_max: Optional[float] = None):
    return compute_rdf(positions=positions, box_size=system.box_size, n_bins=n_bins, r_max=r_max)


def _trapezoid(y: np.ndarray, x: np.ndarray) -> float:
    return float(np.sum(0.5 * (y[1:] + y[:-1]) * np.diff(x))) if len(x) > 1 else 0.0


def rdf_statistics(r: np.ndarray, g_r: np.ndarray, sigma: float = 1.0) -> Dict[str, float]:
    """Position and height of the highest g(r) value beyond 0.9 sigma, and the 2-D coordination integral
    2 pi * int r g(r) dr from 0.9 sigma up to the first local minimum after that peak
    (analysis/rdf.py:111-177; the reference's integral carries no density factor and neither does this)."""
    r, g_r = np.asarray(r), np.asarray(g_r)
    keep = r > 0.9 * sigma
    if not keep.any():
        return {"first_peak_r": 0.0, "first_peak_height": 0.0, "coordination_number": 0.0}
    rv, gv = r[keep], g_r[keep]
    peak = int(np.argmax(gv))
    coordination = 0.0
    if peak + 1 < len(gv) - 1:
        inner = np.arange(peak + 1, len(gv) - 1)
        is_min = (gv[inner] < gv[inner - 1]) & (gv[inner] < gv[inner + 1])
        stop = int(inner[is_min][0]) if is_min.any() else len(gv) - 1
        coordination = 2 * np.pi * _trapezoid(rv[:stop + 1] * gv[:stop + 1], rv[:stop + 1])
    return {"first_peak_r": float(rv[peak]), "first_peak_height": float(gv[peak]),
            "coordination_number": float(coordination)}


def compare_rdfs(r1: np.ndarray, g_r1: np.ndarray, r2: np.ndarray, g_r2: np.ndarray) -> Dict[str, float]:
    """Mean squared and maximum difference of two g(r) curves (linearly interpolated onto the overlap of
    their grids when the grids differ) and the distance between their first peaks (analysis/rdf.py:180-226)."""
    r1, g_r1, r2, g_r2 = (np.asarray(v) for v in (r1, g_r1, r2, g_r2))
    if len(r1) == len(r2) and np.allclose(r1, r2):
        a, b = g_r1, g_r2
    else:
        grid = np.linspace(max(r1.min(), r2.min()), min(r1.max(), r2.max()), min(len(r1), len(r2)))
        a, b = np.interp(grid, r1, g_r1), np.interp(grid, r2, g_r2)
    peak1, peak2 = rdf_statistics(r1, g_r1)["first_peak_r"], rdf_statistics(r2, g_r2)["first_peak_r"]
    return {"mse": float(np.mean((a - b) ** 2)), "max_diff": float(np.max(np.abs(a - b))),
            "peak_diff": float(abs(peak1 - peak2))}
