"""Radial distribution function — engine-backed replacement for compute_rdf / compute_rdf_from_system
(analysis/rdf.py:9-109 of the reference).  rdf_statistics / compare_rdfs of that file are O(n_bins)
host post-processing and stay with the reference."""
from typing import Optional, Tuple

import numpy as np
import torch

from dbs_b200 import abi, runtime


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
