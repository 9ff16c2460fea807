"""Spin-spin correlation function — engine-backed replacement for pair_correlation
(analysis/correlation.py:8-47 of the reference).  The other functions of that file (histograms,
KL, Wasserstein) are O(batch) host post-processing and stay with the reference."""
from typing import Dict, List

import numpy as np
import torch

from dbs_b200 import abi, runtime
from ..systems.ising import IsingModel


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
    """
    x = samples if isinstance(samples, torch.Tensor) else torch.as_tensor(np.asarray(samples))
    n, L, _ = x.shape
    R = L // 2
    if ref_indices is None:
        ref_indices = np.random.choice(L * L, min(100, L * L), replace=False)
    model = IsingModel(L)
    bits = model.pack(x)
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
