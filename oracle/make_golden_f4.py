#!/usr/bin/env python
"""Goldens for the predictor-corrector and ODE samplers, recorded from the UNMODIFIED reference
(imported from /root/reference) with its torch.randn / randn_like draws captured in call order.

Run in the build container only:  python oracle/make_golden_f4.py   ->  tests/golden/samplers_f4.npz
The network is the tiny ScoreNetwork of tests/golden/unet_tiny.npz (same seed; checked below)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DBS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.dont_write_bytecode = True

import torch  # noqa: E402

from backend.ml.models.score_network import ScoreNetwork  # noqa: E402
from backend.ml.samplers.diffusion import DiffusionSampler  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def recorded(fn):
    """Run fn() with torch.randn / randn_like recording every draw."""
    draws = []
    real_randn, real_randn_like = torch.randn, torch.randn_like

    def rec_randn(*a, **k):
        v = real_randn(*a, **k)
        draws.append(v.clone())
        return v

    def rec_randn_like(*a, **k):
        v = real_randn_like(*a, **k)
        draws.append(v.clone())
        return v

    torch.randn, torch.randn_like = rec_randn, rec_randn_like
    try:
        out = fn()
    finally:
        torch.randn, torch.randn_like = real_randn, real_randn_like
    return out, draws


def main():
    torch.manual_seed(42)
    net = ScoreNetwork(in_channels=1, base_channels=8, time_embed_dim=16, num_blocks=2).eval()
    tiny = np.load(os.path.join(OUT, "unet_tiny.npz"))
    for k, v in net.state_dict().items():
        assert np.array_equal(tiny["sd::" + k], v.numpy()), k      # same weights as the forward / sampler goldens
    g = {}
    shape = (2, 1, 8, 8)

    # predictor-corrector: N = 6, two corrector steps, snr 0.2, temperature 0.9
    N, nc, snr, temp = 6, 2, 0.2, 0.9
    smp = DiffusionSampler(net, num_steps=N)
    torch.manual_seed(11)
    out, draws = recorded(lambda: smp.sample_predictor_corrector(shape, temperature=temp, n_corrector_steps=nc,
                                                                 corrector_snr=snr))
    # call order: x0, then per step i = N..2: predictor noise, nc corrector noises; step 1 draws nothing
    assert len(draws) == 1 + (N - 1) * (1 + nc), len(draws)
    g["pc_x0"] = draws[0].numpy()
    pred, corr = [], []
    pos = 1
    for _ in range(N - 1):
        pred.append(draws[pos].numpy()); pos += 1
        corr.append(np.stack([draws[pos + c].numpy() for c in range(nc)])); pos += nc
    g["pc_noises"] = np.stack(pred)            # [N-1][shape]
    g["pc_corrector_noises"] = np.stack(corr)  # [N-1][nc][shape]
    g["pc_out"] = out.numpy()
    g["pc_meta"] = np.array([N, nc, snr, temp])

    # one corrector step, default snr, N = 5 (a second, independent case)
    N2 = 5
    smp2 = DiffusionSampler(net, num_steps=N2)
    torch.manual_seed(12)
    out2, draws2 = recorded(lambda: smp2.sample_predictor_corrector(shape))
    assert len(draws2) == 1 + (N2 - 1) * 2
    g["pc1_x0"] = draws2[0].numpy()
    g["pc1_noises"] = np.stack([draws2[1 + 2 * k].numpy() for k in range(N2 - 1)])
    g["pc1_corrector_noises"] = np.stack([draws2[2 + 2 * k].numpy()[None] for k in range(N2 - 1)])
    g["pc1_out"] = out2.numpy()
    g["pc1_meta"] = np.array([N2, 1, 0.16, 1.0])

    # ODE integrators: N = 8, one recorded start each
    N3 = 8
    smp3 = DiffusionSampler(net, num_steps=N3)
    for method in ("euler", "heun", "rk4"):
        torch.manual_seed(20)
        out3, draws3 = recorded(lambda: smp3.sample_ode(shape, method=method))
        assert len(draws3) == 1
        g[f"ode_{method}_x0"] = draws3[0].numpy()
        g[f"ode_{method}_out"] = out3.numpy()
    g["ode_meta"] = np.array([N3])
    np.savez_compressed(os.path.join(OUT, "samplers_f4.npz"), **g)
    print("wrote samplers_f4.npz:", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
