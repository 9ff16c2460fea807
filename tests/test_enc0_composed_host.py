"""CPU check of the algebra behind enc0_composed_kernel (csrc/unet.cu): conv1 of the first encoder block composed with the
init conv (models/score_network.py:117-151, 203, 238 — both linear, no activation in between) is a function of the
one-channel input alone,
    y1[p][co] = b1[co] + sum_t in(p + t) * ( Bc[co][t] + sum_s Wc[co][t][s] * x~[p + t + s] ),
which the kernel evaluates in two forms: the 81 masked products for tiles on the image's outer ring and, where every in() is 1,
ONE 5x5 kernel W5[u] = sum_{t + s = u} Wc[t][s] with the summed bias.  This restates both forms with plain loops in float64
and compares them with the two convolutions they replace (torch, float64); the kernel itself is pinned on the GPU by
tests/test_diffusion_gpu.py::test_composed_first_layer_tile_forms_match_the_two_launch_path."""
import numpy as np
import torch
import torch.nn.functional as F


def _compose(W1, Wi, bi):
    """Wc[co][t][s] = sum_ci W1[co][ci][t] Wi[ci][s],  Bc[co][t] = sum_ci W1[co][ci][t] bi[ci]   (t, s = 3 kh + kw)"""
    co, ci = W1.shape[:2]
    w1 = W1.reshape(co, ci, 9)
    wi = Wi.reshape(ci, 9)
    return np.einsum("oct,cs->ots", w1, wi), np.einsum("oct,c->ot", w1, bi)


def test_masked_and_collapsed_forms_equal_the_two_convolutions():
    rng = np.random.default_rng(5)
    C, CO, H, W = 6, 5, 9, 12
    Wi, bi = rng.standard_normal((C, 1, 3, 3)), rng.standard_normal(C)
    W1, b1 = rng.standard_normal((CO, C, 3, 3)), rng.standard_normal(CO)
    x = rng.standard_normal((H, W))
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    h = F.conv2d(t64(x)[None, None], t64(Wi), t64(bi), padding=1)
    want = F.conv2d(h, t64(W1), t64(b1), padding=1)[0].numpy()            # [CO][H][W]

    Wc, Bc = _compose(W1, Wi, bi)
    xp = np.zeros((H + 4, W + 4))
    xp[2:-2, 2:-2] = x                                                     # x~: two-pixel zero halo
    inside = lambda i, j: 0 <= i < H and 0 <= j < W
    masked = np.zeros((CO, H, W))
    for i in range(H):
        for j in range(W):
            acc = b1.copy()
            for t in range(9):
                di, dj = t // 3 - 1, t % 3 - 1
                if not inside(i + di, j + dj):                             # conv1 pads the init conv's OUTPUT with zeros
                    continue
                acc = acc + Bc[:, t]
                for s in range(9):
                    ei, ej = s // 3 - 1, s % 3 - 1
                    acc = acc + Wc[:, t, s] * xp[i + di + ei + 2, j + dj + ej + 2]
            masked[:, i, j] = acc
    assert np.abs(masked - want).max() <= 1e-12 * np.abs(want).max()

    W5 = np.zeros((CO, 5, 5))
    for t in range(9):
        for s in range(9):
            W5[:, t // 3 + s // 3, t % 3 + s % 3] += Wc[:, t, s]
    btot = b1 + Bc.sum(1)
    collapsed = F.conv2d(t64(xp)[None, None], t64(W5)[:, None], t64(btot))[0].numpy()   # valid conv over the padded input
    ring = np.ones((H, W), dtype=bool)
    ring[1:-1, 1:-1] = False
    err = np.abs(collapsed - want)
    assert err[:, ~ring].max() <= 1e-12 * np.abs(want).max()              # interior pixels: the 5x5 form is exact
    assert err[:, ring].max() > 1e-3                                       # outer ring: it is not — hence the masked form there
