"""GPU parity of the Lennard-Jones pair kernel through the C ABI: reference golden vectors, the
pinned oracle, and the reference's own property tests (backend/tests/test_lennard_jones.py).
fp32: energies <= 1e-5 relative; forces <= 1e-4 of the largest force component."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import lj as o_lj

pytestmark = pytest.mark.gpu


def _sys(*a, **k):
    from backend.ml.systems.lennard_jones import LennardJonesSystem
    return LennardJonesSystem(*a, **k)


def test_vs_reference_golden():
    g = load_golden("lennard_jones.npz")
    s = _sys(16, 5.0)
    pos = torch.tensor(g["lj16_pos"])
    np.testing.assert_allclose(s.energy(pos).numpy(), g["lj16_E"], rtol=1e-5)
    np.testing.assert_allclose(s.forces(pos).numpy(), g["lj16_F"], rtol=1e-4, atol=1e-4 * np.abs(g["lj16_F"]).max())
    n, box, eps, sig, rc = g["lj64_params"]
    s = _sys(int(n), float(box), epsilon=float(eps), sigma=float(sig), cutoff=float(rc))
    pos = torch.tensor(g["lj64_pos"])
    e, f = s.energy_and_forces(pos)
    np.testing.assert_allclose(e.numpy(), g["lj64_E"], rtol=1e-5)
    np.testing.assert_allclose(f.numpy(), g["lj64_F"], rtol=1e-4, atol=1e-4 * np.abs(g["lj64_F"]).max())
    s2 = _sys(2, 5.0)
    np.testing.assert_allclose(s2.energy(torch.tensor(g["ljpair_pos"])).numpy(), g["ljpair_E"], rtol=1e-6)
    np.testing.assert_allclose(s2.forces(torch.tensor(g["ljpair_pos"])).numpy(), g["ljpair_F"], rtol=1e-5, atol=1e-6)
    assert abs(_sys(2, 10.0).energy(torch.tensor(g["ljfar_pos"])).item()) < 1e-6
    assert np.array_equal(s2.minimum_image(torch.tensor(g["lj_minimg_in"])).numpy(), g["lj_minimg_out"])


@pytest.mark.parametrize("N,B", [(256, 64), (100, 5), (1000, 2), (7, 3)])
def test_vs_oracle_full_size(N, B):
    """BASELINE config-5 shape: N = 256 at rho* = 0.8, jittered lattice."""
    box = float((N / 0.8) ** 0.5)
    s = _sys(N, box)
    torch.manual_seed(0)
    pos = (s.lattice_configuration(B) + 0.05 * torch.randn(B, N, 2)) % box
    p = o_lj.LJParams(N, box)
    e, f = s.energy_and_forces(pos.cuda())
    assert e.is_cuda and f.is_cuda
    want_e, want_f = o_lj.energy(p, pos.numpy()), o_lj.forces(p, pos.numpy())
    np.testing.assert_allclose(e.cpu().numpy(), want_e, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(f.cpu().numpy(), want_f, rtol=1e-4, atol=1e-4 * np.abs(want_f).max())


def test_reference_property_tests():
    """backend/tests/test_lennard_jones.py:13-29, 72-108, 157-165 of the reference."""
    s = _sys(16, 5.0)
    lat = s.lattice_configuration(2)
    e = s.energy(lat)
    assert e.shape == (2,) and torch.isfinite(e).all() and (e < 0).all() and torch.allclose(e[0], e[1])
    f = s.forces(lat + 0.03 * torch.randn(2, 16, 2))
    assert f.shape == (2, 16, 2)
    assert torch.allclose(f.sum(dim=1), torch.zeros(2, 2), atol=1e-3)         # Newton's third law
    # force = -dE/dx by central differences (atol = rtol = 1e-3 as in the reference test)
    pos = (s.lattice_configuration(1) + 0.05 * torch.randn(1, 16, 2)).double()
    F = s.forces(pos.float())[0]
    h = 1e-3
    for (i, d) in [(0, 0), (5, 1), (11, 0)]:
        pp, pm = pos.clone(), pos.clone()
        pp[0, i, d] += h
        pm[0, i, d] -= h
        num = -(s.energy(pp.float()).item() - s.energy(pm.float()).item()) / (2 * h)
        assert abs(F[i, d].item() - num) <= 1e-3 + 1e-2 * abs(num)
    assert torch.allclose(s.score(lat, 2.0), s.forces(lat) / 2.0)
    x = s.to_continuous(lat)
    assert torch.allclose(s.from_continuous(x), lat, atol=1e-5)
    assert s.random_configuration(3).shape == (3, 16, 2) and s.density == 16 / 25.0


def test_overlap_mask_edge_cases():
    """Pairs just above the d > 0.1 mask and coincident particles (adversarial set of SURVEY §8d)."""
    s = _sys(3, 6.0)
    pos = torch.tensor([[[1.0, 1.0], [1.0, 1.0], [1.1001, 1.0]],      # coincident pair is masked out
                        [[0.0, 0.0], [0.09, 0.0], [3.0, 3.0]]])       # d < 0.1 masked; third is beyond rc
    p = o_lj.LJParams(3, 6.0)
    np.testing.assert_allclose(s.energy(pos).numpy(), o_lj.energy(p, pos.numpy()), rtol=1e-5)
    np.testing.assert_allclose(s.forces(pos).numpy(), o_lj.forces(p, pos.numpy()), rtol=1e-4, atol=1e-2)
    assert s.energy(pos)[1].item() == 0.0


@pytest.mark.parametrize("N,B,rho", [(256, 64, 0.8), (64, 16, 0.5), (200, 8, 0.8), (1024, 3, 0.6)])
def test_half_precision_prefilter_is_bit_identical(N, B, rho, monkeypatch):
    """lj_pre_kernel puts a conservative fp16 distance prefilter (two partners per instruction, box-wrapped
    coordinates) in front of the exact fp32 test; everything it admits is re-tested exactly and summed in the same
    ascending partner order, so energies and forces equal the plain kernel's bit for bit — for the config-5 lattice
    + jitter set, for uniform-random positions (near-overlaps), for unwrapped positions several boxes away, for a
    configuration beyond the wrap's precision guard (> 16 boxes: every partner becomes a candidate) and with NaNs."""
    s = _sys(N, float(np.sqrt(N / rho)))
    torch.manual_seed(N)
    lat = s.lattice_configuration(B) + 0.05 * torch.randn(B, N, 2)
    uni = torch.rand(B, N, 2) * s.box_size
    shifted = lat + s.box_size * torch.randint(-6, 7, (B, N, 2)).float()
    far = lat.clone(); far[0] += 40.0 * s.box_size; far[1, 3, 0] = 1e9
    nan = lat.clone(); nan[0, 5, 1] = float("nan"); nan[1, 0, 0] = float("inf")
    for pos in (lat, uni, shifted, far, nan):
        monkeypatch.setenv("DBS_LJ_PREFILTER", "0")
        e0, f0 = s.energy_and_forces(pos.cuda())
        monkeypatch.setenv("DBS_LJ_PREFILTER", "1")
        e1, f1 = s.energy_and_forces(pos.cuda())
        assert torch.equal(e0.cpu().view(torch.int32), e1.cpu().view(torch.int32))
        assert torch.equal(f0.cpu().view(torch.int32), f1.cpu().view(torch.int32))
