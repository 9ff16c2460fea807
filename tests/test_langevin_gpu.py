"""GPU parity of the batched Langevin samplers and the RDF histogram through
the C ABI: the reference's own step() outputs with its recorded draws (tests/golden/langevin_rdf.npz),
the pinned oracle on larger batches, and the reference's physics smoke test (test_rdf.py:146-165).
fp32: positions <= 5e-6 absolute, velocities <= 1e-4; RDF counts are integers: exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import langevin as o_lg
from oracle import lj as o_lj

pytestmark = pytest.mark.gpu


def _mods():
    from backend.ml.systems.lennard_jones import LennardJonesSystem
    from backend.ml.samplers.langevin import LangevinDynamics, UnderdampedLangevin
    return LennardJonesSystem, LangevinDynamics, UnderdampedLangevin


def test_mala_vs_reference_golden():
    LJ, MALA, _ = _mods()
    g = load_golden("langevin_rdf.npz")
    n, box, T, dt = g["mala_params"]
    smp = MALA(LJ(int(n), float(box)), temperature=float(T), dt=float(dt), seed=1)
    pos = torch.tensor(g["mala_traj"][0])
    for k in range(6):
        new = smp.step(pos, noise=torch.tensor(g["mala_noise"][k]), uniform=torch.tensor([g["mala_u"][k]], dtype=torch.float32))
        assert new is not pos and not new.is_cuda and new.shape == (int(n), 2)
        np.testing.assert_allclose(new.numpy(), g["mala_traj"][k + 1], rtol=0, atol=5e-6)
        pos = new


def test_baoab_vs_reference_golden():
    LJ, _, BAOAB = _mods()
    g = load_golden("langevin_rdf.npz")
    n, box, T, dt, mass, fr = g["baoab_params"]
    smp = BAOAB(LJ(int(n), float(box)), temperature=float(T), dt=float(dt), mass=float(mass), friction=float(fr), seed=1)
    pos, vel = torch.tensor(g["baoab_pos"][0]), torch.tensor(g["baoab_vel"][0])
    for k in range(5):
        pos, vel = smp.step(pos, vel, noise=torch.tensor(g["baoab_noise"][k]))
        np.testing.assert_allclose(pos.numpy(), g["baoab_pos"][k + 1], rtol=0, atol=5e-6)
        np.testing.assert_allclose(vel.numpy(), g["baoab_vel"][k + 1], rtol=1e-4, atol=1e-4)


def test_mala_batched_vs_oracle_full_size():
    """BASELINE config-5 shape: N = 256 at rho* = 0.8; 32 chains in lock-step, injected draws."""
    LJ, MALA, _ = _mods()
    N, B, T, dt = 256, 32, 1.0, 1e-3
    box = float((N / 0.8) ** 0.5)
    sys_ = LJ(N, box)
    torch.manual_seed(3)
    pos = (sys_.lattice_configuration(B) + 0.05 * torch.randn(B, N, 2)) % box
    noise = torch.randn(3, B, N, 2)
    u = torch.rand(3, B)
    smp = MALA(sys_, temperature=T, dt=dt, seed=5)
    cur = pos.clone()
    p = o_lj.LJParams(N, box)
    ref = pos.numpy().copy()
    n_acc = 0
    for k in range(3):
        cur = smp.step(cur, noise=noise[k], uniform=u[k])
        for b in range(B):
            ref[b], a = o_lg.mala_step(p, ref[b], T, dt, noise[k, b].numpy(), float(u[k, b]))
            n_acc += int(a)
        np.testing.assert_allclose(cur.numpy(), ref, rtol=0, atol=2e-5)
    assert 0 < n_acc <= 3 * B


def test_langevin_api_and_physics():
    """Shapes of sample / trajectory, batching extension, and the reference's RDF smoke test:
    first g(r) peak of an equilibrated LJ liquid sits in (0.9, 2.0) (test_rdf.py:146-165)."""
    LJ, MALA, BAOAB = _mods()
    from backend.ml.analysis.rdf import compute_rdf_from_system
    np.random.seed(0)
    torch.manual_seed(0)
    sys_ = LJ(16, 5.0)
    smp = MALA(sys_, temperature=1.0, dt=0.001)
    s = smp.sample(n_samples=5, n_steps=50, burn_in=300)
    assert s.shape == (5, 16, 2) and not s.is_cuda and torch.isfinite(s).all() and (s >= 0).all() and (s <= 5.0).all()
    r, gr = compute_rdf_from_system(s, sys_, n_bins=50)
    peak = r[int(np.argmax(gr))]
    assert 0.9 < peak < 2.0
    traj = list(smp.sample_with_trajectory(4))
    assert len(traj) == 5 and all(t.shape == (16, 2) for t in traj)
    many = smp.sample(n_samples=2, n_steps=10, burn_in=20, n_chains=8)
    assert many.shape == (8, 2, 16, 2)
    e = sys_.energy_per_particle(many.reshape(-1, 16, 2))
    assert torch.isfinite(e).all() and (e < 0).all()
    ud = BAOAB(sys_, temperature=1.0, dt=0.002)
    su = ud.sample(n_samples=3, n_steps=20, burn_in=100)
    assert su.shape == (3, 16, 2) and torch.isfinite(su).all()
    p1, v1 = ud.step(sys_.lattice_configuration(1).squeeze(0), torch.zeros(16, 2))
    assert p1.shape == (16, 2) and v1.shape == (16, 2)


def test_rdf_exact_vs_reference_and_oracle():
    from backend.ml.analysis.rdf import compute_rdf
    g = load_golden("langevin_rdf.npz")
    r, gr = compute_rdf(torch.tensor(g["rdf_pos"]), 5.0, n_bins=40)
    assert np.array_equal(r, g["rdf_r"]) and np.array_equal(gr, g["rdf_g"])
    r2, gr2 = compute_rdf(torch.tensor(g["rdf_pos"][0]), 5.0, n_bins=25, r_max=2.0)
    assert np.array_equal(r2, g["rdf2_r"]) and np.array_equal(gr2, g["rdf2_g"])
    N, B = 256, 16
    box = float((N / 0.8) ** 0.5)
    torch.manual_seed(2)
    pos = torch.rand(B, N, 2) * box
    r3, g3 = compute_rdf(pos, box, n_bins=100)
    _, rc, gc = o_lg.rdf_histogram(pos.numpy(), box, n_bins=100)
    assert np.array_equal(r3, rc) and np.array_equal(g3, gc)
