"""Edge cases at the reference-facing API (all through the C ABI): empty batches, single items, degenerate
parameters, dtypes.  The reference's tensor code returns empty tensors for empty inputs; so must the engine."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_empty_batches_return_empty_results():
    from backend.ml.models.score_network import ScoreNetwork
    from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
    from backend.ml.samplers.diffusion import DiffusionSampler
    from backend.ml.samplers.mcmc import MetropolisHastings
    from backend.ml.systems.ising import IsingModel
    from backend.ml.systems.lennard_jones import LennardJonesSystem
    m = IsingModel(16)
    assert tuple(m.energy(torch.empty(0, 16, 16)).shape) == (0,)
    assert tuple(m.magnetization(torch.empty(0, 16, 16)).shape) == (0,)
    assert tuple(m.random_configuration(0).shape) == (0, 16, 16)
    mh = MetropolisHastings(m, 2.0, seed=1)
    assert tuple(mh.sample(2, n_chains=0).shape) == (0, 2, 16, 16)
    net = ScoreNetwork(1, 16, 32, 2).eval()
    with torch.no_grad():
        assert tuple(net(torch.randn(0, 1, 16, 16), torch.rand(0)).shape) == (0, 1, 16, 16)
    for precision in ("fp32", "bf16"):
        smp = DiffusionSampler(net, num_steps=4, precision=precision)
        assert tuple(smp.sample((0, 1, 16, 16)).shape) == (0, 1, 16, 16)
    lj = LennardJonesSystem(16, 5.0)
    assert tuple(lj.energy(torch.empty(0, 16, 2)).shape) == (0,)
    assert tuple(lj.forces(torch.empty(0, 16, 2)).shape) == (0, 16, 2)
    csn = ContinuousScoreNetwork(16).eval()
    with torch.no_grad():
        assert tuple(csn(torch.rand(0, 16, 2), torch.rand(0)).shape) == (0, 16, 2)


def test_abi_accepts_null_pointers_for_empty_batches():
    from dbs_b200 import abi
    lib = abi.lib
    assert lib.dbs_ising_init_packed(0, 16, 0, 1, 0, 0) == 0
    assert lib.dbs_ising_pack(0, 0, 16, 0, 0) == 0 and lib.dbs_ising_unpack(0, 0, 16, 0, 0) == 0
    assert lib.dbs_ising_sweeps(0, 16, 0, 5, 0, 1, 0, 0, 0, 0, 0, 0, 0) == 0
    assert lib.dbs_ising_observables(0, 16, 0, 1.0, 0.0, 0, 0, 0) == 0
    assert lib.dbs_ising_observables_f32(0, 16, 0, 1.0, 0.0, 0, 0, 0) == 0
    assert lib.dbs_lj_energy_forces(0, 0, 16, 5.0, 1.0, 1.0, 2.5, 0.0, 0, 0, 0) == 0
    assert lib.dbs_em_step(0, 0, 0, 0, 256, 1.0, 0.01, 1.0, 1.0, 1, 0, 0, 1, 0) == 0
    # ... but a non-empty batch with a null pointer is still an error
    assert lib.dbs_ising_observables_f32(0, 16, 3, 1.0, 0.0, 0, 0, 0) == abi.ERR_INVALID


def test_single_items_and_degenerate_parameters():
    from backend.ml.analysis import pair_correlation, compute_rdf
    from backend.ml.models.score_network import ScoreNetwork
    from backend.ml.samplers.diffusion import DiffusionSampler
    from backend.ml.samplers.mcmc import MetropolisHastings
    from backend.ml.systems.ising import IsingModel
    from backend.ml.systems.lennard_jones import LennardJonesSystem
    from oracle import unet as o_unet
    m = IsingModel(16)
    assert m.energy(torch.ones(16, 16)).shape == torch.Size([]) and m.energy(torch.ones(16, 16)).item() == -512.0
    assert tuple(m.magnetization(torch.ones(2, 3, 16, 16)).shape) == (2, 3)
    mh = MetropolisHastings(m, 2.0, seed=1)
    first = mh.sample(3, n_sweeps=0, burn_in=0, initial=torch.ones(16, 16))     # no sweeps at all: the initial state
    assert torch.equal(first, torch.ones(3, 16, 16))
    assert len(list(mh.sample_with_trajectory(0))) == 1
    x64 = torch.ones(16, 16, dtype=torch.float64)
    assert mh.sweep(x64) is x64 and x64.dtype == torch.float64 and set(x64.unique().tolist()) <= {-1.0, 1.0}
    cold = MetropolisHastings(m, 0.0, seed=2)                                       # beta = inf: energy never rises
    s = m.random_configuration(4, seed=5)
    e0 = m.energy(s)
    for _ in range(5):
        cold.sweep(s)
        e1 = m.energy(s)
        assert (e1 <= e0).all()
        e0 = e1
    assert len(pair_correlation(torch.ones(1, 16, 16))["C_r"]) == 8
    # U-Net: one sample, a 4x4 lattice (1x1 bottleneck), float64 input, one shared t for a batch
    torch.manual_seed(0)
    net = ScoreNetwork(1, 16, 32, 2).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        for shape in ((1, 1, 16, 16), (2, 1, 4, 4), (3, 1, 8, 16)):
            x, t = torch.randn(shape), torch.rand(shape[0])
            want = o_unet.unet_forward(sd, x, t).numpy()
            assert np.abs(net(x, t).numpy() - want).max() <= 1e-4 * np.abs(want).max()
            assert np.abs(net(x.double(), t).numpy() - want).max() <= 1e-4 * np.abs(want).max()
    smp = DiffusionSampler(net, num_steps=5)
    assert tuple(smp.sample((1, 1, 16, 16)).shape) == (1, 1, 16, 16)
    out = smp.generate_batch(5, lattice_size=12, batch_size=2)                      # ragged last batch, fix-up lattice
    assert tuple(out.shape) == (5, 12, 12) and set(out.unique().tolist()) <= {-1.0, 1.0}
    lj = LennardJonesSystem(16, 5.0)
    assert lj.energy(lj.lattice_configuration(1)[0]).shape == torch.Size([])
    assert LennardJonesSystem(1, 5.0).forces(torch.zeros(1, 1, 2)).abs().max().item() == 0.0
    assert compute_rdf(lj.lattice_configuration(1)[0], 5.0, n_bins=10)[1].shape == (10,)
