"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/dbs_b200.h declares, and refuses to compute without a GPU (no silent fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "dbs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dbs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dbs_b200 import abi
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(abi.lib, s), f"{s} declared in include/dbs_b200.h but not exported"
        assert s in abi.PROTOTYPES, f"{s} has no ctypes prototype"
    assert set(abi.PROTOTYPES) == set(syms)
    assert abi.lib.dbs_abi_version() == 1000 and abi.lib.dbs_built_sm() == 100


def test_argument_validation_needs_no_gpu():
    from dbs_b200 import abi
    assert abi.lib.dbs_ising_words_per_chain(64) == 128
    assert abi.lib.dbs_ising_words_per_chain(16) == 16
    assert abi.lib.dbs_ising_words_per_chain(7) == 7     # odd L: three-colour sweep kernel
    assert abi.lib.dbs_ising_words_per_chain(33) == 66
    assert abi.lib.dbs_ising_words_per_chain(3) < 0 and abi.lib.dbs_ising_words_per_chain(129) < 0
    rc = abi.lib.dbs_ising_sweeps(0, 130, 1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0)
    assert rc == abi.ERR_INVALID and "unsupported" in abi.last_error()
    # C(r) evaluation method: a process-wide switch that returns the previous value (no device work)
    assert abi.lib.dbs_ising_set_correlation_method(1) == 0
    assert abi.lib.dbs_ising_set_correlation_method(0) == 1
    assert abi.lib.dbs_ising_set_correlation_method(7) == abi.ERR_INVALID and "method" in abi.last_error()
    with pytest.raises(ValueError):
        abi.check(rc, "ising_sweeps")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from backend.ml.systems.ising import IsingModel
    from backend.ml.samplers.mcmc import MetropolisHastings
    from backend.ml.systems.lennard_jones import LennardJonesSystem
    from backend.ml.models.score_network import ScoreNetwork
    from backend.ml.samplers.diffusion import DiffusionSampler
    m = IsingModel(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.energy(torch.ones(8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MetropolisHastings(m, 2.0).sweep(torch.ones(8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LennardJonesSystem(4, 5.0).energy(torch.rand(1, 4, 2))
    net = ScoreNetwork(1, 8, 16, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DiffusionSampler(net, num_steps=2).sample((1, 1, 8, 8))


def test_host_side_tables_match_oracle():
    """The product's acceptance table builder (host logic) against the pinned oracle's."""
    import numpy as np
    from backend.ml.systems.ising import acceptance_table
    from oracle import ising as o
    for J, h, T in [(1.0, 0.0, 2.269), (1.0, 0.0, 0.0), (0.7, -0.45, 1.1), (-1.0, 0.3, 5.0), (1.0, 0.0, 1e-3),
                    (1.0, 0.0, 1e12)]:
        assert np.array_equal(acceptance_table(J, h, T), o.acceptance_table(J, h, T)), (J, h, T)


def test_module_surface_and_checkpoint_names():
    from backend.ml.models import ScoreNetwork, SinusoidalTimeEmbedding, ConvBlock, SelfAttention, DiffusionProcess
    from conftest import load_golden
    net = ScoreNetwork()
    assert sum(p.numel() for p in net.parameters()) == 5672609
    tiny = ScoreNetwork(1, 8, 16, 2)
    g = load_golden("unet_tiny.npz")
    ref_keys = sorted(k[4:] for k in g.files if k.startswith("sd::"))
    assert sorted(tiny.state_dict().keys()) == ref_keys
    for k in ref_keys:
        assert tuple(tiny.state_dict()[k].shape) == g["sd::" + k].shape
    tiny.load_state_dict({k: torch.tensor(g["sd::" + k]) for k in ref_keys})
    # the differentiable definition equals the reference forward (autograd route, training only)
    x, t = torch.tensor(g["fwd_x"]), torch.tensor(g["fwd_t"])
    out = tiny._autograd_forward(x, t)
    assert torch.allclose(out, torch.tensor(g["fwd_out"]), rtol=1e-5, atol=1e-6)
    out.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in tiny.parameters())
    emb = SinusoidalTimeEmbedding(16)(torch.tensor(g["temb_t"]))
    assert torch.equal(emb, torch.tensor(g["temb_out"]))
    assert ConvBlock(8, 16, 16)(torch.randn(2, 8, 4, 4), torch.randn(2, 16)).shape == (2, 16, 4, 4)
    assert SelfAttention(16)(torch.randn(1, 16, 4, 4)).shape == (1, 16, 4, 4)
    # sigma table is bit-identical to the reference's per-step scalars
    gs = load_golden("diffusion_scalar.npz")
    assert DiffusionProcess().sigma_table(100) == list(gs["sigma_N100"])
    assert DiffusionProcess(0.05, 12.0).sigma_table(50) == list(gs["sigma_alt_N50"])


def test_plan_release_is_deferred_while_the_thread_captures(monkeypatch):
    """runtime.release_plan: a plan finaliser that runs on a thread in mid-capture must not call cudaFree there; the
    handle is parked and freed by the next release / plan build outside a capture (capture state simulated here)."""
    from dbs_b200 import runtime
    freed = []
    monkeypatch.setattr(runtime, "_capturing", lambda: True)
    runtime.release_plan(freed.append, 11)
    runtime.release_plan(freed.append, 12)
    runtime.flush_released()                       # still capturing: nothing may be freed
    assert freed == [] and len(runtime._parked) == 2
    monkeypatch.setattr(runtime, "_capturing", lambda: False)
    runtime.release_plan(freed.append, 13)         # parked handles go first, then this one
    assert freed == [11, 12, 13] and runtime._parked == []
    runtime.flush_released()
    assert freed == [11, 12, 13]
