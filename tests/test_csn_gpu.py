"""GPU parity of the particle score network through the C ABI: fp32 path, <= 1e-4
relative (north_star) against the reference's own outputs (tests/golden/csn_tiny.npz) and the
pinned torch-fp32 oracle at the configuration-5 shape (N = 256 particles, default widths)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import csn as o_csn

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _sd(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def _from_golden(tag="sd::", **cfg):
    from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
    g = load_golden("csn_tiny.npz")
    net = ContinuousScoreNetwork(**cfg).eval()
    net.load_state_dict({k[len(tag):]: torch.tensor(g[k]) for k in g.files if k.startswith(tag)})
    return net, g


def test_forward_matches_reference_golden():
    net, g = _from_golden(n_particles=12, hidden_dim=32, time_embed_dim=16, pos_embed_dim=16, num_blocks=2, num_heads=4)
    with torch.no_grad():
        out = net(torch.tensor(g["fwd_x"]), torch.tensor(g["fwd_t"]))           # per-sample t
        assert not out.is_cuda and out.shape == (3, 12, 2)                      # CPU in -> CPU out
        assert _rel(out.numpy(), g["fwd_out"]) <= 1e-4
        out = net(torch.tensor(g["fwd_w_x"]), torch.tensor(g["fwd_w_t"]))       # shared t, N != n_particles, |x| > 1
        assert _rel(out.numpy(), g["fwd_w_out"]) <= 1e-4


def test_forward_without_attention_matches_reference_golden():
    net, g = _from_golden("sdna::", n_particles=8, hidden_dim=24, time_embed_dim=8, pos_embed_dim=8, num_blocks=1,
                          use_attention=False)
    with torch.no_grad():
        out = net(torch.tensor(g["na_x"]), torch.tensor(g["na_t"]))
    assert _rel(out.numpy(), g["na_out"]) <= 1e-4


def test_sampler_matches_reference_golden_with_injected_noise():
    from backend.ml.samplers.diffusion import DiffusionSampler
    net, g = _from_golden(n_particles=12, hidden_dim=32, time_embed_dim=16, pos_embed_dim=16, num_blocks=2, num_heads=4)
    N, temp = int(g["smp_meta"][0]), float(g["smp_meta"][1])
    x0, nz = torch.tensor(g["smp_x0"]), torch.tensor(g["smp_noises"])
    out = DiffusionSampler(net, num_steps=N).sample(tuple(x0.shape), temperature=temp, x0=x0, noises=nz)
    assert _rel(out.numpy(), g["smp_out"]) <= 1e-4


@pytest.mark.parametrize("B,N,kw", [
    (3, 256, {}),                                             # configuration-5 shape, default widths (hd = 32)
    (2, 300, dict(hidden_dim=64, num_heads=8)),               # N > one key chunk, hd = 8
    (5, 17, dict(hidden_dim=128, num_heads=2, num_blocks=1)),  # ragged N, hd = 64
    (70, 1, dict(hidden_dim=48, num_heads=3, time_embed_dim=32, pos_embed_dim=32)),  # single particle, hd = 16, > 1 tile
])
def test_forward_matches_oracle_on_shapes(B, N, kw):
    from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
    torch.manual_seed(5)
    net = ContinuousScoreNetwork(N, **kw).eval()
    x = torch.rand(B, N, 2) * 2 - 1
    t = torch.rand(B)
    want = o_csn.csn_forward(_sd(net), x, t, num_heads=net.num_heads).numpy()
    with torch.no_grad():
        got = net(x, t).numpy()
    assert _rel(got, want) <= 1e-4
    with torch.no_grad():
        got1 = net(x, torch.full((B,), 0.25)).numpy()        # all-equal t collapses to one time row
    want1 = o_csn.csn_forward(_sd(net), x, torch.full((B,), 0.25), num_heads=net.num_heads).numpy()
    assert _rel(got1, want1) <= 1e-4


def _rms_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30))


@pytest.mark.parametrize("B,N", [(3, 256), (2, 128), (5, 48), (1, 16), (130, 64)])
def test_bf16_tensor_core_path_within_stated_bound(B, N):
    """tcgen05 path (fc1 / fc2 / qkv / attention / proj in bf16, fp32 accumulate and residual stream).
    Stated bound: rms error <= 2e-2 of the output rms against the fp32 oracle, per forward."""
    from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
    torch.manual_seed(11)
    net = ContinuousScoreNetwork(N).eval()
    x = torch.rand(B, N, 2) * 2 - 1
    t = torch.rand(B)
    want = o_csn.csn_forward(_sd(net), x, t).numpy()
    with torch.no_grad():
        got = net.engine_forward(x, t, precision="bf16").cpu().numpy()
        ref32 = net.engine_forward(x, t, precision="fp32").cpu().numpy()
    assert np.isfinite(got).all()
    assert _rms_rel(got, want) <= 2e-2, _rms_rel(got, want)
    assert _rel(ref32, want) <= 1e-4
    with torch.no_grad():
        got1 = net.engine_forward(x, torch.full((1,), 0.3), precision="bf16").cpu().numpy()
    want1 = o_csn.csn_forward(_sd(net), x, torch.full((B,), 0.3)).numpy()
    assert _rms_rel(got1, want1) <= 2e-2


def test_opt_in_fused_mlp_block_within_stated_bound():
    """DBS_CSN_FUSE_MLP=1 (fc1 -> norm1 -> SiLU -> + time bias -> fc2 -> norm2 -> SiLU -> + x in one tcgen05 launch, measured
    slower than the two launches and therefore off by default) stays inside the bf16 bound; the switch is read once per
    process, so the forward runs in a child process."""
    import os, subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent("""
        import sys, numpy as np, torch
        sys.path[:0] = [%r, %r]
        from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
        from oracle import csn as o_csn
        for B, N in ((3, 256), (130, 64), (1, 16)):
            torch.manual_seed(11)
            net = ContinuousScoreNetwork(N).eval()
            sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
            x, t = torch.rand(B, N, 2) * 2 - 1, torch.rand(B)
            want = o_csn.csn_forward(sd, x, t).numpy().astype(np.float64)
            with torch.no_grad():
                got = net.engine_forward(x, t, precision="bf16").cpu().numpy().astype(np.float64)
            err = float(np.sqrt(((got - want) ** 2).mean()) / np.sqrt((want ** 2).mean()))
            assert np.isfinite(got).all() and err <= 2e-2, (B, N, err)
            print("ok", B, N, err)
    """) % (root, os.path.join(root, "diffusion-boltzmann-sampler_b200"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DBS_CSN_FUSE_MLP="1"), capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 3


def test_bf16_path_rejects_ineligible_shapes():
    from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
    net = ContinuousScoreNetwork(24).eval()
    with torch.no_grad(), pytest.raises(NotImplementedError):
        net.engine_forward(torch.zeros(2, 24, 2), torch.zeros(2), precision="bf16")      # N % 16 != 0
    with torch.no_grad(), pytest.raises(NotImplementedError):
        net.engine_forward(torch.zeros(1, 272, 2), torch.zeros(1), precision="bf16")     # N > 256


def test_bf16_sampler_trajectory_close_to_fp32():
    from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
    from backend.ml.samplers.diffusion import DiffusionSampler
    torch.manual_seed(2)
    N = 64
    net = ContinuousScoreNetwork(N).eval()
    x0, nz = torch.randn(4, N, 2), torch.randn(10, 4, N, 2)
    a = DiffusionSampler(net, num_steps=10, precision="fp32").sample((4, N, 2), x0=x0, noises=nz).numpy()
    b = DiffusionSampler(net, num_steps=10, precision="bf16").sample((4, N, 2), x0=x0, noises=nz).numpy()
    assert _rms_rel(b, a) <= 2e-2


def test_lj_diffusion_sampling_through_the_sampler_api():
    """Config 5's diffusion leg: DiffusionSampler over (B, N, 2) normalised positions, mapped back to the
    box by LennardJonesSystem.from_continuous; same Philox streams for any sharding of the batch."""
    from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
    from backend.ml.samplers.diffusion import DiffusionSampler
    from backend.ml.systems.lennard_jones import LennardJonesSystem
    torch.manual_seed(1)
    N, B = 64, 6
    net = ContinuousScoreNetwork(N, hidden_dim=64, num_blocks=2).eval()
    smp = DiffusionSampler(net, num_steps=8, device="cuda")
    full = smp.sample((B, N, 2), seed=3)
    part = smp.sample((2, N, 2), seed=3, sample_offset=4)
    assert torch.equal(full[4:6], part)
    lj = LennardJonesSystem(N, box_size=float((N / 0.8) ** 0.5))
    pos = lj.from_continuous(full.cpu())
    assert pos.shape == (B, N, 2) and torch.isfinite(lj.energy(pos)).all()
    # injected noise: the engine trajectory equals the oracle's
    x0, nz = torch.randn(2, N, 2), torch.randn(8, 2, N, 2)
    got = smp.sample((2, N, 2), x0=x0, noises=nz).cpu().numpy()
    want = o_csn.sample(_sd(net), (2, N, 2), 8, x0, nz).numpy()
    assert _rel(got, want) <= 1e-4


def test_bad_arguments_raise():
    from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
    net = ContinuousScoreNetwork(8, hidden_dim=32).eval()
    with torch.no_grad():
        with pytest.raises(ValueError):
            net(torch.zeros(2, 8, 3), torch.zeros(2))
        with pytest.raises(ValueError):
            net(torch.zeros(2, 8, 2), torch.zeros(3))
        with pytest.raises(NotImplementedError):
            net.engine_forward(torch.zeros(2, 8, 2), torch.zeros(2), precision="bf16")
