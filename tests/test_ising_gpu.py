"""GPU parity of the Ising path (through the C ABI) against the pinned oracle and the golden
vectors of the unmodified reference.  Integer/bit work: everything here is bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ising as o

pytestmark = pytest.mark.gpu


def _mods():
    from backend.ml.systems.ising import IsingModel
    from backend.ml.samplers.mcmc import MetropolisHastings
    return IsingModel, MetropolisHastings


def _words(t):
    return t.cpu().numpy().view(np.uint32)


def _sweep_cases():
    keys = ["L", "J", "h", "T", "n_sweeps", "seed", "start", "r", "final"]
    out = []
    for name in ("ising_sweeps.npz", "ising_odd.npz"):   # even: checkerboard kernel; odd: three-colour kernel
        g = load_golden(name)
        out += [{k: g[f"c{i}_{k}"] for k in keys} for i in range(int(g["n_cases"]))]
    return out


@pytest.mark.parametrize("L", [4, 8, 12, 16, 32, 48, 64, 96, 128, 5, 9, 33, 65, 127])
def test_hot_start_pack_unpack_bit_exact(L):
    IsingModel, _ = _mods()
    m = IsingModel(L)
    bits = m.random_packed(5, seed=99, chain_offset=3)
    want = o.init_bits_words(99, np.arange(3, 8), L)
    assert np.array_equal(_words(bits), want)
    spins = m.unpack(bits)
    assert np.array_equal(spins.cpu().numpy(), o.unpack(want, L).astype(np.float32))
    assert np.array_equal(_words(m.pack(spins)), want)


@pytest.mark.parametrize("case", _sweep_cases(), ids=lambda c: f"L{int(c['L'])}_J{float(c['J'])}_h{float(c['h'])}_T{float(c['T'])}")
def test_sweeps_bit_exact_vs_reference_golden(case):
    """Both draw sources — the reference's uniforms injected, and in-kernel Philox — must land on
    the lattice the reference's own step() produced (tests/golden/ising_sweeps.npz)."""
    IsingModel, MetropolisHastings = _mods()
    L, seed, ns = int(case["L"]), int(case["seed"]), int(case["n_sweeps"])
    m = IsingModel(L, J=float(case["J"]), h=float(case["h"]))
    start = torch.tensor(case["start"].astype(np.float32))
    n = start.shape[0]
    # (a) injection port
    mh = MetropolisHastings(m, float(case["T"]), seed=seed)
    bits = m.pack(start)
    uni = torch.from_numpy(case["r"].view(np.int32).copy()).cuda()
    mh.sweeps_packed(bits, ns, uniforms=uni, sweep_offset=0)
    assert np.array_equal(m.unpack(bits).cpu().numpy(), case["final"].astype(np.float32))
    # (b) Philox in the kernel, all sweeps in one launch
    bits = m.pack(start)
    mh.sweeps_packed(bits, ns, sweep_offset=0)
    assert np.array_equal(m.unpack(bits).cpu().numpy(), case["final"].astype(np.float32))
    # (c) one launch per sweep continues the same stream
    mh2 = MetropolisHastings(m, float(case["T"]), seed=seed)
    bits = m.pack(start)
    for _ in range(ns):
        mh2.sweeps_packed(bits, 1)
    assert np.array_equal(m.unpack(bits).cpu().numpy(), case["final"].astype(np.float32))
    assert n == 2


@pytest.mark.parametrize("L,n_chains,n_sweeps", [(16, 64, 20), (64, 33, 6), (128, 3, 3), (24, 7, 5), (40, 5, 4), (8, 130, 9),
                                                   (9, 40, 6), (33, 9, 4), (65, 3, 3), (127, 2, 2)])
def test_sweeps_vs_oracle_many_chains_per_chain_temperature(L, n_chains, n_sweeps):
    IsingModel, MetropolisHastings = _mods()
    m = IsingModel(L, J=1.0, h=0.05)
    temps = np.linspace(1.5, 3.5, n_chains)
    mh = MetropolisHastings(m, temps, seed=4242)
    bits = m.random_packed(n_chains, seed=4242, chain_offset=10)
    start = o.unpack(_words(bits), L)
    mh.sweeps_packed(bits, n_sweeps, chain_offset=10, sweep_offset=7)
    want = o.run_sweeps(start.copy(), 1.0, 0.05, temps, n_sweeps, seed=4242, chain_ids=np.arange(10, 10 + n_chains),
                        sweep_offset=7)
    assert np.array_equal(o.unpack(_words(bits), L), want)
    assert not np.array_equal(start, want)


@pytest.mark.parametrize("L", [8, 16, 64])
def test_observables_bit_exact_vs_reference_golden(L):
    IsingModel, _ = _mods()
    g = load_golden("ising_static.npz")
    for tag, (J, h) in {"a": (1.0, 0.0), "b": (0.7, -0.45)}.items():
        m = IsingModel(L, J=J, h=h)
        s = torch.tensor(g[f"obs_L{L}_{tag}_spins"])
        assert np.array_equal(m.energy(s).numpy(), g[f"obs_L{L}_{tag}_E"])
        assert np.array_equal(m.magnetization(s).numpy(), g[f"obs_L{L}_{tag}_M"])
        assert np.array_equal(m.energy_per_spin(s).numpy(), g[f"obs_L{L}_{tag}_Eps"])
        e, mag = m.packed_observables(m.pack(s))
        assert np.array_equal(e.cpu().numpy(), g[f"obs_L{L}_{tag}_E"])
        assert np.array_equal(mag.cpu().numpy(), g[f"obs_L{L}_{tag}_M"])


def test_observables_known_answers_and_shapes():
    """backend/tests/test_ising.py:11-36, 55-79 of the reference."""
    IsingModel, _ = _mods()
    m = IsingModel(8)
    up = torch.ones(8, 8)
    cb = torch.tensor([[1.0 if (i + j) % 2 == 0 else -1.0 for j in range(8)] for i in range(8)])
    assert m.energy(up).item() == -128.0 and m.energy(-up).item() == -128.0 and m.energy(cb).item() == 128.0
    assert m.magnetization(up).item() == 1.0 and m.magnetization(-up).item() == -1.0 and m.magnetization(cb).item() == 0.0
    assert m.local_energy_diff(up, 0, 0).item() == 8.0
    batch = torch.stack([up, cb, -up])
    assert m.energy(batch).shape == (3,) and m.energy(batch).tolist() == [-128.0, 128.0, -128.0]
    assert m.energy(batch.cuda()).is_cuda
    r = m.random_configuration(4)
    assert r.shape == (4, 8, 8) and not r.is_cuda and set(r.unique().tolist()) <= {-1.0, 1.0}
    g = load_golden("ising_static.npz")
    mc = IsingModel(16, J=1.0, h=0.2)
    np.testing.assert_allclose(mc.energy(torch.tensor(g["obs_cont_spins"])).numpy(), g["obs_cont_E"], rtol=2e-6, atol=1e-4)
    np.testing.assert_allclose(mc.magnetization(torch.tensor(g["obs_cont_spins"])).numpy(), g["obs_cont_M"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("L", [8, 16])
def test_pair_correlation_equals_reference(L):
    from backend.ml.analysis.correlation import pair_correlation
    g = load_golden("correlation.npz")
    s = torch.tensor(g[f"corr_L{L}_samples"])
    res = pair_correlation(s, ref_indices=g[f"corr_L{L}_ref_indices"])
    assert res["r"] == list(g[f"corr_L{L}_r"])
    assert np.array_equal(np.array(res["C_r"]), g[f"corr_L{L}_C"])
    # same numpy seed -> same reference sites as the reference draws (analysis/correlation.py:25-26)
    np.random.seed(100 + L)
    assert np.array_equal(np.array(pair_correlation(s)["C_r"]), g[f"corr_L{L}_C"])


@pytest.mark.parametrize("L", [9, 33, 127])
def test_odd_lattice_observables_and_correlation_equal_reference(L):
    """Odd lattices (lattice_size is any integer in 8..128 at the API, routes/sampling.py:26): E, M from fp32 and from
    packed state, and C(r) with the reference's own sites, against tests/golden/ising_odd.npz."""
    from backend.ml.analysis.correlation import pair_correlation
    IsingModel, MetropolisHastings = _mods()
    g = load_golden("ising_odd.npz")
    m = IsingModel(L, J=0.7, h=-0.45)
    s = torch.tensor(g[f"obs_L{L}_spins"].astype(np.float32))
    assert np.array_equal(m.energy(s).numpy(), g[f"obs_L{L}_E"])
    assert np.array_equal(m.magnetization(s).numpy(), g[f"obs_L{L}_M"])
    e, mag = m.packed_observables(m.pack(s))
    assert np.array_equal(e.cpu().numpy(), g[f"obs_L{L}_E"]) and np.array_equal(mag.cpu().numpy(), g[f"obs_L{L}_M"])
    if L == 127:
        L = 15
    c = torch.tensor(g[f"corr_L{L}_samples"].astype(np.float32))
    assert np.array_equal(np.array(pair_correlation(c, ref_indices=g[f"corr_L{L}_ref_indices"])["C_r"]), g[f"corr_L{L}_C"])
    np.random.seed(200 + L)
    assert np.array_equal(np.array(pair_correlation(c)["C_r"]), g[f"corr_L{L}_C"])
    assert np.array_equal(np.array(pair_correlation(c, ref_indices="all")["C_r"]),
                          np.array(o.pair_correlation(c.numpy(), None)["C_r"]))
    # the reference API on an odd lattice: shapes, +/-1 values, ordering at low temperature
    mh = MetropolisHastings(IsingModel(L), 1.0, seed=3)
    out = mh.sample(n_samples=3, n_sweeps=2, burn_in=60, initial=torch.ones(L, L))
    assert out.shape == (3, L, L) and set(out.unique().tolist()) <= {-1.0, 1.0} and out.mean() > 0.9
    x = torch.ones(L, L)
    assert mh.sweep(x) is x


def test_pair_correlation_of_real_valued_lattices_follows_the_reference_product():
    """Real-valued samples (undiscretised diffusion output) are correlated as numbers, as analysis/correlation.py:42 does
    (`ref_spin * sample[ni, nj]`), not binarised: against a direct evaluation of the reference's loops in float64, for
    given reference sites, every site, and a lattice size the packed engine does not take; and the real-valued route
    agrees with the bit-packed kernel on +/-1 input."""
    from backend.ml.analysis import correlation as corr
    rng = np.random.RandomState(3)

    def direct(x, ref):
        n, L, _ = x.shape
        R = L // 2
        c, k = np.zeros(R), np.zeros(R)
        for s in x.astype(np.float64):
            for idx in ref:
                i, j = divmod(int(idx), L)
                for di in range(-R, R + 1):
                    for dj in range(-R, R + 1):
                        r = int(np.sqrt(di * di + dj * dj))
                        if 0 < r < R:
                            c[r] += s[i, j] * s[(i + di) % L, (j + dj) % L]
                            k[r] += 1
        return np.divide(c, k, where=k > 0, out=np.zeros_like(c))

    for L in (8, 6, 3):
        x = rng.randn(5, L, L).astype(np.float32)
        ref = rng.choice(L * L, min(7, L * L), replace=False)
        np.testing.assert_allclose(corr.pair_correlation(torch.tensor(x), ref_indices=ref)["C_r"], direct(x, ref), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(corr.pair_correlation(torch.tensor(x), ref_indices="all")["C_r"], direct(x, np.arange(L * L)),
                                   rtol=1e-9, atol=1e-9)
    s = torch.tensor(rng.choice([-1.0, 1.0], size=(6, 16, 16)).astype(np.float32))
    ref = rng.choice(256, 30, replace=False)
    packed = corr.pair_correlation(s, ref_indices=ref)["C_r"]
    real = corr._pair_correlation_real(s.cuda(), ref)["C_r"]
    np.testing.assert_allclose(real, packed, rtol=0, atol=1e-12)


@pytest.mark.parametrize("L", [12, 32, 64, 128])
def test_pair_correlation_all_sites_vs_oracle(L):
    from backend.ml.analysis.correlation import pair_correlation
    rng = np.random.RandomState(L)
    s = rng.choice([-1.0, 1.0], size=(3, L, L)).astype(np.float32)
    s[0] = 1.0
    got = pair_correlation(torch.tensor(s), ref_indices="all")
    want = o.pair_correlation(s, None)
    assert np.array_equal(np.array(got["C_r"]), np.array(want["C_r"]))
    idx = rng.choice(L * L, 37, replace=False)
    assert np.array_equal(np.array(pair_correlation(torch.tensor(s), ref_indices=idx)["C_r"]),
                          np.array(o.pair_correlation(s, idx)["C_r"]))


def _corr_sums(bits, group, n_groups, method, measure=False, J=1.0, h=0.0):
    """C(r) sums (and, with measure, the fp64 observable block) of packed 64 x 64 chains with a chosen evaluation method."""
    from dbs_b200 import abi, runtime
    prev = abi.lib.dbs_ising_set_correlation_method(method)
    try:
        n = bits.shape[0]
        acc = torch.zeros((n_groups, 32), dtype=torch.int64, device=bits.device)
        gp = group.data_ptr() if group is not None else 0
        if measure:
            block = torch.zeros((n_groups, 6), dtype=torch.float64, device=bits.device)
            abi.check(abi.lib.dbs_ising_measure(bits.data_ptr(), 64, n, J, h, gp, block.data_ptr(), 6, acc.data_ptr(),
                                                runtime.current_stream()))
            return acc.cpu(), block.cpu()
        abi.check(abi.lib.dbs_ising_correlation(bits.data_ptr(), 64, n, 0, gp, acc.data_ptr(), runtime.current_stream()))
        return acc.cpu(), None
    finally:
        abi.lib.dbs_ising_set_correlation_method(prev)


@pytest.mark.parametrize("n_chains", [1, 2, 3, 4, 5, 7, 64, 1001])
def test_spectral_correlation_equals_the_displacement_kernel_bit_for_bit(n_chains):
    """The |FFT2|^2 -> inverse FFT2 evaluation of the C(r) sums at 64 x 64 (csrc/ising_corr_fft.cu) against the XOR / POPC
    kernels: same int64 sums, bin 0 untouched — for hot, ordered, striped, strongly magnetised and equilibrated lattices,
    with no group array, one group per chain, contiguous groups and groups that change inside a quad of chains."""
    IsingModel, MetropolisHastings = _mods()
    m = IsingModel(64)
    rng = np.random.RandomState(n_chains)
    s = rng.choice([-1.0, 1.0], size=(n_chains, 64, 64)).astype(np.float32)
    s[0] = 1.0
    if n_chains > 2:
        s[2] = np.where((np.arange(64)[:, None] // 8 + np.arange(64)[None, :] // 4) % 2, 1.0, -1.0)
    if n_chains > 4:
        s[4] = np.where(rng.rand(64, 64) < 0.93, 1.0, -1.0)
    bits = m.pack(torch.tensor(s)).cuda()
    if n_chains >= 64:
        MetropolisHastings(m, 2.269, seed=9).sweeps_packed(bits[8:], 30, sweep_offset=0)
    for name, group, ng in (("none", None, 1),
                            ("per chain", torch.arange(n_chains, dtype=torch.int32), n_chains),
                            ("runs of 6", (torch.arange(n_chains) // 6).to(torch.int32), (n_chains + 5) // 6),
                            ("alternating", (torch.arange(n_chains) % 3).to(torch.int32), 3)):
        gd = group.cuda() if group is not None else None
        want, _ = _corr_sums(bits, gd, ng, 1)
        got, _ = _corr_sums(bits, gd, ng, 2)
        assert torch.equal(got, want), name
        assert int(got[:, 0].abs().sum()) == 0
    # against the oracle's direct loops on a few chains
    k = min(n_chains, 3)
    want = o.pair_correlation(s[:k], None)
    got, _ = _corr_sums(bits[:k], None, 1, 2)
    from backend.ml.analysis.correlation import _displacement_counts
    cnt = _displacement_counts(64).astype(np.float64) * 4096 * k
    c = np.divide(got[0].numpy().astype(np.float64), cnt, where=cnt > 0, out=np.zeros(32))
    assert np.array_equal(c, np.array(want["C_r"]))


def test_spectral_correlation_at_the_full_config3_ensemble():
    """BASELINE config 3 at full size: 65 536 chains 64 x 64, 64 temperatures in [1.5, 3.5] x 1024 chains, equilibrated for
    50 sweeps — the spectral and the displacement kernels return the same int64 C(r) sums for every temperature, bin 0 stays
    zero, and the sums obey the size-independent properties of a correlation function of +/-1 spins: |C(r)| <= 1 and
    C(1) of the coldest ensemble close to 1, of the hottest well below it."""
    IsingModel, MetropolisHastings = _mods()
    from backend.ml.analysis.correlation import _displacement_counts
    m = IsingModel(64)
    n = 65536
    temps = np.repeat(np.linspace(1.5, 3.5, 64), 1024)
    group = torch.tensor(np.repeat(np.arange(64), 1024).astype(np.int32)).cuda()
    bits = m.random_packed(n, seed=11)
    MetropolisHastings(m, temps, seed=11).sweeps_packed(bits, 50, sweep_offset=0)
    want, wb = _corr_sums(bits, group, 64, 1, measure=True)
    got, gb = _corr_sums(bits, group, 64, 2, measure=True)
    assert torch.equal(got, want)
    assert int(got[:, 0].abs().sum()) == 0
    assert torch.allclose(gb, wb, rtol=1e-12, atol=1e-12)
    cnt = _displacement_counts(64).astype(np.float64) * 4096 * 1024
    c = got.numpy().astype(np.float64)[:, 1:] / cnt[None, 1:]
    assert np.all(np.abs(c) <= 1.0 + 1e-12)
    assert c[0, 0] > 0.9 and c[-1, 0] < 0.5 and c[0, 0] > c[-1, 0]


def test_spectral_measure_block_equals_the_displacement_kernel():
    """dbs_ising_measure through the spectral kernel: identical C(r) sums and the same fp64 observable block (same
    per-chain E and M; the block is a sum of fp64 atomics, compared to rounding of the summation order)."""
    IsingModel, MetropolisHastings = _mods()
    m = IsingModel(64, J=0.8, h=0.0)
    n = 515
    bits = m.random_packed(n, seed=5)
    temps = np.repeat(np.array([1.8, 2.269, 3.1]), 172)[:n]
    MetropolisHastings(m, temps, seed=5).sweeps_packed(bits, 40, sweep_offset=0)
    group = torch.tensor(np.repeat(np.arange(3), 172)[:n].astype(np.int32)).cuda()
    want, wb = _corr_sums(bits, group, 3, 1, measure=True, J=0.8)
    got, gb = _corr_sums(bits, group, 3, 2, measure=True, J=0.8)
    assert torch.equal(got, want)
    assert torch.allclose(gb, wb, rtol=1e-12, atol=1e-12)
    assert torch.equal(gb[:, 0], torch.tensor([172.0, 172.0, 171.0], dtype=torch.float64))


def test_mcmc_api_contract():
    """Reference contract tests (backend/tests/test_mcmc.py:22-71, 256-349): beta, step flips at
    most one spin and mutates/returns its argument, sample/trajectory shapes."""
    IsingModel, MetropolisHastings = _mods()
    np.random.seed(42)
    m = IsingModel(8)
    assert MetropolisHastings(m, 2.0).beta == 0.5 and MetropolisHastings(m, 0.0).beta == float("inf")
    mh = MetropolisHastings(m, 2.27)
    x = m.random_configuration(1).squeeze(0)
    for _ in range(50):
        before = x.clone()
        y = mh.step(x)
        assert y is x and y.shape == (8, 8) and int((y != before).sum()) <= 1
        assert set(y.unique().tolist()) <= {-1.0, 1.0}
    cold = MetropolisHastings(m, 0.01)
    z = m.random_configuration(1).squeeze(0)
    e_prev = m.energy(z).item()
    for _ in range(300):
        z = cold.step(z)
        e = m.energy(z).item()
        assert e <= e_prev + 1e-6
        e_prev = e
    s = mh.sample(n_samples=5, n_sweeps=2, burn_in=10)
    assert s.shape == (5, 8, 8) and not s.is_cuda and set(s.unique().tolist()) <= {-1.0, 1.0}
    init = torch.ones(8, 8)
    s0 = mh.sample(n_samples=1, n_sweeps=0, initial=init, burn_in=0)
    assert torch.equal(s0[0], init) and torch.equal(init, torch.ones(8, 8))
    frames = list(mh.sample_with_trajectory(6, initial=init))
    assert len(frames) == 7 and torch.equal(frames[0], init) and all(f.shape == (8, 8) for f in frames)
    sw = mh.sweep(init.clone())
    assert sw.shape == (8, 8)
    many = mh.sample(n_samples=3, n_sweeps=1, burn_in=5, n_chains=16)
    assert many.shape == (16, 3, 8, 8)


def test_mcmc_physics_envelopes():
    """Statistical envelopes of backend/tests/test_mcmc.py:95-108, 355-488 at 16x16, 256 chains."""
    IsingModel, MetropolisHastings = _mods()
    m = IsingModel(16)
    def run(T, init=None):
        mh = MetropolisHastings(m, T, seed=7)
        x = mh.sample(n_samples=4, n_sweeps=10, burn_in=400, n_chains=256, initial=init, return_device=True)
        flat = x.reshape(-1, 16, 16)
        return m.magnetization(flat).abs().mean().item(), m.energy_per_spin(flat).mean().item()
    mlow, elow = run(1.0, torch.ones(16, 16))
    assert mlow > 0.95 and elow < -1.9
    mhigh, ehigh = run(5.0)
    assert mhigh < 0.2 and -0.6 < ehigh < 0.0
    mc, ec = run(2.269)
    assert 0.3 < mc < 0.9 and -1.6 < ec < -1.2
    # exact 2-D Ising energy per spin at T=2.0 (Onsager): -1.7455...; finite-size + MC error < 1%
    _, e2 = run(2.0, torch.ones(16, 16))
    assert abs(e2 - (-1.7456)) < 0.02


@pytest.mark.parametrize("fixture", ["ising_ensemble.npz", "ising_ensemble_odd.npz"])
def test_ensemble_observables_vs_reference_chain(fixture):
    """north_star: "ensemble observables agree within statistical error".  tests/golden/ising_ensemble.npz holds per-chain
    time averages of E/N, |M|, M^2, M^4 from the UNMODIFIED reference's random-site Metropolis chain
    (oracle/make_golden_ensemble.py: BASELINE config 1 as-is — 16x16, T = 2.269, 64 chains, 1000 sweeps — plus
    T = 1.5 and T = 3.5).  The engine's checkerboard chains follow the same protocol (start, burn-in, one
    measurement per sweep) with 4096 chains; different dynamics, same Boltzmann distribution: the ensemble means agree
    within 3.5 sigma of the combined standard error (chains are independent, so the chain-to-chain spread is the
    error whatever the autocorrelation time), and so does the Binder cumulant."""
    IsingModel, MetropolisHastings = _mods()
    g = load_golden(fixture)       # ..._odd: a 9x9 lattice, swept in three colour classes
    L = int(g["L"])
    m = IsingModel(L)
    n_eng = 4096
    for i, (T, n_ref, hot, burn, meas) in enumerate(g["runs"]):
        ref = g[f"chains_{i}"]
        burn, meas = int(burn), int(meas)
        if hot:
            bits = m.random_packed(n_eng, seed=100 + i)
        else:
            bits = m.pack(torch.ones(n_eng, L, L))
        mh = MetropolisHastings(m, float(T), seed=200 + i)
        mh.sweeps_packed(bits, burn, sweep_offset=0)
        acc = torch.zeros(4, n_eng, dtype=torch.float64, device=bits.device)
        for s in range(meas):
            mh.sweeps_packed(bits, 1, sweep_offset=burn + s)
            e, mag = m.packed_observables(bits)
            mag = mag.double()
            acc[0] += e.double() / (L * L); acc[1] += mag.abs(); acc[2] += mag ** 2; acc[3] += mag ** 4
        eng = (acc / meas).T.cpu().numpy()                       # [chain][4]
        mean_r, sem_r = ref.mean(0), ref.std(0, ddof=1) / np.sqrt(len(ref))
        mean_e, sem_e = eng.mean(0), eng.std(0, ddof=1) / np.sqrt(n_eng)
        z = np.abs(mean_e - mean_r) / np.sqrt(sem_r ** 2 + sem_e ** 2 + 1e-24)
        assert (z <= 3.5).all(), (float(T), mean_e.tolist(), mean_r.tolist(), z.tolist())

        def binder(ch):
            return 1.0 - ch[:, 3].mean() / (3.0 * ch[:, 2].mean() ** 2)
        # jackknife error of the reference's Binder cumulant over its chains
        jk = np.array([binder(np.delete(ref, k, axis=0)) for k in range(len(ref))])
        err = np.sqrt((len(ref) - 1) * ((jk - jk.mean()) ** 2).mean())
        assert abs(binder(eng) - binder(ref)) <= 3.5 * err + 2e-3, (float(T), binder(eng), binder(ref), err)


def test_full_size_sharding_invariance():
    """BASELINE config-3 shape (64x64, per-chain temperatures): chains carry global indices, so a
    rank's shard equals the same rows of the single-GPU run for any world size."""
    IsingModel, MetropolisHastings = _mods()
    L, n = 64, 4096
    m = IsingModel(L)
    temps = np.repeat(np.linspace(1.5, 3.5, 64), n // 64)
    full = m.random_packed(n, seed=5)
    MetropolisHastings(m, temps, seed=5).sweeps_packed(full, 10, sweep_offset=0)
    for rank, world in [(1, 2), (5, 8)]:
        lo, hi = rank * n // world, (rank + 1) * n // world
        part = m.random_packed(hi - lo, seed=5, chain_offset=lo)
        MetropolisHastings(m, temps[lo:hi], seed=5).sweeps_packed(part, 10, chain_offset=lo, sweep_offset=0)
        assert torch.equal(part, full[lo:hi])
    e, mag = m.packed_observables(full)
    spins = m.unpack(full)
    assert torch.equal(e.cpu(), m.energy(spins).cpu()) and torch.equal(mag.cpu(), m.magnetization(spins).cpu())


def test_temperature_sweep_driver_sharding_and_physics():
    """BASELINE config-3 driver at reduced size: per-temperature observables, and the sum of two
    emulated ranks' accumulator blocks equals the single-rank block exactly (integer / fp64 sums of
    the same chains: global chain ids make the shards world-size invariant)."""
    from dbs_b200.temperature_sweep import run_temperature_sweep
    temps = np.array([1.5, 2.0, 2.269, 3.0, 3.5])
    kw = dict(lattice_size=16, temperatures=temps, chains_per_temperature=64, burn_in=300, n_measure=100,
              measure_every=10, seed=77)
    full = run_temperature_sweep(**kw)
    assert full["energy_per_spin"].shape == (5,) and full["C_r"].shape == (5, 8)
    assert np.all(np.diff(full["abs_magnetization"]) < 0)           # order decreases with temperature
    assert np.all(np.diff(full["energy_per_spin"]) > 0)
    assert full["abs_magnetization"][0] > 0.95 and full["abs_magnetization"][-1] < 0.35
    assert abs(full["energy_per_spin"][1] - (-1.7456)) < 0.03       # Onsager at T = 2.0 (finite size 16x16)
    assert np.all(full["C_r"][0, 1:] > 0.9) and abs(full["C_r"][-1, 7]) < 0.1
    assert np.allclose(full["count"], 64 * 10)
    parts = [run_temperature_sweep(rank=r, world_size=2, reduce=False, **kw) for r in range(2)]
    acc = parts[0]["acc"] + parts[1]["acc"]
    corr = parts[0]["corr_sums"] + parts[1]["corr_sums"]
    assert torch.equal(corr, full["corr_sums"])
    assert torch.allclose(acc, full["acc"], rtol=1e-12, atol=1e-9)
