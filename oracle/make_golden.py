#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/make_golden.py
The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it, the CUDA kernels.
Nothing from the reference is copied: only its inputs and outputs are stored.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DBS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.dont_write_bytecode = True

import torch  # noqa: E402

import backend.ml.samplers.mcmc as ref_mcmc  # noqa: E402
from backend.ml.systems.ising import IsingModel  # noqa: E402
from backend.ml.systems.lennard_jones import LennardJonesSystem  # noqa: E402
from backend.ml.models.score_network import ScoreNetwork, SinusoidalTimeEmbedding  # noqa: E402
from backend.ml.models.diffusion import DiffusionProcess  # noqa: E402
from backend.ml.models.score_network_continuous import ContinuousScoreNetwork  # noqa: E402
from backend.ml.models.noise_schedule import get_schedule as ref_get_schedule  # noqa: E402
from backend.ml.samplers.diffusion import DiffusionSampler  # noqa: E402
from backend.ml.analysis.correlation import pair_correlation  # noqa: E402

from oracle import ising as o_ising  # noqa: E402  (only for the Philox draw layout / hot start)

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


class _Shim:
    """Stands in for np.random inside the reference's mcmc.py only."""

    def __init__(self):
        self.q = []
        self.u = 0.0

    def randint(self, n):
        return self.q.pop(0)

    def random(self):
        return self.u


def ising_checkerboard_cases():
    """Drive the reference's own MetropolisHastings.step over colour 0 then colour 1 with the
    uniforms the engine's Philox stream assigns to each site."""
    cases = []
    specs = [
        # L, J, h, T, n_sweeps, seed
        (8, 1.0, 0.0, 2.269, 4, 11),
        (16, 1.0, 0.0, 2.269, 3, 1234),
        (16, 1.0, 0.0, 0.5, 3, 5),
        (16, 1.0, 0.0, 5.0, 2, 6),
        (16, 1.0, 0.0, 0.0, 2, 7),       # T = 0: beta = inf, never accept uphill
        (16, 1.0, 0.3, 2.0, 3, 8),       # external field: 10 distinct dE classes
        (16, -1.0, 0.0, 1.5, 2, 9),      # antiferromagnet
        (16, 0.7, -0.45, 1.1, 2, 10),
        (32, 1.0, 0.0, 2.269, 2, 12),
        (12, 1.0, 0.1, 2.0, 2, 13),      # L not a multiple of 8
    ]
    shim = _Shim()
    saved = ref_mcmc.np
    ref_mcmc.np = types.SimpleNamespace(random=shim, exp=np.exp)
    try:
        for (L, J, h, T, n_sweeps, seed) in specs:
            model = IsingModel(L, J=J, h=h)
            sampler = ref_mcmc.MetropolisHastings(model, T)
            n_chains = 2
            words = o_ising.init_bits_words(seed, np.arange(n_chains), L)
            start = o_ising.unpack(words, L)
            r = np.stack([o_ising.sweep_draws(seed, np.arange(n_chains), s, L) for s in range(n_sweeps)])
            final = np.zeros_like(start)
            for c in range(n_chains):
                x = torch.tensor(start[c].astype(np.float32))
                for s in range(n_sweeps):
                    for colour in (0, 1):
                        for i in range(L):
                            for j in range(L):
                                if (i + j) % 2 != colour:
                                    continue
                                shim.q = [i, j]
                                shim.u = float(r[s, c, i, j]) * 2.0 ** -32
                                x = sampler.step(x)
                final[c] = x.numpy().astype(np.int8)
            cases.append(dict(L=L, J=J, h=h, T=T, n_sweeps=n_sweeps, seed=seed,
                              start=start, r=r, final=final))
    finally:
        ref_mcmc.np = saved
    return cases


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    g = {}

    # ---- Ising: dE classes, energy, magnetisation -------------------------------------------
    for tag, (J, h) in {"a": (1.0, 0.0), "b": (0.7, -0.45), "c": (-1.0, 0.3)}.items():
        model = IsingModel(8, J=J, h=h)
        dE = np.zeros(16, dtype=np.float32)
        for sb in (0, 1):
            for u in range(5):
                lat = -torch.ones(8, 8)
                lat[3, 3] = 2 * sb - 1
                for n_, (a, b) in enumerate([(2, 3), (4, 3), (3, 2), (3, 4)]):
                    if n_ < u:
                        lat[a, b] = 1.0
                dE[8 * sb + u] = model.local_energy_diff(lat, 3, 3).item()
        g[f"dE_{tag}"] = dE
        g[f"dE_{tag}_Jh"] = np.array([J, h])
    rng = np.random.RandomState(3)
    for L in (8, 16, 64):
        lat = torch.tensor(rng.choice([-1.0, 1.0], size=(5, L, L)).astype(np.float32))
        for tag, (J, h) in {"a": (1.0, 0.0), "b": (0.7, -0.45)}.items():
            m = IsingModel(L, J=J, h=h)
            g[f"obs_L{L}_{tag}_spins"] = lat.numpy()
            g[f"obs_L{L}_{tag}_E"] = m.energy(lat).numpy()
            g[f"obs_L{L}_{tag}_M"] = m.magnetization(lat).numpy()
            g[f"obs_L{L}_{tag}_Eps"] = m.energy_per_spin(lat).numpy()
    # continuous lattice (the API accepts real-valued samples)
    latc = torch.randn(3, 16, 16)
    m = IsingModel(16, J=1.0, h=0.2)
    g["obs_cont_spins"] = latc.numpy()
    g["obs_cont_E"] = m.energy(latc).numpy()
    g["obs_cont_M"] = m.magnetization(latc).numpy()
    # known answers of backend/tests/test_ising.py:11-30,55-79
    m8 = IsingModel(8)
    up = torch.ones(8, 8)
    cb = torch.tensor([[1.0 if (i + j) % 2 == 0 else -1.0 for j in range(8)] for i in range(8)])
    g["ka_E_up"] = m8.energy(up).numpy()
    g["ka_E_cb"] = m8.energy(cb).numpy()
    g["ka_dE_up"] = m8.local_energy_diff(up, 0, 0).numpy()
    g["ka_M_cb"] = m8.magnetization(cb).numpy()
    np.savez_compressed(os.path.join(OUT, "ising_static.npz"), **g)

    # ---- Ising: checkerboard sweeps through the reference's step() ---------------------------
    cases = ising_checkerboard_cases()
    flat = {"n_cases": np.array(len(cases))}
    for idx, c in enumerate(cases):
        for k, v in c.items():
            flat[f"c{idx}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "ising_sweeps.npz"), **flat)

    # ---- pair_correlation with the reference's own random reference sites --------------------
    g = {}
    for L, n in ((8, 3), (16, 4)):
        samples = torch.tensor(rng.choice([-1.0, 1.0], size=(n, L, L)).astype(np.float32))
        np.random.seed(100 + L)
        res = pair_correlation(samples)
        np.random.seed(100 + L)
        ref_idx = np.random.choice(L * L, min(100, L * L), replace=False)
        g[f"corr_L{L}_samples"] = samples.numpy()
        g[f"corr_L{L}_ref_indices"] = ref_idx
        g[f"corr_L{L}_C"] = np.array(res["C_r"])
        g[f"corr_L{L}_r"] = np.array(res["r"])
    np.savez_compressed(os.path.join(OUT, "correlation.npz"), **g)

    # ---- VP-SDE noise level + Euler–Maruyama expression --------------------------------------
    g = {}
    dp = DiffusionProcess()
    for N in (10, 100, 1000):
        sig = []
        for i in range(N, 0, -1):
            t = torch.full((2,), i / N)
            _, s_t = dp.noise_level(t)
            sig.append(s_t[0].item())
        g[f"sigma_N{N}"] = np.array(sig, dtype=np.float64)
    dp2 = DiffusionProcess(beta_min=0.05, beta_max=12.0)
    g["sigma_alt_N50"] = np.array([dp2.noise_level(torch.full((1,), i / 50))[1][0].item() for i in range(50, 0, -1)])
    x = torch.randn(2, 1, 16, 16)
    sc = torch.randn(2, 1, 16, 16)
    z = torch.randn(2, 1, 16, 16)
    g["em_x"], g["em_score"], g["em_z"] = x.numpy(), sc.numpy(), z.numpy()
    outs, meta = [], []
    for (i, N, temp) in [(100, 100, 1.0), (57, 100, 1.3), (3, 100, 0.7), (2, 100, 1.0), (1, 100, 1.0), (640, 1000, 1.0)]:
        sigma = dp.noise_level(torch.full((2,), i / N))[1][0].item()
        dt = 1.0 / N
        drift = (sigma ** 2) * sc * dt
        if i > 1:
            diffusion_term = sigma * (dt ** 0.5) * z * temp
        else:
            diffusion_term = 0
        outs.append((x + drift + diffusion_term).numpy())
        meta.append([i, N, temp, sigma])
    g["em_out"] = np.stack(outs)
    g["em_meta"] = np.array(meta, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "diffusion_scalar.npz"), **g)

    # ---- Score U-Net forward + full sampler on a tiny net ------------------------------------
    g = {}
    torch.manual_seed(42)
    net = ScoreNetwork(in_channels=1, base_channels=8, time_embed_dim=16, num_blocks=2).eval()
    sd = net.state_dict()
    for k, v in sd.items():
        g["sd::" + k] = v.numpy()
    emb = SinusoidalTimeEmbedding(16)
    tt = torch.tensor([0.0, 0.013, 0.5, 1.0])
    g["temb_t"], g["temb_out"] = tt.numpy(), emb(tt).numpy()
    x = torch.randn(3, 1, 16, 16)
    t = torch.tensor([1.0, 0.37, 0.02])
    with torch.no_grad():
        g["fwd_x"], g["fwd_t"], g["fwd_out"] = x.numpy(), t.numpy(), net(x, t).numpy()
        xr = torch.randn(2, 1, 8, 16)  # non-square (backend/tests/test_models.py:208-214)
        tr = torch.tensor([0.9, 0.9])
        g["fwd_ns_x"], g["fwd_ns_t"], g["fwd_ns_out"] = xr.numpy(), tr.numpy(), net(xr, tr).numpy()
    # full sampler with recorded noise: patch torch.randn / randn_like used by the reference
    N, shape = 12, (2, 1, 8, 8)
    noises = []
    real_randn, real_randn_like = torch.randn, torch.randn_like

    def rec_randn(*a, **k):
        v = real_randn(*a, **k)
        noises.append(v.clone())
        return v

    def rec_randn_like(*a, **k):
        v = real_randn_like(*a, **k)
        noises.append(v.clone())
        return v

    sampler = DiffusionSampler(net, num_steps=N)
    torch.randn, torch.randn_like = rec_randn, rec_randn_like
    try:
        torch.manual_seed(7)
        out = sampler.sample(shape, temperature=1.1)
    finally:
        torch.randn, torch.randn_like = real_randn, real_randn_like
    g["smp_x0"] = noises[0].numpy()
    g["smp_noises"] = torch.stack(noises[1:]).numpy()
    g["smp_out"] = out.numpy()
    g["smp_meta"] = np.array([N, 1.1])
    np.savez_compressed(os.path.join(OUT, "unet_tiny.npz"), **g)

    # ---- Lennard-Jones -----------------------------------------------------------------------
    g = {}
    lj = LennardJonesSystem(16, 5.0)
    pos = lj.lattice_configuration(3) + 0.05 * torch.randn(3, 16, 2)
    g["lj16_pos"], g["lj16_E"], g["lj16_F"] = pos.numpy(), lj.energy(pos).numpy(), lj.forces(pos).numpy()
    lj2 = LennardJonesSystem(64, float((64 / 0.8) ** 0.5), epsilon=1.3, sigma=0.9, cutoff=2.2)
    pos2 = lj2.lattice_configuration(2) + 0.05 * torch.randn(2, 64, 2)
    pos2 = pos2 % lj2.box_size
    g["lj64_pos"], g["lj64_E"], g["lj64_F"] = pos2.numpy(), lj2.energy(pos2).numpy(), lj2.forces(pos2).numpy()
    g["lj64_params"] = np.array([64, lj2.box_size, 1.3, 0.9, 2.2])
    pair = torch.tensor([[[0.5, 0.5], [4.5, 0.5]]])  # periodic image: distance 1.0 at L = 5
    ljp = LennardJonesSystem(2, 5.0)
    g["ljpair_pos"], g["ljpair_E"], g["ljpair_F"] = pair.numpy(), ljp.energy(pair).numpy(), ljp.forces(pair).numpy()
    far = torch.tensor([[[0.0, 0.0], [3.0, 0.0]]])
    ljf = LennardJonesSystem(2, 10.0)
    g["ljfar_pos"], g["ljfar_E"] = far.numpy(), ljf.energy(far).numpy()
    g["lj_minimg_in"] = np.array([[4.0, 0.0], [-4.0, 0.0], [2.5, -2.5], [7.4, 0.1]], dtype=np.float32)
    g["lj_minimg_out"] = ljp.minimum_image(torch.tensor(g["lj_minimg_in"])).numpy()
    np.savez_compressed(os.path.join(OUT, "lennard_jones.npz"), **g)

    # ---- Langevin samplers (reference step() with recorded draws) and RDF -----------------------
    from backend.ml.samplers.langevin import LangevinDynamics, UnderdampedLangevin
    from backend.ml.analysis.rdf import compute_rdf
    g = {}
    lj = LennardJonesSystem(16, 5.0)
    real_randn_like = torch.randn_like
    rec = []

    def rec_like(*a, **k):
        v = real_randn_like(*a, **k)
        rec.append(v.clone())
        return v

    torch.manual_seed(21)
    np.random.seed(21)
    pos = lj.lattice_configuration(1).squeeze(0) + 0.05 * torch.randn(16, 2)
    smp = LangevinDynamics(lj, temperature=1.0, dt=0.002)
    traj, us = [pos.numpy().copy()], []
    torch.randn_like = rec_like
    try:
        st = np.random.get_state()
        for _ in range(6):
            # the reference draws its uniform only inside step(); replay the same stream to record it
            pos = smp.step(pos)
            traj.append(pos.numpy().copy())
        np.random.set_state(st)
        us = [np.random.random() for _ in range(6)]
    finally:
        torch.randn_like = real_randn_like
    g["mala_traj"], g["mala_noise"], g["mala_u"] = np.stack(traj), torch.stack(rec).numpy(), np.array(us)
    g["mala_params"] = np.array([16, 5.0, 1.0, 0.002])
    rec.clear()
    ud = UnderdampedLangevin(lj, temperature=1.2, dt=0.003, mass=1.5, friction=0.8)
    p0 = lj.lattice_configuration(1).squeeze(0) + 0.05 * torch.randn(16, 2)
    v0 = 0.3 * torch.randn(16, 2)
    ptraj, vtraj = [p0.numpy().copy()], [v0.numpy().copy()]
    torch.randn_like = rec_like
    try:
        pp, vv = p0, v0
        for _ in range(5):
            pp, vv = ud.step(pp, vv)
            ptraj.append(pp.numpy().copy()); vtraj.append(vv.numpy().copy())
    finally:
        torch.randn_like = real_randn_like
    g["baoab_pos"], g["baoab_vel"], g["baoab_noise"] = np.stack(ptraj), np.stack(vtraj), torch.stack(rec).numpy()
    g["baoab_params"] = np.array([16, 5.0, 1.2, 0.003, 1.5, 0.8])
    cfgs = (lj.lattice_configuration(4) + 0.2 * torch.randn(4, 16, 2)) % 5.0
    r, gr = compute_rdf(cfgs, 5.0, n_bins=40)
    g["rdf_pos"], g["rdf_r"], g["rdf_g"] = cfgs.numpy(), r, gr
    r2, gr2 = compute_rdf(cfgs[0], 5.0, n_bins=25, r_max=2.0)
    g["rdf2_r"], g["rdf2_g"] = r2, gr2
    np.savez_compressed(os.path.join(OUT, "langevin_rdf.npz"), **g)

    # ---- particle score network (ContinuousScoreNetwork) + sampler on particle configurations -------
    g = {}
    torch.manual_seed(43)
    cnet = ContinuousScoreNetwork(n_particles=12, hidden_dim=32, time_embed_dim=16, pos_embed_dim=16, num_blocks=2,
                                  num_heads=4).eval()
    for k, v in cnet.state_dict().items():
        g["sd::" + k] = v.numpy()
    x = torch.rand(3, 12, 2) * 2 - 1
    t = torch.tensor([1.0, 0.41, 0.03])
    with torch.no_grad():
        g["fwd_x"], g["fwd_t"], g["fwd_out"] = x.numpy(), t.numpy(), cnet(x, t).numpy()
        xw = torch.randn(2, 40, 2) * 1.5     # outside [-1, 1] and N != n_particles: forward() accepts both
        tw = torch.tensor([0.77, 0.77])
        g["fwd_w_x"], g["fwd_w_t"], g["fwd_w_out"] = xw.numpy(), tw.numpy(), cnet(xw, tw).numpy()
    torch.manual_seed(44)
    cnet_na = ContinuousScoreNetwork(n_particles=8, hidden_dim=24, time_embed_dim=8, pos_embed_dim=8, num_blocks=1,
                                     use_attention=False).eval()
    for k, v in cnet_na.state_dict().items():
        g["sdna::" + k] = v.numpy()
    with torch.no_grad():
        xa = torch.rand(2, 8, 2) * 2 - 1
        ta = torch.tensor([0.6, 0.2])
        g["na_x"], g["na_t"], g["na_out"] = xa.numpy(), ta.numpy(), cnet_na(xa, ta).numpy()
    N, shape = 10, (2, 12, 2)
    noises = []
    real_randn, real_randn_like = torch.randn, torch.randn_like

    def rec2_randn(*a, **k):
        v = real_randn(*a, **k)
        noises.append(v.clone())
        return v

    def rec2_randn_like(*a, **k):
        v = real_randn_like(*a, **k)
        noises.append(v.clone())
        return v

    sampler = DiffusionSampler(cnet, num_steps=N)
    torch.randn, torch.randn_like = rec2_randn, rec2_randn_like
    try:
        torch.manual_seed(8)
        out = sampler.sample(shape, temperature=0.9)
    finally:
        torch.randn, torch.randn_like = real_randn, real_randn_like
    g["smp_x0"] = noises[0].numpy()
    g["smp_noises"] = torch.stack(noises[1:]).numpy()
    g["smp_out"] = out.numpy()
    g["smp_meta"] = np.array([N, 0.9])
    np.savez_compressed(os.path.join(OUT, "csn_tiny.npz"), **g)

    # ---- noise schedules (host-side): sigma tables through DiffusionProcess(schedule=...) ----------
    g = {}
    tt = torch.tensor([0.0, 0.01, 0.1, 0.37, 0.5, 0.9, 1.0])
    g["t"] = tt.numpy()
    for kind, kw in (("linear", {}), ("cosine", {}), ("cosine_s", {"s": 0.02}), ("sigmoid", {}),
                     ("sigmoid_k", {"start": -3.0, "end": 4.0, "beta_min": 0.2, "beta_max": 12.0})):
        sch = ref_get_schedule(kind.split("_")[0], **kw)
        a, sg = sch.noise_level(tt)
        g[kind + "_alpha"], g[kind + "_sigma"], g[kind + "_beta"] = a.numpy(), sg.numpy(), sch.beta(tt).numpy()
        dp = DiffusionProcess(schedule=sch)
        g[kind + "_table"] = np.array([dp.noise_level(torch.full((1,), i / 20))[1][0].item() for i in range(20, 0, -1)])
    np.savez_compressed(os.path.join(OUT, "noise_schedules.npz"), **g)


    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
