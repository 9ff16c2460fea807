import sys, traceback
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from backend.ml.systems.ising import IsingModel
from backend.ml.systems.lennard_jones import LennardJonesSystem
from backend.ml.samplers.mcmc import MetropolisHastings
from backend.ml.samplers.diffusion import DiffusionSampler
from backend.ml.samplers.langevin import LangevinDynamics, UnderdampedLangevin
from backend.ml.models.score_network import ScoreNetwork
from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
from backend.ml.analysis import pair_correlation, compute_rdf, magnetization_distribution

def case(name, fn):
    try:
        r = fn()
        print("OK  ", name, "->", r)
    except Exception as e:
        print("FAIL", name, "->", type(e).__name__, str(e)[:200])

m = IsingModel(16)
case("energy of empty batch", lambda: tuple(m.energy(torch.empty(0, 16, 16)).shape))
case("energy single 2D lattice", lambda: m.energy(torch.ones(16, 16)).shape)
case("magnetization 4D (2,3,16,16)", lambda: tuple(m.magnetization(torch.ones(2, 3, 16, 16)).shape))
case("random_configuration(0)", lambda: tuple(m.random_configuration(0).shape))
mh = MetropolisHastings(m, 2.0, seed=1)
case("sample burn_in=0 n_sweeps=0", lambda: tuple(mh.sample(3, n_sweeps=0, burn_in=0).shape))
case("sample n_samples=1", lambda: tuple(mh.sample(1).shape))
case("sample n_chains=0", lambda: tuple(mh.sample(2, n_chains=0).shape))
case("sweep on (L,L) int64 tensor", lambda: mh.sweep(torch.ones(16, 16)).dtype)
case("sweep on float64 lattice", lambda: mh.sweep(torch.ones(16, 16, dtype=torch.float64)).dtype)
case("step on batch", lambda: tuple(mh.step(torch.ones(4, 16, 16)).shape))
case("trajectory n_steps=0", lambda: len(list(mh.sample_with_trajectory(0))))
case("T=0 sampler", lambda: MetropolisHastings(m, 0.0, seed=2).sample(2, n_sweeps=1, burn_in=5).mean().item())
case("pair_correlation 1 sample", lambda: len(pair_correlation(torch.ones(1, 16, 16))["C_r"]))
case("magdist 1 sample", lambda: sum(magnetization_distribution(torch.ones(1, 16, 16), 10)["P_M"]))
net = ScoreNetwork(1, 16, 32, 2).eval()
with torch.no_grad():
    case("unet B=1", lambda: tuple(net(torch.randn(1, 1, 16, 16), torch.rand(1)).shape))
    case("unet B=0", lambda: tuple(net(torch.randn(0, 1, 16, 16), torch.rand(0)).shape))
    case("unet 8x8 (1x1 bottleneck... 2 blocks -> 2x2)", lambda: tuple(net(torch.randn(2, 1, 8, 8), torch.rand(2)).shape))
    case("unet 4x4 -> 1x1 bottleneck", lambda: tuple(net(torch.randn(2, 1, 4, 4), torch.rand(2)).shape))
    case("unet float64 input", lambda: net(torch.randn(2, 1, 16, 16, dtype=torch.float64), torch.rand(2)).dtype)
    case("unet t scalar-shaped (1,)", lambda: tuple(net(torch.randn(3, 1, 16, 16), torch.rand(1)).shape))
smp = DiffusionSampler(net, num_steps=5)
case("sampler B=1", lambda: tuple(smp.sample((1, 1, 16, 16)).shape))
case("sampler B=0", lambda: tuple(smp.sample((0, 1, 16, 16)).shape))
case("sampler ode heun", lambda: tuple(smp.sample_ode((2, 1, 16, 16), method="heun").shape))
case("sampler pc", lambda: tuple(smp.sample_predictor_corrector((2, 1, 16, 16)).shape))
case("generate_batch 5 of batch 2", lambda: tuple(smp.generate_batch(5, lattice_size=16, batch_size=2).shape))
case("generate_batch lattice 12", lambda: tuple(smp.generate_batch(3, lattice_size=12, batch_size=2).shape))
case("sampler bf16 base16 B=3", lambda: tuple(DiffusionSampler(net, num_steps=3, precision="bf16").sample((3, 1, 16, 16)).shape))
lj = LennardJonesSystem(16, 5.0)
case("lj energy empty batch", lambda: tuple(lj.energy(torch.empty(0, 16, 2)).shape))
case("lj energy single (N,2)", lambda: tuple(lj.energy(lj.lattice_configuration(1)[0]).shape))
case("lj forces N=1", lambda: LennardJonesSystem(1, 5.0).forces(torch.zeros(1, 1, 2)).abs().max().item())
case("lj N=300 (>256)", lambda: tuple(LennardJonesSystem(300, 20.0).energy(LennardJonesSystem(300, 20.0).lattice_configuration(2)).shape))
case("rdf 2D input", lambda: compute_rdf(lj.lattice_configuration(1)[0], 5.0, n_bins=10)[1].shape)
case("langevin sample", lambda: tuple(LangevinDynamics(lj, 1.0, dt=1e-3).sample(2, n_steps=3, burn_in=5).shape))
case("baoab sample", lambda: tuple(UnderdampedLangevin(lj, 1.0, dt=1e-3).sample(2, n_steps=3, burn_in=5).shape))
csn = ContinuousScoreNetwork(16).eval()
with torch.no_grad():
    case("csn B=1", lambda: tuple(csn(torch.rand(1, 16, 2) * 2 - 1, torch.rand(1)).shape))
    case("csn B=0", lambda: tuple(csn(torch.rand(0, 16, 2), torch.rand(0)).shape))
    case("csn N=7 (!= n_particles)", lambda: tuple(csn(torch.rand(2, 7, 2) * 2 - 1, torch.rand(2)).shape))
