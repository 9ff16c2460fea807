"""Second API probe: odd / non-divisible lattices through the higher-level sampler entry points."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from backend.ml.systems.ising import IsingModel
from backend.ml.systems.lennard_jones import LennardJonesSystem
from backend.ml.samplers.mcmc import MetropolisHastings
from backend.ml.samplers.diffusion import DiffusionSampler, PretrainedDiffusionSampler
from backend.ml.samplers.langevin import LangevinDynamics, UnderdampedLangevin
from backend.ml.models.score_network import ScoreNetwork
from backend.ml.analysis import comprehensive_comparison, compute_rdf_from_system, rdf_statistics, autocorrelation_time

def case(name, fn):
    try:
        print("OK  ", name, "->", fn())
    except Exception as e:
        print("FAIL", name, "->", type(e).__name__, str(e)[:300])

for L in (9, 10, 33):
    p = PretrainedDiffusionSampler(lattice_size=L, num_steps=6)
    case(f"pretrained L={L} sample", lambda: tuple(p.sample((2, 1, L, L)).shape))
    case(f"pretrained L={L} heuristic", lambda: tuple(p.sample_heuristic(2, temperature=2.0).shape))
    case(f"pretrained L={L} sample_ising", lambda: tuple(p.sample_ising(2, lattice_size=L).shape))
    case(f"pretrained L={L} trajectory+stats", lambda: len(list(p.sample_trajectory_with_metadata((1, 1, L, L), n_frames=3, include_statistics=True, ising_model=IsingModel(L)))))
    case(f"pretrained L={L} ode rk4", lambda: tuple(p.sample_ode((2, 1, L, L), method="rk4").shape))
    case(f"pretrained L={L} pc", lambda: tuple(p.sample_predictor_corrector((2, 1, L, L)).shape))
    case(f"pretrained L={L} at_temperature", lambda: tuple(p.sample_at_temperature((2, 1, L, L), 3.0).shape))
    mh = MetropolisHastings(IsingModel(L), 2.27)
    case(f"mcmc L={L} sample", lambda: tuple(mh.sample(4, n_sweeps=2, burn_in=5).shape))
    case(f"mcmc L={L} trajectory", lambda: len(list(mh.sample_with_trajectory(3))))
    case(f"mcmc L={L} step", lambda: tuple(mh.step(torch.ones(L, L)).shape))
    a = MetropolisHastings(IsingModel(L), 2.0, seed=1).sample(12, n_sweeps=1, burn_in=5)
    b = MetropolisHastings(IsingModel(L), 3.0, seed=2).sample(12, n_sweeps=1, burn_in=5)
    case(f"comprehensive_comparison L={L}", lambda: sorted(comprehensive_comparison(a, b, IsingModel(L), 10)["summary"]))
    case(f"autocorrelation L={L}", lambda: autocorrelation_time(a))
mh = MetropolisHastings(IsingModel(16), 2.27)
case("mcmc sample from batched initial (3 chains)", lambda: tuple(mh.sample(2, n_sweeps=1, burn_in=1, initial=torch.ones(3, 16, 16)).shape))
case("mcmc L=130 rejected", lambda: MetropolisHastings(IsingModel(130), 2.0).sample(1, burn_in=1).shape)
case("mcmc L=3 rejected", lambda: MetropolisHastings(IsingModel(3), 2.0).sample(1, burn_in=1).shape)
lj = LennardJonesSystem(25, 6.0)
s = LangevinDynamics(lj, 1.0, dt=1e-3).sample(4, n_steps=5, burn_in=20)
case("rdf + statistics", lambda: sorted(rdf_statistics(*compute_rdf_from_system(s, lj, n_bins=30), sigma=lj.sigma)))
case("lj score", lambda: tuple(lj.score(s, 1.0).shape))
case("lj to/from continuous", lambda: float((lj.from_continuous(lj.to_continuous(s)) - s).abs().max()))
case("lj pairwise_distances", lambda: tuple(t.shape for t in lj.pairwise_distances(s)))
case("lj random_configuration", lambda: tuple(lj.random_configuration(3).shape))
case("baoab step", lambda: tuple(t.shape for t in UnderdampedLangevin(lj, 1.0, dt=1e-3).step(s[0], torch.zeros_like(s[0]))))
