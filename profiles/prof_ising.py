"""Driver for ncu: C3 per-GPU shape of the Metropolis sweep (64x64, 8192 chains, 64 temperatures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import numpy as np, torch
from backend.ml.systems.ising import IsingModel
from backend.ml.samplers.mcmc import MetropolisHastings
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
h = float(os.environ.get('ISING_H', '0'))
m = IsingModel(L, h=h)
temps = np.repeat(np.linspace(1.5, 3.5, 64), n // 64)
mh = MetropolisHastings(m, temps, seed=1234)
bits = m.random_packed(n, seed=1234)
for _ in range(3):
    mh.sweeps_packed(bits, sweeps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); mh.sweeps_packed(bits, sweeps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"h={h} L={L} chains={n} sweeps={sweeps}: {ms:.3f} ms  {n*L*L*sweeps/ms/1e6:.1f} Gflips/s")
