"""LJ pair kernel (config-5 shape) with / without the half-precision prefilter, and one MALA step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import numpy as np, torch
from backend.ml.systems.lennard_jones import LennardJonesSystem
N, B = 256, 4096
s = LennardJonesSystem(N, float(np.sqrt(N / 0.8)))
torch.manual_seed(0)
pos = (s.lattice_configuration(B) + 0.05 * torch.randn(B, N, 2)).cuda()
for env in ("0", "1"):
    os.environ["DBS_LJ_PREFILTER"] = env
    for _ in range(3):
        s.energy_and_forces(pos)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        s.energy_and_forces(pos)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("prefilter", env, f"{ms:.4f} ms per launch, {B * N * N / ms * 1e3:.3e} ordered pairs/s")
