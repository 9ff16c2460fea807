import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
from backend.ml.models.score_network import ScoreNetwork
from backend.ml.samplers.diffusion import DiffusionSampler
torch.manual_seed(0)
net = ScoreNetwork(1, 32, 64, 3).eval()
outs = []
for prec, L, B, steps in (("bf16", 64, 5, 12), ("bf16", 32, 130, 8), ("fp32", 16, 3, 6), ("bf16", 64, 5, 12)):
    smp = DiffusionSampler(net, num_steps=steps, precision=prec)
    outs.append(smp.sample((B, 1, L, L), seed=7).cpu().numpy())
np.savez(sys.argv[1], *outs)
