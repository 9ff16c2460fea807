"""Round 2: bf16 tcgen05 path against the fp32 path over a FULL config-4 trajectory (64x64, 1000 Euler-Maruyama steps, same
Philox noise in both): rms error relative to the rms of the fp32 result, sign agreement, and the error growth along the way.
usage: python profiles/r2_bf16_long.py [B] [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from backend.ml.models.score_network import ScoreNetwork
from backend.ml.samplers.diffusion import DiffusionSampler

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
torch.manual_seed(0)
net = ScoreNetwork(1, 32, 64, 3).eval()
shape = (B, 1, 64, 64)
a = DiffusionSampler(net, num_steps=N, device="cuda", precision="fp32")
b = DiffusionSampler(net, num_steps=N, device="cuda", precision="bf16")
ga, gb = a._reverse_sde(shape, 1.0, 123), b._reverse_sde(shape, 1.0, 123)
for (i, xa), (_, xb) in zip(ga, gb):
    if i in (N + 1, N, N - 9, N - 99, N // 2, N // 4, 10, 1) or i % 200 == 0:
        rms = float((xa - xb).pow(2).mean().sqrt() / xa.pow(2).mean().sqrt())
        sign = float((torch.sign(xa) == torch.sign(xb)).float().mean())
        print(f"after step i={i:5d}: rel rms err {rms:.3e}  sign agreement {sign:.5f}  |x| rms {float(xa.pow(2).mean().sqrt()):.3f}", flush=True)
