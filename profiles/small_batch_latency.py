"""Latency of DiffusionSampler.sample for API-sized requests (batch 1..16), eager vs CUDA-graph step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from backend.ml.models.score_network import ScoreNetwork
from backend.ml.samplers.diffusion import DiffusionSampler
torch.manual_seed(0)
net = ScoreNetwork(1, 32, 64, 3).eval()
for graph in (False, True):
    for prec in ("fp32", "bf16"):
        smp = DiffusionSampler(net, num_steps=100, device="cpu", precision=prec)
        if hasattr(smp, "use_graph"):
            smp.use_graph = graph
        elif graph:
            continue
        for B in (1, 8):
            shape = (B, 1, 32, 32)
            smp.sample(shape, seed=1)
            torch.cuda.synchronize()
            t0 = time.time()
            for r in range(3):
                smp.sample(shape, seed=2 + r)
            torch.cuda.synchronize()
            print(f"graph={graph} {prec} B={B}: {(time.time() - t0) / 3 * 1e3:.1f} ms per 100-step sample() call")
