"""Short driver for ncu: N forwards of the default score U-Net (bf16 tcgen05 path) + EM steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from backend.ml.models.score_network import ScoreNetwork
from backend.ml.samplers.diffusion import DiffusionSampler

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
net = ScoreNetwork(1, 32, 64, 3).eval()
smp = DiffusionSampler(net, num_steps=steps, device="cuda", precision="bf16")
if os.environ.get("DBS_NO_GRAPH"):
    smp.use_graph = False     # eager launches: a clean per-kernel list under ncu
x = smp.sample((B, 1, L, L), seed=1)
torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
