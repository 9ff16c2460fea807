"""Large batches in ONE forward (DiffusionSampler.sample does not chunk): 64x64, bf16, B = 4096 and 8192 — the activation
tensors pass 2^31 bytes / 2^31 elements there — against the same samples run 1024 at a time.  Per-sample relative rms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from backend.ml.models.score_network import ScoreNetwork, PRECISIONS
from dbs_b200 import abi
torch.manual_seed(0)
net = ScoreNetwork(1, 32, 64, 3).eval()
L = 64
for B in (4096, 8192):
    need = abi.lib.dbs_unet_workspace_bytes(net.engine_plan(), B, L, L, PRECISIONS["bf16"])
    print(f"B={B}: workspace {need / 2**30:.1f} GiB", flush=True)
    if need > 100 * 2**30:
        print("skipped"); continue
    g = torch.Generator(device="cuda").manual_seed(B)
    x = torch.randn(B, 1, L, L, device="cuda", generator=g)
    t = torch.rand(B, device="cuda", generator=g)
    full = net.engine_forward(x, t, precision="bf16").float()
    parts = torch.cat([net.engine_forward(x[i:i + 1024].contiguous(), t[i:i + 1024].contiguous(), precision="bf16").float() for i in range(0, B, 1024)])
    torch.cuda.synchronize()
    per = ((full - parts) ** 2).flatten(1).mean(1).sqrt() / (parts ** 2).flatten(1).mean(1).sqrt()
    print(f"B={B}: per-sample rel rms max {per.max().item():.3e} at {per.argmax().item()}, mean {per.mean().item():.3e}, finite {bool(torch.isfinite(full).all())}", flush=True)
    del full, parts, x
    net._workspace = None
    torch.cuda.empty_cache()
