"""Per-layer device times of one U-Net forward (in-library CUDA-event profiler, no ncu).
usage: python profiles/layer_times.py [B] [L] [out.tsv]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from backend.ml.models.score_network import ScoreNetwork
from dbs_b200 import abi
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 32
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "layers.tsv")
torch.manual_seed(0)
net = ScoreNetwork(1, 32, 64, 3).eval()
x = torch.randn(B, 1, L, L, device="cuda")
t = torch.full((1,), 0.5, device="cuda")
for _ in range(3):
    net.engine_forward(x, t, precision="bf16")
torch.cuda.synchronize()
abi.profile_read(True)
abi.profile(True)
reps = 5
for _ in range(reps):
    net.engine_forward(x, t, precision="bf16")
torch.cuda.synchronize()
tot = abi.profile_read(True)
abi.profile(False)
abi.check(abi.lib.dbs_profile_dump(out.encode()))
agg = collections.OrderedDict()
for line in open(out):
    k, ms = line.rstrip("\n").split("\t")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += float(ms)
print(f"B={B} L={L}: per forward (mean of {reps})")
for k, (n, ms) in agg.items():
    print(f"{ms / reps * 1e3:9.1f} us  x{n // reps:2d}  {k}")
print({k: round(v[0] / reps, 3) for k, v in tot.items() if v[1]})
flops = bench.unet_flops_per_sample(L) * B
print("conv TFLOP/s:", flops / (tot["conv_tc"][0] / reps * 1e-3) / 1e12)
