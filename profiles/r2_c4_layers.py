"""Round-2 baseline at the config-4 shape: per-layer device times (in-library CUDA events) of one U-Net
forward at 64x64 for several batch sizes, and the graph-replayed step time (forward + EM) per batch size.
usage: python profiles/r2_c4_layers.py [L] [B ...]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from backend.ml.models.score_network import ScoreNetwork
from backend.ml.samplers.diffusion import DiffusionSampler
from dbs_b200 import abi
import bench

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Bs = [int(a) for a in sys.argv[2:] if int(a) > 0]
torch.manual_seed(0)
net = ScoreNetwork(1, 32, 64, 3).eval()
t = torch.full((1,), 0.5, device="cuda")
out = os.path.join(ROOT, "gpurun_out", "layers.tsv")
os.makedirs(os.path.dirname(out), exist_ok=True)
for B in Bs:
    x = torch.randn(B, 1, L, L, device="cuda")
    for _ in range(3):
        net.engine_forward(x, t, precision="bf16")
    torch.cuda.synchronize()
    abi.profile_read(True)
    abi.profile(True)
    reps = 3
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
    print(f"=== B={B} L={L}: per forward (mean of {reps})")
    for k, (n, ms) in agg.items():
        print(f"{ms / reps * 1e3:9.1f} us  x{n // reps:2d}  {k}")
    print({k: round(v[0] / reps, 3) for k, v in tot.items() if v[1]})
    flops = bench.unet_flops_per_sample(L) * B
    print("conv TFLOP/s:", flops / (tot["conv_tc"][0] / reps * 1e-3) / 1e12, flush=True)

print("=== graph-replayed sampler step (forward + EM), ms per EM step and whole-step TFLOP/s")
for B in [64, 128, 256, 512, 1024, 2048]:
    NS = 12
    smp = DiffusionSampler(net, num_steps=NS, device="cuda", precision="bf16")
    if os.environ.get("SPLIT", "1") == "0":
        smp.split_streams = False
    smp.sample((B, 1, L, L), seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(3):
        smp.sample((B, 1, L, L), seed=2 + r)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (3 * NS)
    tf = bench.unet_flops_per_sample(L) * B / (ms * 1e-3) / 1e12
    print(f"B={B:5d}  {ms:8.3f} ms/step  {ms * 1e3 / B:7.3f} us/sample-step  {tf:7.1f} TFLOP/s  frac {tf / 1386.3:.3f}", flush=True)
    del smp
    torch.cuda.empty_cache()
