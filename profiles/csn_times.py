"""Device times of the particle score network at the configuration-5 shape (N = 256), per kernel
class, from the in-library CUDA-event profiler.  usage: python profiles/csn_times.py [B] [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from backend.ml.models.score_network_continuous import ContinuousScoreNetwork
from dbs_b200 import abi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
net = ContinuousScoreNetwork(N).eval()
x = (torch.rand(B, N, 2) * 2 - 1).cuda()
t = torch.full((1,), 0.5).cuda()
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
for _ in range(2):
    net.engine_forward(x, t, precision=prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    net.engine_forward(x, t, precision=prec)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
H, nb, hd = 128, 4, 32
flops = 2.0 * B * N * (66 * H + nb * (6 * H * H) + H * H // 2 + H) + nb * 4.0 * B * 4 * N * N * hd
print(f"{prec} forward B={B} N={N}: {ms:.3f} ms  ({flops / ms / 1e9:.1f} TFLOP/s, {B / ms * 1e3:.0f} configs/s)")
abi.profile_read(reset=True); abi.profile(True)
net.engine_forward(x, t, precision=prec); torch.cuda.synchronize()
print({k: (round(v[0], 3), v[1]) for k, v in abi.profile_read().items() if v[1]})
abi.profile(False)
abi.lib.dbs_profile_dump(b"gpurun_out/csn_layers.tsv")
