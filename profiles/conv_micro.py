"""Micro-benchmark of dbs_conv_tc on one layer shape: python profiles/conv_micro.py kind k B H W C0 C1 Cout [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from dbs_b200 import abi, runtime
kind, k, B, H, W, C0, C1, Cout = [int(v) for v in sys.argv[1:9]]
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 20
stats_on = int(os.environ.get("STATS", "1"))
torch.manual_seed(0)
x0 = torch.randn(B, H, W, C0, device="cuda").to(torch.bfloat16)
x1 = torch.randn(B, H, W, C1, device="cuda").to(torch.bfloat16) if C1 else None
taps = 16 if kind == 2 else k * k
w = (torch.randn(taps, Cout, C0 + C1, device="cuda") / (taps * (C0 + C1)) ** 0.5).to(torch.bfloat16)
b = torch.randn(Cout, device="cuda")
Ho, Wo = (2 * H, 2 * W) if kind == 2 else ((H // 2, W // 2) if kind == 1 else (H, W))
out = torch.zeros(B, Ho, Wo, Cout, dtype=torch.bfloat16, device="cuda")
st = torch.zeros(B, 8, 2, device="cuda") if stats_on else None
def run():
    abi.check(abi.lib.dbs_conv_tc(x0.data_ptr(), C0, runtime.ptr(x1), C1, w.data_ptr(), b.data_ptr(), out.data_ptr(),
                                  runtime.ptr(st), B, H, W, Cout, kind, k, runtime.current_stream()))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
fl = 2.0 * B * H * W * Cout * (C0 + C1) * (4 if kind == 2 else k * k) * (4 if kind == 2 else 1) / (4 if kind == 1 else 1)
print(f"dbg={os.environ.get('DBS_STRIP_DBG','0')} stats={stats_on} {sys.argv[1:9]}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
