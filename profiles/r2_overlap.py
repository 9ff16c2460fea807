"""Round 2: can an HBM-bound kernel run UNDER a tcgen05 strip conv on the same SMs?  Times a conv loop alone, a copy / GroupNorm
-apply-like loop alone, and both concurrently on two streams.  usage: python profiles/r2_overlap.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from dbs_b200 import abi, runtime

B, H, W, C0, Cout = 512, 64, 64, 64, 64
torch.manual_seed(0)
x0 = torch.randn(B, H, W, C0, device="cuda").to(torch.bfloat16)
w = (torch.randn(9, Cout, C0, device="cuda") / (9 * C0) ** 0.5).to(torch.bfloat16)
b = torch.randn(Cout, device="cuda")
out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
st = torch.zeros(B, 8, 2, device="cuda")
src = torch.randn(B, H, W, Cout, device="cuda").to(torch.bfloat16)
dst = torch.empty_like(src)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def conv(n):
    for _ in range(n):
        abi.check(abi.lib.dbs_conv_tc(x0.data_ptr(), C0, 0, 0, w.data_ptr(), b.data_ptr(), out.data_ptr(), st.data_ptr(),
                                      B, H, W, Cout, 0, 3, runtime.current_stream()))
def mem(n, kind):
    for _ in range(n):
        if kind == "copy":
            dst.copy_(src)
        else:
            torch.nn.functional.silu(src, inplace=False)

def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

conv(3); mem(3, "copy"); mem(3, "silu")
for kind in ("copy", "silu"):
    n = 20
    tc = timed(lambda: conv(n))
    tm = timed(lambda: mem(n, kind))
    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1): conv(n)
        with torch.cuda.stream(s2): mem(n, kind)
        cur.wait_stream(s1); cur.wait_stream(s2)
    tb = timed(both)
    print(f"{kind}: conv alone {tc / n * 1e3:7.1f} us  mem alone {tm / n * 1e3:7.1f} us  concurrent {tb / n * 1e3:7.1f} us per pair "
          f"(sum {(tc + tm) / n * 1e3:7.1f}, max {max(tc, tm) / n * 1e3:7.1f})")
