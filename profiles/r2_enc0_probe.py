"""Composed first layer (enc0_composed_kernel) against the two-launch path and the fp32 oracle, per shape.
usage: python profiles/r2_enc0_probe.py   (spawns itself with DBS_ENC0_COMPOSED=1 / 0; DBS_B200_LIB selects the library)"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import numpy as np, torch
SHAPES = ((8, 16, 5), (16, 16, 4), (24, 48, 3), (32, 32, 4), (64, 64, 2))

def inputs(H, W, B):
    g = torch.Generator().manual_seed(H * 100 + W)
    return torch.randn(B, 1, H, W, generator=g), torch.full((B,), 0.61)

if len(sys.argv) > 1:
    from backend.ml.models.score_network import ScoreNetwork
    torch.manual_seed(0)
    net = ScoreNetwork(1, 32, 64, 3).eval()
    outs = {}
    for H, W, B in SHAPES:
        x, t = inputs(H, W, B)
        outs[f"{H}x{W}"] = net.engine_forward(x, t, precision="bf16").cpu().numpy()
        outs[f"{H}x{W}_again"] = net.engine_forward(x, t, precision="bf16").cpu().numpy()
    np.savez(sys.argv[1], **outs)
    sys.exit(0)

from backend.ml.models.score_network import ScoreNetwork
from oracle import unet as o_unet
torch.manual_seed(0)
net = ScoreNetwork(1, 32, 64, 3).eval()
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
res = []
with tempfile.TemporaryDirectory() as tmp:
    for flag in ("1", "0"):
        path = os.path.join(tmp, f"c{flag}.npz")
        r = subprocess.run([sys.executable, __file__, path], env=dict(os.environ, DBS_ENC0_COMPOSED=flag), capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        with np.load(path) as z:
            res.append({k: z[k].astype(np.float64) for k in z.files})
rms = lambda a, b: float(np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean()))
for H, W, B in SHAPES:
    k = f"{H}x{W}"
    x, t = inputs(H, W, B)
    want = o_unet.unet_forward(sd, x, t).numpy().astype(np.float64)
    print(f"{k:6s} composed vs two-launch {rms(res[0][k], res[1][k]):.2e}   composed vs oracle {rms(res[0][k], want):.2e}   "
          f"two-launch vs oracle {rms(res[1][k], want):.2e}   run-to-run composed {rms(res[0][k], res[0][k + '_again']):.2e} "
          f"two-launch {rms(res[1][k], res[1][k + '_again']):.2e}", flush=True)
