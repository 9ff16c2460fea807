import os, sys, subprocess, numpy as np
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
H, B = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3:
    import torch
    from backend.ml.models.score_network import ScoreNetwork
    torch.manual_seed(0)
    net = ScoreNetwork(1, 32, 64, 3).eval()
    torch.manual_seed(1)
    x, t = torch.randn(B, 1, H, H), torch.full((B,), 0.5)
    out = net.engine_forward(x, t, precision="bf16").cpu().numpy()
    np.save(sys.argv[3], out)
    sys.exit(0)
res = {}
for e in ("0", "1"):
    env = dict(os.environ, DBS_GN_PROLOGUE=e)
    subprocess.check_call([sys.executable, __file__, str(H), str(B), f"/tmp/pre_{e}.npy"], env=env)
    res[e] = np.load(f"/tmp/pre_{e}.npy")
d = np.abs(res["1"] - res["0"])
print("H", H, "B", B, "max diff", d.max(), "rms ref", np.sqrt((res["0"] ** 2).mean()), "rms diff", np.sqrt((d ** 2).mean()))
print("per image max", d.reshape(B, -1).max(1)[:8])
print("per row max", d.max(axis=(0, 1, 3)).round(3))
print("per col max", d.max(axis=(0, 1, 2)).round(3))
