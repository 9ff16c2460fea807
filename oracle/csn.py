"""Particle score network (ContinuousScoreNetwork) restated with plain torch fp32 ops on the CPU.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a functional forward driven by a state_dict, plus
the reverse-SDE loop with injectable noise for (B, N, 2) particle configurations.

Restates, citing /root/reference/backend/ml/models/score_network_continuous.py:
  PeriodicPositionEmbedding.forward  :36-75
  SinusoidalTimeEmbedding.forward    :85-101
  MLPBlock.forward                   :123-136
  ParticleAttention.forward          :155-180
  ContinuousScoreNetwork.forward     :256-282
and DiffusionSampler.sample (samplers/diffusion.py:197-245), which is shape-agnostic.
The reference has no test and no caller for this network (SURVEY.md §8c iv); the oracle is pinned
by tests/golden/csn_tiny.npz, generated from the unmodified reference module.
"""
import math

import torch
import torch.nn.functional as Fn

from . import em as _em
from .unet import time_embedding


def config_from_state_dict(sd):
    hidden = sd["input_proj.weight"].shape[0]
    pos_dim = sd["input_proj.weight"].shape[1] - 2
    tdim = sd["time_mlp.0.weight"].shape[1]
    nb = len({k.split(".")[1] for k in sd if k.startswith("blocks.")})
    attn = any(k.startswith("attention_layers.") for k in sd)
    return dict(hidden_dim=hidden, pos_embed_dim=pos_dim, time_embed_dim=tdim, num_blocks=nb, use_attention=attn)


def position_embedding(x, freqs):
    ang = ((x + 1) * math.pi)[..., None] * freqs          # (..., 2, n_freq)
    ax, ay = ang[..., 0, :], ang[..., 1, :]
    return torch.cat([torch.sin(ax), torch.cos(ax), torch.sin(ay), torch.cos(ay)], dim=-1)


def _ln(x, sd, prefix):
    return Fn.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-5)


def _lin(x, sd, prefix):
    return Fn.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


@torch.no_grad()
def csn_forward(sd, x, t, num_heads=4):
    """score = ContinuousScoreNetwork.forward(x, t) from a state_dict of fp32 CPU tensors."""
    cfg = config_from_state_dict(sd)
    H = cfg["hidden_dim"]
    t_emb = _lin(Fn.silu(_lin(time_embedding(t, cfg["time_embed_dim"]), sd, "time_mlp.0")), sd, "time_mlp.2")
    h = _lin(torch.cat([position_embedding(x, sd["pos_embed.freqs"]), x], dim=-1), sd, "input_proj")
    B, N, _ = h.shape
    hd = H // num_heads
    for i in range(cfg["num_blocks"]):
        p = f"blocks.{i}."
        u = Fn.silu(_ln(_lin(h, sd, p + "fc1"), sd, p + "norm1")) + _lin(t_emb, sd, p + "time_mlp")[:, None, :]
        h = h + Fn.silu(_ln(_lin(u, sd, p + "fc2"), sd, p + "norm2"))
        if cfg["use_attention"]:
            a = f"attention_layers.{i}."
            qkv = _lin(_ln(h, sd, a + "norm"), sd, a + "qkv").reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
            w = torch.softmax(torch.matmul(qkv[0], qkv[1].transpose(-2, -1)) * hd ** -0.5, dim=-1)
            o = torch.matmul(w, qkv[2]).transpose(1, 2).reshape(B, N, H)
            h = _lin(o, sd, a + "proj") + h
    return _lin(Fn.silu(_lin(_ln(h, sd, "output_proj.0"), sd, "output_proj.1")), sd, "output_proj.3")


def sample(sd, shape, num_steps, x0, noises, temperature=1.0, beta_min=0.1, beta_max=20.0, num_heads=4):
    """DiffusionSampler.sample (samplers/diffusion.py:197-245) on particle configurations with injected
    x0 and per-step noise (noises[k] at step i = num_steps - k; the last step adds none)."""
    x = x0.clone()
    dt = 1.0 / num_steps
    for k, i in enumerate(range(num_steps, 0, -1)):
        t = torch.full((shape[0],), i / num_steps)
        _, sig = _em.noise_level(t.numpy(), beta_min, beta_max)
        sigma = float(sig[0])
        drift = (sigma ** 2) * csn_forward(sd, x, t, num_heads) * dt
        x = x + drift + (sigma * (dt ** 0.5) * noises[k] * temperature if i > 1 else 0)
    return x
