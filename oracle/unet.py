"""Score U-Net forward and the reverse-SDE sampler restated with plain torch fp32 ops on the CPU.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the "torch fp32 reference" for the
floating-point kernels: a functional forward driven by a ScoreNetwork state_dict (no nn.Module
from the product is involved), plus the sampling loop with injectable noise.

Restates, citing /root/reference/backend/ml:
  SinusoidalTimeEmbedding.forward   models/score_network.py:20-41
  ConvBlock.forward                 models/score_network.py:133-151
  ScoreNetwork.forward              models/score_network.py:224-268
  DiffusionSampler.sample           samplers/diffusion.py:197-245
"""
import math

import torch
import torch.nn.functional as Fn

from . import em as _em


def time_embedding(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half) / (half - 1))
    args = t[:, None] * freqs[None, :] * 1000
    return torch.cat([torch.sin(args), torch.cos(args)], dim=-1)


def _groups(ch):
    return min(8, ch)


def conv_block(sd, prefix, x, t_emb, residual_scale=1.0, taps=None):
    h = Fn.conv2d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=1)
    if taps is not None:
        taps[prefix + "conv1"] = h
    c = h.shape[1]
    h = Fn.silu(Fn.group_norm(h, _groups(c), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], eps=1e-5))
    h = h + Fn.linear(t_emb, sd[prefix + "time_mlp.weight"], sd[prefix + "time_mlp.bias"])[:, :, None, None]
    h = Fn.conv2d(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], padding=1)
    h = Fn.silu(Fn.group_norm(h, _groups(c), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], eps=1e-5))
    if prefix + "residual.weight" in sd:
        r = Fn.conv2d(x, sd[prefix + "residual.weight"], sd[prefix + "residual.bias"])
    else:
        r = x
    return h + residual_scale * r


def unet_config_from_state_dict(sd):
    base = sd["init_conv.weight"].shape[0]
    in_ch = sd["init_conv.weight"].shape[1]
    tdim = sd["time_mlp.0.weight"].shape[1]
    nb = len({k.split(".")[1] for k in sd if k.startswith("encoder.")})
    return dict(in_channels=in_ch, base_channels=base, time_embed_dim=tdim, num_blocks=nb)


@torch.no_grad()
def unet_forward(sd, x, t, output_scale=1.0, normalize_output=False, taps=None):
    """score = ScoreNetwork.forward(x, t) from a state_dict of fp32 CPU tensors."""
    cfg = unet_config_from_state_dict(sd)
    nb = cfg["num_blocks"]
    emb = time_embedding(t, cfg["time_embed_dim"])
    t_emb = Fn.linear(Fn.silu(Fn.linear(emb, sd["time_mlp.0.weight"], sd["time_mlp.0.bias"])),
                      sd["time_mlp.2.weight"], sd["time_mlp.2.bias"])
    h = Fn.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=1)
    if taps is not None:
        taps["init"] = h
    skips = [h]
    for i in range(nb):
        h = conv_block(sd, f"encoder.{i}.", h, t_emb, taps=taps)
        if taps is not None:
            taps[f"enc{i}"] = h
        skips.append(h)
        h = Fn.conv2d(h, sd[f"downsample.{i}.weight"], sd[f"downsample.{i}.bias"], stride=2, padding=1)
        if taps is not None:
            taps[f"down{i}"] = h
    h = conv_block(sd, "middle.", h, t_emb, taps=taps)
    if taps is not None:
        taps["mid"] = h
    for i, skip in enumerate(reversed(skips[1:])):
        h = Fn.conv_transpose2d(h, sd[f"upsample.{i}.weight"], sd[f"upsample.{i}.bias"], stride=2, padding=1)
        if taps is not None:
            taps[f"up{i}"] = h
        if h.shape[-2:] != skip.shape[-2:]:
            h = Fn.interpolate(h, size=skip.shape[-2:], mode="bilinear", align_corners=False)
        h = torch.cat([h, skip], dim=1)
        h = conv_block(sd, f"decoder.{i}.", h, t_emb, taps=taps)
        if taps is not None:
            taps[f"dec{i}"] = h
    score = Fn.conv2d(h, sd["final.weight"], sd["final.bias"], padding=1)
    if normalize_output:
        score = score / (score.std(dim=(-1, -2), keepdim=True) + 1e-8)
    return output_scale * score


@torch.no_grad()
def sample(sd, shape, num_steps, x0, noises, temperature=1.0, beta_min=0.1, beta_max=20.0,
           output_scale=1.0, normalize_output=False):
    """DiffusionSampler.sample (samplers/diffusion.py:197-245) with injected x0 and per-step noise
    (noises[k] is used at step i = num_steps - k; the last step adds none)."""
    x = x0.clone()
    dt = 1.0 / num_steps
    for k, i in enumerate(range(num_steps, 0, -1)):
        t = torch.full((shape[0],), i / num_steps)
        _, sig = _em.noise_level(t.numpy(), beta_min, beta_max)
        sigma = float(sig[0])
        score = unet_forward(sd, x, t, output_scale, normalize_output)
        drift = (sigma ** 2) * score * dt
        if i > 1:
            x = x + drift + sigma * (dt ** 0.5) * noises[k] * temperature
        else:
            x = x + drift + 0
    return x
