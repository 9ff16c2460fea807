"""Score U-Net forward and the reverse-SDE sampler restated with plain torch fp32 ops on the CPU.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the "torch fp32 reference" for the
floating-point kernels: a functional forward driven by a ScoreNetwork state_dict (no nn.Module
from the product is involved), plus the sampling loop with injectable noise.

Restates, citing /root/reference/backend/ml:
  SinusoidalTimeEmbedding.forward   models/score_network.py:20-41
  ConvBlock.forward                 models/score_network.py:133-151
  ScoreNetwork.forward              models/score_network.py:224-268
  DiffusionSampler.sample           samplers/diffusion.py:197-245
  DiffusionSampler.sample_predictor_corrector / sample_ode   samplers/diffusion.py:399-451, 453-521
"""
import math

import numpy as np
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


@torch.no_grad()
def sample_pc(sd, shape, num_steps, x0, noises, corrector_noises, temperature=1.0, n_corrector_steps=1,
              corrector_snr=0.16, beta_min=0.1, beta_max=20.0):
    """DiffusionSampler.sample_predictor_corrector (samplers/diffusion.py:399-451) with injected draws:
    noises[k] is the predictor's noise at step i = num_steps - k (none at i = 1), corrector_noises[k][c] the
    c-th Langevin corrector's noise of that step (no corrector at i = 1)."""
    x = x0.clone()
    dt = 1.0 / num_steps
    for k, i in enumerate(range(num_steps, 0, -1)):
        t = torch.full((shape[0],), i / num_steps)
        _, sig = _em.noise_level(t.numpy(), beta_min, beta_max)
        sigma = float(sig[0])
        score = unet_forward(sd, x, t)
        drift = (sigma ** 2) * score * dt
        x = x + drift + (sigma * (dt ** 0.5) * noises[k] * temperature if i > 1 else 0)
        if i > 1:
            for c in range(n_corrector_steps):
                score = unet_forward(sd, x, t)
                grad_norm = score.flatten(1).norm(dim=1).mean()
                noise_norm = (x.numel() / x.shape[0]) ** 0.5
                step_size = (corrector_snr * noise_norm / grad_norm) ** 2 * 2
                x = x + step_size * score + (2 * step_size) ** 0.5 * corrector_noises[k][c] * temperature
    return x


@torch.no_grad()
def sample_ode(sd, shape, num_steps, x0, method="euler", beta_min=0.1, beta_max=20.0):
    """DiffusionSampler.sample_ode (samplers/diffusion.py:453-521): euler / heun / rk4 from an injected start."""
    def sig2(t_val):
        _, s = _em.noise_level(np.full((1,), t_val, dtype=np.float32), beta_min, beta_max)
        return float(s[0]) ** 2

    def tt(t_val):
        return torch.full((shape[0],), t_val)

    x = x0.clone()
    dt = 1.0 / num_steps
    for i in range(num_steps, 0, -1):
        t_val = i / num_steps
        if method == "euler":
            x = x + sig2(t_val) * unet_forward(sd, x, tt(t_val)) * dt
        elif method == "heun":
            s0 = unet_forward(sd, x, tt(t_val))
            x_pred = x + sig2(t_val) * s0 * dt
            s1 = unet_forward(sd, x_pred, tt((i - 1) / num_steps))
            x = x + 0.5 * (sig2(t_val) * s0 + sig2((i - 1) / num_steps) * s1) * dt
        elif method == "rk4":
            k1 = sig2(t_val) * unet_forward(sd, x, tt(t_val))
            k2 = sig2(t_val - 0.5 * dt) * unet_forward(sd, x + 0.5 * dt * k1, tt(t_val - 0.5 * dt))
            k3 = sig2(t_val - 0.5 * dt) * unet_forward(sd, x + 0.5 * dt * k2, tt(t_val - 0.5 * dt))
            k4 = sig2(t_val - dt) * unet_forward(sd, x + dt * k3, tt(t_val - dt))
            x = x + (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
        else:
            raise ValueError(method)
    return x


def random_state_dict(in_channels=1, base_channels=32, time_embed_dim=64, num_blocks=3, seed=0):
    """Random-init weights of the ScoreNetwork architecture as a state_dict, built from plain torch.nn layers in the
    order ScoreNetwork.__init__ creates them (models/score_network.py:161-222), so torch.manual_seed(seed) gives the
    same tensors as constructing the module — without importing the reference or the product package (the CPU arm of
    bench.py must not touch product code)."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd = {}

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[prefix + "." + k] = v.detach().clone()

    def block(prefix, cin, cout):
        put(prefix + ".conv1", nn.Conv2d(cin, cout, 3, padding=1))
        put(prefix + ".conv2", nn.Conv2d(cout, cout, 3, padding=1))
        put(prefix + ".norm1", nn.GroupNorm(min(8, cout), cout))
        put(prefix + ".norm2", nn.GroupNorm(min(8, cout), cout))
        put(prefix + ".time_mlp", nn.Linear(time_embed_dim, cout))
        if cin != cout:
            put(prefix + ".residual", nn.Conv2d(cin, cout, 1))

    put("time_mlp.0", nn.Linear(time_embed_dim, time_embed_dim * 2))
    put("time_mlp.2", nn.Linear(time_embed_dim * 2, time_embed_dim))
    put("init_conv", nn.Conv2d(in_channels, base_channels, 3, padding=1))
    widths = [base_channels * 2 ** (i + 1) for i in range(num_blocks)]
    ch = base_channels
    for i, w in enumerate(widths):
        block(f"encoder.{i}", ch, w)
        put(f"downsample.{i}", nn.Conv2d(w, w, 3, stride=2, padding=1))
        ch = w
    block("middle", ch, ch)
    for k, level in enumerate(reversed(range(num_blocks))):
        out_ch = base_channels * 2 ** level
        put(f"upsample.{k}", nn.ConvTranspose2d(ch, ch, 4, stride=2, padding=1))
        block(f"decoder.{k}", ch + widths[level], out_ch)
        ch = out_ch
    put("final", nn.Conv2d(base_channels, in_channels, 3, padding=1))
    return sd
