"""Score U-Net — engine-backed replacement for models/score_network.py.

The modules keep the reference's constructor signatures, sub-module names and parameter shapes
(`time_mlp.{0,2}`, `init_conv`, `encoder.i.{conv1,conv2,norm1,norm2,time_mlp,residual}`,
`downsample.i`, `middle.*`, `upsample.i`, `decoder.i.*`, `final`), so reference checkpoints load
with `load_state_dict` unchanged (models/score_network.py:161-222).

Two execution routes, chosen by autograd mode and never by hardware availability:
  * inference (`torch.no_grad()` / `torch.inference_mode()`, which is how every sampler calls the
    network — samplers/diffusion.py:197): the whole forward runs in libdbs_b200.so from a plan
    compiled from the state_dict (fp32 CUDA-core path, or bf16 tcgen05 implicit-GEMM path).
    Without a B200 this raises; nothing is computed on the host.
  * autograd enabled (training, out of this engine's scope): a differentiable definition of the
    same network in PyTorch ops, so `loss.backward()` keeps working (backend/tests/test_models.py:225-260).
"""
import ctypes
import itertools
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from dbs_b200 import abi, runtime


class SinusoidalTimeEmbedding(nn.Module):
    """t in [0,1] -> [sin(1000 t f_k), cos(1000 t f_k)], f_k = 10000^(-k/(half-1))
    (models/score_network.py:9-41)."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def frequencies(self, device=None) -> torch.Tensor:
        half = self.dim // 2
        return torch.exp(-math.log(10000) * torch.arange(half, device=device) / (half - 1))

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        args = t[:, None] * self.frequencies(t.device)[None, :] * 1000
        return torch.cat([torch.sin(args), torch.cos(args)], dim=-1)


class SelfAttention(nn.Module):
    """Importable for API compatibility (models/__init__.py:1); ScoreNetwork never instantiates it
    (SURVEY.md §0 fact 4), so it has no engine kernel — plain PyTorch definition only."""

    def __init__(self, channels: int, num_heads: int = 4):
        super().__init__()
        assert channels % num_heads == 0, "channels must be divisible by num_heads"
        self.channels, self.num_heads, self.head_dim = channels, num_heads, channels // num_heads
        self.norm = nn.GroupNorm(min(8, channels), channels)
        self.qkv = nn.Conv2d(channels, channels * 3, 1)
        self.proj = nn.Conv2d(channels, channels, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, c, h, w = x.shape
        q, k, v = self.qkv(self.norm(x)).reshape(b, 3, self.num_heads, self.head_dim, h * w).unbind(1)
        attn = torch.softmax(torch.einsum("bnci,bncj->bnij", q, k) * self.head_dim ** -0.5, dim=-1)
        out = torch.einsum("bnij,bncj->bnci", attn, v).reshape(b, c, h, w)
        return self.proj(out) + x


class ConvBlock(nn.Module):
    """conv3x3 -> GN(8) -> SiLU -> + Linear(t_emb) -> conv3x3 -> GN -> SiLU -> + scale * skip
    (models/score_network.py:100-151).  Executed by the engine as part of a ScoreNetwork plan; the
    forward below is the differentiable definition."""

    def __init__(self, in_ch: int, out_ch: int, time_dim: int, residual_scale: float = 1.0):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
        self.norm1 = nn.GroupNorm(min(8, out_ch), out_ch)
        self.norm2 = nn.GroupNorm(min(8, out_ch), out_ch)
        self.time_mlp = nn.Linear(time_dim, out_ch)
        self.act = nn.SiLU()
        self.residual_scale = residual_scale
        self.residual = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else nn.Identity()

    def forward(self, x: torch.Tensor, t_emb: torch.Tensor) -> torch.Tensor:
        h = self.act(self.norm1(self.conv1(x))) + self.time_mlp(t_emb)[:, :, None, None]
        h = self.act(self.norm2(self.conv2(h)))
        return h + self.residual_scale * self.residual(x)


_PLAN_GENERATION = itertools.count(1)


class _Plan:
    """Owns a dbs_unet_plan* and frees it with its last holder (the module, and every captured sampler step that
    was recorded against it).  `generation` is unique per process and never re-used, unlike id()."""

    def __init__(self, handle, version):
        self.handle, self.version = handle, version
        self.generation = next(_PLAN_GENERATION)

    def __del__(self):
        try:
            if self.handle:
                runtime.release_plan(abi.lib.dbs_unet_plan_destroy, self.handle)   # deferred if this thread is capturing a graph
        except Exception:
            pass


PRECISIONS = {"fp32": 0, "bf16": 1}


class ScoreNetwork(nn.Module):
    """Time-conditioned conv U-Net s(x, t) (models/score_network.py:154-268)."""

    def __init__(self, in_channels: int = 1, base_channels: int = 32, time_embed_dim: int = 64, num_blocks: int = 3,
                 output_scale: float = 1.0, normalize_output: bool = False):
        super().__init__()
        self.output_scale = output_scale
        self.normalize_output = normalize_output
        self.in_channels, self.base_channels = in_channels, base_channels
        self.time_embed_dim, self.num_blocks = time_embed_dim, num_blocks

        self.time_embed = SinusoidalTimeEmbedding(time_embed_dim)
        self.time_mlp = nn.Sequential(nn.Linear(time_embed_dim, time_embed_dim * 2), nn.SiLU(),
                                      nn.Linear(time_embed_dim * 2, time_embed_dim))
        self.init_conv = nn.Conv2d(in_channels, base_channels, 3, padding=1)

        self.encoder, self.downsample = nn.ModuleList(), nn.ModuleList()
        widths = [base_channels * 2 ** (i + 1) for i in range(num_blocks)]
        ch = base_channels
        for wdt in widths:
            self.encoder.append(ConvBlock(ch, wdt, time_embed_dim))
            self.downsample.append(nn.Conv2d(wdt, wdt, 3, stride=2, padding=1))
            ch = wdt
        self.middle = ConvBlock(ch, ch, time_embed_dim)
        self.decoder, self.upsample = nn.ModuleList(), nn.ModuleList()
        for level in reversed(range(num_blocks)):
            out_ch = base_channels * 2 ** level
            self.upsample.append(nn.ConvTranspose2d(ch, ch, 4, stride=2, padding=1))
            self.decoder.append(ConvBlock(ch + widths[level], out_ch, time_embed_dim))
            ch = out_ch
        self.final = nn.Conv2d(base_channels, in_channels, 3, padding=1)

        self.precision = os.environ.get("DBS_PRECISION", "fp32")
        self._plan = None
        self._workspace = None

    # ---------------------------------------------------------------- engine route
    def _param_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine_plan(self):
        """Compile (or reuse) the engine plan for the current parameter values."""
        runtime.require_cuda()
        ver = self._param_version()
        if self._plan is not None and self._plan.version == ver:
            return self._plan.handle
        runtime.flush_released()    # plans whose finaliser ran during a graph capture
        sd = {k: v.detach().to(torch.float32).contiguous() for k, v in self.state_dict().items()}
        keep = []
        arr = (abi.Param * len(sd))()
        for n, (k, v) in enumerate(sd.items()):
            name = k.encode()
            keep.append(name)
            arr[n] = abi.Param(name, v.data_ptr(), v.numel(), 1 if v.is_cuda else 0)
        cfg = abi.UnetConfig(self.in_channels, self.base_channels, self.time_embed_dim, self.num_blocks,
                             float(self.output_scale), 1 if self.normalize_output else 0)
        freqs = self.time_embed.frequencies().to(torch.float32).contiguous()
        handle = ctypes.c_void_p()
        with runtime.CAPTURE_LOCK:      # the build synchronises the device: never while another thread captures a graph
            abi.check(abi.lib.dbs_unet_plan_create(ctypes.byref(cfg), arr, len(sd), freqs.data_ptr(),
                                                   ctypes.byref(handle)), "unet_plan_create")
        self._plan = _Plan(handle, ver)
        return handle

    def _workspace_bytes(self, plan, B, H, W, prec):
        need = abi.lib.dbs_unet_workspace_bytes(plan, B, H, W, prec)
        if need < 0:
            abi.check(abi.ERR_INVALID, "unet_workspace_bytes")
        return int(need)

    def _get_workspace(self, plan, B, H, W, prec):
        """The module's own scratch for EAGER forwards, regrown on demand.  Captured graphs never use it: a graph has
        its workspace address baked into every kernel node, so each captured step owns one (engine_workspace)."""
        need = self._workspace_bytes(plan, B, H, W, prec)
        ws = self._workspace
        if ws is None or ws.numel() < need or not ws.is_cuda:
            ws = torch.empty(need, dtype=torch.uint8, device=runtime.require_cuda())
            self._workspace = ws
        return ws

    def engine_workspace(self, shape, precision=None) -> torch.Tensor:
        """A fresh workspace for forwards of `shape`, owned by the caller (pass it as engine_forward(workspace=...))."""
        B, _, H, W = shape
        need = self._workspace_bytes(self.engine_plan(), B, H, W, PRECISIONS[precision or self.precision])
        return torch.empty(max(need, 16), dtype=torch.uint8, device=runtime.require_cuda())

    def engine_forward(self, x: torch.Tensor, t: torch.Tensor, precision=None, out=None, tap=None, workspace=None):
        """score (and optionally one named intermediate, fp32 NCHW) from the sm_100a kernels.
        x (B,1,H,W) fp32, t (B,) or (1,) fp32; device tensors in -> device tensor out."""
        prec = PRECISIONS[precision or self.precision]
        plan = self.engine_plan()
        xd = runtime.to_device(x)
        td = runtime.to_device(t).reshape(-1)
        B, Cc, H, W = xd.shape
        if Cc != 1:
            raise NotImplementedError("ScoreNetwork engine: only in_channels == 1 is built")
        if td.numel() not in (1, B):
            raise ValueError(f"t must have 1 or {B} entries, got {td.numel()}")
        if td.numel() == B and B > 1 and bool((td == td[0]).all()):
            td = td[:1].contiguous()  # the sampler's case: one shared time (samplers/diffusion.py:223)
        ws = workspace if workspace is not None else self._get_workspace(plan, B, H, W, prec)
        score = out if out is not None else torch.empty_like(xd)
        if tap is None:
            abi.check(abi.lib.dbs_unet_forward(plan, xd.data_ptr(), td.data_ptr(), td.numel(), score.data_ptr(), B, H,
                                               W, prec, ws.data_ptr(), ws.numel(), runtime.current_stream()),
                      "unet_forward")
            return score
        name, shape = tap
        tap_out = torch.empty(shape, dtype=torch.float32, device=xd.device)
        abi.check(abi.lib.dbs_unet_forward_tap(plan, xd.data_ptr(), td.data_ptr(), td.numel(), score.data_ptr(), B, H,
                                               W, prec, ws.data_ptr(), ws.numel(), name.encode(), tap_out.data_ptr(),
                                               tap_out.numel(), runtime.current_stream()), "unet_forward_tap")
        return score, tap_out

    # ---------------------------------------------------------------- differentiable definition
    def _autograd_forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        t_emb = self.time_mlp(self.time_embed(t))
        h = self.init_conv(x)
        skips = []
        for block, down in zip(self.encoder, self.downsample):
            h = block(h, t_emb)
            skips.append(h)
            h = down(h)
        h = self.middle(h, t_emb)
        for up, block in zip(self.upsample, self.decoder):
            skip = skips.pop()
            h = up(h)
            if h.shape[-2:] != skip.shape[-2:]:
                h = F.interpolate(h, size=skip.shape[-2:], mode="bilinear", align_corners=False)
            h = block(torch.cat([h, skip], dim=1), t_emb)
        score = self.final(h)
        if self.normalize_output:
            score = score / (score.std(dim=(-1, -2), keepdim=True) + 1e-8)
        return self.output_scale * score

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled():
            return self._autograd_forward(x, t)
        return runtime.like_input(self.engine_forward(x, t), x)
