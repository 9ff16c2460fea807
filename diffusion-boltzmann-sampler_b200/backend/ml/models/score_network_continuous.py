"""Particle score network — engine-backed replacement for models/score_network_continuous.py.

Constructor signatures, sub-module names and parameter shapes follow the reference
(`pos_embed.freqs`, `time_mlp.{0,2}`, `input_proj`, `blocks.i.{fc1,fc2,norm1,norm2,time_mlp}`,
`attention_layers.i.{norm,qkv,proj}`, `output_proj.{0,1,3}`; score_network_continuous.py:186-253),
so a reference state_dict loads unchanged.

Execution routes, chosen by autograd mode and never by hardware availability (same rule as
ScoreNetwork): under `torch.no_grad()` — how every sampler calls the network — the forward runs in
libdbs_b200.so (csrc/csn.cu: fused Fourier-embedding/LayerNorm/SiLU linear kernels and a
shared-memory attention kernel); with autograd enabled the differentiable PyTorch definition below is
used (training is outside the engine's scope).  Without a B200 the engine route raises.
"""
import ctypes
import itertools
import math
import os

import torch
import torch.nn as nn

from dbs_b200 import abi, runtime


class PeriodicPositionEmbedding(nn.Module):
    """Fourier features [sin(f a_x) | cos(f a_x) | sin(f a_y) | cos(f a_y)], a = (x + 1) * pi, truncated
    or zero-padded to `dim` (score_network_continuous.py:17-75)."""

    def __init__(self, dim: int, n_frequencies: int = 16, max_freq: float = 10.0):
        super().__init__()
        assert dim % 4 == 0, "Embedding dim must be divisible by 4"
        self.dim, self.n_frequencies = dim, n_frequencies
        self.register_buffer("freqs", torch.linspace(1.0, max_freq, n_frequencies))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ang = ((x + 1) * math.pi)[..., None] * self.freqs
        ax, ay = ang[..., 0, :], ang[..., 1, :]
        emb = torch.cat([torch.sin(ax), torch.cos(ax), torch.sin(ay), torch.cos(ay)], dim=-1)
        have = emb.shape[-1]
        if have > self.dim:
            return emb[..., : self.dim]
        if have < self.dim:
            return torch.cat([emb, emb.new_zeros(*emb.shape[:-1], self.dim - have)], dim=-1)
        return emb


class SinusoidalTimeEmbedding(nn.Module):
    """Same time embedding as the U-Net's (score_network_continuous.py:78-101)."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def frequencies(self, device=None) -> torch.Tensor:
        half = self.dim // 2
        return torch.exp(-math.log(10000) * torch.arange(half, device=device) / (half - 1))

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        args = t[:, None] * self.frequencies(t.device)[None, :] * 1000
        return torch.cat([torch.sin(args), torch.cos(args)], dim=-1)


class MLPBlock(nn.Module):
    """x + SiLU(LN(fc2(SiLU(LN(fc1(x))) + Linear(t_emb)))) (score_network_continuous.py:104-136)."""

    def __init__(self, hidden_dim: int, time_dim: int, dropout: float = 0.0):
        super().__init__()
        self.fc1 = nn.Linear(hidden_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.norm1 = nn.LayerNorm(hidden_dim)
        self.norm2 = nn.LayerNorm(hidden_dim)
        self.time_mlp = nn.Linear(time_dim, hidden_dim)
        self.act = nn.SiLU()
        self.dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()

    def forward(self, x: torch.Tensor, t_emb: torch.Tensor) -> torch.Tensor:
        h = self.act(self.norm1(self.fc1(x))) + self.time_mlp(t_emb)[:, None, :]
        return x + self.dropout(self.act(self.norm2(self.fc2(h))))


class ParticleAttention(nn.Module):
    """Pre-LN multi-head self-attention over the particles of a configuration with a residual
    (score_network_continuous.py:139-180)."""

    def __init__(self, hidden_dim: int, num_heads: int = 4):
        super().__init__()
        assert hidden_dim % num_heads == 0
        self.hidden_dim, self.num_heads, self.head_dim = hidden_dim, num_heads, hidden_dim // num_heads
        self.norm = nn.LayerNorm(hidden_dim)
        self.qkv = nn.Linear(hidden_dim, hidden_dim * 3)
        self.proj = nn.Linear(hidden_dim, hidden_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, n, _ = x.shape
        q, k, v = self.qkv(self.norm(x)).view(b, n, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        w = torch.softmax((q @ k.transpose(-2, -1)) * self.head_dim ** -0.5, dim=-1)
        return self.proj((w @ v).transpose(1, 2).reshape(b, n, self.hidden_dim)) + x


_PLAN_GENERATION = itertools.count(1)


class _Plan:
    def __init__(self, handle, version):
        self.handle, self.version = handle, version
        self.generation = next(_PLAN_GENERATION)   # unique per process, never re-used (unlike id())

    def __del__(self):
        try:
            if self.handle:
                runtime.release_plan(abi.lib.dbs_csn_plan_destroy, self.handle)   # deferred if this thread is capturing a graph
        except Exception:
            pass


class ContinuousScoreNetwork(nn.Module):
    """s(x, t) for particle positions x (B, N, 2) in [-1, 1] (score_network_continuous.py:183-282)."""

    def __init__(self, n_particles: int, hidden_dim: int = 128, time_embed_dim: int = 64, pos_embed_dim: int = 64,
                 num_blocks: int = 4, num_heads: int = 4, dropout: float = 0.0, use_attention: bool = True):
        super().__init__()
        self.n_particles, self.hidden_dim, self.use_attention = n_particles, hidden_dim, use_attention
        self.time_embed_dim, self.pos_embed_dim = time_embed_dim, pos_embed_dim
        self.num_blocks, self.num_heads, self.dropout_p = num_blocks, num_heads, dropout
        self.pos_embed = PeriodicPositionEmbedding(dim=pos_embed_dim, n_frequencies=pos_embed_dim // 4)
        self.time_embed = SinusoidalTimeEmbedding(time_embed_dim)
        self.time_mlp = nn.Sequential(nn.Linear(time_embed_dim, time_embed_dim * 2), nn.SiLU(),
                                      nn.Linear(time_embed_dim * 2, time_embed_dim))
        self.input_proj = nn.Linear(pos_embed_dim + 2, hidden_dim)
        self.blocks = nn.ModuleList(MLPBlock(hidden_dim, time_embed_dim, dropout) for _ in range(num_blocks))
        self.attention_layers = (nn.ModuleList(ParticleAttention(hidden_dim, num_heads) for _ in range(num_blocks))
                                 if use_attention else None)
        self.output_proj = nn.Sequential(nn.LayerNorm(hidden_dim), nn.Linear(hidden_dim, hidden_dim // 2), nn.SiLU(),
                                         nn.Linear(hidden_dim // 2, 2))
        self.precision = os.environ.get("DBS_PRECISION_CSN", "fp32")
        self._plan = None
        self._workspace = None

    # ---------------------------------------------------------------- engine route
    def _param_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine_plan(self):
        runtime.require_cuda()
        ver = self._param_version()
        if self._plan is not None and self._plan.version == ver:
            return self._plan.handle
        runtime.flush_released()    # plans whose finaliser ran during a graph capture
        sd = {k: v.detach().to(torch.float32).contiguous() for k, v in self.state_dict().items()}
        keep, arr = [], (abi.Param * len(sd))()
        for n, (k, v) in enumerate(sd.items()):
            name = k.encode()
            keep.append(name)
            arr[n] = abi.Param(name, v.data_ptr(), v.numel(), 1 if v.is_cuda else 0)
        cfg = abi.CsnConfig(self.n_particles, self.hidden_dim, self.time_embed_dim, self.pos_embed_dim,
                            self.num_blocks, self.num_heads, 1 if self.use_attention else 0)
        pos_f = self.pos_embed.freqs.detach().to("cpu", torch.float32).contiguous()
        t_f = self.time_embed.frequencies().to(torch.float32).contiguous()
        handle = ctypes.c_void_p()
        with runtime.CAPTURE_LOCK:      # the build synchronises the device: never while another thread captures a graph
            abi.check(abi.lib.dbs_csn_plan_create(ctypes.byref(cfg), arr, len(sd), pos_f.data_ptr(), t_f.data_ptr(),
                                                  ctypes.byref(handle)), "csn_plan_create")
        self._plan = _Plan(handle, ver)
        return handle

    def engine_workspace(self, shape, precision=None) -> torch.Tensor:
        """A fresh workspace for forwards of `shape` = (B, N, 2), owned by the caller (captured graphs bake the
        address into their kernel nodes, so each captured sampler step owns one; sized for per-sample t)."""
        B, N, _ = shape
        need = abi.lib.dbs_csn_workspace_bytes(self.engine_plan(), B, N, B)
        if need < 0:
            abi.check(abi.ERR_INVALID, "csn_workspace_bytes")
        return torch.empty(max(int(need), 16), dtype=torch.uint8, device=runtime.require_cuda())

    def engine_forward(self, x: torch.Tensor, t: torch.Tensor, precision=None, out=None, workspace=None):
        """score from the sm_100a kernels.  x (B, N, 2) fp32, t (B,) or (1,); device tensors in -> device
        tensor out.  precision: "fp32" (default; CUDA-core parity path) or "bf16" (tcgen05 path for the MLP and
        attention layers: hidden_dim 128, head_dim 32, N a multiple of 16 up to 256)."""
        prec = {"fp32": 0, "bf16": 1}[precision or self.precision]
        if self.training and self.dropout_p > 0:
            raise NotImplementedError("ContinuousScoreNetwork engine: dropout is inference-only (call .eval())")
        plan = self.engine_plan()
        xd = runtime.to_device(x)
        td = runtime.to_device(t).reshape(-1)
        if xd.dim() != 3 or xd.shape[-1] != 2:
            raise ValueError(f"x must be (batch, n_particles, 2), got {tuple(xd.shape)}")
        B, N, _ = xd.shape
        if td.numel() not in (1, B):
            raise ValueError(f"t must have 1 or {B} entries, got {td.numel()}")
        if td.numel() == B and B > 1 and bool((td == td[0]).all()):
            td = td[:1].contiguous()  # the sampler's case: one shared time (samplers/diffusion.py:223)
        need = abi.lib.dbs_csn_workspace_bytes(plan, B, N, td.numel())
        if need < 0:
            abi.check(abi.ERR_INVALID, "csn_workspace_bytes")
        ws = workspace
        if ws is None:
            ws = self._workspace     # eager forwards only; captured graphs pass the workspace they own
            if ws is None or ws.numel() < need or not ws.is_cuda:
                ws = self._workspace = torch.empty(int(need), dtype=torch.uint8, device=xd.device)
        elif ws.numel() < need:
            raise ValueError(f"engine_forward: workspace of {ws.numel()} bytes, {need} needed")
        score = out if out is not None else torch.empty_like(xd)
        abi.check(abi.lib.dbs_csn_forward(plan, xd.data_ptr(), td.data_ptr(), td.numel(), score.data_ptr(), B, N, prec,
                                          ws.data_ptr(), ws.numel(), runtime.current_stream()), "csn_forward")
        return score

    # ---------------------------------------------------------------- differentiable definition
    def _autograd_forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        t_emb = self.time_mlp(self.time_embed(t))
        h = self.input_proj(torch.cat([self.pos_embed(x), x], dim=-1))
        for i, block in enumerate(self.blocks):
            h = block(h, t_emb)
            if self.use_attention:
                h = self.attention_layers[i](h)
        return self.output_proj(h)

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled():
            return self._autograd_forward(x, t)
        return runtime.like_input(self.engine_forward(x, t), x)
