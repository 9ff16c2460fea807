"""Reverse-diffusion samplers — engine-backed replacement for samplers/diffusion.py.

Every public method of the reference's `DiffusionSampler` / `PretrainedDiffusionSampler`
(samplers/diffusion.py:15-772) is kept with the same signature and return convention.  The state
x lives on the GPU for the whole trajectory: each step is one U-Net forward in libdbs_b200.so
plus one fused Euler–Maruyama kernel (noise drawn in-register from Philox unless injected).
`device` keeps its reference meaning for the caller — where results are returned ("cpu" by
default) — while all arithmetic happens on the current CUDA device; without one the samplers
raise.

Additive keyword extensions: `precision` ("fp32" parity path | "bf16" tcgen05 path), `seed`,
`x0`/`noises` injection ports on `sample` (parity tests), `sample_offset` (multi-GPU sharding with
global sample indices so results do not depend on the world size).
"""
import gc
import threading
from typing import Any, Dict, Generator, Optional, Tuple

import numpy as np
import torch

from dbs_b200 import abi, runtime
from ..models.diffusion import DiffusionProcess

_CAPTURE_LOCK = threading.Lock()   # CUDA-graph captures of different samplers never overlap (see _graph_step)


def _f32(v: float) -> float:
    """A Python float after rounding to fp32 — what `python_float * fp32_tensor` multiplies by."""
    return float(np.float32(v))


class DiffusionSampler:
    def __init__(self, score_network: torch.nn.Module, diffusion: Optional[DiffusionProcess] = None,
                 num_steps: int = 100, device: str = "cpu", precision: Optional[str] = None):
        self.model = score_network
        self.model.eval()
        self.diffusion = diffusion or DiffusionProcess()
        self.num_steps = num_steps
        self.device = device
        self.precision = precision  # None -> the network's default (env DBS_PRECISION or fp32)
        self._training_temperature: Optional[float] = None
        # One U-Net forward is ~70 kernel launches: the forward is captured once per shape into a CUDA
        # graph and replayed every step (measured +7 % at batch 1024, +10 % at batch 1-8).
        # True / False / "auto" (= True; kept as a switch for debugging).
        self.use_graph = "auto"
        self._graphs: Dict[Any, Any] = {}
        self.graph_kernel_launches = 0   # kernels executed through graph replays (for launch accounting)

    # ------------------------------------------------------------------ temperature scaling
    @property
    def training_temperature(self) -> Optional[float]:
        return self._training_temperature

    @training_temperature.setter
    def training_temperature(self, value: float) -> None:
        self._training_temperature = value

    def compute_temperature_scale(self, target_temperature: float, method: str = "linear") -> float:
        """Noise multiplier for sampling at another temperature (samplers/diffusion.py:47-81);
        "boltzmann" is the same ratio as "linear" there, and with no training temperature the raw
        target is returned."""
        if self._training_temperature is None:
            return target_temperature
        ratio = target_temperature / self._training_temperature
        if method in ("linear", "boltzmann"):
            return ratio
        if method == "sqrt":
            return ratio ** 0.5
        raise ValueError(f"Unknown temperature scaling method: {method}")

    def sample_at_temperature(self, shape, target_temperature: float, scaling_method: str = "linear"):
        return self.sample(shape, temperature=self.compute_temperature_scale(target_temperature, scaling_method))

    # ------------------------------------------------------------------ checkpoints
    @classmethod
    def from_checkpoint(cls, checkpoint_path: str, num_steps: int = 100, device: str = "cpu") -> "DiffusionSampler":
        """Accepts the three layouts of samplers/diffusion.py:104-160."""
        from ..models.score_network import ScoreNetwork

        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        cfg = ckpt.get("model_config") if isinstance(ckpt, dict) else None
        model = ScoreNetwork(**cfg) if cfg else ScoreNetwork(in_channels=1, base_channels=32, time_embed_dim=64,
                                                              num_blocks=3)
        for key in ("model_state_dict", "state_dict"):
            if key in ckpt:
                model.load_state_dict(ckpt[key])
                break
        else:
            model.load_state_dict(ckpt)
        diffusion = DiffusionProcess(**ckpt["diffusion_config"]) if "diffusion_config" in ckpt else None
        sampler = cls(score_network=model, diffusion=diffusion, num_steps=num_steps, device=device)
        sampler.training_temperature = ckpt.get("training_temperature")
        return sampler

    def save_checkpoint(self, path: str, model_config=None, diffusion_config=None, training_temperature=None,
                        extra_info=None) -> None:
        """Same dictionary layout as samplers/diffusion.py:162-195."""
        ckpt = {
            "model_state_dict": self.model.state_dict(),
            "num_steps": self.num_steps,
            "diffusion_config": diffusion_config or self.diffusion.get_config(),
        }
        if model_config:
            ckpt["model_config"] = model_config
        if training_temperature is not None:
            ckpt["training_temperature"] = training_temperature
        if extra_info:
            ckpt.update(extra_info)
        torch.save(ckpt, path)

    # ------------------------------------------------------------------ engine plumbing
    def _out(self, x: torch.Tensor) -> torch.Tensor:
        return x if str(self.device).startswith("cuda") else x.cpu()

    def _result(self, x: torch.Tensor) -> torch.Tensor:
        """Detach a result from the sampler's persistent state buffer."""
        return x.clone() if str(self.device).startswith("cuda") else x.cpu()

    @staticmethod
    def _fresh_seed() -> int:
        # drawn from torch's global generator: torch.manual_seed() makes sampling reproducible,
        # like the torch.randn calls it replaces (samplers/diffusion.py:216, 238)
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    def _score(self, x: torch.Tensor, t_val: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        t = torch.full((1,), t_val, dtype=torch.float32, device=x.device)
        return self.model.engine_forward(x, t, precision=self.precision, out=out)

    def _sigma(self, t_val: float) -> float:
        _, s = self.diffusion.noise_level(torch.full((1,), t_val))
        return s[0].item()

    @staticmethod
    def _gaussian(shape, seed: int, step: int, sample_offset: int = 0) -> torch.Tensor:
        dev = runtime.require_cuda()
        x = torch.empty(shape, dtype=torch.float32, device=dev)
        n = shape[0]
        abi.check(abi.lib.dbs_gaussian_fill(x.data_ptr(), n, x.numel() // max(n, 1), seed, sample_offset, step,
                                            runtime.current_stream()), "gaussian_fill")
        return x

    def _em(self, x, score, sigma: float, temperature: float, add_noise: bool, seed: int, step: int,
            noise: Optional[torch.Tensor] = None, sample_offset: int = 0) -> None:
        """x <- x + sigma^2 * score * dt + sigma * sqrt(dt) * z * temperature, in place."""
        dt = 1.0 / self.num_steps
        n = x.shape[0]
        abi.check(abi.lib.dbs_em_step(x.data_ptr(), score.data_ptr(), runtime.ptr(noise), n, x.numel() // max(n, 1),
                                      _f32(sigma ** 2), _f32(dt), _f32(sigma * dt ** 0.5), _f32(temperature),
                                      1 if add_noise else 0, seed, sample_offset, step, runtime.current_stream()),
                  "em_step")

    @staticmethod
    def _lincomb(terms, out: Optional[torch.Tensor] = None, sign_first: bool = False) -> torch.Tensor:
        """out = sum_i a_i * x_i over up to five (a_i, x_i) terms in ONE engine pass (dbs_lincomb); with
        sign_first the first term is a_0 * sign(x_0).  `out` may alias a term."""
        import ctypes
        xs = [t[1] for t in terms]
        out = torch.empty_like(xs[0]) if out is None else out
        ptrs = (ctypes.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
        coef = (ctypes.c_float * len(xs))(*[float(t[0]) for t in terms])
        abi.check(abi.lib.dbs_lincomb(out.data_ptr(), out.numel(), len(xs), ptrs, coef, 1 if sign_first else 0,
                                      runtime.current_stream()), "lincomb")
        return out

    def _graph_step(self, shape):
        """Persistent (x, score, t) buffers and, when enabled, a captured graph of one forward for
        this shape.  Returns (x, score, t_dev, graph_or_None)."""
        dev = runtime.require_cuda()
        want = self.use_graph
        if want == "auto":
            want = True
        plan_version = None
        if hasattr(self.model, "engine_plan"):
            self.model.engine_plan()
            plan_version = id(self.model._plan)
        key = (tuple(shape), self.precision, bool(want), plan_version, torch.cuda.current_device())
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs = {k: v for k, v in self._graphs.items() if k[3] == plan_version}  # drop stale plans
            x = torch.empty(shape, dtype=torch.float32, device=dev)
            score = torch.empty_like(x)
            t_dev = torch.zeros(1, dtype=torch.float32, device=dev)
            graph = None
            if want and hasattr(self.model, "engine_forward") and shape[0] > 0:   # nothing to capture for an empty batch
                x.zero_()
                for _ in range(2):          # warm-up: workspace allocation, one-time function attributes
                    self.model.engine_forward(x, t_dev, precision=self.precision, out=score)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                # thread_local: the capture records kernel launches only (workspace and buffers exist already), so
                # CUDA calls of OTHER threads — an MCMC job on the server's thread pool allocating its lattices
                # (routes/training.py:191) — must neither fail nor invalidate it, as they would in "global" mode.
                # torch.cuda.graph captures on ONE process-wide side stream by default: two samplers capturing from
                # two threads would begin a capture on a stream that is already capturing, so captures are serialised
                # and each gets a stream of its own.
                # The cyclic garbage collector may run finalisers (cudaFree of a dropped plan, ...) on whichever thread
                # happens to allocate — also on this one in mid-capture, where they are not permitted: it is paused for
                # the capture (plan finalisers additionally defer themselves, runtime.release_plan).
                with _CAPTURE_LOCK:
                    gc_was_on = gc.isenabled()
                    gc.disable()
                    try:
                        before = abi.launch_count()
                        with torch.cuda.graph(graph, stream=torch.cuda.Stream(), capture_error_mode="thread_local"):
                            self.model.engine_forward(x, t_dev, precision=self.precision, out=score)
                        graph.n_kernels = abi.launch_count() - before    # kernels recorded in this graph
                    finally:
                        if gc_was_on:
                            gc.enable()
            ent = (x, score, t_dev, graph)
            self._graphs[key] = ent
        return ent

    def _reverse_sde(self, shape, temperature, seed, x0=None, noises=None, sample_offset=0):
        """Generator over the Euler–Maruyama trajectory: first (num_steps + 1, x_1) — the Gaussian
        start — then (i, x) after step i = num_steps .. 1.  x is a persistent buffer updated in place
        (callers copy what they keep)."""
        runtime.require_cuda()
        shape = tuple(shape)
        x, score, t_dev, graph = self._graph_step(shape)
        if x0 is not None:
            x.copy_(runtime.to_device(x0).reshape(shape))
        else:
            n = shape[0]
            abi.check(abi.lib.dbs_gaussian_fill(x.data_ptr(), n, x.numel() // max(n, 1), seed, sample_offset, abi.STEP_X0,
                                                runtime.current_stream()), "gaussian_fill")
        sig = self.diffusion.sigma_table(self.num_steps)
        t_table = torch.tensor([i / self.num_steps for i in range(self.num_steps, 0, -1)], dtype=torch.float32,
                               device=x.device)
        yield self.num_steps + 1, x
        for k, i in enumerate(range(self.num_steps, 0, -1)):
            if graph is not None:
                t_dev.copy_(t_table[k:k + 1])
                graph.replay()
                self.graph_kernel_launches += graph.n_kernels
            else:
                self._score(x, i / self.num_steps, out=score)
            nz = None
            if noises is not None and i > 1:
                nz = runtime.to_device(noises[k])
            self._em(x, score, sig[k], temperature, i > 1, seed, i, noise=nz, sample_offset=sample_offset)
            yield i, x

    # ------------------------------------------------------------------ samplers
    @torch.no_grad()
    def sample(self, shape: Tuple[int, ...], temperature: float = 1.0, seed: Optional[int] = None, x0=None,
               noises=None, sample_offset: int = 0) -> torch.Tensor:
        """Euler–Maruyama reverse SDE (samplers/diffusion.py:197-245)."""
        seed = self._fresh_seed() if seed is None else seed
        x = None
        for _, x in self._reverse_sde(tuple(shape), temperature, seed, x0, noises, sample_offset):
            pass
        return self._result(x)

    @torch.no_grad()
    def sample_with_trajectory(self, shape: Tuple[int, ...], temperature: float = 1.0, yield_every: int = 1,
                               seed: Optional[int] = None) -> Generator[Tuple[torch.Tensor, float], None, None]:
        """Yield (x_t, t): first (x_1, 1.0), then after every `yield_every`-th step
        (samplers/diffusion.py:247-291)."""
        seed = self._fresh_seed() if seed is None else seed
        for i, x in self._reverse_sde(tuple(shape), temperature, seed):
            if i > self.num_steps:
                yield self._result(x), 1.0
            elif i % yield_every == 0 or i == 1:
                yield self._result(x), (i - 1) / self.num_steps

    @torch.no_grad()
    def sample_trajectory_with_metadata(self, shape: Tuple[int, ...], temperature: float = 1.0, n_frames: int = 50,
                                        frame_spacing: str = "linear", include_statistics: bool = False,
                                        ising_model=None, seed: Optional[int] = None
                                        ) -> Generator[Dict[str, Any], None, None]:
        """Frames {x, t, step, progress, sigma[, statistics]} at linear / log / cosine spaced times
        (samplers/diffusion.py:293-397)."""
        N = self.num_steps
        if frame_spacing == "linear":
            times = np.linspace(1.0, 0.0, n_frames + 1)
        elif frame_spacing == "log":
            times = np.sort(np.clip(1.0 - np.logspace(-2, 0, n_frames + 1), 0.0, 1.0))[::-1]
        elif frame_spacing == "cosine":
            times = 0.5 * (1 + np.cos(np.linspace(0, np.pi, n_frames + 1)))
        else:
            raise ValueError(f"Unknown frame_spacing: {frame_spacing}")
        wanted = set((times * N).astype(int)) | {N, 0}
        seed = self._fresh_seed() if seed is None else seed

        def frame(x, step):
            t = step / N
            d = {"x": self._result(x), "t": t, "step": step, "progress": 1.0 - t, "sigma": self._sigma(t)}
            if include_statistics and ising_model:
                d["statistics"] = self.compute_sample_statistics(d["x"], ising_model)
            return d

        for i, x in self._reverse_sde(tuple(shape), temperature, seed):
            if i > N:
                yield frame(x, N)
            elif (i - 1) in wanted or i == 1:
                yield frame(x, i - 1)

    @torch.no_grad()
    def sample_predictor_corrector(self, shape: Tuple[int, ...], temperature: float = 1.0, n_corrector_steps: int = 1,
                                   corrector_snr: float = 0.16, seed: Optional[int] = None) -> torch.Tensor:
        """EM predictor + Langevin corrector (samplers/diffusion.py:399-451).  The corrector's
        step size needs a batch-wide norm, so it is a device reduction followed by an axpy."""
        seed = self._fresh_seed() if seed is None else seed
        shape = tuple(shape)
        N = self.num_steps
        noise_norm = (int(np.prod(shape[1:]))) ** 0.5
        for i, x in self._reverse_sde(shape, temperature, seed):
            if 1 < i <= N:
                for c in range(n_corrector_steps):
                    score = self._score(x, i / N)
                    grad_norm = score.flatten(1).norm(dim=1).mean()
                    step_size = (corrector_snr * noise_norm / grad_norm) ** 2 * 2
                    z = self._gaussian(shape, seed, N + 1 + (N - i) * n_corrector_steps + c)
                    x.add_(step_size * score + (2 * step_size) ** 0.5 * z * temperature)
        return self._result(x)

    @torch.no_grad()
    def sample_ode(self, shape: Tuple[int, ...], method: str = "euler", seed: Optional[int] = None) -> torch.Tensor:
        """Probability-flow style deterministic integrators: euler / heun / rk4
        (samplers/diffusion.py:453-521)."""
        if method not in ("euler", "heun", "rk4"):
            raise ValueError(f"Unknown ODE method: {method}")
        seed = self._fresh_seed() if seed is None else seed
        shape = tuple(shape)
        N = self.num_steps
        dt = 1.0 / N
        x = self._gaussian(shape, seed, abi.STEP_X0)
        score = torch.empty_like(x)

        for i in range(N, 0, -1):
            t_val = i / N
            if method == "euler":
                self._score(x, t_val, out=score)
                self._em(x, score, self._sigma(t_val), 0.0, False, seed, i)
            elif method == "heun":
                # x <- x + dt/2 (sig^2 s(x, t) + sig_n^2 s(x + dt sig^2 s(x, t), t - dt)): two forwards, two fused passes
                c0, c1 = self._sigma(t_val) ** 2, self._sigma((i - 1) / N) ** 2
                s0 = self._score(x, t_val)
                s1 = self._score(self._lincomb([(1.0, x), (c0 * dt, s0)]), (i - 1) / N)
                self._lincomb([(1.0, x), (0.5 * c0 * dt, s0), (0.5 * c1 * dt, s1)], out=x)
            else:
                # classical RK4 on dx/dt = sig(t)^2 s(x, t): four forwards, four fused passes
                ca, cm, cb = self._sigma(t_val) ** 2, self._sigma(t_val - 0.5 * dt) ** 2, self._sigma(t_val - dt) ** 2
                s1 = self._score(x, t_val)
                s2 = self._score(self._lincomb([(1.0, x), (0.5 * dt * ca, s1)]), t_val - 0.5 * dt)
                s3 = self._score(self._lincomb([(1.0, x), (0.5 * dt * cm, s2)]), t_val - 0.5 * dt)
                s4 = self._score(self._lincomb([(1.0, x), (dt * cm, s3)]), t_val - dt)
                self._lincomb([(1.0, x), (dt / 6 * ca, s1), (dt / 3 * cm, s2), (dt / 3 * cm, s3), (dt / 6 * cb, s4)], out=x)
        return self._out(x)

    # ------------------------------------------------------------------ post-processing
    def compute_sample_statistics(self, samples: torch.Tensor, ising_model=None) -> Dict[str, Any]:
        """Summary statistics of a batch (samplers/diffusion.py:523-578); energies go through the
        engine's observable kernel."""
        s2 = samples.squeeze(1) if samples.dim() == 4 else samples
        mag = s2.mean(dim=(-1, -2))
        stats = {
            "mean": samples.mean().item(),
            "std": samples.std().item(),
            "min": samples.min().item(),
            "max": samples.max().item(),
            "magnetization_mean": mag.mean().item(),
            "magnetization_std": mag.std().item(),
            "magnetization_abs_mean": mag.abs().mean().item(),
            "fraction_positive": (samples > 0).float().mean().item(),
            "sample_diversity": samples.var(dim=0).mean().item() if samples.shape[0] > 1 else 0.0,
        }
        if ising_model is not None:
            e = ising_model.energy_per_spin(s2)
            stats["energy_mean"] = e.mean().item()
            stats["energy_std"] = e.std().item()
        return stats

    def discretize_spins(self, x: torch.Tensor, method: str = "sign", threshold: float = 0.0,
                         sharpness: float = 10.0) -> torch.Tensor:
        """sign / tanh / gumbel / stochastic discretisers (samplers/diffusion.py:580-620);
        elementwise on the tensor's own device; sign(0) stays 0 as in the reference."""
        d = x - threshold
        if method == "sign":
            return torch.sign(d)
        if method == "tanh":
            return torch.tanh(sharpness * d)
        if method == "gumbel":
            return 2 * torch.softmax(torch.stack([-d, d], dim=-1) * sharpness, dim=-1)[..., 1] - 1
        if method == "stochastic":
            p = torch.sigmoid(sharpness * d)
            return torch.where(torch.rand_like(p) < p, torch.ones_like(x), -torch.ones_like(x))
        raise ValueError(f"Unknown discretization method: {method}")

    def sample_ising(self, batch_size: int = 1, lattice_size: int = 32, temperature: float = 1.0,
                     discretize: bool = True, discretize_method: str = "sign") -> torch.Tensor:
        """(batch, L, L) Ising samples (samplers/diffusion.py:622-651)."""
        x = self.sample((batch_size, 1, lattice_size, lattice_size), temperature=temperature)
        if discretize:
            x = self.discretize_spins(x, method=discretize_method)
        return x.squeeze(1)

    def generate_batch(self, n_samples: int, lattice_size: int = 32, temperature: float = 1.0, batch_size: int = 16,
                       sampling_method: str = "sde", discretize: bool = True, progress_callback=None) -> torch.Tensor:
        """Chunked generation of n_samples lattices with sde / ode / pc (samplers/diffusion.py:653-709)."""
        if sampling_method not in ("sde", "ode", "pc"):
            raise ValueError(f"Unknown sampling method: {sampling_method}")
        chunks, done = [], 0
        while done < n_samples:
            cur = min(batch_size, n_samples - done)
            shape = (cur, 1, lattice_size, lattice_size)
            if sampling_method == "sde":
                x = self.sample(shape, temperature=temperature)
            elif sampling_method == "ode":
                x = self.sample_ode(shape)
            else:
                x = self.sample_predictor_corrector(shape, temperature=temperature)
            if discretize:
                x = self.discretize_spins(x)
            chunks.append(x.squeeze(1))
            done += cur
            if progress_callback:
                progress_callback(done, n_samples)
        return torch.cat(chunks, dim=0)


class PretrainedDiffusionSampler(DiffusionSampler):
    """Sampler around a small random-init demo network plus the heuristic anneal the routes use
    when no checkpoint is given (samplers/diffusion.py:712-772)."""

    def __init__(self, lattice_size: int = 32, num_steps: int = 100, device: str = "cpu"):
        from ..models.score_network import ScoreNetwork

        super().__init__(score_network=ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2),
                         num_steps=num_steps, device=device)
        self.lattice_size = lattice_size

    @torch.no_grad()
    def sample_heuristic(self, batch_size: int = 1, temperature: float = 2.27, seed: Optional[int] = None
                         ) -> torch.Tensor:
        """x <- (1 - sqrt(t)) * sign(x) + sqrt(t) * 0.5 * z, t = 1 .. 1/N, then sign
        (samplers/diffusion.py:742-772): elementwise on the GPU, Gaussians from the engine's stream."""
        seed = self._fresh_seed() if seed is None else seed
        shape = (batch_size, 1, self.lattice_size, self.lattice_size)
        x = self._gaussian(shape, seed, abi.STEP_X0)
        for i in range(self.num_steps, 0, -1):
            r = (i / self.num_steps) ** 0.5
            self._lincomb([(1 - r, x), (0.5 * r, self._gaussian(shape, seed, i))], out=x, sign_first=True)
        return self._out(torch.sign(x))
