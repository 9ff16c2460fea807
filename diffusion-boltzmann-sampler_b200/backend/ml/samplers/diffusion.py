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

_CAPTURE_LOCK = runtime.CAPTURE_LOCK   # CUDA-graph captures of different samplers never overlap (see _StepProgram)


def _f32(v: float) -> float:
    """A Python float after rounding to fp32 — what `python_float * fp32_tensor` multiplies by."""
    return float(np.float32(v))


class _StepProgram:
    """One sampler step as a captured CUDA graph, with everything the graph's kernel nodes point at.

    The step's scalars (times, sigma^2 dt, noise scales, Philox step ids) are row k of a device table; the graph begins
    with dbs_step_table_next, which copies the row selected by a device counter into `cur` — where the U-Net reads
    its time and the update kernels their coefficients — and advances the counter.  A trajectory is then N replays of
    ONE graph: no per-step host-side parameter updates, no eager launches, no host synchronisation
    (the loops of samplers/diffusion.py:222-243, 425-449, 476-519).

    Row layouts (fp32, STRIDE columns; ids are int32 bit patterns):
      em    [t, c1, dt, c2, temperature, add_noise, step id]
      pc    em columns, then per corrector step c: [temperature, active, step id] at 7 + 3 c
      heun  [t0, t1, 1, a, 1, b0, b1]
      rk4   [t0, t_half, t1, 1, a1, 1, a2, 1, a3, 1, b1, b2, b3, b4]
    """
    STRIDE = 20
    MAX_CORRECTORS = 4

    def __init__(self, sampler: "DiffusionSampler", method: str, shape, n_corr: int, snr: float, plan):
        dev = runtime.require_cuda()
        self.method, self.shape, self.plan = method, shape, plan
        self.model, self.precision = sampler.model, sampler.precision
        n = shape[0]
        self.n, self.elems = n, int(np.prod(shape[1:]))
        f32 = dict(dtype=torch.float32, device=dev)
        self.x = torch.empty(shape, **f32)
        n_aux = {"em": 1, "pc": 1, "heun": 3, "rk4": 5}[method]        # scores (+ the intermediate state)
        self.aux = [torch.empty(shape, **f32) for _ in range(n_aux)]
        self.table = torch.zeros((DiffusionSampler._TABLE_ROWS, self.STRIDE), **f32)
        self.cur = torch.zeros(self.STRIDE, **f32)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.seed_off = torch.zeros(2, dtype=torch.int64, device=dev)
        self.norms = torch.empty(max(n, 1), **f32)
        self.cc = torch.empty(2, **f32)
        self.n_corr, self.snr = n_corr, float(snr)
        self.ws = self.model.engine_workspace(shape, self.precision)
        self._stage = torch.empty((DiffusionSampler._TABLE_ROWS, self.STRIDE), dtype=torch.float32).pin_memory()
        self._seed_stage = torch.zeros(2, dtype=torch.int64).pin_memory()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: the capture records kernel launches only (workspace and buffers exist already), so CUDA calls of
        # OTHER threads — an MCMC job on the server's thread pool allocating its lattices (routes/training.py:191) — must
        # neither fail nor invalidate it, as they would in "global" mode.  torch.cuda.graph captures on ONE process-wide
        # side stream by default: two samplers capturing from two threads would begin a capture on a stream that is
        # already capturing, so captures are serialised and each gets a stream of its own.  The warm-up launch and its
        # synchronisation are inside the same lock (a device-wide synchronise issued by one thread while another thread
        # was in mid-capture failed intermittently with "operation not permitted when stream is capturing"), and wait on
        # this thread's stream only.  The cyclic garbage collector may run finalisers (cudaFree of a dropped plan, ...) on
        # whichever thread happens to allocate — also on this one in mid-capture, where they are not permitted: it is
        # paused for the capture (plan finalisers additionally defer themselves, runtime.release_plan).
        with _CAPTURE_LOCK:
            self._issue()                  # warm-up: one-time function attributes, tensor-map encoder lookup
            torch.cuda.current_stream().synchronize()
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                before = abi.launch_count()
                with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(), capture_error_mode="thread_local"):
                    self._issue()
                self.n_kernels = abi.launch_count() - before
            finally:
                if gc_was_on:
                    gc.enable()

    # ---- the step, as a sequence of engine launches on the current stream (captured once)
    def _coef(self, off: int) -> int:
        return self.cur.data_ptr() + 4 * off

    def _forward(self, x, t_off: int, out):
        self.model.engine_forward(x, self.cur[t_off:t_off + 1], precision=self.precision, out=out, workspace=self.ws)

    def _lincomb(self, out, terms, coef_off: int):
        import ctypes
        ptrs = (ctypes.c_void_p * len(terms))(*[t.data_ptr() for t in terms])
        abi.check(abi.lib.dbs_lincomb_dev(out.data_ptr(), out.numel(), len(terms), ptrs, self._coef(coef_off), 0,
                                          runtime.current_stream()), "lincomb_dev")

    def _issue(self):
        st = runtime.current_stream()
        x, a = self.x, self.aux
        abi.check(abi.lib.dbs_step_table_next(self.cur.data_ptr(), self.table.data_ptr(), self.counter.data_ptr(),
                                              self.STRIDE, st), "step_table_next")
        if self.method in ("em", "pc"):
            self._forward(x, 0, a[0])
            abi.check(abi.lib.dbs_em_step_dev(x.data_ptr(), a[0].data_ptr(), self.n, self.elems, self._coef(1),
                                              self.seed_off.data_ptr(), st), "em_step_dev")
            for c in range(self.n_corr if self.method == "pc" else 0):
                self._forward(x, 0, a[0])
                abi.check(abi.lib.dbs_corrector_coef(a[0].data_ptr(), self.n, self.elems, self.snr, self.norms.data_ptr(),
                                                     self.cc.data_ptr(), st), "corrector_coef")
                abi.check(abi.lib.dbs_corrector_update(x.data_ptr(), a[0].data_ptr(), 0, self.n, self.elems,
                                                       self.cc.data_ptr(), self._coef(7 + 3 * c), self.seed_off.data_ptr(),
                                                       st), "corrector_update")
        elif self.method == "heun":
            s0, s1, x1 = a
            self._forward(x, 0, s0)
            self._lincomb(x1, [x, s0], 2)
            self._forward(x1, 1, s1)
            self._lincomb(x, [x, s0, s1], 4)
        else:   # rk4
            s1, s2, s3, s4, xt = a
            self._forward(x, 0, s1)
            self._lincomb(xt, [x, s1], 3)
            self._forward(xt, 1, s2)
            self._lincomb(xt, [x, s2], 5)
            self._forward(xt, 1, s3)
            self._lincomb(xt, [x, s3], 7)
            self._forward(xt, 2, s4)
            self._lincomb(x, [x, s1, s2, s3, s4], 9)

    # ---- a trajectory: upload the table, reset the counter, replay once per row
    def begin(self, rows: np.ndarray, seed: int, sample_offset: int, x0):
        """Initial state (Gaussian start or the caller's x0), the Philox key, and the first table segment, on the current
        stream."""
        x = self.x
        if x0 is not None:
            x.copy_(runtime.to_device(x0).reshape(self.shape), non_blocking=True)
        else:
            abi.check(abi.lib.dbs_gaussian_fill(x.data_ptr(), self.n, self.elems, seed, sample_offset, abi.STEP_X0,
                                                runtime.current_stream()), "gaussian_fill")
        s64 = seed & 0xFFFFFFFFFFFFFFFF        # the ABI's uint64 seed as the int64 bit pattern the device table holds
        self._seed_stage[0] = s64 - (1 << 64) if s64 >= (1 << 63) else s64
        self._seed_stage[1] = sample_offset
        self.seed_off.copy_(self._seed_stage, non_blocking=True)
        self._rows, self._next = rows, 0

    def step(self):
        """Replay the captured step for the next row (uploading the next table segment when the previous one is used up)."""
        k, cap = self._next, self.table.shape[0]
        if k % cap == 0:
            m = min(cap, self._rows.shape[0] - k)
            torch.cuda.current_stream().synchronize()     # the staging rows of the previous trajectory / segment have been read
            self._stage[:m].copy_(torch.from_numpy(self._rows[k:k + m]))
            self.table[:m].copy_(self._stage[:m], non_blocking=True)
            self.counter.zero_()
        self.graph.replay()
        self.sampler_launches += self.n_kernels
        self._next = k + 1

    def run(self, rows: np.ndarray, seed: int, sample_offset: int, x0, first_index: int, index_of):
        self.begin(rows, seed, sample_offset, x0)
        yield first_index, self.x
        for k in range(rows.shape[0]):
            self.step()
            yield index_of(k), self.x

    sampler_launches = 0     # kernels executed through replays of this program (launch accounting)


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
        # One U-Net forward is ~35 kernel launches: the whole step (forward + fused update, step scalars from a
        # device table) is captured once per shape into a CUDA graph and replayed N times (_StepProgram).
        # True / False / "auto" (= True; False = eager launches, kept for debugging and per-kernel profiling).
        self.use_graph = "auto"
        self._graphs: Dict[Any, Any] = {}
        self._launches_dropped = 0       # replayed kernels of programs that have left the cache
        # True: large bf16 batches run as two half-batch programs on two streams (see _sample_split).  Off by default:
        # measured at the config-4 shape (profiles/README.md, round 2) the halves' kernels co-reside but a strip conv slows
        # down under a co-running bandwidth-bound kernel by about what the overlap saves (12.8 vs 13.1 ms per step).
        self.split_streams = False
        self._streams = None

    @property
    def graph_kernel_launches(self) -> int:
        """Kernels executed through graph replays (for launch accounting: abi.launch_count() sees eager launches only)."""
        return self._launches_dropped + sum(p.sampler_launches for p in self._graphs.values())

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

    # ------------------------------------------------------------------ captured step programs
    _TABLE_ROWS = 4096      # rows of a program's device step table (longer trajectories run in segments)
    _MAX_PROGRAMS = 4       # programs (each owns a U-Net workspace) kept per sampler, least recently used dropped

    def _graph_wanted(self) -> bool:
        want = self.use_graph
        return True if want == "auto" else bool(want)

    def _program(self, method: str, shape, n_corr: int = 0, snr: float = 0.0, slot: int = 0) -> "_StepProgram":
        """The captured step of `method` for this shape / precision / set of weights (built on first use).

        A program owns everything its graph's kernel nodes point at — state and score buffers, the step table, the
        U-Net workspace — and holds the model's plan object, so nothing a cached graph references can be freed or
        re-used while the program is alive (a workspace shared with the model was regrown under cached graphs)."""
        runtime.require_cuda()
        plan = None
        if hasattr(self.model, "engine_plan"):
            self.model.engine_plan()
            plan = self.model._plan
        key = (method, tuple(shape), self.precision, n_corr, float(snr), plan.generation if plan is not None else None,
               torch.cuda.current_device(), slot)
        prog = self._graphs.pop(key, None)
        if prog is None:
            # programs of superseded plans first (their weights are gone), then the least recently used
            gen = key[5]
            for k in [k for k in self._graphs if k[5] != gen]:
                self._launches_dropped += self._graphs.pop(k).sampler_launches
            while len(self._graphs) >= self._MAX_PROGRAMS:
                self._launches_dropped += self._graphs.pop(next(iter(self._graphs))).sampler_launches
            prog = _StepProgram(self, method, tuple(shape), n_corr, snr, plan)
        self._graphs[key] = prog       # (re-)inserted last = most recently used
        return prog

    def _reverse_sde_eager(self, shape, temperature, seed, x0=None, noises=None, sample_offset=0):
        """Euler–Maruyama trajectory with one eager launch sequence per step: the parity port (injected noise tensors)
        and the un-captured reference for the step programs.  Yields like _reverse_sde."""
        dev = runtime.require_cuda()
        x = torch.empty(shape, dtype=torch.float32, device=dev)
        score = torch.empty_like(x)
        if x0 is not None:
            x.copy_(runtime.to_device(x0).reshape(shape))
        else:
            n = shape[0]
            abi.check(abi.lib.dbs_gaussian_fill(x.data_ptr(), n, x.numel() // max(n, 1), seed, sample_offset, abi.STEP_X0,
                                                runtime.current_stream()), "gaussian_fill")
        sig = self.diffusion.sigma_table(self.num_steps)
        yield self.num_steps + 1, x
        for k, i in enumerate(range(self.num_steps, 0, -1)):
            self._score(x, i / self.num_steps, out=score)
            nz = None
            if noises is not None and i > 1:
                nz = runtime.to_device(noises[k])
            self._em(x, score, sig[k], temperature, i > 1, seed, i, noise=nz, sample_offset=sample_offset)
            yield i, x

    def _em_rows(self, temperature: float, stochastic: bool = True) -> np.ndarray:
        """Step table of the Euler–Maruyama loop: row k = step i = N - k."""
        N = self.num_steps
        dt = 1.0 / N
        sig = self.diffusion.sigma_table(N)
        rows = np.zeros((N, _StepProgram.STRIDE), dtype=np.float32)
        steps = np.zeros(N, dtype=np.int32)
        for k, i in enumerate(range(N, 0, -1)):
            rows[k, 0] = i / N
            rows[k, 1:5] = (_f32(sig[k] ** 2), _f32(dt), _f32(sig[k] * dt ** 0.5), _f32(temperature))
            rows[k, 5] = 1.0 if (stochastic and i > 1) else 0.0
            steps[k] = i
        rows[:, 6] = steps.view(np.float32)
        return rows

    def _reverse_sde(self, shape, temperature, seed, x0=None, noises=None, sample_offset=0):
        """Generator over the Euler–Maruyama trajectory: first (num_steps + 1, x_1) — the Gaussian
        start — then (i, x) after step i = num_steps .. 1.  x is a persistent buffer updated in place
        (callers copy what they keep).  One CUDA-graph replay per step: the forward, the fused update and the step's
        scalars (device step table) are all inside the captured step."""
        runtime.require_cuda()
        shape = tuple(shape)
        if noises is not None or not self._graph_wanted() or shape[0] == 0 or not hasattr(self.model, "engine_forward"):
            yield from self._reverse_sde_eager(shape, temperature, seed, x0, noises, sample_offset)
            return
        prog = self._program("em", shape)
        yield from prog.run(self._em_rows(temperature), seed, sample_offset, x0, first_index=self.num_steps + 1,
                            index_of=lambda k: self.num_steps - k)

    # ------------------------------------------------------------------ samplers
    # Two half-batch programs on two streams.  A conv CTA of the tcgen05 kernels owns its SM's shared memory but only
    # half of its registers and a fraction of its issue slots, while GroupNorm apply / Euler-Maruyama / the init conv are
    # HBM-bound kernels that need no shared memory: with the batch split in two independent trajectories, the memory-bound
    # kernels of one half run UNDER the convs of the other half instead of between them (samples are independent, so
    # the halves never synchronise until the end).  Threshold: each half must still fill the GPU with conv tiles.
    SPLIT_MIN_PIXELS = 2 * 148 * 128 * 4      # per half: >= 4 tiles of 128 pixels per SM at full resolution

    def _split_sizes(self, shape):
        if not self.split_streams or not self._graph_wanted() or len(shape) != 4:
            return None
        b = shape[0]
        pix = shape[2] * shape[3]
        if self.precision != "bf16" or b < 2 or (b // 2) * pix < self.SPLIT_MIN_PIXELS:
            return None
        return (b + 1) // 2, b // 2

    def _sample_split(self, shape, sizes, temperature, seed, x0, sample_offset):
        main = torch.cuda.current_stream()
        if self._streams is None:
            self._streams = (torch.cuda.Stream(), torch.cuda.Stream())
        rows = self._em_rows(temperature)
        progs, off = [], 0
        for slot, (b, st) in enumerate(zip(sizes, self._streams)):
            prog = self._program("em", (b,) + tuple(shape[1:]), slot=slot)
            st.wait_stream(main)
            with torch.cuda.stream(st):
                prog.begin(rows, seed, sample_offset + off, None if x0 is None else x0[off:off + b])
            progs.append(prog)
            off += b
        for _ in range(self.num_steps):
            for prog, st in zip(progs, self._streams):
                with torch.cuda.stream(st):
                    prog.step()
        out = torch.empty(shape, dtype=torch.float32, device=progs[0].x.device)
        off = 0
        for prog, st in zip(progs, self._streams):
            with torch.cuda.stream(st):
                out[off:off + prog.n].copy_(prog.x, non_blocking=True)
            off += prog.n
            main.wait_stream(st)
        out.record_stream(self._streams[0])
        out.record_stream(self._streams[1])
        return out

    @torch.no_grad()
    def sample(self, shape: Tuple[int, ...], temperature: float = 1.0, seed: Optional[int] = None, x0=None,
               noises=None, sample_offset: int = 0) -> torch.Tensor:
        """Euler–Maruyama reverse SDE (samplers/diffusion.py:197-245)."""
        seed = self._fresh_seed() if seed is None else seed
        shape = tuple(shape)
        sizes = self._split_sizes(shape) if (noises is None and hasattr(self.model, "engine_forward")) else None
        if sizes is not None:
            out = self._sample_split(shape, sizes, temperature, seed, x0, sample_offset)
            return out if str(self.device).startswith("cuda") else out.cpu()
        x = None
        for _, x in self._reverse_sde(shape, temperature, seed, x0, noises, sample_offset):
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
    def stream_spin_frames(self, shape: Tuple[int, ...], temperature: float = 1.0, yield_every: int = 1,
                           seed: Optional[int] = None, ring_slots: int = 8, ising_model=None
                           ) -> Generator[Dict[str, Any], None, None]:
        """Frame adaptor for streaming consumers (api/websocket.py:163-211 discretises every yielded state with
        torch.sign and asks for its energy and magnetisation).  Same frame schedule as sample_with_trajectory, but a frame
        is produced ON THE DEVICE — sign bits packed 1 bit per spin, E/N and M from the packed words — and copied
        asynchronously into a pinned host ring; the sampling loop never waits for a frame.  A frame is handed out once
        its copy has completed (or when the ring is full), so frames arrive in order, up to `ring_slots` emits late.

        Yields {"bits": uint32 host array (B, L, W) [a ring slot: consume before `ring_slots` more frames],
                "energy_per_spin": (B,), "magnetization": (B,), "t": float, "step": int, "frame": int}."""
        from ..systems.ising import IsingModel
        shape = tuple(shape)
        B, L = shape[0], shape[-1]
        if len(shape) != 4 or shape[1] != 1 or shape[2] != shape[3]:
            raise ValueError("stream_spin_frames: shape must be (batch, 1, L, L)")
        model = ising_model if ising_model is not None else IsingModel(L)
        dev = runtime.require_cuda()
        wpc = abi.lib.dbs_ising_words_per_chain(L)
        if wpc < 0:
            abi.check(abi.ERR_INVALID, "ising_words_per_chain")
        slots = max(2, int(ring_slots))
        bits_d = torch.empty((slots, B, wpc), dtype=torch.int32, device=dev)
        obs_d = torch.empty((slots, 2, B), dtype=torch.float32, device=dev)
        bits_h = torch.empty((slots, B, wpc), dtype=torch.int32).pin_memory()
        obs_h = torch.empty((slots, 2, B), dtype=torch.float32).pin_memory()
        events = [torch.cuda.Event() for _ in range(slots)]
        pending = []                  # (slot, t, step, frame index) in emission order
        seed = self._fresh_seed() if seed is None else seed
        st = runtime.current_stream
        n_frames = 0

        def emit(x, t, step):
            nonlocal n_frames
            slot = n_frames % slots
            abi.check(abi.lib.dbs_ising_pack(x.data_ptr(), bits_d[slot].data_ptr(), L, B, st()), "ising_pack")
            abi.check(abi.lib.dbs_ising_observables(bits_d[slot].data_ptr(), L, B, float(model.J), float(model.h),
                                                    obs_d[slot, 0].data_ptr(), obs_d[slot, 1].data_ptr(), st()),
                      "ising_observables")
            bits_h[slot].copy_(bits_d[slot], non_blocking=True)
            obs_h[slot].copy_(obs_d[slot], non_blocking=True)
            events[slot].record()
            pending.append((slot, t, step, n_frames))
            n_frames += 1

        def take(block):
            slot, t, step, idx = pending[0]
            if not events[slot].query():
                if not block:
                    return None
                events[slot].synchronize()
            pending.pop(0)
            return {"bits": bits_h[slot].numpy().view(np.uint32).reshape(B, L, -1),
                    "energy_per_spin": (obs_h[slot, 0] / float(L * L)).numpy(), "magnetization": obs_h[slot, 1].numpy(),
                    "t": t, "step": step, "frame": idx}

        for i, x in self._reverse_sde(shape, temperature, seed):
            due = i > self.num_steps or i % yield_every == 0 or i == 1
            if due:
                # the slot about to be overwritten must have been handed out
                while len(pending) >= slots:
                    yield take(True)
                emit(x, 1.0 if i > self.num_steps else (i - 1) / self.num_steps, min(i, self.num_steps + 1) - 1)
            while pending:
                fr = take(False)
                if fr is None:
                    break
                yield fr
        while pending:
            yield take(True)

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

    def _pc_rows(self, temperature: float, n_corr: int) -> np.ndarray:
        """Step table of the predictor-corrector loop: the Euler–Maruyama columns, then per corrector step
        {temperature, active, Philox step id}; correctors are inactive at the final predictor step (i = 1)."""
        N = self.num_steps
        rows = self._em_rows(temperature)
        ids = np.zeros(N, dtype=np.int32)
        for c in range(n_corr):
            for k, i in enumerate(range(N, 0, -1)):
                ids[k] = N + 1 + (N - i) * n_corr + c
            rows[:, 7 + 3 * c] = _f32(temperature)
            rows[:, 8 + 3 * c] = [1.0 if i > 1 else 0.0 for i in range(N, 0, -1)]
            rows[:, 9 + 3 * c] = ids.view(np.float32)
        return rows

    @torch.no_grad()
    def sample_predictor_corrector(self, shape: Tuple[int, ...], temperature: float = 1.0, n_corrector_steps: int = 1,
                                   corrector_snr: float = 0.16, seed: Optional[int] = None, x0=None, noises=None,
                                   corrector_noises=None) -> torch.Tensor:
        """EM predictor + Langevin corrector (samplers/diffusion.py:399-451).  The corrector's step size needs a
        batch-wide norm: a device reduction (dbs_corrector_coef) feeding the fused update, no host round trip, and the
        whole predictor + corrector step is one captured graph.  `x0` / `noises` [N] / `corrector_noises`
        [N][n_corrector_steps] are the parity ports (the reference's torch.randn draws in call order)."""
        seed = self._fresh_seed() if seed is None else seed
        shape = tuple(shape)
        N, nc = self.num_steps, int(n_corrector_steps)
        B = shape[0]
        if B == 0:
            return self._result(torch.empty(shape, dtype=torch.float32, device=runtime.require_cuda()))
        if nc > _StepProgram.MAX_CORRECTORS:
            raise NotImplementedError(f"predictor-corrector: at most {_StepProgram.MAX_CORRECTORS} corrector steps are built")
        if noises is None and corrector_noises is None and self._graph_wanted() and hasattr(self.model, "engine_forward"):
            prog = self._program("pc", shape, nc, corrector_snr)
            x = None
            for _, x in prog.run(self._pc_rows(temperature, nc), seed, 0, x0, first_index=N + 1, index_of=lambda k: N - k):
                pass
            return self._result(x)
        # eager path (parity port / use_graph = False): the same kernels, launched step by step
        elems = int(np.prod(shape[1:]))
        dev = runtime.require_cuda()
        norms = torch.empty(B, dtype=torch.float32, device=dev)
        cc = torch.empty(2, dtype=torch.float32, device=dev)
        seed_off = torch.tensor([seed, 0], dtype=torch.int64, device=dev)
        rows = torch.from_numpy(self._pc_rows(temperature, nc)).to(dev)
        score = None
        for i, x in self._reverse_sde_eager(shape, temperature, seed, x0, noises):
            if 1 < i <= N:
                k = N - i
                for c in range(nc):
                    score = self._score(x, i / N, out=score)
                    abi.check(abi.lib.dbs_corrector_coef(score.data_ptr(), B, elems, float(corrector_snr), norms.data_ptr(),
                                                         cc.data_ptr(), runtime.current_stream()), "corrector_coef")
                    nz = None if corrector_noises is None else runtime.to_device(corrector_noises[k][c])
                    abi.check(abi.lib.dbs_corrector_update(x.data_ptr(), score.data_ptr(), runtime.ptr(nz), B, elems,
                                                           cc.data_ptr(), rows[k, 7 + 3 * c:].data_ptr(),
                                                           seed_off.data_ptr(), runtime.current_stream()), "corrector_update")
        return self._result(x)

    def _ode_rows(self, method: str) -> np.ndarray:
        """Step tables of the deterministic integrators; sigma^2 products in fp32 as `python_float * fp32_tensor` does."""
        N = self.num_steps
        dt = 1.0 / N
        rows = np.zeros((N, _StepProgram.STRIDE), dtype=np.float32)
        for k, i in enumerate(range(N, 0, -1)):
            t_val = i / N
            if method == "heun":
                c0, c1 = self._sigma(t_val) ** 2, self._sigma((i - 1) / N) ** 2
                rows[k, :7] = (t_val, (i - 1) / N, 1.0, c0 * dt, 1.0, 0.5 * c0 * dt, 0.5 * c1 * dt)
            else:
                ca, cm, cb = self._sigma(t_val) ** 2, self._sigma(t_val - 0.5 * dt) ** 2, self._sigma(t_val - dt) ** 2
                rows[k, :14] = (t_val, t_val - 0.5 * dt, t_val - dt, 1.0, 0.5 * dt * ca, 1.0, 0.5 * dt * cm, 1.0, dt * cm,
                                1.0, dt / 6 * ca, dt / 3 * cm, dt / 3 * cm, dt / 6 * cb)
        return rows

    @torch.no_grad()
    def sample_ode(self, shape: Tuple[int, ...], method: str = "euler", seed: Optional[int] = None, x0=None
                   ) -> torch.Tensor:
        """Probability-flow style deterministic integrators: euler / heun / rk4
        (samplers/diffusion.py:453-521).  Each step — two (Heun) or four (RK4) forwards and the fused state
        combinations between them — is one captured graph replayed N times."""
        if method not in ("euler", "heun", "rk4"):
            raise ValueError(f"Unknown ODE method: {method}")
        seed = self._fresh_seed() if seed is None else seed
        shape = tuple(shape)
        N = self.num_steps
        dt = 1.0 / N
        if shape[0] > 0 and self._graph_wanted() and hasattr(self.model, "engine_forward"):
            if method == "euler":
                prog, rows = self._program("em", shape), self._em_rows(0.0, stochastic=False)
            else:
                prog, rows = self._program(method, shape), self._ode_rows(method)
            x = None
            for _, x in prog.run(rows, seed, 0, x0, first_index=N + 1, index_of=lambda k: N - k):
                pass
            return self._result(x)
        x = self._gaussian(shape, seed, abi.STEP_X0) if x0 is None else runtime.to_device(x0).reshape(shape).clone()
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
