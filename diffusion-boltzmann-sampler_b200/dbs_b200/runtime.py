"""Device plumbing shared by the reference-facing modules: torch is used for device memory and
streams only.  Every compute entry point calls require_cuda() first, so a box without a B200
fails loudly instead of computing anything on the host."""
import threading

import torch

from . import abi

# CUDA-graph captures (samplers) and the device-wide synchronisations of plan builds never overlap: a thread-local capture
# tolerates other threads' launches and allocations, but a cudaDeviceSynchronize issued elsewhere in mid-capture was seen
# to fail one of the two intermittently ("operation not permitted when stream is capturing").  Re-entrant.
CAPTURE_LOCK = threading.RLock()

_parked = []                       # (destroy_fn, handle) of plans whose release had to wait for a stream capture to end
_parked_lock = threading.Lock()


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "dbs_b200: no CUDA device visible. This engine computes only on sm_100a (B200); "
            "there is no CPU fallback."
        )
    return torch.device("cuda", torch.cuda.current_device())


def current_stream():
    return torch.cuda.current_stream().cuda_stream


def to_device(t, dtype=torch.float32):
    """Host or device tensor -> contiguous tensor of `dtype` on the current CUDA device."""
    dev = require_cuda()
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.to(device=dev, dtype=dtype, non_blocking=False).contiguous()


def like_input(result, ref):
    """Return `result` on the device the caller's tensor lived on (CPU in -> CPU out, as the
    reference's consumers call .tolist()/.numpy() on it directly)."""
    if isinstance(ref, torch.Tensor) and ref.is_cuda:
        return result
    return result.cpu()


def ptr(t):
    return 0 if t is None else t.data_ptr()


def _capturing():
    try:
        return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    except Exception:              # interpreter shutdown
        return False


def flush_released():
    """Free the plans parked by release_plan (no-op while this thread is capturing)."""
    if _capturing():
        return
    with _parked_lock:
        todo, _parked[:] = list(_parked), []
    for destroy_fn, handle in todo:
        destroy_fn(handle)


def release_plan(destroy_fn, handle):
    """Free an engine plan (cudaFree inside).  Python may run a plan's finaliser on ANY thread at ANY time — also on a
    thread that is in the middle of a CUDA-graph capture, where cudaFree fails with "operation not permitted when
    stream is capturing" and invalidates the capture.  Such handles are parked and freed by the next release or
    plan build outside a capture."""
    if _capturing():
        with _parked_lock:
            _parked.append((destroy_fn, handle))
        return
    flush_released()
    destroy_fn(handle)


__all__ = ["abi", "require_cuda", "current_stream", "to_device", "like_input", "ptr", "release_plan", "flush_released",
           "CAPTURE_LOCK"]
