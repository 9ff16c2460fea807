"""Device plumbing shared by the reference-facing modules: torch is used for device memory and
streams only.  Every compute entry point calls require_cuda() first, so a box without a B200
fails loudly instead of computing anything on the host."""
import torch

from . import abi


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


__all__ = ["abi", "require_cuda", "current_stream", "to_device", "like_input", "ptr"]
