"""dbs_b200 — ctypes binding and host-side runtime for libdbs_b200.so (sm_100a kernels).

This package is plumbing: it owns no arithmetic.  The reference-facing API lives in the sibling
`backend/ml/` tree (same module paths and signatures as the reference); those modules call the
C ABI declared in include/dbs_b200.h through `dbs_b200.abi`.
"""
from . import abi  # noqa: F401  (loads the shared library; raises if it is not built)
from .runtime import require_cuda, current_stream, to_device, like_input  # noqa: F401

__all__ = ["abi", "require_cuda", "current_stream", "to_device", "like_input"]
