"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed).

Chains / samples / configurations are independent, so there is no data-path collective: rank r
works on the global index range shard_range(n, r, world) and passes `lo` as chain_offset /
sample_offset so the Philox counters carry GLOBAL indices (any world size gives identical
results).  The one exchange is the all-reduce of the fp64 observable block at the end of a
measurement (NCCL over NVLink on GPUs; gloo in the CPU tests).  SURVEY.md §8(e).
"""
import os

import numpy as np
import torch


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK/WORLD_SIZE/LOCAL_RANK/MASTER_* (torchrun); no-op for
    a single process.  Returns (rank, world_size, local_rank)."""
    import torch.distributed as dist
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, ws, local


def shard_range(n, rank, world_size):
    """Contiguous, balanced [lo, hi) of n units for `rank`."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def striped_temperatures(temperatures, chains_per_temperature, rank, world_size):
    """C3 layout: every temperature's chains are striped across all ranks so that each GPU
    contributes to every temperature bin.  Returns (global_chain_ids, temperature_per_chain,
    temperature_index_per_chain) for this rank."""
    temps = np.asarray(temperatures, dtype=np.float64)
    nT = len(temps)
    ids = np.arange(nT * chains_per_temperature, dtype=np.int64)
    mine = ids[(ids % chains_per_temperature) % world_size == rank] if world_size > 1 else ids
    t_idx = mine // chains_per_temperature
    return mine, temps[t_idx], t_idx.astype(np.int32)


def allreduce_block(acc):
    """Sum an observable accumulator block over all ranks (in place); identity for one process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def finalize_observables(acc):
    """Per-group sums -> means and the Binder cumulant.  Six columns {count, sum E/N, sum (E/N)^2, sum |M|, sum M^2,
    sum M^4} (the temperature sweep's block) or the five without the energy square."""
    a = acc.detach().cpu().numpy()
    a = a.reshape(-1, a.shape[-1]) if a.ndim > 1 else a.reshape(-1, 5)
    if a.shape[1] == 5:
        a = np.concatenate([a[:, :2], np.full((a.shape[0], 1), np.nan), a[:, 2:]], axis=1)
    n = np.maximum(a[:, 0], 1.0)
    m2, m4 = a[:, 4] / n, a[:, 5] / n
    with np.errstate(divide="ignore", invalid="ignore"):
        binder = 1.0 - m4 / (3.0 * m2 * m2)
    return {"count": a[:, 0], "energy_per_spin": a[:, 1] / n, "e2": a[:, 2] / n, "abs_magnetization": a[:, 3] / n,
            "m2": m2, "m4": m4, "binder": binder}
