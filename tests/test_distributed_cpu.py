"""Host-side multi-GPU logic on the CPU: world-size-2 gloo all-reduce of the observable block and
the sharding helpers (the N>1 path of bench.py / temperature_sweep.py without GPUs)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "diffusion-boltzmann-sampler_b200")]
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from dbs_b200 import distributed as d
    import torch.distributed as dist
    r, ws, _ = d.init_from_env(backend="gloo")
    assert (r, ws) == (rank, world) and d.world() == (rank, world)
    # each rank accumulates its shard of 10 "chains"; the all-reduced block must equal the full sum
    e = torch.arange(10, dtype=torch.float64) * 0.1 - 1.0
    m = torch.linspace(-1, 1, 10, dtype=torch.float64)
    lo, hi = d.shard_range(10, rank, world)
    acc = torch.zeros(5, dtype=torch.float64)
    for k in range(lo, hi):
        acc += torch.stack([torch.tensor(1.0, dtype=torch.float64), e[k], m[k].abs(), m[k] ** 2, m[k] ** 4])
    d.allreduce_block(acc)
    q.put((rank, acc.numpy()))
    dist.destroy_process_group()


def test_gloo_world2_observable_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    e = np.arange(10) * 0.1 - 1.0
    m = np.linspace(-1, 1, 10)
    want = np.array([10.0, e.sum(), np.abs(m).sum(), (m ** 2).sum(), (m ** 4).sum()])
    np.testing.assert_allclose(res[0], want, rtol=1e-12)
    np.testing.assert_allclose(res[1], want, rtol=1e-12)


def test_shard_helpers():
    from dbs_b200 import distributed as d
    for n, w in [(10, 3), (65536, 8), (7, 8), (1024, 1)]:
        spans = [d.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    temps = np.linspace(1.5, 3.5, 4)
    seen = []
    for r in range(2):
        ids, t, ti = d.striped_temperatures(temps, 6, r, 2)
        assert len(ids) == 12 and set(ti.tolist()) == {0, 1, 2, 3}          # every rank sees every temperature
        assert np.allclose(t, temps[ids // 6])
        seen.extend(ids.tolist())
    assert sorted(seen) == list(range(24))
    out = d.finalize_observables(torch.tensor([[4.0, -6.0, 2.0, 1.5, 0.8]], dtype=torch.float64))
    assert np.isclose(out["energy_per_spin"][0], -1.5) and np.isclose(out["abs_magnetization"][0], 0.5)
    assert np.isclose(out["binder"][0], 1 - 0.2 / (3 * 0.375 ** 2))
