"""One measurement (E, M, block sums, C(r)) of the config-3 ensemble — 65 536 chains 64x64 — with the XOR / POPC
displacement kernel and with the spectral kernel: device time per launch and equality of the integer sums.
usage: python profiles/r2_corr_methods.py [n_chains]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import numpy as np
import torch
from backend.ml.samplers.mcmc import MetropolisHastings
from backend.ml.systems.ising import IsingModel
from dbs_b200 import abi, runtime

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m = IsingModel(64)
temps = np.repeat(np.linspace(1.5, 3.5, 64), (n + 63) // 64)[:n]
group = torch.tensor(np.repeat(np.arange(64), (n + 63) // 64)[:n].astype(np.int32)).cuda()
bits = m.random_packed(n, seed=1)
MetropolisHastings(m, temps, seed=1).sweeps_packed(bits, 50, sweep_offset=0)
res = {}
for method, name in ((1, "xor/popc"), (2, "spectral")):
    abi.lib.dbs_ising_set_correlation_method(method)
    acc = torch.zeros((64, 32), dtype=torch.int64, device="cuda")
    block = torch.zeros((64, 38), dtype=torch.float64, device="cuda")
    def go():
        abi.check(abi.lib.dbs_ising_measure(bits.data_ptr(), 64, n, 1.0, 0.0, group.data_ptr(), block.data_ptr(), 38,
                                            acc.data_ptr(), runtime.current_stream()))
    go(); torch.cuda.synchronize()
    acc.zero_(); block.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        go()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:9s} {e0.elapsed_time(e1) / 5:.3f} ms per measurement of {n} chains")
    res[method] = (acc.cpu(), block.cpu())
abi.lib.dbs_ising_set_correlation_method(0)
print("C(r) sums identical:", bool(torch.equal(res[1][0], res[2][0])), " bin 0:", int(res[2][0][:, 0].abs().sum()),
      " block max rel diff:", float(((res[1][1] - res[2][1]).abs() / (res[1][1].abs() + 1e-300)).max()))
