"""Per-kernel-class device time of the config-3 temperature-sweep driver (bounded), from the in-library profiler."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
from dbs_b200 import abi
from dbs_b200.temperature_sweep import run_temperature_sweep
kw = dict(lattice_size=64, chains_per_temperature=1024, burn_in=100, n_measure=100, measure_every=10, seed=1234)
run_temperature_sweep(**dict(kw, burn_in=10, n_measure=10))
torch.cuda.synchronize()
abi.profile_read(True); abi.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run_temperature_sweep(**kw); e1.record(); torch.cuda.synchronize()
print("total ms", e0.elapsed_time(e1))
print({k: (round(v[0], 2), v[1]) for k, v in abi.profile_read().items() if v[1]})
