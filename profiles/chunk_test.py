"""Whole-step throughput vs DBS_L2_CHUNK_MB, eager and with the CUDA-graph forward."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-boltzmann-sampler_b200")]
import torch
import bench
B, L, NS = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
for graph in (False, True):
    smp = bench.make_sampler(L, NS, "bf16", "cuda")
    smp.use_graph = graph
    shape = (B, 1, L, L)
    for _ in range(2):
        smp.sample(shape, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(2):
        smp.sample(shape, seed=2 + r)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"chunk_mb={os.environ.get('DBS_L2_CHUNK_MB','24')} graph={graph} B={B} L={L}: {ms / NS:.3f} ms/forward-step, {B * NS / ms * 1e3 / NS:.0f} samples/s-equivalent at {NS} steps => {B / (ms / NS * 100) * 1e3:.0f} samples/s @100 steps")
