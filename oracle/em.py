"""Euler–Maruyama reverse-SDE step and VP-SDE noise level restated in numpy/float32.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, citing /root/reference/backend/ml:
  DiffusionProcess.noise_level      models/diffusion.py:68-97 (default linear schedule)
  DiffusionSampler.sample step      samplers/diffusion.py:219-243
"""
import numpy as np

from .philox import philox4x32_10, STREAM_NOISE

F = np.float32


def noise_level(t, beta_min=0.1, beta_max=20.0):
    """(alpha_t, sigma_t) as fp32 numpy arrays for fp32 t — models/diffusion.py:80-97 op by op with
    torch CPU fp32 ops (the reference's own arithmetic library, so exp/sqrt round identically):
    clamp t; integral = bmin*t + 0.5*(bmax-bmin)*t**2; alpha = exp(-0.5*integral);
    sigma = sqrt(clamp(1-alpha**2, 1e-8)); sigma = clamp(sigma, 1e-4)."""
    import torch

    tt = torch.clamp(torch.as_tensor(np.asarray(t, dtype=F)), min=0.0, max=1.0)
    integral = beta_min * tt + 0.5 * (beta_max - beta_min) * tt ** 2
    alpha = torch.exp(-0.5 * integral)
    sigma = torch.sqrt(torch.clamp(1 - alpha ** 2, min=1e-8))
    sigma = torch.clamp(sigma, min=1e-4)
    return alpha.numpy(), sigma.numpy()


def step_coefficients(sigma, num_steps):
    """Python-float sigma (sigma_t[0].item()) -> (c1, dt, c2) as the fp32 scalars that multiply the
    tensors in samplers/diffusion.py:234,239 (Python floats are rounded to fp32 at the multiply)."""
    dt = 1.0 / num_steps
    return F(sigma ** 2), F(dt), F(sigma * (dt ** 0.5))


def em_step(x, score, z, c1, dt, c2, temperature, add_noise):
    """x + (c1*score)*dt [+ (c2*z)*temperature] in fp32, the reference's evaluation order
    (samplers/diffusion.py:234-243; SURVEY.md Appendix C probe 2)."""
    x = np.asarray(x, dtype=F)
    out = x + (F(c1) * np.asarray(score, dtype=F)) * F(dt)
    if add_noise:
        out = out + (F(c2) * np.asarray(z, dtype=F)) * F(temperature)
    return out.astype(F)


def gaussian(seed, sample_ids, elems, step):
    """fp32 [n][elems] standard normals of the engine's in-kernel stream (csrc/em.cu):
    Philox ctr = {sample, e//4, step, STREAM_NOISE}; Box–Muller on (x,y) -> e%4 in {0,1},
    (z,w) -> {2,3}; u = ((r>>8)+0.5)*2^-24."""
    sample_ids = np.asarray(sample_ids, dtype=np.uint64)
    quads = (elems + 3) // 4
    q = np.arange(quads, dtype=np.uint64)[None, :]
    o = philox4x32_10(sample_ids[:, None], q, np.uint64(step & 0xFFFFFFFF), STREAM_NOISE, seed)

    def bm(a, b):
        u1 = ((a >> np.uint32(8)).astype(F) + F(0.5)) * F(2.0 ** -24)
        u2 = ((b >> np.uint32(8)).astype(F) + F(0.5)) * F(2.0 ** -24)
        rad = np.sqrt(F(-2.0) * np.log(u1)).astype(F)
        ang = (F(2.0) * u2).astype(np.float64) * np.pi
        return (rad * np.cos(ang).astype(F)).astype(F), (rad * np.sin(ang).astype(F)).astype(F)

    z0, z1 = bm(o[0], o[1])
    z2, z3 = bm(o[2], o[3])
    z = np.stack([z0, z1, z2, z3], axis=-1).reshape(len(sample_ids), quads * 4)
    return z[:, :elems]
