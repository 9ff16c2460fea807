"""Ising hot path restated on the CPU in numpy.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, citing /root/reference/backend/ml:
  IsingModel.local_energy_diff       systems/ising.py:68-91
  IsingModel.energy / magnetization  systems/ising.py:47-66, 93-106
  MetropolisHastings.__init__/step   samplers/mcmc.py:15-50   (acceptance rule)
  pair_correlation                   analysis/correlation.py:8-47

The chain itself is the engine's checkerboard order: the reference's own
step() is a random-site chain, but because same-colour sites do not interact, driving step()
over colour 0 then colour 1 with per-site uniforms reproduces a checkerboard sweep bit for bit
(make_golden.py does exactly that with the unmodified reference and stores the result).
"""
import math

import numpy as np

from .philox import philox4x32_10, STREAM_INIT, STREAM_SWEEP, STREAM_SWEEP3


# --------------------------------------------------------------------------- acceptance table
def delta_e_classes(J, h):
    """fp32 dE for the 16 site classes k = 8*s_bit + u (u = number of +1 neighbours, 0..4;
    u = 5..7 unused -> 0).  Same fp32 expression order as systems/ising.py:80-91:
    2 * s * (J * neighbors + h)."""
    dE = np.zeros(16, dtype=np.float32)
    for sb in (0, 1):
        s = np.float32(2 * sb - 1)
        for u in range(5):
            nb = np.float32(2 * u - 4)  # sum of four +/-1 neighbours, exact in fp32
            dE[8 * sb + u] = (np.float32(2) * s) * (np.float32(J) * nb + np.float32(h))
    return dE


def acceptance_table(J, h, temperature):
    """uint32 [16][2] {lo, hi}: accept <=> r < lo or hi, r the site's 32-bit draw (u = r*2^-32).

    samplers/mcmc.py:24   beta = 1/T if T > 0 else inf
    samplers/mcmc.py:43   dE <= 0            -> always accept          (hi = 1)
    samplers/mcmc.py:46   u < exp(-beta*dE)  -> r < ceil(p * 2^32)     (float64 exp, strict <)
    """
    beta = 1.0 / temperature if temperature > 0 else float("inf")
    dE = delta_e_classes(J, h)
    tbl = np.zeros((16, 2), dtype=np.uint32)
    for k in range(16):
        if (k & 7) > 4:
            continue
        d = float(dE[k])  # .item(): fp32 widened to float64
        if d <= 0:
            tbl[k] = (0, 1)
            continue
        with np.errstate(over="ignore", under="ignore"):
            p = float(np.exp(-beta * d))
        t = math.ceil(p * 4294967296.0)
        if t >= 4294967296:
            tbl[k] = (0, 1)  # p rounds to 1.0: u < p holds for every representable u
        else:
            tbl[k] = (t, 0)
    return tbl


# --------------------------------------------------------------------------- RNG streams
def init_bits_words(seed, chain_ids, L):
    """Hot start words, uint32 [n][L][W] — the stream of dbs_ising_init_packed."""
    W = (L + 31) // 32
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    i = np.arange(L, dtype=np.uint64)[None, :, None]
    w = np.arange(W, dtype=np.uint64)[None, None, :]
    x, _, _, _ = philox4x32_10(chain_ids[:, None, None], i, w, STREAM_INIT, seed)
    vb = np.minimum(32, L - 32 * np.arange(W))
    mask = np.where(vb >= 32, 0xFFFFFFFF, (1 << vb) - 1).astype(np.uint32)
    return x & mask[None, None, :]


def unpack(words, L):
    """uint32 [n][L][W] -> int8 [n][L][L] of +/-1."""
    n = words.shape[0]
    j = np.arange(L)
    bits = (words[:, :, j // 32] >> (j % 32).astype(np.uint32)) & np.uint32(1)
    return (2 * bits.astype(np.int8) - 1).reshape(n, L, L)


def pack(spins):
    """+/-1 (any numeric, test x > 0) [n][L][L] -> uint32 [n][L][W]."""
    spins = np.asarray(spins)
    n, L, _ = spins.shape
    W = (L + 31) // 32
    out = np.zeros((n, L, W), dtype=np.uint32)
    up = (spins > 0)
    for j in range(L):
        out[:, :, j // 32] |= up[:, :, j].astype(np.uint32) << np.uint32(j % 32)
    return out


def site_colours(L):
    """(n_colours, colour_of [L][L]).  Even L: the checkerboard (i + j) & 1.  Odd L: the torus is not bipartite
    ((i, L-1) and (i, 0) are neighbours of equal parity), so the cycle is three-coloured c(k) = k & 1 for k < L-1,
    c(L-1) = 2 and colour(i, j) = (c(i) + c(j)) mod 3 — adjacent sites differ in exactly one coordinate's colour."""
    k = np.arange(L)
    if L % 2 == 0:
        return 2, (k[:, None] + k[None, :]) & 1
    c = np.where(k == L - 1, 2, k & 1)
    return 3, (c[:, None] + c[None, :]) % 3


def sweep_draws(seed, chain_ids, sweep_index, L):
    """uint32 [n][L][L]: the draw of every site in sweep `sweep_index` (each site is visited once
    per sweep, in its colour's half-sweep).  Counter layout of dbs_ising_sweeps:
    ctr = {chain, 2*sweep+colour, i*GPR + q//4, STREAM_SWEEP}, word q%4, q = rank of the site
    among the sites of its colour in row i."""
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    n = chain_ids.shape[0]
    if L % 2:
        # odd L (three colour classes): ctr = {chain, sweep, q//4, STREAM_SWEEP3}, word q%4, q = i*L + j
        q = np.arange(L * L).reshape(L, L)
        y = np.uint64(sweep_index & 0xFFFFFFFF)
        o = philox4x32_10(chain_ids[:, None, None], y, (q // 4).astype(np.uint64)[None], STREAM_SWEEP3, seed)
        r = np.choose(np.broadcast_to((q % 4)[None], (n, L, L)), [np.broadcast_to(x, (n, L, L)) for x in o])
        return r.astype(np.uint32)
    gpr = (L + 7) // 8
    i = np.arange(L)[:, None]
    j = np.arange(L)[None, :]
    colour = (i + j) & 1
    q = j // 2
    grp = (i * gpr + q // 4).astype(np.uint64)
    y = (2 * sweep_index + colour).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    o = philox4x32_10(chain_ids[:, None, None], y[None], grp[None], STREAM_SWEEP, seed)
    sel = (q % 4)[None]
    r = np.choose(np.broadcast_to(sel, (n, L, L)), [np.broadcast_to(x, (n, L, L)) for x in o])
    return r.astype(np.uint32)


# --------------------------------------------------------------------------- the sweep
def checkerboard_sweeps(spins, tables, table_index, draws):
    """In-place colour-class Metropolis (checkerboard for even L, three classes for odd L: site_colours).
    spins int8 [n][L][L]; tables uint32 [T][16][2]; table_index int [n] or None;
    draws: callable(sweep)->uint32 [n][L][L], one sweep per call."""
    n, L, _ = spins.shape
    tables = np.asarray(tables, dtype=np.uint32).reshape(-1, 16, 2)
    ti = np.zeros(n, dtype=np.int64) if table_index is None else np.asarray(table_index, dtype=np.int64)
    n_colours, colour_of = site_colours(L)
    s_idx = 0
    while True:
        r = draws(s_idx)
        if r is None:
            break
        for colour in range(n_colours):
            up = (spins > 0).astype(np.int64)
            u = np.roll(up, 1, axis=1) + np.roll(up, -1, axis=1) + np.roll(up, 1, axis=2) + np.roll(up, -1, axis=2)
            k = 8 * up + u
            e = tables[ti[:, None, None], k]  # [n][L][L][2]
            accept = (r < e[..., 0]) | (e[..., 1] != 0)
            flip = accept & (colour_of[None] == colour)
            spins[flip] *= -1
        s_idx += 1
    return spins


def run_sweeps(spins, J, h, temperatures, n_sweeps, seed=0, chain_ids=None, sweep_offset=0, uniforms=None):
    """Convenience driver: temperatures scalar or per-chain array; uniforms uint32
    [n_sweeps][n][L][L] overrides the Philox stream (the injection port)."""
    n, L, _ = spins.shape
    temps = np.broadcast_to(np.asarray(temperatures, dtype=np.float64), (n,))
    uniq, inv = np.unique(temps, return_inverse=True)
    tables = np.stack([acceptance_table(J, h, float(t)) for t in uniq])
    ids = np.arange(n) if chain_ids is None else np.asarray(chain_ids)

    def draws(s):
        if s >= n_sweeps:
            return None
        if uniforms is not None:
            return uniforms[s]
        return sweep_draws(seed, ids, sweep_offset + s, L)

    return checkerboard_sweeps(spins, tables, inv, draws)


# --------------------------------------------------------------------------- observables
def energy(spins, J=1.0, h=0.0):
    """systems/ising.py:47-66 in fp32: -J*sum(s*right + s*down) + (-h)*sum(s)."""
    x = np.asarray(spins, dtype=np.float32)
    right = np.roll(x, -1, axis=-1)
    down = np.roll(x, -1, axis=-2)
    bond = (x * right + x * down).astype(np.float64).sum(axis=(-1, -2)).astype(np.float32)
    tot = x.astype(np.float64).sum(axis=(-1, -2)).astype(np.float32)
    return (np.float32(-J) * bond + np.float32(-h) * tot).astype(np.float32)


def magnetization(spins):
    """systems/ising.py:93-102: mean over the lattice, fp32."""
    x = np.asarray(spins, dtype=np.float32)
    L2 = x.shape[-1] * x.shape[-2]
    tot = x.astype(np.float64).sum(axis=(-1, -2)).astype(np.float32)
    return (tot / np.float32(L2)).astype(np.float32)


def correlation_bins(L):
    """bin index r = int(sqrt(di^2+dj^2)) for (di,dj) in [-R,R]^2 (analysis/correlation.py:34-37);
    -1 where the pair is skipped (r == 0 or r >= R)."""
    R = L // 2
    d = np.arange(-R, R + 1)
    rr = np.sqrt(d[:, None] ** 2 + d[None, :] ** 2).astype(np.int64)
    return np.where((rr > 0) & (rr < R), rr, -1)


def pair_correlation_sums(samples, ref_indices=None):
    """Integer sums and counts of analysis/correlation.py:28-42 for given reference sites
    (flattened indices; None = every site).  Returns (sums int64 [R], counts int64 [R])."""
    s = np.asarray(samples).astype(np.int64)
    n, L, _ = s.shape
    R = L // 2
    bins = correlation_bins(L)
    mask = np.zeros((L, L), dtype=np.int64)
    if ref_indices is None:
        mask[:] = 1
    else:
        for idx in ref_indices:
            mask[idx // L, idx % L] += 1
    sums = np.zeros(R, dtype=np.int64)
    counts = np.zeros(R, dtype=np.int64)
    for a, di in enumerate(range(-R, R + 1)):
        for b, dj in enumerate(range(-R, R + 1)):
            r = bins[a, b]
            if r < 0:
                continue
            shifted = np.roll(s, (-di, -dj), axis=(1, 2))  # shifted[i,j] = s[i+di, j+dj]
            sums[r] += int((s * shifted * mask[None]).sum())
            counts[r] += int(mask.sum()) * n
    return sums, counts


def pair_correlation(samples, ref_indices=None):
    """{'r': [...], 'C_r': [...]} exactly as analysis/correlation.py:43-47."""
    sums, counts = pair_correlation_sums(samples, ref_indices)
    c = np.divide(sums.astype(np.float64), counts.astype(np.float64), where=counts > 0,
                  out=np.zeros(len(sums), dtype=np.float64))
    return {"r": list(range(len(sums))), "C_r": c.tolist()}


# --------------------------------------------------------------------------- the reference's own chain, as is
def random_site_sweeps(spins, J, h, temperature, n_sweeps):
    """MetropolisHastings.step / sweep exactly as the reference runs them (samplers/mcmc.py:26-60 with
    IsingModel.local_energy_diff, systems/ising.py:68-91): one random site per attempt from the GLOBAL NumPy generator
    (two randint calls), dE from five torch tensor reads, `dE <= 0` or `np.random.random() < np.exp(-beta * dE.item())`
    (the uniform is drawn only when dE > 0), flip in place; a sweep is L*L such attempts.  `spins` is a torch fp32
    (L, L) tensor of +/-1, updated in place and returned.  This is the CPU baseline timed for BASELINE configs 1 and 3
    ("reference CPU path as-is"): same operations, same draw order, so under np.random.seed it reproduces the
    reference's chain bit for bit (tests/golden/ising_random_site.npz)."""
    L = spins.shape[-1]
    beta = 1.0 / temperature if temperature > 0 else float("inf")
    for _ in range(n_sweeps * L * L):
        i = np.random.randint(L)
        j = np.random.randint(L)
        s = spins[..., i, j]
        neighbors = (spins[..., (i + 1) % L, j] + spins[..., (i - 1) % L, j]
                     + spins[..., i, (j + 1) % L] + spins[..., i, (j - 1) % L])
        dE = 2 * s * (J * neighbors + h)
        if dE <= 0:
            spins[..., i, j] *= -1
        elif np.random.random() < np.exp(-beta * dE.item()):
            spins[..., i, j] *= -1
    return spins
