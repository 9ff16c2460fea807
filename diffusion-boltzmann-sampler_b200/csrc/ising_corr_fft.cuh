// ising_corr_fft.cuh — per-thread phases of the spectral C(r) kernel (64 x 64 lattices), written as host + device
// functions so that the index bookkeeping can be run on the CPU (oracle-side test harness) and on the GPU from one source.
//
// What is computed (analysis/correlation.py:8-47 with every site a reference site, as dbs_ising_correlation defines it):
//   acc[r] += sum over displacements d = (di, dj), |di|, |dj| < 32, floor(|d|) = r, 0 < r < 32, of
//             A(d) = sum_x s(x) s(x + d)            (periodic)
// Direct evaluation costs a chain 64 x 64-bit XOR / POPC words for each of ~1650 half-plane displacements and is bound by
// the POPC and ALU pipes.  Here A is obtained for ALL displacements at once by the Wiener-Khinchin route:
//   z1 = a + i b, z2 = c + i d  (four chains),  P_k = |FFT2(z_k)|^2,  q = IFFT2(P_1 + i P_2)
//   Re q(d) = A_a(d) + A_b(d) - X_2(d),  Im q(d) = A_c(d) + A_d(d) + X_1(d),  X_k(d) = sum_x im(z_k)(x + d) re(z_k)(x) - (d -> -d)
// X_k is odd in d and a radial bin holds d and -d together, so the cross terms cancel EXACTLY in the integer bin sums:
// every q(d) is an even integer of magnitude <= 16384 carried in fp32 with an absolute error ~1e-2 (measured; bounded by
// ~1e-6 x 16384), rounded to the nearest integer before it is added.  The sums are therefore the same integers the
// XOR / POPC kernel produces — tests/test_ising_gpu.py compares the two bit for bit — at ~3 x 1.1 k FP32 instructions per
// thread and 2-D pass instead of ~130 k integer word operations per chain.
#pragma once
#include <stdint.h>

#include "fft64.cuh"

namespace dbs {
namespace cfft {

using fft::cpx;
using fft::fft64;
using fft::fft64_perm;

constexpr int kL = 64;
constexpr int kStride = 65;                   // row stride of the complex exchange buffer in elements: conflict-free for
                                              // 8-byte accesses both along a row (t * 65 + k) and along a column (k * 65 + t)
constexpr float kScale = 1.0f / 4096.0f;      // the two unnormalised inverse passes
constexpr float kMagic = 12582912.0f;         // 1.5 * 2^23: (x + kMagic) - kMagic = rint(x) for |x| < 2^22, on the FMA pipe

// Phase A: thread t takes row t of the two lattices packed as re / im, transforms along x, leaves Z_row[t][kx] in buf.
__host__ __device__ __forceinline__ void phase_rows_fwd(int t, uint64_t row_re, float amp_re, uint64_t row_im, float amp_im,
                                                        cpx* buf) {
  cpx v[64];
#pragma unroll
  for (int x = 0; x < 64; ++x) {
    v[x].x = ((row_re >> x) & 1ull) ? amp_re : -amp_re;
    v[x].y = ((row_im >> x) & 1ull) ? amp_im : -amp_im;
  }
  fft64<false>(v);
#pragma unroll
  for (int j = 0; j < 64; ++j) buf[t * kStride + fft64_perm(j)] = v[j];
}

// Phase B: thread t takes column kx = t, transforms along y and returns the power spectrum |Z[ky][t]|^2 at register
// index j (ky = perm(j)).  The caller synchronises between the loads (first loop) and any later store into buf.
__host__ __device__ __forceinline__ void phase_cols_load(int t, const cpx* buf, cpx (&v)[64]) {
#pragma unroll
  for (int y = 0; y < 64; ++y) v[y] = buf[y * kStride + t];
}
__host__ __device__ __forceinline__ void phase_cols_power(cpx (&v)[64], float (&p)[64]) {
  fft64<false>(v);
#pragma unroll
  for (int j = 0; j < 64; ++j) p[j] = v[j].x * v[j].x + v[j].y * v[j].y;
}

// Phase C: q = P_1 + i P_2 (both at register index j <-> ky = perm(j)), inverse transform along ky, result rows to buf.
__host__ __device__ __forceinline__ void phase_cols_inv(int t, const float (&p1)[64], const float (&p2)[64], cpx* buf) {
  cpx w[64];
#pragma unroll
  for (int n = 0; n < 64; ++n) w[n] = {p1[fft64_perm(n)], p2[fft64_perm(n)]};   // natural order in ky (perm is an involution)
  fft64<true>(w);
#pragma unroll
  for (int j = 0; j < 64; ++j) buf[fft64_perm(j) * kStride + t] = w[j];
}

// Phase D: thread t takes row di = t, inverse transform along kx; v[j] = 4096 * q(di = t, dj = perm(j)).
__host__ __device__ __forceinline__ void phase_rows_inv(int t, const cpx* buf, cpx (&v)[64]) {
#pragma unroll
  for (int k = 0; k < 64; ++k) v[k] = buf[t * kStride + k];
  fft64<true>(v);
}

__host__ __device__ __forceinline__ float round_scaled(float x, float& maxdev) {
  const float s = x * kScale;
  const float r = (s + kMagic) - kMagic;
  const float d = s - r;
  maxdev = maxdev > (d < 0.f ? -d : d) ? maxdev : (d < 0.f ? -d : d);
  return r;
}

// Binning of thread t's row: part[bin * 64 + t] += w_re * Re q + w_im * Im q over dj = -31 .. 31 (exact small integers
// held as floats); bin_tab[|di| * 32 + |dj|] = floor(sqrt(di^2 + dj^2)) when 0 < r < 32, else 255.
__host__ __device__ __forceinline__ void phase_bin(int t, const cpx (&v)[64], float w_re, float w_im,
                                                   const unsigned char* bin_tab, float* part, float& maxdev) {
  if (t == 32) return;                       // |di| = 32: no displacement with r < 32
  const int adi = t < 32 ? t : 64 - t;
  const unsigned char* brow = bin_tab + adi * 32;
#pragma unroll
  for (int adj = 0; adj < 32; ++adj) {
    const int jp = fft64_perm(adj), jm = fft64_perm((64 - adj) & 63);
    float val = w_re * round_scaled(v[jp].x, maxdev) + w_im * round_scaled(v[jp].y, maxdev);
    if (adj > 0) val += w_re * round_scaled(v[jm].x, maxdev) + w_im * round_scaled(v[jm].y, maxdev);
    const int bin = brow[adj];
    if (bin != 255) part[bin * 64 + t] += val;
  }
}

__host__ __device__ inline unsigned char radial_bin(int di, int dj) {   // floor(sqrt(di^2 + dj^2)) if in (0, 32), else 255
  const int x = di * di + dj * dj;
  int rr = 0;
  while ((rr + 1) * (rr + 1) <= x) ++rr;
  return (rr > 0 && rr < 32) ? (unsigned char)rr : (unsigned char)255;
}

}  // namespace cfft
}  // namespace dbs
