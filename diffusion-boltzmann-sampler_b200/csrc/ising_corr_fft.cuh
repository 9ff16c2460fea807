// ising_corr_fft.cuh — per-thread steps of the spectral C(r) kernel (64 x 64 lattices), written as host + device
// functions so that the index bookkeeping runs on the CPU (tests/host/corr_fft_host.cu) and on the GPU from one source.
//
// What is computed (analysis/correlation.py:8-47 with every site a reference site, as dbs_ising_correlation defines it):
//   acc[r] += sum over displacements d = (di, dj), |di|, |dj| < 32, floor(|d|) = r, 0 < r < 32, of
//             A(d) = sum_x s(x) s(x + d)            (periodic)
// Direct evaluation costs a chain 64 x 64-bit XOR / POPC words for each of ~1650 half-plane displacements and is bound by
// the POPC and ALU pipes.  Here A is obtained for ALL displacements at once by the Wiener-Khinchin route:
//   z1 = a + i b, z2 = c + i d  (four chains),  P_k = |FFT2(z_k)|^2,  q = IFFT2(P_1 + i P_2)
//   Re q(d) = A_a(d) + A_b(d) - X_2(d),  Im q(d) = A_c(d) + A_d(d) + X_1(d),  X_k(d) = sum_x im(z_k)(x + d) re(z_k)(x) - (d -> -d)
// X_k is odd in d and a radial bin holds d and -d together, so the cross terms cancel EXACTLY in the integer bin sums:
// every q(d) is an even integer of magnitude <= 16384 carried in fp32 with an absolute error ~2e-3 (measured; bounded by
// ~1e-6 x 16384), rounded to the nearest integer before it is added.  The sums are therefore the same integers the
// XOR / POPC kernel produces — tests/test_ising_gpu.py compares the two bit for bit — at six 64-point transforms of
// ~1.1 k FP32 instructions per thread and four chains instead of ~130 k integer word operations per chain.
//
// One thread holds one 64-point line in registers in each of the six passes; ONE forward transform body serves all of
// them (the inverse is conj . FFT . conj), so the kernel's code stays inside the instruction cache:
//   pass 0  rows of z1 from the packed bits        -> FFT -> buf (row t)
//   pass 1  column t of buf                        -> FFT -> P1 = |.|^2 parked in a thread-private column of pw
//   pass 2  rows of z2                             -> FFT -> buf
//   pass 3  column t of buf                        -> FFT -> P2 = |.|^2;  v <- conj(P1 + i P2) in natural order
//   pass 4  (v as left by pass 3)                  -> FFT -> buf (column t)
//   pass 5  row t of buf                           -> FFT -> conj -> 4096 q(di = t, dj) -> rounded, radial bins
#pragma once
#include <math.h>
#include <stdint.h>

#include "fft64.cuh"

namespace dbs {
namespace cfft {

using fft::cpx;
using fft::fft64;
using fft::fft64_perm;

constexpr int kStride = 65;                   // row stride of the complex exchange buffer in elements: conflict-free for
                                              // 8-byte accesses both along a row (t * 65 + k) and along a column (k * 65 + t)
constexpr float kScale = 1.0f / 4096.0f;      // the two unnormalised inverse passes
constexpr float kMagic = 12582912.0f;         // 1.5 * 2^23: (x + kMagic) - kMagic = rint(x) for |x| < 2^22, on the FMA pipe

__host__ __device__ constexpr unsigned char radial_bin(int di, int dj) {   // floor(sqrt(di^2 + dj^2)) if in (0, 32), else 255
  const int x = di * di + dj * dj;
  int rr = 0;
  while ((rr + 1) * (rr + 1) <= x) ++rr;
  return (rr > 0 && rr < 32) ? (unsigned char)rr : (unsigned char)255;
}
struct BinTable { unsigned char v[32 * 32]; };   // [|di|][|dj|]
__host__ __device__ constexpr BinTable make_bin_table() {
  BinTable b{};
  for (int a = 0; a < 32; ++a)
    for (int c = 0; c < 32; ++c) b.v[a * 32 + c] = radial_bin(a, c);
  return b;
}

// passes 0 / 2, before the transform: row t of two lattices as re / im; a missing lattice has amplitude 0
__host__ __device__ __forceinline__ void load_bits(cpx (&v)[64], uint64_t row_re, float amp_re, uint64_t row_im, float amp_im) {
#pragma unroll
  for (int x = 0; x < 64; ++x) {
    v[x].x = ((row_re >> x) & 1ull) ? amp_re : -amp_re;
    v[x].y = ((row_im >> x) & 1ull) ? amp_im : -amp_im;
  }
}
// register index j of a transformed line holds element perm(j)
__host__ __device__ __forceinline__ void store_row(int t, const cpx (&v)[64], cpx* buf) {
#pragma unroll
  for (int j = 0; j < 64; ++j) buf[t * kStride + fft64_perm(j)] = v[j];
}
__host__ __device__ __forceinline__ void store_col(int t, const cpx (&v)[64], cpx* buf) {
#pragma unroll
  for (int j = 0; j < 64; ++j) buf[fft64_perm(j) * kStride + t] = v[j];
}
__host__ __device__ __forceinline__ void load_col(int t, const cpx* buf, cpx (&v)[64]) {
#pragma unroll
  for (int y = 0; y < 64; ++y) v[y] = buf[y * kStride + t];
}
__host__ __device__ __forceinline__ void load_row(int t, const cpx* buf, cpx (&v)[64]) {
#pragma unroll
  for (int k = 0; k < 64; ++k) v[k] = buf[t * kStride + k];
}
// pass 1, after the transform: power spectrum of column t into the thread's own column of pw (index j <-> ky = perm(j))
__host__ __device__ __forceinline__ void store_power(int t, const cpx (&v)[64], float* pw) {
#pragma unroll
  for (int j = 0; j < 64; ++j) pw[j * 64 + t] = v[j].x * v[j].x + v[j].y * v[j].y;
}
// pass 3, after the transform: v <- conj(P1 + i P2) in natural order of ky (perm is its own inverse)
__host__ __device__ __forceinline__ void build_q(int t, cpx (&v)[64], const float* pw) {
  float p2[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) p2[j] = v[j].x * v[j].x + v[j].y * v[j].y;
#pragma unroll
  for (int n = 0; n < 64; ++n) v[n] = {pw[fft64_perm(n) * 64 + t], -p2[fft64_perm(n)]};
}

// A sum of up to four values 4096 q(d), each within ~2e-3 of 4096 x an even integer: scaled and rounded to the integer it is
__host__ __device__ __forceinline__ float round_scaled(float x, float& maxdev) {
  const float s = x * kScale;
  const float r = (s + kMagic) - kMagic;
  maxdev = fmaxf(maxdev, fabsf(s - r));
  return r;
}

// pass 5, after the transform: v[j] = conj(4096 q(di = t, dj = perm(j))).  part[bin * 64 + t] += w_re Re q + w_im Im q over
// dj = -31 .. 31 (exact small integers held as floats); bin_tab[|di| * 32 + |dj|] as radial_bin.  The (up to) four values
// of a |dj| are added before the rounding: their errors (~2e-3 each) cannot reach the next integer.
__host__ __device__ __forceinline__ void bin_row(int t, const cpx (&v)[64], float w_re, float w_im,
                                                 const unsigned char* bin_tab, float* part, float& maxdev) {
  if (t == 32) return;                       // |di| = 32: no displacement with r < 32
  const int adi = t < 32 ? t : 64 - t;
  const unsigned char* brow = bin_tab + adi * 32;
  const float w_imc = -w_im;                 // the final conjugation
#pragma unroll
  for (int adj = 0; adj < 32; ++adj) {
    const int jp = fft64_perm(adj), jm = fft64_perm((64 - adj) & 63);
    const float re = adj > 0 ? v[jp].x + v[jm].x : v[jp].x, im = adj > 0 ? v[jp].y + v[jm].y : v[jp].y;
    const float val = round_scaled(w_re * re + w_imc * im, maxdev);
    const int bin = brow[adj];
    if (bin != 255) part[bin * 64 + t] += val;
  }
}

}  // namespace cfft
}  // namespace dbs
