// fft64.cuh — a 64-point complex FFT held entirely in one thread's registers (radix 8 x 8, every index a compile-time
// constant after unrolling, twiddles as immediates).  Used by the spectral C(r) kernel of ising_corr_fft.cu: the
// autocorrelation of a 64 x 64 lattice as |FFT2|^2 -> inverse FFT2 costs ~1.1 k FP32 instructions per 64-point
// transform instead of ~130 k XOR / POPC word operations per chain for the direct sum over displacements.
//
// In-place, decimation in frequency:  input  v[n]       (n = 8 n1 + n2, natural order)
//                                     output v[8 k1 + k2] = X[k1 + 8 k2]      (the two octal digits swapped)
// fft64_perm(j) maps a register index to the frequency (or, for the inverse, sample) index it holds; the map is its own
// inverse.  kInv = false: X[k] = sum x[n] exp(-2 pi i n k / 64); kInv = true: the conjugate kernel, no 1/64 factor.
#pragma once
#include <cuda_runtime.h>

namespace dbs {
namespace fft {

__host__ __device__ constexpr int fft64_perm(int j) { return (j >> 3) + 8 * (j & 7); }

// cos / sin of 2 pi m / 64, m = 0 .. 63 (only products n2 * k1 <= 49 are used)
__host__ __device__ constexpr float tw_cos(int m) {
  constexpr float c[17] = {1.0f, 0.99518472667219688624f, 0.98078528040323044913f, 0.95694033573220886494f,
                           0.92387953251128675613f, 0.88192126434835502971f, 0.83146961230254523708f,
                           0.77301045336273696081f, 0.70710678118654752440f, 0.63439328416364549822f,
                           0.55557023301960222474f, 0.47139673682599764856f, 0.38268343236508977173f,
                           0.29028467725446236764f, 0.19509032201612826785f, 0.09801714032956060199f, 0.0f};
  // quarter-wave symmetry
  const int q = m & 63;
  return q <= 16 ? c[q] : q <= 32 ? -c[32 - q] : q <= 48 ? -c[q - 32] : c[64 - q];
}
__host__ __device__ constexpr float tw_sin(int m) { return tw_cos((m + 48) & 63); }   // sin(x) = cos(x - pi/2)

struct cpx { float x, y; };

__host__ __device__ __forceinline__ cpx cadd(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
__host__ __device__ __forceinline__ cpx csub(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
// multiplication by -i (forward) / +i (inverse)
template <bool kInv>
__host__ __device__ __forceinline__ cpx rot90(cpx a) { return kInv ? cpx{-a.y, a.x} : cpx{a.y, -a.x}; }
// multiplication by exp(-+ 2 pi i m / 64)
template <bool kInv, int M>
__host__ __device__ __forceinline__ cpx twiddle(cpx a) {
  constexpr float c = tw_cos(M), s = tw_sin(M);
  if (M == 0) return a;
  if (kInv) return {a.x * c - a.y * s, a.y * c + a.x * s};
  return {a.x * c + a.y * s, a.y * c - a.x * s};
}

// 4-point DFT, natural order in and out
template <bool kInv>
__host__ __device__ __forceinline__ void dft4(cpx& b0, cpx& b1, cpx& b2, cpx& b3) {
  const cpx t0 = cadd(b0, b2), t1 = csub(b0, b2), t2 = cadd(b1, b3), t3 = rot90<kInv>(csub(b1, b3));
  b0 = cadd(t0, t2); b2 = csub(t0, t2); b1 = cadd(t1, t3); b3 = csub(t1, t3);
}

// 8-point DFT, natural order in and out
template <bool kInv>
__host__ __device__ __forceinline__ void dft8(cpx& a0, cpx& a1, cpx& a2, cpx& a3, cpx& a4, cpx& a5, cpx& a6, cpx& a7) {
  cpx s0 = cadd(a0, a4), s1 = cadd(a1, a5), s2 = cadd(a2, a6), s3 = cadd(a3, a7);
  cpx d0 = csub(a0, a4), d1 = twiddle<kInv, 8>(csub(a1, a5)), d2 = rot90<kInv>(csub(a2, a6)), d3 = twiddle<kInv, 24>(csub(a3, a7));
  dft4<kInv>(s0, s1, s2, s3);   // X0 X2 X4 X6
  dft4<kInv>(d0, d1, d2, d3);   // X1 X3 X5 X7
  a0 = s0; a2 = s1; a4 = s2; a6 = s3;
  a1 = d0; a3 = d1; a5 = d2; a7 = d3;
}

template <bool kInv, int K1>
__host__ __device__ __forceinline__ void twiddle_row(cpx (&v)[64]) {
  v[8 * K1 + 1] = twiddle<kInv, 1 * K1>(v[8 * K1 + 1]);
  v[8 * K1 + 2] = twiddle<kInv, 2 * K1>(v[8 * K1 + 2]);
  v[8 * K1 + 3] = twiddle<kInv, 3 * K1>(v[8 * K1 + 3]);
  v[8 * K1 + 4] = twiddle<kInv, 4 * K1>(v[8 * K1 + 4]);
  v[8 * K1 + 5] = twiddle<kInv, 5 * K1>(v[8 * K1 + 5]);
  v[8 * K1 + 6] = twiddle<kInv, 6 * K1>(v[8 * K1 + 6]);
  v[8 * K1 + 7] = twiddle<kInv, 7 * K1>(v[8 * K1 + 7]);
}

template <bool kInv>
__host__ __device__ __forceinline__ void fft64(cpx (&v)[64]) {
  // step 1: for each n2, DFT-8 over n1 of v[8 n1 + n2]; output index k1 takes the place of n1
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2)
    dft8<kInv>(v[n2], v[8 + n2], v[16 + n2], v[24 + n2], v[32 + n2], v[40 + n2], v[48 + n2], v[56 + n2]);
  // step 2: v[8 k1 + n2] *= W64^(n2 k1)
  twiddle_row<kInv, 1>(v); twiddle_row<kInv, 2>(v); twiddle_row<kInv, 3>(v); twiddle_row<kInv, 4>(v);
  twiddle_row<kInv, 5>(v); twiddle_row<kInv, 6>(v); twiddle_row<kInv, 7>(v);
  // step 3: for each k1, DFT-8 over n2; output index k2 takes the place of n2
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1)
    dft8<kInv>(v[8 * k1], v[8 * k1 + 1], v[8 * k1 + 2], v[8 * k1 + 3], v[8 * k1 + 4], v[8 * k1 + 5], v[8 * k1 + 6], v[8 * k1 + 7]);
}

}  // namespace fft
}  // namespace dbs
