// CPU harness for the spectral C(r) kernel's per-thread steps (csrc/ising_corr_fft.cuh, csrc/fft64.cuh are host + device
// code): the 64 threads of a set are emulated pass by pass and the radial sums are compared with a direct evaluation of
//   sum over chains, displacements d with floor(|d|) = r, sites x of s(x) s(x + d)        (analysis/correlation.py:8-47)
// Built and run by tests/test_corr_fft_host.py with nvcc for the host only; exit code = number of mismatches.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../diffusion-boltzmann-sampler_b200/csrc/ising_corr_fft.cuh"

using namespace dbs::cfft;

static int spin(const uint64_t* lat, int i, int j) { return ((lat[i & 63] >> (j & 63)) & 1) ? 1 : -1; }

int main() {
  // the 64-point transform against a direct DFT
  double worst = 0;
  for (int trial = 0; trial < 8; ++trial) {
    cpx v[64];
    double xr[64], xi[64];
    for (int n = 0; n < 64; ++n) {
      xr[n] = (rand() & 1) ? 1 : -1;
      xi[n] = (rand() % 2001 - 1000) / 100.0;
      v[n] = {(float)xr[n], (float)xi[n]};
    }
    fft64<false>(v);
    for (int j = 0; j < 64; ++j) {
      const int k = fft64_perm(j);
      double sr = 0, si = 0;
      for (int n = 0; n < 64; ++n) {
        const double a = -2 * M_PI * n * k / 64.0;
        sr += xr[n] * cos(a) - xi[n] * sin(a);
        si += xr[n] * sin(a) + xi[n] * cos(a);
      }
      worst = fmax(worst, fmax(fabs(sr - v[j].x), fabs(si - v[j].y)));
    }
  }
  int bad = worst < 1e-3 ? 0 : 1;
  printf("fft64 worst abs err %g\n", worst);

  static const BinTable tab = make_bin_table();
  float worstdev = 0;
  for (int trial = 0; trial < 7; ++trial) {
    uint64_t lat[4][64];
    for (int c = 0; c < 4; ++c)
      for (int i = 0; i < 64; ++i) {
        uint64_t r = 0;
        for (int j = 0; j < 64; ++j) {
          int bit;
          if (trial == 0) bit = 1;                                // ordered
          else if (trial == 1) bit = ((i / 8 + j / 8) & 1);       // domains
          else if (trial == 2) bit = (rand() % 100) < 90;         // magnetised
          else bit = rand() & 1;
          r |= (uint64_t)bit << j;
        }
        lat[c][i] = r;
      }
    const int nvalid = trial == 5 ? 3 : trial == 6 ? 1 : 4;      // missing chains enter with amplitude 0
    // trial 4: "mixed groups" form — lattices 0 and 1 as two real-only inputs, Re q -> bins A, Im q -> bins B
    const bool singles = trial == 4;
    long long want[2][32] = {{0}};
    for (int c = 0; c < (singles ? 2 : nvalid); ++c)
      for (int di = -31; di <= 31; ++di)
        for (int dj = -31; dj <= 31; ++dj) {
          const unsigned char b = radial_bin(abs(di), abs(dj));
          if (b == 255) continue;
          long long s = 0;
          for (int i = 0; i < 64; ++i)
            for (int j = 0; j < 64; ++j) s += spin(lat[c], i, j) * spin(lat[c], i + di + 64, j + dj + 64);
          want[singles ? c : 0][b] += s;
        }
    std::vector<cpx> buf(64 * kStride);
    std::vector<float> pw(64 * 64), partA(32 * 64, 0.f), partB(32 * 64, 0.f);
    static cpx v[64][64];
    float amp[4];
    for (int c = 0; c < 4; ++c) amp[c] = c < nvalid ? 1.f : 0.f;
    const uint64_t* za = lat[0]; const uint64_t* zb = singles ? nullptr : lat[1];
    const uint64_t* zc = singles ? lat[1] : lat[2]; const uint64_t* zd = singles ? nullptr : lat[3];
    const float aa = amp[0], ab = singles ? 0.f : amp[1], ac = singles ? amp[1] : amp[2], ad = singles ? 0.f : amp[3];
    for (int t = 0; t < 64; ++t) { load_bits(v[t], za[t], aa, zb ? zb[t] : 0, ab); fft64<false>(v[t]); store_row(t, v[t], buf.data()); }
    for (int t = 0; t < 64; ++t) load_col(t, buf.data(), v[t]);
    for (int t = 0; t < 64; ++t) { fft64<false>(v[t]); store_power(t, v[t], pw.data()); }
    for (int t = 0; t < 64; ++t) { load_bits(v[t], zc[t], ac, zd ? zd[t] : 0, ad); fft64<false>(v[t]); store_row(t, v[t], buf.data()); }
    for (int t = 0; t < 64; ++t) load_col(t, buf.data(), v[t]);
    for (int t = 0; t < 64; ++t) { fft64<false>(v[t]); build_q(t, v[t], pw.data()); fft64<false>(v[t]); }
    for (int t = 0; t < 64; ++t) store_col(t, v[t], buf.data());
    float maxdev = 0;
    for (int t = 0; t < 64; ++t) {
      load_row(t, buf.data(), v[t]);
      fft64<false>(v[t]);
      if (singles) {
        bin_row(t, v[t], 1.f, 0.f, tab.v, partA.data(), maxdev);
        bin_row(t, v[t], 0.f, 1.f, tab.v, partB.data(), maxdev);
      } else {
        bin_row(t, v[t], 1.f, 1.f, tab.v, partA.data(), maxdev);
      }
    }
    worstdev = fmaxf(worstdev, maxdev);
    for (int b = 1; b < 32; ++b) {
      long long gotA = 0, gotB = 0;
      for (int t = 0; t < 64; ++t) { gotA += llrintf(partA[b * 64 + t]); gotB += llrintf(partB[b * 64 + t]); }
      if (gotA != want[0][b] || (singles && gotB != want[1][b])) {
        ++bad;
        printf("trial %d bin %d: got %lld / %lld, want %lld / %lld\n", trial, b, gotA, gotB, want[0][b], want[1][b]);
      }
    }
  }
  printf("mismatches %d, worst deviation of a value from an integer %g\n", bad, worstdev);
  if (worstdev > 0.05f) ++bad;
  return bad;
}
