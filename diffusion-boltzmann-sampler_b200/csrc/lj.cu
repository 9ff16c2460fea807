// lj.cu — Lennard-Jones pair energy and forces, one CTA per configuration.
//
// Replaces LennardJonesSystem.minimum_image / pairwise_distances / energy / forces
// (systems/lennard_jones.py:58-155).  The reference materialises four (B,N,N[,2]) fp32
// temporaries; here the 2N floats of one configuration sit in shared memory, thread i walks
// j = 0..N-1 and energy + force come out of one pass.  Per-pair arithmetic follows the
// reference operation by operation (IEEE div/sqrt, round-half-even minimum image, the
// 1e-10 epsilons, the 0.1 < d < rc mask); only the summation order differs.
// FP32-ALU bound: N^2 ordered pairs * ~40 flop; 16 N + 4 bytes of HBM traffic per configuration.
#include <math.h>

#include "common.cuh"

namespace dbs {
namespace {

struct LjArgs {
  const float* pos;
  float* energy;
  float* forces;
  int N;
  float box, inv_box, four_eps, twentyfour_eps, sigma, cutoff, shift;
  float d2_lo, d2_hi;   // dist > 0.1f && dist < cutoff  <=>  d2_lo <= (dx^2+dy^2+1e-10) <= d2_hi  (exact, see host)
  int fast_image;       // 1: rint(dx * (1/box)) instead of rint(dx / box) (identical unless |dx| ~ box/2 >= cutoff)
};

// Two phases per 32-wide block of partners j: (1) cheap minimum-image distance test for all 32,
// collecting an in-range bit mask (no sqrt: the mask thresholds are pre-mapped to squared distance);
// (2) the expensive Lennard-Jones arithmetic (IEEE sqrt / divisions, the reference's operation order)
// only for the set bits, ~6 % of the pairs at rho* = 0.8.  Partner order stays ascending, so sums
// are bit-identical to the straightforward loop.
__global__ void __launch_bounds__(256) lj_kernel(LjArgs a) {
  extern __shared__ float sp[];  // [N][2]
  const int N = a.N;
  const float* p = a.pos + (int64_t)blockIdx.x * N * 2;
  for (int k = threadIdx.x; k < 2 * N; k += blockDim.x) sp[k] = p[k];
  __syncthreads();
  float e_acc = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float xi = sp[2 * i], yi = sp[2 * i + 1];
    float fx = 0.f, fy = 0.f, ei = 0.f;
    for (int j0 = 0; j0 < N; j0 += 32) {
      uint32_t mask = 0;
      const int jn = min(32, N - j0);
#pragma unroll 8
      for (int b = 0; b < jn; ++b) {
        const float2 pj = *reinterpret_cast<const float2*>(&sp[2 * (j0 + b)]);
        float dx = __fsub_rn(xi, pj.x), dy = __fsub_rn(yi, pj.y);
        const float kx = a.fast_image ? rintf(__fmul_rn(dx, a.inv_box)) : rintf(__fdiv_rn(dx, a.box));
        const float ky = a.fast_image ? rintf(__fmul_rn(dy, a.inv_box)) : rintf(__fdiv_rn(dy, a.box));
        dx = __fsub_rn(dx, __fmul_rn(a.box, kx));
        dy = __fsub_rn(dy, __fmul_rn(a.box, ky));
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 1e-10f);
        mask |= (uint32_t)(d2 >= a.d2_lo && d2 <= a.d2_hi) << b;
      }
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const float2 pj = *reinterpret_cast<const float2*>(&sp[2 * (j0 + b)]);
        float dx = __fsub_rn(xi, pj.x), dy = __fsub_rn(yi, pj.y);
        const float kx = a.fast_image ? rintf(__fmul_rn(dx, a.inv_box)) : rintf(__fdiv_rn(dx, a.box));
        const float ky = a.fast_image ? rintf(__fmul_rn(dy, a.inv_box)) : rintf(__fdiv_rn(dy, a.box));
        dx = __fsub_rn(dx, __fmul_rn(a.box, kx));
        dy = __fsub_rn(dy, __fmul_rn(a.box, ky));
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 1e-10f);
        const float ds = __fsqrt_rn(d2);
        const float q = __fdiv_rn(a.sigma, ds);
        const float q2 = __fmul_rn(q, q);
        const float r6 = __fmul_rn(__fmul_rn(q2, q2), q2);
        const float r12 = __fmul_rn(r6, r6);
        ei = __fadd_rn(ei, __fsub_rn(__fmul_rn(a.four_eps, __fsub_rn(r12, r6)), a.shift));
        const float fm = __fmul_rn(__fdiv_rn(a.twentyfour_eps, ds), __fsub_rn(__fmul_rn(2.0f, r12), r6));
        const float den = __fadd_rn(ds, 1e-10f);
        fx = __fadd_rn(fx, __fmul_rn(fm, __fdiv_rn(dx, den)));
        fy = __fadd_rn(fy, __fmul_rn(fm, __fdiv_rn(dy, den)));
      }
    }
    e_acc += ei;
    if (a.forces) {
      float* f = a.forces + ((int64_t)blockIdx.x * N + i) * 2;
      f[0] = fx;
      f[1] = fy;
    }
  }
  if (a.energy) {
    __shared__ double red[32];
    double e = warp_sum((double)e_acc);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) red[wid] = e;
    __syncthreads();
    if (wid == 0) {
      const int nw = (blockDim.x + 31) >> 5;
      e = lane < nw ? red[lane] : 0.0;
      e = warp_sum(e);
      if (lane == 0) a.energy[blockIdx.x] = (float)(0.5 * e);
    }
  }
}

// smallest / largest fp32 x with sqrtf(x) > lo resp. sqrtf(x) < hi: maps the reference's mask on
// dist = sqrt(d2 + 1e-10) (systems/lennard_jones.py:104) onto the squared distance exactly.
float d2_threshold_above(float lo) {
  float x = lo * lo;
  while (sqrtf(x) > lo) x = nextafterf(x, 0.f);
  while (!(sqrtf(x) > lo)) x = nextafterf(x, INFINITY);
  return x;
}
float d2_threshold_below(float hi) {
  float x = hi * hi;
  while (sqrtf(x) < hi) x = nextafterf(x, INFINITY);
  while (!(sqrtf(x) < hi)) x = nextafterf(x, 0.f);
  return x;
}

}  // namespace
}  // namespace dbs

using namespace dbs;

extern "C" int dbs_lj_energy_forces(const float* pos, int64_t B, int N, float box, float epsilon, float sigma,
                                    float cutoff, float energy_shift, float* energy, float* forces,
                                    void* stream) {
  DBS_REQUIRE(pos && B >= 0, "lj: null positions");
  DBS_REQUIRE(N >= 1 && N <= 8192, "lj: n_particles %d unsupported (1..8192)", N);
  DBS_REQUIRE(box > 0.f, "lj: box_size must be positive");
  if (B == 0) return DBS_OK;
  LjArgs a;
  a.pos = pos; a.energy = energy; a.forces = forces; a.N = N; a.box = box;
  a.four_eps = (float)(4.0 * (double)epsilon);
  a.twentyfour_eps = (float)(24.0 * (double)epsilon);
  a.sigma = sigma; a.cutoff = cutoff; a.shift = energy_shift;
  a.inv_box = 1.0f / box;
  a.d2_lo = d2_threshold_above(0.1f);
  a.d2_hi = d2_threshold_below(cutoff);
  a.fast_image = cutoff <= 0.5f * box ? 1 : 0;
  int threads = N >= 256 ? 256 : ((N + 31) / 32) * 32;
  const size_t smem = (size_t)N * 2 * sizeof(float);
  if (smem > 48 * 1024)
    DBS_CUDA(cudaFuncSetAttribute(lj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(kProfLj, as_stream(stream));
  lj_kernel<<<(unsigned)B, threads, smem, as_stream(stream)>>>(a);
  DBS_LAUNCHED();
  return DBS_OK;
}
