// langevin.cu — batched Langevin samplers for the Lennard-Jones fluid and the
// radial-distribution histogram (f-2), built on the pair kernel of lj.cu.
//
// Replaces (reference file:line, /root/reference/backend/ml):
//   LangevinDynamics.step        samplers/langevin.py:46-89   (MALA: overdamped proposal + MH test)
//   UnderdampedLangevin.step     samplers/langevin.py:192-225 (BAOAB)
//   compute_rdf                  analysis/rdf.py:9-83         (pair-distance histogram)
// The reference evaluates forces twice and energies twice per MALA step on ONE configuration; here
// the state carries its cached force and energy, so a step is one proposal kernel, ONE pair pass
// on the proposal (dbs_lj_energy_forces) and one accept kernel, for B configurations at once.
#include "common.cuh"
#include "philox.cuh"

#define DBS_STREAM_LANGEVIN 0x4c414e47u /* 'LANG' */
#define DBS_STREAM_ACCEPT 0x41434350u   /* 'ACCP' */

namespace dbs {
namespace {

__device__ __forceinline__ void gauss4(uint32_t seed_lo, uint32_t seed_hi, uint32_t cfg, uint32_t quad, uint32_t step,
                                       float z[4]) {
  u32x4 c;
  c.x = cfg; c.y = quad; c.z = step; c.w = DBS_STREAM_LANGEVIN;
  const u32x4 o = philox4x32_10(c, seed_lo, seed_hi);
  const float k = 5.9604644775390625e-8f;
  const float u1 = ((float)(o.x >> 8) + 0.5f) * k, u2 = ((float)(o.y >> 8) + 0.5f) * k;
  const float u3 = ((float)(o.z >> 8) + 0.5f) * k, u4 = ((float)(o.w >> 8) + 0.5f) * k;
  const float r0 = sqrtf(-2.0f * logf(u1)), r1 = sqrtf(-2.0f * logf(u3));
  float s, cs;
  sincospif(2.0f * u2, &s, &cs);
  z[0] = r0 * cs; z[1] = r0 * s;
  sincospif(2.0f * u4, &s, &cs);
  z[2] = r1 * cs; z[3] = r1 * s;
}

// torch's `%` for floats: result has the sign of the divisor
__device__ __forceinline__ float wrap_box(float x, float box) {
  float r = fmodf(x, box);
  if (r != 0.f && ((r < 0.f) != (box < 0.f))) r += box;
  return r;
}

// x' = ((x + (dt*F)/T) + z*noise_scale) mod box          (samplers/langevin.py:56-63)
__global__ void mala_propose_kernel(const float* __restrict__ pos, const float* __restrict__ force,
                                    const float* __restrict__ noise, float* __restrict__ pos_new, int64_t B,
                                    int64_t elems, float dt, float temperature, float noise_scale, float box,
                                    uint32_t seed_lo, uint32_t seed_hi, int64_t cfg_offset, uint32_t step) {
  const int64_t quads = (elems + 3) / 4;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B * quads) return;
  const int64_t b = g / quads, q = g - b * quads;
  float z[4];
  if (!noise) gauss4(seed_lo, seed_hi, (uint32_t)(cfg_offset + b), (uint32_t)q, step, z);
  for (int k = 0; k < 4; ++k) {
    const int64_t e = 4 * q + k;
    if (e >= elems) break;
    const int64_t i = b * elems + e;
    const float zz = noise ? noise[i] : z[k];
    const float drift = __fdiv_rn(__fmul_rn(dt, force[i]), temperature);
    pos_new[i] = wrap_box(__fadd_rn(__fadd_rn(pos[i], drift), __fmul_rn(zz, noise_scale)), box);
  }
}

// Metropolis–Hastings test of the MALA proposal (samplers/langevin.py:65-87), one CTA per configuration:
//   log_q_fwd = -0.5 * sum((x' - (x + dt F/T))^2) / (2 dt),  log_q_bwd likewise with roles swapped
//   accept  <=>  log(u) < -beta (E' - E) + log_q_bwd - log_q_fwd
// On acceptance the state (positions, cached force, cached energy) is replaced by the proposal's.
__global__ void __launch_bounds__(256)
mala_accept_kernel(float* __restrict__ pos, float* __restrict__ force, float* __restrict__ energy,
                   const float* __restrict__ pos_new, const float* __restrict__ force_new,
                   const float* __restrict__ energy_new, const float* __restrict__ uniforms, int elems, float dt,
                   float temperature, uint32_t seed_lo, uint32_t seed_hi, int64_t cfg_offset, uint32_t step,
                   int* __restrict__ accepted) {
  const int64_t b = blockIdx.x;
  const int64_t base = b * elems;
  float fwd = 0.f, bwd = 0.f;
  for (int e = threadIdx.x; e < elems; e += blockDim.x) {
    const float x = pos[base + e], xn = pos_new[base + e];
    const float df = __fadd_rn(x, __fdiv_rn(__fmul_rn(dt, force[base + e]), temperature));
    const float db = __fadd_rn(xn, __fdiv_rn(__fmul_rn(dt, force_new[base + e]), temperature));
    const float a = __fsub_rn(xn, df), c = __fsub_rn(x, db);
    fwd = fmaf(a, a, fwd);
    bwd = fmaf(c, c, bwd);
  }
  __shared__ float sf[32], sb[32];
  __shared__ int s_acc;
  fwd = warp_sum(fwd); bwd = warp_sum(bwd);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { sf[wid] = fwd; sb[wid] = bwd; }
  __syncthreads();
  if (wid == 0) {
    fwd = lane < nw ? sf[lane] : 0.f;
    bwd = lane < nw ? sb[lane] : 0.f;
    fwd = warp_sum(fwd); bwd = warp_sum(bwd);
    if (lane == 0) {
      const float lq_f = -0.5f * fwd / (2.0f * dt), lq_b = -0.5f * bwd / (2.0f * dt);
      const float beta = 1.0f / temperature;
      const float log_alpha = -beta * (energy_new[b] - energy[b]) + lq_b - lq_f;
      double u;
      if (uniforms) {
        u = (double)uniforms[b];
      } else {
        u32x4 c;
        c.x = (uint32_t)(cfg_offset + b); c.y = 0; c.z = step; c.w = DBS_STREAM_ACCEPT;
        const u32x4 o = philox4x32_10(c, seed_lo, seed_hi);
        u = ((double)o.x + 0.5) * 2.3283064365386963e-10;   // (0,1)
      }
      const int acc = log(u) < (double)log_alpha;
      s_acc = acc;
      if (acc) energy[b] = energy_new[b];
      if (accepted) accepted[b] = acc;
    }
  }
  __syncthreads();
  if (s_acc) {
    for (int e = threadIdx.x; e < elems; e += blockDim.x) {
      pos[base + e] = pos_new[base + e];
      force[base + e] = force_new[base + e];
    }
  }
}

// BAOAB, first part (samplers/langevin.py:204-219): B(half kick) A(half drift) O(thermostat) A, wrap.
__global__ void baoab_first_kernel(float* __restrict__ pos, float* __restrict__ vel, const float* __restrict__ force,
                                   const float* __restrict__ noise, int64_t B, int64_t elems, float dt, float inv_mass,
                                   float c1, float c2, float box, uint32_t seed_lo, uint32_t seed_hi,
                                   int64_t cfg_offset, uint32_t step) {
  const int64_t quads = (elems + 3) / 4;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B * quads) return;
  const int64_t b = g / quads, q = g - b * quads;
  float z[4];
  if (!noise) gauss4(seed_lo, seed_hi, (uint32_t)(cfg_offset + b), (uint32_t)q, step, z);
  const float hdt = 0.5f * dt;
  for (int k = 0; k < 4; ++k) {
    const int64_t e = 4 * q + k;
    if (e >= elems) break;
    const int64_t i = b * elems + e;
    float v = __fadd_rn(vel[i], __fmul_rn(__fmul_rn(hdt, force[i]), inv_mass));   // v + 0.5 dt F / m
    float x = __fadd_rn(pos[i], __fmul_rn(hdt, v));
    v = __fadd_rn(__fmul_rn(c1, v), __fmul_rn(c2, noise ? noise[i] : z[k]));
    x = __fadd_rn(x, __fmul_rn(hdt, v));
    pos[i] = wrap_box(x, box);
    vel[i] = v;
  }
}

// BAOAB, last half kick with the forces of the new positions (samplers/langevin.py:221-223)
__global__ void baoab_second_kernel(float* __restrict__ vel, const float* __restrict__ force, int64_t total, float dt,
                                    float inv_mass) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  vel[i] = __fadd_rn(vel[i], __fmul_rn(__fmul_rn(0.5f * dt, force[i]), inv_mass));
}

// g(r) histogram (analysis/rdf.py:49-68): one CTA per configuration, i < j pairs, minimum image,
// np.histogram bin rule (edges in float64; last bin closed on the right) -> exact integer counts.
__global__ void __launch_bounds__(256)
rdf_hist_kernel(const float* __restrict__ pos, int N, float box, const double* __restrict__ edges, int n_bins,
                unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned char smem_rdf[];
  float* sp = reinterpret_cast<float*>(smem_rdf);                        // [N][2]
  unsigned int* sh = reinterpret_cast<unsigned int*>(sp + 2 * N);        // [n_bins]
  const float* p = pos + (int64_t)blockIdx.x * N * 2;
  for (int k = threadIdx.x; k < 2 * N; k += blockDim.x) sp[k] = p[k];
  for (int k = threadIdx.x; k < n_bins; k += blockDim.x) sh[k] = 0u;
  __syncthreads();
  const double lo = edges[0], hi = edges[n_bins];
  const double scale = (double)n_bins / (hi - lo);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float xi = sp[2 * i], yi = sp[2 * i + 1];
    for (int j = i + 1; j < N; ++j) {
      float dx = __fsub_rn(xi, sp[2 * j]), dy = __fsub_rn(yi, sp[2 * j + 1]);
      dx = __fsub_rn(dx, __fmul_rn(box, rintf(__fdiv_rn(dx, box))));
      dy = __fsub_rn(dy, __fmul_rn(box, rintf(__fdiv_rn(dy, box))));
      const double d = (double)__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      if (d < lo || d > hi) continue;
      int idx = (int)((d - lo) * scale);
      if (idx >= n_bins) idx = n_bins - 1;
      if (d < edges[idx]) --idx;
      else if (idx != n_bins - 1 && d >= edges[idx + 1]) ++idx;
      atomicAdd(&sh[idx], 1u);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_bins; k += blockDim.x)
    if (sh[k]) atomicAdd(hist + k, (unsigned long long)sh[k]);
}

}  // namespace
}  // namespace dbs

using namespace dbs;

extern "C" int dbs_mala_propose(const float* pos, const float* force, const float* noise, float* pos_new, int64_t B,
                                int N, float dt, float temperature, float noise_scale, float box, uint64_t seed,
                                int64_t cfg_offset, int64_t step, void* stream) {
  if (B == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(pos && force && pos_new && B >= 0 && N >= 1, "mala_propose: bad arguments");
  DBS_REQUIRE(temperature > 0.f && box > 0.f, "mala_propose: temperature and box must be positive");
  if (B == 0) return DBS_OK;
  const int64_t elems = 2 * (int64_t)N, total = B * ((elems + 3) / 4);
  mala_propose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      pos, force, noise, pos_new, B, elems, dt, temperature, noise_scale, box, (uint32_t)seed, (uint32_t)(seed >> 32),
      cfg_offset, (uint32_t)step);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_mala_accept(float* pos, float* force, float* energy, const float* pos_new, const float* force_new,
                               const float* energy_new, const float* uniforms, int64_t B, int N, float dt,
                               float temperature, uint64_t seed, int64_t cfg_offset, int64_t step, int* accepted,
                               void* stream) {
  if (B == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(pos && force && energy && pos_new && force_new && energy_new && B >= 0 && N >= 1, "mala_accept: bad arguments");
  if (B == 0) return DBS_OK;
  mala_accept_kernel<<<(unsigned)B, 256, 0, as_stream(stream)>>>(pos, force, energy, pos_new, force_new, energy_new,
                                                               uniforms, 2 * N, dt, temperature, (uint32_t)seed,
                                                               (uint32_t)(seed >> 32), cfg_offset, (uint32_t)step,
                                                               accepted);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_baoab_first(float* pos, float* vel, const float* force, const float* noise, int64_t B, int N,
                               float dt, float mass, float c1, float c2, float box, uint64_t seed, int64_t cfg_offset,
                               int64_t step, void* stream) {
  if (B == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(pos && vel && force && B >= 0 && N >= 1 && mass > 0.f && box > 0.f, "baoab_first: bad arguments");
  if (B == 0) return DBS_OK;
  const int64_t elems = 2 * (int64_t)N, total = B * ((elems + 3) / 4);
  baoab_first_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      pos, vel, force, noise, B, elems, dt, 1.0f / mass, c1, c2, box, (uint32_t)seed, (uint32_t)(seed >> 32), cfg_offset,
      (uint32_t)step);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_baoab_second(float* vel, const float* force, int64_t B, int N, float dt, float mass, void* stream) {
  if (B == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(vel && force && B >= 0 && N >= 1 && mass > 0.f, "baoab_second: bad arguments");
  if (B == 0) return DBS_OK;
  const int64_t total = B * 2 * (int64_t)N;
  baoab_second_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(vel, force, total, dt, 1.0f / mass);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_rdf_histogram(const float* pos, int64_t B, int N, float box, const double* edges, int n_bins,
                                 unsigned long long* hist, void* stream) {
  if (B == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(pos && edges && hist && B >= 0 && N >= 2 && n_bins >= 1, "rdf_histogram: bad arguments");
  DBS_REQUIRE(N <= 8192 && n_bins <= 8192, "rdf_histogram: N and n_bins must be <= 8192");
  if (B == 0) return DBS_OK;
  const size_t smem = (size_t)N * 2 * 4 + (size_t)n_bins * 4;
  if (smem > 48 * 1024)
    DBS_CUDA(cudaFuncSetAttribute(rdf_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rdf_hist_kernel<<<(unsigned)B, 256, smem, as_stream(stream)>>>(pos, N, box, edges, n_bins, hist);
  DBS_LAUNCHED();
  return DBS_OK;
}
