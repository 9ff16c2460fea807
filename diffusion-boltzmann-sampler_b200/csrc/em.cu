// em.cu — fused Euler–Maruyama reverse-SDE update and Gaussian fill.
//
// Replaces the inner step of DiffusionSampler.sample (samplers/diffusion.py:232-243; the same
// expression at :279-287, :375-383, :431-437) and the torch.randn start (:216).
// HBM-bound: 12 B per element (x in, score in, x out); noise is generated in-register from
// Philox4x32-10 + Box–Muller unless the caller injects a noise tensor (parity port).
// Evaluation order is the reference's, with explicit _rn intrinsics so nvcc cannot contract
// mul+add into FMA:  (x + (c1*s)*dt) + ((c2*z)*temperature).
#include <math.h>

#include "common.cuh"
#include "philox.cuh"

namespace dbs {
namespace {

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = ((float)(a >> 8) + 0.5f) * 5.9604644775390625e-8f;  // (0,1), 24 bits
  const float u2 = ((float)(b >> 8) + 0.5f) * 5.9604644775390625e-8f;
  const float rad = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

__device__ __forceinline__ void normal4(uint32_t seed_lo, uint32_t seed_hi, uint32_t sample, uint32_t quad,
                                        uint32_t step, float z[4]) {
  u32x4 c;
  c.x = sample; c.y = quad; c.z = step; c.w = DBS_STREAM_NOISE;
  const u32x4 o = philox4x32_10(c, seed_lo, seed_hi);
  box_muller(o.x, o.y, z[0], z[1]);
  box_muller(o.z, o.w, z[2], z[3]);
}

struct EmArgs {
  float* x;
  const float* score;
  const float* noise;
  int64_t n_samples, elems, quads_per_sample;
  float c1, dt, c2, temperature;
  int add_noise;
  uint32_t seed_lo, seed_hi;
  int64_t sample_offset;
  uint32_t step;
  // graph-captured loops (dbs_em_step_dev / dbs_corrector_update): the per-step scalars live in device memory, so one
  // captured launch serves every step.  coef = {c1, dt, c2, temperature, add_noise, step (int32 bits)};
  // seed_off = {seed, sample_offset}.  NULL = the by-value fields above.
  const float* coef;
  const int64_t* seed_off;
  const float* coef2;   // corrector: {step size, sqrt(2 * step size)} replace c1 / c2 (dt = 1); coef = {temperature, active, step}
};

template <bool kVec, bool kFill>
__global__ void __launch_bounds__(256) em_kernel(EmArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.coef2) {          // Langevin corrector step (samplers/diffusion.py:441-447): (x + ss * s) + ((sqrt(2 ss) * z) * T)
    if (a.coef[1] == 0.f) return;      // inactive at the final predictor step
    a.c1 = a.coef2[0]; a.dt = 1.0f; a.c2 = a.coef2[1]; a.temperature = a.coef[0]; a.add_noise = 1;
    a.step = (uint32_t)__float_as_int(a.coef[2]);
  } else if (a.coef) {
    a.c1 = a.coef[0]; a.dt = a.coef[1]; a.c2 = a.coef[2]; a.temperature = a.coef[3]; a.add_noise = a.coef[4] != 0.f;
    a.step = (uint32_t)__float_as_int(a.coef[5]);
  }
  if (a.seed_off) {
    const uint64_t sd = (uint64_t)a.seed_off[0];
    a.seed_lo = (uint32_t)sd; a.seed_hi = (uint32_t)(sd >> 32);
    a.sample_offset = a.seed_off[1];
  }
  const int64_t total = a.n_samples * a.quads_per_sample;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = g / a.quads_per_sample;
    const int64_t q = g - n * a.quads_per_sample;
    const int64_t e0 = n * a.elems + 4 * q;
    const int cnt = (int)((a.elems - 4 * q) < 4 ? (a.elems - 4 * q) : 4);
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    float xv[4], sv[4];
    if (kFill || (a.add_noise && !a.noise))
      normal4(a.seed_lo, a.seed_hi, (uint32_t)(a.sample_offset + n), (uint32_t)q, a.step, z);
    if (kFill) {
      if (kVec) {
        *reinterpret_cast<float4*>(a.x + e0) = make_float4(z[0], z[1], z[2], z[3]);
      } else {
        for (int k = 0; k < cnt; ++k) a.x[e0 + k] = z[k];
      }
      continue;
    }
    if (kVec) {
      const float4 X = *reinterpret_cast<const float4*>(a.x + e0);
      const float4 S = __ldg(reinterpret_cast<const float4*>(a.score + e0));
      xv[0] = X.x; xv[1] = X.y; xv[2] = X.z; xv[3] = X.w;
      sv[0] = S.x; sv[1] = S.y; sv[2] = S.z; sv[3] = S.w;
      if (a.add_noise && a.noise) {
        const float4 Z = __ldg(reinterpret_cast<const float4*>(a.noise + e0));
        z[0] = Z.x; z[1] = Z.y; z[2] = Z.z; z[3] = Z.w;
      }
    } else {
      for (int k = 0; k < 4; ++k) {
        xv[k] = k < cnt ? a.x[e0 + k] : 0.f;
        sv[k] = k < cnt ? a.score[e0 + k] : 0.f;
        if (a.add_noise && a.noise) z[k] = k < cnt ? a.noise[e0 + k] : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = __fadd_rn(xv[k], __fmul_rn(__fmul_rn(a.c1, sv[k]), a.dt));
      if (a.add_noise) v = __fadd_rn(v, __fmul_rn(__fmul_rn(a.c2, z[k]), a.temperature));
      xv[k] = v;
    }
    if (kVec) {
      *reinterpret_cast<float4*>(a.x + e0) = make_float4(xv[0], xv[1], xv[2], xv[3]);
    } else {
      for (int k = 0; k < cnt; ++k) a.x[e0 + k] = xv[k];
    }
  }
}

template <bool kFill>
int launch_em(const EmArgs& a, cudaStream_t st) {
  const int64_t total = a.n_samples * a.quads_per_sample;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = 148LL * 8 * 4;  // persistent-ish grid: multiples of the SM count
  if (blocks > cap) blocks = cap;
  const bool vec = (a.elems % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0) &&
                   (kFill || ((reinterpret_cast<uintptr_t>(a.score) & 15) == 0 &&
                              (!a.noise || (reinterpret_cast<uintptr_t>(a.noise) & 15) == 0)));
  ProfScope prof(kProfEm, st);
  if (vec) DBS_CUDA(launch_chained(em_kernel<true, kFill>, dim3((unsigned)blocks), dim3(256), 0, st, a));
  else DBS_CUDA(launch_chained(em_kernel<false, kFill>, dim3((unsigned)blocks), dim3(256), 0, st, a));
  DBS_LAUNCHED();
  return DBS_OK;
}

}  // namespace
}  // namespace dbs

using namespace dbs;

extern "C" int dbs_em_step(float* x, const float* score, const float* noise, int64_t n_samples,
                           int64_t elems_per_sample, float c1, float dt, float c2, float temperature,
                           int add_noise, uint64_t seed, int64_t sample_offset, int64_t step, void* stream) {
  if (n_samples == 0 || elems_per_sample == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(x && score, "em_step: null x or score");
  DBS_REQUIRE(n_samples >= 0 && elems_per_sample >= 0, "em_step: negative size");
  if (n_samples == 0 || elems_per_sample == 0) return DBS_OK;
  EmArgs a{};
  a.x = x; a.score = score; a.noise = noise; a.n_samples = n_samples; a.elems = elems_per_sample;
  a.quads_per_sample = (elems_per_sample + 3) / 4;
  a.c1 = c1; a.dt = dt; a.c2 = c2; a.temperature = temperature; a.add_noise = add_noise ? 1 : 0;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  a.sample_offset = sample_offset; a.step = (uint32_t)step;
  return launch_em<false>(a, as_stream(stream));
}

extern "C" int dbs_em_step_dev(float* x, const float* score, int64_t n_samples, int64_t elems_per_sample,
                               const float* coef, const int64_t* seed_offset, void* stream) {
  if (n_samples == 0 || elems_per_sample == 0) return DBS_OK;
  DBS_REQUIRE(x && score && coef && seed_offset, "em_step_dev: null argument");
  DBS_REQUIRE(n_samples >= 0 && elems_per_sample >= 0, "em_step_dev: negative size");
  EmArgs a{};
  a.x = x; a.score = score; a.n_samples = n_samples; a.elems = elems_per_sample;
  a.quads_per_sample = (elems_per_sample + 3) / 4;
  a.add_noise = 1;          // decided per step on the device (coef[4])
  a.coef = coef; a.seed_off = seed_offset;
  return launch_em<false>(a, as_stream(stream));
}

extern "C" int dbs_corrector_update(float* x, const float* score, const float* noise, int64_t n_samples,
                                    int64_t elems_per_sample, const float* step_size2, const float* coef,
                                    const int64_t* seed_offset, void* stream) {
  if (n_samples == 0 || elems_per_sample == 0) return DBS_OK;
  DBS_REQUIRE(x && score && step_size2 && coef, "corrector_update: null argument");
  DBS_REQUIRE(noise || seed_offset, "corrector_update: needs a noise tensor or a seed");
  DBS_REQUIRE(n_samples >= 0 && elems_per_sample >= 0, "corrector_update: negative size");
  EmArgs a{};
  a.x = x; a.score = score; a.noise = noise; a.n_samples = n_samples; a.elems = elems_per_sample;
  a.quads_per_sample = (elems_per_sample + 3) / 4;
  a.add_noise = 1;
  a.coef = coef; a.coef2 = step_size2; a.seed_off = seed_offset;
  return launch_em<false>(a, as_stream(stream));
}

extern "C" int dbs_gaussian_fill(float* x, int64_t n_samples, int64_t elems_per_sample, uint64_t seed,
                                 int64_t sample_offset, int64_t step, void* stream) {
  if (n_samples == 0 || elems_per_sample == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(x, "gaussian_fill: null x");
  DBS_REQUIRE(n_samples >= 0 && elems_per_sample >= 0, "gaussian_fill: negative size");
  if (n_samples == 0 || elems_per_sample == 0) return DBS_OK;
  EmArgs a{};
  a.x = x; a.n_samples = n_samples; a.elems = elems_per_sample;
  a.quads_per_sample = (elems_per_sample + 3) / 4;
  a.add_noise = 1;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  a.sample_offset = sample_offset; a.step = (uint32_t)step;
  return launch_em<true>(a, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// out = sum_i a_i * f_i(x_i), up to five terms, f = identity or sign (torch.sign: sign(0) = 0): the state
// combinations of the deterministic integrators (Heun, RK4: samplers/diffusion.py:488-519), the Langevin
// corrector (:441-447) and the heuristic anneal (:762-768) as one fused pass instead of a chain of elementwise ops.
namespace dbs {
namespace {
struct LinCombArgs {
  float* out;
  const float* x[5];
  float a[5];
  int n_terms;
  int sign_first;
  int64_t n;
  const float* a_dev;   // != NULL: coefficients read from device memory (graph-captured loops)
};
__global__ void __launch_bounds__(256) lincomb_kernel(LinCombArgs q) {
  if (q.a_dev) {
#pragma unroll
    for (int t = 0; t < 5; ++t) q.a[t] = t < q.n_terms ? q.a_dev[t] : 0.f;
  }
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= q.n) return;
  if (i4 + 4 <= q.n) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      if (t < q.n_terms) {
        float4 v = *reinterpret_cast<const float4*>(q.x[t] + i4);
        if (t == 0 && q.sign_first) {
          v.x = (v.x > 0.f) - (v.x < 0.f); v.y = (v.y > 0.f) - (v.y < 0.f);
          v.z = (v.z > 0.f) - (v.z < 0.f); v.w = (v.w > 0.f) - (v.w < 0.f);
        }
        acc.x = fmaf(q.a[t], v.x, acc.x); acc.y = fmaf(q.a[t], v.y, acc.y);
        acc.z = fmaf(q.a[t], v.z, acc.z); acc.w = fmaf(q.a[t], v.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(q.out + i4) = acc;
  } else {
    for (int64_t i = i4; i < q.n; ++i) {
      float acc = 0.f;
      for (int t = 0; t < q.n_terms; ++t) {
        float v = q.x[t][i];
        if (t == 0 && q.sign_first) v = (v > 0.f) - (v < 0.f);
        acc = fmaf(q.a[t], v, acc);
      }
      q.out[i] = acc;
    }
  }
}
}  // namespace
}  // namespace dbs

extern "C" int dbs_lincomb(float* out, int64_t n, int n_terms, const float* const* x, const float* a, int sign_first,
                           void* stream) {
  if (n == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(out && x && a && n >= 0 && n_terms >= 1 && n_terms <= 5, "lincomb: bad arguments (1..5 terms)");
  if (n == 0) return DBS_OK;
  dbs::LinCombArgs q;
  q.out = out; q.n = n; q.n_terms = n_terms; q.sign_first = sign_first ? 1 : 0; q.a_dev = nullptr;
  for (int t = 0; t < 5; ++t) { q.x[t] = t < n_terms ? x[t] : nullptr; q.a[t] = t < n_terms ? a[t] : 0.f; }
  for (int t = 0; t < n_terms; ++t) {
    DBS_REQUIRE(x[t] != nullptr, "lincomb: term %d is null", t);
    DBS_REQUIRE((reinterpret_cast<uintptr_t>(x[t]) & 15) == 0, "lincomb: term %d is not 16-byte aligned", t);
  }
  DBS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "lincomb: out is not 16-byte aligned");
  dbs::ProfScope prof(dbs::kProfEm, dbs::as_stream(stream));
  dbs::lincomb_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, dbs::as_stream(stream)>>>(q);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_lincomb_dev(float* out, int64_t n, int n_terms, const float* const* x, const float* coef_dev,
                               int sign_first, void* stream) {
  if (n == 0) return DBS_OK;
  DBS_REQUIRE(out && x && coef_dev && n >= 0 && n_terms >= 1 && n_terms <= 5, "lincomb_dev: bad arguments (1..5 terms)");
  dbs::LinCombArgs q;
  q.out = out; q.n = n; q.n_terms = n_terms; q.sign_first = sign_first ? 1 : 0; q.a_dev = coef_dev;
  for (int t = 0; t < 5; ++t) { q.x[t] = t < n_terms ? x[t] : nullptr; q.a[t] = 0.f; }
  for (int t = 0; t < n_terms; ++t) {
    DBS_REQUIRE(x[t] != nullptr, "lincomb_dev: term %d is null", t);
    DBS_REQUIRE((reinterpret_cast<uintptr_t>(x[t]) & 15) == 0, "lincomb_dev: term %d is not 16-byte aligned", t);
  }
  DBS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "lincomb_dev: out is not 16-byte aligned");
  dbs::ProfScope prof(dbs::kProfEm, dbs::as_stream(stream));
  dbs::lincomb_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, dbs::as_stream(stream)>>>(q);
  DBS_LAUNCHED();
  return DBS_OK;
}

// ------------------------------------------------------------------------------------------------
// Step tables: the scalars of step k of a sampler loop (times, sigma^2 dt, noise scale, Philox step id ...) are row k of
// a device table; ONE captured graph of a step then serves the whole trajectory.  table_next copies the row selected
// by the device counter to a fixed "current row" buffer the step's kernels read, and advances the counter.
namespace dbs {
namespace {
__global__ void table_next_kernel(float* __restrict__ cur, const float* __restrict__ table, int* __restrict__ counter,
                                  int stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int k = *counter;
  for (int i = threadIdx.x; i < stride; i += blockDim.x) cur[i] = table[(int64_t)k * stride + i];
  __syncthreads();
  if (threadIdx.x == 0) *counter = k + 1;
}

// Langevin corrector step size (samplers/diffusion.py:441-445): per-sample ||score||_2, their mean over the batch,
// step = 2 * (snr * sqrt(elems) / mean)^2; out = {step, sqrt(2 * step)}.  Two launches, no host round trip.
__global__ void __launch_bounds__(256) sample_norm_kernel(const float* __restrict__ s, int64_t elems, float* __restrict__ norms) {
  const float* p = s + (int64_t)blockIdx.x * elems;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < elems; i += blockDim.x) acc = fmaf(p[i], p[i], acc);
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    acc = warp_sum(acc);
    if (threadIdx.x == 0) norms[blockIdx.x] = sqrtf(acc);
  }
}
__global__ void __launch_bounds__(256) corrector_coef_kernel(const float* __restrict__ norms, int64_t B, float snr_noise,
                                                             float* __restrict__ out) {
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) acc += (double)norms[i];
  __shared__ double red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float grad_norm = (float)(tot / (double)B);
    // torch evaluates python_float / tensor as tensor.reciprocal() * python_float
    const float r = __fmul_rn(__frcp_rn(grad_norm), snr_noise);
    const float step = __fmul_rn(__fmul_rn(r, r), 2.0f);
    out[0] = step;
    out[1] = sqrtf(__fmul_rn(2.0f, step));
  }
}
}  // namespace
}  // namespace dbs

extern "C" int dbs_step_table_next(float* cur, const float* table, int* counter, int stride, void* stream) {
  DBS_REQUIRE(cur && table && counter && stride >= 1 && stride <= 4096, "step_table_next: bad arguments");
  DBS_CUDA(dbs::launch_chained(dbs::table_next_kernel, dim3(1), dim3(64), 0, dbs::as_stream(stream), cur, table, counter, stride));
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_corrector_coef(const float* score, int64_t B, int64_t elems_per_sample, double snr, float* norms,
                                  float* coef_out, void* stream) {
  if (B == 0 || elems_per_sample == 0) return DBS_OK;
  DBS_REQUIRE(score && norms && coef_out && B > 0 && elems_per_sample > 0, "corrector_coef: bad arguments");
  dbs::ProfScope prof(dbs::kProfEm, dbs::as_stream(stream));
  dbs::sample_norm_kernel<<<(unsigned)B, 256, 0, dbs::as_stream(stream)>>>(score, elems_per_sample, norms);
  DBS_LAUNCHED();
  // corrector_snr * noise_norm is a product of Python floats (doubles), rounded to fp32 when it meets the fp32 tensor
  const float snr_noise = (float)(snr * sqrt((double)elems_per_sample));
  dbs::corrector_coef_kernel<<<1, 256, 0, dbs::as_stream(stream)>>>(norms, B, snr_noise, coef_out);
  DBS_LAUNCHED();
  return DBS_OK;
}
