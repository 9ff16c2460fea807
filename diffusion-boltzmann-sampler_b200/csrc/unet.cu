// unet.cu — score U-Net executor: plan (repacked weights), layer schedule, and the CUDA-core
// kernels (fp32 parity path; also the fallback for shapes the tcgen05 kernel does not take).
//
// Replaces (reference file:line, /root/reference/backend/ml/models/score_network.py):
//   SinusoidalTimeEmbedding.forward :20-41, ScoreNetwork.time_mlp :185-189,235
//   ConvBlock.forward :133-151 (conv3x3 -> GroupNorm(8) -> SiLU -> +time bias -> conv3x3 ->
//                               GroupNorm -> SiLU -> + residual_scale * (conv1x1(x) | x))
//   ScoreNetwork.forward :224-268 (init conv, encoder/downsample, middle, upsample/concat/decoder,
//                                  final conv, optional std-normalise, output_scale)
//
// Activations are NHWC (channels innermost) so that a pixel's channel vector is one contiguous
// run: fp32 in precision 0, bf16 in precision 1.  With in_channels == 1 the API's NCHW input and
// output tensors are already NHWC.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "unet.cuh"

namespace dbs {

// ------------------------------------------------------------------------------------------------
// geometry builders
ConvGeom geom_conv(int B, int H, int W, int C0, int C1, int Cout, int ksize, int stride) {
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  const int pad = ksize / 2;
  g.B = B; g.Hin = H; g.Win = W;
  g.Hout = (H + 2 * pad - ksize) / stride + 1;
  g.Wout = (W + 2 * pad - ksize) / stride + 1;
  g.Hm = g.Hout; g.Wm = g.Wout;
  g.C0 = C0; g.C1 = C1; g.Cout = Cout;
  g.in_stride = stride; g.out_stride = 1; g.oph = 0; g.opw = 0;
  g.n_taps = ksize * ksize;
  for (int kh = 0; kh < ksize; ++kh)
    for (int kw = 0; kw < ksize; ++kw) {
      const int t = kh * ksize + kw;
      g.dh[t] = (signed char)(kh - pad);
      g.dw[t] = (signed char)(kw - pad);
      g.widx[t] = (unsigned char)t;
    }
  return g;
}

// ConvTranspose2d(k=4, s=2, p=1): out[2a+ph] gathers in[a + dh] with kh = ph + 1 - 2*dh, i.e.
// ph=0: (dh,kh) in {(0,1), (-1,3)};  ph=1: (dh,kh) in {(+1,0), (0,2)}.
ConvGeom geom_convT_phase(int B, int H, int W, int Cin, int Cout, int ph, int pw) {
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.Hin = H; g.Win = W; g.Hm = H; g.Wm = W; g.Hout = 2 * H; g.Wout = 2 * W;
  g.C0 = Cin; g.C1 = 0; g.Cout = Cout;
  g.in_stride = 1; g.out_stride = 2; g.oph = ph; g.opw = pw;
  g.n_taps = 4;
  const int dhs[2][2] = {{0, -1}, {1, 0}};
  const int khs[2][2] = {{1, 3}, {0, 2}};
  int t = 0;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j, ++t) {
      g.dh[t] = (signed char)dhs[ph][i];
      g.dw[t] = (signed char)dhs[pw][j];
      g.widx[t] = (unsigned char)(khs[ph][i] * 4 + khs[pw][j]);
    }
  return g;
}

// All four phases in one geometry (see ConvGeom::n_phases).
ConvGeom geom_convT_all(int B, int H, int W, int Cin, int Cout) {
  ConvGeom g = geom_convT_phase(B, H, W, Cin, Cout, 0, 0);
  g.n_phases = 4;
  for (int ph4 = 1; ph4 < 4; ++ph4) {
    const ConvGeom q = geom_convT_phase(B, H, W, Cin, Cout, ph4 >> 1, ph4 & 1);
    for (int t = 0; t < 4; ++t) { g.dh[4 * ph4 + t] = q.dh[t]; g.dw[4 * ph4 + t] = q.dw[t]; g.widx[4 * ph4 + t] = q.widx[t]; }
  }
  return g;
}

// 3x3 stride-2 pad-1 conv as a 2x2 stride-1 conv on the space-to-depth view (see ConvGeom::s2d).
ConvGeom geom_conv_s2d(int B, int H, int W, int C, int Cout) {
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.Hin = H / 2; g.Win = W / 2; g.Hm = H / 2; g.Wm = W / 2; g.Hout = H / 2; g.Wout = W / 2;
  g.C0 = 4 * C; g.C1 = 0; g.Cout = Cout;
  g.in_stride = 1; g.out_stride = 1; g.s2d = 1;
  g.n_taps = 4;
  int t = 0;
  for (int dh = -1; dh <= 0; ++dh)
    for (int dw = -1; dw <= 0; ++dw, ++t) { g.dh[t] = (signed char)dh; g.dw[t] = (signed char)dw; g.widx[t] = (unsigned char)t; }
  return g;
}

// The two horizontal phases of vertical phase ph in one geometry (see ConvGeom::n_phases).
ConvGeom geom_convT_pair(int B, int H, int W, int Cin, int Cout, int ph) {
  ConvGeom g = geom_convT_phase(B, H, W, Cin, Cout, ph, 0);
  const ConvGeom q = geom_convT_phase(B, H, W, Cin, Cout, ph, 1);
  g.n_phases = 2;
  for (int t = 0; t < 4; ++t) { g.dh[4 + t] = q.dh[t]; g.dw[4 + t] = q.dw[t]; g.widx[4 + t] = q.widx[t]; }
  return g;
}

namespace {

// ------------------------------------------------------------------------------------------------
// element helpers
template <typename T> struct Quad;
template <> struct Quad<float> {
  static __device__ __forceinline__ void load(const float* p, float v[4]) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void store(float* p, const float v[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Quad<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float v[4]) {
    const uint2 q = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float v[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    const __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 q;
    q.x = *reinterpret_cast<const uint32_t*>(&a);
    q.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = q;
  }
};
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
// bf16 path: x*sigmoid(x) with sigmoid = 0.5*tanh(x/2)+0.5, one MUFU.TANH per element (rel. error
// ~2^-11, below bf16 resolution).  The accurate form above costs two MUFU ops + an IEEE division.
__device__ __forceinline__ float silu_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}
template <typename T> __device__ __forceinline__ float silu_t(float x);
template <> __device__ __forceinline__ float silu_t<float>(float x) { return silu(x); }
template <> __device__ __forceinline__ float silu_t<__nv_bfloat16>(float x) { return silu_fast(x); }

// ------------------------------------------------------------------------------------------------
// generic implicit-GEMM convolution on the CUDA cores: 64 pixels x 64 couts per CTA, K chunk 16
constexpr int kGT = 64, kGK = 16, kGPad = 68;

template <typename T, typename TW>
__global__ void __launch_bounds__(256)
conv_generic_kernel(ConvGeom g, const T* __restrict__ in0, const T* __restrict__ in1,
                    const TW* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out) {
  __shared__ __align__(16) float As[kGK][kGPad];
  __shared__ __align__(16) float Bs[kGK][kGPad];
  const int tid = threadIdx.x;
  const int Cin = g.C0 + g.C1;
  const int64_t M = (int64_t)g.B * g.Hm * g.Wm;
  const int64_t m0 = (int64_t)blockIdx.x * kGT;
  const int co0 = blockIdx.y * kGT;

  // loader role: one pixel (resp. one cout) and one channel quad
  const int lp = tid >> 2, lc = (tid & 3) * 4;
  const int64_t lm = m0 + lp;
  const bool lvalid = lm < M;
  int ln = 0, la = 0, lb = 0;
  if (lvalid) {
    ln = (int)(lm / ((int64_t)g.Hm * g.Wm));
    const int rem = (int)(lm - (int64_t)ln * g.Hm * g.Wm);
    la = rem / g.Wm;
    lb = rem - la * g.Wm;
  }
  const int lco = co0 + lp;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < g.n_taps; ++t) {
    const int ih = la * g.in_stride + g.dh[t];
    const int iw = lb * g.in_stride + g.dw[t];
    const bool pix_ok = lvalid && ih >= 0 && ih < g.Hin && iw >= 0 && iw < g.Win;
    const int64_t pix = ((int64_t)ln * g.Hin + ih) * g.Win + iw;
    const TW* wt = w + (int64_t)g.widx[t] * g.Cout * Cin;
    for (int c0 = 0; c0 < Cin; c0 += kGK) {
      const int c = c0 + lc;
      float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (pix_ok && c < Cin) {
        if (c < g.C0) Quad<T>::load(in0 + pix * g.C0 + c, av);
        else Quad<T>::load(in1 + pix * g.C1 + (c - g.C0), av);
      }
      if (lco < g.Cout && c < Cin) Quad<TW>::load(wt + (int64_t)lco * Cin + c, bv);
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        As[lc + q][lp] = av[q];
        Bs[lc + q][lp] = bv[q];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kGK; ++kk) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
  const int co = co0 + tx * 4;
  if (co >= g.Cout) return;
  float bq[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bq[j] = bias ? bias[co + j] : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int n = (int)(m / ((int64_t)g.Hm * g.Wm));
    const int rem = (int)(m - (int64_t)n * g.Hm * g.Wm);
    const int a = rem / g.Wm, b = rem - a * g.Wm;
    const int oh = a * g.out_stride + g.oph, ow = b * g.out_stride + g.opw;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bq[j];
    Quad<T>::store(out + (((int64_t)n * g.Hout + oh) * g.Wout + ow) * g.Cout + co, v);
  }
}

// 16-byte vectors of the activation type
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float v[4]) { Quad<float>::load(p, v); }
  static __device__ __forceinline__ void store(float* p, const float v[4]) { Quad<float>::store(p, v); }
  // the 16 bytes as loaded (kept packed while a load is in flight) and their conversion
  static __device__ __forceinline__ uint4 load_raw(const float* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const uint4& q, float v[4]) {
    v[0] = __uint_as_float(q.x); v[1] = __uint_as_float(q.y); v[2] = __uint_as_float(q.z); v[3] = __uint_as_float(q.w);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float v[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      v[2 * i] = __low2float(h); v[2 * i + 1] = __high2float(h);
    }
  }
  static __device__ __forceinline__ uint4 load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const uint4& q, float v[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      v[2 * i] = __low2float(h); v[2 * i + 1] = __high2float(h);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float v[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// init conv: 1 -> C, 3x3 pad 1 (score_network.py:192, 238).  x fp32 [B][H][W]; w [9][C].
// C/V threads per pixel, one 16-byte channel vector each: a warp's store covers contiguous memory.
template <typename T>
__global__ void __launch_bounds__(256)
init_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 T* __restrict__ out, int64_t n_pix, int H, int W, int C) {
  constexpr int V = Vec16<T>::N;
  const int cv = C / V;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_pix * cv) return;
  const int64_t p = g / cv;
  const int c = (int)(g - p * cv) * V;
  const int ww = (int)(p % W);
  const int hh = (int)((p / W) % H);
  float v[V];
#pragma unroll
  for (int k = 0; k < V; ++k) v[k] = __ldg(bias + c + k);
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ih = hh + kh - 1, iw = ww + kw - 1;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
      const float xv = __ldg(x + p + (int64_t)(kh - 1) * W + (kw - 1));
      const float* wk = w + (kh * 3 + kw) * C + c;
#pragma unroll
      for (int k = 0; k < V; k += 4) {
        const float4 wq = __ldg(reinterpret_cast<const float4*>(wk + k));
        v[k] = fmaf(xv, wq.x, v[k]); v[k + 1] = fmaf(xv, wq.y, v[k + 1]);
        v[k + 2] = fmaf(xv, wq.z, v[k + 2]); v[k + 3] = fmaf(xv, wq.w, v[k + 3]);
      }
    }
  Vec16<T>::store(out + p * C + c, v);
}

// final conv: C -> 1, 3x3 pad 1, then * output_scale (score_network.py:219, 259, 268).
// C/V threads per pixel (a warp's load of one tap is contiguous), partial dot products combined
// with shuffles inside the pixel's thread group.  Requires C/V to be a power of two <= 32.
template <typename T>
__global__ void __launch_bounds__(256)
final_conv_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, int64_t n_pix, int H, int W, int C, float scale) {
  constexpr int V = Vec16<T>::N;
  const int cv = C / V;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p = g / cv;
  const int c = (int)(g - p * cv) * V;
  const bool live = p < n_pix;
  const int ww = (int)(p % W);
  const int hh = (int)((p / W) % H);
  float acc = 0.f;
  if (live) {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ih = hh + kh - 1, iw = ww + kw - 1;
        if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
        float v[V];
        Vec16<T>::load(in + (p + (int64_t)(kh - 1) * W + (kw - 1)) * C + c, v);
        const float* wk = w + (kh * 3 + kw) * C + c;
#pragma unroll
        for (int k = 0; k < V; k += 4) {
          const float4 wq = __ldg(reinterpret_cast<const float4*>(wk + k));
          acc = fmaf(v[k], wq.x, acc); acc = fmaf(v[k + 1], wq.y, acc);
          acc = fmaf(v[k + 2], wq.z, acc); acc = fmaf(v[k + 3], wq.w, acc);
        }
      }
  }
  for (int o = cv >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (live && c == 0) out[p] = scale * (acc + bias[0]);
}

// the same conv for channel counts whose vector count is not a power of two (base_channels 24, 48, ...: checkpoints the
// reference's constructor accepts, not BASELINE configurations): one thread per pixel, plain loops
template <typename T>
__global__ void __launch_bounds__(256)
final_conv_generic_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                          float* __restrict__ out, int64_t n_pix, int H, int W, int C, float scale) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  const int ww = (int)(p % W);
  const int hh = (int)((p / W) % H);
  float acc = 0.f;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const int ih = hh + kh - 1, iw = ww + kw - 1;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
      const T* src = in + (p + (int64_t)(kh - 1) * W + (kw - 1)) * C;
      const float* wk = w + (kh * 3 + kw) * C;
      for (int c = 0; c < C; ++c) acc = fmaf(to_f<T>(src[c]), __ldg(wk + c), acc);
    }
  out[p] = scale * (acc + bias[0]);
}

// Init conv 1 -> C on the bf16 path: one thread per pixel produces all C channels.  The 9 x C weights and
// the bias arrive as a __grid_constant__ kernel parameter, so every FFMA reads its weight from the constant
// bank (warp-uniform index) and the only memory instructions are 9 coalesced x loads and C/8 16-byte stores.
template <int C> struct InitW { float w[9 * C]; float b[C]; };

// time path (score_network.py:20-41, 185-189, 126/146) as a device function: the small embedding MLP is recomputed by
// each of the n_parts CTAs working on time value t_idx, which then split the concatenated per-block Linear layers
// (sumC outputs) between them.  One thread per output row, float4 weight loads issued back to back: three
// dependent stages of ~L2 latency each.  sm: 4*T floats.
struct TimeArgs {
  const float *t, *freqs, *w0, *b0, *w2, *b2, *wt, *bt;
  float* tbias;
  int T, sumC;
  int n_ctas;   // init_conv_fast_kernel: leading CTAs that run the time path instead of pixels (0 = none)
};
__device__ __forceinline__ float dot_row(const float* __restrict__ wrow, const float* __restrict__ v, int n) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k = 0; k < n; k += 4) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(wrow + k));
    a0 = fmaf(q.x, v[k], a0); a1 = fmaf(q.y, v[k + 1], a1);
    a2 = fmaf(q.z, v[k + 2], a2); a3 = fmaf(q.w, v[k + 3], a3);
  }
  return (a0 + a1) + (a2 + a3);
}
__device__ __forceinline__ void time_path(const TimeArgs& a, int t_idx, int part, int n_parts, float* sm) {
  const int T = a.T;
  float* emb = sm;           // [T]
  float* hid = sm + T;       // [2T]
  float* temb = sm + 3 * T;  // [T]
  const int half = T / 2;
  const float tv = a.t[t_idx];
  for (int k = threadIdx.x; k < T; k += blockDim.x) {
    const int f = k < half ? k : k - half;
    const float arg = __fmul_rn(__fmul_rn(tv, a.freqs[f]), 1000.0f);
    emb[k] = k < half ? sinf(arg) : cosf(arg);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * T; j += blockDim.x) hid[j] = silu(dot_row(a.w0 + (int64_t)j * T, emb, T) + a.b0[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) temb[i] = dot_row(a.w2 + (int64_t)i * 2 * T, hid, 2 * T) + a.b2[i];
  __syncthreads();
  const int per = (a.sumC + n_parts - 1) / n_parts;
  const int c_lo = part * per, c_hi = min(a.sumC, c_lo + per);
  for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x)
    a.tbias[(int64_t)t_idx * a.sumC + c] = dot_row(a.wt + (int64_t)c * T, temb, T) + a.bt[c];
}

template <int C>
__global__ void __launch_bounds__(256)
init_conv_fast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n_pix, int H, int W,
                      const __grid_constant__ InitW<C> iw, const __grid_constant__ TimeArgs ta) {
  // horizontal fusion: when all samples share t (the sampler's case) the first ta.n_ctas CTAs compute the time
  // biases of every ConvBlock while the others run the init conv — the time MLP's ~20 us of dependent L2 round
  // trips disappear behind the conv instead of preceding it as a launch of their own
  pdl_launch_dependents();
  pdl_wait();
  if ((int)blockIdx.x < ta.n_ctas) {
    extern __shared__ float sm_time[];
    time_path(ta, 0, blockIdx.x, ta.n_ctas, sm_time);
    return;
  }
  const int64_t p = (int64_t)(blockIdx.x - ta.n_ctas) * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  const int ww = (int)(p % W);
  const int hh = (int)((p / W) % H);
  float xv[9];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ih = hh + kh - 1, iw_ = ww + kw - 1;
      const bool ok = ih >= 0 && ih < H && iw_ >= 0 && iw_ < W;
      xv[kh * 3 + kw] = ok ? __ldg(x + p + (int64_t)(kh - 1) * W + (kw - 1)) : 0.f;
    }
#pragma unroll
  for (int c0 = 0; c0 < C; c0 += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = iw.b[c0 + k];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaf(xv[t], iw.w[t * C + c0 + k], v[k]);
    Vec16<__nv_bfloat16>::store(out + p * C + c0, v);
  }
}

template <int C>
int launch_init_fast(const std::vector<float>& w_host, const std::vector<float>& b_host, const float* x,
                     __nv_bfloat16* out, int64_t n_pix, int H, int W, cudaStream_t st, const TimeArgs& ta) {
  InitW<C> iw;
  for (int i = 0; i < 9 * C; ++i) iw.w[i] = w_host[i];
  for (int i = 0; i < C; ++i) iw.b[i] = b_host[i];
  DBS_CUDA(launch_chained(init_conv_fast_kernel<C>, dim3((unsigned)((n_pix + 255) / 256) + (unsigned)ta.n_ctas), dim3(256),
                          ta.n_ctas ? (size_t)4 * ta.T * 4 : 0, st, x, out, n_pix, H, W, iw, ta));
  DBS_LAUNCHED();
  return DBS_OK;
}

// Final conv C -> 1 (3x3, pad 1) on the bf16 path, optionally fused with the GroupNorm apply that produces its
// input (the last decoder block's  SiLU(GN(conv2)) + residual, score_network.py:148-151, 259, 268).
// A CTA owns TH full-width rows of one image.
//   phase 1: the rows (+ one halo row each side) are brought into shared memory once with coalesced 16-byte loads —
//            transformed on the way when fused, so the normalised tensor never exists in HBM (bf16, C*2 bytes per pixel,
//            the 16-byte chunk index XOR-swizzled with the pixel index: conflict-free for the loads below);
//   phase 2: the nine per-tap dot products of every strip pixel, D[pixel][tap] = sum_c a[pixel][c] * w[tap][c], are a
//            [pixels x C] x [C x 9] product: warps run it as mma.sync m16n8k16 bf16 tiles (16 pixels x 2 x 8 taps, operands
//            by ldmatrix, the weight fragments built once per thread from the __grid_constant__ parameter) — 288 MACs per
//            pixel are ~1.3 warp instructions instead of ~600 thread instructions of unpack + FFMA.  (tcgen05 needs
//            M = 128 x N >= 16 tiles fed from UMMA-layout operands; for a 9-column product the warp-level MMA is the fit.)
//   phase 3: out[p] = scale * (bias + sum_taps D[p + offset(tap)][tap]) — a nine-point stencil on fp32 scalars.
// bfrag: the B-operand fragments of the weights, [C/16 k-steps][2 n-tiles][2 halves][32 lanes] packed bf16 pairs, built once
// per plan (a lane-dependent index into the constant bank would serialise 32 ways)
struct FinalW { const uint32_t* bfrag; float bias, scale; };
constexpr int kFinalDStride = 10;   // floats per pixel of the D table (9 taps, padded to an even count for 8-byte stores)

template <int C, bool kFuseGn>
__global__ void __launch_bounds__(256, 3)
final_fused_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ res,
                   const float* __restrict__ stats, float inv_cnt, float eps, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ out, int H, int W, int TH,
                   const __grid_constant__ FinalW fw) {
  constexpr int NCK = C / 8;                 // 16-byte chunks per pixel
  constexpr int PPL = NCK >= 8 ? 1 : 8 / NCK;  // pixels per 128 bytes
  constexpr int KS = C / 16;                 // k-steps of the m16n8k16 MMA
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __shared__ float s_sc[C], s_sh[C];
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y, h0 = blockIdx.x * TH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wshift = 31 - __clz(W);          // W is a power of two (launch_final_fused): no integer division per chunk / pixel
  const int n_pix = (TH + 2) * W;
  const int n_pix16 = (n_pix + 15) & ~15;    // the MMA walks whole 16-pixel groups (the tail reads unused smem)
  const __nv_bfloat16* const yimg = y + (int64_t)n * H * W * C;      // 32-bit offsets inside the image from here on
  const __nv_bfloat16* const rimg = kFuseGn ? res + (int64_t)n * H * W * C : nullptr;
  float* s_d = reinterpret_cast<float*>(sm_raw + (size_t)n_pix16 * (C * 2));
  if (kFuseGn) {
    if (tid < C) {
      const int gi = tid / (C / 8);
      const float s0 = stats[((int64_t)n * 8 + gi) * 2 + 0], s1 = stats[((int64_t)n * 8 + gi) * 2 + 1];
      const float mean = s0 * inv_cnt;
      const float rstd = rsqrtf(fmaxf(s1 * inv_cnt - mean * mean, 0.f) + eps);
      const float a = rstd * gamma[tid];
      s_sc[tid] = a;
      s_sh[tid] = beta[tid] - mean * a;
    }
    __syncthreads();
  }
  // ---- phase 1: kU 16-byte chunks per thread in flight (all loads of a batch are issued before the first is used: with one
  // chunk per iteration the ten dependent global round trips of a thread were the kernel's critical path, 217 us at
  // B = 1024, 64x64 for 537 MB + 25 % halo)
  const int n_vec = n_pix * NCK;
  constexpr int kU = 5;
  for (int base = tid; base < n_vec; base += (int)blockDim.x * kU) {
    uint4 raw[kU], rq[kU];
    int pis[kU], cks[kU];
    bool ok[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int idx = base + u * (int)blockDim.x;
      const int ck = idx % NCK, pi = idx / NCK;
      const int r = pi >> wshift, wl = pi & (W - 1), ih = h0 + r - 1;
      ok[u] = idx < n_vec && ih >= 0 && ih < H;   // rows outside the image are never used: phase 3 applies the same bounds
      pis[u] = pi; cks[u] = ck;
      if (ok[u]) {
        const int g = ((ih << wshift) + wl) * C + ck * 8;
        raw[u] = *reinterpret_cast<const uint4*>(yimg + g);
        if (kFuseGn) rq[u] = *reinterpret_cast<const uint4*>(rimg + g);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (!ok[u]) continue;
      const int ck = cks[u], pi = pis[u];
      uint4 o = raw[u];
      if (kFuseGn) {
        float v[8], rr[8];
        Vec16<__nv_bfloat16>::unpack(raw[u], v);
        Vec16<__nv_bfloat16>::unpack(rq[u], rr);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf(1.0f, rr[k], silu_fast(fmaf(v[k], s_sc[ck * 8 + k], s_sh[ck * 8 + k])));
        uint32_t w4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          w4[i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        o = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
      const int swz = (pi / PPL) & (NCK - 1);
      *reinterpret_cast<uint4*>(sm_raw + (size_t)pi * (C * 2) + ((ck ^ swz) << 4)) = o;
    }
  }
  // weight fragments of the B operand (K x N = channels x taps, "col" layout): b0 = (k = 2t, 2t+1; n = g), b1 = (k + 8)
  const int fg = lane >> 2, ft = lane & 3;
  uint32_t bfrag[KS][2][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) bfrag[ks][j][hh] = __ldg(fw.bfrag + ((ks * 2 + j) * 2 + hh) * 32 + lane);
  __syncthreads();
  // ---- phase 2: D[pixel][tap] for every strip pixel
  const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm_raw);
  const int nwarps = blockDim.x >> 5;
  for (int grp = warp; grp < n_pix16 / 16; grp += nwarps) {
    // ldmatrix.x4: lane i addresses row (i % 8) of matrix i / 8; matrices = (pixels 0-7 | 8-15) x (k 0-7 | 8-15)
    const int mid = lane >> 3;
    const int pi = grp * 16 + (lane & 7) + 8 * (mid & 1);
    const int swz = (pi / PPL) & (NCK - 1);
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int ck = 2 * ks + (mid >> 1);
      const uint32_t addr = sm_base + (uint32_t)pi * (C * 2) + (uint32_t)((ck ^ swz) << 4);
      uint32_t a0, a1, a2, a3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                   : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
#pragma unroll
      for (int j = 0; j < 2; ++j)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                     "{%0, %1, %2, %3};"
                     : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bfrag[ks][j][0]), "r"(bfrag[ks][j][1]));
    }
    // accumulator fragment: (row g, cols 2t, 2t+1), (row g + 8, same cols); col = tap inside the n-tile
    float* d0 = s_d + (size_t)(grp * 16 + fg) * kFinalDStride;
    float* d1 = d0 + 8 * kFinalDStride;
    *reinterpret_cast<float2*>(d0 + 2 * ft) = make_float2(acc[0][0], acc[0][1]);
    *reinterpret_cast<float2*>(d1 + 2 * ft) = make_float2(acc[0][2], acc[0][3]);
    if (ft == 0) { d0[8] = acc[1][0]; d1[8] = acc[1][2]; }
  }
  __syncthreads();
  // ---- phase 3: nine-point stencil on the scalar tables
  for (int o = tid; o < TH * W; o += blockDim.x) {
    const int hl = o >> wshift, wl = o & (W - 1);
    if (h0 + hl >= H) break;
    float acc = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = h0 + hl + kh - 1;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = wl + kw - 1;
        if (iw < 0 || iw >= W) continue;
        acc += s_d[(((hl + kh) << wshift) + iw) * kFinalDStride + kh * 3 + kw];
      }
    }
    out[((int64_t)n * H + h0 + hl) * W + wl] = fw.scale * (acc + fw.bias);
  }
}

// host-side build of FinalW::bfrag from the [tap][C] fp32 weights (round-to-nearest-even bf16)
static uint32_t bf16_bits(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0u;
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
std::vector<uint32_t> final_bfrag_host(const std::vector<float>& w, int C) {
  std::vector<uint32_t> t((size_t)(C / 16) * 2 * 2 * 32);
  for (int ks = 0; ks < C / 16; ++ks)
    for (int j = 0; j < 2; ++j)
      for (int hh = 0; hh < 2; ++hh)
        for (int lane = 0; lane < 32; ++lane) {
          const int tap = 8 * j + (lane >> 2), c = 16 * ks + 2 * (lane & 3) + 8 * hh;
          const float w0 = tap < 9 ? w[(size_t)tap * C + c] : 0.f, w1 = tap < 9 ? w[(size_t)tap * C + c + 1] : 0.f;
          t[(size_t)(((ks * 2 + j) * 2 + hh) * 32 + lane)] = bf16_bits(w0) | (bf16_bits(w1) << 16);
        }
  return t;
}

template <int C>
int launch_final_fused(const uint32_t* bfrag_dev, float bias, float scale, const __nv_bfloat16* y,
                       const __nv_bfloat16* res, const float* stats, const float* gamma, const float* beta, float* out,
                       int64_t B, int H, int W, cudaStream_t st) {
  FinalW fw;
  fw.bfrag = bfrag_dev; fw.bias = bias; fw.scale = scale;
  // rows per CTA: up to 512 output pixels while the strip + its D table stay within 72 KB (3 CTAs per SM)
  int TH = 512 / W;
  if (TH > H) TH = H;
  auto bytes = [&](int th) { return (size_t)((((th + 2) * W + 15) & ~15)) * (C * 2 + kFinalDStride * 4); };
  while (TH > 1 && bytes(TH) > 72 * 1024) TH >>= 1;
  const size_t smem = bytes(TH);
  const dim3 grid((unsigned)((H + TH - 1) / TH), (unsigned)B);
  const float inv_cnt = 1.0f / (float)((int64_t)H * W * (C / 8));
  static bool attr_done = false;
  if (!attr_done) {
    DBS_CUDA(cudaFuncSetAttribute(final_fused_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    DBS_CUDA(cudaFuncSetAttribute(final_fused_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    attr_done = true;
  }
  if (res) DBS_CUDA(launch_chained(final_fused_kernel<C, true>, grid, dim3(256), smem, st, y, res, stats, inv_cnt, 1e-5f, gamma, beta, out, H, W, TH, fw));
  else DBS_CUDA(launch_chained(final_fused_kernel<C, false>, grid, dim3(256), smem, st, y, (const __nv_bfloat16*)nullptr, (const float*)nullptr, 0.f, 0.f,
                               (const float*)nullptr, (const float*)nullptr, out, H, W, TH, fw));
  DBS_LAUNCHED();
  return DBS_OK;
}

// optional per-sample std normalisation (score_network.py:262-265): score / (std + 1e-8),
// unbiased std over H*W, applied before output_scale (so undo/redo the scale here).
__global__ void normalize_output_kernel(float* __restrict__ score, int HW, float scale) {
  float* s = score + (int64_t)blockIdx.x * HW;
  __shared__ double red[2][32];
  double sum = 0.0, sq = 0.0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const double v = (double)s[i] / (double)scale;
    sum += v; sq += v * v;
  }
  sum = warp_sum(sum); sq = warp_sum(sq);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = sum; red[1][wid] = sq; }
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    sum = lane < nw ? red[0][lane] : 0.0;
    sq = lane < nw ? red[1][lane] : 0.0;
    sum = warp_sum(sum); sq = warp_sum(sq);
    if (lane == 0) {
      const double mean = sum / HW;
      double var = (sq - HW * mean * mean) / (HW > 1 ? HW - 1 : 1);
      if (var < 0) var = 0;
      red[0][0] = sqrt(var) + 1e-8;
    }
  }
  __syncthreads();
  const float inv = (float)(1.0 / red[0][0]);
  for (int i = threadIdx.x; i < HW; i += blockDim.x) s[i] = scale * ((s[i] / scale) * inv);
}

// GroupNorm statistics, two passes, fp64 accumulation: one CTA per (sample, group) -> (mean, rstd)
template <typename T>
__global__ void __launch_bounds__(256)
gn_stats_kernel(const T* __restrict__ y, float* __restrict__ stats, int HW, int C, int G, float eps) {
  const int n = blockIdx.x / G, gi = blockIdx.x - n * G;
  const int Cg = C / G;
  const T* base = y + (int64_t)n * HW * C + gi * Cg;
  const int64_t cnt = (int64_t)HW * Cg;
  __shared__ double red[32];
  __shared__ double s_mean;
  double sum = 0.0;
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t p = i / Cg;
    const int c = (int)(i - p * Cg);
    sum += (double)to_f<T>(base[p * C + c]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  if (wid == 0) {
    sum = lane < nw ? red[lane] : 0.0;
    sum = warp_sum(sum);
    if (lane == 0) s_mean = sum / (double)cnt;
  }
  __syncthreads();
  const double mean = s_mean;
  double sq = 0.0;
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t p = i / Cg;
    const int c = (int)(i - p * Cg);
    const double d = (double)to_f<T>(base[p * C + c]) - mean;
    sq += d * d;
  }
  sq = warp_sum(sq);
  __syncthreads();
  if (lane == 0) red[wid] = sq;
  __syncthreads();
  if (wid == 0) {
    sq = lane < nw ? red[lane] : 0.0;
    sq = warp_sum(sq);
    if (lane == 0) {
      const float var = (float)(sq / (double)cnt);
      stats[(int64_t)blockIdx.x * 2 + 0] = (float)mean;
      stats[(int64_t)blockIdx.x * 2 + 1] = 1.0f / sqrtf(var + eps);
    }
  }
}

// normalise * gamma + beta -> SiLU -> (+ time bias) -> (+ residual_scale * residual)
// stats: (mean, rstd) per (n, g), or raw (sum, sumsq) from the tcgen05 epilogue when raw != 0.
// Grid = (chunks, B): a CTA works on one sample, folds the statistics into per-channel
// scale/shift once (shared memory), then streams 16-byte vectors: HBM-bound, 2 B read + 2 B
// written per element in bf16 (+2 B for the residual).
template <typename T, int UN>
__global__ void __launch_bounds__(256, 4)
gn_apply_kernel(const T* __restrict__ y, T* __restrict__ out, const float* __restrict__ stats, int raw,
                float inv_cnt, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ tbias, int t_stride, const T* __restrict__ res, float res_scale,
                int HW, int C, int G) {
  constexpr int V = Vec16<T>::N;
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y;
  const int Cg = C / G;
  const int cv = C / V;                       // 16-byte vectors per pixel (divides blockDim.x, see launch)
  // a thread keeps one channel vector for its whole life: scale/shift/time-bias stay in registers
  const int c = (threadIdx.x % cv) * V;
  const int prow = threadIdx.x / cv;          // pixel offset of this thread inside a block-row
  const int ppb = blockDim.x / cv;            // pixels per block iteration
  const T* yb = y + (int64_t)n * HW * C + c;
  const T* rb = res ? res + (int64_t)n * HW * C + c : nullptr;
  T* ob = out + (int64_t)n * HW * C + c;
  const int stride = gridDim.x * ppb;
  int p0 = blockIdx.x * ppb + prow;
  // The first batch of data loads is issued BEFORE the statistics -> scale/shift prologue: the prologue is a chain of
  // dependent L2 round trips (stats, then gamma / beta), and an in-order warp would otherwise sit through it with no
  // data load in flight (a CTA lives for only a few microseconds, so that latency was ~15 % of its life).
  uint4 qv[UN], qr[UN];                       // kept packed while in flight (registers: 4 per vector, not V)
  const bool first_full = p0 + (UN - 1) * stride < HW;
  if (first_full) {
#pragma unroll
    for (int u = 0; u < UN; ++u) qv[u] = Vec16<T>::load_raw(yb + (int64_t)(p0 + u * stride) * C);
    if (rb) {
#pragma unroll
      for (int u = 0; u < UN; ++u) qr[u] = Vec16<T>::load_raw(rb + (int64_t)(p0 + u * stride) * C);
    }
  }
  float sc[V], sh[V], tb[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int gi = (c + k) / Cg;
    const float s0 = stats[((int64_t)n * G + gi) * 2 + 0];
    const float s1 = stats[((int64_t)n * G + gi) * 2 + 1];
    float mean, rstd;
    if (raw) {
      mean = s0 * inv_cnt;
      rstd = rsqrtf(fmaxf(s1 * inv_cnt - mean * mean, 0.f) + eps);
    } else {
      mean = s0; rstd = s1;
    }
    const float a = rstd * gamma[c + k];
    sc[k] = a;
    sh[k] = beta[c + k] - mean * a;
    tb[k] = tbias ? tbias[(int64_t)n * t_stride + c + k] : 0.f;
  }
  // UN pixels per iteration: UN (2*UN with a residual) independent 16-byte loads in flight per thread
  if (first_full) {
    while (true) {
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        float v[V], r[V];
        Vec16<T>::unpack(qv[u], v);
        if (rb) Vec16<T>::unpack(qr[u], r);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float h = silu_t<T>(fmaf(v[k], sc[k], sh[k])) + tb[k];
          if (rb) h = fmaf(res_scale, r[k], h);
          v[k] = h;
        }
        Vec16<T>::store(ob + (int64_t)(p0 + u * stride) * C, v);
      }
      p0 += UN * stride;
      if (!(p0 + (UN - 1) * stride < HW)) break;
#pragma unroll
      for (int u = 0; u < UN; ++u) qv[u] = Vec16<T>::load_raw(yb + (int64_t)(p0 + u * stride) * C);
      if (rb) {
#pragma unroll
        for (int u = 0; u < UN; ++u) qr[u] = Vec16<T>::load_raw(rb + (int64_t)(p0 + u * stride) * C);
      }
    }
  }
  for (; p0 < HW; p0 += stride) {
    float v0[V], r0[V];
    Vec16<T>::load(yb + (int64_t)p0 * C, v0);
    if (rb) Vec16<T>::load(rb + (int64_t)p0 * C, r0);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float h0 = silu_t<T>(fmaf(v0[k], sc[k], sh[k])) + tb[k];
      if (rb) h0 = fmaf(res_scale, r0[k], h0);
      v0[k] = h0;
    }
    Vec16<T>::store(ob + (int64_t)p0 * C, v0);
  }
}

// GroupNorm apply of the FIRST encoder block's second conv with the block's residual recomputed from the network input.
// The block's input is the init conv's output h = W_i * x + b_i (score_network.py:238, one input channel, no activation
// in between), so its residual projection r = W_r h + b_r (score_network.py:131,151) is itself a 3x3 convolution of the
// one-channel input: r[p][c] = bc[c] + sum_tap wc[tap][c] * x[p + tap], wc = W_r W_i composed once per plan in fp32.  Nine
// FMAs per output element from a 16 KB-per-sample tensor replace writing r (by conv1's epilogue) and reading it back here:
// 2 x the block's activation bytes, and conv1 runs without its fused 1x1 residual output.  A thread keeps one 8-channel
// vector for UN pixels at a time, so each 32-byte weight row read from shared memory serves UN pixels.
template <int UN>
__global__ void __launch_bounds__(256, UN <= 2 ? 4 : 2)
gn_apply_xres_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, const float* __restrict__ stats,
                     float inv_cnt, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ xin, const float* __restrict__ wc, const float* __restrict__ bc,
                     int H, int W, int C, int TH) {
  // A CTA owns TH full-width rows of one image; the rows of x (+ one halo row / column each side, zero outside the
  // image: the init conv's padding) sit in shared memory, so a tap is one LDS shared by the eight threads of a pixel.
  constexpr int V = 8;
  extern __shared__ __align__(16) float s_dyn[];
  float* s_w = s_dyn;                       // [9][C] composed weights, then [C] composed bias
  float* s_x = s_dyn + 10 * C;              // [(TH + 2)][W + 2]
  const int n = blockIdx.y, h0 = blockIdx.x * TH;
  const int Cg = C / 8;
  const int cv = C / V;
  const int c = (threadIdx.x % cv) * V;
  const int prow = threadIdx.x / cv;
  const int ppb = blockDim.x / cv;
  const int Wp = W + 2;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i] = wc[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_w[9 * C + i] = bc[i];
  const float* xb = xin + (int64_t)n * H * W;
  for (int i = threadIdx.x; i < (TH + 2) * Wp; i += blockDim.x) {
    const int r = i / Wp, cc = i - r * Wp;
    const int ih = h0 + r - 1, iw = cc - 1;
    s_x[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(xb + ih * W + iw) : 0.f;
  }
  const int rows = min(TH, H - h0);
  const int n_pix = rows * W;
  const __nv_bfloat16* yb = y + ((int64_t)n * H + h0) * W * C + c;
  __nv_bfloat16* ob = out + ((int64_t)n * H + h0) * W * C + c;
  float sc[V], sh[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int gi = (c + k) / Cg;
    const float s0 = stats[((int64_t)n * 8 + gi) * 2 + 0], s1 = stats[((int64_t)n * 8 + gi) * 2 + 1];
    // inv_cnt > 0: raw (sum, sum of squares) from a conv epilogue; inv_cnt == 0: (mean, rstd) from gn_stats_kernel
    const float mean = inv_cnt > 0.f ? s0 * inv_cnt : s0;
    const float rstd = inv_cnt > 0.f ? rsqrtf(fmaxf(s1 * inv_cnt - mean * mean, 0.f) + eps) : s1;
    const float a = rstd * gamma[c + k];
    sc[k] = a;
    sh[k] = beta[c + k] - mean * a;
  }
  __syncthreads();
#pragma unroll 1
  for (int p0 = prow; p0 < n_pix; p0 += UN * ppb) {
    uint4 q[UN];
    float r[UN][V];
    int xo[UN];                                // offset of the pixel's top-left tap in s_x
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = p0 + u * ppb;
      const bool live = p < n_pix;
      q[u] = live ? Vec16<__nv_bfloat16>::load_raw(yb + (int64_t)p * C) : make_uint4(0, 0, 0, 0);
      const int pr = live ? p / W : 0, pc = live ? p - pr * W : 0;
      xo[u] = pr * Wp + pc;
    }
    {
      const float4 c0 = *reinterpret_cast<const float4*>(&s_w[9 * C + c]), c1 = *reinterpret_cast<const float4*>(&s_w[9 * C + c + 4]);
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        r[u][0] = c0.x; r[u][1] = c0.y; r[u][2] = c0.z; r[u][3] = c0.w;
        r[u][4] = c1.x; r[u][5] = c1.y; r[u][6] = c1.z; r[u][7] = c1.w;
      }
    }
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[(kh * 3 + kw) * C + c]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[(kh * 3 + kw) * C + c + 4]);
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const float xv = s_x[xo[u] + kh * Wp + kw];
          r[u][0] = fmaf(xv, w0.x, r[u][0]); r[u][1] = fmaf(xv, w0.y, r[u][1]);
          r[u][2] = fmaf(xv, w0.z, r[u][2]); r[u][3] = fmaf(xv, w0.w, r[u][3]);
          r[u][4] = fmaf(xv, w1.x, r[u][4]); r[u][5] = fmaf(xv, w1.y, r[u][5]);
          r[u][6] = fmaf(xv, w1.z, r[u][6]); r[u][7] = fmaf(xv, w1.w, r[u][7]);
        }
      }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = p0 + u * ppb;
      if (p >= n_pix) break;
      float v[V];
      Vec16<__nv_bfloat16>::unpack(q[u], v);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = silu_fast(fmaf(v[k], sc[k], sh[k])) + r[u][k];
      Vec16<__nv_bfloat16>::store(ob + (int64_t)p * C, v);
    }
  }
}

// The same kernel with the nine-tap residual on warp-level tensor-core MMAs (C = 64, W % 16 == 0): per 16 consecutive pixels of
// a row a warp forms  r[16 px][64 ch] = patch[16 px][16 taps (9 used)] x wc[16][64]  as eight mma.sync m16n8k16 (bf16 operands,
// fp32 accumulate, the composed bias as the accumulators' initial value) instead of 16 x 64 x 9 FFMAs — the CUDA-core form
// above spends more than half of its issue slots there and runs at 360-390 us for 1.09 GB (B = 1024, 64x64), i.e. compute
// bound.  The A fragment comes straight from the fp32 x tile in shared memory (thread (g, t) of the warp holds taps 2t, 2t+1
// and, for t = 0, tap 8 of pixels g and g + 8), the B fragments (the weights) are loop-invariant registers.  The
// accumulator fragments (pixel rows g / g + 8, channel pairs) go through a per-warp 2.3 KB shared-memory tile (bf16) to reach the
// (pixel, 8-channel vector) layout in which the raw conv output is loaded, normalised and stored with 16-byte accesses.
// Rounding: x and wc enter the MMA as bf16 (the two-launch path stores the residual itself as bf16).
__global__ void __launch_bounds__(256, 2)
gn_apply_xres_mma_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, const float* __restrict__ stats,
                         float inv_cnt, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ xin, const float* __restrict__ wc, const float* __restrict__ bc,
                         int H, int W, int TH) {
  constexpr int C = 64, V = 8;
  extern __shared__ __align__(16) float s_dyn[];
  // [8 warps][16 px][72] bf16: residual tile of a warp, rounded to bf16 as the two-launch path stores it (row stride 144 B:
  // conflict-free for the fragment stores and the 16-byte loads; as fp32 the staging was 64 of the tile's ~70 shared-memory
  // wavefronts and put the kernel's L1 pipe at 72 %)
  __nv_bfloat16* s_r = reinterpret_cast<__nv_bfloat16*>(s_dyn);
  float* s_x = s_dyn + 8 * 16 * 72 / 2;     // [(TH + 2)][W + 2]
  const int n = blockIdx.y, h0 = blockIdx.x * TH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Wp = W + 2;
  const float* xb = xin + (int64_t)n * H * W;
  for (int i = threadIdx.x; i < (TH + 2) * Wp; i += blockDim.x) {
    const int r = i / Wp, cc = i - r * Wp;
    const int ih = h0 + r - 1, iw = cc - 1;
    s_x[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(xb + ih * W + iw) : 0.f;
  }
  // B fragments: b0 = (k = 2t, 2t+1; n = g), b1 = (k = 2t+8, 2t+9; n = g) of n-tile j  <->  wc[tap = k][channel 8j + g]
  uint32_t bf[8][2];
  float bias0[8], bias1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = 8 * j + g;
    const float w0 = __ldg(wc + (2 * t) * C + ch), w1 = __ldg(wc + (2 * t + 1) * C + ch);
    const float w8 = t == 0 ? __ldg(wc + 8 * C + ch) : 0.f;
    __nv_bfloat162 h = __floats2bfloat162_rn(w0, w1);
    bf[j][0] = *reinterpret_cast<const uint32_t*>(&h);
    h = __floats2bfloat162_rn(w8, 0.f);
    bf[j][1] = *reinterpret_cast<const uint32_t*>(&h);
    bias0[j] = __ldg(bc + 8 * j + 2 * t);
    bias1[j] = __ldg(bc + 8 * j + 2 * t + 1);
  }
  // GroupNorm scale / shift of my 8-channel vector (vector index lane % 8: fixed for the thread's whole life)
  const int c = (lane & 7) * V;
  float sc[V], sh[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int gi = (c + k) / (C / 8);
    const float s0 = stats[((int64_t)n * 8 + gi) * 2 + 0], s1 = stats[((int64_t)n * 8 + gi) * 2 + 1];
    // inv_cnt > 0: raw (sum, sum of squares) from a conv epilogue; inv_cnt == 0: (mean, rstd) from gn_stats_kernel
    const float mean = inv_cnt > 0.f ? s0 * inv_cnt : s0;
    const float rstd = inv_cnt > 0.f ? rsqrtf(fmaxf(s1 * inv_cnt - mean * mean, 0.f) + eps) : s1;
    const float a = rstd * gamma[c + k];
    sc[k] = a;
    sh[k] = beta[c + k] - mean * a;
  }
  __syncthreads();
  const int rows = min(TH, H - h0);
  const int n_tiles = rows * W / 16;                      // W % 16 == 0: a tile is 16 consecutive pixels of one row
  const __nv_bfloat16* yb = y + ((int64_t)n * H + h0) * W * C;
  __nv_bfloat16* ob = out + ((int64_t)n * H + h0) * W * C;
  __nv_bfloat16* rt = s_r + warp * 16 * 72;
  // taps of this thread: k0 = 2t -> (kh, kw), k1 = 2t + 1
  const int k0 = 2 * t, k1 = 2 * t + 1;
  const int o0 = (k0 / 3) * Wp + (k0 % 3), o1 = (k1 / 3) * Wp + (k1 % 3), o8 = 2 * Wp + 2;
  // the raw conv output of my four (pixel, vector) slots is requested one tile ahead: with the loads issued only at the top
  // of the tile they serve, a warp has bytes in flight for about half of its time and sixteen warps per SM do not cover
  // the HBM latency (249 us for 1.07 GB at B = 1024, 64x64)
  uint4 q[4];
  if (warp < n_tiles) {
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = *reinterpret_cast<const uint4*>(yb + (int64_t)(warp * 16 + (lane >> 3) + 4 * i) * C + c);
  }
#pragma unroll 1
  for (int tile = warp; tile < n_tiles; tile += 8) {
    const int p0 = tile * 16;
    const int pr = p0 / W, pc = p0 - pr * W;
    uint4 qn[4];
    {
      const int pn = tile + 8 < n_tiles ? p0 + 128 : p0;   // last round: a harmless re-read of the current slots
#pragma unroll
      for (int i = 0; i < 4; ++i) qn[i] = *reinterpret_cast<const uint4*>(yb + (int64_t)(pn + (lane >> 3) + 4 * i) * C + c);
    }
    // A fragments: rows g and g + 8 of the tile
    const float* xa = s_x + pr * Wp + pc + g;
    const float* xc = xa + 8;
    uint32_t a0, a1, a2, a3;
    {
      __nv_bfloat162 h = __floats2bfloat162_rn(xa[o0], xa[o1]);
      a0 = *reinterpret_cast<const uint32_t*>(&h);
      h = __floats2bfloat162_rn(xc[o0], xc[o1]);
      a1 = *reinterpret_cast<const uint32_t*>(&h);
      h = __floats2bfloat162_rn(t == 0 ? xa[o8] : 0.f, 0.f);
      a2 = *reinterpret_cast<const uint32_t*>(&h);
      h = __floats2bfloat162_rn(t == 0 ? xc[o8] : 0.f, 0.f);
      a3 = *reinterpret_cast<const uint32_t*>(&h);
    }
    __syncwarp();                                         // the previous tile's reads of rt are done
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d0 = bias0[j], d1 = bias1[j], d2 = bias0[j], d3 = bias1[j];
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                   "{%0, %1, %2, %3};"
                   : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bf[j][0]), "r"(bf[j][1]));
      // accumulator fragment: (row g, channels 8j + 2t, + 1), (row g + 8, same channels)
      *reinterpret_cast<__nv_bfloat162*>(rt + g * 72 + 8 * j + 2 * t) = __floats2bfloat162_rn(d0, d1);
      *reinterpret_cast<__nv_bfloat162*>(rt + (g + 8) * 72 + 8 * j + 2 * t) = __floats2bfloat162_rn(d2, d3);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pl = (lane >> 3) + 4 * i;                 // pixel inside the tile
      float rr[V];
      Vec16<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(rt + pl * 72 + c), rr);
      float v[V];
      Vec16<__nv_bfloat16>::unpack(q[i], v);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = silu_fast(fmaf(v[k], sc[k], sh[k])) + rr[k];
      Vec16<__nv_bfloat16>::store(ob + (int64_t)(p0 + pl) * C + c, v);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = qn[i];
  }
}

// First encoder block, conv1 composed with the init conv (C = 32 -> 64 channels, bf16 path).  The init conv's 32-channel
// output has ONE consumer left once the block's residual is recomputed from the network input (above): conv1.  Both are
// linear, so  y1 = conv1(init_conv(x))  is a function of the one-channel input alone:
//   y1[p][co] = b1[co] + sum_t in(p + t) * ( Bc[co][t] + sum_s Wc[co][t][s] * x~[p + t + s] )        t, s over the 3x3 offsets
//   Wc[co][t][s] = sum_ci W1[co][ci][t] Wi[ci][s],   Bc[co][t] = sum_ci W1[co][ci][t] bi[ci]
// x~ is the zero-padded input and in(q) says whether pixel q lies inside the image: conv1 pads the init conv's OUTPUT with
// zeros, so a tap t that leaves the image contributes neither its weights nor its share of the init bias.
// A warp works on 4 x 4-pixel tiles (16 MMA rows), fetched from a CTA-wide counter, in one of two forms:
//   * interior tile (no pixel on the image's outer ring, 77 % of the tiles at 64x64): every in() is 1 and the 81 (t, s)
//     products collapse to ONE 5x5 kernel  W5[co][u] = sum_{t + s = u} Wc[co][t][s]  plus the summed bias taps:
//     [16 px][32] x [32][64] (25 products, two bias columns, 5 zero columns) = 2 x 8 mma.sync m16n8k16;
//   * border tile: the products are kept apart and masked per pixel by in(p + t):  [16 px][96] x [96][64]  (81 products,
//     9 bias taps, two columns for b1, 4 zero columns) = 6 x 8 MMAs.
// The unmasked bias (b1, plus the summed bias taps in the interior form) rides in the MMA as TWO columns against A = 1:
// its bf16 rounding and the bf16 rounding of what that left over (error 2^-17 of the bias), which frees sixteen
// registers per thread for a third resident CTA (measured: 216 -> 212 us from the fold, 211 us with three CTAs).
// bf16 operands, fp32 accumulators, 2 x 32 (or 96) x 64 MACs per pixel instead of conv1's 2 x 288 x 64 + the init conv's
// 2 x 9 x 32, no 32-channel activation in HBM (268 MB written and read at B = 1024, 64x64), one launch instead of two.
// The first form of this kernel ran every tile (16 pixels of a row) through the masked 96-wide form and was bound by
// shared-memory wavefronts (~290 per tile: 160 masked A loads, 96 for the B fragments): 255 us at B = 1024 against 82 us
// for its 537 MB of stores; an interior tile costs ~80.  The A fragments come straight from the fp32 x tile in shared
// memory (two-pixel halo); the B fragments are built once per plan; GroupNorm sums are taken on the fp32 accumulators, an
// n-tile being exactly one group.
__global__ void __launch_bounds__(256, 3)
enc0_composed_kernel(const float* __restrict__ xin, __nv_bfloat16* __restrict__ y, float* __restrict__ stats,
                     const uint2* __restrict__ bfrag, int H, int W, int TH) {
  constexpr int C = 64;
  constexpr int kFragMasked = 6 * 8 * 32, kFragInner = 2 * 8 * 32;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  uint2* s_b = reinterpret_cast<uint2*>(sm_raw);                                   // [6][8][32] masked + [2][8][32] interior
  const uint2* s_b5 = s_b + kFragMasked;
  __nv_bfloat16* s_y = reinterpret_cast<__nv_bfloat16*>(sm_raw + (kFragMasked + kFragInner) * 8);   // [8 warps][16 px][72]
  float* s_x = reinterpret_cast<float*>(sm_raw + (kFragMasked + kFragInner) * 8 + 8 * 16 * 72 * 2); // [(TH + 4)][W + 4]
  __shared__ float s_stat[16];
  __shared__ uint32_t s_ktab[20 * 4];
  __shared__ int s_next;
  const int n = blockIdx.y, h0 = blockIdx.x * TH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int Wp = W + 4;
  for (int i = threadIdx.x; i < kFragMasked + kFragInner; i += blockDim.x) s_b[i] = bfrag[i];
  if (threadIdx.x < 16) s_stat[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) s_next = 0;
  // masked form: the 20 regular k-indices of a lane (k-steps 0..4: k = 16 kk + 2 t4 + {0, 1, 8, 9} = 9 t + s < 80): offset of
  // x~[p + t + s] relative to the pixel and the tap t whose in() bit gates it, packed as (offset + 1024) << 4 | t
  if (threadIdx.x < 80) {
    const int tq = threadIdx.x & 3, e = threadIdx.x >> 2, kk = e >> 2, q = e & 3;
    const int k = 16 * kk + 2 * tq + (q & 1) + 8 * (q >> 1);
    const int t = k / 9, sI = k - 9 * t;
    const int off = (t / 3 - 1 + sI / 3 - 1) * Wp + (t % 3 - 1 + sI % 3 - 1);
    s_ktab[threadIdx.x] = ((uint32_t)(off + 1024) << 4) | (uint32_t)t;
  }
  const float* xb = xin + (int64_t)n * H * W;
  for (int i = threadIdx.x; i < (TH + 4) * Wp; i += blockDim.x) {
    const int r = i / Wp, cc = i - r * Wp;
    const int ih = h0 + r - 2, iw = cc - 2;
    s_x[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(xb + ih * W + iw) : 0.f;
  }
  // interior form: k = 5 (uy + 2) + (ux + 2) < 25; my k-indices 2 t4 + {0, 1, 8, 9} (k-step 0), 16 + 2 t4 + {0, 1} (k-step 1);
  // k = 24 (the last product) and k = 25 (the bias column) sit in lane t4 = 0's upper pair of k-step 1, the rest is zero
  int o5[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const int k = q < 4 ? 2 * t4 + (q & 1) + 8 * (q >> 1) : 16 + 2 * t4 + (q & 1);
    o5[q] = (k / 5 - 2) * Wp + (k % 5 - 2);
  }
  const int o24 = 2 * Wp + 2;
  float st_s[8], st_ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { st_s[j] = 0.f; st_ss[j] = 0.f; }
  __syncthreads();
  const int rows = min(TH, H - h0);
  const int tiles_x = W >> 2;
  const int n_tiles = (rows >> 2) * tiles_x;                // H % 4 == 0, TH % 4 == 0 (host)
  __nv_bfloat16* ob = y + ((int64_t)n * H + h0) * W * C;
  __nv_bfloat16* yt = s_y + warp * 16 * 72;
  auto pack2 = [](float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  };
#pragma unroll 1
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = atomicAdd(&s_next, 1);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) break;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int r0 = 4 * ty, c0 = 4 * tx;                   // tile origin: row inside the CTA's rows, image column
    const int ih0 = h0 + r0;
    // MMA rows g and g + 8 of the tile are the pixels (g >> 2, g & 3) and two rows below
    const float* xc0 = s_x + (r0 + (g >> 2) + 2) * Wp + (c0 + (g & 3) + 2);
    const float* xc1 = xc0 + 2 * Wp;
    float d[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { d[j][0] = 0.f; d[j][1] = 0.f; d[j][2] = 0.f; d[j][3] = 0.f; }
    if (ih0 > 0 && ih0 + 4 < H && c0 > 0 && c0 + 4 < W) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a0, a1, a2, a3;
        if (kk == 0) {
          a0 = pack2(xc0[o5[0]], xc0[o5[1]]);
          a1 = pack2(xc1[o5[0]], xc1[o5[1]]);
          a2 = pack2(xc0[o5[2]], xc0[o5[3]]);
          a3 = pack2(xc1[o5[2]], xc1[o5[3]]);
        } else {
          a0 = pack2(xc0[o5[4]], xc0[o5[5]]);
          a1 = pack2(xc1[o5[4]], xc1[o5[5]]);
          a2 = t4 == 0 ? pack2(xc0[o24], 1.f) : (t4 == 1 ? 0x00003f80u : 0u);   // k = 24 | 25;  k = 26 (bf16 1.0) | 27
          a3 = t4 == 0 ? pack2(xc1[o24], 1.f) : (t4 == 1 ? 0x00003f80u : 0u);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint2 b = s_b5[(kk * 8 + j) * 32 + lane];
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                       "{%0, %1, %2, %3};"
                       : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b.x), "r"(b.y));
        }
      }
    } else {
      // in(p + t) for my two pixels: bit t = 3 (th + 1) + (tw + 1); bit 9 = 0 (zero columns)
      const int iw = c0 + (g & 3);
      const uint32_t colok = (iw > 0 ? 1u : 0u) | 2u | (iw < W - 1 ? 4u : 0u);
      uint32_t inm[2];
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2) {
        const int ih = ih0 + (g >> 2) + 2 * r2;
        const uint32_t rowok = (ih > 0 ? 1u : 0u) | 2u | (ih < H - 1 ? 4u : 0u);
        uint32_t m = 0;
#pragma unroll
        for (int th = 0; th < 3; ++th)
          if ((rowok >> th) & 1u) m |= colok << (3 * th);
        inm[r2] = m;
      }
#pragma unroll
      for (int kk = 0; kk < 6; ++kk) {
        float av[2][4];                                   // [pixel row][q]: k = 16 kk + 2 t4 + {0, 1, 8, 9}
        if (kk < 5) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t e = s_ktab[(kk * 4 + q) * 4 + t4];
            const int off = (int)(e >> 4) - 1024;
            const uint32_t tb = e & 15u;
            av[0][q] = ((inm[0] >> tb) & 1u) ? xc0[off] : 0.f;
            av[1][q] = ((inm[1] >> tb) & 1u) ? xc1[off] : 0.f;
          }
        } else {
          // k = 80: the last product (t = 8, s = 8); k = 81 .. 89: the bias taps (value 1 where the tap is inside); 90, 91: b1; 92 ..: zero
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = 80 + 2 * t4 + (q & 1) + 8 * (q >> 1);
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
              float v = 0.f;
              if (k == 80) v = ((inm[r2] >> 8) & 1u) ? (r2 ? xc1 : xc0)[2 * Wp + 2] : 0.f;
              else if (k < 90) v = ((inm[r2] >> (k - 81)) & 1u) ? 1.f : 0.f;
              else if (k < 92) v = 1.f;                    // conv1's own bias, bf16 high + low part
              av[r2][q] = v;
            }
          }
        }
        const uint32_t a0 = pack2(av[0][0], av[0][1]), a1 = pack2(av[1][0], av[1][1]);
        const uint32_t a2 = pack2(av[0][2], av[0][3]), a3 = pack2(av[1][2], av[1][3]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint2 b = s_b[(kk * 8 + j) * 32 + lane];
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                       "{%0, %1, %2, %3};"
                       : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b.x), "r"(b.y));
        }
      }
    }
    __syncwarp();                                         // the previous tile's reads of yt are done
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st_s[j] += (d[j][0] + d[j][1]) + (d[j][2] + d[j][3]);
      st_ss[j] = fmaf(d[j][0], d[j][0], fmaf(d[j][1], d[j][1], fmaf(d[j][2], d[j][2], fmaf(d[j][3], d[j][3], st_ss[j]))));
      __nv_bfloat162 h = __floats2bfloat162_rn(d[j][0], d[j][1]);
      *reinterpret_cast<__nv_bfloat162*>(yt + g * 72 + 8 * j + 2 * t4) = h;
      h = __floats2bfloat162_rn(d[j][2], d[j][3]);
      *reinterpret_cast<__nv_bfloat162*>(yt + (g + 8) * 72 + 8 * j + 2 * t4) = h;
    }
    __syncwarp();
    // staged row pl = 4 i + (lane >> 3) is pixel (i, lane >> 3) of the tile: four pixels x 128 B contiguous per store
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pl = (lane >> 3) + 4 * i;
      *reinterpret_cast<uint4*>(ob + ((int64_t)(r0 + i) * W + c0 + (lane >> 3)) * C + (lane & 7) * 8) =
          *reinterpret_cast<const uint4*>(yt + pl * 72 + (lane & 7) * 8);
    }
  }
  // GroupNorm sums of this CTA's rows: group j = channels 8 j .. 8 j + 7 = n-tile j
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a = warp_sum(st_s[j]), b = warp_sum(st_ss[j]);
    if (lane == 0) { atomicAdd(&s_stat[2 * j], a); atomicAdd(&s_stat[2 * j + 1], b); }
  }
  __syncthreads();
  if (threadIdx.x < 16) atomicAdd(stats + (int64_t)n * 16 + threadIdx.x, s_stat[threadIdx.x]);
}

// stand-alone time path: gridDim.x = distinct t values, gridDim.y CTAs per value (see time_path above)
__global__ void __launch_bounds__(256)
time_kernel(const __grid_constant__ TimeArgs ta) {
  extern __shared__ float sm[];
  time_path(ta, blockIdx.x, blockIdx.y, gridDim.y, sm);
}

// weight repack: Conv2d [co][ci][kh][kw] (transposed = 0) or ConvTranspose2d [ci][co][kh][kw]
// (transposed = 1)  ->  [tap][co][ci] in fp32 and bf16
__global__ void repack_kernel(const float* __restrict__ src, float* __restrict__ w32, __nv_bfloat16* __restrict__ w16,
                              int co_n, int ci_n, int taps, int transposed) {
  const int64_t total = (int64_t)taps * co_n * ci_n;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = (int)(i % ci_n);
  const int co = (int)((i / ci_n) % co_n);
  const int t = (int)(i / ((int64_t)ci_n * co_n));
  const int64_t s = transposed ? (((int64_t)ci * co_n + co) * taps + t) : (((int64_t)co * ci_n + ci) * taps + t);
  const float v = src[s];
  w32[i] = v;
  w16[i] = __float2bfloat16_rn(v);
}

// [9][Cout][C] (tap = kh * 3 + kw) -> [4][Cout][4C]: tap t = (dh' + 1) * 2 + (dw' + 1), channel k = py * 2C + px * C + c,
// source tap kh = 2 dh' + py + 1, kw = 2 dw' + px + 1 (zero where that falls outside the 3x3 kernel)
__global__ void repack_s2d_kernel(const __nv_bfloat16* __restrict__ w9, __nv_bfloat16* __restrict__ w4, int Cout, int C) {
  const int64_t total = (int64_t)4 * Cout * 4 * C;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % (4 * C));
  const int co = (int)((i / (4 * C)) % Cout);
  const int t = (int)(i / ((int64_t)4 * C * Cout));
  const int py = k / (2 * C), px = (k / C) & 1, c = k % C;
  const int kh = 2 * (t / 2 - 1) + py + 1, kw = 2 * (t % 2 - 1) + px + 1;
  w4[i] = (kh >= 0 && kh < 3 && kw >= 0 && kw < 3) ? w9[((int64_t)(kh * 3 + kw) * Cout + co) * C + c] : __float2bfloat16_rn(0.f);
}

// Bilinear resize (align_corners = false) of an NHWC tensor: the size fix-up ScoreNetwork.forward applies to an
// upsampled map whose side is not the skip connection's (score_network.py:253-255, torch's upsample_bilinear2d rule:
// src = (dst + 0.5) * in/out - 0.5 clamped at 0, neighbour i1 = min(i0 + 1, in - 1)).  One thread per 4 channels of
// one output pixel; only lattices whose side is not divisible by 2^num_blocks ever reach it.
template <typename T>
__global__ void __launch_bounds__(256)
bilinear_resize_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t total_quads, int Hi, int Wi, int Ho, int Wo,
                       int C, float scale_h, float scale_w) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total_quads) return;
  const int cq = C >> 2;
  const int c4 = (int)(q % cq) * 4;
  int64_t pix = q / cq;
  const int ow = (int)(pix % Wo); pix /= Wo;
  const int oh = (int)(pix % Ho);
  const int64_t n = pix / Ho;
  const float sh = fmaxf(__fmaf_rn(scale_h, (float)oh + 0.5f, -0.5f), 0.0f);
  const float sw = fmaxf(__fmaf_rn(scale_w, (float)ow + 0.5f, -0.5f), 0.0f);
  const int h0 = min((int)sh, Hi - 1), w0 = min((int)sw, Wi - 1);
  const int h1 = min(h0 + 1, Hi - 1), w1 = min(w0 + 1, Wi - 1);
  const float lh1 = fminf(fmaxf(sh - (float)h0, 0.0f), 1.0f), lw1 = fminf(fmaxf(sw - (float)w0, 0.0f), 1.0f);
  const float lh0 = 1.0f - lh1, lw0 = 1.0f - lw1;
  const T* base = in + n * Hi * Wi * C + c4;
  float v00[4], v01[4], v10[4], v11[4], r[4];
  Quad<T>::load(base + ((int64_t)h0 * Wi + w0) * C, v00);
  Quad<T>::load(base + ((int64_t)h0 * Wi + w1) * C, v01);
  Quad<T>::load(base + ((int64_t)h1 * Wi + w0) * C, v10);
  Quad<T>::load(base + ((int64_t)h1 * Wi + w1) * C, v11);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    r[k] = lh0 * (lw0 * v00[k] + lw1 * v01[k]) + lh1 * (lw0 * v10[k] + lw1 * v11[k]);
  Quad<T>::store(out + ((n * Ho + oh) * Wo + ow) * C + c4, r);
}

template <typename T>
__global__ void nhwc_to_nchw_f32_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t total, int HW, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // index into NCHW output
  if (i >= total) return;
  const int64_t n = i / ((int64_t)C * HW);
  const int c = (int)((i / HW) % C);
  const int p = (int)(i % HW);
  out[i] = to_f<T>(in[(n * HW + p) * C + c]);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// plan
struct BlockW {
  int cin = 0, cout = 0, tb_off = 0;
  bool has_res = false;
  ConvWeights conv1, conv2, res;
  float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
};

}  // namespace dbs

struct dbs_unet_plan {
  dbs_unet_config cfg;
  dbs::DevArena arena;
  float *freqs = nullptr, *tw0 = nullptr, *tb0 = nullptr, *tw2 = nullptr, *tb2 = nullptr;
  float *twt = nullptr, *tbt = nullptr;  // concatenated per-block time Linear [sumC][T], [sumC]
  int sumC = 0;
  float *init_w = nullptr, *init_b = nullptr, *final_w = nullptr, *final_b = nullptr;
  std::vector<float> final_w_host;   // [tap][C] host copy of the final conv's weights
  uint32_t* final_bfrag = nullptr;   // their mma.sync B fragments (final_bfrag_host), C a multiple of 16 only
  // first encoder block: residual projection composed with the init conv, [9][cout] + [cout] (gn_apply_xres_kernel)
  float *enc0_wc = nullptr, *enc0_bc = nullptr;
  uint2* enc0_bfrag = nullptr;       // enc0_composed_kernel: B fragments of conv1 composed with the init conv, [6][8][32]
  std::vector<float> init_w_host, init_b_host;   // [tap][C], [C]: kernel-parameter copies for the init conv
  float final_b_host = 0.f;
  std::vector<dbs::BlockW> enc, dec;
  dbs::BlockW mid;
  std::vector<dbs::ConvWeights> down, up;
};

namespace dbs {
namespace {

struct ParamTable {
  const dbs_param* p;
  int n;
  const dbs_param* find(const std::string& name) const {
    for (int i = 0; i < n; ++i)
      if (name == p[i].name) return &p[i];
    return nullptr;
  }
};

int dev_alloc(dbs_unet_plan* plan, size_t bytes, void** out) {
  DBS_CUDA(plan->arena.take(bytes, out));
  return DBS_OK;
}

// copy a named fp32 parameter to a fresh device buffer
int upload(dbs_unet_plan* plan, const ParamTable& pt, const std::string& name, int64_t numel, float** out) {
  const dbs_param* q = pt.find(name);
  DBS_REQUIRE(q != nullptr, "unet plan: parameter '%s' missing from state_dict", name.c_str());
  DBS_REQUIRE(q->numel == numel, "unet plan: parameter '%s' has %lld elements, expected %lld", name.c_str(),
              (long long)q->numel, (long long)numel);
  void* d = nullptr;
  int rc = dev_alloc(plan, (size_t)numel * 4, &d);
  if (rc) return rc;
  DBS_CUDA(cudaMemcpy(d, q->data, (size_t)numel * 4, q->is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  *out = (float*)d;
  return DBS_OK;
}

int make_conv(dbs_unet_plan* plan, const ParamTable& pt, const std::string& prefix, int cin, int cout, int ks,
              bool transposed, ConvWeights* cw) {
  cw->cin = cin; cw->cout = cout; cw->taps = ks * ks;
  const int64_t numel = (int64_t)cin * cout * ks * ks;
  float* raw = nullptr;
  int rc = upload(plan, pt, prefix + ".weight", numel, &raw);
  if (rc) return rc;
  rc = upload(plan, pt, prefix + ".bias", cout, &cw->bias);
  if (rc) return rc;
  void *w32 = nullptr, *w16 = nullptr;
  if ((rc = dev_alloc(plan, (size_t)numel * 4, &w32))) return rc;
  if ((rc = dev_alloc(plan, (size_t)numel * 2, &w16))) return rc;
  cw->w32 = (float*)w32; cw->w16 = (__nv_bfloat16*)w16;
  repack_kernel<<<(unsigned)((numel + 255) / 256), 256>>>(raw, cw->w32, cw->w16, cout, cin, ks * ks, transposed ? 1 : 0);
  DBS_LAUNCHED();
  return DBS_OK;
}

int make_block(dbs_unet_plan* plan, const ParamTable& pt, const std::string& prefix, int cin, int cout, BlockW* b) {
  b->cin = cin; b->cout = cout; b->has_res = cin != cout;
  int rc;
  if ((rc = make_conv(plan, pt, prefix + "conv1", cin, cout, 3, false, &b->conv1))) return rc;
  if ((rc = make_conv(plan, pt, prefix + "conv2", cout, cout, 3, false, &b->conv2))) return rc;
  if (b->has_res && (rc = make_conv(plan, pt, prefix + "residual", cin, cout, 1, false, &b->res))) return rc;
  if ((rc = upload(plan, pt, prefix + "norm1.weight", cout, &b->g1))) return rc;
  if ((rc = upload(plan, pt, prefix + "norm1.bias", cout, &b->b1))) return rc;
  if ((rc = upload(plan, pt, prefix + "norm2.weight", cout, &b->g2))) return rc;
  if ((rc = upload(plan, pt, prefix + "norm2.bias", cout, &b->b2))) return rc;
  return DBS_OK;
}

struct Bump {
  char* base; int64_t cap; int64_t off = 0;
  void* take(int64_t bytes) {
    const int64_t a = (off + 1023) & ~(int64_t)1023;
    off = a + bytes;
    return (base && off <= cap) ? base + a : nullptr;
  }
};

template <typename T> struct WSel;
template <> struct WSel<float> { using type = float; static const float* get(const ConvWeights& w) { return w.w32; } };
template <> struct WSel<__nv_bfloat16> { using type = __nv_bfloat16; static const __nv_bfloat16* get(const ConvWeights& w) { return w.w16; } };

// The last decoder block can leave its closing GroupNorm apply to the final conv kernel, which then reads
// the raw conv2 output + residual and normalises on the fly (final_fused_kernel).
struct DeferredGn {
  bool want = false, done = false;
  const void* y = nullptr; const void* res = nullptr;
  const float *gamma = nullptr, *beta = nullptr, *stats = nullptr;
};

struct Ctx {
  dbs_unet_plan* plan;
  cudaStream_t st;
  int64_t B;
  int n_t;            // 1 or B
  float* tbias;       // [n_t][sumC]
  float* stats;       // [B][8][2]: the GroupNorm statistics slice of the conv being run (see next_stats)
  bool use_tc;
  const char* tap; float* tap_out; int64_t tap_numel; bool tapped = false;
  DeferredGn* defer = nullptr;
  const float* xin = nullptr;      // the network input (B, 1, H, W): set while the FIRST encoder block runs (residual from input)
  bool compose0 = false;           // the first block's conv1 runs composed with the init conv (no init-conv activation exists)
  float* stats_base = nullptr; int64_t stats_stride = 0; int* stats_next = nullptr; int stats_slices = 0;
  // bf16 path: every conv+GroupNorm pair gets its own slice of one region zeroed by ONE memset per forward
  // (instead of a memset node in front of each conv); returns false when the region is used up
  bool next_stats() {
    if (!stats_base || *stats_next >= stats_slices) return false;
    stats = stats_base + (int64_t)(*stats_next)++ * stats_stride;
    return true;
  }
};

template <typename T>
int run_conv(Ctx& c, const ConvGeom& g, const T* in0, const T* in1, const ConvWeights& w, T* out, float* raw_stats,
             int groups, bool* stats_done) {
  if (stats_done) *stats_done = false;
  if (c.use_tc && sizeof(T) == 2) {
    const int rc = conv_tc_launch(g, (const __nv_bfloat16*)in0, (const __nv_bfloat16*)in1, w.w16, w.bias,
                                  (__nv_bfloat16*)out, raw_stats, groups, c.st);
    if (rc == DBS_OK) { if (stats_done) *stats_done = raw_stats != nullptr; return rc; }
    if (rc != DBS_ERR_UNSUPPORTED) return rc;
  }
  const int Cin = g.C0 + g.C1;
  DBS_REQUIRE(g.C0 % 4 == 0 && g.C1 % 4 == 0 && g.Cout % 4 == 0 && Cin > 0,
              "unet: channel counts must be multiples of 4 (got %d+%d -> %d)", g.C0, g.C1, g.Cout);
  const int64_t M = (int64_t)g.B * g.Hm * g.Wm;
  dim3 grid((unsigned)((M + kGT - 1) / kGT), (unsigned)((g.Cout + kGT - 1) / kGT));
  ProfScope prof(kProfConvGeneric, c.st);
  conv_generic_kernel<T, typename WSel<T>::type><<<grid, 256, 0, c.st>>>(g, in0, in1, WSel<T>::get(w), w.bias, out);
  DBS_LAUNCHED();
  return DBS_OK;
}

template <typename T>
int maybe_tap(Ctx& c, const char* name, const T* buf, int64_t B, int H, int W, int C) {
  if (!c.tap || c.tapped || strcmp(c.tap, name) != 0) return DBS_OK;
  const int64_t total = B * H * W * C;
  DBS_REQUIRE(total == c.tap_numel, "unet tap '%s': buffer has %lld elements, caller passed %lld", name,
              (long long)total, (long long)c.tap_numel);
  nhwc_to_nchw_f32_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, c.st>>>(buf, c.tap_out, total, H * W, C);
  DBS_LAUNCHED();
  c.tapped = true;
  return DBS_OK;
}

// GroupNorm(min(8,C)) + SiLU (+time bias) (+residual) on y -> out
template <typename T>
int run_gn(Ctx& c, const T* y, T* out, bool have_raw_stats, const float* gamma, const float* beta, int tb_off,
           bool add_time, const T* res, int H, int W, int C) {
  const int G = C < 8 ? C : 8;
  const int HW = H * W;
  if (!have_raw_stats) {
    ProfScope prof(kProfGnStats, c.st);
    gn_stats_kernel<T><<<(unsigned)(c.B * G), 256, 0, c.st>>>(y, c.stats, HW, C, G, 1e-5f);
    DBS_LAUNCHED();
  }
  const int V = 16 / (int)sizeof(T);
  DBS_REQUIRE(C % V == 0, "unet: channel count %d is not a multiple of %d", C, V);
  const int cv = C / V;
  DBS_REQUIRE(cv <= 256, "unet: %d channels exceed the GroupNorm kernel's limit", C);
  static const int tune_threads = [] { const char* e = getenv("DBS_GN_THREADS"); return e ? atoi(e) : 256; }();
  static const int tune_un = [] { const char* e = getenv("DBS_GN_UN"); return e ? atoi(e) : 2; }();
  static const int tune_cap = [] { const char* e = getenv("DBS_GN_CAP"); return e ? atoi(e) : 8; }();
  const int threads = (tune_threads / cv) * cv;   // a whole number of pixels per block iteration
  const int ppb = threads / cv;
  int bx = (HW + tune_un * ppb - 1) / (tune_un * ppb);
  const int cap = (int)std::max<int64_t>(1, (148 * tune_cap) / c.B);   // ~tune_cap CTAs per SM across the batch
  if (bx > cap) bx = cap;
  const float inv_cnt = 1.0f / (float)((int64_t)HW * (C / G));
  prof_set_label("HW%d C%d res%d", HW, C, res ? 1 : 0);
  ProfScope prof(kProfGnApply, c.st);
  const dim3 grid((unsigned)bx, (unsigned)c.B);
#define DBS_GN_LAUNCH(UN)                                                                                   \
  DBS_CUDA(launch_chained(gn_apply_kernel<T, UN>, grid, dim3(threads), 0, c.st, y, out, c.stats, have_raw_stats ? 1 : 0,     \
                          inv_cnt, 1e-5f, gamma, beta, add_time ? c.tbias + tb_off : (const float*)nullptr,          \
                          c.n_t == 1 ? 0 : c.plan->sumC, res, 1.0f, HW, C, G))
  if (tune_un >= 4) DBS_GN_LAUNCH(4);
  else if (tune_un == 2) DBS_GN_LAUNCH(2);
  else DBS_GN_LAUNCH(1);
#undef DBS_GN_LAUNCH
  DBS_LAUNCHED();
  return DBS_OK;
}

// ConvBlock.forward (score_network.py:133-151)
template <typename T>
int run_block(Ctx& c, const BlockW& bw, const char* name, const T* in0, int C0, const T* in1, int C1, int H, int W,
              T* ty, T* th, T* tr, T* out) {
  int rc;
  const int G = bw.cout < 8 ? bw.cout : 8;
  bool sd = false;
  const size_t stat_bytes = (size_t)c.B * G * 2 * sizeof(float);
  ConvGeom g1 = geom_conv((int)c.B, H, W, C0, C1, bw.cout, 3, 1);
  ConvGeom g2 = geom_conv((int)c.B, H, W, bw.cout, 0, bw.cout, 3, 1);
  if (!bw.has_res) DBS_REQUIRE(C1 == 0, "unet: identity residual with a concatenated input");

  // Low-resolution levels (an image fits inside one 128-pixel GEMM tile): the GroupNorm statistics of a tile's
  // images are complete inside the conv epilogue, so normalise + SiLU (+ time bias / + residual) happen there
  // and neither the raw conv output nor a separate GroupNorm pass exists (conv_tc.cu, second epilogue pass).
  static const bool fuse_off = [] { const char* e = getenv("DBS_GN_FUSE"); return e && e[0] == '0'; }();
  if (c.use_tc && sizeof(T) == 2 && !c.tap && !fuse_off && H * W <= 128 && bw.cout >= 128 && bw.cout <= 256) {
    GnFuse f1{bw.g1, bw.b1, c.tbias + bw.tb_off, c.n_t == 1 ? 0 : (int64_t)c.plan->sumC, nullptr};
    rc = conv_tc_launch(g1, (const __nv_bfloat16*)in0, (const __nv_bfloat16*)in1, bw.conv1.w16, bw.conv1.bias,
                        (__nv_bfloat16*)th, c.stats, G, c.st, &f1);
    if (rc == DBS_OK) {
      const T* res = in0;
      if (bw.has_res) {
        ConvGeom gr = geom_conv((int)c.B, H, W, C0, C1, bw.cout, 1, 1);
        if ((rc = run_conv<T>(c, gr, in0, in1, bw.res, tr, nullptr, G, nullptr))) return rc;
        res = tr;
      }
      GnFuse f2{bw.g2, bw.b2, nullptr, 0, (const __nv_bfloat16*)res};
      rc = conv_tc_launch(g2, (const __nv_bfloat16*)th, nullptr, bw.conv2.w16, bw.conv2.bias, (__nv_bfloat16*)out, c.stats, G,
                          c.st, &f2);
      if (rc == DBS_OK) return DBS_OK;
      if (rc != DBS_ERR_UNSUPPORTED) return rc;
      // conv2 not eligible (cannot happen when conv1 was): finish the block on the unfused path below from th
      if (!c.next_stats()) DBS_CUDA(cudaMemsetAsync(c.stats, 0, stat_bytes, c.st));
      if ((rc = run_conv<T>(c, g2, th, (const T*)nullptr, bw.conv2, ty, c.stats, G, &sd))) return rc;
      return run_gn<T>(c, ty, out, sd, bw.g2, bw.b2, 0, false, res, H, W, bw.cout);
    }
    if (rc != DBS_ERR_UNSUPPORTED) return rc;
  }

  if (c.use_tc && !c.next_stats()) DBS_CUDA(cudaMemsetAsync(c.stats, 0, stat_bytes, c.st));
  // First encoder block on the tensor-core path: its residual projection is recomputed from the one-channel network input
  // inside the closing GroupNorm apply, so conv1 runs WITHOUT the fused 1x1 residual output and the residual tensor is
  // neither written nor read.  Measured at 64x64, B = 1024: conv1 474 -> 329 us.  With the residual as 72 FMAs per
  // 8-channel vector (gn_apply_xres_kernel) the apply kernel was compute-bound at 360-396 us against 259 us for the
  // HBM-bound apply it replaces — net zero; with the residual on warp-level MMAs (gn_apply_xres_mma_kernel, 64 channels,
  // W % 16 == 0) it takes 249 us: 733 -> 578 us for the pair, which is why the path is ON where the MMA kernel applies.
  // DBS_XRES=0 keeps the fused 1x1, DBS_XRES=1 forces the recomputation (CUDA-core kernel where the MMA one does not apply).
  static const int xres_env = [] { const char* e = getenv("DBS_XRES"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  static const bool xres_mma_off = [] { const char* e = getenv("DBS_XRES_MMA"); return e && e[0] == '0'; }();
  const bool xres_mma_ok = bw.cout == 64 && W % 16 == 0 && !xres_mma_off;
  const bool xres = c.use_tc && sizeof(T) == 2 && bw.has_res && c.xin && c.plan->enc0_wc && &bw == &c.plan->enc[0] && !c.tap &&
                    (xres_env == 1 || (xres_env == -1 && xres_mma_ok)) && C1 == 0 && bw.cout % 8 == 0 && bw.cout <= 256 && G == 8;
  bool res_done = false;
  if (c.use_tc && sizeof(T) == 2 && bw.has_res && !xres) {
    // conv1 and the 1x1 residual conv read the same input: one pass of the strip kernel does both
    rc = conv_strip_launch(g1, (const __nv_bfloat16*)in0, (const __nv_bfloat16*)in1, bw.conv1.w16, bw.conv1.bias,
                           (__nv_bfloat16*)ty, c.stats, G, c.st, bw.res.w16, bw.res.bias, (__nv_bfloat16*)tr);
    if (rc == DBS_OK) { res_done = true; sd = true; }
    else if (rc != DBS_ERR_UNSUPPORTED) return rc;
  }
  if (c.compose0 && &bw == &c.plan->enc[0]) {
    DBS_REQUIRE(xres && sizeof(T) == 2 && bw.cout == 64 && W % 16 == 0 && H % 4 == 0 && c.plan->enc0_bfrag,
                "unet_forward: composed first layer without its preconditions");
    int TH = 2048 / W / 4 * 4;         // rows per CTA: whole 4 x 4-pixel tiles
    if (TH > H) TH = H;
    if (TH < 4) TH = 4;
    const size_t smem = (size_t)(6 + 2) * 8 * 32 * 8 + (size_t)8 * 16 * 72 * 2 + (size_t)(TH + 4) * (W + 4) * 4;
    static bool attr_done = false;
    if (!attr_done) {
      DBS_CUDA(cudaFuncSetAttribute(enc0_composed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_done = true;
    }
    DBS_REQUIRE(smem <= 96 * 1024, "unet_forward: composed first layer: x tile of %d rows does not fit", TH);
    prof_set_label("enc0 conv1 composed with init conv H%d W%d", H, W);
    ProfScope prof(kProfConvTC, c.st);
    enc0_composed_kernel<<<dim3((unsigned)((H + TH - 1) / TH), (unsigned)c.B), 256, smem, c.st>>>(
        c.xin, (__nv_bfloat16*)ty, c.stats, c.plan->enc0_bfrag, H, W, TH);
    DBS_LAUNCHED();
    res_done = true; sd = true;
  }
  if (!res_done && (rc = run_conv<T>(c, g1, in0, in1, bw.conv1, ty, c.stats, G, &sd))) return rc;
  {
    std::string nm = std::string(name) + ".conv1";
    if ((rc = maybe_tap<T>(c, nm.c_str(), ty, c.B, H, W, bw.cout))) return rc;
  }
  // GroupNorm prologue: where the strip kernel keeps conv2's weights resident, conv2 reads conv1's RAW output and its
  // transform warps normalise the strips in shared memory (SiLU(GN1) + time bias): the GN1 activation and its gn_apply
  // launch vanish.  Needs a fresh statistics slice for conv2 while conv1's is still unread.
  float* const stats1 = c.stats;
  bool gn1_fused = false;
  if (c.use_tc && sizeof(T) == 2 && sd && !c.tap && G == 8 && c.next_stats()) {
    GnFuse pre{bw.g1, bw.b1, c.tbias + bw.tb_off, c.n_t == 1 ? 0 : (int64_t)c.plan->sumC, nullptr};
    rc = conv_strip_launch(g2, (const __nv_bfloat16*)ty, nullptr, bw.conv2.w16, bw.conv2.bias, (__nv_bfloat16*)th, c.stats, G,
                           c.st, nullptr, nullptr, nullptr, &pre, stats1);     // raw conv2 output -> th
    if (rc == DBS_OK) { gn1_fused = true; sd = true; }
    else if (rc != DBS_ERR_UNSUPPORTED) return rc;
  }
  T* const y2 = gn1_fused ? th : ty;       // raw conv2 output
  if (!gn1_fused) {
    // (a declined prologue attempt has already moved c.stats on to conv2's slice: GN1 reads conv1's)
    float* const stats2 = c.stats;
    c.stats = stats1;
    rc = run_gn<T>(c, ty, th, sd, bw.g1, bw.b1, bw.tb_off, true, (const T*)nullptr, H, W, bw.cout);
    c.stats = stats2;
    if (rc) return rc;
    if (c.use_tc && stats2 == stats1 && !c.next_stats()) DBS_CUDA(cudaMemsetAsync(c.stats, 0, stat_bytes, c.st));
    if ((rc = run_conv<T>(c, g2, th, (const T*)nullptr, bw.conv2, ty, c.stats, G, &sd))) return rc;
  }
  if (xres) {
    // conv2 ran on a kernel without the statistics epilogue (a lattice the strip / tile kernels do not tile, e.g. 24x48): a
    // stand-alone pass delivers (mean, rstd) and the apply kernels are told so through inv_cnt = 0.  Falling through to
    // the generic tail instead would read a residual tensor this path never writes.
    if (!sd) {
      ProfScope prof(kProfGnStats, c.st);
      gn_stats_kernel<T><<<(unsigned)(c.B * G), 256, 0, c.st>>>(y2, c.stats, H * W, bw.cout, G, 1e-5f);
      DBS_LAUNCHED();
    }
    const int cv = bw.cout / 8;
    const int threads = (256 / cv) * cv;
    static const int un_env = [] { const char* e = getenv("DBS_XRES_UN"); return e ? atoi(e) : 4; }();
    static const int px_env = [] { const char* e = getenv("DBS_XRES_PX"); return e ? atoi(e) : 2048; }();   // 512: 273 us, 2048: 249 us
    int TH = std::max(1, px_env / W);          // ~512 pixels per CTA
    if (TH > H) TH = H;
    const size_t smem = ((size_t)10 * bw.cout + (size_t)(TH + 2) * (W + 2)) * 4;
    prof_set_label("HW%d C%d xres", H * W, bw.cout);
    ProfScope prof(kProfGnApply, c.st);
    const dim3 xgrid((unsigned)((H + TH - 1) / TH), (unsigned)c.B);
    const float xinv = sd ? 1.0f / (float)((int64_t)H * W * (bw.cout / 8)) : 0.f;
    static const bool mma_off = [] { const char* e = getenv("DBS_XRES_MMA"); return e && e[0] == '0'; }();
    if (bw.cout == 64 && W % 16 == 0 && !mma_off) {
      const size_t smem_m = (size_t)8 * 16 * 72 * 2 + (size_t)(TH + 2) * (W + 2) * 4;
      static bool attr_done = false;
      if (!attr_done) {
        DBS_CUDA(cudaFuncSetAttribute(gn_apply_xres_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_done = true;
      }
      if (smem_m <= 64 * 1024) {
        gn_apply_xres_mma_kernel<<<xgrid, 256, smem_m, c.st>>>((const __nv_bfloat16*)y2, (__nv_bfloat16*)out, c.stats, xinv, 1e-5f, bw.g2,
                                                              bw.b2, c.xin, c.plan->enc0_wc, c.plan->enc0_bc, H, W, TH);
        DBS_LAUNCHED();
        return DBS_OK;
      }
    }
    if (un_env <= 1)
      gn_apply_xres_kernel<1><<<xgrid, threads, smem, c.st>>>((const __nv_bfloat16*)y2, (__nv_bfloat16*)out, c.stats, xinv, 1e-5f, bw.g2,
                                                             bw.b2, c.xin, c.plan->enc0_wc, c.plan->enc0_bc, H, W, bw.cout, TH);
    else if (un_env == 2)
      gn_apply_xres_kernel<2><<<xgrid, threads, smem, c.st>>>((const __nv_bfloat16*)y2, (__nv_bfloat16*)out, c.stats, xinv, 1e-5f, bw.g2,
                                                             bw.b2, c.xin, c.plan->enc0_wc, c.plan->enc0_bc, H, W, bw.cout, TH);
    else
      gn_apply_xres_kernel<4><<<xgrid, threads, smem, c.st>>>((const __nv_bfloat16*)y2, (__nv_bfloat16*)out, c.stats, xinv, 1e-5f, bw.g2,
                                                             bw.b2, c.xin, c.plan->enc0_wc, c.plan->enc0_bc, H, W, bw.cout, TH);
    DBS_LAUNCHED();
    return DBS_OK;
  }
  const T* res = in0;
  if (bw.has_res) {
    if (!res_done) {
      ConvGeom gr = geom_conv((int)c.B, H, W, C0, C1, bw.cout, 1, 1);
      if ((rc = run_conv<T>(c, gr, in0, in1, bw.res, tr, nullptr, G, nullptr))) return rc;
    }
    res = tr;
  }
  if (c.defer && c.defer->want && sd && !c.tap) {
    c.defer->done = true; c.defer->y = y2; c.defer->res = res; c.defer->gamma = bw.g2; c.defer->beta = bw.b2;
    c.defer->stats = c.stats;
    return DBS_OK;
  }
  if ((rc = run_gn<T>(c, y2, out, sd, bw.g2, bw.b2, 0, false, res, H, W, bw.cout))) return rc;
  return maybe_tap<T>(c, name, out, c.B, H, W, bw.cout);
}

struct Sizes {
  int64_t max_act = 0;   // elements of the largest activation tensor
  int64_t total = 0;     // workspace bytes
};

template <typename T>
int forward_impl(dbs_unet_plan* plan, const float* x, const float* t, int64_t n_t, float* score, int64_t B, int H,
                 int W, bool use_tc, void* workspace, int64_t ws_bytes, const char* tap, float* tap_out,
                 int64_t tap_numel, cudaStream_t st, int64_t* need_bytes) {
  const dbs_unet_config& cfg = plan->cfg;
  const int nb = cfg.num_blocks, C = cfg.base_channels, Td = cfg.time_embed_dim;
  Bump ws{(char*)workspace, ws_bytes};
  const bool dry = need_bytes != nullptr;

  // ---- workspace carve-up (identical in the dry run that sizes it)
  float* tbias = (float*)ws.take((int64_t)n_t * plan->sumC * 4);
  const int n_stat_slices = 2 * (2 * nb + 1);          // two GroupNorms per ConvBlock, 2 nb + 1 blocks
  float* stats = (float*)ws.take((int64_t)n_stat_slices * B * 8 * 2 * 4);
  // a lattice side that is not divisible by 2^num_blocks makes some transposed conv overshoot its skip connection by one
  // row / column: that level's output is resized (bilinear fix-up below); buffers are sized for the overshoot
  const bool fixup = (H % (1 << nb)) != 0 || (W % (1 << nb)) != 0;
  auto even_up = [](int v) { return (v + 1) & ~1; };
  int64_t max_act = (int64_t)B * even_up(H) * even_up(W) * C * 2;  // enc0 output has 2C channels at full resolution
  {
    int h = H, w = W, ch = C;
    for (int i = 0; i < nb; ++i) {
      const int co = C << (i + 1);
      max_act = std::max(max_act, (int64_t)B * even_up(h) * even_up(w) * co);
      h = (h - 1) / 2 + 1; w = (w - 1) / 2 + 1; ch = co;
    }
    (void)ch;
  }
  // the tcgen05 / TMA kernels and the fused init / final kernels are built for the divisible lattices every BASELINE
  // config uses; the fix-up shapes run the generic CUDA-core kernels of the same precision
  if (fixup) use_tc = false;
  T* a0 = (T*)ws.take(B * H * W * C * (int64_t)sizeof(T));
  T* ty = (T*)ws.take(max_act * (int64_t)sizeof(T));
  T* th = (T*)ws.take(max_act * (int64_t)sizeof(T));
  T* tr = (T*)ws.take(max_act * (int64_t)sizeof(T));
  // three rotating buffers: with sub-batch chunking a level's output must not alias its input
  T* pp[3] = {(T*)ws.take(max_act * (int64_t)sizeof(T)), (T*)ws.take(max_act * (int64_t)sizeof(T)),
              (T*)ws.take(max_act * (int64_t)sizeof(T))};
  std::vector<T*> skip(nb);
  std::vector<int> sh(nb), sw_(nb), sc(nb);
  {
    int h = H, w = W;
    for (int i = 0; i < nb; ++i) {
      sc[i] = C << (i + 1); sh[i] = h; sw_[i] = w;
      skip[i] = (T*)ws.take((int64_t)B * h * w * sc[i] * (int64_t)sizeof(T));
      h = (h - 1) / 2 + 1; w = (w - 1) / 2 + 1;
    }
  }
  if (dry) { *need_bytes = ws.off + 1024; return DBS_OK; }
  DBS_REQUIRE(workspace != nullptr && ws.off <= ws_bytes, "unet_forward: workspace too small (%lld < %lld bytes)",
              (long long)ws_bytes, (long long)ws.off);

  Ctx c{plan, st, B, (int)n_t, tbias, stats, use_tc, tap, tap_out, tap_numel};
  int rc;
  int stats_next = 0;
  if (use_tc && sizeof(T) == 2) {
    c.stats_base = stats; c.stats_stride = B * 8 * 2; c.stats_next = &stats_next; c.stats_slices = n_stat_slices;
    DBS_CUDA(cudaMemsetAsync(stats, 0, (size_t)n_stat_slices * B * 8 * 2 * 4, st));
  }
  // fused final conv (bf16 path): C in {16, 32, 64}, power-of-two width <= 128, tile within 48 KB of shared memory
  const bool final_fused_ok = sizeof(T) == 2 && (C == 16 || C == 32 || C == 64) && W <= 128 && W >= 4 && (W & (W - 1)) == 0 &&
                              (size_t)((3 * W + 15) & ~15) * (C * 2 + 40) <= 72 * 1024 && !fixup;
  static const bool defer_off = [] { const char* e = getenv("DBS_FINAL_FUSE"); return e && e[0] == '0'; }();
  DeferredGn defer;
  defer.want = final_fused_ok && use_tc && !tap && !defer_off;

  // ---- time path: its own launch, unless the bf16 init conv below can carry it (all samples share t)
  TimeArgs ta{t, plan->freqs, plan->tw0, plan->tb0, plan->tw2, plan->tb2, plan->twt, plan->tbt, tbias, Td, plan->sumC, 0};
  static const bool tfuse_off = [] { const char* e = getenv("DBS_TIME_FUSE"); return e && e[0] == '0'; }();
  // First block's conv1 composed with the init conv (enc0_composed_kernel): the same preconditions as the recomputed
  // residual of that block (run_block: xres through the MMA kernel), so that nothing reads the init conv's activation.
  // DBS_ENC0_COMPOSED=0 keeps the two launches.
  static const bool compose_off = [] { const char* e = getenv("DBS_ENC0_COMPOSED"); return e && e[0] == '0'; }();
  static const bool xres_env_off = [] { const char* e = getenv("DBS_XRES"); return e && e[0] == '0'; }();
  static const bool xres_mma_env_off = [] { const char* e = getenv("DBS_XRES_MMA"); return e && e[0] == '0'; }();
  const bool compose0 = use_tc && sizeof(T) == 2 && !tap && !fixup && !compose_off && !xres_env_off && !xres_mma_env_off && C == 32 &&
                        nb >= 1 && plan->enc0_bfrag && plan->enc0_wc && plan->enc[0].has_res && W % 16 == 0 && H % 4 == 0 && getenv("DBS_L2_CHUNK_MB") == nullptr;
  const bool time_fused = sizeof(T) == 2 && (C == 16 || C == 32 || C == 64) && n_t == 1 && !tfuse_off && !fixup && !compose0;
  if (!time_fused) {
    prof_set_label("time");
    ProfScope prof(kProfUnetMisc, st);
    time_kernel<<<dim3((unsigned)n_t, n_t == 1 ? 16u : 2u), 256, (size_t)4 * Td * 4, st>>>(ta);
    DBS_LAUNCHED();
  } else {
    ta.n_ctas = 16;
  }

  // ---- init conv
  c.compose0 = compose0;
  if (!compose0) {
    const int64_t n_pix = B * H * W;
    prof_set_label(time_fused ? "time+init_conv" : "init_conv");
    ProfScope prof(kProfUnetMisc, st);
    const int64_t nthr = n_pix * (C / (16 / (int)sizeof(T)));
    if (sizeof(T) == 2 && (C == 16 || C == 32 || C == 64) && !fixup) {
      if (C == 16) rc = launch_init_fast<16>(plan->init_w_host, plan->init_b_host, x, (__nv_bfloat16*)a0, n_pix, H, W, st, ta);
      else if (C == 32) rc = launch_init_fast<32>(plan->init_w_host, plan->init_b_host, x, (__nv_bfloat16*)a0, n_pix, H, W, st, ta);
      else rc = launch_init_fast<64>(plan->init_w_host, plan->init_b_host, x, (__nv_bfloat16*)a0, n_pix, H, W, st, ta);
      if (rc) return rc;
    } else {
      init_conv_kernel<T><<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(x, plan->init_w, plan->init_b, a0, n_pix, H, W, C);
      DBS_LAUNCHED();
    }
    if ((rc = maybe_tap<T>(c, "init", a0, B, H, W, C))) return rc;
  }

  // L2-resident chunking: at the high-resolution levels one activation tensor of the whole batch (67-134 MB
  // at B = 1024, 32x32) exceeds the 126 MB L2, so every conv -> GroupNorm -> conv hand-off round-trips HBM.
  // Running a level's layer sequence over sub-batches whose tensors fit in L2 together keeps those
  // hand-offs on chip.  Low-resolution levels stay whole (they need the tiles to fill 148 SMs).
  // MEASURED (B200, B = 1024, 32x32, CUDA-graph forward): 1.94 ms unchunked vs 2.45 / 2.98 ms with 48 / 24 MB
  // chunks — per-launch fixed costs (CTA prologue, re-loading the resident weights, tails) outweigh the
  // L2 hits, so this is OFF by default (DBS_L2_CHUNK_MB=<MB> enables it for experiments).
  static const double chunk_mb = [] { const char* e = getenv("DBS_L2_CHUNK_MB"); return e ? atof(e) : 0.0; }();
  auto chunk_of = [&](int hh, int ww, int cmax) -> int64_t {
    if (chunk_mb <= 0 || tap || fixup) return B;
    const double per_sample = (double)hh * ww * cmax * sizeof(T);
    int64_t cb = (int64_t)(chunk_mb * 1048576.0 / per_sample);
    const int64_t min_cb = (2 * 148 * 128 + (int64_t)hh * ww - 1) / ((int64_t)hh * ww);   // >= 2 waves of tiles
    if (cb < min_cb) cb = min_cb;
    if (cb >= B) return B;
    const int64_t n_chunks = (B + cb - 1) / cb;
    return (B + n_chunks - 1) / n_chunks;    // balanced chunks
  };
  auto sub_ctx = [&](int64_t b0, int64_t nbc) {
    Ctx cc = c;
    cc.B = nbc;
    cc.tbias = c.tbias + (n_t == 1 ? 0 : b0 * plan->sumC);
    return cc;
  };

  // ---- encoder
  const T* cur = a0;
  int ch = C, h = H, w = W;
  int ci = -1;   // index of the rotating buffer that holds `cur` (-1: none)
  auto pick = [&](int avoid0, int avoid1) { for (int k = 0; k < 3; ++k) if (k != avoid0 && k != avoid1) return k; return 0; };
  char name[32];
  for (int i = 0; i < nb; ++i) {
    snprintf(name, sizeof(name), "enc%d", i);
    const int co = sc[i];
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    const int di = pick(ci, -1);
    T* dst = pp[di];
    const int64_t cb = chunk_of(h, w, co);
    for (int64_t b0 = 0; b0 < B; b0 += cb) {
      const int64_t nbc = std::min<int64_t>(cb, B - b0);
      Ctx cc = sub_ctx(b0, nbc);
      const T* in_c = cur + b0 * h * w * ch;
      T* skip_c = skip[i] + b0 * h * w * co;
      cc.xin = (i == 0 && !fixup) ? x + b0 * h * w : nullptr;    // block 0's input is the init conv of x
      if ((rc = run_block<T>(cc, plan->enc[i], name, in_c, ch, (const T*)nullptr, 0, h, w, ty, th, tr, skip_c))) return rc;
      ConvGeom gd = geom_conv((int)nbc, h, w, co, 0, co, 3, 2);
      static const bool s2d_off = [] { const char* e = getenv("DBS_STRIP_S2D"); return e && e[0] == '0'; }();
      rc = DBS_ERR_UNSUPPORTED;
      if (cc.use_tc && sizeof(T) == 2 && !s2d_off && plan->down[i].w16_s2d && !(h & 1) && !(w & 1))
        rc = conv_strip_launch(geom_conv_s2d((int)nbc, h, w, co, co), (const __nv_bfloat16*)skip_c, nullptr, plan->down[i].w16_s2d,
                               plan->down[i].bias, (__nv_bfloat16*)(dst + b0 * ho * wo * co), nullptr, 8, cc.st);
      if (rc != DBS_OK && rc != DBS_ERR_UNSUPPORTED) return rc;
      if (rc == DBS_ERR_UNSUPPORTED &&
          (rc = run_conv<T>(cc, gd, skip_c, (const T*)nullptr, plan->down[i], dst + b0 * ho * wo * co, nullptr, 8, nullptr)))
        return rc;
      c.tapped |= cc.tapped;
    }
    ch = co; h = ho; w = wo;
    snprintf(name, sizeof(name), "down%d", i);
    if ((rc = maybe_tap<T>(c, name, dst, B, h, w, ch))) return rc;
    cur = dst; ci = di;
  }
  // ---- middle
  {
    const int di = pick(ci, -1);
    T* dst = pp[di];
    if ((rc = run_block<T>(c, plan->mid, "mid", cur, ch, (const T*)nullptr, 0, h, w, ty, th, tr, dst))) return rc;
    cur = dst; ci = di;
  }
  // ---- decoder
  for (int i = 0; i < nb; ++i) {
    const int si = nb - 1 - i;
    const int hs = sh[si], wsk = sw_[si];      // the skip connection's size: what the level runs at
    const bool resize = 2 * h != hs || 2 * w != wsk;
    DBS_REQUIRE(!resize || (fixup && ci >= 0), "unet: internal size mismatch at decoder level %d", i);
    const int ui = pick(ci, -1), di = pick(ci, ui);
    T* up = pp[ui];
    T* dst = pp[di];
    const int co = plan->dec[i].cout;
    const int h2 = 2 * h, w2 = 2 * w;
    const int64_t cb = chunk_of(h2, w2, ch + sc[si]);
    T* fixed = resize ? pp[ci] : nullptr;     // the level's input buffer is dead once the transposed conv has read it
    snprintf(name, sizeof(name), "dec%d", i);
    for (int64_t b0 = 0; b0 < B; b0 += cb) {
      const int64_t nbc = std::min<int64_t>(cb, B - b0);
      Ctx cc = sub_ctx(b0, nbc);
      T* up_c = up + b0 * h2 * w2 * ch;
      // ConvTranspose2d(4, 2, 1) = four stride-1 phase convs.  Strip-kernel levels run them as two launches, each the
      // pair of horizontal phases of one vertical phase (shared A rows, four tap slabs per MMA, two outputs per tile);
      // elsewhere the tile kernel takes all four phases in ONE persistent launch (low-resolution phases are less
      // than a wave each: 4 launches of 17-23 us -> one)
      bool up_done = false;
      if (cc.use_tc && sizeof(T) == 2) {
        const __nv_bfloat16* uin = (const __nv_bfloat16*)(cur + b0 * h * w * ch);
        static const bool pair_off = [] { const char* e = getenv("DBS_STRIP_PHASE_PAIR"); return e && e[0] == '0'; }();
        rc = pair_off ? DBS_ERR_UNSUPPORTED
                      : conv_strip_launch(geom_convT_pair((int)nbc, h, w, ch, ch, 0), uin, nullptr, plan->up[i].w16,
                                          plan->up[i].bias, (__nv_bfloat16*)up_c, nullptr, 8, cc.st);
        if (rc == DBS_OK) {
          if ((rc = conv_strip_launch(geom_convT_pair((int)nbc, h, w, ch, ch, 1), uin, nullptr, plan->up[i].w16,
                                      plan->up[i].bias, (__nv_bfloat16*)up_c, nullptr, 8, cc.st)))
            return rc;
          up_done = true;
        } else if (rc != DBS_ERR_UNSUPPORTED) {
          return rc;
        } else {
          rc = conv_strip_launch(geom_convT_phase((int)nbc, h, w, ch, ch, 0, 0), uin, nullptr, plan->up[i].w16,
                                 plan->up[i].bias, (__nv_bfloat16*)up_c, nullptr, 8, cc.st);
          if (rc == DBS_OK) {
            for (int ph4 = 1; ph4 < 4; ++ph4) {
              const ConvGeom gu = geom_convT_phase((int)nbc, h, w, ch, ch, ph4 >> 1, ph4 & 1);
              if ((rc = run_conv<T>(cc, gu, cur + b0 * h * w * ch, (const T*)nullptr, plan->up[i], up_c, nullptr, 8, nullptr)))
                return rc;
            }
            up_done = true;
          } else if (rc != DBS_ERR_UNSUPPORTED) {
            return rc;
          } else {
            rc = conv_tc_launch(geom_convT_all((int)nbc, h, w, ch, ch), uin, nullptr, plan->up[i].w16, plan->up[i].bias,
                                (__nv_bfloat16*)up_c, nullptr, 8, cc.st);
            if (rc == DBS_OK) up_done = true;
            else if (rc != DBS_ERR_UNSUPPORTED) return rc;
          }
        }
      }
      if (!up_done)
        for (int ph = 0; ph < 2; ++ph)
          for (int pw = 0; pw < 2; ++pw) {
            ConvGeom gu = geom_convT_phase((int)nbc, h, w, ch, ch, ph, pw);
            if ((rc = run_conv<T>(cc, gu, cur + b0 * h * w * ch, (const T*)nullptr, plan->up[i], up_c, nullptr, 8, nullptr)))
              return rc;
          }
      if (cb >= B) {
        snprintf(name, sizeof(name), "up%d", i);
        if ((rc = maybe_tap<T>(cc, name, up, B, h2, w2, ch))) return rc;
        snprintf(name, sizeof(name), "dec%d", i);
      }
      if (resize) {   // cb == B here (fix-up lattices are never chunked)
        const int64_t quads = nbc * hs * wsk * (ch / 4);
        prof_set_label("resize");
        ProfScope prof(kProfUnetMisc, st);
        bilinear_resize_kernel<T><<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(
            up_c, fixed, quads, h2, w2, hs, wsk, ch, (float)h2 / (float)hs, (float)w2 / (float)wsk);
        DBS_LAUNCHED();
        up_c = fixed;
      }
      cc.defer = (i == nb - 1 && cb >= B) ? &defer : nullptr;
      if ((rc = run_block<T>(cc, plan->dec[i], name, up_c, ch, skip[si] + b0 * hs * wsk * sc[si], sc[si], hs, wsk, ty, th, tr,
                             dst + b0 * hs * wsk * co)))
        return rc;
      c.tapped |= cc.tapped;
    }
    h = hs; w = wsk;
    ch = co;
    cur = dst; ci = di;
  }
  // ---- final conv (+ optional normalisation, output scale)
  {
    const int64_t n_pix = B * H * W;
    prof_set_label("final_conv");
    {
      const int cvf = C / (16 / (int)sizeof(T));
      const bool cv_pow2 = C % (16 / (int)sizeof(T)) == 0 && cvf >= 1 && cvf <= 32 && (cvf & (cvf - 1)) == 0;
      ProfScope prof(kProfUnetMisc, st);
      const int64_t nthr = n_pix * cvf;
      if (final_fused_ok) {
        const __nv_bfloat16* yin = defer.done ? (const __nv_bfloat16*)defer.y : (const __nv_bfloat16*)cur;
        const __nv_bfloat16* rin = defer.done ? (const __nv_bfloat16*)defer.res : nullptr;
        if (C == 16) rc = launch_final_fused<16>(plan->final_bfrag, plan->final_b_host, cfg.output_scale, yin, rin, defer.stats, defer.gamma, defer.beta, score, B, H, W, st);
        else if (C == 32) rc = launch_final_fused<32>(plan->final_bfrag, plan->final_b_host, cfg.output_scale, yin, rin, defer.stats, defer.gamma, defer.beta, score, B, H, W, st);
        else rc = launch_final_fused<64>(plan->final_bfrag, plan->final_b_host, cfg.output_scale, yin, rin, defer.stats, defer.gamma, defer.beta, score, B, H, W, st);
        if (rc) return rc;
      } else if (cv_pow2) {
        final_conv_kernel<T><<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(cur, plan->final_w, plan->final_b, score, n_pix,
                                                                            H, W, C, cfg.output_scale);
        DBS_LAUNCHED();
      } else {
        final_conv_generic_kernel<T><<<(unsigned)((n_pix + 255) / 256), 256, 0, st>>>(cur, plan->final_w, plan->final_b, score,
                                                                                     n_pix, H, W, C, cfg.output_scale);
        DBS_LAUNCHED();
      }
    }
    if (cfg.normalize_output) {
      normalize_output_kernel<<<(unsigned)B, 256, 0, st>>>(score, H * W, cfg.output_scale);
      DBS_LAUNCHED();
    }
  }
  if (tap && !c.tapped) {
    set_error("unet tap '%s' not found", tap);
    return DBS_ERR_INVALID;
  }
  return DBS_OK;
}

int check_forward_args(const dbs_unet_plan* plan, int64_t B, int H, int W, int precision) {
  DBS_REQUIRE(plan != nullptr, "unet: null plan");
  DBS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "unet: bad shape B=%lld H=%d W=%d", (long long)B, H, W);
  DBS_REQUIRE(precision == 0 || precision == 1, "unet: precision must be 0 (fp32) or 1 (bf16 tcgen05)");
  DBS_REQUIRE(precision == 0 || plan->cfg.base_channels % 8 == 0,
              "unet: the bf16 path needs base_channels to be a multiple of 8 (got %d)", plan->cfg.base_channels);
  return DBS_OK;
}

}  // namespace
}  // namespace dbs

using namespace dbs;

extern "C" int dbs_unet_plan_create(const dbs_unet_config* cfg, const dbs_param* params, int n_params,
                                    const float* freqs_host, dbs_unet_plan** out) {
  DBS_REQUIRE(cfg && params && out && freqs_host, "unet_plan_create: null argument");
  DBS_REQUIRE(cfg->in_channels == 1, "unet: only in_channels == 1 (Ising lattices) is built, got %d", cfg->in_channels);
  DBS_REQUIRE(cfg->base_channels >= 4 && cfg->base_channels % 4 == 0, "unet: base_channels must be a multiple of 4");
  DBS_REQUIRE(cfg->time_embed_dim >= 4 && cfg->time_embed_dim % 4 == 0 && cfg->time_embed_dim <= 1024,
              "unet: time_embed_dim must be a multiple of 4, 4..1024");
  DBS_REQUIRE(cfg->num_blocks >= 1 && cfg->num_blocks <= 6, "unet: num_blocks must be 1..6");
  dbs_unet_plan* plan = new dbs_unet_plan();
  plan->cfg = *cfg;
  const ParamTable pt{params, n_params};
  const int C = cfg->base_channels, T = cfg->time_embed_dim, nb = cfg->num_blocks;
  int rc = DBS_OK;
  auto fail = [&](int code) { dbs_unet_plan_destroy(plan); return code; };

  void* f = nullptr;
  if ((rc = dev_alloc(plan, (size_t)(T / 2) * 4, &f))) return fail(rc);
  plan->freqs = (float*)f;
  if (cudaMemcpy(plan->freqs, freqs_host, (size_t)(T / 2) * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("unet_plan_create: freqs upload failed");
    return fail(DBS_ERR_CUDA);
  }
  if ((rc = upload(plan, pt, "time_mlp.0.weight", (int64_t)2 * T * T, &plan->tw0))) return fail(rc);
  if ((rc = upload(plan, pt, "time_mlp.0.bias", 2 * T, &plan->tb0))) return fail(rc);
  if ((rc = upload(plan, pt, "time_mlp.2.weight", (int64_t)2 * T * T, &plan->tw2))) return fail(rc);
  if ((rc = upload(plan, pt, "time_mlp.2.bias", T, &plan->tb2))) return fail(rc);

  // init / final conv: [co][ci][3][3] -> [tap][C]
  {
    float *raw = nullptr, *rawf = nullptr;
    if ((rc = upload(plan, pt, "init_conv.weight", (int64_t)C * 9, &raw))) return fail(rc);
    if ((rc = upload(plan, pt, "init_conv.bias", C, &plan->init_b))) return fail(rc);
    if ((rc = upload(plan, pt, "final.weight", (int64_t)C * 9, &rawf))) return fail(rc);
    if ((rc = upload(plan, pt, "final.bias", 1, &plan->final_b))) return fail(rc);
    void *iw = nullptr, *fw = nullptr, *scratch = nullptr;
    if ((rc = dev_alloc(plan, (size_t)C * 9 * 4, &iw))) return fail(rc);
    if ((rc = dev_alloc(plan, (size_t)C * 9 * 4, &fw))) return fail(rc);
    if ((rc = dev_alloc(plan, (size_t)C * 9 * 2, &scratch))) return fail(rc);
    plan->init_w = (float*)iw; plan->final_w = (float*)fw;
    // init: co = C, ci = 1 -> [tap][C][1]; final: co = 1, ci = C -> [tap][1][C]
    repack_kernel<<<(unsigned)((C * 9 + 255) / 256), 256>>>(raw, plan->init_w, (__nv_bfloat16*)scratch, C, 1, 9, 0);
    repack_kernel<<<(unsigned)((C * 9 + 255) / 256), 256>>>(rawf, plan->final_w, (__nv_bfloat16*)scratch, 1, C, 9, 0);
    if (cudaGetLastError() != cudaSuccess) { set_error("unet_plan_create: repack launch failed"); return fail(DBS_ERR_CUDA); }
    plan->final_w_host.resize((size_t)C * 9);
    plan->init_w_host.resize((size_t)C * 9);
    plan->init_b_host.resize((size_t)C);
    if (cudaMemcpy(plan->init_w_host.data(), plan->init_w, (size_t)C * 9 * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(plan->init_b_host.data(), plan->init_b, (size_t)C * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(plan->final_w_host.data(), plan->final_w, (size_t)C * 9 * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&plan->final_b_host, plan->final_b, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
      set_error("unet_plan_create: reading back the final conv weights failed");
      return fail(DBS_ERR_CUDA);
    }
    if (C % 16 == 0) {
      const std::vector<uint32_t> bf = final_bfrag_host(plan->final_w_host, C);
      void* d = nullptr;
      if ((rc = dev_alloc(plan, bf.size() * 4, &d))) return fail(rc);
      if (cudaMemcpy(d, bf.data(), bf.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("unet_plan_create: final conv fragment upload failed");
        return fail(DBS_ERR_CUDA);
      }
      plan->final_bfrag = (uint32_t*)d;
    }
  }

  // blocks, with the per-block time Linear layers concatenated into one [sumC][T] matrix
  struct TL { std::string prefix; int cout; };
  std::vector<TL> tls;
  plan->enc.resize(nb); plan->dec.resize(nb); plan->down.resize(nb); plan->up.resize(nb);
  int ch = C, off = 0;
  std::vector<int> enc_ch;
  for (int i = 0; i < nb; ++i) {
    const int co = C << (i + 1);
    const std::string p = "encoder." + std::to_string(i) + ".";
    if ((rc = make_block(plan, pt, p, ch, co, &plan->enc[i]))) return fail(rc);
    plan->enc[i].tb_off = off; off += co; tls.push_back({p, co});
    if ((rc = make_conv(plan, pt, "downsample." + std::to_string(i), co, co, 3, false, &plan->down[i]))) return fail(rc);
    {
      void* w4 = nullptr;
      const int64_t n4 = (int64_t)4 * co * 4 * co;
      if ((rc = dev_alloc(plan, (size_t)n4 * 2, &w4))) return fail(rc);
      plan->down[i].w16_s2d = (__nv_bfloat16*)w4;
      repack_s2d_kernel<<<(unsigned)((n4 + 255) / 256), 256>>>(plan->down[i].w16, plan->down[i].w16_s2d, co, co);
    }
    ch = co; enc_ch.push_back(co);
  }
  {
    // encoder.0's residual projection composed with the init conv: wc[tap][co] = sum_ci Wr[co][ci] Wi[ci][tap],
    // bc[co] = br[co] + sum_ci Wr[co][ci] bi[ci]  (fp32 on the host; see gn_apply_xres_kernel)
    const int co0 = C << 1;
    const dbs_param* wr = pt.find("encoder.0.residual.weight");
    const dbs_param* br = pt.find("encoder.0.residual.bias");
    if (wr && br && wr->numel == (int64_t)co0 * C && br->numel == co0) {
      std::vector<float> hwr((size_t)co0 * C), hbr((size_t)co0), wc((size_t)9 * co0), bc((size_t)co0);
      if (cudaMemcpy(hwr.data(), wr->data, hwr.size() * 4, wr->is_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost) != cudaSuccess ||
          cudaMemcpy(hbr.data(), br->data, hbr.size() * 4, br->is_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost) != cudaSuccess) {
        set_error("unet_plan_create: reading encoder.0.residual failed");
        return fail(DBS_ERR_CUDA);
      }
      for (int o = 0; o < co0; ++o) {
        double b = hbr[o];
        for (int ci = 0; ci < C; ++ci) b += (double)hwr[(size_t)o * C + ci] * plan->init_b_host[ci];
        bc[o] = (float)b;
        for (int t = 0; t < 9; ++t) {
          double a = 0.0;
          for (int ci = 0; ci < C; ++ci) a += (double)hwr[(size_t)o * C + ci] * plan->init_w_host[(size_t)t * C + ci];
          wc[(size_t)t * co0 + o] = (float)a;
        }
      }
      void *dw = nullptr, *db = nullptr;
      if ((rc = dev_alloc(plan, wc.size() * 4, &dw))) return fail(rc);
      if ((rc = dev_alloc(plan, bc.size() * 4, &db))) return fail(rc);
      if (cudaMemcpy(dw, wc.data(), wc.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(db, bc.data(), bc.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("unet_plan_create: composed residual upload failed");
        return fail(DBS_ERR_CUDA);
      }
      plan->enc0_wc = (float*)dw; plan->enc0_bc = (float*)db;
    }
    // conv1 of the first block composed with the init conv (enc0_composed_kernel): Wk[k][co], k = 9 t + s products, then
    // the nine bias taps, laid out as mma.sync B fragments (lane = 4 g + t4: b0 = k 2 t4, 2 t4 + 1; b1 = k + 8; n = g)
    const dbs_param* w1 = pt.find("encoder.0.conv1.weight");
    const dbs_param* bb1 = pt.find("encoder.0.conv1.bias");
    if (C == 32 && w1 && bb1 && w1->numel == (int64_t)co0 * C * 9 && bb1->numel == co0) {
      std::vector<float> hw1((size_t)co0 * C * 9), hb1((size_t)co0), wk((size_t)96 * co0, 0.f);
      if (cudaMemcpy(hw1.data(), w1->data, hw1.size() * 4, w1->is_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost) != cudaSuccess ||
          cudaMemcpy(hb1.data(), bb1->data, hb1.size() * 4, bb1->is_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost) != cudaSuccess) {
        set_error("unet_plan_create: reading encoder.0.conv1 failed");
        return fail(DBS_ERR_CUDA);
      }
      for (int o = 0; o < co0; ++o)
        for (int t = 0; t < 9; ++t) {
          double bsum = 0.0;
          for (int ci = 0; ci < C; ++ci) bsum += (double)hw1[((size_t)o * C + ci) * 9 + t] * plan->init_b_host[ci];
          wk[(size_t)(81 + t) * co0 + o] = (float)bsum;
          for (int s2 = 0; s2 < 9; ++s2) {
            double a = 0.0;
            for (int ci = 0; ci < C; ++ci) a += (double)hw1[((size_t)o * C + ci) * 9 + t] * plan->init_w_host[(size_t)s2 * C + ci];
            wk[(size_t)(9 * t + s2) * co0 + o] = (float)a;
          }
        }
      // interior tiles: the products of one input offset u = t + s summed into a 5x5 kernel (k = 5 uy + ux < 25), k = 25 the
      // summed bias taps, 26 .. 31 zero
      auto bf16_round = [](float v) { const uint32_t b = bf16_bits(v) << 16; float f; memcpy(&f, &b, 4); return f; };
      std::vector<double> w5d((size_t)32 * co0, 0.0);
      for (int o = 0; o < co0; ++o) {
        wk[(size_t)90 * co0 + o] = hb1[o];                               // masked form: b1 as bf16 high + low part
        wk[(size_t)91 * co0 + o] = hb1[o] - bf16_round(hb1[o]);
        w5d[(size_t)25 * co0 + o] = (double)hb1[o];
      }
      for (int o = 0; o < co0; ++o)
        for (int t = 0; t < 9; ++t) {
          w5d[(size_t)25 * co0 + o] += (double)wk[(size_t)(81 + t) * co0 + o];
          for (int s2 = 0; s2 < 9; ++s2)
            w5d[(size_t)(5 * (t / 3 + s2 / 3) + (t % 3 + s2 % 3)) * co0 + o] += (double)wk[(size_t)(9 * t + s2) * co0 + o];
        }
      for (int o = 0; o < co0; ++o)                                      // interior form: b1 + bias taps, high + low part
        w5d[(size_t)26 * co0 + o] = w5d[(size_t)25 * co0 + o] - (double)bf16_round((float)w5d[(size_t)25 * co0 + o]);
      std::vector<uint32_t> frag((size_t)(6 + 2) * 8 * 32 * 2);
      for (int kk = 0; kk < 8; ++kk)
        for (int j = 0; j < 8; ++j)
          for (int lane = 0; lane < 32; ++lane) {
            const int gg = lane >> 2, tt = lane & 3, ch = 8 * j + gg, k0 = 16 * (kk < 6 ? kk : kk - 6) + 2 * tt;
            auto W_ = [&](int k) { return kk < 6 ? wk[(size_t)k * co0 + ch] : (float)w5d[(size_t)k * co0 + ch]; };
            const size_t idx = ((size_t)(kk * 8 + j) * 32 + lane) * 2;
            frag[idx] = bf16_bits(W_(k0)) | (bf16_bits(W_(k0 + 1)) << 16);
            frag[idx + 1] = bf16_bits(W_(k0 + 8)) | (bf16_bits(W_(k0 + 9)) << 16);
          }
      void* df = nullptr;                                                // (conv1's bias rides in the fragments)
      if ((rc = dev_alloc(plan, frag.size() * 4, &df))) return fail(rc);
      if (cudaMemcpy(df, frag.data(), frag.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("unet_plan_create: composed conv1 upload failed");
        return fail(DBS_ERR_CUDA);
      }
      plan->enc0_bfrag = (uint2*)df;
    }
  }
  if ((rc = make_block(plan, pt, "middle.", ch, ch, &plan->mid))) return fail(rc);
  plan->mid.tb_off = off; off += ch; tls.push_back({"middle.", ch});
  for (int k = 0; k < nb; ++k) {
    const int i = nb - 1 - k;
    const int co = i > 0 ? (C << i) : C;
    const std::string p = "decoder." + std::to_string(k) + ".";
    if ((rc = make_conv(plan, pt, "upsample." + std::to_string(k), ch, ch, 4, true, &plan->up[k]))) return fail(rc);
    if ((rc = make_block(plan, pt, p, ch + enc_ch[i], co, &plan->dec[k]))) return fail(rc);
    plan->dec[k].tb_off = off; off += co; tls.push_back({p, co});
    ch = co;
  }
  plan->sumC = off;
  {
    void *wt = nullptr, *bt = nullptr;
    if ((rc = dev_alloc(plan, (size_t)off * T * 4, &wt))) return fail(rc);
    if ((rc = dev_alloc(plan, (size_t)off * 4, &bt))) return fail(rc);
    plan->twt = (float*)wt; plan->tbt = (float*)bt;
    int o = 0;
    for (const TL& tl : tls) {
      const dbs_param* w = pt.find(tl.prefix + "time_mlp.weight");
      const dbs_param* b = pt.find(tl.prefix + "time_mlp.bias");
      if (!w || !b || w->numel != (int64_t)tl.cout * T || b->numel != tl.cout) {
        set_error("unet plan: parameter '%stime_mlp.*' missing or mis-sized", tl.prefix.c_str());
        return fail(DBS_ERR_INVALID);
      }
      const cudaMemcpyKind kw = w->is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
      const cudaMemcpyKind kb = b->is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
      if (cudaMemcpy(plan->twt + (size_t)o * T, w->data, (size_t)tl.cout * T * 4, kw) != cudaSuccess ||
          cudaMemcpy(plan->tbt + o, b->data, (size_t)tl.cout * 4, kb) != cudaSuccess) {
        set_error("unet plan: time_mlp upload failed");
        return fail(DBS_ERR_CUDA);
      }
      o += tl.cout;
    }
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("unet_plan_create: device error after repack"); return fail(DBS_ERR_CUDA); }
  *out = plan;
  return DBS_OK;
}

extern "C" void dbs_unet_plan_destroy(dbs_unet_plan* plan) {
  if (!plan) return;
  plan->arena.release();
  delete plan;
}

extern "C" int64_t dbs_unet_workspace_bytes(const dbs_unet_plan* plan, int64_t B, int H, int W, int precision) {
  if (plan && B == 0) return 0;   // empty batch
  if (check_forward_args(plan, B, H, W, precision)) return -1;
  int64_t need = 0;
  // n_t = B sizes the time-bias block for the general per-sample-t call
  int rc = precision == 0
               ? forward_impl<float>(const_cast<dbs_unet_plan*>(plan), nullptr, nullptr, B, nullptr, B, H, W, false,
                                     nullptr, 0, nullptr, nullptr, 0, nullptr, &need)
               : forward_impl<__nv_bfloat16>(const_cast<dbs_unet_plan*>(plan), nullptr, nullptr, B, nullptr, B, H, W,
                                             true, nullptr, 0, nullptr, nullptr, 0, nullptr, &need);
  return rc ? -1 : need;
}

static int unet_forward_common(dbs_unet_plan* plan, const float* x, const float* t, int64_t n_t, float* score,
                               int64_t B, int H, int W, int precision, void* workspace, int64_t workspace_bytes,
                               const char* tap, float* tap_out, int64_t tap_numel, void* stream) {
  if (plan && B == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  int rc = check_forward_args(plan, B, H, W, precision);
  if (rc) return rc;
  DBS_REQUIRE(x && t && score, "unet_forward: null tensor");
  DBS_REQUIRE(n_t == 1 || n_t == B, "unet_forward: n_t must be 1 or B");
  // carve the workspace exactly as dbs_unet_workspace_bytes did (time-bias block sized for B rows)
  if (precision == 0)
    return forward_impl<float>(plan, x, t, n_t, score, B, H, W, false, workspace, workspace_bytes, tap, tap_out,
                               tap_numel, as_stream(stream), nullptr);
  return forward_impl<__nv_bfloat16>(plan, x, t, n_t, score, B, H, W, true, workspace, workspace_bytes, tap, tap_out,
                                     tap_numel, as_stream(stream), nullptr);
}

extern "C" int dbs_unet_forward(dbs_unet_plan* plan, const float* x, const float* t, int64_t n_t, float* score,
                                int64_t B, int H, int W, int precision, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  return unet_forward_common(plan, x, t, n_t, score, B, H, W, precision, workspace, workspace_bytes, nullptr, nullptr,
                             0, stream);
}

extern "C" int dbs_unet_forward_tap(dbs_unet_plan* plan, const float* x, const float* t, int64_t n_t, float* score,
                                    int64_t B, int H, int W, int precision, void* workspace, int64_t workspace_bytes,
                                    const char* tap, float* tap_out, int64_t tap_numel, void* stream) {
  DBS_REQUIRE(tap && tap_out, "unet_forward_tap: null tap");
  return unet_forward_common(plan, x, t, n_t, score, B, H, W, precision, workspace, workspace_bytes, tap, tap_out,
                             tap_numel, stream);
}
