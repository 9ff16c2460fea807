// csn.cu — executor of the particle score network for continuous systems.
//
// Replaces (reference file:line, /root/reference/backend/ml/models/score_network_continuous.py):
//   PeriodicPositionEmbedding.forward :36-75   (Fourier features of positions in [-1, 1])
//   SinusoidalTimeEmbedding.forward   :85-101  + ContinuousScoreNetwork.time_mlp :229-233
//   MLPBlock.forward                  :123-136 (fc1 -> LN -> SiLU -> + Linear(t_emb) -> fc2 -> LN -> SiLU -> + x)
//   ParticleAttention.forward         :155-180 (LN -> qkv -> softmax(q k^T / sqrt(hd)) v -> proj -> + x)
//   ContinuousScoreNetwork.forward    :256-282 (embed + raw coords -> input_proj -> blocks -> LN/MLP head)
//
// Data layout: tokens = (configuration, particle) rows, features innermost: fp32 [B*N][width].
// Kernels (fp32 CUDA-core path; this is the parity path of the row, <= 1e-4 relative):
//   csn_time_kernel       one CTA per distinct t: embedding MLP + every block's time Linear
//   csn_linear_kernel     y = epi(pro(x) W^T + b) on a 64-token tile.  pro = identity | LayerNorm |
//                         on-the-fly Fourier embedding; epi = [LayerNorm] -> [SiLU] -> [+ time bias]
//                         -> [+ residual].  The CTA owns whole rows, so both LayerNorms are fused and
//                         no normalised / activated intermediate ever round-trips HBM.
//   csn_attention_kernel  one CTA per (configuration, head): K and V of the head staged in shared
//                         memory, one query per thread, two-pass softmax (max, then exp-sum + PV);
//                         the N x N score matrix never exists in memory.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include <cuda_bf16.h>

#include "common.cuh"

namespace dbs {
// csn_tc.cu: tcgen05 kernels of the bf16 path
int csn_tc_convert_weights(const float* src, __nv_bfloat16* dst, int64_t n);
int csn_tc_linear(const void* X, int in_bf16, const __nv_bfloat16* W, const float* bias, int N_total, const float* pre_g,
                  const float* pre_b, const float* post_g, const float* post_b, int silu, const float* tbias,
                  int64_t tb_stride, int tokens_per_cfg, const float* res, void* out, int out_bf16, int64_t M,
                  cudaStream_t st, const float* pos_freqs = nullptr, const float* head_w2 = nullptr,
                  const float* head_b2 = nullptr, const __nv_bfloat16* W2 = nullptr, const float* bias2 = nullptr,
                  const float* post2_g = nullptr, const float* post2_b = nullptr);
int csn_tc_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int64_t B, int N, int heads, cudaStream_t st);
}  // namespace dbs

struct dbs_csn_plan {
  dbs_csn_config cfg;
  dbs::DevArena arena;
  float *pos_freqs = nullptr, *t_freqs = nullptr;
  float *tw0 = nullptr, *tb0 = nullptr, *tw2 = nullptr, *tb2 = nullptr;
  float *twt = nullptr, *tbt = nullptr;   // every block's time Linear, concatenated: [nb*H][T], [nb*H]
  float *in_w = nullptr, *in_b = nullptr;
  struct Block {
    float *fc1_w, *fc1_b, *fc2_w, *fc2_b, *n1_g, *n1_b, *n2_g, *n2_b;
    float *an_g, *an_b, *qkv_w, *qkv_b, *proj_w, *proj_b;
    __nv_bfloat16 *fc1_h, *fc2_h, *qkv_h, *proj_h;   // bf16 copies for the tcgen05 path (null if not eligible)
  };
  bool tc_ok = false;
  __nv_bfloat16 *in_h = nullptr, *o1_h = nullptr;   // input_proj / output_proj.1 zero-padded to [128][128] bf16
  float* o1_bpad = nullptr;                         // output_proj.1 bias padded to 128
  bool tc_io_ok = false;                            // pos_embed_dim == 64: embedding + head also run on the tensor cores
  std::vector<Block> blocks;
  float *on_g = nullptr, *on_b = nullptr, *o1_w = nullptr, *o1_b = nullptr, *o2_w = nullptr, *o2_b = nullptr;
};

namespace dbs {
namespace {

constexpr int kTM = 64;   // tokens per CTA
constexpr int kTN = 64;   // output columns per pass
constexpr int kKC = 32;   // K chunk of the weights staged in shared memory
constexpr int kLdW = kTN + 4;
constexpr float kLnEps = 1e-5f;

__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }

enum { kProNone = 0, kProLN = 1, kProEmbed = 2 };

struct LinArgs {
  const float* X;       // [M][K] rows (kProNone / kProLN) or positions [M][2] (kProEmbed)
  int K;                // reduction length (kProEmbed: 4 * n_freq + 2)
  const float* W;       // [N][K]
  const float* bias;    // [N]
  int N;
  const float *pre_g, *pre_b;     // kProLN: LayerNorm over K applied to the input rows
  const float* pos_freqs;         // kProEmbed
  int n_freq;
  const float *post_g, *post_b;   // LayerNorm over N applied to the output rows (null = none)
  int silu;
  const float* tbias;             // time bias [n_t][tb_stride] (+ column), null = none
  int64_t tb_stride;              // 0 when all configurations share t
  int tokens_per_cfg;
  const float* res;               // [M][N] residual, null = none
  float* out;                     // [M][N]
  int64_t M;
  int cols_per_cta;               // columns one CTA produces (N when a post-LayerNorm needs whole rows)
};

template <int PRO>
__global__ void __launch_bounds__(256)
csn_linear_kernel(const LinArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int K = a.K, N = a.N;
  const int Kp = (K + 3) & ~3;
  const int lda = Kp + 4;
  const int ncols = a.cols_per_cta;
  const int ldo = ncols + 1;
  float* A_s = sm;                       // [kTM][lda]
  float* W_s = A_s + kTM * lda;          // [kKC][kLdW]
  float* O_s = W_s + kKC * kLdW;         // [kTM][ldo]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * kTM;
  const int col0 = blockIdx.y * ncols;
  const int rows = (int)min((int64_t)kTM, a.M - m0);

  // ---- A tile: the CTA's token rows, transformed by the prologue
  if (PRO == kProEmbed) {
    // feature k of a token at (px, py), models/score_network_continuous.py:46-62:
    // [sin(ax f) | cos(ax f) | sin(ay f) | cos(ay f) | px | py], ax = fl(fl(px + 1) * pi)
    const int nf = a.n_freq;
    for (int idx = tid; idx < kTM * Kp; idx += 256) {
      const int r = idx / Kp, k = idx - r * Kp;
      float v = 0.f;
      if (r < rows && k < K) {
        const float px = a.X[(m0 + r) * 2 + 0], py = a.X[(m0 + r) * 2 + 1];
        if (k >= 4 * nf) {
          v = k == 4 * nf ? px : py;
        } else {
          const int q = k / nf, j = k - q * nf;
          const float p = q < 2 ? px : py;
          const float arg = __fmul_rn(__fmul_rn(__fadd_rn(p, 1.0f), 3.14159265358979323846f), a.pos_freqs[j]);
          v = (q & 1) ? cosf(arg) : sinf(arg);
        }
      }
      A_s[r * lda + k] = v;
    }
  } else {
    for (int r = warp; r < kTM; r += 8) {
      const float* xr = a.X + (m0 + r) * (int64_t)K;
      if (r >= rows) {
        for (int k = lane; k < Kp; k += 32) A_s[r * lda + k] = 0.f;
        continue;
      }
      if (PRO == kProLN) {
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s += xr[k];
        const float mean = warp_sum(s) / (float)K;
        float ss = 0.f;
        for (int k = lane; k < K; k += 32) { const float d = xr[k] - mean; ss = fmaf(d, d, ss); }
        const float rstd = 1.0f / sqrtf(warp_sum(ss) / (float)K + kLnEps);
        for (int k = lane; k < Kp; k += 32)
          A_s[r * lda + k] = k < K ? (xr[k] - mean) * rstd * a.pre_g[k] + a.pre_b[k] : 0.f;
      } else {
        for (int k = lane; k < Kp; k += 32) A_s[r * lda + k] = k < K ? xr[k] : 0.f;
      }
    }
  }

  // ---- GEMM: 64 tokens x 64 columns per pass, 4 x 4 outputs per thread
  const int ty = tid >> 4, tx = tid & 15;
  for (int n0 = 0; n0 < ncols; n0 += kTN) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < Kp; k0 += kKC) {
      __syncthreads();   // A tile complete / previous chunk consumed
      for (int idx = tid; idx < kTN * kKC; idx += 256) {
        const int nn = idx / kKC, kk = idx - nn * kKC;
        const int n = col0 + n0 + nn, k = k0 + kk;
        W_s[kk * kLdW + nn] = (n < N && n0 + nn < ncols && k < K) ? __ldg(a.W + (int64_t)n * K + k) : 0.f;
      }
      __syncthreads();
      const int kc = min(kKC, Kp - k0);
      for (int kk = 0; kk < kc; kk += 4) {
        float4 av[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(&A_s[(ty * 4 + i) * lda + k0 + kk]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w4 = *reinterpret_cast<const float4*>(&W_s[(kk + j) * kLdW + tx * 4]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float ai = j == 0 ? av[i].x : (j == 1 ? av[i].y : (j == 2 ? av[i].z : av[i].w));
            acc[i][0] = fmaf(ai, w4.x, acc[i][0]);
            acc[i][1] = fmaf(ai, w4.y, acc[i][1]);
            acc[i][2] = fmaf(ai, w4.z, acc[i][2]);
            acc[i][3] = fmaf(ai, w4.w, acc[i][3]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = n0 + tx * 4 + j;
        if (c < ncols && col0 + c < N) O_s[(ty * 4 + i) * ldo + c] = acc[i][j] + a.bias[col0 + c];
      }
  }
  __syncthreads();

  // ---- epilogue: a warp owns whole rows of the CTA's column range
  const int nc = min(ncols, N - col0);
  for (int r = warp; r < rows; r += 8) {
    const int64_t m = m0 + r;
    float* orow = O_s + r * ldo;
    float mean = 0.f, rstd = 1.f;
    if (a.post_g) {
      float s = 0.f;
      for (int c = lane; c < nc; c += 32) s += orow[c];
      mean = warp_sum(s) / (float)nc;
      float ss = 0.f;
      for (int c = lane; c < nc; c += 32) { const float d = orow[c] - mean; ss = fmaf(d, d, ss); }
      rstd = 1.0f / sqrtf(warp_sum(ss) / (float)nc + kLnEps);
    }
    const float* tb = a.tbias ? a.tbias + (m / a.tokens_per_cfg) * a.tb_stride : nullptr;
    for (int c = lane; c < nc; c += 32) {
      float v = orow[c];
      if (a.post_g) v = (v - mean) * rstd * a.post_g[c] + a.post_b[c];
      if (a.silu) v = silu_acc(v);
      if (tb) v += tb[col0 + c];
      if (a.res) v += a.res[m * N + col0 + c];
      a.out[m * N + col0 + c] = v;
    }
  }
}

// One CTA per (configuration, head).  qkv rows are [3][heads][HD] per token (the reshape at
// score_network_continuous.py:169-171); the head's K and V are staged once per key chunk, every thread
// owns one query.  softmax as torch computes it: p = exp(s - max s) / sum, s = (q . k) * scale.
template <int HD>
__global__ void __launch_bounds__(256)
csn_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int N, int heads, float scale) {
  constexpr int KT = 256;
  extern __shared__ __align__(16) float sm[];
  float* K_s = sm;              // [KT][HD]
  float* V_s = sm + KT * HD;    // [KT][HD]
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
  const int H = heads * HD;
  const int qi = blockIdx.y * 256 + threadIdx.x;
  const bool live = qi < N;
  const float* base = qkv + (int64_t)b * N * 3 * H + h * HD;
  float q[HD];
  if (live) {
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 v = *reinterpret_cast<const float4*>(base + (int64_t)qi * 3 * H + d);
      q[d] = v.x; q[d + 1] = v.y; q[d + 2] = v.z; q[d + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = 0.f;
  }
  auto stage = [&](int j0, bool with_v) {
    __syncthreads();
    const int cnt = min(KT, N - j0);
    for (int idx = threadIdx.x; idx < cnt * (HD / 4); idx += 256) {
      const int j = idx / (HD / 4), d = (idx - j * (HD / 4)) * 4;
      const float* tok = base + (int64_t)(j0 + j) * 3 * H;
      *reinterpret_cast<float4*>(&K_s[j * HD + d]) = *reinterpret_cast<const float4*>(tok + H + d);
      if (with_v) *reinterpret_cast<float4*>(&V_s[j * HD + d]) = *reinterpret_cast<const float4*>(tok + 2 * H + d);
    }
    __syncthreads();
    return cnt;
  };
  auto dot = [&](int j) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int d = 0; d < HD; d += 8) {
      const float4 k0 = *reinterpret_cast<const float4*>(&K_s[j * HD + d]);
      const float4 k1 = *reinterpret_cast<const float4*>(&K_s[j * HD + d + 4]);
      s0 = fmaf(q[d], k0.x, s0); s0 = fmaf(q[d + 1], k0.y, s0); s0 = fmaf(q[d + 2], k0.z, s0); s0 = fmaf(q[d + 3], k0.w, s0);
      s1 = fmaf(q[d + 4], k1.x, s1); s1 = fmaf(q[d + 5], k1.y, s1); s1 = fmaf(q[d + 6], k1.z, s1); s1 = fmaf(q[d + 7], k1.w, s1);
    }
    return (s0 + s1) * scale;
  };
  // pass 1: row maximum
  float mx = -INFINITY;
  for (int j0 = 0; j0 < N; j0 += KT) {
    const int cnt = stage(j0, N <= KT);
    for (int j = 0; j < cnt; ++j) mx = fmaxf(mx, dot(j));
  }
  // pass 2: exp-sum and P V
  float l = 0.f, o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  for (int j0 = 0; j0 < N; j0 += KT) {
    const int cnt = N <= KT ? N : stage(j0, true);   // a single chunk is still resident from pass 1
    for (int j = 0; j < cnt; ++j) {
      const float p = expf(dot(j) - mx);
      l += p;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&V_s[j * HD + d]);
        o[d] = fmaf(p, v.x, o[d]); o[d + 1] = fmaf(p, v.y, o[d + 1]);
        o[d + 2] = fmaf(p, v.z, o[d + 2]); o[d + 3] = fmaf(p, v.w, o[d + 3]);
      }
    }
  }
  if (live) {
    const float inv = 1.0f / l;
    float* dst = out + ((int64_t)b * N + qi) * H + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 4)
      *reinterpret_cast<float4*>(dst + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
  }
}

__device__ __forceinline__ float dot_row(const float* __restrict__ wrow, const float* __restrict__ v, int n) {
  float a0 = 0.f, a1 = 0.f;
  for (int k = 0; k < n; k += 2) { a0 = fmaf(__ldg(wrow + k), v[k], a0); a1 = fmaf(__ldg(wrow + k + 1), v[k + 1], a1); }
  return a0 + a1;
}

// time path (score_network_continuous.py:85-101, 229-233, 117/132): one CTA per distinct t
__global__ void __launch_bounds__(256)
csn_time_kernel(const float* __restrict__ t, const float* __restrict__ freqs, int T, const float* __restrict__ w0,
                const float* __restrict__ b0, const float* __restrict__ w2, const float* __restrict__ b2,
                const float* __restrict__ wt, const float* __restrict__ bt, int sumH, float* __restrict__ tbias) {
  extern __shared__ float sm[];
  float* emb = sm;           // [T]
  float* hid = sm + T;       // [2T]
  float* temb = sm + 3 * T;  // [T]
  const int half = T / 2;
  const float tv = t[blockIdx.x];
  for (int k = threadIdx.x; k < T; k += blockDim.x) {
    const int f = k < half ? k : k - half;
    const float arg = __fmul_rn(__fmul_rn(tv, freqs[f]), 1000.0f);
    emb[k] = k < half ? sinf(arg) : cosf(arg);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * T; j += blockDim.x) hid[j] = silu_acc(dot_row(w0 + (int64_t)j * T, emb, T) + b0[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) temb[i] = dot_row(w2 + (int64_t)i * 2 * T, hid, 2 * T) + b2[i];
  __syncthreads();
  for (int c = threadIdx.x; c < sumH; c += blockDim.x)
    tbias[(int64_t)blockIdx.x * sumH + c] = dot_row(wt + (int64_t)c * T, temb, T) + bt[c];
}

struct ParamTable {
  const dbs_param* p;
  int n;
  const dbs_param* find(const std::string& name) const {
    for (int i = 0; i < n; ++i)
      if (name == p[i].name) return &p[i];
    return nullptr;
  }
};

int upload(dbs_csn_plan* plan, const ParamTable& pt, const std::string& name, int64_t numel, float** out) {
  const dbs_param* q = pt.find(name);
  DBS_REQUIRE(q != nullptr, "csn plan: parameter '%s' missing from state_dict", name.c_str());
  DBS_REQUIRE(q->numel == numel, "csn plan: parameter '%s' has %lld elements, expected %lld", name.c_str(),
              (long long)q->numel, (long long)numel);
  void* d = nullptr;
  DBS_CUDA(plan->arena.take((size_t)(numel ? numel : 4) * 4, &d));
  DBS_CUDA(cudaMemcpy(d, q->data, (size_t)numel * 4, q->is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  *out = (float*)d;
  return DBS_OK;
}

int upload_host(dbs_csn_plan* plan, const float* src, int64_t numel, float** out) {
  void* d = nullptr;
  DBS_CUDA(plan->arena.take((size_t)(numel ? numel : 4) * 4, &d));
  DBS_CUDA(cudaMemcpy(d, src, (size_t)numel * 4, cudaMemcpyHostToDevice));
  *out = (float*)d;
  return DBS_OK;
}

size_t linear_smem(int K, int ncols) {
  const int Kp = (K + 3) & ~3;
  return (size_t)(kTM * (Kp + 4) + kKC * kLdW + kTM * (ncols + 1)) * sizeof(float);
}

int run_linear(int pro, LinArgs a, cudaStream_t st) {
  a.cols_per_cta = a.post_g ? a.N : (a.N < kTN ? a.N : kTN);
  DBS_REQUIRE(a.K <= 512 && a.cols_per_cta <= 512, "csn: layer width %d x %d exceeds the linear kernel's tile", a.K, a.N);
  const size_t smem = linear_smem(a.K, a.cols_per_cta);
  DBS_REQUIRE(smem <= 200 * 1024, "csn: layer width %d x %d needs %zu B of shared memory", a.K, a.N, smem);
  static bool attr_done = false;
  if (!attr_done) {
    DBS_CUDA(cudaFuncSetAttribute(csn_linear_kernel<kProNone>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DBS_CUDA(cudaFuncSetAttribute(csn_linear_kernel<kProLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DBS_CUDA(cudaFuncSetAttribute(csn_linear_kernel<kProEmbed>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  const dim3 grid((unsigned)((a.M + kTM - 1) / kTM), (unsigned)((a.N + a.cols_per_cta - 1) / a.cols_per_cta));
  prof_set_label("linear pro%d K%d N%d ln%d", pro, a.K, a.N, a.post_g ? 1 : 0);
  ProfScope prof(kProfCsnLinear, st);
  if (pro == kProNone) csn_linear_kernel<kProNone><<<grid, 256, smem, st>>>(a);
  else if (pro == kProLN) csn_linear_kernel<kProLN><<<grid, 256, smem, st>>>(a);
  else csn_linear_kernel<kProEmbed><<<grid, 256, smem, st>>>(a);
  DBS_LAUNCHED();
  return DBS_OK;
}

template <int HD>
int launch_attention(const float* qkv, float* out, int64_t B, int N, int heads, cudaStream_t st) {
  const size_t smem = (size_t)2 * 256 * HD * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    DBS_CUDA(cudaFuncSetAttribute(csn_attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const dim3 grid((unsigned)(B * heads), (unsigned)((N + 255) / 256));
  prof_set_label("attention N%d heads%d hd%d", N, heads, HD);
  ProfScope prof(kProfCsnAttention, st);
  csn_attention_kernel<HD><<<grid, 256, smem, st>>>(qkv, out, N, heads, 1.0f / sqrtf((float)HD));
  DBS_LAUNCHED();
  return DBS_OK;
}

struct Bump {
  char* base; int64_t cap; int64_t off = 0;
  void* take(int64_t bytes) {
    const int64_t a = (off + 255) & ~(int64_t)255;
    off = a + bytes;
    return (base && off <= cap) ? base + a : nullptr;
  }
};

int forward_impl(dbs_csn_plan* plan, const float* x, const float* t, int64_t n_t, float* score, int64_t B, int N,
                 int precision, void* workspace, int64_t ws_bytes, cudaStream_t st, int64_t* need_bytes) {
  const dbs_csn_config& c = plan->cfg;
  const int H = c.hidden_dim, T = c.time_embed_dim, nb = c.num_blocks;
  const int64_t M = B * N;
  Bump ws{(char*)workspace, ws_bytes};
  float* tbias = (float*)ws.take(n_t * (int64_t)nb * H * 4);
  float* h = (float*)ws.take(M * H * 4);
  float* h1 = (float*)ws.take(M * H * 4);
  float* qkv = (float*)ws.take(c.use_attention ? M * 3 * H * 4 : 16);
  float* o1 = (float*)ws.take(M * (H / 2) * 4);
  if (need_bytes) { *need_bytes = ws.off + 256; return DBS_OK; }
  DBS_REQUIRE(workspace != nullptr && ws.off <= ws_bytes, "csn_forward: workspace too small (%lld < %lld bytes)",
              (long long)ws_bytes, (long long)ws.off);
  int rc;
  const bool tc = precision == 1;
  if (tc) {
    if (!plan->tc_ok || N % 16 != 0 || N > 256) {
      set_error("csn_forward: the bf16 tcgen05 path needs hidden_dim 128, head_dim 32, attention on and N in {16, 32, ..., 256} "
                "(got hidden %d, heads %d, N %d)", c.hidden_dim, c.num_heads, N);
      return DBS_ERR_UNSUPPORTED;
    }
  }
  {
    prof_set_label("csn time");
    ProfScope prof(kProfCsnLinear, st);
    csn_time_kernel<<<(unsigned)n_t, 256, (size_t)4 * T * 4, st>>>(t, plan->t_freqs, T, plan->tw0, plan->tb0, plan->tw2,
                                                                  plan->tb2, plan->twt, plan->tbt, nb * H, tbias);
    DBS_LAUNCHED();
  }
  LinArgs a;
  // input projection of [Fourier features | raw coordinates]  (:262-266)
  if (tc && plan->tc_io_ok) {
    if ((rc = csn_tc_linear(x, 0, plan->in_h, plan->in_b, H, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, N, nullptr, h, 0, M, st,
                            plan->pos_freqs)))
      return rc;
  } else {
    memset(&a, 0, sizeof(a));
    a.X = x; a.K = c.pos_embed_dim + 2; a.W = plan->in_w; a.bias = plan->in_b; a.N = H;
    a.pos_freqs = plan->pos_freqs; a.n_freq = c.pos_embed_dim / 4; a.out = h; a.M = M; a.tokens_per_cfg = N;
    if ((rc = run_linear(kProEmbed, a, st))) return rc;
  }
  for (int i = 0; i < nb; ++i) {
    const dbs_csn_plan::Block& b = plan->blocks[i];
    if (tc) {
      // bf16 tensor-core route: the same five steps, LayerNorm / SiLU / time bias / residual fused into the GEMM
      // kernels' prologue and epilogue; h (the residual stream) stays fp32, h1 / qkv / attention output are bf16
      const float* tb = tbias + (int64_t)i * H;
      const int64_t tbs = n_t == 1 ? 0 : (int64_t)nb * H;
      // MLPBlock as ONE launch: fc1 -> norm1 -> SiLU -> + time bias -> fc2 -> norm2 -> SiLU -> + x, in place on the
      // residual stream (a CTA reads and writes only its own token rows).  Parity-green but measured SLOWER than the two
      // launches (4096 x 256: linears 9.89 vs 9.30 ms per forward): the CTA runs stage -> MMA -> drain -> MMA -> drain as
      // serial phases and the third weight tile drops it from 3 to 2 CTAs per SM, so the saved HBM round trip of the bf16
      // intermediate is paid back in latency.  Opt-in: DBS_CSN_FUSE_MLP=1.
      static const bool fuse_mlp = [] { const char* e = getenv("DBS_CSN_FUSE_MLP"); return e && e[0] == '1'; }();
      if (fuse_mlp) {
        if ((rc = csn_tc_linear(h, 0, b.fc1_h, b.fc1_b, H, nullptr, nullptr, b.n1_g, b.n1_b, 1, tb, tbs, N, nullptr, h, 0, M, st,
                                nullptr, nullptr, nullptr, b.fc2_h, b.fc2_b, b.n2_g, b.n2_b)))
          return rc;
      } else {
        if ((rc = csn_tc_linear(h, 0, b.fc1_h, b.fc1_b, H, nullptr, nullptr, b.n1_g, b.n1_b, 1, tb, tbs, N, nullptr, h1, 1, M, st))) return rc;
        if ((rc = csn_tc_linear(h1, 1, b.fc2_h, b.fc2_b, H, nullptr, nullptr, b.n2_g, b.n2_b, 1, nullptr, 0, N, h, h, 0, M, st))) return rc;
      }
      if ((rc = csn_tc_linear(h, 0, b.qkv_h, b.qkv_b, 3 * H, b.an_g, b.an_b, nullptr, nullptr, 0, nullptr, 0, N, nullptr, qkv, 1, M, st))) return rc;
      if ((rc = csn_tc_attention((const __nv_bfloat16*)qkv, (__nv_bfloat16*)h1, B, N, c.num_heads, st))) return rc;
      if ((rc = csn_tc_linear(h1, 1, b.proj_h, b.proj_b, H, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, N, h, h, 0, M, st))) return rc;
      continue;
    }
    // MLPBlock (:123-136)
    memset(&a, 0, sizeof(a));
    a.X = h; a.K = H; a.W = b.fc1_w; a.bias = b.fc1_b; a.N = H; a.post_g = b.n1_g; a.post_b = b.n1_b; a.silu = 1;
    a.tbias = tbias + (int64_t)i * H; a.tb_stride = n_t == 1 ? 0 : (int64_t)nb * H; a.tokens_per_cfg = N;
    a.out = h1; a.M = M;
    if ((rc = run_linear(kProNone, a, st))) return rc;
    memset(&a, 0, sizeof(a));
    a.X = h1; a.K = H; a.W = b.fc2_w; a.bias = b.fc2_b; a.N = H; a.post_g = b.n2_g; a.post_b = b.n2_b; a.silu = 1;
    a.res = h; a.out = h; a.M = M; a.tokens_per_cfg = N;   // in place: a CTA reads exactly the residual rows it writes
    if ((rc = run_linear(kProNone, a, st))) return rc;
    if (c.use_attention) {
      // ParticleAttention (:155-180)
      memset(&a, 0, sizeof(a));
      a.X = h; a.K = H; a.W = b.qkv_w; a.bias = b.qkv_b; a.N = 3 * H; a.pre_g = b.an_g; a.pre_b = b.an_b;
      a.out = qkv; a.M = M; a.tokens_per_cfg = N;
      if ((rc = run_linear(kProLN, a, st))) return rc;
      const int hd = H / c.num_heads;
      if (hd == 8) rc = launch_attention<8>(qkv, h1, B, N, c.num_heads, st);
      else if (hd == 16) rc = launch_attention<16>(qkv, h1, B, N, c.num_heads, st);
      else if (hd == 32) rc = launch_attention<32>(qkv, h1, B, N, c.num_heads, st);
      else if (hd == 64) rc = launch_attention<64>(qkv, h1, B, N, c.num_heads, st);
      else { set_error("csn: head_dim %d is not built (8, 16, 32, 64)", hd); return DBS_ERR_UNSUPPORTED; }
      if (rc) return rc;
      memset(&a, 0, sizeof(a));
      a.X = h1; a.K = H; a.W = b.proj_w; a.bias = b.proj_b; a.N = H; a.res = h; a.out = h; a.M = M; a.tokens_per_cfg = N;
      // proj is column-tiled (no post-LayerNorm): each CTA reads/writes only its own (rows, columns) of h
      if ((rc = run_linear(kProNone, a, st))) return rc;
    }
  }
  // output head: LN -> Linear -> SiLU -> Linear (:248-253)
  if (tc && plan->tc_io_ok)
    return csn_tc_linear(h, 0, plan->o1_h, plan->o1_bpad, H, plan->on_g, plan->on_b, nullptr, nullptr, 1, nullptr, 0, N, nullptr, score,
                         0, M, st, nullptr, plan->o2_w, plan->o2_b);
  memset(&a, 0, sizeof(a));
  a.X = h; a.K = H; a.W = plan->o1_w; a.bias = plan->o1_b; a.N = H / 2; a.pre_g = plan->on_g; a.pre_b = plan->on_b;
  a.silu = 1; a.out = o1; a.M = M; a.tokens_per_cfg = N;
  if ((rc = run_linear(kProLN, a, st))) return rc;
  memset(&a, 0, sizeof(a));
  a.X = o1; a.K = H / 2; a.W = plan->o2_w; a.bias = plan->o2_b; a.N = 2; a.out = score; a.M = M; a.tokens_per_cfg = N;
  return run_linear(kProNone, a, st);
}

}  // namespace
}  // namespace dbs

using namespace dbs;

extern "C" int dbs_csn_plan_create(const dbs_csn_config* cfg, const dbs_param* params, int n_params,
                                   const float* pos_freqs_host, const float* t_freqs_host, dbs_csn_plan** out) {
  DBS_REQUIRE(cfg && params && pos_freqs_host && t_freqs_host && out, "csn_plan_create: null argument");
  const int H = cfg->hidden_dim, T = cfg->time_embed_dim, P = cfg->pos_embed_dim, nb = cfg->num_blocks;
  DBS_REQUIRE(H >= 8 && H <= 256 && H % 8 == 0, "csn: hidden_dim %d must be a multiple of 8 in [8, 256]", H);
  DBS_REQUIRE(T >= 4 && T % 2 == 0 && T <= 512, "csn: time_embed_dim %d must be even, in [4, 512]", T);
  DBS_REQUIRE(P >= 4 && P % 4 == 0 && P <= 256, "csn: pos_embed_dim %d must be a multiple of 4 in [4, 256]", P);
  DBS_REQUIRE(nb >= 1 && nb <= 64, "csn: num_blocks %d out of range", nb);
  DBS_REQUIRE(cfg->num_heads >= 1 && H % cfg->num_heads == 0, "csn: hidden_dim %d not divisible by %d heads", H, cfg->num_heads);
  dbs_csn_plan* plan = new dbs_csn_plan();
  plan->cfg = *cfg;
  const ParamTable pt{params, n_params};
  int rc = DBS_OK;
  auto up = [&](const std::string& name, int64_t numel, float** dst) { if (!rc) rc = upload(plan, pt, name, numel, dst); };
  if (!rc) rc = upload_host(plan, pos_freqs_host, P / 4, &plan->pos_freqs);
  if (!rc) rc = upload_host(plan, t_freqs_host, T / 2, &plan->t_freqs);
  up("time_mlp.0.weight", (int64_t)2 * T * T, &plan->tw0); up("time_mlp.0.bias", 2 * T, &plan->tb0);
  up("time_mlp.2.weight", (int64_t)2 * T * T, &plan->tw2); up("time_mlp.2.bias", T, &plan->tb2);
  up("input_proj.weight", (int64_t)H * (P + 2), &plan->in_w); up("input_proj.bias", H, &plan->in_b);
  plan->blocks.resize(nb);
  std::vector<float> wt((size_t)nb * H * T), bt((size_t)nb * H);
  for (int i = 0; i < nb && !rc; ++i) {
    dbs_csn_plan::Block& b = plan->blocks[i];
    memset(&b, 0, sizeof(b));
    const std::string p = "blocks." + std::to_string(i) + ".";
    up(p + "fc1.weight", (int64_t)H * H, &b.fc1_w); up(p + "fc1.bias", H, &b.fc1_b);
    up(p + "fc2.weight", (int64_t)H * H, &b.fc2_w); up(p + "fc2.bias", H, &b.fc2_b);
    up(p + "norm1.weight", H, &b.n1_g); up(p + "norm1.bias", H, &b.n1_b);
    up(p + "norm2.weight", H, &b.n2_g); up(p + "norm2.bias", H, &b.n2_b);
    // the block's time Linear goes into the concatenated table the time kernel evaluates
    const dbs_param* w = pt.find(p + "time_mlp.weight");
    const dbs_param* bb = pt.find(p + "time_mlp.bias");
    if (!rc && (!w || !bb || w->numel != (int64_t)H * T || bb->numel != H)) {
      set_error("csn plan: parameter '%stime_mlp.*' missing or of the wrong size", p.c_str());
      rc = DBS_ERR_INVALID;
    }
    if (!rc) {
      const cudaMemcpyKind kind = w->is_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
      if (cudaMemcpy(wt.data() + (size_t)i * H * T, w->data, (size_t)H * T * 4, kind) != cudaSuccess ||
          cudaMemcpy(bt.data() + (size_t)i * H, bb->data, (size_t)H * 4,
                     bb->is_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost) != cudaSuccess) {
        set_error("csn plan: copying '%stime_mlp' failed", p.c_str());
        rc = DBS_ERR_CUDA;
      }
    }
    if (cfg->use_attention) {
      const std::string q = "attention_layers." + std::to_string(i) + ".";
      up(q + "norm.weight", H, &b.an_g); up(q + "norm.bias", H, &b.an_b);
      up(q + "qkv.weight", (int64_t)3 * H * H, &b.qkv_w); up(q + "qkv.bias", 3 * H, &b.qkv_b);
      up(q + "proj.weight", (int64_t)H * H, &b.proj_w); up(q + "proj.bias", H, &b.proj_b);
    }
  }
  plan->tc_ok = H == 128 && cfg->use_attention && H / cfg->num_heads == 32;
  if (plan->tc_ok && !rc) {
    for (int i = 0; i < nb && !rc; ++i) {
      dbs_csn_plan::Block& b = plan->blocks[i];
      struct { const float* src; __nv_bfloat16** dst; int64_t n; } cv[4] = {
          {b.fc1_w, &b.fc1_h, (int64_t)H * H}, {b.fc2_w, &b.fc2_h, (int64_t)H * H},
          {b.qkv_w, &b.qkv_h, (int64_t)3 * H * H}, {b.proj_w, &b.proj_h, (int64_t)H * H}};
      for (auto& e : cv) {
        void* d = nullptr;
        if (plan->arena.take((size_t)e.n * 2, &d) != cudaSuccess) { set_error("csn plan: cudaMalloc failed"); rc = DBS_ERR_CUDA; break; }
        *e.dst = (__nv_bfloat16*)d;
        if ((rc = csn_tc_convert_weights(e.src, *e.dst, e.n))) break;
      }
    }
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) { set_error("csn plan: weight conversion failed"); rc = DBS_ERR_CUDA; }
  }
  if (!rc) rc = upload_host(plan, wt.data(), (int64_t)nb * H * T, &plan->twt);
  if (!rc) rc = upload_host(plan, bt.data(), (int64_t)nb * H, &plan->tbt);
  up("output_proj.0.weight", H, &plan->on_g); up("output_proj.0.bias", H, &plan->on_b);
  up("output_proj.1.weight", (int64_t)(H / 2) * H, &plan->o1_w); up("output_proj.1.bias", H / 2, &plan->o1_b);
  up("output_proj.3.weight", (int64_t)2 * (H / 2), &plan->o2_w); up("output_proj.3.bias", 2, &plan->o2_b);
  plan->tc_io_ok = plan->tc_ok && P == 64;
  if (!rc && plan->tc_io_ok) {
    // zero-padded [128][128] copies: input_proj [128][66] -> K 128; output_proj.1 [64][128] -> N 128
    std::vector<float> hw((size_t)H * (P + 2)), pad((size_t)128 * 128, 0.f), hb(H / 2), bpad(128, 0.f);
    float* tmp = nullptr;
    auto to_bf16 = [&](const std::vector<float>& src, __nv_bfloat16** dst) {
      if (rc) return;
      void* d = nullptr;
      if (plan->arena.take(src.size() * 2, &d) != cudaSuccess || plan->arena.take(src.size() * 4, (void**)&tmp) != cudaSuccess) { set_error("csn plan: cudaMalloc failed"); rc = DBS_ERR_CUDA; return; }
      if (cudaMemcpy(tmp, src.data(), src.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("csn plan: upload failed"); rc = DBS_ERR_CUDA; return; }
      *dst = (__nv_bfloat16*)d;
      rc = csn_tc_convert_weights(tmp, *dst, (int64_t)src.size());
    };
    if (cudaMemcpy(hw.data(), plan->in_w, hw.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("csn plan: readback failed"); rc = DBS_ERR_CUDA; }
    if (!rc) {
      for (int n = 0; n < H; ++n) for (int k = 0; k < P + 2; ++k) pad[(size_t)n * 128 + k] = hw[(size_t)n * (P + 2) + k];
      to_bf16(pad, &plan->in_h);
    }
    std::vector<float> ow((size_t)(H / 2) * H);
    if (!rc && (cudaMemcpy(ow.data(), plan->o1_w, ow.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(hb.data(), plan->o1_b, hb.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess)) { set_error("csn plan: readback failed"); rc = DBS_ERR_CUDA; }
    if (!rc) {
      std::fill(pad.begin(), pad.end(), 0.f);
      for (size_t i = 0; i < ow.size(); ++i) pad[i] = ow[i];          // rows 0..63 of [128][128]
      to_bf16(pad, &plan->o1_h);
      for (int i = 0; i < H / 2; ++i) bpad[i] = hb[i];
      if (!rc) rc = upload_host(plan, bpad.data(), 128, &plan->o1_bpad);
    }
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) { set_error("csn plan: padded weight conversion failed"); rc = DBS_ERR_CUDA; }
  }
  if (rc) { dbs_csn_plan_destroy(plan); return rc; }
  *out = plan;
  return DBS_OK;
}

extern "C" void dbs_csn_plan_destroy(dbs_csn_plan* plan) {
  if (!plan) return;
  plan->arena.release();
  delete plan;
}

extern "C" int64_t dbs_csn_workspace_bytes(const dbs_csn_plan* plan, int64_t B, int N, int64_t n_t) {
  if (plan && B == 0) return 0;   // empty batch
  if (!plan || B <= 0 || N <= 0 || n_t <= 0) { set_error("csn_workspace_bytes: bad argument"); return -1; }
  int64_t need = 0;
  if (forward_impl(const_cast<dbs_csn_plan*>(plan), nullptr, nullptr, n_t, nullptr, B, N, 0, nullptr, 0, nullptr, &need)) return -1;
  return need;
}

extern "C" int dbs_csn_forward(dbs_csn_plan* plan, const float* x, const float* t, int64_t n_t, float* score,
                               int64_t B, int N, int precision, void* workspace, int64_t workspace_bytes, void* stream) {
  if (plan && B == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(plan && x && t && score, "csn_forward: null argument");
  DBS_REQUIRE(B > 0 && N > 0, "csn_forward: empty batch (%lld x %d)", (long long)B, N);
  DBS_REQUIRE(n_t == 1 || n_t == B, "csn_forward: t must have 1 or B entries (got %lld)", (long long)n_t);
  DBS_REQUIRE(precision == 0 || precision == 1, "csn_forward: precision must be 0 (fp32) or 1 (bf16 tcgen05)");
  return forward_impl(plan, x, t, n_t, score, B, N, precision, workspace, workspace_bytes, as_stream(stream), nullptr);
}
