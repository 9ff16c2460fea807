// csn_tc.cu — tensor-core (tcgen05 / TMEM) kernels of the particle score network, bf16 operands with
// fp32 accumulation (reference models/score_network_continuous.py:123-180).
//
//   csn_tc_linear_kernel     Y[128 tokens x 128] = epi(pro(X)[128 x 128] . W[128 x 128]^T) per tile.
//       The prologue (LayerNorm of the input rows, ParticleAttention.norm :166) runs while the rows are
//       converted to bf16 and written straight into the 128-byte-swizzled K-major operand layout the
//       UMMA descriptor names — no TMA here because the operand is produced, not copied.  Eight
//       tcgen05.mma (M=128, N=128, K=16) accumulate in TMEM; in the epilogue a thread owns one token row
//       (TMEM lane = row), so MLPBlock's LayerNorm -> SiLU -> + time bias / + residual (:130-136) is
//       thread-local arithmetic on tcgen05.ld chunks.  HBM traffic per layer = one read of X, one write
//       of Y: the normalised / activated intermediates never leave the SM.
//   csn_tc_attention_kernel  one CTA per (configuration, head), N <= 256 particles, head_dim 32:
//       S = Q K^T (two 128-row tiles, N columns) into TMEM; softmax per query row is thread-local
//       (lane = query); P goes to shared memory as the bf16 A operand of O = P V (V transposed into the
//       K-major B layout on the way in); the N x N matrix never reaches HBM  (:169-178).
// Both kernels are phase-serial inside a CTA (produce operands -> MMA -> drain TMEM); overlap comes from
// several CTAs per SM (linear) and from the persistent loop (attention).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dbs {
namespace {

using namespace tc;

constexpr uint32_t kAtomBytes = 128 * 128;   // one 64-element K slab of 128 rows (128-byte rows)

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float silu_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// K-major operand descriptor words for 128-byte rows, SWIZZLE_128B (same encoding as conv_tc.cu, KC = 64)
constexpr uint64_t kDesc128 = (1ull << 16) | ((uint64_t)((8 * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
// 64-byte rows, SWIZZLE_64B (conv_tc.cu, KC = 32)
constexpr uint64_t kDesc64 = (1ull << 16) | ((uint64_t)((8 * 64) >> 4) << 32) | (1ull << 46) | (4ull << 61);
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcLinArgs {
  const void* X;               // [M][128] fp32 (in_bf16 = 0) or bf16 (in_bf16 = 1)
  const __nv_bfloat16* W;      // [N_total][128] bf16
  const float* bias;           // [N_total]
  int N_total;                 // multiple of 128; blockIdx.y selects the 128-column block
  const float *pre_g, *pre_b;  // LayerNorm over the 128 inputs (null = none)
  const float *post_g, *post_b;  // LayerNorm over the 128 outputs (null = none; needs N_total == 128)
  int silu;
  const float* tbias; int64_t tb_stride; int tokens_per_cfg;
  const float* res;            // fp32 [M][N_total] or null
  void* out;                   // [M][N_total] fp32 (out_bf16 = 0) or bf16
  int in_bf16, out_bf16;
  int64_t M;
  // embedding prologue (input_proj, :262-266): X = positions [M][2]; the operand row is
  // [sin(a_x f) | cos(a_x f) | sin(a_y f) | cos(a_y f) | x | y | 0 ...], 16 frequencies, padded to K = 128
  const float* pos_freqs;       // null = not the embedding layer
  // fused output head (:248-253): this layer is Linear(128 -> 64) + SiLU and the final Linear(64 -> 2) is
  // applied to the thread's row in the epilogue; out = score fp32 [M][2]
  const float* head_w2;         // [2][64], null = no head
  const float* head_b2;         // [2]
  // fused MLPBlock (:123-136): this launch is fc1 (post LayerNorm norm1, SiLU, + time bias) AND fc2 (LayerNorm norm2, SiLU)
  // + the block's residual x.  The activated fc1 rows go straight from the TMEM drain into the bf16 A-operand tile of the
  // second GEMM: the intermediate never reaches HBM and the residual stream is read once and written once per block.
  const __nv_bfloat16* W2;      // [128][128] bf16, null = single layer
  const float *bias2, *post2_g, *post2_b;
};

// a warp converts 128-wide rows to bf16 and stores them into the swizzled K-major operand tile; the loads of
// kRows rows are issued together so that their latencies overlap
template <bool kBf16In>
__device__ __forceinline__ void load_row(const void* src_row, bool live, int lane, float v[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (live) {
    if (kBf16In) {
      const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(src_row) + 4 * lane);
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
      const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
      v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(c); v[3] = __high2float(c);
    } else {
      const float4 q = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src_row) + 4 * lane);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
  }
}
__device__ __forceinline__ void finish_row(uint8_t* tile, int r, int lane, float v[4], const float* g, const float* b) {
  if (g) {
    const float mean = warp_sum((v[0] + v[1]) + (v[2] + v[3])) * (1.0f / 128.0f);
    const float d0 = v[0] - mean, d1 = v[1] - mean, d2 = v[2] - mean, d3 = v[3] - mean;
    const float var = warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) * (1.0f / 128.0f);
    const float rstd = rsqrtf(var + 1e-5f);
    const float4 g4 = *reinterpret_cast<const float4*>(g + 4 * lane);
    const float4 b4 = *reinterpret_cast<const float4*>(b + 4 * lane);
    v[0] = d0 * rstd * g4.x + b4.x; v[1] = d1 * rstd * g4.y + b4.y;
    v[2] = d2 * rstd * g4.z + b4.z; v[3] = d3 * rstd * g4.w + b4.w;
  }
  // element k = 4*lane: slab k/64, 16-byte chunk (k%64)/8 XOR (row % 8), 8-byte half (lane & 1)
  const uint32_t off = (uint32_t)(lane >> 4) * kAtomBytes + (uint32_t)r * 128u +
                       (uint32_t)((((lane & 15) >> 1) ^ (r & 7)) << 4) + (uint32_t)(lane & 1) * 8u;
  *reinterpret_cast<uint2*>(tile + off) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
}
// Fourier-embedding operand rows: lane l produces features 4l .. 4l+3 of its token
__device__ __forceinline__ void stage_embed(uint8_t* tile, const float* pos, int64_t rows_live, int warp, int lane,
                                            const float* freqs) {
  const int qd = lane >> 2;                 // 0..3: sin x, cos x, sin y, cos y;  4: raw coords (lane 16);  >4: zero padding
  float fr[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane < 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) fr[i] = freqs[4 * (lane & 3) + i];
  }
#pragma unroll 1
  for (int r = warp; r < 128; r += 8) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < rows_live) {
      const float2 pq = *reinterpret_cast<const float2*>(pos + (int64_t)r * 2);
      if (lane < 16) {
        const float base = __fmul_rn(__fadd_rn(qd < 2 ? pq.x : pq.y, 1.0f), 3.14159265358979323846f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float arg = __fmul_rn(base, fr[i]);
          v[i] = (qd & 1) ? cosf(arg) : sinf(arg);
        }
      } else if (lane == 16) {
        v[0] = pq.x; v[1] = pq.y;
      }
    }
    finish_row(tile, r, lane, v, nullptr, nullptr);
  }
}

// rows warp, warp + 8, ... of a 128-row tile; src = first row of the tile, rows_live = rows that exist
template <bool kBf16In>
__device__ __forceinline__ void stage_tile(uint8_t* tile, const void* src, int64_t rows_live, int warp, int lane,
                                           const float* g, const float* b) {
  constexpr int kRows = 4;
  const size_t esz = kBf16In ? 2 : 4;
#pragma unroll 1
  for (int r0 = warp; r0 < 128; r0 += 8 * kRows) {
    float v[kRows][4];
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int r = r0 + 8 * i;
      load_row<kBf16In>(reinterpret_cast<const uint8_t*>(src) + (size_t)r * 128 * esz, r < rows_live, lane, v[i]);
    }
#pragma unroll
    for (int i = 0; i < kRows; ++i) finish_row(tile, r0 + 8 * i, lane, v[i], g, b);
  }
}

__global__ void __launch_bounds__(256, 3)
csn_tc_linear_kernel(const TcLinArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* As = smem;                   // 2 slabs x 16 KB
  uint8_t* Ws = smem + 2 * kAtomBytes;  // 2 slabs x 16 KB
  uint8_t* Ws2 = smem + 4 * kAtomBytes; // fused MLPBlock: fc2's weights, 2 slabs x 16 KB (allocated only then)
  __shared__ __align__(16) float s_bias2[128], s_pg2[128], s_pb2[128];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  // per-column epilogue vectors of this CTA's 128 columns: warp-uniform reads, so they are served as
  // shared-memory broadcasts (float4) instead of ~4 global loads per output element
  __shared__ __align__(16) float s_bias[128], s_pg[128], s_pb[128], s_tb[128];
  __shared__ __align__(16) float s_w2[2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.y * 128;
  const bool tb_shared = a.tbias != nullptr && a.tb_stride == 0;
  if (threadIdx.x < 128) {
    const int c = threadIdx.x;
    s_bias[c] = a.bias[col0 + c];
    s_pg[c] = a.post_g ? a.post_g[c] : 1.f;
    s_pb[c] = a.post_b ? a.post_b[c] : 0.f;
    s_tb[c] = tb_shared ? a.tbias[col0 + c] : 0.f;
    if (a.head_w2) s_w2[c >> 6][c & 63] = a.head_w2[c];
    if (a.W2) { s_bias2[c] = a.bias2[c]; s_pg2[c] = a.post2_g[c]; s_pb2[c] = a.post2_b[c]; }
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 128);
  // this CTA's 128 weight rows, once
  stage_tile<true>(Ws, a.W + (int64_t)col0 * 128, 128, warp, lane, nullptr, nullptr);
  if (a.W2) stage_tile<true>(Ws2, a.W2, 128, warp, lane, nullptr, nullptr);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t a_lo = (uint32_t)kDesc128 | ((smem_u32(As) & 0x3FFFFu) >> 4);
  const uint32_t w_lo = (uint32_t)kDesc128 | ((smem_u32(Ws) & 0x3FFFFu) >> 4);
  const uint32_t w2_lo = (uint32_t)kDesc128 | ((smem_u32(Ws2) & 0x3FFFFu) >> 4);
  const uint32_t hi32 = (uint32_t)(kDesc128 >> 32);
  const int64_t n_tiles = (a.M + 127) / 128;
  uint32_t phase = 0;
  const int q = warp & 3, ch = warp >> 2;        // epilogue: TMEM lane quadrant, 64-column half
  const int N = a.N_total;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t m0 = tile * 128;
    // ---- phase 1: operand rows (prologue fused)
    if (a.pos_freqs) stage_embed(As, reinterpret_cast<const float*>(a.X) + m0 * 2, a.M - m0, warp, lane, a.pos_freqs);
    else if (a.in_bf16) stage_tile<true>(As, reinterpret_cast<const __nv_bfloat16*>(a.X) + m0 * 128, a.M - m0, warp, lane, a.pre_g, a.pre_b);
    else stage_tile<false>(As, reinterpret_cast<const float*>(a.X) + m0 * 128, a.M - m0, warp, lane, a.pre_g, a.pre_b);
    fence_async_smem();
    __syncthreads();
    // ---- phase 2: eight MMAs (K = 128), completion signalled on the mbarrier
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t o = (uint32_t)(k >> 2) * (kAtomBytes >> 4) + (uint32_t)(k & 3) * 2u;
        umma_bf16_lohi(tmem, a_lo + o, w_lo + o, hi32, idesc_bf16(128, 128), k ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    tc_fence_after();
    // ---- phase 3: drain TMEM, one token row per thread
    const int row = q * 32 + lane;
    const int64_t m = m0 + row;
    const bool live = m < a.M;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
    float mean = 0.f, rstd = 1.f;
    if (a.post_g) {
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int c = 0; c < 128; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c + 4 * j4]);
          const float f0 = __uint_as_float(v[4 * j4]) + b4.x, f1 = __uint_as_float(v[4 * j4 + 1]) + b4.y;
          const float f2 = __uint_as_float(v[4 * j4 + 2]) + b4.z, f3 = __uint_as_float(v[4 * j4 + 3]) + b4.w;
          s += (f0 + f1) + (f2 + f3);
          ss = fmaf(f0, f0, ss); ss = fmaf(f1, f1, ss); ss = fmaf(f2, f2, ss); ss = fmaf(f3, f3, ss);
        }
      }
      mean = s * (1.0f / 128.0f);
      rstd = rsqrtf(fmaxf(ss * (1.0f / 128.0f) - mean * mean, 0.f) + 1e-5f);
    }
    const float* tb = (a.tbias && !tb_shared) ? a.tbias + (live ? m / a.tokens_per_cfg : 0) * a.tb_stride : nullptr;
    if (a.W2) {
      // ---- fused MLPBlock, stage 1 drain: norm1 -> SiLU -> + time bias, written as the bf16 A operand of fc2
#pragma unroll 1
      for (int c = ch * 64; c < ch * 64 + 64; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j2 = 0; j2 < 8; ++j2) {
          float y[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int j = 2 * j2 + k;
            float x = (__uint_as_float(v[j]) + s_bias[c + j] - mean) * rstd * s_pg[c + j] + s_pb[c + j];
            x = silu_fast(x) + s_tb[c + j];
            if (tb) x += __ldg(tb + c + j);
            y[k] = x;
          }
          pk[j2] = pack_bf16(y[0], y[1]);
        }
        // element k = c .. c + 15 of row `row`: slab k / 64, 16-byte chunks (k % 64) / 8 and + 1, XOR (row % 8)
        uint8_t* rowp = As + (uint32_t)(c >> 6) * kAtomBytes + (uint32_t)row * 128u;
        const int c8 = (c & 63) >> 3;
        *reinterpret_cast<uint4*>(rowp + (((c8) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rowp + (((c8 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();          // every row of the operand tile written, TMEM drained
      if (threadIdx.x == 0) {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t o = (uint32_t)(k >> 2) * (kAtomBytes >> 4) + (uint32_t)(k & 3) * 2u;
          umma_bf16_lohi(tmem, a_lo + o, w2_lo + o, hi32, idesc_bf16(128, 128), k ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar));
      }
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      tc_fence_after();
      // ---- stage 2 drain: norm2 -> SiLU -> + x (the block's residual, read back from L2) -> out
      float s2 = 0.f, ss2 = 0.f;
#pragma unroll
      for (int c = 0; c < 128; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float f = __uint_as_float(v[j]) + s_bias2[c + j];
          s2 += f;
          ss2 = fmaf(f, f, ss2);
        }
      }
      const float mean2 = s2 * (1.0f / 128.0f);
      const float rstd2 = rsqrtf(fmaxf(ss2 * (1.0f / 128.0f) - mean2 * mean2, 0.f) + 1e-5f);
#pragma unroll 1
      for (int c = ch * 64; c < ch * 64 + 64; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        if (live) {
          // a thread's sixteen fp32 columns are two 32-byte sectors of its row: 32-byte loads / stores (tc_ptx.cuh)
          const float* rp = reinterpret_cast<const float*>(a.X) + m * 128 + c;
          float* dst = reinterpret_cast<float*>(a.out) + m * 128 + c;
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            float rr[8], o[8];
            ld_global_256f(rp + 8 * j8, rr);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int j = 8 * j8 + k;
              const float x = (__uint_as_float(v[j]) + s_bias2[c + j] - mean2) * rstd2 * s_pg2[c + j] + s_pb2[c + j];
              o[k] = silu_fast(x) + rr[k];
            }
            st_global_256f(dst + 8 * j8, o);
          }
        }
      }
      tc_fence_before();
      __syncthreads();
      continue;
    }
    if (a.head_w2) {
      // output head: columns 0..63 hold Linear(128 -> 64) of the row (64..127 are zero padding); SiLU, then the
      // 64 -> 2 projection as two thread-local dot products
      if (ch == 0) {
        float y0 = 0.f, y1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 64; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c + 4 * j4]);
            const float4 u4 = *reinterpret_cast<const float4*>(&s_w2[0][c + 4 * j4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&s_w2[1][c + 4 * j4]);
            const float x0 = silu_fast(__uint_as_float(v[4 * j4]) + b4.x), x1 = silu_fast(__uint_as_float(v[4 * j4 + 1]) + b4.y);
            const float x2 = silu_fast(__uint_as_float(v[4 * j4 + 2]) + b4.z), x3 = silu_fast(__uint_as_float(v[4 * j4 + 3]) + b4.w);
            y0 = fmaf(x0, u4.x, y0); y0 = fmaf(x1, u4.y, y0); y0 = fmaf(x2, u4.z, y0); y0 = fmaf(x3, u4.w, y0);
            y1 = fmaf(x0, w4.x, y1); y1 = fmaf(x1, w4.y, y1); y1 = fmaf(x2, w4.z, y1); y1 = fmaf(x3, w4.w, y1);
          }
        }
        if (live) *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.out) + m * 2) = make_float2(y0 + a.head_b2[0], y1 + a.head_b2[1]);
      }
      tc_fence_before();
      __syncthreads();
      continue;
    }
#pragma unroll 1
    for (int c = ch * 64; c < ch * 64 + 64; c += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c, v);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c + 4 * j4]);
        const float4 g4 = *reinterpret_cast<const float4*>(&s_pg[c + 4 * j4]);
        const float4 p4 = *reinterpret_cast<const float4*>(&s_pb[c + 4 * j4]);
        const float4 t4 = *reinterpret_cast<const float4*>(&s_tb[c + 4 * j4]);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
        const float pp[4] = {p4.x, p4.y, p4.z, p4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = 4 * j4 + k;
          float x = __uint_as_float(v[j]) + bb[k];
          if (a.post_g) x = (x - mean) * rstd * gg[k] + pp[k];
          if (a.silu) x = silu_fast(x);
          x += tt[k];
          if (tb) x += __ldg(tb + col0 + c + j);
          f[j] = x;
        }
      }
      if (live) {
        // 32-byte loads / stores: a thread's chunk is one (bf16) or two (fp32) whole sectors of its row
        if (a.res) {
          const float* rp = a.res + m * N + col0 + c;
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            float rr[8];
            ld_global_256f(rp + 8 * j8, rr);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[8 * j8 + k] += rr[k];
          }
        }
        if (a.out_bf16) {
          const uint32_t pk[8] = {pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]),
                                  pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15])};
          st_global_256(reinterpret_cast<__nv_bfloat16*>(a.out) + m * N + col0 + c, pk);
        } else {
          float* dst = reinterpret_cast<float*>(a.out) + m * N + col0 + c;
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            const float o[8] = {f[8 * j8], f[8 * j8 + 1], f[8 * j8 + 2], f[8 * j8 + 3], f[8 * j8 + 4], f[8 * j8 + 5], f[8 * j8 + 6], f[8 * j8 + 7]};
            st_global_256f(dst + 8 * j8, o);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();     // TMEM drained and the operand tile consumed: next tile may overwrite both
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised, persistent form of the same layer (the default; DBS_CSN_WS=0 selects the phase-serial kernel above).
// The phase-serial kernel leaves every resource idle two thirds of the time (fc1 at 3.1 TB/s although it moves 768 B per
// token and nothing else is near a limit); three CTAs per SM only partly interleave the phases.  Here one CTA per SM keeps
// all three going at once:
//   8 producer warps   rows -> (LayerNorm) -> bf16 -> swizzled operand tile, three 32 KB stages (full / empty mbarriers);
//                      eight rows (4 KB) in flight per warp, so the loads of a stage overlap the previous stages' MMAs and drains
//   1 MMA warp         one elected thread: per staged tile and 128-column block eight tcgen05.mma into one of FOUR TMEM
//                      accumulators; tcgen05.commit releases the operand stage and publishes the accumulator
//   8 epilogue warps   two per TMEM lane quadrant (64 columns each): LayerNorm statistics as half-row partial sums
//                      exchanged through shared memory (one pass over TMEM instead of the old kernel's 1.5 passes per warp),
//                      SiLU / time bias / output head as before; results leave through a padded staging tile in shared
//                      memory so that global memory sees whole rows per warp instruction instead of 32 rows x 16 B
// Measured (4096 x 256 tokens, per layer): qkv 0.92 -> 0.66 ms, fc1 0.26 -> 0.25 ms.  The residual layers (fc2, proj) were
// measured in three forms — row-per-thread residual reads with a one-chunk prefetch (0.49 / 0.44 ms, as the old kernel),
// residual rows staged by the epilogue warps (0.79 ms: the load latency lands on the epilogue's serial path), residual and
// results moved by a bulk-copy (cp.async.bulk) warp through the single staging tile (0.64 ms: the tile's read-out, the
// residual load and the drain form one latency chain per tile) — and stay on the phase-serial kernel, whose three CTAs per
// SM hide that latency better; so does the Fourier-embedding layer (sinf / cosf staging wants all the SM's warps).
// The layer's whole weight matrix (32 KB, 96 KB for qkv) is resident, and a qkv tile is staged ONCE for its three column
// blocks (the old grid.y = 3 staged and LayerNorm-ed every row three times).
constexpr int kWsProd = 8, kWsEpi = 8;
constexpr int kWsThreads = 32 * (kWsProd + 1 + kWsEpi);
constexpr int kWsStages = 3, kWsAcc = 4;

template <bool kBf16In>
__device__ __forceinline__ void ws_stage_rows(uint8_t* tile, const void* src, int64_t rows_live, int pw, int lane,
                                              const float* g, const float* b) {
  constexpr int kRows = 8;
  const size_t esz = kBf16In ? 2 : 4;
#pragma unroll 1
  for (int r0 = pw; r0 < 128; r0 += kWsProd * kRows) {
    float v[kRows][4];
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int r = r0 + kWsProd * i;
      load_row<kBf16In>(reinterpret_cast<const uint8_t*>(src) + (size_t)r * 128 * esz, r < rows_live, lane, v[i]);
    }
#pragma unroll
    for (int i = 0; i < kRows; ++i) finish_row(tile, r0 + kWsProd * i, lane, v[i], g, b);
  }
}

__global__ void __launch_bounds__(kWsThreads, 1)
csn_tc_linear_ws_kernel(const TcLinArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nblk = a.N_total >> 7;
  const int n_stages = nblk > 1 ? 2 : kWsStages;        // qkv: 96 KB of weights leave room for two operand stages
  uint8_t* Ws = smem;                                   // nblk x (2 slabs x 16 KB)
  uint8_t* As = smem + (size_t)nblk * 2 * kAtomBytes;   // n_stages x (2 slabs x 16 KB)
  // output staging tile [128 rows][row bytes + 16]: the drain is one token row per thread (TMEM lane = row), i.e. 32 rows
  // x 16 B per warp instruction if it went to global memory directly (32 sectors per request; the residual read likewise).
  // Results — and the residual on the way in — pass through this tile instead, and global memory sees whole 256 / 512-byte
  // rows per warp instruction.  The 16-byte pad makes the per-row 16-byte accesses of a quarter-warp conflict-free.
  uint8_t* Ys = As + (size_t)n_stages * 2 * kAtomBytes;
  const uint32_t y_row = (a.out_bf16 ? 256u : 512u) + 16u;
  __shared__ __align__(8) uint64_t bars[2 * kWsStages + 2 * kWsAcc];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[384], s_tb[384], s_pg[128], s_pb[128];
  __shared__ float2 s_stat[2][2][128];                  // [parity][column half][row] partial (sum, sum of squares)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kWsStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kWsStages]), bar_tempty = smem_u32(&bars[2 * kWsStages + kWsAcc]);
  const bool tb_shared = a.tbias != nullptr && a.tb_stride == 0;
  for (int c = threadIdx.x; c < a.N_total; c += kWsThreads) {
    s_bias[c] = a.bias[c];
    s_tb[c] = tb_shared ? a.tbias[c] : 0.f;
    if (c < 128) {
      s_pg[c] = a.post_g ? a.post_g[c] : 1.f;
      s_pb[c] = a.post_b ? a.post_b[c] : 0.f;
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWsStages; ++s) { mbar_init(bar_full + 8 * s, kWsProd); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < kWsAcc; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, kWsEpi); }
    fence_barrier_init();
  }
  if (warp == kWsProd) tmem_alloc(smem_u32(&s_tmem), kWsAcc * 128);
  // the layer's weights, once: row r of W -> column block r / 128
  for (int r = warp; r < a.N_total; r += kWsThreads / 32) {
    float v[4];
    load_row<true>(a.W + (int64_t)r * 128, true, lane, v);
    finish_row(Ws + (size_t)(r >> 7) * 2 * kAtomBytes, r & 127, lane, v, nullptr, nullptr);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int64_t n_tiles = (a.M + 127) / 128;

  if (warp < kWsProd) {
    // ===================== producers =====================
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int s = it % n_stages;
      const uint32_t n = (uint32_t)(it / n_stages);
      mbar_wait(bar_empty + 8 * s, (n & 1u) ^ 1u);
      uint8_t* dst = As + (size_t)s * 2 * kAtomBytes;
      const int64_t m0 = tile * 128;
      if (a.pos_freqs) stage_embed(dst, reinterpret_cast<const float*>(a.X) + m0 * 2, a.M - m0, warp, lane, a.pos_freqs);
      else if (a.in_bf16) ws_stage_rows<true>(dst, reinterpret_cast<const __nv_bfloat16*>(a.X) + m0 * 128, a.M - m0, warp, lane, a.pre_g, a.pre_b);
      else ws_stage_rows<false>(dst, reinterpret_cast<const float*>(a.X) + m0 * 128, a.M - m0, warp, lane, a.pre_g, a.pre_b);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
    }
  } else if (warp == kWsProd) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t hi32 = (uint32_t)(kDesc128 >> 32);
      const uint32_t a_lo0 = (uint32_t)kDesc128 | ((smem_u32(As) & 0x3FFFFu) >> 4);
      const uint32_t w_lo0 = (uint32_t)kDesc128 | ((smem_u32(Ws) & 0x3FFFFu) >> 4);
      int it = 0;
      uint32_t acc_it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = it % n_stages;
        const uint32_t n = (uint32_t)(it / n_stages);
        mbar_wait(bar_full + 8 * s, n & 1u);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)s * ((2 * kAtomBytes) >> 4);
        for (int nb = 0; nb < nblk; ++nb, ++acc_it) {
          const uint32_t acc = acc_it & (kWsAcc - 1), n2 = acc_it / kWsAcc;
          mbar_wait(bar_tempty + 8 * acc, (n2 & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t w_lo = w_lo0 + (uint32_t)nb * ((2 * kAtomBytes) >> 4);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t o = (uint32_t)(k >> 2) * (kAtomBytes >> 4) + (uint32_t)(k & 3) * 2u;
            umma_bf16_lohi(tmem + acc * 128, a_lo + o, w_lo + o, hi32, idesc_bf16(128, 128), k ? 1u : 0u);
          }
          umma_commit(bar_tfull + 8 * acc);
        }
        umma_commit(bar_empty + 8 * s);     // the operand stage is free once the tile's MMAs have read it
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3, ch = (warp - kWsProd - 1) >> 2;   // TMEM lane quadrant (= warp % 4), 64-column half
    const int row = q * 32 + lane;
    const int N = a.N_total;
    uint32_t acc_it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t m = tile * 128 + row;
      const bool live = m < a.M;
      const float* tb = (a.tbias && !tb_shared) ? a.tbias + (live ? m / a.tokens_per_cfg : 0) * a.tb_stride : nullptr;
      for (int nb = 0; nb < nblk; ++nb, ++acc_it) {
        const uint32_t acc = acc_it & (kWsAcc - 1), n2 = acc_it / kWsAcc;
        const int col0 = nb * 128;
        mbar_wait(bar_tfull + 8 * acc, n2 & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * 128;
        float mean = 0.f, rstd = 1.f;
        if (a.post_g) {
          // LayerNorm over the 128 outputs of the row: this warp's 64 columns, the other half from its partner warp
          float s = 0.f, ss = 0.f;
#pragma unroll
          for (int c = ch * 64; c < ch * 64 + 64; c += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float f = __uint_as_float(v[j]) + s_bias[c + j];
              s += f;
              ss = fmaf(f, f, ss);
            }
          }
          s_stat[acc_it & 1u][ch][row] = make_float2(s, ss);
        }
        // (1) partial sums published, and every warp is done copying the PREVIOUS tile out of the staging rows it owns
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kWsEpi) : "memory");
        if (a.post_g) {
          const float2 o0 = s_stat[acc_it & 1u][0][row], o1 = s_stat[acc_it & 1u][1][row];
          const float s = o0.x + o1.x, ss = o0.y + o1.y;
          mean = s * (1.0f / 128.0f);
          rstd = rsqrtf(fmaxf(ss * (1.0f / 128.0f) - mean * mean, 0.f) + 1e-5f);
        }
        uint8_t* yrow = Ys + (size_t)row * y_row;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = ch * 64 + cc * 16;
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[col0 + c + 4 * j4]);
            const float4 g4 = *reinterpret_cast<const float4*>(&s_pg[c + 4 * j4]);
            const float4 p4 = *reinterpret_cast<const float4*>(&s_pb[c + 4 * j4]);
            const float4 t4 = *reinterpret_cast<const float4*>(&s_tb[col0 + c + 4 * j4]);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
            const float pp[4] = {p4.x, p4.y, p4.z, p4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int j = 4 * j4 + k;
              float x = __uint_as_float(v[j]) + bb[k];
              if (a.post_g) x = (x - mean) * rstd * gg[k] + pp[k];
              if (a.silu) x = silu_fast(x);
              x += tt[k];
              if (tb) x += __ldg(tb + col0 + c + j);
              f[j] = x;
            }
          }
          if (a.out_bf16) {
            uint4* dst = reinterpret_cast<uint4*>(yrow + 2 * c);
            dst[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            dst[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
          } else {
            float4* dst = reinterpret_cast<float4*>(yrow + 4 * c);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) dst[j4] = make_float4(f[4 * j4], f[4 * j4 + 1], f[4 * j4 + 2], f[4 * j4 + 3]);
          }
        }
        // the accumulator is drained: the MMA warp may reuse it while the tile is copied out
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        // (2) every row of the staging tile is complete: warp ew copies rows 16 ew .. 16 ew + 15 out, whole rows per instruction
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kWsEpi) : "memory");
        {
          const int ew = warp - kWsProd - 1;
          const int64_t rows_live = a.M - tile * 128;
          uint8_t* obase = reinterpret_cast<uint8_t*>(a.out);
          const size_t out_row = (size_t)N * (a.out_bf16 ? 2 : 4), col_off = (size_t)col0 * (a.out_bf16 ? 2 : 4);
#pragma unroll 4
          for (int rr = 0; rr < 16; ++rr) {
            const int r = ew * 16 + rr;
            if (r < rows_live && (lane < 16 || !a.out_bf16))
              *reinterpret_cast<uint4*>(obase + (size_t)(tile * 128 + r) * out_row + col_off + 16 * lane) =
                  *reinterpret_cast<const uint4*>(Ys + (size_t)r * y_row + 16 * lane);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWsProd) {
    tc_fence_after();
    tmem_dealloc(tmem, kWsAcc * 128);
  }
}

// ------------------------------------------------------------------------------------------------
// attention: qkv bf16 [B][N][3][heads][32]  ->  out bf16 [B][N][heads*32]
constexpr int kHD = 32;
constexpr uint32_t kQBytes = 256 * 64;         // Q (two 128-row tiles), K: 64-byte rows, SWIZZLE_64B
constexpr uint32_t kVtSlab = 32 * 128;         // V^T: 32 rows x 64 keys per slab, SWIZZLE_128B
constexpr uint32_t kPTile = 4 * kAtomBytes;    // P tile: 128 rows x 256 keys

// One CTA = one (configuration, head, 128-query tile) unit at a time, 128 threads (thread = query row = TMEM lane), 104 KB of
// shared memory and 256 TMEM columns, so TWO CTAs share an SM and their phases interleave: while one is in its softmax
// (MUFU / TMEM-read work) the other loads, transposes or runs its MMAs.  The first cut (one 256-thread CTA per SM working
// on both query tiles of a pair in lock step) ran every phase serially: 1.23 ms per layer at 4096 x 256.  K and V of a
// pair are fetched once per query tile (L2 hits the second time).
__global__ void __launch_bounds__(128, 2)
csn_tc_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int64_t n_units, int N,
                        int heads, int n_qt, float scale_log2e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Ks = smem;                       // 256 keys x 64 B, SWIZZLE_64B                 16 KB
  uint8_t* Qs = Ks + kQBytes;               // 128 queries x 64 B                             8 KB
  uint8_t* Vt = Qs + kQBytes / 2;           // V^T: 4 slabs of 32 rows x 64 keys             16 KB
  uint8_t* Ps = Vt + 4 * kVtSlab;           // P: 128 rows x 256 keys, 4 K-slabs             64 KB
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int H = heads * kHD;
  const int row = threadIdx.x;              // my query row inside the tile (= warp * 32 + lane: the warp's TMEM lane quadrant)
  uint32_t phase = 0;
  for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int64_t pair = unit / n_qt;
    const int t_mine = (int)(unit - pair * n_qt);
    const int64_t b = pair / heads;
    const int h = (int)(pair - b * heads);
    const int qi = t_mine * 128 + row;      // my query
    const __nv_bfloat16* base = qkv + b * N * 3 * H + h * kHD;
    // ---- phase 1: K rows and my tile's Q rows (64 B each) into the SWIZZLE_64B K-major layout; V transposed into V^T slabs.
    // All of a thread's global loads (8 K chunks, 4 Q chunks, the V rows of two keys: 20 x 16 B) are issued before the first
    // shared-memory store, so the unit pays one global round trip instead of one per loop iteration.
    {
      uint4 kq[8], qq[4], vq[2][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = threadIdx.x + 128 * i, r = idx >> 2, c = idx & 3;
        kq[i] = make_uint4(0, 0, 0, 0);
        if (r < N) kq[i] = *reinterpret_cast<const uint4*>(base + (int64_t)r * 3 * H + H + c * 8);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + 128 * i, r = idx >> 2, c = idx & 3, qr = t_mine * 128 + r;
        qq[i] = make_uint4(0, 0, 0, 0);
        if (qr < N) qq[i] = *reinterpret_cast<const uint4*>(base + (int64_t)qr * 3 * H + c * 8);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = threadIdx.x + 128 * kk;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          vq[kk][i] = make_uint4(0, 0, 0, 0);
          if (k < N) vq[kk][i] = reinterpret_cast<const uint4*>(base + (int64_t)k * 3 * H + 2 * H)[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = threadIdx.x + 128 * i, r = idx >> 2, c = idx & 3;
        *reinterpret_cast<uint4*>(Ks + (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4)) = kq[i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + 128 * i, r = idx >> 2, c = idx & 3;
        *reinterpret_cast<uint4*>(Qs + (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4)) = qq[i];
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = threadIdx.x + 128 * kk;     // key
        const uint32_t w[16] = {vq[kk][0].x, vq[kk][0].y, vq[kk][0].z, vq[kk][0].w, vq[kk][1].x, vq[kk][1].y, vq[kk][1].z, vq[kk][1].w,
                                vq[kk][2].x, vq[kk][2].y, vq[kk][2].z, vq[kk][2].w, vq[kk][3].x, vq[kk][3].y, vq[kk][3].z, vq[kk][3].w};
        uint8_t* slab = Vt + (uint32_t)(k >> 6) * kVtSlab + (uint32_t)(k & 7) * 2u;
        const int kc = (k & 63) >> 3;
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
          const int n0 = 2 * n2, n1 = 2 * n2 + 1;
          *reinterpret_cast<uint16_t*>(slab + n0 * 128 + ((kc ^ (n0 & 7)) << 4)) = (uint16_t)(w[n2] & 0xFFFFu);
          *reinterpret_cast<uint16_t*>(slab + n1 * 128 + ((kc ^ (n1 & 7)) << 4)) = (uint16_t)(w[n2] >> 16);
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    // ---- phase 2: S = Q K^T into an N-column accumulator
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t q_lo = (uint32_t)kDesc64 | ((smem_u32(Qs) & 0x3FFFFu) >> 4);
      const uint32_t k_lo = (uint32_t)kDesc64 | ((smem_u32(Ks) & 0x3FFFFu) >> 4);
      const uint32_t hi = (uint32_t)(kDesc64 >> 32);
      const uint32_t idesc = idesc_bf16(128, N);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) umma_bf16_lohi(tmem, q_lo + ks * 2u, k_lo + ks * 2u, hi, idesc, ks ? 1u : 0u);
      umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    tc_fence_after();
    // ---- phase 3: softmax of my query row, P -> shared memory (bf16, A operand of P V)
    float l = 0.f;
    {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      // (two chunks in flight with four partial maxima / sums per thread was measured: 1.25 vs 1.19 ms per layer, no gain)
      float mx = -INFINITY;
      for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      const float mxs = mx * scale_log2e;
      uint8_t* prow = Ps + (uint32_t)row * 128u;
      for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        float p[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { p[j] = exp2f(fmaf(__uint_as_float(v[j]), scale_log2e, -mxs)); }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // the row sum uses the bf16-rounded probabilities that the P V product will see
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(p[2 * j], p[2 * j + 1]);
          l += __low2float(h2) + __high2float(h2);
          pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        uint8_t* slab = prow + (uint32_t)(c >> 6) * kAtomBytes;
        const int kc = (c & 63) >> 3;
        *reinterpret_cast<uint4*>(slab + (((kc) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(slab + (((kc + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- phase 4: O = P V (32 columns, over the S columns just consumed)
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t p_lo = (uint32_t)kDesc128 | ((smem_u32(Ps) & 0x3FFFFu) >> 4);
      const uint32_t v_lo = (uint32_t)kDesc128 | ((smem_u32(Vt) & 0x3FFFFu) >> 4);
      const uint32_t hi = (uint32_t)(kDesc128 >> 32);
      const uint32_t idesc = idesc_bf16(128, kHD);
      for (int ks = 0; ks < N / 16; ++ks) {
        const uint32_t po = (uint32_t)(ks >> 2) * (kAtomBytes >> 4) + (uint32_t)(ks & 3) * 2u;
        const uint32_t vo = (uint32_t)(ks >> 2) * (kVtSlab >> 4) + (uint32_t)(ks & 3) * 2u;
        umma_bf16_lohi(tmem, p_lo + po, v_lo + vo, hi, idesc, ks ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    tc_fence_after();
    // ---- phase 5: normalise and store my query's head slice
    {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      uint32_t v0[16], v1[16];
      tmem_ld16(taddr, v0);
      tmem_ld16(taddr + 16, v1);
      tmem_ld_wait();
      if (qi < N) {
        const float inv = 1.0f / l;
        __nv_bfloat16* dst = out + (b * N + qi) * H + h * kHD;
        uint32_t pa[8], pb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          pa[j] = pack_bf16(__uint_as_float(v0[2 * j]) * inv, __uint_as_float(v0[2 * j + 1]) * inv);
          pb[j] = pack_bf16(__uint_as_float(v1[2 * j]) * inv, __uint_as_float(v1[2 * j + 1]) * inv);
        }
        st_global_256(dst, pa);          // the head slice of a query row: two 32-byte sectors
        st_global_256(dst + 16, pb);
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

}  // namespace

// ---- host launchers (used by csn.cu) --------------------------------------------------------------
int csn_tc_convert_weights(const float* src, __nv_bfloat16* dst, int64_t n) {
  f32_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256>>>(src, dst, n);
  DBS_LAUNCHED();
  return DBS_OK;
}

int csn_tc_linear(const void* X, int in_bf16, const __nv_bfloat16* W, const float* bias, int N_total, const float* pre_g,
                  const float* pre_b, const float* post_g, const float* post_b, int silu, const float* tbias,
                  int64_t tb_stride, int tokens_per_cfg, const float* res, void* out, int out_bf16, int64_t M,
                  cudaStream_t st, const float* pos_freqs, const float* head_w2, const float* head_b2,
                  const __nv_bfloat16* W2, const float* bias2, const float* post2_g, const float* post2_b) {
  DBS_REQUIRE(N_total % 128 == 0 && (!post_g || N_total == 128), "csn_tc_linear: unsupported width %d", N_total);
  DBS_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(X)) & 31) == 0,
              "csn_tc_linear: tensors must be 32-byte aligned");
  DBS_REQUIRE(!W2 || (N_total == 128 && post_g && post2_g && bias2 && !in_bf16 && !out_bf16 && !res && !head_w2 && !pos_freqs),
              "csn_tc_linear: the fused MLPBlock form needs fp32 in / out, both LayerNorms and no other fusion");
  TcLinArgs a;
  a.X = X; a.W = W; a.bias = bias; a.N_total = N_total; a.pre_g = pre_g; a.pre_b = pre_b; a.post_g = post_g; a.post_b = post_b;
  a.silu = silu; a.tbias = tbias; a.tb_stride = tb_stride; a.tokens_per_cfg = tokens_per_cfg; a.res = res; a.out = out;
  a.in_bf16 = in_bf16; a.out_bf16 = out_bf16; a.M = M;
  a.pos_freqs = pos_freqs; a.head_w2 = head_w2; a.head_b2 = head_b2;
  a.W2 = W2; a.bias2 = bias2; a.post2_g = post2_g; a.post2_b = post2_b;
  static const bool ws_off = [] { const char* e = getenv("DBS_CSN_WS"); return e && e[0] == '0'; }();
  // (the Fourier-embedding layer stays on the phase-serial kernel: its staging is sinf / cosf work that wants all the
  // SM's warps, measured 0.52 ms there against 1.04 ms with eight producer warps)
  if (!W2 && !ws_off && N_total <= 384 && !pos_freqs && !res && !head_w2) {
    const int nblk = N_total / 128;
    const size_t smem_ws = (size_t)(nblk + (nblk > 1 ? 2 : kWsStages)) * 2 * kAtomBytes + 128 * ((out_bf16 ? 256 : 512) + 16) + 1024;
    static bool ws_attr_done = false;
    static int n_sm = 148;
    if (!ws_attr_done) {
      int dev = 0;
      DBS_CUDA(cudaGetDevice(&dev));
      DBS_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
      DBS_CUDA(cudaFuncSetAttribute(csn_tc_linear_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 10 * 1024));
      ws_attr_done = true;
    }
    const int64_t tiles = (M + 127) / 128;
    prof_set_label("tc linear ws N%d ln%d/%d", N_total, pre_g ? 1 : 0, post_g ? 1 : 0);
    ProfScope prof(kProfCsnLinear, st);
    csn_tc_linear_ws_kernel<<<(unsigned)(tiles < n_sm ? tiles : n_sm), kWsThreads, smem_ws, st>>>(a);
    DBS_LAUNCHED();
    return DBS_OK;
  }
  const size_t smem = (W2 ? 6 : 4) * kAtomBytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    DBS_CUDA(cudaFuncSetAttribute(csn_tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(6 * kAtomBytes + 1024)));
    attr_done = true;
  }
  const int64_t n_tiles = (M + 127) / 128;
  const int by = N_total / 128;
  int64_t bx = (148 * 3 + by - 1) / by;
  if (bx > n_tiles) bx = n_tiles;
  prof_set_label("tc linear N%d ln%d/%d%s", N_total, pre_g ? 1 : 0, post_g ? 1 : 0, W2 ? " +fc2 (MLPBlock)" : "");
  ProfScope prof(kProfCsnLinear, st);
  csn_tc_linear_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, smem, st>>>(a);
  DBS_LAUNCHED();
  return DBS_OK;
}

int csn_tc_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int64_t B, int N, int heads, cudaStream_t st) {
  DBS_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "csn_tc_attention: N = %d not in {16, 32, ..., 256}", N);
  const size_t smem = kQBytes + kQBytes / 2 + 4 * kVtSlab + kPTile + 1024;
  static bool attr_done = false;
  static int n_sm = 148;
  if (!attr_done) {
    int dev = 0;
    DBS_CUDA(cudaGetDevice(&dev));
    DBS_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    DBS_CUDA(cudaFuncSetAttribute(csn_tc_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int n_qt = (N + 127) / 128;
  const int64_t n_units = B * heads * n_qt;
  const int grid = (int)(n_units < 2 * n_sm ? n_units : 2 * n_sm);
  prof_set_label("tc attention N%d heads%d", N, heads);
  ProfScope prof(kProfCsnAttention, st);
  csn_tc_attention_kernel<<<grid, 128, smem, st>>>(qkv, out, n_units, N, heads, n_qt, (1.0f / sqrtf((float)kHD)) * 1.4426950408889634f);
  DBS_LAUNCHED();
  return DBS_OK;
}

}  // namespace dbs
