// unet.cuh — shared declarations of the score U-Net executor (unet.cu) and the tcgen05
// implicit-GEMM convolution (conv_tc.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace dbs {

// One convolution as an implicit GEMM over a "pixel grid" (Hm x Wm per sample):
//   out[n, a*os+oph, b*os+opw, co] = bias[co] +
//       sum_t sum_c in[n, a*is+dh[t], b*is+dw[t], c] * w[widx[t]][co][c]      (zero outside the image)
// which covers Conv2d k in {1,3} stride {1,2} and, phase by phase, ConvTranspose2d(4, 2, 1).
// The input is the channel concat [in0 (C0) | in1 (C1)] (score_network.py:256), NHWC.
struct ConvGeom {
  int B, Hin, Win;    // input spatial size
  int Hm, Wm;         // GEMM pixel grid
  int Hout, Wout;     // output spatial size
  int C0, C1, Cout;
  int in_stride;      // 1 or 2
  int out_stride;     // 1 or 2 (transposed conv)
  int oph, opw;       // output phase
  int n_taps;
  signed char dh[16], dw[16];
  unsigned char widx[16];
  // n_phases == 4: all four output phases of a ConvTranspose2d(4, 2, 1) in one geometry — phase ph4 = 2*ph + pw
  // uses taps [ph4*n_taps, (ph4+1)*n_taps) and writes output pixels (2a + ph, 2b + pw); oph / opw are ignored.
  // n_phases == 2: the two horizontal phases pw = 0, 1 of vertical phase oph — taps [0, n_taps) belong to pw = 0,
  // [n_taps, 2 n_taps) to pw = 1 (strip kernel only: the pair shares its A rows, conv_strip.cu).
  // 0 or 1: a single phase / an ordinary conv.  The tile kernel (conv_tc_launch) takes only the 4-phase form.
  int n_phases;
  // s2d == 1 (strip kernel only): a 3x3 stride-2 pad-1 conv run as a 2x2 stride-1 conv on the space-to-depth view of its
  // input: Hin x Win are the HALVED sizes, C0 = 4 * C with channel k = py * 2C + px * C + c standing for input pixel
  // (2a + py, 2b + px), taps (dh', dw') in {-1, 0}^2, weights repacked (with zeros) by repack_s2d_kernel.  The strips then
  // cost 9 * Win rows of TMA per 4C-chunk and tile instead of one box per tap (the stride-2 convs are TMA-row-bound).
  int s2d;
};

struct ConvWeights {
  int cin = 0, cout = 0, taps = 0;
  float* w32 = nullptr;           // [taps][cout][cin]
  __nv_bfloat16* w16 = nullptr;   // same layout
  float* bias = nullptr;          // [cout]
  __nv_bfloat16* w16_s2d = nullptr;   // stride-2 3x3 convs: [4 taps (dh', dw')][cout][4 * cin], see ConvGeom::s2d
};

// GroupNorm apply fused into the conv epilogue (layers whose images fit inside one 128-pixel tile):
//   out = SiLU(GN(conv + bias) * gamma + beta) [+ time bias] [+ residual]      (score_network.py:143-151)
struct GnFuse {
  const float *gamma, *beta;   // [Cout]
  const float* tb;             // time bias row of sample 0 (already offset to this block), null = none
  int64_t tb_stride;           // floats between samples' rows; 0 = one shared row
  const __nv_bfloat16* res;    // residual, same NHWC shape as out, null = none
};

// tcgen05 path.  Returns DBS_OK, or DBS_ERR_UNSUPPORTED if this geometry is not eligible
// (caller then uses the CUDA-core kernel).  stats: fp32 [B][G][2] raw (sum, sumsq) or null.
int conv_tc_launch(const ConvGeom& g, const __nv_bfloat16* in0, const __nv_bfloat16* in1,
                   const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out, float* stats,
                   int groups, cudaStream_t st, const GnFuse* fuse = nullptr);

// Strip variant of the tcgen05 path for stride-1 convs with Cout <= 64 on full-width row tiles: one
// halo strip of the input per channel chunk serves every tap, weights stay resident in shared memory
// when they fit.  Same contract as conv_tc_launch (DBS_ERR_UNSUPPORTED if not eligible).
// With w_res != nullptr the ConvBlock's 1x1 residual conv of the SAME input (score_network.py:131,151) is
// computed in the same pass (extra weight slab on the dh = 0 group, own accumulator, own output).
// With pre != nullptr the input is the RAW output of a conv whose GroupNorm(8) sums are pre_stats [B][8][2]: dedicated
// transform warps rewrite each landed strip in shared memory as SiLU(GN(.)) + time bias [+ pre->res] (pre->gamma / beta /
// tb / res) before the MMAs read it — the GroupNorm prologue: the normalised activation never exists in HBM.
int conv_strip_launch(const ConvGeom& g, const __nv_bfloat16* in0, const __nv_bfloat16* in1,
                      const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out, float* stats,
                      int groups, cudaStream_t st, const __nv_bfloat16* w_res = nullptr,
                      const float* bias_res = nullptr, __nv_bfloat16* out_res = nullptr,
                      const GnFuse* pre = nullptr, const float* pre_stats = nullptr);

// geometry builders
ConvGeom geom_conv(int B, int H, int W, int C0, int C1, int Cout, int ksize, int stride);
ConvGeom geom_convT_phase(int B, int H, int W, int Cin, int Cout, int ph, int pw);
ConvGeom geom_convT_all(int B, int H, int W, int Cin, int Cout);
ConvGeom geom_convT_pair(int B, int H, int W, int Cin, int Cout, int ph);
ConvGeom geom_conv_s2d(int B, int H, int W, int C, int Cout);   // H, W: the conv's (even) input size

}  // namespace dbs
