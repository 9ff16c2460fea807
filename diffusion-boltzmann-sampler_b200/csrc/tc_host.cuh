// tc_host.cuh — host-side helpers shared by the tensor-core convolution kernels: the
// cuTensorMapEncodeTiled entry point (fetched through the runtime, so the library does not link
// libcuda) and the tensor-map builders for NHWC activations and repacked weights.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dbs {
namespace tch {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();
int num_sms();
bool is_pow2(int v);
// activation map: stride 1 -> {C, W, 1, H, N}; stride 2 -> space-to-depth view {2C, W/2, 2, H/2, N};
// box {KC, Wt, 1, Ht, Nt}
int encode_act(EncodeTiledFn enc, CUtensorMap* m, const void* ptr, int C, int H, int W, int B, int stride, int KC, int Wt,
               int Ht, int Nt, CUtensorMapSwizzle sw);
// weight map {Cin, Cout, taps}, box {KC, Nb, 1}
int encode_w(EncodeTiledFn enc, CUtensorMap* m, const void* ptr, int Cin, int Cout, int taps, int KC, int Nb,
             CUtensorMapSwizzle sw);

}  // namespace tch
}  // namespace dbs
