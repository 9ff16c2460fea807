// conv_tc.cu — placeholder until the tcgen05 kernel lands in this file (next commit).
#include "unet.cuh"
namespace dbs {
int conv_tc_launch(const ConvGeom&, const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, const float*,
                   __nv_bfloat16*, float*, int, cudaStream_t) {
  return DBS_ERR_UNSUPPORTED;
}
}  // namespace dbs
extern "C" int dbs_conv_tc(const void*, int, const void*, int, const void*, const float*, void*, float*, int64_t, int,
                           int, int, int, int, void*) {
  dbs::set_error("dbs_conv_tc: tcgen05 kernel not built yet");
  return DBS_ERR_UNSUPPORTED;
}
