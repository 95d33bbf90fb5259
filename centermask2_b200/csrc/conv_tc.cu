// placeholder until the tcgen05 engine lands (replaced in the next commit)
#include "common.cuh"
namespace cm2 {
int conv_tc_launch(const cm2_conv_desc* d, cudaStream_t stream) {
  set_error("conv2d: tensor-core engine not built");
  return CM2_ERR_UNSUPPORTED;
}
}  // namespace cm2
extern "C" int64_t cm2_conv_tc_klen(int32_t kh, int32_t kw, int32_t num_src, const int32_t* src_c) {
  int64_t k = 0;
  for (int i = 0; i < num_src; ++i) k += (src_c[i] + 63) / 64 * 64;
  return k * kh * kw;
}
extern "C" int cm2_conv_tc_supported(const cm2_conv_desc* d) { return 0; }
