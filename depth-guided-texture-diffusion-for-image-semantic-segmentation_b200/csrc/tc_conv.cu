// tcgen05 implicit-GEMM convolution over NHWC bf16 (decoder convs).  Placeholder entry until the
// TMA-im2col producer lands: fails loudly rather than falling back.
#include "common.cuh"
namespace dgtd {
int tc_conv_nhwc(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int,
                 int, int, int, int, int, int, cudaStream_t) {
  set_error("conv_nhwc(bf16): tcgen05 implicit-GEMM convolution is not built in this version");
  return -4;
}
}  // namespace dgtd
