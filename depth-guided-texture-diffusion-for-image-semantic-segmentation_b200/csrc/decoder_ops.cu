// a10/a11: ShapePropDecoder convolutions (cod.py:1216-1222) as implicit GEMM over NHWC, and the
// last conv folded with the bilinear down-sample of the prompt injection (cod.py:1471).
// fp32 = CUDA-core exact path (im2col gather inside the GEMM's A loader, nothing materialised);
// bf16 = tcgen05 implicit GEMM (tc_conv.cu).
#include "simt_gemm.cuh"

namespace dgtd {
int tc_conv_nhwc(const void* x, const void* w, const float* bias, void* out, int B, int h, int wd, int Cin,
                 int ldx, int oh, int ow, int Cout, int ldo, int ks, int stride, int off, int act,
                 int dtype_out, int groups, int64_t x_group_stride, int w_group_rows, int64_t out_group_stride,
                 cudaStream_t s);
}
using namespace dgtd;

extern "C" {

int dgtd_conv_nhwc_grouped_fwd(const void* x, const void* w, const float* bias, void* out, int B, int h,
                               int wd, int Cin, int ldx, int oh, int ow, int Cout, int ldo, int ks,
                               int stride, int off, int act, int dtype_in, int dtype_out, int groups,
                               int64_t x_group_stride, int w_group_rows, int64_t out_group_stride,
                               dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && w && out, "conv_nhwc: null pointer");
  DGTD_CHECK_ARG(B > 0 && h > 0 && wd > 0 && oh > 0 && ow > 0 && Cin > 0 && Cout > 0 && ks > 0 && stride > 0 &&
                     groups > 0,
                 "conv_nhwc: bad shape");
  const bool out_split = groups == 1 && out_group_stride > 0 && ldo < Cout;   // 32-channel output chunks scattered group-major
  DGTD_CHECK_ARG(ldx >= Cin && (ldo >= Cout || (out_split && ldo >= 32)), "conv_nhwc: leading dims too small");
  DGTD_CHECK_ARG(!out_split || dtype_in == DGTD_BF16, "conv_nhwc: the chunk-scattered output is a bf16 (tcgen05) mode");
  DGTD_CHECK_ARG(act == DGTD_ACT_NONE || act == DGTD_ACT_RELU, "conv_nhwc: activation must be none or relu");
  DGTD_CHECK_ARG(groups == 1 || w_group_rows >= Cout, "conv_nhwc: w_group_rows < Cout");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == DGTD_BF16) {
    DGTD_CHECK_ARG(groups == 1 || (x_group_stride >= 32 && x_group_stride % 8 == 0),
                   "conv_nhwc(bf16): x_group_stride must be 32 (interleaved slices) or a multiple of 8 >= 32 (group-major)");
    int rc = tc_conv_nhwc(x, w, bias, out, B, h, wd, Cin, ldx, oh, ow, Cout, ldo, ks, stride, off, act, dtype_out,
                          groups, x_group_stride, groups == 1 ? Cout : w_group_rows, out_group_stride, s);
    if (rc) return rc;
    DGTD_LAUNCH_CHECK("conv_nhwc(tcgen05)");
    return 0;
  }
  DGTD_CHECK_ARG(dtype_in == DGTD_F32 && dtype_out == DGTD_F32, "conv_nhwc(fp32): fp32 in/out only");
  DGTD_CHECK_ARG(Cin % 4 == 0 && ldx % 4 == 0 && Cout % 4 == 0 && ldo % 4 == 0 && x_group_stride % 4 == 0 &&
                     out_group_stride % 4 == 0,
                 "conv_nhwc(fp32): channel counts / strides must be multiples of 4");
  const int64_t M64 = (int64_t)B * oh * ow;
  DGTD_CHECK_ARG(M64 < (1ll << 31), "conv_nhwc: too many output pixels");
  const int M = (int)M64, K = ks * ks * Cin;
  for (int g = 0; g < groups; ++g) {
    Im2colLoader al{(const float*)x + (int64_t)g * x_group_stride, h, wd, ldx, Cin, oh, ow, ks, stride, off, M, K};
    RowMajorLoader bl{(const float*)w + (int64_t)g * w_group_rows * K, K, 0, Cout, K};
    const float* bg = bias ? bias + (int64_t)g * w_group_rows : nullptr;
    float* og = (float*)out + g * out_group_stride;
    if (act == DGTD_ACT_RELU) {
      EpiBiasAct<float, DGTD_ACT_RELU> ep{og, bg, ldo};
      launch_simt_gemm<true, true>(al, bl, ep, M, Cout, K, 1, s);
    } else {
      EpiBiasAct<float, DGTD_ACT_NONE> ep{og, bg, ldo};
      launch_simt_gemm<true, true>(al, bl, ep, M, Cout, K, 1, s);
    }
    DGTD_LAUNCH_CHECK("conv_nhwc(fp32)");
  }
  return 0;
}

int dgtd_conv_nhwc_fwd(const void* x, const void* w, const float* bias, void* out, int B, int h, int wd,
                       int Cin, int ldx, int oh, int ow, int Cout, int ldo, int ks, int stride, int off,
                       int act, int dtype_in, int dtype_out, dgtd_stream_t stream) {
  return dgtd_conv_nhwc_grouped_fwd(x, w, bias, out, B, h, wd, Cin, ldx, oh, ow, Cout, ldo, ks, stride, off,
                                    act, dtype_in, dtype_out, 1, 0, Cout, 0, stream);
}

}  // extern "C"
