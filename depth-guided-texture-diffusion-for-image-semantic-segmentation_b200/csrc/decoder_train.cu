// Training backward of the ShapePropDecoder bank (cod.py:1210-1226 + the folded prompt injection,
// cod.py:1471): data-movement halves of the convolution gradients.  The contractions themselves
// run in the GEMM families (simt_gemm.cuh exact fp32 / tcgen05 bf16):
//   input gradient   dcol[M_out, ks*ks*Ct] = g[M_out, E] . W[E, ks*ks*Ct]     then col2im (this file)
//   weight gradient  dW^T[ks*ks*Ct, E]     = im2col(x)^T . g                   im2col^T (this file) + split-K GEMM
#include "common.cuh"

namespace dgtd {

// out[b,iy,ix,c] = mask > 0 ? sum_{ty,tx : iy = oy*stride+off+ty, ix = ox*stride+off+tx}
//                                  dcol[(b,oy,ox)][(ty*ks+tx)*Ct + c] : 0
// gather form (deterministic, no atomics).  ks = 1, stride = 1, off = 0 is a strided masked copy
// (the ReLU backward into a channel slice).
template <typename IT, typename OT>
__global__ void __launch_bounds__(256)
col2im_kernel(const IT* __restrict__ dcol, int ldc, int Ct, const OT* __restrict__ mask, int ldm,
              OT* __restrict__ out, int ldo, int h, int w, int C, int ks, int stride, int off, int oh, int ow,
              int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t pix = i / C;
  const int ix = (int)(pix % w);
  const int64_t t = pix / w;
  const int iy = (int)(t % h);
  const int b = (int)(t / h);
  float acc = 0.f;
  const bool live = !mask || to_float(mask[pix * ldm + c]) > 0.f;
  if (live) {
    for (int ty = 0; ty < ks; ++ty) {
      const int ny = iy - off - ty;
      if (ny < 0 || ny % stride) continue;
      const int oy = ny / stride;
      if (oy >= oh) continue;
      for (int tx = 0; tx < ks; ++tx) {
        const int nx = ix - off - tx;
        if (nx < 0 || nx % stride) continue;
        const int ox = nx / stride;
        if (ox >= ow) continue;
        acc += to_float(dcol[(((int64_t)b * oh + oy) * ow + ox) * ldc + (ty * ks + tx) * Ct + c]);
      }
    }
  }
  store1(out + pix * ldo + c, acc);
}

// outT[(tap*32 + c)][m] = x[b, oy*stride+off+ty, ox*stride+off+tx, c]   (zero outside the map),
// m = (b*oh + oy)*ow + ox, c < 32: the K-major operand of the tcgen05 weight-gradient GEMM.
// CTA = 64 pixels x 32 channels of one tap, transposed through shared memory.
__global__ void __launch_bounds__(256)
im2col_t_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ outT, int64_t M, int h,
                int w, int ks, int stride, int off, int oh, int ow) {
  __shared__ float t[64][33];
  const int tap = blockIdx.y, ty = tap / ks, tx = tap - ty * ks;
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  {
    const int c = threadIdx.x & 31;
    for (int mi = threadIdx.x >> 5; mi < 64; mi += 8) {
      const int64_t m = m0 + mi;
      float v = 0.f;
      if (m < M) {
        const int ox = (int)(m % ow);
        const int64_t r = m / ow;
        const int oy = (int)(r % oh);
        const int b = (int)(r / oh);
        const int iy = oy * stride + off + ty, ix = ox * stride + off + tx;
        if ((unsigned)iy < (unsigned)h && (unsigned)ix < (unsigned)w)
          v = __bfloat162float(x[(((int64_t)b * h + iy) * w + ix) * ldx + c]);
      }
      t[mi][c] = v;
    }
  }
  __syncthreads();
  {
    const int mp = threadIdx.x & 31;
    const int64_t m = m0 + 2 * mp;
    for (int c = threadIdx.x >> 5; c < 32; c += 8) {
      __nv_bfloat16* o = outT + ((int64_t)tap * 32 + c) * M + m;
      if (m + 1 < M && (M & 1) == 0)
        *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(t[2 * mp][c], t[2 * mp + 1][c]);
      else {
        if (m < M) o[0] = __float2bfloat16_rn(t[2 * mp][c]);
        if (m + 1 < M) o[1] = __float2bfloat16_rn(t[2 * mp + 1][c]);
      }
    }
  }
}

// out[m][c] = sum_g x[m][g*gs + c], c < C   (input gradients of the decoders' first convs, summed over decoders)
template <typename IT>
__global__ void group_sum_kernel(const IT* __restrict__ x, float* __restrict__ out, int64_t M, int G, int gs, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const int c = (int)(i % C);
  const int64_t m = i / C;
  const IT* p = x + m * (int64_t)G * gs + c;
  float s = 0.f;
  for (int g = 0; g < G; ++g) s += to_float(p[(int64_t)g * gs]);
  out[i] = s;
}

}  // namespace dgtd

using namespace dgtd;

extern "C" {

int dgtd_col2im_nhwc(const void* dcol, int dcol_dtype, int ldc, int Ct, const void* mask, int ldm, void* out,
                     int out_dtype, int ldo, int B, int h, int w, int C, int ks, int stride, int off, int oh, int ow,
                     dgtd_stream_t stream) {
  DGTD_CHECK_ARG(dcol && out && B > 0 && h > 0 && w > 0 && C > 0 && C <= Ct && ks >= 1 && stride >= 1 && oh > 0 && ow > 0,
                 "col2im_nhwc: bad args");
  DGTD_CHECK_ARG(ldc >= ks * ks * Ct && ldo >= C && (!mask || ldm >= C), "col2im_nhwc: bad pitches");
  const int64_t total = (int64_t)B * h * w * C;
  const unsigned blocks = (unsigned)cdiv(total, (int64_t)256);
  cudaStream_t s = (cudaStream_t)stream;
#define DGTD_C2I(IT, OT)                                                                                              \
  col2im_kernel<IT, OT><<<blocks, 256, 0, s>>>((const IT*)dcol, ldc, Ct, (const OT*)mask, ldm, (OT*)out, ldo, h, w, C, \
                                                ks, stride, off, oh, ow, total)
  if (dcol_dtype == DGTD_F32 && out_dtype == DGTD_F32) DGTD_C2I(float, float);
  else if (dcol_dtype == DGTD_F32 && out_dtype == DGTD_BF16) DGTD_C2I(float, __nv_bfloat16);
  else if (dcol_dtype == DGTD_BF16 && out_dtype == DGTD_BF16) DGTD_C2I(__nv_bfloat16, __nv_bfloat16);
  else if (dcol_dtype == DGTD_BF16 && out_dtype == DGTD_F32) DGTD_C2I(__nv_bfloat16, float);
  else DGTD_CHECK_ARG(false, "col2im_nhwc: bad dtypes %d -> %d", dcol_dtype, out_dtype);
#undef DGTD_C2I
  DGTD_LAUNCH_CHECK("col2im_nhwc");
  return 0;
}

// x: NHWC bf16, the 32-channel slice starting at the pointer (pixel pitch ldx); outT: (ks*ks*32) rows x M, M = B*oh*ow
int dgtd_im2col_t(const void* x, int ldx, void* outT, int B, int h, int w, int ks, int stride, int off, int oh, int ow,
                  dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && outT && B > 0 && h > 0 && w > 0 && ks >= 1 && ks <= 8 && stride >= 1 && oh > 0 && ow > 0 && ldx >= 32,
                 "im2col_t: bad args");
  const int64_t M = (int64_t)B * oh * ow;
  im2col_t_kernel<<<dim3((unsigned)cdiv(M, (int64_t)64), ks * ks), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)outT, M, h, w, ks, stride, off, oh, ow);
  DGTD_LAUNCH_CHECK("im2col_t");
  return 0;
}

int dgtd_group_sum(const void* x, int dtype, float* out, int64_t M, int groups, int group_stride, int C,
                   dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out && M > 0 && groups > 0 && C > 0 && C <= group_stride, "group_sum: bad args");
  const unsigned blocks = (unsigned)cdiv(M * C, (int64_t)256);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DGTD_BF16)
    group_sum_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, out, M, groups, group_stride, C);
  else if (dtype == DGTD_F32)
    group_sum_kernel<float><<<blocks, 256, 0, s>>>((const float*)x, out, M, groups, group_stride, C);
  else DGTD_CHECK_ARG(false, "group_sum: bad dtype %d", dtype);
  DGTD_LAUNCH_CHECK("group_sum");
  return 0;
}

}  // extern "C"
