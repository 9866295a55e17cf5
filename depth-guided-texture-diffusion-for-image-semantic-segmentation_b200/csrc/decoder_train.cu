// Training backward of the ShapePropDecoder bank (cod.py:1210-1226 + the folded prompt injection,
// cod.py:1471): data-movement halves of the convolution gradients.  The contractions themselves
// run in the GEMM families (simt_gemm.cuh exact fp32 / tcgen05 bf16):
//   input gradient   dcol[M_out, ks*ks*Ct] = g[M_out, E] . W[E, ks*ks*Ct]     then col2im (this file)
//   weight gradient  dW^T[ks*ks*Ct, E]     = im2col(x)^T . g                   im2col (this file) + split-K GEMM (MN-major)
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

// Vector form: thread = 8 consecutive channels of one pixel (16-byte loads of bf16 dcol, one index decode per 8
// outputs instead of per output -- the scalar kernel above ran at ~5 % of its HBM floor, 2.6 ms of the training step).
__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void ld8f(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void st8f(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void st8f(__nv_bfloat16* p, const float (&f)[8]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = u;
}
template <typename IT, typename OT>
__global__ void __launch_bounds__(256)
col2im_vec8_kernel(const IT* __restrict__ dcol, int ldc, int Ct, const OT* __restrict__ mask, int ldm,
                   OT* __restrict__ out, int ldo, int h, int w, int C, int ks, int stride, int off, int oh, int ow,
                   int64_t total8) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int cq = C >> 3;
  const int c = (int)(i % cq) * 8;
  const int64_t pix = i / cq;
  const int ix = (int)(pix % w);
  const int64_t t = pix / w;
  const int iy = (int)(t % h);
  const int b = (int)(t / h);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
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
      float v[8];
      ld8f(dcol + (((int64_t)b * oh + oy) * ow + ox) * ldc + (ty * ks + tx) * Ct + c, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
  if (mask) {
    float m[8];
    ld8f(mask + pix * ldm + c, m);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = m[e] > 0.f ? acc[e] : 0.f;
  }
  st8f(out + pix * ldo + c, acc);
}

// col[m][tap*C + c] = x[b, oy*stride+off+ty, ox*stride+off+tx, c]   (zero outside the map),
// m = (b*oh + oy)*ow + ox, c < C (C a multiple of 8): the row-major im2col matrix -- an MN-major operand of
// dgtd_wgrad_tc_mn (decoder weight gradients, C = 32) or the A operand of a patch-embed GEMM.
// Thread = 8 channels (one 16-byte load and store).
__global__ void __launch_bounds__(256)
im2col_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ col, int64_t total, int h,
              int w, int C, int ks, int stride, int off, int oh, int ow) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cq = C >> 3;
  const int chunk = (int)(i % cq);
  const int taps = ks * ks;
  const int tap = (int)((i / cq) % taps);
  const int64_t m = (i / cq) / taps;
  const int ty = tap / ks, tx = tap - ty * ks;
  const int ox = (int)(m % ow);
  const int64_t r = m / ow;
  const int oy = (int)(r % oh);
  const int b = (int)(r / oh);
  const int iy = oy * stride + off + ty, ix = ox * stride + off + tx;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if ((unsigned)iy < (unsigned)h && (unsigned)ix < (unsigned)w)
    v = *reinterpret_cast<const uint4*>(x + (((int64_t)b * h + iy) * w + ix) * ldx + chunk * 8);
  *reinterpret_cast<uint4*>(col + (m * taps + tap) * C + chunk * 8) = v;
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
  {   // 8-channel vector form when every pitch / pointer allows 16-byte (bf16) or 32-byte (fp32) accesses
    const size_t ie = dcol_dtype == DGTD_BF16 ? 2 : 4, oe = out_dtype == DGTD_BF16 ? 2 : 4;
    const bool vec = C % 8 == 0 && Ct % 8 == 0 && ldc % 8 == 0 && ldo % 8 == 0 && (!mask || ldm % 8 == 0) &&
                     (reinterpret_cast<uintptr_t>(dcol) % (8 * ie)) == 0 && (reinterpret_cast<uintptr_t>(out) % (8 * oe)) == 0 &&
                     (!mask || (reinterpret_cast<uintptr_t>(mask) % (8 * oe)) == 0) &&
                     (dcol_dtype == DGTD_BF16 || dcol_dtype == DGTD_F32) && (out_dtype == DGTD_BF16 || out_dtype == DGTD_F32);
    if (vec) {
      const int64_t total8 = total / 8;
      const unsigned vb = (unsigned)cdiv(total8, (int64_t)256);
#define DGTD_C2V(IT, OT)                                                                                               \
  col2im_vec8_kernel<IT, OT><<<vb, 256, 0, s>>>((const IT*)dcol, ldc, Ct, (const OT*)mask, ldm, (OT*)out, ldo, h, w, C, \
                                                ks, stride, off, oh, ow, total8)
      if (dcol_dtype == DGTD_F32 && out_dtype == DGTD_F32) DGTD_C2V(float, float);
      else if (dcol_dtype == DGTD_F32 && out_dtype == DGTD_BF16) DGTD_C2V(float, __nv_bfloat16);
      else if (dcol_dtype == DGTD_BF16 && out_dtype == DGTD_BF16) DGTD_C2V(__nv_bfloat16, __nv_bfloat16);
      else DGTD_C2V(__nv_bfloat16, float);
#undef DGTD_C2V
      DGTD_LAUNCH_CHECK("col2im_nhwc");
      return 0;
    }
  }
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

// x: NHWC bf16, the C-channel slice starting at the pointer (pixel pitch ldx, multiple of 8, 16-byte aligned);
// col: (B*oh*ow) rows x (ks*ks*C)
int dgtd_im2col_nhwc(const void* x, int ldx, void* col, int B, int h, int w, int C, int ks, int stride, int off, int oh,
                     int ow, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && col && B > 0 && h > 0 && w > 0 && ks >= 1 && ks <= 8 && stride >= 1 && oh > 0 && ow > 0,
                 "im2col_nhwc: bad args");
  DGTD_CHECK_ARG(C >= 8 && C % 8 == 0 && ldx >= C && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                 "im2col_nhwc: C and the pixel pitch must be multiples of 8, x 16-byte aligned");
  const int64_t total = (int64_t)B * oh * ow * ks * ks * (C / 8);
  im2col_kernel<<<(unsigned)cdiv(total, (int64_t)256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)col, total, h, w, C, ks, stride, off, oh, ow);
  DGTD_LAUNCH_CHECK("im2col_nhwc");
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
