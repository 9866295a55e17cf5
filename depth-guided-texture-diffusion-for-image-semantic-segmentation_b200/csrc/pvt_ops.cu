// PVT-v2 blocks that consume the texture prompts (cod.py:824-961, 1455-1509): the
// non-GEMM pieces.  Token streams are (B, N, C) = NHWC, fp32 residual stream, GEMM operands bf16 (or fp32
// in exact mode); every projection runs on the GEMM kernels of the hot path.
//   ln_tokens      x (+ prompt) -> LayerNorm over C      (cod.py:1471-1472 `x + prompt[i]`, :958/:960 norm1/2)
//   patchify       spatial-reduction conv k = stride = sr as a patch gather + GEMM (cod.py:887,903)
//   attention      softmax(q k^T / sqrt(d)) v with N_kv <= a few hundred keys (cod.py:911-915), fp32 math
//   dwconv3_gelu   Mlp's depthwise 3x3 + GELU(erf) on the hidden tokens (cod.py:852-854, 1520-1531)
#include "common.cuh"

namespace dgtd {

template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// One warp per token row.  sum_out (nullable, fp32) receives x + add; out = LN(x + add) * w + b.
template <typename AT, typename OT, int VPL>
__global__ void __launch_bounds__(256)
ln_tokens_kernel(const float* __restrict__ x, const AT* __restrict__ add, float* __restrict__ sum_out,
                 const float* __restrict__ ln_w, const float* __restrict__ ln_b, OT* __restrict__ out, int64_t rows,
                 int C, float eps) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nq = C >> 2;
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int q = j * 32 + lane;
    if (q < nq) {
      v[j] = *reinterpret_cast<const float4*>(x + row * C + q * 4);
      if (add) {
        const float4 a = ld4<AT>(add + row * C + q * 4);
        v[j].x += a.x; v[j].y += a.y; v[j].z += a.z; v[j].w += a.w;
      }
      if (sum_out) *reinterpret_cast<float4*>(sum_out + row * C + q * 4) = v[j];
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  }
  const float mean = warp_sum(s) / C;
  float qq = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    if (j * 32 + lane < nq) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      qq += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(qq) / C + eps);
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int q = j * 32 + lane;
    if (q < nq) {
      const float4 g = *reinterpret_cast<const float4*>(ln_w + q * 4), be = *reinterpret_cast<const float4*>(ln_b + q * 4);
      store4(out + row * C + q * 4, (v[j].x - mean) * rstd * g.x + be.x, (v[j].y - mean) * rstd * g.y + be.y,
             (v[j].z - mean) * rstd * g.z + be.z, (v[j].w - mean) * rstd * g.w + be.w);
    }
  }
}

// out[(b,oy,ox)][(ty*sr+tx)*C + c] = x[b, oy*sr+ty, ox*sr+tx, c]; thread = 4 channels
template <typename T>
__global__ void patchify_tokens_kernel(const T* __restrict__ x, T* __restrict__ out, int h, int w, int C, int sr,
                                       int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cq = C >> 2;
  const int c4 = (int)(i % cq);
  int64_t t = i / cq;
  const int tap = (int)(t % (sr * sr)); t /= sr * sr;
  const int ow = w / sr, oh = h / sr;
  const int ox = (int)(t % ow); t /= ow;
  const int oy = (int)(t % oh);
  const int b = (int)(t / oh);
  const int ty = tap / sr, tx = tap - ty * sr;
  const T* src = x + (((int64_t)b * h + oy * sr + ty) * w + ox * sr + tx) * C + c4 * 4;
  const float4 v = ld4<T>(src);
  store4(out + i * 4, v.x, v.y, v.z, v.w);
}

// depthwise 3x3 (pad 1) + bias + GELU(erf) on NHWC tokens; wT is (9, C); thread = 4 channels of one pixel
template <typename T>
__global__ void dwconv3_gelu_kernel(const T* __restrict__ x, const float* __restrict__ wT, const float* __restrict__ bias,
                                    T* __restrict__ out, int h, int w, int C, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cq = C >> 2;
  const int c = (int)(i % cq) * 4;
  int64_t t = i / cq;
  const int ox = (int)(t % w); t /= w;
  const int oy = (int)(t % h);
  const int b = (int)(t / h);
  float4 acc = *reinterpret_cast<const float4*>(bias + c);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy + ky - 1;
    if ((unsigned)iy >= (unsigned)h) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox + kx - 1;
      if ((unsigned)ix >= (unsigned)w) continue;
      const float4 v = ld4<T>(x + (((int64_t)b * h + iy) * w + ix) * C + c);
      const float4 k = *reinterpret_cast<const float4*>(wT + (ky * 3 + kx) * C + c);
      acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y);
      acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
    }
  }
  store4(out + i * 4, gelu_erf(acc.x), gelu_erf(acc.y), gelu_erf(acc.z), gelu_erf(acc.w));
}

// Attention core, head_dim = 64, fp32 math with an online softmax over key tiles of 64.
// q: (B*N, C) rows, head hd at columns [hd*64, hd*64+64); kv: (B*Nk, 2C) rows, k at [hd*64, ...), v at
// [C + hd*64, ...).  CTA = (query block of 32, head, image), 8 warps, one query at a time per warp.
// K / V tiles sit in shared memory with a 65-float row pitch: lanes = keys read K conflict-free for the
// scores, lanes = channels read V conflict-free for the weighted sum.
constexpr int ATT_D = 64, ATT_TK = 64, ATT_QB = 32;
template <typename T>
__global__ void __launch_bounds__(256)
attention_kernel(const T* __restrict__ q, const T* __restrict__ kv, T* __restrict__ out, int N, int Nk, int C,
                 float scale) {
  __shared__ float Ks[ATT_TK][ATT_D + 1];
  __shared__ float Vs[ATT_TK][ATT_D + 1];
  __shared__ float Qs[8][ATT_D];
  const int hd = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * ATT_QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int QPW = ATT_QB / 8;   // queries per warp
  float m[QPW], l[QPW], o0[QPW], o1[QPW];
#pragma unroll
  for (int i = 0; i < QPW; ++i) { m[i] = -INFINITY; l[i] = 0.f; o0[i] = 0.f; o1[i] = 0.f; }
  const T* kvb = kv + (int64_t)b * Nk * 2 * C + hd * ATT_D;
  for (int k0 = 0; k0 < Nk; k0 += ATT_TK) {
    __syncthreads();
    for (int i = threadIdx.x; i < ATT_TK * (ATT_D / 4); i += 256) {   // stage K and V tiles (zero beyond Nk)
      const int key = i / (ATT_D / 4), dq = (i % (ATT_D / 4)) * 4;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (k0 + key < Nk) {
        kk = ld4<T>(kvb + (int64_t)(k0 + key) * 2 * C + dq);
        vv = ld4<T>(kvb + (int64_t)(k0 + key) * 2 * C + C + dq);
      }
      Ks[key][dq] = kk.x; Ks[key][dq + 1] = kk.y; Ks[key][dq + 2] = kk.z; Ks[key][dq + 3] = kk.w;
      Vs[key][dq] = vv.x; Vs[key][dq + 1] = vv.y; Vs[key][dq + 2] = vv.z; Vs[key][dq + 3] = vv.w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
      const int qi = q0 + warp * QPW + i;
      if (qi >= N) continue;   // warp-uniform
      const T* qp = q + ((int64_t)b * N + qi) * C + hd * ATT_D;
      Qs[warp][lane] = to_float(qp[lane]) * scale;
      Qs[warp][lane + 32] = to_float(qp[lane + 32]) * scale;
      __syncwarp();
      float s0 = 0.f, s1 = 0.f;   // scores of keys lane, lane + 32
#pragma unroll 16
      for (int d = 0; d < ATT_D; ++d) {
        const float qd = Qs[warp][d];
        s0 = fmaf(qd, Ks[lane][d], s0);
        s1 = fmaf(qd, Ks[lane + 32][d], s1);
      }
      if (k0 + lane >= Nk) s0 = -INFINITY;
      if (k0 + lane + 32 >= Nk) s1 = -INFINITY;
      float tm = fmaxf(s0, s1);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, off));
      const float mn = fmaxf(m[i], tm);
      const float corr = __expf(m[i] - mn);
      const float p0 = __expf(s0 - mn), p1 = __expf(s1 - mn);
      float ps = p0 + p1;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l[i] = l[i] * corr + ps;
      m[i] = mn;
      float a0 = o0[i] * corr, a1 = o1[i] * corr;   // channels lane, lane + 32
#pragma unroll 8
      for (int key = 0; key < 32; ++key) {
        const float pk0 = __shfl_sync(0xffffffffu, p0, key), pk1 = __shfl_sync(0xffffffffu, p1, key);
        a0 = fmaf(pk0, Vs[key][lane], a0);
        a1 = fmaf(pk0, Vs[key][lane + 32], a1);
        a0 = fmaf(pk1, Vs[key + 32][lane], a0);
        a1 = fmaf(pk1, Vs[key + 32][lane + 32], a1);
      }
      o0[i] = a0; o1[i] = a1;
      __syncwarp();
    }
  }
#pragma unroll
  for (int i = 0; i < QPW; ++i) {
    const int qi = q0 + warp * QPW + i;
    if (qi >= N) continue;
    T* op = out + ((int64_t)b * N + qi) * C + hd * ATT_D;
    const float inv = 1.0f / l[i];
    store1(op + lane, o0[i] * inv);
    store1(op + lane + 32, o1[i] * inv);
  }
}

}  // namespace dgtd

using namespace dgtd;

extern "C" {

int dgtd_ln_tokens_fwd(const float* x, const void* add, int add_dtype, float* sum_out, const float* ln_w,
                       const float* ln_b, void* out, int out_dtype, int64_t rows, int C, float eps,
                       dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && ln_w && ln_b && out && rows > 0 && C > 0 && C % 4 == 0 && C <= 2048,
                 "ln_tokens: bad args (C must be a multiple of 4, <= 2048)");
  DGTD_CHECK_ARG(!add || add_dtype == DGTD_F32 || add_dtype == DGTD_BF16, "ln_tokens: bad add dtype");
  DGTD_CHECK_ARG(out_dtype == DGTD_F32 || out_dtype == DGTD_BF16, "ln_tokens: bad out dtype");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)cdiv(rows, (int64_t)8);
#define DGTD_LNT(AT, OT, V)                                                                                       \
  ln_tokens_kernel<AT, OT, V><<<blocks, 256, 0, s>>>(x, (const AT*)add, sum_out, ln_w, ln_b, (OT*)out, rows, C, eps)
#define DGTD_LNT_V(AT, OT)                         \
  do {                                             \
    if (C <= 128) DGTD_LNT(AT, OT, 1);             \
    else if (C <= 512) DGTD_LNT(AT, OT, 4);        \
    else DGTD_LNT(AT, OT, 16);                     \
  } while (0)
  const bool abf = add && add_dtype == DGTD_BF16;
  if (out_dtype == DGTD_BF16) {
    if (abf) DGTD_LNT_V(__nv_bfloat16, __nv_bfloat16);
    else DGTD_LNT_V(float, __nv_bfloat16);
  } else {
    if (abf) DGTD_LNT_V(__nv_bfloat16, float);
    else DGTD_LNT_V(float, float);
  }
#undef DGTD_LNT_V
#undef DGTD_LNT
  DGTD_LAUNCH_CHECK("ln_tokens");
  return 0;
}

int dgtd_patchify_tokens_fwd(const void* x, void* out, int dtype, int B, int h, int w, int C, int sr,
                             dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out && B > 0 && sr >= 1 && h % sr == 0 && w % sr == 0 && C % 4 == 0,
                 "patchify_tokens: h, w must be multiples of sr and C of 4");
  const int64_t total = (int64_t)B * h * w * (C / 4);
  const unsigned blocks = (unsigned)cdiv(total, (int64_t)256);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DGTD_BF16)
    patchify_tokens_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, h, w, C, sr, total);
  else if (dtype == DGTD_F32)
    patchify_tokens_kernel<<<blocks, 256, 0, s>>>((const float*)x, (float*)out, h, w, C, sr, total);
  else DGTD_CHECK_ARG(false, "patchify_tokens: bad dtype %d", dtype);
  DGTD_LAUNCH_CHECK("patchify_tokens");
  return 0;
}

int dgtd_dwconv3_gelu_fwd(const void* x, const float* wT, const float* bias, void* out, int dtype, int B, int h, int w,
                          int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && wT && bias && out && B > 0 && h > 0 && w > 0 && C % 4 == 0, "dwconv3_gelu: bad args");
  const int64_t total = (int64_t)B * h * w * (C / 4);
  const unsigned blocks = (unsigned)cdiv(total, (int64_t)256);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DGTD_BF16)
    dwconv3_gelu_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, wT, bias, (__nv_bfloat16*)out, h, w, C, total);
  else if (dtype == DGTD_F32)
    dwconv3_gelu_kernel<<<blocks, 256, 0, s>>>((const float*)x, wT, bias, (float*)out, h, w, C, total);
  else DGTD_CHECK_ARG(false, "dwconv3_gelu: bad dtype %d", dtype);
  DGTD_LAUNCH_CHECK("dwconv3_gelu");
  return 0;
}

int dgtd_attention_fwd(const void* q, const void* kv, void* out, int dtype, int B, int N, int Nk, int heads,
                       float scale, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(q && kv && out && B > 0 && N > 0 && Nk > 0 && heads > 0 && heads <= 65535 && B <= 65535,
                 "attention: bad args");
  const int C = heads * ATT_D;   // head_dim 64 (pvt_v2: 64/1, 128/2, 320/5, 512/8)
  dim3 grid(cdiv(N, ATT_QB), heads, B);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DGTD_BF16)
    attention_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)kv, (__nv_bfloat16*)out, N, Nk,
                                          C, scale);
  else if (dtype == DGTD_F32)
    attention_kernel<<<grid, 256, 0, s>>>((const float*)q, (const float*)kv, (float*)out, N, Nk, C, scale);
  else DGTD_CHECK_ARG(false, "attention: bad dtype %d", dtype);
  DGTD_LAUNCH_CHECK("attention");
  return 0;
}

}  // extern "C"
