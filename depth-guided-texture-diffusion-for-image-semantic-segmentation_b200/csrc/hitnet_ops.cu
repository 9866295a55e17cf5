// the Hitnet iterative decoder (cod.py:685-807) on NHWC fp32, inference semantics.
//
//   conv (+ folded BatchNorm affine) (+ PReLU) (+ residual)   implicit GEMM, operands gathered in the A loader,
//                                                             results written into a channel slice of a wider
//                                                             tensor (torch.cat of cod.py:778,785,789,791 never
//                                                             materialises as a copy of the conv output)
//   CALayer / SAM gates (cod.py:413-429, 454-506)             fixed-order channel sums -> tiny MLP -> sigmoid
//   gated sum                                                 res * gate + x  /  x_h*y_h*w_h + x_l*y_l*w_l
//   bilinear resize, both align_corners conventions           nn.Upsample(align_corners=True) :709,:733,:737
//   1-channel heads out_CFM / out_SAM (cod.py:710-711)        per-pixel dot product
//
// Every reduction is fixed order (no atomics): the masks are bit-stable run to run.
#include "simt_gemm.cuh"

namespace dgtd {

// out = prelu(scale[n] * acc + shift[n]) + residual[m, n]      (every piece optional)
struct EpiAffinePReLU {
  float* out;
  const float* scale;
  const float* shift;
  const float* prelu;      // one slope (nn.PReLU() default), nullable
  const float* residual;   // nullable
  int64_t ldo, ldr;
  __device__ __forceinline__ void operator()(int m, int n, int, float4 v) const {
    if (scale) {
      const float4 s = load4(scale + n);
      v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
    }
    if (shift) {
      const float4 t = load4(shift + n);
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    if (prelu) {
      const float a = __ldg(prelu);
      v.x = v.x >= 0.f ? v.x : v.x * a;
      v.y = v.y >= 0.f ? v.y : v.y * a;
      v.z = v.z >= 0.f ? v.z : v.z * a;
      v.w = v.w >= 0.f ? v.w : v.w * a;
    }
    if (residual) {
      const float4 r = load4(residual + (int64_t)m * ldr + n);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    store4(out + (int64_t)m * ldo + n, v.x, v.y, v.z, v.w);
  }
};

// ---- channel sums: partial[b][chunk][c] = sum over the chunk's rows, fixed order -------------------
constexpr int SUM_ROWS = 256;   // rows per chunk
__global__ void __launch_bounds__(256) channel_sums_kernel(const float* __restrict__ x, int ldx, int hw, int C,
                                                           float* __restrict__ partial, int nchunks) {
  extern __shared__ float4 red[];   // [groups][C/4]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int q = C >> 2;
  const int groups = 256 / q;                       // row groups working in parallel
  const int cq = threadIdx.x % q, g = threadIdx.x / q;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g < groups) {
    const int r0 = chunk * SUM_ROWS, r1 = min(hw, r0 + SUM_ROWS);
    const float* base = x + (int64_t)b * hw * ldx + cq * 4;
    for (int r = r0 + g; r < r1; r += groups) {
      const float4 v = load4(base + (int64_t)r * ldx);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    red[g * q + cq] = acc;
  }
  __syncthreads();
  if (threadIdx.x < q) {
    float4 s = red[threadIdx.x];
    for (int gg = 1; gg < groups; ++gg) {
      const float4 v = red[gg * q + threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(partial + ((int64_t)b * nchunks + chunk) * C + threadIdx.x * 4) = s;
  }
}

// gate[b][o] = sigmoid(sum_j W2[o][j] relu(sum_c W1[j][c] mean[b][c]))
__global__ void __launch_bounds__(128) channel_gate_kernel(const float* __restrict__ partial, int nchunks, float inv_hw,
                                                           const float* __restrict__ w1, const float* __restrict__ w2,
                                                           float* __restrict__ gate, int C, int Cr, int Co) {
  extern __shared__ float sm[];   // mean[C] | hidden[Cr]
  float* mean = sm;
  float* hid = sm + C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nchunks; ++k) s += partial[((int64_t)b * nchunks + k) * C + c];
    mean[c] = s * inv_hw;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(w1[j * C + c], mean[c], s);
    hid[j] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < Co; o += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(w2[o * Cr + j], hid[j], s);
    gate[(int64_t)b * Co + o] = sigmoidf_acc(s);
  }
}

// out = a * ga[b,c] * sa[b] + bb * gb[b,c] * sb[b]
__global__ void __launch_bounds__(256) gated_sum_kernel(const float* __restrict__ a, int lda, const float* __restrict__ ga,
                                                        const float* __restrict__ sa, const float* __restrict__ bb,
                                                        int ldb, const float* __restrict__ gb,
                                                        const float* __restrict__ sb, float* __restrict__ out, int ldo,
                                                        int64_t total_q, int hw, int C) {
  const int q = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_q; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = (int)(i % q);
    const int64_t row = i / q;
    const int b = (int)(row / hw);
    float4 v = load4(a + row * lda + cq * 4);
    if (ga) {
      const float4 g = load4(ga + (int64_t)b * C + cq * 4);
      v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    }
    if (sa) {
      const float s = sa[b];
      v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    }
    if (bb) {
      float4 u = load4(bb + row * ldb + cq * 4);
      if (gb) {
        const float4 g = load4(gb + (int64_t)b * C + cq * 4);
        u.x *= g.x; u.y *= g.y; u.z *= g.z; u.w *= g.w;
      }
      if (sb) {
        const float s = sb[b];
        u.x *= s; u.y *= s; u.z *= s; u.w *= s;
      }
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    store4(out + row * ldo + cq * 4, v.x, v.y, v.z, v.w);
  }
}

// two-tap source of one output coordinate (ATen's area_pixel_compute_source_index)
__device__ __forceinline__ void bilinear_tap(int d, int n_in, float scale, bool align, int& i0, int& i1, float& f) {
  float src = align ? d * scale : fmaxf((d + 0.5f) * scale - 0.5f, 0.f);
  i0 = min((int)src, n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  f = src - (float)i0;
}

__global__ void __launch_bounds__(256) resize_nhwc_ld_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out,
                                                             int ldo, int h, int w, int C, int oh, int ow,
                                                             float sy, float sx, int align, int64_t total_q) {
  const int q = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_q; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = (int)(i % q);
    int64_t p = i / q;
    const int ox = (int)(p % ow);
    p /= ow;
    const int oy = (int)(p % oh);
    const int b = (int)(p / oh);
    int y0, y1, x0, x1;
    float fy, fx;
    bilinear_tap(oy, h, sy, align, y0, y1, fy);
    bilinear_tap(ox, w, sx, align, x0, x1, fx);
    const float* base = x + (int64_t)b * h * w * ldx + cq * 4;
    const float4 v00 = load4(base + ((int64_t)y0 * w + x0) * ldx), v01 = load4(base + ((int64_t)y0 * w + x1) * ldx);
    const float4 v10 = load4(base + ((int64_t)y1 * w + x0) * ldx), v11 = load4(base + ((int64_t)y1 * w + x1) * ldx);
    const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
    store4(out + ((int64_t)(b * oh + oy) * ow + ox) * ldo + cq * 4,
           w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x, w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y,
           w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z, w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w);
  }
}

__global__ void __launch_bounds__(256) copy_channels_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out,
                                                            int ldo, int C, int64_t total_q) {
  const int q = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_q; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = (int)(i % q);
    const int64_t row = i / q;
    const float4 v = load4(x + row * ldx + cq * 4);
    store4(out + row * ldo + cq * 4, v.x, v.y, v.z, v.w);
  }
}

// out[row] (+)= bias + sum_c x[row, c] * w[c]        8 lanes per row, float4 each, shuffle tree inside the octet
__global__ void __launch_bounds__(256) head1_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                                    const float* __restrict__ bias, float* __restrict__ out,
                                                    int64_t rows, int C, int accumulate) {
  const int lane8 = threadIdx.x & 7;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  float s = 0.f;
  if (row < rows) {
    for (int c = lane8 * 4; c < C; c += 32) {
      const float4 v = load4(x + row * ldx + c), ww = load4(w + c);
      s = fmaf(v.x, ww.x, s);
      s = fmaf(v.y, ww.y, s);
      s = fmaf(v.z, ww.z, s);
      s = fmaf(v.w, ww.w, s);
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (row < rows && lane8 == 0) {
    s += bias ? __ldg(bias) : 0.f;
    out[row] = accumulate ? out[row] + s : s;
  }
}

// bf16 mode: operand of the tcgen05 GEMM.  col[m, (tap, c)] = bf16(prelu(x[b, oy*stride+off+ty, ox*stride+off+tx, c])),
// zero outside the image; 8 channels (one 16-byte store) per thread, consecutive threads walk K.
__global__ void __launch_bounds__(256) im2col_act_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ col,
                                                         const float* __restrict__ prelu, int64_t total, int h, int w, int C,
                                                         int ks, int stride, int off, int oh, int ow) {
  const int cg = C >> 3, taps = ks * ks;
  const float a = prelu ? __ldg(prelu) : 1.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg);
    int64_t t = i / cg;
    const int tap = (int)(t % taps);
    const int64_t m = t / taps;
    const int ox = (int)(m % ow);
    const int64_t t2 = m / ow;
    const int oy = (int)(t2 % oh);
    const int b = (int)(t2 / oh);
    const int ty = tap / ks, tx = tap - ty * ks;
    const int iy = oy * stride + off + ty, ix = ox * stride + off + tx;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if ((unsigned)iy < (unsigned)h && (unsigned)ix < (unsigned)w) {
      const float* src = x + ((int64_t)(b * h + iy) * w + ix) * ldx + c8 * 8;
      v0 = load4(src);
      v1 = load4(src + 4);
      if (prelu) {
        v0.x = v0.x >= 0.f ? v0.x : v0.x * a; v0.y = v0.y >= 0.f ? v0.y : v0.y * a;
        v0.z = v0.z >= 0.f ? v0.z : v0.z * a; v0.w = v0.w >= 0.f ? v0.w : v0.w * a;
        v1.x = v1.x >= 0.f ? v1.x : v1.x * a; v1.y = v1.y >= 0.f ? v1.y : v1.y * a;
        v1.z = v1.z >= 0.f ? v1.z : v1.z * a; v1.w = v1.w >= 0.f ? v1.w : v1.w * a;
      }
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v0.x, v0.y), p1 = __floats2bfloat162_rn(v0.z, v0.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v1.x, v1.y), p3 = __floats2bfloat162_rn(v1.z, v1.w);
    uint4 pk;
    pk.x = *reinterpret_cast<unsigned int*>(&p0); pk.y = *reinterpret_cast<unsigned int*>(&p1);
    pk.z = *reinterpret_cast<unsigned int*>(&p2); pk.w = *reinterpret_cast<unsigned int*>(&p3);
    *reinterpret_cast<uint4*>(col + i * 8) = pk;
  }
}

// bf16 operand of the implicit 3x3 conv: out[pixel, 0:Cp] = bf16(prelu(x[pixel, 0:C])), channels C..Cp-1 zero
__global__ void __launch_bounds__(256) cast_pad_act_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ out,
                                                           const float* __restrict__ prelu, int64_t total, int C, int Cp) {
  const int cg = Cp >> 3;
  const float a = prelu ? __ldg(prelu) : 1.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * 8;
    const int64_t row = i / cg;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (c < C) {
      v0 = load4(x + row * ldx + c);
      v1 = load4(x + row * ldx + c + 4);
      if (prelu) {
        v0.x = v0.x >= 0.f ? v0.x : v0.x * a; v0.y = v0.y >= 0.f ? v0.y : v0.y * a;
        v0.z = v0.z >= 0.f ? v0.z : v0.z * a; v0.w = v0.w >= 0.f ? v0.w : v0.w * a;
        v1.x = v1.x >= 0.f ? v1.x : v1.x * a; v1.y = v1.y >= 0.f ? v1.y : v1.y * a;
        v1.z = v1.z >= 0.f ? v1.z : v1.z * a; v1.w = v1.w >= 0.f ? v1.w : v1.w * a;
      }
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v0.x, v0.y), p1 = __floats2bfloat162_rn(v0.z, v0.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v1.x, v1.y), p3 = __floats2bfloat162_rn(v1.z, v1.w);
    uint4 pk;
    pk.x = *reinterpret_cast<unsigned int*>(&p0); pk.y = *reinterpret_cast<unsigned int*>(&p1);
    pk.z = *reinterpret_cast<unsigned int*>(&p2); pk.w = *reinterpret_cast<unsigned int*>(&p3);
    *reinterpret_cast<uint4*>(out + i * 8) = pk;
  }
}

__global__ void __launch_bounds__(256) sigmoid_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = sigmoidf_acc(x[i]);
}

}  // namespace dgtd
using namespace dgtd;

static inline int ew_grid(int64_t total, int threads = 256) {
  int64_t g = (total + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" {

int dgtd_conv_nhwc_affine_fwd(const float* x, const float* w, const float* scale, const float* shift,
                              const float* prelu, const float* residual, int ldr, float* out, int B, int h, int wd,
                              int Cin, int ldx, int oh, int ow, int Cout, int ldo, int ks, int stride, int off,
                              dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && w && out, "conv_nhwc_affine: null pointer");
  DGTD_CHECK_ARG(B > 0 && h > 0 && wd > 0 && oh > 0 && ow > 0 && Cin > 0 && Cout > 0 && ks > 0 && stride > 0,
                 "conv_nhwc_affine: bad shape");
  DGTD_CHECK_ARG(ldx >= Cin && ldo >= Cout && (!residual || ldr >= Cout), "conv_nhwc_affine: leading dims too small");
  DGTD_CHECK_ARG(Cin % 4 == 0 && ldx % 4 == 0 && Cout % 4 == 0 && ldo % 4 == 0 && ldr % 4 == 0,
                 "conv_nhwc_affine: channel counts / strides must be multiples of 4");
  const int64_t M64 = (int64_t)B * oh * ow;
  DGTD_CHECK_ARG(M64 < (1ll << 31), "conv_nhwc_affine: too many output pixels");
  const int M = (int)M64, K = ks * ks * Cin;
  Im2colLoader al{x, h, wd, ldx, Cin, oh, ow, ks, stride, off, M, K};
  RowMajorLoader bl{w, K, 0, Cout, K};
  EpiAffinePReLU ep{out, scale, shift, prelu, residual, ldo, ldr};
  launch_simt_gemm<true, true>(al, bl, ep, M, Cout, K, 1, (cudaStream_t)stream);
  DGTD_LAUNCH_CHECK("conv_nhwc_affine");
  return 0;
}

int dgtd_channel_sums_chunks(int hw) { return cdiv(hw, SUM_ROWS); }

int dgtd_channel_sums_fwd(const float* x, int ldx, float* partial, int B, int hw, int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && partial, "channel_sums: null pointer");
  DGTD_CHECK_ARG(B > 0 && hw > 0 && C > 0 && C % 4 == 0 && C <= 1024 && ldx >= C && ldx % 4 == 0,
                 "channel_sums: C must be a multiple of 4, <= 1024, ldx >= C");
  const int nch = cdiv(hw, SUM_ROWS);
  const int q = C / 4, groups = 256 / q;
  channel_sums_kernel<<<dim3(nch, B), 256, (size_t)groups * q * sizeof(float4), (cudaStream_t)stream>>>(x, ldx, hw, C,
                                                                                                      partial, nch);
  DGTD_LAUNCH_CHECK("channel_sums");
  return 0;
}

int dgtd_channel_gate_fwd(const float* partial, int nchunks, int hw, const float* w1, const float* w2, float* gate,
                          int B, int C, int Cr, int Co, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(partial && w1 && w2 && gate, "channel_gate: null pointer");
  DGTD_CHECK_ARG(B > 0 && C > 0 && Cr > 0 && Co > 0 && nchunks > 0 && hw > 0 && C + Cr <= 8192, "channel_gate: bad shape");
  channel_gate_kernel<<<B, 128, (size_t)(C + Cr) * sizeof(float), (cudaStream_t)stream>>>(partial, nchunks, 1.0f / hw, w1,
                                                                                        w2, gate, C, Cr, Co);
  DGTD_LAUNCH_CHECK("channel_gate");
  return 0;
}

int dgtd_gated_sum_fwd(const float* a, int lda, const float* ga, const float* sa, const float* b, int ldb,
                       const float* gb, const float* sb, float* out, int ldo, int B, int hw, int C,
                       dgtd_stream_t stream) {
  DGTD_CHECK_ARG(a && out, "gated_sum: null pointer");
  DGTD_CHECK_ARG(B > 0 && hw > 0 && C > 0 && C % 4 == 0 && lda % 4 == 0 && ldo % 4 == 0 && (!b || ldb % 4 == 0) &&
                     lda >= C && ldo >= C && (!b || ldb >= C),
                 "gated_sum: channels / strides must be multiples of 4 and strides >= C");
  const int64_t total = (int64_t)B * hw * (C / 4);
  gated_sum_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(a, lda, ga, sa, b, ldb, gb, sb, out, ldo, total, hw, C);
  DGTD_LAUNCH_CHECK("gated_sum");
  return 0;
}

int dgtd_resize_nhwc_ld_fwd(const float* x, int ldx, float* out, int ldo, int B, int h, int w, int C, int oh, int ow,
                            int align_corners, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out, "resize_nhwc_ld: null pointer");
  DGTD_CHECK_ARG(B > 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 &&
                     ldx >= C && ldo >= C,
                 "resize_nhwc_ld: channels / strides must be multiples of 4 and strides >= C");
  const float sy = align_corners ? (oh > 1 ? (float)(h - 1) / (float)(oh - 1) : 0.f) : (float)h / (float)oh;
  const float sx = align_corners ? (ow > 1 ? (float)(w - 1) / (float)(ow - 1) : 0.f) : (float)w / (float)ow;
  const int64_t total = (int64_t)B * oh * ow * (C / 4);
  resize_nhwc_ld_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, h, w, C, oh, ow, sy, sx,
                                                                         align_corners, total);
  DGTD_LAUNCH_CHECK("resize_nhwc_ld");
  return 0;
}

int dgtd_copy_channels_fwd(const float* x, int ldx, float* out, int ldo, int64_t rows, int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out, "copy_channels: null pointer");
  DGTD_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ldx >= C && ldo >= C,
                 "copy_channels: channels / strides must be multiples of 4 and strides >= C");
  const int64_t total = rows * (C / 4);
  copy_channels_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, C, total);
  DGTD_LAUNCH_CHECK("copy_channels");
  return 0;
}

int dgtd_head1_fwd(const float* x, int ldx, const float* w, const float* bias, float* out, int64_t rows, int C,
                   int accumulate, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && w && out, "head1: null pointer");
  DGTD_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldx >= C, "head1: C, ldx multiples of 4, ldx >= C");
  const int64_t threads = rows * 8;
  head1_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, out, rows, C, accumulate);
  DGTD_LAUNCH_CHECK("head1");
  return 0;
}

int dgtd_im2col_act_fwd(const float* x, int ldx, void* col, const float* prelu, int B, int h, int w, int C, int ks,
                        int stride, int off, int oh, int ow, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && col && B > 0 && h > 0 && w > 0 && ks >= 1 && stride >= 1 && oh > 0 && ow > 0, "im2col_act: bad args");
  DGTD_CHECK_ARG(C >= 8 && C % 8 == 0 && ldx >= C && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(col) & 15) == 0,
                 "im2col_act: C multiple of 8, pitch multiple of 4, 16-byte aligned pointers");
  const int64_t total = (int64_t)B * oh * ow * ks * ks * (C / 8);
  im2col_act_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, (__nv_bfloat16*)col, prelu, total, h, w, C, ks,
                                                                      stride, off, oh, ow);
  DGTD_LAUNCH_CHECK("im2col_act");
  return 0;
}

int dgtd_cast_pad_act_fwd(const float* x, int ldx, void* out, const float* prelu, int64_t rows, int C, int Cp,
                          dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out && rows > 0, "cast_pad_act: bad args");
  DGTD_CHECK_ARG(C >= 8 && C % 8 == 0 && Cp >= C && Cp % 8 == 0 && ldx >= C && ldx % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "cast_pad_act: C, Cp multiples of 8, pitch multiple of 4, 16-byte aligned pointers");
  const int64_t total = rows * (Cp / 8);
  cast_pad_act_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, (__nv_bfloat16*)out, prelu, total, C, Cp);
  DGTD_LAUNCH_CHECK("cast_pad_act");
  return 0;
}

int dgtd_sigmoid_fwd(const float* x, float* out, int64_t n, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out && n > 0, "sigmoid: bad arguments");
  sigmoid_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, out, n);
  DGTD_LAUNCH_CHECK("sigmoid");
  return 0;
}

}  // extern "C"
