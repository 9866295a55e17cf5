// Training kernels of the Hitnet iterative decoder (SURVEY.md 8f-2 under autograd): what the inference kernels of
// hitnet_ops.cu lack for `cod.forward(mode='loss')` to train every parameter.
//
//   BasicConv2d (cod.py:355-368)   train-mode BatchNorm2d: per-channel batch statistics (biased variance for the
//                                  normalisation, unbiased for the running estimate), apply, and the backward
//                                  dx = gamma rstd (dy - mean(dy) - xhat mean(dy xhat))
//   CAB (cod.py:434-451)           shared-slope PReLU forward / backward (the ONE nn.PReLU() of cod.py:686)
//   CALayer / SAM (cod.py:413-506) backward of x * sigmoid(W2 relu(W1 mean(x))) [* sigmoid(v2 relu(V1 mean(x)))]:
//                                  per-image channel dots sum_p g x, the tiny two-layer gate MLPs, the broadcast of
//                                  d mean back onto the pixels
//   out_CFM / out_SAM (:710-711)   backward of the 1-channel 1x1 conv
//   nn.Upsample(align_corners=True) (:709,733,737)   adjoint of the two-tap bilinear resize (gather form)
//
// Every reduction runs in a fixed order (per-CTA partials in double, one finalising pass): no atomics, results are
// bit-stable run to run.  All maps NHWC fp32; `ld*` = pixel pitch (a channel slice of a wider tensor is allowed).
#include "common.cuh"

namespace dgtd {
namespace {

constexpr int ST_ROWS = 512;    // rows per CTA of a column-statistics pass
constexpr int DOT_ROWS = 256;   // rows per chunk of the per-image channel dots (== SUM_ROWS of hitnet_ops.cu)

// ---- column statistics: ws[cta][2][C] (double) -----------------------------------------------------------------
//   MODE 0: (sum a, sum a^2)                         batch statistics of BatchNorm
//   MODE 1: (sum a, sum a * (b - mean) * rstd)       d beta, d gamma  (a = dy, b = conv output)
//   MODE 2: (sum a * b_row, sum b_row)               weight / bias gradient of a 1-channel head (b: one value per row)
template <int MODE>
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b,
                                                        int ldb, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, int64_t M, int C,
                                                        double* __restrict__ ws) {
  __shared__ double red[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * ST_ROWS;
  const int64_t r1 = r0 + ST_ROWS < M ? r0 + ST_ROWS : M;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + tx;
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
      float mu = 0.f, rs = 0.f;
      if (MODE == 1) {
        mu = mean[c];
        rs = rstd[c];
      }
      for (int64_t r = r0 + ty; r < r1; r += 8) {
        const float av = a[r * lda + c];
        if (MODE == 0) {
          s1 += (double)av;
          s2 += (double)av * (double)av;
        } else if (MODE == 1) {
          s1 += (double)av;
          s2 += (double)av * (double)((b[r * ldb + c] - mu) * rs);
        } else {
          const float bv = b[r * ldb];
          s1 += (double)av * (double)bv;
          s2 += (double)bv;
        }
      }
    }
    red[0][ty][tx] = s1;
    red[1][ty][tx] = s2;
    __syncthreads();
    if (ty < 2 && c < C) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[ty][k][tx];
      ws[((int64_t)blockIdx.x * 2 + ty) * C + c] = s;
    }
    __syncthreads();
  }
}

// BatchNorm forward finalise: mean, rstd of the batch; running estimates updated like nn.BatchNorm2d (momentum form,
// unbiased variance).  run_mean == nullptr: track_running_stats off.
__global__ void bn_finalize_kernel(const double* __restrict__ ws, int nblk, int64_t M, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ run_mean,
                                   float* __restrict__ run_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < nblk; ++k) {
    s1 += ws[((int64_t)k * 2 + 0) * C + c];
    s2 += ws[((int64_t)k * 2 + 1) * C + c];
  }
  const double mu = s1 / (double)M;
  double var = s2 / (double)M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (run_mean) {
    const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
    run_mean[c] = (float)((1.0 - (double)momentum) * (double)run_mean[c] + (double)momentum * mu);
    run_var[c] = (float)((1.0 - (double)momentum) * (double)run_var[c] + (double)momentum * unb);
  }
}

// out1[c] = sum_k ws[k][0][c], out2[c] = sum_k ws[k][1][c] (c < n2)
__global__ void stats_finalize_kernel(const double* __restrict__ ws, int nblk, int C, float* __restrict__ out1,
                                      float* __restrict__ out2, int n2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < nblk; ++k) {
    s1 += ws[((int64_t)k * 2 + 0) * C + c];
    s2 += ws[((int64_t)k * 2 + 1) * C + c];
  }
  out1[c] = (float)s1;
  if (c < n2) out2[c] = (float)s2;
}

// out = (y - mean) * rstd * gamma + beta
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ out, int ldo,
                                                       int C, int64_t total_q) {
  const int q = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_q; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % q) * 4;
    const int64_t row = i / q;
    const float4 v = load4(y + row * ldy + c), mu = load4(mean + c), rs = load4(rstd + c), ga = load4(gamma + c),
                 be = load4(beta + c);
    store4(out + row * ldo + c, fmaf((v.x - mu.x) * rs.x, ga.x, be.x), fmaf((v.y - mu.y) * rs.y, ga.y, be.y),
           fmaf((v.z - mu.z) * rs.z, ga.z, be.z), fmaf((v.w - mu.w) * rs.w, ga.w, be.w));
  }
}

// batch statistics: dx = gamma rstd (dy - dbeta / M - xhat dgamma / M);  fixed statistics: dx = gamma rstd dy
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y,
                                                           int ldy, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                                                           float* __restrict__ dx, int lddx, int C, float inv_m,
                                                           int batch_stats, int64_t total_q) {
  const int q = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_q; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % q) * 4;
    const int64_t row = i / q;
    const float4 g = load4(dy + row * lddy + c), rs = load4(rstd + c), ga = load4(gamma + c);
    float4 o;
    if (batch_stats) {
      const float4 v = load4(y + row * ldy + c), mu = load4(mean + c), dg = load4(dgamma + c), db = load4(dbeta + c);
      o.x = ga.x * rs.x * (g.x - db.x * inv_m - (v.x - mu.x) * rs.x * dg.x * inv_m);
      o.y = ga.y * rs.y * (g.y - db.y * inv_m - (v.y - mu.y) * rs.y * dg.y * inv_m);
      o.z = ga.z * rs.z * (g.z - db.z * inv_m - (v.z - mu.z) * rs.z * dg.z * inv_m);
      o.w = ga.w * rs.w * (g.w - db.w * inv_m - (v.w - mu.w) * rs.w * dg.w * inv_m);
    } else {
      o.x = ga.x * rs.x * g.x; o.y = ga.y * rs.y * g.y; o.z = ga.z * rs.z * g.z; o.w = ga.w * rs.w * g.w;
    }
    store4(dx + row * lddx + c, o.x, o.y, o.z, o.w);
  }
}

// ---- PReLU with ONE slope -------------------------------------------------------------------------------------
__device__ __forceinline__ float prelu1(float v, float a) { return v >= 0.f ? v : v * a; }

__global__ void __launch_bounds__(256) prelu_fwd_kernel(const float* __restrict__ u, const float* __restrict__ slope,
                                                        float* __restrict__ v, int64_t n4) {
  const float a = __ldg(slope);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = load4(u + i * 4);
    store4(v + i * 4, prelu1(x.x, a), prelu1(x.y, a), prelu1(x.z, a), prelu1(x.w, a));
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// du = g * (u >= 0 ? 1 : a);  ws[cta] = sum over the CTA's elements of g * u * [u < 0]
__global__ void __launch_bounds__(256) prelu_bwd_kernel(const float* __restrict__ u, const float* __restrict__ g,
                                                        const float* __restrict__ slope, float* __restrict__ du, int64_t n4,
                                                        double* __restrict__ ws) {
  __shared__ double red[8];
  const float a = __ldg(slope);
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = load4(u + i * 4), gv = load4(g + i * 4);
    store4(du + i * 4, x.x >= 0.f ? gv.x : gv.x * a, x.y >= 0.f ? gv.y : gv.y * a, x.z >= 0.f ? gv.z : gv.z * a,
           x.w >= 0.f ? gv.w : gv.w * a);
    float t = 0.f;
    t += x.x < 0.f ? gv.x * x.x : 0.f;
    t += x.y < 0.f ? gv.y * x.y : 0.f;
    t += x.z < 0.f ? gv.z * x.z : 0.f;
    t += x.w < 0.f ? gv.w * x.w : 0.f;
    s += (double)t;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    ws[blockIdx.x] = t;
  }
}

__global__ void sum_partials_kernel(const double* __restrict__ ws, int n, float* __restrict__ out, int accumulate) {
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += 32) s += ws[k];
  s = warp_sum_d(s);
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + (float)s : (float)s;
}

// ---- per-image channel dots: partial[b][chunk][c] = sum over the chunk's rows of a * b --------------------------
__global__ void __launch_bounds__(256) channel_dot_kernel(const float* __restrict__ a, int lda, const float* __restrict__ bb,
                                                          int ldb, int hw, int C, float* __restrict__ partial, int nchunks) {
  extern __shared__ float4 red4[];   // [groups][C/4]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int q = C >> 2;
  const int groups = 256 / q;
  const int cq = threadIdx.x % q, g = threadIdx.x / q;
  if (g < groups) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int r0 = chunk * DOT_ROWS, r1 = min(hw, r0 + DOT_ROWS);
    const float* pa = a + (int64_t)b * hw * lda + cq * 4;
    const float* pb = bb + (int64_t)b * hw * ldb + cq * 4;
    for (int r = r0 + g; r < r1; r += groups) {
      const float4 v = load4(pa + (int64_t)r * lda), w = load4(pb + (int64_t)r * ldb);
      acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
      acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
    }
    red4[g * q + cq] = acc;
  }
  __syncthreads();
  if (threadIdx.x < q) {
    float4 s = red4[threadIdx.x];
    for (int gg = 1; gg < groups; ++gg) {
      const float4 v = red4[gg * q + threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(partial + ((int64_t)b * nchunks + chunk) * C + threadIdx.x * 4) = s;
  }
}

// ---- backward of the gates ---------------------------------------------------------------------------------------
// out = x * gc[c] * gs with gc = sigmoid(W2 relu(W1 m)), gs = sigmoid(v2 relu(V1 m)) (gs = 1 without V), m = mean(x).
// Given D[c] = sum_p g x (the channel dots): d gc = D gs, d gs = sum_c D gc; back through both MLPs to d m and the
// weight gradients.  One CTA per image writes d m and the image's share of the weight gradients (ws[b][dw1 | dw2 | dv1 |
// dv2]); gate_bwd_reduce_kernel sums the shares over the images in order.
__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ part, int nch, float inv_hw,
                                                       const float* __restrict__ dpart, int nchd,
                                                       const float* __restrict__ w1, const float* __restrict__ w2,
                                                       const float* __restrict__ v1, const float* __restrict__ v2,
                                                       float* __restrict__ dmean, float* __restrict__ ws, int C, int Cr,
                                                       int Cs) {
  extern __shared__ float sm[];
  float* m = sm;              // [C]
  float* D = m + C;           // [C]
  float* gate = D + C;        // [C]
  float* dz = gate + C;       // [C]
  float* h = dz + C;          // [Cr]
  float* dh = h + Cr;         // [Cr]
  float* hs = dh + Cr;        // [Cs]
  float* dhs = hs + Cs;       // [Cs]
  float* sc = dhs + Cs;       // [2]: gs, dzs
  const int tid = threadIdx.x, nt = blockDim.x;
  const int b = blockIdx.x;
  const int nW = 2 * C * Cr + Cs * C + Cs;
  float* pw1 = ws + (int64_t)b * nW;
  float* pw2 = pw1 + C * Cr;
  float* pv1 = pw2 + C * Cr;
  float* pv2 = pv1 + Cs * C;
  for (int c = tid; c < C; c += nt) {
    float s = 0.f, d = 0.f;
    for (int k = 0; k < nch; ++k) s += part[((int64_t)b * nch + k) * C + c];
    for (int k = 0; k < nchd; ++k) d += dpart[((int64_t)b * nchd + k) * C + c];
    m[c] = s * inv_hw;
    D[c] = d;
  }
  __syncthreads();
  for (int j = tid; j < Cr + Cs; j += nt) {
    const float* w = j < Cr ? w1 + (int64_t)j * C : v1 + (int64_t)(j - Cr) * C;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(w[c], m[c], s);
    if (j < Cr) h[j] = fmaxf(s, 0.f);
    else hs[j - Cr] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int o = tid; o < C; o += nt) {
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(w2[(int64_t)o * Cr + j], h[j], s);
    gate[o] = sigmoidf_acc(s);
  }
  if (tid == 0) {
    float gs = 1.f;
    if (Cs > 0) {
      float s = 0.f;
      for (int j = 0; j < Cs; ++j) s = fmaf(v2[j], hs[j], s);
      gs = sigmoidf_acc(s);
    }
    sc[0] = gs;
  }
  __syncthreads();
  for (int o = tid; o < C; o += nt) dz[o] = D[o] * sc[0] * gate[o] * (1.f - gate[o]);
  if (tid == 32 && Cs > 0) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(D[c], gate[c], s);
    sc[1] = s * sc[0] * (1.f - sc[0]);
  }
  __syncthreads();
  for (int j = tid; j < Cr + Cs; j += nt) {
    if (j < Cr) {
      float s = 0.f;
      for (int o = 0; o < C; ++o) s = fmaf(w2[(int64_t)o * Cr + j], dz[o], s);
      dh[j] = h[j] > 0.f ? s : 0.f;
    } else {
      const int k = j - Cr;
      dhs[k] = hs[k] > 0.f ? v2[k] * sc[1] : 0.f;
    }
  }
  __syncthreads();
  for (int e = tid; e < C * Cr; e += nt) {       // dW2[o][j] = dz[o] h[j];  dW1[j][c] = dh[j] m[c]
    const int o = e / Cr, j = e - o * Cr;
    pw2[e] = dz[o] * h[j];
    const int jj = e / C, c = e - jj * C;
    pw1[e] = dh[jj] * m[c];
  }
  for (int e = tid; e < Cs * C; e += nt) {
    const int j = e / C, c = e - j * C;
    pv1[e] = dhs[j] * m[c];
  }
  for (int j = tid; j < Cs; j += nt) pv2[j] = sc[1] * hs[j];
  for (int c = tid; c < C; c += nt) {
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(w1[(int64_t)j * C + c], dh[j], s);
    for (int j = 0; j < Cs; ++j) s = fmaf(v1[(int64_t)j * C + c], dhs[j], s);
    dmean[(int64_t)b * C + c] = s;
  }
}

// dw1 | dw2 | dv1 | dv2 (+)= sum over the images, in order
__global__ void __launch_bounds__(256) gate_bwd_reduce_kernel(const float* __restrict__ ws, int B, int C, int Cr, int Cs,
                                                              float* __restrict__ dw1, float* __restrict__ dw2,
                                                              float* __restrict__ dv1, float* __restrict__ dv2,
                                                              int accumulate) {
  const int nW = 2 * C * Cr + Cs * C + Cs;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nW) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += ws[(int64_t)b * nW + e];
  float* dst;
  if (e < C * Cr) dst = dw1 + e;
  else if (e < 2 * C * Cr) dst = dw2 + (e - C * Cr);
  else if (e < 2 * C * Cr + Cs * C) dst = dv1 + (e - 2 * C * Cr);
  else dst = dv2 + (e - 2 * C * Cr - Cs * C);
  *dst = accumulate ? *dst + s : s;
}

// out = g * gate[b,c] * scal[b] + dmean[b,c] * inv_hw
__global__ void __launch_bounds__(256) gated_bwd_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ gate,
                                                        const float* __restrict__ scal, const float* __restrict__ dmean,
                                                        float inv_hw, float* __restrict__ out, int ldo, int64_t total_q,
                                                        int hw, int C) {
  const int q = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_q; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % q) * 4;
    const int64_t row = i / q;
    const int b = (int)(row / hw);
    float4 v = load4(g + row * ldg + c);
    const float4 ga = load4(gate + (int64_t)b * C + c);
    const float s = scal ? scal[b] : 1.f;
    v.x *= ga.x * s; v.y *= ga.y * s; v.z *= ga.z * s; v.w *= ga.w * s;
    if (dmean) {
      const float4 d = load4(dmean + (int64_t)b * C + c);
      v.x = fmaf(d.x, inv_hw, v.x); v.y = fmaf(d.y, inv_hw, v.y);
      v.z = fmaf(d.z, inv_hw, v.z); v.w = fmaf(d.w, inv_hw, v.w);
    }
    store4(out + row * ldo + c, v.x, v.y, v.z, v.w);
  }
}

// dx[row, c] = g[row] * w[c]
__global__ void __launch_bounds__(256) head1_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                          float* __restrict__ dx, int lddx, int C, int64_t total_q) {
  const int q = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_q; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % q) * 4;
    const int64_t row = i / q;
    const float gv = g[row];
    const float4 ww = load4(w + c);
    store4(dx + row * lddx + c, gv * ww.x, gv * ww.y, gv * ww.z, gv * ww.w);
  }
}

// ---- adjoint of the two-tap bilinear resize, both conventions (gather form, 4 channels per thread) --------------
__device__ __forceinline__ void tap2(int d, int n_in, float scale, bool align, int& i0, int& i1, float& f) {
  float src = align ? d * scale : fmaxf((d + 0.5f) * scale - 0.5f, 0.f);   // same expression as the forward kernel
  i0 = min((int)src, n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  f = src - (float)i0;
}

__device__ __forceinline__ void out_window(int i, int n_in, int n_out, float scale, bool align, int& lo, int& hi) {
  if (scale <= 0.f) {
    lo = 0;
    hi = n_out - 1;
    return;
  }
  const float shift = align ? 0.f : 0.5f;
  lo = (int)floorf((i - 1 + shift) / scale - shift) - 1;
  hi = (int)ceilf((i + 1 + shift) / scale - shift) + 1;
  if (i == 0) lo = 0;
  if (i == n_in - 1) hi = n_out - 1;
  lo = max(lo, 0);
  hi = min(hi, n_out - 1);
}

__global__ void __launch_bounds__(256) resize_ld_bwd_kernel(const float* __restrict__ g, int ldg, float* __restrict__ dx,
                                                            int lddx, int h, int w, int C, int oh, int ow, float sy, float sx,
                                                            int align, int64_t total_q) {
  const int q = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_q; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = (int)(i % q);
    int64_t t = i / q;
    const int ix = (int)(t % w);
    t /= w;
    const int iy = (int)(t % h);
    const int b = (int)(t / h);
    int oy_lo, oy_hi, ox_lo, ox_hi;
    out_window(iy, h, oh, sy, align, oy_lo, oy_hi);
    out_window(ix, w, ow, sx, align, ox_lo, ox_hi);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1;
      float ly;
      tap2(oy, h, sy, align, y0, y1, ly);
      const float wy = (y0 == iy ? 1.f - ly : 0.f) + (y1 == iy ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1;
        float lx;
        tap2(ox, w, sx, align, x0, x1, lx);
        const float wx = (x0 == ix ? 1.f - lx : 0.f) + (x1 == ix ? lx : 0.f);
        if (wx == 0.f) continue;
        const float4 gv = load4(g + (((int64_t)b * oh + oy) * ow + ox) * ldg + cq * 4);
        const float ww = wy * wx;
        acc.x = fmaf(ww, gv.x, acc.x); acc.y = fmaf(ww, gv.y, acc.y);
        acc.z = fmaf(ww, gv.z, acc.z); acc.w = fmaf(ww, gv.w, acc.w);
      }
    }
    store4(dx + (((int64_t)b * h + iy) * w + ix) * lddx + cq * 4, acc.x, acc.y, acc.z, acc.w);
  }
}

inline int ew_grid(int64_t total, int threads = 256) {
  int64_t g = (total + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

inline int stat_blocks(int64_t M) { return cdiv(M, ST_ROWS); }
inline int prelu_blocks(int64_t n4) {
  const int64_t g = (n4 + 255) / 256;
  return (int)(g < 592 ? (g > 0 ? g : 1) : 592);
}

}  // namespace
}  // namespace dgtd
using namespace dgtd;

extern "C" {

int64_t dgtd_col_stats_ws_bytes(int64_t M, int C) { return (int64_t)stat_blocks(M) * 2 * C * (int64_t)sizeof(double); }

int dgtd_bn_train_fwd(const float* y, int ldy, const float* gamma, const float* beta, float* run_mean, float* run_var,
                      float momentum, float eps, float* out, int ldo, float* mean, float* rstd, void* ws, int64_t M,
                      int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(y && gamma && beta && out && mean && rstd && ws, "bn_train_fwd: null pointer");
  DGTD_CHECK_ARG((run_mean == nullptr) == (run_var == nullptr), "bn_train_fwd: running mean / var go together");
  DGTD_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0 && ldy % 4 == 0 && ldo % 4 == 0 && ldy >= C && ldo >= C,
                 "bn_train_fwd: C / strides must be multiples of 4 and strides >= C");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = stat_blocks(M);
  col_stats_kernel<0><<<nblk, 256, 0, st>>>(y, ldy, nullptr, 0, nullptr, nullptr, M, C, (double*)ws);
  DGTD_LAUNCH_CHECK("bn_train_fwd.stats");
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>((const double*)ws, nblk, M, C, eps, momentum, mean, rstd, run_mean,
                                                  run_var);
  DGTD_LAUNCH_CHECK("bn_train_fwd.finalize");
  const int64_t total = M * (C / 4);
  bn_apply_kernel<<<ew_grid(total), 256, 0, st>>>(y, ldy, mean, rstd, gamma, beta, out, ldo, C, total);
  DGTD_LAUNCH_CHECK("bn_train_fwd.apply");
  return 0;
}

int dgtd_bn_apply_fwd(const float* y, int ldy, const float* mean, const float* rstd, const float* gamma,
                      const float* beta, float* out, int ldo, int64_t M, int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(y && mean && rstd && gamma && beta && out, "bn_apply: null pointer");
  DGTD_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0 && ldy % 4 == 0 && ldo % 4 == 0 && ldy >= C && ldo >= C,
                 "bn_apply: C / strides must be multiples of 4 and strides >= C");
  const int64_t total = M * (C / 4);
  bn_apply_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(y, ldy, mean, rstd, gamma, beta, out, ldo, C, total);
  DGTD_LAUNCH_CHECK("bn_apply");
  return 0;
}

int dgtd_bn_train_bwd(const float* dy, int lddy, const float* y, int ldy, const float* mean, const float* rstd,
                      const float* gamma, float* dx, int lddx, float* dgamma, float* dbeta, void* ws, int batch_stats,
                      int64_t M, int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(dy && y && mean && rstd && gamma && dx && dgamma && dbeta && ws, "bn_train_bwd: null pointer");
  DGTD_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0 && ldy % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && ldy >= C &&
                     lddy >= C && lddx >= C,
                 "bn_train_bwd: C / strides must be multiples of 4 and strides >= C");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = stat_blocks(M);
  col_stats_kernel<1><<<nblk, 256, 0, st>>>(dy, lddy, y, ldy, mean, rstd, M, C, (double*)ws);
  DGTD_LAUNCH_CHECK("bn_train_bwd.stats");
  stats_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>((const double*)ws, nblk, C, dbeta, dgamma, C);
  DGTD_LAUNCH_CHECK("bn_train_bwd.finalize");
  const int64_t total = M * (C / 4);
  bn_bwd_apply_kernel<<<ew_grid(total), 256, 0, st>>>(dy, lddy, y, ldy, mean, rstd, gamma, dgamma, dbeta, dx, lddx, C,
                                                     1.0f / (float)M, batch_stats, total);
  DGTD_LAUNCH_CHECK("bn_train_bwd.apply");
  return 0;
}

int dgtd_prelu_fwd(const float* u, const float* slope, float* v, int64_t n, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(u && slope && v && n > 0 && n % 4 == 0, "prelu_fwd: null pointer or n not a multiple of 4");
  prelu_fwd_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(u, slope, v, n / 4);
  DGTD_LAUNCH_CHECK("prelu_fwd");
  return 0;
}

int64_t dgtd_prelu_bwd_ws_bytes(int64_t n) { return (int64_t)prelu_blocks(n / 4) * (int64_t)sizeof(double); }

int dgtd_prelu_bwd(const float* u, const float* g, const float* slope, float* du, float* dslope, void* ws, int64_t n,
                   dgtd_stream_t stream) {
  DGTD_CHECK_ARG(u && g && slope && du && dslope && ws && n > 0 && n % 4 == 0,
                 "prelu_bwd: null pointer or n not a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = prelu_blocks(n / 4);
  prelu_bwd_kernel<<<nblk, 256, 0, st>>>(u, g, slope, du, n / 4, (double*)ws);
  DGTD_LAUNCH_CHECK("prelu_bwd");
  sum_partials_kernel<<<1, 32, 0, st>>>((const double*)ws, nblk, dslope, 0);
  DGTD_LAUNCH_CHECK("prelu_bwd.finalize");
  return 0;
}

int dgtd_channel_dot_fwd(const float* a, int lda, const float* b, int ldb, float* partial, int B, int hw, int C,
                         dgtd_stream_t stream) {
  DGTD_CHECK_ARG(a && b && partial, "channel_dot: null pointer");
  DGTD_CHECK_ARG(B > 0 && hw > 0 && C > 0 && C % 4 == 0 && C <= 1024 && lda >= C && lda % 4 == 0 && ldb >= C &&
                     ldb % 4 == 0,
                 "channel_dot: C must be a multiple of 4, <= 1024, strides >= C");
  const int nch = cdiv(hw, DOT_ROWS);
  DGTD_CHECK_ARG(nch == dgtd_channel_sums_chunks(hw), "channel_dot: chunking differs from channel_sums");
  const int q = C / 4, groups = 256 / q;
  channel_dot_kernel<<<dim3(nch, B), 256, (size_t)groups * q * sizeof(float4), (cudaStream_t)stream>>>(a, lda, b, ldb, hw,
                                                                                                     C, partial, nch);
  DGTD_LAUNCH_CHECK("channel_dot");
  return 0;
}

int64_t dgtd_gate_bwd_ws_floats(int B, int C, int Cr, int Cs) { return (int64_t)B * (2 * C * Cr + Cs * C + Cs); }

int dgtd_gate_bwd(const float* part, int nch, int hw, const float* dpart, int nchd, const float* w1, const float* w2,
                  const float* v1, const float* v2, float* dmean, float* dw1, float* dw2, float* dv1, float* dv2,
                  float* ws, int accumulate, int B, int C, int Cr, int Cs, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(part && dpart && w1 && w2 && dmean && dw1 && dw2 && ws, "gate_bwd: null pointer");
  DGTD_CHECK_ARG((Cs == 0) || (v1 && v2 && dv1 && dv2), "gate_bwd: scalar gate needs v1, v2, dv1, dv2");
  DGTD_CHECK_ARG(B > 0 && C > 0 && Cr > 0 && Cs >= 0 && nch > 0 && nchd > 0 && hw > 0 && 4 * C + 2 * Cr + 2 * Cs + 2 <= 8192,
                 "gate_bwd: bad shape");
  const size_t smem = (size_t)(4 * C + 2 * Cr + 2 * Cs + 2) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  gate_bwd_kernel<<<B, 256, smem, st>>>(part, nch, 1.0f / hw, dpart, nchd, w1, w2, v1, v2, dmean, ws, C, Cr, Cs);
  DGTD_LAUNCH_CHECK("gate_bwd");
  const int nW = 2 * C * Cr + Cs * C + Cs;
  gate_bwd_reduce_kernel<<<cdiv(nW, 256), 256, 0, st>>>(ws, B, C, Cr, Cs, dw1, dw2, dv1, dv2, accumulate);
  DGTD_LAUNCH_CHECK("gate_bwd.reduce");
  return 0;
}

int dgtd_gated_bwd(const float* g, int ldg, const float* gate, const float* scal, const float* dmean, float* out, int ldo,
                   int B, int hw, int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(g && gate && out, "gated_bwd: null pointer");
  DGTD_CHECK_ARG(B > 0 && hw > 0 && C > 0 && C % 4 == 0 && ldg % 4 == 0 && ldo % 4 == 0 && ldg >= C && ldo >= C,
                 "gated_bwd: channels / strides must be multiples of 4 and strides >= C");
  const int64_t total = (int64_t)B * hw * (C / 4);
  gated_bwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(g, ldg, gate, scal, dmean, 1.0f / hw, out, ldo, total,
                                                                    hw, C);
  DGTD_LAUNCH_CHECK("gated_bwd");
  return 0;
}

int dgtd_head1_bwd(const float* g, const float* x, int ldx, const float* w, float* dx, int lddx, float* dw, float* db,
                   void* ws, int64_t rows, int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(g && x && w && dw && db && ws, "head1_bwd: null pointer");
  DGTD_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldx >= C && (!dx || (lddx % 4 == 0 && lddx >= C)),
                 "head1_bwd: C, strides multiples of 4, strides >= C");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = stat_blocks(rows);
  col_stats_kernel<2><<<nblk, 256, 0, st>>>(x, ldx, g, 1, nullptr, nullptr, rows, C, (double*)ws);
  DGTD_LAUNCH_CHECK("head1_bwd.stats");
  stats_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>((const double*)ws, nblk, C, dw, db, 1);
  DGTD_LAUNCH_CHECK("head1_bwd.finalize");
  if (dx) {
    const int64_t total = rows * (C / 4);
    head1_dgrad_kernel<<<ew_grid(total), 256, 0, st>>>(g, w, dx, lddx, C, total);
    DGTD_LAUNCH_CHECK("head1_bwd.dgrad");
  }
  return 0;
}

int dgtd_resize_nhwc_ld_bwd(const float* g, int ldg, float* dx, int lddx, int B, int h, int w, int C, int oh, int ow,
                            int align_corners, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(g && dx, "resize_nhwc_ld_bwd: null pointer");
  DGTD_CHECK_ARG(B > 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && C > 0 && C % 4 == 0 && ldg % 4 == 0 && lddx % 4 == 0 &&
                     ldg >= C && lddx >= C,
                 "resize_nhwc_ld_bwd: channels / strides must be multiples of 4 and strides >= C");
  const float sy = align_corners ? (oh > 1 ? (float)(h - 1) / (float)(oh - 1) : 0.f) : (float)h / (float)oh;
  const float sx = align_corners ? (ow > 1 ? (float)(w - 1) / (float)(ow - 1) : 0.f) : (float)w / (float)ow;
  const int64_t total = (int64_t)B * h * w * (C / 4);
  resize_ld_bwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(g, ldg, dx, lddx, h, w, C, oh, ow, sy, sx,
                                                                        align_corners, total);
  DGTD_LAUNCH_CHECK("resize_nhwc_ld_bwd");
  return 0;
}

}  // extern "C"
