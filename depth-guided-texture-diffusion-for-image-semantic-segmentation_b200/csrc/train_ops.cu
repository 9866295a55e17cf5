// Backward kernels of the trunk / decoders (training path, config/sod.yml).  Exact fp32 on the
// CUDA cores: gradients GEMMs reuse simt_gemm_kernel with transposed operand loaders
// (dgrad: A K-major, B N-major; wgrad: both M-major with split-K over the rows), reductions are
// fixed-order (no atomics) so gradients are bit-stable run to run.
#include "simt_gemm.cuh"

namespace dgtd {

int dwconv7_tma(const float* x, const float* wT, const float* dw_b, const float* add, float* y, int B, int h, int w,
                int C, cudaStream_t s);
int dwconv7_wgrad_tma(const float* x, const float* dy, float* part, int max_parts, int* parts, int B, int h, int w,
                      int C, cudaStream_t s);

__device__ __forceinline__ float gelu_erf_grad(float x) {
  // d/dx [0.5 x (1 + erf(x/sqrt2))] = Phi(x) + x phi(x)
  const float phi = 0.3989422804014327f * expf(-0.5f * x * x);
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * phi;
}

// ---- epilogues ------------------------------------------------------------------------------
// out = acc * gelu'(pre[m,n])   (pre nullable: plain store);  optional accumulate into out
struct EpiDgrad {
  float* out;
  const float* pre;   // nullable
  int64_t ld;
  int accumulate;
  __device__ __forceinline__ void operator()(int m, int n, int, float4 v) const {
    float* o = out + (int64_t)m * ld + n;
    if (pre) {
      float4 p = load4(pre + (int64_t)m * ld + n);
      v.x *= gelu_erf_grad(p.x); v.y *= gelu_erf_grad(p.y);
      v.z *= gelu_erf_grad(p.z); v.w *= gelu_erf_grad(p.w);
    }
    if (accumulate) {
      float4 a = load4(o);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    store4(o, v.x, v.y, v.z, v.w);
  }
};
struct EpiSplitStore {   // partial[z][m][n]
  float* out;
  int64_t ld, zs;
  __device__ __forceinline__ void operator()(int m, int n, int z, float4 v) const {
    store4(out + z * zs + (int64_t)m * ld + n, v.x, v.y, v.z, v.w);
  }
};

// sum over splits: out[i] = sum_z partial[z][i]
__global__ void sum_splits_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n, int S) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += part[(int64_t)z * n + i];
  out[i] = s;
}

// column sums with optional per-sample row scale: part[blk][n] then summed by sum_splits_kernel
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, const float* __restrict__ keep, int rows_per_sample,
                      float* __restrict__ part, int M, int N, int rows_per_block) {
  __shared__ float red[8][32];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + (threadIdx.x >> 5); r < r1; r += 8) {
      float v = x[(int64_t)r * N + n];
      if (keep) v *= keep[r / rows_per_sample];
      s += v;
    }
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && n < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    part[(int64_t)blockIdx.y * N + n] = t;
  }
}

// layer-scale finalize: dW2 = gamma_n G_nk, db2 = gamma s, dgamma_n = sum_k W2_nk G_nk + b2_n s_n
__global__ void __launch_bounds__(256)
layer_scale_finalize_kernel(const float* __restrict__ G, const float* __restrict__ s, const float* __restrict__ W2,
                            const float* __restrict__ b2, const float* __restrict__ gamma,
                            float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ dgamma, int K) {
  __shared__ float red[256];
  const int n = blockIdx.x;
  const float g = gamma ? gamma[n] : 1.f;
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) {
    const float gv = G[(int64_t)n * K + k];
    acc = fmaf(W2[(int64_t)n * K + k], gv, acc);
    dW2[(int64_t)n * K + k] = g * gv;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    db2[n] = g * s[n];
    if (dgamma) dgamma[n] = red[0] + b2[n] * s[n];
  }
}

__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 v = load4(x + i);
  store4(out + i, gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
}
// g *= (out > 0)   (ReLU backward on the saved post-activation)
__global__ void relu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ out, float* __restrict__ dx,
                                int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 a = load4(g + i), o = load4(out + i);
  store4(dx + i, o.x > 0.f ? a.x : 0.f, o.y > 0.f ? a.y : 0.f, o.z > 0.f ? a.z : 0.f, o.w > 0.f ? a.w : 0.f);
}

// ---- LayerNorm over the trailing channel dim, backward ------------------------------------------
// one warp per row; per-block partials of dw = sum g*xhat, db = sum g go to part[blk][2][C]
__global__ void __launch_bounds__(256)
ln_rows_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ w,
                   float* __restrict__ dy, float* __restrict__ part, int64_t rows, int C, float eps) {
  extern __shared__ float sm[];   // [8 warps][2][C]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* mydw = sm + (wid * 2) * C;
  float* mydb = mydw + C;
  for (int c = lane; c < C; c += 32) { mydw[c] = 0.f; mydb[c] = 0.f; }
  for (int64_t row = (int64_t)blockIdx.x * 8 + wid; row < rows; row += (int64_t)gridDim.x * 8) {
    const float* yp = y + row * C;
    const float* gp = g + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += yp[c];
    const float mean = warp_sum(s) / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { float d = yp[c] - mean; q = fmaf(d, d, q); }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
    float a1 = 0.f, a2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (yp[c] - mean) * rstd, gx = gp[c] * w[c];
      a1 += gx; a2 = fmaf(gx, xh, a2);
    }
    const float m1 = warp_sum(a1) / C, m2 = warp_sum(a2) / C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (yp[c] - mean) * rstd, gv = gp[c];
      dy[row * C + c] = rstd * (gv * w[c] - m1 - xh * m2);
      mydw[c] = fmaf(gv, xh, mydw[c]);
      mydb[c] += gv;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i * 2 * C + c];
    part[(int64_t)blockIdx.x * 2 * C + c] = t;
  }
}

// Register-resident form for C = 128 * VPL (every trunk width): a lane owns the 4 * VPL channels (j * 32 + lane) * 4 + e of
// every row its warp visits; the row's y and g are read ONCE (16-byte loads) and stay in registers through the three
// passes (statistics, projections, result), and the dw / db partial sums stay in registers across all rows of the warp
// -- the generic kernel above re-read y four times and paid two shared-memory read-modify-writes per element
// (r2 profile: 2.7 ms of the 33 ms training step at ~4x its HBM floor of 12 bytes per element).
template <int VPL>
__global__ void __launch_bounds__(256)
ln_rows_bwd_v_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ w,
                     float* __restrict__ dy, float* __restrict__ part, int64_t rows, float eps) {
  constexpr int C = 128 * VPL;
  extern __shared__ float sm[];   // [8 warps][2][C]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float4 wv[VPL], adw[VPL], adb[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    wv[j] = *reinterpret_cast<const float4*>(w + (j * 32 + lane) * 4);
    adw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    adb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = (int64_t)blockIdx.x * 8 + wid; row < rows; row += (int64_t)gridDim.x * 8) {
    const float* yp = y + row * C;
    const float* gp = g + row * C;
    float4 yv[VPL], gv[VPL];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      yv[j] = *reinterpret_cast<const float4*>(yp + (j * 32 + lane) * 4);
      gv[j] = *reinterpret_cast<const float4*>(gp + (j * 32 + lane) * 4);
      s += (yv[j].x + yv[j].y) + (yv[j].z + yv[j].w);
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      yv[j].x -= mean; yv[j].y -= mean; yv[j].z -= mean; yv[j].w -= mean;
      q += (yv[j].x * yv[j].x + yv[j].y * yv[j].y) + (yv[j].z * yv[j].z + yv[j].w * yv[j].w);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {   // yv <- xhat
      yv[j].x *= rstd; yv[j].y *= rstd; yv[j].z *= rstd; yv[j].w *= rstd;
      const float gx0 = gv[j].x * wv[j].x, gx1 = gv[j].y * wv[j].y, gx2 = gv[j].z * wv[j].z, gx3 = gv[j].w * wv[j].w;
      a1 += (gx0 + gx1) + (gx2 + gx3);
      a2 += (gx0 * yv[j].x + gx1 * yv[j].y) + (gx2 * yv[j].z + gx3 * yv[j].w);
    }
    const float m1 = warp_sum(a1) / C, m2 = warp_sum(a2) / C;
    float* dp = dy + row * C;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      float4 o;
      o.x = rstd * (gv[j].x * wv[j].x - m1 - yv[j].x * m2);
      o.y = rstd * (gv[j].y * wv[j].y - m1 - yv[j].y * m2);
      o.z = rstd * (gv[j].z * wv[j].z - m1 - yv[j].z * m2);
      o.w = rstd * (gv[j].w * wv[j].w - m1 - yv[j].w * m2);
      *reinterpret_cast<float4*>(dp + (j * 32 + lane) * 4) = o;
      adw[j].x = fmaf(gv[j].x, yv[j].x, adw[j].x); adw[j].y = fmaf(gv[j].y, yv[j].y, adw[j].y);
      adw[j].z = fmaf(gv[j].z, yv[j].z, adw[j].z); adw[j].w = fmaf(gv[j].w, yv[j].w, adw[j].w);
      adb[j].x += gv[j].x; adb[j].y += gv[j].y; adb[j].z += gv[j].z; adb[j].w += gv[j].w;
    }
  }
  float* mydw = sm + (wid * 2) * C;
  float* mydb = mydw + C;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    *reinterpret_cast<float4*>(mydw + (j * 32 + lane) * 4) = adw[j];
    *reinterpret_cast<float4*>(mydb + (j * 32 + lane) * 4) = adb[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i * 2 * C + c];
    part[(int64_t)blockIdx.x * 2 * C + c] = t;
  }
}

// ---- depthwise 7x7: plain forward (optionally + add) and weight gradient -------------------------
// thread = channel, CTA = (image, 8-row band, 128-channel group), sliding 7-row window
__global__ void __launch_bounds__(128)
dwconv7_plain_kernel(const float* __restrict__ x, const float* __restrict__ wT, const float* __restrict__ bias,
                     const float* __restrict__ add, float* __restrict__ y, int h, int w, int C, int flip) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int band = blockIdx.y, b = blockIdx.z;
  float wr[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) wr[k] = wT[(int64_t)(flip ? 48 - k : k) * C + c];
  const float bc = bias ? bias[c] : 0.f;
  const float* xb = x + (int64_t)b * h * w * C + c;
  for (int oy = band * 8; oy < min(h, band * 8 + 8); ++oy)
    for (int ox = 0; ox < w; ++ox) {
      float acc = bc;
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const int iy = oy + ky - 3;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int ix = ox + kx - 3;
          if (ix >= 0 && ix < w) acc = fmaf(wr[ky * 7 + kx], xb[((int64_t)iy * w + ix) * C], acc);
        }
      }
      const int64_t o = (((int64_t)b * h + oy) * w + ox) * C + c;
      y[o] = add ? acc + add[o] : acc;
    }
}

__global__ void __launch_bounds__(128)
dwconv7_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int h,
                     int w, int C) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int band = blockIdx.y, b = blockIdx.z;
  float acc[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) acc[k] = 0.f;
  float gsum = 0.f;
  const float* xb = x + (int64_t)b * h * w * C + c;
  const float* gb = dy + (int64_t)b * h * w * C + c;
  for (int oy = band * 8; oy < min(h, band * 8 + 8); ++oy)
    for (int ox = 0; ox < w; ++ox) {
      const float g = gb[((int64_t)oy * w + ox) * C];
      gsum += g;
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const int iy = oy + ky - 3;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int ix = ox + kx - 3;
          if (ix >= 0 && ix < w) acc[ky * 7 + kx] = fmaf(g, xb[((int64_t)iy * w + ix) * C], acc[ky * 7 + kx]);
        }
      }
    }
  // part[(b*bands + band)][50][C]: 49 taps + bias
  float* p = part + ((int64_t)(b * gridDim.y + band) * 50) * C + c;
#pragma unroll
  for (int k = 0; k < 49; ++k) p[(int64_t)k * C] = acc[k];
  p[(int64_t)49 * C] = gsum;
}

// ---- stem patch gather (training forward; bf16 operand of the tcgen05 stem GEMM) and its adjoint ----------
// patches[(b,oy,ox)][ci*16+ky*4+kx] = image[b,ci,4oy+ky,4ox+kx] + up(grid)[...]; thread = 4 consecutive kx
template <typename OT>
__global__ void stem_patchify_kernel(const float* __restrict__ image, const float* __restrict__ grid, int G,
                                     OT* __restrict__ patches, int H, int W, int oh, int ow, int64_t total4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int k4 = (int)(i % 12);            // (ci, ky)
  int64_t t = i / 12;
  const int ox = (int)(t % ow); t /= ow;
  const int oy = (int)(t % oh);
  const int b = (int)(t / oh);
  const int ci = k4 >> 2, ky = k4 & 3;
  const int iy = oy * 4 + ky, ix0 = ox * 4;
  const float* ip = image + (((int64_t)b * 3 + ci) * H + iy) * W + ix0;
  float v[4];
  if ((W & 3) == 0) {
    const float4 q = *reinterpret_cast<const float4*>(ip);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) v[kx] = ip[kx];
  }
  if (grid) {
    int y0, y1;
    float ly;
    bilinear_src(iy, (float)G / H, G, y0, y1, ly);
    const float* g = grid + ((int64_t)b * 3 + ci) * G * G;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      int x0, x1;
      float lx;
      bilinear_src(ix0 + kx, (float)G / W, G, x0, x1, lx);
      v[kx] += (1.f - ly) * ((1.f - lx) * g[y0 * G + x0] + lx * g[y0 * G + x1]) +
               ly * ((1.f - lx) * g[y1 * G + x0] + lx * g[y1 * G + x1]);
    }
  }
  store4(patches + i * 4, v[0], v[1], v[2], v[3]);
}
// dimg[b,ci,iy,ix] = dpatches[(b,iy/4,ix/4)][ci*16+(iy%4)*4+ix%4]  (0 outside the covered area)
__global__ void stem_unpatchify_kernel(const float* __restrict__ dp, float* __restrict__ dimg, int H, int W, int oh,
                                       int ow, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ix = (int)(i % W);
  int64_t t = i / W;
  const int iy = (int)(t % H); t /= H;
  const int ci = (int)(t % 3);
  const int b = (int)(t / 3);
  const int oy = iy >> 2, ox = ix >> 2;
  float v = 0.f;
  if (oy < oh && ox < ow) v = dp[(((int64_t)b * oh + oy) * ow + ox) * 48 + ci * 16 + (iy & 3) * 4 + (ix & 3)];
  dimg[i] = v;
}

// 2x2 patch scatter back: dx[b,y,x,c] = dpatch[(b,y/2,x/2)][((y&1)*2+(x&1))*C + c]
__global__ void unpatchify2_kernel(const float* __restrict__ dp, float* __restrict__ dx, int h, int w, int C,
                                   int64_t total) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= total) return;
  const int c = (int)(i % C);
  int64_t t = i / C;
  const int x = (int)(t % w); t /= w;
  const int y = (int)(t % h);
  const int b = (int)(t / h);
  const int h2 = h >> 1, w2 = w >> 1;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((y >> 1) < h2 && (x >> 1) < w2)
    v = load4(dp + (((int64_t)b * h2 + (y >> 1)) * w2 + (x >> 1)) * (4 * (int64_t)C) + ((y & 1) * 2 + (x & 1)) * C + c);
  store4(dx + i, v.x, v.y, v.z, v.w);
}
// forward counterpart without LayerNorm (training path keeps LN separate)
__global__ void patchify2_kernel(const float* __restrict__ x, float* __restrict__ out, int h, int w, int C,
                                 int64_t total) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= total) return;   // total = B*h2*w2*4C
  const int64_t C4 = 4 * (int64_t)C;
  const int k = (int)(i % C4);
  int64_t t = i / C4;
  const int h2 = h >> 1, w2 = w >> 1;
  const int x2 = (int)(t % w2); t /= w2;
  const int y2 = (int)(t % h2);
  const int b = (int)(t / h2);
  const int q = k / C, c = k - q * C;
  float4 v = load4(x + (((int64_t)b * h + 2 * y2 + (q >> 1)) * w + 2 * x2 + (q & 1)) * C + c);
  store4(out + i, v.x, v.y, v.z, v.w);
}

// ---- generic NHWC bilinear resize adjoint (gather form), 4 channels per thread ---------------------
__global__ void resize_nhwc_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, int h, int w, int C,
                                       int oh, int ow, int accumulate, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = C >> 2;
  const int cq = (int)(i % c4);
  int64_t t = i / c4;
  const int ix = (int)(t % w); t /= w;
  const int iy = (int)(t % h);
  const int b = (int)(t / h);
  const float sy = (float)h / oh, sx = (float)w / ow;
  int oy_lo = (int)floorf((iy - 1 + 0.5f) / sy - 0.5f) - 1, oy_hi = (int)ceilf((iy + 1 + 0.5f) / sy - 0.5f) + 1;
  int ox_lo = (int)floorf((ix - 1 + 0.5f) / sx - 0.5f) - 1, ox_hi = (int)ceilf((ix + 1 + 0.5f) / sx - 0.5f) + 1;
  if (iy == 0) oy_lo = 0;
  if (ix == 0) ox_lo = 0;
  if (iy == h - 1) oy_hi = oh - 1;
  if (ix == w - 1) ox_hi = ow - 1;
  oy_lo = max(oy_lo, 0); ox_lo = max(ox_lo, 0);
  oy_hi = min(oy_hi, oh - 1); ox_hi = min(ox_hi, ow - 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    int y0, y1;
    float ly;
    bilinear_src(oy, sy, h, y0, y1, ly);
    const float wy = (y0 == iy ? 1.f - ly : 0.f) + (y1 == iy ? ly : 0.f);
    if (wy == 0.f) continue;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      int x0, x1;
      float lx;
      bilinear_src(ox, sx, w, x0, x1, lx);
      const float wx = (x0 == ix ? 1.f - lx : 0.f) + (x1 == ix ? lx : 0.f);
      if (wx == 0.f) continue;
      const float4 gv = load4(g + (((int64_t)b * oh + oy) * ow + ox) * C + cq * 4);
      const float ww = wy * wx;
      acc.x = fmaf(ww, gv.x, acc.x); acc.y = fmaf(ww, gv.y, acc.y);
      acc.z = fmaf(ww, gv.z, acc.z); acc.w = fmaf(ww, gv.w, acc.w);
    }
  }
  float* o = dx + (((int64_t)b * h + iy) * w + ix) * C + cq * 4;
  if (accumulate) {
    float4 a = load4(o);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  store4(o, acc.x, acc.y, acc.z, acc.w);
}

// Split-K factor of the CUDA-core weight-gradient GEMM (reduction over the M rows).  Round 1 used M / 4096 whatever the
// output size, which left the thin gradients of the head / fusion / stem convs (N x K = 24 x 128 ... 128 x 48: ONE
// 128 x 128 output tile) on 36 CTAs streaming 4096 rows each: 2.7 ms of the 33 ms training step at 2 TFLOP/s.  Now the
// splits fill the machine (~2 CTAs per SM in total) as long as a split keeps >= 256 rows.
static int pick_splits(int M, int N, int K) {
  const int tiles = cdiv(N, 128) * cdiv(K, 128);
  int s = cdiv(2 * 148, tiles);
  const int by_rows = M / 256;
  if (s > by_rows) s = by_rows;
  if (s < M / 4096) s = M / 4096;
  if (s > 512) s = 512;
  if (s < 1) s = 1;
  return s;
}

}  // namespace dgtd

using namespace dgtd;

extern "C" {

// dx[M,K] (+)= (keep . gamma . g)[M,N] . W[N,K]  (. gelu'(pre[M,K]) when pre != NULL)
int dgtd_linear_dgrad(const float* g, const float* w, float* dx, const float* pre, const float* keep,
                      const float* gamma, int rows_per_sample, int M, int N, int K, int accumulate,
                      dgtd_stream_t stream) {
  DGTD_CHECK_ARG(g && w && dx && M > 0 && N > 0 && K > 0 && N % 4 == 0 && K % 4 == 0, "linear_dgrad: bad args");
  SplitRowLoader al{g, N, M, N, M, keep, gamma, rows_per_sample > 0 ? rows_per_sample : 1};
  RowMajorLoader bl{w, K, 0, N, K};            // (k' = n, n' = k) -> W[n][k]
  EpiDgrad ep{dx, pre, K, accumulate};
  launch_simt_gemm<true, false>(al, bl, ep, M, K, N, 1, (cudaStream_t)stream);
  DGTD_LAUNCH_CHECK("linear_dgrad");
  return 0;
}

// dw[N,K] = (keep . g)^T[N,M] . a[M,K]; ws: pick_splits(M)*N*K floats.  a may be an im2col view
// (conv weight gradient): pass conv geometry with ks > 0, then `a` is the NHWC input.
int dgtd_linear_wgrad(const float* g, const float* a, float* dw, float* ws, const float* keep, int rows_per_sample,
                      int M, int N, int K, int ks, int h, int wd, int Cin, int ldx, int oh, int ow, int stride,
                      int off, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(g && a && dw && ws && M > 0 && N > 0 && K > 0 && N % 4 == 0 && K % 4 == 0, "linear_wgrad: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int S = pick_splits(M, N, K);
  const int zrows = cdiv(M, S);
  SplitRowLoader al{g, N, M, N, zrows, keep, nullptr, rows_per_sample > 0 ? rows_per_sample : 1};
  EpiSplitStore ep{ws, K, (int64_t)N * K};
  if (ks > 0) {
    Im2colLoader bl{a, h, wd, ldx, Cin, oh, ow, ks, stride, off, M, K, zrows};
    launch_simt_gemm<false, false>(al, bl, ep, N, K, zrows, S, s);
  } else {
    SplitRowLoader bl{a, K, M, K, zrows, nullptr, nullptr, 1};
    launch_simt_gemm<false, false>(al, bl, ep, N, K, zrows, S, s);
  }
  DGTD_LAUNCH_CHECK("linear_wgrad");
  const int64_t n = (int64_t)N * K;
  sum_splits_kernel<<<cdiv(n, 256), 256, 0, s>>>(ws, dw, n, S);
  DGTD_LAUNCH_CHECK("linear_wgrad.reduce");
  return 0;
}
int dgtd_linear_wgrad_ws_floats(int M, int N, int K) { return pick_splits(M, N, K) * N * K; }

// out[N] = sum_m keep[m/rows] x[m,n];  ws: cdiv(M,1024)*N floats
int dgtd_colsum(const float* x, const float* keep, int rows_per_sample, float* ws, float* out, int M, int N,
                dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && ws && out && M > 0 && N > 0, "colsum: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = cdiv(M, 1024);
  colsum_partial_kernel<<<dim3(cdiv(N, 32), nb), 256, 0, s>>>(x, keep, rows_per_sample > 0 ? rows_per_sample : 1, ws, M,
                                                               N, 1024);
  DGTD_LAUNCH_CHECK("colsum");
  sum_splits_kernel<<<cdiv(N, 256), 256, 0, s>>>(ws, out, N, nb);
  DGTD_LAUNCH_CHECK("colsum.reduce");
  return 0;
}

int dgtd_layer_scale_finalize(const float* G, const float* s, const float* W2, const float* b2, const float* gamma,
                              float* dW2, float* db2, float* dgamma, int N, int K, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(G && s && W2 && b2 && dW2 && db2 && N > 0 && K > 0, "layer_scale_finalize: bad args");
  layer_scale_finalize_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(G, s, W2, b2, gamma, dW2, db2, dgamma, K);
  DGTD_LAUNCH_CHECK("layer_scale_finalize");
  return 0;
}

int dgtd_gelu_fwd(const float* x, float* out, int64_t n, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out && n > 0 && n % 4 == 0, "gelu: n must be a positive multiple of 4");
  gelu_fwd_kernel<<<cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, out, n);
  DGTD_LAUNCH_CHECK("gelu_fwd");
  return 0;
}
int dgtd_relu_bwd(const float* g, const float* out, float* dx, int64_t n, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(g && out && dx && n > 0 && n % 4 == 0, "relu_bwd: n must be a positive multiple of 4");
  relu_bwd_kernel<<<cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(g, out, dx, n);
  DGTD_LAUNCH_CHECK("relu_bwd");
  return 0;
}

// LayerNorm(rows of C) backward.  ws: blocks*2*C floats with blocks = min(cdiv(rows,8), 592)
int dgtd_ln_rows_bwd(const float* g, const float* y, const float* w, float* dy, float* ws, float* dw, float* db,
                     int64_t rows, int C, float eps, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(g && y && w && dy && ws && dw && db && rows > 0 && C > 0, "ln_rows_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = cdiv(rows, 8);
  if (blocks > 592) blocks = 592;
  size_t smem = (size_t)8 * 2 * C * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ln_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DGTD_CHECK_ARG(e == cudaSuccess, "ln_rows_bwd: cannot opt in to %zu B smem", smem);
  }
  const bool aligned = ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy) |
                         reinterpret_cast<uintptr_t>(w)) & 15) == 0;
#define DGTD_LNB(V)                                                                                             \
  {                                                                                                             \
    if (smem > 48 * 1024) {                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(ln_rows_bwd_v_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      DGTD_CHECK_ARG(e == cudaSuccess, "ln_rows_bwd: cannot opt in to %zu B smem", smem);                       \
    }                                                                                                           \
    ln_rows_bwd_v_kernel<V><<<blocks, 256, smem, s>>>(g, y, w, dy, ws, rows, eps);                              \
  }
  if (aligned && C == 128) DGTD_LNB(1)
  else if (aligned && C == 256) DGTD_LNB(2)
  else if (aligned && C == 512) DGTD_LNB(4)
  else if (aligned && C == 1024) DGTD_LNB(8)
  else ln_rows_bwd_kernel<<<blocks, 256, smem, s>>>(g, y, w, dy, ws, rows, C, eps);
#undef DGTD_LNB
  DGTD_LAUNCH_CHECK("ln_rows_bwd");
  // ws holds [blocks][2C]; dw and db are the two halves of the reduced vector
  sum_splits_kernel<<<cdiv(2 * C, 256), 256, 0, s>>>(ws, ws + (int64_t)blocks * 2 * C, 2 * C, blocks);
  DGTD_LAUNCH_CHECK("ln_rows_bwd.reduce");
  cudaError_t e1 = cudaMemcpyAsync(dw, ws + (int64_t)blocks * 2 * C, C * sizeof(float), cudaMemcpyDeviceToDevice, s);
  cudaError_t e2 = cudaMemcpyAsync(db, ws + (int64_t)blocks * 2 * C + C, C * sizeof(float), cudaMemcpyDeviceToDevice, s);
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    set_error("ln_rows_bwd: gradient copy failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return -4;
  }
  return 0;
}
int dgtd_ln_rows_bwd_ws_floats(int64_t rows, int C) {
  int blocks = cdiv(rows, 8);
  if (blocks > 592) blocks = 592;
  return (blocks + 1) * 2 * C;
}

// y = dwconv7(x; wT (49,C), bias) (+ add); flip = 1 uses the 180-degree rotated taps (input gradient)
int dgtd_dwconv7_fwd(const float* x, const float* wT, const float* bias, const float* add, float* y, int B, int h,
                     int w, int C, int flip, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && wT && y && B > 0 && h > 0 && w > 0 && C % 128 == 0, "dwconv7: bad args (C multiple of 128)");
  if (!flip && (int64_t)B * h * w >= 2048) {   // TMA-staged kernel (the caller passes rotated taps for dgrad)
    int rc = dwconv7_tma(x, wT, bias, add, y, B, h, w, C, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  dwconv7_plain_kernel<<<dim3(C / 128, cdiv(h, 8), B), 128, 0, (cudaStream_t)stream>>>(x, wT, bias, add, y, h, w, C, flip);
  DGTD_LAUNCH_CHECK("dwconv7");
  return 0;
}
// dwT (49,C) and db (C) from x and dy;  ws: (B*cdiv(h,8) + 1) * 50 * C floats
int dgtd_dwconv7_wgrad(const float* x, const float* dy, float* ws, float* dwT, float* db, int B, int h, int w, int C,
                       dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && dy && ws && dwT && db && C % 128 == 0, "dwconv7_wgrad: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int bands = cdiv(h, 8);
  const int64_t n = (int64_t)50 * C;
  float* red = ws + (int64_t)B * bands * n;
  int parts = 0;
  const int rc = dwconv7_wgrad_tma(x, dy, ws, B * bands, &parts, B, h, w, C, s);
  if (rc < 0) return rc;
  if (rc > 0) {   // small maps: thread-per-channel kernel
    dwconv7_wgrad_kernel<<<dim3(C / 128, bands, B), 128, 0, s>>>(x, dy, ws, h, w, C);
    DGTD_LAUNCH_CHECK("dwconv7_wgrad");
    parts = B * bands;
  }
  sum_splits_kernel<<<cdiv(n, 256), 256, 0, s>>>(ws, red, n, parts);
  DGTD_LAUNCH_CHECK("dwconv7_wgrad.reduce");
  cudaError_t e1 = cudaMemcpyAsync(dwT, red, (size_t)49 * C * sizeof(float), cudaMemcpyDeviceToDevice, s);
  cudaError_t e2 = cudaMemcpyAsync(db, red + (int64_t)49 * C, (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, s);
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    set_error("dwconv7_wgrad: gradient copy failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return -4;
  }
  return 0;
}

int dgtd_stem_patchify(const float* image, const float* grid, int G, void* patches, int out_dtype, int B, int H, int W,
                       dgtd_stream_t stream) {
  DGTD_CHECK_ARG(image && patches && B > 0 && H >= 4 && W >= 4, "stem_patchify: bad args");
  const int oh = H / 4, ow = W / 4;
  const int64_t total4 = (int64_t)B * oh * ow * 12;
  if (out_dtype == DGTD_BF16)
    stem_patchify_kernel<<<cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>(image, grid, G, (__nv_bfloat16*)patches, H,
                                                                              W, oh, ow, total4);
  else
    stem_patchify_kernel<<<cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>(image, grid, G, (float*)patches, H, W, oh,
                                                                              ow, total4);
  DGTD_LAUNCH_CHECK("stem_patchify");
  return 0;
}
int dgtd_stem_unpatchify(const float* dpatches, float* dimg, int B, int H, int W, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(dpatches && dimg && B > 0, "stem_unpatchify: bad args");
  const int64_t total = (int64_t)B * 3 * H * W;
  stem_unpatchify_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(dpatches, dimg, H, W, H / 4, W / 4, total);
  DGTD_LAUNCH_CHECK("stem_unpatchify");
  return 0;
}
int dgtd_patchify2(const float* x, float* out, int B, int h, int w, int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out && C % 4 == 0, "patchify2: bad args");
  const int64_t total = (int64_t)B * (h / 2) * (w / 2) * 4 * C;
  patchify2_kernel<<<cdiv(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, out, h, w, C, total);
  DGTD_LAUNCH_CHECK("patchify2");
  return 0;
}
int dgtd_unpatchify2(const float* dp, float* dx, int B, int h, int w, int C, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(dp && dx && C % 4 == 0, "unpatchify2: bad args");
  const int64_t total = (int64_t)B * h * w * C;
  unpatchify2_kernel<<<cdiv(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(dp, dx, h, w, C, total);
  DGTD_LAUNCH_CHECK("unpatchify2");
  return 0;
}
int dgtd_resize_nhwc_bwd(const float* g, float* dx, int B, int h, int w, int C, int oh, int ow, int accumulate,
                         dgtd_stream_t stream) {
  DGTD_CHECK_ARG(g && dx && C % 4 == 0, "resize_nhwc_bwd: bad args");
  const int64_t total = (int64_t)B * h * w * (C / 4);
  resize_nhwc_bwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(g, dx, h, w, C, oh, ow, accumulate, total);
  DGTD_LAUNCH_CHECK("resize_nhwc_bwd");
  return 0;
}

}  // extern "C"
