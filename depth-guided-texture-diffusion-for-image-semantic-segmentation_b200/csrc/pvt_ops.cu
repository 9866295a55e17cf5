// PVT-v2 blocks that consume the texture prompts (cod.py:824-961, 1455-1509): the
// non-GEMM pieces.  Token streams are (B, N, C) = NHWC, fp32 residual stream, GEMM operands bf16 (or fp32
// in exact mode); every projection runs on the GEMM kernels of the hot path.
//   ln_tokens      x (+ prompt) -> LayerNorm over C      (cod.py:1471-1472 `x + prompt[i]`, :958/:960 norm1/2)
//   patchify       spatial-reduction conv k = stride = sr as a patch gather + GEMM (cod.py:887,903)
//   attention      softmax(q k^T / sqrt(d)) v with N_kv <= a few hundred keys (cod.py:911-915), fp32 math
//   dwconv3_gelu   Mlp's depthwise 3x3 + GELU(erf) on the hidden tokens (cod.py:852-854, 1520-1531)
#include "tc_common.cuh"

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

// LPR lanes per token row (32 / LPR rows per warp: narrow stage-1 rows of 64 channels use half a warp each).
// sum_out (nullable, fp32) receives x + add; out = LN(x + add) * w + b.
template <int LPR>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
  for (int off = LPR / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
// U row groups per warp: all loads of the U rows a lane touches are issued before the first reduction (with one row per
// lane group the C <= 128 stages had a single 16-byte load per operand in flight per lane and ran at half the HBM rate).
template <typename AT, typename OT, int VPL, int LPR, int U>
__global__ void __launch_bounds__(256)
ln_tokens_kernel(const float* __restrict__ x, const AT* __restrict__ add, float* __restrict__ sum_out,
                 const float* __restrict__ ln_w, const float* __restrict__ ln_b, OT* __restrict__ out, int64_t rows,
                 int C, float eps) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & (LPR - 1);
  const int64_t row0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * (RPW * U) + ((threadIdx.x & 31) / LPR);
  const int nq = C >> 2;
  float4 v[U][VPL];
  int64_t rowu[U];
  bool liveu[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    int64_t row = row0 + u * RPW;
    liveu[u] = row < rows;
    if (!liveu[u]) row = rows - 1;      // keep the whole warp in the shuffles; stores are predicated
    rowu[u] = row;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int q = j * LPR + lane;
      if (q < nq) v[u][j] = *reinterpret_cast<const float4*>(x + row * C + q * 4);
    }
  }
  if (add) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 a[VPL];
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int q = j * LPR + lane;
        if (q < nq) a[j] = ld4<AT>(add + rowu[u] * C + q * 4);
      }
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int q = j * LPR + lane;
        if (q < nq) { v[u][j].x += a[j].x; v[u][j].y += a[j].y; v[u][j].z += a[j].z; v[u][j].w += a[j].w; }
      }
    }
  }
  float4 g[VPL], be[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int q = j * LPR + lane;
    if (q < nq) {
      g[j] = *reinterpret_cast<const float4*>(ln_w + q * 4);
      be[j] = *reinterpret_cast<const float4*>(ln_b + q * 4);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t row = rowu[u];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int q = j * LPR + lane;
      if (q < nq) {
        if (sum_out && liveu[u]) *reinterpret_cast<float4*>(sum_out + row * C + q * 4) = v[u][j];
        s += (v[u][j].x + v[u][j].y) + (v[u][j].z + v[u][j].w);
      }
    }
    const float mean = seg_sum<LPR>(s) / C;
    float qq = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      if (j * LPR + lane < nq) {
        const float a = v[u][j].x - mean, b = v[u][j].y - mean, c = v[u][j].z - mean, d = v[u][j].w - mean;
        qq += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = 1.0f / sqrtf(seg_sum<LPR>(qq) / C + eps);
    if (liveu[u]) {
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int q = j * LPR + lane;
        if (q < nq)
          store4(out + row * C + q * 4, (v[u][j].x - mean) * rstd * g[j].x + be[j].x, (v[u][j].y - mean) * rstd * g[j].y + be[j].y,
                 (v[u][j].z - mean) * rstd * g[j].z + be[j].z, (v[u][j].w - mean) * rstd * g[j].w + be[j].w);
      }
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

// depthwise 3x3 (pad 1) + bias + GELU on NHWC tokens; wT is (9, C).
// thread = 8 channels x 4 consecutive pixels of one row: the 3 x 6 input window is loaded once (16-byte loads
// for bf16) and feeds 4 outputs.  bf16 storage uses the packed polynomial GELU of the GEMM epilogues
// (|err| <= 2.8e-5), fp32 storage the exact erf form.
__device__ __forceinline__ void ld8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void st8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = u;
}

// CTA = 64 channels (8 threads x 8 channels) x 32 groups of 4 pixels; the 9 x 64 taps of the CTA's channel group
// sit in shared memory and are re-read per kernel row (24 registers instead of 72).
template <typename T, bool GELU = true>
__global__ void __launch_bounds__(256)
dwconv3_gelu_kernel(const T* __restrict__ x, const float* __restrict__ wT, const float* __restrict__ bias,
                    T* __restrict__ out, int h, int w, int C) {
  constexpr int PX = 4;
  __shared__ __align__(16) float ws[10][64];   // 9 taps + bias
  const int c0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 640; i += 256) {
    const int j = i >> 6, cc = i & 63;
    ws[j][cc] = c0 + cc < C ? (j < 9 ? wT[j * C + c0 + cc] : bias[c0 + cc]) : 0.f;
  }
  __syncthreads();
  const int co = (threadIdx.x & 7) * 8;
  const int c = c0 + co;
  // CTA = one 8-row x 16-pixel patch (32 groups of 4 pixels, 4 across x 8 down): the rows oy-1 / oy+1 a thread
  // needs are the rows its neighbours in the CTA read anyway, so the vertical re-reads hit L1 instead of L2
  // (10 input rows per 8 output rows instead of 3 per 1).
  const int g = threadIdx.x >> 3;
  const int patches_x = (w + 15) / 16, patches_y = (h + 7) / 8;
  int t = blockIdx.y;
  const int ox0 = ((t % patches_x) * 4 + (g & 3)) * PX; t /= patches_x;
  const int oy = (t % patches_y) * 8 + (g >> 2);
  const int b = t / patches_y;
  if (ox0 >= w || oy >= h || c >= C) return;
  // packed f32x2 arithmetic: 8 channels = 4 register pairs (half the FMA / GELU issue slots)
  uint64_t acc[PX][4];
  {
    float bs[8];
    ld8(&ws[9][co], bs);
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[p][e] = pk2(bs[2 * e], bs[2 * e + 1]);
  }
  if constexpr (sizeof(T) == 2) {
    // bf16 storage.  Addressing is hoisted: six clamped column offsets (shared by the three rows) and three clamped
    // row pointers; a warp is one 16-pixel row of the patch, so the row test is warp-uniform (a skipped row costs
    // nothing) and only the two outer columns of the 6-wide window can fall outside the image in a full group --
    // the masks are applied to those packed words only.
    int coff[PX + 2];
#pragma unroll
    for (int j = 0; j < PX + 2; ++j) coff[j] = min(max(ox0 - 1 + j, 0), w - 1) * C;
    const bool partial = ox0 + PX > w;            // last group of a row when w is not a multiple of 4
    const bool left_ok = ox0 > 0, right_ok = ox0 + PX < w;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy + ky - 1;
      if ((unsigned)iy >= (unsigned)h) continue;  // warp-uniform
      const T* row = x + (((int64_t)b * h + iy) * w) * C + c;
      uint4 raw[PX + 2];
#pragma unroll
      for (int j = 0; j < PX + 2; ++j) raw[j] = __ldg(reinterpret_cast<const uint4*>(row + coff[j]));
      if (!left_ok) raw[0] = make_uint4(0u, 0u, 0u, 0u);
      if (!right_ok) raw[PX + 1] = make_uint4(0u, 0u, 0u, 0u);
      if (partial) {
#pragma unroll
        for (int j = 1; j <= PX; ++j)
          if (ox0 - 1 + j >= w) raw[j] = make_uint4(0u, 0u, 0u, 0u);
      }
      uint64_t k[3][4];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float kf[8];
        ld8(&ws[ky * 3 + kx][co], kf);
#pragma unroll
        for (int e = 0; e < 4; ++e) k[kx][e] = pk2(kf[2 * e], kf[2 * e + 1]);
      }
#pragma unroll
      for (int j = 0; j < PX + 2; ++j) {
        const uint32_t u[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
        uint64_t v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)   // bf16 pair -> fp32 pair: low half << 16, high half masked
          v[e] = pk2(__uint_as_float(u[e] << 16), __uint_as_float(u[e] & 0xffff0000u));
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int p = j - kx;
          if (p < 0 || p >= PX) continue;
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[p][e] = fma2(v[e], k[kx][e], acc[p][e]);
        }
      }
    }
  } else {
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy + ky - 1;
    if ((unsigned)iy >= (unsigned)h) continue;
    uint64_t k[3][4];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      float kf[8];
      ld8(&ws[ky * 3 + kx][co], kf);
#pragma unroll
      for (int e = 0; e < 4; ++e) k[kx][e] = pk2(kf[2 * e], kf[2 * e + 1]);
    }
    const T* row = x + (((int64_t)b * h + iy) * w) * C + c;
    // all six column loads of the row are issued back to back (clamped address, zeroed afterwards): no
    // branch between them, so they overlap instead of paying one memory round trip each
    float vf[PX + 2][8];
#pragma unroll
    for (int j = 0; j < PX + 2; ++j) {
      const int ix = ox0 - 1 + j;
      ld8(row + (int64_t)min(max(ix, 0), w - 1) * C, vf[j]);
    }
#pragma unroll
    for (int j = 0; j < PX + 2; ++j) {   // input column ox0 - 1 + j feeds outputs p = j - kx, kx in 0..2
      const int ix = ox0 - 1 + j;
      const float ok = (unsigned)ix < (unsigned)w ? 1.f : 0.f;
      uint64_t v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = pk2(vf[j][2 * e] * ok, vf[j][2 * e + 1] * ok);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int p = j - kx;
        if (p < 0 || p >= PX) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[p][e] = fma2(v[e], k[kx][e], acc[p][e]);
      }
    }
  }
  }
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    if (ox0 + p >= w) break;
    float r[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      up2(acc[p][e], r[2 * e], r[2 * e + 1]);
      if (GELU) {
        if (sizeof(T) == 2) gelu_fast2(r[2 * e], r[2 * e + 1]);   // the 2.8e-5 polynomial: the tanh form measured here
        // cost 0.29 ms less per full-model step and moved the 128-pixel logit error 1.7e-2 -> 2.05e-2 (r2): not taken
        else { r[2 * e] = gelu_erf(r[2 * e]); r[2 * e + 1] = gelu_erf(r[2 * e + 1]); }
      }
    }
    st8(out + (((int64_t)b * h + oy) * w + ox0 + p) * C + c, r);
  }
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


// ---- bf16 attention on the warp-level tensor-core path (mma.sync m16n8k16, fp32 accumulate) -----------------
// K/V sets are tiny (144 keys x 64 channels per head), so the whole product is two small GEMMs per 16-query
// warp tile with an online softmax in between (flash-attention register layout: the S accumulators of two
// adjacent key tiles ARE the A fragment of the P.V product).  CTA = 64 queries of one (image, head), 4 warps;
// keys stream through shared memory in tiles of 48 (144 = 3 x 48; 72-element row pitch = conflict-free ldmatrix).
constexpr int FA_TK = 48, FA_PITCH = 72, FA_QB = 64;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(128)
attention_mma_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                     __nv_bfloat16* __restrict__ out, int N, int Nk, int C, float scale) {
  __shared__ __align__(16) __nv_bfloat16 Ks[FA_TK][FA_PITCH];
  __shared__ __align__(16) __nv_bfloat16 Vs[FA_TK][FA_PITCH];
  const int hd = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int qrow0 = blockIdx.x * FA_QB + warp * 16 + g, qrow1 = qrow0 + 8;   // the two query rows of this thread

  // Q fragments (A operand of S = Q K^T), pre-scaled: 4 k-steps over d = 64
  uint32_t qa[4][4];
  {
    const __nv_bfloat16* q0 = q + ((int64_t)b * N + min(qrow0, N - 1)) * C + hd * 64;
    const __nv_bfloat16* q1 = q + ((int64_t)b * N + min(qrow1, N - 1)) * C + hd * 64;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c0 = ks * 16 + t4 * 2;
      const float2 a0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q0 + c0));
      const float2 a1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q1 + c0));
      const float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q0 + c0 + 8));
      const float2 a3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q1 + c0 + 8));
      qa[ks][0] = pack_bf16(a0.x * scale, a0.y * scale);
      qa[ks][1] = pack_bf16(a1.x * scale, a1.y * scale);
      qa[ks][2] = pack_bf16(a2.x * scale, a2.y * scale);
      qa[ks][3] = pack_bf16(a3.x * scale, a3.y * scale);
    }
  }
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const __nv_bfloat16* kvb = kv + (int64_t)b * Nk * 2 * C + hd * 64;
  for (int k0 = 0; k0 < Nk; k0 += FA_TK) {
    __syncthreads();
    for (int i = threadIdx.x; i < FA_TK * 8; i += 128) {   // 16-byte chunks; keys beyond Nk are zero
      const int key = i >> 3, ch = (i & 7) * 8;
      uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
      if (k0 + key < Nk) {
        kk = *reinterpret_cast<const uint4*>(kvb + (int64_t)(k0 + key) * 2 * C + ch);
        vv = *reinterpret_cast<const uint4*>(kvb + (int64_t)(k0 + key) * 2 * C + C + ch);
      }
      *reinterpret_cast<uint4*>(&Ks[key][ch]) = kk;
      *reinterpret_cast<uint4*>(&Vs[key][ch]) = vv;
    }
    __syncthreads();

    // S = Q K^T for 6 key tiles of 8
    float sacc[6][4];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[j][e] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {   // one ldmatrix.x4 = B fragments of two k-steps (32 channels)
        uint32_t kb[4];
        ldsm_x4(kb, &Ks[j * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]);
        mma_bf16_16816(sacc[j], qa[kp * 2], kb[0], kb[1]);
        mma_bf16_16816(sacc[j], qa[kp * 2 + 1], kb[2], kb[3]);
      }
    }
    // mask keys beyond Nk, online softmax (rows g and g + 8)
    float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int key = k0 + j * 8 + t4 * 2;
      if (key >= Nk) { sacc[j][0] = -INFINITY; sacc[j][2] = -INFINITY; }
      if (key + 1 >= Nk) { sacc[j][1] = -INFINITY; sacc[j][3] = -INFINITY; }
      tm0 = fmaxf(tm0, fmaxf(sacc[j][0], sacc[j][1]));
      tm1 = fmaxf(tm1, fmaxf(sacc[j][2], sacc[j][3]));
    }
    tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1)); tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
    tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1)); tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
    const float mn0 = fmaxf(m0, tm0), mn1 = fmaxf(m1, tm1);
    const float c0 = __expf(m0 - mn0), c1 = __expf(m1 - mn1);
    float ps0 = 0.f, ps1 = 0.f;
    uint32_t pa[3][4];   // P as A fragments of the three 16-key k-steps
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const float p0 = __expf(sacc[j][0] - mn0), p1 = __expf(sacc[j][1] - mn0);
      const float p2 = __expf(sacc[j][2] - mn1), p3 = __expf(sacc[j][3] - mn1);
      ps0 += p0 + p1; ps1 += p2 + p3;
      pa[j >> 1][(j & 1) * 2] = pack_bf16(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1); ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
    ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1); ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
    l0 = l0 * c0 + ps0; l1 = l1 * c1 + ps1;
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
    // O += P V : 3 k-steps of 16 keys x 8 channel tiles of 8 (ldmatrix.trans serves two channel tiles)
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t vb[4];
        ldsm_x4_t(vb, &Vs[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][jp * 16 + (lane >> 4) * 8]);
        mma_bf16_16816(o[jp * 2], pa[ks], vb[0], vb[1]);
        mma_bf16_16816(o[jp * 2 + 1], pa[ks], vb[2], vb[3]);
      }
    }
  }
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = hd * 64 + j * 8 + t4 * 2;
    if (qrow0 < N)
      *reinterpret_cast<uint32_t*>(out + ((int64_t)b * N + qrow0) * C + col) = pack_bf16(o[j][0] * i0, o[j][1] * i0);
    if (qrow1 < N)
      *reinterpret_cast<uint32_t*>(out + ((int64_t)b * N + qrow1) * C + col) = pack_bf16(o[j][2] * i1, o[j][3] * i1);
  }
}

}  // namespace dgtd

namespace dgtd {
int dwconv3_gelu_tma(const void* x, const float* wT, const float* bias, void* out, int B, int h, int w, int C,
                     cudaStream_t s);   // dwconv3_tma.cu
}
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
#define DGTD_LNT(AT, OT, V, L, U)                                                                                 \
  ln_tokens_kernel<AT, OT, V, L, U><<<(unsigned)cdiv(rows, (int64_t)(8 * (32 / L) * U)), 256, 0, s>>>(             \
      x, (const AT*)add, sum_out, ln_w, ln_b, (OT*)out, rows, C, eps)
#define DGTD_LNT_V(AT, OT)                         \
  do {                                             \
    if (C <= 64) DGTD_LNT(AT, OT, 1, 16, 4);       \
    else if (C <= 128) DGTD_LNT(AT, OT, 1, 32, 4); \
    else if (C <= 512) DGTD_LNT(AT, OT, 4, 32, 1); \
    else DGTD_LNT(AT, OT, 16, 32, 1);              \
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
  DGTD_CHECK_ARG(x && wT && bias && out && B > 0 && h > 0 && w > 0 && C % 8 == 0, "dwconv3_gelu: bad args (C % 8)");
  const int64_t total = (int64_t)B * cdiv(h, 8) * cdiv(w, 16);   // 8 x 16 pixel patches
  DGTD_CHECK_ARG(total <= 65535, "dwconv3_gelu: too many pixels for one launch");
  dim3 blocks(cdiv(C, 64), (unsigned)total);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DGTD_BF16) {
    const int rc = dwconv3_gelu_tma(x, wT, bias, out, B, h, w, C, s);   // persistent TMA-staged kernel (dwconv3_tma.cu)
    if (rc < 0) return rc;
    if (rc == 0) {
      DGTD_LAUNCH_CHECK("dwconv3_gelu(tma)");
      return 0;
    }
    dwconv3_gelu_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, wT, bias, (__nv_bfloat16*)out, h, w, C);
  } else if (dtype == DGTD_F32)
    dwconv3_gelu_kernel<<<blocks, 256, 0, s>>>((const float*)x, wT, bias, (float*)out, h, w, C);
  else DGTD_CHECK_ARG(false, "dwconv3_gelu: bad dtype %d", dtype);
  DGTD_LAUNCH_CHECK("dwconv3_gelu");
  return 0;
}

// the depthwise 3x3 alone (`DWConv.forward`, cod.py:1520-1531, called on its own by a maintainer)
int dgtd_dwconv3_fwd(const void* x, const float* wT, const float* bias, void* out, int dtype, int B, int h, int w, int C,
                     dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && wT && bias && out && B > 0 && h > 0 && w > 0 && C % 8 == 0, "dwconv3: bad args (C % 8)");
  const int64_t total = (int64_t)B * cdiv(h, 8) * cdiv(w, 16);
  DGTD_CHECK_ARG(total <= 65535, "dwconv3: too many pixels for one launch");
  dim3 blocks(cdiv(C, 64), (unsigned)total);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DGTD_BF16)
    dwconv3_gelu_kernel<__nv_bfloat16, false><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, wT, bias, (__nv_bfloat16*)out, h, w, C);
  else if (dtype == DGTD_F32)
    dwconv3_gelu_kernel<float, false><<<blocks, 256, 0, s>>>((const float*)x, wT, bias, (float*)out, h, w, C);
  else DGTD_CHECK_ARG(false, "dwconv3: bad dtype %d", dtype);
  DGTD_LAUNCH_CHECK("dwconv3");
  return 0;
}

int dgtd_attention_fwd(const void* q, const void* kv, void* out, int dtype, int B, int N, int Nk, int heads,
                       float scale, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(q && kv && out && B > 0 && N > 0 && Nk > 0 && heads > 0 && heads <= 65535 && B <= 65535,
                 "attention: bad args");
  const int C = heads * ATT_D;   // head_dim 64 (pvt_v2: 64/1, 128/2, 320/5, 512/8)
  dim3 grid(cdiv(N, ATT_QB), heads, B);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DGTD_BF16)   // tensor-core path (scores rounded to bf16 before P.V, fp32 accumulate)
    attention_mma_kernel<<<dim3(cdiv(N, FA_QB), heads, B), 128, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)kv,
                                                                        (__nv_bfloat16*)out, N, Nk, C, scale);
  else if (dtype == DGTD_F32)
    attention_kernel<<<grid, 256, 0, s>>>((const float*)q, (const float*)kv, (float*)out, N, Nk, C, scale);
  else DGTD_CHECK_ARG(false, "attention: bad dtype %d", dtype);
  DGTD_LAUNCH_CHECK("attention");
  return 0;
}

}  // extern "C"
