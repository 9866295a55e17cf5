// Backward of the PVT-v2 attention core (cod.py:911-915) for bf16 q / kv on the warp-level tensor-core path
// (mma.sync m16n8k16, fp32 accumulate) -- the training-side counterpart of attention_mma_kernel (pvt_ops.cu).
//
// With S = scale Q K^T, P = softmax(S), O = P V and D = rowsum(dO . O):   dV = P^T dO,  dP = dO V^T,
// dS = P . (dP - D),  dQ = scale dS K,  dK = scale dS^T Q.
// A warp-level MMA hands its accumulator back in a layout that can feed the NEXT product only as the A operand with
// the same row index, so the two families of products are computed in the orientation that suits them:
//
//   attn_bwd_dq_kernel    rows = queries (CTA = 64 queries of one image / head, 4 warps x 16): pass 1 over the key tiles
//                         gives the log-sum-exp of each row (and D from the saved O), pass 2 recomputes S, forms P, dP and
//                         dS in registers and accumulates dQ = dS K (dS accumulators ARE the A fragments).  Writes dQ,
//                         and lse / D for the second kernel.
//   attn_bwd_dkv_kernel   rows = keys (CTA = 48 keys, 3 warps x 16, x a slice of the queries): S^T = K Q^T and
//                         dP^T = V dO^T per 64-query tile staged in shared memory, P^T / dS^T in registers are the A
//                         fragments of dV += P^T dO and dK += dS^T Q.  The queries are split over CTAs so that the
//                         machine is full when B x heads x key blocks is small (stage 1: 48); every CTA owns its
//                         (split, key rows) slice of a partial buffer and attn_bwd_reduce_kernel sums the splits in
//                         order -- no atomics, bit-stable (the fp32 kernel of pvt_train.cu uses fp32 atomics).
//
// Work: 7 GEMM units instead of the minimal 5 (S twice more, once for the row statistics and once transposed) in exchange
// for no shared-memory transposes; at N_k = 144 the whole thing is ~0.4 GFLOP per image and head.
#include "common.cuh"

namespace dgtd {
namespace {

constexpr int TB_P = 72;     // shared-memory row pitch in bf16 (144 B: conflict-free ldmatrix)
constexpr int TB_TK = 48;    // keys per tile (dq kernel) / per CTA (dkv kernel)
constexpr int TB_QB = 64;    // queries per CTA (dq kernel) / per staged tile (dkv kernel)

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 ldbf2(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

// A fragments (4 k-steps over d = 64) of the 16 rows {r0, r0 + 8} x 64 of a bf16 matrix, times `mul`; rows >= nrows: zero
__device__ __forceinline__ void load_afrag_bf16(uint32_t (&a)[4][4], const __nv_bfloat16* base, int64_t ld, int r0, int r1,
                                                int nrows, int t4, float mul) {
  const bool ok0 = r0 < nrows, ok1 = r1 < nrows;
  const __nv_bfloat16* p0 = base + (int64_t)(ok0 ? r0 : 0) * ld;
  const __nv_bfloat16* p1 = base + (int64_t)(ok1 ? r1 : 0) * ld;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int c0 = ks * 16 + t4 * 2;
    float2 a0 = ldbf2(p0 + c0), a1 = ldbf2(p1 + c0), a2 = ldbf2(p0 + c0 + 8), a3 = ldbf2(p1 + c0 + 8);
    const float m0 = ok0 ? mul : 0.f, m1 = ok1 ? mul : 0.f;
    a[ks][0] = pack2(a0.x * m0, a0.y * m0);
    a[ks][1] = pack2(a1.x * m1, a1.y * m1);
    a[ks][2] = pack2(a2.x * m0, a2.y * m0);
    a[ks][3] = pack2(a3.x * m1, a3.y * m1);
  }
}

// K and V tiles of TB_TK keys (zero beyond Nk) -> shared memory
__device__ __forceinline__ void stage_kv(__nv_bfloat16 (*Ks)[TB_P], __nv_bfloat16 (*Vs)[TB_P], const __nv_bfloat16* kvb, int k0,
                                         int Nk, int C, int nthreads) {
  for (int i = threadIdx.x; i < TB_TK * 8; i += nthreads) {
    const int key = i >> 3, ch = (i & 7) * 8;
    uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
    if (k0 + key < Nk) {
      kk = *reinterpret_cast<const uint4*>(kvb + (int64_t)(k0 + key) * 2 * C + ch);
      vv = *reinterpret_cast<const uint4*>(kvb + (int64_t)(k0 + key) * 2 * C + C + ch);
    }
    *reinterpret_cast<uint4*>(&Ks[key][ch]) = kk;
    *reinterpret_cast<uint4*>(&Vs[key][ch]) = vv;
  }
}

__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                   const __nv_bfloat16* __restrict__ o, const float* __restrict__ dout, float* __restrict__ lse,
                   float* __restrict__ dsum, float* __restrict__ dq, int N, int Nk, int C, float scale) {
  __shared__ __align__(16) __nv_bfloat16 Ks[TB_TK][TB_P];
  __shared__ __align__(16) __nv_bfloat16 Vs[TB_TK][TB_P];
  const int hd = blockIdx.y, b = blockIdx.z, heads = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int r0 = blockIdx.x * TB_QB + warp * 16 + g, r1 = r0 + 8;
  const bool ok0 = r0 < N, ok1 = r1 < N;

  uint32_t qa[4][4], da[4][4];
  load_afrag_bf16(qa, q + (int64_t)b * N * C + hd * 64, C, r0, r1, N, t4, scale);
  float D0 = 0.f, D1 = 0.f;
  {
    const int64_t o0 = ((int64_t)b * N + (ok0 ? r0 : 0)) * C + hd * 64, o1 = ((int64_t)b * N + (ok1 ? r1 : 0)) * C + hd * 64;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c0 = ks * 16 + t4 * 2;
      float2 d00 = *reinterpret_cast<const float2*>(dout + o0 + c0), d10 = *reinterpret_cast<const float2*>(dout + o1 + c0);
      float2 d01 = *reinterpret_cast<const float2*>(dout + o0 + c0 + 8), d11 = *reinterpret_cast<const float2*>(dout + o1 + c0 + 8);
      if (!ok0) { d00 = make_float2(0.f, 0.f); d01 = d00; }
      if (!ok1) { d10 = make_float2(0.f, 0.f); d11 = d10; }
      const float2 x00 = ldbf2(o + o0 + c0), x10 = ldbf2(o + o1 + c0), x01 = ldbf2(o + o0 + c0 + 8), x11 = ldbf2(o + o1 + c0 + 8);
      D0 += d00.x * x00.x + d00.y * x00.y + d01.x * x01.x + d01.y * x01.y;
      D1 += d10.x * x10.x + d10.y * x10.y + d11.x * x11.x + d11.y * x11.y;
      da[ks][0] = pack2(d00.x, d00.y);
      da[ks][1] = pack2(d10.x, d10.y);
      da[ks][2] = pack2(d01.x, d01.y);
      da[ks][3] = pack2(d11.x, d11.y);
    }
    D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
    D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
  }

  const __nv_bfloat16* kvb = kv + (int64_t)b * Nk * 2 * C + hd * 64;
  // ---- pass 1: row log-sum-exp ----
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  for (int k0 = 0; k0 < Nk; k0 += TB_TK) {
    __syncthreads();
    stage_kv(Ks, Vs, kvb, k0, Nk, C, 128);
    __syncthreads();
    float s[6][4];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t kb[4];
        ldsm4(kb, &Ks[j * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]);
        mma16816(s[j], qa[kp * 2], kb[0], kb[1]);
        mma16816(s[j], qa[kp * 2 + 1], kb[2], kb[3]);
      }
    }
    float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int key = k0 + j * 8 + t4 * 2;
      if (key >= Nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (key + 1 >= Nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      tm0 = fmaxf(tm0, fmaxf(s[j][0], s[j][1]));
      tm1 = fmaxf(tm1, fmaxf(s[j][2], s[j][3]));
    }
    tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1)); tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
    tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1)); tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
    const float mn0 = fmaxf(m0, tm0), mn1 = fmaxf(m1, tm1);
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      ps0 += __expf(s[j][0] - mn0) + __expf(s[j][1] - mn0);
      ps1 += __expf(s[j][2] - mn1) + __expf(s[j][3] - mn1);
    }
    ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1); ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
    ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1); ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
    l0 = l0 * __expf(m0 - mn0) + ps0;
    l1 = l1 * __expf(m1 - mn1) + ps1;
    m0 = mn0; m1 = mn1;
  }
  const float lse0 = m0 + __logf(l0), lse1 = m1 + __logf(l1);
  if (t4 == 0) {
    const int64_t base = ((int64_t)b * heads + hd) * N;
    if (ok0) { lse[base + r0] = lse0; dsum[base + r0] = D0; }
    if (ok1) { lse[base + r1] = lse1; dsum[base + r1] = D1; }
  }

  // ---- pass 2: dQ = scale (P . (dO V^T - D)) K ----
  float dqa[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) dqa[j][e] = 0.f;
  for (int k0 = 0; k0 < Nk; k0 += TB_TK) {
    if (Nk > TB_TK || k0 > 0) {          // a single tile is still staged from pass 1
      __syncthreads();
      stage_kv(Ks, Vs, kvb, k0, Nk, C, 128);
      __syncthreads();
    }
    uint32_t dsa[3][4];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t kb[4], vb[4];
        ldsm4(kb, &Ks[j * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]);
        ldsm4(vb, &Vs[j * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]);
        mma16816(s, qa[kp * 2], kb[0], kb[1]);
        mma16816(s, qa[kp * 2 + 1], kb[2], kb[3]);
        mma16816(dp, da[kp * 2], vb[0], vb[1]);
        mma16816(dp, da[kp * 2 + 1], vb[2], vb[3]);
      }
      const int key = k0 + j * 8 + t4 * 2;
      const bool v0 = key < Nk, v1 = key + 1 < Nk;
      const float p0 = v0 ? __expf(s[0] - lse0) : 0.f, p1 = v1 ? __expf(s[1] - lse0) : 0.f;
      const float p2 = v0 ? __expf(s[2] - lse1) : 0.f, p3 = v1 ? __expf(s[3] - lse1) : 0.f;
      dsa[j >> 1][(j & 1) * 2] = pack2(p0 * (dp[0] - D0) * scale, p1 * (dp[1] - D0) * scale);
      dsa[j >> 1][(j & 1) * 2 + 1] = pack2(p2 * (dp[2] - D1) * scale, p3 * (dp[3] - D1) * scale);
    }
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t kb[4];
        ldsm4t(kb, &Ks[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][jp * 16 + (lane >> 4) * 8]);
        mma16816(dqa[jp * 2], dsa[ks], kb[0], kb[1]);
        mma16816(dqa[jp * 2 + 1], dsa[ks], kb[2], kb[3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = hd * 64 + j * 8 + t4 * 2;
    if (ok0) *reinterpret_cast<float2*>(dq + ((int64_t)b * N + r0) * C + col) = make_float2(dqa[j][0], dqa[j][1]);
    if (ok1) *reinterpret_cast<float2*>(dq + ((int64_t)b * N + r1) * C + col) = make_float2(dqa[j][2], dqa[j][3]);
  }
}

// out: (nsplit, B, Nk, 2C) partial [dK | dV] (nsplit == 1: the final dkv)
__global__ void __launch_bounds__(96)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv, const float* __restrict__ dout,
                    const float* __restrict__ lse, const float* __restrict__ dsum, float* __restrict__ out, int N, int Nk,
                    int C, float scale, int nsplit, int tiles_per_split, int64_t split_stride) {
  __shared__ __align__(16) __nv_bfloat16 Qs[TB_QB][TB_P];
  __shared__ __align__(16) __nv_bfloat16 Os[TB_QB][TB_P];
  __shared__ float lses[TB_QB], dss[TB_QB];
  const int kblock = blockIdx.x / nsplit, split = blockIdx.x - kblock * nsplit;
  const int hd = blockIdx.y, b = blockIdx.z, heads = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int kr0 = kblock * TB_TK + warp * 16 + g, kr1 = kr0 + 8;
  const bool kok0 = kr0 < Nk, kok1 = kr1 < Nk;

  uint32_t ka[4][4], va[4][4];
  const __nv_bfloat16* kvb = kv + (int64_t)b * Nk * 2 * C + hd * 64;
  load_afrag_bf16(ka, kvb, 2 * C, kr0, kr1, Nk, t4, scale);
  load_afrag_bf16(va, kvb + C, 2 * C, kr0, kr1, Nk, t4, 1.0f);
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[j][e] = 0.f; dv[j][e] = 0.f; }

  const int ntq = (N + TB_QB - 1) / TB_QB;
  const int t_begin = split * tiles_per_split, t_end = min(ntq, t_begin + tiles_per_split);
  const int64_t sbase = ((int64_t)b * heads + hd) * N;
  for (int qt = t_begin; qt < t_end; ++qt) {
    const int q0 = qt * TB_QB;
    __syncthreads();
    for (int i = threadIdx.x; i < TB_QB * 8; i += 96) {
      const int row = i >> 3, ch = (i & 7) * 8;
      uint4 qq = make_uint4(0u, 0u, 0u, 0u), dd = qq;
      if (q0 + row < N) {
        const int64_t off = ((int64_t)b * N + q0 + row) * C + hd * 64 + ch;
        qq = *reinterpret_cast<const uint4*>(q + off);
        const float4 f0 = *reinterpret_cast<const float4*>(dout + off), f1 = *reinterpret_cast<const float4*>(dout + off + 4);
        dd.x = pack2(f0.x, f0.y); dd.y = pack2(f0.z, f0.w); dd.z = pack2(f1.x, f1.y); dd.w = pack2(f1.z, f1.w);
      }
      *reinterpret_cast<uint4*>(&Qs[row][ch]) = qq;
      *reinterpret_cast<uint4*>(&Os[row][ch]) = dd;
    }
    for (int i = threadIdx.x; i < TB_QB; i += 96) {
      const bool ok = q0 + i < N;
      lses[i] = ok ? lse[sbase + q0 + i] : INFINITY;     // exp(s - inf) = 0: rows beyond N contribute nothing
      dss[i] = ok ? dsum[sbase + q0 + i] : 0.f;
    }
    __syncthreads();

    uint32_t pa[4][4], dsa[4][4];   // P^T and dS^T as A fragments: 4 k-steps of 16 queries
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // 8 query tiles of 8
      float st[4] = {0.f, 0.f, 0.f, 0.f}, dpt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t qb[4], ob[4];
        ldsm4(qb, &Qs[j * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]);
        ldsm4(ob, &Os[j * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]);
        mma16816(st, ka[kp * 2], qb[0], qb[1]);
        mma16816(st, ka[kp * 2 + 1], qb[2], qb[3]);
        mma16816(dpt, va[kp * 2], ob[0], ob[1]);
        mma16816(dpt, va[kp * 2 + 1], ob[2], ob[3]);
      }
      const int qc = j * 8 + t4 * 2;
      const float la = lses[qc], lb = lses[qc + 1], da_ = dss[qc], db_ = dss[qc + 1];
      const float p0 = kok0 ? __expf(st[0] - la) : 0.f, p1 = kok0 ? __expf(st[1] - lb) : 0.f;
      const float p2 = kok1 ? __expf(st[2] - la) : 0.f, p3 = kok1 ? __expf(st[3] - lb) : 0.f;
      pa[j >> 1][(j & 1) * 2] = pack2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack2(p2, p3);
      dsa[j >> 1][(j & 1) * 2] = pack2(p0 * (dpt[0] - da_) * scale, p1 * (dpt[1] - db_) * scale);
      dsa[j >> 1][(j & 1) * 2 + 1] = pack2(p2 * (dpt[2] - da_) * scale, p3 * (dpt[3] - db_) * scale);
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t ob[4], qb[4];
        ldsm4t(ob, &Os[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][jp * 16 + (lane >> 4) * 8]);
        ldsm4t(qb, &Qs[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][jp * 16 + (lane >> 4) * 8]);
        mma16816(dv[jp * 2], pa[ks], ob[0], ob[1]);
        mma16816(dv[jp * 2 + 1], pa[ks], ob[2], ob[3]);
        mma16816(dk[jp * 2], dsa[ks], qb[0], qb[1]);
        mma16816(dk[jp * 2 + 1], dsa[ks], qb[2], qb[3]);
      }
    }
  }
  float* dst = out + (int64_t)split * split_stride + (int64_t)b * Nk * 2 * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = hd * 64 + j * 8 + t4 * 2;
    if (kok0) {
      *reinterpret_cast<float2*>(dst + (int64_t)kr0 * 2 * C + col) = make_float2(dk[j][0], dk[j][1]);
      *reinterpret_cast<float2*>(dst + (int64_t)kr0 * 2 * C + C + col) = make_float2(dv[j][0], dv[j][1]);
    }
    if (kok1) {
      *reinterpret_cast<float2*>(dst + (int64_t)kr1 * 2 * C + col) = make_float2(dk[j][2], dk[j][3]);
      *reinterpret_cast<float2*>(dst + (int64_t)kr1 * 2 * C + C + col) = make_float2(dv[j][2], dv[j][3]);
    }
  }
}

__global__ void __launch_bounds__(256) attn_bwd_reduce_kernel(const float* __restrict__ part, int nsplit, int64_t n4,
                                                              float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = *reinterpret_cast<const float4*>(part + i * 4);
    for (int k = 1; k < nsplit; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(part + (int64_t)k * n4 * 4 + i * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i * 4) = s;
  }
}

int tc_nsplit(int B, int N, int Nk, int heads) {
  const int ctas = B * heads * cdiv(Nk, TB_TK);
  const int ntq = cdiv(N, TB_QB);
  int ns = cdiv(2 * 148, ctas);
  if (ns > ntq) ns = ntq;
  if (ns > 32) ns = 32;
  return ns < 1 ? 1 : ns;
}

}  // namespace

int64_t attention_bwd_tc_ws_floats(int B, int N, int Nk, int heads) {
  const int ns = tc_nsplit(B, N, Nk, heads);
  return 2ll * B * N * heads + (ns > 1 ? (int64_t)ns * B * Nk * 2 * heads * 64 : 0);
}

// bf16 q / kv / o.  Launches three kernels (two when the queries are not split); the caller checks the launch.
int attention_bwd_tc(const void* q, const void* kv, const void* o, const float* dout, float* dq, float* dkv, float* ws, int B,
                     int N, int Nk, int heads, float scale, cudaStream_t s) {
  const int C = heads * 64;
  float* lse = ws;
  float* dsum = ws + (int64_t)B * N * heads;
  float* part = dsum + (int64_t)B * N * heads;
  attn_bwd_dq_kernel<<<dim3(cdiv(N, TB_QB), heads, B), 128, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)kv,
                                                                   (const __nv_bfloat16*)o, dout, lse, dsum, dq, N, Nk, C, scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("attention_bwd(tc): dq launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  count_launch();
  const int ns = tc_nsplit(B, N, Nk, heads);
  const int tps = cdiv(cdiv(N, TB_QB), ns);
  const int64_t stride = (int64_t)B * Nk * 2 * C;
  attn_bwd_dkv_kernel<<<dim3(cdiv(Nk, TB_TK) * ns, heads, B), 96, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)kv, dout,
                                                                         lse, dsum, ns > 1 ? part : dkv, N, Nk, C, scale, ns,
                                                                         tps, stride);
  if (ns > 1) {
    e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("attention_bwd(tc): dkv launch failed: %s", cudaGetErrorString(e));
      return -2;
    }
    count_launch();
    const int64_t n4 = stride / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    attn_bwd_reduce_kernel<<<(unsigned)blocks, 256, 0, s>>>(part, ns, n4, dkv);
  }
  return 0;
}

}  // namespace dgtd
