// fp32 CUDA-core GEMM family: the exact path (<=1e-4 of the reference through 36 residual
// blocks needs true fp32 products; TF32/bf16 tensor-core products do not hold it), and the
// dense projector products of the FFT high-pass.
//
//   C[m,n] = sum_k A[m,k] * B[n,k]        128x128x16 tile, 256 threads, 8x8 per thread,
//                                         two smem stages, operand loaders / epilogue as functors
#pragma once
#include "common.cuh"

namespace dgtd {

// ---------------------------------------------------------------- operand loaders
// A loader contract:  load(m, k, z) -> 4 consecutive k of row m   (KMAJOR = true)
//                     load(k, m, z) -> 4 consecutive m of row k   (KMAJOR = false)
struct RowMajorLoader {   // element (r, c) at p[z*bs + r*ld + c], 4 consecutive c
  const float* p;
  int64_t ld, bs;
  int R, Ccols;
  __device__ __forceinline__ float4 load(int r, int c, int z) const {
    if (r < R && c < Ccols) return load4(p + z * bs + (int64_t)r * ld + c);
    return make_float4(0.f, 0.f, 0.f, 0.f);
  }
};

// Split-K view for weight gradients: batch z covers rows [z*zrows, (z+1)*zrows) of an R-row matrix.
// Optional per-row / per-column scaling folds `keep[m / rows_per_sample] * gamma[c]` into the load
// (the DropPath / layer-scale factors of cod.py:1112-1116).
struct SplitRowLoader {
  const float* p;
  int64_t ld;
  int R, Ccols, zrows;
  const float* keep;    // nullable
  const float* gamma;   // nullable
  int rows_per_sample;
  __device__ __forceinline__ float4 load(int r, int c, int z) const {
    const int gr = z * zrows + r;
    if (r >= zrows || gr >= R || c >= Ccols) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = load4(p + (int64_t)gr * ld + c);
    if (keep) {
      const float k = keep[gr / rows_per_sample];
      v.x *= k; v.y *= k; v.z *= k; v.w *= k;
    }
    if (gamma) {
      const float4 g = load4(gamma + c);
      v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    }
    return v;
  }
};

// NHWC convolution gather: row m = (b, oy, ox), col k = (tap, c), tap-major.
struct Im2colLoader {
  const float* x;
  int h, w, ldx, Cin, oh, ow, ks, stride, off, M, K;
  int zrows = 0;   // > 0: split-K view, batch z covers rows [z*zrows, (z+1)*zrows)
  __device__ __forceinline__ float4 load(int m, int k, int z) const {
    if (zrows > 0) {
      if (m >= zrows) return make_float4(0.f, 0.f, 0.f, 0.f);
      m += z * zrows;
    }
    if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    int tap = k / Cin, c = k - tap * Cin;
    int ty = tap / ks, tx = tap - ty * ks;
    int ox = m % ow;
    int t = m / ow;
    int oy = t % oh;
    int b = t / oh;
    int iy = oy * stride + off + ty, ix = ox * stride + off + tx;
    if ((unsigned)iy >= (unsigned)h || (unsigned)ix >= (unsigned)w)
      return make_float4(0.f, 0.f, 0.f, 0.f);
    return load4(x + ((int64_t)(b * h + iy) * w + ix) * ldx + c);
  }
};

// ---------------------------------------------------------------- kernel
template <bool A_KMAJOR, bool B_KMAJOR, class AL, class BL, class EP>
__global__ void __launch_bounds__(256) simt_gemm_kernel(AL al, BL bl, EP ep, int M, int N, int K) {
  constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, z = blockIdx.z;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;  // 512 float4 per operand tile
      if (A_KMAJOR) {
        int r = idx >> 2, kq = (idx & 3) * 4;
        ra[i] = al.load(m0 + r, k0 + kq, z);
      } else {
        int kk = idx >> 5, mq = (idx & 31) * 4;
        ra[i] = al.load(k0 + kk, m0 + mq, z);
      }
      if (B_KMAJOR) {
        int r = idx >> 2, kq = (idx & 3) * 4;
        rb[i] = bl.load(n0 + r, k0 + kq, z);
      } else {
        int kk = idx >> 5, nq = (idx & 31) * 4;
        rb[i] = bl.load(k0 + kk, n0 + nq, z);
      }
    }
  };
  auto sstore = [&](int st) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;
      if (A_KMAJOR) {
        int r = idx >> 2, kq = (idx & 3) * 4;
        As[st][kq + 0][r] = ra[i].x;
        As[st][kq + 1][r] = ra[i].y;
        As[st][kq + 2][r] = ra[i].z;
        As[st][kq + 3][r] = ra[i].w;
      } else {
        int kk = idx >> 5, mq = (idx & 31) * 4;
        *reinterpret_cast<float4*>(&As[st][kk][mq]) = ra[i];
      }
      if (B_KMAJOR) {
        int r = idx >> 2, kq = (idx & 3) * 4;
        Bs[st][kq + 0][r] = rb[i].x;
        Bs[st][kq + 1][r] = rb[i].y;
        Bs[st][kq + 2][r] = rb[i].z;
        Bs[st][kq + 3][r] = rb[i].w;
      } else {
        int kk = idx >> 5, nq = (idx & 31) * 4;
        *reinterpret_cast<float4*>(&Bs[st][kk][nq]) = rb[i];
      }
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int st = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[st][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[st][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[st][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[st][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(st ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int n = n0 + jh * 64 + tx * 4;
      if (n >= N) continue;
      ep(m, n, z, make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2],
                              acc[i][jh * 4 + 3]));
    }
  }
}

template <bool A_KMAJOR, bool B_KMAJOR, class AL, class BL, class EP>
static inline void launch_simt_gemm(AL al, BL bl, EP ep, int M, int N, int K, int batches,
                                    cudaStream_t s) {
  dim3 grid(cdiv(N, 128), cdiv(M, 128), batches);
  simt_gemm_kernel<A_KMAJOR, B_KMAJOR, AL, BL, EP><<<grid, 256, 0, s>>>(al, bl, ep, M, N, K);
}

// ---------------------------------------------------------------- epilogues
// out = act(acc + bias[n])
template <typename OT, int ACT>
struct EpiBiasAct {
  OT* out;
  const float* bias;  // nullable
  int64_t ldo;
  __device__ __forceinline__ void operator()(int m, int n, int, float4 v) const {
    if (bias) {
      float4 bb = load4(bias + n);
      v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
    }
    store4(out + (int64_t)m * ldo + n, apply_act<ACT>(v.x), apply_act<ACT>(v.y),
           apply_act<ACT>(v.z), apply_act<ACT>(v.w));
  }
};

// out = residual + keep[m / rows_per_sample] * gamma[n] * (acc + bias[n])
struct EpiResidual {
  float* out;
  const float* bias;
  const float* gamma;  // nullable
  const float* keep;   // nullable
  const float* residual;
  int rows_per_sample;
  int64_t ld;
  __device__ __forceinline__ void operator()(int m, int n, int, float4 v) const {
    float4 bb = load4(bias + n);
    float4 g = gamma ? load4(gamma + n) : make_float4(1.f, 1.f, 1.f, 1.f);
    float ks = keep ? keep[m / rows_per_sample] : 1.f;
    float4 r = load4(residual + (int64_t)m * ld + n);
    store4(out + (int64_t)m * ld + n, r.x + ks * (g.x * (v.x + bb.x)), r.y + ks * (g.y * (v.y + bb.y)),
           r.z + ks * (g.z * (v.z + bb.z)), r.w + ks * (g.w * (v.w + bb.w)));
  }
};

}  // namespace dgtd
