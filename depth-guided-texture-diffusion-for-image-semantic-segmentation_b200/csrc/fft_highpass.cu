// a2: prompt_encoder.fft (cod.py:1256-1271) without an FFT.
//
// Zeroing the centred (2*line)^2 block of the shifted spectrum removes the projection of the
// image onto the frequency band k in [-line, line-1] of each axis, so
//     out = | x - Re( P_h x P_w^T ) |,   P = E E^H / n  (n x n, Hermitian, circulant)
//         = | x - (A_h x A_w^T - B_h x B_w^T) |,   A = Re P,  B = Im P.
// A is a real symmetric circulant (dense product on the fp32 FMA pipe: two GEMMs per plane);
// B comes only from the unpaired bin k = -line:  B[a,b] = -sin(th_a - th_b)/n with
// th_a = 2 pi line a / n, i.e. rank 2, so B_h x B_w^T collapses to four scalars per plane.
#include "simt_gemm.cuh"

namespace dgtd {

// row 0 of A: r[d] = (1/n) sum_{k=-line}^{line-1} cos(2 pi k d / n); sc = [sin th | cos th].
__global__ void projector_row_kernel(float* __restrict__ P, float* __restrict__ sc, int n, int line) {
  int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const double two_pi = 6.283185307179586476925286766559;
  double acc = 0.0;
  for (int k = -line; k < line; ++k) {
    long long r = ((long long)k * d) % n;  // exact angle reduction
    if (r < 0) r += n;
    acc += cos(two_pi * (double)r / (double)n);
  }
  P[d] = (float)(acc / n);
  long long r = ((long long)line * d) % n;
  double th = two_pi * (double)r / (double)n;
  sc[d] = (float)sin(th);
  sc[n + d] = (float)cos(th);
}
// A[a,b] = r[(a-b) mod n] (r even, so row 0 is r itself)
__global__ void projector_fill_kernel(float* __restrict__ P, int n) {
  int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y + 1;
  if (b >= n || a >= n) return;
  int d = a - b;
  if (d < 0) d += n;
  P[(int64_t)a * n + b] = P[d];
}

// coef[z] = {Qcc, Qcs, Qsc, Qss} / (H*W),  Qcs = sum_{b,c} cos_h[b] x[b,c] sin_w[c], ...
__global__ void __launch_bounds__(256)
imag_coef_kernel(const float* __restrict__ x, const float* __restrict__ sc_h,
                 const float* __restrict__ sc_w, float* __restrict__ coef, int H, int W) {
  __shared__ float red[4][8];
  const float* p = x + (int64_t)blockIdx.x * H * W;
  float q[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < H * W; i += 256) {
    int r = i / W, c = i - r * W;
    float v = p[i];
    float sh = sc_h[r], ch = sc_h[H + r], sw = sc_w[c], cw = sc_w[W + c];
    q[0] = fmaf(ch * cw, v, q[0]);
    q[1] = fmaf(ch * sw, v, q[1]);
    q[2] = fmaf(sh * cw, v, q[2]);
    q[3] = fmaf(sh * sw, v, q[3]);
  }
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float s = warp_sum(q[j]);
    if (lane == 0) red[j][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    coef[blockIdx.x * 4 + threadIdx.x] = s / ((float)H * (float)W);
  }
}

struct EpiStore {
  float* out;
  int64_t ld, bs;
  __device__ __forceinline__ void operator()(int m, int n, int z, float4 v) const {
    store4(out + z * bs + (int64_t)m * ld + n, v.x, v.y, v.z, v.w);
  }
};

// out = | x - acc + (s_a s'_d Qcc - s_a c'_d Qcs - c_a s'_d Qsc + c_a c'_d Qss) |
struct EpiHighpass {
  const float* x;
  const float* sc_h;
  const float* sc_w;
  const float* coef;
  float* out;
  int H, W;
  __device__ __forceinline__ void operator()(int m, int n, int z, float4 v) const {
    const int64_t o = ((int64_t)z * H + m) * W + n;
    float4 xv = load4(x + o);
    float sa = sc_h[m], ca = sc_h[H + m];
    float4 q = load4(coef + z * 4);
    float4 sd = load4(sc_w + n), cd = load4(sc_w + W + n);
    float u = sa * q.x - ca * q.z;   // multiplies s'_d
    float t = ca * q.w - sa * q.y;   // multiplies c'_d
    store4(out + o, fabsf(xv.x - v.x + (u * sd.x + t * cd.x)), fabsf(xv.y - v.y + (u * sd.y + t * cd.y)),
           fabsf(xv.z - v.z + (u * sd.z + t * cd.z)), fabsf(xv.w - v.w + (u * sd.w + t * cd.w)));
  }
};

}  // namespace dgtd

using namespace dgtd;

extern "C" {

int dgtd_lowpass_projector(float* P, float* sc, int n, int line, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(P && sc && n > 0 && line >= 0 && 2 * line <= n, "lowpass_projector: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  projector_row_kernel<<<cdiv(n, 128), 128, 0, s>>>(P, sc, n, line);
  DGTD_LAUNCH_CHECK("lowpass_projector.row");
  if (n > 1) {
    projector_fill_kernel<<<dim3(cdiv(n, 128), n - 1), 128, 0, s>>>(P, n);
    DGTD_LAUNCH_CHECK("lowpass_projector.fill");
  }
  return 0;
}

int dgtd_fft_highpass_fwd(const float* x, const float* Ph, const float* Pw, const float* sc_h,
                          const float* sc_w, float* tmp, float* coef, float* out, int planes,
                          int H, int W, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && Ph && Pw && sc_h && sc_w && tmp && coef && out, "fft_highpass: null pointer");
  DGTD_CHECK_ARG(planes > 0 && H > 0 && W > 0 && (W % 4) == 0 && (H % 4) == 0,
                 "fft_highpass: H and W must be multiples of 4 (got %dx%d)", H, W);
  DGTD_CHECK_ARG(planes <= 65535, "fft_highpass: too many planes");
  cudaStream_t s = (cudaStream_t)stream;
  imag_coef_kernel<<<planes, 256, 0, s>>>(x, sc_h, sc_w, coef, H, W);
  DGTD_LAUNCH_CHECK("fft_highpass.coef");
  {  // tmp = x . A_w^T over all (plane,row) rows at once
    RowMajorLoader al{x, W, 0, planes * H, W};
    RowMajorLoader bl{Pw, W, 0, W, W};
    EpiStore ep{tmp, W, 0};
    launch_simt_gemm<true, true>(al, bl, ep, planes * H, W, W, 1, s);
    DGTD_LAUNCH_CHECK("fft_highpass.rows");
  }
  {  // per plane: A_h . tmp, fused with the rank-2 imaginary term, the subtraction and |.|
    RowMajorLoader al{Ph, H, 0, H, H};
    RowMajorLoader bl{tmp, W, (int64_t)H * W, H, W};  // (K x N) row major, N contiguous
    EpiHighpass ep{x, sc_h, sc_w, coef, out, H, W};
    launch_simt_gemm<true, false>(al, bl, ep, H, W, H, planes, s);
    DGTD_LAUNCH_CHECK("fft_highpass.cols");
  }
  return 0;
}

}  // extern "C"
