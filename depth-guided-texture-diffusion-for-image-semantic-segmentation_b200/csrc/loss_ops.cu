// Structure loss of the reference (cod.py:75-84): boundary-weighted BCE + IoU.
//   weit = 1 + 5 |avgpool31x31(gt) - gt|   depends on the label only: computed once and shared by the five
//                                           supervised maps of a step (the reference recomputes it per map)
//   forward   per (image, channel) plane: four weighted sums in ONE pass over logits / label / weights
//   backward  element-wise from the four saved sums
// All reductions are two-level with a fixed order (no atomics): bit-stable run to run.
#include "common.cuh"

namespace dgtd {

constexpr int BW_T = 32, BW_R = 15, BW_P = BW_T + 2 * BW_R;   // 32x32 output tile, 62x62 halo

// CTA = one 32x32 tile of one plane: halo tile -> horizontal 31-sums -> vertical 31-sums (separable box filter)
__global__ void __launch_bounds__(256)
boundary_weight_kernel(const float* __restrict__ gt, float* __restrict__ weit, int H, int W) {
  __shared__ float tile[BW_P][BW_P + 1];
  __shared__ float hs[BW_P][BW_T + 1];
  const int plane = blockIdx.z, y0 = blockIdx.y * BW_T, x0 = blockIdx.x * BW_T;
  const float* g = gt + (int64_t)plane * H * W;
  for (int i = threadIdx.x; i < BW_P * BW_P; i += 256) {
    const int r = i / BW_P, c = i - r * BW_P;
    const int y = y0 + r - BW_R, x = x0 + c - BW_R;
    tile[r][c] = ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) ? g[(int64_t)y * W + x] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BW_P * BW_T; i += 256) {
    const int r = i / BW_T, c = i - r * BW_T;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * BW_R + 1; ++k) s += tile[r][c + k];
    hs[r][c] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BW_T * BW_T; i += 256) {
    const int r = i / BW_T, c = i - r * BW_T;
    const int y = y0 + r, x = x0 + c;
    if (y >= H || x >= W) continue;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * BW_R + 1; ++k) s += hs[r + k][c];
    weit[(int64_t)plane * H * W + (int64_t)y * W + x] = 1.f + 5.f * fabsf(s * (1.f / 961.f) - tile[r + BW_R][c + BW_R]);
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) t += red[i];
  __syncthreads();
  return t;   // valid on thread 0
}

// part[plane][blk][4] = (sum w*bce, sum w, sum sig*g*w, sum (sig+g)*w) over the block's pixels
__global__ void __launch_bounds__(256)
structure_loss_partial_kernel(const float* __restrict__ p, const float* __restrict__ g, const float* __restrict__ wt,
                              float* __restrict__ part, int64_t HW, int chunk) {
  __shared__ float red[8];
  const int plane = blockIdx.y;
  const int64_t base = (int64_t)plane * HW;
  const int64_t i0 = (int64_t)blockIdx.x * chunk, i1 = min(HW, i0 + chunk);
  float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    const float x = p[base + i], t = g[base + i], w = wt[base + i];
    const float bce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
    const float s = 1.f / (1.f + expf(-x));
    a += w * bce; b += w; c += s * t * w; d += (s + t) * w;
  }
  float* o = part + ((int64_t)plane * gridDim.x + blockIdx.x) * 4;
  float r;
  r = block_sum_256(a, red); if (threadIdx.x == 0) o[0] = r;
  r = block_sum_256(b, red); if (threadIdx.x == 0) o[1] = r;
  r = block_sum_256(c, red); if (threadIdx.x == 0) o[2] = r;
  r = block_sum_256(d, red); if (threadIdx.x == 0) o[3] = r;
}

// sums[plane][4] from the partials, loss = mean_planes(wbce + wiou); one CTA
__global__ void structure_loss_finalize_kernel(const float* __restrict__ part, float* __restrict__ sums,
                                               float* __restrict__ loss, int planes, int nblk) {
  __shared__ float acc[256];
  float l = 0.f;
  for (int pl = threadIdx.x; pl < planes; pl += 256) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < nblk; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] += part[((int64_t)pl * nblk + k) * 4 + e];
#pragma unroll
    for (int e = 0; e < 4; ++e) sums[pl * 4 + e] = s[e];
    l += s[0] / s[1] + 1.f - (s[2] + 1.f) / (s[3] - s[2] + 1.f);
  }
  acc[threadIdx.x] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 256; ++i) t += acc[i];
    loss[0] = t / planes;
  }
}

// d loss / d logit, scaled by grad_out[0] / planes
__global__ void structure_loss_bwd_kernel(const float* __restrict__ p, const float* __restrict__ g,
                                          const float* __restrict__ wt, const float* __restrict__ sums,
                                          const float* __restrict__ grad_out, float* __restrict__ dp, int64_t HW,
                                          int planes, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int pl = (int)(i / HW);
  const float sw = sums[pl * 4 + 1], I = sums[pl * 4 + 2], U = sums[pl * 4 + 3];
  const float x = p[i], t = g[i], w = wt[i];
  const float s = 1.f / (1.f + expf(-x));
  const float ds = s * (1.f - s);
  const float den = U - I + 1.f;
  const float dI = ds * t * w, dU = ds * w;
  const float diou = -(dI * den - (I + 1.f) * (dU - dI)) / (den * den);
  dp[i] = grad_out[0] / planes * (w * (s - t) / sw + diou);
}

}  // namespace dgtd

using namespace dgtd;


// ---- SSIM constant of the training loss (cod.py:142-144, SSIM :316-351) ---------------------------------------
// loss_3 = mean( clamp((1 - SSIM(x, y)) / 2, 0, 1) ), x = (embedding1 - min) / (max - min + 1e-8) with the min / max
// taken over the WHOLE batch tensor (cod.py:143), y = the input image; 3x3 means over a reflection-padded window.
// No parameter is upstream of it (embedding1 is the parameter-free FFT high-pass), so it has no backward.
// Two fixed-order reduction levels (min/max, then the sum): bit-stable.
namespace dgtd {
constexpr int SSIM_THREADS = 256, SSIM_PER_CTA = 4096;

__global__ void __launch_bounds__(SSIM_THREADS) minmax_partial_kernel(const float* __restrict__ x, int64_t n,
                                                                      float* __restrict__ part) {
  __shared__ float smin[8], smax[8];
  float lo = INFINITY, hi = -INFINITY;
  const int64_t i0 = (int64_t)blockIdx.x * SSIM_PER_CTA;
  for (int64_t i = i0 + threadIdx.x; i < min(n, i0 + SSIM_PER_CTA); i += SSIM_THREADS) {
    const float v = x[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { lo = fminf(lo, smin[w]); hi = fmaxf(hi, smax[w]); }
    part[2 * blockIdx.x] = fminf(lo, smin[0]);
    part[2 * blockIdx.x + 1] = fmaxf(hi, smax[0]);
  }
}

__global__ void minmax_final_kernel(const float* __restrict__ part, int nblk, float* __restrict__ mm) {
  __shared__ float smin[8], smax[8];
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) { lo = fminf(lo, part[2 * i]); hi = fmaxf(hi, part[2 * i + 1]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < 8; ++w) { lo = fminf(lo, smin[w]); hi = fmaxf(hi, smax[w]); }
    mm[0] = lo;
    mm[1] = hi;
  }
}

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(SSIM_THREADS) ssim_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                    const float* __restrict__ mm, int H, int W, int64_t n,
                                                                    float* __restrict__ part) {
  __shared__ float red[8];
  const float lo = mm[0], inv = 1.0f / (mm[1] - mm[0] + 1e-8f);
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  float acc = 0.f;
  const int64_t i0 = (int64_t)blockIdx.x * SSIM_PER_CTA;
  for (int64_t i = i0 + threadIdx.x; i < min(n, i0 + SSIM_PER_CTA); i += SSIM_THREADS) {
    const int c = (int)(i % W);
    const int64_t t = i / W;
    const int r = (int)(t % H);
    const int64_t plane = (t / H) * H * W;
    float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int64_t row = plane + (int64_t)reflect1(r + dy, H) * W;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int64_t j = row + reflect1(c + dx, W);
        const float a = (x[j] - lo) * inv, b = y[j];
        sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
      }
    }
    const float mx = sx / 9.f, my = sy / 9.f;
    const float vx = sxx / 9.f - mx * mx, vy = syy / 9.f - my * my, vxy = sxy / 9.f - mx * my;
    const float num = (2.f * mx * my + C1) * (2.f * vxy + C2), den = (mx * mx + my * my + C1) * (vx + vy + C2);
    acc += fminf(fmaxf((1.f - num / den) * 0.5f, 0.f), 1.f);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    part[blockIdx.x] = s;
  }
}

__global__ void sum_final_kernel(const float* __restrict__ part, int nblk, float scale, float* __restrict__ out) {
  // one thread: fixed order, double accumulation (nblk is a few thousand)
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nblk; ++i) s += (double)part[i];
    out[0] = (float)(s * (double)scale);
  }
}
}  // namespace dgtd

extern "C" {

int dgtd_boundary_weight_fwd(const float* gt, float* weit, int planes, int H, int W, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(gt && weit && planes > 0 && planes <= 65535 && H > 0 && W > 0, "boundary_weight: bad args");
  boundary_weight_kernel<<<dim3(cdiv(W, BW_T), cdiv(H, BW_T), planes), 256, 0, (cudaStream_t)stream>>>(gt, weit, H, W);
  DGTD_LAUNCH_CHECK("boundary_weight");
  return 0;
}

int dgtd_structure_loss_ws_floats(int planes, int64_t HW) { return planes * cdiv(HW, (int64_t)8192) * 4; }

// sums (planes x 4) and loss (1) are outputs; ws: dgtd_structure_loss_ws_floats floats
int dgtd_structure_loss_fwd(const float* preds, const float* gt, const float* weit, float* ws, float* sums, float* loss,
                            int planes, int64_t HW, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(preds && gt && weit && ws && sums && loss && planes > 0 && planes <= 65535 && HW > 0,
                 "structure_loss: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int nblk = cdiv(HW, (int64_t)8192);
  structure_loss_partial_kernel<<<dim3(nblk, planes), 256, 0, s>>>(preds, gt, weit, ws, HW, 8192);
  DGTD_LAUNCH_CHECK("structure_loss.partial");
  structure_loss_finalize_kernel<<<1, 256, 0, s>>>(ws, sums, loss, planes, nblk);
  DGTD_LAUNCH_CHECK("structure_loss.finalize");
  return 0;
}

int dgtd_structure_loss_bwd(const float* preds, const float* gt, const float* weit, const float* sums,
                            const float* grad_out, float* dpreds, int planes, int64_t HW, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(preds && gt && weit && sums && grad_out && dpreds && planes > 0 && HW > 0, "structure_loss_bwd: bad args");
  const int64_t total = (int64_t)planes * HW;
  structure_loss_bwd_kernel<<<(unsigned)cdiv(total, (int64_t)256), 256, 0, (cudaStream_t)stream>>>(
      preds, gt, weit, sums, grad_out, dpreds, HW, planes, total);
  DGTD_LAUNCH_CHECK("structure_loss_bwd");
  return 0;
}


int dgtd_ssim_loss_ws_floats(int64_t n) { return n > 0 ? (int)(3 * ((n + dgtd::SSIM_PER_CTA - 1) / dgtd::SSIM_PER_CTA) + 2) : 0; }

int dgtd_ssim_loss_fwd(const float* emb1, const float* image, float* ws, float* out, int planes, int H, int W,
                       dgtd_stream_t stream) {
  DGTD_CHECK_ARG(emb1 && image && ws && out, "ssim_loss: null pointer");
  DGTD_CHECK_ARG(planes > 0 && H >= 2 && W >= 2, "ssim_loss: bad shape (reflection padding needs H, W >= 2)");
  const int64_t n = (int64_t)planes * H * W;
  const int nblk = (int)((n + dgtd::SSIM_PER_CTA - 1) / dgtd::SSIM_PER_CTA);
  cudaStream_t s = (cudaStream_t)stream;
  float* part_mm = ws;                 // [nblk][2]
  float* mm = ws + 2 * (int64_t)nblk;  // [2]
  float* part_s = mm + 2;              // [nblk]
  dgtd::minmax_partial_kernel<<<nblk, dgtd::SSIM_THREADS, 0, s>>>(emb1, n, part_mm);
  DGTD_LAUNCH_CHECK("ssim_loss(minmax)");
  dgtd::minmax_final_kernel<<<1, 256, 0, s>>>(part_mm, nblk, mm);
  DGTD_LAUNCH_CHECK("ssim_loss(minmax final)");
  dgtd::ssim_partial_kernel<<<nblk, dgtd::SSIM_THREADS, 0, s>>>(emb1, image, mm, H, W, n, part_s);
  DGTD_LAUNCH_CHECK("ssim_loss(ssim)");
  dgtd::sum_final_kernel<<<1, 32, 0, s>>>(part_s, nblk, 1.0f / (float)n, out);
  DGTD_LAUNCH_CHECK("ssim_loss(sum)");
  return 0;
}

}  // extern "C"
