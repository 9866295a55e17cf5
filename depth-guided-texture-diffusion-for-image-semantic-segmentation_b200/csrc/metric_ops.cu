// MAE and S-measure of the reference's test path (twig/metric/MAE.py:18-36, Smeasure.py:18-36;
// arithmetic = pysodmetrics 1.3.1, restated in oracle/metrics_ref.py) as batched per-image reductions.
//
// HBM-bound integer/byte work: the wrappers quantise prediction and label to uint8 first, so every sum the two
// metrics need is an exact integer.  Pass 1 quantises (2 bytes per pixel kept) and reduces min / max / foreground
// count / foreground index sums per image; pass 2 accumulates, per centroid quadrant, the integer moments
// S(d), S(g), S(d^2), S(d g), S(d^2 g) of d = pred_u8 - min; a one-thread-per-image tail evaluates both metrics in
// float64 from the exact integers (variances as N*S2 - S1^2: no cancellation).  No atomics, no floating-point
// reduction: results are bit-stable and independent of the launch geometry.
#include "common.cuh"

namespace dgtd {

constexpr int MROWS = 8;     // image rows per CTA
constexpr int NS1 = 5;       // min, max, count, sum(row), sum(col)
constexpr int NS2 = 20;      // 4 quadrants x {S(d), S(g), S(d^2), S(d g), S(d^2 g)}

struct MetricWs {
  unsigned char* pu8;
  unsigned char* gu8;
  long long* part1;   // [B][chunks][NS1]
  int* info;          // [B][4] = dmin, R, x, y
  unsigned long long* part2;  // [B][chunks][NS2]
  unsigned int* part3;        // [B][chunks][512]: 256-bin histograms of the re-quantised prediction, fg | bg
};

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static MetricWs carve(void* ws, int B, int H, int W, size_t* total) {
  const size_t hw = (size_t)H * W, chunks = (size_t)cdiv(H, MROWS);
  size_t off = 0;
  MetricWs m;
  char* base = (char*)ws;
  m.pu8 = (unsigned char*)(base + off); off = align256(off + B * hw);
  m.gu8 = (unsigned char*)(base + off); off = align256(off + B * hw);
  m.part1 = (long long*)(base + off); off = align256(off + B * chunks * NS1 * sizeof(long long));
  m.info = (int*)(base + off); off = align256(off + (size_t)B * 4 * sizeof(int));
  m.part2 = (unsigned long long*)(base + off); off = align256(off + B * chunks * NS2 * sizeof(unsigned long long));
  m.part3 = (unsigned int*)(base + off); off = align256(off + B * chunks * 512 * sizeof(unsigned int));
  if (total) *total = off;
  return m;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// pass 1: `(x * 255).astype(np.uint8)` (Smeasure.py:25-26), gt > 128, per-band statistics
__global__ void __launch_bounds__(256) metric_quantise_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                              MetricWs m, int H, int W) {
  __shared__ long long red[8][NS1];
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int r0 = chunk * MROWS, r1 = min(H, r0 + MROWS);
  const int64_t img = (int64_t)b * H * W;
  int vmin = 255, vmax = 0;
  long long cnt = 0, srow = 0, scol = 0;
  for (int i = r0 * W + threadIdx.x; i < r1 * W; i += 256) {
    const float p = pred[img + i], g = gt[img + i];
    const int pu = (int)(unsigned char)(int)(__fmul_rn(p, 255.0f));     // truncation toward zero, like astype
    const int gu = (int)(unsigned char)(int)(__fmul_rn(g, 255.0f));
    const int fg = gu > 128;
    m.pu8[img + i] = (unsigned char)pu;
    m.gu8[img + i] = (unsigned char)fg;
    vmin = min(vmin, pu);
    vmax = max(vmax, pu);
    if (fg) {
      const int r = i / W;
      cnt += 1;
      srow += r;
      scol += i - r * W;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  cnt = warp_sum_ll(cnt);
  srow = warp_sum_ll(srow);
  scol = warp_sum_ll(scol);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = vmin; red[warp][1] = vmax; red[warp][2] = cnt; red[warp][3] = srow; red[warp][4] = scol;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long a = red[0][0], c = red[0][1], n = red[0][2], sr = red[0][3], sc = red[0][4];
    for (int w = 1; w < 8; ++w) {
      a = min(a, red[w][0]); c = max(c, red[w][1]); n += red[w][2]; sr += red[w][3]; sc += red[w][4];
    }
    long long* o = m.part1 + ((int64_t)b * nchunks + chunk) * NS1;
    o[0] = a; o[1] = c; o[2] = n; o[3] = sr; o[4] = sc;
  }
}

// per image: min-max normalisation constants (`_prepare_data`) and the centroid split (`Smeasure.centroid`)
__global__ void metric_info_kernel(MetricWs m, int nchunks, int B, int H, int W) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long long* p = m.part1 + (int64_t)b * nchunks * NS1;
  long long a = 255, c = 0, n = 0, sr = 0, sc = 0;
  for (int k = 0; k < nchunks; ++k) {
    a = min(a, p[k * NS1 + 0]); c = max(c, p[k * NS1 + 1]); n += p[k * NS1 + 2]; sr += p[k * NS1 + 3]; sc += p[k * NS1 + 4];
  }
  int dmin = (int)a, R = (int)(c - a);
  if (R == 0) { dmin = 0; R = 255; }          // constant prediction: pred / 255 is left un-normalised
  double x, y;
  if (n == 0) {
    x = rint((double)W / 2.0); y = rint((double)H / 2.0);     // np.round = half to even
  } else {
    y = rint((double)sr / (double)n); x = rint((double)sc / (double)n);
  }
  int* o = m.info + b * 4;
  o[0] = dmin; o[1] = R; o[2] = (int)x + 1; o[3] = (int)y + 1;
}

// pass 2: integer moments per centroid quadrant
__global__ void __launch_bounds__(256) metric_moments_kernel(MetricWs m, int H, int W) {
  __shared__ unsigned long long red[8][NS2];
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int r0 = chunk * MROWS, r1 = min(H, r0 + MROWS);
  const int64_t img = (int64_t)b * H * W;
  const int dmin = m.info[b * 4 + 0], cx = m.info[b * 4 + 2], cy = m.info[b * 4 + 3];
  unsigned int s[NS2];
#pragma unroll
  for (int k = 0; k < NS2; ++k) s[k] = 0u;
  for (int i = r0 * W + threadIdx.x; i < r1 * W; i += 256) {
    const unsigned int d = (unsigned int)((int)m.pu8[img + i] - dmin), g = m.gu8[img + i];
    const int r = i / W, c = i - r * W;
    const int q = (r >= cy ? 2 : 0) + (c >= cx ? 1 : 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned int sel = (q == k) ? 1u : 0u;
      s[k * 5 + 0] += sel * d;
      s[k * 5 + 1] += sel * g;
      s[k * 5 + 2] += sel * d * d;
      s[k * 5 + 3] += sel * d * g;
      s[k * 5 + 4] += sel * d * d * g;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NS2; ++k) {
    unsigned long long v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NS2) {
    unsigned long long v = 0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    m.part2[((int64_t)b * nchunks + chunk) * NS2 + threadIdx.x] = v;
  }
}

__device__ double ssim_from_moments(double N, double Sd, double Sg, double Sd2, double Sdg, double R) {
  // pysodmetrics Smeasure.ssim with x = pred = d / R, y = gt; unbiased (N - 1) moments from exact integers
  const double x = Sd / (R * N), y = Sg / N;
  const double den = N * (N - 1.0);
  const double sx = (N * Sd2 - Sd * Sd) / (den * R * R);
  const double sy = (N * Sg - Sg * Sg) / den;
  const double sxy = (N * Sdg - Sd * Sg) / (den * R);
  const double alpha = 4.0 * x * y * sxy;
  const double beta = (x * x + y * y) * (sx + sy);
  if (alpha != 0.0) return alpha / (beta + 2.220446049250313e-16);
  return beta == 0.0 ? 1.0 : 0.0;
}

__device__ double s_object_from_moments(double n, double S1, double S2, double R) {
  // values v = e / R with integer e: mean and std(ddof = 1)
  const double x = S1 / (R * n);
  const double var = (n * S2 - S1 * S1) / (n * (n - 1.0) * R * R);
  const double sigma = sqrt(var > 0.0 ? var : 0.0);
  return 2.0 * x / (x * x + 1.0 + sigma + 2.220446049250313e-16);
}

__global__ void metric_final_kernel(MetricWs m, int nchunks, int B, int H, int W, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double S[NS2];
  for (int k = 0; k < NS2; ++k) {
    unsigned long long v = 0;
    for (int c = 0; c < nchunks; ++c) v += m.part2[((int64_t)b * nchunks + c) * NS2 + k];
    S[k] = (double)v;      // < 2^53 for any image up to 16k x 16k
  }
  const double R = (double)m.info[b * 4 + 1];
  const int cx = m.info[b * 4 + 2], cy = m.info[b * 4 + 3];
  const double n = (double)H * (double)W;
  double Sd = 0, Sg = 0, Sd2 = 0, Sdg = 0, Sd2g = 0;
  for (int q = 0; q < 4; ++q) {
    Sd += S[q * 5 + 0]; Sg += S[q * 5 + 1]; Sd2 += S[q * 5 + 2]; Sdg += S[q * 5 + 3]; Sd2g += S[q * 5 + 4];
  }
  const double n_fg = Sg, n_bg = n - Sg;
  // MAE.step: mean |pred - gt| = (sum_bg d + sum_fg (R - d)) / (R n)
  out[b * 2 + 0] = ((Sd - Sdg) + (n_fg * R - Sdg)) / (R * n);
  double sm;
  if (n_fg == 0.0) {
    sm = 1.0 - Sd / (R * n);
  } else if (n_bg == 0.0) {
    sm = Sd / (R * n);
  } else {
    const double Nq[4] = {(double)cy * cx, (double)cy * (W - cx), (double)(H - cy) * cx, (double)(H - cy) * (W - cx)};
    const double w1 = Nq[0] / n, w2 = Nq[1] / n, w3 = Nq[2] / n, w4 = 1.0 - w1 - w2 - w3;
    const double wq[4] = {w1, w2, w3, w4};
    bool degenerate = (n_fg < 2.0) || (n_bg < 2.0);     // the library's NaN cases: max(0, nan) == 0
    double region = 0.0;
    for (int q = 0; q < 4; ++q) {
      if (Nq[q] < 2.0) { degenerate = true; continue; }
      region += wq[q] * ssim_from_moments(Nq[q], S[q * 5 + 0], S[q * 5 + 1], S[q * 5 + 2], S[q * 5 + 3], R);
    }
    if (degenerate) {
      sm = 0.0;
    } else {
      // Smeasure.object: fg values pred[gt], bg values (1 - pred)[~gt] = (R - d) / R
      const double bd = Sd - Sdg, bd2 = Sd2 - Sd2g;
      const double e1 = n_bg * R - bd, e2 = n_bg * R * R - 2.0 * R * bd + bd2;
      const double u = n_fg / n;
      const double object = u * s_object_from_moments(n_fg, Sdg, Sd2g, R) + (1.0 - u) * s_object_from_moments(n_bg, e1, e2, R);
      sm = 0.5 * object + 0.5 * region;
      sm = sm > 0.0 ? sm : 0.0;
    }
  }
  out[b * 2 + 1] = sm;
}

// F- / E-measure (pysodmetrics `Fmeasure.cal_pr`, `Emeasure.cal_em_with_cumsumhistogram`): both are functions of the
// foreground / background histograms of the min-max normalised prediction re-quantised to uint8.  The library
// does that re-quantisation in float64 -- (u/255 - min/255) / (max/255 - min/255) * 255, truncated -- and a value
// that lands on an integer can fall either side of it, so the 256-entry map is evaluated with the same IEEE
// operations in the same order (explicit round-to-nearest intrinsics: no contraction).
__global__ void __launch_bounds__(256) metric_hist_kernel(MetricWs m, int H, int W) {
  __shared__ unsigned int hist[512];
  __shared__ unsigned char lut[256];
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int dmin = m.info[b * 4 + 0], R = m.info[b * 4 + 1];
  {
    const int u = threadIdx.x;
    const double p0 = __ddiv_rn((double)u, 255.0);
    const double mn = __ddiv_rn((double)dmin, 255.0), mx = __ddiv_rn((double)(dmin + R), 255.0);
    // dmin = 0, R = 255 is both the constant-prediction convention of metric_info_kernel (no normalisation) and
    // the full-range case, where (p0 - 0) / (1 - 0) == p0 exactly
    const double p = (dmin == 0 && R == 255) ? p0 : __ddiv_rn(__dsub_rn(p0, mn), __dsub_rn(mx, mn));
    const int q = (u >= dmin && u <= dmin + R) ? (int)__dmul_rn(p, 255.0) : 0;
    lut[u] = (unsigned char)q;
    hist[u] = 0u;
    hist[256 + u] = 0u;
  }
  __syncthreads();
  const int r0 = chunk * MROWS, r1 = min(H, r0 + MROWS);
  const int64_t img = (int64_t)b * H * W;
  for (int i = r0 * W + threadIdx.x; i < r1 * W; i += 256)
    atomicAdd(&hist[(m.gu8[img + i] ? 0 : 256) + lut[m.pu8[img + i]]], 1u);   // integer: exact in any order
  __syncthreads();
  unsigned int* o = m.part3 + ((int64_t)b * nchunks + chunk) * 512;
  o[threadIdx.x] = hist[threadIdx.x];
  o[256 + threadIdx.x] = hist[256 + threadIdx.x];
}

// curves[b][0][i] = changeable F-measure, curves[b][1][i] = E-measure at threshold 255 - i (the library's order)
__global__ void __launch_bounds__(256) metric_curves_kernel(MetricWs m, int nchunks, int H, int W, double* __restrict__ curves) {
  __shared__ double cf[256], cb[256];
  const int b = blockIdx.x, t = threadIdx.x;
  unsigned long long f = 0, g = 0;
  for (int c = 0; c < nchunks; ++c) {
    const unsigned int* o = m.part3 + ((int64_t)b * nchunks + c) * 512;
    f += o[255 - t];          // flipped: entry t holds bin 255 - t
    g += o[256 + 255 - t];
  }
  cf[t] = (double)f;
  cb[t] = (double)g;
  __syncthreads();
  if (t == 0) {               // cumulative sums (exact: integers below 2^53)
    for (int i = 1; i < 256; ++i) { cf[i] += cf[i - 1]; cb[i] += cb[i - 1]; }
  }
  __syncthreads();
  const double eps = 2.220446049250313e-16;
  const double size = (double)H * (double)W, gt_fg = cf[255];
  const double fg_fg = cf[t], fg_bg = cb[t];
  // Fmeasure.cal_pr, beta^2 = 0.3
  {
    double ps = fg_fg + fg_bg;
    if (ps == 0.0) ps = 1.0;
    const double T = gt_fg > 1.0 ? gt_fg : 1.0;
    const double prec = fg_fg / ps, rec = fg_fg / T;
    const double num = 1.3 * prec * rec;
    const double den = num == 0.0 ? 1.0 : 0.3 * prec + rec;
    curves[((int64_t)b * 2 + 0) * 256 + t] = num / den;
  }
  // Emeasure.cal_em_with_cumsumhistogram
  {
    const double pred_fg = fg_fg + fg_bg, pred_bg = size - pred_fg;
    double total;
    if (gt_fg == 0.0) {
      total = pred_bg;
    } else if (gt_fg == size) {
      total = pred_fg;
    } else {
      const double bg_fg = gt_fg - fg_fg, bg_bg = pred_bg - bg_fg;
      const double mp = pred_fg / size, mg = gt_fg / size;
      const double a[2] = {1.0 - mp, 0.0 - mp}, c[2] = {1.0 - mg, 0.0 - mg};
      const double numel[4] = {fg_fg, fg_bg, bg_fg, bg_bg};
      total = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double x = a[i >> 1], y = c[i & 1];
        const double align = 2.0 * (x * y) / (x * x + y * y + eps);
        total += (align + 1.0) * (align + 1.0) / 4.0 * numel[i];
      }
    }
    curves[((int64_t)b * 2 + 1) * 256 + t] = total / (size - 1.0 + eps);
  }
}

}  // namespace dgtd
using namespace dgtd;

extern "C" {

int64_t dgtd_sod_metrics_ws_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  size_t total = 0;
  carve(nullptr, B, H, W, &total);
  return (int64_t)total;
}

int dgtd_sod_metrics_fwd(const float* pred, const float* gt, void* ws, double* out, double* curves, int B, int H,
                         int W, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(pred && gt && ws && out, "sod_metrics: null pointer");
  DGTD_CHECK_ARG(B > 0 && H > 1 && W > 1 && (int64_t)H * W < (1ll << 28), "sod_metrics: bad shape");
  DGTD_CHECK_ARG(((uintptr_t)ws & 255) == 0, "sod_metrics: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  MetricWs m = carve(ws, B, H, W, nullptr);
  const int nchunks = cdiv(H, MROWS);
  metric_quantise_kernel<<<dim3(nchunks, B), 256, 0, s>>>(pred, gt, m, H, W);
  DGTD_LAUNCH_CHECK("sod_metrics(quantise)");
  metric_info_kernel<<<cdiv(B, 32), 32, 0, s>>>(m, nchunks, B, H, W);
  DGTD_LAUNCH_CHECK("sod_metrics(info)");
  metric_moments_kernel<<<dim3(nchunks, B), 256, 0, s>>>(m, H, W);
  DGTD_LAUNCH_CHECK("sod_metrics(moments)");
  metric_final_kernel<<<cdiv(B, 32), 32, 0, s>>>(m, nchunks, B, H, W, out);
  DGTD_LAUNCH_CHECK("sod_metrics(final)");
  if (curves) {
    metric_hist_kernel<<<dim3(nchunks, B), 256, 0, s>>>(m, H, W);
    DGTD_LAUNCH_CHECK("sod_metrics(hist)");
    metric_curves_kernel<<<B, 256, 0, s>>>(m, nchunks, H, W, curves);
    DGTD_LAUNCH_CHECK("sod_metrics(curves)");
  }
  return 0;
}

}  // extern "C"
