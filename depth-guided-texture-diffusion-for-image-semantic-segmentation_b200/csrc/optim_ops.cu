// Optimizer step of the training configuration (config/sod.yml:56-76: AdamW lr 5e-4, weight_decay 0.1, per-prefix
// lr multipliers through paramwise_cfg.custom_keys) as ONE launch over the flat parameter / gradient / moment
// buffers of the hot path (twig/graphs.py keeps every .grad as a view of one flat fp32 buffer).
//
// torch.optim.AdamW semantics (the optimizer mmengine's OptimWrapper constructs), per element:
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// HBM-bound: 16 bytes read + 12 bytes written per parameter, float4 accesses.  The per-parameter lr / weight decay
// come from a block table (one entry per <= 4096-element slice of one parameter).  The flat layout
// (twig/flat.py) starts every parameter on a 128-byte boundary -- the other kernels read parameters with float4 /
// TMA accesses -- and the pad elements belong to no slice.
#include "common.cuh"

namespace dgtd {

struct AdamwSlice {   // 24 bytes
  long long off;      // first element of the slice in the flat buffers
  int n;              // elements (<= 4096)
  float lr, wd;
  int pad;
};

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, const AdamwSlice* __restrict__ table,
                                                    float b1, float b2, float eps, float inv_bc1, float inv_sqrt_bc2,
                                                    float grad_scale, float lr_scale) {
  const AdamwSlice s = table[blockIdx.x];
  const float lr = s.lr * lr_scale;   // lr_scale: the scheduler's factor (CosineAnnealingLR, config/sod.yml param_scheduler)
  const float decay = 1.0f - lr * s.wd, step = lr * inv_bc1;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= grad_scale;
    pp *= decay;
    mm = b1 * mm + (1.0f - b1) * gg;
    vv = b2 * vv + (1.0f - b2) * gg * gg;
    pp -= step * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
  };
  const long long o = s.off;
  // head up to 16-byte alignment, float4 body, scalar tail (slices start anywhere in the flat buffer)
  const int head = min(s.n, (int)((4 - (o & 3)) & 3));
  if (threadIdx.x < head) {
    const long long i = o + threadIdx.x;
    float pp = p[i], mm = m[i], vv = v[i];
    upd(pp, g[i], mm, vv);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
  const int body = (s.n - head) >> 2;
  for (int q = threadIdx.x; q < body; q += 256) {
    const long long i = o + head + 4ll * q;
    float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    const float4 gg = *reinterpret_cast<const float4*>(g + i);
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    *reinterpret_cast<float4*>(p + i) = pp; *reinterpret_cast<float4*>(m + i) = mm; *reinterpret_cast<float4*>(v + i) = vv;
  }
  const int tail0 = head + 4 * body;
  if (threadIdx.x < s.n - tail0) {
    const long long i = o + tail0 + threadIdx.x;
    float pp = p[i], mm = m[i], vv = v[i];
    upd(pp, g[i], mm, vv);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

}  // namespace dgtd
using namespace dgtd;

extern "C" {

int dgtd_adamw_slice_bytes(void) { return (int)sizeof(AdamwSlice); }

int dgtd_adamw_step(float* p, const float* g, float* m, float* v, const void* table, int nslices, float beta1, float beta2,
                    float eps, int step, float grad_scale, float lr_scale, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(p && g && m && v && table && nslices > 0 && step >= 1, "adamw_step: bad arguments");
  DGTD_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                   reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                 "adamw_step: flat buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  adamw_kernel<<<nslices, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (const AdamwSlice*)table, beta1, beta2, eps,
                                                          (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)), grad_scale, lr_scale);
  DGTD_LAUNCH_CHECK("adamw_step");
  return 0;
}

}  // extern "C"
