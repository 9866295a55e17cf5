// a6 at microbench scale (BASELINE configs[3]: 1024 x 1024 x 256, k = 7, T sweep):
// MessagePassing.forward core (cod.py:1190-1205) with shared weights (wc == 1) on NHWC maps.
//
// One launch per diffusion step (at C = 256 a step is already at the crossover between the HBM
// and the fp32-FMA roof, so fusing steps would only add halo recomputation).  Per CTA:
//   * 8 x 32 output pixels; the 49 raw weights of each pixel are read once (coalesced planes),
//     random-walk normalised (W / (sum W + eps), cod.py:1201) and kept in shared memory for ALL
//     channel chunks -- the reference re-reads a 49x unfolded copy of x instead;
//   * x streams through in 32-channel chunks: one 4-D TMA box (14 x 38 pixels x 128 B, zero
//     fill outside the map == unfold's zero padding) per chunk, double buffered;
//   * warp = 4 x 8 pixel sub-tile, lane = channel: each input row is read once from shared
//     memory (conflict-free) and feeds up to 4 output rows; weights come as broadcast 16-byte
//     shared loads; 32 accumulators per thread.
// Algorithmic HBM traffic per step: (2*C + 49) * H * W * 4 bytes.
#include "blackwell.cuh"
#include "common.cuh"

namespace dgtd {

constexpr int MP_TH = 8, MP_TW = 32, MP_PH = MP_TH + 6, MP_PW = MP_TW + 6;
constexpr int MP_TILE_FLOATS = MP_PH * MP_PW * 32;
constexpr int MP_WN_FLOATS = MP_TH * MP_TW * 56;   // [pixel][ky][8] (7 taps + 1 pad)
constexpr int MP_SMEM = (2 * MP_TILE_FLOATS + MP_WN_FLOATS) * 4 + 128;

template <typename OT>
__global__ void __launch_bounds__(256, 1)
mp_tiled_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ weight, OT* __restrict__ out,
                int h, int w, int C, float eps) {
  extern __shared__ uint8_t smem_raw[];
  float* xs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  float* wn = xs + 2 * MP_TILE_FLOATS;
  __shared__ uint64_t bar[2];

  const int tiles_x = (w + MP_TW - 1) / MP_TW, tiles_y = (h + MP_TH - 1) / MP_TH;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int n = bid / tiles_y;
  const int x0 = tx * MP_TW, y0 = ty * MP_TH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sy = warp >> 2, sx = warp & 3;
  const int nchunks = C >> 5;
  constexpr uint32_t TILE_BYTES = MP_TILE_FLOATS * 4;

  if (threadIdx.x == 0) {
    bw::prefetch_tmap(&tmX);
    bw::mbar_init(&bar[0], 1);
    bw::mbar_init(&bar[1], 1);
    bw::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    bw::mbar_arrive_expect_tx(&bar[0], TILE_BYTES);
    bw::tma_load_4d(&tmX, &bar[0], xs, 0, x0 - 3, y0 - 3, n);
  }
  {  // normalised weights of this tile: thread = pixel (row-major 8 x 32)
    const int py = threadIdx.x >> 5, px = threadIdx.x & 31;
    const int gy = y0 + py, gx = x0 + px;
    float wv[49];
    float s = 0.f;
    const bool ok = gy < h && gx < w;
    const float* wp = weight + ((int64_t)n * 49 * h + gy) * w + gx;
#pragma unroll
    for (int k = 0; k < 49; ++k) {
      wv[k] = ok ? __ldg(wp + (int64_t)k * h * w) : 0.f;
      s += wv[k];
    }
    const float inv = 1.0f / (s + eps);
    float* dst = wn + threadIdx.x * 56;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) dst[ky * 8 + kx] = wv[ky * 7 + kx] * inv;
      dst[ky * 8 + 7] = 0.f;
    }
  }
  __syncthreads();

  const float* wbase = wn + ((4 * sy) * MP_TW + 8 * sx) * 56;
#pragma unroll 1
  for (int i = 0; i < nchunks; ++i) {
    if (threadIdx.x == 0 && i + 1 < nchunks) {
      bw::fence_proxy_async_smem();
      bw::mbar_arrive_expect_tx(&bar[(i + 1) & 1], TILE_BYTES);
      bw::tma_load_4d(&tmX, &bar[(i + 1) & 1], xs + ((i + 1) & 1) * MP_TILE_FLOATS, (i + 1) * 32, x0 - 3, y0 - 3, n);
    }
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
    bw::mbar_wait(&bar[i & 1], (i >> 1) & 1);
    const float* base = xs + (i & 1) * MP_TILE_FLOATS + ((4 * sy) * MP_PW + 8 * sx) * 32 + lane;
#pragma unroll
    for (int iy = 0; iy < 10; ++iy) {
      float in[14];
#pragma unroll
      for (int j = 0; j < 14; ++j) in[j] = base[(iy * MP_PW + j) * 32];
#pragma unroll
      for (int oy = 0; oy < 4; ++oy) {
        const int ky = iy - oy;
        if (ky < 0 || ky >= 7) continue;
#pragma unroll
        for (int ox = 0; ox < 8; ++ox) {
          const float4 wa = *reinterpret_cast<const float4*>(wbase + (oy * MP_TW + ox) * 56 + ky * 8);
          const float4 wb = *reinterpret_cast<const float4*>(wbase + (oy * MP_TW + ox) * 56 + ky * 8 + 4);
          float a = acc[oy][ox];
          a = fmaf(wa.x, in[ox + 0], a);
          a = fmaf(wa.y, in[ox + 1], a);
          a = fmaf(wa.z, in[ox + 2], a);
          a = fmaf(wa.w, in[ox + 3], a);
          a = fmaf(wb.x, in[ox + 4], a);
          a = fmaf(wb.y, in[ox + 5], a);
          a = fmaf(wb.z, in[ox + 6], a);
          acc[oy][ox] = a;
        }
      }
      asm volatile("" ::: "memory");
    }
    const int c = i * 32 + lane;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int oy = y0 + 4 * sy + a;
      if (oy >= h) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ox = x0 + 8 * sx + j;
        if (ox < w) out[(((int64_t)n * h + oy) * w + ox) * C + c] = from_float<OT>(acc[a][j]);
      }
    }
    __syncthreads();
  }
}

static int mp_step(const float* x, const float* weight, float* out, int n, int h, int w, int c, float eps,
                   cudaStream_t s) {
  CUtensorMap tm;
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return -3;
  cuuint64_t gd[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gs[3] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 4};
  cuuint32_t bx[4] = {32, MP_PW, MP_PH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("message_passing_tiled: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -3;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mp_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, MP_SMEM);
    if (e != cudaSuccess) {
      set_error("message_passing_tiled: cannot opt in to %d B smem: %s", MP_SMEM, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  const int64_t blocks = (int64_t)n * cdiv(h, MP_TH) * cdiv(w, MP_TW);
  mp_tiled_kernel<float><<<(unsigned)blocks, 256, MP_SMEM, s>>>(tm, weight, out, h, w, c, eps);
  return 0;
}

}  // namespace dgtd

using namespace dgtd;

extern "C" int dgtd_message_passing_tiled_fwd(const void* x, const float* weight, void* out, void* tmp, int n,
                                              int h, int w, int c, int T, float eps, int dtype,
                                              dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && weight && out, "message_passing_tiled: null pointer");
  DGTD_CHECK_ARG(n > 0 && h > 0 && w > 0 && c >= 32 && c % 32 == 0 && T >= 1,
                 "message_passing_tiled: bad shape n=%d h=%d w=%d c=%d T=%d (c must be a multiple of 32)", n, h, w, c, T);
  DGTD_CHECK_ARG(dtype == DGTD_F32, "message_passing_tiled: fp32 storage only in this build");
  DGTD_CHECK_ARG(T == 1 || tmp, "message_passing_tiled: tmp buffer required for T > 1");
  DGTD_CHECK_ARG(!(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15),
                 "message_passing_tiled: buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  // ping-pong so that the last step lands in `out`
  const float* src = (const float*)x;
  for (int t = 0; t < T; ++t) {
    float* dst = ((T - 1 - t) % 2 == 0) ? (float*)out : (float*)tmp;
    int rc = mp_step(src, weight, dst, n, h, w, c, eps, s);
    if (rc) return rc;
    DGTD_LAUNCH_CHECK("message_passing_tiled");
    src = dst;
  }
  return 0;
}
