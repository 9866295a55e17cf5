// a6 at microbench scale, weight mode W2: MessagePassing.forward core
// (cod.py:1190-1205) with the per-channel propagation weights of the model (cod.py:1296,
// ShapePropWeightRegressor :1051-1060) generated ON CHIP:
//     W[c,k,p] = sigmoid(Wr[c*49+k, :] . guide[:, p] + br[c*49+k]),   Wn = W / (sum_k W + eps)
//     y[c,p]   = sum_k Wn[c,k,p] * x[c, p + delta_k]                   (zero padding, no border renorm)
// Materialising W at 1024^2 x 256 would be 49 GiB (fp32); here the HBM traffic per iteration is
// x in, y out and the 3-channel guide:  (2*C*sizeof(T) + 3*4) * H * W bytes.
//
// CTA = 8 x 16 pixels, looping over chunks of 32 channels; lane = channel, warp = 2 x 8 pixel sub-tile.
//   * x halo tile (14 x 22 pixels x 32 channels) per chunk by one 4-D TMA box, double buffered;
//   * the regressor rows of the chunk come from a packed [chunk][tap][lane] float4 table (coalesced
//     512-byte loads, L1/L2 resident: 200 KB for C = 256);
//   * per tap: 4 parameter registers, then for each of the warp's 16 pixels 3 FMA + sigmoid + 2 FMA
//     + 1 conflict-free shared load; only the 3 guide values per pixel are warp-uniform (24 broadcast
//     loads per 8 pixels x 49 taps).
// The kernel is bound by the MUFU pipe (49 sigmoids per output element: ex2 + rcp, or one tanh in the
// fast mode), not by HBM; T iterations regenerate the weights T times (caching them would cost the
// same 26 GB of traffic per iteration as recomputing).
#include "blackwell.cuh"
#include "common.cuh"

namespace dgtd {

constexpr int MR_TH = 8, MR_TW = 16, MR_PH = MR_TH + 6, MR_PW = MR_TW + 6;

template <bool FAST>
__device__ __forceinline__ float sigmoid_dev(float z) {
  if (FAST) {   // 0.5 tanh(z/2) + 0.5: one MUFU, |err| <= 2.5e-4
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    return fmaf(t, 0.5f, 0.5f);
  }
  return __fdividef(1.0f, 1.0f + __expf(-z));   // ex2.approx + rcp.approx: ~2 ulp
}

template <typename T, bool FAST>
__global__ void __launch_bounds__(256, 2)
mp_regress_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ guide,
                  const float4* __restrict__ packed, T* __restrict__ out, int h, int w, int C, float eps) {
  constexpr int TILE_ELEMS = MR_PH * MR_PW * 32;
  constexpr uint32_t TILE_BYTES = TILE_ELEMS * sizeof(T);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar[2];

  const int tiles_x = (w + MR_TW - 1) / MR_TW, tiles_y = (h + MR_TH - 1) / MR_TH;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int n = bid / tiles_y;
  const int x0 = tx * MR_TW, y0 = ty * MR_TH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sy = warp >> 1, sx = warp & 1;   // 4 x 2 sub-tiles of 2 x 8 pixels
  const int nchunks = C >> 5;

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
  // guide values of the tile (3 x 128 floats), read back as warp-uniform broadcasts per half sub-tile
  __shared__ float gs[3][MR_TH * MR_TW];
  if (threadIdx.x < MR_TH * MR_TW) {
    const int gy = min(y0 + (threadIdx.x >> 4), h - 1), gx = min(x0 + (threadIdx.x & 15), w - 1);
    const float* gp = guide + ((int64_t)n * 3 * h + gy) * w + gx;
    gs[0][threadIdx.x] = __ldg(gp);
    gs[1][threadIdx.x] = __ldg(gp + (int64_t)h * w);
    gs[2][threadIdx.x] = __ldg(gp + 2 * (int64_t)h * w);
  }
  __syncthreads();

#pragma unroll 1
  for (int i = 0; i < nchunks; ++i) {
    if (threadIdx.x == 0 && i + 1 < nchunks) {   // buffer (i+1)&1 was released by the barrier below
      bw::fence_proxy_async_smem();
      bw::mbar_arrive_expect_tx(&bar[(i + 1) & 1], TILE_BYTES);
      bw::tma_load_4d(&tmX, &bar[(i + 1) & 1], xs + ((i + 1) & 1) * TILE_ELEMS, (i + 1) * 32, x0 - 3, y0 - 3, n);
    }
    const float4* prm = packed + (int64_t)i * 49 * 32 + lane;
    const int c = i * 32 + lane;
    bw::mbar_wait(&bar[i & 1], (i >> 1) & 1);
#pragma unroll 1
    for (int a = 0; a < 2; ++a) {   // the two pixel rows of the warp's 2 x 8 sub-tile, 8 pixels at a time
      const int prow = 2 * sy + a;
      float g0[8], g1[8], g2[8], acc[8], sum[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        g0[p] = gs[0][prow * MR_TW + 8 * sx + p];
        g1[p] = gs[1][prow * MR_TW + 8 * sx + p];
        g2[p] = gs[2][prow * MR_TW + 8 * sx + p];
        acc[p] = sum[p] = 0.f;
      }
      const T* base = xs + (i & 1) * TILE_ELEMS + (prow * MR_PW + 8 * sx) * 32 + lane;
#pragma unroll 1
      for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float4 pr = __ldg(prm + (ky * 7 + kx) * 32);
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const float z = fmaf(pr.x, g0[p], fmaf(pr.y, g1[p], fmaf(pr.z, g2[p], pr.w)));
            const float s = sigmoid_dev<FAST>(z);
            sum[p] += s;
            acc[p] = fmaf(s, to_float(base[(ky * MR_PW + p + kx) * 32]), acc[p]);
          }
        }
      }
      const int oy = y0 + prow;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int ox = x0 + 8 * sx + p;
        if (oy < h && ox < w) store1(out + (((int64_t)n * h + oy) * w + ox) * C + c, acc[p] / (sum[p] + eps));
      }
    }
    __syncthreads();   // every warp is done with buffer i&1 before it is refilled
  }
}

// packed[(chunk*49 + k)*32 + lane] = (Wr[c*49+k][0..2], br[c*49+k]),  c = chunk*32 + lane
__global__ void pack_regressor_kernel(const float* __restrict__ wr, const float* __restrict__ br,
                                      float4* __restrict__ packed, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * 49) return;
  const int lane = i & 31, k = (i >> 5) % 49, chunk = (i >> 5) / 49;
  const int row = (chunk * 32 + lane) * 49 + k;
  packed[i] = make_float4(wr[row * 3], wr[row * 3 + 1], wr[row * 3 + 2], br[row]);
}

template <typename T, bool FAST>
static int mr_step(const void* x, const float* guide, const float4* packed, void* out, int n, int h, int w, int c,
                   float eps, cudaStream_t s) {
  CUtensorMap tm;
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return -3;
  const size_t es = sizeof(T);
  cuuint64_t gd[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gs[3] = {(cuuint64_t)c * es, (cuuint64_t)w * c * es, (cuuint64_t)h * w * c * es};
  cuuint32_t bx[4] = {32, MR_PW, MR_PH, 1};
  cuuint32_t est[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                   const_cast<void*>(x), gd, gs, bx, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("message_passing_regress: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -3;
  }
  const int smem = 2 * MR_PH * MR_PW * 32 * (int)es + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mp_regress_kernel<T, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("message_passing_regress: cannot opt in to %d B smem: %s", smem, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  const int64_t blocks = (int64_t)n * cdiv(h, MR_TH) * cdiv(w, MR_TW);
  mp_regress_kernel<T, FAST><<<(unsigned)blocks, 256, smem, s>>>(tm, guide, packed, (T*)out, h, w, c, eps);
  return 0;
}

}  // namespace dgtd

using namespace dgtd;

extern "C" {

int dgtd_pack_regressor(const float* wr, const float* br, float* packed, int c, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(wr && br && packed && c > 0 && c % 32 == 0, "pack_regressor: c must be a multiple of 32");
  DGTD_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "pack_regressor: packed must be 16-byte aligned");
  pack_regressor_kernel<<<cdiv(c * 49, 256), 256, 0, (cudaStream_t)stream>>>(wr, br, reinterpret_cast<float4*>(packed), c);
  DGTD_LAUNCH_CHECK("pack_regressor");
  return 0;
}

int dgtd_message_passing_regress_fwd(const void* x, const float* guide, const float* packed, void* out, void* tmp,
                                     int n, int h, int w, int c, int T, float eps, int dtype, int fast_sigmoid,
                                     dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && guide && packed && out, "message_passing_regress: null pointer");
  DGTD_CHECK_ARG(dtype == DGTD_F32 || dtype == DGTD_BF16, "message_passing_regress: bad dtype %d", dtype);
  DGTD_CHECK_ARG(n > 0 && h > 0 && w > 0 && c >= 32 && c % 32 == 0 && T >= 1,
                 "message_passing_regress: bad shape n=%d h=%d w=%d c=%d T=%d (c must be a multiple of 32)", n, h, w, c, T);
  DGTD_CHECK_ARG(T == 1 || tmp, "message_passing_regress: tmp buffer required for T > 1");
  DGTD_CHECK_ARG(!(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15) &&
                     !(reinterpret_cast<uintptr_t>(packed) & 15),
                 "message_passing_regress: buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const float4* pk = reinterpret_cast<const float4*>(packed);
  const void* src = x;   // ping-pong so that the last step lands in `out`
  for (int t = 0; t < T; ++t) {
    void* dst = ((T - 1 - t) % 2 == 0) ? out : tmp;
    int rc;
    if (dtype == DGTD_F32)
      rc = fast_sigmoid ? mr_step<float, true>(src, guide, pk, dst, n, h, w, c, eps, s)
                        : mr_step<float, false>(src, guide, pk, dst, n, h, w, c, eps, s);
    else
      rc = fast_sigmoid ? mr_step<__nv_bfloat16, true>(src, guide, pk, dst, n, h, w, c, eps, s)
                        : mr_step<__nv_bfloat16, false>(src, guide, pk, dst, n, h, w, c, eps, s);
    if (rc) return rc;
    DGTD_LAUNCH_CHECK("message_passing_regress");
    src = dst;
  }
  return 0;
}

}  // extern "C"
