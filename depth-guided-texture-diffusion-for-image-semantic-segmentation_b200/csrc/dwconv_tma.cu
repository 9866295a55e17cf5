// a8 (cod.py:1106-1108), large-batch path: depthwise 7x7 on NHWC fp32 with TMA-staged halo
// tiles, followed by a row LayerNorm kernel.
//
// K1 dwconv7_tma_kernel: CTA = (4*SY x 8*SX) output pixels x 128 channels.  The (TH+6)x(TW+6)
//    halo tile of one 32-channel chunk (128 B per pixel) is fetched by ONE 4-D TMA box load
//    (zero fill outside the image == conv padding), double buffered across the 4 chunks.
//    warp = one 4x8 pixel sub-tile, lane = channel: 49 taps + 32 accumulators in registers, every
//    input row is read once from shared memory (conflict-free 128 B per warp) and feeds up to
//    4 output rows x 7 taps (LDS : FMA = 1 : 11).  Writes y (pre-norm) fp32, coalesced.
// K2 ln_rows_kernel: one warp per pixel, two-pass statistics in registers, writes bf16 / fp32.
//    y is consumed straight out of L2 for the 24x24 and 12x12 stages (75 MB / 19 MB at B=64).
#include "blackwell.cuh"
#include "common.cuh"

namespace dgtd {

int sm_count();   // tc_gemm.cu
static inline int sm_count_dw() { return sm_count(); }

template <int SY, int SX, bool ADD, typename OT = float>
__global__ void __launch_bounds__(SY * SX * 32, 2)
dwconv7_tma_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ wT,
                   const float* __restrict__ bias, const float* __restrict__ add, OT* __restrict__ y, int h,
                   int w, int C, int tiles_x, int tiles_y) {
  constexpr int TH = 4 * SY, TW = 8 * SX, PH = TH + 6, PW = TW + 6, NCH = 4;
  constexpr int TILE_FLOATS = PH * PW * 32;
  constexpr uint32_t TILE_BYTES = TILE_FLOATS * 4;
  extern __shared__ __align__(128) float xs[];   // TMA destination; keeps ld.shared addressing
  __shared__ uint64_t bar[2];

  int bid = blockIdx.x;
  const int cgs = C >> 7;
  const int cg = bid % cgs; bid /= cgs;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sy = warp / SX, sx = warp - sy * SX;
  const int x0 = tx * TW, y0 = ty * TH;

  if (threadIdx.x == 0) {
    bw::prefetch_tmap(&tmX);
    bw::mbar_init(&bar[0], 1);
    bw::mbar_init(&bar[1], 1);
    bw::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    bw::mbar_arrive_expect_tx(&bar[0], TILE_BYTES);
    bw::tma_load_4d(&tmX, &bar[0], xs, cg * 128, x0 - 3, y0 - 3, b);
  }

#pragma unroll 1
  for (int i = 0; i < NCH; ++i) {
    const int c = cg * 128 + i * 32 + lane;
    if (threadIdx.x == 0 && i + 1 < NCH) {   // buffer (i+1)&1 was released by the barrier below
      bw::fence_proxy_async_smem();
      bw::mbar_arrive_expect_tx(&bar[(i + 1) & 1], TILE_BYTES);
      bw::tma_load_4d(&tmX, &bar[(i + 1) & 1], xs + ((i + 1) & 1) * TILE_FLOATS, cg * 128 + (i + 1) * 32, x0 - 3,
                      y0 - 3, b);
    }
    // taps duplicated into both halves of a 64-bit register: fma.rn.f32x2 then updates two
    // vertically adjacent output pixels per instruction (half the FMA issue slots)
    uint64_t w2[49];
#pragma unroll
    for (int k = 0; k < 49; ++k) {
      const float t = __ldg(wT + (int64_t)k * C + c);
      w2[k] = pk2(t, t);
    }
    const float bc = bias ? __ldg(bias + c) : 0.f;
    // acc[q][j] = output pixels (row 2q, col j) and (row 2q+1, col j): the two halves of one f32x2 register.
    // Pairing ROWS (not columns) means the matching input pair (in[iy][j+kx], in[iy+1][j+kx]) is two plain
    // shared loads into the halves of a register pair for every kx -- no re-packing moves.
    uint64_t acc[2][8];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = pk2(bc, bc);

    bw::mbar_wait(&bar[i & 1], (i >> 1) & 1);
    const float* base = xs + (i & 1) * TILE_FLOATS + ((4 * sy) * PW + 8 * sx) * 32 + lane;
#pragma unroll
    for (int iy = 0; iy < 9; ++iy) {
      uint64_t pr[14];   // (input row iy, input row iy+1) at the 14 columns of the sub-tile's halo
#pragma unroll
      for (int j = 0; j < 14; ++j) {   // volatile: two fresh loads per pair (the compiler would otherwise reuse row
        // iy+1 from the previous iteration and pay two moves per pair to re-pack it)
        const volatile float* p0 = base + (iy * PW + j) * 32;
        pr[j] = pk2(p0[0], p0[PW * 32]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int ky = iy - 2 * q;
        if (ky < 0 || ky >= 7) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[q][j] = fma2(w2[ky * 7 + kx], pr[j + kx], acc[q][j]);
      }
    }
    if (ADD) {
      // residual operand: all 32 loads issued before the first dependent store (one round trip, not 32)
      float av[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int oy = y0 + 4 * sy + a, ox = x0 + 8 * sx + j;
          av[a][j] = (oy < h && ox < w) ? __ldg(add + (((int64_t)b * h + oy) * w + ox) * C + c) : 0.f;
        }
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[q][j] = add2(acc[q][j], pk2(av[2 * q][j], av[2 * q + 1][j]));
    }
    {
      const int oy0 = y0 + 4 * sy, ox0 = x0 + 8 * sx;
      OT* yp = y + (((int64_t)b * h + oy0) * w + ox0) * C + c;
      const int rs = w * C;   // one image row; a sub-tile spans < 2^31 elements
      if (oy0 + 4 <= h && ox0 + 8 <= w) {   // interior sub-tile (warp-uniform): no per-store predicates
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v0, v1;
            up2(acc[q][j], v0, v1);
            store1(yp + (2 * q) * rs + j * C, v0);
            store1(yp + (2 * q + 1) * rs + j * C, v1);
          }
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v0, v1;
            up2(acc[q][j], v0, v1);
            if (ox0 + j < w) {
              if (oy0 + 2 * q < h) store1(yp + (2 * q) * rs + j * C, v0);
              if (oy0 + 2 * q + 1 < h) store1(yp + (2 * q + 1) * rs + j * C, v1);
            }
          }
      }
    }
    __syncthreads();   // every warp is done with buffer i&1 before it is refilled
  }
}


// Persistent form of the kernel above (round 2).  The one-shot kernel ran the FMA pipe at 44 %: its grid is a fixed
// number of (tile, 128-channel group) CTAs -- 768 at stage 2 on 296 CTA slots = 2.6 waves, the last one 60 % full -- and
// every CTA pays its own barrier setup, first-tile TMA latency and, per 32-channel chunk, 49 global weight loads in front
// of the first FMA.  Here 2 CTAs per SM walk (tile, 32-channel chunk) ITEMS round-robin (3072 at stage 2: 10.4 per
// CTA, a quantisation loss of 5 % instead of 13 %), the next item's halo tile AND its 49 x 32 weights arrive by TMA while
// the current item computes (the weights box lands in a single 6 KB buffer that is free again as soon as every warp has
// copied its taps into registers), and nothing but the bias is read with a plain global load -- issued after the
// proxy fence of the iteration, whose membar would otherwise wait for it.
template <int SY, int SX, bool ADD, typename OT>
__global__ void __launch_bounds__(SY * SX * 32, 2)
dwconv7_tma_persist_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                           const float* __restrict__ bias, const float* __restrict__ add, OT* __restrict__ y, int h,
                           int w, int C, int tiles_x, int tiles_y, int num_items) {
  constexpr int TH = 4 * SY, TW = 8 * SX, PH = TH + 6, PW = TW + 6;
  constexpr int TILE_FLOATS = PH * PW * 32;
  constexpr uint32_t TILE_BYTES = TILE_FLOATS * 4, W_BYTES = 49 * 32 * 4;
  extern __shared__ __align__(128) float xs[];   // [2][TILE_FLOATS] halo tiles, then [49][32] weights
  float* wbuf = xs + 2 * TILE_FLOATS;
  __shared__ uint64_t bar[3];                    // tile buffer 0 / 1, weights

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sy = warp / SX, sx = warp - sy * SX;
  const int nch = C >> 5;
  auto issue_tile = [&](int item, int buf) {
    const int chunk = item % nch;
    int t = item / nch;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    bw::mbar_arrive_expect_tx(&bar[buf], TILE_BYTES);
    bw::tma_load_4d(&tmX, &bar[buf], xs + buf * TILE_FLOATS, chunk * 32, tx * TW - 3, ty * TH - 3, b);
  };
  auto issue_weights = [&](int item) {
    bw::mbar_arrive_expect_tx(&bar[2], W_BYTES);
    bw::tma_load_2d(&tmW, &bar[2], wbuf, (item % nch) * 32, 0);
  };
  if (threadIdx.x == 0) {
    bw::prefetch_tmap(&tmX);
    bw::prefetch_tmap(&tmW);
    bw::mbar_init(&bar[0], 1);
    bw::mbar_init(&bar[1], 1);
    bw::mbar_init(&bar[2], 1);
    bw::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0 && (int)blockIdx.x < num_items) {
    issue_tile(blockIdx.x, 0);
    issue_weights(blockIdx.x);
  }
  int it = 0;
#pragma unroll 1
  for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
    const int next = item + gridDim.x;
    if (threadIdx.x == 0 && next < num_items) {   // buffer (it+1)&1 was released by the barrier that ended the last item
      bw::fence_proxy_async_smem();
      issue_tile(next, (it + 1) & 1);
    }
    const int chunk = item % nch;
    int t = item / nch;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int x0 = tx * TW, y0 = ty * TH;
    const int c = chunk * 32 + lane;
    // taps duplicated into both halves of a 64-bit register (see the one-shot kernel)
    bw::mbar_wait(&bar[2], it & 1);
    uint64_t w2[49];
#pragma unroll
    for (int k = 0; k < 49; ++k) {
      const float tv = wbuf[k * 32 + lane];
      w2[k] = pk2(tv, tv);
    }
    __syncthreads();                              // every warp holds its taps: the weights buffer is free
    if (threadIdx.x == 0 && next < num_items) {
      bw::fence_proxy_async_smem();
      issue_weights(next);
    }
    const float bc = bias ? __ldg(bias + c) : 0.f;
    uint64_t acc[2][8];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = pk2(bc, bc);

    bw::mbar_wait(&bar[it & 1], (it >> 1) & 1);
    const float* base = xs + (it & 1) * TILE_FLOATS + ((4 * sy) * PW + 8 * sx) * 32 + lane;
#pragma unroll
    for (int iy = 0; iy < 9; ++iy) {
      uint64_t pr[14];
#pragma unroll
      for (int j = 0; j < 14; ++j) {
        const volatile float* p0 = base + (iy * PW + j) * 32;
        pr[j] = pk2(p0[0], p0[PW * 32]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int ky = iy - 2 * q;
        if (ky < 0 || ky >= 7) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[q][j] = fma2(w2[ky * 7 + kx], pr[j + kx], acc[q][j]);
      }
    }
    if (ADD) {
      float av[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int oy = y0 + 4 * sy + a, ox = x0 + 8 * sx + j;
          av[a][j] = (oy < h && ox < w) ? __ldg(add + (((int64_t)b * h + oy) * w + ox) * C + c) : 0.f;
        }
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[q][j] = add2(acc[q][j], pk2(av[2 * q][j], av[2 * q + 1][j]));
    }
    {
      const int oy0 = y0 + 4 * sy, ox0 = x0 + 8 * sx;
      OT* yp = y + (((int64_t)b * h + oy0) * w + ox0) * C + c;
      const int rs = w * C;
      if (oy0 + 4 <= h && ox0 + 8 <= w) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v0, v1;
            up2(acc[q][j], v0, v1);
            store1(yp + (2 * q) * rs + j * C, v0);
            store1(yp + (2 * q + 1) * rs + j * C, v1);
          }
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v0, v1;
            up2(acc[q][j], v0, v1);
            if (ox0 + j < w) {
              if (oy0 + 2 * q < h) store1(yp + (2 * q) * rs + j * C, v0);
              if (oy0 + 2 * q + 1 < h) store1(yp + (2 * q + 1) * rs + j * C, v1);
            }
          }
      }
    }
    __syncthreads();   // every warp is done with tile buffer it&1 before it is refilled
  }
}

// Weight/bias gradient of the depthwise 7x7 (training backward of cod.py:1106):
//   dW[ky,kx,c] = sum_{b,oy,ox} dy[b,oy,ox,c] * x[b,oy+ky-3,ox+kx-3,c],   db[c] = sum dy
// CTA = one 32-channel chunk, looping over (image, tile) pairs with stride gridDim.y; x halo tiles
// arrive by TMA (double buffered, zero fill = padding), dy of the warp's 4x8 sub-tile sits in
// registers, the 49 + 1 sums stay in registers across all tiles of the CTA.  One partial (50 x 32)
// per CTA, reduced across warps in shared memory in a fixed order (deterministic).
template <int SY, int SX>
__global__ void __launch_bounds__(SY * SX * 32, 2)
dwconv7_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ dy,
                         float* __restrict__ part, int h, int w, int C, int tiles_x, int tiles_y, int ntiles) {
  constexpr int TH = 4 * SY, TW = 8 * SX, PH = TH + 6, PW = TW + 6, NW = SY * SX;
  constexpr int TILE_FLOATS = PH * PW * 32;
  constexpr uint32_t TILE_BYTES = TILE_FLOATS * 4;
  static_assert(NW * 50 * 32 <= 2 * TILE_FLOATS, "reduction scratch must fit in the tile buffers");
  extern __shared__ __align__(128) float xs[];
  __shared__ uint64_t bar[2];

  const int chunk = blockIdx.x, P = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sy = warp / SX, sx = warp - sy * SX;
  const int c = chunk * 32 + lane;

  auto issue = [&](int t, int buf) {
    const int tx = t % tiles_x;
    const int r = t / tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    bw::mbar_arrive_expect_tx(&bar[buf], TILE_BYTES);
    bw::tma_load_4d(&tmX, &bar[buf], xs + buf * TILE_FLOATS, chunk * 32, tx * TW - 3, ty * TH - 3, b);
  };
  if (threadIdx.x == 0) {
    bw::prefetch_tmap(&tmX);
    bw::mbar_init(&bar[0], 1);
    bw::mbar_init(&bar[1], 1);
    bw::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) issue(blockIdx.y, 0);

  float acc[50];
#pragma unroll
  for (int k = 0; k < 50; ++k) acc[k] = 0.f;

  int it = 0;
#pragma unroll 1
  for (int t = blockIdx.y; t < ntiles; t += P, ++it) {
    if (threadIdx.x == 0 && t + P < ntiles) {   // buffer (it+1)&1 was released by the barrier below
      bw::fence_proxy_async_smem();
      issue(t + P, (it + 1) & 1);
    }
    const int tx = t % tiles_x;
    const int r = t / tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    const int oy0 = ty * TH + 4 * sy, ox0 = tx * TW + 8 * sx;
    float g[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int oy = oy0 + a, ox = ox0 + j;
        g[a][j] = (oy < h && ox < w) ? __ldg(dy + (((int64_t)b * h + oy) * w + ox) * C + c) : 0.f;
        acc[49] += g[a][j];
      }
    bw::mbar_wait(&bar[it & 1], (it >> 1) & 1);
    const float* base = xs + (it & 1) * TILE_FLOATS + ((4 * sy) * PW + 8 * sx) * 32 + lane;
#pragma unroll
    for (int iy = 0; iy < 10; ++iy) {
      float in[14];
#pragma unroll
      for (int j = 0; j < 14; ++j) in[j] = base[(iy * PW + j) * 32];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int ky = iy - a;
        if (ky < 0 || ky >= 7) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[ky * 7 + kx] = fmaf(g[a][j], in[j + kx], acc[ky * 7 + kx]);
      }
    }
    __syncthreads();   // every warp is done with buffer it&1 before it is refilled
  }
  // all issued loads were waited on: the tile buffers are free for the cross-warp reduction
  float* red = xs;
#pragma unroll
  for (int k = 0; k < 50; ++k) red[(warp * 50 + k) * 32 + lane] = acc[k];
  __syncthreads();
  for (int i = threadIdx.x; i < 50 * 32; i += NW * 32) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < NW; ++q) v += red[q * 50 * 32 + i];
    part[((int64_t)blockIdx.y * 50 + (i >> 5)) * C + chunk * 32 + (i & 31)] = v;
  }
}

template <typename OT, int VPL>
__global__ void __launch_bounds__(256)
ln_rows_kernel(const float* __restrict__ y, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
               OT* __restrict__ out, int64_t rows, int C, float eps) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = y + row * C;
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    v[j] = *reinterpret_cast<const float4*>(p + (j * 32 + lane) * 4);
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
  OT* o = out + row * C;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c0 = (j * 32 + lane) * 4;
    float4 g = *reinterpret_cast<const float4*>(ln_w + c0), be = *reinterpret_cast<const float4*>(ln_b + c0);
    store4(o + c0, (v[j].x - mean) * rstd * g.x + be.x, (v[j].y - mean) * rstd * g.y + be.y,
           (v[j].z - mean) * rstd * g.z + be.z, (v[j].w - mean) * rstd * g.w + be.w);
  }
}

// (mean, rstd) of every row of a bf16 matrix (rows x C), statistics in fp32 from the stored (rounded) values -- the
// values the LayerNorm-folded GEMM multiplies, so the centring in its epilogue is exact for them.
// LPR lanes share a row (16-byte loads, 8 values each), V loads per lane and row, U row groups per warp: every lane has
// U * V = 4 independent 16-byte loads in flight (the one-row-per-warp form left 256 B per warp in flight and ran the
// C = 128 stage at 1.8 TB/s); the two-pass variance runs on the registers.
template <int LPR, int V, int U>
__global__ void __launch_bounds__(256)
row_stats_bf16_kernel(const __nv_bfloat16* __restrict__ y, float2* __restrict__ stats, int64_t rows, int C, float eps) {
  constexpr int RPP = 32 / LPR;          // rows per pass of a warp
  constexpr int RPW = RPP * U;           // rows per warp
  const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
  const int64_t row0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + sub;
  uint4 raw[U][V];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t row = row0 + u * RPP;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      raw[u][j] = make_uint4(0u, 0u, 0u, 0u);
      if (row < rows) raw[u][j] = __ldg(reinterpret_cast<const uint4*>(y + row * C + (j * LPR + l) * 8));
    }
  }
  const float inv_c = 1.0f / C;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float f[V][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t w4[4] = {raw[u][j].x, raw[u][j].y, raw[u][j].z, raw[u][j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {   // bf16 -> fp32 is a 16-bit shift
        f[j][2 * e] = __uint_as_float(w4[e] << 16);
        f[j][2 * e + 1] = __uint_as_float(w4[e] & 0xffff0000u);
      }
      s += ((f[j][0] + f[j][1]) + (f[j][2] + f[j][3])) + ((f[j][4] + f[j][5]) + (f[j][6] + f[j][7]));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float a = f[j][e] - mean, b = f[j][e + 1] - mean;
        q += a * a + b * b;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const int64_t row = row0 + u * RPP;
    if (l == 0 && row < rows) stats[row] = make_float2(mean, 1.0f / sqrtf(q * inv_c + eps));
  }
}

// one-shot grid (a persistent, register-prefetching form of this kernel measured 5-15 % slower)
static int row_stats_grid(int64_t rows, int rows_per_block) { return (int)cdiv(rows, rows_per_block); }

int row_stats_bf16(const __nv_bfloat16* y, float2* stats, int64_t rows, int C, float eps, cudaStream_t s) {
  if (reinterpret_cast<uintptr_t>(y) & 15) {
    set_error("row_stats: the matrix must be 16-byte aligned");
    return -1;
  }
  switch (C / 128) {   // rows per block = 8 warps x rows per warp
    case 1: row_stats_bf16_kernel<16, 1, 4><<<row_stats_grid(rows, 64), 256, 0, s>>>(y, stats, rows, C, eps); break;
    case 2: row_stats_bf16_kernel<32, 1, 4><<<row_stats_grid(rows, 32), 256, 0, s>>>(y, stats, rows, C, eps); break;
    case 4: row_stats_bf16_kernel<32, 2, 2><<<row_stats_grid(rows, 16), 256, 0, s>>>(y, stats, rows, C, eps); break;
    case 8: row_stats_bf16_kernel<32, 4, 1><<<row_stats_grid(rows, 8), 256, 0, s>>>(y, stats, rows, C, eps); break;
    default:
      set_error("row_stats: C=%d must be 128*{1,2,4,8}", C);
      return -1;
  }
  return 0;
}

// persistent launch (both output types, with / without the residual operand); 0 ok, < 0 error
template <int SY, int SX, bool ADD, typename OT>
static int dw_launch_persist(const CUtensorMap& tm, const float* wT, const float* bias, const float* add, OT* y, int B, int h,
                             int w, int C, cudaStream_t s) {
  constexpr int PH = 4 * SY + 6, PW = 8 * SX + 6;
  constexpr int SMEM = 2 * PH * PW * 128 + 49 * 32 * 4 + 128;
  auto kern = dwconv7_tma_persist_kernel<SY, SX, ADD, OT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("dwconv7_tma(persistent): cannot opt in to %d B smem: %s", SMEM, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  CUtensorMap tmW;
  {
    const uint64_t dims[2] = {(uint64_t)C, 49}, strides[1] = {(uint64_t)C * 4};
    const uint32_t box[2] = {32, 49};
    int rc = make_tmap(&tmW, wT, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  const int tiles_x = cdiv(w, 8 * SX), tiles_y = cdiv(h, 4 * SY);
  const int64_t items = (int64_t)B * tiles_x * tiles_y * (C / 32);
  if (items > 0x7fffffff) {
    set_error("dwconv7_tma(persistent): too many items");
    return -1;
  }
  const int slots = 2 * sm_count_dw();
  const int grid = items < slots ? (int)items : slots;
  kern<<<grid, SY * SX * 32, SMEM, s>>>(tm, tmW, bias, add, y, h, w, C, tiles_x, tiles_y, (int)items);
  return 0;
}

template <int SY, int SX>
static int dw_launch_bf16(const CUtensorMap& tm, const float* wT, const float* bias, __nv_bfloat16* y, int B, int h, int w,
                          int C, cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(wT) & 15) == 0)
    return dw_launch_persist<SY, SX, false, __nv_bfloat16>(tm, wT, bias, nullptr, y, B, h, w, C, s);
  constexpr int PH = 4 * SY + 6, PW = 8 * SX + 6;
  constexpr int SMEM = 2 * PH * PW * 128 + 128;
  auto kern = dwconv7_tma_kernel<SY, SX, false, __nv_bfloat16>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("dwconv7_tma(bf16): cannot opt in to %d B smem: %s", SMEM, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  const int tiles_x = cdiv(w, 8 * SX), tiles_y = cdiv(h, 4 * SY);
  const int64_t blocks = (int64_t)B * tiles_x * tiles_y * (C / 128);
  kern<<<(unsigned)blocks, SY * SX * 32, SMEM, s>>>(tm, wT, bias, nullptr, y, h, w, C, tiles_x, tiles_y);
  return 0;
}

template <int SY, int SX>
static int dw_launch(const CUtensorMap& tm, const float* wT, const float* bias, const float* add, float* y, int B,
                     int h, int w, int C, cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(wT) & 15) == 0) {
    if (add) return dw_launch_persist<SY, SX, true, float>(tm, wT, bias, add, y, B, h, w, C, s);
    return dw_launch_persist<SY, SX, false, float>(tm, wT, bias, nullptr, y, B, h, w, C, s);
  }
  constexpr int PH = 4 * SY + 6, PW = 8 * SX + 6;
  constexpr int SMEM = 2 * PH * PW * 128 + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv7_tma_kernel<SY, SX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dwconv7_tma_kernel<SY, SX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("dwconv7_tma: cannot opt in to %d B smem: %s", SMEM, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  const int tiles_x = cdiv(w, 8 * SX), tiles_y = cdiv(h, 4 * SY);
  const int64_t blocks = (int64_t)B * tiles_x * tiles_y * (C / 128);
  if (add)
    dwconv7_tma_kernel<SY, SX, true><<<(unsigned)blocks, SY * SX * 32, SMEM, s>>>(tm, wT, bias, add, y, h, w, C, tiles_x, tiles_y);
  else
    dwconv7_tma_kernel<SY, SX, false><<<(unsigned)blocks, SY * SX * 32, SMEM, s>>>(tm, wT, bias, add, y, h, w, C, tiles_x, tiles_y);
  return 0;
}

template <typename OT>
static int ln_rows_launch(const float* y, const float* ln_w, const float* ln_b, OT* out, int64_t rows, int C,
                          float eps, cudaStream_t s) {
  const int blocks = cdiv(rows, 8);
  switch (C / 128) {
    case 1: ln_rows_kernel<OT, 1><<<blocks, 256, 0, s>>>(y, ln_w, ln_b, out, rows, C, eps); break;
    case 2: ln_rows_kernel<OT, 2><<<blocks, 256, 0, s>>>(y, ln_w, ln_b, out, rows, C, eps); break;
    case 4: ln_rows_kernel<OT, 4><<<blocks, 256, 0, s>>>(y, ln_w, ln_b, out, rows, C, eps); break;
    case 8: ln_rows_kernel<OT, 8><<<blocks, 256, 0, s>>>(y, ln_w, ln_b, out, rows, C, eps); break;
    default:
      set_error("dwconv7_ln(tma): C=%d must be 128*{1,2,4,8}", C);
      return -1;
  }
  return 0;
}

// TMA depthwise conv only: y = conv(x; wT (49,C), bias) (+ add).  Returns 1 when the shape is not
// handled here (C not a multiple of 128 / misaligned x).
int dwconv7_tma(const float* x, const float* wT, const float* dw_b, const float* add, float* y, int B, int h, int w,
                int C, cudaStream_t s) {
  if (C % 128 || C > 1024 || (reinterpret_cast<uintptr_t>(x) & 15)) return 1;
  const int SX = (w % 24 == 0 && w % 16 != 0) ? 3 : 2;                 // 24-wide maps: 8x24 tiles
  const int SY = (h % 8 == 0 || h > 12) ? 2 : 3;                       // 12-high maps: 12x16 tiles
  CUtensorMap tm;
  {
    PFN_tmapEncodeTiled enc = get_tmap_encoder();
    if (!enc) return -3;
    cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
    cuuint64_t gs[3] = {(cuuint64_t)C * 4, (cuuint64_t)w * C * 4, (cuuint64_t)h * w * C * 4};
    cuuint32_t bx[4] = {32, (cuuint32_t)(8 * SX + 6), (cuuint32_t)(4 * SY + 6), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("dwconv7(tma): cuTensorMapEncodeTiled failed (%d) for x (%d,%d,%d,%d)", (int)r, B, h, w, C);
      return -3;
    }
  }
  int rc;
  if (SY == 2 && SX == 2) rc = dw_launch<2, 2>(tm, wT, dw_b, add, y, B, h, w, C, s);
  else if (SY == 2 && SX == 3) rc = dw_launch<2, 3>(tm, wT, dw_b, add, y, B, h, w, C, s);
  else if (SY == 3 && SX == 2) rc = dw_launch<3, 2>(tm, wT, dw_b, add, y, B, h, w, C, s);
  else rc = dw_launch<3, 3>(tm, wT, dw_b, add, y, B, h, w, C, s);
  if (rc) return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("dwconv7_tma: launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  count_launch();
  return 0;
}


template <int SY, int SX>
static int dwg_launch(const CUtensorMap& tm, const float* dy, float* part, int P, int B, int h, int w, int C,
                      cudaStream_t s) {
  constexpr int PH = 4 * SY + 6, PW = 8 * SX + 6;
  constexpr int SMEM = 2 * PH * PW * 128 + 128;
  auto kern = dwconv7_wgrad_tma_kernel<SY, SX>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("dwconv7_wgrad_tma: cannot opt in to %d B smem: %s", SMEM, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  const int tiles_x = cdiv(w, 8 * SX), tiles_y = cdiv(h, 4 * SY);
  kern<<<dim3(C / 32, P), SY * SX * 32, SMEM, s>>>(tm, dy, part, h, w, C, tiles_x, tiles_y, B * tiles_x * tiles_y);
  return 0;
}

static int dw_tile_shape(int h, int w, int* SY, int* SX) {
  *SX = (w % 24 == 0 && w % 16 != 0) ? 3 : 2;   // 24-wide maps: 8x24 tiles
  *SY = (h % 8 == 0 || h > 12) ? 2 : 3;         // 12-high maps: 12x16 tiles
  return 0;
}

static int dw_make_tmap(CUtensorMap* tm, const float* x, int B, int h, int w, int C, int SY, int SX) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return -3;
  cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
  cuuint64_t gs[3] = {(cuuint64_t)C * 4, (cuuint64_t)w * C * 4, (cuuint64_t)h * w * C * 4};
  cuuint32_t bx[4] = {32, (cuuint32_t)(8 * SX + 6), (cuuint32_t)(4 * SY + 6), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("dwconv7(tma): cuTensorMapEncodeTiled failed (%d) for x (%d,%d,%d,%d)", (int)r, B, h, w, C);
    return -3;
  }
  return 0;
}

// Partial weight/bias gradients: part[P][50][C] (49 taps + bias), P <= max_parts returned in *parts.
// Returns 1 when the shape is not handled here.
int dwconv7_wgrad_tma(const float* x, const float* dy, float* part, int max_parts, int* parts, int B, int h, int w,
                      int C, cudaStream_t s) {
  if (C % 128 || (reinterpret_cast<uintptr_t>(x) & 15) || (int64_t)B * h * w < 2048) return 1;
  int SY, SX;
  dw_tile_shape(h, w, &SY, &SX);
  const int ntiles = B * cdiv(w, 8 * SX) * cdiv(h, 4 * SY);
  int P = cdiv(2 * 148, C / 32);
  if (P > ntiles) P = ntiles;
  if (P > max_parts) P = max_parts;
  if (P < 1) return 1;
  CUtensorMap tm;
  int rc = dw_make_tmap(&tm, x, B, h, w, C, SY, SX);
  if (rc) return rc;
  if (SY == 2 && SX == 2) rc = dwg_launch<2, 2>(tm, dy, part, P, B, h, w, C, s);
  else if (SY == 2 && SX == 3) rc = dwg_launch<2, 3>(tm, dy, part, P, B, h, w, C, s);
  else if (SY == 3 && SX == 2) rc = dwg_launch<3, 2>(tm, dy, part, P, B, h, w, C, s);
  else rc = dwg_launch<3, 3>(tm, dy, part, P, B, h, w, C, s);
  if (rc) return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("dwconv7_wgrad_tma: launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  count_launch();
  *parts = P;
  return 0;
}

// bf16-mode front of a ConvNeXt block with the LayerNorm folded into pwconv1: y = dwconv(x) stored ONCE as bf16
// (no fp32 scratch), then (mean, rstd) per pixel from the stored values.  Returns 1 when the shape is not handled.
int dwconv7_stats_tma(const float* x, const float* wT, const float* dw_b, __nv_bfloat16* y, float2* stats, int B, int h,
                      int w, int C, float eps, cudaStream_t s) {
  if (C % 128 || C > 1024 || (reinterpret_cast<uintptr_t>(x) & 15)) return 1;
  int SY, SX;
  dw_tile_shape(h, w, &SY, &SX);
  CUtensorMap tm;
  int rc = dw_make_tmap(&tm, x, B, h, w, C, SY, SX);
  if (rc) return rc;
  if (SY == 2 && SX == 2) rc = dw_launch_bf16<2, 2>(tm, wT, dw_b, y, B, h, w, C, s);
  else if (SY == 2 && SX == 3) rc = dw_launch_bf16<2, 3>(tm, wT, dw_b, y, B, h, w, C, s);
  else if (SY == 3 && SX == 2) rc = dw_launch_bf16<3, 2>(tm, wT, dw_b, y, B, h, w, C, s);
  else rc = dw_launch_bf16<3, 3>(tm, wT, dw_b, y, B, h, w, C, s);
  if (rc) return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("dwconv7_stats_tma: launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  count_launch();
  rc = row_stats_bf16(y, stats, (int64_t)B * h * w, C, eps, s);
  if (rc) return rc;
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("row_stats: launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  count_launch();
  return 0;
}

// Two launches: TMA depthwise conv into `ws` (B*h*w*C fp32), then LayerNorm rows -> out.
int dwconv7_ln_tma(const float* x, const float* wT, const float* dw_b, const float* ln_w, const float* ln_b,
                   float* ws, void* out, int out_dtype, int B, int h, int w, int C, float eps, cudaStream_t s) {
  int rc = dwconv7_tma(x, wT, dw_b, nullptr, ws, B, h, w, C, s);
  if (rc) return rc;
  const int64_t rows = (int64_t)B * h * w;
  return out_dtype == DGTD_BF16 ? ln_rows_launch(ws, ln_w, ln_b, (__nv_bfloat16*)out, rows, C, eps, s)
                                : ln_rows_launch(ws, ln_w, ln_b, (float*)out, rows, C, eps, s);
}

// LayerNorm rows only (fp32 in, fp32|bf16 out); C multiple of 128
int ln_rows_any(const float* y, const float* ln_w, const float* ln_b, void* out, int out_dtype, int64_t rows, int C,
                float eps, cudaStream_t s) {
  return out_dtype == DGTD_BF16 ? ln_rows_launch(y, ln_w, ln_b, (__nv_bfloat16*)out, rows, C, eps, s)
                                : ln_rows_launch(y, ln_w, ln_b, (float*)out, rows, C, eps, s);
}

}  // namespace dgtd
