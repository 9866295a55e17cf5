// a8, stage-0 blocks (C = 128) in bf16 inference: pwconv1 -> GELU -> pwconv2 -> layer scale -> residual as ONE kernel
// (cod.py:1109-1116), the 4C hidden tensor never leaving the SM.
//
//   out[m, :] = x[m, :] + gamma * ( GELU( LN(y[m, :]) W1^T + b1 ) W2^T + b2 )
//
// with the LayerNorm folded into the first GEMM as in dgtd_linear_lnfold_fwd (W1' = W1 gamma_LN, col_s = sum_k W1',
// cbias = W1 beta_LN + b1, (mean, rstd) per row from the stored bf16 conv output).
//
// Why: un-fused, a stage-0 block writes the hidden tensor (M x 512 bf16 = 604 MB at B = 64) and reads it back; the two
// GEMMs take 257 + 185 us for 2 x 77 GFLOP, i.e. they are HBM / epilogue bound, not tensor bound.  Fused, the block moves
// y (151 MB) + x in / out (2 x 302 MB) and the tensor pipe needs 68 us.
//
// Per 128-row tile the hidden tensor is produced 128 columns (one "chunk") at a time:
//   GEMM1(c): acc1[c & 1] (TMEM, 128 columns) = Y[128 x 128] . W1'[chunk c]^T            8 MMAs 128 x 128 x 16
//   epilogue-1(c): acc1 -> LN fold + bias + GELU -> bf16 -> H[c & 1] in shared memory, written directly in the
//                  K-major SWIZZLE_128B layout an A operand has (what a TMA box would have produced)
//   GEMM2(c): acc2[tile & 1] (TMEM, 128 columns) += H[c & 1] . W2[:, chunk c]^T          8 MMAs 128 x 128 x 16
//   epilogue-2: acc2 -> gamma * (. + b2) + x -> out; the residual tile comes in and the result leaves through a
//               128B-swizzled 32 x 32 staging tile per warp by TMA (a lane owns a ROW of the accumulator: direct global
//               accesses would be 32 different lines per instruction -- measured first, 563 us per launch)
// issued in the order G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) G2(3), so the GELU of chunk c runs under G1(c+1) and
// G2(c-1).  TMEM: 2 x 128 (acc1) + 2 x 128 (acc2) = 512 columns.  Shared memory: Y tile 32 KB, H 2 x 32 KB, a ring of
// four 16 KB weight blocks ([128 rows x 64 K], streamed from L2 in consumption order: 256 KB per tile), 16 x 4 KB staging.
//
// The GELU epilogue is what bounds a stage-0 block (the un-fused pwconv1 spends 257 us on 77 GFLOP): with two warps per
// scheduler it issues at a quarter of the slot rate (MUFU / TMEM-load / LDG latencies in a dependent chain), so this
// kernel runs SIXTEEN epilogue warps -- four per TMEM lane quadrant, 32 columns each.
//
// Warp roles (640 threads, 1 CTA / SM, persistent): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w4-19 epilogues.
#include "tc_common.cuh"

namespace dgtd {

namespace mlpf {

constexpr int C = 128, HID = 4 * C, HC = 128, NCH = HID / HC;   // hidden chunks per tile
constexpr int BLK = 128 * 64 * 2;                               // one [128 x 64] bf16 block, 16 KB
constexpr int KB1 = C / 64, KB2 = HC / 64;                      // K blocks of GEMM1 / of one GEMM2 chunk
constexpr int A_BYTES = KB1 * BLK, H_BYTES = KB2 * BLK;
constexpr int WST = 4;                                          // weight ring
constexpr int NEPI = 16;                                        // epilogue warps
constexpr int STG_BYTES = NEPI * 4096;
constexpr int OFF_A = 0, OFF_H = A_BYTES, OFF_W = OFF_H + 2 * H_BYTES, OFF_STG = OFF_W + WST * BLK, OFF_BAR = OFF_STG + STG_BYTES;
constexpr int NBARS = 2 + 2 * WST + 4 + 4 + 4 + NEPI;
constexpr int SMEM = OFF_BAR + NBARS * 8 + 16;
constexpr int THREADS = 128 + NEPI * 32;
static_assert(SMEM <= 232448, "shared memory budget");

struct Params {
  const float* row_stats;   // (M, 2): mean, rstd of the stored bf16 conv output
  const float* col_s;       // (HID)
  const float* cbias;       // (HID)
  const float* b2;          // (C)
  const float* gamma;       // (C) nullable
  const float* res;         // (M, C) fp32
  float* out;               // (M, C) fp32 (may alias res: every element is read and written by the same thread)
  int M, tiles;
};

__device__ __forceinline__ uint32_t swz_l(int lane) { return (uint32_t)(lane & 7); }

__global__ void __launch_bounds__(THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmRes,
                 const __grid_constant__ CUtensorMap tmOut, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem + OFF_A;
  uint8_t* sH = smem + OFF_H;
  uint8_t* sW = smem + OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint8_t* sStg = smem + OFF_STG;
  uint64_t* a_full = bars;            // [1]
  uint64_t* a_empty = a_full + 1;     // [1]
  uint64_t* w_full = a_empty + 1;     // [WST]
  uint64_t* w_empty = w_full + WST;   // [WST]
  uint64_t* c1_full = w_empty + WST;  // [2] acc1 complete
  uint64_t* c1_empty = c1_full + 2;   // [2] acc1 drained
  uint64_t* h_full = c1_empty + 2;    // [2] hidden chunk written
  uint64_t* h_empty = h_full + 2;     // [2] hidden chunk consumed
  uint64_t* c2_full = h_empty + 2;    // [2]
  uint64_t* c2_empty = c2_full + 2;   // [2]
  uint64_t* r_full = c2_empty + 2;    // [NEPI] residual tile of the warp landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_full + NEPI);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    bw::prefetch_tmap(&tmY);
    bw::prefetch_tmap(&tmW1);
    bw::prefetch_tmap(&tmW2);
    bw::prefetch_tmap(&tmRes);
    bw::prefetch_tmap(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    bw::mbar_init(a_full, 1);
    bw::mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      bw::mbar_init(&c1_full[i], 1);
      bw::mbar_init(&c1_empty[i], NEPI);   // one arrival per epilogue warp
      bw::mbar_init(&h_full[i], NEPI);
      bw::mbar_init(&h_empty[i], 1);
      bw::mbar_init(&c2_full[i], 1);
      bw::mbar_init(&c2_empty[i], NEPI);
    }
    for (int i = 0; i < NEPI; ++i) bw::mbar_init(&r_full[i], 1);
    for (int i = 0; i < WST; ++i) {
      bw::mbar_init(&w_full[i], 1);
      bw::mbar_init(&w_empty[i], 1);
    }
    bw::fence_mbar_init();
  }
  if (warp == 2) bw::tmem_alloc(tmem_slot, 512);
  bw::tc_fence_before();
  __syncthreads();
  bw::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer: the Y tile, then the weight blocks in the order the MMAs use them ==========
      int ws = 0;
      uint32_t wph = 0;
      auto load_w = [&](const CUtensorMap* tm, int k0, int r0) {
        bw::mbar_wait(&w_empty[ws], wph ^ 1);
        bw::mbar_arrive_expect_tx(&w_full[ws], BLK);
        bw::tma_load_2d(tm, &w_full[ws], sW + ws * BLK, k0, r0);
        if (++ws == WST) { ws = 0; wph ^= 1; }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        bw::mbar_wait(a_empty, (it & 1) ^ 1);
        bw::mbar_arrive_expect_tx(a_full, A_BYTES);
        for (int kb = 0; kb < KB1; ++kb) bw::tma_load_2d(&tmY, a_full, sA + kb * BLK, kb * 64, tile * 128);
        for (int s = 0; s <= NCH; ++s) {
          if (s < NCH)
            for (int kb = 0; kb < KB1; ++kb) load_w(&tmW1, kb * 64, s * HC);            // W1' rows of hidden chunk s
          if (s >= 1)
            for (int kb = 0; kb < KB2; ++kb) load_w(&tmW2, (s - 1) * HC + kb * 64, 0);   // W2 columns of chunk s - 1
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = bw::umma_idesc_bf16(128, 128);
      int ws = 0;
      uint32_t wph = 0;
      uint32_t q1 = 0, q2 = 0;   // hidden chunks started by GEMM1 / by GEMM2 (global counters)
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        bw::mbar_wait(a_full, it & 1);
        for (int s = 0; s <= NCH; ++s) {
          if (s < NCH) {   // GEMM1 of chunk s
            const uint32_t b = q1 & 1;
            bw::mbar_wait(&c1_empty[b], ((q1 >> 1) & 1) ^ 1);
            bw::tc_fence_after();
            const uint32_t d = tmem_base + b * HC;
            for (int kb = 0; kb < KB1; ++kb) {
              bw::mbar_wait(&w_full[ws], wph);
              bw::tc_fence_after();
              const uint64_t da = bw::umma_smem_desc_kmajor(bw::smem_u32(sA + kb * BLK), 128);
              const uint64_t db = bw::umma_smem_desc_kmajor(bw::smem_u32(sW + ws * BLK), 128);
#pragma unroll
              for (int k = 0; k < 4; ++k) bw::umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
              bw::umma_commit(&w_empty[ws]);
              if (++ws == WST) { ws = 0; wph ^= 1; }
            }
            bw::umma_commit(&c1_full[b]);
            if (s == NCH - 1) bw::umma_commit(a_empty);   // the Y tile is no longer read
            ++q1;
          }
          if (s >= 1) {    // GEMM2 of chunk s - 1
            const int c = s - 1;
            const uint32_t b = q2 & 1;
            if (c == 0) {
              bw::mbar_wait(&c2_empty[ab], ((it >> 1) & 1) ^ 1);
              bw::tc_fence_after();
            }
            bw::mbar_wait(&h_full[b], (q2 >> 1) & 1);
            bw::tc_fence_after();
            const uint32_t d = tmem_base + 2 * HC + ab * C;
            for (int kb = 0; kb < KB2; ++kb) {
              bw::mbar_wait(&w_full[ws], wph);
              bw::tc_fence_after();
              const uint64_t da = bw::umma_smem_desc_kmajor(bw::smem_u32(sH + b * H_BYTES + kb * BLK), 128);
              const uint64_t db = bw::umma_smem_desc_kmajor(bw::smem_u32(sW + ws * BLK), 128);
#pragma unroll
              for (int k = 0; k < 4; ++k) bw::umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (c | kb | k) != 0);
              bw::umma_commit(&w_empty[ws]);
              if (++ws == WST) { ws = 0; wph ^= 1; }
            }
            bw::umma_commit(&h_empty[b]);
            if (c == NCH - 1) bw::umma_commit(&c2_full[ab]);
            ++q2;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogues (16 warps): quadrant = TMEM lanes, part = 32 of the 128 columns =====================
    const int quad = warp & 3, part = (warp - 4) >> 2, ew = warp - 4;
    const int r = quad * 32 + lane;                       // row of the tile
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    // this row of K block part / 2 of H; the warp's 32 hidden columns are 16-byte chunks (part & 1) * 4 .. + 3 of the line
    uint8_t* hrow_p = sH + (part >> 1) * BLK + r * 128;
    const uint32_t swz = (uint32_t)(r & 7);
    uint8_t* stg = sStg + ew * 4096;                      // 32 rows x 32 fp32 columns, 128B swizzle
    uint8_t* srow_p = stg + lane * 128;
    uint64_t* rbar = &r_full[ew];
    uint32_t q = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const int row0 = tile * 128 + quad * 32;
      if (lane == 0) {   // residual tile of this warp: requested now, needed after the four hidden chunks
        bw::tma_store_wait_read<0>();                     // the previous tile's store has read the staging tile
        bw::mbar_arrive_expect_tx(rbar, 4096);
        bw::tma_load_2d(&tmRes, rbar, stg, part * 32, row0);
      }
      const float2 rs = __ldg(reinterpret_cast<const float2*>(p.row_stats) + row0 + lane);
      const uint64_t nm = pk2(-rs.x, -rs.x), rr = pk2(rs.y, rs.y);
      for (int c = 0; c < NCH; ++c, ++q) {
        const uint32_t b = q & 1;
        bw::mbar_wait(&c1_full[b], (q >> 1) & 1);
        bw::tc_fence_after();
        uint32_t v[32];
        bw::tmem_ld_32x32(tmem_base + lane_addr + b * HC + part * 32, v);
        bw::tmem_ld_wait();
        bw::tc_fence_before();
        __syncwarp();
        if (lane == 0) bw::mbar_arrive(&c1_empty[b]);      // acc1[b] may be overwritten by GEMM1 of chunk q + 2
        bw::mbar_wait(&h_empty[b], ((q >> 1) & 1) ^ 1);    // GEMM2 of chunk q - 2 has read H[b]
        const int n0 = c * HC + part * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.col_s + n0 + j));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.col_s + n0 + j + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.cbias + n0 + j));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.cbias + n0 + j + 4));
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]);
          // LayerNorm fold: rstd * (acc - mean * s) + c  (tc_common.cuh, same arithmetic as the un-fused epilogue)
          uint64_t x0 = fma2(nm, pk2(s0.x, s0.y), pk2(f[0], f[1])), x1 = fma2(nm, pk2(s0.z, s0.w), pk2(f[2], f[3]));
          uint64_t x2 = fma2(nm, pk2(s1.x, s1.y), pk2(f[4], f[5])), x3 = fma2(nm, pk2(s1.z, s1.w), pk2(f[6], f[7]));
          up2(fma2(rr, x0, pk2(b0.x, b0.y)), f[0], f[1]);
          up2(fma2(rr, x1, pk2(b0.z, b0.w)), f[2], f[3]);
          up2(fma2(rr, x2, pk2(b1.x, b1.y)), f[4], f[5]);
          up2(fma2(rr, x3, pk2(b1.z, b1.w)), f[6], f[7]);
#pragma unroll
          for (int e = 0; e < 8; e += 2) DGTD_GELU2(f[e], f[e + 1]);
          const __nv_bfloat162 a0 = __floats2bfloat162_rn(f[0], f[1]), a1 = __floats2bfloat162_rn(f[2], f[3]);
          const __nv_bfloat162 a2 = __floats2bfloat162_rn(f[4], f[5]), a3 = __floats2bfloat162_rn(f[6], f[7]);
          // 16-byte chunk ch of the row sits at chunk ch ^ (row & 7) of its 128-byte line
          const uint32_t ch = (uint32_t)((part & 1) * 4 + (j >> 3));
          // (a plain store, not asm with a "memory" clobber: the clobber pinned the NEXT group's col_s / cbias loads behind
          // this store and exposed their latency four times per chunk -- r2 profile, long-scoreboard stalls on the FFMA2s)
          *reinterpret_cast<uint4*>(hrow_p + b * H_BYTES + ((ch ^ swz) << 4)) =
              make_uint4(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1),
                         *reinterpret_cast<const uint32_t*>(&a2), *reinterpret_cast<const uint32_t*>(&a3));
        }
        bw::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) bw::mbar_arrive(&h_full[b]);
      }
      // ---- epilogue-2: out = x + gamma * (acc2 + b2), 32 columns of this lane's row, through the staging tile ----
      bw::mbar_wait(&c2_full[ab], (it >> 1) & 1);
      bw::tc_fence_after();
      uint32_t v[32];
      bw::tmem_ld_32x32(tmem_base + lane_addr + 2 * HC + ab * C + part * 32, v);
      bw::tmem_ld_wait();
      bw::tc_fence_before();
      __syncwarp();
      if (lane == 0) bw::mbar_arrive(&c2_empty[ab]);
      bw::mbar_wait(rbar, it & 1);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4* sp = reinterpret_cast<float4*>(srow_p + ((((uint32_t)(j >> 2)) ^ swz_l(lane)) << 4));
        float4 x = *sp;
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + part * 32 + j));
        float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.gamma) g = __ldg(reinterpret_cast<const float4*>(p.gamma + part * 32 + j));
        x.x += g.x * (__uint_as_float(v[j]) + bb.x);
        x.y += g.y * (__uint_as_float(v[j + 1]) + bb.y);
        x.z += g.z * (__uint_as_float(v[j + 2]) + bb.z);
        x.w += g.w * (__uint_as_float(v[j + 3]) + bb.w);
        *sp = x;
      }
      bw::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        bw::tma_store_2d(&tmOut, stg, part * 32, row0);
        bw::tma_store_commit();
      }
    }
    if (lane == 0) bw::tma_store_wait_all<0>();
  }

  bw::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    bw::tc_fence_after();
    bw::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace mlpf
}  // namespace dgtd

using namespace dgtd;

extern "C" int dgtd_convnext_mlp_fused_fwd(const void* y, const float* row_stats, const void* w1, const float* col_s,
                                           const float* cbias, const void* w2, const float* b2, const float* gamma,
                                           const float* residual, float* out, int64_t M, int C, dgtd_stream_t stream) {
  using namespace mlpf;
  DGTD_CHECK_ARG(y && row_stats && w1 && col_s && cbias && w2 && b2 && residual && out, "convnext_mlp_fused: null pointer");
  DGTD_CHECK_ARG(C == mlpf::C, "convnext_mlp_fused: C = %d is not built (stage 0, C = 128, only)", C);
  DGTD_CHECK_ARG(M > 0 && M % 128 == 0 && M < ((int64_t)1 << 31), "convnext_mlp_fused: M = %lld must be a positive multiple of 128",
                 (long long)M);
  cudaStream_t s = (cudaStream_t)stream;
  CUtensorMap tmY, tmW1, tmW2;
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)M}, str[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {64, 128};
    if (make_tmap(&tmY, y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
  }
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)HID}, str[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {64, 128};
    if (make_tmap(&tmW1, w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
  }
  {
    uint64_t dims[2] = {(uint64_t)HID, (uint64_t)C}, str[1] = {(uint64_t)HID * 2};
    uint32_t box[2] = {64, 128};
    if (make_tmap(&tmW2, w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
  }
  CUtensorMap tmRes, tmOut;
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)M}, str[1] = {(uint64_t)C * 4};
    uint32_t box[2] = {32, 32};
    if (make_tmap(&tmRes, residual, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
    if (make_tmap(&tmOut, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("convnext_mlp_fused: cannot opt in to %d B of shared memory: %s", SMEM, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  Params p;
  p.row_stats = row_stats; p.col_s = col_s; p.cbias = cbias; p.b2 = b2; p.gamma = gamma; p.res = residual; p.out = out;
  p.M = (int)M; p.tiles = (int)(M / 128);
  const int grid = p.tiles < sm_count() ? p.tiles : sm_count();
  mlp_fused_kernel<<<grid, THREADS, SMEM, s>>>(tmY, tmW1, tmW2, tmRes, tmOut, p);
  DGTD_LAUNCH_CHECK("convnext_mlp_fused");
  return 0;
}
