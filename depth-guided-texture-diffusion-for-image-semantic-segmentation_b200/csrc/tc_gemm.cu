// Dense contractions of the path on the 5th-gen tensor cores (tcgen05 + TMEM + TMA):
//   C[M,N] = A[M,K] . B[N,K]^T   bf16 operands (K-major), fp32 accumulation in TMEM,
// with the epilogues of convnext_Block (cod.py:1109-1116), the down-sample convs (:1134) and
// the head / decoder convs fused on the accumulator read-back.
//
// One persistent CTA per SM, 256 threads, warp-specialised:
//   warp 0 (one lane)  TMA producer: 128xBK A tile + BNxBK B tile per stage, 128B swizzle
//   warp 1 (one lane)  tcgen05.mma issuer: 128 x BN x 16 UMMAs, accumulator in TMEM
//   warp 2             TMEM allocation / release
//   warps 4-11         epilogue: tcgen05.ld (32 lanes x 32 columns, double buffered; two warps per
//                      TMEM lane quadrant split the columns) -> registers -> bias / GELU (packed
//                      fp32x2) / gamma / DropPath scale / residual -> 16-byte global stores
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
#include <mutex>

#include "blackwell.cuh"
#include "simt_gemm.cuh"

namespace dgtd {

// ---------------------------------------------------------------- tensor-map plumbing
PFN_tmapEncodeTiled get_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_tmapEncodeTiled)p;
  });
  if (!fn) set_error("cuTensorMapEncodeTiled is not available from this driver");
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
              CUtensorMapSwizzle swz) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return -3;
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu] box=[%u,%u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
              rank > 1 ? box[1] : 0);
    return -3;
  }
  return 0;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---------------------------------------------------------------- fast GELU for the bf16 path
// Phi(x) = 0.5 + x Q(x^2) on |x| <= 4.5 (odd minimax-style fit of the erf form, clamped to
// [0,1]); max |gelu_fast - gelu_erf| = 2.8e-5 over the reals, i.e. ~1/100 of a bf16 ulp at 1.
__device__ __forceinline__ float gelu_fast(float x) {
  float xc = fminf(fmaxf(x, -4.5f), 4.5f);
  float t = xc * xc;
  float q = -1.400070736e-12f;
  q = fmaf(q, t, 1.697307069e-10f);
  q = fmaf(q, t, -9.193762573e-09f);
  q = fmaf(q, t, 2.958901695e-07f);
  q = fmaf(q, t, -6.365260363e-06f);
  q = fmaf(q, t, 9.787139965e-05f);
  q = fmaf(q, t, -1.122678685e-03f);
  q = fmaf(q, t, 9.833185488e-03f);
  q = fmaf(q, t, -6.633705714e-02f);
  q = fmaf(q, t, 3.988837948e-01f);
  float phi = fminf(fmaxf(fmaf(xc, q, 0.5f), 0.f), 1.f);
  return x * phi;
}

// Packed fp32x2 arithmetic (Blackwell FFMA2/FMUL2/FADD2): halves the issue slots of the epilogue.
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void up2(uint64_t r, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_fast on two values at once (same polynomial, 8 issue slots per element)
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  const float c0 = fminf(fmaxf(x0, -4.5f), 4.5f), c1 = fminf(fmaxf(x1, -4.5f), 4.5f);
  const uint64_t xc = pk2(c0, c1);
  const uint64_t t = mul2(xc, xc);
  uint64_t q = pk2(-1.400070736e-12f, -1.400070736e-12f);
  q = fma2(q, t, pk2(1.697307069e-10f, 1.697307069e-10f));
  q = fma2(q, t, pk2(-9.193762573e-09f, -9.193762573e-09f));
  q = fma2(q, t, pk2(2.958901695e-07f, 2.958901695e-07f));
  q = fma2(q, t, pk2(-6.365260363e-06f, -6.365260363e-06f));
  q = fma2(q, t, pk2(9.787139965e-05f, 9.787139965e-05f));
  q = fma2(q, t, pk2(-1.122678685e-03f, -1.122678685e-03f));
  q = fma2(q, t, pk2(9.833185488e-03f, 9.833185488e-03f));
  q = fma2(q, t, pk2(-6.633705714e-02f, -6.633705714e-02f));
  q = fma2(q, t, pk2(3.988837948e-01f, 3.988837948e-01f));
  const uint64_t phi = fma2(xc, q, pk2(0.5f, 0.5f));   // within 3e-5 of [0,1]: no clamp needed
  up2(mul2(pk2(x0, x1), phi), x0, x1);
}

__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

template <int ACT>
__device__ __forceinline__ float tc_act(float x) {
  if (ACT == DGTD_ACT_GELU) return gelu_fast(x);
  if (ACT == DGTD_ACT_RELU) return fmaxf(x, 0.f);
  return x;
}

struct TcParams {
  int M, N, K;
  int tiles_m, tiles_n;
  const float* bias;      // [N] nullable
  const float* gamma;     // [N] nullable          (RESIDUAL)
  const float* keep;      // [M / rows_per_sample]  nullable (RESIDUAL)
  const float* residual;  // [M, ldo] fp32          (RESIDUAL)
  int rows_per_sample;
  void* out;
  int64_t ldo;
};

template <int BN>
struct TcCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGES = BN >= 256 ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + 1024;  // +align slack
};

template <int BN, int ACT, typename OT, bool RESIDUAL>
__global__ void __launch_bounds__(384, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = bars + STAGES;          // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;      // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    bw::prefetch_tmap(&tmA);
    bw::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      bw::mbar_init(&full[i], 1);
      bw::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      bw::mbar_init(&tfull[i], 1);
      bw::mbar_init(&tempty[i], 256);
    }
    bw::fence_mbar_init();
  }
  if (warp == 2) bw::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  bw::tc_fence_before();
  __syncthreads();
  bw::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          bw::mbar_wait(&empty[stage], phase ^ 1);
          bw::mbar_arrive_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          bw::tma_load_2d(&tmA, &full[stage], sA + stage * Cfg::A_BYTES, kb * BK, m_blk * BM);
          bw::tma_load_2d(&tmB, &full[stage], sB + stage * Cfg::B_BYTES, kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = bw::umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1;
        bw::mbar_wait(&tempty[as], aphase ^ 1);  // epilogue has drained this accumulator
        bw::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          bw::mbar_wait(&full[stage], phase);
          bw::tc_fence_after();
          const uint64_t da = bw::umma_smem_desc_kmajor(bw::smem_u32(sA + stage * Cfg::A_BYTES), 128);
          const uint64_t db = bw::umma_smem_desc_kmajor(bw::smem_u32(sB + stage * Cfg::B_BYTES), 128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)  // +32 B per K=16 step inside the 128B swizzle atom
            bw::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          bw::umma_commit(&empty[stage]);            // smem slot free once these MMAs retire
          if (kb == num_kb - 1) bw::umma_commit(&tfull[as]);  // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // warp w may read TMEM lanes [32*(w%4), +32); the two warps of a quadrant split the columns.
    const int quad = warp & 3, half = (warp - 4) >> 2;
    constexpr int HALF_COLS = BN >= 64 ? BN / 2 : BN;          // columns per warp
    constexpr int NCH = HALF_COLS / 32;                         // 32-column chunks per warp
    const bool has_work = (BN >= 64) || half == 0;
    OT* out = reinterpret_cast<OT*>(p.out);
    int iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      bw::mbar_wait(&tfull[as], aphase);
      bw::tc_fence_after();
      const int row = m_blk * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      float ks = 1.f;
      if (RESIDUAL && p.keep && row_ok) ks = p.keep[row / p.rows_per_sample];
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + half * HALF_COLS;
      uint32_t v[2][32];
      if (has_work) bw::tmem_ld_32x32(t0, v[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (has_work) {
          bw::tmem_ld_wait();
          if (c + 1 < NCH) bw::tmem_ld_32x32(t0 + (c + 1) * 32, v[(c + 1) & 1]);
        }
        if (c == NCH - 1) {   // every TMEM read of this tile has landed: release the accumulator
          bw::tc_fence_before();
          bw::mbar_arrive(&tempty[as]);
        }
        const int col0 = n_blk * BN + half * HALF_COLS + c * 32;
        if (has_work && row_ok && col0 < p.N) {
          const uint32_t(&vv)[32] = v[c & 1];
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j >= p.N) break;  // N is a multiple of 8
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(vv[j + e]);
            if (p.bias) {
              float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j + 4));
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if (RESIDUAL) {
              float g[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
              if (p.gamma) {
                float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + j));
                float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + j + 4));
                g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w;
                g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
              }
              const float* rp = p.residual + (int64_t)row * p.ldo + col0 + j;
              float4 r0 = *reinterpret_cast<const float4*>(rp);
              float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
              float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = r[e] + ks * (g[e] * f[e]);
            } else if (ACT == DGTD_ACT_GELU) {
#pragma unroll
              for (int e = 0; e < 8; e += 2) gelu_fast2(f[e], f[e + 1]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = tc_act<ACT>(f[e]);
            }
            store8(out + (int64_t)row * p.ldo + col0 + j, f);
          }
        }
      }
    }
  }

  bw::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    bw::tc_fence_after();
    bw::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int ACT, typename OT, bool RESIDUAL>
static int tc_launch(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* B, int64_t ldb,
                     TcParams p, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  auto kern = tc_gemm_kernel<BN, ACT, OT, RESIDUAL>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("tc_gemm: cannot opt in to %d B of shared memory: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.M}, str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N}, str[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {64, (uint32_t)BN};
    int rc = make_tmap(&tmB, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  p.tiles_m = cdiv(p.M, 128);
  p.tiles_n = cdiv(p.N, BN);
  int tiles = p.tiles_m * p.tiles_n;
  int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, 384, Cfg::SMEM_BYTES, s>>>(tmA, tmB, p);
  return 0;
}

template <int ACT, typename OT, bool RESIDUAL>
static int tc_dispatch_bn(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* B, int64_t ldb,
                          const TcParams& p, cudaStream_t s) {
  // widest tile that divides the work sensibly; N <= 32 keeps a single narrow tile
  if (p.N <= 32) return tc_launch<32, ACT, OT, RESIDUAL>(A, lda, B, ldb, p, s);
  if (p.N <= 64) return tc_launch<64, ACT, OT, RESIDUAL>(A, lda, B, ldb, p, s);
  if (p.N % 256 == 0 && (int64_t)cdiv(p.M, 128) * (p.N / 256) >= sm_count())
    return tc_launch<256, ACT, OT, RESIDUAL>(A, lda, B, ldb, p, s);
  return tc_launch<128, ACT, OT, RESIDUAL>(A, lda, B, ldb, p, s);
}

int tc_linear(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* B, int64_t ldb, const float* bias,
              void* out, int M, int N, int K, int64_t ldo, int dtype_out, int act, cudaStream_t s) {
  TcParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = out; p.ldo = ldo; p.rows_per_sample = 1;
  if (dtype_out == DGTD_BF16) {
    if (act == DGTD_ACT_GELU) return tc_dispatch_bn<DGTD_ACT_GELU, __nv_bfloat16, false>(A, lda, B, ldb, p, s);
    if (act == DGTD_ACT_RELU) return tc_dispatch_bn<DGTD_ACT_RELU, __nv_bfloat16, false>(A, lda, B, ldb, p, s);
    return tc_dispatch_bn<DGTD_ACT_NONE, __nv_bfloat16, false>(A, lda, B, ldb, p, s);
  }
  if (act == DGTD_ACT_GELU) return tc_dispatch_bn<DGTD_ACT_GELU, float, false>(A, lda, B, ldb, p, s);
  if (act == DGTD_ACT_RELU) return tc_dispatch_bn<DGTD_ACT_RELU, float, false>(A, lda, B, ldb, p, s);
  return tc_dispatch_bn<DGTD_ACT_NONE, float, false>(A, lda, B, ldb, p, s);
}

int tc_linear_residual(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* B, int64_t ldb,
                       const float* bias, const float* gamma, const float* keep, int rows_per_sample,
                       const float* residual, float* out, int M, int N, int K, cudaStream_t s) {
  TcParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.gamma = gamma; p.keep = keep;
  p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
  p.residual = residual; p.out = out; p.ldo = N;
  return tc_dispatch_bn<DGTD_ACT_NONE, float, true>(A, lda, B, ldb, p, s);
}

}  // namespace dgtd

using namespace dgtd;

extern "C" {

int dgtd_linear_fwd(const void* a, const void* w, const float* bias, void* out, int M, int N, int K,
                    int ldo, int dtype_in, int dtype_out, int act, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(a && w && out, "linear: null pointer");
  DGTD_CHECK_ARG(M > 0 && N > 0 && K > 0 && ldo >= N, "linear: bad shape M=%d N=%d K=%d ldo=%d", M, N, K, ldo);
  DGTD_CHECK_ARG(act >= 0 && act <= 2, "linear: bad activation %d", act);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == DGTD_BF16) {
    DGTD_CHECK_ARG(K % 8 == 0 && N % 8 == 0 && ldo % 8 == 0, "linear(bf16): K, N, ldo must be multiples of 8");
    int rc = tc_linear((const __nv_bfloat16*)a, K, (const __nv_bfloat16*)w, K, bias, out, M, N, K, ldo,
                       dtype_out, act, s);
    if (rc) return rc;
    DGTD_LAUNCH_CHECK("linear(tcgen05)");
    return 0;
  }
  DGTD_CHECK_ARG(dtype_in == DGTD_F32, "linear: bad dtype_in %d", dtype_in);
  DGTD_CHECK_ARG(K % 4 == 0 && N % 4 == 0 && ldo % 4 == 0, "linear(fp32): K, N, ldo must be multiples of 4");
  RowMajorLoader al{(const float*)a, K, 0, M, K};
  RowMajorLoader bl{(const float*)w, K, 0, N, K};
#define DGTD_SIMT_LIN(OT, ACT)                                        \
  {                                                                   \
    EpiBiasAct<OT, ACT> ep{(OT*)out, bias, ldo};                      \
    launch_simt_gemm<true, true>(al, bl, ep, M, N, K, 1, s);          \
  }
  if (dtype_out == DGTD_F32) {
    if (act == DGTD_ACT_GELU) DGTD_SIMT_LIN(float, DGTD_ACT_GELU)
    else if (act == DGTD_ACT_RELU) DGTD_SIMT_LIN(float, DGTD_ACT_RELU)
    else DGTD_SIMT_LIN(float, DGTD_ACT_NONE)
  } else {
    if (act == DGTD_ACT_GELU) DGTD_SIMT_LIN(__nv_bfloat16, DGTD_ACT_GELU)
    else if (act == DGTD_ACT_RELU) DGTD_SIMT_LIN(__nv_bfloat16, DGTD_ACT_RELU)
    else DGTD_SIMT_LIN(__nv_bfloat16, DGTD_ACT_NONE)
  }
#undef DGTD_SIMT_LIN
  DGTD_LAUNCH_CHECK("linear(fp32)");
  return 0;
}

int dgtd_linear_residual_fwd(const void* a, const void* w, const float* bias, const float* gamma,
                             const float* keep, int rows_per_sample, const float* residual,
                             float* out, int M, int N, int K, int dtype_in, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(a && w && bias && residual && out, "linear_residual: null pointer");
  DGTD_CHECK_ARG(M > 0 && N > 0 && K > 0, "linear_residual: bad shape");
  DGTD_CHECK_ARG(!keep || rows_per_sample > 0, "linear_residual: rows_per_sample required with keep");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == DGTD_BF16) {
    DGTD_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "linear_residual(bf16): K, N must be multiples of 8");
    int rc = tc_linear_residual((const __nv_bfloat16*)a, K, (const __nv_bfloat16*)w, K, bias, gamma, keep,
                                rows_per_sample, residual, out, M, N, K, s);
    if (rc) return rc;
    DGTD_LAUNCH_CHECK("linear_residual(tcgen05)");
    return 0;
  }
  DGTD_CHECK_ARG(dtype_in == DGTD_F32, "linear_residual: bad dtype_in %d", dtype_in);
  DGTD_CHECK_ARG(K % 4 == 0 && N % 4 == 0, "linear_residual(fp32): K, N must be multiples of 4");
  RowMajorLoader al{(const float*)a, K, 0, M, K};
  RowMajorLoader bl{(const float*)w, K, 0, N, K};
  EpiResidual ep{out, bias, gamma, keep, residual, rows_per_sample > 0 ? rows_per_sample : 1, N};
  launch_simt_gemm<true, true>(al, bl, ep, M, N, K, 1, s);
  DGTD_LAUNCH_CHECK("linear_residual(fp32)");
  return 0;
}

}  // extern "C"
