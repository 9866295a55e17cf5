// a6 at microbench scale, fp32 storage, on the TENSOR pipe with fp32-level accuracy: the banded GEMM of mp_tc.cu
//
//   Y[128 px, C] = A[128 px, 336 halo px] . X[336 halo px, C]            (cod.py:1201-1205, shared weights)
//
// evaluated as THREE bf16 products with fp32 accumulation in TMEM,
//
//   Y = A_hi X_hi + A_hi X_lo + A_lo X_hi,     v_hi = bf16(v),  v_lo = bf16(v - v_hi),
//
// i.e. both operands carry 16 significant bits (error 2^-17 per factor, the dropped A_lo X_lo term 2^-18): <= 1e-5 of
// max|y| against the float64 oracle, where one bf16 product gives 4e-3 and a TF32 product 3e-4 (the tensor core
// truncates fp32 operands to 10 mantissa bits).  3 x 6.9 = 21x the useful FLOPs of the stencil are executed -- on a
// pipe that is ~30x faster than the fp32 FMA pipe the SIMT kernel (mp_tiled.cu) saturates at 35 %.
//
// Unlike the bf16-storage kernel, X cannot go from HBM to the MMA untouched: fp32 pixels have to be split into the two
// bf16 planes.  16 converter warps read the halo rows straight from global memory (16 bytes per lane, six loads in
// flight per thread, re-issued for the next chunk as soon as a register is consumed), split them and write both planes
// in the MN-major SWIZZLE_128B layout the MMA reads.  The weights operand is rebuilt per chunk (hi by warps 8-11, lo
// by warps 12-15) in 3 rotating stages, because the resident ring of mp_tc.cu (2 x 84 KB) no longer fits next to
// X_hi / X_lo stages; a pixel's row is zero-filled and its taps scattered by the same thread.
//
// Warp roles (1024 threads, 1 CTA / SM, persistent over (tile, 256-channel block) items):
//   w0 raw-weights TMA ({16 px, 8 rows, 49 taps} box per tile)   w1 MMA issuer (9 MMAs per chunk)   w2 TMEM allocator
//   w4-7 epilogue (TMEM -> fp32 staging -> 4-D TMA store)   w8-15 A builders   w16-31 X converters
// Algorithmic HBM traffic per step: (2 * C + 49) * 4 * H * W bytes.
#include "blackwell.cuh"
#include "common.cuh"

namespace dgtd {
int sm_count();   // tc_gemm.cu

namespace mptc32 {

constexpr int TH = 8, TW = 16, PW = 24;
constexpr int NCHUNK = 7, CK = 48, NB = 256;
constexpr int A_HALF = 3 * 128 * 32;           // one plane (hi or lo) of a chunk: 3 K-step blocks of 128 x 32 B, SWIZZLE_32B
constexpr int A_STAGE = 2 * A_HALF, NA = 3;
constexpr int XBLK = CK * 128;                 // 64-channel block of a chunk: 48 K rows x 128 B
constexpr int X_HALF = (NB / 64) * XBLK;       // one plane of a chunk, 24 KB
constexpr int X_STAGE = 2 * X_HALF, NX = 2;    // (staging X per K step -- 6 stages of 16 KB -- was measured: 0.89 ms
                                               // against 0.74 ms, three times the fences and barrier round trips)
constexpr int W_BYTES = 49 * TH * TW * 4, W_SLOT = 25600;
constexpr int STG = 4 * 4096;
constexpr int OFF_X = NA * A_STAGE;
constexpr int OFF_W = OFF_X + NX * X_STAGE;
constexpr int OFF_STG = OFF_W + W_SLOT;
constexpr int OFF_BAR = OFF_STG + STG;
constexpr int NBARS = 2 * NX + 2 * NA + 2 + 4;
constexpr int SMEM = OFF_BAR + NBARS * 8 + 16 + 1024;
constexpr int THREADS = 1024;
constexpr int NCONV = 512, NBUILD = 256;

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(bw::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ uint32_t a_off(uint32_t m, uint32_t kl) {   // SWIZZLE_32B K-major, see mp_tc.cu
  return (kl >> 4) * 4096 + m * 32 + (((((kl >> 3) & 1)) ^ ((m >> 2) & 1)) << 4) + ((kl & 7) << 1);
}
__host__ __device__ constexpr uint32_t idesc() {   // bf16 x bf16 -> fp32, A K-major, B MN-major, M 128, N 256
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Params {
  const float* x;        // (n, h, w, C)
  int n, h, w, C;
  int tiles_x, tiles_y, num_tiles, nblk;
  float eps;
};

struct Item {
  int tx, ty, img, nb;
};
__device__ __forceinline__ Item decode_item(const Params& p, int item) {   // item = local index of (tile, channel block)
  const int tile = blockIdx.x + (item / p.nblk) * gridDim.x;
  Item it;
  it.nb = item % p.nblk;
  int t = tile;
  it.tx = t % p.tiles_x; t /= p.tiles_x;
  it.ty = t % p.tiles_y;
  it.img = t / p.tiles_y;
  return it;
}

__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
mp_tc_f32_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmX, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sX = smem + OFF_X;
  float* sW = reinterpret_cast<float*>(smem + OFF_W);
  uint8_t* sStg = smem + OFF_STG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + NX;
  uint64_t* a_full = x_empty + NX;
  uint64_t* a_empty = a_full + NA;
  uint64_t* w_full = a_empty + NA;
  uint64_t* w_empty = w_full + 1;
  uint64_t* t_full = w_empty + 1;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // items of this CTA: tiles blockIdx.x, + gridDim.x, ... times the channel blocks
  const int my_tiles = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int num_items = my_tiles * p.nblk;

  if (warp == 0 && lane == 0) {
    bw::prefetch_tmap(&tmW);
    bw::prefetch_tmap(&tmOut);
    bw::prefetch_tmap(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NX; ++i) {
      bw::mbar_init(&x_full[i], NCONV / 32);   // one arrival per converter warp
      bw::mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < NA; ++i) {
      bw::mbar_init(&a_full[i], NBUILD);       // every builder thread
      bw::mbar_init(&a_empty[i], 1);
    }
    bw::mbar_init(w_full, 1);
    bw::mbar_init(w_empty, NBUILD);
    for (int i = 0; i < 2; ++i) {
      bw::mbar_init(&t_full[i], 1);
      bw::mbar_init(&t_empty[i], 128);
    }
    bw::fence_mbar_init();
  }
  if (warp == 2) bw::tmem_alloc(tmem_slot, 512);
  bw::tc_fence_before();
  __syncthreads();
  bw::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== raw weights of the tile: one TMA box {16 px, 8 rows, 49 taps} =====================
    for (int item = 0; item < num_items; ++item) {
      const Item it = decode_item(p, item);
      bw::mbar_wait(w_empty, (item & 1) ^ 1);
      if (bw::elect_one()) {
        bw::mbar_arrive_expect_tx(w_full, W_BYTES);
        bw::tma_load_4d(&tmW, w_full, sW, it.tx * TW, it.ty * TH, 0, it.img);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ===================== L2 prefetcher: the converters' register pipeline holds 48 KB per SM in flight, which at
    // the ~1.8 us of a loaded DRAM round trip caps the kernel at 4 TB/s of L2->SM traffic (r2 measurement, 0.75 ms).
    // One TMA L2-prefetch of the {256 ch, 24 px, 2 rows} box per chunk, PF chunks ahead of the conversion, turns the
    // converters' loads into L2 hits without a register or a byte of shared memory. =====================
    constexpr int PF = 3;
    const int total_chunks = num_items * NCHUNK;
    for (int c = -PF; c < total_chunks - PF; ++c) {
      if (c >= 0) bw::mbar_wait(&x_empty[(uint32_t)c % NX], (((uint32_t)c / NX) & 1) ^ 1);   // pace with the consumption
      const int t = c + PF;
      const Item it = decode_item(p, t / NCHUNK);
      if (bw::elect_one())
        tma_prefetch_l2_4d(&tmX, it.nb * NB, it.tx * TW - 3, it.ty * TH - 3 + 2 * (t % NCHUNK), it.img);
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 3 products x 3 K steps per chunk =====================
    constexpr uint32_t IDESC = idesc();
    const uint64_t da0 = bw::umma_smem_desc_kmajor(bw::smem_u32(sA), 32);
    const uint64_t db0 = bw::umma_smem_desc_mnmajor_sw128(bw::smem_u32(sX), XBLK, 1024);
    uint32_t cc = 0;
    for (int item = 0; item < num_items; ++item) {
      const int as = item & 1;
      bw::mbar_wait(&t_empty[as], ((item >> 1) & 1) ^ 1);
      const uint32_t d_tmem = tmem_base + as * NB;
#pragma unroll 1
      for (int j = 0; j < NCHUNK; ++j, ++cc) {
        const uint32_t sa = cc % NA, xs = cc % NX;
        bw::mbar_wait(&x_full[xs], (cc / NX) & 1);
        bw::mbar_wait(&a_full[sa], (cc / NA) & 1);
        bw::tc_fence_after();
        if (bw::elect_one()) {
          const uint64_t da = da0 + (uint64_t)((sa * A_STAGE) >> 4), db = db0 + (uint64_t)((xs * X_STAGE) >> 4);
          constexpr uint64_t ALO = A_HALF >> 4, XLO = X_HALF >> 4;
#pragma unroll
          for (int k = 0; k < CK / 16; ++k) {
            const uint64_t ak = da + (uint64_t)(k * (4096 >> 4)), xk = db + 128u * k;
            bw::umma_bf16(d_tmem, ak, xk, IDESC, (j | k) != 0);   // A_hi X_hi
            bw::umma_bf16(d_tmem, ak, xk + XLO, IDESC, 1);        // A_hi X_lo
            bw::umma_bf16(d_tmem, ak + ALO, xk, IDESC, 1);        // A_lo X_hi
          }
          bw::umma_commit(&x_empty[xs]);
          bw::umma_commit(&a_empty[sa]);
          if (j == NCHUNK - 1) bw::umma_commit(&t_full[as]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: TMEM -> fp32 staging (32 px x 32 ch) -> 4-D TMA store =====================
    const int quad = warp & 3;
    uint8_t* tl = sStg + quad * 4096;
    const uint32_t swz = (uint32_t)(lane & 7);
    for (int item = 0; item < num_items; ++item) {
      const Item it = decode_item(p, item);
      const int as = item & 1;
      bw::mbar_wait(&t_full[as], (item >> 1) & 1);
      bw::tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + as * NB;
#pragma unroll 1
      for (int c = 0; c < NB / 32; ++c) {
        uint32_t v[32];
        bw::tmem_ld_32x32(t0 + c * 32, v);
        bw::tmem_ld_wait();
        if (c == NB / 32 - 1) {   // all TMEM reads of this accumulator have landed: hand it back
          bw::tc_fence_before();
          bw::mbar_arrive(&t_empty[as]);
        }
        if (lane == 0) bw::tma_store_wait_read<0>();   // the store that last read the staging tile is done
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(tl + lane * 128 + ((((uint32_t)q) ^ swz) << 4)) =
              make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        bw::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmOut, tl, it.nb * NB + c * 32, it.tx * TW, it.ty * TH + 2 * quad, it.img);
          bw::tma_store_commit();
        }
      }
    }
    if (lane == 0) bw::tma_store_wait_all<0>();
  } else if (warp >= 8 && warp < 16) {
    // ===================== A builders: thread = (pixel, plane); warps 8-11 build A_hi, warps 12-15 A_lo ==========
    // warp (8 + w) / (12 + w) builds plane hi / lo of tile rows w and w + 4 (same parity: the chunk transitions inside
    // the ky loop are warp-uniform, see mp_tc.cu)
    const int plane = (warp - 8) >> 2;
    const int py = ((warp - 8) & 3) + 4 * (lane >> 4), px = lane & 15;
    const int m = py * 16 + px;
    uint32_t off_e[7], off_o[7];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      off_e[kx] = a_off((uint32_t)m, (uint32_t)(px + kx));
      off_o[kx] = a_off((uint32_t)m, (uint32_t)(24 + px + kx));
    }
    const uint32_t row_base = bw::smem_u32(sA) + plane * A_HALF;
    uint32_t cc = 0;
    auto acquire = [&](uint32_t c) {   // stage of chunk counter c: wait until its previous MMAs are done, clear own row
      const uint32_t sa = c % NA;
      bw::mbar_wait(&a_empty[sa], ((c / NA) & 1) ^ 1);
      const uint32_t r = row_base + sa * A_STAGE + m * 32;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(r + b * 4096), "r"(0u) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(r + b * 4096 + 16), "r"(0u) : "memory");
      }
    };
    auto publish = [&](uint32_t c) {
      bw::fence_proxy_async_smem();
      bw::mbar_arrive(&a_full[c % NA]);
    };
    for (int item = 0; item < num_items; ++item) {
      bw::mbar_wait(w_full, item & 1);
      const float* wsrc = sW + py * TW + px;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int k = 0; k < 48; k += 4) {
        s0 += wsrc[k * 128]; s1 += wsrc[(k + 1) * 128]; s2 += wsrc[(k + 2) * 128]; s3 += wsrc[(k + 3) * 128];
      }
      const float inv = 1.0f / (((s0 + s1) + (s2 + s3)) + wsrc[48 * 128] + p.eps);
      uint32_t wp2[25];   // this plane of the normalised weights as bf16 pairs (tap 2i in the low half)
#pragma unroll
      for (int k = 0; k < 25; ++k) {
        float a = wsrc[(2 * k) * 128] * inv, b = k < 24 ? wsrc[(2 * k + 1) * 128] * inv : 0.f;
        __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
        if (plane) {
          const float2 hf = __bfloat1622float2(hi);
          hi = __floats2bfloat162_rn(a - hf.x, b - hf.y);
        }
        wp2[k] = *reinterpret_cast<const uint32_t*>(&hi);
      }
      bw::mbar_arrive(w_empty);   // raw weights consumed: the producer may fetch the next tile's
      uint32_t cur = 0;
      acquire(cc);
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const uint32_t j = (uint32_t)(py + ky) >> 1;
        while (cur < j) {
          publish(cc + cur);
          ++cur;
          acquire(cc + cur);
        }
        const uint32_t cbase = row_base + ((cc + j) % NA) * A_STAGE;
        const bool odd_row = ((py + ky) & 1) != 0;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int t = ky * 7 + kx;
          const uint32_t v = (t & 1) ? (wp2[t >> 1] >> 16) : wp2[t >> 1];
          const uint32_t addr = cbase + (odd_row ? off_o[kx] : off_e[kx]);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
        }
      }
      while (cur < NCHUNK - 1) {
        publish(cc + cur);
        ++cur;
        acquire(cc + cur);
      }
      publish(cc + NCHUNK - 1);
      cc += NCHUNK;
    }
  } else if (warp >= 16) {
    // ===================== X converters: fp32 halo rows -> (hi, lo) bf16 planes, MN-major SWIZZLE_128B ===========
    // r2 profile of the first version: 410 instructions per chunk per warp (64-bit address arithmetic, item decoding
    // and bounds checks per 16-byte piece) made this role issue-bound at 3x the MMA time.  Everything that does not
    // depend on the chunk is hoisted: a thread's six pieces are rows rg, rg+8, rg+16 of halo row 0 and of halo row 1,
    // their shared-memory slots are 1 KB apart, the item is decoded once per seven chunks.
    const int cid = threadIdx.x - 512;
    const int pc = cid & 63, rg = cid >> 6;                 // 16-byte piece (4 channels) of a pixel, row group 0..7
    const int total_chunks = num_items * NCHUNK;
    const int rowstride = p.w * p.C;                        // floats per image row (host: < 2^31)
    const int o0 = rg * p.C, o1 = (rg + 8) * p.C, o2 = (rg + 16) * p.C;
    // slot of piece i: row rg + 8i of the 48-row stage (1 KB apart); (rg + 8i) & 7 == rg, so the swizzle term is fixed
    const uint32_t dst0 = bw::smem_u32(sX) + (uint32_t)(pc >> 4) * XBLK + (uint32_t)(pc & 1) * 8 + (uint32_t)rg * 128 +
                          (((((uint32_t)(pc & 15)) >> 1) ^ (uint32_t)rg) << 4);
    // state of the chunk that is fetched NEXT
    int item_n = 0, j_n = 0, gy_n = 0;
    const float* p_n = p.x;
    bool col0 = false, col1 = false, col2 = false, live = total_chunks > 0;
    auto start_item = [&]() {
      const Item it = decode_item(p, item_n);
      const int gx0 = it.tx * TW - 3;
      gy_n = it.ty * TH - 3;
      col0 = (unsigned)(gx0 + rg) < (unsigned)p.w;
      col1 = (unsigned)(gx0 + rg + 8) < (unsigned)p.w;
      col2 = (unsigned)(gx0 + rg + 16) < (unsigned)p.w;
      p_n = p.x + ((int64_t)it.img * p.h + gy_n) * rowstride + (int64_t)gx0 * p.C + it.nb * NB + pc * 4;
    };
    auto advance = [&]() {
      if (++j_n == NCHUNK) {
        j_n = 0;
        live = ++item_n < num_items;
        if (live) start_item();
      } else {
        p_n += 2 * rowstride;
        gy_n += 2;
      }
    };
    auto fetch = [&](int i) {   // piece i of the next chunk (i < 3: halo row 0, else halo row 1)
      const bool rowok = (unsigned)(gy_n + (i >= 3 ? 1 : 0)) < (unsigned)p.h;
      const bool colok = (i % 3 == 0) ? col0 : (i % 3 == 1) ? col1 : col2;
      const float* q = p_n + (i >= 3 ? rowstride : 0) + ((i % 3 == 0) ? o0 : (i % 3 == 1) ? o1 : o2);
      return (live && rowok && colok) ? __ldg(reinterpret_cast<const float4*>(q)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 v[6];
    if (live) {
      start_item();
#pragma unroll
      for (int i = 0; i < 6; ++i) v[i] = fetch(i);
      advance();
    }
    for (int c = 0; c < total_chunks; ++c) {
      const uint32_t xs = (uint32_t)c % NX;
      bw::mbar_wait(&x_empty[xs], (((uint32_t)c / NX) & 1) ^ 1);
      const uint32_t sbase = dst0 + xs * X_STAGE;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float4 cur = v[i];
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(cur.x, cur.y), h23 = __floats2bfloat162_rn(cur.z, cur.w);
        const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
        // bf16 -> fp32 is a 16-bit shift: low half << 16, high half masked
        const float l0 = cur.x - __uint_as_float(u01 << 16), l1 = cur.y - __uint_as_float(u01 & 0xffff0000u);
        const float l2 = cur.z - __uint_as_float(u23 << 16), l3 = cur.w - __uint_as_float(u23 & 0xffff0000u);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(l0, l1), l23 = __floats2bfloat162_rn(l2, l3);
        const uint32_t addr = sbase + (uint32_t)i * 1024;
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(u01), "r"(u23) : "memory");
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr + X_HALF), "r"(*reinterpret_cast<const uint32_t*>(&l01)),
                     "r"(*reinterpret_cast<const uint32_t*>(&l23))
                     : "memory");
      }
      bw::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) bw::mbar_arrive(&x_full[xs]);
      // the next chunk's loads are issued AFTER the fence: fence.proxy.async contains a MEMBAR.ALL.CTA that waits for
      // every outstanding global load of the thread, so loads requested before it were not a prefetch at all (r2 profile:
      // the converters' long-scoreboard stalls sat on the fence).  They now fly while this thread waits for x_empty.
#pragma unroll
      for (int i = 0; i < 6; ++i) v[i] = fetch(i);
      advance();
    }
  }
  bw::tc_fence_before();
  __syncthreads();
  if (warp == 2) bw::tmem_dealloc(tmem_base, 512);
}

}  // namespace mptc32

// One fp32-storage diffusion step on the tensor pipe; 0 on success, 1 when the shape is not handled here.
int mp_tc_step_f32(const void* x, const float* weight, void* out, int n, int h, int w, int c, float eps, cudaStream_t s) {
  using namespace mptc32;
  if (c % NB != 0 || w % 4 != 0 || (int64_t)h * w * 49 >= (int64_t)1 << 31 || (int64_t)w * c >= (int64_t)1 << 29) return 1;   // TMA strides: multiples of 16 B
  CUtensorMap tmW, tmOut, tmX;
  {
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)c * 4, (uint64_t)w * c * 4, (uint64_t)h * w * c * 4};
    const uint32_t box[4] = {NB, PW, 2, 1};
    int rc = make_tmap(&tmX, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)w, (uint64_t)h, 49, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)w * 4, (uint64_t)h * w * 4, (uint64_t)49 * h * w * 4};
    const uint32_t box[4] = {TW, TH, 49, 1};
    int rc = make_tmap(&tmW, weight, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)c * 4, (uint64_t)w * c * 4, (uint64_t)h * w * c * 4};
    const uint32_t box[4] = {32, TW, 2, 1};
    int rc = make_tmap(&tmOut, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e1 = cudaFuncSetAttribute(mp_tc_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e1 != cudaSuccess) {
      set_error("message_passing_tc(f32): cannot opt in to %d B smem: %s", SMEM, cudaGetErrorString(e1));
      return -2;
    }
    configured = true;
  }
  Params p;
  p.x = (const float*)x;
  p.n = n; p.h = h; p.w = w; p.C = c;
  p.tiles_x = cdiv(w, TW);
  p.tiles_y = cdiv(h, TH);
  p.num_tiles = n * p.tiles_x * p.tiles_y;
  p.nblk = c / NB;
  p.eps = eps;
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  mp_tc_f32_kernel<<<grid, THREADS, SMEM, s>>>(tmW, tmOut, tmX, p);
  return 0;
}

}  // namespace dgtd
