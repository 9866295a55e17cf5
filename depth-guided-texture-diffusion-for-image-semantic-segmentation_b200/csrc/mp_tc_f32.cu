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
// X: the halo rows arrive by TMA as fp32 ({128 ch, 24 px, 2 rows} boxes = 24 KB, out-of-image pixels zero-filled by
// the unit) and are split IN PLACE: 4 bytes of fp32 become 2 + 2 bytes of (hi, lo), so a stage holds first the raw
// box and then both planes in the MN-major SWIZZLE_128B layout the MMA reads.  The converter threads of a stage read
// their six 16-byte pieces, meet on a named barrier (every raw byte is in a register), and write the planes; a chunk
// of two halo rows is issued as two 128-channel halves.  Two groups of 8 converter warps work on alternate stages.
//
// A: the weights operand is rebuilt per chunk in 3 rotating stages (hi and lo planes, SWIZZLE_32B K-major).
//
// What paces it (profiles/r2_ncu_w1_diffusion.md: ablation of every role, then a clock64 timeline of every role of one
// CTA): not HBM, L2, the tensor pipe (34 % active) or shared-memory bandwidth, but the LATENCY of the role chains.
// A warp's instruction takes ~8-10 clocks to issue in this 32-warp CTA, a poll of a ready mbarrier ~200, so
//   * one builder thread per (pixel, plane) walking all seven chunks (v1) took 3500 clocks per chunk -> the builders
//     are two groups owning the even / the odd chunks (two chunks are built concurrently), a thread computes both
//     planes of its taps, reads them from the raw-weights slot when it scatters them instead of holding 25 registers
//     (no spills under the 64-register cap; the slot is double-buffered so the next tile's TMA overlaps), and a
//     half-warp zero-fills its 16 rows cooperatively (512 contiguous bytes per K step, no 2-way bank conflict);
//   * one converter group handling every stage in turn took ~1500 clocks per stage WHATEVER the work per thread (8
//     warps x 6 pieces and 16 warps x 3 pieces measure the same: barrier, LDS, barrier, convert, STS, proxy fence,
//     arrive are fixed latencies) -> two groups convert two stages concurrently;
//   * what remains is the loop latency of an X stage -- TMA ~1100-1800 clocks for a 24 KB box, conversion ~1400, MMA
//     batch ~1100, release ~300 -- over the three stages that fit next to A, the weights slots and the staging tiles.
//
// Warp roles (1024 threads, 1 CTA / SM, persistent over (tile, 256-channel block) items):
//   w0 raw-weights TMA ({16 px, 8 rows, 49 taps} box per tile, two slots)   w1 MMA issuer (2 x 9 MMAs per chunk)
//   w2 TMEM allocator   w3 X TMA producer   w4-7 epilogue (TMEM -> fp32 staging -> 4-D TMA store)
//   w8-15 A builders (chunk parity, row pair; both planes)   w16-31 X converters (two groups of 8)
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
constexpr int NH = 128;                        // channels per X stage: a chunk is issued as NB / NH halves
constexpr int X_HALF = (NH / 64) * XBLK;       // one plane of a stage, 12 KB
constexpr int X_STAGE = 2 * X_HALF, NX = 3;    // = CK * NH * 4: the raw fp32 box and the two planes have the same size
constexpr int W_BYTES = 49 * TH * TW * 4, W_SLOT = 25600, NW = 2;
constexpr int STG = 4 * 4096;
constexpr int OFF_X = NA * A_STAGE;
constexpr int OFF_W = OFF_X + NX * X_STAGE;
constexpr int OFF_STG = OFF_W + NW * W_SLOT;
constexpr int OFF_BAR = OFF_STG + STG;
constexpr int NBARS = 3 * NX + 2 * NA + 2 * NW + 4;
constexpr int SMEM = OFF_BAR + NBARS * 8 + 16 + 1024;
constexpr int THREADS = 1024;
constexpr int NCONV = 256, NBUILD = 256;   // threads of ONE converter group / of all builders
static_assert(SMEM <= 232448, "shared memory budget");

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(bw::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ uint32_t a_off(uint32_t m, uint32_t kl) {   // SWIZZLE_32B K-major, see mp_tc.cu
  return (kl >> 4) * 4096 + m * 32 + (((((kl >> 3) & 1)) ^ ((m >> 2) & 1)) << 4) + ((kl & 7) << 1);
}
__host__ __device__ constexpr uint32_t idesc() {   // bf16 x bf16 -> fp32, A K-major, B MN-major, M 128, N 128
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Params {
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

// ring position: stage index and the parity of the pass over the ring
struct Ring {
  uint32_t s = 0, ph = 0;
  template <int N>
  __device__ __forceinline__ void step() {
    if (++s == N) { s = 0; ph ^= 1; }
  }
};

__global__ void __launch_bounds__(THREADS, 1)
mp_tc_f32_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmX, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sX = smem + OFF_X;
  uint8_t* sW = smem + OFF_W;
  uint8_t* sStg = smem + OFF_STG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* x_raw = bars;          // TMA box landed (fp32)
  uint64_t* x_full = x_raw + NX;   // planes written
  uint64_t* x_empty = x_full + NX;
  uint64_t* a_full = x_empty + NX;
  uint64_t* a_empty = a_full + NA;
  uint64_t* w_full = a_empty + NA;
  uint64_t* w_empty = w_full + NW;
  uint64_t* t_full = w_empty + NW;
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
      bw::mbar_init(&x_raw[i], 1);
      bw::mbar_init(&x_full[i], NCONV / 32);   // one arrival per converter warp
      bw::mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < NA; ++i) {
      bw::mbar_init(&a_full[i], 8);            // a chunk is built by ONE group: 4 warps x 2 half-warps (= tile rows)
      bw::mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < NW; ++i) {
      bw::mbar_init(&w_full[i], 1);
      bw::mbar_init(&w_empty[i], NBUILD / 32);
    }
    for (int i = 0; i < 2; ++i) {
      bw::mbar_init(&t_full[i], 1);
      bw::mbar_init(&t_empty[i], 4);
    }
    bw::fence_mbar_init();
  }
  if (warp == 2) bw::tmem_alloc(tmem_slot, 512);
  bw::tc_fence_before();
  __syncthreads();
  bw::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== raw weights of the tile: one TMA box {16 px, 8 rows, 49 taps}, two slots =====================
    for (int item = 0; item < num_items; ++item) {
      const Item it = decode_item(p, item);
      const int slot = item & 1;
      bw::mbar_wait_relaxed(&w_empty[slot], ((item >> 1) & 1) ^ 1, 1000);
      if (bw::elect_one()) {
        bw::mbar_arrive_expect_tx(&w_full[slot], W_BYTES);
        bw::tma_load_4d(&tmW, &w_full[slot], sW + slot * W_SLOT, it.tx * TW, it.ty * TH, 0, it.img);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ===================== X producer: one fp32 box {128 ch, 24 px, 2 rows} per (chunk, half).  (A TMA L2 prefetch of the
    // box eight stages ahead was measured twice: no effect -- the loads are L2 hits or bandwidth-paced anyway.) ==========
    constexpr int PER = NCHUNK * (NB / NH);   // x-items per (tile, channel block)
    Ring rx;
    for (int item = 0; item < num_items; ++item) {
      const Item cur = decode_item(p, item);
#pragma unroll 1
      for (int q = 0; q < PER; ++q) {
        bw::mbar_wait(&x_empty[rx.s], rx.ph ^ 1);
        if (bw::elect_one()) {
          bw::mbar_arrive_expect_tx(&x_raw[rx.s], X_STAGE);
          bw::tma_load_4d(&tmX, &x_raw[rx.s], sX + rx.s * X_STAGE, cur.nb * NB + (q & 1) * NH, cur.tx * TW - 3,
                          cur.ty * TH - 3 + 2 * (q >> 1), cur.img);
        }
        __syncwarp();
        rx.step<NX>();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: per chunk and 128-channel half, 3 products x 3 K steps =====================
    constexpr uint32_t IDESC = idesc();
    const uint64_t da0 = bw::umma_smem_desc_kmajor(bw::smem_u32(sA), 32);
    const uint64_t db0 = bw::umma_smem_desc_mnmajor_sw128(bw::smem_u32(sX), XBLK, 1024);
    Ring ra, rx;
    for (int item = 0; item < num_items; ++item) {
      const int as = item & 1;
      bw::mbar_wait(&t_empty[as], ((item >> 1) & 1) ^ 1);
      const uint32_t d_tmem = tmem_base + as * NB;
#pragma unroll 1
      for (int j = 0; j < NCHUNK; ++j) {
#pragma unroll 1
        for (int hf = 0; hf < NB / NH; ++hf) {
          bw::mbar_wait(&x_full[rx.s], rx.ph);
          if (hf == 0) bw::mbar_wait(&a_full[ra.s], ra.ph);
          bw::tc_fence_after();
          if (bw::elect_one()) {
            const uint64_t da = da0 + (uint64_t)((ra.s * A_STAGE) >> 4), db = db0 + (uint64_t)((rx.s * X_STAGE) >> 4);
            constexpr uint64_t ALO = A_HALF >> 4, XLO = X_HALF >> 4;
            const uint32_t d = d_tmem + hf * NH;
#pragma unroll
            for (int k = 0; k < CK / 16; ++k) {
              const uint64_t ak = da + (uint64_t)(k * (4096 >> 4)), xk = db + 128u * k;
              bw::umma_bf16(d, ak, xk, IDESC, (j | k) != 0);   // A_hi X_hi
              bw::umma_bf16(d, ak, xk + XLO, IDESC, 1);        // A_hi X_lo
              bw::umma_bf16(d, ak + ALO, xk, IDESC, 1);        // A_lo X_hi
            }
            bw::umma_commit(&x_empty[rx.s]);
            if (hf == NB / NH - 1) {
              bw::umma_commit(&a_empty[ra.s]);
              if (j == NCHUNK - 1) bw::umma_commit(&t_full[as]);
            }
          }
          __syncwarp();
          rx.step<NX>();
        }
        ra.step<NA>();
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
      bw::mbar_wait_relaxed(&t_full[as], (item >> 1) & 1, 300);
      bw::tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + as * NB;
#pragma unroll 1
      for (int c = 0; c < NB / 32; ++c) {
        uint32_t v[32];
        bw::tmem_ld_32x32(t0 + c * 32, v);
        bw::tmem_ld_wait();
        if (c == NB / 32 - 1) {   // all TMEM reads of this accumulator have landed: hand it back
          bw::tc_fence_before();
          __syncwarp();
          if (lane == 0) bw::mbar_arrive(&t_empty[as]);
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
    // ===================== A builders: thread = (chunk parity, pixel), both planes =====================
    // warp 8 + 4 g + r builds tile rows r and r + 4 -- one per half-warp -- in the chunks
    // j = g, g + 2, ...  A tile row py meets halo row hy = py + ky; chunk j holds halo rows 2j and 2j + 1, so the row
    // contributes at most two tap rows ky = 2j + e - py (e = 0, 1) to it: uniform per half-warp.
    const int gid = warp - 8;
    const int grp = gid >> 2;
    const int hl = lane & 15;
    const int py = (gid & 3) + 4 * (lane >> 4), px = hl;
    const int m = py * 16 + px;
    const uint32_t half_mask = lane < 16 ? 0x0000ffffu : 0xffff0000u;
    uint32_t off_e[7], off_o[7];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      off_e[kx] = a_off((uint32_t)m, (uint32_t)(px + kx));
      off_o[kx] = a_off((uint32_t)m, (uint32_t)(24 + px + kx));
    }
    const uint32_t plane_base = bw::smem_u32(sA);
    const uint32_t zoff = (uint32_t)py * 512 + (uint32_t)hl * 16;   // this lane's share of the half-warp's 16 rows
    Ring ra;                                                        // position of chunk 0 of the current tile
    for (int item = 0; item < num_items; ++item) {
      const int slot = item & 1;
      bw::mbar_wait_relaxed(&w_full[slot], (item >> 1) & 1, 200);
      const float* wsrc = reinterpret_cast<const float*>(sW + slot * W_SLOT) + py * TW + px;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int k = 0; k < 48; k += 4) {
        s0 += wsrc[k * 128]; s1 += wsrc[(k + 1) * 128]; s2 += wsrc[(k + 2) * 128]; s3 += wsrc[(k + 3) * 128];
      }
      const float inv = 1.0f / (((s0 + s1) + (s2 + s3)) + wsrc[48 * 128] + p.eps);
      Ring rc = ra;
      if (grp) rc.step<NA>();
#pragma unroll 1
      for (int j = grp; j < NCHUNK; j += 2) {
        bw::mbar_wait_relaxed(&a_empty[rc.s], rc.ph ^ 1, 200);   // the MMAs that read this stage three chunks ago are done
        const uint32_t base = plane_base + rc.s * A_STAGE;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + b * 4096 + zoff), "r"(0u) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + b * 4096 + zoff + 256), "r"(0u) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + A_HALF + b * 4096 + zoff), "r"(0u) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + A_HALF + b * 4096 + zoff + 256), "r"(0u) : "memory");
        }
        __syncwarp(half_mask);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ky = 2 * j + e - py;
          if ((unsigned)ky < 7u) {
            const float* wrow = wsrc + ky * 7 * 128;
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
              const float v = wrow[kx * 128] * inv;
              const __nv_bfloat16 hb = __float2bfloat16_rn(v);
              const __nv_bfloat16 lb = __float2bfloat16_rn(v - __bfloat162float(hb));
              const uint32_t addr = base + (e ? off_o[kx] : off_e[kx]);
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short*>(&hb)) : "memory");
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr + A_HALF), "h"(*reinterpret_cast<const unsigned short*>(&lb))
                           : "memory");
            }
          }
        }
        bw::fence_proxy_async_smem();
        __syncwarp(half_mask);
        if (hl == 0) bw::mbar_arrive(&a_full[rc.s]);
        rc.step<NA>();
        rc.step<NA>();
      }
      __syncwarp();
      if (lane == 0) bw::mbar_arrive(&w_empty[slot]);   // raw weights consumed: the producer may fetch the tile after next
#pragma unroll
      for (int i = 0; i < NCHUNK % NA; ++i) ra.step<NA>();   // 7 chunks = 2 full passes of the ring + 1
      ra.ph ^= (NCHUNK / NA) & 1;
    }
  } else if (warp >= 16) {
    // ===================== X converters: raw fp32 box -> (hi, lo) bf16 planes, in place =====================
    // Two groups of 8 warps convert alternate stages concurrently: a stage's conversion is a chain of fixed latencies
    // (barrier, LDS, barrier, convert, STS, proxy fence, arrive: ~1500 clocks whatever the work per thread, r2 trace),
    // so it is the number of stages in conversion that sets the rate.  Warp cw of a group owns K rows cw + 8 i (i < 6)
    // of the stage (K row = halo pixel: 128 channels = 512 B = one 16-byte piece per lane).  Piece (k, pc) goes to
    // 64-channel block pc >> 4, row k, 16-byte column ((pc & 15) >> 1) ^ (k & 7), 8-byte half pc & 1; k & 7 is the same
    // for a thread's six rows.
    const int cgrp = (warp - 16) >> 3;
    const int cid = (threadIdx.x - 512) & 255;
    const uint32_t k0 = (uint32_t)cid >> 5, pc = (uint32_t)cid & 31;
    const uint32_t src0 = bw::smem_u32(sX) + (uint32_t)cid * 16;
    const uint32_t dst0 = bw::smem_u32(sX) + (pc >> 4) * XBLK + k0 * 128 + ((((pc & 15) >> 1) ^ k0) << 4) + (pc & 1) * 8;
    const int total = num_items * NCHUNK * (NB / NH);
    const bool leader = ((warp - 16) & 7) == 0;
    Ring rx;
    if (cgrp) rx.step<NX>();
    for (int xi = cgrp; xi < total; xi += 2) {
      if (leader) {   // one poller per group; the other warps sleep on a hardware barrier (no issue slots)
        // A stage is converted by the two groups in turn (3 stages, stride 2).  A parity wait can tell "phase n done" from
        // "phase n running" but not from "phase n - 1 running": before waiting for THIS use's box, wait until the other
        // group has finished the previous use of the stage (x_full of one ring pass ago; passes at once on the first pass).
        bw::mbar_wait(&x_full[rx.s], rx.ph ^ 1);
        bw::mbar_wait(&x_raw[rx.s], rx.ph);
      }
      if (cgrp) asm volatile("bar.sync 4, %0;" ::"n"(NCONV) : "memory");   // hardware barrier (no issue slots)
      else asm volatile("bar.sync 2, %0;" ::"n"(NCONV) : "memory");
      float4 v[6];
#pragma unroll
      for (int i = 0; i < 6; ++i)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                     : "r"(src0 + rx.s * X_STAGE + (uint32_t)i * 4096)
                     : "memory");
      if (cgrp) asm volatile("bar.sync 3, %0;" ::"n"(NCONV) : "memory");   // every raw byte of the stage is in a register
      else asm volatile("bar.sync 1, %0;" ::"n"(NCONV) : "memory");
      const uint32_t sbase = dst0 + rx.s * X_STAGE;
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
      if (lane == 0) bw::mbar_arrive(&x_full[rx.s]);
      rx.step<NX>();
      rx.step<NX>();
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
    const uint32_t box[4] = {NH, PW, 2, 1};
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
