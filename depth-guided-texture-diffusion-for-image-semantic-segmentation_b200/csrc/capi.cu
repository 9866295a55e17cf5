// Library plumbing: version, thread-local error string, launch counter, dtype/layout kernels.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace dgtd {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = from_float<D>(to_float(src[i]));
}

// NHWC (B,h,w,ldx) -> NCHW (B,C,h,w): 32x32 smem transpose per (b, pixel tile, channel tile)
template <typename S>
__global__ void nhwc_to_nchw_kernel(const S* __restrict__ x, float* __restrict__ out, int HW, int C,
                                    int ldx) {
  __shared__ float t[32][33];
  int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (p < HW && c < C) ? to_float(x[((int64_t)b * HW + p) * ldx + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) out[((int64_t)b * C + c) * HW + p] = t[threadIdx.x][i];
  }
}

template <typename D>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, D* __restrict__ out, int HW, int C,
                                    int ldo) {
  __shared__ float t[32][33];
  int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    t[i][threadIdx.x] = (p < HW && c < C) ? x[((int64_t)b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < ldo) out[((int64_t)b * HW + p) * ldo + c] = from_float<D>(c < C ? t[threadIdx.x][i] : 0.f);
  }
}
// few channels (the RGB image in front of the first patch embed): thread = pixel, the C plane reads are coalesced
// and a warp's writes are one contiguous run of 32 * ldo values
template <typename D>
__global__ void nchw_to_nhwc_small_kernel(const float* __restrict__ x, D* __restrict__ out, int HW, int C, int ldo,
                                          int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t b = i / HW;
  const int p = (int)(i - b * HW);
  for (int c = 0; c < ldo; ++c)
    out[i * ldo + c] = from_float<D>(c < C ? x[(b * C + c) * HW + p] : 0.f);
}
}  // namespace dgtd

using namespace dgtd;

extern "C" {

int dgtd_version(void) { return DGTD_VERSION; }
const char* dgtd_last_error(void) { return g_err; }
int64_t dgtd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int dgtd_cast_fwd(const void* src, void* dst, int64_t n, int dtype_src, int dtype_dst,
                  dgtd_stream_t stream) {
  DGTD_CHECK_ARG(src && dst && n >= 0, "cast: bad args");
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype_src == DGTD_F32 && dtype_dst == DGTD_BF16)
    cast_kernel<<<blocks, 256, 0, s>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (dtype_src == DGTD_BF16 && dtype_dst == DGTD_F32)
    cast_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (dtype_src == DGTD_F32 && dtype_dst == DGTD_F32)
    cast_kernel<<<blocks, 256, 0, s>>>((const float*)src, (float*)dst, n);
  else {
    set_error("cast: unsupported dtype pair %d -> %d", dtype_src, dtype_dst);
    return -1;
  }
  DGTD_LAUNCH_CHECK("cast");
  return 0;
}

int dgtd_nhwc_to_nchw_fwd(const void* x, float* out, int B, int h, int w, int C, int ldx,
                          int dtype_in, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out && B > 0 && h > 0 && w > 0 && C > 0 && ldx >= C, "nhwc_to_nchw: bad args");
  dim3 grid(cdiv(h * w, 32), cdiv(C, 32), B), block(32, 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == DGTD_F32)
    nhwc_to_nchw_kernel<<<grid, block, 0, s>>>((const float*)x, out, h * w, C, ldx);
  else
    nhwc_to_nchw_kernel<<<grid, block, 0, s>>>((const __nv_bfloat16*)x, out, h * w, C, ldx);
  DGTD_LAUNCH_CHECK("nhwc_to_nchw");
  return 0;
}

int dgtd_nchw_to_nhwc_fwd(const float* x, void* out, int B, int h, int w, int C, int ldo,
                          int dtype_out, dgtd_stream_t stream) {
  DGTD_CHECK_ARG(x && out && B > 0 && h > 0 && w > 0 && C > 0 && ldo >= C, "nchw_to_nhwc: bad args");
  dim3 grid(cdiv(h * w, 32), cdiv(ldo, 32), B), block(32, 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (ldo <= 8) {
    const int64_t total = (int64_t)B * h * w;
    if (dtype_out == DGTD_F32)
      nchw_to_nhwc_small_kernel<<<(unsigned)cdiv(total, (int64_t)256), 256, 0, s>>>(x, (float*)out, h * w, C, ldo, total);
    else
      nchw_to_nhwc_small_kernel<<<(unsigned)cdiv(total, (int64_t)256), 256, 0, s>>>(x, (__nv_bfloat16*)out, h * w, C, ldo,
                                                                                   total);
    DGTD_LAUNCH_CHECK("nchw_to_nhwc");
    return 0;
  }
  if (dtype_out == DGTD_F32)
    nchw_to_nhwc_kernel<<<grid, block, 0, s>>>(x, (float*)out, h * w, C, ldo);
  else
    nchw_to_nhwc_kernel<<<grid, block, 0, s>>>(x, (__nv_bfloat16*)out, h * w, C, ldo);
  DGTD_LAUNCH_CHECK("nchw_to_nhwc");
  return 0;
}

}  // extern "C"
