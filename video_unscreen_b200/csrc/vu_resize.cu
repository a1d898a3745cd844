// cv2.resize on uint8: the default INTER_LINEAR in cv2's
// fixed-point form (11-bit coefficients, horizontal fraction reset at the
// borders, vertical index clamp only, two-stage >>4 / *beta>>16 / +2>>2
// vertical pass), its silent INTER_AREA substitution for exact 2x
// down-scales, and INTER_NEAREST.  Reference call sites:
// unscreen/colorfiltering/agent.py:315-316,342; unscreen/trimap/agent.py:52,59;
// unscreen/utils/imgprocess.py:36; unscreen/utils/fgfuncs.py:198.
#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int THREADS = 256;

struct Axis {
  int i0, i1, w0, w1;
};

// coefficient of one destination coordinate; `reset` = horizontal rule
__device__ __forceinline__ Axis linear_axis(int d, int dst, int src, bool reset) {
  const double inv_scale = (double)dst / (double)src;
  const double scale = 1.0 / inv_scale;
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  int i0 = (int)floorf(f);
  float fr = __fsub_rn(f, (float)i0);
  Axis a;
  if (reset) {
    if (i0 < 0) { i0 = 0; fr = 0.f; }
    if (i0 >= src - 1) { i0 = src - 1; fr = 0.f; }
    a.i0 = i0;
    a.i1 = min(i0 + 1, src - 1);
  } else {
    a.i0 = min(max(i0, 0), src - 1);
    a.i1 = min(max(i0 + 1, 0), src - 1);
  }
  a.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fr), 2048.f));
  a.w1 = __float2int_rn(__fmul_rn(fr, 2048.f));
  return a;
}

template <int C>
__global__ void __launch_bounds__(THREADS) resize_linear_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst,
                                                                int dh, int dw, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int dx = (int)(i % dw);
    const int64_t t = i / dw;
    const int dy = (int)(t % dh);
    const int64_t n = t / dh;
    const Axis ax = linear_axis(dx, dw, sw, true);
    const Axis ay = linear_axis(dy, dh, sh, false);
    const uint8_t* r0 = src + ((n * sh + ay.i0) * (int64_t)sw) * C;
    const uint8_t* r1 = src + ((n * sh + ay.i1) * (int64_t)sw) * C;
    uint8_t* o = dst + i * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int R0 = __ldg(r0 + ax.i0 * C + c) * ax.w0 + __ldg(r0 + ax.i1 * C + c) * ax.w1;
      const int R1 = __ldg(r1 + ax.i0 * C + c) * ax.w0 + __ldg(r1 + ax.i1 * C + c) * ax.w1;
      const int v = (((ay.w0 * (R0 >> 4)) >> 16) + ((ay.w1 * (R1 >> 4)) >> 16) + 2) >> 2;
      o[c] = (uint8_t)min(255, max(0, v));
    }
  }
}

template <int C>
__global__ void __launch_bounds__(THREADS) resize_area2_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst,
                                                               int dh, int dw, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int dx = (int)(i % dw);
    const int64_t t = i / dw;
    const int dy = (int)(t % dh);
    const int64_t n = t / dh;
    const uint8_t* r0 = src + ((n * sh + 2 * dy) * (int64_t)sw + 2 * dx) * C;
    const uint8_t* r1 = r0 + (int64_t)sw * C;
#pragma unroll
    for (int c = 0; c < C; ++c) dst[i * C + c] = (uint8_t)((__ldg(r0 + c) + __ldg(r0 + C + c) + __ldg(r1 + c) + __ldg(r1 + C + c) + 2) >> 2);
  }
}

template <int C>
__global__ void __launch_bounds__(THREADS) resize_nearest_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst,
                                                                 int dh, int dw, int64_t total) {
  const double ifx = 1.0 / ((double)dw / (double)sw), ify = 1.0 / ((double)dh / (double)sh);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int dx = (int)(i % dw);
    const int64_t t = i / dw;
    const int dy = (int)(t % dh);
    const int64_t n = t / dh;
    const int sx = min((int)floor(__dmul_rn((double)dx, ifx)), sw - 1);
    const int sy = min((int)floor(__dmul_rn((double)dy, ify)), sh - 1);
    const uint8_t* p = src + ((n * sh + sy) * (int64_t)sw + sx) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) dst[i * C + c] = __ldg(p + c);
  }
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_resize_linear_u8(const uint8_t* src, int n, int sh, int sw, int channels, uint8_t* dst, int dh, int dw,
                                   vu_stream_t stream) {
  VU_REQUIRE(src && dst && n >= 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0);
  if (channels != 1 && channels != 3) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  const int64_t total = (int64_t)n * dh * dw;
  if (dh == sh && dw == sw) {
    return record_cuda(cudaMemcpyAsync(dst, src, (size_t)total * channels, cudaMemcpyDeviceToDevice, S(stream)));
  }
  const int grid = grid_for(total, THREADS, 8);
  const bool area = (sw == 2 * dw && sh == 2 * dh);
  if (channels == 1) {
    if (area) resize_area2_kernel<1><<<grid, THREADS, 0, S(stream)>>>(src, sh, sw, dst, dh, dw, total);
    else resize_linear_kernel<1><<<grid, THREADS, 0, S(stream)>>>(src, sh, sw, dst, dh, dw, total);
  } else {
    if (area) resize_area2_kernel<3><<<grid, THREADS, 0, S(stream)>>>(src, sh, sw, dst, dh, dw, total);
    else resize_linear_kernel<3><<<grid, THREADS, 0, S(stream)>>>(src, sh, sw, dst, dh, dw, total);
  }
  VU_RETURN_LAUNCH();
}

extern "C" int vu_resize_nearest_u8(const uint8_t* src, int n, int sh, int sw, int channels, uint8_t* dst, int dh, int dw,
                                    vu_stream_t stream) {
  VU_REQUIRE(src && dst && n >= 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0);
  if (channels != 1 && channels != 3) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  const int64_t total = (int64_t)n * dh * dw;
  const int grid = grid_for(total, THREADS, 8);
  if (channels == 1) resize_nearest_kernel<1><<<grid, THREADS, 0, S(stream)>>>(src, sh, sw, dst, dh, dw, total);
  else resize_nearest_kernel<3><<<grid, THREADS, 0, S(stream)>>>(src, sh, sw, dst, dh, dw, total);
  VU_RETURN_LAUNCH();
}
