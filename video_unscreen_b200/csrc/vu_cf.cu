// ColorFilteringAgent.get_alpha_by_gmm (unscreen/colorfiltering/agent.py:
// 232-257) with the six 1-D Gaussian mixtures folded into 256-entry float32
// tables (the tables are built on the host with the
// reference's own torch op sequence, so table entries are bitwise what the
// reference computes per pixel).  Per pixel:
//   bg = ((1*lutbH[h])*lutbS[s])*lutbV[v]        float32, left to right
//   fg likewise; bg = pow(bg, 1/3.f), fg = pow(fg, 1/3.f)
//   p  = fg / ((bg + fg) + 1e-6f);  alpha = u8(clip(p*255, 0, 255))
// torch.pow(float32 tensor, python 1/3.) uses the float32 exponent
// 0.3333333432674408 through SLEEF's 1-ULP powf; here the power is taken in
// float64 and rounded once (the correctly rounded value).  No FTZ: the
// products reach 1e-39.
#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ float pow_third(float x) {
  if (x == 0.f) return 0.f;
  return (float)pow((double)x, 0.3333333432674408);
}

__device__ __forceinline__ int cf_alpha_px(const float* lut, int h, int s, int v) {
  const float bg = __fmul_rn(__fmul_rn(__fmul_rn(1.f, lut[h]), lut[256 + s]), lut[512 + v]);
  const float fg = __fmul_rn(__fmul_rn(__fmul_rn(1.f, lut[768 + h]), lut[1024 + s]), lut[1280 + v]);
  const float pb = pow_third(bg), pf = pow_third(fg);
  const float den = __fadd_rn(__fadd_rn(pb, pf), 1e-6f);
  const float p = __fmul_rn(__fdiv_rn(pf, den), 255.f);
  // np.clip(..., 0, 255).astype(uint8): truncation toward zero
  return (int)fminf(fmaxf(p, 0.f), 255.f);
}

__global__ void __launch_bounds__(THREADS) cf_alpha_kernel(const uint8_t* __restrict__ hsv, int64_t npix, const float* __restrict__ luts,
                                                           uint8_t* __restrict__ alpha) {
  __shared__ float lut[6 * 256];
  for (int i = threadIdx.x; i < 6 * 256; i += THREADS) lut[i] = luts[i];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride)
    alpha[p] = (uint8_t)cf_alpha_px(lut, __ldg(hsv + 3 * p), __ldg(hsv + 3 * p + 1), __ldg(hsv + 3 * p + 2));
}

// one CTA per (h, s) row of 256 v entries
__global__ void __launch_bounds__(256) cf_lut3d_kernel(const float* __restrict__ luts, uint8_t* __restrict__ lut3d) {
  __shared__ float lut[6 * 256];
  for (int i = threadIdx.x; i < 6 * 256; i += 256) lut[i] = luts[i];
  __syncthreads();
  for (int row = blockIdx.x; row < 180 * 256; row += gridDim.x) {
    const int h = row >> 8, s = row & 255;
    lut3d[(int64_t)row * 256 + threadIdx.x] = (uint8_t)cf_alpha_px(lut, h, s, threadIdx.x);
  }
}

__global__ void __launch_bounds__(THREADS) cf_alpha_lut3d_kernel(const uint8_t* __restrict__ hsv, int64_t npix, const uint8_t* __restrict__ lut3d,
                                                                 uint8_t* __restrict__ alpha) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
    const int h = min((int)__ldg(hsv + 3 * p), 179), s = __ldg(hsv + 3 * p + 1), v = __ldg(hsv + 3 * p + 2);
    alpha[p] = __ldg(lut3d + (((int64_t)h << 16) | (s << 8) | v));
  }
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_cf_alpha_u8(const uint8_t* hsv, int64_t npix, const float* luts, uint8_t* alpha, vu_stream_t stream) {
  VU_REQUIRE(hsv && luts && alpha && npix >= 0);
  if (npix == 0) return VU_OK;
  cf_alpha_kernel<<<grid_for(npix, THREADS, 8), THREADS, 0, S(stream)>>>(hsv, npix, luts, alpha);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_cf_build_lut3d(const float* luts, uint8_t* lut3d, vu_stream_t stream) {
  VU_REQUIRE(luts && lut3d);
  cf_lut3d_kernel<<<device_sms() * 8, 256, 0, S(stream)>>>(luts, lut3d);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_cf_alpha_lut3d_u8(const uint8_t* hsv, int64_t npix, const uint8_t* lut3d, uint8_t* alpha, vu_stream_t stream) {
  VU_REQUIRE(hsv && lut3d && alpha && npix >= 0);
  if (npix == 0) return VU_OK;
  cf_alpha_lut3d_kernel<<<grid_for(npix, THREADS, 8), THREADS, 0, S(stream)>>>(hsv, npix, lut3d, alpha);
  VU_RETURN_LAUNCH();
}
