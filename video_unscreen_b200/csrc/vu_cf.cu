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

// ---------------------------------------------------------------------------------
// vu_cf_samples: the training samples of the mixtures, gathered on the device
// (colorfiltering/agent.py:139-141, 165-167, 192-194 and the histogram of :142-143).
//
//   samples = channel[selection]                      row-major order
//   if len(samples) > max_samples: samples = samples[:: len(samples) // max_samples]
//
// for the three HSV channels over ONE selection
//   selection(p) = mask_test(mask[p]) && prior_test(H[p])
//   mask_test: mask < 128 (op 0) or mask > 128 (op 1);  prior_test: none, lo < H < hi, or its negation
// plus the 256-bin histogram of the strided H samples (get_color_prior).  Order-exact: the rank of a selected pixel
// is (selected pixels in the rows above) + (selected pixels to its left), computed with ballots and a scan.
// Three small launches: per-row counts, scan of the row counts (one CTA), gather.
// ---------------------------------------------------------------------------------
namespace vu {
namespace {

struct SelParams {
  int mask_op;            // 0: mask < 128, 1: mask > 128
  int prior_mode;         // 0: none, 1: lo < H < hi, 2: !(lo < H < hi)
  int lo, hi;
};
__device__ __forceinline__ bool selected(const SelParams& sp, unsigned mask, unsigned hval) {
  const bool m = sp.mask_op == 0 ? mask < 128u : mask > 128u;
  if (sp.prior_mode == 0) return m;
  const bool in = (int)hval > sp.lo && (int)hval < sp.hi;
  return m && (sp.prior_mode == 1 ? in : !in);
}

constexpr int ST = 256;
__global__ void __launch_bounds__(ST) sel_rowcount_kernel(const uint8_t* __restrict__ hsv, const uint8_t* __restrict__ mask, int w, SelParams sp,
                                                          int* __restrict__ rowcnt) {
  __shared__ int part[ST / 32];
  const int y = blockIdx.x;
  int c = 0;
  for (int x = threadIdx.x; x < w; x += ST) c += selected(sp, __ldg(mask + (int64_t)y * w + x), __ldg(hsv + ((int64_t)y * w + x) * 3));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < ST / 32; ++i) t += part[i];
    rowcnt[y] = t;
  }
}

// exclusive scan of the row counts in place; meta = {total, step, strided count}
__global__ void __launch_bounds__(1024) sel_scan_kernel(int* __restrict__ rowcnt, int h, int max_samples, int* __restrict__ meta) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < h; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < h ? rowcnt[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) >= o) s += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      int ws = wsum[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, ws, o);
        if (threadIdx.x >= o) ws += t;
      }
      wsum[threadIdx.x] = ws;
    }
    __syncthreads();
    const int before = carry + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0) + s - v;
    if (i < h) rowcnt[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int total = carry;
    const int step = total > max_samples ? total / max_samples : 1;
    meta[0] = total;
    meta[1] = step;
    meta[2] = total ? (total + step - 1) / step : 0;
  }
}

__global__ void __launch_bounds__(ST) sel_gather_kernel(const uint8_t* __restrict__ hsv, const uint8_t* __restrict__ mask, int w, SelParams sp,
                                                        const int* __restrict__ rowoff, const int* __restrict__ meta, int cap,
                                                        uint8_t* __restrict__ samples, unsigned* __restrict__ hist) {
  __shared__ int wcnt[ST / 32];
  __shared__ int running;
  const int y = blockIdx.x;
  const int step = meta[1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = rowoff[y];
  __syncthreads();
  for (int x0 = 0; x0 < w; x0 += ST) {
    const int x = x0 + threadIdx.x;
    unsigned hv = 0, sv = 0, vv = 0;
    bool sel = false;
    if (x < w) {
      const uint8_t* p = hsv + ((int64_t)y * w + x) * 3;
      hv = __ldg(p); sv = __ldg(p + 1); vv = __ldg(p + 2);
      sel = selected(sp, __ldg(mask + (int64_t)y * w + x), hv);
    }
    const unsigned b = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) wcnt[warp] = __popc(b);
    __syncthreads();
    int before = running;
    for (int i = 0; i < warp; ++i) before += wcnt[i];
    if (sel) {
      const int k = before + __popc(b & ((1u << lane) - 1u));   // rank among the selected pixels of the image
      if (k % step == 0) {
        const int idx = k / step;
        if (idx < cap) {
          samples[idx] = (uint8_t)hv;
          samples[cap + idx] = (uint8_t)sv;
          samples[2 * cap + idx] = (uint8_t)vv;
          atomicAdd(hist + hv, 1u);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int i = 0; i < ST / 32; ++i) t += wcnt[i];
      running += t;
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace vu

extern "C" size_t vu_cf_samples_workspace_bytes(int h) { return sizeof(int) * ((size_t)(h > 0 ? h : 0) + 4); }

extern "C" int vu_cf_samples(const uint8_t* hsv, const uint8_t* mask, int h, int w, int mask_op, int prior_mode, int prior_lo, int prior_hi,
                             int max_samples, uint8_t* samples, int cap, int32_t* meta3, uint32_t* hist256, void* workspace,
                             size_t workspace_bytes, vu_stream_t stream) {
  VU_REQUIRE(hsv && mask && samples && meta3 && hist256 && workspace && h > 0 && w > 0 && max_samples > 0 && cap > 0);
  VU_REQUIRE((mask_op == 0 || mask_op == 1) && prior_mode >= 0 && prior_mode <= 2);
  if (workspace_bytes < vu_cf_samples_workspace_bytes(h) || (reinterpret_cast<uintptr_t>(workspace) & 3)) return VU_ERR_WORKSPACE;
  if (cap < 2 * max_samples) return VU_ERR_INVALID_ARG;   // len // max_samples strides leave fewer than 2 * max_samples samples
  const SelParams sp{mask_op, prior_mode, prior_lo, prior_hi};
  int* rowcnt = static_cast<int*>(workspace);
  int e = record_cuda(cudaMemsetAsync(hist256, 0, 256 * sizeof(uint32_t), S(stream)));
  if (e) return e;
  sel_rowcount_kernel<<<h, ST, 0, S(stream)>>>(hsv, mask, w, sp, rowcnt);
  sel_scan_kernel<<<1, 1024, 0, S(stream)>>>(rowcnt, h, max_samples, meta3);
  sel_gather_kernel<<<h, ST, 0, S(stream)>>>(hsv, mask, w, sp, rowcnt, meta3, cap, samples, hist256);
  note_launch(2);
  VU_RETURN_LAUNCH();
}
