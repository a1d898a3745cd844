// Colour-space kernels: cv2.cvtColor BGR2HSV / HSV2BGR / BGR2GRAY and the
// HSV in-range tests of unscreen/utils/fgfuncs.py:9-65.  Streaming,
// HBM-bound: each thread moves 4 pixels (three 32-bit words in) per step, a
// warp touches 384 contiguous bytes, grids are SM-count multiples.
#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int kThreads = 256;

struct Bgr2Hsv {
  const uint8_t* src; uint8_t* dst;
  __device__ void px(int64_t p, const HsvTab& t) const {
    int h, s, v; bgr2hsv_px(src[3 * p], src[3 * p + 1], src[3 * p + 2], t, h, s, v);
    dst[3 * p] = h; dst[3 * p + 1] = s; dst[3 * p + 2] = v;
  }
  __device__ void px4(int64_t g, const HsvTab& t) const {
    const unsigned* s4 = reinterpret_cast<const unsigned*>(src) + 3 * g;
    int c[12], o[12];
    unpack12(__ldg(s4), __ldg(s4 + 1), __ldg(s4 + 2), c);
#pragma unroll
    for (int i = 0; i < 4; ++i) bgr2hsv_px(c[3 * i], c[3 * i + 1], c[3 * i + 2], t, o[3 * i], o[3 * i + 1], o[3 * i + 2]);
    unsigned w0, w1, w2; pack12(o, w0, w1, w2);
    unsigned* d4 = reinterpret_cast<unsigned*>(dst) + 3 * g;
    d4[0] = w0; d4[1] = w1; d4[2] = w2;
  }
};

struct Hsv2Bgr {
  const uint8_t* src; uint8_t* dst;
  __device__ void px(int64_t p, const HsvTab& t) const {
    int b, g, r; hsv2bgr_px(src[3 * p], src[3 * p + 1], src[3 * p + 2], t, b, g, r);
    dst[3 * p] = b; dst[3 * p + 1] = g; dst[3 * p + 2] = r;
  }
  __device__ void px4(int64_t g, const HsvTab& t) const {
    const unsigned* s4 = reinterpret_cast<const unsigned*>(src) + 3 * g;
    int c[12], o[12];
    unpack12(__ldg(s4), __ldg(s4 + 1), __ldg(s4 + 2), c);
#pragma unroll
    for (int i = 0; i < 4; ++i) hsv2bgr_px(c[3 * i], c[3 * i + 1], c[3 * i + 2], t, o[3 * i], o[3 * i + 1], o[3 * i + 2]);
    unsigned w0, w1, w2; pack12(o, w0, w1, w2);
    unsigned* d4 = reinterpret_cast<unsigned*>(dst) + 3 * g;
    d4[0] = w0; d4[1] = w1; d4[2] = w2;
  }
};

struct Bgr2Gray {
  const uint8_t* src; uint8_t* dst;
  __device__ void px(int64_t p, const HsvTab&) const { dst[p] = bgr2gray_px(src[3 * p], src[3 * p + 1], src[3 * p + 2]); }
  __device__ void px4(int64_t g, const HsvTab&) const {
    const unsigned* s4 = reinterpret_cast<const unsigned*>(src) + 3 * g;
    int c[12];
    unpack12(__ldg(s4), __ldg(s4 + 1), __ldg(s4 + 2), c);
    unsigned w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) w |= (unsigned)bgr2gray_px(c[3 * i], c[3 * i + 1], c[3 * i + 2]) << (8 * i);
    reinterpret_cast<unsigned*>(dst)[g] = w;
  }
};

struct InRangeColor {
  const uint8_t* src; uint8_t* dst; int lo[3], hi[3];
  __device__ int test(int b, int g, int r, const HsvTab& t) const {
    int h, s, v; bgr2hsv_px(b, g, r, t, h, s, v);
    return (h >= lo[0]) & (h <= hi[0]) & (s >= lo[1]) & (s <= hi[1]) & (v >= lo[2]) & (v <= hi[2]);
  }
  __device__ void px(int64_t p, const HsvTab& t) const { dst[p] = test(src[3 * p], src[3 * p + 1], src[3 * p + 2], t); }
  __device__ void px4(int64_t g, const HsvTab& t) const {
    const unsigned* s4 = reinterpret_cast<const unsigned*>(src) + 3 * g;
    int c[12];
    unpack12(__ldg(s4), __ldg(s4 + 1), __ldg(s4 + 2), c);
    unsigned w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) w |= (unsigned)test(c[3 * i], c[3 * i + 1], c[3 * i + 2], t) << (8 * i);
    reinterpret_cast<unsigned*>(dst)[g] = w;
  }
};

struct InRangeImage {
  const uint8_t* src; const uint8_t* bg; uint8_t* dst; int64_t bg_npix; int half[3];
  __device__ int test(const int* c, const int* q, const HsvTab& t) const {
    int h, s, v, bh, bs, bv;
    bgr2hsv_px(c[0], c[1], c[2], t, h, s, v);
    bgr2hsv_px(q[0], q[1], q[2], t, bh, bs, bv);
    // torch.clamp(bg -/+ half, 10, 255), fgfuncs.py:44-45
    const int lh = min(255, max(10, bh - half[0])), uh = min(255, max(10, bh + half[0]));
    const int ls = min(255, max(10, bs - half[1])), us = min(255, max(10, bs + half[1]));
    const int lv = min(255, max(10, bv - half[2])), uv = min(255, max(10, bv + half[2]));
    return (h >= lh) & (h <= uh) & (s >= ls) & (s <= us) & (v >= lv) & (v <= uv);
  }
  __device__ void px(int64_t p, const HsvTab& t) const {
    const int64_t q = p % bg_npix;
    int c[3] = {src[3 * p], src[3 * p + 1], src[3 * p + 2]};
    int b[3] = {bg[3 * q], bg[3 * q + 1], bg[3 * q + 2]};
    dst[p] = test(c, b, t);
  }
  __device__ void px4(int64_t g, const HsvTab& t) const {
    const unsigned* s4 = reinterpret_cast<const unsigned*>(src) + 3 * g;
    const unsigned* b4 = reinterpret_cast<const unsigned*>(bg) + 3 * (g % (bg_npix >> 2));
    int c[12], q[12];
    unpack12(__ldg(s4), __ldg(s4 + 1), __ldg(s4 + 2), c);
    unpack12(__ldg(b4), __ldg(b4 + 1), __ldg(b4 + 2), q);
    unsigned w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) w |= (unsigned)test(c + 3 * i, q + 3 * i, t) << (8 * i);
    reinterpret_cast<unsigned*>(dst)[g] = w;
  }
};

template <class F>
__global__ void __launch_bounds__(kThreads) px_kernel(F f, int64_t ngroups, int64_t tail_begin, int64_t npix) {
  __shared__ HsvTab tab;
  hsv_tab_init(tab);
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) f.px4(g, tab);
  if (blockIdx.x == 0)
    for (int64_t p = tail_begin + threadIdx.x; p < npix; p += blockDim.x) f.px(p, tab);
}

inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; }

template <class F>
int launch_px(const F& f, int64_t npix, bool vector_ok, vu_stream_t stream) {
  if (npix <= 0) return VU_OK;
  const int64_t ngroups = vector_ok ? npix / 4 : 0;
  const int64_t tail = ngroups * 4;
  const int64_t work = ngroups > 0 ? ngroups : npix;
  px_kernel<F><<<grid_for(work, kThreads, 8), kThreads, 0, S(stream)>>>(f, ngroups, tail, npix);
  VU_RETURN_LAUNCH();
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_bgr2hsv_u8(const uint8_t* bgr, uint8_t* hsv, int64_t npix, vu_stream_t stream) {
  VU_REQUIRE(bgr && hsv && npix >= 0);
  return launch_px(Bgr2Hsv{bgr, hsv}, npix, aligned4(bgr) && aligned4(hsv), stream);
}

extern "C" int vu_hsv2bgr_u8(const uint8_t* hsv, uint8_t* bgr, int64_t npix, vu_stream_t stream) {
  VU_REQUIRE(bgr && hsv && npix >= 0);
  return launch_px(Hsv2Bgr{hsv, bgr}, npix, aligned4(bgr) && aligned4(hsv), stream);
}

extern "C" int vu_bgr2gray_u8(const uint8_t* bgr, uint8_t* gray, int64_t npix, vu_stream_t stream) {
  VU_REQUIRE(bgr && gray && npix >= 0);
  return launch_px(Bgr2Gray{bgr, gray}, npix, aligned4(bgr) && aligned4(gray), stream);
}

extern "C" int vu_inrange_color(const uint8_t* bgr, int64_t npix, const int32_t lo[3], const int32_t hi[3],
                                uint8_t* mask01, vu_stream_t stream) {
  VU_REQUIRE(bgr && mask01 && lo && hi && npix >= 0);
  InRangeColor f{bgr, mask01, {lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}};
  return launch_px(f, npix, aligned4(bgr) && aligned4(mask01), stream);
}

extern "C" int vu_inrange_image(const uint8_t* bgr, const uint8_t* bgimg, int64_t npix, int64_t bg_npix,
                                const int32_t half[3], uint8_t* mask01, vu_stream_t stream) {
  VU_REQUIRE(bgr && bgimg && mask01 && half && npix >= 0 && bg_npix > 0);
  InRangeImage f{bgr, bgimg, mask01, bg_npix, {half[0], half[1], half[2]}};
  const bool vec = aligned4(bgr) && aligned4(bgimg) && aligned4(mask01) && (bg_npix % 4 == 0);
  return launch_px(f, npix, vec, stream);
}

// ---- frame I/O glue: nvJPEG hands out planar RGB, the reference's arrays are interleaved BGR ----
namespace vu {
namespace {
// planes [3][npix] (R, G, B) -> pixels [npix][3] (B, G, R), or back
template <bool TO_PLANAR>
__global__ void __launch_bounds__(kThreads) planar_rgb_bgr_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t npix) {
  for (int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x; p < npix; p += (int64_t)gridDim.x * kThreads) {
    if (TO_PLANAR) {
      dst[p] = src[3 * p + 2]; dst[npix + p] = src[3 * p + 1]; dst[2 * npix + p] = src[3 * p];
    } else {
      dst[3 * p] = src[2 * npix + p]; dst[3 * p + 1] = src[npix + p]; dst[3 * p + 2] = src[p];
    }
  }
}
}  // namespace
}  // namespace vu

extern "C" int vu_planar_rgb_to_bgr(const uint8_t* src, uint8_t* dst, int64_t npix, vu_stream_t stream) {
  VU_REQUIRE(src && dst && npix >= 0);
  if (npix == 0) return VU_OK;
  vu::planar_rgb_bgr_kernel<false><<<vu::grid_for(npix, vu::kThreads, 8), vu::kThreads, 0, vu::S(stream)>>>(src, dst, npix);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_bgr_to_planar_rgb(const uint8_t* src, uint8_t* dst, int64_t npix, vu_stream_t stream) {
  VU_REQUIRE(src && dst && npix >= 0);
  if (npix == 0) return VU_OK;
  vu::planar_rgb_bgr_kernel<true><<<vu::grid_for(npix, vu::kThreads, 8), vu::kThreads, 0, vu::S(stream)>>>(src, dst, npix);
  VU_RETURN_LAUNCH();
}
