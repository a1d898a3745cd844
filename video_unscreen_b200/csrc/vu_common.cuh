// Shared device helpers of the B200 (sm_100a) video_unscreen hot path.
// Arithmetic here restates the third-party primitives the reference calls
// (SURVEY.md Appendix A); it is compiled with -fmad=false and without
// fast-math so every float op rounds once, like the CPU libraries.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vu_b200.h"

namespace vu {

int record_cuda(cudaError_t e);  // stores the message for vu_last_cuda_error
int device_sms();
void note_launch(int kernels = 1);  // feeds vu_launch_count (bench.py reports it as gpu_launches)
inline cudaStream_t S(vu_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define VU_RETURN_LAUNCH()                        \
  do {                                            \
    ::vu::note_launch();                          \
    return ::vu::record_cuda(cudaGetLastError()); \
  } while (0)
#define VU_REQUIRE(cond) \
  do {                   \
    if (!(cond)) return VU_ERR_INVALID_ARG; \
  } while (0)

// grid sizing: a multiple of the SM count, capped by the work available
inline int grid_for(int64_t work_items, int threads, int ctas_per_sm) {
  int64_t need = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)device_sms() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- 128-bit streaming loads / stores (read-once data: keep it out of L1) ----
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- cv2 BGR2HSV (uint8, H range 180): integer fixed point, SURVEY A.2 ----
// sdiv[i] = rint((255<<12)/i), hdiv[i] = rint((180<<12)/(6 i)); neither
// quotient can tie (the numerators hold too few factors of two), so
// round-half-even == round-half-up == the integer division below.
struct HsvTab {
  int sdiv[256];
  int hdiv[256];
};
__device__ __forceinline__ void hsv_tab_init(HsvTab& t) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    t.sdiv[i] = i ? (2 * 1044480 + i) / (2 * i) : 0;
    t.hdiv[i] = i ? (2 * 122880 + i) / (2 * i) : 0;
  }
}
__device__ __forceinline__ void bgr2hsv_px(int b, int g, int r, const HsvTab& t, int& h, int& s, int& v) {
  v = max(b, max(g, r));
  const int mn = min(b, min(g, r));
  const int d = v - mn;
  s = (d * t.sdiv[v] + 2048) >> 12;
  int hh = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * d) : (r - g + 4 * d));
  hh = (hh * t.hdiv[d] + 2048) >> 12;  // arithmetic shift on a signed value
  h = hh < 0 ? hh + 180 : hh;
}

// ---- cv2 BGR2GRAY (uint8): 15-bit coefficients, SURVEY A.4 ----
__device__ __forceinline__ int bgr2gray_px(int b, int g, int r) { return (3735 * b + 19235 * g + 9798 * r + 16384) >> 15; }

// ---- cv2 HSV2BGR (uint8), float32 formula with the truncating cast of the
// whole-image SIMD path, SURVEY A.5 ----
__device__ __forceinline__ void hsv2bgr_px(int hi, int si, int vi, int& b, int& g, int& r) {
  const float v = __fmul_rn((float)vi, 1.0f / 255.0f);
  float bf, gf, rf;
  if (si == 0) {
    bf = gf = rf = v;
  } else {
    const float s = __fmul_rn((float)si, 1.0f / 255.0f);
    float h = __fmul_rn((float)hi, 6.0f / 180.0f);
    h = fmodf(h, 6.0f);
    int sec = (int)floorf(h);
    h = __fsub_rn(h, (float)sec);
    if ((unsigned)sec >= 6u) {
      sec = 0;
      h = 0.f;
    }
    const float t0 = v;
    const float t1 = __fmul_rn(v, __fsub_rn(1.f, s));
    const float t2 = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, h)));
    const float t3 = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, h))));
    // sector_data = {{1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}}
    switch (sec) {
      case 0: bf = t1; gf = t3; rf = t0; break;
      case 1: bf = t1; gf = t0; rf = t2; break;
      case 2: bf = t3; gf = t0; rf = t1; break;
      case 3: bf = t0; gf = t2; rf = t1; break;
      case 4: bf = t0; gf = t1; rf = t3; break;
      default: bf = t2; gf = t1; rf = t0; break;
    }
  }
  b = min(255, max(0, (int)__fmul_rn(bf, 255.f)));
  g = min(255, max(0, (int)__fmul_rn(gf, 255.f)));
  r = min(255, max(0, (int)__fmul_rn(rf, 255.f)));
}

// unpack / pack 4 BGR pixels held in three little-endian words
__device__ __forceinline__ void unpack12(unsigned w0, unsigned w1, unsigned w2, int (&c)[12]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c[i] = (w0 >> (8 * i)) & 255;
    c[4 + i] = (w1 >> (8 * i)) & 255;
    c[8 + i] = (w2 >> (8 * i)) & 255;
  }
}
__device__ __forceinline__ void pack12(const int (&c)[12], unsigned& w0, unsigned& w1, unsigned& w2) {
  w0 = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
  w1 = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
  w2 = c[8] | (c[9] << 8) | (c[10] << 16) | (c[11] << 24);
}

}  // namespace vu
