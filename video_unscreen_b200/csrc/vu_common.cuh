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
  // HSV2BGR, per hue byte: the sector of cv2's float formula and the factor f (odd sectors) / 1 - f (even sectors)
  float hfac[256];
  unsigned char hsec[256];
};
__device__ __forceinline__ void hsv_tab_init(HsvTab& t) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    t.sdiv[i] = i ? (2 * 1044480 + i) / (2 * i) : 0;
    t.hdiv[i] = i ? (2 * 122880 + i) / (2 * i) : 0;
    // h *= 6/180; h = fmod(h, 6); sector = floor(h); f = h - sector   (SURVEY.md A.5), hue bytes >= 180 wrap
    float h = __fmul_rn((float)i, 6.0f / 180.0f);
    h = fmodf(h, 6.0f);
    int sec = (int)floorf(h);
    h = __fsub_rn(h, (float)sec);
    if ((unsigned)sec >= 6u) {
      sec = 0;
      h = 0.f;
    }
    t.hsec[i] = (unsigned char)sec;
    t.hfac[i] = (sec & 1) ? h : __fsub_rn(1.f, h);
  }
}
// the BGR2HSV half alone (kernels that never convert back)
__device__ __forceinline__ void hsv_tab_init_fwd(HsvTab& t) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    t.sdiv[i] = i ? (2 * 1044480 + i) / (2 * i) : 0;
    t.hdiv[i] = i ? (2 * 122880 + i) / (2 * i) : 0;
  }
}
// value, saturation (and the chroma d the hue needs) of one pixel
__device__ __forceinline__ void bgr2hsv_sv(int b, int g, int r, const HsvTab& t, int& s, int& v, int& d) {
  v = max(b, max(g, r));
  d = v - min(b, min(g, r));
  s = (d * t.sdiv[v] + 2048) >> 12;
}
// its hue.  The numerator by priority r, g, b of the maximum: the three candidates cost subtractions and multiply-adds
// on the (idle) FMA pipe, the choice two selects on the (busy) ALU pipe - instead of five selects for x, y and k of x - y + k*d
__device__ __forceinline__ int bgr2hsv_hue(int b, int g, int r, int v, int d, const HsvTab& t) {
  const int nr = g - b, ng = (b - r) + 2 * d, nb = (r - g) + 4 * d;
  int hh = v == r ? nr : (v == g ? ng : nb);
  hh = (hh * t.hdiv[d] + 2048) >> 12;  // arithmetic shift on a signed value
  return hh - 180 * (hh >> 31);        // hh < 0 ? hh + 180 : hh
}
__device__ __forceinline__ void bgr2hsv_px(int b, int g, int r, const HsvTab& t, int& h, int& s, int& v) {
  int d;
  bgr2hsv_sv(b, g, r, t, s, v, d);
  h = bgr2hsv_hue(b, g, r, v, d, t);
}
// The same from channel values that arrive multiplied by 4 (an IDP.4A byte extraction with the selector 4 << 8k costs what
// the plain one does): 4v and 4d ARE the byte offsets into the two tables, which saves the two address computations per
// pixel on the ALU pipe, and (4x + 8192) >> 14 == (x + 2048) >> 12 exactly.  v comes back unscaled.
__device__ __forceinline__ void bgr2hsv_px4(int b4, int g4, int r4, const HsvTab& t, int& h, int& s, int& v) {
  const int v4 = max(b4, max(g4, r4));
  const int d4 = v4 - min(b4, min(g4, r4));
  const int sd = *reinterpret_cast<const int*>(reinterpret_cast<const char*>(t.sdiv) + v4);
  const int hd = *reinterpret_cast<const int*>(reinterpret_cast<const char*>(t.hdiv) + d4);
  s = (d4 * sd + 8192) >> 14;
  const int nr = g4 - b4, ng = (b4 - r4) + 2 * d4, nb = (r4 - g4) + 4 * d4;
  int hh = v4 == r4 ? nr : (v4 == g4 ? ng : nb);
  hh = (hh * hd + 8192) >> 14;
  h = hh - 180 * (hh >> 31);
  v = v4 >> 2;
}
__device__ __forceinline__ int byte_fma4(unsigned w, int k) { return (int)__dp4a(w, 4u << (8 * k), 0u); }   // 4 * byte k of w

// uint8 <-> float32 without the conversion unit (16 lanes/clk/SM against 64 for an FADD): 2^23 + i has i in its
// mantissa, and x + 2^23 rounded toward zero has trunc(x) there (0 <= x < 2^23)
__device__ __forceinline__ float u8_to_f32(int i) { return __fadd_rn(__int_as_float(0x4B000000 | i), -8388608.0f); }
__device__ __forceinline__ int f32_trunc_nonneg(float x) { return __float_as_int(__fadd_rz(x, 8388608.0f)) & 0x007FFFFF; }

// ---- cv2 BGR2GRAY (uint8): 15-bit coefficients, SURVEY A.4 ----
__device__ __forceinline__ int bgr2gray_px(int b, int g, int r) { return (3735 * b + 19235 * g + 9798 * r + 16384) >> 15; }

// ---- cv2 HSV2BGR (uint8), float32 formula with the truncating cast of the
// whole-image SIMD path, SURVEY A.5:  tab = {v, v(1-s), v(1-s f), v(1-s(1-f))},
// (b,g,r) = tab[sector_data[sector]], sector_data = {{1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}}.
// Sector and f come from the per-hue table; only one of v(1-s f) / v(1-s(1-f)) is used by a sector (odd / even).
// s == 0 needs no branch: every entry is then v * (1 - 0) = v exactly. ----
__device__ __forceinline__ void hsv2bgr_px(int hi, int si, int vi, const HsvTab& t, int& b, int& g, int& r) {
  const float v = __fmul_rn(u8_to_f32(vi), 1.0f / 255.0f);
  const float s = __fmul_rn(u8_to_f32(si), 1.0f / 255.0f);
  const int sec = t.hsec[hi];
  const float p = __fmul_rn(v, __fsub_rn(1.f, s));
  const float x = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, t.hfac[hi])));
  // sector: 0 (p,x,v)  1 (p,v,x)  2 (x,v,p)  3 (v,x,p)  4 (v,p,x)  5 (x,p,v)
  const float bf = sec < 2 ? p : ((sec == 2 || sec == 5) ? x : v);
  const float gf = (sec == 1 || sec == 2) ? v : ((sec == 0 || sec == 3) ? x : p);
  const float rf = (sec == 0 || sec == 5) ? v : ((sec == 1 || sec == 4) ? x : p);
  b = min(255, f32_trunc_nonneg(__fmul_rn(bf, 255.f)));
  g = min(255, f32_trunc_nonneg(__fmul_rn(gf, 255.f)));
  r = min(255, f32_trunc_nonneg(__fmul_rn(rf, 255.f)));
}

// unpack / pack 4 BGR pixels held in three little-endian words
// byte k of w on the FMA pipe (IDP.4A with a one-hot selector) instead of the ALU pipe (SHF + LOP3): the per-pixel
// kernels run the ALU pipe at 85 % and the FMA pipe at 15 % (profiles/r01_cf_lowres2_wide_ncu_keys.txt); moving the
// byte extraction over took 16 % off cf_lowres2_wide
__device__ __forceinline__ int byte_fma(unsigned w, int k) { return (int)__dp4a(w, 1u << (8 * k), 0u); }

__device__ __forceinline__ void unpack12(unsigned w0, unsigned w1, unsigned w2, int (&c)[12]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c[i] = byte_fma(w0, i);
    c[4 + i] = byte_fma(w1, i);
    c[8 + i] = byte_fma(w2, i);
  }
}
__device__ __forceinline__ void pack12(const int (&c)[12], unsigned& w0, unsigned& w1, unsigned& w2) {
  w0 = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
  w1 = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
  w2 = c[8] | (c[9] << 8) | (c[10] << 16) | (c[11] << 24);
}

}  // namespace vu
