// Compositing family: get_fg / get_bg in
// float32 HSV arithmetic (unscreen/utils/fgfuncs.py:84-137), fused with the
// predicated background patch of the pipeline scripts; the float64 blends of
// fgfuncs.py:68-81,172-214, visualize.py:7-24 and tools/replace/replace.py:
// 74-76; the background fusion and difference gate of bg_offline.py:150-160.
// Every float op is a single IEEE rounding (__f*_rn / __d*_rn) in the
// reference's operation order, so the stages before cv2's HSV2BGR are
// bit-exact; HSV2BGR itself is the truncating whole-image variant (+-1 LSB
// against cv2 by cv2's own inconsistency).
#include <cstdlib>
#include <initializer_list>

#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ int trunc_clamp255(float x) { return f32_trunc_nonneg(fminf(fmaxf(x, 0.f), 255.f)); }

// 4 pixels per thread; alpha word holds the 4 alphas.
// BGMODE: how the background repeats under the frames (no 64-bit modulo in the pixel loop):
//   0  bg as large as the input            1  bg is ONE 4-pixel group (constant colour)
//   2  the input is gridDim.y repetitions of bg (a clip over one background image): g runs over one repetition
//   3  anything else: g % bg_groups
template <int PATCH, bool WRITE_BG, int BGMODE>
__global__ void __launch_bounds__(THREADS) get_fg_kernel(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ alpha,
                                                         const uint8_t* __restrict__ bg, int64_t ngroups, int64_t bg_groups,
                                                         uint8_t* __restrict__ fg_out, uint8_t* __restrict__ bg_out) {
  __shared__ HsvTab tab;
  __shared__ float ktab[256];   // 1 - alpha/255. for every alpha byte
  hsv_tab_init(tab);
  for (int a = threadIdx.x; a < 256; a += blockDim.x) ktab[a] = __fsub_rn(1.f, __fdiv_rn((float)a, 255.f));
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rep = BGMODE == 2 ? (int64_t)blockIdx.y * bg_groups : 0;   // first group of this repetition
  const int64_t gend = BGMODE == 2 ? bg_groups : ngroups;
  int q0[12];
  if (BGMODE == 1) {
    const unsigned* b4 = reinterpret_cast<const unsigned*>(bg);
    unpack12(__ldg(b4), __ldg(b4 + 1), __ldg(b4 + 2), q0);
  }
  for (int64_t gl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gl < gend; gl += stride) {
    const int64_t g = rep + gl;
    const unsigned* f4 = reinterpret_cast<const unsigned*>(frame) + 3 * g;
    const unsigned aw = __ldg(reinterpret_cast<const unsigned*>(alpha) + g);
    if (PATCH != VU_PATCH_NONE) {
      // a patched pixel has bg == frame, hence fg = HSV2BGR(hsv - (1 - a/255) * hsv): at alpha == 0 - patched under both
      // rules (== 0, < 128) - that is HSV2BGR(0,0,0) = black.  Whole warps of such pixels (everything outside the matte)
      // skip the arithmetic, and the loads too when the patched background is not asked for.
      const bool allzero = aw == 0u;
      if (__all_sync(__activemask(), allzero)) {
        unsigned* d4 = reinterpret_cast<unsigned*>(fg_out) + 3 * g;
        d4[0] = 0u; d4[1] = 0u; d4[2] = 0u;
        if (WRITE_BG) {
          unsigned* e4 = reinterpret_cast<unsigned*>(bg_out) + 3 * g;
          e4[0] = __ldg(f4); e4[1] = __ldg(f4 + 1); e4[2] = __ldg(f4 + 2);
        }
        continue;
      }
    }
    int c[12], q[12], o[12];
    unpack12(__ldg(f4), __ldg(f4 + 1), __ldg(f4 + 2), c);
    if (BGMODE == 1) {
#pragma unroll
      for (int i = 0; i < 12; ++i) q[i] = q0[i];
    } else {
      const int64_t gb = BGMODE == 0 ? g : (BGMODE == 2 ? gl : g % bg_groups);
      const unsigned* b4 = reinterpret_cast<const unsigned*>(bg) + 3 * gb;
      unpack12(__ldg(b4), __ldg(b4 + 1), __ldg(b4 + 2), q);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = (aw >> (8 * i)) & 255;
      const bool patch = (PATCH == VU_PATCH_ALPHA_LT128) ? (a < 128) : (PATCH == VU_PATCH_ALPHA_EQ0 ? (a == 0) : false);
      q[3 * i] = patch ? c[3 * i] : q[3 * i];
      q[3 * i + 1] = patch ? c[3 * i + 1] : q[3 * i + 1];
      q[3 * i + 2] = patch ? c[3 * i + 2] : q[3 * i + 2];
      int ih, is, iv, bh, bs, bv;
      bgr2hsv_px(c[3 * i], c[3 * i + 1], c[3 * i + 2], tab, ih, is, iv);
      bgr2hsv_px(q[3 * i], q[3 * i + 1], q[3 * i + 2], tab, bh, bs, bv);
      const float k = ktab[a];   // 1 - alpha/255.
      const int fh = trunc_clamp255(__fsub_rn(u8_to_f32(ih), __fmul_rn(k, u8_to_f32(bh))));
      const int fs = trunc_clamp255(__fsub_rn(u8_to_f32(is), __fmul_rn(k, u8_to_f32(bs))));
      const int fv = trunc_clamp255(__fsub_rn(u8_to_f32(iv), __fmul_rn(k, u8_to_f32(bv))));
      hsv2bgr_px(fh, fs, fv, tab, o[3 * i], o[3 * i + 1], o[3 * i + 2]);
    }
    unsigned w0, w1, w2;
    pack12(o, w0, w1, w2);
    unsigned* d4 = reinterpret_cast<unsigned*>(fg_out) + 3 * g;
    d4[0] = w0; d4[1] = w1; d4[2] = w2;
    if (WRITE_BG) {
      pack12(q, w0, w1, w2);
      unsigned* e4 = reinterpret_cast<unsigned*>(bg_out) + 3 * g;
      e4[0] = w0; e4[1] = w1; e4[2] = w2;
    }
  }
}

// get_fg at 16 pixels per thread: 128-bit streaming loads / stores (enough bytes in flight per thread for HBM), the
// alpha == 0 early-out per thread.  Same arithmetic as get_fg_kernel, one 4-pixel group at a time.
// groups16 = 16-pixel groups of the input; bg as in get_fg_kernel (BGMODE 0, 1, 2), counted in 16-pixel groups too.
template <int PATCH, bool WRITE_BG, int BGMODE>
__global__ void __launch_bounds__(THREADS) get_fg16_kernel(const uint4* __restrict__ frame, const uint4* __restrict__ alpha, const uint4* __restrict__ bg,
                                                           int64_t ngroups, int64_t bg_groups, uint4* __restrict__ fg_out, uint4* __restrict__ bg_out) {
  __shared__ HsvTab tab;
  __shared__ float ktab[256];   // 1 - alpha/255. for every alpha byte
  hsv_tab_init(tab);
  for (int a = threadIdx.x; a < 256; a += blockDim.x) ktab[a] = __fsub_rn(1.f, __fdiv_rn((float)a, 255.f));
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rep = BGMODE == 2 ? (int64_t)blockIdx.y * bg_groups : 0;
  const int64_t gend = BGMODE == 2 ? bg_groups : ngroups;
  int q0[12], h0[4], s0[4], v0[4];   // BGMODE 1: the constant background group and its HSV, converted once
  if (BGMODE == 1) {
    const unsigned* b4 = reinterpret_cast<const unsigned*>(bg);
    unpack12(__ldg(b4), __ldg(b4 + 1), __ldg(b4 + 2), q0);
#pragma unroll
    for (int i = 0; i < 4; ++i) bgr2hsv_px(q0[3 * i], q0[3 * i + 1], q0[3 * i + 2], tab, h0[i], s0[i], v0[i]);
  }
  for (int64_t gl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gl < gend; gl += stride) {
    const int64_t g = rep + gl;
    const uint4 av = ldg_stream16(alpha + g);
    uint4 fv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) fv[k] = ldg_stream16(frame + 3 * g + k);
    if (PATCH != VU_PATCH_NONE && (av.x | av.y | av.z | av.w) == 0u) {
      // alpha == 0 is patched under both rules (== 0, < 128): every pixel's background is the pixel itself and
      // fg = HSV2BGR(hsv - 1.0 * hsv) = black
#pragma unroll
      for (int k = 0; k < 3; ++k) stg_stream16(fg_out + 3 * g + k, make_uint4(0u, 0u, 0u, 0u));
      if (WRITE_BG) {
#pragma unroll
        for (int k = 0; k < 3; ++k) stg_stream16(bg_out + 3 * g + k, fv[k]);
      }
      continue;
    }
    uint4 qv[3];
    if (BGMODE != 1) {
      const int64_t gb = BGMODE == 0 ? g : gl;
#pragma unroll
      for (int k = 0; k < 3; ++k) qv[k] = BGMODE == 0 ? ldg_stream16(bg + 3 * gb + k) : __ldg(bg + 3 * gb + k);
    }
    const unsigned* fw = reinterpret_cast<const unsigned*>(fv);
    const unsigned* qw = reinterpret_cast<const unsigned*>(qv);
    const unsigned aws[4] = {av.x, av.y, av.z, av.w};
    uint4 ov[3], bv[3];
    unsigned* ow = reinterpret_cast<unsigned*>(ov);
    unsigned* bw = reinterpret_cast<unsigned*>(bv);
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      int c[12], q[12], o[12];
      unpack12(fw[3 * s4], fw[3 * s4 + 1], fw[3 * s4 + 2], c);
      if (BGMODE == 1) {
#pragma unroll
        for (int i = 0; i < 12; ++i) q[i] = q0[i];
      } else {
        unpack12(qw[3 * s4], qw[3 * s4 + 1], qw[3 * s4 + 2], q);
      }
      const unsigned aw = aws[s4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int a = (aw >> (8 * i)) & 255;
        const bool patch = (PATCH == VU_PATCH_ALPHA_LT128) ? (a < 128) : (PATCH == VU_PATCH_ALPHA_EQ0 ? (a == 0) : false);
        q[3 * i] = patch ? c[3 * i] : q[3 * i];
        q[3 * i + 1] = patch ? c[3 * i + 1] : q[3 * i + 1];
        q[3 * i + 2] = patch ? c[3 * i + 2] : q[3 * i + 2];
        int ih, is, iv, bh, bs, bvv;
        bgr2hsv_px(c[3 * i], c[3 * i + 1], c[3 * i + 2], tab, ih, is, iv);
        if (BGMODE == 1) {   // a patched pixel's background IS the frame pixel: no second conversion
          bh = patch ? ih : h0[i];
          bs = patch ? is : s0[i];
          bvv = patch ? iv : v0[i];
        } else if (a != 255) {
          bgr2hsv_px(q[3 * i], q[3 * i + 1], q[3 * i + 2], tab, bh, bs, bvv);
        } else {
          bh = bs = bvv = 0;   // k = 1 - 255/255 = 0 exactly: the background's HSV is multiplied by it
        }
        int fh = ih, fs = is, fv2 = iv;   // alpha == 255 (the inside of a matte): k = 0 exactly and trunc(clamp(x - 0 * y)) = x
        if (a != 255) {
          const float k = ktab[a];   // 1 - alpha/255.
          fh = trunc_clamp255(__fsub_rn(u8_to_f32(ih), __fmul_rn(k, u8_to_f32(bh))));
          fs = trunc_clamp255(__fsub_rn(u8_to_f32(is), __fmul_rn(k, u8_to_f32(bs))));
          fv2 = trunc_clamp255(__fsub_rn(u8_to_f32(iv), __fmul_rn(k, u8_to_f32(bvv))));
        }
        hsv2bgr_px(fh, fs, fv2, tab, o[3 * i], o[3 * i + 1], o[3 * i + 2]);
      }
      pack12(o, ow[3 * s4], ow[3 * s4 + 1], ow[3 * s4 + 2]);
      if (WRITE_BG) pack12(q, bw[3 * s4], bw[3 * s4 + 1], bw[3 * s4 + 2]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) stg_stream16(fg_out + 3 * g + k, ov[k]);
    if (WRITE_BG) {
#pragma unroll
      for (int k = 0; k < 3; ++k) stg_stream16(bg_out + 3 * g + k, bv[k]);
    }
  }
}

__global__ void __launch_bounds__(THREADS) get_bg_kernel(const uint8_t* __restrict__ alpha, const uint8_t* __restrict__ bg, int64_t ngroups,
                                                         uint8_t* __restrict__ out) {
  __shared__ HsvTab tab;
  __shared__ float ktab[256];
  hsv_tab_init(tab);
  for (int a = threadIdx.x; a < 256; a += blockDim.x) ktab[a] = __fsub_rn(1.f, __fdiv_rn((float)a, 255.f));
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    const unsigned* b4 = reinterpret_cast<const unsigned*>(bg) + 3 * g;
    int q[12], o[12];
    unpack12(__ldg(b4), __ldg(b4 + 1), __ldg(b4 + 2), q);
    const unsigned aw = __ldg(reinterpret_cast<const unsigned*>(alpha) + g);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = (aw >> (8 * i)) & 255;
      int bh, bs, bv;
      bgr2hsv_px(q[3 * i], q[3 * i + 1], q[3 * i + 2], tab, bh, bs, bv);
      const float k = ktab[a];
      hsv2bgr_px(trunc_clamp255(__fmul_rn(k, u8_to_f32(bh))), trunc_clamp255(__fmul_rn(k, u8_to_f32(bs))),
                 trunc_clamp255(__fmul_rn(k, u8_to_f32(bv))), tab, o[3 * i], o[3 * i + 1], o[3 * i + 2]);
    }
    unsigned w0, w1, w2;
    pack12(o, w0, w1, w2);
    unsigned* d4 = reinterpret_cast<unsigned*>(out) + 3 * g;
    d4[0] = w0; d4[1] = w1; d4[2] = w2;
  }
}

// float64 blends, one thread per 4 pixels.  AC = alpha channels (1 or 3)
template <int MODE, int AC>
__global__ void __launch_bounds__(THREADS) blend_kernel(const uint8_t* __restrict__ fg, const uint8_t* __restrict__ alpha,
                                                        const uint8_t* __restrict__ bg, int64_t ngroups, int64_t bg_groups,
                                                        uint8_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    const unsigned* f4 = reinterpret_cast<const unsigned*>(fg) + 3 * g;
    int c[12], q[12], a[12], o[12];
    unpack12(__ldg(f4), __ldg(f4 + 1), __ldg(f4 + 2), c);
    if (MODE != VU_BLEND_NAIVE) {
      const unsigned* b4 = reinterpret_cast<const unsigned*>(bg) + 3 * (g % bg_groups);
      unpack12(__ldg(b4), __ldg(b4 + 1), __ldg(b4 + 2), q);
    }
    if (AC == 3) {
      const unsigned* a4 = reinterpret_cast<const unsigned*>(alpha) + 3 * g;
      unpack12(__ldg(a4), __ldg(a4 + 1), __ldg(a4 + 2), a);
    } else {
      const unsigned aw = __ldg(reinterpret_cast<const unsigned*>(alpha) + g);
#pragma unroll
      for (int i = 0; i < 4; ++i) a[3 * i] = a[3 * i + 1] = a[3 * i + 2] = (aw >> (8 * i)) & 255;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      double m = __ddiv_rn((double)a[i], 255.0);
      double r;
      if (MODE == VU_BLEND_NAIVE) {
        r = __dmul_rn((double)c[i], m);
      } else if (MODE == VU_BLEND_FUSE) {
        r = __dadd_rn(__dmul_rn(m, (double)c[i]), __dmul_rn(__dsub_rn(1.0, m), (double)q[i]));
      } else if (MODE == VU_BLEND_COMPOSITE) {
        if (m > 0.9) m = 1.0;
        r = __dadd_rn((double)c[i], __dmul_rn((double)q[i], __dsub_rn(1.0, m)));
        r = fmin(fmax(r, 0.0), 255.0);
      } else {
        r = __dadd_rn(__dmul_rn((double)c[i], m), __dmul_rn((double)q[i], __dsub_rn(1.0, m)));
      }
      o[i] = (int)r;
    }
    unsigned w0, w1, w2;
    pack12(o, w0, w1, w2);
    unsigned* d4 = reinterpret_cast<unsigned*>(out) + 3 * g;
    d4[0] = w0; d4[1] = w1; d4[2] = w2;
  }
}

// The same blends at 16 pixels (48 bytes) per thread with 128-bit streaming loads / stores.  The float64 division
// alpha/255 and 1 - alpha/255 (and COMPOSITE's a > 0.9 -> 1) are tabulated per CTA for the 256 possible alphas with
// the very operations of the per-pixel kernel above, so the results are identical; uint8 -> float64 is one exact DADD
// (2^52 + x, minus 2^52) and the truncating cast one DADD.RZ (x + 2^52: the low word is trunc(x)), which leaves six
// float64 pipe operations per output byte and nothing on the slow conversion unit.  (Tried: the blends in integers -
// PRMT + IDP.4A + two IMAD per byte - with the float64 sequence only for the bytes whose quotient is exact, the only
// ones where truncation can differ: bit-exact as well, but ~19 instructions per byte and a fix-up loop made it slower,
// 1.33 ms against 1.15 ms per 300 1080p frames.  blend16_int_kernel below is the integer variant that IS faster; it takes
// the replace / fuse blends with one alpha channel, this kernel the rest.)
__device__ __forceinline__ double u8_to_f64(unsigned x) { return __dsub_rn(__hiloint2double(0x43300000, (int)x), 4503599627370496.0); }
__device__ __forceinline__ int f64_trunc_nonneg(double x) { return __double2loint(__dadd_rz(x, 4503599627370496.0)); }

template <int MODE, int AC>
__global__ void __launch_bounds__(THREADS) blend16_kernel(const uint4* __restrict__ fg, const uint4* __restrict__ alpha, const uint4* __restrict__ bg,
                                                          int64_t ngroups, int64_t bg_groups, uint4* __restrict__ out) {
  __shared__ double mtab[256], otab[256];
  for (int a = threadIdx.x; a < 256; a += THREADS) {
    double m = __ddiv_rn((double)a, 255.0);
    if (MODE == VU_BLEND_COMPOSITE && m > 0.9) m = 1.0;
    mtab[a] = m;
    otab[a] = __dsub_rn(1.0, m);
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    uint4 f[3], q[3], av[3], o[3];
    if (AC == 3) {
#pragma unroll
      for (int k = 0; k < 3; ++k) av[k] = ldg_stream16(alpha + 3 * g + k);
    } else {
      av[0] = ldg_stream16(alpha + g);
    }
    if (MODE == VU_BLEND_REPLACE && AC == 1) {
      // mattes are mostly 0 or 255: m = 1 gives fl(c * 1) + fl(q * 0) = c and m = 0 gives q, exactly: sixteen such pixels are
      // a copy of one of the two images, and the other one is not even read
      const unsigned all1 = av[0].x & av[0].y & av[0].z & av[0].w, any = av[0].x | av[0].y | av[0].z | av[0].w;
      if (all1 == 0xFFFFFFFFu) {
#pragma unroll
        for (int k = 0; k < 3; ++k) stg_stream16(out + 3 * g + k, ldg_stream16(fg + 3 * g + k));
        continue;
      }
      if (any == 0u) {
        const int64_t gb = g % bg_groups;
#pragma unroll
        for (int k = 0; k < 3; ++k) stg_stream16(out + 3 * g + k, (bg_groups == ngroups) ? ldg_stream16(bg + 3 * gb + k) : __ldg(bg + 3 * gb + k));
        continue;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) f[k] = ldg_stream16(fg + 3 * g + k);
    if (MODE != VU_BLEND_NAIVE) {
      const int64_t gb = g % bg_groups;
#pragma unroll
      for (int k = 0; k < 3; ++k) q[k] = (bg_groups == ngroups) ? ldg_stream16(bg + 3 * gb + k) : __ldg(bg + 3 * gb + k);
    }
    const unsigned* fw = reinterpret_cast<const unsigned*>(f);
    const unsigned* qw = reinterpret_cast<const unsigned*>(q);
    const unsigned* aw = reinterpret_cast<const unsigned*>(av);
    unsigned* ow = reinterpret_cast<unsigned*>(o);
#pragma unroll
    for (int w = 0; w < 12; ++w) {
      unsigned vb[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = 4 * w + b;                       // byte of the 48
        const int ai = AC == 3 ? i : i / 3;            // its alpha byte
        // one PRMT per byte extraction (selector 4 = a zero byte), one per pair of results: the kernel is bound by
        // instruction issue as much as by the float64 pipe
        const unsigned a = __byte_perm(aw[ai >> 2], 0u, 0x4440 | (ai & 3));
        const double m = mtab[a];
        const double c = u8_to_f64(__byte_perm(fw[w], 0u, 0x4440 | b));
        double r;
        if (MODE == VU_BLEND_NAIVE) {
          r = __dmul_rn(c, m);
        } else {
          const double qq = u8_to_f64(__byte_perm(qw[w], 0u, 0x4440 | b));
          if (MODE == VU_BLEND_COMPOSITE) r = __dadd_rn(c, __dmul_rn(qq, otab[a]));
          else r = __dadd_rn(__dmul_rn(c, m), __dmul_rn(qq, otab[a]));   // FUSE == REPLACE (products commute)
        }
        int v = f64_trunc_nonneg(r);
        if (MODE == VU_BLEND_COMPOSITE) v = min(v, 255);
        vb[b] = (unsigned)v;
      }
      ow[w] = __byte_perm(__byte_perm(vb[0], vb[1], 0x0040), __byte_perm(vb[2], vb[3], 0x0040), 0x5410);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) stg_stream16(out + 3 * g + k, o[k]);
  }
}

// REPLACE / FUSE with one alpha per pixel, in integers.  The float64 sequence trunc(fl(fl(c m) + fl(q fl(1 - m)))), m = fl(a / 255),
// equals floor(t / 255), t = c a + q (255 - a), whenever 255 does not divide t: its rounding errors (~1e-13) cannot cross an
// integer that lies at least 1 / 255 away.  Where 255 | t the exact value IS an integer K and the float64 result is K or a hair
// below it (truncated: K - 1): 12 397 of the 2^24 (a, c, q) triples, with no closed form.
//   main path: two channels of a pixel that lie in one word share an IMAD pair as 16-bit lanes ([c0, c1] * a + [q0, q1] * (255 - a)
//   cannot carry: <= 65025), the third one is an IDP.4A; t / 255 = (t + 1 + (t >> 8)) >> 8 on both lanes at once, and the LOW
//   byte of that sum is zero exactly when 255 | t and t > 0 - the candidates.  Candidate bytes of soft pixels (a not 0 / 255: those
//   two are exact copies) are collected as bits, by position;
//   fix-up: a thread with candidates (1.9 % of the bytes on random data) parks its 112 input bytes in shared memory - registers
//   cannot be indexed by a run-time position -, walks the bits, evaluates the float64 sequence of blend_kernel for those bytes and
//   stores the byte again where it differs.  (A first version looked the triples up in a 2 MB bit table and re-read the bytes
//   from global memory: two dependent L2 round trips per candidate made it latency-bound, 1.44 ms against 1.17 ms for the
//   float64 kernel on 300 x 1080p.)
// s = t + 1 + (t >> 8) on two 16-bit lanes: byte 1 / 3 = floor(t / 255), byte 0 / 2 = 0 iff 255 | t and t > 0
__device__ __forceinline__ unsigned div255_lanes(unsigned t) { return t + __byte_perm(t, 0u, 0x4341) + 0x00010001u; }

// 80 registers, three CTAs per SM: capping them at 64 for a fourth CTA was no faster (0.954 against 0.948 ms)
__global__ void __launch_bounds__(THREADS, 3) blend16_int_kernel(const uint4* __restrict__ fg, const uint4* __restrict__ alpha, const uint4* __restrict__ bg,
                                                              int64_t ngroups, int64_t bg_groups, uint4* __restrict__ out) {
  __shared__ double mtab[256];
  __shared__ uint4 park[7][THREADS];   // fg 0..2, bg 3..5, alpha 6 of the thread's group
  mtab[threadIdx.x] = __ddiv_rn((double)threadIdx.x, 255.0);   // THREADS == 256
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  // what a group reads follows from its alphas: 0 = all 255 (sixteen pixels of the foreground: the background is not read),
  // 1 = all 0 (the background alone), 2 = both
  auto kind_of = [](const uint4& a) { return (a.x & a.y & a.z & a.w) == 0xFFFFFFFFu ? 0 : ((a.x | a.y | a.z | a.w) == 0u ? 1 : 2); };
  uint4 f0, f1, f2, q0, q1, q2;   // scalars, not arrays: with arrays behind a lambda the compiler kept them on the stack
#define VU_FETCH(gg, gbb, kind)                                                              \
  do {                                                                                       \
    if ((kind) != 1) {                                                                       \
      f0 = ldg_stream16(fg + 3 * (gg)); f1 = ldg_stream16(fg + 3 * (gg) + 1); f2 = ldg_stream16(fg + 3 * (gg) + 2); \
    }                                                                                        \
    if ((kind) != 0) {                                                                       \
      const uint4* bp = bg + 3 * (gbb);                                                      \
      if (bg_groups == ngroups) { q0 = ldg_stream16(bp); q1 = ldg_stream16(bp + 1); q2 = ldg_stream16(bp + 2); } \
      else { q0 = __ldg(bp); q1 = __ldg(bp + 1); q2 = __ldg(bp + 2); }                       \
    }                                                                                        \
  } while (0)
  // software pipeline over the thread's groups: the alphas run two groups ahead and the pixels one group ahead of the
  // arithmetic, so that the next group's pixels are in flight while this group's candidates are walked and no group waits
  // for its alphas before it can ask for its pixels
  const int64_t g0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g0 >= ngroups) return;
  uint4 av = ldg_stream16(alpha + g0);
  uint4 av1 = g0 + stride < ngroups ? ldg_stream16(alpha + g0 + stride) : zero4;
  // the background's group index, g mod bg_groups, advances with the loop: one 64-bit modulo per thread, not one per group
  const int64_t sb = stride % bg_groups;
  int64_t gb1 = g0 % bg_groups;
  VU_FETCH(g0, gb1, kind_of(av));
  for (int64_t g = g0; g < ngroups; g += stride) {
    const int64_t g1 = g + stride, g2 = g1 + stride;
    gb1 += sb;
    if (gb1 >= bg_groups) gb1 -= bg_groups;
    const uint4 av2 = g2 < ngroups ? ldg_stream16(alpha + g2) : zero4;
    const int kind = kind_of(av);
    if (kind != 2) {
      stg_stream16(out + 3 * g, kind == 0 ? f0 : q0);
      stg_stream16(out + 3 * g + 1, kind == 0 ? f1 : q1);
      stg_stream16(out + 3 * g + 2, kind == 0 ? f2 : q2);
      if (g1 < ngroups) VU_FETCH(g1, gb1, kind_of(av1));
      av = av1; av1 = av2;
      continue;
    }
    uint4 o[3];
    const unsigned fw[12] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w};
    const unsigned qw[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    const unsigned aw[4] = {av.x, av.y, av.z, av.w};
    unsigned* ow = reinterpret_cast<unsigned*>(o);
    unsigned accA = 0u, accB = 0u;   // candidate bits: byte b of output word w at bit 8 b + w (w < 8) / 8 b + w - 4 (w >= 8)
#pragma unroll
    for (int j = 0; j < 4; ++j) {    // four pixels = three words: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
      const unsigned a4 = aw[j];
      // bit 7 of byte p: pixel p is soft (a differs from its own sign byte, 0x00 / 0xFF)
      unsigned sg;                                                      // every byte's sign spread over the byte
      asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(sg) : "r"(a4));
      const unsigned ax = a4 ^ sg;
      const unsigned soft = (((ax & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | ax) & 0x80808080u;
      unsigned m[4], om[4], mw[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        m[p] = __byte_perm(a4, 0u, 0x4440 | p);
        om[p] = 255u - m[p];
        mw[p] = om[p] * 256u + m[p];                                  // IDP.4A weights [a, 255 - a, 0, 0]
      }
      const unsigned fa = fw[3 * j], fb = fw[3 * j + 1], fc = fw[3 * j + 2];   // the three words of the four pixels
      const unsigned qa = qw[3 * j], qb = qw[3 * j + 1], qc = qw[3 * j + 2];
      // pairs (two channels of one pixel, one word): lanes [x, 0, y, 0]
      const unsigned tP0 = __byte_perm(fa, 0u, 0x4140) * m[0] + __byte_perm(qa, 0u, 0x4140) * om[0];   // B0 G0
      const unsigned tP1 = __byte_perm(fb, 0u, 0x4140) * m[1] + __byte_perm(qb, 0u, 0x4140) * om[1];   // G1 R1
      const unsigned tP2 = __byte_perm(fb, 0u, 0x4342) * m[2] + __byte_perm(qb, 0u, 0x4342) * om[2];   // B2 G2
      const unsigned tP3 = __byte_perm(fc, 0u, 0x4241) * m[3] + __byte_perm(qc, 0u, 0x4241) * om[3];   // B3 G3
      // singles: [c, q, *, *] . [a, 255 - a, 0, 0]
      const unsigned tS0 = __dp4a(__byte_perm(fa, qa, 0x2262), mw[0], 0u);                              // R0
      const unsigned tS1 = __dp4a(__byte_perm(fa, qa, 0x3373), mw[1], 0u);                              // B1
      const unsigned tS2 = __dp4a(__byte_perm(fc, qc, 0x0040), mw[2], 0u);                              // R2
      const unsigned tS3 = __dp4a(__byte_perm(fc, qc, 0x3373), mw[3], 0u);                              // R3
      const unsigned sP0 = div255_lanes(tP0), sP1 = div255_lanes(tP1), sP2 = div255_lanes(tP2), sP3 = div255_lanes(tP3);
      const unsigned sS01 = div255_lanes(tS1 * 65536u + tS0), sS23 = div255_lanes(tS3 * 65536u + tS2);
      ow[3 * j] = __byte_perm(sP0, sS01, 0x7531);
      ow[3 * j + 1] = __byte_perm(sP1, sP2, 0x7531);
      ow[3 * j + 2] = __byte_perm(sS23, sP3, 0x3751);
      // the low bytes in the same order, zero bytes of soft pixels -> bit 7 (a borrow can flag the byte above a zero byte as
      // well: a needless evaluation, never a wrong one)
      const unsigned l0 = __byte_perm(sP0, sS01, 0x6420), l1 = __byte_perm(sP1, sP2, 0x6420), l2 = __byte_perm(sS23, sP3, 0x2640);
      const unsigned z0 = (l0 - 0x01010101u) & ~l0 & __byte_perm(soft, 0u, 0x1000);
      const unsigned z1 = (l1 - 0x01010101u) & ~l1 & __byte_perm(soft, 0u, 0x2211);
      const unsigned z2 = (l2 - 0x01010101u) & ~l2 & __byte_perm(soft, 0u, 0x3332);
      if (3 * j < 8) accA = (accA >> 1) + z0; else accB = (accB >> 1) + z0;
      if (3 * j + 1 < 8) accA = (accA >> 1) + z1; else accB = (accB >> 1) + z1;
      if (3 * j + 2 < 8) accA = (accA >> 1) + z2; else accB = (accB >> 1) + z2;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) stg_stream16(out + 3 * g + k, o[k]);
    const bool cand = (accA | accB) != 0u;
    if (cand) {
      // parked only now: parking every group up front (registers die early, 40 instead of 58, six CTAs per SM) was slower,
      // 1.14 against 1.07 ms
      park[0][threadIdx.x] = f0; park[1][threadIdx.x] = f1; park[2][threadIdx.x] = f2;
      park[3][threadIdx.x] = q0; park[4][threadIdx.x] = q1; park[5][threadIdx.x] = q2;
      park[6][threadIdx.x] = av;
    }
    if (g1 < ngroups) VU_FETCH(g1, gb1, kind_of(av1));   // the registers are free: the next group's pixels travel during the walk below
    av = av1; av1 = av2;
    if (cand) {
      const uint8_t* pk = reinterpret_cast<const uint8_t*>(&park[0][threadIdx.x]);   // the thread's own bytes: no barrier
      uint8_t* o8 = reinterpret_cast<uint8_t*>(out + 3 * g);
      do {
        unsigned n, w0;
        if (accA) { n = __ffs(accA) - 1; accA &= accA - 1; w0 = 0; }
        else { n = __ffs(accB) - 1; accB &= accB - 1; w0 = 4; }
        const unsigned i = 4 * ((n & 7) + w0) + (n >> 3);              // byte of the 48
        const unsigned pi = (i * 171u) >> 9;                           // its pixel, i / 3
        const unsigned c = pk[(i >> 4) * (THREADS * 16) + (i & 15)], b = pk[(3 + (i >> 4)) * (THREADS * 16) + (i & 15)];
        const unsigned a = pk[6 * (THREADS * 16) + pi];
        const double md = mtab[a];
        const int v = f64_trunc_nonneg(__dadd_rn(__dmul_rn(u8_to_f64(c), md), __dmul_rn(u8_to_f64(b), __dsub_rn(1.0, md))));
        const int k = (int)(((c * a + b * (255u - a)) * 0x8081u) >> 23);
        if (v != k) o8[i] = (uint8_t)v;
      } while (accA | accB);
    }
  }
}

#undef VU_FETCH

__global__ void __launch_bounds__(THREADS) fuse_bg_kernel(const unsigned* __restrict__ bg, const unsigned* __restrict__ always, int64_t nwords,
                                                          int64_t always_words, float beta, float omb, unsigned* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
    const unsigned a = __ldg(bg + i), b = __ldg(always + (i % always_words));
    unsigned w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float x = (float)((a >> (8 * k)) & 255), y = (float)((b >> (8 * k)) & 255);
      const float r = __fadd_rn(__fmul_rn(x, beta), __fmul_rn(omb, y));
      w |= (unsigned)((int)r & 255) << (8 * k);
    }
    out[i] = w;
  }
}

__global__ void __launch_bounds__(THREADS) bgdiff_gray_kernel(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ bg, int64_t ngroups,
                                                              int64_t bg_groups, int thr, uint8_t* __restrict__ gray) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    const unsigned* f4 = reinterpret_cast<const unsigned*>(frame) + 3 * g;
    const unsigned* b4 = reinterpret_cast<const unsigned*>(bg) + 3 * (g % bg_groups);
    int c[12], q[12];
    unpack12(__ldg(f4), __ldg(f4 + 1), __ldg(f4 + 2), c);
    unpack12(__ldg(b4), __ldg(b4 + 1), __ldg(b4 + 2), q);
    unsigned w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int y = bgr2gray_px(abs(c[3 * i] - q[3 * i]), abs(c[3 * i + 1] - q[3 * i + 1]), abs(c[3 * i + 2] - q[3 * i + 2]));
      if (y > thr) y = 255;  // values <= thr are kept, not zeroed
      w |= (unsigned)y << (8 * i);
    }
    reinterpret_cast<unsigned*>(gray)[g] = w;
  }
}

// ---- one pixel per thread, byte accesses: frames whose pixel count is not a multiple of 4 (odd x odd sizes) or whose
// buffers are not 4-byte aligned (views into larger arrays).  Same arithmetic as the vector kernels above. ----
template <int PATCH, bool WRITE_BG>
__global__ void __launch_bounds__(THREADS) get_fg_px_kernel(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ alpha,
                                                            const uint8_t* __restrict__ bg, int64_t npix, int64_t bg_npix,
                                                            uint8_t* __restrict__ fg_out, uint8_t* __restrict__ bg_out) {
  __shared__ HsvTab tab;
  hsv_tab_init(tab);
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = alpha[i];
    const uint8_t* f = frame + 3 * i;
    const uint8_t* b = bg + 3 * (i % bg_npix);
    const bool patch = (PATCH == VU_PATCH_ALPHA_LT128) ? (a < 128) : (PATCH == VU_PATCH_ALPHA_EQ0 ? (a == 0) : false);
    const int c0 = f[0], c1 = f[1], c2 = f[2];
    const int q0 = patch ? c0 : b[0], q1 = patch ? c1 : b[1], q2 = patch ? c2 : b[2];
    int ih, is, iv, bh, bs, bv, o0, o1, o2;
    bgr2hsv_px(c0, c1, c2, tab, ih, is, iv);
    bgr2hsv_px(q0, q1, q2, tab, bh, bs, bv);
    const float k = __fsub_rn(1.f, __fdiv_rn((float)a, 255.f));
    hsv2bgr_px(trunc_clamp255(__fsub_rn(u8_to_f32(ih), __fmul_rn(k, u8_to_f32(bh)))), trunc_clamp255(__fsub_rn(u8_to_f32(is), __fmul_rn(k, u8_to_f32(bs)))),
               trunc_clamp255(__fsub_rn(u8_to_f32(iv), __fmul_rn(k, u8_to_f32(bv)))), tab, o0, o1, o2);
    fg_out[3 * i] = (uint8_t)o0; fg_out[3 * i + 1] = (uint8_t)o1; fg_out[3 * i + 2] = (uint8_t)o2;
    if (WRITE_BG) { bg_out[3 * i] = (uint8_t)q0; bg_out[3 * i + 1] = (uint8_t)q1; bg_out[3 * i + 2] = (uint8_t)q2; }
  }
}

__global__ void __launch_bounds__(THREADS) get_bg_px_kernel(const uint8_t* __restrict__ alpha, const uint8_t* __restrict__ bg, int64_t npix,
                                                            uint8_t* __restrict__ out) {
  __shared__ HsvTab tab;
  hsv_tab_init(tab);
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    int bh, bs, bv, o0, o1, o2;
    bgr2hsv_px(bg[3 * i], bg[3 * i + 1], bg[3 * i + 2], tab, bh, bs, bv);
    const float k = __fsub_rn(1.f, __fdiv_rn((float)alpha[i], 255.f));
    hsv2bgr_px(trunc_clamp255(__fmul_rn(k, u8_to_f32(bh))), trunc_clamp255(__fmul_rn(k, u8_to_f32(bs))), trunc_clamp255(__fmul_rn(k, u8_to_f32(bv))),
               tab, o0, o1, o2);
    out[3 * i] = (uint8_t)o0; out[3 * i + 1] = (uint8_t)o1; out[3 * i + 2] = (uint8_t)o2;
  }
}

template <int MODE, int AC>
__global__ void __launch_bounds__(THREADS) blend_px_kernel(const uint8_t* __restrict__ fg, const uint8_t* __restrict__ alpha,
                                                           const uint8_t* __restrict__ bg, int64_t npix, int64_t bg_npix, uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const double c = (double)fg[3 * i + ch];
      const double q = MODE == VU_BLEND_NAIVE ? 0.0 : (double)bg[3 * (i % bg_npix) + ch];
      double m = __ddiv_rn((double)(AC == 3 ? alpha[3 * i + ch] : alpha[i]), 255.0);
      double r;
      if (MODE == VU_BLEND_NAIVE) {
        r = __dmul_rn(c, m);
      } else if (MODE == VU_BLEND_COMPOSITE) {
        if (m > 0.9) m = 1.0;
        r = fmin(fmax(__dadd_rn(c, __dmul_rn(q, __dsub_rn(1.0, m))), 0.0), 255.0);
      } else if (MODE == VU_BLEND_FUSE) {
        r = __dadd_rn(__dmul_rn(m, c), __dmul_rn(__dsub_rn(1.0, m), q));
      } else {
        r = __dadd_rn(__dmul_rn(c, m), __dmul_rn(q, __dsub_rn(1.0, m)));
      }
      out[3 * i + ch] = (uint8_t)(int)r;
    }
  }
}

__global__ void __launch_bounds__(THREADS) fuse_bg_px_kernel(const uint8_t* __restrict__ bg, const uint8_t* __restrict__ always, int64_t nbytes,
                                                             int64_t always_bytes, float beta, float omb, uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (uint8_t)((int)__fadd_rn(__fmul_rn((float)bg[i], beta), __fmul_rn(omb, (float)always[i % always_bytes])) & 255);
}

__global__ void __launch_bounds__(THREADS) bgdiff_gray_px_kernel(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ bg, int64_t npix,
                                                                 int64_t bg_npix, int thr, uint8_t* __restrict__ gray) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t* f = frame + 3 * i;
    const uint8_t* b = bg + 3 * (i % bg_npix);
    int y = bgr2gray_px(abs((int)f[0] - (int)b[0]), abs((int)f[1] - (int)b[1]), abs((int)f[2] - (int)b[2]));
    if (y > thr) y = 255;
    gray[i] = (uint8_t)y;
  }
}

inline bool al4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; }
// true when the 4-pixel vector kernels cannot take the call: the byte kernels above do
inline bool overlaps(const void* a, int64_t na, const void* b, int64_t nb) {
  const uintptr_t x = reinterpret_cast<uintptr_t>(a), y = reinterpret_cast<uintptr_t>(b);
  return x < y + (uintptr_t)nb && y < x + (uintptr_t)na;
}

// VU_BLEND_FP64=1 keeps the float64 kernel for the replace / fuse blend (A/B measurements)
inline bool blend_fp64_forced() {
  static const bool forced = [] { const char* e = getenv("VU_BLEND_FP64"); return e && e[0] == '1'; }();
  return forced;
}

inline bool needs_px(int64_t npix, int64_t other_npix, std::initializer_list<const void*> ptrs) {
  if (npix % 4 != 0 || (other_npix > 0 && other_npix % 4 != 0)) return true;
  for (const void* p : ptrs)
    if (p && !al4(p)) return true;
  return false;
}

}  // namespace
}  // namespace vu

using namespace vu;

// The vector kernels need 4-byte aligned pointers and pixel counts that are multiples of 4 (true for every even-sized
// frame); anything else (odd x odd frames, unaligned views) goes to the one-pixel-per-thread kernels: same results.

extern "C" int vu_get_fg(const uint8_t* frame, const uint8_t* alpha, const uint8_t* bg, int64_t npix, int64_t bg_npix, int patch_mode,
                         uint8_t* fg_out, uint8_t* bg_out, vu_stream_t stream) {
  VU_REQUIRE(frame && alpha && bg && fg_out && npix >= 0 && bg_npix > 0);
  VU_REQUIRE(patch_mode >= VU_PATCH_NONE && patch_mode <= VU_PATCH_ALPHA_EQ0);
  if (npix == 0) return VU_OK;
  if (needs_px(npix, bg_npix == 1 ? 0 : bg_npix, {frame, alpha, bg, fg_out, bg_out}) || bg_npix == 1) {
    const int g = grid_for(npix, THREADS, 8);
#define VU_PX(P)                                                                                                   \
  do {                                                                                                             \
    if (bg_out) get_fg_px_kernel<P, true><<<g, THREADS, 0, S(stream)>>>(frame, alpha, bg, npix, bg_npix, fg_out, bg_out); \
    else get_fg_px_kernel<P, false><<<g, THREADS, 0, S(stream)>>>(frame, alpha, bg, npix, bg_npix, fg_out, bg_out);       \
  } while (0)
    if (patch_mode == VU_PATCH_NONE) VU_PX(VU_PATCH_NONE);
    else if (patch_mode == VU_PATCH_ALPHA_LT128) VU_PX(VU_PATCH_ALPHA_LT128);
    else VU_PX(VU_PATCH_ALPHA_EQ0);
#undef VU_PX
    VU_RETURN_LAUNCH();
  }
  const int64_t ng = npix / 4, bgg = bg_npix / 4;
  int mode = 3;
  dim3 grid(grid_for(ng, THREADS, 8));
  if (bgg == ng) mode = 0;
  else if (bgg == 1) mode = 1;
  else if (ng % bgg == 0 && ng / bgg <= 65535) {
    mode = 2;
    const int64_t reps = ng / bgg;
    int64_t gx = ((int64_t)device_sms() * 8 + reps - 1) / reps;
    const int64_t need = (bgg + THREADS - 1) / THREADS;
    if (gx > need) gx = need;
    grid = dim3((unsigned)(gx < 1 ? 1 : gx), (unsigned)reps);
  }
  // 16 pixels per thread whenever the sizes and alignments allow (BGMODE 3 has no wide variant)
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool wide = mode != 3 && npix % 16 == 0 && (mode == 1 || bg_npix % 16 == 0) && a16(frame) && a16(alpha) && a16(fg_out) && a16(bg_out) &&
                    (mode == 1 || a16(bg));
  const int64_t ng16 = ng / 4, bgg16 = mode == 1 ? 1 : bgg / 4;
  if (wide) {
    grid = dim3(grid_for(ng16, THREADS, 8));
    if (mode == 2) {
      const int64_t reps = ng16 / bgg16;
      int64_t gx = ((int64_t)device_sms() * 8 + reps - 1) / reps;
      const int64_t need = (bgg16 + THREADS - 1) / THREADS;
      if (gx > need) gx = need;
      grid = dim3((unsigned)(gx < 1 ? 1 : gx), (unsigned)reps);
    }
  }
#define LAUNCH3(P, W, M)                                                                                                                  \
  do {                                                                                                                                    \
    if (wide && M != 3)                                                                                                                   \
      get_fg16_kernel<P, W, (M == 3 ? 0 : M)><<<grid, THREADS, 0, S(stream)>>>(reinterpret_cast<const uint4*>(frame), reinterpret_cast<const uint4*>(alpha), \
                                                                               reinterpret_cast<const uint4*>(bg), ng16, bgg16,          \
                                                                               reinterpret_cast<uint4*>(fg_out), reinterpret_cast<uint4*>(bg_out)); \
    else                                                                                                                                  \
      get_fg_kernel<P, W, M><<<grid, THREADS, 0, S(stream)>>>(frame, alpha, bg, ng, bgg, fg_out, bg_out);                                 \
  } while (0)
#define LAUNCH(P, W)                      \
  do {                                    \
    if (mode == 0) LAUNCH3(P, W, 0);      \
    else if (mode == 1) LAUNCH3(P, W, 1); \
    else if (mode == 2) LAUNCH3(P, W, 2); \
    else LAUNCH3(P, W, 3);                \
  } while (0)
  if (bg_out) {
    if (patch_mode == VU_PATCH_NONE) LAUNCH(VU_PATCH_NONE, true);
    else if (patch_mode == VU_PATCH_ALPHA_LT128) LAUNCH(VU_PATCH_ALPHA_LT128, true);
    else LAUNCH(VU_PATCH_ALPHA_EQ0, true);
  } else {
    if (patch_mode == VU_PATCH_NONE) LAUNCH(VU_PATCH_NONE, false);
    else if (patch_mode == VU_PATCH_ALPHA_LT128) LAUNCH(VU_PATCH_ALPHA_LT128, false);
    else LAUNCH(VU_PATCH_ALPHA_EQ0, false);
  }
#undef LAUNCH
#undef LAUNCH3
  VU_RETURN_LAUNCH();
}

extern "C" int vu_get_bg(const uint8_t* alpha, const uint8_t* bg, int64_t npix, uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(alpha && bg && out && npix >= 0);
  if (npix == 0) return VU_OK;
  if (needs_px(npix, 0, {alpha, bg, out})) {
    get_bg_px_kernel<<<grid_for(npix, THREADS, 8), THREADS, 0, S(stream)>>>(alpha, bg, npix, out);
    VU_RETURN_LAUNCH();
  }
  get_bg_kernel<<<grid_for(npix / 4, THREADS, 8), THREADS, 0, S(stream)>>>(alpha, bg, npix / 4, out);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_blend(int mode, const uint8_t* fg, const uint8_t* alpha, int alpha_channels, const uint8_t* bg, int64_t npix,
                        int64_t bg_npix, uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(fg && alpha && out && npix >= 0);
  VU_REQUIRE(mode >= VU_BLEND_NAIVE && mode <= VU_BLEND_REPLACE);
  VU_REQUIRE(alpha_channels == 1 || alpha_channels == 3);
  VU_REQUIRE(mode == VU_BLEND_NAIVE || (bg && bg_npix > 0));
  if (npix == 0) return VU_OK;
  if (needs_px(npix, mode == VU_BLEND_NAIVE ? 0 : bg_npix, {fg, alpha, bg, out})) {
    const int g = grid_for(npix, THREADS, 8);
    const int64_t bn = mode == VU_BLEND_NAIVE ? 1 : bg_npix;
#define VU_PX(M)                                                                                              \
  do {                                                                                                        \
    if (alpha_channels == 1) blend_px_kernel<M, 1><<<g, THREADS, 0, S(stream)>>>(fg, alpha, bg, npix, bn, out); \
    else blend_px_kernel<M, 3><<<g, THREADS, 0, S(stream)>>>(fg, alpha, bg, npix, bn, out);                  \
  } while (0)
    switch (mode) {
      case VU_BLEND_NAIVE: VU_PX(VU_BLEND_NAIVE); break;
      case VU_BLEND_FUSE: VU_PX(VU_BLEND_FUSE); break;
      case VU_BLEND_COMPOSITE: VU_PX(VU_BLEND_COMPOSITE); break;
      default: VU_PX(VU_BLEND_REPLACE); break;
    }
#undef VU_PX
    VU_RETURN_LAUNCH();
  }
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool wide = npix % 16 == 0 && (mode == VU_BLEND_NAIVE || bg_npix % 16 == 0) && al16(fg) && al16(alpha) && al16(out) && (!bg || al16(bg));
  if (wide) {
    const int64_t ng = npix / 16, bgg = (mode == VU_BLEND_NAIVE) ? 1 : bg_npix / 16;
    const int grid = grid_for(ng, THREADS, 8);
    if ((mode == VU_BLEND_REPLACE || mode == VU_BLEND_FUSE) && alpha_channels == 1 && !blend_fp64_forced() &&
        !overlaps(out, npix * 3, fg, npix * 3) && !overlaps(out, npix * 3, bg, bg_npix * 3) && !overlaps(out, npix * 3, alpha, npix)) {
      blend16_int_kernel<<<grid, THREADS, 0, S(stream)>>>(reinterpret_cast<const uint4*>(fg), reinterpret_cast<const uint4*>(alpha),
                                                          reinterpret_cast<const uint4*>(bg), ng, bgg, reinterpret_cast<uint4*>(out));
      VU_RETURN_LAUNCH();
    }
#define LAUNCH16(M)                                                                                                          \
  do {                                                                                                                       \
    if (alpha_channels == 1)                                                                                                 \
      blend16_kernel<M, 1><<<grid, THREADS, 0, S(stream)>>>(reinterpret_cast<const uint4*>(fg), reinterpret_cast<const uint4*>(alpha), \
                                                            reinterpret_cast<const uint4*>(bg), ng, bgg, reinterpret_cast<uint4*>(out)); \
    else                                                                                                                     \
      blend16_kernel<M, 3><<<grid, THREADS, 0, S(stream)>>>(reinterpret_cast<const uint4*>(fg), reinterpret_cast<const uint4*>(alpha), \
                                                            reinterpret_cast<const uint4*>(bg), ng, bgg, reinterpret_cast<uint4*>(out)); \
  } while (0)
    switch (mode) {
      case VU_BLEND_NAIVE: LAUNCH16(VU_BLEND_NAIVE); break;
      case VU_BLEND_COMPOSITE: LAUNCH16(VU_BLEND_COMPOSITE); break;
      default: LAUNCH16(VU_BLEND_REPLACE); break;   // FUSE and REPLACE are the same arithmetic
    }
#undef LAUNCH16
    VU_RETURN_LAUNCH();
  }
  const int64_t ng = npix / 4, bgg = (mode == VU_BLEND_NAIVE) ? 1 : bg_npix / 4;
  const int grid = grid_for(ng, THREADS, 8);
#define LAUNCH(M)                                                                                        \
  do {                                                                                                   \
    if (alpha_channels == 1) blend_kernel<M, 1><<<grid, THREADS, 0, S(stream)>>>(fg, alpha, bg, ng, bgg, out); \
    else blend_kernel<M, 3><<<grid, THREADS, 0, S(stream)>>>(fg, alpha, bg, ng, bgg, out);               \
  } while (0)
  switch (mode) {
    case VU_BLEND_NAIVE: LAUNCH(VU_BLEND_NAIVE); break;
    case VU_BLEND_FUSE: LAUNCH(VU_BLEND_FUSE); break;
    case VU_BLEND_COMPOSITE: LAUNCH(VU_BLEND_COMPOSITE); break;
    default: LAUNCH(VU_BLEND_REPLACE); break;
  }
#undef LAUNCH
  VU_RETURN_LAUNCH();
}

extern "C" int vu_fuse_bg(const uint8_t* bg, const uint8_t* bg_always, int64_t npix, int64_t always_npix, float beta, float one_minus_beta,
                          uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(bg && bg_always && out && npix >= 0 && always_npix > 0);
  if (npix == 0) return VU_OK;
  if (needs_px(npix, always_npix, {bg, bg_always, out})) {
    fuse_bg_px_kernel<<<grid_for(npix * 3, THREADS, 8), THREADS, 0, S(stream)>>>(bg, bg_always, npix * 3, always_npix * 3, beta, one_minus_beta, out);
    VU_RETURN_LAUNCH();
  }
  const int64_t nwords = npix * 3 / 4, aw = always_npix * 3 / 4;
  fuse_bg_kernel<<<grid_for(nwords, THREADS, 8), THREADS, 0, S(stream)>>>(reinterpret_cast<const unsigned*>(bg), reinterpret_cast<const unsigned*>(bg_always),
                                                                           nwords, aw, beta, one_minus_beta, reinterpret_cast<unsigned*>(out));
  VU_RETURN_LAUNCH();
}

extern "C" int vu_bgdiff_gray(const uint8_t* frame, const uint8_t* bg, int64_t npix, int64_t bg_npix, int thr, uint8_t* gray,
                              vu_stream_t stream) {
  VU_REQUIRE(frame && bg && gray && npix >= 0 && bg_npix > 0);
  if (npix == 0) return VU_OK;
  if (needs_px(npix, bg_npix, {frame, bg, gray})) {
    bgdiff_gray_px_kernel<<<grid_for(npix, THREADS, 8), THREADS, 0, S(stream)>>>(frame, bg, npix, bg_npix, thr, gray);
    VU_RETURN_LAUNCH();
  }
  bgdiff_gray_kernel<<<grid_for(npix / 4, THREADS, 8), THREADS, 0, S(stream)>>>(frame, bg, npix / 4, bg_npix / 4, thr, gray);
  VU_RETURN_LAUNCH();
}
