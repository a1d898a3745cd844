// Fused kernels of the batched per-frame path (clip pipelines):
//
//  vu_cross_chain_u8   chains of 3x3-cross dilations / erosions (the reference
//                      only ever iterates MORPH_ELLIPSE(3,3), i.e. L1 diamonds:
//                      maskprocess.py:7-34 called from colorfiltering/agent.py:
//                      281-282, trimap/agent.py:55-56, bg.py:77) executed
//                      entirely in shared memory: the tile plus a halo of one
//                      pixel per pass is staged once, every pass is a
//                      VIMNMX.U16x2 stencil over even/odd pixel planes, and
//                      only the final tile goes back to HBM.  Optional fused
//                      prologue: postprocess' adaptive threshold (agent.py:
//                      277-280).  Optional fused epilogue: the trimap
//                      classification of trimap/agent.py:54-58 from a dilate
//                      chain and an erode chain of the same tile.
//  vu_cf_lowres        BGR2HSV + exact 2x / 4x cv2 down-scale + tabulated
//                      mixture evaluation + postprocess statistics in one pass
//                      over the full-resolution frame (agent.py:310-320, 277-279).
//  vu_resize_up_u8     cv2 fixed-point bilinear up-scale of single-channel
//                      maps with the coefficient math hoisted out of the pixel
//                      loop; optional fused snap (trimap/agent.py:60) and
//                      fuzzy override (:100).
//  vu_fuzzy_count      is_pixel_inrange (bg colour) + fuzzy area + the two
//                      counts of trimap/agent.py:90-94 in one pass.
#include "vu_common.cuh"

namespace vu {
namespace {

// ---------------------------------------------------------------------------------
// cross chains
// ---------------------------------------------------------------------------------
constexpr int CC_THREADS = 256;
constexpr int CC_TW = 128;             // output tile width  (pixels)
constexpr int CC_TH = 32;              // output tile height (rows)
constexpr int CC_MAXPASS = 12;         // most passes (= halo rows) one launch can chain

// tile geometry for a chain of `halo` passes: halo rows above/below, ceil(halo/4) 4-pixel groups left/right
struct Geo {
  int halo, hg, gw, rows, cells;
};
__host__ __device__ inline Geo make_geo(int halo) {
  Geo g;
  g.halo = halo;
  g.hg = (halo + 3) / 4;
  g.gw = CC_TW / 4 + 2 * g.hg;
  g.rows = CC_TH + 2 * halo;
  g.cells = g.gw * g.rows;
  return g;
}

struct Chain {
  int nseg;
  int op[4];     // VU_DILATE / VU_ERODE
  int iters[4];
};

// one 4-pixel group = (E, O): E = pixels 0,2 and O = pixels 1,3 in 16-bit lanes
__device__ __forceinline__ uint2 split4(unsigned w) { return make_uint2(w & 0x00FF00FFu, (w >> 8) & 0x00FF00FFu); }
__device__ __forceinline__ unsigned merge4(uint2 g) { return g.x | (g.y << 8); }

template <bool DIL>
__device__ __forceinline__ unsigned mm2(unsigned a, unsigned b) { return DIL ? __vmaxu2(a, b) : __vminu2(a, b); }

// one pass over the whole staged tile: out = cross-max/min(in); cells outside the image are reset to the identity
// `ident` = identity of the NEXT pass' operation: that is what its out-of-image neighbours must read
template <bool DIL>
__device__ __forceinline__ void cross_pass(const uint2* __restrict__ in, uint2* __restrict__ out, const Geo g, int x0, int y0, int h, int w,
                                           unsigned ident) {
  for (int i = threadIdx.x; i < g.cells; i += CC_THREADS) {
    const int row = i / g.gw, gx = i - row * g.gw;
    const int up = row > 0 ? i - g.gw : i, dn = row < g.rows - 1 ? i + g.gw : i;
    const uint2 c = in[i], u = in[up], d = in[dn];
    const unsigned prevO = in[gx > 0 ? i - 1 : i].y, nextE = in[gx < g.gw - 1 ? i + 1 : i].x;
    // left neighbours of pixels (0,2) are (prev.3, 1); right neighbours of pixels (1,3) are (2, next.0)
    const unsigned le = __byte_perm(prevO, c.y, 0x5432);   // [prevO.hi16, O.lo16]
    const unsigned ro = __byte_perm(c.x, nextE, 0x5432);   // [E.hi16, nextE.lo16]
    unsigned e = mm2<DIL>(mm2<DIL>(c.x, u.x), mm2<DIL>(d.x, mm2<DIL>(le, c.y)));
    unsigned o = mm2<DIL>(mm2<DIL>(c.y, u.y), mm2<DIL>(d.y, mm2<DIL>(c.x, ro)));
    const int gy = y0 + row, px = x0 + 4 * gx;
    if ((unsigned)gy >= (unsigned)h || px < 0 || px + 3 >= w) {
      const bool rv = (unsigned)gy < (unsigned)h;
      const unsigned m0 = (rv && px >= 0 && px < w) ? 0xFFFFu : 0u, m1 = (rv && px + 1 >= 0 && px + 1 < w) ? 0xFFFFu : 0u;
      const unsigned m2 = (rv && px + 2 >= 0 && px + 2 < w) ? 0xFFFF0000u : 0u, m3 = (rv && px + 3 >= 0 && px + 3 < w) ? 0xFFFF0000u : 0u;
      const unsigned me = m0 | m2, mo = m1 | m3;
      e = (e & me) | (ident & ~me);
      o = (o & mo) | (ident & ~mo);
    }
    out[i] = make_uint2(e, o);
  }
}

__device__ __forceinline__ uint2* run_chain(const Chain& ch, uint2* a, uint2* b, const Geo g, int x0, int y0, int h, int w) {
  for (int s = 0; s < ch.nseg; ++s)
    for (int it = 0; it < ch.iters[s]; ++it) {
      // operation of the pass after this one (same segment, or the next non-empty segment)
      int nop = ch.op[s];
      if (it + 1 == ch.iters[s])
        for (int t = s + 1; t < ch.nseg; ++t)
          if (ch.iters[t] > 0) { nop = ch.op[t]; break; }
      const unsigned ident = nop == VU_DILATE ? 0u : 0x00FF00FFu;
      if (ch.op[s] == VU_DILATE) cross_pass<true>(a, b, g, x0, y0, h, w, ident);
      else cross_pass<false>(a, b, g, x0, y0, h, w, ident);
      __syncthreads();
      uint2* t = a; a = b; b = t;
    }
  return a;  // buffer holding the result
}

// MODE 0: dst = chain(src)          MODE 1: dst = trimap classify(dilate^r(src), erode^r(src))
template <int MODE>
__global__ void __launch_bounds__(CC_THREADS) cross_chain_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w, Chain ch,
                                                                 const unsigned long long* __restrict__ stats2, double thr_ratio, int halo) {
  extern __shared__ __align__(16) unsigned char cc_smem[];
  const Geo g = make_geo(halo);
  uint2* buf0 = reinterpret_cast<uint2*>(cc_smem);
  uint2* buf1 = buf0 + g.cells;
  uint2* orig = buf1 + g.cells;   // MODE 1 only
  uint2* resA = orig + g.cells;   // MODE 1 only
  const int64_t frame = (int64_t)blockIdx.z * h * w;
  const int x0 = blockIdx.x * CC_TW - 4 * g.hg, y0 = blockIdx.y * CC_TH - g.halo;
  // postprocess threshold fused into the load (colorfiltering/agent.py:277-280)
  bool have = false;
  double thr = 0.0;
  if (stats2) {
    const unsigned long long sum = stats2[2 * blockIdx.z], cnt = stats2[2 * blockIdx.z + 1];
    have = cnt != 0;
    if (have) thr = __dmul_rn(__ddiv_rn((double)sum, (double)cnt), thr_ratio);
  }
  const bool vec = (w & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
  int first_op = VU_DILATE;
  for (int t = ch.nseg - 1; t >= 0; --t)
    if (ch.iters[t] > 0) first_op = ch.op[t];
  const int ident_first = first_op == VU_DILATE ? 0 : 255;
  for (int i = threadIdx.x; i < g.cells; i += CC_THREADS) {
    const int row = i / g.gw, gx = i - row * g.gw;
    const int gy = y0 + row, px = x0 + 4 * gx;
    unsigned word;
    if ((unsigned)gy < (unsigned)h && px >= 0 && px + 3 < w && vec) {
      word = __ldg(reinterpret_cast<const unsigned*>(src + frame + (int64_t)gy * w + px));
    } else {
      word = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = px + k;
        const unsigned v = ((unsigned)gy < (unsigned)h && x >= 0 && x < w) ? __ldg(src + frame + (int64_t)gy * w + x) : (unsigned)ident_first;
        word |= v << (8 * k);
      }
    }
    if (have) {
      unsigned r = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned v = (word >> (8 * k)) & 255u;
        r |= (((double)v < thr) ? 0u : v) << (8 * k);
      }
      word = r;
    }
    const uint2 sp = split4(word);
    buf0[i] = sp;
    if (MODE == 1) orig[i] = sp;
  }
  __syncthreads();
  uint2* res;
  if (MODE == 0) {
    res = run_chain(ch, buf0, buf1, g, x0, y0, h, w);
  } else {
    Chain d{1, {VU_DILATE, 0, 0, 0}, {ch.iters[0], 0, 0, 0}};
    Chain e{1, {VU_ERODE, 0, 0, 0}, {ch.iters[0], 0, 0, 0}};
    uint2* ra = run_chain(d, buf0, buf1, g, x0, y0, h, w);
    for (int i = threadIdx.x; i < g.cells; i += CC_THREADS) {
      resA[i] = ra[i];
      // out-of-image cells must start at the erosion identity
      const int row = i / g.gw, gx = i - row * g.gw;
      const int gy = y0 + row, px = x0 + 4 * gx;
      uint2 og = orig[i];
      if ((unsigned)gy >= (unsigned)h || px < 0 || px + 3 >= w) {
        unsigned wv = merge4(og), r = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int x = px + k;
          const bool in = (unsigned)gy < (unsigned)h && x >= 0 && x < w;
          r |= (in ? ((wv >> (8 * k)) & 255u) : 255u) << (8 * k);
        }
        og = split4(r);
      }
      buf0[i] = og;
    }
    __syncthreads();
    res = run_chain(e, buf0, buf1, g, x0, y0, h, w);
  }
  // write the inner tile
  for (int i = threadIdx.x; i < (CC_TW / 4) * CC_TH; i += CC_THREADS) {
    const int ty = i / (CC_TW / 4), tg = i - ty * (CC_TW / 4);
    const int gy = blockIdx.y * CC_TH + ty, px = blockIdx.x * CC_TW + 4 * tg;
    if (gy >= h || px >= w) continue;
    const int cell = (ty + g.halo) * g.gw + tg + g.hg;
    unsigned word = merge4(res[cell]);
    if (MODE == 1) {
      const unsigned dil = merge4(resA[cell]);
      unsigned r = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned dv = (dil >> (8 * k)) & 255u, ev = (word >> (8 * k)) & 255u;
        r |= (dv < 128u ? 0u : (ev > 127u ? 255u : 128u)) << (8 * k);
      }
      word = r;
    }
    uint8_t* o = dst + frame + (int64_t)gy * w + px;
    if (px + 3 < w && vec && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
      *reinterpret_cast<unsigned*>(o) = word;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (px + k < w) o[k] = (uint8_t)(word >> (8 * k));
    }
  }
}

}  // namespace
}  // namespace vu

using namespace vu;

// ops[i] / iters[i]: up to 4 chained segments, every segment `iters` passes of
// the 3x3 cross (== cv2 MORPH_ELLIPSE(3,3) with iterations=iters).  stats2 (may
// be NULL): per-frame {sum, count}; elements below thr_ratio*sum/count are zeroed
// on load.  Total passes <= 12.
extern "C" int vu_cross_chain_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int nseg, const int32_t* ops, const int32_t* iters,
                                 const uint64_t* stats2, double thr_ratio, vu_stream_t stream) {
  VU_REQUIRE(src && dst && ops && iters && n >= 0 && h > 0 && w > 0 && nseg >= 1 && nseg <= 4);
  Chain ch{};
  ch.nseg = nseg;
  int total = 0;
  for (int i = 0; i < nseg; ++i) {
    VU_REQUIRE(ops[i] == VU_DILATE || ops[i] == VU_ERODE);
    VU_REQUIRE(iters[i] >= 0);
    ch.op[i] = ops[i];
    ch.iters[i] = iters[i];
    total += iters[i];
  }
  if (total > CC_MAXPASS) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  dim3 grid((w + CC_TW - 1) / CC_TW, (h + CC_TH - 1) / CC_TH, n);
  const size_t smem = 2 * sizeof(uint2) * make_geo(total).cells;
  cross_chain_kernel<0><<<grid, CC_THREADS, smem, S(stream)>>>(src, dst, h, w, ch, reinterpret_cast<const unsigned long long*>(stats2), thr_ratio,
                                                               total);
  VU_RETURN_LAUNCH();
}

// trimap/agent.py:53-58 at working resolution: dst = classify(dilate^iters(src), erode^iters(src))
extern "C" int vu_trimap_core_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int iters, vu_stream_t stream) {
  VU_REQUIRE(src && dst && n >= 0 && h > 0 && w > 0 && iters >= 0);
  if (iters > CC_MAXPASS) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  Chain ch{1, {VU_DILATE, 0, 0, 0}, {iters, 0, 0, 0}};
  dim3 grid((w + CC_TW - 1) / CC_TW, (h + CC_TH - 1) / CC_TH, n);
  const size_t smem = 4 * sizeof(uint2) * make_geo(iters).cells;
  static bool configured = false;
  if (!configured) {
    int e = record_cuda(cudaFuncSetAttribute(cross_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(4 * sizeof(uint2) * make_geo(CC_MAXPASS).cells)));
    if (e) return e;
    configured = true;
  }
  cross_chain_kernel<1><<<grid, CC_THREADS, smem, S(stream)>>>(src, dst, h, w, ch, nullptr, 0.0, iters);
  VU_RETURN_LAUNCH();
}
