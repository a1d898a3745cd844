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
#include <cstdlib>

#include "vu_common.cuh"
#include "vu_cross_march.cuh"

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
  if (n > 65535) return VU_ERR_UNSUPPORTED;
  // the chains the reference runs go to the register-marching kernels (vu_cross_march.cuh); anything else to the
  // generic shared-memory kernel below
  {
    unsigned emask = 0;
    int k = 0;
    for (int i = 0; i < nseg; ++i)
      for (int j = 0; j < iters[i]; ++j, ++k)
        if (ops[i] == VU_ERODE) emask |= 1u << k;
    const auto* st = reinterpret_cast<const unsigned long long*>(stats2);
    const unsigned all = total > 0 ? (1u << total) - 1u : 0u;
#define VU_MARCH(NP, EM) \
  if (total == NP && emask == (EM)) return march::launch<NP, (EM), 0>(src, dst, n, h, w, st, thr_ratio, S(stream));
    VU_MARCH(8, 0x3Cu)   // colour filter postprocess: dilate 2, erode 2, erode 2, dilate 2
    VU_MARCH(1, 0x0u) VU_MARCH(2, 0x0u) VU_MARCH(3, 0x0u) VU_MARCH(4, 0x0u) VU_MARCH(5, 0x0u)
    VU_MARCH(1, 0x1u) VU_MARCH(2, 0x3u) VU_MARCH(3, 0x7u) VU_MARCH(4, 0xFu) VU_MARCH(5, 0x1Fu)
#undef VU_MARCH
    (void)all;
  }
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
  if (n > 65535) return VU_ERR_UNSUPPORTED;
  switch (iters) {   // register-marching kernels for the usual radii
    case 1: return march::launch<1, 0u, 1>(src, dst, n, h, w, nullptr, 0.0, S(stream));
    case 2: return march::launch<2, 0u, 1>(src, dst, n, h, w, nullptr, 0.0, S(stream));
    case 3: return march::launch<3, 0u, 1>(src, dst, n, h, w, nullptr, 0.0, S(stream));
    case 4: return march::launch<4, 0u, 1>(src, dst, n, h, w, nullptr, 0.0, S(stream));
    case 5: return march::launch<5, 0u, 1>(src, dst, n, h, w, nullptr, 0.0, S(stream));
    default: break;
  }
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

// =================================================================================
// cf_lowres, resize_up, fuzzy_count, trimap_src_lo
// =================================================================================
namespace vu {
namespace {

constexpr int FT = 256;

__device__ __forceinline__ void warp_block_atomic2(unsigned long long a, unsigned long long b, unsigned long long* dst) {
  // reduce two counters over the block, one atomic pair per block
  __shared__ unsigned long long sh[2][FT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < FT / 32 ? sh[0][threadIdx.x] : 0;
    b = threadIdx.x < FT / 32 ? sh[1][threadIdx.x] : 0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, o);
      b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if (threadIdx.x == 0) {
      if (a) atomicAdd(dst, a);
      if (b) atomicAdd(dst + 1, b);
    }
  }
}

// S = 2: thread = two low-res pixels (a 2x4 block of the frame); S = 4: one low-res pixel (centre 2x2 of a 4x4 block)
template <int S>
__global__ void __launch_bounds__(FT) cf_lowres_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ masks, int h, int w, int th, int tw,
                                                       const uint8_t* __restrict__ lut3d, uint8_t* __restrict__ alpha_lo,
                                                       unsigned long long* __restrict__ stats2) {
  __shared__ HsvTab tab;
  hsv_tab_init(tab);
  __syncthreads();
  const int n = blockIdx.y;
  const int per_row = (S == 2) ? tw / 2 : tw;           // threads per low-res row
  const int64_t items = (int64_t)th * per_row;
  const uint8_t* fr = frames + (int64_t)n * h * w * 3;
  const uint8_t* mk = masks + (int64_t)n * h * w;
  uint8_t* out = alpha_lo + (int64_t)n * th * tw;
  unsigned long long sum = 0, cnt = 0;
  for (int64_t it = (int64_t)blockIdx.x * FT + threadIdx.x; it < items; it += (int64_t)gridDim.x * FT) {
    const int y = (int)(it / per_row), xg = (int)(it - (int64_t)y * per_row);
    const int r0 = (S == 2) ? 2 * y : 4 * y + 1;
    const int c0 = 4 * xg;                              // first full-res column of this thread's 4-pixel span
    int hs[4][3];                                       // S=2: [row][...] handled below
    int px[2][12];
    unsigned mw[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const unsigned* p = reinterpret_cast<const unsigned*>(fr + ((int64_t)(r0 + r) * w + c0) * 3);
      unpack12(__ldg(p), __ldg(p + 1), __ldg(p + 2), px[r]);
      mw[r] = __ldg(reinterpret_cast<const unsigned*>(mk + (int64_t)(r0 + r) * w + c0));
    }
    (void)hs;
    if (S == 2) {
      unsigned res = 0;
#pragma unroll
      for (int o = 0; o < 2; ++o) {                     // two outputs: columns (0,1) and (2,3)
        int acc[3] = {0, 0, 0};
        int macc = 0;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int q = 2 * o + k;
            int hh, ss, vv;
            bgr2hsv_px(px[r][3 * q], px[r][3 * q + 1], px[r][3 * q + 2], tab, hh, ss, vv);
            acc[0] += hh; acc[1] += ss; acc[2] += vv;
            macc += (mw[r] >> (8 * q)) & 255;
          }
        const int hh = (acc[0] + 2) >> 2, ss = (acc[1] + 2) >> 2, vv = (acc[2] + 2) >> 2, mm = (macc + 2) >> 2;
        const unsigned a = __ldg(lut3d + ((hh << 16) | (ss << 8) | vv));
        if (a > 128 && mm > 0) { sum += a; ++cnt; }
        res |= a << (8 * o);
      }
      *reinterpret_cast<unsigned short*>(out + (int64_t)y * tw + 2 * xg) = (unsigned short)res;
    } else {
      int acc[3] = {0, 0, 0};
      int macc = 0;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 1; q <= 2; ++q) {
          int hh, ss, vv;
          bgr2hsv_px(px[r][3 * q], px[r][3 * q + 1], px[r][3 * q + 2], tab, hh, ss, vv);
          acc[0] += hh; acc[1] += ss; acc[2] += vv;
          macc += (mw[r] >> (8 * q)) & 255;
        }
      const int hh = (acc[0] + 2) >> 2, ss = (acc[1] + 2) >> 2, vv = (acc[2] + 2) >> 2, mm = (macc + 2) >> 2;
      const unsigned a = __ldg(lut3d + ((hh << 16) | (ss << 8) | vv));
      if (a > 128 && mm > 0) { sum += a; ++cnt; }
      out[(int64_t)y * tw + xg] = (uint8_t)a;
    }
  }
  warp_block_atomic2(sum, cnt, stats2 + 2 * n);
}

// S = 2, 16 full-resolution columns per thread (w % 16 == 0, 16-byte aligned rows): 128-bit loads, eight outputs,
// all the loads of a thread in flight before the first conversion
__global__ void __launch_bounds__(FT, 4) cf_lowres2_wide_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ masks, int h, int w, int th,
                                                             int tw, const uint8_t* __restrict__ lut3d, uint8_t* __restrict__ alpha_lo,
                                                             unsigned long long* __restrict__ stats2, unsigned long long* __restrict__ mcounts2) {
  __shared__ HsvTab tab;
  hsv_tab_init(tab);
  __syncthreads();
  const int n = blockIdx.y;
  const int per_row = tw / 8;
  const int64_t items = (int64_t)th * per_row;
  const uint8_t* fr = frames + (int64_t)n * h * w * 3;
  const uint8_t* mk = masks + (int64_t)n * h * w;
  uint8_t* out = alpha_lo + (int64_t)n * th * tw;
  unsigned long long sum = 0, cnt = 0;
  unsigned mgt = 0, mlt = 0;
  for (int64_t it = (int64_t)blockIdx.x * FT + threadIdx.x; it < items; it += (int64_t)gridDim.x * FT) {
    const int y = (int)(it / per_row), xg = (int)(it - (int64_t)y * per_row);
    uint4 fv[2][3], mv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint4* p = reinterpret_cast<const uint4*>(fr + ((int64_t)(2 * y + r) * w + 16 * xg) * 3);
#pragma unroll
      for (int k = 0; k < 3; ++k) fv[r][k] = ldg_stream16(p + k);
      mv[r] = ldg_stream16(mk + (int64_t)(2 * y + r) * w + 16 * xg);
    }
    if (mcounts2) {   // the early-out counts of agent.py:303-307 from the mask bytes that pass through anyway
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const unsigned ws[4] = {mv[r].x, mv[r].y, mv[r].z, mv[r].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // bytes > 128 and bytes < 128, four at a time: bit 7 of (x + 127) is set iff x > 128 for x < 129 + 128; handle
          // by the two cases of bit 7 of x
          const unsigned x = ws[k], hi7 = x & 0x80808080u, lo7 = x & 0x7F7F7F7Fu;
          const unsigned gt = hi7 & ((lo7 + 0x7F7F7F7Fu));          // bit 7 set and low 7 bits non-zero: x > 128
          mgt += __popc(gt & 0x80808080u);
          mlt += 4 - __popc(hi7);                                   // bit 7 clear: x < 128
        }
      }
    }
    unsigned res[2] = {0u, 0u};
#pragma unroll
    for (int o = 0; o < 8; ++o) {   // output o: full-resolution columns 2o, 2o+1 of both rows
      int acc0 = 0, acc1 = 0, acc2 = 0, macc = 0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const unsigned* fw = reinterpret_cast<const unsigned*>(fv[r]);
        const unsigned* mw = reinterpret_cast<const unsigned*>(&mv[r]);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int q = 2 * o + k;   // pixel of the 16
          const int b0 = 3 * q, b1 = 3 * q + 1, b2 = 3 * q + 2;
          const int B = byte_fma4(fw[b0 >> 2], b0 & 3), G = byte_fma4(fw[b1 >> 2], b1 & 3), R = byte_fma4(fw[b2 >> 2], b2 & 3);
          int hh, ss, vv;
          bgr2hsv_px4(B, G, R, tab, hh, ss, vv);
          acc0 += hh; acc1 += ss; acc2 += vv;
          if (k == 0) macc = (int)__dp4a(mw[q >> 2], 0x0101u << (8 * (q & 3)), (unsigned)macc);   // both columns of the pair: one word
        }
      }
      const int hh = (acc0 + 2) >> 2, ss = (acc1 + 2) >> 2, vv = (acc2 + 2) >> 2, mm = (macc + 2) >> 2;
      const unsigned a = __ldg(lut3d + ((hh << 16) | (ss << 8) | vv));
      if (a > 128 && mm > 0) { sum += a; ++cnt; }
      res[o >> 2] |= a << (8 * (o & 3));
    }
    *reinterpret_cast<uint2*>(out + (int64_t)y * tw + 8 * xg) = make_uint2(res[0], res[1]);
  }
  warp_block_atomic2(sum, cnt, stats2 + 2 * n);
  if (mcounts2) warp_block_atomic2(mgt, mlt, mcounts2 + 2 * n);
}

// S = 4 (4K at the default working resolution), 16 full-resolution columns x 4 rows per thread: four outputs, each the
// rounded mean of the centre 2x2 of its 4x4 block (rows 4y+1, 4y+2, columns 4x+1, 4x+2: what cv2's bilinear comes to at
// exactly 4x).  Only the two centre rows of the frame are read; with mcounts2 all four rows of the mask
// are, for the early-out counts of agent.py:303-307.
__global__ void __launch_bounds__(FT, 4) cf_lowres4_wide_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ masks, int h, int w, int th,
                                                             int tw, const uint8_t* __restrict__ lut3d, uint8_t* __restrict__ alpha_lo,
                                                             unsigned long long* __restrict__ stats2, unsigned long long* __restrict__ mcounts2) {
  __shared__ HsvTab tab;
  hsv_tab_init(tab);
  __syncthreads();
  const int n = blockIdx.y;
  const int per_row = tw / 4;
  const int64_t items = (int64_t)th * per_row;
  const uint8_t* fr = frames + (int64_t)n * h * w * 3;
  const uint8_t* mk = masks + (int64_t)n * h * w;
  uint8_t* out = alpha_lo + (int64_t)n * th * tw;
  unsigned long long sum = 0, cnt = 0;
  unsigned mgt = 0, mlt = 0;
  for (int64_t it = (int64_t)blockIdx.x * FT + threadIdx.x; it < items; it += (int64_t)gridDim.x * FT) {
    const int y = (int)(it / per_row), xg = (int)(it - (int64_t)y * per_row);
    uint4 fv[2][3], mv[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint4* p = reinterpret_cast<const uint4*>(fr + ((int64_t)(4 * y + 1 + r) * w + 16 * xg) * 3);
#pragma unroll
      for (int k = 0; k < 3; ++k) fv[r][k] = ldg_stream16(p + k);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (mcounts2 || r == 1 || r == 2) mv[r] = ldg_stream16(mk + (int64_t)(4 * y + r) * w + 16 * xg);
    if (mcounts2) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const unsigned ws[4] = {mv[r].x, mv[r].y, mv[r].z, mv[r].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned x = ws[k], hi7 = x & 0x80808080u, lo7 = x & 0x7F7F7F7Fu;
          mgt += __popc(hi7 & (lo7 + 0x7F7F7F7Fu) & 0x80808080u);   // bit 7 set and low 7 bits non-zero: x > 128
          mlt += 4 - __popc(hi7);                                   // bit 7 clear: x < 128
        }
      }
    }
    unsigned res = 0u;
#pragma unroll
    for (int o = 0; o < 4; ++o) {   // output o: full-resolution columns 4o+1, 4o+2 of rows 4y+1, 4y+2
      int acc0 = 0, acc1 = 0, acc2 = 0, macc = 0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const unsigned* fw = reinterpret_cast<const unsigned*>(fv[r]);
        const unsigned* mw = reinterpret_cast<const unsigned*>(&mv[1 + r]);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int q = 4 * o + 1 + k;
          const int b0 = 3 * q, b1 = 3 * q + 1, b2 = 3 * q + 2;
          const int B = byte_fma4(fw[b0 >> 2], b0 & 3), G = byte_fma4(fw[b1 >> 2], b1 & 3), R = byte_fma4(fw[b2 >> 2], b2 & 3);
          int hh, ss, vv;
          bgr2hsv_px4(B, G, R, tab, hh, ss, vv);
          acc0 += hh; acc1 += ss; acc2 += vv;
        }
        macc = (int)__dp4a(mw[o], 0x00010100u, (unsigned)macc);   // bytes 1 and 2 of the block's mask word
      }
      const int hh = (acc0 + 2) >> 2, ss = (acc1 + 2) >> 2, vv = (acc2 + 2) >> 2, mm = (macc + 2) >> 2;
      const unsigned a = __ldg(lut3d + ((hh << 16) | (ss << 8) | vv));
      if (a > 128 && mm > 0) { sum += a; ++cnt; }
      res |= a << (8 * o);
    }
    *reinterpret_cast<unsigned*>(out + (int64_t)y * tw + 4 * xg) = res;
  }
  warp_block_atomic2(sum, cnt, stats2 + 2 * n);
  if (mcounts2) warp_block_atomic2(mgt, mlt, mcounts2 + 2 * n);
}

struct AxisC {
  int i0, i1, w0, w1;
};
__device__ __forceinline__ double up_scale(int dst, int src) {
  const double inv_scale = (double)dst / (double)src;
  return 1.0 / inv_scale;
}
__device__ __forceinline__ AxisC up_axis(int d, double scale, int src, bool reset) {
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  int i0 = (int)floorf(f);
  float fr = __fsub_rn(f, (float)i0);
  AxisC a;
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

constexpr int RU_TW = 128, RU_TH = 128;   // dst tile; 256 threads = 32 column groups (4 pixels) x 8 runs of 16 consecutive rows
constexpr int RU_RUN = RU_TH / 8;
// cv2's fixed-point bilinear: row[x] = S[y][x0]*a0 + S[y][x1]*a1;
// dst = (((b0*(R0>>4))>>16) + ((b1*(R1>>4))>>16) + 2) >> 2.  A thread owns 4 destination columns and walks down 8
// consecutive destination rows: the horizontal pass of a source row is computed once and kept in registers for all the
// destination rows that use it (up-scaling: most consecutive rows share their source rows).  Source taps come through
// the read-only path: the low-resolution maps are L2 / L1 resident.
// MODE 0: plain   MODE 1: snap (0<v<255 -> 128), then 128 where flags[n]==0 && fuzzy
template <int MODE>
__global__ void __launch_bounds__(FT) resize_up_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst, int dh, int dw,
                                                       const uint8_t* __restrict__ fuzzy, const uint8_t* __restrict__ flags,
                                                       const uint8_t* __restrict__ alt_src, const uint8_t* __restrict__ alt_flags) {
  __shared__ AxisC ytab[RU_TH];
  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * RU_TH, tx0 = blockIdx.x * RU_TW;
  const double ysc = up_scale(dh, sh), xsc = up_scale(dw, sw);
  if (threadIdx.x < RU_TH && ty0 + threadIdx.x < dh) ytab[threadIdx.x] = up_axis(ty0 + threadIdx.x, ysc, sh, false);
  const int gx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int x = tx0 + 4 * gx;
  AxisC xa[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) xa[k] = up_axis(min(x + k, dw - 1), xsc, sw, true);
  __syncthreads();
  if (x >= dw) return;
  const uint8_t* s = src + (int64_t)n * sh * sw;
  uint8_t* d = dst + (int64_t)n * dh * dw;
  const bool use_alt = alt_flags && alt_flags[n] != 0;
  const bool ens = (MODE == 1) && fuzzy && flags && flags[n] == 0;
  const bool vec = (dw & 3) == 0 && x + 3 < dw && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0);
  // horizontal pass of source row sy, >> 4 as the vertical pass wants it.  Returns the common value of the eight taps, or
  // -1: where all the taps of both source rows agree the interpolation is that value exactly (the two weighted
  // halves lose less than 2 of 4v + 2 together), which is most of a matte or a trimap
  auto hrow = [&](int sy, int (&R)[4]) -> int {
    const uint8_t* r = s + (int64_t)sy * sw;
    int first = 0;
    bool same = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int a = (int)__ldg(r + xa[k].i0), b = (int)__ldg(r + xa[k].i1);
      if (k == 0) first = a;
      same = same && a == first && b == first;
      R[k] = (a * xa[k].w0 + b * xa[k].w1) >> 4;
    }
    return same ? first : -1;
  };
  // (the shortcut needs weights that sum to 2048 on both axes: true for every scale the pipelines use, checked anyway)
  const bool xsum_ok = xa[0].w0 + xa[0].w1 == 2048 && xa[1].w0 + xa[1].w1 == 2048 && xa[2].w0 + xa[2].w1 == 2048 && xa[3].w0 + xa[3].w1 == 2048;
  int R0[4], R1[4];
  int u0 = -1, u1 = -1;     // common tap value of the rows in R0 / R1 (or -1)
  int y0c = -1, y1c = -1;   // source rows held in R0 / R1
#pragma unroll 1
  for (int rr = 0; rr < RU_RUN; ++rr) {
    const int r = ry * RU_RUN + rr;
    const int y = ty0 + r;
    if (y >= dh) break;
    unsigned word = 0;
    if (use_alt) {
      if (vec) word = __ldg(reinterpret_cast<const unsigned*>(alt_src + ((int64_t)n * dh + y) * dw + x));
      else
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (x + k < dw) word |= (unsigned)__ldg(alt_src + ((int64_t)n * dh + y) * dw + x + k) << (8 * k);
    } else {
      const AxisC ya = ytab[r];
      if (ya.i0 != y0c) {
        if (ya.i0 == y1c) {
#pragma unroll
          for (int k = 0; k < 4; ++k) R0[k] = R1[k];
          u0 = u1;
        } else {
          u0 = hrow(ya.i0, R0);
        }
        y0c = ya.i0;
      }
      if (ya.i1 != y1c) {
        if (ya.i1 == y0c) {
#pragma unroll
          for (int k = 0; k < 4; ++k) R1[k] = R0[k];
          u1 = u0;
        } else {
          u1 = hrow(ya.i1, R1);
        }
        y1c = ya.i1;
      }
      const bool flat = u0 >= 0 && u0 == u1 && xsum_ok && ya.w0 + ya.w1 == 2048;
      unsigned fz = 0;
      if (ens) {
        if (vec) fz = __ldg(reinterpret_cast<const unsigned*>(fuzzy + ((int64_t)n * dh + y) * dw + x));
        else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (x + k < dw) fz |= (unsigned)__ldg(fuzzy + ((int64_t)n * dh + y) * dw + x + k) << (8 * k);
        }
      }
      if (flat) {
        int v = u0;
        if (MODE == 1 && v > 0 && v < 255) v = 128;
        word = (unsigned)v * 0x01010101u;
        if (MODE == 1 && fz) {
          const unsigned nz = ((fz & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | fz;       // bit 7 of every non-zero fuzzy byte
          const unsigned sel = ((nz >> 7) & 0x01010101u) * 255u;
          word = (word & ~sel) | (0x80808080u & sel);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int v = (((ya.w0 * R0[k]) >> 16) + ((ya.w1 * R1[k]) >> 16) + 2) >> 2;
          v = min(255, max(0, v));
          if (MODE == 1) {
            if (v > 0 && v < 255) v = 128;
            if ((fz >> (8 * k)) & 255u) v = 128;
          }
          word |= (unsigned)v << (8 * k);
        }
      }
    }
    uint8_t* o = d + (int64_t)y * dw + x;
    if (vec) {
      *reinterpret_cast<unsigned*>(o) = word;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x + k < dw) o[k] = (uint8_t)(word >> (8 * k));
    }
  }
}

// Exact 2x / 4x up-scales (1080p and 4K from the 540 x 960 working resolution), plain mode: the interpolation phases
// and weights are compile-time constants (2x: 512/1536; 4x: 256/768/1280/1792 of 2048), the taps of a pixel are the 2x2
// neighbourhood towards its quadrant with replicate clamping (at the borders cv2 resets the fraction horizontally and
// clamps the index vertically: with both taps on the same pixel and weights that sum to 2048 that is the same
// number), and a thread owns 8 destination columns for a run of rows: one 64-bit store per row, the horizontal pass of
// a source row computed once for all the destination rows that use it.
template <int SC>
__global__ void __launch_bounds__(FT) resize_up_int_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst,
                                                           const uint8_t* __restrict__ alt_src, const uint8_t* __restrict__ alt_flags, int run) {
  constexpr int NS = 8 / SC;   // source columns under the 8 destination columns
  const int n = blockIdx.z;
  const int dh = SC * sh, dw = SC * sw;
  const int tx = blockIdx.x * 32 + (threadIdx.x & 31);   // 8-column group
  const int ry = blockIdx.y * (FT / 32) + (threadIdx.x >> 5);
  const int x0 = 8 * tx, y0 = ry * run;
  if (x0 >= dw || y0 >= dh) return;
  const int y1 = min(dh, y0 + run);
  uint8_t* d = dst + (int64_t)n * dh * dw;
  if (alt_flags && alt_flags[n] != 0) {   // early-out frame: copied from the full-resolution alternative
    for (int y = y0; y < y1; ++y)
      *reinterpret_cast<uint2*>(d + (int64_t)y * dw + x0) = __ldg(reinterpret_cast<const uint2*>(alt_src + ((int64_t)n * dh + y) * dw + x0));
    return;
  }
  const uint8_t* s = src + (int64_t)n * sh * sw;
  const int c0 = x0 / SC;
  // weights of (left tap, right tap) per phase, in 1/2048
  auto w0 = [](int ph) { return SC == 2 ? (ph == 0 ? 512 : 1536) : (ph == 0 ? 768 : (ph == 1 ? 256 : (ph == 2 ? 1792 : 1280))); };
  auto hrow = [&](int sy, int (&R)[8]) {
    int t[NS + 2];   // source columns c0-1 .. c0+NS, replicated at the borders
    const uint8_t* r = s + (int64_t)sy * sw;
#pragma unroll
    for (int k = 0; k < NS + 2; ++k) t[k] = (int)__ldg(r + min(max(c0 - 1 + k, 0), sw - 1));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = k / SC, ph = k % SC;                     // column c0 + c, phase ph
      const int il = ph < SC / 2 ? c : c + 1;                // index into t of the left tap (t[c+1] is column c0 + c)
      R[k] = (t[il] * w0(ph) + t[il + 1] * (2048 - w0(ph))) >> 4;
    }
  };
  int Ra[8], Rb[8];
  int ya = -1, yb = -1;
  for (int y = y0; y < y1; ++y) {
    const int r = y / SC, ph = y % SC;
    const int ia = ph < SC / 2 ? max(r - 1, 0) : r, ib = ph < SC / 2 ? r : min(r + 1, sh - 1);
    if (ia != ya) {
      if (ia == yb) {
#pragma unroll
        for (int k = 0; k < 8; ++k) Ra[k] = Rb[k];
      } else {
        hrow(ia, Ra);
      }
      ya = ia;
    }
    if (ib != yb) {
      if (ib == ya) {
#pragma unroll
        for (int k = 0; k < 8; ++k) Rb[k] = Ra[k];
      } else {
        hrow(ib, Rb);
      }
      yb = ib;
    }
    const int b0 = SC == 2 ? (ph == 0 ? 512 : 1536) : (ph == 0 ? 768 : (ph == 1 ? 256 : (ph == 2 ? 1792 : 1280)));
    const int b1 = 2048 - b0;
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned v = (unsigned)((((b0 * Ra[k]) >> 16) + ((b1 * Rb[k]) >> 16) + 2) >> 2);   // <= 255 by construction
      if (k < 4) lo |= v << (8 * k);
      else hi |= v << (8 * (k - 4));
    }
    *reinterpret_cast<uint2*>(d + (int64_t)y * dw + x0) = make_uint2(lo, hi);
  }
}

// counts2[n] = {#(v > thr), #(v < thr)} in one vectorised pass (the two early-out tests of colorfiltering/agent.py:303-307)
__global__ void __launch_bounds__(FT) count_gt_lt_kernel(const uint8_t* __restrict__ src, int64_t per_item, int thr, unsigned long long* __restrict__ counts2) {
  const int n = blockIdx.y;
  const uint8_t* p = src + (int64_t)n * per_item;
  unsigned long long gt = 0, lt = 0;
  const int64_t nvec = ((reinterpret_cast<uintptr_t>(p) & 15) == 0) ? per_item / 16 : 0;
  const uint4* p16 = reinterpret_cast<const uint4*>(p);
  for (int64_t i = (int64_t)blockIdx.x * FT + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * FT) {
    const uint4 v = ldg_stream16(p16 + i);
    const unsigned ws[4] = {v.x, v.y, v.z, v.w};
    unsigned g = 0, l = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int x = (ws[k] >> (8 * b)) & 255;
        g += x > thr;
        l += x < thr;
      }
    gt += g;
    lt += l;
  }
  if (blockIdx.x == 0)
    for (int64_t i = nvec * 16 + threadIdx.x; i < per_item; i += FT) {
      const int x = __ldg(p + i);
      gt += x > thr;
      lt += x < thr;
    }
  warp_block_atomic2(gt, lt, counts2 + 2 * n);
}

// fuzzy01 = (alpha > 0) && lo <= HSV(frame) <= hi; counts2[n] = {#fuzzy, #alpha>0}
__global__ void __launch_bounds__(FT) fuzzy_count_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ alpha, int64_t ngroups,
                                                         int lo0, int lo1, int lo2, int hi0, int hi1, int hi2, uint8_t* __restrict__ fuzzy,
                                                         unsigned long long* __restrict__ counts2) {
  __shared__ HsvTab tab;
  hsv_tab_init(tab);
  __syncthreads();
  const int n = blockIdx.y;
  const unsigned* f4 = reinterpret_cast<const unsigned*>(frames) + (int64_t)n * ngroups * 3;
  const unsigned* a4 = reinterpret_cast<const unsigned*>(alpha) + (int64_t)n * ngroups;
  unsigned* o4 = reinterpret_cast<unsigned*>(fuzzy) + (int64_t)n * ngroups;
  unsigned long long nf = 0, np = 0;
  for (int64_t g = (int64_t)blockIdx.x * FT + threadIdx.x; g < ngroups; g += (int64_t)gridDim.x * FT) {
    const unsigned aw = __ldg(a4 + g);
    // fuzzy = (alpha > 0) && in-range: a warp whose 128 pixels are all outside the matte neither loads nor converts the frame
    if (__all_sync(__activemask(), aw == 0u)) {
      o4[g] = 0u;
      continue;
    }
    int c[12];
    unpack12(__ldg(f4 + 3 * g), __ldg(f4 + 3 * g + 1), __ldg(f4 + 3 * g + 2), c);
    unsigned w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int hh, ss, vv;
      bgr2hsv_px(c[3 * i], c[3 * i + 1], c[3 * i + 2], tab, hh, ss, vv);
      const bool in = (hh >= lo0) & (hh <= hi0) & (ss >= lo1) & (ss <= hi1) & (vv >= lo2) & (vv <= hi2);
      const bool pos = ((aw >> (8 * i)) & 255u) != 0;
      np += pos;
      nf += pos && in;
      w |= (unsigned)(pos && in) << (8 * i);
    }
    o4[g] = w;
  }
  warp_block_atomic2(nf, np, counts2 + 2 * n);
}

// the same at 16 pixels per thread (npix % 16 == 0, 16-byte aligned): 128-bit loads keep enough bytes in flight for
// a streaming kernel (4-byte loads, one per thread at a time, cap it near 1.2 TB/s whatever the arithmetic)
__global__ void __launch_bounds__(FT) fuzzy_count16_kernel(const uint4* __restrict__ frames, const uint4* __restrict__ alpha, int64_t ngroups, int lo0,
                                                           int lo1, int lo2, int hi0, int hi1, int hi2, uint4* __restrict__ fuzzy,
                                                           unsigned long long* __restrict__ counts2) {
  __shared__ HsvTab tab;
  hsv_tab_init(tab);
  __syncthreads();
  const int n = blockIdx.y;
  const uint4* f16 = frames + (int64_t)n * ngroups * 3;
  const uint4* a16 = alpha + (int64_t)n * ngroups;
  uint4* o16 = fuzzy + (int64_t)n * ngroups;
  unsigned long long nf = 0, np = 0;
  for (int64_t g = (int64_t)blockIdx.x * FT + threadIdx.x; g < ngroups; g += (int64_t)gridDim.x * FT) {
    const uint4 av = ldg_stream16(a16 + g);
    if ((av.x | av.y | av.z | av.w) == 0u) {   // 16 pixels outside the matte: neither load nor convert the frame
      o16[g] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    uint4 fv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) fv[k] = ldg_stream16(f16 + 3 * g + k);
    const unsigned* fw = reinterpret_cast<const unsigned*>(fv);
    const unsigned aw[4] = {av.x, av.y, av.z, av.w};
    unsigned ow[4];
    unsigned cf = 0, cp = 0;
#pragma unroll
    for (int w4 = 0; w4 < 4; ++w4) {
      unsigned w = 0;
      if (aw[w4] != 0u) {
        int c[12];
        unpack12(fw[3 * w4], fw[3 * w4 + 1], fw[3 * w4 + 2], c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int hh, ss, vv;
          bgr2hsv_px(c[3 * i], c[3 * i + 1], c[3 * i + 2], tab, hh, ss, vv);
          const bool in = (hh >= lo0) & (hh <= hi0) & (ss >= lo1) & (ss <= hi1) & (vv >= lo2) & (vv <= hi2);
          const bool pos = ((aw[w4] >> (8 * i)) & 255u) != 0;
          cp += pos;
          cf += pos && in;
          w |= (unsigned)(pos && in) << (8 * i);
        }
      }
      ow[w4] = w;
    }
    nf += cf;
    np += cp;
    o16[g] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
  warp_block_atomic2(nf, np, counts2 + 2 * n);
}

// nearest down-scale of the trimap source mask with the ensemble clearing fused:
// out[y][x] = (flags[n] == 0 && fuzzy[sy][sx]) ? 0 : mask[sy][sx]
__global__ void __launch_bounds__(FT) trimap_src_lo_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ fuzzy,
                                                           const uint8_t* __restrict__ flags, int h, int w, int th, int tw, uint8_t* __restrict__ out) {
  const int n = blockIdx.y;
  const double ifx = 1.0 / ((double)tw / (double)w), ify = 1.0 / ((double)th / (double)h);
  const bool ens = fuzzy && flags && flags[n] == 0;
  const int64_t total = (int64_t)th * tw;
  for (int64_t i = (int64_t)blockIdx.x * FT + threadIdx.x; i < total; i += (int64_t)gridDim.x * FT) {
    const int y = (int)(i / tw), x = (int)(i - (int64_t)y * tw);
    const int sx = min((int)floor(__dmul_rn((double)x, ifx)), w - 1);
    const int sy = min((int)floor(__dmul_rn((double)y, ify)), h - 1);
    const int64_t si = ((int64_t)n * h + sy) * w + sx;
    int v = __ldg(mask + si);
    if (ens && __ldg(fuzzy + si)) v = 0;
    out[(int64_t)n * total + i] = (uint8_t)v;
  }
}

// exact integer scales S = 2, 4 (w == S*tw, h == S*th, tw % 4 == 0): floor(x * S) == S*x, four outputs per thread from
// one 8- or 16-byte load of the source row (and of the fuzzy row), no float64 coordinates
template <int SC>
__global__ void __launch_bounds__(FT) trimap_src_lo_int_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ fuzzy,
                                                               const uint8_t* __restrict__ flags, int h, int w, int th, int tw,
                                                               uint8_t* __restrict__ out) {
  const int n = blockIdx.y;
  const bool ens = fuzzy && flags && flags[n] == 0;
  const int gpr = tw / 4;   // 4-output groups per row
  const int64_t total = (int64_t)th * gpr;
  auto pick = [](const uint8_t* p) -> unsigned {   // bytes 0, S, 2S, 3S of the 4S bytes at p
    if (SC == 2) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
      return __byte_perm(v.x, v.y, 0x6420);
    } else {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
      return (v.x & 255u) | ((v.y & 255u) << 8) | ((v.z & 255u) << 16) | (v.w << 24);
    }
  };
  for (int64_t i = (int64_t)blockIdx.x * FT + threadIdx.x; i < total; i += (int64_t)gridDim.x * FT) {
    const int y = (int)(i / gpr), xg = (int)(i - (int64_t)y * gpr);
    const int64_t si = ((int64_t)n * h + (int64_t)SC * y) * w + (int64_t)4 * SC * xg;
    unsigned v = pick(mask + si);
    if (ens) {
      const unsigned f = pick(fuzzy + si);
      // zero the bytes whose fuzzy byte is non-zero
      const unsigned nz = ((f & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | f;          // bit 7 of every non-zero byte
      const unsigned keep = ~(((nz >> 7) & 0x01010101u) * 255u);
      v &= keep;
    }
    *reinterpret_cast<unsigned*>(out + ((int64_t)n * th + y) * tw + 4 * xg) = v;
  }
}

inline dim3 frame_grid(int n, int64_t items_per_frame) {
  int64_t bx = (items_per_frame + FT - 1) / FT;
  int64_t cap = ((int64_t)device_sms() * 8 + n - 1) / n;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, n);
}

}  // namespace
}  // namespace vu

// BGR2HSV + cv2.resize (exact 2x or 4x) of the HSV image and of the mask +
// tabulated mixture evaluation + postprocess statistics (colorfiltering/
// agent.py:310-320, 277-279).  h == s*th and w == s*tw with s in {2, 4}.
extern "C" int vu_cf_lowres(const uint8_t* frames, const uint8_t* masks, int n, int h, int w, int th, int tw, const uint8_t* lut3d,
                            uint8_t* alpha_lo, uint64_t* stats2, uint64_t* mask_counts2, vu_stream_t stream) {
  VU_REQUIRE(frames && masks && lut3d && alpha_lo && stats2 && n >= 0 && h > 0 && w > 0 && th > 0 && tw > 0);
  const int s = (h == 2 * th && w == 2 * tw) ? 2 : ((h == 4 * th && w == 4 * tw) ? 4 : 0);
  if (s == 0 || (w & 3) || (tw & 1)) return VU_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(frames) & 3) || (reinterpret_cast<uintptr_t>(masks) & 3) || (reinterpret_cast<uintptr_t>(alpha_lo) & 1))
    return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  int e = record_cuda(cudaMemsetAsync(stats2, 0, sizeof(uint64_t) * 2 * n, S(stream)));
  if (e) return e;
  auto* st = reinterpret_cast<unsigned long long*>(stats2);
  const bool wide = s == 2 && (w % 16 == 0) && (tw % 8 == 0) && ((reinterpret_cast<uintptr_t>(frames) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(masks) & 15) == 0) && ((reinterpret_cast<uintptr_t>(alpha_lo) & 7) == 0);
  const bool wide4 = s == 4 && (w % 16 == 0) && (tw % 4 == 0) && ((reinterpret_cast<uintptr_t>(frames) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(masks) & 15) == 0) && ((reinterpret_cast<uintptr_t>(alpha_lo) & 3) == 0);
  if (mask_counts2) {   // only the 16-column kernels see every mask byte
    if (!wide && !wide4) return VU_ERR_UNSUPPORTED;
    e = record_cuda(cudaMemsetAsync(mask_counts2, 0, sizeof(uint64_t) * 2 * n, S(stream)));
    if (e) return e;
  }
  if (wide) cf_lowres2_wide_kernel<<<frame_grid(n, (int64_t)th * (tw / 8)), FT, 0, S(stream)>>>(frames, masks, h, w, th, tw, lut3d, alpha_lo, st,
                                                                                       reinterpret_cast<unsigned long long*>(mask_counts2));
  else if (wide4) cf_lowres4_wide_kernel<<<frame_grid(n, (int64_t)th * (tw / 4)), FT, 0, S(stream)>>>(frames, masks, h, w, th, tw, lut3d, alpha_lo, st,
                                                                                             reinterpret_cast<unsigned long long*>(mask_counts2));
  else if (s == 2) cf_lowres_kernel<2><<<frame_grid(n, (int64_t)th * (tw / 2)), FT, 0, S(stream)>>>(frames, masks, h, w, th, tw, lut3d, alpha_lo, st);
  else cf_lowres_kernel<4><<<frame_grid(n, (int64_t)th * tw), FT, 0, S(stream)>>>(frames, masks, h, w, th, tw, lut3d, alpha_lo, st);
  VU_RETURN_LAUNCH();
}

// cv2.resize (bilinear) of single-channel maps.  mode 1 fuses trimap/agent.py:60
// (snap) and :100 (128 where fuzzy, only for frames with flags[i] == 0).
// alt_src/alt_flags (nullable): frames with alt_flags[i] != 0 are copied from
// alt_src (full resolution) instead (the early-outs of colorfiltering/agent.py:303-307).
extern "C" int vu_resize_up_u8(const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw, int mode, const uint8_t* fuzzy,
                               const uint8_t* flags, const uint8_t* alt_src, const uint8_t* alt_flags, vu_stream_t stream) {
  VU_REQUIRE(src && dst && n >= 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && (mode == 0 || mode == 1));
  VU_REQUIRE((alt_src == nullptr) == (alt_flags == nullptr));
  if (sw == 2 * dw && sh == 2 * dh) return VU_ERR_UNSUPPORTED;  // cv2 switches to INTER_AREA there
  if (n == 0) return VU_OK;
  {
    // exact 2x / 4x, plain mode, 8-byte aligned rows: the constant-weight kernel
    const int sc = (dw == 2 * sw && dh == 2 * sh) ? 2 : ((dw == 4 * sw && dh == 4 * sh) ? 4 : 0);
    const bool al = ((reinterpret_cast<uintptr_t>(dst) & 7) == 0) && (!alt_src || (reinterpret_cast<uintptr_t>(alt_src) & 7) == 0);
    if (mode == 0 && sc != 0 && dw % 8 == 0 && al && n <= 65535) {
      const int run = 16 * sc;   // destination rows per thread: 16 source rows
      dim3 g((dw / 8 + 31) / 32, ((dh + run - 1) / run + FT / 32 - 1) / (FT / 32), n);
      if (sc == 2) resize_up_int_kernel<2><<<g, FT, 0, S(stream)>>>(src, sh, sw, dst, alt_src, alt_flags, run);
      else resize_up_int_kernel<4><<<g, FT, 0, S(stream)>>>(src, sh, sw, dst, alt_src, alt_flags, run);
      VU_RETURN_LAUNCH();
    }
  }
  for (int n0 = 0; n0 < n; n0 += 65535) {   // grid.z holds at most 65535 frames
    const int nn = n - n0 < 65535 ? n - n0 : 65535;
    dim3 grid((dw + RU_TW - 1) / RU_TW, (dh + RU_TH - 1) / RU_TH, nn);
    const uint8_t* s0 = src + (int64_t)n0 * sh * sw;
    uint8_t* d0 = dst + (int64_t)n0 * dh * dw;
    const uint8_t* fz = fuzzy ? fuzzy + (int64_t)n0 * dh * dw : nullptr;
    const uint8_t* as = alt_src ? alt_src + (int64_t)n0 * dh * dw : nullptr;
    if (mode == 0) resize_up_kernel<0><<<grid, FT, 0, S(stream)>>>(s0, sh, sw, d0, dh, dw, fz, flags ? flags + n0 : nullptr, as, alt_flags ? alt_flags + n0 : nullptr);
    else resize_up_kernel<1><<<grid, FT, 0, S(stream)>>>(s0, sh, sw, d0, dh, dw, fz, flags ? flags + n0 : nullptr, as, alt_flags ? alt_flags + n0 : nullptr);
  }
  VU_RETURN_LAUNCH();
}

extern "C" int vu_count_gt_lt_u8(const uint8_t* src, int n, int64_t per_item, int thr, uint64_t* counts2, vu_stream_t stream) {
  VU_REQUIRE(src && counts2 && n >= 0 && per_item >= 0);
  if (n == 0) return VU_OK;
  int e = record_cuda(cudaMemsetAsync(counts2, 0, sizeof(uint64_t) * 2 * n, S(stream)));
  if (e) return e;
  if (per_item == 0) return VU_OK;
  count_gt_lt_kernel<<<frame_grid(n, per_item / 16 + 1), FT, 0, S(stream)>>>(src, per_item, thr, reinterpret_cast<unsigned long long*>(counts2));
  VU_RETURN_LAUNCH();
}

extern "C" int vu_fuzzy_count(const uint8_t* frames, const uint8_t* alpha, int n, int64_t npix, const int32_t lo[3], const int32_t hi[3],
                              uint8_t* fuzzy01, uint64_t* counts2, vu_stream_t stream) {
  VU_REQUIRE(frames && alpha && lo && hi && fuzzy01 && counts2 && n >= 0 && npix >= 0);
  if (npix % 4) return VU_ERR_UNSUPPORTED;
  const void* ptrs[] = {frames, alpha, fuzzy01};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 3) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  int e = record_cuda(cudaMemsetAsync(counts2, 0, sizeof(uint64_t) * 2 * n, S(stream)));
  if (e) return e;
  if (npix == 0) return VU_OK;
  const bool wide = npix % 16 == 0 && ((reinterpret_cast<uintptr_t>(frames) & 15) == 0) && ((reinterpret_cast<uintptr_t>(alpha) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(fuzzy01) & 15) == 0);
  if (wide)
    fuzzy_count16_kernel<<<frame_grid(n, npix / 16), FT, 0, S(stream)>>>(reinterpret_cast<const uint4*>(frames), reinterpret_cast<const uint4*>(alpha),
                                                                         npix / 16, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2],
                                                                         reinterpret_cast<uint4*>(fuzzy01), reinterpret_cast<unsigned long long*>(counts2));
  else
    fuzzy_count_kernel<<<frame_grid(n, npix / 4), FT, 0, S(stream)>>>(frames, alpha, npix / 4, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], fuzzy01,
                                                                      reinterpret_cast<unsigned long long*>(counts2));
  VU_RETURN_LAUNCH();
}

extern "C" int vu_trimap_src_lo(const uint8_t* mask, const uint8_t* fuzzy, const uint8_t* flags, int n, int h, int w, int th, int tw, uint8_t* out,
                                vu_stream_t stream) {
  VU_REQUIRE(mask && out && n >= 0 && h > 0 && w > 0 && th > 0 && tw > 0);
  VU_REQUIRE((fuzzy == nullptr) == (flags == nullptr));
  if (n == 0) return VU_OK;
  const int sc = (w == 2 * tw && h == 2 * th) ? 2 : ((w == 4 * tw && h == 4 * th) ? 4 : 0);
  const bool al = ((reinterpret_cast<uintptr_t>(mask) & 15) == 0) && (!fuzzy || (reinterpret_cast<uintptr_t>(fuzzy) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
  if (sc != 0 && tw % 4 == 0 && al) {   // w is then a multiple of 16 (or 8): the wide loads are aligned
    if (sc == 2) trimap_src_lo_int_kernel<2><<<frame_grid(n, (int64_t)th * (tw / 4)), FT, 0, S(stream)>>>(mask, fuzzy, flags, h, w, th, tw, out);
    else trimap_src_lo_int_kernel<4><<<frame_grid(n, (int64_t)th * (tw / 4)), FT, 0, S(stream)>>>(mask, fuzzy, flags, h, w, th, tw, out);
  } else {
    trimap_src_lo_kernel<<<frame_grid(n, (int64_t)th * tw), FT, 0, S(stream)>>>(mask, fuzzy, flags, h, w, th, tw, out);
  }
  VU_RETURN_LAUNCH();
}
