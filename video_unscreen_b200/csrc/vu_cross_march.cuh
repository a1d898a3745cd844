// Chains of 3x3-cross dilations / erosions (cv2 MORPH_ELLIPSE(3,3) iterated:
// unscreen/utils/maskprocess.py:7-34 as called from colorfiltering/agent.py:281-282,
// trimap/agent.py:55-56 and tools/unscreen/bg.py:77), register-resident.
//
// A warp owns a vertical strip of 128 pixels (lane l: 4 pixels as two
// VIMNMX.U16x2 planes E = pixels 0,2 and O = pixels 1,3) and marches down the
// rows.  The NP passes of the chain are pipelined: pass k keeps the last two
// rows of ITS input in registers, and when the next one arrives it emits one
// row of output (centre row: min/max of up, down, self, left, right; the
// horizontal neighbours come from the adjacent lanes by shuffle), which is the
// row that arrives at pass k+1 in the same step.  One row of the image enters
// and one leaves per step; nothing touches shared memory and nothing
// synchronises.  The strip loses one pixel of validity per pass on either
// side, so only the inner 128 - 2*4*ceil(NP/4) pixels are written; rows
// likewise (NP rows of lead-in above and below the band of rows a warp owns).
//
// "Taps outside the image are ignored" (SURVEY.md A.1): cells outside the
// image hold the identity of the pass that reads them (0 for a dilation, 255
// for an erosion) - on load, and again after every pass.
//
// MODE 0: dst = chain(src), pass k erodes if bit k of EMASK is set, with the
//         colour filter's adaptive threshold applied on load when stats2 is
//         given (colorfiltering/agent.py:277-280).
// MODE 1: dst = trimap classification (trimap/agent.py:54-58) of dilate^NP(src)
//         and erode^NP(src), both chains marching together.
#pragma once
#include "vu_common.cuh"

namespace vu {
namespace march {

constexpr int WARPS = 4;

template <bool ERODE>
__device__ __forceinline__ unsigned mm(unsigned a, unsigned b) { return ERODE ? __vminu2(a, b) : __vmaxu2(a, b); }

// centre row c with its vertical neighbours u, d -> cross min/max
template <bool ERODE>
__device__ __forceinline__ uint2 cross_row(uint2 u, uint2 c, uint2 d) {
  const unsigned prevO = __shfl_up_sync(0xffffffffu, c.y, 1), nextE = __shfl_down_sync(0xffffffffu, c.x, 1);
  // left neighbours of pixels (0,2) are (prev.3, 1); right neighbours of pixels (1,3) are (2, next.0)
  const unsigned le = __byte_perm(prevO, c.y, 0x5432), ro = __byte_perm(c.x, nextE, 0x5432);
  uint2 r;
  r.x = mm<ERODE>(mm<ERODE>(c.x, u.x), mm<ERODE>(d.x, mm<ERODE>(le, c.y)));
  r.y = mm<ERODE>(mm<ERODE>(c.y, u.y), mm<ERODE>(d.y, mm<ERODE>(c.x, ro)));
  return r;
}

template <int NP, unsigned EMASK, int MODE>
struct Marcher {
  static constexpr int CH = MODE == 1 ? 2 : 1;
  uint2 hist[CH][NP][2];
  unsigned me, mo;   // planes' masks of this lane's pixels that are inside the image

  __device__ __forceinline__ static constexpr bool erodes(int ch, int k) { return MODE == 1 ? ch == 1 : ((EMASK >> k) & 1u) != 0; }
  __device__ __forceinline__ static constexpr unsigned ident(int ch, int k) { return erodes(ch, k) ? 0x00FF00FFu : 0u; }

  // out-of-image cells of a row that pass k of chain ch is about to read
  __device__ __forceinline__ uint2 reset(uint2 v, bool row_in, int ch, int k) const {
    const unsigned id = ident(ch, k);
    if (!row_in) return make_uint2(id, id);
    return make_uint2((v.x & me) | (id & ~me), (v.y & mo) | (id & ~mo));
  }

  // one step: `in` = the new input row (already reset for pass 0) of every chain, y_in its row index;
  // returns the row y_in - NP of the last pass
  template <int PH>
  __device__ __forceinline__ void step(uint2 (&cur)[CH], int y_in, int h) {
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const bool row_in = (unsigned)(y_in - 1 - k) < (unsigned)h;   // the row this pass emits
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const uint2 up = hist[ch][k][PH], ce = hist[ch][k][PH ^ 1], dn = cur[ch];
        hist[ch][k][PH] = dn;
        uint2 o = erodes(ch, k) ? cross_row<true>(up, ce, dn) : cross_row<false>(up, ce, dn);
        if (k + 1 < NP) o = reset(o, row_in, ch, k + 1);
        cur[ch] = o;
      }
    }
  }
};

template <int NP, unsigned EMASK, int MODE>
__global__ void __launch_bounds__(WARPS * 32) cross_march_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w, int strips,
                                                                 int bands, int band_rows, const unsigned long long* __restrict__ stats2,
                                                                 double thr_ratio) {
  constexpr int HG = (NP + 3) / 4;        // halo, in 4-pixel groups
  constexpr int OG = 32 - 2 * HG;         // groups of the strip that are written
  using M = Marcher<NP, EMASK, MODE>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int unit = blockIdx.x * WARPS + warp;
  if (unit >= strips * bands) return;
  const int strip = unit % strips, band = unit / strips;
  const int64_t frame = (int64_t)blockIdx.y * h * w;
  const int px = (strip * OG - HG + lane) * 4;
  const int Y0 = band * band_rows, Y1 = min(h, Y0 + band_rows);
  M m;
  {
    const bool v0 = (unsigned)px < (unsigned)w, v1 = (unsigned)(px + 1) < (unsigned)w, v2 = (unsigned)(px + 2) < (unsigned)w,
               v3 = (unsigned)(px + 3) < (unsigned)w;
    m.me = (v0 ? 0x0000FFFFu : 0u) | (v2 ? 0xFFFF0000u : 0u);
    m.mo = (v1 ? 0x0000FFFFu : 0u) | (v3 ? 0xFFFF0000u : 0u);
  }
#pragma unroll
  for (int ch = 0; ch < M::CH; ++ch)
#pragma unroll
    for (int k = 0; k < NP; ++k) m.hist[ch][k][0] = m.hist[ch][k][1] = make_uint2(0u, 0u);
  // colour filter threshold (agent.py:277-280): alpha < thr -> 0, thr = thr_ratio * mean; v < thr <=> v < ceil(thr)
  int ithr = 0;
  if (MODE == 0 && stats2) {
    const unsigned long long sum = stats2[2 * blockIdx.y], cnt = stats2[2 * blockIdx.y + 1];
    if (cnt != 0) ithr = (int)ceil(__dmul_rn(__ddiv_rn((double)sum, (double)cnt), thr_ratio));
  }
  const bool vec = (w & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0);
  const bool all_in = px >= 0 && px + 3 < w;
  const bool writer = lane >= HG && lane < 32 - HG && px < w;
  const int T = (Y1 - Y0) + 2 * NP;
  // the row of step t: loaded PF steps ahead (a step is a long dependent chain of shuffles and min/max: without the
  // look-ahead every step would start with a full L2 round trip)
  constexpr int PF = 4;
  auto load_row = [&](int t) -> unsigned {
    const int y_in = Y0 - NP + t;
    unsigned word = 0;
    if (t < T && (unsigned)y_in < (unsigned)h && px + 3 >= 0 && px < w) {
      const uint8_t* p = src + frame + (int64_t)y_in * w + px;
      if (all_in && vec) {
        word = __ldg(reinterpret_cast<const unsigned*>(p));
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((unsigned)(px + k) < (unsigned)w) word |= (unsigned)__ldg(p + k) << (8 * k);
      }
    }
    return word;
  };
  // ---- flat units: a unit whose input (after the threshold) is one value c leaves c behind, whatever the chain (a
  //      max / min of equal values; cells outside the image are ignored), and a trimap classification of a constant
  //      is 0 or 255.  Most of a matte is exactly that - zero outside the person - so a min / max sweep over the unit's
  //      rows (7 instructions per row against ~20 per pass and row of the chain) pays for itself many times over. ----
  if (vec) {
    unsigned mn = 0x00FF00FFu, mx = 0u;
    for (int t = 0; t < T; ++t) {
      const int y_in = Y0 - NP + t;
      if ((unsigned)y_in < (unsigned)h && all_in) {
        const unsigned word = __ldg(reinterpret_cast<const unsigned*>(src + frame + (int64_t)y_in * w + px));
        const unsigned e = word & 0x00FF00FFu, o = (word >> 8) & 0x00FF00FFu;
        mn = __vminu2(mn, __vminu2(e, o));
        mx = __vmaxu2(mx, __vmaxu2(e, o));
      }
    }
    int lo = (int)min(mn & 0xFFFFu, mn >> 16), hi = (int)max(mx & 0xFFFFu, mx >> 16);
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (hi < ithr || lo >= hi) {   // (all below the threshold -> all zero) or (one value; lo > hi: no pixel of the unit is inside the image)
      int c = hi < ithr ? 0 : hi;
      if (MODE == 1) c = c >= 128 ? 255 : 0;
      if (writer) {
        const unsigned out = (unsigned)c * 0x01010101u;
        for (int y = Y0; y < Y1; ++y) *reinterpret_cast<unsigned*>(dst + frame + (int64_t)y * w + px) = out;
      }
      return;
    }
  }
  unsigned q[PF];
#pragma unroll
  for (int i = 0; i < PF; ++i) q[i] = load_row(i);
  auto one = [&](auto ph, int t, unsigned word) {
    constexpr int PH = decltype(ph)::value;
    const int y_in = Y0 - NP + t;
    const bool row_in = (unsigned)y_in < (unsigned)h;
    if (MODE == 0 && ithr > 0) {
      unsigned r = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned v = (word >> (8 * k)) & 255u;
        r |= ((int)v < ithr ? 0u : v) << (8 * k);
      }
      word = r;
    }
    const uint2 sp = make_uint2(word & 0x00FF00FFu, (word >> 8) & 0x00FF00FFu);
    uint2 cur[M::CH];
#pragma unroll
    for (int ch = 0; ch < M::CH; ++ch) cur[ch] = m.reset(sp, row_in, ch, 0);
    m.template step<PH>(cur, y_in, h);
    const int y_out = y_in - NP;
    if (y_out >= Y0 && y_out < Y1 && writer) {
      unsigned out;
      if (MODE == 0) {
        out = cur[0].x | (cur[0].y << 8);
      } else {
        // dilated < 128 -> 0, else eroded > 127 -> 255, else 128  (per 16-bit lane, values 0..255)
        uint2 r;
        {
          const unsigned md = cur[0].x & 0x00800080u, mer = cur[1].x & md;
          r.x = md | (mer - (mer >> 7));
        }
        {
          const unsigned md = cur[0].y & 0x00800080u, mer = cur[1].y & md;
          r.y = md | (mer - (mer >> 7));
        }
        out = r.x | (r.y << 8);
      }
      uint8_t* o = dst + frame + (int64_t)y_out * w + px;
      if (px + 3 < w && vec) {
        *reinterpret_cast<unsigned*>(o) = out;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (px + k < w) o[k] = (uint8_t)(out >> (8 * k));
      }
    }
  };
  static_assert(PF == 4, "the loop below is unrolled over the look-ahead queue");
  for (int t = 0; t < T; t += 4) {
    const unsigned w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
#pragma unroll
    for (int i = 0; i < PF; ++i) q[i] = load_row(t + PF + i);
    one(std::integral_constant<int, 0>{}, t, w0);
    if (t + 1 < T) one(std::integral_constant<int, 1>{}, t + 1, w1);
    if (t + 2 < T) one(std::integral_constant<int, 0>{}, t + 2, w2);
    if (t + 3 < T) one(std::integral_constant<int, 1>{}, t + 3, w3);
  }
}

// rows a warp owns: enough warps to fill the machine, little enough lead-in overhead
inline int pick_band_rows(int h, int np) {
  // short bands: more warps in flight (a 540 x 960 matte in 64-row bands left the SMs at 16 % occupancy with a ragged
  // last wave) and more of them flat (see the shortcut above), for 2 * np rows of lead-in per band
  int rows = 32;
  if (rows < 4 * np) rows = 4 * np;
  if (rows > h) rows = h;
  return rows;
}

template <int NP, unsigned EMASK, int MODE>
int launch(const uint8_t* src, uint8_t* dst, int n, int h, int w, const unsigned long long* stats2, double thr_ratio, cudaStream_t stream) {
  constexpr int HG = (NP + 3) / 4, OG = 32 - 2 * HG;
  const int strips = (w + 4 * OG - 1) / (4 * OG);
  const int band_rows = pick_band_rows(h, NP);
  const int bands = (h + band_rows - 1) / band_rows;
  dim3 grid((strips * bands + WARPS - 1) / WARPS, n);
  cross_march_kernel<NP, EMASK, MODE><<<grid, WARPS * 32, 0, stream>>>(src, dst, h, w, strips, bands, band_rows, stats2, thr_ratio);
  note_launch();
  return record_cuda(cudaGetLastError());
}

}  // namespace march
}  // namespace vu
