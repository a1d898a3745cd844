// Background-difference gate of the bg / bg_step loops, fused
// (tools/unscreen/bg.py:85-92 == bg_offline.py:154-160):
//
//     raw  = u8(|f32(frame) - f32(bg)|)          exact integers: == absdiff
//     g    = BGR2GRAY(raw);  g[g > thr] = 255    values <= thr are KEPT
//     g    = dilate_mask(g, 4, 2)                MORPH_ELLIPSE(4,4), 2 iterations
//     out  = mask * (g // 255)
//
// Only g == 255 survives the integer division, and a dilation (a max) is 255
// exactly where one of its taps is: the gate is a BINARY dilation of
// B = (gray > thr || gray == 255).  The kernel therefore never materialises g:
// a CTA computes B for an output tile plus its halo straight from the frame and
// the background, packs it to one bit per pixel in shared memory (REDUX.OR of
// the lanes' nibbles), runs both dilations there on 32-pixel words (funnel
// shifts and ORs; cells outside the image are reset to 0 between the iterations,
// which is what cv2's "taps outside the image are ignored" means under
// iteration), and gates the mask on the way out.  Per frame it moves
// 3P (frame) + P (mask) + P (out) bytes; the background is shared by the
// frames of a clip and stays in L2.
//
// MORPH_ELLIPSE(4,4) = rows 0010 / 1111 / 1111 / 1111, anchor (2,2):
// dst(y,x) = src(y-2,x) | OR_{dy in -1..1, dx in -2..1} src(y+dy, x+dx)  (SURVEY.md A.1).
#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int GT_THREADS = 256;
constexpr int GT_TW = 120;           // output tile width: 128 staged pixels minus the 4 + 2 (+2 spare) halo columns
constexpr int GT_TH = 64;            // output tile height
constexpr int GT_ROWS = GT_TH + 6;   // staged rows: 4 above, 2 below
constexpr int GT_WORDS = 4;          // 128 staged pixels per row

// pixel x+dx of a bit-packed row (bit i of word j = pixel 32 j + i): dx in -2..1
__device__ __forceinline__ unsigned hor_or(const unsigned* row, int j) {
  const unsigned c = row[j];
  const unsigned p = j > 0 ? row[j - 1] : 0u, n = j < GT_WORDS - 1 ? row[j + 1] : 0u;
  const unsigned l1 = __funnelshift_l(p, c, 1), l2 = __funnelshift_l(p, c, 2);   // pixels x-1, x-2
  const unsigned r1 = __funnelshift_r(c, n, 1);                                   // pixel x+1
  return c | l1 | l2 | r1;
}

__global__ void __launch_bounds__(GT_THREADS) bgdiff_gate_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ bg,
                                                                 const uint8_t* __restrict__ masks, int h, int w, int64_t bg_frame_stride,
                                                                 int thr, uint8_t* __restrict__ out) {
  __shared__ unsigned B[GT_ROWS][GT_WORDS];
  __shared__ unsigned H[GT_ROWS][GT_WORDS];
  __shared__ unsigned D[GT_ROWS][GT_WORDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t fpix = (int64_t)blockIdx.z * h * w;
  const uint8_t* fr = frames + fpix * 3;
  const uint8_t* bgp = bg + (int64_t)blockIdx.z * bg_frame_stride;
  const int X0 = blockIdx.x * GT_TW, Y0 = blockIdx.y * GT_TH;
  const int px = X0 - 4 + 4 * lane;                 // first of this lane's 4 staged pixels
  const int teff = thr < 254 ? thr : 254;           // gray > thr || gray == 255
  // ---- B = (gray(|frame - bg|) > thr), one bit per pixel, 0 outside the image ----
  // 4 pixels = 3 words.  |frame - bg| of 4 bytes is one VABSDIFF4; BGR2GRAY's 15-bit coefficients are split into bytes
  // (3735 = 14*256 + 151, 19235 = 75*256 + 35, 9798 = 38*256 + 70), so that the weighted sum of a pixel is two IDP.4A over
  // the word that holds its three bytes (a PRMT gathers the two pixels that straddle words): 6 instructions per pixel
  // instead of ~20.  A warp takes its rows three at a time so that 18 loads per lane are in flight.
  const bool lane_full = px >= 0 && px + 3 < w;
  const bool lane_any = px + 3 >= 0 && px < w;
  constexpr unsigned WH = 14u | (75u << 8) | (38u << 16), WL = 151u | (35u << 8) | (70u << 16);
  const unsigned limit = ((unsigned)teff + 1u) << 15;   // gray > teff  <=>  weighted sum + 16384 >= (teff + 1) << 15
  constexpr int NW = GT_THREADS / 32, RB = 3;
  for (int r0 = warp; r0 < GT_ROWS; r0 += NW * RB) {
    unsigned fw[RB][3], bw[RB][3];
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      const int r = r0 + j * NW, gy = Y0 - 4 + r;
#pragma unroll
      for (int k = 0; k < 3; ++k) fw[j][k] = bw[j][k] = 0u;
      if (r < GT_ROWS && (unsigned)gy < (unsigned)h && lane_full) {
        const int64_t o = ((int64_t)gy * w + px) * 3;
        const unsigned* f4 = reinterpret_cast<const unsigned*>(fr + o);
        const unsigned* b4 = reinterpret_cast<const unsigned*>(bgp + o);
#pragma unroll
        for (int k = 0; k < 3; ++k) { fw[j][k] = __ldg(f4 + k); bw[j][k] = __ldg(b4 + k); }
      }
    }
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      const int r = r0 + j * NW, gy = Y0 - 4 + r;
      if (r >= GT_ROWS) break;   // warp-uniform
      unsigned nib = 0;
      if ((unsigned)gy < (unsigned)h) {
        if (lane_full) {
          const unsigned d0 = __vabsdiffu4(fw[j][0], bw[j][0]), d1 = __vabsdiffu4(fw[j][1], bw[j][1]), d2 = __vabsdiffu4(fw[j][2], bw[j][2]);
          const unsigned ps[4] = {d0,                             // B G R x
                                  __byte_perm(d0, d1, 0x0543),    // d0.3 d1.0 d1.1 x
                                  __byte_perm(d1, d2, 0x0432),    // d1.2 d1.3 d2.0 x
                                  d2 >> 8};                       // d2.1 d2.2 d2.3 0
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const unsigned num = __dp4a(ps[k], WH, 0u) * 256u + __dp4a(ps[k], WL, 16384u);
            nib |= (unsigned)(num >= limit) << k;
          }
        } else if (lane_any) {   // the group straddles the image border: pixel by pixel
          const int64_t o = ((int64_t)gy * w + px) * 3;
          for (int k = 0; k < 4; ++k) {
            const int x = px + k;
            if (x < 0 || x >= w) continue;
            const uint8_t* f = fr + o + 3 * k;
            const uint8_t* b = bgp + o + 3 * k;
            const int y = bgr2gray_px(abs((int)__ldg(f) - (int)__ldg(b)), abs((int)__ldg(f + 1) - (int)__ldg(b + 1)),
                                      abs((int)__ldg(f + 2) - (int)__ldg(b + 2)));
            nib |= (unsigned)(y > teff) << k;
          }
        }
      }
      const unsigned word = __reduce_or_sync(0xFFu << (lane & 24), nib << (4 * (lane & 7)));
      if ((lane & 7) == 0) B[r][lane >> 3] = word;
    }
  }
  __syncthreads();
  // in-image mask of a staged word (bit i = pixel X0 - 4 + 32 j + i)
  auto inside = [&](int r, int j) -> unsigned {
    const int gy = Y0 - 4 + r;
    if ((unsigned)gy >= (unsigned)h) return 0u;
    const int x0 = X0 - 4 + 32 * j;
    const int lo = max(0, -x0), hi = min(32, w - x0);   // valid bits [lo, hi)
    if (hi <= lo) return 0u;
    const unsigned m_hi = hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
    return m_hi & ~((1u << lo) - 1u);
  };
  for (int it = 0; it < 2; ++it) {
    for (int i = threadIdx.x; i < GT_ROWS * GT_WORDS; i += GT_THREADS) H[i >> 2][i & 3] = hor_or(B[i >> 2], i & 3);
    __syncthreads();
    for (int i = threadIdx.x; i < GT_ROWS * GT_WORDS; i += GT_THREADS) {
      const int r = i >> 2, j = i & 3;
      unsigned v = H[r][j];
      if (r >= 1) v |= H[r - 1][j];
      if (r + 1 < GT_ROWS) v |= H[r + 1][j];
      if (r >= 2) v |= B[r - 2][j];
      D[r][j] = v & inside(r, j);
    }
    __syncthreads();
    if (it == 0) {
      for (int i = threadIdx.x; i < GT_ROWS * GT_WORDS; i += GT_THREADS) B[i >> 2][i & 3] = D[i >> 2][i & 3];
      __syncthreads();
    }
  }
  // ---- out = mask where the dilated bit is set: 8 pixels (64-bit loads / stores) per thread when the rows allow ----
  const uint8_t* mk = masks + fpix;
  uint8_t* op = out + fpix;
  auto keep4 = [](unsigned bits) -> unsigned {
    return (((bits & 1u) | ((bits & 2u) << 7) | ((bits & 4u) << 14) | ((bits & 8u) << 21)) * 255u);
  };
  if ((w & 7) == 0) {
    for (int i = threadIdx.x; i < GT_TH * (GT_TW / 8); i += GT_THREADS) {
      const int ty = i / (GT_TW / 8), tg = i - ty * (GT_TW / 8);
      const int gy = Y0 + ty, gx = X0 + 8 * tg;
      if (gy >= h || gx >= w) continue;
      const int sx = 4 + 8 * tg;   // staged pixel index of gx: the 8 bits may straddle two words
      const unsigned lo = D[ty + 4][sx >> 5], hi = (sx >> 5) + 1 < GT_WORDS ? D[ty + 4][(sx >> 5) + 1] : 0u;
      const unsigned bits = __funnelshift_r(lo, hi, sx & 31) & 255u;
      const int64_t o = (int64_t)gy * w + gx;
      uint2 m = __ldg(reinterpret_cast<const uint2*>(mk + o));
      m.x &= keep4(bits & 15u);
      m.y &= keep4(bits >> 4);
      *reinterpret_cast<uint2*>(op + o) = m;
    }
  } else {
    for (int i = threadIdx.x; i < GT_TH * (GT_TW / 4); i += GT_THREADS) {
      const int ty = i / (GT_TW / 4), tg = i - ty * (GT_TW / 4);
      const int gy = Y0 + ty, gx = X0 + 4 * tg;
      if (gy >= h || gx >= w) continue;
      const int sx = 4 + 4 * tg;   // staged pixel index of gx
      const unsigned bits = (D[ty + 4][sx >> 5] >> (sx & 31)) & 15u;
      const int64_t o = (int64_t)gy * w + gx;
      *reinterpret_cast<unsigned*>(op + o) = __ldg(reinterpret_cast<const unsigned*>(mk + o)) & keep4(bits);
    }
  }
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_bgdiff_gate(const uint8_t* frames, const uint8_t* bg, const uint8_t* masks, int n, int h, int w, int bg_frames, int thr,
                              uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(frames && bg && masks && out && n >= 0 && h > 0 && w > 0);
  VU_REQUIRE(bg_frames == 1 || bg_frames == n);
  if (w % 4 != 0) return VU_ERR_UNSUPPORTED;
  const void* ptrs[] = {frames, bg, masks, out};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & ((w & 7) == 0 ? 7 : 3)) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  if (n > 65535) return VU_ERR_UNSUPPORTED;
  dim3 grid((w + GT_TW - 1) / GT_TW, (h + GT_TH - 1) / GT_TH, n);
  bgdiff_gate_kernel<<<grid, GT_THREADS, 0, S(stream)>>>(frames, bg, masks, h, w, bg_frames == 1 ? 0 : (int64_t)h * w * 3, thr, out);
  VU_RETURN_LAUNCH();
}
