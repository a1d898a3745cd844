// The trimap tail in bit logic, for frames that are an exact 2x or 4x of the working resolution
// (1080p and 4K with input_long_side = 960) and MORPH_ELLIPSE(3,3) (unscreen/trimap/agent.py:35-61, 97-100):
//
//     m    = resize(mask, NEAREST)             m[fuzzy] = 0 first in the ensemble branch (:97)
//     dil  = dilate_mask(m, 3, r);  ero = erode_mask(m, 3, r)
//     t    = 128;  t[ero > 127] = 255;  t[dil < 128] = 0
//     out  = resize(t, (W,H))  (bilinear, :59);  out[(out > 0) & (out < 255)] = 128;  out[fuzzy] = 128 (:100)
//
// Everything after the nearest down-scale depends on m only through B = (m >= 128): a max is < 128 iff all its taps
// are, a min is > 127 iff all its taps are.  So t == 0 where the binary dilation of B is empty, t == 255 where the
// binary erosion of B is full.  And for exact 2x / 4x scales the snapped bilinear up-scale of a {0,128,255} map is tap
// logic too: cv2's fixed-point interpolation gives exactly 0 iff every tap with a non-zero weight is 0 and exactly
// 255 iff every such tap is 255 (the smallest weight product, 1/64, still moves the result by 2; checked
// exhaustively over all weight pairs and tap values, see tests), and the taps with non-zero weights are the 2x2
// neighbourhood towards the pixel's quadrant, replicated at the borders.
//
// Kernel A: one CTA per 96 x 64 tile of the working resolution: B packed to bits straight from the full-resolution
// mask (and fuzzy map), r cross dilations and r cross erosions on 32-pixel words in shared memory (cells outside the
// image reset to the identity after every pass), then the four quadrant ANDs of Z = "t == 0" and F = "t == 255":
// eight bit planes of th x ceil(tw/32) words per frame.  Kernel B: one pass over the output: four plane words per 4
// pixels, the fuzzy override, one store.  The working-resolution trimap never exists as bytes.
#include <cstdlib>

#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int TB_THREADS = 256;
constexpr int TB_OW = 96, TB_OH = 64;     // output tile (working-resolution pixels)
constexpr int TB_HX = 16;                 // staged halo columns on either side (>= passes + 1)
constexpr int TB_MAXR = 12;               // most passes
constexpr int TB_WORDS = (TB_OW + 2 * TB_HX) / 32;   // 4
constexpr int TB_ROWS_MAX = TB_OH + 2 * (TB_MAXR + 1);

// bytes 0, S, 2S, 3S of the 4S bytes at p
template <int SC>
__device__ __forceinline__ unsigned pick4(const uint8_t* p) {
  if (SC == 2) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    return __byte_perm(v.x, v.y, 0x6420);
  } else {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    return (v.x & 255u) | ((v.y & 255u) << 8) | ((v.z & 255u) << 16) | (v.w << 24);
  }
}

// bits 0, 2, 4 .. 30 of x -> bits 0 .. 15;  bits 0, 4, 8 .. 28 -> bits 0 .. 7
__device__ __forceinline__ unsigned compress2(unsigned x) {
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  return (x | (x >> 8)) & 0xFFFFu;
}
__device__ __forceinline__ unsigned compress4(unsigned x) {
  x &= 0x11111111u;
  x = (x | (x >> 3)) & 0x03030303u;
  x = (x | (x >> 6)) & 0x000F000Fu;
  return (x | (x >> 12)) & 0xFFu;
}

// plane p of frame n: planes[((p * nframes + n) * th + y) * wpr + j]
// PACKED: the source arrives as bit planes written by vu_cf_alpha_up_fuzzy: mask = B bits at the working resolution
// [n][th][tw/8] (nearest-sampled alpha >= 128), fuzzy = one bit per FULL-resolution pixel [n][h][w/8]; tw % 16 == 0.
template <int SC, bool PACKED>
__global__ void __launch_bounds__(TB_THREADS) trimap_bits_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ fuzzy,
                                                                 const uint8_t* __restrict__ flags, int h, int w, int th, int tw, int passes,
                                                                 unsigned* __restrict__ planes, int nframes, int wpr, int oh) {
  __shared__ unsigned Db[2][TB_ROWS_MAX][TB_WORDS], Eb[2][TB_ROWS_MAX][TB_WORDS], In[TB_ROWS_MAX][TB_WORDS];
  __shared__ unsigned ZL[TB_ROWS_MAX][TB_WORDS], ZR[TB_ROWS_MAX][TB_WORDS], FL[TB_ROWS_MAX][TB_WORDS], FR[TB_ROWS_MAX][TB_WORDS];
  const int n = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int halo = passes + 1;
  const int rows = oh + 2 * halo;   // oh <= TB_OH output rows per tile
  const int X0 = blockIdx.x * TB_OW - TB_HX, Y0 = blockIdx.y * oh - halo;   // staged origin
  const bool ens = fuzzy && flags && flags[n] == 0;
  const uint8_t* mk = mask + (int64_t)n * h * w;
  const uint8_t* fz = fuzzy ? fuzzy + (int64_t)n * h * w : nullptr;
  // ---- B = (nearest-sampled mask >= 128, fuzzy pixels cleared), one bit per working-resolution pixel ----
  if (PACKED) {
    const uint8_t* mb = mask + (int64_t)n * th * (tw >> 3);
    const uint8_t* fb = fuzzy ? fuzzy + (int64_t)n * h * (w >> 3) : nullptr;
    for (int i = threadIdx.x; i < rows * TB_WORDS * 2; i += TB_THREADS) {   // 16 working-resolution pixels per item
      const int r = i / (TB_WORDS * 2), hw = i % (TB_WORDS * 2);
      const int y = Y0 + r, x = X0 + 16 * hw;
      unsigned b = 0, in = 0;
      if ((unsigned)y < (unsigned)th && x >= 0 && x < tw) {   // tw % 16 == 0: inside or outside as a whole
        b = __ldg(reinterpret_cast<const unsigned short*>(mb + (int64_t)y * (tw >> 3) + (x >> 3)));
        if (ens) {   // fuzzy bits of the full-resolution pixels (SC*y, SC*x .. SC*(x+15)): SC*16 bits, every SC-th one
          const unsigned* f = reinterpret_cast<const unsigned*>(fb + (int64_t)SC * y * (w >> 3) + ((SC * x) >> 3));
          const unsigned fz = SC == 2 ? compress2(__ldg(f)) : (compress4(__ldg(f)) | (compress4(__ldg(f + 1)) << 8));
          b &= ~fz;
        }
        in = 0xFFFFu;
      }
      reinterpret_cast<unsigned short*>(&Db[0][r][0])[hw] = (unsigned short)b;
      reinterpret_cast<unsigned short*>(&Eb[0][r][0])[hw] = (unsigned short)(b | (~in & 0xFFFFu));
      reinterpret_cast<unsigned short*>(&In[r][0])[hw] = (unsigned short)in;
    }
  } else
  for (int r = warp; r < rows; r += TB_THREADS / 32) {
    const int y = Y0 + r, x = X0 + 4 * lane;
    unsigned nib = 0, inb = 0;
    if ((unsigned)y < (unsigned)th && x >= 0 && x < tw) {   // tw % 4 == 0: a group is inside or outside as a whole
      const int64_t si = (int64_t)SC * y * w + (int64_t)SC * x;
      unsigned v = pick4<SC>(mk + si);
      if (ens) {
        const unsigned f = pick4<SC>(fz + si);
        const unsigned nz = ((f & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | f;   // bit 7 of every non-zero fuzzy byte
        v &= ~(((nz >> 7) & 0x01010101u) * 255u);
      }
      nib = ((v >> 7) & 1u) | ((v >> 14) & 2u) | ((v >> 21) & 4u) | ((v >> 28) & 8u);   // byte >= 128
      inb = 15u;
    }
    const unsigned grp = 0xFFu << (lane & 24);
    const unsigned bw = __reduce_or_sync(grp, nib << (4 * (lane & 7)));
    const unsigned iw = __reduce_or_sync(grp, inb << (4 * (lane & 7)));
    if ((lane & 7) == 0) {
      Db[0][r][lane >> 3] = bw;            // outside the image: 0, the identity of a dilation
      Eb[0][r][lane >> 3] = bw | ~iw;      // outside the image: 1, the identity of an erosion
      In[r][lane >> 3] = iw;
    }
  }
  __syncthreads();
  // ---- passes of the 3x3 cross on both chains ----
  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    for (int i = threadIdx.x; i < rows * TB_WORDS; i += TB_THREADS) {
      const int r = i / TB_WORDS, j = i % TB_WORDS;
      const unsigned in = In[r][j];
      {
        const unsigned c = Db[cur][r][j];
        const unsigned lf = __funnelshift_l(j > 0 ? Db[cur][r][j - 1] : 0u, c, 1);                 // pixel x-1
        const unsigned rt = __funnelshift_r(c, j < TB_WORDS - 1 ? Db[cur][r][j + 1] : 0u, 1);     // pixel x+1
        const unsigned up = r > 0 ? Db[cur][r - 1][j] : 0u, dn = r + 1 < rows ? Db[cur][r + 1][j] : 0u;
        Db[cur ^ 1][r][j] = (c | lf | rt | up | dn) & in;
      }
      {
        const unsigned c = Eb[cur][r][j];
        const unsigned lf = __funnelshift_l(j > 0 ? Eb[cur][r][j - 1] : 0xFFFFFFFFu, c, 1);
        const unsigned rt = __funnelshift_r(c, j < TB_WORDS - 1 ? Eb[cur][r][j + 1] : 0xFFFFFFFFu, 1);
        const unsigned up = r > 0 ? Eb[cur][r - 1][j] : 0xFFFFFFFFu, dn = r + 1 < rows ? Eb[cur][r + 1][j] : 0xFFFFFFFFu;
        Eb[cur ^ 1][r][j] = (c & lf & rt & up & dn) | ~in;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  // ---- Z = (t == 0) = no dilated bit, F = (t == 255) = eroded bit; their ANDs with the left / right neighbour,
  //      replicated at the image border (the neighbour outside the image is the pixel itself) ----
  for (int i = threadIdx.x; i < rows * TB_WORDS; i += TB_THREADS) {
    const int r = i / TB_WORDS, j = i % TB_WORDS;
    const unsigned in = In[r][j];
    const unsigned inl = __funnelshift_l(j > 0 ? In[r][j - 1] : 0u, in, 1), inr = __funnelshift_r(in, j < TB_WORDS - 1 ? In[r][j + 1] : 0u, 1);
    auto lr = [&](unsigned c, unsigned pw, unsigned nw, unsigned& L, unsigned& R) {
      const unsigned lf = __funnelshift_l(pw, c, 1), rt = __funnelshift_r(c, nw, 1);
      L = c & (lf | ~inl);   // no pixel to the left: the pixel itself
      R = c & (rt | ~inr);
    };
    const unsigned z = ~Db[cur][r][j], zp = j > 0 ? ~Db[cur][r][j - 1] : 0u, zn = j < TB_WORDS - 1 ? ~Db[cur][r][j + 1] : 0u;
    const unsigned f = Eb[cur][r][j], fp = j > 0 ? Eb[cur][r][j - 1] : 0u, fn = j < TB_WORDS - 1 ? Eb[cur][r][j + 1] : 0u;
    unsigned L, R;
    lr(z, zp, zn, L, R);
    ZL[r][j] = L; ZR[r][j] = R;
    lr(f, fp, fn, L, R);
    FL[r][j] = L; FR[r][j] = R;
  }
  __syncthreads();
  // ---- the eight quadrant planes of the tile's rows: {Z,F} x {left,right} x {row above, row below}, 3 words per row ----
  for (int i = threadIdx.x; i < oh * 3 * 8; i += TB_THREADS) {
    const int pl = i / (oh * 3), rem = i - pl * (oh * 3);
    const int ty = rem / 3, jo = rem - 3 * ty;
    const int y = blockIdx.y * oh + ty;
    const int xw = blockIdx.x * 3 + jo;      // word of the plane row
    if (y >= th || xw >= wpr) continue;
    const int r = ty + halo;
    const int ro = (pl & 1) ? (y + 1 < th ? r + 1 : r) : (y > 0 ? r - 1 : r);   // the other row of the pair, replicated at the border
    const unsigned(*src)[TB_WORDS] = (pl >> 1) == 0 ? ZL : ((pl >> 1) == 1 ? ZR : ((pl >> 1) == 2 ? FL : FR));
    // output word jo = staged bits 16 + 32 jo .. 47 + 32 jo
    const unsigned a = __funnelshift_r(src[r][jo], src[r][jo + 1], 16), b = __funnelshift_r(src[ro][jo], src[ro][jo + 1], 16);
    planes[(((int64_t)pl * nframes + n) * th + y) * wpr + xw] = a & b;
  }
}

// The same planes by MARCHING (packed inputs, P passes, tw % 32 == 0 and at most 32 words per row: every working resolution
// of input_long_side <= 1024): one warp owns a band of rows over the whole width, lane = word column, so the horizontal
// neighbours are the neighbouring lanes' words (two shuffles per row and pass, no halo columns) and the P passes are a
// systolic pipeline down the rows held in registers: when input row t arrives, pass p + 1 of row t - p - 1 is complete.
// No shared memory, no barrier, and ~6x fewer instructions than the tile kernel above, which spends them on indices,
// bounds and barriers for one word per thread and pass.
constexpr int TM_WARPS = 4;      // bands per CTA
template <int SC, int P>
__global__ void __launch_bounds__(32 * TM_WARPS) trimap_bits_march_kernel(const uint8_t* __restrict__ mask_bits, const uint8_t* __restrict__ fuzzy_bits,
                                                                          const uint8_t* __restrict__ flags, int h, int w, int th, int tw,
                                                                          unsigned* __restrict__ planes, int nframes, int wpr, int band_rows) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int y0 = (blockIdx.x * TM_WARPS + (threadIdx.x >> 5)) * band_rows;
  if (y0 >= th) return;                      // warp-uniform; nothing below synchronises the CTA
  const int y1 = min(y0 + band_rows, th);
  const bool ens = fuzzy_bits && flags && flags[n] == 0;
  const bool col = lane < wpr;
  const unsigned inw = col ? FULL : 0u;      // tw % 32 == 0: whole words
  const unsigned* mb = reinterpret_cast<const unsigned*>(mask_bits + (int64_t)n * th * (tw >> 3));
  const uint8_t* fb = fuzzy_bits + (int64_t)n * h * (w >> 3);
  const int fstride = SC * (w >> 3);         // bytes between the fuzzy rows of consecutive working rows
  auto load_row = [&](int y) -> unsigned {   // B = (nearest-sampled alpha >= 128) & ~fuzzy of working row y; 0 outside the image
    if (!col || y < 0 || y >= th) return 0u;
    unsigned b = __ldg(mb + (int64_t)y * wpr + lane);
    if (ens) {
      const uint8_t* f = fb + (int64_t)y * fstride;
      unsigned fz;
      if (SC == 2) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(f) + lane);
        fz = compress2(v.x) | (compress2(v.y) << 16);
      } else {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(f) + lane);
        fz = compress4(v.x) | (compress4(v.y) << 8) | (compress4(v.z) << 16) | (compress4(v.w) << 24);
      }
      b &= ~fz;
    }
    return b;
  };
  // x | pixel x-1 | pixel x+1 (dilation; nothing beyond lanes 0 / 31) and x & x-1 & x+1 (erosion; ones beyond: lanes past the
  // last column hold the erosion's identity anyway)
  auto hor_or = [&](unsigned x) {
    unsigned l = __shfl_up_sync(FULL, x, 1), r = __shfl_down_sync(FULL, x, 1);
    if (lane == 0) l = 0u;
    if (lane == 31) r = 0u;
    return x | __funnelshift_l(l, x, 1) | __funnelshift_r(x, r, 1);
  };
  auto hor_and = [&](unsigned x) {
    unsigned l = __shfl_up_sync(FULL, x, 1), r = __shfl_down_sync(FULL, x, 1);
    if (lane == 0) l = FULL;
    if (lane == 31) r = FULL;
    return x & __funnelshift_l(l, x, 1) & __funnelshift_r(x, r, 1);
  };
  // in-image masks of the pixels left / right of every pixel of this lane's word
  const unsigned inl = __funnelshift_l(lane > 0 ? inw : 0u, inw, 1);                              // the previous lane is a column whenever this one is
  const unsigned inr = __funnelshift_r(inw, (lane + 1 < wpr) ? FULL : 0u, 1);
  unsigned aD[P], bD[P], hD[P], aE[P], bE[P], hE[P];     // rows tp-2, tp-1 of X_p and the horizontal combination of row tp-1
#pragma unroll
  for (int p = 0; p < P; ++p) { aD[p] = bD[p] = hD[p] = 0u; aE[p] = bE[p] = hE[p] = FULL; }
  unsigned above[4] = {0u, 0u, 0u, 0u}, cur[4] = {0u, 0u, 0u, 0u};   // ZL, ZR, FL, FR of rows y-1 and y
  const int t0 = y0 - (P + 1), t1 = y1 + P;              // input rows t0 .. t1 inclusive
  unsigned nextB = load_row(t0);
#pragma unroll 1
  for (int t = t0; t <= t1; ++t) {
    const unsigned B = nextB;
    nextB = load_row(t + 1);                             // one row ahead of the arithmetic
    const unsigned in_t = ((unsigned)t < (unsigned)th) ? inw : 0u;
    unsigned cD = B, cE = B | ~in_t;                     // X_0(t)
#pragma unroll
    for (int p = 0; p < P; ++p) {
      // X_(p+1)(tp - 1) from X_p(tp - 2), X_p(tp - 1), X_p(tp), tp = t - p; rows outside the image are reset to the identity
      const unsigned in_o = ((unsigned)(t - p - 1) < (unsigned)th) ? inw : 0u;
      const unsigned oD = (hD[p] | aD[p] | cD) & in_o;
      const unsigned oE = (hE[p] & aE[p] & cE) | ~in_o;
      aD[p] = bD[p]; bD[p] = cD; hD[p] = hor_or(cD);
      aE[p] = bE[p]; bE[p] = cE; hE[p] = hor_and(cE);
      cD = oD; cE = oE;
    }
    // cD / cE = X_P(yP), yP = t - P: Z = no dilated bit, F = eroded bit, ANDed with the pixel on the left / right
    // (replicated at the image border: the pixel itself)
    unsigned nw[4];
    {
      const unsigned z = ~cD, f = cE;
      unsigned zl = __shfl_up_sync(FULL, z, 1), zr = __shfl_down_sync(FULL, z, 1), fl = __shfl_up_sync(FULL, f, 1), fr = __shfl_down_sync(FULL, f, 1);
      if (lane == 0) { zl = 0u; fl = 0u; }
      if (lane == 31) { zr = 0u; fr = 0u; }
      nw[0] = z & (__funnelshift_l(zl, z, 1) | ~inl);
      nw[1] = z & (__funnelshift_r(z, zr, 1) | ~inr);
      nw[2] = f & (__funnelshift_l(fl, f, 1) | ~inl);
      nw[3] = f & (__funnelshift_r(f, fr, 1) | ~inr);
    }
    const int y = t - P - 1;                             // the row whose three X_P rows are now known
    if (y >= y0 && col) {                                // y < y1 by the loop bound
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        planes[(((int64_t)(2 * k) * nframes + n) * th + y) * wpr + lane] = cur[k] & (y > 0 ? above[k] : cur[k]);
        planes[(((int64_t)(2 * k + 1) * nframes + n) * th + y) * wpr + lane] = cur[k] & (y + 1 < th ? nw[k] : cur[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { above[k] = cur[k]; cur[k] = nw[k]; }
  }
}

// plane index: (Z left, Z right, F left, F right) x (pair with the row above, pair with the row below) -> 2 * k + below.
// NG groups of 4 output pixels per thread (4: 128-bit fuzzy loads and stores - a streaming kernel needs the bytes in
// flight; 1: widths that are not a multiple of 16).  The NG groups of a thread read the same four plane words.
// PACKED (NG == 4 only): fuzzy is the bit plane [n][h][w/8] of vu_cf_alpha_up_fuzzy
template <int SC, int NG, bool PACKED>
__global__ void __launch_bounds__(TB_THREADS) trimap_up_bits_kernel(const unsigned* __restrict__ planes, int nframes, int th, int tw, int wpr,
                                                                    const uint8_t* __restrict__ fuzzy, const uint8_t* __restrict__ flags,
                                                                    uint8_t* __restrict__ out) {
  const int n = blockIdx.y;
  const int h = SC * th, w = SC * tw;
  const bool ens = fuzzy && flags && flags[n] == 0;
  const int tpr = w / (4 * NG);   // threads per row
  const int64_t psz = (int64_t)nframes * th * wpr;   // words per plane
  constexpr int CPG = 4 / SC;                        // working-resolution columns per group: 1 (4x) or 2 (2x)
  // a thread owns 4 * NG columns of the SC output rows of ONE working-resolution row: the eight plane words (and the index
  // arithmetic - the kernel is bound by instruction issue) are fetched once for SC stores; a warp = 8 such column groups x 4
  // working rows: a compact footprint, because the flat test below diverges per warp and a compact warp straddles the
  // unknown band far less often than 32 groups of one row.  Warp tiles in a grid-stride loop, one 32-bit division each.
  const int lane = threadIdx.x & 31;
  const int tiles_x = (tpr + 7) >> 3, tiles = tiles_x * ((th + 3) >> 2);
  // The tile loop is a software pipeline: the plane words (and the packed fuzzy bits) of the warp's NEXT tile are requested
  // before this tile is worked on - a tile is ~60 instructions behind an L2 round trip, and with the loads inside the
  // iteration the warps spent most of it waiting (ncu: long-scoreboard stalls 7 per issued instruction, 62 % of the issue slots)
  const int tstep = gridDim.x * (TB_THREADS / 32);
  unsigned nxt[8], nfz[PACKED ? SC : 1];
  auto request = [&](int tile) {
    const int tyy = tile / tiles_x, txx = tile - tyy * tiles_x;
    const int r = tyy * 4 + (lane >> 3), t = txx * 8 + (lane & 7);
    if (tile >= tiles || r >= th || t >= tpr) return;
    const unsigned* row = planes + ((int64_t)n * th + r) * wpr + ((t * NG * CPG) >> 5);
#pragma unroll
    for (int q = 0; q < 8; ++q) nxt[q] = __ldg(row + q * psz);
    if (PACKED && ens) {
#pragma unroll
      for (int yi = 0; yi < SC; ++yi)
        nfz[yi] = __ldg(reinterpret_cast<const unsigned short*>(fuzzy + ((((int64_t)n * h + SC * r + yi) * w + (int64_t)t * 4 * NG) >> 3)));
    }
  };
  int tile = blockIdx.x * (TB_THREADS / 32) + (threadIdx.x >> 5);
  request(tile);
  for (; tile < tiles; tile += tstep) {
    const int tyy = tile / tiles_x, txx = tile - tyy * tiles_x;
    const int r = tyy * 4 + (lane >> 3), t = txx * 8 + (lane & 7);
    const int c0 = t * NG * CPG;                     // first working-resolution column of the thread: NG * CPG <= 8 bits, one word
    const int b = c0 & 31;
    unsigned pw[8], fzb[PACKED ? SC : 1];
#pragma unroll
    for (int q = 0; q < 8; ++q) pw[q] = nxt[q] >> b;
#pragma unroll
    for (int yi = 0; yi < (PACKED ? SC : 1); ++yi) fzb[yi] = nfz[yi];
    request(tile + tstep);
    if (r >= th || t >= tpr) continue;
#pragma unroll
    for (int yi = 0; yi < SC; ++yi) {
    const int y = SC * r + yi, below = yi >= SC / 2;
    const unsigned zl = pw[0 + below], zr = pw[2 + below], fl = pw[4 + below], fr = pw[6 + below];
    const int64_t o = ((int64_t)n * h + y) * w + (int64_t)t * 4 * NG;
    unsigned fz[NG];
    if (ens && PACKED) {
      const unsigned bits = fzb[yi];
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        const unsigned nib = (bits >> (4 * k)) & 15u;
        fz[k] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);   // 0 / 1 bytes, like the byte map
      }
    } else if (ens) {
      if (NG == 4) {
        const uint4 v = ldg_stream16(fuzzy + o);
        fz[0] = v.x; fz[1 % NG] = v.y; fz[2 % NG] = v.z; fz[3 % NG] = v.w;
      } else {
        fz[0] = __ldg(reinterpret_cast<const unsigned*>(fuzzy + o));
      }
    }
    // flat runs (most of a trimap): every pixel of the thread's 4 * NG is 0 / 255 / 128 as soon as both halves' bits of
    // its NG * CPG working-resolution columns agree - one test instead of the bit shuffling below (the kernel is bound
    // by instruction issue, not by the bytes it writes)
    constexpr unsigned CM = (1u << (NG * CPG)) - 1u;
    bool anyfz = false;
    if (ens) {
#pragma unroll
      for (int k = 0; k < NG; ++k) anyfz = anyfz || fz[k] != 0u;
    }
    if (!anyfz) {
      const unsigned za = zl & zr & CM, fa = fl & fr & CM, un = (zl | zr | fl | fr) & CM;
      if (za == CM || fa == CM || un == 0u) {
        const unsigned v = za == CM ? 0u : (fa == CM ? 0xFFFFFFFFu : 0x80808080u);
        if (NG == 4) stg_stream16(out + o, make_uint4(v, v, v, v));
        else *reinterpret_cast<unsigned*>(out + o) = v;
        continue;
      }
    }
    unsigned words[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      unsigned zb, fb;   // bit q = output pixel q of the group: Z / F of its quadrant
      if (SC == 4) {     // the 4 pixels share column c0 + k: two on its left half, two on its right half
        zb = ((zl >> k) & 1u) * 3u | ((zr >> k) & 1u) * 12u;
        fb = ((fl >> k) & 1u) * 3u | ((fr >> k) & 1u) * 12u;
      } else {           // pixels: (c, left) (c, right) (c+1, left) (c+1, right), c = c0 + 2k
        const unsigned a = zl >> (2 * k), bb = zr >> (2 * k), c = fl >> (2 * k), d = fr >> (2 * k);
        zb = (a & 1u) | ((bb & 1u) << 1) | ((a & 2u) << 1) | ((bb & 2u) << 2);
        fb = (c & 1u) | ((d & 1u) << 1) | ((c & 2u) << 1) | ((d & 2u) << 2);
      }
      // bits -> bytes: 0 where Z, 255 where F, 128 elsewhere
      const unsigned zm = ((zb & 1u) | ((zb & 2u) << 7) | ((zb & 4u) << 14) | ((zb & 8u) << 21)) * 255u;
      const unsigned fm = ((fb & 1u) | ((fb & 2u) << 7) | ((fb & 4u) << 14) | ((fb & 8u) << 21)) * 255u;
      unsigned word = (0x80808080u & ~zm) | fm;
      if (ens && fz[k]) {
        const unsigned nz = ((fz[k] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | fz[k];
        const unsigned sel = ((nz >> 7) & 0x01010101u) * 255u;
        word = (word & ~sel) | (0x80808080u & sel);
      }
      words[k] = word;
    }
    if (NG == 4) stg_stream16(out + o, make_uint4(words[0], words[1 % NG], words[2 % NG], words[3 % NG]));
    else *reinterpret_cast<unsigned*>(out + o) = words[0];
    }
  }
}

// output rows per tile: with 64 staged rows (x 4 words) every pass of the cross is exactly one cell per thread; tiles of
// 64 output rows (76 staged rows for 5 passes) left 208 of the 256 threads idle in every second sweep
inline int tile_rows(int passes) {
  const int oh = 64 - 2 * (passes + 1);
  return oh >= 32 ? oh : TB_OH;
}

// rows per band of the marching kernel: a band re-computes 12 rows of halo, so bands are as tall as the machine allows -
// about 12 warps per SM over the whole launch, never shorter than 12 rows (2x redundancy), never taller than 64
inline int march_band_rows(int n, int th) {
  const int64_t warps = (int64_t)device_sms() * 12;   // 6 / 24 / 48 warps per SM: 1.944 / 1.904 / 1.903 ms for config 1 against 1.911
  int bands = (int)((warps + n - 1) / n);
  if (bands < 1) bands = 1;
  int rows = (th + bands - 1) / bands;
  rows = rows < 12 ? 12 : (rows > 64 ? 64 : rows);
  return rows;
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" size_t vu_trimap_bits_workspace_bytes(int n, int th, int tw) {
  if (n <= 0 || th <= 0 || tw <= 0) return 0;
  return (size_t)8 * n * th * ((tw + 31) / 32) * sizeof(unsigned);
}

extern "C" int vu_trimap_bits(const uint8_t* mask, const uint8_t* fuzzy, const uint8_t* flags, int n, int h, int w, int th, int tw, int iters,
                              uint8_t* out, void* workspace, size_t workspace_bytes, vu_stream_t stream) {
  VU_REQUIRE(mask && out && workspace && n >= 0 && h > 0 && w > 0 && th > 0 && tw > 0 && iters >= 0);
  VU_REQUIRE((fuzzy == nullptr) == (flags == nullptr));
  const int sc = (w == 2 * tw && h == 2 * th) ? 2 : ((w == 4 * tw && h == 4 * th) ? 4 : 0);
  if (sc == 0 || tw % 4 != 0 || iters > TB_MAXR || n > 65535) return VU_ERR_UNSUPPORTED;
  const void* ptrs[] = {mask, fuzzy, out};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return VU_ERR_UNSUPPORTED;
  if (workspace_bytes < vu_trimap_bits_workspace_bytes(n, th, tw) || (reinterpret_cast<uintptr_t>(workspace) & 3)) return VU_ERR_WORKSPACE;
  if (n == 0) return VU_OK;
  const int wpr = (tw + 31) / 32;
  unsigned* planes = static_cast<unsigned*>(workspace);
  const int oh = tile_rows(iters);
  dim3 ga((tw + TB_OW - 1) / TB_OW, (th + oh - 1) / oh, n);
  const int ng = (w % 16 == 0) ? 4 : 1;
  const int64_t items = (int64_t)h * (w / (4 * ng));
  int64_t bx = (items + TB_THREADS - 1) / TB_THREADS;
  const int64_t cap = ((int64_t)device_sms() * 8 + n - 1) / n;
  if (bx > cap) bx = cap;
  dim3 gb((unsigned)(bx < 1 ? 1 : bx), n);
#define VU_UP(SCV)                                                                                                                   \
  do {                                                                                                                               \
    trimap_bits_kernel<SCV, false><<<ga, TB_THREADS, 0, S(stream)>>>(mask, fuzzy, flags, h, w, th, tw, iters, planes, n, wpr, oh);      \
    if (ng == 4) trimap_up_bits_kernel<SCV, 4, false><<<gb, TB_THREADS, 0, S(stream)>>>(planes, n, th, tw, wpr, fuzzy, flags, out);   \
    else trimap_up_bits_kernel<SCV, 1, false><<<gb, TB_THREADS, 0, S(stream)>>>(planes, n, th, tw, wpr, fuzzy, flags, out);           \
  } while (0)
  if (sc == 2) VU_UP(2);
  else VU_UP(4);
#undef VU_UP
  note_launch();
  VU_RETURN_LAUNCH();
}

// The same tail from the bit planes vu_cf_alpha_up_fuzzy leaves behind: mask_bits [n][th][tw/8] = nearest-sampled
// alpha >= 128 at the working resolution, fuzzy_bits [n][h][w/8] = one bit per full-resolution pixel (may be NULL
// together with flags: mask-only trimap).  No full-resolution byte map is read at all.  tw % 16 == 0.
extern "C" int vu_trimap_bits_packed(const uint8_t* mask_bits, const uint8_t* fuzzy_bits, const uint8_t* flags, int n, int h, int w, int th, int tw,
                                     int iters, uint8_t* out, void* workspace, size_t workspace_bytes, vu_stream_t stream) {
  VU_REQUIRE(mask_bits && out && workspace && n >= 0 && h > 0 && w > 0 && th > 0 && tw > 0 && iters >= 0);
  VU_REQUIRE((fuzzy_bits == nullptr) == (flags == nullptr));
  const int sc = (w == 2 * tw && h == 2 * th) ? 2 : ((w == 4 * tw && h == 4 * th) ? 4 : 0);
  if (sc == 0 || tw % 16 != 0 || iters > TB_MAXR || n > 65535) return VU_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(mask_bits) & 1) || (reinterpret_cast<uintptr_t>(fuzzy_bits) & 3))
    return VU_ERR_UNSUPPORTED;
  if (workspace_bytes < vu_trimap_bits_workspace_bytes(n, th, tw) || (reinterpret_cast<uintptr_t>(workspace) & 3)) return VU_ERR_WORKSPACE;
  if (n == 0) return VU_OK;
  const int wpr = (tw + 31) / 32;
  unsigned* planes = static_cast<unsigned*>(workspace);
  const int oh = tile_rows(iters);
  dim3 ga((tw + TB_OW - 1) / TB_OW, (th + oh - 1) / oh, n);
  const int64_t items = (int64_t)h * (w / 16);
  int64_t bx = (items + TB_THREADS - 1) / TB_THREADS;
  const int64_t cap = ((int64_t)device_sms() * 8 + n - 1) / n;
  if (bx > cap) bx = cap;
  dim3 gb((unsigned)(bx < 1 ? 1 : bx), n);
  // the reference's 5 passes at a working resolution of whole 32-pixel words: the marching kernel
  static const bool march_off = [] { const char* e = getenv("VU_TRIMAP_MARCH"); return e && e[0] == '0'; }();   // A/B switch
  const bool march = !march_off && iters == 5 && tw % 32 == 0 && wpr <= 32 && (reinterpret_cast<uintptr_t>(mask_bits) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(fuzzy_bits) & 15) == 0;
  const int band = march_band_rows(n, th);
  dim3 gm(((th + band - 1) / band + TM_WARPS - 1) / TM_WARPS, n);
  if (sc == 2) {
    if (march) trimap_bits_march_kernel<2, 5><<<gm, 32 * TM_WARPS, 0, S(stream)>>>(mask_bits, fuzzy_bits, flags, h, w, th, tw, planes, n, wpr, band);
    else trimap_bits_kernel<2, true><<<ga, TB_THREADS, 0, S(stream)>>>(mask_bits, fuzzy_bits, flags, h, w, th, tw, iters, planes, n, wpr, oh);
    trimap_up_bits_kernel<2, 4, true><<<gb, TB_THREADS, 0, S(stream)>>>(planes, n, th, tw, wpr, fuzzy_bits, flags, out);
  } else {
    if (march) trimap_bits_march_kernel<4, 5><<<gm, 32 * TM_WARPS, 0, S(stream)>>>(mask_bits, fuzzy_bits, flags, h, w, th, tw, planes, n, wpr, band);
    else trimap_bits_kernel<4, true><<<ga, TB_THREADS, 0, S(stream)>>>(mask_bits, fuzzy_bits, flags, h, w, th, tw, iters, planes, n, wpr, oh);
    trimap_up_bits_kernel<4, 4, true><<<gb, TB_THREADS, 0, S(stream)>>>(planes, n, th, tw, wpr, fuzzy_bits, flags, out);
  }
  note_launch();
  VU_RETURN_LAUNCH();
}
