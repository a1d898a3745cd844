// One pass per frame of the bg_step live loop (tools/unscreen/bg_offline.py:150-172 == bg.py:85-100), the CNN stage
// skipped (alpha := gated mask):
//
//     raw   = u8(|f32(frame) - f32(bg)|);  g = BGR2GRAY(raw);  g[g > thr] = 255;  g = dilate_mask(g, 4, 2)
//     alpha = mask * (g // 255)                                                   the difference gate, :154-160
//     bgimg[alpha == 0] = frame[alpha == 0];  fg = get_fg(frame, alpha, bgimg)    :171-172
//     B     = nearest-down(alpha) >= 128                                          what the trimap of :166 needs (bits)
//
// vu_bgdiff_gate + vu_get_fg read the frame twice (and the gate's 4-byte loads are bound by instruction issue, not by
// HBM).  Here a CTA owns a 224 x 32 output tile: the TMA unit fetches the tile plus its halo (4 rows above, 2 below for the
// two 4x4-ellipse dilations; 16 columns either side: the unit wants 16-byte aligned box rows, and 16 pixels are 48
// bytes) of the frame AND of the background into shared memory as two boxes of 38 rows x 768 bytes - coordinates
// outside the image are zero-filled by the unit, which is exactly "no difference" for the gate - and everything else
// happens on chip: the binary difference map (VABSDIFF4 + two IDP.4A per pixel, REDUX.OR to 32-pixel words), both
// dilations on the words, the gated matte, and get_fg for the pixels the gate let through, with the frame and
// background pixels taken from the staged tile.  Per frame: frame 3P (x 1.36 for the halo) + mask P in, alpha P +
// fg 3P out; the background stays in L2.  Tiles whose masks are all zero skip everything but the stores.
#include "vu_common.cuh"
#include "vu_tma.cuh"

namespace vu {
namespace {

constexpr int BS_THREADS = 512;   // three CTAs per SM by shared memory: 48 warps in flight
constexpr int BS_TW = 224, BS_TH = 32;          // output tile: 67 KB of shared memory per CTA, three CTAs per SM
constexpr int BS_SW = 256;                      // staged pixels per row: 16 + 224 + 16
constexpr int BS_ROWS = BS_TH + 6;              // staged rows: 4 above, 2 below
constexpr int BS_ROWB = BS_SW * 3;              // bytes per staged row
constexpr int BS_WORDS = BS_SW / 32;            // 8 bit-words per staged row
constexpr int BS_TILE_BYTES = BS_ROWS * BS_ROWB;   // 41472

__device__ __forceinline__ int trunc_clamp255b(float x) { return f32_trunc_nonneg(fminf(fmaxf(x, 0.f), 255.f)); }

// pixel x+dx of a bit-packed row (bit i of word j = staged pixel 32 j + i): dx in -2..1 (the 4x4 ellipse's rows 1..3)
__device__ __forceinline__ unsigned hor_or8(const unsigned* row, int j) {
  const unsigned c = row[j];
  const unsigned p = j > 0 ? row[j - 1] : 0u, n = j < BS_WORDS - 1 ? row[j + 1] : 0u;
  return c | __funnelshift_l(p, c, 1) | __funnelshift_l(p, c, 2) | __funnelshift_r(c, n, 1);
}

// SC: 0 = no trimap bits, 2 / 4 = frame size / working size (mbits [n][h/SC][w/SC/8])
template <int SC>
__global__ void __launch_bounds__(BS_THREADS, 3) bgstep_frame_kernel(const __grid_constant__ CUtensorMap fmap, const __grid_constant__ CUtensorMap bmap,
                                                                  int bg_per_frame, const uint8_t* __restrict__ masks, int h, int w, int thr,
                                                                  uint8_t* __restrict__ alpha_out, uint8_t* __restrict__ fg_out,
                                                                  uint8_t* __restrict__ mbits) {
  extern __shared__ __align__(128) uint8_t tiles[];   // frame tile, then background tile: [BS_ROWS][BS_ROWB] each
  __shared__ unsigned B[BS_ROWS][BS_WORDS], H[BS_ROWS][BS_WORDS], D[BS_ROWS][BS_WORDS];
  __shared__ HsvTab tab;
  __shared__ float ktab[256];
  __shared__ __align__(8) unsigned long long bar;
  uint8_t* ft = tiles;
  uint8_t* bt = tiles + BS_TILE_BYTES;
  const int n = blockIdx.z;
  const int X0 = blockIdx.x * BS_TW, Y0 = blockIdx.y * BS_TH;
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(&bar);
  // the masks of this thread's output items come first: alpha = mask * gate, so a tile whose masks are all zero (most of
  // a frame: the person covers a fraction of it) has alpha = 0, fg = black and no trimap bits whatever the frame holds,
  // and neither the frame nor the background is fetched for it
  constexpr int GPR = BS_TW / 16;                                          // 14 items of 16 pixels per tile row
  constexpr int ITEMS = BS_TH * GPR, IPT = (ITEMS + BS_THREADS - 1) / BS_THREADS;
  const int64_t fpix = (int64_t)n * h * w;
  const uint8_t* mk = masks + fpix;
  uint4 mreg[IPT];
  unsigned anym = 0;
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    const int i = threadIdx.x + k * BS_THREADS;
    const int ty = i / GPR, tg = i - ty * GPR;
    const int gy = Y0 + ty, gx = X0 + 16 * tg;
    mreg[k] = (i < ITEMS && gy < h && gx < w) ? ldg_stream16(mk + (int64_t)gy * w + gx) : make_uint4(0u, 0u, 0u, 0u);
    anym |= mreg[k].x | mreg[k].y | mreg[k].z | mreg[k].w;
  }
  if (!__syncthreads_or(anym != 0u)) {
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const int i = threadIdx.x + k * BS_THREADS;
      const int ty = i / GPR, tg = i - ty * GPR;
      const int gy = Y0 + ty, gx = X0 + 16 * tg;
      if (i >= ITEMS || gy >= h || gx >= w) continue;
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      stg_stream16(alpha_out + fpix + (int64_t)gy * w + gx, z);
      uint8_t* o = fg_out + (fpix + (int64_t)gy * w + gx) * 3;
      stg_stream16(o, z); stg_stream16(o + 16, z); stg_stream16(o + 32, z);
      if constexpr (SC != 0) {
        const int tw8 = (w / SC) >> 3;
        if ((gy % SC) == 0) {
          if (SC == 2) mbits[((int64_t)n * (h / SC) + gy / SC) * tw8 + (gx >> 4)] = 0;
          else if (!(tg & 1)) mbits[((int64_t)n * (h / SC) + gy / SC) * tw8 + (gx >> 5)] = 0;
        }
      }
    }
    return;
  }
  if (threadIdx.x == 0) {
    tma::mbar_init(mbar, 1);
    tma::mbar_fence_init();
    tma::mbar_expect_tx(mbar, 2 * BS_TILE_BYTES);
    // innermost coordinate in 32-bit elements: pixel X0 - 16 starts at byte 3 (X0 - 16), a multiple of 48
    const int c0 = ((X0 - 16) * 3) / 4;
    tma::load_3d((unsigned)__cvta_generic_to_shared(ft), &fmap, c0, Y0 - 4, n, mbar);
    tma::load_3d((unsigned)__cvta_generic_to_shared(bt), &bmap, c0, Y0 - 4, bg_per_frame ? n : 0, mbar);
  }
  hsv_tab_init(tab);
  for (int a = threadIdx.x; a < 256; a += BS_THREADS) ktab[a] = __fsub_rn(1.f, __fdiv_rn((float)a, 255.f));
  __syncthreads();
  tma::mbar_wait(mbar, 0);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int teff = thr < 254 ? thr : 254;           // gray > thr || gray == 255
  // ---- B = (gray(|frame - bg|) > thr), one bit per staged pixel; outside the image both tiles are zero: no difference ----
  {
    constexpr unsigned WH = 14u | (75u << 8) | (38u << 16), WL = 151u | (35u << 8) | (70u << 16);   // BGR2GRAY's 15-bit weights, split in bytes
    const unsigned limit = ((unsigned)teff + 1u) << 15;   // gray > teff  <=>  weighted sum + 16384 >= (teff + 1) << 15
    for (int it = warp; it < BS_ROWS * 2; it += BS_THREADS / 32) {   // a warp: 32 groups of 4 pixels = half a staged row
      const int r = it >> 1, g = ((it & 1) << 5) + lane;
      const unsigned* f4 = reinterpret_cast<const unsigned*>(ft + r * BS_ROWB + 12 * g);
      const unsigned* b4 = reinterpret_cast<const unsigned*>(bt + r * BS_ROWB + 12 * g);
      const unsigned d0 = __vabsdiffu4(f4[0], b4[0]), d1 = __vabsdiffu4(f4[1], b4[1]), d2 = __vabsdiffu4(f4[2], b4[2]);
      const unsigned ps[4] = {d0, __byte_perm(d0, d1, 0x0543), __byte_perm(d1, d2, 0x0432), d2 >> 8};
      unsigned nib = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned num = __dp4a(ps[k], WH, 0u) * 256u + __dp4a(ps[k], WL, 16384u);
        nib |= (unsigned)(num >= limit) << k;
      }
      const unsigned word = __reduce_or_sync(0xFFu << (lane & 24), nib << (4 * (lane & 7)));
      if ((lane & 7) == 0) B[r][((it & 1) << 2) + (lane >> 3)] = word;
    }
  }
  __syncthreads();
  // in-image mask of a staged word (bit i = pixel X0 - 16 + 32 j + i): cells outside the image are reset to 0 between the
  // iterations, which is what cv2's "taps outside the image are ignored" means under iteration
  auto inside = [&](int r, int j) -> unsigned {
    const int gy = Y0 - 4 + r;
    if ((unsigned)gy >= (unsigned)h) return 0u;
    const int x0 = X0 - 16 + 32 * j;
    const int lo = max(0, -x0), hi = min(32, w - x0);
    if (hi <= lo) return 0u;
    const unsigned m_hi = hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
    return m_hi & ~((1u << lo) - 1u);
  };
  // MORPH_ELLIPSE(4,4) = rows 0010 / 1111 / 1111 / 1111, anchor (2,2): dst(y,x) = src(y-2,x) | OR_{dy -1..1, dx -2..1} src(y+dy,x+dx)
  for (int it = 0; it < 2; ++it) {   // B -> H -> D: one sweep per iteration (the horizontal ORs of the three rows recomputed per cell)
    const unsigned(*src)[BS_WORDS] = it == 0 ? B : H;
    unsigned(*dst)[BS_WORDS] = it == 0 ? H : D;
    for (int i = threadIdx.x; i < BS_ROWS * BS_WORDS; i += BS_THREADS) {
      const int r = i >> 3, j = i & 7;
      unsigned v = hor_or8(src[r], j);
      if (r >= 1) v |= hor_or8(src[r - 1], j);
      if (r + 1 < BS_ROWS) v |= hor_or8(src[r + 1], j);
      if (r >= 2) v |= src[r - 2][j];
      dst[r][j] = v & inside(r, j);
    }
    __syncthreads();
  }
  // ---- output: 16 pixels per item.  alpha = mask where the dilated bit is set; fg = get_fg(frame, alpha, bg patched where
  //      alpha == 0): black where the gate closed (the patched background is the pixel itself), the HSV arithmetic of
  //      utils/fgfuncs.py:84-110 elsewhere, frame and background pixels from the staged tiles ----
  uint8_t* ao = alpha_out + fpix;
  uint8_t* fo = fg_out + fpix * 3;
#pragma unroll
  for (int kk = 0; kk < IPT; ++kk) {   // every thread runs every round: the shuffle below wants whole warps
    const int i = threadIdx.x + kk * BS_THREADS;
    const int ty = i / GPR, tg = i - ty * GPR;
    const int gy = Y0 + ty, gx = X0 + 16 * tg;
    const bool act = i < ITEMS && gy < h && gx < w;
    unsigned aw[4] = {0u, 0u, 0u, 0u};
    if (act) {
      const int sx = 16 + 16 * tg;   // staged pixel index of gx: bits sx .. sx+15 = one half of a word
      const unsigned bits = (D[ty + 4][sx >> 5] >> (sx & 31)) & 0xFFFFu;
      const unsigned mw[4] = {mreg[kk].x, mreg[kk].y, mreg[kk].z, mreg[kk].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned nb = (bits >> (4 * k)) & 15u;
        aw[k] = mw[k] & (((nb & 1u) | ((nb & 2u) << 7) | ((nb & 4u) << 14) | ((nb & 8u) << 21)) * 255u);
      }
      stg_stream16(ao + (int64_t)gy * w + gx, make_uint4(aw[0], aw[1], aw[2], aw[3]));
    }
    if constexpr (SC != 0) {   // B bits of the trimap source: pixels (SC*r, SC*c), rows with gy % SC == 0
      constexpr int NS = 16 / SC;
      unsigned mb = 0;
      if ((aw[0] | aw[1] | aw[2] | aw[3]) & 0x80808080u) {
#pragma unroll
        for (int c = 0; c < NS; ++c) {
          const int k = SC * c;
          mb |= ((aw[k >> 2] >> (8 * (k & 3) + 7)) & 1u) << c;
        }
      }
      const int tw8 = (w / SC) >> 3;
      if (SC == 2) {
        if (act && (gy % SC) == 0) mbits[((int64_t)n * (h / SC) + gy / SC) * tw8 + (gx >> 4)] = (uint8_t)mb;
      } else {
        const unsigned hi = __shfl_down_sync(0xffffffffu, mb, 1);   // items 2j, 2j + 1 of a row are neighbouring lanes (GPR is even)
        if (act && (gy % SC) == 0 && !(tg & 1)) mbits[((int64_t)n * (h / SC) + gy / SC) * tw8 + (gx >> 5)] = (uint8_t)(mb | (hi << 4));
      }
    }
    if (!act) continue;
    uint8_t* o = fo + ((int64_t)gy * w + gx) * 3;
    if ((aw[0] | aw[1] | aw[2] | aw[3]) == 0u) {
#pragma unroll
      for (int k = 0; k < 3; ++k) stg_stream16(o + 16 * k, make_uint4(0u, 0u, 0u, 0u));
      continue;
    }
    const unsigned* fw = reinterpret_cast<const unsigned*>(ft + (ty + 4) * BS_ROWB + (16 + 16 * tg) * 3);
    const unsigned* qw = reinterpret_cast<const unsigned*>(bt + (ty + 4) * BS_ROWB + (16 + 16 * tg) * 3);
    uint4 ov[3];
    unsigned* ow = reinterpret_cast<unsigned*>(ov);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (aw[g] == 0u) {
        ow[3 * g] = ow[3 * g + 1] = ow[3 * g + 2] = 0u;
        continue;
      }
      int c[12], q[12], oo[12];
      unpack12(fw[3 * g], fw[3 * g + 1], fw[3 * g + 2], c);
      unpack12(qw[3 * g], qw[3 * g + 1], qw[3 * g + 2], q);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int a = (int)((aw[g] >> (8 * p)) & 255u);
        const bool patch = a == 0;   // bg_offline.py:171
        int ih, is, iv, bh = 0, bs = 0, bv = 0;
        bgr2hsv_px(c[3 * p], c[3 * p + 1], c[3 * p + 2], tab, ih, is, iv);
        // a == 255 (almost every pixel the gate lets through: the masks are binary): k = 1 - 255/255 = 0 exactly, the
        // background's HSV is multiplied by it: no need to compute it
        // ... nor the float arithmetic: trunc(clamp(x - 0 * y)) = x
        int fh = ih, fs = is, fv = iv;
        if (a != 255) {
          bgr2hsv_px(patch ? c[3 * p] : q[3 * p], patch ? c[3 * p + 1] : q[3 * p + 1], patch ? c[3 * p + 2] : q[3 * p + 2], tab, bh, bs, bv);
          const float k = ktab[a];
          fh = trunc_clamp255b(__fsub_rn(u8_to_f32(ih), __fmul_rn(k, u8_to_f32(bh))));
          fs = trunc_clamp255b(__fsub_rn(u8_to_f32(is), __fmul_rn(k, u8_to_f32(bs))));
          fv = trunc_clamp255b(__fsub_rn(u8_to_f32(iv), __fmul_rn(k, u8_to_f32(bv))));
        }
        hsv2bgr_px(fh, fs, fv, tab, oo[3 * p], oo[3 * p + 1], oo[3 * p + 2]);
      }
      pack12(oo, ow[3 * g], ow[3 * g + 1], ow[3 * g + 2]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) stg_stream16(o + 16 * k, ov[k]);
  }
}

bool make_map(tma::EncodeTiledFn enc, CUtensorMap* map, const uint8_t* base, int n, int h, int w) {
  const cuuint64_t dims[3] = {(cuuint64_t)w * 3 / 4, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[2] = {(cuuint64_t)w * 3, (cuuint64_t)w * 3 * h};
  const cuuint32_t box[3] = {BS_ROWB / 4, BS_ROWS, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_bgstep_frames(const uint8_t* frames, const uint8_t* bg, const uint8_t* masks, int n, int h, int w, int bg_frames, int thr,
                                uint8_t* alpha, uint8_t* fg, uint8_t* mask_bits, int scale, vu_stream_t stream) {
  VU_REQUIRE(frames && bg && masks && alpha && fg && n >= 0 && h > 0 && w > 0);
  VU_REQUIRE(bg_frames == 1 || bg_frames == n);
  VU_REQUIRE(mask_bits ? (scale == 2 || scale == 4) : true);
  if (w % 16 != 0 || n > 65535) return VU_ERR_UNSUPPORTED;
  if (mask_bits && (h % scale != 0 || (w / scale) % 8 != 0)) return VU_ERR_UNSUPPORTED;
  const void* ptrs[] = {frames, bg, masks, alpha, fg};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  tma::EncodeTiledFn enc = tma::encode_tiled_fn();
  if (!enc) return VU_ERR_UNSUPPORTED;
  CUtensorMap fmap, bmap;
  if (!make_map(enc, &fmap, frames, n, h, w) || !make_map(enc, &bmap, bg, bg_frames, h, w)) return VU_ERR_UNSUPPORTED;
  dim3 grid((w + BS_TW - 1) / BS_TW, (h + BS_TH - 1) / BS_TH, n);
  if (grid.y > 65535) return VU_ERR_UNSUPPORTED;
  const size_t smem = 2 * (size_t)BS_TILE_BYTES;
  static bool configured = false;
  if (!configured) {
    int e = record_cuda(cudaFuncSetAttribute(bgstep_frame_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!e) e = record_cuda(cudaFuncSetAttribute(bgstep_frame_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!e) e = record_cuda(cudaFuncSetAttribute(bgstep_frame_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (e) return e;
    configured = true;
  }
  const int sc = mask_bits ? scale : 0;
  if (sc == 0) bgstep_frame_kernel<0><<<grid, BS_THREADS, smem, S(stream)>>>(fmap, bmap, bg_frames == n && n > 1 ? 1 : 0, masks, h, w, thr, alpha, fg, mask_bits);
  else if (sc == 2) bgstep_frame_kernel<2><<<grid, BS_THREADS, smem, S(stream)>>>(fmap, bmap, bg_frames == n && n > 1 ? 1 : 0, masks, h, w, thr, alpha, fg, mask_bits);
  else bgstep_frame_kernel<4><<<grid, BS_THREADS, smem, S(stream)>>>(fmap, bmap, bg_frames == n && n > 1 ? 1 : 0, masks, h, w, thr, alpha, fg, mask_bits);
  VU_RETURN_LAUNCH();
}
