// Geometric pre-steps of the replacement path (tools/replace/replace.py:69-72):
//
//   shift_fg   unscreen/utils/imgprocess.py:55-64   cv2.warpAffine(img, [[1,0,dx],[0,1,dy]], (W,H))
//   rescale_fg unscreen/utils/imgprocess.py:40-52   cv2.resize(fx=fy=s, INTER_CUBIC) + centre crop to the input size
//
// shift: cv2's fixed-point warp.  For a pure translation the source position of every destination
// pixel is (x + ox, y + oy) plus ONE 5-bit sub-pixel fraction (fx, fy) for the whole image, so the warp is a constant
// 2x2 filter: dst = (32*[(32-fy)*((32-fx)*a + fx*b) + fy*((32-fx)*c + fx*d)] + 16384) >> 15 with taps outside the
// image reading 0.  The tile kernel lets the TMA unit do both the translation and the border: the box is fetched at
// byte coordinate (x0 + ox)*C, row y0 + oy - any integers - and the unit zero-fills what lies outside the frame.
//
// rescale: separable float bicubic (a = -0.75) with replicated borders, rounded half to even - the model of the
// Intel-IPP path cv2 4.13 dispatches this call to (oracle/cvmodel.py:resize_cubic_crop holds the parity note).  Only
// the cropped window is computed.
#include <cmath>
#include <cstdlib>

#include "vu_common.cuh"
#include "vu_tma.cuh"

namespace vu {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// shift
// ---------------------------------------------------------------------------------------------------------------

// Y0(y) = rint((y + b2) * 1024) + 16: integer row y + oy and fraction fy of destination row y  (b2 = -dy)
__host__ __device__ inline void shift_row(int y, double b2, int& sy, int& fy) {
#ifdef __CUDA_ARCH__
  const long long Y = ((long long)rint(((double)y + b2) * 1024.0) + 16) >> 5;
#else
  const long long Y = ((long long)std::nearbyint(((double)y + b2) * 1024.0) + 16) >> 5;
#endif
  long long s = Y >> 5;
  s = s < -32768 ? -32768 : (s > 32767 ? 32767 : s);   // cv2 keeps the integer part as a saturated int16
  sy = (int)s;
  fy = (int)(Y & 31);
}

// any shape, any alignment: one thread per destination byte
template <int C>
__global__ void __launch_bounds__(256) shift_generic_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int h, int w, int ox,
                                                            int fx, double b2) {
  const int64_t wc = (int64_t)w * C, per = (int64_t)h * wc, total = per * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / per, r = i - f * per;
    const int y = (int)(r / wc), j = (int)(r - (int64_t)y * wc);
    const int x = j / C, ch = j - x * C;
    int sy, fy;
    shift_row(y, b2, sy, fy);
    int sx = x + ox;
    sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);
    const uint8_t* fr = src + f * per;
    auto tap = [&](int yy, int xx) -> int {
      return ((unsigned)yy < (unsigned)h && (unsigned)xx < (unsigned)w) ? (int)__ldg(fr + ((int64_t)yy * w + xx) * C + ch) : 0;
    };
    const int a = tap(sy, sx), b = tap(sy, sx + 1), c = tap(sy + 1, sx), d = tap(sy + 1, sx + 1);
    const int acc = 32 * ((32 - fy) * ((32 - fx) * a + fx * b) + fy * ((32 - fx) * c + fx * d));
    dst[i] = (uint8_t)((acc + 16384) >> 15);
  }
}

constexpr int SH_TB = 232;         // destination bytes of a tile row: 29 lanes x 8 bytes
constexpr int SH_TR = 32;          // destination rows of a tile: 8 warps x 4 consecutive rows
constexpr int SH_RPW = 4;
constexpr int SH_BOX_W = 256, SH_BOX_H = SH_TR + 1;

// horizontal pass of 8 destination bytes of one source row: H = (32-fx)*a + fx*b (13 bits) for every byte, as four
// words of two 16-bit lanes: he[k] = bytes (4k, 4k+2), ho[k] = bytes (4k+1, 4k+3).  `p` points at the 8-byte group of
// the tile row that holds the first source byte, `m` (0..7) is where in the group it sits.
template <int C>
__device__ __forceinline__ void shift_hrow(const uint8_t* p, int m, unsigned wx0, unsigned wx1, unsigned (&s)[3], unsigned (&he)[2],
                                           unsigned (&ho)[2]) {
  const uint2 a = *reinterpret_cast<const uint2*>(p), b = *reinterpret_cast<const uint2*>(p + 8);
  const unsigned w4 = *reinterpret_cast<const unsigned*>(p + 16);
  const bool q = (m & 4) != 0;
  const unsigned x0 = q ? a.y : a.x, x1 = q ? b.x : a.y, x2 = q ? b.y : b.x, x3 = q ? w4 : b.y;
  const unsigned sel = 0x3210u + 0x1111u * (unsigned)(m & 3);
  s[0] = __byte_perm(x0, x1, sel);
  s[1] = __byte_perm(x1, x2, sel);
  s[2] = __byte_perm(x2, x3, sel);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const unsigned t = __byte_perm(s[k], s[k + 1], C == 1 ? 0x4321 : 0x6543);   // the same bytes, C further right
    he[k] = (s[k] & 0x00FF00FFu) * wx0 + (t & 0x00FF00FFu) * wx1;
    ho[k] = __byte_perm(s[k], 0u, 0x4341) * wx0 + __byte_perm(t, 0u, 0x4341) * wx1;
  }
}

// The TMA unit wants every row of a box to start on a 16-byte boundary of global memory, so the box is fetched from
// the aligned column below the (arbitrary) source column and the remainder (0..15 bytes) is taken out when the
// threads pick their bytes from shared memory: 232 + 15 + C <= 256, the widest box there is.
template <int C>
__global__ void __launch_bounds__(256) shift_tile_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ dst, int h, int64_t wc, int oxc,
                                                         int oy, int fx, int fy) {
  __shared__ __align__(128) uint8_t tile[SH_BOX_H * SH_BOX_W];
  __shared__ __align__(8) unsigned long long bar;
  const int j0 = blockIdx.x * SH_TB, y0 = blockIdx.y * SH_TR, f = blockIdx.z;
  const int start = j0 + oxc, astart = start & ~15, m = start - astart;
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    tma::mbar_init(mbar, 1);
    tma::mbar_fence_init();
    tma::mbar_expect_tx(mbar, SH_BOX_H * SH_BOX_W);
    tma::load_3d((unsigned)__cvta_generic_to_shared(tile), &tmap, astart, y0 + oy, f, mbar);
  }
  __syncthreads();
  tma::mbar_wait(mbar, 0);
  const int lane = threadIdx.x & 31, ly = SH_RPW * (threadIdx.x >> 5);
  const int j = j0 + 8 * lane, y = y0 + ly;
  if (lane >= SH_TB / 8 || j >= wc || y >= h) return;
  uint8_t* out = dst + ((int64_t)f * h + y) * wc + j;
  const uint8_t* p = tile + ly * SH_BOX_W + 8 * lane + (m & 8);
  const unsigned wx0 = 32 - fx, wx1 = fx, wy = (unsigned)(32 - fy) | ((unsigned)fy << 8);
  const bool copy = fx == 0 && fy == 0;   // whole-pixel shift
  unsigned s[3], het[2], hot[2];
  shift_hrow<C>(p, m & 7, wx0, wx1, s, het, hot);
#pragma unroll
  for (int i = 0; i < SH_RPW; ++i) {
    if (y + i >= h) break;
    uint2 o;
    if (copy) {
      o = make_uint2(s[0], s[1]);
      shift_hrow<C>(p + (i + 1) * SH_BOX_W, m & 7, wx0, wx1, s, het, hot);
    } else {
      unsigned heb[2], hob[2], r[2];
      shift_hrow<C>(p + (i + 1) * SH_BOX_W, m & 7, wx0, wx1, s, heb, hob);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        // (H_top*(32-fy) + H_bottom*fy + 512) >> 10 per byte: the two H of a byte side by side, one two-way dot product
        const unsigned b0 = __dp2a_lo(__byte_perm(het[k], heb[k], 0x5410), wy, 512u) >> 10;
        const unsigned b1 = __dp2a_lo(__byte_perm(hot[k], hob[k], 0x5410), wy, 512u) >> 10;
        const unsigned b2 = __dp2a_lo(__byte_perm(het[k], heb[k], 0x7632), wy, 512u) >> 10;
        const unsigned b3 = __dp2a_lo(__byte_perm(hot[k], hob[k], 0x7632), wy, 512u) >> 10;
        r[k] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        het[k] = heb[k];
        hot[k] = hob[k];
      }
      o = make_uint2(r[0], r[1]);
    }
    *reinterpret_cast<uint2*>(out + i * wc) = o;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// rescale (bicubic up-scale + centre crop)
// ---------------------------------------------------------------------------------------------------------------

// source position of destination coordinate d: first tap is s - 1; coefficients in double from the double
// position, rounded to float once (oracle/cvmodel.py:cubic_axis; compiled with -fmad=false: every op rounds)
__device__ __forceinline__ int cubic_taps(int d, double factor, float (&c)[4]) {
  const double p = ((double)d + 0.5) / factor - 0.5;
  const double s = floor(p), x = p - s;
  const double A = -0.75, x1 = x + 1.0, y = 1.0 - x;
  const double c0 = ((A * x1 - 5.0 * A) * x1 + 8.0 * A) * x1 - 4.0 * A;
  const double c1 = ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0;
  const double c2 = ((A + 2.0) * y - (A + 3.0)) * y * y + 1.0;
  const double c3 = 1.0 - c0 - c1 - c2;
  c[0] = (float)c0; c[1] = (float)c1; c[2] = (float)c2; c[3] = (float)c3;
  return (int)s;
}
__device__ __forceinline__ int cubic_pos(int d, double factor) { return (int)floor(((double)d + 0.5) / factor - 0.5); }

constexpr int RS_TH = 32;            // destination rows between two refills of the source tile
constexpr int RS_SRH = RS_TH + 4;    // source rows under them (factor >= 1), at most
constexpr int RS_STRIP = 128;        // destination rows a CTA walks down
template <int C>
struct RsCfg {
  static constexpr int JW = C == 3 ? 192 : 256;            // destination bytes of a strip row = threads of the CTA
  static constexpr int TW = JW / C;                        // destination pixels of a strip row
  static constexpr int SRW = ((TW + 4) * C + 6 + 3) / 4 * 4;   // floats of a source tile row: (TW + 4) pixels, widened to whole words
  static constexpr int LW = C == 3 ? 64 : 128;             // loader threads per source row (>= SRW / 4)
  static_assert(SRW / 4 <= LW && JW % LW == 0, "loader layout");
};

// np.clip(np.rint(v), 0, 255) in the low byte of the result: saturating first is the same (rint is monotone), and the
// add of 1.5 * 2^23 leaves rint(v) - ties to even - in the low mantissa bits without touching the conversion unit
__device__ __forceinline__ unsigned round_u8(float v) { return __float_as_uint(fminf(fmaxf(v, 0.f), 255.f) + 12582912.f); }

// A CTA owns a strip of TW destination pixels and walks down RS_STRIP destination rows, one thread per destination
// byte column.  The four horizontal taps of the column (replicated at the image border) and their weights live in
// registers for the whole walk; the vertical taps of every row of the strip are tabulated once.  Every 32 destination
// rows the source rows under them are staged in shared memory as floats (aligned 32-bit loads, each byte converted
// once).  A thread then marches: the horizontal pass of a source row is computed when the vertical window reaches it
// and kept in a four-deep register window (up-scaling: the window moves by at most one row per destination row), the
// vertical pass runs from those registers, and the byte goes straight out - a warp stores 32 consecutive bytes.
// Both passes are one multiply and three fused multiply-adds per value, like the oracle.  While a group is marched the
// loader threads already hold the words of the next group in registers.  (Tried and dropped: two columns per thread
// with the packed FFMA2 / FMUL2 of sm_100 - fewer issue slots, but half the warps per SM; 1.8 ms against 1.1 ms per
// 120 x 1080p x 3.  Note for anyone who wants unfused arithmetic there: ptxas contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2 even under --fmad=false.)
template <int C>
__global__ void __launch_bounds__(RsCfg<C>::JW) rescale_cubic_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w,
                                                                      double factor, int yoff, int xoff, int words) {
  using K = RsCfg<C>;
  __shared__ __align__(16) float srcf[RS_SRH * K::SRW];
  __shared__ int ys[RS_STRIP];
  __shared__ __align__(16) float ytab[RS_STRIP][8];   // per destination row: 4 vertical weights, then `due` of the NEXT row (-1: none)
  const int t = threadIdx.x;
  const int x0 = blockIdx.x * K::TW, ystrip = blockIdx.y * RS_STRIP;
  const int rows = min(RS_STRIP, h - ystrip);
  const int64_t fbase = (int64_t)blockIdx.z * h * w * C;
  for (int ly = t; ly < rows; ly += K::JW) {
    float c[4];
    ys[ly] = cubic_taps(ystrip + ly + yoff, factor, c);
    *reinterpret_cast<float4*>(ytab[ly]) = make_float4(c[0], c[1], c[2], c[3]);
  }
  // source bytes of a row this strip reads: pixels clamp(rx0) .. clamp(rxl), from the 4-byte boundary below
  const int xlast = min(x0 + K::TW, w) - 1;
  const int rx0 = min(max(cubic_pos(x0 + xoff, factor) - 1, 0), w - 1), rxl = min(max(cubic_pos(xlast + xoff, factor) + 2, 0), w - 1);
  const int abase = words ? (rx0 * C) & ~3 : rx0 * C;
  const int nbytes = rxl * C + C - abase;            // <= SRW
  // this thread's column
  const bool live = x0 + t / C < w;
  const int px = min(x0 + t / C, w - 1), ch = t % C;
  float cx[4];
  int o[4];
  {
    const int sx = cubic_taps(px + xoff, factor, cx);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = min(max(sx - 1 + k, 0), w - 1) * C + ch - abase;
  }
  const int64_t pitch = (int64_t)w * C;
  uint8_t* orow = dst + fbase + ((int64_t)ystrip * w + x0) * C + t;
  const float *s0 = srcf + o[0], *s1 = srcf + o[1], *s2 = srcf + o[2], *s3 = srcf + o[3];
  const int lk = t % K::LW, lr = t / K::LW;
  const uint8_t* lbase = src + fbase + abase + 4 * lk;
  __syncthreads();
  // a destination row is due when the source row that completes its window (first source row of its 32-row group
  // = row 0) has entered
  for (int ly = t; ly < rows; ly += K::JW) {
    const int nx = ly + 1;
    ytab[ly][4] = __int_as_float((nx < rows && (nx % RS_TH) != 0) ? ys[nx] + 2 - (ys[ly - ly % RS_TH] - 1) : -1);
  }
  // the loader threads hold the words of the NEXT group of source rows in registers while the current one is marched
  constexpr int RL = K::JW / K::LW, NPF = (RS_SRH + RL - 1) / RL;
  unsigned pf[NPF];
  const bool loader = words && 4 * lk < nbytes;
  auto fetch = [&](int l0) {
    const int l1 = min(l0 + RS_TH, rows);
    const int ry0 = ys[l0] - 1, nrows = ys[l1 - 1] + 2 - ry0 + 1;
#pragma unroll
    for (int i = 0; i < NPF; ++i) {
      const int r = lr + i * RL;
      if (loader && r < nrows) pf[i] = __ldg(reinterpret_cast<const unsigned*>(lbase + (int64_t)min(max(ry0 + r, 0), h - 1) * pitch));
    }
  };
  if (words) fetch(0);
  for (int l0 = 0; l0 < rows; l0 += RS_TH) {
    const int l1 = min(l0 + RS_TH, rows);
    const int ry0 = ys[l0] - 1, nrows = ys[l1 - 1] + 2 - ry0 + 1;
    // stage the source rows ry0 .. ry0 + nrows - 1 (replicated outside the image)
    if (words) {
#pragma unroll
      for (int i = 0; i < NPF; ++i) {
        const int r = lr + i * RL;
        if (loader && r < nrows) {
          const unsigned v = pf[i];
          float4 f;   // byte k under the exponent of 2^23, minus 2^23
          f.x = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440)) - 8388608.f;
          f.y = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7441)) - 8388608.f;
          f.z = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7442)) - 8388608.f;
          f.w = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7443)) - 8388608.f;
          *reinterpret_cast<float4*>(srcf + r * K::SRW + 4 * lk) = f;
        }
      }
    } else {
      for (int r = 0; r < nrows; ++r) {
        const uint8_t* srow = src + fbase + (int64_t)min(max(ry0 + r, 0), h - 1) * w * C + abase;
        for (int q = t; q < nbytes; q += K::JW) srcf[r * K::SRW + q] = u8_to_f32(__ldg(srow + q));
      }
    }
    __syncthreads();
    if (words && l0 + RS_TH < rows) fetch(l0 + RS_TH);
    if (live) {
      // one source row enters the window per step (r is uniform: the tap addresses are pointer + constant); the
      // destination rows whose window ends at that row leave with it
      float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;
      int due = 3;                       // the first row of the group: its window is source rows 0..3
      const float* yp = ytab[l0];
#pragma unroll 4
      for (int r = 0; r < nrows; ++r) {
        h0 = h1; h1 = h2; h2 = h3;
        h3 = s0[r * K::SRW] * cx[0];
        h3 = __fmaf_rn(s1[r * K::SRW], cx[1], h3);
        h3 = __fmaf_rn(s2[r * K::SRW], cx[2], h3);
        h3 = __fmaf_rn(s3[r * K::SRW], cx[3], h3);
        while (due == r) {
          const float4 c = *reinterpret_cast<const float4*>(yp);
          float v = h0 * c.x;
          v = __fmaf_rn(h1, c.y, v);
          v = __fmaf_rn(h2, c.z, v);
          v = __fmaf_rn(h3, c.w, v);
          *orow = (uint8_t)round_u8(v);
          orow += pitch;
          due = __float_as_int(yp[4]);
          yp += 8;
        }
      }
    }
    __syncthreads();
  }
}

template <int C>
int launch_rescale(const uint8_t* src, uint8_t* dst, int n, int h, int w, double factor, int yoff, int xoff, cudaStream_t stream) {
  using K = RsCfg<C>;
  const int words = (((int64_t)w * C) % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) ? 1 : 0;
  dim3 grid((w + K::TW - 1) / K::TW, (h + RS_STRIP - 1) / RS_STRIP, n);
  rescale_cubic_kernel<C><<<grid, K::JW, 0, stream>>>(src, dst, h, w, factor, yoff, xoff, words);
  note_launch();
  return record_cuda(cudaGetLastError());
}

inline int rescaled_size(int n, double factor) { return (int)std::nearbyint((double)n * factor); }   // cv2: saturate_cast<int>(n * fx)

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_shift_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int channels, float dx, float dy, vu_stream_t stream) {
  VU_REQUIRE(src && dst && n >= 0 && h > 0 && w > 0 && h <= 32767 && w <= 32767);
  VU_REQUIRE(std::fabs((double)dx) < 1048576.0 && std::fabs((double)dy) < 1048576.0);   // also rejects NaN
  if (channels != 1 && channels != 3) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  const double b1 = -(double)dx, b2 = -(double)dy;
  const long long X0 = (long long)std::nearbyint(b1 * 1024.0) + 16;
  const int ox = (int)(X0 >> 10), fx = (int)((X0 >> 5) & 31);
  // the tile kernel needs one row offset and one row fraction for the whole image (true unless dy is so small
  // against the row index that y - dy rounds differently from row to row) and rows the TMA unit can address
  int oy = 0, fy = 0;
  shift_row(0, b2, oy, fy);
  bool uniform = true;
  for (int y = 1; y < h && uniform; ++y) {
    int sy, f;
    shift_row(y, b2, sy, f);
    uniform = (sy - y == oy) && f == fy;
  }
  const int64_t wc = (int64_t)w * channels;
  tma::EncodeTiledFn enc = tma::encode_tiled_fn();
  const bool aligned = wc % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  static const bool force_generic = std::getenv("VU_SHIFT_GENERIC") != nullptr;   // test hook
  if (!force_generic && uniform && aligned && enc && std::abs(oy) < 32767 && std::abs(ox) < 32767) {
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)wc, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)wc, (cuuint64_t)(wc * h)};
    const cuuint32_t box[3] = {SH_BOX_W, SH_BOX_H, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
      dim3 grid((unsigned)((wc + SH_TB - 1) / SH_TB), (h + SH_TR - 1) / SH_TR, n);
      if (grid.y <= 65535 && grid.z <= 65535) {
        if (channels == 1) shift_tile_kernel<1><<<grid, 256, 0, S(stream)>>>(tmap, dst, h, wc, ox * channels, oy, fx, fy);
        else shift_tile_kernel<3><<<grid, 256, 0, S(stream)>>>(tmap, dst, h, wc, ox * channels, oy, fx, fy);
        VU_RETURN_LAUNCH();
      }
    }
  }
  const int grid = grid_for((int64_t)n * h * wc, 256, 8);
  if (channels == 1) shift_generic_kernel<1><<<grid, 256, 0, S(stream)>>>(src, dst, n, h, w, ox, fx, b2);
  else shift_generic_kernel<3><<<grid, 256, 0, S(stream)>>>(src, dst, n, h, w, ox, fx, b2);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_rescale_cubic_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int channels, double factor, vu_stream_t stream) {
  VU_REQUIRE(src && dst && n >= 0 && h > 0 && w > 0);
  VU_REQUIRE(factor >= 1.0 && factor <= 16.0);
  if (channels != 1 && channels != 3) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  if (n > 65535 || (h + RS_STRIP - 1) / RS_STRIP > 65535) return VU_ERR_UNSUPPORTED;
  const int dh = rescaled_size(h, factor), dw = rescaled_size(w, factor);
  const int yoff = (dh - h) / 2, xoff = (dw - w) / 2;   // int((dh - h) / 2) of imgprocess.py:49-50
  if (channels == 1) return launch_rescale<1>(src, dst, n, h, w, factor, yoff, xoff, S(stream));
  return launch_rescale<3>(src, dst, n, h, w, factor, yoff, xoff, S(stream));
}
