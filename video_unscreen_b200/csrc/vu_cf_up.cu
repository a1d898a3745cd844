// The second full-resolution pass of the green-screen path, for frames that are an exact 2x or 4x of the working
// resolution (1080p and 4K at input_long_side 960):
//
//   alpha  = cv2.resize(alpha_lo, (W, H))                                   colorfiltering/agent.py:342 (early-out frames,
//                                                                           :303-307, are copied from the segmentation mask)
//   bgmask = is_pixel_inrange(frame, bg_colour, color_winsize)              trimap/agent.py:84 -> utils/fgfuncs.py:55-64
//   fuzzy  = (alpha > 0) & bgmask;  counts = (#fuzzy, #(alpha > 0))         trimap/agent.py:90-94
//   B      = nearest-down(alpha) >= 128                                     trimap/agent.py:52 (what the bit-logic trimap needs)
//   [FG]   bgimg[alpha < 128] = frame[alpha < 128];  fg = get_fg(frame, alpha, bgimg)      tools/unscreen/green.py:125-126
//
// in ONE pass: the working-resolution alpha is L2-resident (0.5 MB per frame), the full-resolution alpha is written
// once and never read back, fuzzy and B leave as bit planes (1/8 and 1/32 .. 1/128 byte per pixel), and the frame is
// only read where the matte is non-zero (everywhere in the FG variant, whose patched background needs it).  This
// replaces resize_up + fuzzy_count (+ get_fg) and the strided full-resolution reads of the trimap kernel: per frame
// the chunk moves about 7P bytes instead of 14P (P = pixels), see DESIGN.md section 4.
//
// A thread owns 16 destination columns of ONE row (a warp: 128 columns x 4 rows), so that every load of the kernel is
// independent of every other thread's (a first version walked 16 rows per thread to reuse the horizontal pass and
// spent its time waiting: one DRAM round trip per row, 830 GB/s); where every tap of both source rows has one value
// the interpolation is that value (most of a matte is 0 or 255).  Arithmetic = cv2's fixed-point bilinear, as in
// resize_up_int_kernel; HSV / get_fg arithmetic as in vu_composite.cu.
#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int UT = 256;
constexpr int UP_ROWS = 4;   // destination rows per warp of alpha_up_fuzzy_kernel

__device__ __forceinline__ int trunc_clamp255_(float x) { return f32_trunc_nonneg(fminf(fmaxf(x, 0.f), 255.f)); }

template <int SC>
__device__ __forceinline__ int wleft(int ph) {   // weight of the left / upper tap of phase ph, in 1/2048
  return SC == 2 ? (ph == 0 ? 512 : 1536) : (ph == 0 ? 768 : (ph == 1 ? 256 : (ph == 2 ? 1792 : 1280)));
}

__device__ __forceinline__ void block_add2(unsigned a, unsigned b, unsigned long long* dst) {
  __shared__ unsigned sh[2][UT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long sa = 0, sb = 0;
#pragma unroll
    for (int i = 0; i < UT / 32; ++i) { sa += sh[0][i]; sb += sh[1][i]; }
    if (sa) atomicAdd(dst, sa);
    if (sb) atomicAdd(dst + 1, sb);
  }
}

// 16-byte asynchronous copy global -> shared (LDGSTS): the FG variant prefetches the next row's frame bytes while it works
// on this one, without holding them in registers
__device__ __forceinline__ void cp_async16(void* smem, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gptr));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int SC, bool FG>
__global__ void __launch_bounds__(UT, FG ? 4 : 5) alpha_up_fuzzy_kernel(const uint8_t* __restrict__ alpha_lo, int th, int tw, const uint8_t* __restrict__ alt_src,
                                                            const uint8_t* __restrict__ alt_flags, const uint8_t* __restrict__ frames, int lo0, int lo1,
                                                            int lo2, int hi0, int hi1, int hi2, uint8_t* __restrict__ alpha,
                                                            uint8_t* __restrict__ fzbits, uint8_t* __restrict__ mbits,
                                                            unsigned long long* __restrict__ counts2, int bgB, int bgG, int bgR,
                                                            uint8_t* __restrict__ fg_out, uint8_t* __restrict__ bg_out) {
  __shared__ HsvTab tab;
  __shared__ float ktab[FG ? 256 : 1];   // 1 - alpha/255. for every alpha byte
  __shared__ uint4 stage[FG ? 2 : 1][FG ? 3 : 1][FG ? UT : 1];   // FG: the frame bytes of this row and the next, [stage][16-byte piece][thread]
  if (FG) hsv_tab_init(tab);
  else hsv_tab_init_fwd(tab);
  if (FG)
    for (int a = threadIdx.x; a < 256; a += UT) ktab[a] = __fsub_rn(1.f, __fdiv_rn((float)a, 255.f));
  __syncthreads();
  constexpr int NS = 16 / SC;   // source columns under the 16 destination columns
  const int n = blockIdx.z;
  const int h = SC * th, w = SC * tw;
  // a warp = 8 column groups x 4 rows (128 x 4 pixels), not 512 pixels of one row: the matte-dependent branches below
  // diverge per warp, and a compact footprint leaves far fewer warps straddling the matte's outline
  const int lane = threadIdx.x & 31;
  const int tx = blockIdx.x * 8 + (lane & 7);
  const int x0 = 16 * tx;
  const bool alt = alt_flags && alt_flags[n] != 0;
  const uint8_t* s = alpha_lo + (int64_t)n * th * tw;
  const int c0 = NS * tx;
  int h0v = 0, s0v = 0, v0v = 0;
  if (FG) bgr2hsv_px(bgB, bgG, bgR, tab, h0v, s0v, v0v);
  unsigned cnt_f = 0, cnt_p = 0;
  auto row_of = [&](int it) { return ((blockIdx.y * UP_ROWS + it) * (UT / 32) + (threadIdx.x >> 5)) * 4 + (lane >> 3); };
  // FG: the patched background needs the frame everywhere, so its 48 bytes per thread and row are fetched one row ahead with
  // asynchronous copies (a thread only ever reads back its own pieces: no barrier).  The kernel waited on these loads with
  // half of its issue slots empty when every row fetched its own.
  auto prefetch = [&](int it) {
    if (!FG) return;
    const int y = row_of(it);
    if (it < UP_ROWS && x0 < w && y < h) {
      const uint8_t* src = frames + (((int64_t)n * h + y) * w + x0) * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) cp_async16(&stage[it & 1][k][threadIdx.x], src + 16 * k);
    }
    cp_async_commit();
  };
  prefetch(0);
  // a warp takes one destination row at a time (UP_ROWS of them, interleaved with the CTA's other warps: the table set-up
  // above is paid once per UP_ROWS * 8 rows); the rows of a thread do not depend on each other
#pragma unroll 1
  for (int it = 0; it < UP_ROWS; ++it) {
  const int y = row_of(it);
  if (y - (lane >> 3) >= h) break;   // warp-uniform
  prefetch(it + 1);
  const bool act = x0 < w && y < h;
  const int r = y / SC, ph = y % SC;
  unsigned aw[4] = {0u, 0u, 0u, 0u};
  const int64_t fo = (((int64_t)n * h + y) * w + x0) * 3;
  uint4 fv[3];
  if (act) {
    if (alt) {
      const uint4 v = ldg_stream16(alt_src + ((int64_t)n * h + y) * w + x0);
      aw[0] = v.x; aw[1] = v.y; aw[2] = v.z; aw[3] = v.w;
    } else {
      // taps: source columns c0-1 .. c0+NS (replicated at the borders) of the two source rows, NS bytes as one word (pair)
      const int ia = ph < SC / 2 ? max(r - 1, 0) : r, ib = ph < SC / 2 ? r : min(r + 1, th - 1);
      unsigned wv[2][2], ev[2][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint8_t* row = s + (int64_t)(j ? ib : ia) * tw;
        if (SC == 2) {
          const uint2 v = __ldg(reinterpret_cast<const uint2*>(row + c0));
          wv[j][0] = v.x; wv[j][1] = v.y;
        } else {
          wv[j][0] = wv[j][1] = __ldg(reinterpret_cast<const unsigned*>(row + c0));
        }
        ev[j][0] = __ldg(row + max(c0 - 1, 0));
        ev[j][1] = __ldg(row + min(c0 + NS, tw - 1));
      }
      // one value everywhere (most of a matte): a handful of word compares, no unpacking
      const unsigned u = ev[0][0], splat = u * 0x01010101u;
      const bool same = (wv[0][0] == splat) & (wv[0][1] == splat) & (wv[1][0] == splat) & (wv[1][1] == splat) & (ev[0][1] == u) & (ev[1][0] == u) &
                        (ev[1][1] == u);
      if (same) {
        // every tap has one value and the weights sum to 2048 on both axes: the two halves lose less than 2 of 4v + 2 together
        aw[0] = aw[1] = aw[2] = aw[3] = splat;
      } else {
        int t[2][NS + 2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          t[j][0] = (int)ev[j][0];
          t[j][NS + 1] = (int)ev[j][1];
#pragma unroll
          for (int k = 0; k < NS; ++k) t[j][1 + k] = (int)((wv[j][k >> 2] >> (8 * (k & 3))) & 255u);
        }
        const int b0 = wleft<SC>(ph), b1 = 2048 - b0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int c = k / SC, p = k % SC;
          const int il = p < SC / 2 ? c : c + 1;   // index into t of the left tap (t[c+1] is column c0 + c)
          const int Ra = (t[0][il] * wleft<SC>(p) + t[0][il + 1] * (2048 - wleft<SC>(p))) >> 4;
          const int Rb = (t[1][il] * wleft<SC>(p) + t[1][il + 1] * (2048 - wleft<SC>(p))) >> 4;
          const unsigned v = (unsigned)((((b0 * Ra) >> 16) + ((b1 * Rb) >> 16) + 2) >> 2);   // <= 255 by construction
          aw[k >> 2] |= v << (8 * (k & 3));
        }
      }
    }
    stg_stream16(alpha + ((int64_t)n * h + y) * w + x0, make_uint4(aw[0], aw[1], aw[2], aw[3]));
  }
  // ---- B bits of the trimap source: pixels (SC*r, SC*c) of rows with y % SC == 0 (the shuffle runs on every lane) ----
  {
    unsigned mb = 0;
    if ((aw[0] | aw[1] | aw[2] | aw[3]) & 0x80808080u) {   // some pixel >= 128 (else all bits are zero)
#pragma unroll
      for (int c = 0; c < NS; ++c) {
        const int k = SC * c;
        mb |= ((aw[k >> 2] >> (8 * (k & 3) + 7)) & 1u) << c;
      }
    }
    if (SC == 2) {
      if (act && ph == 0) mbits[((int64_t)n * th + r) * (tw >> 3) + tx] = (uint8_t)mb;
    } else {
      const unsigned hi = __shfl_down_sync(0xffffffffu, mb, 1);   // the neighbour in x: lanes 2j, 2j + 1 share a row
      if (act && ph == 0 && !(lane & 1)) mbits[((int64_t)n * th + r) * (tw >> 3) + (tx >> 1)] = (uint8_t)(mb | (hi << 4));
    }
  }
  // ---- fuzzy bits (+ fg / patched bg) ----
  if (act) {
    const unsigned any = aw[0] | aw[1] | aw[2] | aw[3];
    unsigned fz = 0;
    if (FG) {
      cp_async_wait<1>();   // everything but the row just requested has landed
#pragma unroll
      for (int k = 0; k < 3; ++k) fv[k] = stage[it & 1][k][threadIdx.x];
    } else if (any) {
#pragma unroll
      for (int k = 0; k < 3; ++k) fv[k] = ldg_stream16(frames + fo + 16 * k);
    }
    if (FG && !any) {
      // alpha == 0 everywhere: patched background = the frame, fg = HSV2BGR(hsv - 1.0 * hsv) = black
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        stg_stream16(fg_out + fo + 16 * k, make_uint4(0u, 0u, 0u, 0u));
        stg_stream16(bg_out + fo + 16 * k, fv[k]);
      }
    } else if (any) {
      const unsigned* fw = reinterpret_cast<const unsigned*>(fv);
      uint4 ov[3], bv[3];
      unsigned* ow = reinterpret_cast<unsigned*>(ov);
      unsigned* bw = reinterpret_cast<unsigned*>(bv);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (aw[g] == 0u) {
          if (FG) {
            ow[3 * g] = ow[3 * g + 1] = ow[3 * g + 2] = 0u;
            bw[3 * g] = fw[3 * g]; bw[3 * g + 1] = fw[3 * g + 1]; bw[3 * g + 2] = fw[3 * g + 2];
          }
          continue;
        }
        int c[12], o[12], q[12];
        unpack12(fw[3 * g], fw[3 * g + 1], fw[3 * g + 2], c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int a = (int)((aw[g] >> (8 * i)) & 255u);
          int ih = 0, is, iv, id;
          bgr2hsv_sv(c[3 * i], c[3 * i + 1], c[3 * i + 2], tab, is, iv, id);
          bool in = (is >= lo1) & (is <= hi1) & (iv >= lo2) & (iv <= hi2);
          if (FG || in) {   // the hue is the expensive third: a pixel whose saturation or value is out of range needs none
            ih = bgr2hsv_hue(c[3 * i], c[3 * i + 1], c[3 * i + 2], iv, id, tab);
            in = in & (ih >= lo0) & (ih <= hi0);
          }
          const bool pos = a != 0;
          cnt_p += pos;
          fz |= (unsigned)(pos && in) << (4 * g + i);
          if (FG) {
            const bool patch = a < 128;   // green.py:125
            q[3 * i] = patch ? c[3 * i] : bgB;
            q[3 * i + 1] = patch ? c[3 * i + 1] : bgG;
            q[3 * i + 2] = patch ? c[3 * i + 2] : bgR;
            const int bh = patch ? ih : h0v, bs = patch ? is : s0v, bvv = patch ? iv : v0v;
            const float k = ktab[a];
            const int fh = trunc_clamp255_(__fsub_rn(u8_to_f32(ih), __fmul_rn(k, u8_to_f32(bh))));
            const int fs = trunc_clamp255_(__fsub_rn(u8_to_f32(is), __fmul_rn(k, u8_to_f32(bs))));
            const int fv2 = trunc_clamp255_(__fsub_rn(u8_to_f32(iv), __fmul_rn(k, u8_to_f32(bvv))));
            hsv2bgr_px(fh, fs, fv2, tab, o[3 * i], o[3 * i + 1], o[3 * i + 2]);
          }
        }
        if (FG) {
          pack12(o, ow[3 * g], ow[3 * g + 1], ow[3 * g + 2]);
          pack12(q, bw[3 * g], bw[3 * g + 1], bw[3 * g + 2]);
        }
      }
      if (FG) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          stg_stream16(fg_out + fo + 16 * k, ov[k]);
          stg_stream16(bg_out + fo + 16 * k, bv[k]);
        }
      }
    }
    cnt_f += __popc(fz);
    *reinterpret_cast<unsigned short*>(fzbits + (((int64_t)n * h + y) * w + x0) / 8) = (unsigned short)fz;
  }
  }
  if (FG) cp_async_wait<0>();
  block_add2(cnt_f, cnt_p, counts2 + 2 * n);
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_cf_alpha_up_fuzzy(const uint8_t* alpha_lo, int n, int th, int tw, int h, int w, const uint8_t* alt_src, const uint8_t* alt_flags,
                                    const uint8_t* frames, const int32_t lo[3], const int32_t hi[3], uint8_t* alpha, uint8_t* fuzzy_bits,
                                    uint8_t* mask_bits, uint64_t* counts2, const uint8_t* bg_bgr, uint8_t* fg_out, uint8_t* bg_out,
                                    vu_stream_t stream) {
  VU_REQUIRE(alpha_lo && frames && lo && hi && alpha && fuzzy_bits && mask_bits && counts2 && n >= 0 && th > 0 && tw > 0);
  VU_REQUIRE((alt_src == nullptr) == (alt_flags == nullptr));
  VU_REQUIRE((fg_out == nullptr) == (bg_out == nullptr) && (fg_out == nullptr) == (bg_bgr == nullptr));
  const int sc = (w == 2 * tw && h == 2 * th) ? 2 : ((w == 4 * tw && h == 4 * th) ? 4 : 0);
  if (sc == 0 || tw % 16 != 0 || n > 65535) return VU_ERR_UNSUPPORTED;
  const void* p16[] = {frames, alpha, alt_src, fg_out, bg_out};
  for (const void* p : p16)
    if (reinterpret_cast<uintptr_t>(p) & 15) return VU_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(alpha_lo) & 7) || (reinterpret_cast<uintptr_t>(fuzzy_bits) & 1)) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  int e = record_cuda(cudaMemsetAsync(counts2, 0, sizeof(uint64_t) * 2 * n, S(stream)));
  if (e) return e;
  dim3 g((w / 16 + 7) / 8, (h + UP_ROWS * (UT / 32) * 4 - 1) / (UP_ROWS * (UT / 32) * 4), n);
  auto* c2 = reinterpret_cast<unsigned long long*>(counts2);
  const int B = bg_bgr ? bg_bgr[0] : 0, G = bg_bgr ? bg_bgr[1] : 0, R = bg_bgr ? bg_bgr[2] : 0;
#define VU_CALL(SCV, FGV)                                                                                                                       \
  alpha_up_fuzzy_kernel<SCV, FGV><<<g, UT, 0, S(stream)>>>(alpha_lo, th, tw, alt_src, alt_flags, frames, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], \
                                                           alpha, fuzzy_bits, mask_bits, c2, B, G, R, fg_out, bg_out)
  if (sc == 2) {
    if (fg_out) VU_CALL(2, true);
    else VU_CALL(2, false);
  } else {
    if (fg_out) VU_CALL(4, true);
    else VU_CALL(4, false);
  }
#undef VU_CALL
  VU_RETURN_LAUNCH();
}
