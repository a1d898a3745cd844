// Exact temporal median of up to 608 uint8 frames, data resident in registers
// (new specification, no reference code).
//
// The median of x_1..x_n minimises the convex S(m) = sum_f |x_f - m|.
// VABSDIFF4 with accumulate evaluates S for 4 frames of one element per
// instruction, which makes "one probe of S" cost n/4 ALU-pipe instructions per
// element, and the ALU pipe (64 lanes/clk/SM) is what bounds this kernel.  The
// design goal is therefore the smallest number of probes:
//
//   stage A  the lower median of 16 frames spread over the whole clip is the
//            estimate (bisection on the slope of S, 8 steps x 8 instructions;
//            every frame-part does it for its share of the lane's elements);
//   stage B  Fibonacci search of the 20 values around the estimate on ALL
//            frames: 6 probes.  The bracket ends keep their S values, so the
//            search ends knowing S at both neighbours of the minimiser: a
//            strict minimum IS the median (odd n) / both middle order
//            statistics (even n), with no further probe;
//   even n   a neighbour with the same S belongs to the plateau [x_lo, x_hi]
//            of minimisers: its end is walked one probe at a time (one probe
//            for the usual plateau of two adjacent values), binary search for
//            long plateaus;
//   fallback if the minimiser of any element of the warp sits on the edge of
//            its window (estimate off by more than 8), the whole warp repeats
//            stage B over 0..255 (12 probes).  Always exact; only the speed
//            depends on the data.
//
// Layout.  A warp owns SEG = 128/SPLIT consecutive bytes of every frame.  The
// frames are dealt to SPLIT lane groups ("parts"): part p holds frames
// f = SPLIT*i + p.  Lane l of a part holds its 4 bytes of all those frames: one
// coalesced LDG.32 per frame, every load of the warp in flight at once.  Slot
// (g, r) of a lane (register 4g+r) holds frame index i = r*Q + g with
// Q = ceil(ceil(n/SPLIT)/4), so that after the 4x4 byte transposes register
// 4g+j holds four frames of element j a quarter of the clip apart, and any few
// groups form a sample spread over the whole clip.  Unused slots are padded
// with 0 in one half of the parts and 255 in the other, which leaves the middle
// order statistics where they are; when the two pad counts differ by one (odd
// n) the surplus pad is taken out of S arithmetically (S += delta * m).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include "vu_tma.cuh"

namespace vu {
namespace msad {

__device__ __forceinline__ unsigned sad_acc(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// 4x4 byte transposes: d[4g+j] <- byte j of the four registers of group g
// (tried: the same transposition with IDP.4A one-hot byte extraction + IMAD shifts to get it off the ALU pipe:
// 28 instead of 8 instructions per group, and the kernel got slower with every group moved over)
template <int G>
__device__ __forceinline__ void transpose_groups(unsigned (&d)[4 * G]) {
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const unsigned a = d[4 * g], b = d[4 * g + 1], c = d[4 * g + 2], e = d[4 * g + 3];
    const unsigned ab_lo = __byte_perm(a, b, 0x5140), ab_hi = __byte_perm(a, b, 0x7362);
    const unsigned ce_lo = __byte_perm(c, e, 0x5140), ce_hi = __byte_perm(c, e, 0x7362);
    d[4 * g] = __byte_perm(ab_lo, ce_lo, 0x5410);
    d[4 * g + 1] = __byte_perm(ab_lo, ce_lo, 0x7632);
    d[4 * g + 2] = __byte_perm(ab_hi, ce_hi, 0x5410);
    d[4 * g + 3] = __byte_perm(ab_hi, ce_hi, 0x7632);
  }
}

// ld.global.nc.u32 of [base + g * stride]: the 64-bit address is formed by ONE IMAD.WIDE.U32 with g as an immediate
// (FMA pipe), which keeps the address arithmetic of the ~150 loads per lane off the ALU pipe that bounds the search
template <int GI>
__device__ __forceinline__ unsigned ldg_strided(const uint8_t* base, unsigned stride) {
  unsigned long long addr;
  unsigned v;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(stride), "n"(GI), "l"(base));
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(addr));
  return v;
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

constexpr unsigned BIG = 0x7fffffffu;
constexpr unsigned FULL = 0xffffffffu;

// S(m[j]) over all frames of the warp's segment for the 4 elements of the lane.
// RANGE: probes outside 0..255 are allowed and evaluate to BIG.
template <int SPLIT, int G, bool RANGE>
__device__ __forceinline__ void eval_s(const unsigned (&d)[4 * G], const int (&m)[4], int delta, unsigned (&S)[4]) {
  constexpr int LPS = 32 / SPLIT;
  unsigned q[4], s[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q[j] = (unsigned)(RANGE ? min(max(m[j], 0), 255) : m[j]) * 0x01010101u;
    s[j] = 0;
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = sad_acc(d[4 * g + j], q[j], s[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int o = LPS; o < 32; o <<= 1) s[j] += __shfl_xor_sync(FULL, s[j], o);
    s[j] += (unsigned)(delta * m[j]);
    S[j] = (RANGE && (m[j] < 0 || m[j] > 255)) ? BIG : s[j];
  }
}

// Fibonacci search of the interior points a+1 .. a+F(K0)-1 of the bracket
// (a, a+F(K0)).  On return r is the minimiser among them, smin = S(r), and
// sl / sr are S(r-1) / S(r+1), or BIG where that neighbour is the bracket's
// original end (never evaluated).  2 + (K0 - 4) evaluations of S.
template <int SPLIT, int G, bool RANGE>
__device__ __forceinline__ void fib_search(const unsigned (&d)[4 * G], int delta, int k0, int fa, int fb, int fc, int (&a)[4], int (&r)[4],
                                           unsigned (&smin)[4], unsigned (&sl)[4], unsigned (&sr)[4]) {
  unsigned S1[4], S2[4];
  {
    int i1[4], i2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      i1[j] = a[j] + fb;
      i2[j] = a[j] + fa;
      sl[j] = sr[j] = BIG;
    }
    eval_s<SPLIT, G, RANGE>(d, i1, delta, S1);
    eval_s<SPLIT, G, RANGE>(d, i2, delta, S2);
  }
#pragma unroll 1
  for (int k = k0; k >= 5; --k) {
    int nidx[4];
    bool left[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      left[j] = S1[j] <= S2[j];
      if (left[j]) {
        sr[j] = S2[j];
        S2[j] = S1[j];
        nidx[j] = a[j] + fc;
      } else {
        sl[j] = S1[j];
        a[j] += fb;
        S1[j] = S2[j];
        nidx[j] = a[j] + fb;
      }
    }
    unsigned Sn[4];
    eval_s<SPLIT, G, RANGE>(d, nidx, delta, Sn);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (left[j]) S1[j] = Sn[j];
      else S2[j] = Sn[j];
    }
    const int t = fb - fc;
    fa = fb;
    fb = fc;
    fc = t;
  }
  // k == 4: bracket (a, a+3), S1 = S(a+1), S2 = S(a+2)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (S1[j] <= S2[j]) {
      r[j] = a[j] + 1;
      smin[j] = S1[j];
      sr[j] = S2[j];
    } else {
      r[j] = a[j] + 2;
      smin[j] = S2[j];
      sl[j] = S1[j];
    }
  }
}

// SS = stride of the 4 sample groups of stage A (0: no stage A, straight to the full search)
template <int SPLIT, int G, int SS>
__device__ __forceinline__ unsigned sad_median(const unsigned (&d)[4 * G], int n, int delta) {
  int a[4], r[4];
  unsigned smin[4], sl[4], sr[4];
  bool full = (SS == 0);
  if (SS > 0) {
    // ---- stage A: every part estimates 4/SPLIT of the lane's elements as the lower median of 16 of ITS frames
    // (bisection on the slope of S), then the parts swap their estimates ----
    constexpr int EPP = 4 / SPLIT;   // elements per part
    const int part = (threadIdx.x & 31) / (32 / SPLIT);
    unsigned smp[EPP][4];
#pragma unroll
    for (int e = 0; e < EPP; ++e)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        unsigned v = d[4 * (t * SS) + e];
#pragma unroll
        for (int p = 1; p < SPLIT; ++p) v = (part == p) ? d[4 * (t * SS) + p * EPP + e] : v;
        smp[e][t] = v;
      }
    int est[EPP];
#pragma unroll
    for (int e = 0; e < EPP; ++e) est[e] = 0;
#pragma unroll 1
    for (int bit = 128; bit > 0; bit >>= 1) {
#pragma unroll
      for (int e = 0; e < EPP; ++e) {
        const unsigned q1 = (unsigned)(est[e] | bit) * 0x01010101u, q0 = q1 - 0x01010101u;
        unsigned s0 = 0, s1 = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          s0 = sad_acc(smp[e][t], q0, s0);
          s1 = sad_acc(smp[e][t], q1, s1);
        }
        if (s1 < s0) est[e] |= bit;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = __shfl_sync(FULL, est[j % EPP], (j / EPP) * (32 / SPLIT) + (threadIdx.x & (32 / SPLIT - 1)));
      a[j] = min(max(c - 10, -1), 235);   // interior a+1 .. a+20 within 0..255
    }
    // ---- stage B: the 20 values around the estimate ----
    fib_search<SPLIT, G, false>(d, delta, 8, 13, 8, 5, a, r, smin, sl, sr);
    bool fail = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) fail |= (sl[j] == BIG && r[j] > 0) || (sr[j] == BIG && r[j] < 255);
    full = __any_sync(FULL, fail);
  }
  if (full) {
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = -1;
    fib_search<SPLIT, G, true>(d, delta, 14, 233, 144, 89, a, r, smin, sl, sr);   // interior -1+1 .. 375: covers 0..255
  }
  int lo[4], hi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) lo[j] = hi[j] = r[j];
  if (!(n & 1)) {
    // even n: minimisers form the plateau [x_lo, x_hi]
    bool openL[4], openR[4];
    bool any = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      openL[j] = openR[j] = false;
      if (sl[j] == smin[j]) { lo[j] = r[j] - 1; openL[j] = lo[j] > 0; }
      if (sr[j] == smin[j]) { hi[j] = r[j] + 1; openR[j] = hi[j] < 255; }
      any |= openL[j] | openR[j];
    }
    // walk the open ends, one probe per lane and round
#pragma unroll 1
    for (int it = 0; it < 4 && __any_sync(FULL, any); ++it) {
      int idx[4];
      unsigned S[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) idx[j] = openL[j] ? lo[j] - 1 : (openR[j] ? hi[j] + 1 : r[j]);
      eval_s<SPLIT, G, false>(d, idx, delta, S);
      any = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (openL[j]) {
          if (S[j] == smin[j]) { --lo[j]; openL[j] = lo[j] > 0; }
          else openL[j] = false;
        } else if (openR[j]) {
          if (S[j] == smin[j]) { ++hi[j]; openR[j] = hi[j] < 255; }
          else openR[j] = false;
        }
        any |= openL[j] | openR[j];
      }
    }
    // long plateaus (e.g. two-valued data): binary search for the ends (S == smin is monotone on either side)
    if (__any_sync(FULL, any)) {
      int L[4], R[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { R[j] = lo[j]; L[j] = openL[j] ? 0 : lo[j]; }
      while (__any_sync(FULL, (L[0] < R[0]) | (L[1] < R[1]) | (L[2] < R[2]) | (L[3] < R[3]))) {
        int idx[4];
        unsigned S[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = (L[j] + R[j]) >> 1;
        eval_s<SPLIT, G, false>(d, idx, delta, S);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (L[j] < R[j]) {
            const int mid = (L[j] + R[j]) >> 1;
            if (S[j] == smin[j]) R[j] = mid;
            else L[j] = mid + 1;
          }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { lo[j] = R[j]; L[j] = hi[j]; R[j] = openR[j] ? 255 : hi[j]; }
      while (__any_sync(FULL, (L[0] < R[0]) | (L[1] < R[1]) | (L[2] < R[2]) | (L[3] < R[3]))) {
        int idx[4];
        unsigned S[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = (L[j] + R[j] + 1) >> 1;
        eval_s<SPLIT, G, false>(d, idx, delta, S);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (L[j] < R[j]) {
            const int mid = (L[j] + R[j] + 1) >> 1;
            if (S[j] == smin[j]) L[j] = mid;
            else R[j] = mid - 1;
          }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) hi[j] = L[j];
    }
  }
  unsigned res = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) res |= (unsigned)((lo[j] + hi[j]) >> 1) << (8 * j);
  return res;
}

// pads: 0 in parts {0} (SPLIT 2) / {0,3} (SPLIT 4), 255 in the others: the two pad counts differ by at most one
template <int SPLIT>
__device__ __forceinline__ bool pad_high(int part) { return SPLIT == 2 ? part == 1 : (part == 1 || part == 2); }

// GFULL: groups whose rows 0..2 hold real frames for every n the variant is dispatched for (no bounds test on those loads).
// SPLIT * m must fit 32 bits.
template <int SPLIT, int G, int GFULL, int SS, int CTAS>
__global__ void __launch_bounds__(128, CTAS) median_sad_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, int n, long long m,
                                                               int nseg, unsigned zero) {
  constexpr int LPS = 32 / SPLIT;  // lanes per frame-part
  constexpr int SEG = LPS * 4;     // bytes of a frame one warp owns
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = lane / LPS, li = lane % LPS;
  const int seg = blockIdx.x * 4 + warp;
  if (seg >= nseg) return;
  const int cnt0 = (n + SPLIT - 1) / SPLIT;          // frames of part 0 (the largest part)
  const int cnt = (n - part + SPLIT - 1) / SPLIT;    // frames of this part
  const int Q = (cnt0 + 3) >> 2;
  const unsigned pad = pad_high<SPLIT>(part) ? 0xFFFFFFFFu : 0u;
  // surplus of 255-pads over 0-pads (-1, 0 or +1), taken out of S arithmetically
  int n0 = 0, n255 = 0;
#pragma unroll
  for (int p = 0; p < SPLIT; ++p) {
    const int c = 4 * G - (n - p + SPLIT - 1) / SPLIT;
    if (pad_high<SPLIT>(p)) n255 += c;
    else n0 += c;
  }
  const int delta = n255 - n0;
  // frame i of this part starts (SPLIT*i + part) * m bytes into the clip; slot (g, r) holds i = r*Q + g
  // (`zero` is 0: it gives every row its own copy of the stride, or the compiler shares g * stride between the
  // rows and goes back to 64-bit adds on the ALU pipe)
  unsigned gstride[4];
  const uint8_t* rowbase[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    rowbase[r] = frames + (long long)seg * SEG + li * 4 + ((long long)SPLIT * (r * Q) + part) * m;
    gstride[r] = (unsigned)(SPLIT * m) + (unsigned)r * zero;
  }
  unsigned d[4 * G];
  static_for<0, G>([&](auto gi) {
    constexpr int g = decltype(gi)::value;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const bool ok = (r < 3 && g < GFULL) ? true : (g < Q && r * Q + g < cnt);
      d[4 * g + r] = pad;
      if (ok) d[4 * g + r] = ldg_strided<g>(rowbase[r], gstride[r]);
    }
  });
  transpose_groups<G>(d);
  const unsigned res = sad_median<SPLIT, G, SS>(d, n, delta);
  if (part == 0) reinterpret_cast<unsigned*>(out + (long long)seg * SEG)[li] = res;
}


// ---- persistent tile kernel: TMA fetches the next tile while the current one is searched ---------------------------
//
// Measured on the B200 (tools/ldgsts_ubench.cu, fetch only): "one row per frame" streams at 2.8 TB/s when a warp
// fetches 64-byte rows on its own, at 5.9 TB/s with 128-byte rows and at 7.2 TB/s with 512-byte rows: requests must
// cover whole 128-byte lines, and the direct kernel above (64 bytes per frame and warp) cannot get there.  Here a CTA
// of WARPS warps owns tiles of WARPS adjacent segments (8 warps, SPLIT 2: 512 bytes of every frame).  The clip is
// described to the TMA unit as a 2-D tensor [n frames][m bytes], and SPLIT boxes of {TILE bytes x 4G frames}
// (cp.async.bulk.tensor.2d, one elected thread, completion counted in bytes on an mbarrier) bring the tile into
// shared memory as rows of TILE bytes: no address arithmetic and no load instructions in the warps.  Frames past the end of the clip are zero-filled by the unit: pads are all 0 here and leave S through
// delta = -(number of pads).  The warps are only loosely coupled: a warp waits for the tile on the mbarrier, moves its
// segment to registers and transposes, bumps a counter and goes searching; the warp that bumps it last knows the
// buffer is free and launches the fetch of the next tile, which then lands while everybody searches.
using tma::mbar_expect_tx;
using tma::mbar_init;
using tma::mbar_wait;
__device__ __forceinline__ void tma_load_2d(unsigned dst, const void* tmap, int c0, int c1, unsigned mbar) { tma::load_2d(dst, tmap, c0, c1, mbar); }

// GQ: groups g < GQ are below Q = ceil(ceil(n/SPLIT)/4) for every n the variant is dispatched for
template <int SPLIT, int G, int GQ, int SS, int WARPS, int CTAS>
__global__ void __launch_bounds__(WARPS * 32, CTAS) median_sad_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ out, int n,
                                                                          int ntiles) {
  extern __shared__ __align__(128) uint8_t smem_tma[];
  constexpr int LPS = 32 / SPLIT;
  constexpr int SEG = LPS * 4;
  constexpr int TILE = WARPS * SEG;          // bytes of a frame one CTA owns = pitch of the rows in shared memory
  constexpr int ROWS = SPLIT * 4 * G;        // rows of the buffer (frames, the last ones past the clip)
  constexpr unsigned BOX_BYTES = 4 * G * TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = lane / LPS, li = lane % LPS;
  const int cnt0 = (n + SPLIT - 1) / SPLIT;
  const int Q = (cnt0 + 3) >> 2;
  const int delta = -(ROWS - n);
  uint64_t* mbar_p = reinterpret_cast<uint64_t*>(smem_tma + ROWS * TILE);
  unsigned* count_p = reinterpret_cast<unsigned*>(mbar_p + 1);
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(mbar_p);
  const unsigned sdata = (unsigned)__cvta_generic_to_shared(smem_tma);
  auto fetch = [&](int tile) {   // one thread
    mbar_expect_tx(mbar, SPLIT * BOX_BYTES);
#pragma unroll
    for (int b = 0; b < SPLIT; ++b) tma_load_2d(sdata + b * BOX_BYTES, &tmap, tile * (TILE > 256 ? TILE / 4 : TILE), b * 4 * G, mbar);
  };
  if (threadIdx.x == 0) {
    mbar_init(mbar, 1);
    *count_p = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < ntiles) fetch(tile);
  // slot (g, r) of this lane: frame SPLIT*(r*Q + g) + part, row pitch TILE
  const uint8_t* rowaddr[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) rowaddr[r] = smem_tma + (SPLIT * (r * Q) + part) * TILE + warp * SEG + li * 4;
  unsigned parity = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    mbar_wait(mbar, parity);
    parity ^= 1;
    unsigned d[4 * G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        d[4 * g + r] = 0;
        if (g < GQ || g < Q) d[4 * g + r] = *reinterpret_cast<const unsigned*>(rowaddr[r] + g * (SPLIT * TILE));
      }
    }
    transpose_groups<G>(d);   // consumes every load
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const unsigned old = atomicAdd(count_p, 1u);
      if (old % WARPS == WARPS - 1 && tile + (int)gridDim.x < ntiles) {
        __threadfence_block();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fetch(tile + gridDim.x);
      }
    }
    const unsigned res = sad_median<SPLIT, G, SS>(d, n, delta);
    if (part == 0) reinterpret_cast<unsigned*>(out + ((long long)tile * WARPS + warp) * SEG)[li] = res;
  }
}

// ---- n > 608: estimate on a subsample, then ONE streaming pass that proves it ---------------------------------------
//
// More frames than the registers of a warp can hold.  Pass 1 is the tile kernel above on every s-th frame (a tensor
// map with row pitch s*m: at most 304 frames), which leaves the subsample's median in `out`.  Pass 2 (this kernel)
// streams ALL frames once, TMA chunk by chunk through a ring of shared-memory stages, and accumulates S(m) at the
// eight values m = e-4 .. e+3 around the estimate e of every element: 8 VABSDIFF4 + 2 PRMT per input word, just
// under what the ALU pipe can do at the HBM rate.  The minimiser of a convex function that is strictly inside the
// probed range is the global one, so a strict interior minimum is the median (odd n), and a plateau that ends
// inside the range gives both middle order statistics (even n).  Elements whose minimum touches the edge of the
// range are not decided here: the segment is flagged and the histogram kernel redoes the flagged segments.
// Frames are dealt to the two half-warps alternately; rows past the end of the clip are zero-filled by the TMA and
// leave S arithmetically (S -= pads * m).
constexpr int RF_WARPS = 8, RF_ROWS = 64, RF_STAGES = 4, RF_PROBES = 8;
constexpr int RF_TILE = RF_WARPS * 64;
constexpr int RF_STAGE_BYTES = RF_ROWS * RF_TILE;
constexpr int RF_SMEM = RF_STAGES * RF_STAGE_BYTES + RF_STAGES * 16;

__global__ void __launch_bounds__(RF_WARPS * 32, 1) median_refine_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ out,
                                                                         uint8_t* __restrict__ flags, int n, int ntiles, int nchunks) {
  extern __shared__ __align__(128) uint8_t smem_rf[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = lane >> 4, li = lane & 15;
  uint64_t* mbar_p = reinterpret_cast<uint64_t*>(smem_rf + RF_STAGES * RF_STAGE_BYTES);
  unsigned* count_p = reinterpret_cast<unsigned*>(mbar_p + RF_STAGES);
  const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(mbar_p);
  const unsigned sdata = (unsigned)__cvta_generic_to_shared(smem_rf);
  const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int J = my_tiles * nchunks;   // chunks this CTA streams, in order
  auto issue = [&](int j) {           // one thread
    const int tile = blockIdx.x + (j / nchunks) * gridDim.x, chunk = j % nchunks, st = j % RF_STAGES;
    mbar_expect_tx(mbar0 + 8 * st, RF_STAGE_BYTES);
    tma_load_2d(sdata + st * RF_STAGE_BYTES, &tmap, tile * (RF_TILE / 4), chunk * RF_ROWS, mbar0 + 8 * st);
  };
  if (threadIdx.x == 0) {
    for (int st = 0; st < RF_STAGES; ++st) {
      mbar_init(mbar0 + 8 * st, 1);
      count_p[st] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int j = 0; j < RF_STAGES && j < J; ++j) issue(j);
  const int delta = -(nchunks * RF_ROWS - n);   // zero rows past the clip
  unsigned acc[RF_PROBES][4], q[RF_PROBES][4];
  int a0[4];
  const uint8_t* lane_row = smem_rf + part * RF_TILE + warp * 64 + li * 4;   // rows 2i + part of a stage
  for (int j = 0; j < J; ++j) {
    const int tile = blockIdx.x + (j / nchunks) * gridDim.x, chunk = j % nchunks, st = j % RF_STAGES;
    const int64_t seg = (int64_t)tile * RF_WARPS + warp;
    if (chunk == 0) {
      const unsigned e = *reinterpret_cast<const unsigned*>(out + seg * 64 + li * 4);   // pass 1's estimates
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a0[k] = min(max((int)((e >> (8 * k)) & 255u) - 4, 0), 256 - RF_PROBES);
#pragma unroll
        for (int p = 0; p < RF_PROBES; ++p) {
          q[p][k] = (unsigned)(a0[k] + p) * 0x01010101u;
          acc[p][k] = 0u;
        }
      }
    }
    mbar_wait(mbar0 + 8 * st, (unsigned)((j / RF_STAGES) & 1));
    const uint8_t* base = lane_row + st * RF_STAGE_BYTES;
#pragma unroll
    for (int h = 0; h < 2; ++h) {   // 2 x 16 rows of this part: 16 words, 4 groups
      unsigned d[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) d[i] = *reinterpret_cast<const unsigned*>(base + (2 * (16 * h + i)) * RF_TILE);
      transpose_groups<4>(d);
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int p = 0; p < RF_PROBES; ++p)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[p][k] = sad_acc(d[4 * g + k], q[p][k], acc[p][k]);
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const unsigned old = atomicAdd(count_p + st, 1u);
      if (old % RF_WARPS == RF_WARPS - 1 && j + RF_STAGES < J) {
        __threadfence_block();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(j + RF_STAGES);
      }
    }
    if (chunk == nchunks - 1) {
      unsigned res = 0;
      bool bad = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned S[RF_PROBES];
        unsigned smin = 0xFFFFFFFFu;
#pragma unroll
        for (int p = 0; p < RF_PROBES; ++p) {
          S[p] = acc[p][k] + __shfl_xor_sync(FULL, acc[p][k], 16) + (unsigned)(delta * (a0[k] + p));
          smin = min(smin, S[p]);
        }
        int lo = -1, hi = -1;   // first / last probe at the minimum
#pragma unroll
        for (int p = 0; p < RF_PROBES; ++p)
          if (S[p] == smin) {
            if (lo < 0) lo = p;
            hi = p;
          }
        // decided only if the minimum does not touch the edge of the probed range (or the edge of the byte range)
        bad |= (lo == 0 && a0[k] > 0) || (hi == RF_PROBES - 1 && a0[k] + RF_PROBES - 1 < 255);
        res |= (unsigned)((2 * a0[k] + lo + hi) >> 1) << (8 * k);
      }
      const bool any_bad = __any_sync(FULL, bad);
      if (part == 0) *reinterpret_cast<unsigned*>(out + seg * 64 + li * 4) = res;
      if (lane == 0) flags[seg] = any_bad ? 1 : 0;
    }
  }
}

}  // namespace msad
}  // namespace vu
