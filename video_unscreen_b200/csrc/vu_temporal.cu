// Temporal background estimators.
//
// vu_temporal_median_u8: exact per-element median over n frames.
//
// n <= 608: register-resident bit-wise binary search on the median value.  A
// warp owns a 64-byte (SPLIT=2) or 32-byte (SPLIT=4) segment of every frame;
// each lane loads its 4 bytes of up to 152 frames straight into registers (one
// coalesced LDG.32 per frame, all of them in flight at once), transposes 4x4
// byte blocks with PRMT so that a register holds 4 frames of ONE element, and
// then minimises the convex S(m) = sum_f |x_f - m| with a Fibonacci search: S
// costs one VABSDIFF4-with-accumulate per 4 frames per probe, 12 probes find
// the minimiser among 0..255 (see sad_search).  Frames are split over half /
// quarter warps and the partial sums meet in a shuffle.  Unused slots are
// padded with 0 on one side and 255 on the other, which keeps the middle order
// statistics where they are.
//
// n > 608: streaming 256-bin histograms in shared memory.  One warp owns a segment of
// 32*PX consecutive bytes of every frame; lane l owns PX of them and a private
// column of 256 packed counters laid out so that its bank is always `lane`
// (bin stride = WARPS*128 bytes): every increment is a conflict-free shared
// atomic, whatever the pixel values are.  Warps never synchronise with each
// other, so one warp's histogram scan overlaps the other warps' streaming.
// Two 16-bit counters per word (PX = 2, 64-byte segments), n <= 65535.
// The median is read back with a two-level scan (16 coarse groups, then 16
// bins); even n returns (lo + hi) >> 1 like np.median(...).astype(uint8).
//
// vu_masked_temporal_mean: tools/unscreen/bg_offline.py:106-125 as integer
// sums and counts per pixel, one float64 divide at the end.
#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int WARPS = 7;
constexpr int MTHREADS = WARPS * 32;
constexpr int BIN_STRIDE = WARPS * 128;  // bytes between bins
constexpr int HIST_BYTES = 256 * BIN_STRIDE;

struct PolU16x2 {
  static constexpr int PX = 2;
  __device__ static __forceinline__ void add(unsigned char* base, unsigned w) {
#pragma unroll
    for (int j = 0; j < 2; ++j) atomicAdd(reinterpret_cast<unsigned*>(base + ((w >> (8 * j)) & 0xFFu) * BIN_STRIDE), 1u << (16 * j));
  }
  __device__ static __forceinline__ void unpack(unsigned w, unsigned* c) { c[0] = w & 0xFFFFu; c[1] = w >> 16; }
};

template <int PX>
__device__ __forceinline__ unsigned load_px(const uint8_t* p) {
  if (PX == 4) return __ldg(reinterpret_cast<const unsigned*>(p));
  return __ldg(reinterpret_cast<const unsigned short*>(p));
}

template <class P, int U>
__global__ void __launch_bounds__(MTHREADS, 1) median_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, int n, long long m,
                                                             int nseg) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int PX = P::PX;
  constexpr int SEG = 32 * PX;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* base = smem + warp * 128 + lane * 4;
  const unsigned klo = (unsigned)(n - 1) >> 1, khi = (unsigned)n >> 1;
  const int nb = n / U;
  for (int seg = blockIdx.x * WARPS + warp; seg < nseg; seg += gridDim.x * WARPS) {
#pragma unroll 8
    for (int b = 0; b < 256; ++b) *reinterpret_cast<unsigned*>(base + b * BIN_STRIDE) = 0u;
    __syncwarp();
    {
      const uint8_t* p = frames + (long long)seg * SEG + lane * PX;
      unsigned ra[U], rb[U];
      auto load = [&](unsigned(&r)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) { r[u] = load_px<PX>(p); p += m; }
      };
      auto proc = [&](unsigned(&r)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) P::add(base, r[u]);
      };
      if (nb > 0) load(ra);
      for (int i = 0; i < nb; i += 2) {
        if (i + 1 < nb) load(rb);
        proc(ra);
        if (i + 1 < nb) {
          if (i + 2 < nb) load(ra);
          proc(rb);
        }
      }
      for (int f = nb * U; f < n; ++f) { P::add(base, load_px<PX>(p)); p += m; }
    }
    __syncwarp();
    // two-level scan
    unsigned cum[PX], gsel[PX][2], before[PX][2];
#pragma unroll
    for (int j = 0; j < PX; ++j) { cum[j] = 0; gsel[j][0] = gsel[j][1] = 0; before[j][0] = before[j][1] = 0; }
#pragma unroll 1
    for (int g = 0; g < 16; ++g) {
      unsigned c[PX];
#pragma unroll
      for (int j = 0; j < PX; ++j) c[j] = 0;
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        unsigned t[PX];
        P::unpack(*reinterpret_cast<unsigned*>(base + (g * 16 + b) * BIN_STRIDE), t);
#pragma unroll
        for (int j = 0; j < PX; ++j) c[j] += t[j];
      }
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        const unsigned nc = cum[j] + c[j];
        if (nc <= klo) { gsel[j][0] = g + 1; before[j][0] = nc; }
        if (nc <= khi) { gsel[j][1] = g + 1; before[j][1] = nc; }
        cum[j] = nc;
      }
    }
    unsigned res = 0;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      unsigned med[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const unsigned k = q ? khi : klo;
        if (q == 1 && khi == klo) { med[1] = med[0]; continue; }
        const unsigned g = min(gsel[j][q], 15u);
        unsigned v = g * 16, c2 = before[j][q];
#pragma unroll 4
        for (int b = 0; b < 16; ++b) {
          unsigned t[PX];
          P::unpack(*reinterpret_cast<unsigned*>(base + (g * 16 + b) * BIN_STRIDE), t);
          c2 += t[j];
          if (c2 <= k) v = g * 16 + b + 1;
        }
        med[q] = min(v, 255u);
      }
      res |= ((med[0] + med[1]) >> 1) << (8 * j);
    }
    uint8_t* dst = out + (long long)seg * SEG;
    if (PX == 4) reinterpret_cast<unsigned*>(dst)[lane] = res;
    else reinterpret_cast<unsigned short*>(dst)[lane] = (unsigned short)res;
    __syncwarp();
  }
}

// ---- register-resident SAD search ------------------------------------------
__device__ __forceinline__ unsigned sad_acc(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// frames of one element are split over SPLIT lane groups ("parts").  Parts on
// the low side (part < SPLIT/2) pad their unused slots with 0, parts on the
// high side with 255, and the surplus frames go to the sides alternately, so
// the padded multiset keeps the data's middle order statistics.
template <int SPLIT>
__device__ __forceinline__ void part_frames(int n, int part, int& f_begin, int& f_cnt) {
  const int base = n / SPLIT, rem = n % SPLIT;
  if (SPLIT == 4) {
    // surplus order: parts 0, 2, 1, 3
    const int c0 = base + (rem > 0), c1 = base + (rem > 2), c2 = base + (rem > 1), c3 = base;
    f_cnt = part == 0 ? c0 : (part == 1 ? c1 : (part == 2 ? c2 : c3));
    f_begin = part == 0 ? 0 : (part == 1 ? c0 : (part == 2 ? c0 + c1 : c0 + c1 + c2));
  } else {
    f_cnt = base + (part < rem ? 1 : 0);
    f_begin = part * base + min(part, rem);
  }
}

// 4x4 byte transposes: d[4g+j] <- 4 consecutive frame slots of element j
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

// The median minimises the convex S(m) = sum_f |x_f - m| (one VABSDIFF4.ACC per
// 4 frames per probe).  Fibonacci search over m = 0..255 needs 12 evaluations
// of S per element (each round re-uses one of its two probes).  For odd n one
// high-side pad "floats" (it is overwritten with the probe itself and so adds
// nothing): the minimiser is unique and is the median.  For even n the
// minimisers form the plateau [x_lo, x_hi] of the two middle order statistics;
// its ends are found by probing outwards (S(m) == S_min is monotone on either
// side), a few linear steps first, binary search for pathological plateaus.
template <int SPLIT, int G>
__device__ __forceinline__ unsigned sad_search(const unsigned (&d)[4 * G], int n, int part) {
  constexpr int LPS = 32 / SPLIT;
  constexpr unsigned INF = 0x7fffffffu;
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned fmask = ((n & 1) && part == SPLIT - 1) ? 0xFF000000u : 0u;
  // idx = m + 1 (the search runs on indices 1..376; m > 255 evaluates to +inf)
  auto evalS = [&](const int(&idx)[4], unsigned(&S)[4]) {
    unsigned q[4], s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[j] = (unsigned)min(max(idx[j] - 1, 0), 255) * 0x01010101u;
      s[j] = 0;
    }
#pragma unroll
    for (int g = 0; g < G - 1; ++g) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] = sad_acc(d[4 * g + j], q[j], s[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned w = (d[4 * (G - 1) + j] & ~fmask) | (q[j] & fmask);
      s[j] = sad_acc(w, q[j], s[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int o = LPS; o < 32; o <<= 1) s[j] += __shfl_xor_sync(FULL, s[j], o);
      S[j] = (idx[j] < 1 || idx[j] > 256) ? INF : s[j];
    }
  };
  int a[4] = {0, 0, 0, 0};
  unsigned S1[4], S2[4];
  int fa = 233, fb = 144, fc = 89;  // F(k-1), F(k-2), F(k-3) for k = 14
  {
    int i1[4], i2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { i1[j] = fb; i2[j] = fa; }
    evalS(i1, S1);
    evalS(i2, S2);
  }
#pragma unroll 1
  for (int k = 14; k >= 5; --k) {
    int nidx[4];
    bool left[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      left[j] = S1[j] <= S2[j];
      if (left[j]) {
        S2[j] = S1[j];
        nidx[j] = a[j] + fc;
      } else {
        a[j] += fb;
        S1[j] = S2[j];
        nidx[j] = a[j] + fb;
      }
    }
    unsigned Sn[4];
    evalS(nidx, Sn);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (left[j]) S1[j] = Sn[j];
      else S2[j] = Sn[j];
    }
    const int t = fb - fc;
    fa = fb; fb = fc; fc = t;
  }
  int lo[4], hi[4];
  unsigned Smin[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool left = S1[j] <= S2[j];
    lo[j] = hi[j] = (left ? a[j] + fb : a[j] + fa) - 1;
    Smin[j] = left ? S1[j] : S2[j];
  }
  if (!(n & 1)) {
    bool actL[4], actR[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { actL[j] = lo[j] > 0; actR[j] = hi[j] < 255; }
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
      const bool anyL = actL[0] | actL[1] | actL[2] | actL[3];
      if (__any_sync(FULL, anyL)) {
        int idx[4];
        unsigned S[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = lo[j];  // m = lo - 1
        evalS(idx, S);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (actL[j]) {
            if (S[j] == Smin[j]) { --lo[j]; actL[j] = lo[j] > 0; }
            else actL[j] = false;
          }
      }
      const bool anyR = actR[0] | actR[1] | actR[2] | actR[3];
      if (__any_sync(FULL, anyR)) {
        int idx[4];
        unsigned S[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = hi[j] + 2;  // m = hi + 1
        evalS(idx, S);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (actR[j]) {
            if (S[j] == Smin[j]) { ++hi[j]; actR[j] = hi[j] < 255; }
            else actR[j] = false;
          }
      }
    }
    // plateaus longer than 3 on a side (e.g. two-valued data): binary search for the end
    if (__any_sync(FULL, actL[0] | actL[1] | actL[2] | actL[3])) {
      int L[4], R[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { R[j] = lo[j]; L[j] = actL[j] ? 0 : lo[j]; }
      while (__any_sync(FULL, (L[0] < R[0]) | (L[1] < R[1]) | (L[2] < R[2]) | (L[3] < R[3]))) {
        int idx[4];
        unsigned S[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = ((L[j] + R[j]) >> 1) + 1;
        evalS(idx, S);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (L[j] < R[j]) {
            const int mid = (L[j] + R[j]) >> 1;
            if (S[j] == Smin[j]) R[j] = mid;
            else L[j] = mid + 1;
          }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) lo[j] = R[j];
    }
    if (__any_sync(FULL, actR[0] | actR[1] | actR[2] | actR[3])) {
      int L[4], R[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { L[j] = hi[j]; R[j] = actR[j] ? 255 : hi[j]; }
      while (__any_sync(FULL, (L[0] < R[0]) | (L[1] < R[1]) | (L[2] < R[2]) | (L[3] < R[3]))) {
        int idx[4];
        unsigned S[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = ((L[j] + R[j] + 1) >> 1) + 1;
        evalS(idx, S);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (L[j] < R[j]) {
            const int mid = (L[j] + R[j] + 1) >> 1;
            if (S[j] == Smin[j]) L[j] = mid;
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

// GMIN = register groups that are full for every n the variant is dispatched for (no bounds test on those loads)
template <int SPLIT, int G, int GMIN, int CTAS>
__global__ void __launch_bounds__(128, CTAS) median_sad_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, int n, long long m,
                                                               int nseg) {
  constexpr int LPS = 32 / SPLIT;  // lanes per frame-part
  constexpr int SEG = LPS * 4;     // bytes of a frame one warp owns
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = lane / LPS, li = lane % LPS;
  const int seg = blockIdx.x * 4 + warp;
  if (seg >= nseg) return;
  int f_begin, f_cnt;
  part_frames<SPLIT>(n, part, f_begin, f_cnt);
  const unsigned pad = part >= SPLIT / 2 ? 0xFFFFFFFFu : 0u;
  const uint8_t* p = frames + (long long)seg * SEG + li * 4 + (long long)f_begin * m;
  unsigned d[4 * G];
#pragma unroll
  for (int k = 0; k < 4 * GMIN; ++k) {
    d[k] = __ldg(reinterpret_cast<const unsigned*>(p));
    p += m;
  }
#pragma unroll
  for (int k = 4 * GMIN; k < 4 * G; ++k) {
    d[k] = (k < f_cnt) ? __ldg(reinterpret_cast<const unsigned*>(p)) : pad;
    p += m;
  }
  transpose_groups<G>(d);
  const unsigned res = sad_search<SPLIT, G>(d, n, part);
  if (part == 0) reinterpret_cast<unsigned*>(out + (long long)seg * SEG)[li] = res;
}

template <int SPLIT, int G, int GMIN, int CTAS>
int launch_median_sad(const uint8_t* frames, int n, int64_t m, int64_t nseg, uint8_t* out, vu_stream_t stream) {
  median_sad_kernel<SPLIT, G, GMIN, CTAS><<<(unsigned)((nseg + 3) / 4), 128, 0, S(stream)>>>(frames, out, n, m, (int)nseg);
  note_launch();
  return record_cuda(cudaGetLastError());
}

// elements that do not fill a whole segment (and unaligned inputs): one CTA
// per element, 256-bin shared histogram
__global__ void __launch_bounds__(256) median_tail_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, int n, long long m,
                                                          long long first) {
  __shared__ unsigned hist[256];
  const long long e = first + blockIdx.x;
  hist[threadIdx.x] = 0;
  __syncthreads();
  for (int f = threadIdx.x; f < n; f += 256) atomicAdd(&hist[__ldg(frames + (long long)f * m + e)], 1u);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned klo = (unsigned)(n - 1) >> 1, khi = (unsigned)n >> 1;
    unsigned cum = 0, lo = 255, hi = 255;
    bool flo = false, fhi = false;
    for (int b = 0; b < 256; ++b) {
      cum += hist[b];
      if (!flo && cum > klo) { lo = b; flo = true; }
      if (!fhi && cum > khi) { hi = b; fhi = true; }
    }
    out[e] = (uint8_t)((lo + hi) >> 1);
  }
}

// ---- masked temporal mean -------------------------------------------------
constexpr int TT = 256;
__global__ void __launch_bounds__(TT) masked_mean_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ masks, int n,
                                                         int64_t ngroups, int64_t npix, int min_count, uint8_t* __restrict__ bg_out,
                                                         uint8_t* __restrict__ always_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    unsigned sum[12], cnt[4];
#pragma unroll
    for (int i = 0; i < 12; ++i) sum[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) cnt[i] = 0;
    const unsigned* f4 = reinterpret_cast<const unsigned*>(frames) + 3 * g;
    const unsigned* m4 = reinterpret_cast<const unsigned*>(masks) + g;
    const int64_t fstride = npix * 3 / 4, mstride = npix / 4;
#pragma unroll 4
    for (int f = 0; f < n; ++f) {
      int c[12];
      unpack12(__ldg(f4), __ldg(f4 + 1), __ldg(f4 + 2), c);
      const unsigned mw = __ldg(m4);
      f4 += fstride;
      m4 += mstride;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned mv = (mw >> (8 * i)) & 255u;
        const unsigned keep = 1u - mv / 255u;  // frame * (1 - mask // 255)
        sum[3 * i] += c[3 * i] * keep;
        sum[3 * i + 1] += c[3 * i + 1] * keep;
        sum[3 * i + 2] += c[3 * i + 2] * keep;
        cnt[i] += mv < 250u;               // count += (mask < 250)
      }
    }
    int o[12];
    unsigned aw = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool always = cnt[i] <= (unsigned)min_count;
      const double den = (double)(cnt[i] ? cnt[i] : 1u);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double q = fmin(fmax(__ddiv_rn((double)sum[3 * i + k], den), 0.0), 255.0);
        o[3 * i + k] = always ? 0 : (int)q;
      }
      aw |= (always ? 255u : 0u) << (8 * i);
    }
    unsigned w0, w1, w2;
    pack12(o, w0, w1, w2);
    unsigned* d4 = reinterpret_cast<unsigned*>(bg_out) + 3 * g;
    d4[0] = w0; d4[1] = w1; d4[2] = w2;
    reinterpret_cast<unsigned*>(always_out)[g] = aw;
  }
}

template <class P, int U>
int launch_median(const uint8_t* frames, int n, int64_t m, int64_t nseg, uint8_t* out, vu_stream_t stream) {
  static bool configured = false;
  if (!configured) {
    int e = record_cuda(cudaFuncSetAttribute(median_kernel<P, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, HIST_BYTES));
    if (e) return e;
    configured = true;
  }
  int grid = device_sms();
  const int64_t need = (nseg + WARPS - 1) / WARPS;
  if (need < grid) grid = (int)need;
  median_kernel<P, U><<<grid, MTHREADS, HIST_BYTES, S(stream)>>>(frames, out, n, m, (int)nseg);
  note_launch();
  return record_cuda(cudaGetLastError());
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_temporal_median_u8(const uint8_t* frames, int n, int64_t m, uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(frames && out && m >= 0);
  if (n < 1 || n > 65535) return VU_ERR_UNSUPPORTED;
  if (m == 0) return VU_OK;
  // path: 0 = SAD search, half-warp split (n <= 304); 1 = SAD search, quarter-warp split (n <= 608);
  //       2 = shared-memory histograms with 16-bit counters
  const int path = n <= 304 ? 0 : (n <= 608 ? 1 : 2);
  const int seg = path == 0 ? 64 : (path == 1 ? 32 : 64);
  const int align = path == 2 ? 2 : 4;
  // vector paths need element-aligned frame rows
  const bool ok = (reinterpret_cast<uintptr_t>(frames) % align == 0) && (reinterpret_cast<uintptr_t>(out) % align == 0) && (m % align == 0);
  int64_t nseg = ok ? m / seg : 0;
  if (nseg > 0x7fffffff) return VU_ERR_UNSUPPORTED;
  if (nseg > 0) {
    int e;
    // G = register groups of 4 frames per lane: the smallest variant that holds ceil(n / SPLIT) frames;
    // GMIN = groups every lane fills for the whole n-range of the variant
    if (path == 0) {
      if (n <= 80) e = launch_median_sad<2, 10, 0, 4>(frames, n, m, nseg, out, stream);
      else if (n <= 152) e = launch_median_sad<2, 19, 10, 4>(frames, n, m, nseg, out, stream);
      else if (n <= 232) e = launch_median_sad<2, 29, 19, 2>(frames, n, m, nseg, out, stream);
      else e = launch_median_sad<2, 38, 29, 2>(frames, n, m, nseg, out, stream);
    } else if (path == 1) {
      e = n <= 464 ? launch_median_sad<4, 29, 19, 2>(frames, n, m, nseg, out, stream) : launch_median_sad<4, 38, 29, 2>(frames, n, m, nseg, out, stream);
    }
    else e = launch_median<PolU16x2, 32>(frames, n, m, nseg, out, stream);
    if (e) return e;
  }
  const int64_t first = nseg * seg;
  if (first < m) {
    if (m - first > 0x7fffffff) return VU_ERR_UNSUPPORTED;
    median_tail_kernel<<<(unsigned)(m - first), 256, 0, S(stream)>>>(frames, out, n, m, first);
    note_launch();
    return record_cuda(cudaGetLastError());
  }
  return VU_OK;
}

extern "C" int vu_masked_temporal_mean(const uint8_t* frames, const uint8_t* masks, int n, int64_t npix, int min_count, uint8_t* bg_out,
                                       uint8_t* mask_always_out, vu_stream_t stream) {
  VU_REQUIRE(frames && masks && bg_out && mask_always_out && n >= 1 && npix >= 0);
  if (n > 16000000) return VU_ERR_UNSUPPORTED;  // 32-bit sums: 255 * n must fit
  if (npix % 4 != 0) return VU_ERR_UNSUPPORTED;
  const void* ptrs[] = {frames, masks, bg_out, mask_always_out};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 3) return VU_ERR_UNSUPPORTED;
  if (npix == 0) return VU_OK;
  masked_mean_kernel<<<grid_for(npix / 4, TT, 8), TT, 0, S(stream)>>>(frames, masks, n, npix / 4, npix, min_count, bg_out, mask_always_out);
  VU_RETURN_LAUNCH();
}
