// Temporal background estimators.
//
// vu_temporal_median_u8: exact per-element median over n frames.
//
// n <= 608: register-resident bit-wise binary search on the median value.  A
// warp owns a 64-byte (SPLIT=2) or 32-byte (SPLIT=4) segment of every frame;
// each lane loads its 4 bytes of up to 152 frames straight into registers (one
// coalesced LDG.32 per frame, all of them in flight at once), transposes 4x4
// byte blocks with PRMT so that a register holds 4 frames of ONE element, and
// then resolves the 8 bits of the median MSB-first.  The counting primitive is
// VABSDIFF4 with accumulate: S(m) = sum_f |x_f - m| costs one instruction per
// 4 frames, and count(x <= m) = (S(m+1) - S(m) + N) / 2.  Frames are split over
// half / quarter warps and the partial sums meet in a shuffle.  Unused slots are
// padded with 255, which no probe (<= 254) ever counts.  For even n both middle
// order statistics are searched; the second search only costs extra from the
// round where some lane's two searches part.
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

template <int SPLIT, int G, int CTAS>
__global__ void __launch_bounds__(128, CTAS) median_sad_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, int n, long long m,
                                                               int nseg) {
  constexpr int LPS = 32 / SPLIT;  // lanes per frame-part
  constexpr int SEG = LPS * 4;     // bytes of a frame one warp owns
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = lane / LPS, li = lane % LPS;
  const int seg = blockIdx.x * 4 + warp;
  if (seg >= nseg) return;
  const int base_cnt = n / SPLIT, rem = n % SPLIT;
  const int f_cnt = base_cnt + (part < rem ? 1 : 0);
  const int f_begin = part * base_cnt + min(part, rem);
  const uint8_t* p = frames + (long long)seg * SEG + li * 4 + (long long)f_begin * m;
  unsigned d[4 * G];
#pragma unroll
  for (int k = 0; k < 4 * G; ++k) {
    d[k] = (k < f_cnt) ? __ldg(reinterpret_cast<const unsigned*>(p)) : 0xFFFFFFFFu;
    p += m;
  }
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
  constexpr int ntot = SPLIT * 4 * G;  // slots per element, pads are 255
  const int ta = ((n - 1) >> 1) + 1, tb = (n >> 1) + 1;
  unsigned ma[4] = {0, 0, 0, 0}, mb[4] = {0, 0, 0, 0};
  bool diverged = false;
  auto count_le = [&](const unsigned(&mm)[4], unsigned step, int(&cnt)[4]) {
    unsigned s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, q0[4], q1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q0[j] = (mm[j] + step - 1) * 0x01010101u;
      q1[j] = q0[j] + 0x01010101u;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s0[j] = sad_acc(d[4 * g + j], q0[j], s0[j]);
        s1[j] = sad_acc(d[4 * g + j], q1[j], s1[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ds = (int)s1[j] - (int)s0[j];
#pragma unroll
      for (int o = LPS; o < 32; o <<= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
      cnt[j] = (ds + ntot) >> 1;
    }
  };
#pragma unroll 1
  for (int bit = 7; bit >= 0; --bit) {
    const unsigned step = 1u << bit;
    int ca[4], cb[4];
    count_le(ma, step, ca);
    if (diverged) {
      count_le(mb, step, cb);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) cb[j] = ca[j];
    }
    bool dv = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (ca[j] < ta) ma[j] += step;
      if (cb[j] < tb) mb[j] += step;
      dv |= (ma[j] != mb[j]);
    }
    diverged = __any_sync(0xffffffffu, dv);
  }
  if (part == 0) {
    unsigned res = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) res |= ((ma[j] + mb[j]) >> 1) << (8 * j);
    reinterpret_cast<unsigned*>(out + (long long)seg * SEG)[li] = res;
  }
}

// Persistent variant for the half-warp split: every warp walks segments with a
// grid stride and, while it searches segment i in registers, the next
// segment's rows (64 bytes of each of its <= 8G frames) stream into a private
// shared-memory buffer with cp.async.  The register fill is then conflict-free
// LDS instead of a dependent-latency global load phase.
constexpr int PF_WARPS = 4;
template <int G>
struct PfLayout {
  static constexpr int FT = 4 * G;              // frame slots per half
  static constexpr int HALF_STRIDE = FT * 64 + 64;  // +64 B: the two halves land on different banks
  static constexpr int WARP_BYTES = 2 * HALF_STRIDE;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int G, int CTAS>
__global__ void __launch_bounds__(PF_WARPS * 32, CTAS) median_sad_pf_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, int n,
                                                                             long long m, int nseg) {
  using L = PfLayout<G>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* buf = smem + warp * L::WARP_BYTES;
  const int part = lane >> 4, li = lane & 15;
  const int n0 = (n + 1) >> 1;                   // frames of half 0; half 1 has n - n0
  // pads (slots past a half's frame count) hold 255 for the whole kernel
  for (int i = lane; i < L::WARP_BYTES / 4; i += 32) reinterpret_cast<unsigned*>(buf)[i] = 0xFFFFFFFFu;
  __syncwarp();
  const int total_warps = gridDim.x * PF_WARPS;
  int seg = blockIdx.x * PF_WARPS + warp;
  // lane -> (frame lane>>2 of every group of 8, 16-byte chunk lane&3)
  auto prefetch = [&](int sg) {
    const uint8_t* src = frames + (long long)sg * 64 + (lane & 3) * 16 + (long long)(lane >> 2) * m;
    for (int f = lane >> 2; f < n; f += 8) {
      const int slot_off = f < n0 ? f * 64 : L::HALF_STRIDE + (f - n0) * 64;
      cp_async16(buf + slot_off + (lane & 3) * 16, src);
      src += 8 * m;
    }
  };
  if (seg < nseg) prefetch(seg);
  constexpr int ntot = 2 * 4 * G;
  const int ta = ((n - 1) >> 1) + 1, tb = (n >> 1) + 1;
  for (; seg < nseg; seg += total_warps) {
    cp_async_wait_all();
    __syncwarp();
    unsigned d[4 * G];
    const unsigned* rd = reinterpret_cast<const unsigned*>(buf + part * L::HALF_STRIDE) + li;
#pragma unroll
    for (int k = 0; k < 4 * G; ++k) d[k] = rd[k * 16];
    __syncwarp();
    if (seg + total_warps < nseg) prefetch(seg + total_warps);
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
    unsigned ma[4] = {0, 0, 0, 0}, mb[4] = {0, 0, 0, 0};
    bool diverged = false;
    auto count_le = [&](const unsigned(&mm)[4], unsigned step, int(&cnt)[4]) {
      unsigned s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, q0[4], q1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q0[j] = (mm[j] + step - 1) * 0x01010101u;
        q1[j] = q0[j] + 0x01010101u;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s0[j] = sad_acc(d[4 * g + j], q0[j], s0[j]);
          s1[j] = sad_acc(d[4 * g + j], q1[j], s1[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int ds = (int)s1[j] - (int)s0[j];
        ds += __shfl_xor_sync(0xffffffffu, ds, 16);
        cnt[j] = (ds + ntot) >> 1;
      }
    };
#pragma unroll 1
    for (int bit = 7; bit >= 0; --bit) {
      const unsigned step = 1u << bit;
      int ca[4], cb[4];
      count_le(ma, step, ca);
      if (diverged) {
        count_le(mb, step, cb);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) cb[j] = ca[j];
      }
      bool dv = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ca[j] < ta) ma[j] += step;
        if (cb[j] < tb) mb[j] += step;
        dv |= (ma[j] != mb[j]);
      }
      diverged = __any_sync(0xffffffffu, dv);
    }
    if (part == 0) {
      unsigned res = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) res |= ((ma[j] + mb[j]) >> 1) << (8 * j);
      reinterpret_cast<unsigned*>(out + (long long)seg * 64)[li] = res;
    }
  }
}

template <int G, int CTAS>
int launch_median_sad_pf(const uint8_t* frames, int n, int64_t m, int64_t nseg, uint8_t* out, vu_stream_t stream) {
  constexpr int smem = PF_WARPS * PfLayout<G>::WARP_BYTES;
  static bool configured = false;
  if (!configured) {
    int e = record_cuda(cudaFuncSetAttribute(median_sad_pf_kernel<G, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (e) return e;
    configured = true;
  }
  int64_t grid = (int64_t)device_sms() * CTAS;
  const int64_t need = (nseg + PF_WARPS - 1) / PF_WARPS;
  if (need < grid) grid = need;
  median_sad_pf_kernel<G, CTAS><<<(unsigned)grid, PF_WARPS * 32, smem, S(stream)>>>(frames, out, n, m, (int)nseg);
  note_launch();
  return record_cuda(cudaGetLastError());
}

template <int SPLIT, int G, int CTAS>
int launch_median_sad(const uint8_t* frames, int n, int64_t m, int64_t nseg, uint8_t* out, vu_stream_t stream) {
  median_sad_kernel<SPLIT, G, CTAS><<<(unsigned)((nseg + 3) / 4), 128, 0, S(stream)>>>(frames, out, n, m, (int)nseg);
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
    // G = register groups of 4 frames per lane: the smallest variant that holds ceil(n / SPLIT) frames
    if (path == 0) {
      if (n <= 80) e = launch_median_sad<2, 10, 4>(frames, n, m, nseg, out, stream);
      else if (n <= 152) e = launch_median_sad<2, 19, 4>(frames, n, m, nseg, out, stream);
      else {
        // cp.async moves 16-byte chunks: needs 16-byte aligned frame rows
        // (measured on B200, 300x1080p: the prefetching variant wins when both middle order statistics are
        // searched (even n: 0.81 vs 0.86 ms) and loses slightly when not (odd n: 0.72 vs 0.69 ms))
        const bool pf = (n % 2 == 0) && (reinterpret_cast<uintptr_t>(frames) % 16 == 0) && (m % 16 == 0);
        if (n <= 232) e = pf ? launch_median_sad_pf<29, 2>(frames, n, m, nseg, out, stream) : launch_median_sad<2, 29, 2>(frames, n, m, nseg, out, stream);
        else e = pf ? launch_median_sad_pf<38, 2>(frames, n, m, nseg, out, stream) : launch_median_sad<2, 38, 2>(frames, n, m, nseg, out, stream);
      }
    } else if (path == 1) {
      e = n <= 464 ? launch_median_sad<4, 29, 2>(frames, n, m, nseg, out, stream) : launch_median_sad<4, 38, 2>(frames, n, m, nseg, out, stream);
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
