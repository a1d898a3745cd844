// color_correct (unscreen/utils/imgprocess.py:263-300; called from tools/unscreen/green.py:120) for a clip, the part
// after the two cv2.resize calls (those are vu_resize_linear_u8): chroma distance to the background colour in Lab at
// the working resolution, normalised to [0,1] per frame, square-rooted until its mean over the matte reaches
// `mean_exp`, zeroed outside the matte, nearest-neighbour up-sampled and multiplied into alpha.
//
//   lab_dist   BGR2Lab (cv2's integer tables: SURVEY.md section 8f-1, oracle/cvmodel.py:bgr2lab) -> a, b ->
//              sqrt((a/255 - bg_a/255)^2 + (b/255 - bg_b/255)^2) in float32, + per-frame min / max
//   sums       S_k = sum over the matte of the normalised distance after k square roots, k = 0..23, + the count
//              (float64; the reference's loop `while mean < mean_exp: dist = sqrt(dist)` evaluated for every k at once,
//              so that nothing comes back to the host)
//   final      k* = first k with S_k / count >= mean_exp; working-resolution map = sqrt^k*(normalised distance), 0
//              where the working-resolution alpha is 0
//   apply      out = trunc(float(alpha) * map[nearest])   (torch.nn.functional.interpolate 'nearest' indices)
//
// Every float operation is the reference's float32 operation (torch CPU), in its order; -fmad=false keeps them apart.
#include "vu_common.cuh"
#include "vu_lab_tables.inc"

namespace vu {
namespace {

constexpr int CC_K = 24;          // square roots tabulated: d^(1/2^23) > 0.99998 for every positive float32 d
constexpr int CC_K1 = 8;          // ... of which the first 8 for every frame
constexpr int CC_THREADS = 256;
constexpr int LAB_CBRT_N = 3072;

__device__ const unsigned short d_lab_gamma[256] = {VU_LAB_GAMMA_TABLE};
__device__ const unsigned short d_lab_cbrt[LAB_CBRT_N] = {VU_LAB_CBRT_TABLE};
const unsigned short h_lab_gamma[256] = {VU_LAB_GAMMA_TABLE};
const unsigned short h_lab_cbrt[LAB_CBRT_N] = {VU_LAB_CBRT_TABLE};

// 12-bit sRGB -> XYZ / white point matrix of cv2's RGB2Lab_b (rows X, Y, Z; columns R, G, B)
#define VU_LAB_COEFFS {1777, 1541, 778, 871, 2929, 296, 73, 448, 3575}

struct LabStats {
  unsigned lo, hi;                 // float bits of min / max of the distance (non-negative floats order like their bits)
  unsigned long long count;        // matte pixels with a positive normalised distance
  double sums[CC_K];
};

// a and b of cv2's BGR2Lab (uint8): gamma table, 12-bit matrix, cube-root table, 15-bit descale
template <typename G, typename T>
__host__ __device__ inline void lab_ab(int b8, int g8, int r8, const G* gamma, const T* cbrt, int& a, int& b) {
  const int C[9] = VU_LAB_COEFFS;
  const int R = gamma[r8], Gv = gamma[g8], B = gamma[b8];
  const int fx = cbrt[(R * C[0] + Gv * C[1] + B * C[2] + 2048) >> 12];
  const int fy = cbrt[(R * C[3] + Gv * C[4] + B * C[5] + 2048) >> 12];
  const int fz = cbrt[(R * C[6] + Gv * C[7] + B * C[8] + 2048) >> 12];
  a = (500 * (fx - fy) + (128 << 15) + 16384) >> 15;
  b = (200 * (fy - fz) + (128 << 15) + 16384) >> 15;
  a = a < 0 ? 0 : (a > 255 ? 255 : a);
  b = b < 0 ? 0 : (b > 255 ? 255 : b);
}

__global__ void __launch_bounds__(CC_THREADS) cc_init_kernel(LabStats* stats, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  stats[i].lo = 0x7F800000u;       // +inf
  stats[i].hi = 0u;
  stats[i].count = 0ull;
  for (int k = 0; k < CC_K; ++k) stats[i].sums[k] = 0.0;
}

__global__ void __launch_bounds__(CC_THREADS) cc_lab_dist_kernel(const uint8_t* __restrict__ frames_lo, int64_t per, float bg_a, float bg_b,
                                                                 float* __restrict__ dist, LabStats* __restrict__ stats) {
  __shared__ unsigned short gamma[256], cbrt[LAB_CBRT_N];
  __shared__ float q255[256];      // i / 255.f
  __shared__ unsigned red[2][CC_THREADS / 32];
  for (int i = threadIdx.x; i < 256; i += CC_THREADS) {
    gamma[i] = d_lab_gamma[i];
    q255[i] = __fdiv_rn((float)i, 255.f);
  }
  for (int i = threadIdx.x; i < LAB_CBRT_N; i += CC_THREADS) cbrt[i] = d_lab_cbrt[i];
  __syncthreads();
  const int f = blockIdx.y;
  const uint8_t* px = frames_lo + (int64_t)f * per * 3;
  float* out = dist + (int64_t)f * per;
  unsigned lo = 0x7F800000u, hi = 0u;
  for (int64_t i = (int64_t)blockIdx.x * CC_THREADS + threadIdx.x; i < per; i += (int64_t)gridDim.x * CC_THREADS) {
    int a, b;
    lab_ab(px[3 * i], px[3 * i + 1], px[3 * i + 2], gamma, cbrt, a, b);
    const float da = q255[a] - bg_a, db = q255[b] - bg_b;
    const float d = sqrtf(da * da + db * db);
    out[i] = d;
    const unsigned u = __float_as_uint(d);
    lo = min(lo, u);
    hi = max(hi, u);
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = lo;
    red[1][threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < CC_THREADS / 32; ++k) {
      lo = min(lo, red[0][k]);
      hi = max(hi, red[1][k]);
    }
    atomicMin(&stats[f].lo, lo);
    atomicMax(&stats[f].hi, hi);
  }
}

// The same with cv2's down-scale in front, for frames that are exactly S = 2 or 4 times the working resolution (1080p
// and 4K at the default 960): 2x is INTER_AREA's rounded 2x2 mean, 4x the rounded mean of the centre 2x2 of every 4x4
// block (what the fixed-point bilinear comes to there).  A thread takes 16 full-resolution columns of two
// rows (128-bit loads, all in flight before the first conversion) of the frame and of alpha, and writes 16 / S
// working-resolution distances and alphas: the resized frame never exists.
template <int S>
__global__ void __launch_bounds__(CC_THREADS) cc_lab_dist_lowres_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ alpha, int h, int w,
                                                                        int th, int tw, float bg_a, float bg_b, float* __restrict__ dist,
                                                                        uint8_t* __restrict__ alpha_lo, LabStats* __restrict__ stats) {
  constexpr int NO = 16 / S;       // outputs per thread
  __shared__ unsigned short gamma[256], cbrt[LAB_CBRT_N];
  __shared__ float q255[256];
  __shared__ unsigned red[2][CC_THREADS / 32];
  for (int i = threadIdx.x; i < 256; i += CC_THREADS) {
    gamma[i] = d_lab_gamma[i];
    q255[i] = __fdiv_rn((float)i, 255.f);
  }
  for (int i = threadIdx.x; i < LAB_CBRT_N; i += CC_THREADS) cbrt[i] = d_lab_cbrt[i];
  __syncthreads();
  const int f = blockIdx.y;
  const int per_row = tw / NO;
  const int64_t items = (int64_t)th * per_row;
  const uint8_t* fr = frames + (int64_t)f * h * w * 3;
  const uint8_t* al = alpha + (int64_t)f * h * w;
  float* dout = dist + (int64_t)f * th * tw;
  uint8_t* aout = alpha_lo + (int64_t)f * th * tw;
  unsigned lo = 0x7F800000u, hi = 0u;
  for (int64_t it = (int64_t)blockIdx.x * CC_THREADS + threadIdx.x; it < items; it += (int64_t)gridDim.x * CC_THREADS) {
    const int y = (int)(it / per_row), xg = (int)(it - (int64_t)y * per_row);
    const int r0 = S == 2 ? 2 * y : 4 * y + 1;
    uint4 fv[2][3], av[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint4* p = reinterpret_cast<const uint4*>(fr + ((int64_t)(r0 + r) * w + 16 * xg) * 3);
#pragma unroll
      for (int k = 0; k < 3; ++k) fv[r][k] = ldg_stream16(p + k);
      av[r] = ldg_stream16(al + (int64_t)(r0 + r) * w + 16 * xg);
    }
    float dv[NO];
    unsigned ares = 0, ares2 = 0;
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      const int c0 = S == 2 ? 2 * o : 4 * o + 1;     // first of the two columns averaged
      int acc[3] = {2, 2, 2}, aacc = 2;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const unsigned* fw = reinterpret_cast<const unsigned*>(fv[r]);
        const unsigned* aw = reinterpret_cast<const unsigned*>(&av[r]);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int q = c0 + k;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int bi = 3 * q + c;
            acc[c] = (int)__dp4a(fw[bi >> 2], 1u << (8 * (bi & 3)), (unsigned)acc[c]);   // byte extraction + add on the FMA pipe
          }
          aacc = (int)__dp4a(aw[q >> 2], 1u << (8 * (q & 3)), (unsigned)aacc);
        }
      }
      int a, b;
      lab_ab(acc[0] >> 2, acc[1] >> 2, acc[2] >> 2, gamma, cbrt, a, b);
      const float da = q255[a] - bg_a, db = q255[b] - bg_b;
      const float d = sqrtf(da * da + db * db);
      dv[o] = d;
      const unsigned u = __float_as_uint(d);
      lo = min(lo, u);
      hi = max(hi, u);
      if (o < 4) ares |= (unsigned)(aacc >> 2) << (8 * o);
      else ares2 |= (unsigned)(aacc >> 2) << (8 * (o - 4));
    }
    float4* dp = reinterpret_cast<float4*>(dout + (int64_t)y * tw + NO * xg);
    dp[0] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    if (NO == 8) {
      dp[1] = make_float4(dv[NO - 4], dv[NO - 3], dv[NO - 2], dv[NO - 1]);
      *reinterpret_cast<uint2*>(aout + (int64_t)y * tw + NO * xg) = make_uint2(ares, ares2);
    } else {
      *reinterpret_cast<unsigned*>(aout + (int64_t)y * tw + NO * xg) = ares;
    }
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = lo;
    red[1][threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < CC_THREADS / 32; ++k) {
      lo = min(lo, red[0][k]);
      hi = max(hi, red[1][k]);
    }
    atomicMin(&stats[f].lo, lo);
    atomicMax(&stats[f].hi, hi);
  }
}

__device__ __forceinline__ float cc_normalise(float d, float lo, float hi) { return __fdiv_rn(d - lo, hi - lo); }

// number of square roots the reference's loop takes: the first k with S_k / count >= mean_exp (`while mean < mean_exp`;
// an empty matte has mean NaN: no iteration).  Lane k tests S_k.
__device__ __forceinline__ int cc_iters(const LabStats& st, double mean_exp, int kmax) {
  const int lane = threadIdx.x & 31;
  const bool below = st.count != 0 && lane < kmax && st.sums[lane < CC_K ? lane : 0] / (double)st.count < mean_exp;
  const unsigned m = __ballot_sync(0xffffffffu, below || lane >= kmax);
  return min(__ffs(~m) - 1 < 0 ? 32 : __ffs(~m) - 1, kmax);
}

// S_k for k in [K0, K0 + NK).  The first range runs for every frame; the second one only for the frames whose mean is
// still below mean_exp after CC_K1 - 1 square roots (hardly ever: d^(1/128) is above 0.95 for d > 0.0014).
template <int K0, int NK>
__global__ void __launch_bounds__(CC_THREADS) cc_sums_kernel(const float* __restrict__ dist, const uint8_t* __restrict__ alpha_lo, int64_t per,
                                                             LabStats* __restrict__ stats, double mean_exp, int vec) {
  __shared__ double part[CC_THREADS / 32][NK];
  __shared__ unsigned cpart[CC_THREADS / 32];
  const int f = blockIdx.y;
  if (K0 > 0 && cc_iters(stats[f], mean_exp, K0) < K0) return;   // decided already (uniform over the CTA)
  const float lo = __uint_as_float(stats[f].lo), hi = __uint_as_float(stats[f].hi);
  const float* d = dist + (int64_t)f * per;
  const uint8_t* al = alpha_lo + (int64_t)f * per;
  double s[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) s[k] = 0.0;
  unsigned cnt = 0;
  auto one = [&](float dv, unsigned a) {
    float v = cc_normalise(dv, lo, hi);
    if (a > 0 && v > 0.f) {
      ++cnt;
#pragma unroll 1
      for (int k = 0; k < K0; ++k) v = sqrtf(v);
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        s[k] += (double)v;
        v = sqrtf(v);
      }
    }
  };
  if (vec) {   // per % 4 == 0, aligned: 16 + 4 bytes per load
    for (int64_t i = 4 * ((int64_t)blockIdx.x * CC_THREADS + threadIdx.x); i < per; i += 4 * (int64_t)gridDim.x * CC_THREADS) {
      const unsigned a = __ldg(reinterpret_cast<const unsigned*>(al + i));
      if (a == 0) continue;        // outside the matte (most of a frame): the distances are not even read
      const float4 dv = __ldg(reinterpret_cast<const float4*>(d + i));
      one(dv.x, a & 255u);
      one(dv.y, (a >> 8) & 255u);
      one(dv.z, (a >> 16) & 255u);
      one(dv.w, a >> 24);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * CC_THREADS + threadIdx.x; i < per; i += (int64_t)gridDim.x * CC_THREADS) one(d[i], al[i]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    double v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) part[warp][k] = v;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) cpart[warp] = cnt;
  __syncthreads();
  if (threadIdx.x < NK) {
    double v = 0.0;
    for (int wv = 0; wv < CC_THREADS / 32; ++wv) v += part[wv][threadIdx.x];
    if (v != 0.0) atomicAdd(&stats[f].sums[K0 + threadIdx.x], v);
  }
  if (K0 == 0 && threadIdx.x == 32) {
    unsigned c = 0;
    for (int wv = 0; wv < CC_THREADS / 32; ++wv) c += cpart[wv];
    if (c) atomicAdd(&stats[f].count, (unsigned long long)c);
  }
}

__global__ void __launch_bounds__(CC_THREADS) cc_final_kernel(float* __restrict__ dist, const uint8_t* __restrict__ alpha_lo, int64_t per,
                                                              const LabStats* __restrict__ stats, double mean_exp, int vec) {
  const int f = blockIdx.y;
  const LabStats& st = stats[f];
  const float lo = __uint_as_float(st.lo), hi = __uint_as_float(st.hi);
  const int iters = cc_iters(st, mean_exp, CC_K);
  float* d = dist + (int64_t)f * per;
  const uint8_t* al = alpha_lo + (int64_t)f * per;
  auto one = [&](float dv, unsigned a) -> float {
    float v = cc_normalise(dv, lo, hi);
    for (int k = 0; k < iters; ++k) v = sqrtf(v);
    return a == 0 ? 0.f : v;
  };
  if (vec) {
    for (int64_t i = 4 * ((int64_t)blockIdx.x * CC_THREADS + threadIdx.x); i < per; i += 4 * (int64_t)gridDim.x * CC_THREADS) {
      const unsigned a = __ldg(reinterpret_cast<const unsigned*>(al + i));
      if (a == 0) {                // outside the matte: zero, whatever the distance
        *reinterpret_cast<float4*>(d + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      const float4 dv = *reinterpret_cast<const float4*>(d + i);
      *reinterpret_cast<float4*>(d + i) = make_float4(one(dv.x, a & 255u), one(dv.y, (a >> 8) & 255u), one(dv.z, (a >> 16) & 255u), one(dv.w, a >> 24));
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * CC_THREADS + threadIdx.x; i < per; i += (int64_t)gridDim.x * CC_THREADS) d[i] = one(d[i], al[i]);
  }
}

// nearest source index of torch's CPU interpolate: identity, i >> 1 for an exact doubling, else floor(float(i) * scale)
__device__ __forceinline__ int nearest_idx(int i, int dst, int src, float scale) {
  if (dst == src) return i;
  if (dst == 2 * src) return i >> 1;
  return min((int)floorf((float)i * scale), src - 1);
}

// out = uint8(float(alpha) * map): truncation; NaN (a frame of one single chroma: 0 / 0 above) casts to 0
__device__ __forceinline__ unsigned cc_mul(unsigned a, float m) {
  const float v = u8_to_f32((int)a) * m;
  return v == v ? (unsigned)f32_trunc_nonneg(v) : 0u;
}

// 16 pixels per thread (128-bit loads and stores) when the row allows it, else 4 or pixel by pixel
__global__ void __launch_bounds__(CC_THREADS) cc_apply_kernel(const uint8_t* __restrict__ alpha, const float* __restrict__ map, int h, int w, int th,
                                                              int tw, uint8_t* __restrict__ out, int vec) {
  const int f = blockIdx.z, y = blockIdx.y;
  const float ys = __fdiv_rn((float)th, (float)h), xs = __fdiv_rn((float)tw, (float)w);
  const float* mrow = map + ((int64_t)f * th + nearest_idx(y, h, th, ys)) * tw;
  const int64_t base = ((int64_t)f * h + y) * w;
  if (vec == 2) {          // w % 16 == 0, w == 2 * tw: eight map values per 16 pixels
    for (int x = 16 * (blockIdx.x * CC_THREADS + threadIdx.x); x < w; x += 16 * gridDim.x * CC_THREADS) {
      const uint4 a = ldg_stream16(alpha + base + x);
      if ((a.x | a.y | a.z | a.w) == 0) {          // 0 * anything (NaN included) ends as 0
        stg_stream16(out + base + x, a);
        continue;
      }
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(mrow + (x >> 1))), p1 = __ldg(reinterpret_cast<const float4*>(mrow + (x >> 1) + 4));
      const float m[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      const unsigned aw[4] = {a.x, a.y, a.z, a.w};
      unsigned r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        r[k] = cc_mul(aw[k] & 255u, m[2 * k]) | (cc_mul((aw[k] >> 8) & 255u, m[2 * k]) << 8) | (cc_mul((aw[k] >> 16) & 255u, m[2 * k + 1]) << 16) |
               (cc_mul(aw[k] >> 24, m[2 * k + 1]) << 24);
      stg_stream16(out + base + x, make_uint4(r[0], r[1], r[2], r[3]));
    }
    return;
  }
  const int groups = (w + 3) / 4;
  for (int g = blockIdx.x * CC_THREADS + threadIdx.x; g < groups; g += gridDim.x * CC_THREADS) {
    const int x = 4 * g;
    if (vec && x + 3 < w) {
      const unsigned a = __ldg(reinterpret_cast<const unsigned*>(alpha + base + x));
      float m[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = __ldg(mrow + nearest_idx(x + k, w, tw, xs));
      const unsigned r = cc_mul(a & 255u, m[0]) | (cc_mul((a >> 8) & 255u, m[1]) << 8) | (cc_mul((a >> 16) & 255u, m[2]) << 16) |
                         (cc_mul(a >> 24, m[3]) << 24);
      *reinterpret_cast<unsigned*>(out + base + x) = r;
    } else {
      for (int k = 0; k < 4 && x + k < w; ++k) out[base + x + k] = (uint8_t)cc_mul(alpha[base + x + k], __ldg(mrow + nearest_idx(x + k, w, tw, xs)));
    }
  }
}

}  // namespace
}  // namespace vu

using namespace vu;

namespace {

size_t cc_map_bytes(int n, int th, int tw) { return ((size_t)n * th * tw * sizeof(float) + 255) / 256 * 256; }
size_t cc_stats_bytes(int n) { return ((size_t)n * sizeof(LabStats) + 255) / 256 * 256; }

void cc_background(const uint8_t* bg_bgr, float& bg_a, float& bg_b) {
  // the background colour's a / 255, b / 255 (imgprocess.py:286-289)
  int ba, bb;
  lab_ab(bg_bgr[0], bg_bgr[1], bg_bgr[2], h_lab_gamma, h_lab_cbrt, ba, bb);
  bg_a = (float)ba / 255.f;
  bg_b = (float)bb / 255.f;
}

// about `per_sm` CTAs per SM over the whole clip
dim3 cc_grid(int n, int64_t items, int per_sm) {
  int bx = (per_sm * device_sms() + n - 1) / n;
  const int64_t most = (items + CC_THREADS - 1) / CC_THREADS;
  if (bx > most) bx = (int)most;
  if (bx < 1) bx = 1;
  return dim3(bx, n);
}

// everything after the distance map: sums, final map, apply
int cc_tail(float* map, LabStats* stats, const uint8_t* alpha_lo, const uint8_t* alpha, int n, int h, int w, int th, int tw, double mean_exp,
            uint8_t* out, cudaStream_t s) {
  const int64_t per = (int64_t)th * tw;
  // the streaming passes over the working-resolution map: thin CTAs, 4 pixels per load
  const int pvec = (per % 4 == 0 && (reinterpret_cast<uintptr_t>(alpha_lo) & 3) == 0) ? 1 : 0;
  const dim3 sgrid = cc_grid(n, (per + 3) / 4, 16);
  cc_sums_kernel<0, CC_K1><<<sgrid, CC_THREADS, 0, s>>>(map, alpha_lo, per, stats, mean_exp, pvec);
  cc_sums_kernel<CC_K1, CC_K - CC_K1><<<sgrid, CC_THREADS, 0, s>>>(map, alpha_lo, per, stats, mean_exp, pvec);
  cc_final_kernel<<<sgrid, CC_THREADS, 0, s>>>(map, alpha_lo, per, stats, mean_exp, pvec);
  int vec = (w % 4 == 0 && (reinterpret_cast<uintptr_t>(alpha) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) ? 1 : 0;
  if (w % 16 == 0 && w == 2 * tw && (reinterpret_cast<uintptr_t>(alpha) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) vec = 2;
  const int per_thread = vec == 2 ? 16 : 4;
  dim3 agrid((((w + per_thread - 1) / per_thread) + CC_THREADS - 1) / CC_THREADS, h, n);
  cc_apply_kernel<<<agrid, CC_THREADS, 0, s>>>(alpha, map, h, w, th, tw, out, vec);
  note_launch(4);
  return record_cuda(cudaGetLastError());
}

}  // namespace

extern "C" size_t vu_color_correct_workspace_bytes(int n, int th, int tw) {
  if (n <= 0 || th <= 0 || tw <= 0) return 0;
  return cc_map_bytes(n, th, tw) + cc_stats_bytes(n) + ((size_t)n * th * tw + 255) / 256 * 256;   // map, stats, alpha_lo (frames entry)
}

extern "C" int vu_color_correct(const uint8_t* frames_lo, const uint8_t* alpha_lo, const uint8_t* alpha, int n, int h, int w, int th, int tw,
                                const uint8_t* bg_bgr, double mean_exp, uint8_t* out, void* workspace, size_t workspace_bytes,
                                vu_stream_t stream) {
  VU_REQUIRE(frames_lo && alpha_lo && alpha && out && bg_bgr && n >= 0 && h > 0 && w > 0 && th > 0 && tw > 0);
  if (n == 0) return VU_OK;
  VU_REQUIRE(workspace && workspace_bytes >= vu_color_correct_workspace_bytes(n, th, tw));
  if (n > 65535 || h > 65535) return VU_ERR_UNSUPPORTED;
  float* map = static_cast<float*>(workspace);
  LabStats* stats = reinterpret_cast<LabStats*>(static_cast<uint8_t*>(workspace) + cc_map_bytes(n, th, tw));
  float bg_a, bg_b;
  cc_background(bg_bgr, bg_a, bg_b);
  const int64_t per = (int64_t)th * tw;
  cudaStream_t s = S(stream);
  cc_init_kernel<<<(n + CC_THREADS - 1) / CC_THREADS, CC_THREADS, 0, s>>>(stats, n);
  // the Lab kernel copies 7 KB of tables per CTA: few, fat CTAs
  cc_lab_dist_kernel<<<cc_grid(n, per, 4), CC_THREADS, 0, s>>>(frames_lo, per, bg_a, bg_b, map, stats);
  note_launch(2);
  return cc_tail(map, stats, alpha_lo, alpha, n, h, w, th, tw, mean_exp, out, s);
}

extern "C" int vu_color_correct_frames(const uint8_t* frames, const uint8_t* alpha, int n, int h, int w, int th, int tw, const uint8_t* bg_bgr,
                                       double mean_exp, uint8_t* out, void* workspace, size_t workspace_bytes, vu_stream_t stream) {
  VU_REQUIRE(frames && alpha && out && bg_bgr && n >= 0 && h > 0 && w > 0 && th > 0 && tw > 0);
  if (n == 0) return VU_OK;
  VU_REQUIRE(workspace && workspace_bytes >= vu_color_correct_workspace_bytes(n, th, tw));
  const int scale = (h == 2 * th && w == 2 * tw) ? 2 : ((h == 4 * th && w == 4 * tw) ? 4 : 0);
  if (!scale || w % 16 != 0 || n > 65535 || h > 65535 || (reinterpret_cast<uintptr_t>(frames) & 15) || (reinterpret_cast<uintptr_t>(alpha) & 15))
    return VU_ERR_UNSUPPORTED;   // the caller resizes (vu_resize_linear_u8) and calls vu_color_correct
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  float* map = reinterpret_cast<float*>(wsb);
  LabStats* stats = reinterpret_cast<LabStats*>(wsb + cc_map_bytes(n, th, tw));
  uint8_t* alpha_lo = wsb + cc_map_bytes(n, th, tw) + cc_stats_bytes(n);
  float bg_a, bg_b;
  cc_background(bg_bgr, bg_a, bg_b);
  cudaStream_t s = S(stream);
  cc_init_kernel<<<(n + CC_THREADS - 1) / CC_THREADS, CC_THREADS, 0, s>>>(stats, n);
  const int64_t items = (int64_t)th * (tw / (16 / scale));
  if (scale == 2) cc_lab_dist_lowres_kernel<2><<<cc_grid(n, items, 8), CC_THREADS, 0, s>>>(frames, alpha, h, w, th, tw, bg_a, bg_b, map, alpha_lo, stats);
  else cc_lab_dist_lowres_kernel<4><<<cc_grid(n, items, 8), CC_THREADS, 0, s>>>(frames, alpha, h, w, th, tw, bg_a, bg_b, map, alpha_lo, stats);
  note_launch(2);
  return cc_tail(map, stats, alpha_lo, alpha, n, h, w, th, tw, mean_exp, out, s);
}
