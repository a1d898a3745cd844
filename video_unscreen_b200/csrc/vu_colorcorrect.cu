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

__device__ __forceinline__ float cc_normalise(float d, float lo, float hi) { return __fdiv_rn(d - lo, hi - lo); }

__global__ void __launch_bounds__(CC_THREADS) cc_sums_kernel(const float* __restrict__ dist, const uint8_t* __restrict__ alpha_lo, int64_t per,
                                                             LabStats* __restrict__ stats) {
  __shared__ double part[CC_THREADS / 32][CC_K];
  __shared__ unsigned cpart[CC_THREADS / 32];
  const int f = blockIdx.y;
  const float lo = __uint_as_float(stats[f].lo), hi = __uint_as_float(stats[f].hi);
  const float* d = dist + (int64_t)f * per;
  const uint8_t* al = alpha_lo + (int64_t)f * per;
  double s[CC_K];
#pragma unroll
  for (int k = 0; k < CC_K; ++k) s[k] = 0.0;
  unsigned cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * CC_THREADS + threadIdx.x; i < per; i += (int64_t)gridDim.x * CC_THREADS) {
    float v = cc_normalise(d[i], lo, hi);
    if (al[i] > 0 && v > 0.f) {
      ++cnt;
#pragma unroll
      for (int k = 0; k < CC_K; ++k) {
        s[k] += (double)v;
        v = sqrtf(v);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < CC_K; ++k) {
    double v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) part[warp][k] = v;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) cpart[warp] = cnt;
  __syncthreads();
  if (threadIdx.x < CC_K) {
    double v = 0.0;
    for (int wv = 0; wv < CC_THREADS / 32; ++wv) v += part[wv][threadIdx.x];
    if (v != 0.0) atomicAdd(&stats[f].sums[threadIdx.x], v);
  }
  if (threadIdx.x == 32) {
    unsigned c = 0;
    for (int wv = 0; wv < CC_THREADS / 32; ++wv) c += cpart[wv];
    if (c) atomicAdd(&stats[f].count, (unsigned long long)c);
  }
}

__global__ void __launch_bounds__(CC_THREADS) cc_final_kernel(float* __restrict__ dist, const uint8_t* __restrict__ alpha_lo, int64_t per,
                                                              const LabStats* __restrict__ stats, double mean_exp) {
  const int f = blockIdx.y;
  const LabStats& st = stats[f];
  const float lo = __uint_as_float(st.lo), hi = __uint_as_float(st.hi);
  // the reference's loop: `while mean < mean_exp: sqrt` (an empty matte has mean NaN: no iteration)
  int iters = 0;
  if (st.count) {
    const double cnt = (double)st.count;
    while (iters < CC_K && st.sums[iters] / cnt < mean_exp) ++iters;
  }
  float* d = dist + (int64_t)f * per;
  const uint8_t* al = alpha_lo + (int64_t)f * per;
  for (int64_t i = (int64_t)blockIdx.x * CC_THREADS + threadIdx.x; i < per; i += (int64_t)gridDim.x * CC_THREADS) {
    float v = cc_normalise(d[i], lo, hi);
    for (int k = 0; k < iters; ++k) v = sqrtf(v);
    d[i] = al[i] == 0 ? 0.f : v;
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

__global__ void __launch_bounds__(CC_THREADS) cc_apply_kernel(const uint8_t* __restrict__ alpha, const float* __restrict__ map, int h, int w, int th,
                                                              int tw, uint8_t* __restrict__ out, int vec) {
  const int f = blockIdx.z, y = blockIdx.y;
  const float ys = __fdiv_rn((float)th, (float)h), xs = __fdiv_rn((float)tw, (float)w);
  const float* mrow = map + ((int64_t)f * th + nearest_idx(y, h, th, ys)) * tw;
  const int64_t base = ((int64_t)f * h + y) * w;
  const int groups = (w + 3) / 4;
  for (int g = blockIdx.x * CC_THREADS + threadIdx.x; g < groups; g += gridDim.x * CC_THREADS) {
    const int x = 4 * g;
    if (vec && x + 3 < w) {
      const unsigned a = __ldg(reinterpret_cast<const unsigned*>(alpha + base + x));
      float m[4];
      if (w == 2 * tw) {
        const float2 p = __ldg(reinterpret_cast<const float2*>(mrow + (x >> 1)));
        m[0] = m[1] = p.x;
        m[2] = m[3] = p.y;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = __ldg(mrow + nearest_idx(x + k, w, tw, xs));
      }
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

extern "C" size_t vu_color_correct_workspace_bytes(int n, int th, int tw) {
  if (n <= 0 || th <= 0 || tw <= 0) return 0;
  const size_t map = ((size_t)n * th * tw * sizeof(float) + 255) / 256 * 256;
  return map + (size_t)n * sizeof(LabStats);
}

extern "C" int vu_color_correct(const uint8_t* frames_lo, const uint8_t* alpha_lo, const uint8_t* alpha, int n, int h, int w, int th, int tw,
                                const uint8_t* bg_bgr, double mean_exp, uint8_t* out, void* workspace, size_t workspace_bytes,
                                vu_stream_t stream) {
  VU_REQUIRE(frames_lo && alpha_lo && alpha && out && bg_bgr && n >= 0 && h > 0 && w > 0 && th > 0 && tw > 0);
  if (n == 0) return VU_OK;
  VU_REQUIRE(workspace && workspace_bytes >= vu_color_correct_workspace_bytes(n, th, tw));
  if (n > 65535 || h > 65535) return VU_ERR_UNSUPPORTED;
  float* map = static_cast<float*>(workspace);
  LabStats* stats = reinterpret_cast<LabStats*>(static_cast<uint8_t*>(workspace) + ((size_t)n * th * tw * sizeof(float) + 255) / 256 * 256);
  // the background colour's a / 255, b / 255 (imgprocess.py:286-289)
  int ba, bb;
  lab_ab(bg_bgr[0], bg_bgr[1], bg_bgr[2], h_lab_gamma, h_lab_cbrt, ba, bb);
  const float bg_a = (float)ba / 255.f, bg_b = (float)bb / 255.f;
  const int64_t per = (int64_t)th * tw;
  cudaStream_t s = S(stream);
  cc_init_kernel<<<(n + CC_THREADS - 1) / CC_THREADS, CC_THREADS, 0, s>>>(stats, n);
  // about four CTAs per SM over the whole clip (the Lab kernel copies 7 KB of tables per CTA: few, fat CTAs)
  int bx = (4 * device_sms() + n - 1) / n;
  const int64_t most = (per + CC_THREADS - 1) / CC_THREADS;
  if (bx > most) bx = (int)most;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
  cc_lab_dist_kernel<<<grid, CC_THREADS, 0, s>>>(frames_lo, per, bg_a, bg_b, map, stats);
  cc_sums_kernel<<<grid, CC_THREADS, 0, s>>>(map, alpha_lo, per, stats);
  cc_final_kernel<<<grid, CC_THREADS, 0, s>>>(map, alpha_lo, per, stats, mean_exp);
  const int vec = (w % 4 == 0 && tw % 2 == 0 && (reinterpret_cast<uintptr_t>(alpha) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) ? 1 : 0;
  dim3 agrid((((w + 3) / 4) + CC_THREADS - 1) / CC_THREADS, h, n);
  cc_apply_kernel<<<agrid, CC_THREADS, 0, s>>>(alpha, map, h, w, th, tw, out, vec);
  note_launch(5);
  return record_cuda(cudaGetLastError());
}
