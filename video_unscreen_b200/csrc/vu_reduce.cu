// Per-frame integer reductions and mask algebra around the trimap / colour
// filtering branches: exist_foreground (maskprocess.py:56-60), the early-outs
// of colorfiltering/agent.py:303-307, the fuzzy-area ratio and the two masked
// assignments of trimap/agent.py:88-100, postprocess' adaptive threshold
// (colorfiltering/agent.py:277-280).  Counts are exact 64-bit integers.
#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ unsigned long long block_sum(unsigned long long v) {
  __shared__ unsigned long long warp_sums[THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
  __syncthreads();
  v = 0;
  if (threadIdx.x < THREADS / 32) v = warp_sums[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

__device__ __forceinline__ bool cmp(int op, int a, int thr) {
  switch (op) {
    case VU_CMP_GE: return a >= thr;
    case VU_CMP_GT: return a > thr;
    case VU_CMP_LT: return a < thr;
    case VU_CMP_EQ: return a == thr;
    default: return a != thr;
  }
}

__global__ void __launch_bounds__(THREADS) count_cmp_kernel(const uint8_t* __restrict__ src, int64_t per_item, int op, int thr,
                                                            unsigned long long* counts) {
  const uint8_t* p = src + (int64_t)blockIdx.y * per_item;
  unsigned long long c = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_item; i += stride) c += cmp(op, __ldg(p + i), thr);
  c = block_sum(c);
  if (threadIdx.x == 0 && c) atomicAdd(counts + blockIdx.y, c);
}

__global__ void __launch_bounds__(THREADS) count_and_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int64_t per_item,
                                                            unsigned long long* counts2) {
  const int64_t base = (int64_t)blockIdx.y * per_item;
  unsigned long long both = 0, pos = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_item; i += stride) {
    const int av = __ldg(a + base + i), bv = __ldg(b + base + i);
    pos += av > 0;
    both += (av > 0) && (bv > 0);
  }
  both = block_sum(both);
  pos = block_sum(pos);
  if (threadIdx.x == 0) {
    if (both) atomicAdd(counts2 + 2 * blockIdx.y, both);
    if (pos) atomicAdd(counts2 + 2 * blockIdx.y + 1, pos);
  }
}

__global__ void __launch_bounds__(THREADS) cf_stats_kernel(const uint8_t* __restrict__ alpha, const uint8_t* __restrict__ mask, int64_t per_item,
                                                           unsigned long long* stats2) {
  const int64_t base = (int64_t)blockIdx.y * per_item;
  unsigned long long sum = 0, cnt = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_item; i += stride) {
    const int av = __ldg(alpha + base + i), mv = __ldg(mask + base + i);
    if (av > 128 && mv > 0) { sum += av; ++cnt; }
  }
  sum = block_sum(sum);
  cnt = block_sum(cnt);
  if (threadIdx.x == 0 && cnt) {
    atomicAdd(stats2 + 2 * blockIdx.y, sum);
    atomicAdd(stats2 + 2 * blockIdx.y + 1, cnt);
  }
}

__global__ void __launch_bounds__(THREADS) cf_apply_kernel(const uint8_t* __restrict__ alpha, int64_t per_item, const unsigned long long* stats2,
                                                           double ratio, uint8_t* __restrict__ out) {
  const int64_t base = (int64_t)blockIdx.y * per_item;
  const unsigned long long sum = stats2[2 * blockIdx.y], cnt = stats2[2 * blockIdx.y + 1];
  // numpy: mean (f64 sum / n) * thr_ratio; an empty selection gives NaN and
  // "alpha < NaN" is false everywhere
  const bool have = cnt != 0;
  const double thr = have ? __dmul_rn(__ddiv_rn((double)sum, (double)cnt), ratio) : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_item; i += stride) {
    const int av = __ldg(alpha + base + i);
    out[base + i] = (have && (double)av < thr) ? 0 : av;
  }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS) mask_where_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out,
                                                             int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const int av = __ldg(a + i), bv = __ldg(b + i);
    int o;
    if (MODE == 0) o = bv ? 0 : av;                 // a[b] = 0
    else if (MODE == 1) o = bv ? 128 : av;          // a[b] = 128
    else if (MODE == 2) o = av * (bv / 255);        // mask * (g // 255)
    else if (MODE == 3) o = (av - bv) & 255;        // uint8 wrap-around subtraction
    else if (MODE == 4) o = (av < 128) ? 0 : (bv > 127 ? 255 : 128);  // a = dilated, b = eroded
    else o = (av > 0 && bv > 0) ? 1 : 0;            // fuzzy area
    out[i] = (uint8_t)o;
  }
}

__global__ void __launch_bounds__(THREADS) snap_kernel(const uint8_t* __restrict__ a, uint8_t* __restrict__ out, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const int v = __ldg(a + i);
    out[i] = (v > 0 && v < 255) ? 128 : v;
  }
}

__global__ void __launch_bounds__(THREADS) binarise_kernel(const uint8_t* __restrict__ a, uint8_t* __restrict__ out, int64_t count, int thr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) out[i] = __ldg(a + i) > thr ? 255 : 0;
}

inline dim3 item_grid(int n, int64_t per_item) {
  int bx = (int)((per_item + THREADS * 16 - 1) / (THREADS * 16));
  int cap = (device_sms() * 8 + n - 1) / n;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3(bx, n);
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_count_cmp_u8(const uint8_t* src, int n, int64_t per_item, int op, int thr, uint64_t* counts, vu_stream_t stream) {
  VU_REQUIRE(src && counts && n >= 0 && per_item >= 0 && op >= VU_CMP_GE && op <= VU_CMP_NE);
  if (n == 0) return VU_OK;
  int e = record_cuda(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * n, S(stream)));
  if (e) return e;
  if (per_item == 0) return VU_OK;
  count_cmp_kernel<<<item_grid(n, per_item), THREADS, 0, S(stream)>>>(src, per_item, op, thr, reinterpret_cast<unsigned long long*>(counts));
  VU_RETURN_LAUNCH();
}

extern "C" int vu_count_and_u8(const uint8_t* a, const uint8_t* b, int n, int64_t per_item, uint64_t* counts2, vu_stream_t stream) {
  VU_REQUIRE(a && b && counts2 && n >= 0 && per_item >= 0);
  if (n == 0) return VU_OK;
  int e = record_cuda(cudaMemsetAsync(counts2, 0, sizeof(uint64_t) * 2 * n, S(stream)));
  if (e) return e;
  if (per_item == 0) return VU_OK;
  count_and_kernel<<<item_grid(n, per_item), THREADS, 0, S(stream)>>>(a, b, per_item, reinterpret_cast<unsigned long long*>(counts2));
  VU_RETURN_LAUNCH();
}

extern "C" int vu_cf_threshold_stats(const uint8_t* alpha, const uint8_t* mask, int n, int64_t per_item, uint64_t* stats2,
                                     vu_stream_t stream) {
  VU_REQUIRE(alpha && mask && stats2 && n >= 0 && per_item >= 0);
  if (n == 0) return VU_OK;
  int e = record_cuda(cudaMemsetAsync(stats2, 0, sizeof(uint64_t) * 2 * n, S(stream)));
  if (e) return e;
  if (per_item == 0) return VU_OK;
  cf_stats_kernel<<<item_grid(n, per_item), THREADS, 0, S(stream)>>>(alpha, mask, per_item, reinterpret_cast<unsigned long long*>(stats2));
  VU_RETURN_LAUNCH();
}

extern "C" int vu_cf_threshold_apply(const uint8_t* alpha, int n, int64_t per_item, const uint64_t* stats2, double thr_ratio,
                                     uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(alpha && out && stats2 && n >= 0 && per_item >= 0);
  if (n == 0 || per_item == 0) return VU_OK;
  cf_apply_kernel<<<item_grid(n, per_item), THREADS, 0, S(stream)>>>(alpha, per_item, reinterpret_cast<const unsigned long long*>(stats2), thr_ratio, out);
  VU_RETURN_LAUNCH();
}

#define VU_WHERE(NAME, MODE)                                                                                              \
  extern "C" int NAME(const uint8_t* a, const uint8_t* b, uint8_t* out, int64_t count, vu_stream_t stream) {              \
    VU_REQUIRE(a && b && out && count >= 0);                                                                              \
    if (count == 0) return VU_OK;                                                                                         \
    mask_where_kernel<MODE><<<grid_for(count, THREADS * 4, 8), THREADS, 0, S(stream)>>>(a, b, out, count);                \
    VU_RETURN_LAUNCH();                                                                                                   \
  }
VU_WHERE(vu_mask_clear_where, 0)
VU_WHERE(vu_mask_set128_where, 1)
VU_WHERE(vu_trimap_classify, 4)
VU_WHERE(vu_mask_and01, 5)
#undef VU_WHERE

extern "C" int vu_gate(const uint8_t* mask, const uint8_t* g, int64_t count, uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(mask && g && out && count >= 0);
  if (count == 0) return VU_OK;
  mask_where_kernel<2><<<grid_for(count, THREADS * 4, 8), THREADS, 0, S(stream)>>>(mask, g, out, count);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_sub_wrap_u8(const uint8_t* a, const uint8_t* b, int64_t count, uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(a && b && out && count >= 0);
  if (count == 0) return VU_OK;
  mask_where_kernel<3><<<grid_for(count, THREADS * 4, 8), THREADS, 0, S(stream)>>>(a, b, out, count);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_binarise(const uint8_t* alpha, int64_t count, int thr, uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(alpha && out && count >= 0);
  if (count == 0) return VU_OK;
  binarise_kernel<<<grid_for(count, THREADS * 4, 8), THREADS, 0, S(stream)>>>(alpha, out, count, thr);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_trimap_snap(const uint8_t* a, int64_t count, uint8_t* out, vu_stream_t stream) {
  VU_REQUIRE(a && out && count >= 0);
  if (count == 0) return VU_OK;
  snap_kernel<<<grid_for(count, THREADS * 4, 8), THREADS, 0, S(stream)>>>(a, out, count);
  VU_RETURN_LAUNCH();
}

// ---- per-frame branches without a host round trip ----------------------------
namespace vu {
namespace {
// flags[i] = 1 when the frame takes the "trust the mask" branch of
// generate_trimap_withbg (trimap/agent.py:94): float(fuzzy)/pos > thr, or no
// positive pixel at all (:88, the mask is returned as is)
__global__ void ratio_flags_kernel(const unsigned long long* counts2, int n, double thr, uint8_t* flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long fz = counts2[2 * i], pos = counts2[2 * i + 1];
  flags[i] = (pos == 0) ? 2 : ((__ddiv_rn((double)fz, (double)pos) > thr) ? 1 : 0);
}
// degenerate-mask early-outs of ColorFilteringAgent.forward (agent.py:303-307):
// flags[i] = 1 "no foreground" (alpha := mask), 2 "no background" (alpha := mask), else 0
__global__ void cf_degenerate_flags_kernel(const unsigned long long* nfg, const unsigned long long* nbg, int n, int stride,
                                           unsigned long long fg_min, unsigned long long bg_min, uint8_t* flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = nfg[(int64_t)i * stride] < fg_min ? 1 : (nbg[(int64_t)i * stride] < bg_min ? 2 : 0);
}
// out[i][:] = flags[i] ? a[i][:] : b[i][:]
__global__ void __launch_bounds__(THREADS) select_frames_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                                const uint8_t* __restrict__ flags, int64_t per_item, uint8_t* __restrict__ out) {
  const int64_t base = (int64_t)blockIdx.y * per_item;
  const bool fa = flags[blockIdx.y] != 0;
  const uint8_t* src = fa ? a : b;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_item; i += stride) out[base + i] = __ldg(src + base + i);
}
// out = (flags[i] == 0 && b != 0) ? 128 : a    (trimap[fuzzy] = 128 only on the ensemble branch)
__global__ void __launch_bounds__(THREADS) set128_unflagged_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                                   const uint8_t* __restrict__ flags, int64_t per_item, uint8_t* __restrict__ out) {
  const int64_t base = (int64_t)blockIdx.y * per_item;
  const bool ens = flags[blockIdx.y] == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_item; i += stride) {
    const int av = __ldg(a + base + i);
    out[base + i] = (ens && __ldg(b + base + i)) ? 128 : av;
  }
}
}  // namespace
}  // namespace vu

extern "C" int vu_ratio_flags(const uint64_t* counts2, int n, double thr, uint8_t* flags, vu_stream_t stream) {
  VU_REQUIRE(counts2 && flags && n >= 0);
  if (n == 0) return VU_OK;
  ratio_flags_kernel<<<(n + 127) / 128, 128, 0, S(stream)>>>(reinterpret_cast<const unsigned long long*>(counts2), n, thr, flags);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_cf_degenerate_flags(const uint64_t* nfg, const uint64_t* nbg, int n, int stride, uint64_t fg_min, uint64_t bg_min,
                                      uint8_t* flags, vu_stream_t stream) {
  VU_REQUIRE(nfg && nbg && flags && n >= 0 && stride >= 1);
  if (n == 0) return VU_OK;
  cf_degenerate_flags_kernel<<<(n + 127) / 128, 128, 0, S(stream)>>>(reinterpret_cast<const unsigned long long*>(nfg),
                                                                       reinterpret_cast<const unsigned long long*>(nbg), n, stride, fg_min, bg_min,
                                                                       flags);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_select_frames(const uint8_t* a, const uint8_t* b, const uint8_t* flags, int n, int64_t per_item, uint8_t* out,
                                vu_stream_t stream) {
  VU_REQUIRE(a && b && flags && out && n >= 0 && per_item >= 0);
  if (n == 0 || per_item == 0) return VU_OK;
  select_frames_kernel<<<item_grid(n, per_item), THREADS, 0, S(stream)>>>(a, b, flags, per_item, out);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_set128_unflagged(const uint8_t* a, const uint8_t* b, const uint8_t* flags, int n, int64_t per_item, uint8_t* out,
                                   vu_stream_t stream) {
  VU_REQUIRE(a && b && flags && out && n >= 0 && per_item >= 0);
  if (n == 0 || per_item == 0) return VU_OK;
  set128_unflagged_kernel<<<item_grid(n, per_item), THREADS, 0, S(stream)>>>(a, b, flags, per_item, out);
  VU_RETURN_LAUNCH();
}
