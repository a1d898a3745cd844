// Single-image background inpainting of BackgroundAgent (unscreen/bgmodel/agent.py):
// the pieces that are not already kernels of the hot path.
//
//   vu_mask_bbox     bounding box of mask > 0 (get_fgbox, utils/maskprocess.py:37-53)
//   vu_masked_sum3   per-channel sums of an image over mask > 0 and the pixel count (get_mean_bg, agent.py:80-88)
//   vu_pcov_round    one round of get_bg_by_pcov (agent.py:118-129): normalised k x k box filter (cv2.boxFilter,
//                    BORDER_REFLECT_101) of the image and of the validity map, mean / validity where the window saw a
//                    valid pixel.  Rounds are enqueued back to back; a round that finds the previous one left no invalid
//                    pixel returns at once, so the host only has to look at the flags every few rounds.
#include <climits>

#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int BT = 256;

__global__ void __launch_bounds__(BT) mask_bbox_kernel(const uint8_t* __restrict__ mask, int h, int w, int* __restrict__ out4) {
  int r0 = INT_MAX, r1 = -1, c0 = INT_MAX, c1 = -1;
  const int64_t total = (int64_t)h * w;
  for (int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x; i < total; i += (int64_t)gridDim.x * BT) {
    if (__ldg(mask + i)) {
      const int y = (int)(i / w), x = (int)(i - (int64_t)y * w);
      r0 = min(r0, y); r1 = max(r1, y); c0 = min(c0, x); c1 = max(c1, x);
    }
  }
  r0 = __reduce_min_sync(0xffffffffu, r0); c0 = __reduce_min_sync(0xffffffffu, c0);
  r1 = __reduce_max_sync(0xffffffffu, r1); c1 = __reduce_max_sync(0xffffffffu, c1);
  if ((threadIdx.x & 31) == 0 && r1 >= 0) {
    atomicMin(out4 + 0, r0); atomicMax(out4 + 1, r1); atomicMin(out4 + 2, c0); atomicMax(out4 + 3, c1);
  }
}

__global__ void __launch_bounds__(BT) masked_sum3_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask, int64_t npix,
                                                         unsigned long long* __restrict__ out4) {
  unsigned long long s0 = 0, s1 = 0, s2 = 0, c = 0;
  for (int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x; i < npix; i += (int64_t)gridDim.x * BT) {
    if (!mask || __ldg(mask + i)) {
      s0 += __ldg(img + 3 * i); s1 += __ldg(img + 3 * i + 1); s2 += __ldg(img + 3 * i + 2); ++c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_down_sync(0xffffffffu, s0, o); s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o); c += __shfl_down_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0 && c) {
    atomicAdd(out4 + 0, s0); atomicAdd(out4 + 1, s1); atomicAdd(out4 + 2, s2); atomicAdd(out4 + 3, c);
  }
}

__device__ __forceinline__ int reflect101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// FIRST: the inputs are the (resized) image and the dilated mask themselves, addressed with the image's row pitch:
// hole pixels (mask > 0) count as 0 / invalid (agent.py:109-110)
template <bool FIRST>
__global__ void __launch_bounds__(BT) pcov_round_kernel(const uint8_t* __restrict__ img_in, const uint8_t* __restrict__ valid_in, int in_pitch_px,
                                                        int rh, int rw, int k, uint8_t* __restrict__ img_out, uint8_t* __restrict__ valid_out,
                                                        unsigned* __restrict__ flags, int round) {
  if (round > 0 && flags[round] == 0) return;   // the round before left no invalid pixel: done (agent.py:127-128)
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * (BT / 32) + (threadIdx.x >> 5);
  bool invalid = false;
  if (x < rw && y < rh) {
    const int r0 = k / 2;
    int s0 = 0, s1 = 0, s2 = 0, cnt = 0;
    for (int dy = 0; dy < k; ++dy) {
      const int yy = reflect101(y + dy - r0, rh);
      for (int dx = 0; dx < k; ++dx) {
        const int xx = reflect101(x + dx - r0, rw);
        const int64_t o = (int64_t)yy * in_pitch_px + xx;
        const bool ok = FIRST ? (__ldg(valid_in + o) == 0) : (__ldg(valid_in + o) != 0);
        if (!FIRST || ok) {
          s0 += __ldg(img_in + 3 * o); s1 += __ldg(img_in + 3 * o + 1); s2 += __ldg(img_in + 3 * o + 2);
        }
        cnt += ok;
      }
    }
    const int kk = k * k;
    int m0 = (2 * s0 + kk) / (2 * kk), m1 = (2 * s1 + kk) / (2 * kk), m2 = (2 * s2 + kk) / (2 * kk);   // cv2.boxFilter: rounded mean, never a tie
    if (cnt > 0) {
      const double c = __dmul_rn((double)cnt, 1.0 / (double)kk);   // the float64 box filter of the 0/1 validity map
      m0 = (int)fmin(fmax(__ddiv_rn((double)m0, c), 0.0), 255.0);
      m1 = (int)fmin(fmax(__ddiv_rn((double)m1, c), 0.0), 255.0);
      m2 = (int)fmin(fmax(__ddiv_rn((double)m2, c), 0.0), 255.0);
    } else {
      invalid = true;
    }
    const int64_t o = (int64_t)y * rw + x;
    img_out[3 * o] = (uint8_t)m0; img_out[3 * o + 1] = (uint8_t)m1; img_out[3 * o + 2] = (uint8_t)m2;
    valid_out[o] = cnt > 0;
  }
  if (__any_sync(0xffffffffu, invalid) && (threadIdx.x & 31) == 0) atomicOr(flags + round + 1, 1u);
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" int vu_mask_bbox(const uint8_t* mask, int h, int w, int32_t* out4, vu_stream_t stream) {
  VU_REQUIRE(mask && out4 && h > 0 && w > 0);
  const int init[4] = {INT_MAX, -1, INT_MAX, -1};
  int e = record_cuda(cudaMemcpyAsync(out4, init, sizeof(init), cudaMemcpyHostToDevice, S(stream)));
  if (e) return e;
  mask_bbox_kernel<<<grid_for((int64_t)h * w, BT, 4), BT, 0, S(stream)>>>(mask, h, w, out4);
  VU_RETURN_LAUNCH();
}

extern "C" int vu_masked_sum3(const uint8_t* img, const uint8_t* mask, int64_t npix, uint64_t* out4, vu_stream_t stream) {
  VU_REQUIRE(img && out4 && npix >= 0);
  int e = record_cuda(cudaMemsetAsync(out4, 0, 4 * sizeof(uint64_t), S(stream)));
  if (e) return e;
  if (npix == 0) return VU_OK;
  masked_sum3_kernel<<<grid_for(npix, BT, 4), BT, 0, S(stream)>>>(img, mask, npix, reinterpret_cast<unsigned long long*>(out4));
  VU_RETURN_LAUNCH();
}

extern "C" int vu_pcov_round(const uint8_t* img_in, const uint8_t* valid_in, int in_pitch_px, int first, int rh, int rw, int ksize, uint8_t* img_out,
                             uint8_t* valid_out, uint32_t* flags, int round, vu_stream_t stream) {
  VU_REQUIRE(img_in && valid_in && img_out && valid_out && flags && rh > 0 && rw > 0 && in_pitch_px >= rw && round >= 0 && round < 100);
  if (ksize < 1 || ksize > 15 || !(ksize & 1) || rh < ksize || rw < ksize) return VU_ERR_UNSUPPORTED;
  dim3 grid((rw + 31) / 32, (rh + BT / 32 - 1) / (BT / 32));
  if (first) pcov_round_kernel<true><<<grid, BT, 0, S(stream)>>>(img_in, valid_in, in_pitch_px, rh, rw, ksize, img_out, valid_out, flags, round);
  else pcov_round_kernel<false><<<grid, BT, 0, S(stream)>>>(img_in, valid_in, in_pitch_px, rh, rw, ksize, img_out, valid_out, flags, round);
  VU_RETURN_LAUNCH();
}
