// dilate_mask / erode_mask (unscreen/utils/maskprocess.py:7-34):
// grey-scale max / min under cv2.getStructuringElement(MORPH_ELLIPSE,(k,k)),
// `iterations` sequential applications, taps outside the image ignored
// (SURVEY.md A.1).  All iterations run inside ONE kernel: a CTA stages an
// output tile plus a halo of iters*reach pixels in shared memory and
// ping-pongs the passes there, so the image is read once and written once
// however many iterations are asked for.  Out-of-image cells are re-set to the
// identity after every pass, which is what "ignored" means under iteration.
#include "vu_common.cuh"

namespace vu {
namespace {

constexpr int TW = 64, TH = 32, THREADS = 256;
constexpr int MAXK = 7;

struct SE {
  int k, anchor;
  unsigned rowmask[MAXK];  // bit j of rowmask[i] = tap at (i - anchor, j - anchor)
};

// cv2's ellipse rasterisation (getStructuringElement, MORPH_ELLIPSE)
SE make_se(int k) {
  SE se{};
  se.k = k;
  se.anchor = k / 2;
  const int r = k / 2, c = k / 2;
  const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
  for (int i = 0; i < k; ++i) {
    const int dy = i - r;
    unsigned m = 0;
    if (abs(dy) <= r) {
      const int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));
      const int j1 = (c - dx) > 0 ? (c - dx) : 0;
      const int j2 = (c + dx + 1) < k ? (c + dx + 1) : k;
      for (int j = j1; j < j2; ++j) m |= 1u << j;
    }
    se.rowmask[i] = m;
  }
  return se;
}

template <bool DILATE>
__global__ void __launch_bounds__(THREADS) morph_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w,
                                                        SE se, int iters, int halo_lo, int halo_hi) {
  extern __shared__ uint8_t sm[];
  const int sw = TW + halo_lo + halo_hi, sh = TH + halo_lo + halo_hi;
  uint8_t* buf0 = sm;
  uint8_t* buf1 = sm + sw * sh;
  const int64_t frame = (int64_t)blockIdx.z * h * w;
  const int x0 = blockIdx.x * TW - halo_lo, y0 = blockIdx.y * TH - halo_lo;
  const uint8_t ident = DILATE ? 0 : 255;
  for (int i = threadIdx.x; i < sw * sh; i += THREADS) {
    const int ly = i / sw, lx = i - ly * sw;
    const int gy = y0 + ly, gx = x0 + lx;
    buf0[i] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? __ldg(src + frame + (int64_t)gy * w + gx) : ident;
  }
  __syncthreads();
  uint8_t* in = buf0;
  uint8_t* out = buf1;
  const int a = se.anchor, k = se.k;
  for (int it = 0; it < iters; ++it) {
    for (int i = threadIdx.x; i < sw * sh; i += THREADS) {
      const int ly = i / sw, lx = i - ly * sw;
      const int gy = y0 + ly, gx = x0 + lx;
      int acc = ident;
      if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
        for (int r = 0; r < k; ++r) {
          const int yy = ly + r - a;
          if (yy < 0 || yy >= sh) continue;  // cell in the (discarded) invalid ring
          const unsigned m = se.rowmask[r];
          for (int c = 0; c < k; ++c) {
            const int xx = lx + c - a;
            if (!((m >> c) & 1u) || xx < 0 || xx >= sw) continue;
            const int v = in[yy * sw + xx];
            acc = DILATE ? max(acc, v) : min(acc, v);
          }
        }
      }
      out[i] = (uint8_t)acc;
    }
    __syncthreads();
    uint8_t* t = in; in = out; out = t;
  }
  for (int i = threadIdx.x; i < TW * TH; i += THREADS) {
    const int ty = i / TW, tx = i - ty * TW;
    const int gy = blockIdx.y * TH + ty, gx = blockIdx.x * TW + tx;
    if (gy < h && gx < w) dst[frame + (int64_t)gy * w + gx] = in[(ty + halo_lo) * sw + tx + halo_lo];
  }
}

}  // namespace
}  // namespace vu

using namespace vu;

extern "C" size_t vu_morph_workspace_bytes(int, int, int, int, int) { return 0; }

extern "C" int vu_morph_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int ksize, int iters, int op,
                           void*, size_t, vu_stream_t stream) {
  VU_REQUIRE(src && dst && n >= 0 && h > 0 && w > 0 && iters >= 0);
  VU_REQUIRE(op == VU_DILATE || op == VU_ERODE);
  if (ksize < 1 || ksize > MAXK) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  const SE se = make_se(ksize);
  const int reach_lo = se.anchor, reach_hi = ksize - 1 - se.anchor;
  const int halo_lo = iters * reach_lo, halo_hi = iters * reach_hi;
  const int sw = TW + halo_lo + halo_hi, sh = TH + halo_lo + halo_hi;
  const size_t smem = 2 * (size_t)sw * sh;
  if (smem > 200 * 1024) return VU_ERR_UNSUPPORTED;
  dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH, n);
  if (op == VU_DILATE) {
    if (smem > 48 * 1024) {
      int e = record_cuda(cudaFuncSetAttribute(morph_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (e) return e;
    }
    morph_kernel<true><<<grid, THREADS, smem, S(stream)>>>(src, dst, h, w, se, iters, halo_lo, halo_hi);
  } else {
    if (smem > 48 * 1024) {
      int e = record_cuda(cudaFuncSetAttribute(morph_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (e) return e;
    }
    morph_kernel<false><<<grid, THREADS, smem, S(stream)>>>(src, dst, h, w, se, iters, halo_lo, halo_hi);
  }
  VU_RETURN_LAUNCH();
}
