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
constexpr int MAXK = 15;

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

namespace {
constexpr size_t SMEM_LIMIT = 200 * 1024;
// iterations one launch can chain: the staged tile (output tile + a halo of iterations * reach) twice in shared memory
int iters_per_launch(int ksize, int iters) {
  const int a = ksize / 2, b = ksize - 1 - a;
  int k = iters;
  while (k > 1 && 2 * (size_t)(TW + k * (a + b)) * (TH + k * (a + b)) > SMEM_LIMIT) --k;
  return k;
}
template <bool DILATE>
int launch_morph(const uint8_t* src, uint8_t* dst, int n, int h, int w, const SE& se, int iters, cudaStream_t st) {
  const int reach_lo = se.anchor, reach_hi = se.k - 1 - se.anchor;
  const int halo_lo = iters * reach_lo, halo_hi = iters * reach_hi;
  const size_t smem = 2 * (size_t)(TW + halo_lo + halo_hi) * (TH + halo_lo + halo_hi);
  if (smem > SMEM_LIMIT) return VU_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    int e = record_cuda(cudaFuncSetAttribute(morph_kernel<DILATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (e) return e;
  }
  for (int n0 = 0; n0 < n; n0 += 65535) {   // grid.z
    const int nn = n - n0 < 65535 ? n - n0 : 65535;
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH, nn);
    morph_kernel<DILATE><<<grid, THREADS, smem, st>>>(src + (int64_t)n0 * h * w, dst + (int64_t)n0 * h * w, h, w, se, iters, halo_lo, halo_hi);
  }
  note_launch();
  return record_cuda(cudaGetLastError());
}
}  // namespace

// bytes of scratch vu_morph_u8 needs: none while all iterations fit one launch (every combination the reference uses),
// else one image-sized buffer for the ping-pong between launches
extern "C" size_t vu_morph_workspace_bytes(int n, int h, int w, int ksize, int iters) {
  if (n <= 0 || h <= 0 || w <= 0 || ksize < 1 || ksize > MAXK || iters <= 0) return 0;
  return iters_per_launch(ksize, iters) < iters ? (size_t)n * h * w : 0;
}

extern "C" int vu_morph_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int ksize, int iters, int op,
                           void* workspace, size_t workspace_bytes, vu_stream_t stream) {
  VU_REQUIRE(src && dst && n >= 0 && h > 0 && w > 0 && iters >= 0);
  VU_REQUIRE(op == VU_DILATE || op == VU_ERODE);
  if (ksize < 1 || ksize > MAXK) return VU_ERR_UNSUPPORTED;
  if (n == 0) return VU_OK;
  const SE se = make_se(ksize);
  cudaStream_t st = S(stream);
  const int per = iters > 0 ? iters_per_launch(ksize, iters) : 0;
  if (per >= iters) return op == VU_DILATE ? launch_morph<true>(src, dst, n, h, w, se, iters, st) : launch_morph<false>(src, dst, n, h, w, se, iters, st);
  // more iterations than one launch can stage: several launches (iterated morphology composes exactly: cells outside the
  // image are ignored by every pass either way), ping-ponging between dst and the workspace so that the last one lands in dst
  if (!workspace || workspace_bytes < (size_t)n * h * w) return VU_ERR_WORKSPACE;
  uint8_t* tmp = static_cast<uint8_t*>(workspace);
  const int launches = (iters + per - 1) / per;
  const uint8_t* in = src;
  int left = iters;
  for (int l = 0; l < launches; ++l) {
    uint8_t* out = ((launches - 1 - l) % 2 == 0) ? dst : tmp;
    const int k = left < per ? left : per;
    const int e = op == VU_DILATE ? launch_morph<true>(in, out, n, h, w, se, k, st) : launch_morph<false>(in, out, n, h, w, se, k, st);
    if (e) return e;
    in = out;
    left -= k;
  }
  return VU_OK;
}
