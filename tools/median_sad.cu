// Development micro-benchmark: exact temporal median by binary search on the
// value with VABSDIFF4 (sum of absolute differences) as the counting
// primitive, data resident in registers.  count(x <= m) = (S(m+1) - S(m) + N)/2
// with S(m) = sum_f |x_f - m|.  See DESIGN.md "temporal median".
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned sad_acc(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int SPLIT, int G, int CTAS>
__global__ void __launch_bounds__(128, CTAS) median_sad_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, int n, long long m, int nseg) {
  constexpr int LPS = 32 / SPLIT;          // lanes per frame-part
  constexpr int SEG = LPS * 4;             // bytes of a frame one warp owns
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = lane / LPS, li = lane % LPS;
  const int seg = blockIdx.x * 4 + warp;
  if (seg >= nseg) return;
  // frames of this part: contiguous range, sizes differ by at most one
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
  // 4x4 byte transposes: d[4g+j] <- 4 frames of px-ch j
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
  const int ntot = SPLIT * 4 * G;          // slots per px-ch, pads are 255
  const int ta = ((n - 1) >> 1) + 1, tb = (n >> 1) + 1;
  unsigned ma[4] = {0, 0, 0, 0}, mb[4] = {0, 0, 0, 0};
  bool diverged = false;
#pragma unroll 1
  for (int bit = 7; bit >= 0; --bit) {
    const unsigned step = 1u << bit;
    int ca[4], cb[4];
    {
      unsigned s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, q0[4], q1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { q0[j] = (ma[j] + step - 1) * 0x01010101u; q1[j] = q0[j] + 0x01010101u; }
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s0[j] = sad_acc(d[4 * g + j], q0[j], s0[j]); s1[j] = sad_acc(d[4 * g + j], q1[j], s1[j]); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int ds = (int)s1[j] - (int)s0[j];
#pragma unroll
        for (int o = LPS; o < 32; o <<= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
        ca[j] = (ds + ntot) >> 1;
      }
    }
    if (!diverged) {
#pragma unroll
      for (int j = 0; j < 4; ++j) cb[j] = ca[j];
    } else {
      unsigned s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, q0[4], q1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { q0[j] = (mb[j] + step - 1) * 0x01010101u; q1[j] = q0[j] + 0x01010101u; }
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s0[j] = sad_acc(d[4 * g + j], q0[j], s0[j]); s1[j] = sad_acc(d[4 * g + j], q1[j], s1[j]); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int ds = (int)s1[j] - (int)s0[j];
#pragma unroll
        for (int o = LPS; o < 32; o <<= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
        cb[j] = (ds + ntot) >> 1;
      }
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

template <int SPLIT, int G, int CTAS>
void run(const char* name, const uint8_t* d_frames, uint8_t* d_out, int n, long long m, const std::vector<uint8_t>& ref, long long ncheck) {
  constexpr int SEG = (32 / SPLIT) * 4;
  int nseg = (int)(m / SEG);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemset(d_out, 0xEE, m));
  int grid = (nseg + 3) / 4;
  for (int i = 0; i < 2; ++i) median_sad_kernel<SPLIT, G, CTAS><<<grid, 128>>>(d_frames, d_out, n, m, nseg);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  float best = 1e30f, sum = 0; const int reps = 5;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(e0));
    median_sad_kernel<SPLIT, G, CTAS><<<grid, 128>>>(d_frames, d_out, n, m, nseg);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms); sum += ms;
  }
  std::vector<uint8_t> got(ncheck);
  CK(cudaMemcpy(got.data(), d_out, ncheck, cudaMemcpyDeviceToHost));
  long long bad = 0; for (long long i = 0; i < ncheck; ++i) bad += got[i] != ref[i];
  printf("%-12s SPLIT=%d G=%d CTAs/SM=%d n=%d  best %.3f ms avg %.3f ms  %.0f GB/s  mismatches=%lld\n", name, SPLIT, G, CTAS, n, best, sum / reps,
         (double)(n + 1) * nseg * SEG / best / 1e6, bad);
  fflush(stdout);
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 300;
  int H = argc > 2 ? atoi(argv[2]) : 1080, W = argc > 3 ? atoi(argv[3]) : 1920;
  int mode = argc > 4 ? atoi(argv[4]) : 0;   // 0 noisy background, 1 uniform random, 2 two-valued
  long long m = (long long)H * W * 3;
  std::vector<uint8_t> h((size_t)n * m);
  uint32_t x = 12345;
  std::vector<uint8_t> basev(m);
  for (long long i = 0; i < m; ++i) { x = x * 1664525u + 1013904223u; basev[i] = x >> 24; }
  for (int f = 0; f < n; ++f)
    for (long long i = 0; i < m; ++i) {
      x = x * 1664525u + 1013904223u;
      int v;
      if (mode == 0) { v = basev[i] + (int)((x >> 16) % 13) - 6; if (((x >> 8) & 63) == 0) v = (x >> 20) & 255; }
      else if (mode == 1) v = x >> 24;
      else v = ((x >> 13) & 1) ? 255 : 0;
      h[(size_t)f * m + i] = (uint8_t)std::min(255, std::max(0, v));
    }
  const long long ncheck = 1 << 16;
  std::vector<uint8_t> ref(ncheck), col(n);
  for (long long i = 0; i < ncheck; ++i) {
    for (int f = 0; f < n; ++f) col[f] = h[(size_t)f * m + i];
    std::sort(col.begin(), col.end());
    ref[i] = (uint8_t)((col[(n - 1) / 2] + col[n / 2]) >> 1);
  }
  uint8_t *d_frames, *d_out;
  CK(cudaMalloc(&d_frames, (size_t)n * m)); CK(cudaMalloc(&d_out, m));
  CK(cudaMemcpy(d_frames, h.data(), (size_t)n * m, cudaMemcpyHostToDevice));
  printf("n=%d m=%lld mode=%d\n", n, m, mode);
  if (n <= 304) { run<2, 38, 2>("sad", d_frames, d_out, n, m, ref, ncheck); run<4, 19, 4>("sad", d_frames, d_out, n, m, ref, ncheck); run<4, 19, 3>("sad", d_frames, d_out, n, m, ref, ncheck); }
  if (n <= 152) { run<2, 19, 4>("sad", d_frames, d_out, n, m, ref, ncheck); run<1, 38, 2>("sad", d_frames, d_out, n, m, ref, ncheck); }
  return 0;
}
