// Micro-benchmark of shared-memory histogram update strategies for the exact
// temporal median.  Development tool only: it
// answers "what does one histogram increment cost on sm_100a" before the
// product kernel in video_unscreen_b200/csrc/vu_temporal.cu is fixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o median_ubench median_ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int WARPS = 7;
constexpr int THREADS = WARPS * 32;
constexpr int BIN_STRIDE = WARPS * 128;   // bytes between consecutive bins: bank == lane for every access

// ---------------------------------------------------------------- policies
// Each policy: PX = px-ch per lane, how a lane's PX bytes of one frame are
// accumulated, and how the 256-bin scan unpacks the counters.
struct PolU8Rmw {          // 4 saturating u8 counters per word, LDS/VIADDMNMX/STS
  static constexpr int PX = 4;
  static constexpr int MAXN = 509;
  __device__ static void add(unsigned char* base, unsigned w) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned bin = (w >> (8 * j)) & 0xFFu;
      unsigned char* p = base + bin * BIN_STRIDE + j;
      unsigned c = *p;
      *p = (unsigned char)__viaddmin_u32(c, 1u, 255u);
    }
  }
  __device__ static void unpack(unsigned w, unsigned* c) { c[0] = w & 255u; c[1] = (w >> 8) & 255u; c[2] = (w >> 16) & 255u; c[3] = w >> 24; }
};
struct PolU8Atom {         // 4 wrapping u8 counters per word, one ATOMS per byte (exact only for N<=255)
  static constexpr int PX = 4;
  static constexpr int MAXN = 255;
  __device__ static void add(unsigned char* base, unsigned w) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned bin = (w >> (8 * j)) & 0xFFu;
      atomicAdd((unsigned*)(base + bin * BIN_STRIDE), 1u << (8 * j));
    }
  }
  __device__ static void unpack(unsigned w, unsigned* c) { c[0] = w & 255u; c[1] = (w >> 8) & 255u; c[2] = (w >> 16) & 255u; c[3] = w >> 24; }
};
struct PolU16Atom {        // 2 u16 counters per word, one ATOMS per byte, N<=65535
  static constexpr int PX = 2;
  static constexpr int MAXN = 65535;
  __device__ static void add(unsigned char* base, unsigned w) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      unsigned bin = (w >> (8 * j)) & 0xFFu;
      atomicAdd((unsigned*)(base + bin * BIN_STRIDE), 1u << (16 * j));
    }
  }
  __device__ static void unpack(unsigned w, unsigned* c) { c[0] = w & 0xFFFFu; c[1] = w >> 16; }
};
struct PolU10Atom {        // 3 ten-bit counters per word (one BGR pixel per lane), N<=1023
  static constexpr int PX = 3;
  static constexpr int MAXN = 1023;
  __device__ static void add(unsigned char* base, unsigned w) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      unsigned bin = (w >> (8 * j)) & 0xFFu;
      atomicAdd((unsigned*)(base + bin * BIN_STRIDE), 1u << (10 * j));
    }
  }
  __device__ static void unpack(unsigned w, unsigned* c) { c[0] = w & 1023u; c[1] = (w >> 10) & 1023u; c[2] = (w >> 20) & 1023u; }
};
struct PolU16Rmw {         // 2 u16 counters per word, LDS.U16/IADD/STS.U16
  static constexpr int PX = 2;
  static constexpr int MAXN = 65535;
  __device__ static void add(unsigned char* base, unsigned w) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      unsigned bin = (w >> (8 * j)) & 0xFFu;
      unsigned short* p = (unsigned short*)(base + bin * BIN_STRIDE + 2 * j);
      *p = (unsigned short)(*p + 1);
    }
  }
  __device__ static void unpack(unsigned w, unsigned* c) { c[0] = w & 0xFFFFu; c[1] = w >> 16; }
};

// raw per-lane load for one frame (PX==3: 24 lanes fetch the 96-byte segment)
template <int PX>
__device__ __forceinline__ unsigned load_raw(const uint8_t* p, int lane) {
  if (PX == 4) return __ldg((const unsigned*)p);
  if (PX == 2) return __ldg((const unsigned short*)p);
  return (lane < 24) ? __ldg((const unsigned*)p) : 0u;
}
// turn the raw word into "this lane's PX bytes in the low bytes"
template <int PX>
__device__ __forceinline__ unsigned fix_px(unsigned wv, int lane) {
  if (PX != 3) return wv;
  int b = 3 * lane;
  unsigned lo = __shfl_sync(0xffffffffu, wv, b >> 2);
  unsigned hi = __shfl_sync(0xffffffffu, wv, ((b >> 2) + 1) & 31);
  return __funnelshift_r(lo, hi, 8 * (b & 3));
}

template <class P, int U, int mode>
__global__ void __launch_bounds__(THREADS, 1)
median_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, int n, long long m, int nseg) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int PX = P::PX;
  constexpr int SEG = 32 * PX;                         // bytes of one frame a warp owns
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* base = smem + warp * 128 + lane * 4;  // + bin*BIN_STRIDE
  const int klo = (n - 1) >> 1, khi = n >> 1;
  const int lane_off = (PX == 4) ? lane * 4 : (PX == 2 ? lane * 2 : (lane < 24 ? lane * 4 : 0));
  const int nb = n / U;
  for (int seg = blockIdx.x * WARPS + warp; seg < nseg; seg += gridDim.x * WARPS) {
    const uint8_t* src = frames + (long long)seg * SEG + lane_off;
#pragma unroll 8
    for (int b = 0; b < 256; ++b) *(unsigned*)(base + b * BIN_STRIDE) = 0u;
    __syncwarp();
    if (mode != 2) {
      unsigned ra[U], rb[U];
      const uint8_t* p = src;
      auto load = [&](unsigned (&r)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) { r[u] = load_raw<PX>(p, lane); p += m; }
      };
      auto proc = [&](unsigned (&r)[U]) {
        if (mode == 0) {
#pragma unroll
          for (int u = 0; u < U; ++u) P::add(base, fix_px<PX>(r[u], lane));
        } else {                                       // mode 1: loads only
          unsigned acc = 0;
#pragma unroll
          for (int u = 0; u < U; ++u) acc += r[u];
          if (acc == 0x12345678u) *(unsigned*)base = acc;
        }
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
      for (int f = nb * U; f < n; ++f) {
        unsigned w = load_raw<PX>(p, lane); p += m;
        if (mode == 0) P::add(base, fix_px<PX>(w, lane));
      }
    } else {                                           // mode 2: histogram updates only, no global loads
      unsigned x = seg * 2654435761u + lane * 40503u;
      for (int f = 0; f < n; ++f) { x = x * 1664525u + 1013904223u; P::add(base, x >> 8); }
    }
    __syncwarp();
    // ---- two-level scan: 16 coarse groups, then 16 bins inside the group ----
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
        unsigned w = *(unsigned*)(base + (g * 16 + b) * BIN_STRIDE);
        unsigned t[PX]; P::unpack(w, t);
#pragma unroll
        for (int j = 0; j < PX; ++j) c[j] += t[j];
      }
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        unsigned nc = cum[j] + c[j];
        if (nc <= (unsigned)klo) { gsel[j][0] = g + 1; before[j][0] = nc; }
        if (nc <= (unsigned)khi) { gsel[j][1] = g + 1; before[j][1] = nc; }
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
        if (q == 1 && gsel[j][1] == gsel[j][0] && khi == klo) { med[1] = med[0]; continue; }
        unsigned g = min(gsel[j][q], 15u);
        unsigned v = g * 16, c2 = before[j][q];
#pragma unroll 4
        for (int b = 0; b < 16; ++b) {
          unsigned w = *(unsigned*)(base + (g * 16 + b) * BIN_STRIDE);
          unsigned t[PX]; P::unpack(w, t);
          c2 += t[j];
          if (c2 <= k) v = g * 16 + b + 1;
        }
        med[q] = min(v, 255u);
      }
      res |= ((med[0] + med[1]) >> 1) << (8 * j);
    }
    uint8_t* dst = out + (long long)seg * SEG;
    if (PX == 4) ((unsigned*)dst)[lane] = res;
    else if (PX == 2) ((unsigned short*)dst)[lane] = (unsigned short)res;
    else { dst[3 * lane] = res & 255u; dst[3 * lane + 1] = (res >> 8) & 255u; dst[3 * lane + 2] = (res >> 16) & 255u; }
    __syncwarp();
  }
}

template <class P, int U, int mode>
float run(const char* name, const uint8_t* d_frames, uint8_t* d_out, int n, long long m, int sms, int reps,
          const std::vector<uint8_t>& ref, long long ncheck) {
  constexpr int SEG = 32 * P::PX;
  int nseg = (int)(m / SEG);
  size_t smem = 256 * BIN_STRIDE;
  CK(cudaFuncSetAttribute(median_kernel<P, U, mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemset(d_out, 0xEE, m));
  for (int i = 0; i < 2; ++i) median_kernel<P, U, mode><<<sms, THREADS, smem>>>(d_frames, d_out, n, m, nseg);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  float best = 1e30f, sum = 0;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(e0));
    median_kernel<P, U, mode><<<sms, THREADS, smem>>>(d_frames, d_out, n, m, nseg);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms); sum += ms;
  }
  double bytes = (double)(n + 1) * nseg * SEG;
  long long bad = -1;
  if (mode == 0 && n <= P::MAXN) {
    std::vector<uint8_t> got(ncheck);
    CK(cudaMemcpy(got.data(), d_out, ncheck, cudaMemcpyDeviceToHost));
    bad = 0; for (long long i = 0; i < ncheck; ++i) bad += got[i] != ref[i];
  }
  printf("%-14s U=%2d mode=%d n=%d  best %.3f ms  avg %.3f ms  %.0f GB/s (best)  mismatches=%lld\n", name, U, mode, n, best, sum / reps,
         bytes / best / 1e6, bad);
  fflush(stdout);
  return best;
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 300;
  int H = argc > 2 ? atoi(argv[2]) : 1080, W = argc > 3 ? atoi(argv[3]) : 1920;
  long long m = (long long)H * W * 3;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, n=%d m=%lld (%.2f GB)\n", prop.name, sms, n, m, (double)n * m / 1e9);
  std::vector<uint8_t> h((size_t)n * m);
  uint32_t x = 12345;
  // background-like data: per px-ch base value + noise in [-6,6] + occasional outliers
  std::vector<uint8_t> basev(m);
  for (long long i = 0; i < m; ++i) { x = x * 1664525u + 1013904223u; basev[i] = x >> 24; }
  for (int f = 0; f < n; ++f)
    for (long long i = 0; i < m; ++i) {
      x = x * 1664525u + 1013904223u;
      int v = basev[i] + (int)((x >> 16) % 13) - 6;
      if (((x >> 8) & 63) == 0) v = (x >> 20) & 255;
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
  const int reps = 5;
#define RUN(P, U, MODE, NAME) run<P, U, MODE>(NAME, d_frames, d_out, n, m, sms, reps, ref, ncheck)
  RUN(PolU8Rmw, 16, 0, "u8x4-rmw"); RUN(PolU8Rmw, 8, 0, "u8x4-rmw"); RUN(PolU8Rmw, 32, 0, "u8x4-rmw");
  RUN(PolU8Atom, 16, 0, "u8x4-atom");
  RUN(PolU16Atom, 16, 0, "u16x2-atom"); RUN(PolU16Atom, 32, 0, "u16x2-atom");
  RUN(PolU10Atom, 16, 0, "u10x3-atom"); RUN(PolU10Atom, 32, 0, "u10x3-atom");
  RUN(PolU16Rmw, 16, 0, "u16x2-rmw");
  RUN(PolU8Rmw, 16, 1, "u8x4 loads"); RUN(PolU8Rmw, 32, 1, "u8x4 loads"); RUN(PolU16Atom, 32, 1, "u16x2 loads"); RUN(PolU10Atom, 32, 1, "u10x3 loads");
  RUN(PolU8Rmw, 16, 2, "u8x4-rmw"); RUN(PolU8Atom, 16, 2, "u8x4-atom"); RUN(PolU16Atom, 16, 2, "u16x2-atom"); RUN(PolU10Atom, 16, 2, "u10x3-atom"); RUN(PolU16Rmw, 16, 2, "u16x2-rmw");
  // plain copy ceiling for context
  {
    uint8_t* d2; CK(cudaMalloc(&d2, (size_t)n * m / 2));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int i = 0; i < 5; ++i) { CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(d2, d_frames, (size_t)n * m / 2, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms); }
    printf("memcpy D2D %.0f GB/s (read+write)\n", (double)n * m / best / 1e6);
  }
  return 0;
}
