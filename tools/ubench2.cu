// Development micro-benchmarks (not product code):
//  (1) ALU throughput of candidate per-byte primitives on sm_100a
//  (2) achievable bandwidth of the "N frames x small segment" read pattern of the temporal median
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int OP>
__global__ void alu_kernel(unsigned* out, unsigned seed, int iters) {
  unsigned a[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i + 1); acc[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) acc[i] = __vsadu4(a[i], acc[i]) + acc[i];          // VABSDIFF4.ACC
      if (OP == 1) acc[i] = __dp4a(a[i], acc[i], acc[i]);              // IDP.4A
      if (OP == 2) acc[i] = (acc[i] & a[i]) ^ (acc[i] >> 1);          // LOP3+SHF
      if (OP == 3) acc[i] = acc[i] * 3 + a[i];                         // IMAD
      if (OP == 4) acc[i] = __byte_perm(acc[i], a[i], 0x5140 + (acc[i] & 1)); // PRMT
      if (OP == 5) acc[i] = __viaddmin_u32(acc[i], a[i], 0x7fffffffu);  // VIADDMNMX
      if (OP == 6) acc[i] = __popc(acc[i]) + a[i];                     // POPC
    }
  }
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run_alu(const char* name, unsigned* d_out, int sms) {
  const int iters = 4096, threads = 512, blocks = sms * 4;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  alu_kernel<OP><<<blocks, threads>>>(d_out, 12345u, iters); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); alu_kernel<OP><<<blocks, threads>>>(d_out, 12345u, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  double ops = (double)blocks * threads * iters * 8;
  printf("ALU %-22s %.1f Gop/s  = %.1f lane-ops/ns/SM\n", name, ops / ms / 1e6, ops / ms / 1e6 / sms);
}

// read pattern: each warp owns SEGW bytes of every frame; loads V-byte vectors
template <int V, int U>
__global__ void __launch_bounds__(256) read_kernel(const uint8_t* __restrict__ frames, unsigned* out, int n, long long stride, int nseg, int warps_per_cta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int SEGW = 32 * V;
  unsigned acc = 0;
  for (int seg = blockIdx.x * warps_per_cta + warp; seg < nseg; seg += gridDim.x * warps_per_cta) {
    const uint8_t* p = frames + (long long)seg * SEGW + lane * V;
    for (int f = 0; f + U <= n; f += U) {
      if (V == 4) { unsigned r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { r[u] = __ldg((const unsigned*)p); p += stride; }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += r[u];
      } else { uint4 r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { r[u] = __ldg((const uint4*)p); p += stride; }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += r[u].x ^ r[u].y ^ r[u].z ^ r[u].w;
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <int V, int U>
void run_read(const char* name, const uint8_t* d, unsigned* d_out, int n, long long stride, long long m, int sms, int warps, int ctas_per_sm) {
  int nseg = (int)(m / (32 * V));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  read_kernel<V, U><<<sms * ctas_per_sm, warps * 32>>>(d, d_out, n, stride, nseg, warps); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < 3; ++i) {
    CK(cudaEventRecord(e0)); read_kernel<V, U><<<sms * ctas_per_sm, warps * 32>>>(d, d_out, n, stride, nseg, warps); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  printf("READ %-10s V=%2d U=%2d warps/CTA=%d CTAs/SM=%d stride=%lld m=%lld: %.3f ms %.0f GB/s\n", name, V, U, warps, ctas_per_sm, stride, m, best, (double)(n / U * U) * nseg * 32 * V / best / 1e6);
  fflush(stdout);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  unsigned* d_out; CK(cudaMalloc(&d_out, 1 << 24));
  run_alu<0>("VABSDIFF4.ACC+IADD", d_out, sms); run_alu<1>("IDP.4A", d_out, sms); run_alu<2>("LOP3+SHF (2 ops)", d_out, sms);
  run_alu<3>("IMAD", d_out, sms); run_alu<4>("LOP3+IADD+PRMT (3 ops)", d_out, sms); run_alu<5>("VIADDMNMX", d_out, sms); run_alu<6>("POPC+IADD", d_out, sms);
  const int n = 300; const long long m = 1080LL * 1920 * 3;
  uint8_t* d; CK(cudaMalloc(&d, (size_t)n * m)); CK(cudaMemset(d, 1, (size_t)n * m));
  // real layout: frame stride = m
  run_read<4, 16>("frames", d, d_out, n, m, m, sms, 7, 1);
  run_read<4, 32>("frames", d, d_out, n, m, m, sms, 7, 1);
  run_read<4, 32>("frames", d, d_out, n, m, m, sms, 8, 2);
  run_read<4, 32>("frames", d, d_out, n, m, m, sms, 8, 4);
  run_read<4, 16>("frames", d, d_out, n, m, m, sms, 8, 8);
  run_read<16, 8>("frames", d, d_out, n, m, m, sms, 7, 1);
  run_read<16, 16>("frames", d, d_out, n, m, m, sms, 7, 1);
  run_read<16, 8>("frames", d, d_out, n, m, m, sms, 8, 4);
  // TLB-friendly control: same volume, but "frames" are only 64 KB apart inside slabs (tile-major layout)
  run_read<4, 32>("tilemajor", d, d_out, n, 128, m, sms, 7, 1);
  run_read<4, 32>("64KB", d, d_out, n, 65536, 65536, sms, 7, 1);
  run_read<4, 32>("2MB", d, d_out, n, 2097152, 2097152, sms, 7, 1);
  run_read<4, 32>("256KB", d, d_out, n, 262144, 262144, sms, 7, 1);
  return 0;
}
