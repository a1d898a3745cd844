// Development micro-benchmark: how fast can one SM stream "n rows of W bytes, one per frame" into shared memory
// with cp.async (LDGSTS.128), as a function of the contiguous width fetched at a time?  Fetch only, no compute.
//   mode 0: every warp fetches its own 64-byte column segment (8 rows per instruction), warps independent
//   mode 1: the CTA's 8 warps fetch one 512-byte-wide tile together (one 512-byte row per instruction), barrier per tile
//   mode 2: like 0 but 128-byte segments per warp (4 rows per instruction)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void cp16(unsigned s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) fetch_kernel(const uint8_t* __restrict__ frames, unsigned* out, int n, long long m, int ntiles) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  unsigned acc = 0;
  if (MODE == 0 || MODE == 2) {
    constexpr int SEGB = MODE == 0 ? 64 : 128;
    constexpr int RPI = 512 / SEGB;           // rows per instruction
    constexpr int CPRW = SEGB / 16;
    const int nseg = ntiles * (512 / SEGB);
    constexpr int NW = 512 / SEGB;            // warps that fit the n*512-byte buffer
    if (warp >= NW) return;
    const unsigned buf = sbase + warp * (n * SEGB);
    for (int seg = blockIdx.x * NW + warp; seg < nseg; seg += gridDim.x * NW) {
      const uint8_t* src = frames + (long long)seg * SEGB + (lane % CPRW) * 16 + (long long)(lane / CPRW) * m;
      for (int f = 0; f < n; f += RPI) {
        if (f + lane / CPRW < n) cp16(buf + f * SEGB + lane * 16, src + (long long)f * m);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      acc += *reinterpret_cast<const unsigned*>(sm + warp * (n * SEGB) + lane * 4);
    }
  } else {
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const uint8_t* src = frames + (long long)tile * 512 + lane * 16;
      for (int f = warp; f < n; f += 8) cp16(sbase + f * 512 + lane * 16, src + (long long)f * m);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      acc += *reinterpret_cast<const unsigned*>(sm + threadIdx.x * 4);
      __syncthreads();
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
void run(const char* name, const uint8_t* d, unsigned* out, int n, long long m) {
  const int ntiles = (int)(m / 512);
  const int smem = n * 512;
  CK(cudaFuncSetAttribute(fetch_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) fetch_kernel<MODE><<<148, 256, smem>>>(d, out, n, m, ntiles);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < 5; ++i) {
    CK(cudaEventRecord(e0));
    fetch_kernel<MODE><<<148, 256, smem>>>(d, out, n, m, ntiles);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  printf("%-40s n=%d: %.3f ms  %.0f GB/s\n", name, n, best, (double)n * ntiles * 512 / best / 1e6);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 300;
  const long long m = 1080LL * 1920 * 3;
  uint8_t* d; unsigned* out;
  CK(cudaMalloc(&d, (size_t)n * m)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(d, 1, (size_t)n * m));
  run<0>("warp-private 64 B rows (8 rows/instr)", d, out, n, m);
  run<2>("warp-private 128 B rows (4 rows/instr)", d, out, n, m);
  run<1>("CTA tile, 512 B rows (1 row/instr)", d, out, n, m);
  return 0;
}
