// Microbenchmark: tcgen05.mma issue/throughput floor for M=128, N in {64,128,256}, K=16 steps, SS mode.
// One CTA per SM, no TMA, operands are whatever is in smem. Variants: number of independent accumulators,
// MMAs per commit. Prints cycles per MMA.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../calciumgan_b200/csrc/cg_kernels_tc.cuh"
int cg_tc_set_err(const char* m) { fprintf(stderr, "%s\n", m); return 1; }
using namespace tc;

__global__ void __launch_bounds__(128, 1) k_rate(int N, int nacc, int per_commit, int total, int a_kmajor, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(128, N, a_kmajor ? 0 : 1, a_kmajor ? 0 : 1);
    const uint32_t hi = desc_hi(1024);
    const uint32_t a_lo = desc_lo(smem_u32(smem), a_kmajor ? 16 : 8192), b_lo = desc_lo(smem_u32(smem + 16384), a_kmajor ? 16 : 8192);
    uint32_t ph = 0;
    long long t0 = clock64();
    int issued = 0;
    while (issued < total) {
      if (elect_one()) {
        for (int j = 0; j < per_commit; ++j) {
          const int k = j & 3;
          const int acc = (j >> 2) % nacc;
          umma_bf16_lohi(tb + acc * N, a_lo + (a_kmajor ? 2 * k : 128 * k), b_lo + (a_kmajor ? 2 * k : 128 * k), hi, idesc, 1);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      issued += per_commit;
      mbar_wait(&bar, ph);   // wait for completion of this batch (per_commit large => amortised)
      ph ^= 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int Ns[] = {64, 128, 256};
  printf("%-8s %-6s %-5s %-10s %-10s\n", "layout", "N", "nacc", "per_commit", "clk/MMA");
  for (int kmaj = 1; kmaj >= 0; --kmaj)
    for (int n = 0; n < 3; ++n)
      for (int nacc = 1; nacc <= 2; ++nacc)
        for (int pc : {4, 8, 32, 256}) {
          const int total = 4096;
          k_rate<<<148, 128, 100 * 1024>>>(Ns[n], nacc, pc, total, kmaj, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long c;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("%-8s %-6d %-5d %-10d %-10.1f (math %d clk)\n", kmaj ? "K-major" : "MN-major", Ns[n], nacc, pc, (double)c / total, 128 * Ns[n] * 16 / 4096);
        }
  return 0;
}
