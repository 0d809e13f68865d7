// Developer micro-benchmark: issue cost and execution time of tcgen05.mma kind::f16 (bf16 operands) by shape and operand source,
// one thread issuing back-to-back MMAs into one accumulator (what the attention kernels do).  Operands are zeros (timing only).
// Build: nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -I../../quantized_vit_b200/csrc -o mma_rate.bin mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_ptx.cuh"

using namespace qvit;

__host__ __device__ constexpr uint32_t idesc(int M, int N, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode 0: SS, K-major B; 1: SS, MN-major B (N = 64 only); 2: TS (A from TMEM), MN-major B (N = 64); 3: TS, K-major B
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int n_mma, int distinct_acc, long long* out, int alt) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  for (int i = threadIdx.x; i < (16384 + 32768 + 64) / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0;
  const uint32_t bar = base + 16384 + 32768;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + 16384 + 32768 + 32);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(bar + 8u * i, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t*>(slot)), 512); ptx::tmem_relinquish<1>(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  // `distinct_acc` = number of issuing warps (1, 2 or 4): lane 0 of warp w issues n_mma MMAs into its own accumulator
  const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // warp-uniform for the compiler
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  if (w < distinct_acc) {
    const uint32_t tmem = tmem_u;
    const uint32_t id = idesc(128, N, mode == 1 || mode == 2);
    const uint64_t ad = ptx::make_kmajor_sw128_desc(base), bd = ptx::make_kmajor_sw128_desc(base + 16384);
    const uint32_t mybar = bar + 8u * w;
    const uint32_t dd = tmem + 256 + (uint32_t)(w * 64);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        if (ptx::elect_one()) {
          const uint32_t d = dd + (uint32_t)((i & alt) * 64);       // alt = 1: two accumulators alternate, 3: four
          if (mode <= 1) ptx::mma_f16_ss(d, ad + (uint64_t)((i & 3) * 2), bd + (uint64_t)((i & 3) * 2), id, 1u);
          else ptx::mma_f16_ts(d, tmem + (uint32_t)((i & 3) * 8), bd + (uint64_t)((i & 3) * (mode == 2 ? 128 : 2)), id, 1u);
        }
      }
      const long long t1 = clock64();
      if (ptx::elect_one()) ptx::mma_commit(mybar);
      __syncwarp();
      ptx::mbar_wait(mybar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      if (w == 0 && (threadIdx.x & 31) == 0) {
        out[rep * 2] = t1 - t0;
        out[rep * 2 + 1] = t2 - t0;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc<1>(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  const int n = 64;
  struct { int mode, N; const char* what; } cfg[] = {
      {0, 16, "SS K-major B  N=16"}, {0, 32, "SS K-major B  N=32"}, {0, 64, "SS K-major B  N=64"}, {0, 96, "SS K-major B  N=96"},
      {0, 128, "SS K-major B  N=128"}, {0, 208, "SS K-major B  N=208"}, {0, 256, "SS K-major B  N=256"}, {1, 64, "SS MN-major B N=64"},
      {2, 64, "TS MN-major B N=64"},  {3, 64, "TS K-major B  N=64"},  {3, 128, "TS K-major B  N=128"}, {3, 208, "TS K-major B  N=208"}};
  for (int na = 1; na <= 4; na *= 2)
    for (int mode = 0; mode <= 2; mode += 2) {
      k<<<1, 128, 60000>>>(mode, 64, n, na, d, 0);
      long long h[6];
      cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
      printf("%s N=64, %d issuing warps (one accumulator each): issue %6.1f cyc/MMA, issue+complete %6.1f cyc/MMA\n", mode ? "TS" : "SS", na, h[4] / (double)n, h[5] / (double)n);
    }
  for (int alt = 1; alt <= 3; alt += 2)
    for (int mode = 0; mode <= 2; mode += 2) {
      k<<<1, 128, 60000>>>(mode, 64, n, 1, d, alt);
      long long h[6];
      cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
      printf("%s N=64, ONE issuing warp alternating over %d accumulators: issue %6.1f cyc/MMA, issue+complete %6.1f cyc/MMA\n", mode ? "TS" : "SS", alt + 1, h[4] / (double)n, h[5] / (double)n);
    }
  for (auto& c : cfg) {
    k<<<1, 128, 60000>>>(c.mode, c.N, n, 1, d, 0);
    long long h[6];
    cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: %s\n", c.what, cudaGetErrorString(e)); return 1; }
    printf("%-22s M=128 K=16 x%d: issue %6.1f cyc/MMA, issue+complete %6.1f cyc/MMA (third repetition; first %6.1f)\n", c.what, n,
           h[4] / (double)n, h[5] / (double)n, h[1] / (double)n);
  }
  return 0;
}
