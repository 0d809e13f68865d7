// Developer check: tcgen05.mma kind::f16 with BOTH operands MN-major and several 64-element swizzle atoms along M / N (what a
// gradient GEMM g^T x would use to read row-major g planes and codes without a transposed copy).  Operand images are written the
// way TMA SWIZZLE_128B boxes of [64 columns x 64 rows] would land: box j at j * 8192 B, row k at k * 128 B, 16-byte piece c at
// ((c ^ (k & 7)) << 4).  Two descriptor variants are tried: (LBO = atom stride 8192, SBO = 1024) and the swapped one.
// Build: nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../quantized_vit_b200/csrc -I../../include -o mnmajor_check.bin mnmajor_check.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_ptx.cuh"

using namespace qvit;

constexpr int M = 128, N = 256, K = 64;

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// at: [K][M] bf16 (A transposed = MN-major A), bt: [K][N] bf16 (MN-major B), out: [M][N] fp32
__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* at, const __nv_bfloat16* bt, float* out, int variant) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  uint8_t* a_sm = gen;                 // 2 boxes x 8 KB
  uint8_t* b_sm = gen + 16384;         // 4 boxes x 8 KB
  const uint32_t bar = base + 49152;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + 49152 + 16);
  for (int i = threadIdx.x; i < K * (M / 8); i += 128) {          // 16-byte pieces of A^T
    const int kk = i / (M / 8), piece = i % (M / 8), box = piece / 8, c = piece % 8;
    const uint4 v = *reinterpret_cast<const uint4*>(at + kk * M + piece * 8);
    *reinterpret_cast<uint4*>(a_sm + box * 8192 + kk * 128 + ((c ^ (kk & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < K * (N / 8); i += 128) {
    const int kk = i / (N / 8), piece = i % (N / 8), box = piece / 8, c = piece % 8;
    const uint4 v = *reinterpret_cast<const uint4*>(bt + kk * N + piece * 8);
    *reinterpret_cast<uint4*>(b_sm + box * 8192 + kk * 128 + ((c ^ (kk & 7)) << 4)) = v;
  }
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) { ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t*>(slot)), 256); ptx::tmem_relinquish<1>(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *slot, 0);
  if (warp == 1) {
    if (ptx::elect_one()) {
      // D = f32, A = B = bf16, A and B MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const uint32_t lbo = variant == 0 ? 8192u : 1024u, sbo = variant == 0 ? 1024u : 8192u;
      for (int ks = 0; ks < K / 16; ++ks)
        ptx::mma_f16_ss(tmem, desc_mn(base + ks * 2048, lbo, sbo), desc_mn(base + 16384 + ks * 2048, lbo, sbo), idesc, ks ? 1u : 0u);
      ptx::mma_commit(bar);
    }
    __syncwarp();
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  const int row = threadIdx.x;                                      // warp w reads TMEM lanes 32 w ..
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[row * N + c0 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<1>(tmem, 256); }
}

int main() {
  std::vector<__nv_bfloat16> at(K * M), bt(K * N);
  std::vector<float> af(K * M), bf(K * N), ref(M * N, 0.f), got(M * N);
  srand(1);
  for (int i = 0; i < K * M; ++i) { af[i] = (float)(rand() % 15 - 7); at[i] = __float2bfloat16(af[i]); }
  for (int i = 0; i < K * N; ++i) { bf[i] = (float)(rand() % 15 - 7); bt[i] = __float2bfloat16(bf[i]); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int kk = 0; kk < K; ++kk) s += af[kk * M + m] * bf[kk * N + n];
      ref[m * N + n] = s;
    }
  __nv_bfloat16 *dat, *dbt;
  float* dout;
  cudaMalloc(&dat, at.size() * 2); cudaMalloc(&dbt, bt.size() * 2); cudaMalloc(&dout, got.size() * 4);
  cudaMemcpy(dat, at.data(), at.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dbt, bt.data(), bt.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dout, 0, got.size() * 4);
    k<<<1, 128, 52000>>>(dat, dbt, dout, variant);
    cudaError_t e = cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
    int bad = 0, bad_m64 = 0, bad_n64 = 0;
    for (int i = 0; i < M * N; ++i)
      if (got[i] != ref[i]) { ++bad; if ((i / N) >= 64) ++bad_m64; if ((i % N) >= 64) ++bad_n64; }
    printf("variant %d (LBO %s): %d of %d wrong (rows >= 64: %d, cols >= 64: %d)  D[0][0]=%g ref %g  D[70][200]=%g ref %g\n", variant,
           variant == 0 ? "= atom stride 8192, SBO = 1024" : "= 1024, SBO = atom stride 8192", bad, M * N, bad_m64, bad_n64, got[0], ref[0],
           got[70 * N + 200], ref[70 * N + 200]);
  }
  return 0;
}
