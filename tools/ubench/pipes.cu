// Developer micro-benchmark: issue / pipe throughput of the CUDA-core instructions the GEMM epilogue is made of (sm_100a).
// Build: nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o pipes.bin pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define ITERS 4096
#define NV 16   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, float seed, int n_iter) {
  float x[NV];
  uint32_t u[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) { x[j] = seed + j + threadIdx.x; u[j] = threadIdx.x * 7 + j; }
  const float c0 = seed * 0.5f, c1 = seed * 0.25f;
  unsigned long long cc0, cc1;
  asm("mov.b64 %0, {%1, %1};" : "=l"(cc0) : "f"(c0));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cc1) : "f"(c1));
  for (int it = 0; it < n_iter; ++it) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (MODE == 0) x[j] = fmaf(x[j], c0, c1);                                      // FFMA
      if (MODE == 1 && (j & 1) == 0) {                                               // FFMA2 (2 elements per instruction)
        unsigned long long p;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x[j]), "f"(x[j + 1]));
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(cc0), "l"(cc1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x[j]), "=f"(x[j + 1]) : "l"(p));
      }
      if (MODE == 2) x[j] = fminf(x[j], c0 + j);                                     // FMNMX
      if (MODE == 3) u[j] = u[j] + 0x4B400000u + j;                                  // IADD
      if (MODE == 4) u[j] = __byte_perm(u[j], u[(j + 1) % NV], 0x0040 + j);           // PRMT
      if (MODE == 5) asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));                 // MUFU.EX2
      if (MODE == 6) x[j] = (float)(int)u[j], u[j] += 3;                             // I2F (+IADD)
      if (MODE == 7) u[j] = __vmins2(u[j], 0x00070007u + j);                         // VIMNMX.S16x2 ?
      if (MODE == 8) { x[j] = fmaf(x[j], c0, c1); u[j] = u[j] + 0x4B400000u + j; }   // FFMA + IADD 1:1
      if (MODE == 9) {                                                               // FFMA2 + IADD (same flops as 8)
        if ((j & 1) == 0) {
          unsigned long long p;
          asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x[j]), "f"(x[j + 1]));
          asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(cc0), "l"(cc1));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(x[j]), "=f"(x[j + 1]) : "l"(p));
        }
        u[j] = u[j] + 0x4B400000u + j;
      }
      if (MODE == 10) { x[j] = fmaf(x[j], c0, c1); x[j] = fminf(x[j], c0 + j); }     // FFMA + FMNMX 1:1
      if (MODE == 11) {                                                              // FFMA2 + FMNMX
        if ((j & 1) == 0) {
          unsigned long long p;
          asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x[j]), "f"(x[j + 1]));
          asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(cc0), "l"(cc1));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(x[j]), "=f"(x[j + 1]) : "l"(p));
        }
        x[j] = fminf(x[j], c0 + j);
      }
      if (MODE == 12) u[j] = (u[j] ^ u[(j + 1) % NV]) | (uint32_t)j;                   // LOP3
      if (MODE == 13) x[j] = rintf(x[j]) + 0.5f;                                       // FRND (+FADD)
      if (MODE == 14) { x[j] = fmaf(x[j], c0, c1); asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j])); }  // FFMA + MUFU
      if (MODE == 15) u[j] = (uint32_t)__float2int_rn(x[j]) , x[j] += 1.0f;           // F2I (+FADD)
      if (MODE == 16) x[j] = x[j] + c0;                                                // FADD
      if (MODE == 17 && (j & 1) == 0) {                                                // F2FP.F16.F32.PACK_AB (2 elements per instruction)
        const __half2 h = __floats2half2_rn(x[j], x[j + 1]);
        u[j] ^= *reinterpret_cast<const uint32_t*>(&h);
        x[j] += 1.0f;
      }
      if (MODE == 18) {                                                                // HADD2.F32 (fp16 -> fp32)
        const __half2 h = *reinterpret_cast<const __half2*>(&u[j]);
        x[j] += __low2float(h);
        u[j] += 0x3c01u;
      }
      if (MODE == 19) x[j] = fmaf(x[j], 1.0009765625f, 0.25f);                         // FFMA, immediate operands
      if (MODE == 20 && (j & 1) == 0) {                                                // FADD2
        unsigned long long p;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x[j]), "f"(x[j + 1]));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(cc0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x[j]), "=f"(x[j + 1]) : "l"(p));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) s += x[j] + __uint_as_float(u[j]);
  if (s == 123.456f) out[threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter_elem) {
  float* d;
  cudaMalloc(&d, 4096);
  const int grid = 148 * 1;
  k<MODE><<<grid, 512>>>(d, 1.0f, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 512>>>(d, 1.0f, ITERS);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // warp-level "operations" (as listed in the mode) per second per SM
  const double warp_ops = (double)grid * 16 /*warps*/ * ITERS * NV * ops_per_iter_elem;
  printf("%-28s %8.3f ms   %6.2f warp-ops/ns/SM  (= %5.2f per clk per SM at 1.9 GHz)\n", name, ms, warp_ops / (ms * 1e6) / 148,
         warp_ops / (ms * 1e6) / 148 / 1.9);
  cudaFree(d);
}

int main() {
  run<0>("FFMA", 1);
  run<1>("FFMA2 (per fp32 fma)", 1);
  run<16>("FADD", 1);
  run<2>("FMNMX", 1);
  run<3>("IADD", 1);
  run<4>("PRMT", 1);
  run<12>("LOP3", 1);
  run<7>("VIMNMX.S16x2 (__vmins2)", 1);
  run<5>("MUFU.EX2", 1);
  run<6>("I2F + IADD", 1);
  run<13>("FRND + FADD", 1);
  run<15>("F2I + FADD", 1);
  run<8>("FFMA + IADD (pairs)", 1);
  run<9>("FFMA2 + IADD (pairs)", 1);
  run<10>("FFMA + FMNMX (pairs)", 1);
  run<11>("FFMA2 + FMNMX (pairs)", 1);
  run<14>("FFMA + MUFU (pairs)", 1);
  run<17>("F2FP.F16 pack (+FADD, IADD; per element)", 1);
  run<18>("HADD2.F32 unpack (+FADD, IADD)", 1);
  run<19>("FFMA imm", 1);
  run<20>("FADD2 (per fp32 add)", 1);
  return 0;
}
