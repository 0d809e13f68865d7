// Shared device helpers for libqvit_b200 (sm_100a only).
//
// Exact-arithmetic contract (see DESIGN.md "bit-exactness"): every quantizer kernel uses IEEE
// round-to-nearest division (__fdiv_rn) and round-half-to-even (rintf -> cvt.rni / FRND), never a
// reciprocal multiply, never --use_fast_math; that is what makes the integer codes equal to
// torch.round(x.div(d)) of the reference (quant_layers.py:157).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/qvit_b200.h"

namespace qvit {

// ---- error plumbing (api.cu) -----------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define QVIT_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::qvit::set_error(__VA_ARGS__);           \
      return QVIT_ERR_INVALID;                  \
    }                                           \
  } while (0)

constexpr int kFlagNaN = 1;        // a NaN reached a quantizer (the reference would propagate it)
constexpr int kFlagOverflow = 2;   // |code| > 127 : int8 pipe not applicable (caller must use the wide path)
constexpr int kFlagNaNGrad = 4;    // NaN in a reduced gradient (reference raises NanInGradientError)

// ---- quantizer parameters resolved once per thread ---------------------------------------------
struct SymParams {
  float d;        // |d_quant|
  float qm;       // signed q_m (the reference compares |x| >= q_m with the signed value, QL:159)
  float t;        // exponent (non-linear) or 1
  float sat;      // |round(r / d)|, r = |q_m| (linear) or exp(t*log(|q_m|+1e-6)) (non-linear)
  int nonlinear;
};

__device__ __forceinline__ SymParams load_sym_params(const float* d, const float* qm, const float* t) {
  SymParams p;
  p.d = fabsf(__ldg(d));
  p.qm = __ldg(qm);
  p.nonlinear = (t != nullptr);
  p.t = p.nonlinear ? __ldg(t) : 1.0f;
  float r = fabsf(p.qm);
  if (p.nonlinear) r = expf(p.t * logf(fabsf(p.qm) + 1e-6f));
  p.sat = fabsf(rintf(__fdiv_rn(r, p.d)));
  return p;
}

// magnitude code of one element: |round(p/d)| with the reference's zero / saturation overrides
// (QL:157-159, QL:65-67).  Returns a float so that NaN / overflow can be detected by the caller.
__device__ __forceinline__ float sym_mag(float x, const SymParams& p) {
  const float a = fabsf(x);
  float pw = a;
  if (p.nonlinear) pw = expf(p.t * logf(a));
  float k = fabsf(rintf(__fdiv_rn(pw, p.d)));
  if (a <= 0.0f) k = 0.0f;
  if (a >= p.qm) k = p.sat;
  return k;
}

// signed int code, clamped to int8; sets flag bits
__device__ __forceinline__ int sym_code(float x, const SymParams& p, int& flags) {
  float k = sym_mag(x, p);
  if (x != x) { flags |= kFlagNaN; return 0; }      // sign(NaN) = NaN in the reference
  if (!(k <= 127.0f)) { flags |= (k != k) ? kFlagNaN : kFlagOverflow; k = (k != k) ? 0.0f : 127.0f; }
  int c = (int)k;
  return x < 0.0f ? -c : (x > 0.0f ? c : 0);
}

// Same code, ~4x fewer instructions for the hot epilogue: k = rint(y * (1/d)) agrees with rint(RN(y/d)) unless the
// quotient lies within 2^-21 relative of a rounding boundary, in which case the exact IEEE division decides.
// (|RN(y * RN(1/d)) - RN(y/d)| <= 1.5 * 2^-23 |q|, so the 2^-21 guard band is conservative.)  Saturation:
// for q_m > 0, "|y| >= q_m -> sat" equals clamping to +-sat because RN division and rint are monotone.
struct FastQ {
  float inv_d, d, sat;
  int generic;     // non-linear quantizer, q_m <= 0 or codes beyond int8: use the general path
};
__device__ __forceinline__ FastQ make_fastq(const SymParams& p) {
  FastQ f;
  f.d = p.d;
  f.inv_d = __fdiv_rn(1.0f, p.d);
  f.sat = p.sat;
  f.generic = (p.nonlinear || !(p.qm > 0.0f) || !(p.sat <= 127.0f) || !(p.d > 0.0f)) ? 1 : 0;
  return f;
}
// branch-free fast code; `doubt` is OR-ed with 1 when the exact division has to decide (caller redoes the element)
__device__ __forceinline__ int sym_code_fast(float y, const FastQ& f, int& doubt) {
  const float q = y * f.inv_d;
  float k = rintf(q);
  const float m = fmaf(fabsf(q), 4.8e-7f, fabsf(q - k));      // distance to the rounding boundary is 0.5 - |q - k|
  doubt |= !(m < 0.5f) ? 1 : 0;                                // also true for NaN
  k = fminf(fmaxf(k, -f.sat), f.sat);
  return __float2int_rn(k);
}

// four codes packed little-endian into one word: fast path + exact redo when any of the four is in doubt
__device__ __forceinline__ uint32_t pack4_i8_fwd(int a, int b, int c, int d);
__device__ __forceinline__ uint32_t sym_codes4(float x0, float x1, float x2, float x3, const SymParams& p, const FastQ& f,
                                               int& flags) {
  int doubt = f.generic;
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  if (!f.generic) {
    c0 = sym_code_fast(x0, f, doubt);
    c1 = sym_code_fast(x1, f, doubt);
    c2 = sym_code_fast(x2, f, doubt);
    c3 = sym_code_fast(x3, f, doubt);
  }
  if (doubt) {
    c0 = sym_code(x0, p, flags);
    c1 = sym_code(x1, p, flags);
    c2 = sym_code(x2, p, flags);
    c3 = sym_code(x3, p, flags);
  }
  return pack4_i8_fwd(c0, c1, c2, c3);
}

// fake-quantized value exactly as the reference returns it: sign(x) * (d * round(p/d))
__device__ __forceinline__ float sym_value(float x, const SymParams& p) {
  float k = sym_mag(x, p);
  float v = p.d * k;
  float s = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : ((x == 0.0f) ? 0.0f : x /*NaN*/));
  return s * v;
}

__device__ __forceinline__ float gelu_erf(float x) {
  // torch.nn.GELU() default ('none'): x * Phi(x), Phi(x) = 0.5 * (1 + erf(x / sqrt(2)))   (vit_model.py:173).
  // Single-branch evaluation: h(|x|) = 0.5 * erfc(|x| / sqrt(2)) = exp2(P(t)), t = min(|x| / sqrt(2), 4.2), P a degree-8
  // minimax fit of log2(h) weighted by h * max(1, |x|); Phi = h for x < 0 and 1 - h otherwise.  Max |error| of the
  // result against float64 over [-8, 8]: 3.8e-7 absolute, 1.14e-7 relative to max(1, |x|) - the same as the
  // erf form evaluated in fp32 (4.5e-7 / 1.07e-7) at half the instructions (8 FFMA + 1 MUFU.EX2 instead of ~28).
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.2f);
  float p = -3.6413832276593894e-05f;
  p = fmaf(p, t, 0.000372989394236356f);
  p = fmaf(p, t, -0.0012582261115312576f);
  p = fmaf(p, t, -0.0011454012710601091f);
  p = fmaf(p, t, 0.02857113443315029f);
  p = fmaf(p, t, -0.1486237645149231f);
  p = fmaf(p, t, -0.9183861017227173f);
  p = fmaf(p, t, -1.62791109085083f);
  p = fmaf(p, t, -0.9999999403953552f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(p));
  const float phi = (x < 0.0f) ? h : 1.0f - h;
  return x * phi;
}

// ---- warp / block reductions -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_or(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- 128-bit / 256-bit streaming loads & stores ------------------------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_v4_b32(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void stg_v8_b32(void* p, const uint32_t* r) {   // 32-byte aligned
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack4_i8(int a, int b, int c, int d) {
  return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}

__device__ __forceinline__ uint32_t pack4_i8_fwd(int a, int b, int c, int d) { return pack4_i8(a, b, c, d); }

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
int sm_count();

}  // namespace qvit
