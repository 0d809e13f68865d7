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

// Fast-path constants shared by the packed quantizers below: adding 1.5 * 2^23 to a value of magnitude < 2^22 rounds it to
// the nearest integer, ties to even, exactly like rintf; the low BYTE of the sum's bit pattern is then the two's-complement
// int8 code of that integer.
constexpr float kRoundMagic = 12582912.0f;      // 1.5 * 2^23
// low bytes of four 32-bit words -> one little-endian word (3 PRMT)
__device__ __forceinline__ uint32_t pack4_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
__device__ __forceinline__ uint32_t pack4_i8_fwd(int a, int b, int c, int d);

// fake-quantized value exactly as the reference returns it: sign(x) * (d * round(p/d))
__device__ __forceinline__ float sym_value(float x, const SymParams& p) {
  float k = sym_mag(x, p);
  float v = p.d * k;
  float s = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : ((x == 0.0f) ? 0.0f : x /*NaN*/));
  return s * v;
}

__device__ __forceinline__ float gelu_erf(float x) {
  // torch.nn.GELU() default ('none'): x * Phi(x), Phi(x) = 0.5 * (1 + erf(x / sqrt(2)))   (vit_model.py:173).
  // Single-branch evaluation: h(|x|) = 0.5 * erfc(|x| / sqrt(2)) = exp2(P(t)), t = min(|x|, 4.2 * sqrt(2)), P a degree-8
  // minimax fit of log2(h) weighted by h * max(1, |x|) (coefficients pre-scaled by powers of 1/sqrt(2)); then
  // x * Phi(x) = max(x, 0) - |x| * h for either sign.  Max |error| against float64 over [-8, 8] (tools/gelu_fit_check.py):
  // 2.5e-7 absolute, 8.4e-8 relative to max(1, |x|) - better than the erf form evaluated in fp32 (4.5e-7 / 1.06e-7) at
  // 11 instructions (FMNMX, 8 FFMA, MUFU.EX2, FMNMX, FFMA) instead of ~28.
  const float t = fminf(fabsf(x), 5.939696788787842f);
  float p = -2.2758645172871184e-06f;
  p = fmaf(p, t, 3.296791692264378e-05f);
  p = fmaf(p, t, -0.0001572782639414072f);
  p = fmaf(p, t, -0.00020248025248292834f);
  p = fmaf(p, t, 0.007142783608287573f);
  p = fmaf(p, t, -0.052546434104442596f);
  p = fmaf(p, t, -0.45919305086135864f);
  p = fmaf(p, t, -1.1511069536209106f);
  p = fmaf(p, t, -0.9999999403953552f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(p));
  float r;                                    // max(x, 0) that keeps NaN (fmaxf would drop it) and +inf
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(0.0f));
  return fmaf(-t, h, r);                      // t == |x| wherever h is not negligible (|x| <= 5.94)
}

// gelu(x) exactly as gelu_erf computes it, plus d gelu / dx = Phi(x) + x * phi(x) from the SAME h = 0.5 * erfc(|x| / sqrt(2)):
// Phi(x) = 1 - h (x >= 0) or h (x < 0); phi(x) = exp2(-x^2 * log2(e) / 2) / sqrt(2 pi).  One polynomial and two MUFU.EX2
// instead of erff + expf (the fused GELU + quantizer backward was compute-bound at 2 x its HBM time with those).
__device__ __forceinline__ void gelu_erf_grad(float x, float& y, float& dy) {
  const float t = fminf(fabsf(x), 5.939696788787842f);
  float p = -2.2758645172871184e-06f;
  p = fmaf(p, t, 3.296791692264378e-05f);
  p = fmaf(p, t, -0.0001572782639414072f);
  p = fmaf(p, t, -0.00020248025248292834f);
  p = fmaf(p, t, 0.007142783608287573f);
  p = fmaf(p, t, -0.052546434104442596f);
  p = fmaf(p, t, -0.45919305086135864f);
  p = fmaf(p, t, -1.1511069536209106f);
  p = fmaf(p, t, -0.9999999403953552f);
  float h, e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(p));
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(0.0f));
  y = fmaf(-t, h, r);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.7213475204444817f));
  const float cdf = (x >= 0.0f) ? 1.0f - h : h;
  dy = fmaf(x * 0.3989422804014327f, e, cdf);          // (NaN x: y and dy are NaN)
}

// ---- packed fp32 pairs: FFMA2 / FADD2 / FMUL2 of sm_100 process two elements per issue slot ------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2 pk1(float a) { return pk2(a, a); }
__device__ __forceinline__ void unpk2(f32x2 p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// gelu_erf() on pairs, bit-identical to the scalar form: the polynomial runs in t' = -min(|x|, 5.94) with the odd
// coefficients negated (every Horner step is the exact mirror image), so that the last step is fma(t', h, max(x, 0)).
// Four pairs at once, coefficient-major (four independent Horner chains per thread for the scheduler to interleave).
__device__ __forceinline__ void gelu_erf2x4(f32x2 (&x)[4]) {
  f32x2 t[4], p[4];
  float x0[4], x1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unpk2(x[i], x0[i], x1[i]);
    t[i] = pk2(fmaxf(-fabsf(x0[i]), -5.939696788787842f), fmaxf(-fabsf(x1[i]), -5.939696788787842f));
    p[i] = fma2(pk1(-2.2758645172871184e-06f), t[i], pk1(-3.296791692264378e-05f));
  }
  constexpr float c[7] = {-0.0001572782639414072f, 0.00020248025248292834f, 0.007142783608287573f, 0.052546434104442596f,
                          -0.45919305086135864f, 1.1511069536209106f, -0.9999999403953552f};
#pragma unroll
  for (int k = 0; k < 7; ++k) {
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], t[i], pk1(c[k]));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float p0, p1, h0, h1, r0, r1;
    unpk2(p[i], p0, p1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h0) : "f"(p0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h1) : "f"(p1));
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r0) : "f"(x0[i]), "f"(0.0f));
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r1) : "f"(x1[i]), "f"(0.0f));
    x[i] = fma2(t[i], pk2(h0, h1), pk2(r0, r1));
  }
}

// Pair form of the fast quantizer for the GEMM epilogue (no conversion-unit instruction, 2 FFMA2-slots per element):
//   ta = fma(y, inv_hi, 1.5 * 2^23), tb = fma(y, inv_lo, 1.5 * 2^23),  inv_hi/lo = RN(1/d) * (1 +- 3e-7)
// each fma rounds the exact product to the nearest integer (ties to even).  The reference's quotient RN(y / d) lies
// strictly between y * inv_lo and y * inv_hi (both margins exceed 2^-24 after the roundings of the constants), so
// ta == tb means rint(RN(y / d)) is that integer; ta != tb (or NaN / inf, caught because (ta - tb)^2 is then NaN) sends
// the whole 32-column chunk of the warp to the exact path.  Codes = low byte of clamp(ta) (saturation == clamp, q_m > 0).
struct FastQ2 {
  f32x2 inv_hi, inv_lo;
  float t_lo, t_hi;        // 1.5 * 2^23 -+ sat
  float inv_d, d;
  int generic;             // no fast path at all: every element through sym_code()
  int nl;                  // non-linear quantizer on the fast path: |y|^t through MUFU lg2 / ex2, wider doubt band
  float t, qm, satd;       // exponent, q_m, sat * d (the power-domain value every saturated element is replaced by)
};
__device__ __forceinline__ FastQ2 make_fastq2(const SymParams& p) {
  FastQ2 f;
  const float inv_d = __fdiv_rn(1.0f, p.d);
  const bool base_ok = (p.qm > 0.0f) && (p.sat <= 127.0f) && (p.d > 1.0e-10f) && (p.d < 1.0e10f);
  // Non-linear quantizer (QL:40-69): the code is rint(exp(t log|y|) / d).  Fast path: |y|^t = ex2(t * lg2|y|) on the MUFU
  // (relative error ~0.7 * t * |lg2 y| * 2^-22 + 2^-22, i.e. < 1e-5 wherever the code is not trivially 0), the same
  // interval test with a band of +-1.2e-5; elements it cannot decide go to sym_code() (expf / logf / IEEE division), so
  // the codes are those of the exact path bit for bit.  Saturation compares |y| with q_m BEFORE the power (QL:65-67):
  // saturated elements are replaced by sat * d, which rounds to sat under both multipliers.
  f.nl = (p.nonlinear && base_ok && p.t > 0.0f && p.t < 8.0f && p.d >= 1.0e-6f) ? 1 : 0;
  const float band = f.nl ? 1.2e-5f : 3.0e-7f;
  f.inv_hi = pk1(fmaf(inv_d, band, inv_d));
  f.inv_lo = pk1(fmaf(inv_d, -band, inv_d));
  f.t_lo = kRoundMagic - p.sat;
  f.t_hi = kRoundMagic + p.sat;
  f.inv_d = inv_d;
  f.d = p.d;
  f.t = p.t;
  f.qm = p.qm;
  f.satd = p.sat * p.d;
  f.generic = (!base_ok || (p.nonlinear && !f.nl)) ? 1 : 0;
  return f;
}
// power-domain value of one element for the non-linear fast path, sign of y attached (NaN stays NaN -> lands in the doubt set)
__device__ __forceinline__ float nl_power(float y, const FastQ2& f) {
  const float a = fabsf(y);
  float l, pw;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(a));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pw) : "f"(f.t * l));
  pw = (a >= f.qm) ? f.satd : pw;
  return copysignf(pw, y);
}
// Exact code of one element without the division unit, for the elements the interval test could not decide:
// Markstein's correction (q' = q + (y - q d) * RN(1/d), residual exact in an fma) applied twice turns q0 = RN(y * RN(1/d))
// into the correctly rounded quotient RN(y / d) - the first step makes q faithful, the second is then exact by
// Markstein's theorem (no overflow / underflow: callers guarantee 1e-10 < d < 1e10 and |y| < 1e30) - and the add of
// 1.5 * 2^23 rounds it to the nearest integer, ties to even, like torch.round.  Returns the clamped t (code in the low byte).
__device__ __forceinline__ float sym_t_exact(float y, const FastQ2& f) {
  const float q0 = y * f.inv_d;
  const float q1 = fmaf(fmaf(-q0, f.d, y), f.inv_d, q0);
  const float q2 = fmaf(fmaf(-q1, f.d, y), f.inv_d, q1);
  return fminf(fmaxf(q2 + kRoundMagic, f.t_lo), f.t_hi);
}
// four elements (two pairs) -> one word of four int8 codes; dacc accumulates (ta - tb)^2
// (NL is a template parameter and callers branch OUTSIDE their element loops: with the test inside, the linear hot loop
// of the GEMM epilogue carried the non-linear fields in registers and lost ~10 %)
template <bool NL>
__device__ __forceinline__ uint32_t sym_codes4_fast2(f32x2 y01, f32x2 y23, const FastQ2& f, f32x2& dacc) {
  const f32x2 magic = pk1(kRoundMagic), mone = pk1(-1.0f);
  if (NL) {
    float y0, y1, y2, y3;
    unpk2(y01, y0, y1);
    unpk2(y23, y2, y3);
    y01 = pk2(nl_power(y0, f), nl_power(y1, f));
    y23 = pk2(nl_power(y2, f), nl_power(y3, f));
  }
  const f32x2 ta01 = fma2(y01, f.inv_hi, magic), tb01 = fma2(y01, f.inv_lo, magic);
  const f32x2 ta23 = fma2(y23, f.inv_hi, magic), tb23 = fma2(y23, f.inv_lo, magic);
  const f32x2 d01 = fma2(tb01, mone, ta01), d23 = fma2(tb23, mone, ta23);
  dacc = fma2(d01, d01, dacc);
  dacc = fma2(d23, d23, dacc);
  float a0, a1, a2, a3;
  unpk2(ta01, a0, a1);
  unpk2(ta23, a2, a3);
  a0 = fminf(fmaxf(a0, f.t_lo), f.t_hi);
  a1 = fminf(fmaxf(a1, f.t_lo), f.t_hi);
  a2 = fminf(fmaxf(a2, f.t_lo), f.t_hi);
  a3 = fminf(fmaxf(a3, f.t_lo), f.t_hi);
  return pack4_low_bytes(__float_as_uint(a0), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(a3));
}

// 16 consecutive values of one row -> 16 int8 codes (four words), for kernel epilogues that quantize what they just
// computed (attention context -> `proj` codes).  Hot path: packed interval test; rows in doubt: exact Markstein division;
// NaN / inf / generic quantizers: the scalar reference sequence, out of line so that the hot code stays small.
static __device__ __noinline__ uint4 sym_codes16_slow(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7,
                                               float a8, float a9, float a10, float a11, float a12, float a13, float a14,
                                               float a15, SymParams p, int* flags) {
  int fl = 0;
  uint4 w;
  w.x = pack4_i8_fwd(sym_code(a0, p, fl), sym_code(a1, p, fl), sym_code(a2, p, fl), sym_code(a3, p, fl));
  w.y = pack4_i8_fwd(sym_code(a4, p, fl), sym_code(a5, p, fl), sym_code(a6, p, fl), sym_code(a7, p, fl));
  w.z = pack4_i8_fwd(sym_code(a8, p, fl), sym_code(a9, p, fl), sym_code(a10, p, fl), sym_code(a11, p, fl));
  w.w = pack4_i8_fwd(sym_code(a12, p, fl), sym_code(a13, p, fl), sym_code(a14, p, fl), sym_code(a15, p, fl));
  *flags |= fl;
  return w;
}
__device__ __forceinline__ uint4 sym_codes16(const float (&v)[16], const SymParams& p, const FastQ2& f, int& flags) {
  uint32_t w[4];
  bool slow = f.generic != 0;
  if (!slow) {
    f32x2 dacc = pk1(0.0f);
    if (f.nl) {
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = sym_codes4_fast2<true>(pk2(v[4 * j], v[4 * j + 1]), pk2(v[4 * j + 2], v[4 * j + 3]), f, dacc);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = sym_codes4_fast2<false>(pk2(v[4 * j], v[4 * j + 1]), pk2(v[4 * j + 2], v[4 * j + 3]), f, dacc);
    }
    float d0, d1;
    unpk2(dacc, d0, d1);
    if (!(d0 + d1 == 0.0f) && f.nl) {
      slow = true;                                           // non-linear: the scalar sequence decides (out of line)
    } else if (!(d0 + d1 == 0.0f)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        slow = slow || !(fabsf(v[4 * j]) < 1.0e30f) || !(fabsf(v[4 * j + 1]) < 1.0e30f) || !(fabsf(v[4 * j + 2]) < 1.0e30f) ||
               !(fabsf(v[4 * j + 3]) < 1.0e30f);
        w[j] = pack4_low_bytes(__float_as_uint(sym_t_exact(v[4 * j], f)), __float_as_uint(sym_t_exact(v[4 * j + 1], f)),
                               __float_as_uint(sym_t_exact(v[4 * j + 2], f)), __float_as_uint(sym_t_exact(v[4 * j + 3], f)));
      }
    }
  }
  if (slow)
    return sym_codes16_slow(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15], p,
                            &flags);
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// four values -> one word of codes with the same three levels (elementwise kernels: LayerNorm + quantize)
static __device__ __noinline__ uint32_t sym_codes4_slow(float a0, float a1, float a2, float a3, SymParams p, int* flags) {
  int fl = 0;
  const uint32_t w = pack4_i8_fwd(sym_code(a0, p, fl), sym_code(a1, p, fl), sym_code(a2, p, fl), sym_code(a3, p, fl));
  *flags |= fl;
  return w;
}
__device__ __forceinline__ uint32_t sym_codes4_v2(float a0, float a1, float a2, float a3, const SymParams& p, const FastQ2& f,
                                                  int& flags) {
  bool slow = f.generic != 0;
  uint32_t w = 0;
  if (!slow) {
    f32x2 dacc = pk1(0.0f);
    w = f.nl ? sym_codes4_fast2<true>(pk2(a0, a1), pk2(a2, a3), f, dacc) : sym_codes4_fast2<false>(pk2(a0, a1), pk2(a2, a3), f, dacc);
    float d0, d1;
    unpk2(dacc, d0, d1);
    if (!(d0 + d1 == 0.0f) && f.nl) {
      slow = true;
    } else if (!(d0 + d1 == 0.0f)) {
      slow = !(fabsf(a0) < 1.0e30f) || !(fabsf(a1) < 1.0e30f) || !(fabsf(a2) < 1.0e30f) || !(fabsf(a3) < 1.0e30f);
      w = pack4_low_bytes(__float_as_uint(sym_t_exact(a0, f)), __float_as_uint(sym_t_exact(a1, f)),
                          __float_as_uint(sym_t_exact(a2, f)), __float_as_uint(sym_t_exact(a3, f)));
    }
  }
  if (slow) return sym_codes4_slow(a0, a1, a2, a3, p, &flags);
  return w;
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
