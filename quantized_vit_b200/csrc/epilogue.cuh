// Fused GEMM/conv epilogue shared by every integer-contraction kernel (tcgen05, SIMT, direct conv):
//   y = act( fma(float(acc), (|d_a| * |d_w| * scale_const) * col_scale[n], bias[n]) ) + residual[m, n]
// followed by the store in the requested kind (raw int32 / fp32 / bf16 / int8 re-quantised with the
// consumer layer's quantizer).  Replaces the fp32 tail of QuantizeLinear.forward (quant_layers.py:499:
// F.linear adds the bias), nn.GELU of Mlp.forward (vit_model.py:173), the residual adds of Block.forward
// (vit_model.py:206-207) and the next layer's quantize_act (quant_layers.py:356-381).
#pragma once
#include "common.cuh"

namespace qvit {

struct EpiParams {
  int out_kind;
  int act;
  float scale_const;
  int acc_abs_max;
  const float* scale_a;
  const float* scale_w;
  const float* col_scale;
  const float* bias;
  const float* residual;
  int64_t ld_res;
  const float* next_d;
  const float* next_qm;
  const float* next_t;
  int32_t* flags;
  void* out;
  int64_t ldo;
  int M, N;
};

__device__ __forceinline__ float epi_scale(const EpiParams& e) {
  float s = e.scale_const;
  if (e.scale_a) s *= fabsf(__ldg(e.scale_a));
  if (e.scale_w) s *= fabsf(__ldg(e.scale_w));
  return s;
}

__device__ __forceinline__ float epi_value_f(const EpiParams& e, float acc, float scale, int64_t m, int n) {
  // one canonical rounding sequence for every backend: s = scale * col_scale[n], y = fma(acc, s, bias[n])
  float s = scale;
  if (e.col_scale) s *= __ldg(e.col_scale + n);
  float y = e.bias ? fmaf(acc, s, __ldg(e.bias + n)) : acc * s;
  if (e.act == QVIT_ACT_GELU) y = gelu_erf(y);
  else if (e.act == QVIT_ACT_RELU) y = fmaxf(y, 0.0f);
  if (e.residual) y += e.residual[m * e.ld_res + n];
  return y;
}

__device__ __forceinline__ float epi_value(const EpiParams& e, int acc, float scale, int64_t m, int n) {
  return epi_value_f(e, (float)acc, scale, m, n);
}

// scalar store of one element (any kind)
__device__ __forceinline__ void epi_store_one(const EpiParams& e, const SymParams* nq, int acc, float scale, int64_t m,
                                              int n, int& fl) {
  if (e.out_kind == QVIT_OUT_NONE) return;
  if (e.out_kind == QVIT_OUT_I32) {
    reinterpret_cast<int32_t*>(e.out)[m * e.ldo + n] = acc;
    return;
  }
  const float y = epi_value(e, acc, scale, m, n);
  if (e.out_kind == QVIT_OUT_F32) reinterpret_cast<float*>(e.out)[m * e.ldo + n] = y;
  else if (e.out_kind == QVIT_OUT_BF16) reinterpret_cast<__nv_bfloat16*>(e.out)[m * e.ldo + n] = __float2bfloat16_rn(y);
  else reinterpret_cast<int8_t*>(e.out)[m * e.ldo + n] = (int8_t)sym_code(y, *nq, fl);
}

// 32 consecutive columns [n0, n0+32) of row m held in registers (the tcgen05.ld 32x32b shape).
// Vector path when the whole chunk is in range and 16-byte aligned, scalar path otherwise.
__device__ __forceinline__ void epi_store_chunk32(const EpiParams& e, const SymParams* nq, const uint32_t (&acc)[32],
                                                  float scale, int64_t m, int n0, int& fl) {
  const int N = e.N;
  if (n0 >= N) return;
  const bool full = (n0 + 32 <= N);
  if (e.out_kind == QVIT_OUT_I32) {
    int32_t* o = reinterpret_cast<int32_t*>(e.out) + m * e.ldo + n0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) stg_v4_b32(o + 4 * j, acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) o[j] = (int32_t)acc[j];
    }
    return;
  }
  float y[32];
  const bool vec_in = full && (!e.bias || ((reinterpret_cast<uintptr_t>(e.bias + n0) & 15) == 0)) &&
                      (!e.col_scale || ((reinterpret_cast<uintptr_t>(e.col_scale + n0) & 15) == 0)) &&
                      (!e.residual || ((reinterpret_cast<uintptr_t>(e.residual + m * e.ld_res + n0) & 15) == 0));
  if (vec_in) {
    float sc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) sc[j] = scale;
    if (e.col_scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(e.col_scale + n0) + j);
        sc[4 * j] *= c.x; sc[4 * j + 1] *= c.y; sc[4 * j + 2] *= c.z; sc[4 * j + 3] *= c.w;
      }
    }
    if (e.bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(e.bias + n0) + j);
        y[4 * j] = fmaf((float)(int32_t)acc[4 * j], sc[4 * j], c.x);
        y[4 * j + 1] = fmaf((float)(int32_t)acc[4 * j + 1], sc[4 * j + 1], c.y);
        y[4 * j + 2] = fmaf((float)(int32_t)acc[4 * j + 2], sc[4 * j + 2], c.z);
        y[4 * j + 3] = fmaf((float)(int32_t)acc[4 * j + 3], sc[4 * j + 3], c.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = (float)(int32_t)acc[j] * sc[j];
    }
    if (e.act == QVIT_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = gelu_erf(y[j]);
    } else if (e.act == QVIT_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.0f);
    }
    if (e.residual) {
      const float* r = e.residual + m * e.ld_res + n0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 c = *reinterpret_cast<const float4*>(r + 4 * j);   // plain load: `out` may alias `residual`
        y[4 * j] += c.x; y[4 * j + 1] += c.y; y[4 * j + 2] += c.z; y[4 * j + 3] += c.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = (n0 + j < N) ? epi_value(e, (int32_t)acc[j], scale, m, n0 + j) : 0.0f;
  }
  if (e.out_kind == QVIT_OUT_F32) {
    float* o = reinterpret_cast<float*>(e.out) + m * e.ldo + n0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        stg_v4_b32(o + 4 * j, __float_as_uint(y[4 * j]), __float_as_uint(y[4 * j + 1]), __float_as_uint(y[4 * j + 2]),
                   __float_as_uint(y[4 * j + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) o[j] = y[j];
    }
  } else if (e.out_kind == QVIT_OUT_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + m * e.ldo + n0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __nv_bfloat162 p = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&p);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) stg_v4_b32(o + 8 * j, w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) o[j] = __float2bfloat16_rn(y[j]);
    }
  } else {  // QVIT_OUT_I8: the consumer layer's activation quantizer
    int8_t* o = reinterpret_cast<int8_t*>(e.out) + m * e.ldo + n0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        w[j] = pack4_i8(sym_code(y[4 * j], *nq, fl), sym_code(y[4 * j + 1], *nq, fl), sym_code(y[4 * j + 2], *nq, fl),
                        sym_code(y[4 * j + 3], *nq, fl));
      stg_v4_b32(o, w[0], w[1], w[2], w[3]);
      stg_v4_b32(o + 16, w[4], w[5], w[6], w[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) o[j] = (int8_t)sym_code(y[j], *nq, fl);
    }
  }
}

}  // namespace qvit
