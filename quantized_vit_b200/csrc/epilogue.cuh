// Fused GEMM/conv epilogue shared by every integer-contraction kernel (tcgen05, SIMT, direct conv):
//   y = act( fma(float(acc), (|d_a| * |d_w| * scale_const) * col_scale[n], bias[n]) ) + residual[m, n]
// followed by the store in the requested kind (raw int32 / fp32 / bf16 / int8 re-quantised with the
// consumer layer's quantizer).  Replaces the fp32 tail of QuantizeLinear.forward (quant_layers.py:499:
// F.linear adds the bias), nn.GELU of Mlp.forward (vit_model.py:173), the residual adds of Block.forward
// (vit_model.py:206-207) and the next layer's quantize_act (quant_layers.py:356-381).
#pragma once
#include <cuda_fp16.h>

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
  if (e.out_kind == QVIT_OUT_F16X2) {
    const __half hi = __float2half_rn(y);
    __half* o = reinterpret_cast<__half*>(e.out);
    o[m * e.ldo + n] = hi;
    o[m * e.ldo + e.ldo / 2 + n] = __float2half_rn(y - __half2float(hi));
    return;
  }
  if (e.out_kind == QVIT_OUT_F32) reinterpret_cast<float*>(e.out)[m * e.ldo + n] = y;
  else if (e.out_kind == QVIT_OUT_BF16) reinterpret_cast<__nv_bfloat16*>(e.out)[m * e.ldo + n] = __float2bfloat16_rn(y);
  else reinterpret_cast<int8_t*>(e.out)[m * e.ldo + n] = (int8_t)sym_code(y, *nq, fl);
}

}  // namespace qvit
