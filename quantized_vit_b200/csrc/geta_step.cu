// One-launch update of every quantizer scalar of a model: the quant-parameter half of GETA.step()
// (QViT_with_GETA/only_train_once/optimizer/geta.py:571-772, 787-804; base_optimizer.py:17-86), SURVEY.md section 8f rank 2.
//
// The reference walks ~6 (1,)-shaped nn.Parameters per layer in nested Python loops with substring matching, issues a
// handful of ATen kernels per parameter and calls .item() for every projection bound.  Here a table of device pointers to
// those parameters and their gradients is walked by one small kernel: moment update -> bias-corrected Adam direction (or
// SGD / momentum) -> decoupled weight decay -> step with lr_quant -> projection of d_quant onto
// [d(max_bit), d(min_bit)] (or onto the fixed bit width), d(b) = exp(t * log(max(|q_m|, 1e-10))) / (2^(b-1) - 1).
// Slots per layer: 0 d_quant_wt, 1 q_m_wt, 2 t_quant_wt, 3 d_quant_act, 4 q_m_act, 5 t_quant_act (NULL = absent).
#include "common.cuh"

namespace qvit {

struct GetaStepArgs {
  int variant;            // 0 sgd, 1 adam, 2 adamw
  int mode;               // 0 plain descent (stage 1), 1 range projection (stage 2), 2 fixed bit widths (stage 3)
  int has_wd;
  float lr, lr_quant, wd, beta1, beta2, dampening, safe_guard;
  double bc1, bc2;        // 1 - beta^t (host, double as in the reference)
  int clip;               // clamp gradients to [clip_min, clip_max] first (GETA.grad_clipping, geta.py:160-165)
  float clip_min, clip_max;
  float min_bit_wt, max_bit_wt, min_bit_act, max_bit_act;
};

__device__ __forceinline__ float d_of_bits(float bits, float q_m, const float* t) {
  // _d_quant_helper (geta.py:787-804): q_m = max(|q_m|, 1e-10); exp(t * log(q_m)) / (2^(bits-1) - 1).  The reference
  // evaluates log / exp / the division in double; with a tensor t the product t * log(q_m) is an fp32 tensor op.
  const double q = fmax(fabs((double)q_m), 1e-10);
  double e;
  if (t) e = exp((double)(__ldg(t) * (float)log(q)));
  else e = exp(1.0 * log(q));
  return (float)(e / (exp2((double)bits - 1.0) - 1.0));
}

__global__ void geta_quant_step_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads,
                                       float* __restrict__ m1, float* __restrict__ m2, uint8_t* __restrict__ inited,
                                       const float* __restrict__ fix_bits_wt, const float* __restrict__ fix_bits_act, int layers,
                                       GetaStepArgs a, int32_t* __restrict__ flags) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= layers) return;
  int fl = 0;
#pragma unroll 1
  for (int s = 0; s < 6; ++s) {
    float* p = params[l * 6 + s];
    const float* gp = grads[l * 6 + s];
    if (!p || !gp) continue;                                   // "if p.grad is None: continue" / p_name not in grad_variant
    float g = *gp;
    if (a.clip) g = fminf(fmaxf(g, a.clip_min), a.clip_max);
    if (g != g) fl |= kFlagNaNGrad;
    float pv = *p;
    // ---- compute_grad_variant (base_optimizer.py:40-86)
    if (a.has_wd && a.variant != 2) g += a.wd * pv;
    float gv;
    const int k = l * 6 + s;
    if (a.variant == 0) {
      if (a.beta1 > 0.0f || a.dampening > 0.0f) {
        if (a.beta1 > 0.0f) {
          m1[k] = inited[k] ? m1[k] * a.beta1 + (1.0f - a.dampening) * g : g;
          gv = m1[k];
        } else {
          gv = g;
        }
      } else {
        gv = g;
      }
    } else {
      const float f = (a.beta1 > 0.0f) ? (inited[k] ? m1[k] * a.beta1 + (1.0f - a.beta1) * g : g) : g;
      const float v = (a.beta2 > 0.0f) ? (inited[k] ? m2[k] * a.beta2 + (1.0f - a.beta2) * (g * g) : g * g) : g * g;
      m1[k] = f;
      m2[k] = v;
      const float fh = __fdiv_rn(f, (float)a.bc1), vh = __fdiv_rn(v, (float)a.bc2);   // tensor / python float: fp32 division
      gv = fh / (sqrtf(vh) + a.safe_guard);
    }
    inited[k] = 1;
    // ---- descent (geta.py:571-596 / 598-629 / 667-698 / 723-746)
    const bool is_act = s >= 3;
    if (a.mode == 1 && is_act) {
      // stage 2 runs ..._range_wt first, whose `else` branch also moves the activation quantizer scalars with the MODEL
      // learning rate before ..._range_act moves them again with lr_quant (geta.py:620-629, 689-698): reproduced as is
      if (a.has_wd && a.variant == 2) pv += -a.lr * (a.wd * pv);
      pv += -a.lr * gv;
    }
    if (a.has_wd && a.variant == 2) pv += -a.lr_quant * (a.wd * pv);
    pv += -a.lr_quant * gv;
    *p = pv;
  }
  // ---- projection of the step sizes (after q_m / t of the layer have moved)
  if (a.mode != 0) {
    for (int side = 0; side < 2; ++side) {
      float* d = params[l * 6 + 3 * side];
      const float* qm = params[l * 6 + 3 * side + 1];
      const float* t = params[l * 6 + 3 * side + 2];
      if (!d || !qm) continue;
      if (a.mode == 1) {
        const float lo = d_of_bits(side ? a.max_bit_act : a.max_bit_wt, *qm, t);
        const float hi = d_of_bits(side ? a.min_bit_act : a.min_bit_wt, *qm, t);
        *d = fminf(fmaxf(*d, lo), hi);                         // clamp_(min, max): min first, then max (torch semantics)
      } else {
        const float* fb = side ? fix_bits_act : fix_bits_wt;
        if (fb) *d = d_of_bits(fb[l], *qm, t);
      }
    }
  }
  if (fl && flags) atomicOr(flags, fl);
}

// Saturation codes round(r / |d|) of every layer's weight / activation quantizer (what QuantizeMixin needs to decide
// whether the codes fit the int8 tensor-core pipe): out[2 * l] = weight, out[2 * l + 1] = activation (-1 = absent).
__global__ void quant_sat_levels_kernel(const float* const* __restrict__ params, int layers, float* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= layers) return;
  for (int side = 0; side < 2; ++side) {
    const float* d = params[l * 6 + 3 * side];
    const float* qm = params[l * 6 + 3 * side + 1];
    const float* t = params[l * 6 + 3 * side + 2];
    float v = -1.0f;
    if (d && qm) v = load_sym_params(d, qm, t).sat;            // same fp32 sequence as the quantize kernels
    out[2 * l + side] = v;
  }
}

}  // namespace qvit

using namespace qvit;

extern "C" int qvit_quant_sat_levels(const float* const* params, int layers, float* out, qvit_stream_t stream) {
  QVIT_REQUIRE(params && out && layers >= 0, "qvit_quant_sat_levels: null pointer");
  if (layers == 0) return QVIT_OK;
  quant_sat_levels_kernel<<<(layers + 127) / 128, 128, 0, (cudaStream_t)stream>>>(params, layers, out);
  return check_launch("qvit_quant_sat_levels");
}

extern "C" int qvit_geta_quant_step(float* const* params, const float* const* grads, float* m1, float* m2, uint8_t* inited,
                                    const float* fix_bits_wt, const float* fix_bits_act, int layers, int variant, int mode,
                                    float lr, float lr_quant, int has_wd, float wd, float beta1, float beta2, float dampening,
                                    double bc1, double bc2, float safe_guard, int clip, float clip_min, float clip_max,
                                    float min_bit_wt, float max_bit_wt, float min_bit_act, float max_bit_act, int32_t* flags,
                                    qvit_stream_t stream) {
  QVIT_REQUIRE(params && grads && m1 && m2 && inited && layers >= 0, "qvit_geta_quant_step: null pointer");
  QVIT_REQUIRE(variant >= 0 && variant <= 2 && mode >= 0 && mode <= 2, "qvit_geta_quant_step: bad variant / mode");
  if (layers == 0) return QVIT_OK;
  GetaStepArgs a;
  a.variant = variant; a.mode = mode; a.has_wd = has_wd; a.lr = lr; a.lr_quant = lr_quant; a.wd = wd; a.beta1 = beta1;
  a.beta2 = beta2; a.dampening = dampening; a.safe_guard = safe_guard; a.bc1 = bc1; a.bc2 = bc2; a.clip = clip;
  a.clip_min = clip_min; a.clip_max = clip_max; a.min_bit_wt = min_bit_wt; a.max_bit_wt = max_bit_wt;
  a.min_bit_act = min_bit_act; a.max_bit_act = max_bit_act;
  geta_quant_step_kernel<<<(layers + 127) / 128, 128, 0, (cudaStream_t)stream>>>(params, grads, m1, m2, inited, fix_bits_wt,
                                                                                  fix_bits_act, layers, a, flags);
  return check_launch("qvit_geta_quant_step");
}
