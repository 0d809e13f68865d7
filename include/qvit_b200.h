/* libqvit_b200 - C ABI of the B200-native (sm_100a) quantized Conv2d/Linear hot path.
 *
 * The reference (LongAoTianxia/Quantized_ViT) is pure Python/PyTorch and has NO FFI: its "operator
 * interface" for this path is a set of nn.Module / autograd.Function classes.  Each entry point below
 * therefore names the reference function whose ATen op chain it replaces (file:line relative to the
 * reference root; QL = QViT_with_GETA/only_train_once/quantization/quant_layers.py,
 * QU = "4-bit quantization/quant_ultra.py", QZ = "4-bit quantization/quantization.py",
 * MEM = "4-bit quantization/qnn_mem_process.py").  The Python binding a reference maintainer would add is
 * shown in INTEGRATION.md and implemented in quantized_vit_b200/_lib.py (ctypes).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated; the caller (PyTorch) owns all memory;
 *  - quantizer parameters (d_quant, q_m, t_quant) are passed as device pointers to the (1,) fp32
 *    nn.Parameter storage, so no call ever synchronises with the host;
 *  - `stream` is a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); functions only enqueue
 *    work, never synchronise, never allocate -> they are CUDA-graph capturable;
 *  - return value: 0 = ok, otherwise an error code; the message is qvit_last_error() (thread-local);
 *  - `flags` (optional, may be NULL) is an int32 device word that kernels OR bits into:
 *      QVIT_FLAG_NAN (1)       a NaN reached a quantizer (the reference would propagate NaN, QL:160)
 *      QVIT_FLAG_OVERFLOW (2)  a code magnitude exceeded 127 (int8 pipe not applicable)
 *      QVIT_FLAG_NAN_GRAD (4)  NaN in a reduced gradient (reference raises NanInGradientError, QL:189-204)
 */
#ifndef QVIT_B200_H_
#define QVIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QVIT_ABI_VERSION 1

#define QVIT_OK 0
#define QVIT_ERR_INVALID 1
#define QVIT_ERR_CUDA 2
#define QVIT_ERR_UNSUPPORTED 3

#define QVIT_FLAG_NAN 1
#define QVIT_FLAG_OVERFLOW 2
#define QVIT_FLAG_NAN_GRAD 4

typedef void* qvit_stream_t; /* cudaStream_t */

/* ------------------------------------------------------------------ library */
int qvit_abi_version(void);
const char* qvit_last_error(void);
/* SM count and compute capability of the current device (host call). */
int qvit_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ GETA symmetric quantizers
 * Replaces SymQuantizerLinear.forward (QL:136-161) when t == NULL and SymQuantizerNonLinear.forward
 * (QL:40-69) otherwise.  x is a [rows, cols] fp32 matrix with row pitch ld_x (elements).            */

/* integer codes: codes[r*ld_codes + c] = sign(x) * |round(p/d)| (saturated to round(r/d) where |x| >= q_m),
 * int8; columns cols..ld_codes-1 of every row are written as 0 (GEMM K padding).                      */
int qvit_quantize_sym(const float* x, int64_t rows, int64_t cols, int64_t ld_x,
                      const float* d, const float* q_m, const float* t /* NULL = linear */,
                      int8_t* codes, int64_t ld_codes, int32_t* flags, qvit_stream_t stream);

/* fake-quantized fp32 values, bit-for-bit what the reference autograd.Function returns (n elements). */
int qvit_fake_quantize_sym(const float* x, int64_t n, const float* d, const float* q_m, const float* t,
                           float* out, qvit_stream_t stream);

/* Backward of both Functions (QL:163-205 / QL:71-125) in ONE pass over (x, g):
 *   grad_x[i]      = g[i] if clip_lo < x[i] < clip_hi else 0            (STE through saturation)
 *   grad_scalars[0] += sum g*sign(x)*(round(p/d) - p/d | saturated residual | 0)      (d_quant)
 *   grad_scalars[1] += sum g*sign(x)*[|x| > q_m] (* t*exp((t-1)log(|q_m|+1e-6)))      (q_m)
 *   grad_scalars[2] += sum g*sign(x)*(p*log|x| | r*log(|q_m|+1e-6) | 0)   (t_quant; only if t != NULL)
 * grad_scalars must be zeroed by the caller; block partials are combined with fp32 atomics after an fp32
 * warp-shuffle/shared-memory tree.  NaN in a reduced scalar sets QVIT_FLAG_NAN_GRAD in *flags.
 * grad_x may be NULL (weight-side call when only the scalars and a masked copy are wanted separately).   */
int qvit_sym_backward(const float* x, const float* g, int64_t n,
                      const float* d, const float* q_m, const float* t,
                      float clip_lo, float clip_hi,
                      float* grad_x, float* grad_scalars, int32_t* flags, qvit_stream_t stream);

/* The same two kernels with Mlp.forward's nn.GELU (VIT:173) fused in, for the QAT step of the layer BEHIND the GELU (fc2):
 * codes = quantize_act(gelu(x)) from the pre-activation x (contiguous, n elements; the fp32 activation is never written), and the
 * backward that recomputes gelu(pre), applies the STE mask / scalar-gradient sums of qvit_sym_backward to it and writes the
 * gradient with respect to the PRE-activation (x gelu'(pre)).                                                                   */
int qvit_gelu_quantize_sym(const float* x, int64_t n, const float* d, const float* q_m, const float* t, int8_t* codes,
                           int32_t* flags, qvit_stream_t stream);
int qvit_gelu_sym_backward(const float* pre, const float* g, int64_t n, const float* d, const float* q_m, const float* t,
                           float clip_lo, float clip_hi, float* grad_pre, float* grad_scalars, int32_t* flags,
                           qvit_stream_t stream);

/* max|x| over n elements -> out[0] (initialize_quant_layer, QL:423: q_m = max|W|).  out is overwritten. */
int qvit_absmax(const float* x, int64_t n, float* out, qvit_stream_t stream);

/* ------------------------------------------------------------------ conv lowering
 * QuantizeConv2d.forward (QL:575-587): activation quantize (as qvit_quantize_sym) fused with im2col.
 * x is NCHW fp32; cols is [B*OH*OW, ld_cols] int8 with K = C*kh*kw ordered (c, kh, kw) =
 * weight.reshape(O, -1); zero padding and columns K..ld_cols-1 get code 0.                            */
int qvit_im2col_quantize_sym(const float* x, int B, int C, int H, int W,
                             int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                             const float* d, const float* q_m, const float* t,
                             int8_t* cols, int64_t ld_cols, int32_t* flags, qvit_stream_t stream);

/* ------------------------------------------------------------------ UltraNet (DoReFa) quantizers */
/* max|tanh(w)| -> out[0]  (weight_quantize_fn.forward, QU:50-53). */
int qvit_ultra_tanh_absmax(const float* w, int64_t n, float* out, qvit_stream_t stream);
/* codes = round(tanh(w)/max * (2^(w_bit-1)-1)); 2 <= w_bit <= 8.  export_rounding = 0: the torch forward
 * (uniform_quantize(k=1) is sign() at w_bit == 2, QU:15-16); 1: the NumPy export, which always rounds (QZ:24-31). */
int qvit_ultra_quantize_weight(const float* w, int64_t n, int w_bit, int export_rounding, const float* max_tanh,
                               int8_t* codes, qvit_stream_t stream);
/* activation_quantize_fn.forward (QU:66-73): codes = round(clamp(x,0,1) * (2^a_bit-1)), uint8;
 * values = codes / (2^a_bit-1) written to out_values if non-NULL (either output may be NULL).        */
int qvit_ultra_quantize_act(const float* x, int64_t n, int a_bit, uint8_t* codes, float* out_values,
                            qvit_stream_t stream);
/* uniform_quantize(k).forward (QU:12-20): identity (k = 32), sign (k = 1), else round(x*n)/n with n = 2^k-1
 * (k = 0 gives n = 0 -> NaN, the reference's own 1-bit-weight quirk, QU:36 with QU:18-19).             */
int qvit_uniform_quantize(const float* x, int64_t n, int k, float* out, qvit_stream_t stream);
/* nn.BatchNorm2d(eval) as (scale, bias) + activation_quantize_fn + optional 2x2 max-pool (MM:74-76) on an NCHW
 * fp32 map, written as NHWC uint8 codes with channel pitch ldc (padding channels are NOT written).
 * scale = bias = NULL: codes = round(clamp(x,0,1)*levels) only (image -> 8-bit input codes).           */
int qvit_ultra_bn_act_pool_nchw(const float* x, int B, int C, int H, int W, const float* scale, const float* bias,
                                int levels, int pool, uint8_t* out_codes, int ldc, qvit_stream_t stream);
/* Conv2d_Q.forward (QU:85-89) with arbitrary fp32 input: y = conv2d(x, codes / w_levels) + bias, NCHW fp32,
 * groups == 1, w_levels = 2^(w_bit-1)-1 (the fp32 quotient codes/w_levels is bit-for-bit the reference's w_q,
 * QU:18-19).  The input is NOT quantised by the reference layer (the first UltraNet layer sees the image). */
int qvit_conv2d_f32_wcodes(const float* x, int B, int C, int H, int W,
                           const int8_t* w_codes, int O, int kh, int kw,
                           int sh, int sw, int ph, int pw, int dh, int dw,
                           float w_levels, const float* bias, float* y, qvit_stream_t stream);
/* One fused integer UltraNet layer (MM:71-125 pattern conv -> BatchNorm2d(eval) -> act-quant [-> 2x2 max-pool]):
 * in_codes  [B, H, W, C]  uint8 NHWC activation codes (0..2^a_bit-1)
 * w_codes   [O, kh, kw, C] int8 (K ordered (kh, kw, c) - the order MEM:152-154 exports)
 * y = acc * acc_scale * bn_scale[o] + bn_bias[o];  out_codes = round(clamp(y,0,1) * out_levels), max-pooled 2x2
 * if pool != 0 (pooling commutes with the monotone code map).  out_codes [B, OH, OW, O] uint8.
 * If out_f32 != NULL the un-quantised y is written there instead as NCHW fp32 (last layer, MM:123).   */
int qvit_ultra_conv_bn_act(const uint8_t* in_codes, int B, int H, int W, int C,
                           const int8_t* w_codes, int O, int kh, int kw, int pad,
                           float acc_scale, const float* bn_scale, const float* bn_bias,
                           int out_levels, int pool, uint8_t* out_codes, float* out_f32,
                           qvit_stream_t stream);

/* The same fused layer as an IMPLICIT GEMM on the tcgen05 int8 tensor-core pipe (M = output pixels, N = O, K = taps * C,
 * unsigned activation codes x signed weight codes -> int32 in TMEM), for C in {16, 32, 64, 128}, stride 1.  w_packed: int8
 * [O_pad, K_pad], k = (ky * kw + kx) * C + c (the [O, kh, kw, C] codes flattened), O_pad = O rounded up to 16, K_pad = kh*kw*C
 * rounded up to 128, zero padded.  Weights stay resident in shared memory while a persistent CTA walks its pixel tiles; the
 * im2col tile is gathered in shared memory, never in HBM.  Codes are identical to qvit_ultra_conv_bn_act.   */
int qvit_ultra_conv_tc(const uint8_t* in_codes, int B, int H, int W, int C, const int8_t* w_packed, int O, int kh, int kw,
                       int pad, float acc_scale, const float* bn_scale, const float* bn_bias, int out_levels, int pool,
                       uint8_t* out_codes, float* out_f32, qvit_stream_t stream);
/* QuantizeConv2d.forward (QL:575-587) as the same implicit GEMM on SIGNED int8 activation codes in NHWC [B, H, W, C], C in
 * {16, 32, 64, 128}, any stride / symmetric padding / dilation, groups == 1: out fp32 NCHW [B, O, OH, OW] =
 * acc * |*scale_a| * |*scale_w| + bias[o].  w_packed as for qvit_ultra_conv_tc ([O, kh, kw, C] codes flattened and padded). */
int qvit_conv2d_i8_tc(const int8_t* a_codes_nhwc, int B, int H, int W, int C, const int8_t* w_packed, int O, int kh, int kw,
                      int sh, int sw, int pad, int dh, int dw, const float* scale_a, const float* scale_w, const float* bias,
                      float* out_nchw, qvit_stream_t stream);
/* ------------------------------------------------------------------ BN fold / pack (one-time)
 * mode 0: nn.BatchNorm2d eval  scale = gamma/sqrt(var+eps),   bias = beta - mean*scale   (MM:74.. + F.batch_norm)
 * mode 1: export fold          scale = gamma/(sqrt(var)+eps), bias = beta - mean/(sqrt(var)+eps)*gamma (QZ:34-46) */
int qvit_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                 int mode, int C, float* scale, float* bias, qvit_stream_t stream);
/* bn_act_quantize_int (QZ:68-89): int32 (inc, bias) thresholds.  The NumPy reference computes in the dtype of
 * the arrays it is handed (float32 from an .npz export, float64 in its own demo): is_f64 selects which; the
 * four inputs are device arrays of that dtype.                                                         */
int qvit_bn_act_quantize_int(const void* gamma, const void* beta, const void* mean, const void* var, int is_f64,
                             double eps, int w_bit, int in_bit, int out_bit, int l_shift, int C,
                             int32_t* inc, int32_t* bias, qvit_stream_t stream);
/* 4-bit pack, element i of a run in bits [4i, 4i+4), two's complement (array_to_string, MEM:11-24).
 * n must be even; packed has n/2 bytes.                                                               */
int qvit_pack_int4(const int8_t* codes, int64_t n, uint8_t* packed, qvit_stream_t stream);
int qvit_unpack_int4(const uint8_t* packed, int64_t n, int is_signed, int8_t* codes, qvit_stream_t stream);
/* FPGA (HLS) weight layout, QNNLayerMemProcess.conv / w_to_hls_array (MEM:84-130, 152-157): codes [O, I, kh, kw] ->
 * rows reordered to (kh, kw, I), cut into runs of `simd` codes (ragged last run allowed), each run packed as above into one
 * word of simd * w_bit <= 64 bits; word (oc, j) is stored at words[(oc % pe) * tiles + (oc / pe) * runs + j],
 * runs = ceil(kh*kw*I / simd), tiles = runs * O / pe.  O must be a multiple of pe.                       */
int qvit_pack_hls_weights(const int8_t* codes, int O, int I, int kh, int kw, int w_bit, int simd, int pe,
                          unsigned long long* words, qvit_stream_t stream);

/* ------------------------------------------------------------------ GETA quant-parameter step
 * The quantizer-scalar half of GETA.step() (optimizer/geta.py:571-772, 787-804; base_optimizer.py:17-86) in ONE launch:
 * optional gradient clamp (geta.py:160-165), SGD / momentum / Adam / AdamW direction with the reference's moment
 * initialisation (first buffer = grad) and bias corrections bc = 1 - beta^t, decoupled weight decay, step with lr_quant,
 * and the projection of every d_quant onto [d(max_bit), d(min_bit)] (mode 1) or onto d(fixed bit) (mode 2),
 * d(b) = exp(t * log(max(|q_m|, 1e-10))) / (2^(b-1) - 1).  mode 0 = plain descent (stage 1).
 * params / grads: device arrays of `layers * 6` device pointers to the (1,) fp32 scalars, slot order d_quant_wt, q_m_wt,
 * t_quant_wt, d_quant_act, q_m_act, t_quant_act; NULL = absent (or no gradient this step).  m1 / m2 / inited: persistent
 * state, `layers * 6` elements, zero-initialised by the caller.  variant: 0 sgd, 1 adam, 2 adamw.  In mode 1 the
 * activation scalars are first moved with the MODEL learning rate `lr` as well - the reference's range_wt `else` branch
 * does exactly that (geta.py:620-629).  fix_bits_*: [layers] fp32 (mode 2) or NULL. */
int qvit_geta_quant_step(float* const* params, const float* const* grads, float* m1, float* m2, uint8_t* inited,
                         const float* fix_bits_wt, const float* fix_bits_act, int layers, int variant, int mode,
                         float lr, float lr_quant, int has_wd, float wd, float beta1, float beta2, float dampening,
                         double bc1, double bc2, float safe_guard, int clip, float clip_min, float clip_max,
                         float min_bit_wt, float max_bit_wt, float min_bit_act, float max_bit_act, int32_t* flags,
                         qvit_stream_t stream);

/* Saturation codes round(r / |d|), r = |q_m| or exp(t log(|q_m| + 1e-6)) (QL:159 / QL:62-67), of `layers` weight /
 * activation quantizers through the same pointer table: out[2l] = weight, out[2l+1] = activation, -1 where absent.
 * Lets QuantizeMixin follow GETA's bit-width walk (geta.py:895-900) and pick the int8 or the wide path without a host
 * synchronisation (the result is read back asynchronously, one step late).                                   */
int qvit_quant_sat_levels(const float* const* params, int layers, float* out, qvit_stream_t stream);

/* ------------------------------------------------------------------ QuantLinear GEMM
 * Replaces nn.functional.linear(x_q, w_q, bias) on fake-quant values (QL:499, QU:220) by the exact integer
 * contraction  acc[m,n] = sum_k A[m,k] * Wc[n,k]  (int8 x int8 -> int32, tcgen05.mma kind::i8, TMEM
 * accumulators) and a fused epilogue.  A is [M, lda] int8 (or uint8 if a_unsigned), Wc is [N, ldw] int8,
 * both K-major; lda, ldw multiples of 16 and base pointers 16-byte aligned for the tensor-core backend. */
enum {
  QVIT_OUT_I32 = 0,   /* raw accumulators (parity tests)                           out: int32  [M, ldo] */
  QVIT_OUT_F32 = 1,   /* y                                                       out: fp32   [M, ldo] */
  QVIT_OUT_BF16 = 2,  /* y rounded to bf16                                         out: bf16   [M, ldo] */
  QVIT_OUT_I8 = 3,    /* y re-quantised with (next_d, next_qm[, next_t]) -> int8 codes  out: int8 [M, ldo] */
  QVIT_OUT_NONE = 4,  /* nothing is stored (out may be NULL): times the TMA/MMA main loop alone (bench only)  */
  QVIT_OUT_F16X2 = 5  /* y as TWO fp16 planes, hi = fp16(y), lo = fp16(y - hi): 22 significant bits in the bytes of one
                         fp32, directly consumable by fp16 tensor-core MMAs (qvit_attention_f16x2).  out: fp16 [M, ldo],
                         hi in columns [0, N), lo in columns [ldo/2, ldo/2 + N); N % 32 == 0, ldo % 16 == 0.  The caller
                         keeps |y| < 65504 through col_scale (a power of two per column scales y exactly).      */
};
enum { QVIT_ACT_NONE = 0, QVIT_ACT_GELU = 1, QVIT_ACT_RELU = 2 };
enum { QVIT_GEMM_AUTO = 0, QVIT_GEMM_TCGEN05 = 1, QVIT_GEMM_SIMT = 2 };

typedef struct qvit_epilogue {
  int32_t out_kind;        /* QVIT_OUT_*                                                              */
  int32_t act;             /* QVIT_ACT_* applied to y before residual / requant                        */
  const float* scale_a;    /* (1,) device: |.| is taken  (d_quant_act)  or NULL = 1                    */
  const float* scale_w;    /* (1,) device: |.| is taken  (d_quant_wt)   or NULL = 1                    */
  float scale_const;       /* host multiplier (1.0; 1/(7*15) for UltraNet codes)                      */
  int32_t acc_abs_max;     /* caller's promise |accumulator| <= acc_abs_max (sat_a * sat_w * K), 0 = unknown.
                              Below 2^22 the epilogue converts int32 -> fp32 without the conversion unit.  */
  const float* col_scale;  /* [N] per-output-channel multiplier (folded BN) or NULL                    */
  const float* bias;       /* [N] or NULL                                                              */
  const float* residual;   /* [M, ld_res] fp32 added after act, or NULL                                */
  int64_t ld_res;
  const float* next_d;     /* QVIT_OUT_I8: quantizer of the consumer layer                             */
  const float* next_qm;
  const float* next_t;     /* NULL = linear                                                            */
  int32_t* flags;          /* optional                                                                 */
} qvit_epilogue_t;
/* y = act( float(acc) * (|scale_a|*|scale_w|*scale_const) * col_scale[n] + bias[n] ) + residual[m,n]  */

/* tile mode of the tensor-core backend: 0 = automatic (CTA pairs, tcgen05 cta_group::2, once the problem has a full
 * wave of [256 x 256] tiles), 1 = single-CTA [128 x BN] tiles, 2 = CTA pairs whenever BN = 256.  Process-wide; tests / benches. */
int qvit_gemm_set_cta_group(int cta_group);
/* developer hook: with mode + 50 the tensor-core kernel stamps clock64 timelines of CTA 0 (8 slots per tile, see
 * gemm_tc.cu); copies up to n stamps to host memory (synchronises the device), returns the count or -1. */
int qvit_gemm_read_profile(long long* host, int n);

int qvit_gemm_i8(const void* a, int64_t lda, int a_unsigned,
                 const int8_t* w, int64_t ldw,
                 int M, int N, int K,
                 void* out, int64_t ldo,
                 const qvit_epilogue_t* epi, int backend, qvit_stream_t stream);

/* ------------------------------------------------------------------ QAT gradient GEMMs
 * Backward of F.linear(x_q, w_q) in QuantizeLinear.forward (QL:499): grad_x_q = g @ w_q, grad_w_q = g^T @ x_q.  The
 * reference never quantizes g, so the int8 pipe does not apply; g is split EXACTLY into three bf16 planes and the
 * integer codes become one bf16 plane (|code| <= 127 is exact), and (g1 + g2 + g3) * codes runs on tcgen05 kind::f16
 * with fp32 accumulation in TMEM.
 *
 * qvit_split3_bf16: x [rows, cols] fp32 -> out bf16, three planes side by side.
 *   transpose = 0: out [rows, 3 * plane_cols], plane p in columns [p*plane_cols, ...), plane_cols >= cols, zero padded;
 *   transpose = 1: out [cols, 3 * plane_cols], element (c, r) = plane_p(x[r][c]), plane_cols >= rows, zero padded.
 * qvit_codes_to_bf16_t: codes [rows, cols] int8 -> out [cols, out_cols] bf16 (transposed), out_cols >= rows, zero padded.
 * plane_cols / out_cols must be multiples of 64.
 * qvit_gemm_bf16_split: out[M, N] (fp32) = epilogue( sum_p A_p[M, K] * B[N, K]^T ); A = [M, lda >= planes*Kp] bf16,
 *   B = [N, ldb >= Kp] bf16, Kp = K rounded up to 64; epilogue as qvit_gemm_i8 with QVIT_OUT_F32 / QVIT_ACT_NONE. */
int qvit_split3_bf16(const float* x, int64_t rows, int64_t cols, int64_t ld_x, int transpose, void* out, int64_t plane_cols,
                     qvit_stream_t stream);

/* Both forms of the gradient operand of a QAT linear layer, and its bias gradient, from one read of g (fp32 [rows, cols], pitch
 * ld_g): rows_out = bf16 [rows, 3 * row_plane_cols] (as qvit_split3_bf16 transpose = 0; NULL = not needed), trans_out = bf16
 * [cols, 3 * trans_plane_cols] (as transpose = 1; NULL = not needed), colsum[c] = sum over rows of g[:, c] (NULL = not needed; summed in a fixed order).
 * partial: workspace fp32 [ceil(trans_plane_cols / 256), cols], required with colsum.  Plane widths: multiples of 64.          */
int qvit_grad_prep(const float* g, int64_t rows, int64_t cols, int64_t ld_g, void* rows_out, int64_t row_plane_cols, void* trans_out,
                   int64_t trans_plane_cols, float* partial, float* colsum, qvit_stream_t stream);
int qvit_codes_to_bf16_t(const int8_t* codes, int64_t rows, int64_t cols, int64_t ld, void* out, int64_t out_cols,
                         qvit_stream_t stream);
int qvit_gemm_bf16_split(const void* a_planes, int64_t lda, int planes, const void* b, int64_t ldb, int M, int N, int K,
                         float* out, int64_t ldo, const qvit_epilogue_t* epi, qvit_stream_t stream);

/* grad_w_q = g^T x_q without a transposed copy of either operand: out[N_out, K_in] (fp32) = |scale| * sum_p G_p^T X, G = `planes`
 * bf16 planes of g side by side ([tokens, planes * plane_cols] row-major, as qvit_split3_bf16 / qvit_grad_prep write the row form),
 * X = bf16 [tokens, ld_x >= K_in] (qvit_codes_to_bf16); both are read as MN-major tcgen05 operands.  Plain fp32 output
 * (scale_const / scale_a / scale_w of the epilogue apply; no activation, no residual).                                            */
int qvit_gemm_bf16_split_t(const void* g_planes, int64_t ld_g, int planes, int64_t plane_cols, const void* x, int64_t ld_x,
                           int64_t tokens, int N_out, int K_in, float* out, int64_t ldo,
                           float* workspace /* optional, fp32 [N_out, ldo]: lets small outputs split the contraction four ways */,
                           const qvit_epilogue_t* epi, qvit_stream_t stream);
/* int8 codes [rows, ld] -> bf16 [rows, out_cols] (same orientation, columns >= cols zero; out_cols a multiple of 64). */
int qvit_codes_to_bf16(const int8_t* codes, int64_t rows, int64_t cols, int64_t ld, void* out, int64_t out_cols, qvit_stream_t stream);

/* ------------------------------------------------------------------ glue fused with the quantizer
 * ("next" rows of SURVEY.md section 8f, built on the same quantizer device function)
 * LayerNorm (vit_model.py:206-207, eps inside sqrt, biased variance) followed by the consumer layer's
 * activation quantizer: codes = Q(LN(x) * gamma + beta).  x [rows, cols] fp32 contiguous.             */
int qvit_layernorm_quantize(const float* x, int64_t rows, int cols, const float* gamma, const float* beta,
                            float eps, const float* d, const float* q_m, const float* t,
                            int8_t* codes, int64_t ld_codes, float* ln_out /* optional fp32 copy */,
                            int32_t* flags, qvit_stream_t stream);
/* Token sequence of the ViT (cat(cls_token, x) + pos_embed, VIT:295-305) from the patch-embedding output: h[b, 0, :] = cls + pos[0],
 * h[b, 1 + p, :] = tok[b * P + p, :] + pos[1 + p, :].  tok fp32 [B * P, D], pos fp32 [P + 1, D], cls fp32 [D], h fp32 [B, P + 1, D];
 * D a multiple of 4, 16-byte aligned tensors.                                                                                   */
int qvit_embed_assemble(const float* tok, const float* pos, const float* cls, int B, int P, int D, float* h, qvit_stream_t stream);

/* LayerNorm forward with saved statistics, and its backward, for the caller of the QAT step (Block.norm1 / norm2, VIT:202-208):
 * y = (x - mean) * rstd * gamma + beta; gx = rstd * (g' - mean(g') - xhat * mean(g' * xhat)), g' = gy * gamma; dgamma = sum gy * xhat,
 * dbeta = sum gy (both overwritten).  cols must be a multiple of 128, <= 1024; all tensors fp32, 16-byte aligned.           */
int qvit_layernorm_fwd(const float* x, int64_t rows, int cols, const float* gamma, const float* beta, float eps, float* y,
                       float* mean, float* rstd, qvit_stream_t stream);
int qvit_layernorm_bwd(const float* x, const float* gy, int64_t rows, int cols, const float* gamma, const float* mean,
                       const float* rstd, const float* add /* optional: gradient reaching x over the residual path */, float* gx,
                       float* dgamma, float* dbeta, qvit_stream_t stream);
/* bf16 -> codes (attention output feeding `proj`): same quantizer on bf16 input widened to fp32. */
int qvit_quantize_sym_bf16(const void* x_bf16, int64_t rows, int64_t cols, int64_t ld_x,
                           const float* d, const float* q_m, const float* t,
                           int8_t* codes, int64_t ld_codes, int32_t* flags, qvit_stream_t stream);

/* softmax(Q K^T * scale) V of ViTAttention.forward (vit_model.py:141-149; NOT quantized by the reference, fp32) in
 * fp32-equivalent precision on the tensor cores (3xTF32 split, tcgen05 kind::tf32, fp32 accumulation in TMEM).
 * qkv: [B, T, 3, H, head_dim] fp32 contiguous (the output of the qkv QuantizeLinear, vit_model.py:133);
 * out: [B, T, H*head_dim] fp32 (the layout `proj` consumes, vit_model.py:149).  head_dim == 64 and T <= 208.   */
int qvit_attention_f32(const float* qkv, int B, int T, int H, int head_dim, float scale, float* out,
                       qvit_stream_t stream);
/* test hook: same, and dumps raw scores (cols 0..207) and un-normalised probabilities (cols 208..415) of every query
 * row into dbg [B, H, 256, 512] fp32 (caller-zeroed; cols 416..479 raw O, 480..482 row sums and 1/sum).                                                            */
/* Same product with the consumer layer's quantize_act (QL:356-381; `proj` in ViTAttention.forward, VIT:151) fused into
 * the epilogue: codes [B*T, ld_codes] int8, column h * head_dim + i, ld_codes >= H * head_dim and a multiple of 16 (padding
 * columns are not written).  out (fp32 context) is optional.  d / q_m / t as qvit_quantize_sym. */
int qvit_attention_quantize_sym(const float* qkv, int B, int T, int H, int head_dim, float scale, const float* d,
                                const float* q_m, const float* t, int8_t* codes, int64_t ld_codes, float* out,
                                int32_t* flags, qvit_stream_t stream);
int qvit_attention_f32_debug(const float* qkv, int B, int T, int H, int head_dim, float scale, float* out,
                             float* dbg, int diag, qvit_stream_t stream);

/* Attention core from the TWO-PLANE fp16 form of qkv (what qvit_gemm_i8 writes with QVIT_OUT_F16X2): planes fp16 [B, T, ld],
 * hi plane in columns [0, 3*H*64) (q | k | v, head-major inside a part), lo plane in [plane_off, plane_off + 3*H*64);
 * exp_q / exp_k / exp_v = the powers of two q / k / v were multiplied by.  Products are evaluated as hi*hi' + hi*lo' + lo*hi' on
 * tcgen05 kind::f16 with fp32 accumulation (fp32-equivalent accuracy, half the tensor-core work of the 3 x bf16 split and no
 * conversion work: operand tiles arrive by TMA).  codes (the consumer's quantize_act, QL:356-381) and / or out (fp32 context
 * [B, T, H*64]) as in qvit_attention_quantize_sym; prof: optional device buffer of 256 clock stamps (developer timeline).
 * head_dim == 64, T <= 208.                                                                                              */
int qvit_attention_f16x2(const void* planes, int64_t ld, int plane_off, int B, int T, int H, int head_dim, float scale,
                         int exp_q, int exp_k, int exp_v, const float* d, const float* q_m, const float* t, int8_t* codes,
                         int64_t ld_codes, float* out, int32_t* flags, long long* prof, qvit_stream_t stream);
/* fp32 [rows, ldx] -> the two-plane fp16 form, x * 2^col_exp[c] = hi + lo: out fp16 [rows, ld], hi in columns [0, cols), lo in
 * [plane_off, plane_off + cols).  For callers that hold qkv in fp32 (drop-in modules); sets QVIT_FLAG_OVERFLOW on |x| >= 65504. */
int qvit_split2_f16(const float* x, int64_t rows, int cols, int64_t ldx, const int* col_exp, void* out, int64_t ld,
                    int plane_off, int32_t* flags, qvit_stream_t stream);

/* Attention core of the QAT step (ViTAttention.forward under autograd, VIT:133-149; config 3).  qkv: fp32 [B, T, 3, H, 64] - the
 * output of the qkv layer as it lies in memory; out / dout: fp32 [B, T, H, 64]; lse (saved by the forward) and dstat (workspace):
 * fp32 [B, H, 256]; planes: workspace, fp16 [B * T, 2 * 3 * H * 64]; dqkv: fp32 like qkv, every element written.
 * head_dim must be 64 and T <= 208 (QVIT_ERR_UNSUPPORTED otherwise).  Forward: two fp16 planes, 22 significant bits; backward:
 * two bf16 planes, 16 significant bits (fp32 range), both with fp32 accumulation on tcgen05.                                  */
int qvit_attention_train_fwd(const float* qkv, int B, int T, int H, int head_dim, float scale, void* planes, float* out,
                             float* lse, qvit_stream_t stream);
int qvit_attention_train_bwd(const float* qkv, const float* out, const float* dout, const float* lse, int B, int T, int H,
                             int head_dim, float scale, float* dstat, float* dqkv, qvit_stream_t stream);
/* same; additionally records cycle stamps of the first CTA's first 8 tiles into prof (int64 [256]; NULL = off).  Diagnostic. */
int qvit_attention_train_bwd_prof(const float* qkv, const float* out, const float* dout, const float* lse, int B, int T, int H,
                                  int head_dim, float scale, float* dstat, float* dqkv, long long* prof, qvit_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* QVIT_B200_H_ */
