// DoReFa-style fixed-grid path of UltraNet ("4-bit quantization/quant_ultra.py", quantization.py, mymodel.py):
//   K1'  weight quantizer: tanh -> /max|tanh| -> round(. * (2^(b-1)-1))           QU:38-56
//   K2'  activation quantizer: clamp[0,1] -> round(. * (2^a-1))                    QU:66-73
//   K5   BN fold (both formulas) + integer (inc, bias) thresholds                  MM:74.. / QZ:34-46, QZ:68-89
//   Conv2d_Q with a raw fp32 input (first layer sees the image)                    QU:85-89
//   one fused integer layer  conv -> BatchNorm2d(eval) -> act-quant [-> 2x2 pool]  MM:71-125
// The elementwise kernels are HBM-bound (4 B in, 1 B out per element); UltraNet at batch 1 is
// launch/latency-bound (0.4 GOP, 105 KB of weights), so the fused layer kernel is a CUDA-core dp4a
// direct convolution with the whole epilogue fused, meant to be replayed from a CUDA graph.
#include "common.cuh"

namespace qvit {

constexpr int kUT = 256;

static inline int ultra_grid(int64_t items, int per_block) {
  int64_t b = (items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------- weight quantizer
__global__ void __launch_bounds__(kUT)
tanh_absmax_kernel(const float* __restrict__ w, int64_t n, uint32_t* __restrict__ out_bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(tanhf(w[i])));
  m = warp_max(m);
  __shared__ float sm[kUT / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < kUT / 32 ? sm[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) atomicMax(out_bits, __float_as_uint(m));
  }
}
__global__ void zero_u32_kernel(uint32_t* p) { *p = 0u; }

__global__ void __launch_bounds__(kUT)
ultra_quantize_weight_kernel(const float* __restrict__ w, int64_t n, int w_bit, int export_rounding,
                             const float* __restrict__ max_tanh, int8_t* __restrict__ codes) {
  const float mx = __ldg(max_tanh);
  const float levels = (float)((1 << (w_bit - 1)) - 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __fdiv_rn(tanhf(w[i]), mx);            // QU:50-53
    float c;
    if (w_bit == 2 && !export_rounding) c = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);   // uniform_quantize(k=1) = sign  (QU:15-16)
    else c = rintf(v * levels);                            // QU:18
    if (!(c == c)) c = 0.f;
    codes[i] = (int8_t)(int)c;
  }
}

// ---------------------------------------------------------------- activation quantizer
__global__ void __launch_bounds__(kUT)
ultra_quantize_act_kernel(const float* __restrict__ x, int64_t n, int a_bit, uint8_t* __restrict__ codes,
                          float* __restrict__ values) {
  const float levels = (float)((1 << a_bit) - 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    const float cl = fminf(fmaxf(xv, 0.0f), 1.0f);         // torch.clamp(x, 0, 1); NaN handled below
    float c = rintf(cl * levels);                          // QU:18
    if (xv != xv) c = xv;
    if (values) values[i] = __fdiv_rn(c, levels);          // QU:19
    if (codes) codes[i] = (xv != xv) ? (uint8_t)0 : (uint8_t)(int)c;
  }
}

// ---------------------------------------------------------------- Conv2d_Q on an fp32 input (groups == 1)
struct ConvF32Geom {
  int B, C, H, W, O, kh, kw, sh, sw, ph, pw, dh, dw, OH, OW;
};

// grid = (pixel blocks, O): one CTA = one output channel x 256 output pixels.  The channel's fake-quant weights
// w_q = code / levels (bit-for-bit the reference's round(v*n)/n, QU:18-19) are expanded to fp32 in shared memory
// once, so the inner loop is one broadcast LDS + one coalesced LDG + one FFMA per tap.
__global__ void __launch_bounds__(kUT)
conv2d_f32_wcodes_kernel(const float* __restrict__ x, const int8_t* __restrict__ wc, ConvF32Geom g, float w_levels,
                         const float* __restrict__ bias, float* __restrict__ y) {
  extern __shared__ float s_wq[];
  const int o = blockIdx.y;
  const int taps = g.C * g.kh * g.kw;
  for (int i = threadIdx.x; i < taps; i += blockDim.x)
    s_wq[i] = __fdiv_rn((float)__ldg(wc + (int64_t)o * taps + i), w_levels);
  __syncthreads();
  const int64_t npix = (int64_t)g.B * g.OH * g.OW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const int ow = (int)(i % g.OW);
    const int oh = (int)((i / g.OW) % g.OH);
    const int b = (int)(i / ((int64_t)g.OW * g.OH));
    float acc = 0.f;
    for (int c = 0; c < g.C; ++c) {
      const float* xc = x + ((int64_t)b * g.C + c) * g.H * g.W;
      const float* wq = s_wq + c * g.kh * g.kw;
      for (int ki = 0; ki < g.kh; ++ki) {
        const int ih = oh * g.sh - g.ph + ki * g.dh;
        if (ih < 0 || ih >= g.H) continue;
        for (int kj = 0; kj < g.kw; ++kj) {
          const int iw = ow * g.sw - g.pw + kj * g.dw;
          if (iw < 0 || iw >= g.W) continue;
          acc = fmaf(__ldg(xc + (int64_t)ih * g.W + iw), wq[ki * g.kw + kj], acc);
        }
      }
    }
    if (bias) acc += __ldg(bias + o);
    y[(((int64_t)b * g.O + o) * g.OH + oh) * g.OW + ow] = acc;
  }
}

// ---------------------------------------------------------------- uniform_quantize(k) forward (QU:12-20)
__global__ void __launch_bounds__(kUT)
uniform_quantize_kernel(const float* __restrict__ x, int64_t n, int k, float* __restrict__ out) {
  const float levels = (k >= 1 && k < 32) ? (float)((1u << k) - 1u) : 0.0f;      // k = 0 -> n = 0 (NaN, as upstream)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    float o;
    if (k == 32) o = v;
    else if (k == 1) o = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : ((v == 0.f) ? 0.f : v));
    else o = __fdiv_rn(rintf(v * levels), levels);
    out[i] = o;
  }
}

// ---------------------------------------------------------------- BN(eval) + act-quant (+ 2x2 max-pool): NCHW fp32 -> NHWC codes
// Used after the fp32 first-layer conv (MM:73-76) and, with scale = bias = NULL, to turn an image into 8-bit codes.
__global__ void __launch_bounds__(kUT)
bn_act_pool_nchw_kernel(const float* __restrict__ x, int B, int C, int H, int W, const float* __restrict__ scale,
                        const float* __restrict__ bias, int levels, int pool, uint8_t* __restrict__ out, int ldc) {
  const int OH = pool ? H / 2 : H, OW = pool ? W / 2 : W;
  const int64_t total = (int64_t)B * C * OH * OW;
  const float lv = (float)levels;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ow = (int)(i % OW);
    const int oh = (int)((i / OW) % OH);
    const int c = (int)((i / ((int64_t)OW * OH)) % C);
    const int b = (int)(i / ((int64_t)OW * OH * C));
    const float s = scale ? __ldg(scale + c) : 1.0f, bb = bias ? __ldg(bias + c) : 0.0f;
    const float* src = x + ((int64_t)b * C + c) * H * W;
    int best = 0;
    const int reps = pool ? 2 : 1;
    for (int dy = 0; dy < reps; ++dy)
      for (int dx = 0; dx < reps; ++dx) {
        const float v = src[(int64_t)(oh * reps + dy) * W + (ow * reps + dx)];
        const float yv = scale ? (v * s + bb) : v;
        const int code = (int)rintf(fminf(fmaxf(yv, 0.0f), 1.0f) * lv);
        best = max(best, code);
      }
    out[(((int64_t)b * OH + oh) * OW + ow) * ldc + c] = (uint8_t)best;
  }
}

// ---------------------------------------------------------------- fused integer layer
// CTA = 256 threads = (pixels in a TH x 16 tile) x (channel groups of 16 output channels).
// Shared memory: input halo tile [(TH+kh-1) x (16+kw-1)] pixels x Cw words (+1 word pad per pixel: no bank
// conflicts for the per-thread pixel stride), weights [O][kh*kw*Cw] words, output code tile for pooling.
constexpr int kTW = 16;        // tile width in pixels

// kOPT = output channels per thread: 16 for large maps, 4 for the small late layers (4x more threads per pixel, so a
// 10 x 20 map still spreads over >= 13 CTAs instead of 4)
template <int kOPT>
__global__ void __launch_bounds__(kUT)
ultra_conv_bn_act_kernel(const uint8_t* __restrict__ in, int B, int H, int W, int C, const int8_t* __restrict__ wc, int O,
                         int kh, int kw, int pad, float acc_scale, const float* __restrict__ bn_scale,
                         const float* __restrict__ bn_bias, int out_levels, int pool, uint8_t* __restrict__ out_codes,
                         float* __restrict__ out_f32, int tiles_x, int tiles_y, int TH) {
  extern __shared__ uint32_t smem[];
  const int Cw = (C + 3) >> 2;                    // input words per pixel
  const int pix_stride = Cw + 1;
  const int in_w = kTW + kw - 1, in_h = TH + kh - 1;
  const int KW = kh * kw * Cw;                    // words per output channel
  const int groups = (O + kOPT - 1) / kOPT;
  uint32_t* s_in = smem;                                    // in_h*in_w*pix_stride (rounded up to 16 bytes)
  uint32_t* s_w = s_in + ((in_h * in_w * pix_stride + 3) & ~3);   // O*KW
  uint8_t* s_out = reinterpret_cast<uint8_t*>(s_w + O * KW);   // TH*kTW*O  (pool only)

  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int ty0 = ((tile / tiles_x) % tiles_y) * TH;
  const int tx0 = (tile % tiles_x) * kTW;

  // ---- stage weights: global [O][kh][kw][C] int8 -> smem words (zero padded to 4 channels)
  if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(wc) & 15) == 0 && ((O * KW) & 3) == 0) {
    // channels already fill whole words: the global layout IS the shared layout -> straight 128-bit copies
    const uint4* src = reinterpret_cast<const uint4*>(wc);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < (O * KW) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  } else
  for (int i = threadIdx.x; i < O * KW; i += blockDim.x) {
    const int o = i / KW, r = i - o * KW;
    const int tap = r / Cw, cw = r - tap * Cw;
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cw * 4 + j;
      if (c < C) v |= (uint32_t)(uint8_t)__ldg(wc + ((int64_t)o * kh * kw + tap) * C + c) << (8 * j);
    }
    s_w[i] = v;
  }
  // ---- stage the input halo tile (zero padding -> code 0)
  for (int i = threadIdx.x; i < in_h * in_w * Cw; i += blockDim.x) {
    const int p = i / Cw, cw = i - p * Cw;
    const int iy = ty0 - pad + p / in_w, ix = tx0 - pad + p % in_w;
    uint32_t v = 0;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const uint8_t* src = in + (((int64_t)b * H + iy) * W + ix) * C + cw * 4;
      if ((C & 3) == 0) v = __ldg(reinterpret_cast<const uint32_t*>(src));
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (cw * 4 + j < C) v |= (uint32_t)__ldg(src + j) << (8 * j);
      }
    }
    s_in[p * pix_stride + cw] = v;
  }
  __syncthreads();

  const int npix = TH * kTW;
  const int pix = threadIdx.x % npix, grp = threadIdx.x / npix;
  const int py = pix / kTW, px = pix % kTW;
  const int oy = ty0 + py, ox = tx0 + px;
  const bool active = (grp < groups);
  int acc[kOPT];
#pragma unroll
  for (int j = 0; j < kOPT; ++j) acc[j] = 0;
  if (active) {
    const int o0 = grp * kOPT;
    for (int ki = 0; ki < kh; ++ki) {
      for (int kj = 0; kj < kw; ++kj) {
        const uint32_t* ip = s_in + ((py + ki) * in_w + (px + kj)) * pix_stride;
        const uint32_t* wp = s_w + (ki * kw + kj) * Cw;
        for (int cw = 0; cw < Cw; ++cw) {
          const uint32_t a = ip[cw];
#pragma unroll
          for (int j = 0; j < kOPT; ++j) {
            if (o0 + j < O) {
              int d;
              const uint32_t wv = wp[(o0 + j) * KW + cw];
              asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(wv), "r"(acc[j]));
              acc[j] = d;
            }
          }
        }
      }
    }
  }
  const int OHf = H + 2 * pad - kh + 1, OWf = W + 2 * pad - kw + 1;   // stride 1
  const bool in_img = active && oy < OHf && ox < OWf;
  const float lv = (float)out_levels;
  if (out_f32) {
    if (in_img) {
#pragma unroll
      for (int j = 0; j < kOPT; ++j) {
        const int o = grp * kOPT + j;
        if (o < O) {
          float yv = (float)acc[j] * acc_scale;
          if (bn_scale) yv *= __ldg(bn_scale + o);
          if (bn_bias) yv += __ldg(bn_bias + o);
          out_f32[(((int64_t)b * O + o) * OHf + oy) * OWf + ox] = yv;
        }
      }
    }
    return;
  }
  uint32_t packed[kOPT / 4];
#pragma unroll
  for (int q = 0; q < kOPT / 4; ++q) packed[q] = 0;
#pragma unroll
  for (int j = 0; j < kOPT; ++j) {
    const int o = grp * kOPT + j;
    float yv = (float)acc[j] * acc_scale;
    if (active && o < O) {
      if (bn_scale) yv *= __ldg(bn_scale + o);
      if (bn_bias) yv += __ldg(bn_bias + o);
    }
    const float cl = fminf(fmaxf(yv, 0.0f), 1.0f);
    const int code = (int)rintf(cl * lv);
    packed[j >> 2] |= (uint32_t)(code & 0xff) << (8 * (j & 3));
  }
  if (!pool) {
    if (in_img) {
      uint8_t* dst = out_codes + (((int64_t)b * OHf + oy) * OWf + ox) * O + grp * kOPT;
      if ((O % kOPT) == 0 && kOPT == 16) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[kOPT / 4 - 1]);
      } else if ((O % kOPT) == 0 && kOPT == 4) {
        *reinterpret_cast<uint32_t*>(dst) = packed[0];
      } else {
#pragma unroll
        for (int j = 0; j < kOPT; ++j)
          if (grp * kOPT + j < O) dst[j] = (uint8_t)(packed[j >> 2] >> (8 * (j & 3)));
      }
    }
    return;
  }
  // 2x2 max-pool on codes (monotone code map => pool(codes) == codes(pool), MM:76)
  if (active) {
    uint32_t* so = reinterpret_cast<uint32_t*>(s_out + ((size_t)pix * groups + grp) * kOPT);
#pragma unroll
    for (int q = 0; q < kOPT / 4; ++q) so[q] = in_img ? packed[q] : 0u;
  }
  __syncthreads();
  const int PH = OHf / 2, PW = OWf / 2;
  const int ppix = (TH / 2) * (kTW / 2);
  const int Ow = groups * kOPT;                   // padded channel count in s_out
  for (int i = threadIdx.x; i < ppix * (Ow / 4); i += blockDim.x) {
    const int pp = i / (Ow / 4), q = i - pp * (Ow / 4);
    const int qy = pp / (kTW / 2), qx = pp % (kTW / 2);
    const int gy = ty0 / 2 + qy, gx = tx0 / 2 + qx;
    if (gy >= PH || gx >= PW) continue;
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s_out);
    const int p00 = (2 * qy) * kTW + 2 * qx;
    const uint32_t v0 = s32[(p00)*(Ow / 4) + q], v1 = s32[(p00 + 1) * (Ow / 4) + q];
    const uint32_t v2 = s32[(p00 + kTW) * (Ow / 4) + q], v3 = s32[(p00 + kTW + 1) * (Ow / 4) + q];
    const uint32_t m = __vmaxu4(__vmaxu4(v0, v1), __vmaxu4(v2, v3));
    uint8_t* dst = out_codes + (((int64_t)b * PH + gy) * PW + gx) * O + q * 4;
    if (q * 4 + 4 <= O && (O & 3) == 0) *reinterpret_cast<uint32_t*>(dst) = m;
    else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (q * 4 + j < O) dst[j] = (uint8_t)(m >> (8 * j));
    }
  }
}

// ---------------------------------------------------------------- BN fold + integer thresholds
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                               int mode, int C, float* scale, float* bias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (mode == 0) {   // nn.BatchNorm2d eval: (v - mean) / sqrt(var + eps) * gamma + beta
    const float s = __fdiv_rn(gamma[c], sqrtf(var[c] + eps));
    scale[c] = s;
    bias[c] = beta[c] - mean[c] * s;
  } else {           // QZ:43-45: eps OUTSIDE the sqrt
    const float den = sqrtf(var[c]) + eps;
    scale[c] = __fdiv_rn(gamma[c], den);
    bias[c] = beta[c] - __fdiv_rn(mean[c], den) * gamma[c];
  }
}

template <typename T>
__global__ void bn_act_quantize_int_kernel(const T* gamma, const T* beta, const T* mean, const T* var, double eps,
                                           int w_bit, int in_bit, int out_bit, int l_shift, int C, int32_t* inc,
                                           int32_t* bias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // QZ:34-46 in the array dtype (NumPy keeps float32 arrays float32 against Python scalars)
  const T den = (T)sqrt((double)var[c]) + (T)eps;
  const T w = gamma[c] / den;
  const T bb = beta[c] - (mean[c] / den * gamma[c]);
  // QZ:76-86
  const double wl = (double)((1 << (w_bit - 1)) - 1), il = (double)((1 << in_bit) - 1), ol = (double)((1 << out_bit) - 1);
  const double n = ldexp(1.0, w_bit - 1 + in_bit + l_shift) / (wl * il);
  const T inc_f = (T)(ol * n) * w;
  const T bias_f = (T)(wl * il * ol * n) * bb;
  inc[c] = (int32_t)rint((double)inc_f);
  bias[c] = (int32_t)rint((double)bias_f);
}

}  // namespace qvit

using namespace qvit;

extern "C" {

int qvit_ultra_tanh_absmax(const float* w, int64_t n, float* out, qvit_stream_t stream) {
  QVIT_REQUIRE(w && out && n >= 0, "qvit_ultra_tanh_absmax: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  zero_u32_kernel<<<1, 1, 0, s>>>(reinterpret_cast<uint32_t*>(out));
  if (n > 0) tanh_absmax_kernel<<<ultra_grid(n, kUT * 4), kUT, 0, s>>>(w, n, reinterpret_cast<uint32_t*>(out));
  return check_launch("qvit_ultra_tanh_absmax");
}

int qvit_ultra_quantize_weight(const float* w, int64_t n, int w_bit, int export_rounding, const float* max_tanh,
                               int8_t* codes, qvit_stream_t stream) {
  QVIT_REQUIRE(w && max_tanh && codes && n >= 0, "qvit_ultra_quantize_weight: bad argument");
  QVIT_REQUIRE(w_bit >= 2 && w_bit <= 8, "qvit_ultra_quantize_weight: w_bit must be in [2, 8] (got %d)", w_bit);
  if (n == 0) return QVIT_OK;
  ultra_quantize_weight_kernel<<<ultra_grid(n, kUT * 4), kUT, 0, (cudaStream_t)stream>>>(w, n, w_bit, export_rounding, max_tanh,
                                                                                         codes);
  return check_launch("qvit_ultra_quantize_weight");
}

int qvit_ultra_quantize_act(const float* x, int64_t n, int a_bit, uint8_t* codes, float* out_values,
                            qvit_stream_t stream) {
  QVIT_REQUIRE(x && n >= 0 && (codes || out_values), "qvit_ultra_quantize_act: bad argument");
  QVIT_REQUIRE(a_bit >= 1 && a_bit <= 8, "qvit_ultra_quantize_act: a_bit must be in [1, 8] (got %d)", a_bit);
  if (n == 0) return QVIT_OK;
  ultra_quantize_act_kernel<<<ultra_grid(n, kUT * 4), kUT, 0, (cudaStream_t)stream>>>(x, n, a_bit, codes, out_values);
  return check_launch("qvit_ultra_quantize_act");
}

int qvit_conv2d_f32_wcodes(const float* x, int B, int C, int H, int W, const int8_t* w_codes, int O, int kh, int kw,
                           int sh, int sw, int ph, int pw, int dh, int dw, float w_levels, const float* bias, float* y,
                           qvit_stream_t stream) {
  QVIT_REQUIRE(x && w_codes && y, "qvit_conv2d_f32_wcodes: null pointer");
  QVIT_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && O > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 &&
                   ph >= 0 && pw >= 0 && w_levels > 0.f, "qvit_conv2d_f32_wcodes: bad geometry");
  ConvF32Geom g{B, C, H, W, O, kh, kw, sh, sw, ph, pw, dh, dw, 0, 0};
  g.OH = (H + 2 * ph - dh * (kh - 1) - 1) / sh + 1;
  g.OW = (W + 2 * pw - dw * (kw - 1) - 1) / sw + 1;
  QVIT_REQUIRE(g.OH > 0 && g.OW > 0, "qvit_conv2d_f32_wcodes: empty output");
  const int64_t npix = (int64_t)B * g.OH * g.OW;
  const size_t smem = sizeof(float) * (size_t)C * kh * kw;
  QVIT_REQUIRE(smem <= 48 * 1024 && O <= 65535, "qvit_conv2d_f32_wcodes: C*kh*kw <= 12288 and O <= 65535");
  int gx = (int)((npix + kUT - 1) / kUT);
  const int cap = (sm_count() * 8 + O - 1) / O;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  conv2d_f32_wcodes_kernel<<<dim3((unsigned)gx, (unsigned)O), kUT, smem, (cudaStream_t)stream>>>(x, w_codes, g, w_levels,
                                                                                                bias, y);
  return check_launch("qvit_conv2d_f32_wcodes");
}

int qvit_ultra_conv_bn_act(const uint8_t* in_codes, int B, int H, int W, int C, const int8_t* w_codes, int O, int kh,
                           int kw, int pad, float acc_scale, const float* bn_scale, const float* bn_bias, int out_levels,
                           int pool, uint8_t* out_codes, float* out_f32, qvit_stream_t stream) {
  QVIT_REQUIRE(in_codes && w_codes && (out_codes || out_f32), "qvit_ultra_conv_bn_act: null pointer");
  QVIT_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && O > 0 && kh > 0 && kw > 0 && pad >= 0, "qvit_ultra_conv_bn_act: bad geometry");
  QVIT_REQUIRE(O <= 256, "qvit_ultra_conv_bn_act: O <= 256");
  const int OH = H + 2 * pad - kh + 1, OW = W + 2 * pad - kw + 1;
  QVIT_REQUIRE(OH > 0 && OW > 0, "qvit_ultra_conv_bn_act: empty output");
  QVIT_REQUIRE(!pool || (out_codes && !out_f32), "qvit_ultra_conv_bn_act: pooling applies to the code output only");
  QVIT_REQUIRE(out_f32 || (out_levels >= 1 && out_levels <= 255), "qvit_ultra_conv_bn_act: out_levels in [1,255]");
  // channels per thread: 16 when that already gives >= 2 CTAs per SM, else 4 (small late layers, batch-1 latency)
  int opt = 16;
  {
    const int g16 = (O + 15) / 16;
    int th16 = kUT / (g16 * kTW);
    if (th16 < 1) th16 = 1;
    const int64_t ctas16 = (int64_t)B * ((OW + kTW - 1) / kTW) * ((OH + th16 - 1) / th16);
    if (ctas16 < 2 * sm_count() && O <= 64 && !(pool && kUT / (((O + 3) / 4) * kTW) < 2)) opt = 4;
  }
  const int groups = (O + opt - 1) / opt;
  QVIT_REQUIRE(groups <= 16, "qvit_ultra_conv_bn_act: O too large");
  int TH = kUT / (groups * kTW);                  // 16, 8, 4, 2 or 1 rows of 16 pixels
  if (TH < 1) TH = 1;
  if (pool) QVIT_REQUIRE(TH >= 2 && (TH % 2) == 0, "qvit_ultra_conv_bn_act: pooling needs O <= 128");
  const int Cw = (C + 3) / 4;
  const size_t in_words = ((size_t)(TH + kh - 1) * (kTW + kw - 1) * (Cw + 1) + 3) & ~(size_t)3;
  const size_t smem = sizeof(uint32_t) * (in_words + (size_t)O * kh * kw * Cw) + (pool ? (size_t)TH * kTW * groups * opt : 0);
  QVIT_REQUIRE(smem <= 200 * 1024, "qvit_ultra_conv_bn_act: layer too large for the fused kernel (%zu B smem)", smem);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(ultra_conv_bn_act_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(ultra_conv_bn_act_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  const int tiles_x = (OW + kTW - 1) / kTW, tiles_y = (OH + TH - 1) / TH;
  const int64_t grid = (int64_t)B * tiles_x * tiles_y;
  QVIT_REQUIRE(grid < (1ll << 31), "qvit_ultra_conv_bn_act: grid too large");
  if (opt == 16)
    ultra_conv_bn_act_kernel<16><<<(unsigned)grid, kUT, smem, (cudaStream_t)stream>>>(
        in_codes, B, H, W, C, w_codes, O, kh, kw, pad, acc_scale, bn_scale, bn_bias, out_levels, pool, out_codes, out_f32,
        tiles_x, tiles_y, TH);
  else
    ultra_conv_bn_act_kernel<4><<<(unsigned)grid, kUT, smem, (cudaStream_t)stream>>>(
        in_codes, B, H, W, C, w_codes, O, kh, kw, pad, acc_scale, bn_scale, bn_bias, out_levels, pool, out_codes, out_f32,
        tiles_x, tiles_y, TH);
  return check_launch("qvit_ultra_conv_bn_act");
}

int qvit_uniform_quantize(const float* x, int64_t n, int k, float* out, qvit_stream_t stream) {
  QVIT_REQUIRE(x && out && n >= 0 && k >= 0 && k <= 32, "qvit_uniform_quantize: bad argument");
  if (n == 0) return QVIT_OK;
  uniform_quantize_kernel<<<ultra_grid(n, kUT * 4), kUT, 0, (cudaStream_t)stream>>>(x, n, k, out);
  return check_launch("qvit_uniform_quantize");
}

int qvit_ultra_bn_act_pool_nchw(const float* x, int B, int C, int H, int W, const float* scale, const float* bias,
                                int levels, int pool, uint8_t* out_codes, int ldc, qvit_stream_t stream) {
  QVIT_REQUIRE(x && out_codes && B > 0 && C > 0 && H > 0 && W > 0 && ldc >= C, "qvit_ultra_bn_act_pool_nchw: bad argument");
  QVIT_REQUIRE(levels >= 1 && levels <= 255, "qvit_ultra_bn_act_pool_nchw: levels in [1,255]");
  QVIT_REQUIRE((scale == nullptr) == (bias == nullptr), "qvit_ultra_bn_act_pool_nchw: scale and bias go together");
  const int64_t total = (int64_t)B * C * (pool ? H / 2 : H) * (pool ? W / 2 : W);
  if (total == 0) return QVIT_OK;
  bn_act_pool_nchw_kernel<<<ultra_grid(total, kUT * 2), kUT, 0, (cudaStream_t)stream>>>(x, B, C, H, W, scale, bias, levels,
                                                                                       pool, out_codes, ldc);
  return check_launch("qvit_ultra_bn_act_pool_nchw");
}

int qvit_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int mode, int C,
                 float* scale, float* bias, qvit_stream_t stream) {
  QVIT_REQUIRE(gamma && beta && mean && var && scale && bias && C >= 0, "qvit_bn_fold: bad argument");
  QVIT_REQUIRE(mode == 0 || mode == 1, "qvit_bn_fold: mode must be 0 (BatchNorm2d eval) or 1 (export fold)");
  if (C == 0) return QVIT_OK;
  bn_fold_kernel<<<div_up(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, mode, C, scale, bias);
  return check_launch("qvit_bn_fold");
}

int qvit_bn_act_quantize_int(const void* gamma, const void* beta, const void* mean, const void* var, int is_f64,
                             double eps, int w_bit, int in_bit, int out_bit, int l_shift, int C, int32_t* inc,
                             int32_t* bias, qvit_stream_t stream) {
  QVIT_REQUIRE(gamma && beta && mean && var && inc && bias && C >= 0, "qvit_bn_act_quantize_int: bad argument");
  QVIT_REQUIRE(w_bit >= 2 && w_bit <= 8 && in_bit >= 1 && in_bit <= 8 && out_bit >= 1 && out_bit <= 8 && l_shift >= 0 &&
                   l_shift <= 16, "qvit_bn_act_quantize_int: bad bit widths");
  if (C == 0) return QVIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (is_f64)
    bn_act_quantize_int_kernel<double><<<div_up(C, 128), 128, 0, s>>>(
        (const double*)gamma, (const double*)beta, (const double*)mean, (const double*)var, eps, w_bit, in_bit, out_bit,
        l_shift, C, inc, bias);
  else
    bn_act_quantize_int_kernel<float><<<div_up(C, 128), 128, 0, s>>>((const float*)gamma, (const float*)beta,
                                                                     (const float*)mean, (const float*)var, eps, w_bit,
                                                                     in_bit, out_bit, l_shift, C, inc, bias);
  return check_launch("qvit_bn_act_quantize_int");
}

}  // extern "C"
