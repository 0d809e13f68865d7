// K1/K2: GETA symmetric quantizers -> int8 codes / fp32 fake-quant values; im2col+quantize;
// LayerNorm+quantize; bf16 quantize; absmax; int4 pack/unpack.
//
// All of these are HBM-bound elementwise/row kernels (roofline: 4 B read + 1 B written per element
// for fp32 -> int8).  Layout rule: one warp owns 512 consecutive elements per step; lane l loads the
// float4 at element 128*j + 4*l (j = 0..3) so every load instruction is a fully coalesced 512 B
// request, and stores one packed 4-code word per j (128 B coalesced per store instruction).
// Grids are sized as multiples of the SM count (148 on B200) with grid-stride loops.
#include "common.cuh"

namespace qvit {

constexpr int kThreads = 256;
constexpr int kWarpElems = 512;                        // elements one warp converts per step
constexpr int kBlockElems = kWarpElems * (kThreads / 32);

static inline int stream_grid(int64_t work_items, int per_block, int ctas_per_sm = 8) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------------------------------------
// flat fp32 -> int8 codes, n % 512 handled by a scalar tail
// ------------------------------------------------------------------------------------------------
// GELU: the codes of quantize_act(gelu(x)) - Mlp.forward's nn.GELU (vit_model.py:173) fused into fc2's activation quantizer
// for the QAT step (the fp32 activation is never written; the backward recomputes it, backward.cu)
template <bool GELU>
__global__ void __launch_bounds__(kThreads)
quantize_sym_flat_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ d,
                         const float* __restrict__ qm, const float* __restrict__ t,
                         int8_t* __restrict__ codes, int32_t* __restrict__ flags) {
  const SymParams p = load_sym_params(d, qm, t);
  const FastQ2 fq = make_fastq2(p);
  int fl = 0;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t full = n / kWarpElems;
  for (int64_t w = warp_global; w < full; w += warp_stride) {
    const float* src = x + w * kWarpElems;
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ldg_stream4(src + j * 128 + lane * 4);
    if (GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = make_float4(gelu_erf(v[j].x), gelu_erf(v[j].y), gelu_erf(v[j].z), gelu_erf(v[j].w));
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(codes + w * kWarpElems);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dst[j * 32 + lane] = sym_codes4_v2(v[j].x, v[j].y, v[j].z, v[j].w, p, fq, fl);
    }
  }
  // tail (< 512 elements): first warp of block 0
  if (warp_global == 0) {
    for (int64_t i = full * kWarpElems + lane; i < n; i += 32) codes[i] = (int8_t)sym_code(GELU ? gelu_erf(x[i]) : x[i], p, fl);
  }
  fl = warp_or(fl);
  if (fl && flags && lane == 0) atomicOr(flags, fl);
}

// general 2-D form with row pitches and zero padding of columns cols..ld_codes-1
__global__ void __launch_bounds__(kThreads)
quantize_sym_rows_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld_x,
                         const float* __restrict__ d, const float* __restrict__ qm, const float* __restrict__ t,
                         int8_t* __restrict__ codes, int64_t ld_codes, int32_t* __restrict__ flags) {
  const SymParams p = load_sym_params(d, qm, t);
  int fl = 0;
  const int64_t total = rows * ld_codes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_codes, c = i - r * ld_codes;
    codes[i] = (c < cols) ? (int8_t)sym_code(x[r * ld_x + c], p, fl) : (int8_t)0;
  }
  fl = warp_or(fl);
  if (fl && flags && (threadIdx.x & 31) == 0) atomicOr(flags, fl);
}

__global__ void __launch_bounds__(kThreads)
fake_quantize_sym_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ d,
                         const float* __restrict__ qm, const float* __restrict__ t, float* __restrict__ out) {
  const SymParams p = load_sym_params(d, qm, t);
  const int64_t n4 = n >> 2;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (aligned) {
    for (int64_t i = tid; i < n4; i += stride) {
      float4 v = ldg_stream4(x + i * 4);
      float4 o = make_float4(sym_value(v.x, p), sym_value(v.y, p), sym_value(v.z, p), sym_value(v.w, p));
      reinterpret_cast<float4*>(out)[i] = o;
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += stride) out[i] = sym_value(x[i], p);
  } else {
    for (int64_t i = tid; i < n; i += stride) out[i] = sym_value(x[i], p);
  }
}

// ------------------------------------------------------------------------------------------------
// bf16 -> codes (flat or rows)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
quantize_sym_bf16_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int64_t cols, int64_t ld_x,
                              const float* __restrict__ d, const float* __restrict__ qm, const float* __restrict__ t,
                              int8_t* __restrict__ codes, int64_t ld_codes, int32_t* __restrict__ flags) {
  const SymParams p = load_sym_params(d, qm, t);
  const FastQ2 fq = make_fastq2(p);
  int fl = 0;
  // one thread = 8 consecutive columns (16 B in, 8 B out) when everything is 8-aligned
  const bool vec = (cols % 8 == 0) && (ld_x % 8 == 0) && (ld_codes % 8 == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(codes) & 7) == 0);
  if (vec) {
    const int64_t groups_per_row = ld_codes / 8;
    const int64_t total = rows * groups_per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i / groups_per_row, g = i - r * groups_per_row;
      uint2 o = make_uint2(0u, 0u);
      if (g * 8 < cols) {
        const uint4 raw = *reinterpret_cast<const uint4*>(x + r * ld_x + g * 8);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[2 * j] = __uint_as_float(w[j] << 16);
          f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
        }
        o.x = sym_codes4_v2(f[0], f[1], f[2], f[3], p, fq, fl);
        o.y = sym_codes4_v2(f[4], f[5], f[6], f[7], p, fq, fl);
      }
      *reinterpret_cast<uint2*>(codes + r * ld_codes + g * 8) = o;
    }
  } else {
    const int64_t total = rows * ld_codes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i / ld_codes, c = i - r * ld_codes;
      codes[i] = (c < cols) ? (int8_t)sym_code(__bfloat162float(x[r * ld_x + c]), p, fl) : (int8_t)0;
    }
  }
  fl = warp_or(fl);
  if (fl && flags && (threadIdx.x & 31) == 0) atomicOr(flags, fl);
}

// ------------------------------------------------------------------------------------------------
// im2col + quantize (NCHW fp32 -> [B*OH*OW, ld_cols] int8, K ordered (c, kh, kw))
// ------------------------------------------------------------------------------------------------
struct ConvGeom {
  int B, C, H, W, kh, kw, sh, sw, ph, pw, dh, dw, OH, OW, K;
};

__global__ void __launch_bounds__(kThreads)
im2col_quantize_kernel(const float* __restrict__ x, ConvGeom g, const float* __restrict__ d,
                       const float* __restrict__ qm, const float* __restrict__ t,
                       int8_t* __restrict__ cols, int64_t ld_cols, int32_t* __restrict__ flags) {
  const SymParams p = load_sym_params(d, qm, t);
  int fl = 0;
  const int64_t words_per_row = ld_cols / 4;            // ld_cols % 4 == 0 enforced by the host
  const int64_t rows = (int64_t)g.B * g.OH * g.OW;
  const int64_t total = rows * words_per_row;
  const int khw = g.kh * g.kw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / words_per_row;
    const int k0 = (int)(i - r * words_per_row) * 4;
    const int ow = (int)(r % g.OW);
    const int oh = (int)((r / g.OW) % g.OH);
    const int b = (int)(r / ((int64_t)g.OW * g.OH));
    int c4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j;
      int code = 0;
      if (k < g.K) {
        const int c = k / khw;
        const int rem = k - c * khw;
        const int ki = rem / g.kw, kj = rem - ki * g.kw;
        const int ih = oh * g.sh - g.ph + ki * g.dh;
        const int iw = ow * g.sw - g.pw + kj * g.dw;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
          code = sym_code(__ldg(x + (((int64_t)b * g.C + c) * g.H + ih) * g.W + iw), p, fl);
      }
      c4[j] = code;
    }
    reinterpret_cast<uint32_t*>(cols)[i] = pack4_i8(c4[0], c4[1], c4[2], c4[3]);
  }
  fl = warp_or(fl);
  if (fl && flags && (threadIdx.x & 31) == 0) atomicOr(flags, fl);
}

// Non-overlapping patches (kernel == stride, no padding, no dilation, kw % 4 == 0, W % 4 == 0): ViT PatchEmbed
// (vit_model.py:94-100).  One WARP = one (image, channel, patch): its kh x kw pixels are kh runs of kw * 4 contiguous bytes in
// the image and ONE run of kh * kw contiguous bytes in the patch matrix (K index = (c * kh + ki) * kw + kj), so lane j takes the
// j-th group of 4 pixels: 128-bit loads in full 32-byte sectors, 128-byte coalesced stores.  (The first version gave every thread
// 4 pixels of an image row: its 4-byte stores landed 16 bytes per patch row, half a sector at a time - 81 us for the 154 MB batch.)
__global__ void __launch_bounds__(kThreads)
patchify_quantize_kernel(const float* __restrict__ x, ConvGeom g, const float* __restrict__ d, const float* __restrict__ qm,
                         const float* __restrict__ t, int8_t* __restrict__ cols, int64_t ld_cols, int32_t* __restrict__ flags) {
  const SymParams p = load_sym_params(d, qm, t);
  const FastQ2 fq = make_fastq2(p);
  int fl = 0;
  const int lane = threadIdx.x & 31;
  const int kw4 = g.kw / 4;                                // float4 groups per patch row
  const int groups = g.kh * kw4;                           // float4 groups per (channel, patch)
  const int64_t total = (int64_t)g.B * g.C * g.OH * g.OW;  // warps' work items
  const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5), wstride = (int64_t)gridDim.x * (kThreads / 32);
  const uint32_t per_plane = (uint32_t)(g.OH * g.OW);
  for (int64_t w = warp0; w < total; w += wstride) {
    const uint32_t plane = (uint32_t)(w / per_plane), rem = (uint32_t)(w - (int64_t)plane * per_plane);
    const int b = (int)(plane / (uint32_t)g.C), c = (int)(plane - (uint32_t)b * (uint32_t)g.C);
    const int oh = (int)(rem / (uint32_t)g.OW), ow = (int)(rem - (uint32_t)oh * (uint32_t)g.OW);
    const float* src = x + (((int64_t)b * g.C + c) * g.H + (int64_t)oh * g.kh) * g.W + ow * g.kw;
    int8_t* dst = cols + (((int64_t)b * g.OH + oh) * g.OW + ow) * ld_cols + c * g.kh * g.kw;
    for (int j = lane; j < groups; j += 32) {
      const int ki = j / kw4, kj = (j - ki * kw4) * 4;
      const float4 v = ldg_stream4(src + (int64_t)ki * g.W + kj);
      *reinterpret_cast<uint32_t*>(dst + j * 4) = sym_codes4_v2(v.x, v.y, v.z, v.w, p, fq, fl);
    }
  }
  fl = warp_or(fl);
  if (fl && flags && lane == 0) atomicOr(flags, fl);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm + quantize: one warp per row, the row lives in registers (cols = 128 * V, V <= 16)
// ------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kThreads)
layernorm_quantize_kernel(const float* __restrict__ x, int64_t rows, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, const float* __restrict__ d,
                          const float* __restrict__ qm, const float* __restrict__ t,
                          int8_t* __restrict__ codes, int64_t ld_codes, float* __restrict__ ln_out,
                          int32_t* __restrict__ flags) {
  constexpr int cols = V * 128;
  const SymParams p = load_sym_params(d, qm, t);
  const FastQ2 fq = make_fastq2(p);
  int fl = 0;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * (kThreads / 32);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  // software pipeline: the loads of the warp's NEXT row are in flight while the current row is reduced, normalised and
  // stored (two dependent warp reductions per row would otherwise leave the memory system idle in between)
  float4 nxt[V];
  if (warp_global < rows) {
#pragma unroll
    for (int j = 0; j < V; ++j) nxt[j] = ldg_stream4(x + warp_global * cols + j * 128 + lane * 4);
  }
  for (int64_t r = warp_global; r < rows; r += warp_stride) {
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = nxt[j];
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    if (r + warp_stride < rows) {
#pragma unroll
      for (int j = 0; j < V; ++j) nxt[j] = ldg_stream4(x + (r + warp_stride) * cols + j * 128 + lane * 4);
    }
    const float mean = warp_sum(s) * (1.0f / cols);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / cols) + eps);
    uint32_t* dst = reinterpret_cast<uint32_t*>(codes + r * ld_codes);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 gm = __ldg(g4 + j * 32 + lane), bt = __ldg(b4 + j * 32 + lane);   // L1-resident, shared by all rows
      float4 y;
      y.x = (v[j].x - mean) * rstd * gm.x + bt.x;
      y.y = (v[j].y - mean) * rstd * gm.y + bt.y;
      y.z = (v[j].z - mean) * rstd * gm.z + bt.z;
      y.w = (v[j].w - mean) * rstd * gm.w + bt.w;
      if (ln_out) reinterpret_cast<float4*>(ln_out + r * cols)[j * 32 + lane] = y;
      dst[j * 32 + lane] = sym_codes4_v2(y.x, y.y, y.z, y.w, p, fq, fl);
    }
    // zero the K padding, if any
    for (int64_t c = cols + lane; c < ld_codes; c += 32) codes[r * ld_codes + c] = 0;
  }
  fl = warp_or(fl);
  if (fl && flags && lane == 0) atomicOr(flags, fl);
}

// h[b, 0, :] = cls + pos[0]; h[b, 1 + p, :] = tok[b * P + p, :] + pos[1 + p]   (cat(cls_token, x) + pos_embed, vit_model.py:295-305)
__global__ void __launch_bounds__(kThreads)
embed_assemble_kernel(const float* __restrict__ tok, const float* __restrict__ pos, const float* __restrict__ cls, int B, int P, int D4,
                      float* __restrict__ h) {
  const int64_t n4 = (int64_t)B * (P + 1) * D4;
  const int64_t row4 = (int64_t)(P + 1) * D4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / row4, r = i - b * row4;                 // r = token * D4 + column group
    const float4 pe = __ldg(reinterpret_cast<const float4*>(pos) + r);
    float4 v;
    if (r < D4) v = __ldg(reinterpret_cast<const float4*>(cls) + r);
    else v = ldg_stream4(tok + (b * (int64_t)P * D4 + (r - D4)) * 4);
    reinterpret_cast<float4*>(h)[i] = make_float4(v.x + pe.x, v.y + pe.y, v.z + pe.z, v.w + pe.w);
  }
}

// generic width: one warp per row, three passes over the (L1/L2-resident) row
__global__ void __launch_bounds__(kThreads)
layernorm_quantize_generic_kernel(const float* __restrict__ x, int64_t rows, int cols, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float eps, const float* __restrict__ d,
                                  const float* __restrict__ qm, const float* __restrict__ t,
                                  int8_t* __restrict__ codes, int64_t ld_codes, float* __restrict__ ln_out,
                                  int32_t* __restrict__ flags) {
  const SymParams p = load_sym_params(d, qm, t);
  int fl = 0;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * (kThreads / 32);
  for (int64_t r = warp_global; r < rows; r += warp_stride) {
    const float* src = x + r * cols;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += src[c];
    const float mean = warp_sum(s) / (float)cols;
    float q = 0.f;
    for (int c = lane; c < cols; c += 32) { const float a = src[c] - mean; q += a * a; }
    const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
    for (int c = lane; c < ld_codes; c += 32) {
      int code = 0;
      if (c < cols) {
        const float y = (src[c] - mean) * rstd * gamma[c] + beta[c];
        if (ln_out) ln_out[r * cols + c] = y;
        code = sym_code(y, p, fl);
      }
      codes[r * ld_codes + c] = (int8_t)code;
    }
  }
  fl = warp_or(fl);
  if (fl && flags && lane == 0) atomicOr(flags, fl);
}

// ------------------------------------------------------------------------------------------------
// absmax, int4 pack / unpack
// ------------------------------------------------------------------------------------------------
__global__ void zero_word_kernel(uint32_t* p) { *p = 0u; }

__global__ void __launch_bounds__(kThreads)
absmax_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ out_bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  __shared__ float sm[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < kThreads / 32 ? sm[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order like uints
  }
}

__global__ void __launch_bounds__(kThreads)
pack_int4_kernel(const int8_t* __restrict__ codes, int64_t nbytes, uint8_t* __restrict__ packed) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (int64_t)gridDim.x * blockDim.x) {
    const int lo = codes[2 * i], hi = codes[2 * i + 1];
    packed[i] = (uint8_t)((lo & 0xF) | ((hi & 0xF) << 4));
  }
}

// FPGA weight layout of the reference (qnn_mem_process.py:84-130, 152-157): one thread per output word.
// codes [O, I, kh, kw] int8 -> word (oc, j) = runs of `simd` codes of the (kh, kw, I)-ordered row, element e in bits
// [w_bit * e, +w_bit) two's complement; stored at out[(oc % pe) * tiles + (oc / pe) * runs + j].
__global__ void __launch_bounds__(kThreads)
pack_hls_weights_kernel(const int8_t* __restrict__ codes, int O, int I, int kh, int kw, int w_bit, int simd, int pe, int runs,
                        unsigned long long* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)O * runs) return;
  const int oc = (int)(idx / runs), j = (int)(idx - (int64_t)oc * runs);
  const int h = kh * kw * I;
  const unsigned long long mask = (1ull << w_bit) - 1ull;
  unsigned long long word = 0;
  for (int e = 0; e < simd; ++e) {
    const int pos = j * simd + e;                              // position in the (kh, kw, I) order
    if (pos >= h) break;                                       // ragged last run
    const int ic = pos % I, rc = pos / I, r = rc / kw, c = rc - r * kw;
    const int v = codes[(((int64_t)oc * I + ic) * kh + r) * kw + c];
    word |= ((unsigned long long)(long long)v & mask) << (w_bit * e);
  }
  const int64_t tiles = (int64_t)runs * (O / pe);
  out[(int64_t)(oc % pe) * tiles + (int64_t)(oc / pe) * runs + j] = word;
}

__global__ void __launch_bounds__(kThreads)
unpack_int4_kernel(const uint8_t* __restrict__ packed, int64_t nbytes, int is_signed, int8_t* __restrict__ codes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (int64_t)gridDim.x * blockDim.x) {
    int lo = packed[i] & 0xF, hi = (packed[i] >> 4) & 0xF;
    if (is_signed) { lo = (lo ^ 8) - 8; hi = (hi ^ 8) - 8; }
    codes[2 * i] = (int8_t)lo;
    codes[2 * i + 1] = (int8_t)hi;
  }
}

}  // namespace qvit

using namespace qvit;

extern "C" {

int qvit_quantize_sym(const float* x, int64_t rows, int64_t cols, int64_t ld_x, const float* d, const float* q_m,
                      const float* t, int8_t* codes, int64_t ld_codes, int32_t* flags, qvit_stream_t stream) {
  QVIT_REQUIRE(x && d && q_m && codes, "qvit_quantize_sym: null pointer");
  QVIT_REQUIRE(rows >= 0 && cols >= 0 && ld_x >= cols && ld_codes >= cols, "qvit_quantize_sym: bad shape");
  if (rows == 0 || ld_codes == 0) return QVIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool flat = (ld_x == cols) && (ld_codes == cols) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(codes) & 3) == 0);
  if (flat) {
    const int64_t n = rows * cols;
    quantize_sym_flat_kernel<false><<<stream_grid(n, kBlockElems), kThreads, 0, s>>>(x, n, d, q_m, t, codes, flags);
  } else {
    const int64_t n = rows * ld_codes;
    quantize_sym_rows_kernel<<<stream_grid(n, kThreads * 4), kThreads, 0, s>>>(x, rows, cols, ld_x, d, q_m, t, codes,
                                                                             ld_codes, flags);
  }
  return check_launch("qvit_quantize_sym");
}

int qvit_gelu_quantize_sym(const float* x, int64_t n, const float* d, const float* q_m, const float* t, int8_t* codes, int32_t* flags,
                           qvit_stream_t stream) {
  QVIT_REQUIRE(x && d && q_m && codes && n >= 0, "qvit_gelu_quantize_sym: bad argument");
  QVIT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(codes) & 3) == 0,
               "qvit_gelu_quantize_sym: x must be 16-byte and codes 4-byte aligned");
  if (n == 0) return QVIT_OK;
  quantize_sym_flat_kernel<true><<<stream_grid(n, kBlockElems), kThreads, 0, (cudaStream_t)stream>>>(x, n, d, q_m, t, codes, flags);
  return check_launch("qvit_gelu_quantize_sym");
}

int qvit_fake_quantize_sym(const float* x, int64_t n, const float* d, const float* q_m, const float* t, float* out,
                           qvit_stream_t stream) {
  QVIT_REQUIRE(x && d && q_m && out && n >= 0, "qvit_fake_quantize_sym: bad argument");
  if (n == 0) return QVIT_OK;
  fake_quantize_sym_kernel<<<stream_grid(n, kThreads * 4), kThreads, 0, (cudaStream_t)stream>>>(x, n, d, q_m, t, out);
  return check_launch("qvit_fake_quantize_sym");
}

int qvit_quantize_sym_bf16(const void* x, int64_t rows, int64_t cols, int64_t ld_x, const float* d, const float* q_m,
                           const float* t, int8_t* codes, int64_t ld_codes, int32_t* flags, qvit_stream_t stream) {
  QVIT_REQUIRE(x && d && q_m && codes, "qvit_quantize_sym_bf16: null pointer");
  QVIT_REQUIRE(rows >= 0 && cols >= 0 && ld_x >= cols && ld_codes >= cols, "qvit_quantize_sym_bf16: bad shape");
  if (rows == 0 || ld_codes == 0) return QVIT_OK;
  const int64_t n = rows * ld_codes;
  quantize_sym_bf16_rows_kernel<<<stream_grid(n, kThreads * 8), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), rows, cols, ld_x, d, q_m, t, codes, ld_codes, flags);
  return check_launch("qvit_quantize_sym_bf16");
}

int qvit_im2col_quantize_sym(const float* x, int B, int C, int H, int W, int kh, int kw, int sh, int sw, int ph,
                             int pw, int dh, int dw, const float* d, const float* q_m, const float* t, int8_t* cols,
                             int64_t ld_cols, int32_t* flags, qvit_stream_t stream) {
  QVIT_REQUIRE(x && d && q_m && cols, "qvit_im2col_quantize_sym: null pointer");
  QVIT_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 &&
                   ph >= 0 && pw >= 0, "qvit_im2col_quantize_sym: bad geometry");
  ConvGeom g{B, C, H, W, kh, kw, sh, sw, ph, pw, dh, dw, 0, 0, C * kh * kw};
  g.OH = (H + 2 * ph - dh * (kh - 1) - 1) / sh + 1;
  g.OW = (W + 2 * pw - dw * (kw - 1) - 1) / sw + 1;
  QVIT_REQUIRE(g.OH > 0 && g.OW > 0, "qvit_im2col_quantize_sym: empty output");
  QVIT_REQUIRE(ld_cols >= g.K && ld_cols % 4 == 0 && (reinterpret_cast<uintptr_t>(cols) & 3) == 0,
               "qvit_im2col_quantize_sym: ld_cols must be >= C*kh*kw and a multiple of 4");
  cudaStream_t s = (cudaStream_t)stream;
  if (kh == sh && kw == sw && ph == 0 && pw == 0 && dh == 1 && dw == 1 && (kw % 4) == 0 && (W % 4) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    if (ld_cols > g.K)      // K padding columns (none for ViT: K = 768)
      cudaMemsetAsync(cols, 0, (size_t)B * g.OH * g.OW * ld_cols, s);
    const int64_t patches = (int64_t)B * C * g.OH * g.OW;   // one warp each
    patchify_quantize_kernel<<<stream_grid(patches, kThreads / 32, 8), kThreads, 0, s>>>(x, g, d, q_m, t, cols, ld_cols, flags);
    return check_launch("qvit_im2col_quantize_sym(patchify)");
  }
  const int64_t words = (int64_t)B * g.OH * g.OW * (ld_cols / 4);
  im2col_quantize_kernel<<<stream_grid(words, kThreads * 2), kThreads, 0, (cudaStream_t)stream>>>(x, g, d, q_m, t, cols,
                                                                                                  ld_cols, flags);
  return check_launch("qvit_im2col_quantize_sym");
}

int qvit_layernorm_quantize(const float* x, int64_t rows, int cols, const float* gamma, const float* beta, float eps,
                            const float* d, const float* q_m, const float* t, int8_t* codes, int64_t ld_codes,
                            float* ln_out, int32_t* flags, qvit_stream_t stream) {
  QVIT_REQUIRE(x && gamma && beta && d && q_m && codes, "qvit_layernorm_quantize: null pointer");
  QVIT_REQUIRE(rows >= 0 && cols > 0 && ld_codes >= cols, "qvit_layernorm_quantize: bad shape");
  if (rows == 0) return QVIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = stream_grid(rows, kThreads / 32, 8);
  const bool fast = (cols % 128 == 0) && (cols / 128 <= 8) && (ld_codes % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(codes) & 3) == 0) &&
                    ((reinterpret_cast<uintptr_t>(gamma) & 15) == 0) && ((reinterpret_cast<uintptr_t>(beta) & 15) == 0) &&
                    (!ln_out || (reinterpret_cast<uintptr_t>(ln_out) & 15) == 0);
  // persistent sizing: as many CTAs as are resident at once (the row loop is software pipelined, a second wave would
  // only pay the pipeline prologue again)
#define QVIT_LN_CASE(V)                                                                                            \
  case V: {                                                                                                        \
    static int per_sm = 0;                                                                                         \
    if (per_sm == 0 &&                                                                                             \
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, layernorm_quantize_kernel<V>, kThreads, 0) !=      \
             cudaSuccess || per_sm < 1))                                                                           \
      per_sm = 1;                                                                                                  \
    const int g = stream_grid(rows, kThreads / 32, per_sm);                                                        \
    layernorm_quantize_kernel<V><<<g, kThreads, 0, s>>>(x, rows, gamma, beta, eps, d, q_m, t, codes, ld_codes,     \
                                                        ln_out, flags);                                            \
  } break;
  if (fast) {
    switch (cols / 128) {
      QVIT_LN_CASE(1) QVIT_LN_CASE(2) QVIT_LN_CASE(3) QVIT_LN_CASE(4) QVIT_LN_CASE(5) QVIT_LN_CASE(6) QVIT_LN_CASE(7)
      QVIT_LN_CASE(8)
    }
  } else {
    layernorm_quantize_generic_kernel<<<grid, kThreads, 0, s>>>(x, rows, cols, gamma, beta, eps, d, q_m, t, codes,
                                                                ld_codes, ln_out, flags);
  }
#undef QVIT_LN_CASE
  return check_launch("qvit_layernorm_quantize");
}

int qvit_embed_assemble(const float* tok, const float* pos, const float* cls, int B, int P, int D, float* h, qvit_stream_t stream) {
  QVIT_REQUIRE(tok && pos && cls && h && B > 0 && P > 0 && D > 0 && D % 4 == 0, "qvit_embed_assemble: bad argument (D must be a multiple of 4)");
  QVIT_REQUIRE(((reinterpret_cast<uintptr_t>(tok) | reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(cls) |
                 reinterpret_cast<uintptr_t>(h)) & 15) == 0, "qvit_embed_assemble: 16-byte aligned tensors");
  const int64_t n4 = (int64_t)B * (P + 1) * (D / 4);
  embed_assemble_kernel<<<stream_grid(n4, kThreads * 4), kThreads, 0, (cudaStream_t)stream>>>(tok, pos, cls, B, P, D / 4, h);
  return check_launch("qvit_embed_assemble");
}

int qvit_absmax(const float* x, int64_t n, float* out, qvit_stream_t stream) {
  QVIT_REQUIRE(x && out && n >= 0, "qvit_absmax: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  zero_word_kernel<<<1, 1, 0, s>>>(reinterpret_cast<uint32_t*>(out));
  if (n > 0)
    absmax_kernel<<<stream_grid(n, kThreads * 8, 4), kThreads, 0, s>>>(x, n, reinterpret_cast<uint32_t*>(out));
  return check_launch("qvit_absmax");
}

int qvit_pack_int4(const int8_t* codes, int64_t n, uint8_t* packed, qvit_stream_t stream) {
  QVIT_REQUIRE(codes && packed && n >= 0 && n % 2 == 0, "qvit_pack_int4: n must be even");
  if (n == 0) return QVIT_OK;
  pack_int4_kernel<<<stream_grid(n / 2, kThreads * 4), kThreads, 0, (cudaStream_t)stream>>>(codes, n / 2, packed);
  return check_launch("qvit_pack_int4");
}

int qvit_pack_hls_weights(const int8_t* codes, int O, int I, int kh, int kw, int w_bit, int simd, int pe,
                          unsigned long long* words, qvit_stream_t stream) {
  QVIT_REQUIRE(codes && words && O > 0 && I > 0 && kh > 0 && kw > 0, "qvit_pack_hls_weights: bad argument");
  QVIT_REQUIRE(w_bit >= 1 && simd >= 1 && simd * w_bit <= 64, "qvit_pack_hls_weights: simd * w_bit must fit 64 bits");
  QVIT_REQUIRE(pe >= 1 && O % pe == 0, "qvit_pack_hls_weights: out_ch mod pe must be 0 (qnn_mem_process.py:86)");
  const int runs = (kh * kw * I + simd - 1) / simd;
  const int64_t n = (int64_t)O * runs;
  pack_hls_weights_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      codes, O, I, kh, kw, w_bit, simd, pe, runs, words);
  return check_launch("qvit_pack_hls_weights");
}

int qvit_unpack_int4(const uint8_t* packed, int64_t n, int is_signed, int8_t* codes, qvit_stream_t stream) {
  QVIT_REQUIRE(codes && packed && n >= 0 && n % 2 == 0, "qvit_unpack_int4: n must be even");
  if (n == 0) return QVIT_OK;
  unpack_int4_kernel<<<stream_grid(n / 2, kThreads * 4), kThreads, 0, (cudaStream_t)stream>>>(packed, n / 2, is_signed,
                                                                                             codes);
  return check_launch("qvit_unpack_int4");
}

}  // extern "C"
