// LayerNorm forward / backward for the QAT step's caller (Block.norm1 / norm2 and the final norm, vit_model.py:202-208, 309):
// glue beside the hot path.  The inference engine fuses LayerNorm with the consumer's quantizer
// (quantize.cu); training needs the fp32 output and a backward.  One warp per row, the row lives in registers (cols = 128 * V),
// next row's loads in flight while the current one is reduced.
//   forward : y = (x - mean) * rstd * gamma + beta, saves mean / rstd per row          (8 B/elem of traffic)
//   backward: gx = rstd * (g' - mean_j(g') - xhat * mean_j(g' * xhat)),  g' = gy * gamma, xhat = (x - mean) * rstd;
//             dgamma = sum_rows gy * xhat, dbeta = sum_rows gy - accumulated per lane in registers over the warp's rows, reduced
//             across the block in shared memory, one atomicAdd per column and block                (12 B/elem of traffic)
#include "common.cuh"

namespace qvit {

constexpr int kLnThreads = 256;

template <int V>
__global__ void __launch_bounds__(kLnThreads) layernorm_fwd_kernel(const float* __restrict__ x, int64_t rows, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float eps, float* __restrict__ y,
                                                                    float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  constexpr int cols = V * 128;
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5), ws = (int64_t)gridDim.x * (kLnThreads / 32);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float4 nxt[V];
  if (w0 < rows) {
#pragma unroll
    for (int j = 0; j < V; ++j) nxt[j] = ldg_stream4(x + w0 * cols + j * 128 + lane * 4);
  }
  for (int64_t r = w0; r < rows; r += ws) {
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = nxt[j];
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    if (r + ws < rows) {
#pragma unroll
      for (int j = 0; j < V; ++j) nxt[j] = ldg_stream4(x + (r + ws) * cols + j * 128 + lane * 4);
    }
    const float mean = warp_sum(s) * (1.0f / cols);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / cols) + eps);
    if (lane == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 gm = __ldg(g4 + j * 32 + lane), bt = __ldg(b4 + j * 32 + lane);
      float4 o;
      o.x = (v[j].x - mean) * rstd * gm.x + bt.x;
      o.y = (v[j].y - mean) * rstd * gm.y + bt.y;
      o.z = (v[j].z - mean) * rstd * gm.z + bt.z;
      o.w = (v[j].w - mean) * rstd * gm.w + bt.w;
      reinterpret_cast<float4*>(y + r * cols)[j * 32 + lane] = o;
    }
  }
}

template <int V>
__global__ void __launch_bounds__(kLnThreads) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, int64_t rows,
                                                                    const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                                    const float* __restrict__ rstd_in, float* __restrict__ gx,
                                                                    float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                    const float* __restrict__ add) {
  constexpr int cols = V * 128;
  __shared__ float red[kLnThreads / 32][128];                    // one 128-column slice at a time
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t w0 = (int64_t)blockIdx.x * (kLnThreads / 32) + warp, ws = (int64_t)gridDim.x * (kLnThreads / 32);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 dg[V], db[V];
#pragma unroll
  for (int j = 0; j < V; ++j) dg[j] = db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = w0; r < rows; r += ws) {
    const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
    float4 xh[V], gp[V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 xv = ldg_stream4(x + r * cols + j * 128 + lane * 4);
      const float4 gv = ldg_stream4(gy + r * cols + j * 128 + lane * 4);
      const float4 gm = __ldg(g4 + j * 32 + lane);
      xh[j] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      gp[j] = make_float4(gv.x * gm.x, gv.y * gm.y, gv.z * gm.z, gv.w * gm.w);
      s1 += (gp[j].x + gp[j].y) + (gp[j].z + gp[j].w);
      s2 += (gp[j].x * xh[j].x + gp[j].y * xh[j].y) + (gp[j].z * xh[j].z + gp[j].w * xh[j].w);
      dg[j].x += gv.x * xh[j].x; dg[j].y += gv.y * xh[j].y; dg[j].z += gv.z * xh[j].z; dg[j].w += gv.w * xh[j].w;
      db[j].x += gv.x; db[j].y += gv.y; db[j].z += gv.z; db[j].w += gv.w;
    }
    const float c1 = warp_sum(s1) * (1.0f / cols), c2 = warp_sum(s2) * (1.0f / cols);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 o;
      o.x = __fmul_rn(rstd, gp[j].x - c1 - xh[j].x * c2);
      o.y = __fmul_rn(rstd, gp[j].y - c1 - xh[j].y * c2);
      o.z = __fmul_rn(rstd, gp[j].z - c1 - xh[j].z * c2);
      o.w = __fmul_rn(rstd, gp[j].w - c1 - xh[j].w * c2);
      if (add) {                                             // gradient arriving over the residual connection (same rounding as a separate add)
        const float4 a = ldg_stream4(add + r * cols + j * 128 + lane * 4);
        o.x = __fadd_rn(a.x, o.x);
        o.y = __fadd_rn(a.y, o.y);
        o.z = __fadd_rn(a.z, o.z);
        o.w = __fadd_rn(a.w, o.w);
      }
      reinterpret_cast<float4*>(gx + r * cols)[j * 32 + lane] = o;
    }
  }
  // column sums of this block: warps -> shared memory -> one atomicAdd per column
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 v = pass == 0 ? dg[j] : db[j];
      __syncthreads();
      reinterpret_cast<float4*>(&red[warp][0])[lane] = v;
      __syncthreads();
      if (threadIdx.x < 128) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kLnThreads / 32; ++w) s += red[w][threadIdx.x];
        atomicAdd((pass == 0 ? dgamma : dbeta) + j * 128 + threadIdx.x, s);
      }
    }
  }
}

}  // namespace qvit

using namespace qvit;

static int ln_grid(int64_t rows, int waves = 4) {
  const int64_t want = (rows + (kLnThreads / 32) - 1) / (kLnThreads / 32);
  const int64_t cap = (int64_t)sm_count() * waves;
  return (int)(want < cap ? want : cap);
}

// y = LayerNorm(x) over the last dimension (cols = 128 * V, V <= 8) with the statistics the backward needs.
extern "C" int qvit_layernorm_fwd(const float* x, int64_t rows, int cols, const float* gamma, const float* beta, float eps, float* y,
                                  float* mean, float* rstd, qvit_stream_t stream) {
  QVIT_REQUIRE(x && gamma && beta && y && mean && rstd && rows >= 0, "qvit_layernorm_fwd: null pointer");
  if (cols <= 0 || cols % 128 != 0 || cols > 1024) {
    set_error("qvit_layernorm_fwd: cols must be a multiple of 128, <= 1024 (got %d)", cols);
    return QVIT_ERR_UNSUPPORTED;
  }
  QVIT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                 reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "qvit_layernorm_fwd: 16-byte aligned tensors");
  if (rows == 0) return QVIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = ln_grid(rows);
#define QVIT_LNF(V) case V: layernorm_fwd_kernel<V><<<g, kLnThreads, 0, s>>>(x, rows, gamma, beta, eps, y, mean, rstd); break;
  switch (cols / 128) { QVIT_LNF(1) QVIT_LNF(2) QVIT_LNF(3) QVIT_LNF(4) QVIT_LNF(5) QVIT_LNF(6) QVIT_LNF(7) QVIT_LNF(8) }
#undef QVIT_LNF
  return check_launch("qvit_layernorm_fwd");
}

// gx, dgamma, dbeta of LayerNorm from x, grad_output and the saved statistics (dgamma / dbeta are overwritten).  add (optional,
// like x): a gradient that reaches x over another path (the residual connection around the block) - gx = add + LayerNorm'(gy).
extern "C" int qvit_layernorm_bwd(const float* x, const float* gy, int64_t rows, int cols, const float* gamma, const float* mean,
                                  const float* rstd, const float* add, float* gx, float* dgamma, float* dbeta, qvit_stream_t stream) {
  QVIT_REQUIRE(x && gy && gamma && mean && rstd && gx && dgamma && dbeta && rows >= 0, "qvit_layernorm_bwd: null pointer");
  if (cols <= 0 || cols % 128 != 0 || cols > 1024) {
    set_error("qvit_layernorm_bwd: cols must be a multiple of 128, <= 1024 (got %d)", cols);
    return QVIT_ERR_UNSUPPORTED;
  }
  QVIT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gx) |
                 reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(add)) & 15) == 0, "qvit_layernorm_bwd: 16-byte aligned tensors");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(dgamma, 0, sizeof(float) * cols, s) != cudaSuccess || cudaMemsetAsync(dbeta, 0, sizeof(float) * cols, s) != cudaSuccess) {
    set_error("qvit_layernorm_bwd: cudaMemsetAsync failed");
    return QVIT_ERR_CUDA;
  }
  if (rows == 0) return QVIT_OK;
  const int g = ln_grid(rows, 2);   // the wide instances hold one block per SM (150-190 registers); fewer blocks = fewer column atomics
#define QVIT_LNB(V) case V: layernorm_bwd_kernel<V><<<g, kLnThreads, 0, s>>>(x, gy, rows, gamma, mean, rstd, gx, dgamma, dbeta, add); break;
  switch (cols / 128) { QVIT_LNB(1) QVIT_LNB(2) QVIT_LNB(3) QVIT_LNB(4) QVIT_LNB(5) QVIT_LNB(6) QVIT_LNB(7) QVIT_LNB(8) }
#undef QVIT_LNB
  return check_launch("qvit_layernorm_bwd");
}
