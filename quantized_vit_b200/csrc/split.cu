// Operand preparation for the QAT gradient GEMMs (grad_x_q = g @ w_q, grad_w_q = g^T @ x_q; the backward of the
// F.linear in QuantizeLinear.forward, quant_layers.py:499).  The reference never quantizes the gradient g, so the int8
// pipe cannot be used; instead g is split EXACTLY into three bf16 planes (8 + 8 + 8 mantissa bits) and the integer
// codes (|code| <= 127, exact in bf16) become one bf16 plane, so that  (g1 + g2 + g3) * codes  on the bf16 tensor cores
// with fp32 accumulation reproduces the fp32 product.  HBM-bound elementwise / transpose kernels.
#include "common.cuh"

namespace qvit {

__device__ __forceinline__ void split3_pair_bf16(float x, float y, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(x, y);
  const float2 af = __bfloat1622float2(a);
  const float rx = x - af.x, ry = y - af.y;                    // exact
  const __nv_bfloat162 b = __floats2bfloat162_rn(rx, ry);
  const float2 bf = __bfloat1622float2(b);
  const __nv_bfloat162 c = __floats2bfloat162_rn(rx - bf.x, ry - bf.y);
  p1 = *reinterpret_cast<const uint32_t*>(&a);
  p2 = *reinterpret_cast<const uint32_t*>(&b);
  p3 = *reinterpret_cast<const uint32_t*>(&c);
}

// x [R, C] fp32 (pitch ld_x) -> out [R, 3 * Cp] bf16, plane p in columns [p * Cp, p * Cp + Cp); columns >= C are zero
__global__ void __launch_bounds__(256)
split3_rows_kernel(const float* __restrict__ x, int64_t R, int C, int64_t ld_x, int Cp, __nv_bfloat16* __restrict__ out) {
  const int chunks = Cp / 8;
  const int64_t total = R * chunks;
  const bool vec = ((ld_x & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / chunks;
    const int c0 = (int)(i - r * chunks) * 8;
    float v[8];
    const float* src = x + r * ld_x + c0;
    if (vec && c0 + 8 <= C) {
      const float4 a = ldg_stream4(src), b = ldg_stream4(src + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? src[j] : 0.0f;
    }
    uint4 p1, p2, p3;
    split3_pair_bf16(v[0], v[1], p1.x, p2.x, p3.x);
    split3_pair_bf16(v[2], v[3], p1.y, p2.y, p3.y);
    split3_pair_bf16(v[4], v[5], p1.z, p2.z, p3.z);
    split3_pair_bf16(v[6], v[7], p1.w, p2.w, p3.w);
    __nv_bfloat16* dst = out + r * (3ll * Cp) + c0;
    *reinterpret_cast<uint4*>(dst) = p1;
    *reinterpret_cast<uint4*>(dst + Cp) = p2;
    *reinterpret_cast<uint4*>(dst + 2 * Cp) = p3;
  }
}

// x [R, C] fp32 -> out [C, 3 * Rp] bf16 (transposed), plane p in columns [p * Rp, p * Rp + Rp); columns >= R are zero.
// Tile = 64 rows (r) x 32 columns (c) through shared memory: coalesced 128-byte reads and 128-byte writes.
__global__ void __launch_bounds__(256)
split3_transpose_kernel(const float* __restrict__ x, int64_t R, int C, int64_t ld_x, int64_t Rp, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[64][33];
  const int64_t r0 = (int64_t)blockIdx.y * 64;
  const int c0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;      // 8 warps
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = r0 + wy * 8 + i;
    const int c = c0 + lane;
    tile[wy * 8 + i][lane] = (r < R && c < C) ? x[r * ld_x + c] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cl = wy * 4 + i;                                   // output row inside the tile
    const int c = c0 + cl;
    if (c >= C) continue;
    uint32_t p1, p2, p3;
    split3_pair_bf16(tile[2 * lane][cl], tile[2 * lane + 1][cl], p1, p2, p3);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (int64_t)c * (3 * Rp) + r0) + lane;   // r0 is a multiple of 64
    dst[0] = p1;
    dst[Rp / 2] = p2;
    dst[Rp] = p3;
  }
}

// Both forms of the gradient operand and the bias gradient from ONE read of g [R, C]: the row planes [R, 3 * Cp] (A operand of
// grad_x_q = g @ w_q), the transposed planes [C, 3 * Rp] (A operand of grad_w_q = g^T @ x_q) and per-block column sums
// partial[blockIdx.y, c] over the block's kPrepTiles x 64 rows (summed by colsum_reduce_kernel in a fixed order: the bias
// gradient is reproducible run to run).  Tile = 64 rows x 32 columns through shared memory.
constexpr int kPrepTiles = 4;
__global__ void __launch_bounds__(256)
grad_prep_kernel(const float* __restrict__ x, int64_t R, int C, int64_t ld_x, int Cp, __nv_bfloat16* __restrict__ rows_out, int64_t Rp,
                 __nv_bfloat16* __restrict__ trans_out, float* __restrict__ partial) {
  __shared__ float tile[64][33];
  const int c0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;      // 8 warps
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int rt = 0; rt < kPrepTiles; ++rt) {
    const int64_t r0 = ((int64_t)blockIdx.y * kPrepTiles + rt) * 64;
    if (r0 >= Rp) break;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = r0 + wy * 8 + i;
      const int c = c0 + lane;
      tile[wy * 8 + i][lane] = (r < R && c < C) ? __ldcs(x + r * ld_x + c) : 0.0f;
    }
    __syncthreads();
    // transposed planes + column sums: this warp owns 4 columns, the lane 2 rows
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cl = wy * 4 + i;
      const int c = c0 + cl;
      const float a = tile[2 * lane][cl], b = tile[2 * lane + 1][cl];
      csum[i] += a + b;
      if (trans_out && c < C) {
        uint32_t p1, p2, p3;
        split3_pair_bf16(a, b, p1, p2, p3);
        uint32_t* dst = reinterpret_cast<uint32_t*>(trans_out + (int64_t)c * (3 * Rp) + r0) + lane;   // r0 is a multiple of 64
        dst[0] = p1;
        dst[Rp / 2] = p2;
        dst[Rp] = p3;
      }
    }
    // row planes: thread = (row, 8-column piece); bank = (row + column) mod 32 is distinct across the warp
    if (rows_out) {
      const int r = threadIdx.x >> 2, pc = (threadIdx.x & 3) * 8;
      if (r0 + r < R && c0 + pc < Cp) {
        uint4 p1, p2, p3;
        split3_pair_bf16(tile[r][pc], tile[r][pc + 1], p1.x, p2.x, p3.x);
        split3_pair_bf16(tile[r][pc + 2], tile[r][pc + 3], p1.y, p2.y, p3.y);
        split3_pair_bf16(tile[r][pc + 4], tile[r][pc + 5], p1.z, p2.z, p3.z);
        split3_pair_bf16(tile[r][pc + 6], tile[r][pc + 7], p1.w, p2.w, p3.w);
        __nv_bfloat16* dst = rows_out + (r0 + r) * (3ll * Cp) + c0 + pc;
        *reinterpret_cast<uint4*>(dst) = p1;
        *reinterpret_cast<uint4*>(dst + Cp) = p2;
        *reinterpret_cast<uint4*>(dst + 2 * Cp) = p3;
      }
    }
    __syncthreads();
  }
  if (partial) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float s = warp_sum(csum[i]);
      const int c = c0 + wy * 4 + i;
      if (lane == 0 && c < C) partial[(int64_t)blockIdx.y * C + c] = s;
    }
  }
}

// Row planes + per-block column sums WITHOUT the transposed planes (the weight-gradient GEMM reads the row planes MN-major): no
// shared-memory tile, every thread streams 8 columns of its rows (32-byte loads, three 16-byte stores) and keeps the 8 column
// sums in registers.  Block = 32 column groups (256 columns) x 8 row lanes over kPrepTiles * 64 = 256 rows.
__global__ void __launch_bounds__(256)
grad_rows_kernel(const float* __restrict__ x, int64_t R, int C, int64_t ld_x, int Cp, __nv_bfloat16* __restrict__ rows_out,
                 float* __restrict__ partial) {
  __shared__ float red[8][257];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + tx * 8;
  const int64_t r0 = (int64_t)blockIdx.y * (kPrepTiles * 64);
  const bool vec = ((ld_x & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && c0 + 8 <= C;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < Cp) {
    for (int64_t r = r0 + ty; r < r0 + kPrepTiles * 64 && r < R; r += 8) {
      float v[8];
      const float* src = x + r * ld_x + c0;
      if (vec) {
        const float4 a = ldg_stream4(src), b = ldg_stream4(src + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? src[j] : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[j] += v[j];
      uint4 p1, p2, p3;
      split3_pair_bf16(v[0], v[1], p1.x, p2.x, p3.x);
      split3_pair_bf16(v[2], v[3], p1.y, p2.y, p3.y);
      split3_pair_bf16(v[4], v[5], p1.z, p2.z, p3.z);
      split3_pair_bf16(v[6], v[7], p1.w, p2.w, p3.w);
      __nv_bfloat16* dst = rows_out + r * (3ll * Cp) + c0;
      *reinterpret_cast<uint4*>(dst) = p1;
      *reinterpret_cast<uint4*>(dst + Cp) = p2;
      *reinterpret_cast<uint4*>(dst + 2 * Cp) = p3;
    }
  }
  if (partial) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = cs[j];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < C) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += red[i][threadIdx.x];
      partial[(int64_t)blockIdx.y * C + c] = sum;
    }
  }
}

__global__ void colsum_reduce_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int i = 0; i < nblk; ++i) s += partial[(int64_t)i * C + c];
  out[c] = s;
}

// codes [R, C] int8 (pitch ld) -> out [C, Rp] bf16 (transposed); columns >= R are zero.
// Tile = 64 rows (r) x 128 columns (c) through shared memory: a warp reads one 128-byte row segment per instruction (4 bytes per
// lane) and writes 64-byte runs of one output row (lane = r, then r + 32: bank = (33 r + c / 4) mod 32 is distinct across lanes).
__global__ void __launch_bounds__(256)
codes_transpose_bf16_kernel(const int8_t* __restrict__ codes, int64_t R, int C, int64_t ld, int64_t Rp,
                            __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) int8_t tile[64][132];
  const int64_t r0 = (int64_t)blockIdx.y * 64;
  const int c0 = blockIdx.x * 128;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;      // 8 warps
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(codes) & 3) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = wy * 8 + i;
    const int64_t r = r0 + rl;
    const int c = c0 + lane * 4;
    uint32_t w = 0;
    if (r < R) {
      if (vec && c + 4 <= C) {
        w = __ldg(reinterpret_cast<const uint32_t*>(codes + r * ld + c));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < C) w |= (uint32_t)(uint8_t)codes[r * ld + c + j] << (8 * j);
      }
    }
    *reinterpret_cast<uint32_t*>(&tile[rl][lane * 4]) = w;
  }
  __syncthreads();
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int cl = wy * 16 + i;                                  // output row inside the tile
    const int c = c0 + cl;
    if (c >= C) break;
    __nv_bfloat16* dst = out + (int64_t)c * Rp + r0;             // r0 + 63 < Rp (Rp is a multiple of 64)
    dst[lane] = __float2bfloat16_rn((float)tile[lane][cl]);
    dst[lane + 32] = __float2bfloat16_rn((float)tile[lane + 32][cl]);
  }
}

// codes [R, C] int8 (pitch ld) -> out [R, Cp] bf16 (same orientation); columns >= C are zero.  16 codes per thread.
__global__ void __launch_bounds__(256)
codes_to_bf16_kernel(const int8_t* __restrict__ codes, int64_t R, int C, int64_t ld, int Cp, __nv_bfloat16* __restrict__ out) {
  const int pieces = Cp / 16;
  const int64_t total = R * pieces;
  const bool vec = ((ld & 15) == 0) && ((reinterpret_cast<uintptr_t>(codes) & 15) == 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / pieces;
    const int c0 = (int)(i - r * pieces) * 16;
    int8_t v[16];
    if (vec && c0 + 16 <= C) {
      *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(codes + r * ld + c0));
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (c0 + j < C) ? codes[r * ld + c0 + j] : (int8_t)0;
    }
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat162 p = __floats2bfloat162_rn((float)v[2 * j], (float)v[2 * j + 1]);
      w[j] = *reinterpret_cast<const uint32_t*>(&p);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + r * (int64_t)Cp + c0);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

int gemm_tc_launch_bf16_split(const void* a, int64_t lda, int planes, const void* b, int64_t ldb, const struct EpiParams& ep, int K,
                              cudaStream_t s);
int gemm_tc_launch_bf16_split_t(const void* g, int64_t ldg, int planes, int64_t plane_cols, const void* x, int64_t ldx, int64_t tokens,
                                const struct EpiParams& ep, float* workspace, cudaStream_t s);

}  // namespace qvit

#include "epilogue.cuh"
using namespace qvit;

extern "C" {

int qvit_split3_bf16(const float* x, int64_t rows, int64_t cols, int64_t ld_x, int transpose, void* out, int64_t plane_cols,
                     qvit_stream_t stream) {
  QVIT_REQUIRE(x && out && rows > 0 && cols > 0 && ld_x >= cols, "qvit_split3_bf16: bad argument");
  QVIT_REQUIRE(plane_cols % 64 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "qvit_split3_bf16: plane_cols must be a multiple of 64");
  QVIT_REQUIRE(cols < (1ll << 31) && rows < (1ll << 40), "qvit_split3_bf16: too large");
  cudaStream_t s = (cudaStream_t)stream;
  if (!transpose) {
    QVIT_REQUIRE(plane_cols >= cols, "qvit_split3_bf16: plane_cols < cols");
    const int64_t total = rows * (plane_cols / 8);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    split3_rows_kernel<<<(int)blocks, 256, 0, s>>>(x, rows, (int)cols, ld_x, (int)plane_cols, reinterpret_cast<__nv_bfloat16*>(out));
  } else {
    QVIT_REQUIRE(plane_cols >= rows, "qvit_split3_bf16: plane_cols < rows (transposed)");
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)(plane_cols / 64));
    QVIT_REQUIRE(grid.y <= 65535u, "qvit_split3_bf16: too many rows for the transposed form");
    split3_transpose_kernel<<<grid, 256, 0, s>>>(x, rows, (int)cols, ld_x, plane_cols, reinterpret_cast<__nv_bfloat16*>(out));
  }
  return check_launch("qvit_split3_bf16");
}

int qvit_grad_prep(const float* g, int64_t rows, int64_t cols, int64_t ld_g, void* rows_out, int64_t row_plane_cols, void* trans_out,
                   int64_t trans_plane_cols, float* partial, float* colsum, qvit_stream_t stream) {
  QVIT_REQUIRE(g && (trans_out || rows_out) && rows > 0 && cols > 0 && ld_g >= cols, "qvit_grad_prep: bad argument");
  QVIT_REQUIRE(trans_plane_cols % 64 == 0 && trans_plane_cols >= rows && (reinterpret_cast<uintptr_t>(trans_out) & 15) == 0,
               "qvit_grad_prep: trans_plane_cols must be a multiple of 64, >= rows (also without trans_out: it sizes the row blocks)");
  QVIT_REQUIRE(!rows_out || (row_plane_cols % 64 == 0 && row_plane_cols >= cols && (reinterpret_cast<uintptr_t>(rows_out) & 15) == 0),
               "qvit_grad_prep: row_plane_cols must be a multiple of 64, >= cols");
  QVIT_REQUIRE((colsum == nullptr) == (partial == nullptr), "qvit_grad_prep: colsum needs the partial workspace (and vice versa)");
  QVIT_REQUIRE(cols < (1ll << 31) && rows < (1ll << 40), "qvit_grad_prep: too large");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t cp = rows_out ? row_plane_cols : (cols + 63) / 64 * 64;
  const int64_t tiles_y = trans_plane_cols / 64;
  dim3 grid((unsigned)(cp / 32), (unsigned)((tiles_y + kPrepTiles - 1) / kPrepTiles));
  QVIT_REQUIRE(grid.y <= 65535u, "qvit_grad_prep: too many rows");
  if (!trans_out) {
    grid.x = (unsigned)((cp + 255) / 256);
    grad_rows_kernel<<<grid, 256, 0, s>>>(g, rows, (int)cols, ld_g, (int)cp, reinterpret_cast<__nv_bfloat16*>(rows_out), partial);
  } else
  grad_prep_kernel<<<grid, 256, 0, s>>>(g, rows, (int)cols, ld_g, (int)cp, reinterpret_cast<__nv_bfloat16*>(rows_out), trans_plane_cols,
                                        reinterpret_cast<__nv_bfloat16*>(trans_out), partial);
  int rc = check_launch("qvit_grad_prep");
  if (rc || !colsum) return rc;
  colsum_reduce_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, s>>>(partial, (int)grid.y, (int)cols, colsum);
  return check_launch("qvit_grad_prep (colsum)");
}

int qvit_codes_to_bf16_t(const int8_t* codes, int64_t rows, int64_t cols, int64_t ld, void* out, int64_t out_cols,
                         qvit_stream_t stream) {
  QVIT_REQUIRE(codes && out && rows > 0 && cols > 0 && ld >= cols, "qvit_codes_to_bf16_t: bad argument");
  QVIT_REQUIRE(out_cols % 64 == 0 && out_cols >= rows && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
               "qvit_codes_to_bf16_t: out_cols must be a multiple of 64 and >= rows");
  dim3 grid((unsigned)((cols + 127) / 128), (unsigned)(out_cols / 64));
  QVIT_REQUIRE(grid.y <= 65535u, "qvit_codes_to_bf16_t: too many rows");
  codes_transpose_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(codes, rows, (int)cols, ld, out_cols,
                                                                     reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("qvit_codes_to_bf16_t");
}

int qvit_codes_to_bf16(const int8_t* codes, int64_t rows, int64_t cols, int64_t ld, void* out, int64_t out_cols, qvit_stream_t stream) {
  QVIT_REQUIRE(codes && out && rows > 0 && cols > 0 && ld >= cols, "qvit_codes_to_bf16: bad argument");
  QVIT_REQUIRE(out_cols % 64 == 0 && out_cols >= cols && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "qvit_codes_to_bf16: out_cols must be a multiple of 64 and >= cols");
  const int64_t total = rows * (out_cols / 16);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  codes_to_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(codes, rows, (int)cols, ld, (int)out_cols,
                                                                     reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("qvit_codes_to_bf16");
}

int qvit_gemm_bf16_split_t(const void* g_planes, int64_t ld_g, int planes, int64_t plane_cols, const void* x, int64_t ld_x, int64_t tokens,
                           int N_out, int K_in, float* out, int64_t ldo, float* workspace, const qvit_epilogue_t* epi, qvit_stream_t stream) {
  QVIT_REQUIRE(g_planes && x && out && epi, "qvit_gemm_bf16_split_t: null pointer");
  QVIT_REQUIRE(tokens > 0 && N_out > 0 && K_in > 0 && planes >= 1 && planes <= 3 && ldo >= K_in, "qvit_gemm_bf16_split_t: bad shape");
  QVIT_REQUIRE(plane_cols % 64 == 0 && plane_cols >= N_out && ld_g >= (int64_t)planes * plane_cols && (ld_g % 8) == 0 && (ld_x % 8) == 0 &&
                   ld_x >= K_in, "qvit_gemm_bf16_split_t: plane_cols a multiple of 64 >= N_out, ld_g >= planes * plane_cols, pitches multiples of 8");
  QVIT_REQUIRE(((reinterpret_cast<uintptr_t>(g_planes) | reinterpret_cast<uintptr_t>(x)) & 15) == 0, "qvit_gemm_bf16_split_t: alignment");
  QVIT_REQUIRE(epi->out_kind == QVIT_OUT_F32 && epi->act == QVIT_ACT_NONE && !epi->residual, "qvit_gemm_bf16_split_t: plain fp32 output only");
  QVIT_REQUIRE(tokens < (1ll << 31) - 64 && (int64_t)planes * plane_cols * 2 < (1ll << 31), "qvit_gemm_bf16_split_t: too large");
  int dev = 0, maj = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  if (maj != 10) {
    set_error("qvit_gemm_bf16_split_t: needs sm_100");
    return QVIT_ERR_UNSUPPORTED;
  }
  EpiParams ep;
  ep.out_kind = QVIT_OUT_F32;
  ep.act = QVIT_ACT_NONE;
  ep.scale_const = epi->scale_const;
  ep.acc_abs_max = 0;
  ep.scale_a = epi->scale_a;
  ep.scale_w = epi->scale_w;
  ep.col_scale = epi->col_scale;
  ep.bias = epi->bias;
  ep.residual = nullptr;
  ep.ld_res = 0;
  ep.next_d = ep.next_qm = ep.next_t = nullptr;
  ep.flags = epi->flags;
  ep.out = out;
  ep.ldo = ldo;
  ep.M = N_out;
  ep.N = K_in;
  return gemm_tc_launch_bf16_split_t(g_planes, ld_g, planes, plane_cols, x, ld_x, tokens, ep, workspace, (cudaStream_t)stream);
}

int qvit_gemm_bf16_split(const void* a_planes, int64_t lda, int planes, const void* b, int64_t ldb, int M, int N, int K,
                         float* out, int64_t ldo, const qvit_epilogue_t* epi, qvit_stream_t stream) {
  QVIT_REQUIRE(a_planes && b && out && epi, "qvit_gemm_bf16_split: null pointer");
  QVIT_REQUIRE(M > 0 && N > 0 && K > 0 && planes >= 1 && planes <= 3 && ldo >= N, "qvit_gemm_bf16_split: bad shape");
  const int Kp = (K + 63) / 64 * 64;
  QVIT_REQUIRE(lda >= (int64_t)planes * Kp && ldb >= Kp && (lda % 8) == 0 && (ldb % 8) == 0,
               "qvit_gemm_bf16_split: lda >= planes*Kp, ldb >= Kp (Kp = K rounded up to 64) and both multiples of 8 elements");
  QVIT_REQUIRE(((reinterpret_cast<uintptr_t>(a_planes) | reinterpret_cast<uintptr_t>(b)) & 15) == 0, "qvit_gemm_bf16_split: alignment");
  QVIT_REQUIRE(epi->out_kind == QVIT_OUT_F32 && epi->act == QVIT_ACT_NONE, "qvit_gemm_bf16_split: fp32 output without activation only");
  int dev = 0, maj = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  if (maj != 10) {
    set_error("qvit_gemm_bf16_split: needs sm_100");
    return QVIT_ERR_UNSUPPORTED;
  }
  EpiParams ep;
  ep.out_kind = QVIT_OUT_F32;
  ep.act = QVIT_ACT_NONE;
  ep.scale_const = epi->scale_const;
  ep.acc_abs_max = 0;
  ep.scale_a = epi->scale_a;
  ep.scale_w = epi->scale_w;
  ep.col_scale = epi->col_scale;
  ep.bias = epi->bias;
  ep.residual = epi->residual;
  ep.ld_res = epi->ld_res;
  ep.next_d = ep.next_qm = ep.next_t = nullptr;
  ep.flags = epi->flags;
  ep.out = out;
  ep.ldo = ldo;
  ep.M = M;
  ep.N = N;
  return gemm_tc_launch_bf16_split(a_planes, lda, planes, b, ldb, ep, K, (cudaStream_t)stream);
}

}  // extern "C"
