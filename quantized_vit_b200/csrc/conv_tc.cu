// QuantConv2d as an IMPLICIT GEMM on the tcgen05 int8 tensor-core pipe (BASELINE north_star (4)): the fused UltraNet layer
//   uint8 NHWC activation codes --3x3 / 1x1 conv with int8 weight codes--> int32 (TMEM) --BN, clamp, round (, 2x2 max-pool)--> codes
// (Conv2d_Q.forward quant_ultra.py:85-89 + nn.BatchNorm2d(eval) + activation_quantize_fn + MaxPool2d, mymodel.py:71-125).
//
// GEMM view: M = output pixels (tiles of 128), N = output channels (padded to 16), K = taps * C ordered (tap, channel) - the
// order the NHWC input delivers 16-byte pieces in.  Nothing is materialised in global memory:
//   * weights [O_pad, K_pad] int8 (packed once per model) are copied once per CTA into K-major, 128B-swizzled shared-memory
//     tiles and stay resident while the persistent CTA walks its pixel tiles;
//   * the im2col tile A [128 pixels x K_pad] is GATHERED per tile straight from the NHWC codes (16-byte pieces = 16 channels
//     of one tap of one pixel, zero for padding / out-of-image pixels) into the same swizzled layout;
//   * one thread issues K_pad / 32 tcgen05.mma.kind::i8 (unsigned A x signed B, M = 128, N = O_pad); the accumulator lives
//     in TMEM; four warps (lane = pixel) read it back with tcgen05.ld and run the layer's integer-to-code epilogue - the
//     same arithmetic sequence as the CUDA-core kernel of ultra.cu, so the codes are identical - pooling on packed codes with
//     warp shuffles (a warp holds two 16-pixel rows of the tile).
// Tiles: 8 x 16 pixels when the layer pools (neighbours must share a warp), 128 consecutive pixels of the flattened image
// otherwise (small maps, e.g. 10 x 20, fill their tiles).
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace qvit {

namespace ctc {
constexpr int kThreads = 256;
constexpr int kTileM = 128;
constexpr int kTW = 16, kTH = 8;

struct Params {
  const uint8_t* in;
  const int8_t* wpk;
  const float* bn_scale;
  const float* bn_bias;
  uint8_t* out_codes;
  float* out_f32;
  const float* scale_a;      // optional (1,) device scalars multiplied into acc_scale (|.| taken): GETA's d_quant_act / d_quant_wt
  const float* scale_w;
  int B, H, W, C, O, O_pad, kh, kw, pad, K, K_pad;
  int sh, sw, dh, dw, a_signed;
  int OH, OW, pool, out_levels, linear, tiles_per_img, tiles_x, total_tiles, tmem_cols;
  float acc_scale;
};

__device__ __forceinline__ uint4 ldg16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
}  // namespace ctc

__global__ void __launch_bounds__(ctc::kThreads, 1) ultra_conv_tc_kernel(const ctc::Params p) {
  using namespace ctc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int chunks = p.K_pad >> 7;                                  // 128-byte k-chunks
  const uint32_t w_sm = base;                                       // chunks x [O_pad x 128 B]
  const uint32_t a_sm = w_sm + (uint32_t)(chunks * p.O_pad * 128);  // chunks x [128 x 128 B]   (O_pad % 8 == 0: 1024-byte aligned)
  const uint32_t bar = a_sm + (uint32_t)(chunks * kTileM * 128);
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (bar - base) + 16);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (visibly warp-uniform)

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish<1>();
  }
  // weights -> swizzled K-major tiles (once per CTA); zero the A tiles once (the k-padding columns stay zero)
  {
    const int pieces_per_row = p.K_pad >> 4;
    const int total = p.O_pad * pieces_per_row;
    for (int i = threadIdx.x; i < total; i += kThreads) {
      const int o = i / pieces_per_row, q = i - o * pieces_per_row;
      const uint4 v = ldg16(p.wpk + (int64_t)o * p.K_pad + q * 16);
      sts16(w_sm + (uint32_t)((q >> 3) * p.O_pad * 128 + o * 128 + (((q & 7) ^ (o & 7)) << 4)), v);
    }
    const int a_pieces = chunks * kTileM * 8;
    for (int i = threadIdx.x; i < a_pieces; i += kThreads) sts16(a_sm + (uint32_t)(i * 16), make_uint4(0, 0, 0, 0));
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int taps = p.kh * p.kw;
  const int cp = p.C >> 4;                                           // 16-byte pieces per (pixel, tap): 1, 2, 4 or 8
  const int cp_shift = 31 - __clz(cp);
  const uint32_t idesc = ptx::make_idesc_i8(kTileM, p.O_pad, p.a_signed != 0, true);
  float acc_scale = p.acc_scale;
  if (p.scale_a) acc_scale *= fabsf(__ldg(p.scale_a));
  if (p.scale_w) acc_scale *= fabsf(__ldg(p.scale_w));
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int b = tile / p.tiles_per_img;
    const int tin = tile - b * p.tiles_per_img;
    // pixel of this tile's row r -> (oy, ox); oy = -1 marks a row outside the image
    auto pixel = [&](int r, int& oy, int& ox) {
      if (p.linear) {
        const int idx = tin * kTileM + r;
        oy = idx / p.OW;
        ox = idx - oy * p.OW;
        if (idx >= p.OH * p.OW) oy = -1;
      } else {
        oy = (tin / p.tiles_x) * kTH + (r >> 4);
        ox = (tin % p.tiles_x) * kTW + (r & 15);
        if (oy >= p.OH || ox >= p.OW) oy = -1;
      }
    };
    // ---- gather the im2col tile: piece index = (tap * 128 + row) * cp + c16 (consecutive threads read consecutive bytes)
    {
      const int total = taps * kTileM * cp;
      for (int i = threadIdx.x; i < total; i += kThreads) {
        const int c16 = i & (cp - 1);
        const int tr = i >> cp_shift;
        const int r = tr & (kTileM - 1), tap = tr >> 7;
        int oy, ox;
        pixel(r, oy, ox);
        const int ky = tap / p.kw, kx = tap - ky * p.kw;
        const int iy = oy * p.sh + ky * p.dh - p.pad, ix = ox * p.sw + kx * p.dw - p.pad;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (oy >= 0 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
          v = ldg16(p.in + (((int64_t)b * p.H + iy) * p.W + ix) * p.C + c16 * 16);
        const int kb = tap * p.C + c16 * 16;                          // byte offset inside the row's K
        sts16(a_sm + (uint32_t)((kb >> 7) * kTileM * 128 + r * 128 + ((((kb >> 4) & 7) ^ (r & 7)) << 4)), v);
      }
    }
    ptx::fence_proxy_async_smem();                                    // generic-proxy writes -> visible to the tensor core
    ptx::tc_fence_before();                                           // (previous tile's TMEM reads are complete)
    __syncthreads();
    if (warp == 4) {
      // converged warp, one elected lane issues: the descriptors stay in uniform registers and every MMA is one instruction
      // (a `threadIdx.x == 128` guard wraps each in an elect / broadcast loop, tools/ubench/mma_rate.cu)
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        uint32_t acc = 0;
        for (int kb = 0; kb < p.K; kb += 32) {                        // 32 bytes of K per MMA; chunks beyond K hold zeros (skipped)
          const uint32_t ch = (uint32_t)(kb >> 7), within = (uint32_t)(kb & 127);
          ptx::mma_i8<1>(tmem, ptx::make_kmajor_sw128_desc(a_sm + ch * kTileM * 128 + within),
                         ptx::make_kmajor_sw128_desc(w_sm + ch * (uint32_t)p.O_pad * 128 + within), idesc, acc);
          acc = 1;
        }
        ptx::mma_commit(bar);
      }
      __syncwarp();
    }
    // every thread waits for the MMAs: the gather of the next tile overwrites the operand tile they read
    ptx::mbar_wait(bar, phase);
    // ---- epilogue: warps 0..3, lane = pixel row of the tile
    if (warp < 4) {
      ptx::tc_fence_after();
      const int r = warp * 32 + lane;
      int oy, ox;
      pixel(r, oy, ox);
      const float lv = (float)p.out_levels;
      const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
      for (int o0 = 0; o0 < p.O_pad; o0 += 16) {
        uint32_t acc[16];
        tmem_ld16(trow + (uint32_t)o0, acc);
        ptx::tmem_ld_wait();
        if (p.out_f32) {
          if (oy >= 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int o = o0 + j;
              if (o < p.O) {
                float yv = (float)(int32_t)acc[j] * acc_scale;
                if (p.bn_scale) yv *= __ldg(p.bn_scale + o);
                if (p.bn_bias) yv += __ldg(p.bn_bias + o);
                p.out_f32[(((int64_t)b * p.O + o) * p.OH + oy) * p.OW + ox] = yv;
              }
            }
          }
          continue;
        }
        uint32_t packed[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int o = o0 + j;
          float yv = (float)(int32_t)acc[j] * acc_scale;
          if (o < p.O) {
            if (p.bn_scale) yv *= __ldg(p.bn_scale + o);
            if (p.bn_bias) yv += __ldg(p.bn_bias + o);
          }
          const float cl = fminf(fmaxf(yv, 0.0f), 1.0f);
          const int code = (int)rintf(cl * lv);
          packed[j >> 2] |= (uint32_t)(code & 0xff) << (8 * (j & 3));
        }
        if (p.pool) {
          // 2x2 max-pool on codes (monotone code map => pool(codes) == codes(pool), MM:76): partners are lanes ^1 (x) and ^16 (y)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t v = oy >= 0 ? packed[q] : 0u;
            v = __vmaxu4(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = __vmaxu4(v, __shfl_xor_sync(0xffffffffu, v, 16));
            packed[q] = v;
          }
          if (oy >= 0 && !(lane & 17) && (oy >> 1) < (p.OH >> 1) && (ox >> 1) < (p.OW >> 1)) {
            uint8_t* dst = p.out_codes + ((((int64_t)b * (p.OH >> 1) + (oy >> 1)) * (p.OW >> 1)) + (ox >> 1)) * p.O + o0;
            if (o0 + 16 <= p.O && (p.O & 15) == 0) *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            else
              for (int j = 0; j < 16 && o0 + j < p.O; ++j) dst[j] = (uint8_t)(packed[j >> 2] >> (8 * (j & 3)));
          }
        } else if (oy >= 0) {
          uint8_t* dst = p.out_codes + (((int64_t)b * p.OH + oy) * p.OW + ox) * p.O + o0;
          if (o0 + 16 <= p.O && (p.O & 15) == 0) *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          else
            for (int j = 0; j < 16 && o0 + j < p.O; ++j) dst[j] = (uint8_t)(packed[j >> 2] >> (8 * (j & 3)));
        }
      }
    }
    phase ^= 1u;
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem, (uint32_t)p.tmem_cols);
  }
}

}  // namespace qvit

using namespace qvit;

// Fused UltraNet layer on the tensor cores.  in_codes: uint8 NHWC [B, H, W, C] (C % 16 == 0, C <= 128); w_packed: int8
// [O_pad, K_pad] with k = (ky * kw + kx) * C + c, O_pad = O rounded up to 16, K_pad = kh * kw * C rounded up to 128, zero padded
// (the [O, kh, kw, C] codes of the CUDA-core kernel, flattened and padded).  Stride 1.  Outputs as qvit_ultra_conv_bn_act.
static int conv_tc_launch(const void* in_codes, int a_signed, int B, int H, int W, int C, const int8_t* w_packed, int O, int kh, int kw,
                          int sh, int sw, int pad, int dh, int dw, float acc_scale, const float* scale_a, const float* scale_w,
                          const float* bn_scale, const float* bn_bias, int out_levels, int pool, uint8_t* out_codes, float* out_f32,
                          qvit_stream_t stream) {
  QVIT_REQUIRE(in_codes && w_packed && (out_codes || out_f32), "qvit_ultra_conv_tc: null pointer");
  QVIT_REQUIRE(B > 0 && H > 0 && W > 0 && O > 0 && kh > 0 && kw > 0 && pad >= 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0,
               "qvit_ultra_conv_tc: bad geometry");
  QVIT_REQUIRE(C == 16 || C == 32 || C == 64 || C == 128, "qvit_ultra_conv_tc: C must be 16, 32, 64 or 128 (got %d)", C);
  QVIT_REQUIRE(O <= 256, "qvit_ultra_conv_tc: O <= 256");
  QVIT_REQUIRE((reinterpret_cast<uintptr_t>(in_codes) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
               "qvit_ultra_conv_tc: 16-byte aligned operands");
  ctc::Params p;
  p.in = reinterpret_cast<const uint8_t*>(in_codes); p.wpk = w_packed; p.bn_scale = bn_scale;
  p.scale_a = scale_a; p.scale_w = scale_w; p.sh = sh; p.sw = sw; p.dh = dh; p.dw = dw; p.a_signed = a_signed; p.bn_bias = bn_bias; p.out_codes = out_codes; p.out_f32 = out_f32;
  p.B = B; p.H = H; p.W = W; p.C = C; p.O = O; p.O_pad = (O + 15) / 16 * 16; p.kh = kh; p.kw = kw; p.pad = pad;
  p.K = kh * kw * C; p.K_pad = (p.K + 127) / 128 * 128;
  p.OH = (H + 2 * pad - dh * (kh - 1) - 1) / sh + 1; p.OW = (W + 2 * pad - dw * (kw - 1) - 1) / sw + 1;
  QVIT_REQUIRE(p.OH > 0 && p.OW > 0, "qvit_ultra_conv_tc: empty output");
  QVIT_REQUIRE(!pool || (out_codes && !out_f32), "qvit_ultra_conv_tc: pooling applies to the code output only");
  QVIT_REQUIRE(out_f32 || (out_levels >= 1 && out_levels <= 255), "qvit_ultra_conv_tc: out_levels in [1,255]");
  p.pool = pool; p.out_levels = out_levels; p.acc_scale = acc_scale;
  p.linear = pool ? 0 : 1;
  if (p.linear) {
    p.tiles_x = 1;
    p.tiles_per_img = (p.OH * p.OW + ctc::kTileM - 1) / ctc::kTileM;
  } else {
    p.tiles_x = (p.OW + ctc::kTW - 1) / ctc::kTW;
    p.tiles_per_img = p.tiles_x * ((p.OH + ctc::kTH - 1) / ctc::kTH);
  }
  const int64_t total = (int64_t)B * p.tiles_per_img;
  QVIT_REQUIRE(total < (1ll << 31), "qvit_ultra_conv_tc: grid too large");
  p.total_tiles = (int)total;
  p.tmem_cols = p.O_pad <= 32 ? 32 : (p.O_pad <= 64 ? 64 : (p.O_pad <= 128 ? 128 : 256));
  const int chunks = p.K_pad / 128;
  const size_t smem = (size_t)chunks * (p.O_pad + ctc::kTileM) * 128 + 64 + 1024;
  QVIT_REQUIRE(smem <= 227 * 1024, "qvit_ultra_conv_tc: layer too large for the resident-weight kernel (%zu B of shared memory)", smem);
  int dev = 0, maj = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  if (maj != 10) {
    set_error("qvit_ultra_conv_tc: needs sm_100 (tcgen05)");
    return QVIT_ERR_UNSUPPORTED;
  }
  static bool attr_set[64] = {false};
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(ultra_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("qvit_ultra_conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return QVIT_ERR_CUDA;
    }
    attr_set[dev] = true;
  }
  const int grid = (int)(total < sm_count() ? total : sm_count());
  ultra_conv_tc_kernel<<<grid, ctc::kThreads, smem, (cudaStream_t)stream>>>(p);
  return check_launch("qvit_ultra_conv_tc");
}

extern "C" int qvit_ultra_conv_tc(const uint8_t* in_codes, int B, int H, int W, int C, const int8_t* w_packed, int O, int kh, int kw,
                                  int pad, float acc_scale, const float* bn_scale, const float* bn_bias, int out_levels, int pool,
                                  uint8_t* out_codes, float* out_f32, qvit_stream_t stream) {
  return conv_tc_launch(in_codes, 0, B, H, W, C, w_packed, O, kh, kw, 1, 1, pad, 1, 1, acc_scale, nullptr, nullptr, bn_scale, bn_bias,
                        out_levels, pool, out_codes, out_f32, stream);
}

// QuantizeConv2d.forward (quant_layers.py:575-587) as the same implicit GEMM: SIGNED int8 activation codes in NHWC [B, H, W, C]
// (C in {16, 32, 64, 128}), weights packed as above, any stride / padding / dilation (symmetric padding, groups == 1);
// y[b, o, oy, ox] = acc * |d_a| * |d_w| + bias[o] as fp32 NCHW.  No im2col matrix is materialised.
extern "C" int qvit_conv2d_i8_tc(const int8_t* a_codes_nhwc, int B, int H, int W, int C, const int8_t* w_packed, int O, int kh, int kw,
                                 int sh, int sw, int pad, int dh, int dw, const float* scale_a, const float* scale_w, const float* bias,
                                 float* out_nchw, qvit_stream_t stream) {
  QVIT_REQUIRE(out_nchw != nullptr, "qvit_conv2d_i8_tc: null output");
  return conv_tc_launch(a_codes_nhwc, 1, B, H, W, C, w_packed, O, kh, kw, sh, sw, pad, dh, dw, 1.0f, scale_a, scale_w, nullptr, bias, 1, 0,
                        nullptr, out_nchw, stream);
}
