// softmax(Q K^T * scale) V in fp32-equivalent precision on the tcgen05 tensor cores - "2 x FP16" split, pipelined.
//
// Glue BESIDE the hot path (ViTAttention.forward, vit_model.py:133-149): the reference keeps the
// attention core in fp32 and the 4-bit `proj` quantizer behind it turns 1e-5-level errors into flipped codes, so the
// operands must carry (nearly) all fp32 bits.  The producer - the qkv GEMM epilogue (QVIT_OUT_F16X2, gemm_tc.cu) - writes
// every value as TWO fp16 numbers x * 2^s = hi + lo (11 + 11 significant bits; the power of two 2^s, derived from a static
// bound of |x|, keeps hi below 65504 and lo out of the subnormal range), in the bytes one fp32 would take.  Each product is
// then hi*hi' + hi*lo' + lo*hi' (dropped: lo*lo' <= 2^-22 relative) with fp32 accumulation in TMEM: THREE tensor-core terms
// instead of the six of the 3 x bf16 split (attention.cu), no conversion work in this kernel at all, and the operand
// tiles arrive by TMA straight in the layout the MMAs read (K-major 128B-swizzled Q / K tiles, V as an MN-major B operand:
// no transposition).  Measured on the reference's own qkv values (tools/att_split_study.py): max-norm error 6e-7 (the fp32
// CPU reference itself: 6-7e-7 against float64), 0 `proj`-input code flips of 302 592.
//
// Structure: persistent CTA per SM, work unit = (batch, head) with its <= 2 tiles of 128 queries (T <= 208 keys).
//   control warp   one thread: TMA loads (Q double-buffered per tile, K single, V double-buffered per pair) and all MMAs
//   16 compute warps (thread = query row, four warps share a row's key columns): softmax of tile t, then the epilogue
//                  of tile t-1 - while the tensor core runs P V of tile t and S = Q K^T of tile t+2
// TMEM (512 columns): S/P buffer 0 [0,208), S/P buffer 1 [208,416), O accumulator [416,480).  The probabilities go back
// into their S buffer as two packed-fp16 planes (A operand of P V, TS mode).
#include <cuda.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace qvit {

namespace a2 {
constexpr int kHd = 64;               // head dim
constexpr int kMQ = 128;              // queries per tile
constexpr int kNK = 208;              // keys per unit (13 x 16)
constexpr int kComputeWarps = 16;
constexpr int kThreads = 32 * (kComputeWarps + 1);
constexpr int kColQ = kNK / 4;        // 52 key columns per thread in the softmax
constexpr int kQPlane = kMQ * 128;    // 16 KiB
constexpr int kKVPlane = kNK * 128;   // 26 KiB
constexpr int kOffQ = 0;                                  // [slot 2][plane 2]
constexpr int kOffK = kOffQ + 4 * kQPlane;                // [plane 2]
constexpr int kOffV = kOffK + 2 * kKVPlane;               // [slot 2][plane 2]
constexpr int kOffStat = kOffV + 4 * kKVPlane;            // red_max [4][128], red_sum [4][128]
constexpr int kOffBar = kOffStat + 2 * 4 * kMQ * 4;
constexpr int kOffQuant = kOffBar + 256;                   // consumer quantizer constants (SymParams, FastQ2)
constexpr int kSmem = kOffQuant + 256 + 1024;
constexpr int kPCols = kNK / 2;       // TMEM columns of one packed-fp16 P plane (104)
constexpr int kSB1 = 208, kOCol = 416;
constexpr bool kPolyExp = false;      // experiment (measured: exp phase 2430 -> 3050 cycles per tile - the phase is issue-bound, not MUFU-bound): off
constexpr float kPScaleLog2 = 10.0f;  // probabilities are carried as p * 2^10 (keeps the lo plane of small p normal)

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, bool b_mn_major) {
  return (1u << 4) /*D = f32*/ | (0u << 7) /*A = f16*/ | (0u << 10) /*B = f16*/ | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
using ptx::mma_f16_ss;
using ptx::mma_f16_ts;
using ptx::tma_load_3d;
using ptx::tmem_ld16;
using ptx::tmem_ld4;
using ptx::tmem_st;
using ptx::tmem_st_wait;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a packed pair on the FMA / ALU pipes (no MUFU): x = n + f with n = rint(x), f in [-0.5, 0.5]; 2^f by a degree-6
// polynomial (max relative error 1.0e-7 in fp32 Horner form - ex2.approx: 1.7e-7), 2^n by adding n to the exponent field.
// The softmax evaluates every other pair this way: the MUFU pipe (4 lanes per clock and sub-partition, 8 cycles per warp
// instruction) is what bounds the exp phase, and the FMA pipe is nearly idle there.
__device__ __forceinline__ f32x2 exp2_poly2(f32x2 x) {
  float x0, x1;
  unpk2(x, x0, x1);
  const f32x2 xc = pk2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));   // (masked keys arrive as -inf: 2^-125 rounds to 0 in both fp16 planes)
  const f32x2 t = add2(xc, pk1(kRoundMagic));                     // 1.5 * 2^23 + rint(x): the integer sits in the low mantissa bits
  const f32x2 f = add2(xc, fma2(t, pk1(-1.0f), pk1(kRoundMagic))); // x - rint(x)
  f32x2 p = fma2(pk1(0.00015461444854736328f), f, pk1(0.0013400427997112274f));
  p = fma2(p, f, pk1(0.009618056938052177f));
  p = fma2(p, f, pk1(0.05550327152013779f));
  p = fma2(p, f, pk1(0.24022650718688965f));
  p = fma2(p, f, pk1(0.6931471824645996f));
  p = fma2(p, f, pk1(1.0f));
  float p0, p1, t0, t1;
  unpk2(p, p0, p1);
  unpk2(t, t0, t1);
  return pk2(__uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23)), __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23)));
}
// two fp32 values -> packed fp16 hi pair and lo pair (hi + lo = value to 22 significant bits; the residual is exact)
__device__ __forceinline__ void split2_pair(f32x2 v, uint32_t& hi, uint32_t& lo) {
  float a, b;
  unpk2(v, a, b);
  const __half2 h2 = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h2);
  float ra, rb;
  unpk2(fma2(pk2(hf.x, hf.y), pk1(-1.0f), v), ra, rb);
  const __half2 l2 = __floats2half2_rn(ra, rb);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}
}  // namespace a2

// planes: fp16 [B, T, ld] with hi in columns [0, 3*H*64) and lo in [plane_off, plane_off + 3*H*64) (part-major: q | k | v,
// head-major inside a part), as the qkv GEMM writes them.  s_scale = softmax scale * log2(e) * 2^-(sq + sk); o_scale = 2^-sv.
// (17 warps are allocated as 20 - registers come in units of four warps - so 96 registers per thread is the ceiling; the
// softmax below is written in column groups so that it fits)
// kLse (training forward, attention_train.cu): also writes the base-2 log-sum-exp of every row of scaled scores,
// lse[(b * H + h) * 256 + t] = max * s_scale + log2(sum 2^(s * s_scale - max * s_scale)), which the backward recomputes P from.
template <bool kLse>
__global__ void __launch_bounds__(a2::kThreads, 1)
attention_f16x2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, float* __restrict__ out,
                       int T, int H, int plane_off, int total_pairs, float s_scale, float o_scale, int8_t* __restrict__ codes,
                       int64_t ld_codes, const float* __restrict__ q_d, const float* __restrict__ q_qm, const float* __restrict__ q_t,
                       int32_t* __restrict__ q_flags, long long* __restrict__ prof, float* __restrict__ lse) {
  using namespace a2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  float* red_max = reinterpret_cast<float*>(gen + kOffStat);     // [4][128]
  float* red_sum = red_max + 4 * kMQ;
  const uint32_t bar0 = base + kOffBar;
  auto q_full = [&](int s) { return bar0 + 8u * s; };
  const uint32_t k_full = bar0 + 16u;
  auto v_full = [&](int s) { return bar0 + 24u + 8u * s; };
  auto s_done = [&](int b) { return bar0 + 40u + 8u * b; };
  auto p_ready = [&](int b) { return bar0 + 56u + 8u * b; };
  const uint32_t o_done = bar0 + 72u, o_free = bar0 + 80u;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + kOffBar + 128);

  // warp index / TMEM base as values the compiler can SEE are warp-uniform (shuffle from lane 0): the control warp below then
  // keeps descriptors in uniform registers and issues each tcgen05.mma as one instruction.  (Under a plain `lane == 0` guard
  // every MMA sat in an ELECT / R2UR.BROADCAST loop; tools/ubench/mma_rate.cu: 134 -> 83 cycles per issued MMA, and the 39
  // N = 64 P V MMAs of a tile are issue-bound.)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int D = H * kHd;
  const int q_tiles = (T + kMQ - 1) / kMQ;
  const int n_pairs = ((int)blockIdx.x < total_pairs) ? (total_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int nt = n_pairs * q_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(q_full(i), 1);
      ptx::mbar_init(v_full(i), 1);
      ptx::mbar_init(s_done(i), 1);
      ptx::mbar_init(p_ready(i), kComputeWarps);
    }
    ptx::mbar_init(k_full, 1);
    ptx::mbar_init(o_done, 1);
    ptx::mbar_init(o_free, kComputeWarps);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    ptx::tmem_relinquish<1>();
  }
  if (threadIdx.x == 64 && codes) {
    const SymParams qp = load_sym_params(q_d, q_qm, q_t);
    *reinterpret_cast<SymParams*>(gen + kOffQuant) = qp;
    *reinterpret_cast<FastQ2*>(gen + kOffQuant + 64) = make_fastq2(qp);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == kComputeWarps) {
    // ------------------------------------------------------------------------------------------ control warp (converged; one
    // elected lane issues the copies and the MMAs)
    if (nt > 0) {
      if (lane == 0) {
        ptx::prefetch_tmap(&tm_q);
        ptx::prefetch_tmap(&tm_kv);
      }
      auto pair_idx = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
      auto load_q = [&](int t) {
        const int pr = pair_idx(t / q_tiles), b = pr / H, h = pr % H, qt = t % q_tiles, s = t & 1;
        const uint32_t dst = base + kOffQ + (uint32_t)(s * 2 * kQPlane);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(q_full(s), 2u * kQPlane);
          tma_load_3d(dst, &tm_q, q_full(s), h * kHd, qt * kMQ, b);
          tma_load_3d(dst + kQPlane, &tm_q, q_full(s), plane_off + h * kHd, qt * kMQ, b);
        }
        __syncwarp();
      };
      auto load_k = [&](int i) {
        const int pr = pair_idx(i), b = pr / H, h = pr % H;
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(k_full, 2u * kKVPlane);
          tma_load_3d(base + kOffK, &tm_kv, k_full, D + h * kHd, 0, b);
          tma_load_3d(base + kOffK + kKVPlane, &tm_kv, k_full, plane_off + D + h * kHd, 0, b);
        }
        __syncwarp();
      };
      auto load_v = [&](int i) {
        const int pr = pair_idx(i), b = pr / H, h = pr % H, s = i & 1;
        const uint32_t dst = base + kOffV + (uint32_t)(s * 2 * kKVPlane);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(v_full(s), 2u * kKVPlane);
          tma_load_3d(dst, &tm_kv, v_full(s), 2 * D + h * kHd, 0, b);
          tma_load_3d(dst + kKVPlane, &tm_kv, v_full(s), plane_off + 2 * D + h * kHd, 0, b);
        }
        __syncwarp();
      };
      // S_t = Q K^T: cross terms first (small), then hi * hi; 4 k-steps of 16 head-dim values (32 B inside the swizzle atom)
      auto issue_s = [&](int t) {
        const int s = t & 1, i = t / q_tiles;
        ptx::mbar_wait(q_full(s), (uint32_t)((t >> 1) & 1));
        if (t % q_tiles == 0) ptx::mbar_wait(k_full, (uint32_t)(i & 1));
        ptx::tc_fence_after();
        const uint32_t qh = base + kOffQ + (uint32_t)(s * 2 * kQPlane), ql = qh + kQPlane;
        const uint32_t kh = base + kOffK, kl = kh + kKVPlane;
        const uint32_t d = tmem + (uint32_t)(s ? kSB1 : 0);
        constexpr uint32_t idesc = idesc_f16(kMQ, kNK, false);
        if (ptx::elect_one()) {
          uint32_t acc = 0;
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint64_t da = ptx::make_kmajor_sw128_desc(term == 0 ? ql : qh), db = ptx::make_kmajor_sw128_desc(term == 1 ? kl : kh);
#pragma unroll
            for (int ks = 0; ks < kHd / 16; ++ks) {
              mma_f16_ss(d, da + 2 * ks, db + 2 * ks, idesc, acc);
              acc = 1;
            }
          }
          ptx::mma_commit(s_done(s));
        }
        __syncwarp();
      };
      // O_t = P V: A = the packed-fp16 P planes in TMEM, B = the V planes as they lie in memory ([key][head dim], i.e. MN-major:
      // 8-key groups of 1024 B); 13 k-steps of 16 keys, three terms each into the one accumulator
      auto issue_pv = [&](int t) {
        const int s = t & 1, i = t / q_tiles;
        ptx::mbar_wait(p_ready(s), (uint32_t)((t >> 1) & 1));
        if (t % q_tiles == 0) ptx::mbar_wait(v_full(i & 1), (uint32_t)((i >> 1) & 1));
        if (t > 0) ptx::mbar_wait(o_free, (uint32_t)((t - 1) & 1));
        ptx::tc_fence_after();
        const uint32_t vh = base + kOffV + (uint32_t)((i & 1) * 2 * kKVPlane), vl = vh + kKVPlane;
        const uint32_t pa = tmem + (uint32_t)(s ? kSB1 : 0);
        const uint32_t d = tmem + kOCol;
        constexpr uint32_t idesc = idesc_f16(kMQ, kHd, true);
        if (ptx::elect_one()) {
          const uint64_t bh0 = ptx::make_kmajor_sw128_desc(vh), bl0 = ptx::make_kmajor_sw128_desc(vl);
#pragma unroll
          for (int ks = 0; ks < kNK / 16; ++ks) {
            const uint64_t bh = bh0 + (uint64_t)(ks * 128), bl = bl0 + (uint64_t)(ks * 128);
            mma_f16_ts(d, pa + (uint32_t)(kPCols + ks * 8), bh, idesc, ks ? 1u : 0u);   // P_lo V_hi
            mma_f16_ts(d, pa + (uint32_t)(ks * 8), bl, idesc, 1u);                      // P_hi V_lo
            mma_f16_ts(d, pa + (uint32_t)(ks * 8), bh, idesc, 1u);                      // P_hi V_hi
          }
          ptx::mma_commit(o_done);
        }
        __syncwarp();
      };
      auto last_of_pair = [&](int t) { return (t % q_tiles) == q_tiles - 1; };

      // prologue: first operands, the first two score products, and the refills they free
      load_k(0);
      load_q(0);
      if (nt > 1) load_q(1);
      load_v(0);
      if (n_pairs > 1) load_v(1);
      for (int t = 0; t < 2 && t < nt; ++t) {
        issue_s(t);
        ptx::mbar_wait(s_done(t & 1), 0u);
        if (t + 2 < nt) load_q(t + 2);
        if (last_of_pair(t) && t / q_tiles + 1 < n_pairs) load_k(t / q_tiles + 1);
      }
      for (int t = 0; t < nt; ++t) {
        issue_pv(t);
        if (prof && blockIdx.x == 0 && t < 16 && lane == 0) prof[64 + t] = clock64();
        if (t + 2 < nt || last_of_pair(t)) {
          ptx::mbar_wait(o_done, (uint32_t)(t & 1));             // P V of tile t complete: its S/P buffer and (end of pair) V slot are free
          const int i = t / q_tiles;
          if (last_of_pair(t) && i + 2 < n_pairs) load_v(i + 2);
        }
        if (t + 2 < nt) {
          issue_s(t + 2);
          ptx::mbar_wait(s_done(t & 1), (uint32_t)(((t + 2) >> 1) & 1));
          if (t + 4 < nt) load_q(t + 4);
          const int i2 = (t + 2) / q_tiles;
          if (last_of_pair(t + 2) && i2 + 1 < n_pairs) load_k(i2 + 1);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ compute warps
    // the consumer quantizer's constants live in shared memory (written once by thread 0 below, read back in the epilogue):
    // ~20 registers that would otherwise stay live across the softmax
    const SymParams* qp_s = reinterpret_cast<const SymParams*>(gen + kOffQuant);
    const FastQ2* qf_s = reinterpret_cast<const FastQ2*>(gen + kOffQuant + 64);
    int qfl = 0;
    const int row = (warp & 3) * 32 + lane;
    const int cq = warp >> 2;
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
    const int col0 = cq * kColQ;
    float inv_prev = 0.0f;

    auto epilogue = [&](int t, float inv) {
      const int pr = (int)blockIdx.x + (t / q_tiles) * (int)gridDim.x, b = pr / H, h = pr % H, q0 = (t % q_tiles) * kMQ;
      const bool rows_live = q0 + (warp & 3) * 32 < T;
      ptx::mbar_wait(o_done, (uint32_t)(t & 1));
      ptx::tc_fence_after();
      if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[192 + t * 4 + 0] = clock64();
      uint32_t r[16];
      if (rows_live) {
        tmem_ld16(tmem + kOCol + lane_addr + (uint32_t)(cq * 16), r);
        ptx::tmem_ld_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_free);                   // the accumulator may be overwritten by the next P V
      if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[192 + t * 4 + 1] = clock64();
      const int tq = q0 + row;
      if (rows_live && tq < T) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * inv;
        if (out) {
          float* dst = out + ((int64_t)b * T + tq) * ((int64_t)H * kHd) + h * kHd + cq * 16;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            stg_v4_b32(dst + 4 * j, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                       __float_as_uint(v[4 * j + 3]));
        }
        if (codes) {                                             // quantize_act of the consumer layer (quant_layers.py:356-381)
          int8_t* dst = codes + ((int64_t)b * T + tq) * ld_codes + h * kHd + cq * 16;
          const SymParams qp = *qp_s;
          const FastQ2 qf = *qf_s;
          const uint4 w = sym_codes16(v, qp, qf, qfl);
          stg_v4_b32(dst, w.x, w.y, w.z, w.w);
        }
      }
    };

#pragma unroll 1
    for (int t = 0; t < nt; ++t) {
      const int q0 = (t % q_tiles) * kMQ;
      const bool rows_live = q0 + (warp & 3) * 32 < T;
      const uint32_t sb = tmem + (uint32_t)((t & 1) ? kSB1 : 0);
      if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[t * 4 + 0] = clock64();
      ptx::mbar_wait(s_done(t & 1), (uint32_t)((t >> 1) & 1));
      ptx::tc_fence_after();
      if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[t * 4 + 1] = clock64();
      float inv = 0.0f;
      if (rows_live) {
        float p[kColQ];
        {
          uint32_t r[kColQ];
          ptx::tmem_ld_32x32(sb + lane_addr + (uint32_t)col0, reinterpret_cast<uint32_t(&)[32]>(r[0]));
          tmem_ld16(sb + lane_addr + (uint32_t)(col0 + 32), r + 32);
          tmem_ld4(sb + lane_addr + (uint32_t)(col0 + 48), r + 48);
          ptx::tmem_ld_wait();
          if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[128 + t * 4 + 0] = clock64();
#pragma unroll
          for (int j = 0; j < kColQ; ++j) p[j] = __uint_as_float(r[j]);
          if (col0 + kColQ > T) {                                // only the last column quarter holds padding keys
#pragma unroll
            for (int j = 0; j < kColQ; ++j)
              if (col0 + j >= T) p[j] = -INFINITY;               // masked keys: exp2(-inf) = 0
          }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kColQ; ++j) mx = fmaxf(mx, p[j]);
        red_max[cq * kMQ + row] = mx;
        ptx::tc_fence_before();
        // the four warps of a lane quarter share their rows' statistics and TMEM lanes: a 128-thread named barrier is enough
        // (it also orders "all of them have READ S" before the P planes overwrite those columns)
        asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp & 3)) : "memory");
        ptx::tc_fence_after();
        if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[128 + t * 4 + 1] = clock64();
        mx = fmaxf(fmaxf(red_max[row], red_max[kMQ + row]), fmaxf(red_max[2 * kMQ + row], red_max[3 * kMQ + row]));
        // p' = 2^10 * exp((s - max) * scale): the exponent offset is folded into the constant term
        const f32x2 sc2 = pk1(s_scale), nb2 = pk1(kPScaleLog2 - mx * s_scale);
        f32x2 sum2 = pk1(0.0f);
        // probabilities -> two packed-fp16 planes, written back over the scores: plane q occupies TMEM columns [q * 104, +104) of
        // the buffer and this thread owns 26 of them.  Done in column groups of 32 / 16 / 4 scores (16 / 8 / 2 packed words per
        // plane = one tcgen05.st each) so that only one group's words are live at a time.
        const uint32_t cbase = sb + lane_addr + (uint32_t)(cq * 26);
        auto group = [&](auto n_words, int first_pair) {
          constexpr int NW = decltype(n_words)::value;
          uint32_t w1[NW], w2[NW];
#pragma unroll
          for (int j = 0; j < NW; ++j) {
            const f32x2 arg = fma2(pk2(p[2 * (first_pair + j)], p[2 * (first_pair + j) + 1]), sc2, nb2);   // <= 10
            f32x2 e;
            if ((j & 1) && kPolyExp) {
              e = exp2_poly2(arg);                                 // FMA / ALU pipes
            } else {
              float a0, a1;
              unpk2(arg, a0, a1);
              e = pk2(ex2_approx(a0), ex2_approx(a1));             // MUFU, <= 2 ulp
            }
            sum2 = add2(sum2, e);
            split2_pair(e, w1[j], w2[j]);
          }
          tmem_st<NW>(cbase + (uint32_t)first_pair, w1);
          tmem_st<NW>(cbase + (uint32_t)(kPCols + first_pair), w2);
        };
        group(std::integral_constant<int, 16>{}, 0);
        group(std::integral_constant<int, 8>{}, 16);
        group(std::integral_constant<int, 2>{}, 24);
        tmem_st_wait();
        if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[128 + t * 4 + 2] = clock64();
        float sum, sum_hi;
        unpk2(sum2, sum, sum_hi);
        red_sum[cq * kMQ + row] = sum + sum_hi;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp & 3)) : "memory");
        if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[128 + t * 4 + 3] = clock64();
        const float rs = (red_sum[row] + red_sum[kMQ + row]) + (red_sum[2 * kMQ + row] + red_sum[3 * kMQ + row]);
        inv = __fdiv_rn(o_scale, rs);                            // O' / sum(p') * 2^-sv  (the 2^10 of p' cancels)
        if (kLse && cq == 0 && q0 + row < T)
          lse[((int64_t)blockIdx.x + (int64_t)(t / q_tiles) * gridDim.x) * 256 + q0 + row] = mx * s_scale + (log2f(rs) - kPScaleLog2);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_ready(t & 1));
      if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[t * 4 + 2] = clock64();
      if (t > 0) epilogue(t - 1, inv_prev);
      if (prof && blockIdx.x == 0 && threadIdx.x == 0 && t < 16) prof[t * 4 + 3] = clock64();
      inv_prev = inv;
    }
    if (nt > 0) epilogue(nt - 1, inv_prev);
    if (codes) {
      qfl = warp_or(qfl);
      if (qfl && q_flags && lane == 0) atomicOr(q_flags, qfl);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem, 512);
  }
}

}  // namespace qvit

using namespace qvit;

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 encode_fn2() {
  static EncodeTiledFn2 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

// fp16 [B, T, ld] viewed as a 3-D tensor (column, token, batch); box = [64 columns (one head, 128 B) x rows x 1], 128B swizzle;
// tokens beyond T are zero filled per batch element
static int make_tmap_planes(CUtensorMap* map, const void* base, int B, int T, int64_t ld, int cols, int box_rows) {
  EncodeTiledFn2 enc = encode_fn2();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return QVIT_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)(ld * 2), (cuuint64_t)((int64_t)T * ld * 2)};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(planes) failed (CUresult %d) B=%d T=%d ld=%lld", (int)r, B, T, (long long)ld);
    return QVIT_ERR_CUDA;
  }
  return QVIT_OK;
}

// softmax(q k^T * scale) v from the two-plane fp16 form of qkv.  planes: fp16 [B, T, ld], hi plane in columns [0, 3*H*64), lo
// plane in [plane_off, plane_off + 3*H*64); exp_q / exp_k / exp_v: the powers of two the producer multiplied q / k / v by.
namespace qvit {
int attention_f16x2_launch(const void* planes, int64_t ld, int plane_off, int B, int T, int H, int head_dim, float scale, int exp_q,
                           int exp_k, int exp_v, const float* d, const float* q_m, const float* t, int8_t* codes, int64_t ld_codes,
                           float* out, int32_t* flags, long long* prof, float* lse, cudaStream_t stream) {
  QVIT_REQUIRE(planes && (out || codes) && B > 0 && T > 0 && H > 0, "qvit_attention_f16x2: bad argument");
  QVIT_REQUIRE(!codes || (d && q_m && ld_codes >= (int64_t)H * head_dim && (ld_codes & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(codes) & 15) == 0),
               "qvit_attention_f16x2: codes need 16-byte alignment, a pitch >= H * head_dim that is a multiple of 16, and d / q_m");
  if (head_dim != a2::kHd || T > a2::kNK) {
    set_error("qvit_attention_f16x2: supports head_dim == 64 and T <= 208 (got head_dim=%d, T=%d)", head_dim, T);
    return QVIT_ERR_UNSUPPORTED;
  }
  const int D3 = 3 * H * head_dim;
  QVIT_REQUIRE((reinterpret_cast<uintptr_t>(planes) & 15) == 0 && (ld & 7) == 0 && (plane_off & 7) == 0 && plane_off >= D3 &&
                   ld >= plane_off + D3 && (!out || (reinterpret_cast<uintptr_t>(out) & 15) == 0),
               "qvit_attention_f16x2: planes must be 16-byte aligned with ld, plane_off multiples of 8 and ld >= plane_off + 3*H*64");
  int dev = 0, maj = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  if (maj != 10) {
    set_error("qvit_attention_f16x2: needs sm_100 (tcgen05)");
    return QVIT_ERR_UNSUPPORTED;
  }
  static bool attr_set[64] = {false};
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attention_f16x2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, a2::kSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_f16x2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, a2::kSmem);
    if (e != cudaSuccess) {
      set_error("qvit_attention_f16x2: cudaFuncSetAttribute(%d): %s", a2::kSmem, cudaGetErrorString(e));
      return QVIT_ERR_CUDA;
    }
    attr_set[dev] = true;
  }
  CUtensorMap tm_q, tm_kv;
  int rc = make_tmap_planes(&tm_q, planes, B, T, ld, plane_off + D3, a2::kMQ);
  if (rc) return rc;
  rc = make_tmap_planes(&tm_kv, planes, B, T, ld, plane_off + D3, a2::kNK);
  if (rc) return rc;
  const int64_t total_pairs = (int64_t)H * B;
  QVIT_REQUIRE(total_pairs < (1ll << 30), "qvit_attention_f16x2: problem too large");
  const int grid = (int)(total_pairs < sm_count() ? total_pairs : sm_count());
  const float s_scale = ldexpf(scale * 1.4426950408889634f, -(exp_q + exp_k));
  const float o_scale = ldexpf(1.0f, -exp_v);
  if (lse)
    attention_f16x2_kernel<true><<<grid, a2::kThreads, a2::kSmem, stream>>>(tm_q, tm_kv, out, T, H, plane_off, (int)total_pairs, s_scale,
                                                                              o_scale, codes, ld_codes, d, q_m, t, flags, prof, lse);
  else
    attention_f16x2_kernel<false><<<grid, a2::kThreads, a2::kSmem, stream>>>(tm_q, tm_kv, out, T, H, plane_off, (int)total_pairs, s_scale,
                                                                               o_scale, codes, ld_codes, d, q_m, t, flags, prof, nullptr);
  return check_launch("qvit_attention_f16x2");
}
}  // namespace qvit

extern "C" int qvit_attention_f16x2(const void* planes, int64_t ld, int plane_off, int B, int T, int H, int head_dim, float scale,
                                    int exp_q, int exp_k, int exp_v, const float* d, const float* q_m, const float* t, int8_t* codes,
                                    int64_t ld_codes, float* out, int32_t* flags, long long* prof, qvit_stream_t stream) {
  return attention_f16x2_launch(planes, ld, plane_off, B, T, H, head_dim, scale, exp_q, exp_k, exp_v, d, q_m, t, codes, ld_codes, out,
                                flags, prof, nullptr, (cudaStream_t)stream);
}

// fp32 -> the two-plane fp16 form (x * 2^e = hi + lo), for callers that hold qkv in fp32: out fp16 [rows, ld], hi in columns
// [0, cols), lo in [plane_off, plane_off + cols); col_exp: one power of two per column (NULL = no scaling).
namespace qvit {
__global__ void split2_f16_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ldx, const int* __restrict__ col_exp,
                                  __half* __restrict__ out, int64_t ld, int plane_off, int32_t* __restrict__ flags) {
  const int64_t n = rows * (int64_t)(cols / 2);
  int fl = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (cols / 2);
    const int c = (int)(i - r * (cols / 2)) * 2;
    const float2 v = *reinterpret_cast<const float2*>(x + r * ldx + c);
    const float a = col_exp ? ldexpf(v.x, col_exp[c]) : v.x, b = col_exp ? ldexpf(v.y, col_exp[c + 1]) : v.y;
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(a - hf.x, b - hf.y);
    if (!(fabsf(hf.x) <= 65504.0f) || !(fabsf(hf.y) <= 65504.0f)) fl |= kFlagOverflow;
    *reinterpret_cast<__half2*>(out + r * ld + c) = h2;
    *reinterpret_cast<__half2*>(out + r * ld + plane_off + c) = l2;
  }
  fl = warp_or(fl);
  if (fl && flags && (threadIdx.x & 31) == 0) atomicOr(flags, fl);
}
}  // namespace qvit

namespace qvit {
// 8 values per thread (two 16-byte loads, one 16-byte store per plane); needs cols, pitches and plane_off multiples of 8 and
// 16-byte aligned bases
__global__ void __launch_bounds__(256) split2_f16_v8_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ldx,
                                                             const int* __restrict__ col_exp, __half* __restrict__ out, int64_t ld,
                                                             int plane_off, int32_t* __restrict__ flags) {
  const int pieces = cols / 8;
  const int64_t n = rows * (int64_t)pieces;
  int fl = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / pieces;
    const int c = (int)(i - r * pieces) * 8;
    const float4 a = ldg_stream4(x + r * ldx + c), b = ldg_stream4(x + r * ldx + c + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (col_exp) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ldexpf(v[j], col_exp[c + j]);
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 h2 = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      const float2 hf = __half22float2(h2);
      const __half2 l2 = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
      if (!(fabsf(hf.x) <= 65504.0f) || !(fabsf(hf.y) <= 65504.0f)) fl |= kFlagOverflow;
      hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    stg_v4_b32(out + r * ld + c, hi[0], hi[1], hi[2], hi[3]);
    stg_v4_b32(out + r * ld + plane_off + c, lo[0], lo[1], lo[2], lo[3]);
  }
  fl = warp_or(fl);
  if (fl && flags && (threadIdx.x & 31) == 0) atomicOr(flags, fl);
}
}  // namespace qvit

extern "C" int qvit_split2_f16(const float* x, int64_t rows, int cols, int64_t ldx, const int* col_exp, void* out, int64_t ld,
                               int plane_off, int32_t* flags, qvit_stream_t stream) {
  QVIT_REQUIRE(x && out && rows >= 0 && cols > 0 && (cols & 1) == 0 && (ldx & 1) == 0 && (ld & 1) == 0 && (plane_off & 1) == 0 &&
                   plane_off >= cols && ld >= plane_off + cols,
               "qvit_split2_f16: bad argument (cols, pitches and plane_off must be even, ld >= plane_off + cols)");
  if (rows == 0) return QVIT_OK;
  if ((cols & 7) == 0 && (ldx & 7) == 0 && (ld & 7) == 0 && (plane_off & 7) == 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int64_t n8 = rows * (int64_t)(cols / 8);
    const int blocks8 = (int)((n8 + 255) / 256 < 16 * 148 ? (n8 + 255) / 256 : 16 * 148);
    split2_f16_v8_kernel<<<blocks8, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ldx, col_exp, reinterpret_cast<__half*>(out), ld, plane_off,
                                                                   flags);
    return check_launch("qvit_split2_f16");
  }
  const int64_t n = rows * (int64_t)(cols / 2);
  const int blocks = (int)((n + 255) / 256 < 8 * 148 ? (n + 255) / 256 : 8 * 148);
  split2_f16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ldx, col_exp, reinterpret_cast<__half*>(out), ld, plane_off, flags);
  return check_launch("qvit_split2_f16");
}
