// softmax(Q K^T * scale) V in fp32-equivalent precision on the tcgen05 tensor cores ("3 x BF16" split).
//
// This is glue BESIDE the hot path: the reference never quantizes the attention core
// (vit_model.py:141-149) and runs it in fp32, and the 4-bit quantizer that follows (`proj`) turns any 1e-5-level
// error of its input into flipped activation codes, so bf16 / single-pass TF32 attention is not usable.
// Instead every fp32 operand is split EXACTLY into three bf16 numbers x = b1 + b2 + b3 (8 + 8 + 8 mantissa bits)
// and each product is evaluated as b1*b1' + b1*b2' + b2*b1' + b2*b2' + b1*b3' + b3*b1' with fp32 accumulation in
// TMEM (dropped terms <= 2^-24 relative): fp32-level accuracy at tensor-core speed, same MMA count as 3xTF32.
//
// One work unit = one (batch, head): its K planes [208 x 64] and V^T planes [64 x 208] stay resident in shared memory
// for both tiles of 128 queries (T <= 208).  All conversions read the fp32 qkv matrix in place (global loads issued ahead
// of a running MMA batch) and write K-major, 128-byte-swizzled bf16 planes.  Per query tile:
//   phase A  S = Q K^T: 4 k-steps x 6 terms of tcgen05.mma kind::f16 (M=128, N=208, K=16), accumulator in TMEM
//            || V -> three TRANSPOSED bf16 planes V^T (first tile of the pair)
//   phase B  softmax over the row held by each thread (thread = query row; four warps share a row's columns);
//            P = exp2((S - max) * scale * log2e) goes back to TMEM as three packed-bf16 planes (A operand of phase C)
//   phase C  O = P V: 13 k-steps x 6 terms, A from TMEM, B = V^T planes stacked to N = 192 / 128 / 64
//            || Q planes of the next tile (and K planes of the next pair)
//   phase D  O / rowsum -> (quantize ->) global [B, T, H*64]
// TMEM: S 208 columns, overwritten by P (3 x 104), O partials 192 columns.  Shared memory: 222 KiB of planes.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace qvit {

constexpr int kAttHd = 64;            // head dim
constexpr int kAttMQ = 128;           // queries per CTA
constexpr int kAttNK = 208;           // keys per CTA (13 x 16): UMMA N of the S tile, K extent of the PV product
constexpr int kAttThreads = 512;      // 16 warps: four per TMEM lane quarter, each owning a quarter (52) of the key columns
constexpr int kColQ = kAttNK / 4;     // 52 key columns per thread in the softmax
constexpr int kQPlane = kAttMQ * 128;             // bf16 Q plane   [128 rows x 64 bf16]            16 KiB
constexpr int kKVPlane = 4 * kAttHd * 128;        // bf16 V^T plane: 4 sub-tiles [64 head-dim rows x 64 keys]        32 KiB
constexpr int kVtSub = kAttHd * 128;               // one V^T sub-tile of one plane: [64 head-dim rows x 64 keys] 8 KiB
constexpr int kKPlane = kAttNK * 128;             // bf16 K plane [208 key rows x 64 bf16]             26 KiB
constexpr int kAttSmem = 3 * kQPlane + 3 * kKPlane + 3 * kKVPlane + 128 + 1024;   // 222 KiB of planes + barriers + alignment slack
constexpr int kPCols = kAttNK / 2;                // TMEM columns of one packed-bf16 P plane (104)
constexpr int kOCol = 320;                        // TMEM columns [320, 512): three partial O accumulators of 64 columns

// tcgen05.mma kind::f16 (bf16 inputs, fp32 accumulate), A and B from shared memory
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// A from tensor memory (lane = row, two bf16 k-values per 32-bit column), B from shared memory
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t* r);
template <>
__device__ __forceinline__ void tmem_st_n<32>(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_st_n<16>(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_st_n<4>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_st_n<8>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_st_n<2>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 3-D tile load: coordinates (column float index, token, batch)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) /*D = f32*/ | (1u << 7) /*A = bf16*/ | (1u << 10) /*B = bf16*/ | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exact three-way split of two fp32 values into packed bf16 pairs (low half = first value)
__device__ __forceinline__ uint32_t bf16x2_rn(float lo, float hi) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&a);
}
__device__ __forceinline__ void split3_pair(float x, float y, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  // per level: one F2FP (round both to bf16), two ALU ops to widen them again, one packed subtract (FFMA2, exact)
  const f32x2 mone = pk1(-1.0f);
  p1 = bf16x2_rn(x, y);
  const f32x2 r1 = fma2(pk2(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u)), mone, pk2(x, y));
  float r1x, r1y;
  unpk2(r1, r1x, r1y);
  p2 = bf16x2_rn(r1x, r1y);
  const f32x2 r2 = fma2(pk2(__uint_as_float(p2 << 16), __uint_as_float(p2 & 0xffff0000u)), mone, r1);
  float r2x, r2y;
  unpk2(r2, r2x, r2y);
  p3 = bf16x2_rn(r2x, r2y);                                     // exact remainder, <= 8 significant bits
}

// 8 consecutive fp32 -> one 16-byte chunk (8 bf16) per plane
__device__ __forceinline__ void split3_chunk(const float4& u, const float4& v, uint4& c1, uint4& c2, uint4& c3) {
  split3_pair(u.x, u.y, c1.x, c2.x, c3.x);
  split3_pair(u.z, u.w, c1.y, c2.y, c3.y);
  split3_pair(v.x, v.y, c1.z, c2.z, c3.z);
  split3_pair(v.z, v.w, c1.w, c2.w, c3.w);
}

// the six product terms kept (plane of A, plane of B): all pairs with index sum <= 2 plus (1,1)
__device__ __constant__ const int kTermA[6] = {0, 0, 1, 1, 0, 2};
__device__ __constant__ const int kTermB[6] = {0, 1, 0, 1, 2, 0};

__global__ void __launch_bounds__(kAttThreads, 1)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int H, int total_pairs, float scale_log2e,
                     float* __restrict__ dbg, int dump, int8_t* __restrict__ codes, int64_t ld_codes,
                     const float* __restrict__ q_d, const float* __restrict__ q_qm, const float* __restrict__ q_t,
                     int32_t* __restrict__ q_flags) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  // layout: Q planes (3 x 16 KiB) | K planes (3 x 26 KiB) | V^T planes (3 x 32 KiB) | barriers.  The softmax row statistics
  // (4 KiB) alias the Q planes, which are dead between the S product and the conversion of the next Q tile.
  const uint32_t q_pl = base, k_pl = base + 3 * kQPlane, v_pl = k_pl + 3 * kKPlane, misc_a = v_pl + 3 * kKVPlane;
  uint8_t* g_q = gen;
  uint8_t* g_k = gen + 3 * kQPlane;
  uint8_t* g_v = g_k + 3 * kKPlane;
  uint8_t* misc = g_v + 3 * kKVPlane;
  const uint32_t bar_s = misc_a, bar_o = misc_a + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(misc + 16);
  float* red_max = reinterpret_cast<float*>(g_q);                 // [4][128]
  float* red_sum = red_max + 4 * kAttMQ;                         // [4][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row_floats = 3ll * H * kAttHd;
  const int q_tiles = (T + kAttMQ - 1) / kAttMQ;
  // optional fused activation quantizer of the consumer layer (`proj`): codes instead of / beside the fp32 context
  SymParams qp;
  FastQ2 qf;
  int qfl = 0;
  if (codes) {
    qp = load_sym_params(q_d, q_qm, q_t);
    qf = make_fastq2(qp);
  }

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_s = tmem, t_o = tmem + kOCol;

  // ---- operand conversions, straight from global memory (the qkv matrix is read in place): fp32 -> three exact bf16 planes,
  // K-major with the 128-byte swizzle.  All loads of a pass are issued before the first use: one memory round trip, which
  // the caller hides behind a running MMA batch.
  // Q tile (128 queries x 64) of (bb, hh, qt)
  auto convert_q = [&](int bb, int hh, int qt) {
    constexpr int kIt = kAttMQ * 8 / kAttThreads;
    float4 u[kIt], v[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int item = threadIdx.x + it * kAttThreads;
      const int r = item >> 3, c = item & 7;                   // row of the tile, chunk of 8 head-dim values
      const int t = qt * kAttMQ + r;
      u[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      v[it] = u[it];
      if (t < T) {
        const float* src = qkv + ((int64_t)bb * T + t) * row_floats + (0 * H + hh) * kAttHd + c * 8;
        u[it] = ldg_stream4(src);
        v[it] = ldg_stream4(src + 4);
      }
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int item = threadIdx.x + it * kAttThreads;
      const int r = item >> 3, c = item & 7;
      uint4 c1, c2, c3;
      split3_chunk(u[it], v[it], c1, c2, c3);
      const int off = r * 128 + ((c ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(g_q + off) = c1;
      *reinterpret_cast<uint4*>(g_q + kQPlane + off) = c2;
      *reinterpret_cast<uint4*>(g_q + 2 * kQPlane + off) = c3;
    }
  };
  // K (208 key rows x 64; rows >= T are zero) of (bb, hh): two passes of two pieces per thread
  auto convert_k = [&](int bb, int hh) {
    constexpr int kIt = (kAttNK * 8 + kAttThreads - 1) / kAttThreads;   // 4 (the last one covers 128 threads)
    float4 u[kIt], v[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int item = threadIdx.x + it * kAttThreads;
      const int r = item >> 3, c = item & 7;
      u[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      v[it] = u[it];
      if (item < kAttNK * 8 && r < T) {
        const float* src = qkv + ((int64_t)bb * T + r) * row_floats + (1 * H + hh) * kAttHd + c * 8;
        u[it] = ldg_stream4(src);
        v[it] = ldg_stream4(src + 4);
      }
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int item = threadIdx.x + it * kAttThreads;
      if (item < kAttNK * 8) {
        const int r = item >> 3, c = item & 7;
        uint4 c1, c2, c3;
        split3_chunk(u[it], v[it], c1, c2, c3);
        const int off = r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(g_k + off) = c1;
        *reinterpret_cast<uint4*>(g_k + kKPlane + off) = c2;
        *reinterpret_cast<uint4*>(g_k + 2 * kKPlane + off) = c3;
      }
    }
  };
  // Q tile 0 and K of a NEW pair in one go: the loads of both are in flight together (one round trip instead of two)
  auto convert_qk = [&](int bb, int hh) {
    constexpr int kItQ = kAttMQ * 8 / kAttThreads, kItK = (kAttNK * 8 + kAttThreads - 1) / kAttThreads;
    float4 uq[kItQ], vq[kItQ], uk[kItK], vk[kItK];
#pragma unroll
    for (int it = 0; it < kItK; ++it) {
      const int item = threadIdx.x + it * kAttThreads;
      const int r = item >> 3, c = item & 7;
      uk[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      vk[it] = uk[it];
      if (item < kAttNK * 8 && r < T) {
        const float* src = qkv + ((int64_t)bb * T + r) * row_floats + (1 * H + hh) * kAttHd + c * 8;
        uk[it] = ldg_stream4(src);
        vk[it] = ldg_stream4(src + 4);
      }
    }
#pragma unroll
    for (int it = 0; it < kItQ; ++it) {
      const int item = threadIdx.x + it * kAttThreads;
      const int r = item >> 3, c = item & 7;
      uq[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      vq[it] = uq[it];
      if (r < T) {
        const float* src = qkv + ((int64_t)bb * T + r) * row_floats + (0 * H + hh) * kAttHd + c * 8;
        uq[it] = ldg_stream4(src);
        vq[it] = ldg_stream4(src + 4);
      }
    }
#pragma unroll
    for (int it = 0; it < kItK; ++it) {
      const int item = threadIdx.x + it * kAttThreads;
      if (item < kAttNK * 8) {
        const int r = item >> 3, c = item & 7;
        uint4 c1, c2, c3;
        split3_chunk(uk[it], vk[it], c1, c2, c3);
        const int off = r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(g_k + off) = c1;
        *reinterpret_cast<uint4*>(g_k + kKPlane + off) = c2;
        *reinterpret_cast<uint4*>(g_k + 2 * kKPlane + off) = c3;
      }
    }
#pragma unroll
    for (int it = 0; it < kItQ; ++it) {
      const int item = threadIdx.x + it * kAttThreads;
      const int r = item >> 3, c = item & 7;
      uint4 c1, c2, c3;
      split3_chunk(uq[it], vq[it], c1, c2, c3);
      const int off = r * 128 + ((c ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(g_q + off) = c1;
      *reinterpret_cast<uint4*>(g_q + kQPlane + off) = c2;
      *reinterpret_cast<uint4*>(g_q + 2 * kQPlane + off) = c3;
    }
  };
  // V (keys x 64) of (bb, hh) -> three TRANSPOSED planes V^T [64 head-dim rows x keys].
  // Layout [64-key sub-tile][plane][64 rows x 128 B]: the three planes of a sub-tile are contiguous, so one MMA with
  // N = 192 / 128 / 64 multiplies a P plane with V1|V2|V3, V1|V2 or V1 at once.
  auto convert_vt = [&](int bb, int hh) {
    // one warp item = all 64 head-dim rows x 8 consecutive keys; lane = head-dim pair (2l, 2l+1): every load instruction
    // reads the 256 contiguous bytes of one key row, 16 loads per lane in all (26 items over 16 warps)
    constexpr int kItems = kAttNK / 8, kWarps = kAttThreads / 32, kIt = (kItems + kWarps - 1) / kWarps;
    float2 x[kIt][8];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int kg = warp + it * kWarps;
      const float* src = qkv + ((int64_t)bb * T + 8 * kg) * row_floats + (2 * H + hh) * kAttHd + 2 * lane;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[it][i] = make_float2(0.f, 0.f);
        if (kg < kItems && 8 * kg + i < T) x[it][i] = __ldg(reinterpret_cast<const float2*>(src + i * row_floats));
      }
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int kg = warp + it * kWarps;
      if (kg < kItems) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int hd = 2 * lane + half;
          uint4 c1, c2, c3;
          if (half == 0)
            split3_chunk(make_float4(x[it][0].x, x[it][1].x, x[it][2].x, x[it][3].x),
                         make_float4(x[it][4].x, x[it][5].x, x[it][6].x, x[it][7].x), c1, c2, c3);
          else
            split3_chunk(make_float4(x[it][0].y, x[it][1].y, x[it][2].y, x[it][3].y),
                         make_float4(x[it][4].y, x[it][5].y, x[it][6].y, x[it][7].y), c1, c2, c3);
          const int off = (kg >> 3) * (3 * kVtSub) + hd * 128 + (((kg & 7) ^ (hd & 7)) << 4);
          *reinterpret_cast<uint4*>(g_v + off) = c1;
          *reinterpret_cast<uint4*>(g_v + kVtSub + off) = c2;
          *reinterpret_cast<uint4*>(g_v + 2 * kVtSub + off) = c3;
        }
      }
    }
  };
  // L2 prefetch of the 128-byte lines a later conversion reads (one line per thread): the conversion's loads then see an
  // L2 hit instead of an HBM round trip.  part: 0 = Q tile qt, 1 = K, 2 = V of (bb, hh).
  auto prefetch_rows = [&](int part, int bb, int hh, int qt) {
    const int rows = part == 0 ? kAttMQ : kAttNK;
    const int r = threadIdx.x >> 1, half = threadIdx.x & 1;
    const int t = (part == 0 ? qt * kAttMQ : 0) + r;
    if (r < rows && t < T) {
      const float* src = qkv + ((int64_t)bb * T + t) * row_floats + (part * H + hh) * kAttHd + half * 32;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
    }
  };

  // S = Q K^T  (6 bf16 terms) of the tile whose Q / K planes are in shared memory; completion on bar_s
  auto issue_s = [&]() {
    if (threadIdx.x == 32) {
      ptx::tc_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16(kAttMQ, kAttNK);
      uint32_t acc = 0;
#pragma unroll 1
      for (int term = 0; term < 6; ++term) {
        const uint32_t a_base = q_pl + kTermA[term] * kQPlane;
        const uint32_t b_base = k_pl + kTermB[term] * kKPlane;
#pragma unroll
        for (int ks = 0; ks < kAttHd / 16; ++ks) {             // 16 head-dim values (32 B) per MMA
          mma_bf16_ss(t_s, ptx::make_kmajor_sw128_desc(a_base + ks * 32), ptx::make_kmajor_sw128_desc(b_base + ks * 32), idesc, acc);
          acc = 1;
        }
      }
      ptx::mma_commit(bar_s);
    }
  };

  // Persistent over (batch, head); both query tiles of a pair run back to back on the same K / V^T planes.
  // Per tile:  A  S = Q K^T on the tensor core        || epilogue D of the PREVIOUS tile, V^T conversion (first tile of a pair)
  //            B  softmax, P planes -> TMEM
  //            C  O = P V on the tensor core           || Q (and K) conversion of the NEXT tile / pair
  //            D  normalise, (quantize,) store         (runs under the next tile's S product)
  if ((int)blockIdx.x < total_pairs) {
    convert_q(blockIdx.x / H, blockIdx.x % H, 0);
    convert_k(blockIdx.x / H, blockIdx.x % H);
  }
  ptx::fence_proxy_async_smem();
  __syncthreads();
  if ((int)blockIdx.x < total_pairs) issue_s();
  uint32_t par = 0;                                            // every mbarrier completes exactly once per tile
  int tile_no = 0;
#pragma unroll 1
  for (int pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
  const int h = pair % H, b = pair / H;
#pragma unroll 1
  for (int q_tile = 0; q_tile < q_tiles; ++q_tile, par ^= 1u, ++tile_no) {
  const int q0 = q_tile * kAttMQ;
  long long ts[8];
  const bool prof = dbg && blockIdx.x == 0 && threadIdx.x == 64 && (tile_no == 2 || tile_no == 3);   // both tiles of the CTA's second pair
  if (prof) ts[0] = clock64();

  // ---- phase A: S = Q K^T was issued before the previous tile's epilogue (or in the prologue); V^T conversion underneath
  if (q_tile == 0) convert_vt(b, h);                           // (the previous tile's P V product is complete: phase C waited)
  if (prof) ts[1] = clock64();
  ptx::mbar_wait(bar_s, par);
  ptx::tc_fence_after();
  if (prof) ts[2] = clock64();

  // ---- phase B: softmax.  thread = row (lane quarter = warp & 3), column quarter = warp >> 2 (52 key columns each)
  // (first: pull what phase C will convert into L2)
  if (q_tile + 1 < q_tiles) {
    prefetch_rows(0, b, h, q_tile + 1);
  } else if (pair + (int)gridDim.x < total_pairs) {
    const int np = pair + gridDim.x;
    prefetch_rows(0, np / H, np % H, 0);
    prefetch_rows(1, np / H, np % H, 0);
    prefetch_rows(2, np / H, np % H, 0);
  }
  const int row = (warp & 3) * 32 + lane;
  const int cq = warp >> 2;
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
  const int col0 = cq * kColQ;
  // a lane quarter whose 32 query rows all lie beyond T (the tail of the last tile) has nothing to do in phases B and D:
  // its TMEM rows keep stale P values, which only ever reach output rows that are never stored (MMA rows are independent)
  const bool rows_live = q0 + (warp & 3) * 32 < T;
  float inv = 0.0f, rs0 = 0.0f, rs1 = 0.0f, rs2 = 0.0f, rs3 = 0.0f;
  if (rows_live) {
  float p[kColQ];                                              // this thread's 52 scores, then probabilities
  {
    uint32_t r[kColQ];
    ptx::tmem_ld_32x32(t_s + lane_addr + (uint32_t)col0, reinterpret_cast<uint32_t(&)[32]>(r[0]));
    tmem_ld_32x16(t_s + lane_addr + (uint32_t)(col0 + 32), r + 32);
    tmem_ld_32x4(t_s + lane_addr + (uint32_t)(col0 + 48), r + 48);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < kColQ; ++j) p[j] = __uint_as_float(r[j]);
    if (col0 + kColQ > T) {                                    // only the last column quarter holds padding keys
#pragma unroll
      for (int j = 0; j < kColQ; ++j)
        if (col0 + j >= T) p[j] = -INFINITY;                   // masked keys: exp2(-inf) = 0
    }
  }
  if (dump) {
    float* d = dbg + ((((int64_t)b * H + h) * 256 + q0 + row) * 512) + col0;
#pragma unroll
    for (int j = 0; j < kColQ; ++j) d[j] = p[j];
  }
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kColQ; ++j) mx = fmaxf(mx, p[j]);
  red_max[cq * kAttMQ + row] = mx;
  ptx::tc_fence_before();
  // the four warps of a lane quarter share their rows' statistics and TMEM lanes: a 128-thread named barrier is enough
  // (it also orders "all of them have READ S" before the P planes overwrite those columns)
  asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp & 3)) : "memory");
  ptx::tc_fence_after();
  mx = fmaxf(fmaxf(red_max[row], red_max[kAttMQ + row]), fmaxf(red_max[2 * kAttMQ + row], red_max[3 * kAttMQ + row]));
  const f32x2 sc2 = pk1(scale_log2e), nb2 = pk1(-(mx * scale_log2e));
  f32x2 sum2 = pk1(0.0f);
#pragma unroll
  for (int j = 0; j < kColQ / 2; ++j) {
    float a0, a1;
    unpk2(fma2(pk2(p[2 * j], p[2 * j + 1]), sc2, nb2), a0, a1);
    const float e0 = ex2_approx(a0), e1 = ex2_approx(a1);         // <= 2 ulp, argument <= 0
    sum2 = add2(sum2, pk2(e0, e1));
    p[2 * j] = e0;
    p[2 * j + 1] = e1;
  }
  float sum, sum_hi;
  unpk2(sum2, sum, sum_hi);
  sum += sum_hi;
  if (dump) {
    float* d = dbg + ((((int64_t)b * H + h) * 256 + q0 + row) * 512) + 208 + col0;
#pragma unroll
    for (int j = 0; j < kColQ; ++j) d[j] = p[j];
  }
  red_sum[cq * kAttMQ + row] = sum;
  {
    // three packed-bf16 planes: plane q occupies TMEM columns [q*104, q*104 + 104); this thread owns 26 of them
    uint32_t w1[26], w2[26], w3[26];
#pragma unroll
    for (int j = 0; j < 26; ++j) split3_pair(p[2 * j], p[2 * j + 1], w1[j], w2[j], w3[j]);
    const uint32_t cbase = t_s + lane_addr + (uint32_t)(cq * 26);
    tmem_st_n<16>(cbase, w1);              tmem_st_n<8>(cbase + 16, w1 + 16);              tmem_st_n<2>(cbase + 24, w1 + 24);
    tmem_st_n<16>(cbase + kPCols, w2);     tmem_st_n<8>(cbase + kPCols + 16, w2 + 16);     tmem_st_n<2>(cbase + kPCols + 24, w2 + 24);
    tmem_st_n<16>(cbase + 2 * kPCols, w3); tmem_st_n<8>(cbase + 2 * kPCols + 16, w3 + 16); tmem_st_n<2>(cbase + 2 * kPCols + 24, w3 + 24);
    tmem_st_wait();
  }
  // row sums of the lane quarter are complete after its second barrier; 1 / sum stays in a register (the statistics
  // buffer aliases the Q planes, which phase C overwrites)
  asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp & 3)) : "memory");
  rs0 = red_sum[row], rs1 = red_sum[kAttMQ + row], rs2 = red_sum[2 * kAttMQ + row], rs3 = red_sum[3 * kAttMQ + row];
  inv = __fdiv_rn(1.0f, (rs0 + rs1) + (rs2 + rs3));
  }  // rows_live
  ptx::fence_proxy_async_smem();                               // V^T planes (generic-proxy stores of phase A) -> tensor core
  ptx::tc_fence_before();
  __syncthreads();
  if (prof) ts[3] = clock64();

  // ---- phase C: O = P V, conversions of the next tile underneath
  if (threadIdx.x == 32) {
    // O partials: acc[0:64) += P1 V1 + P2 V1 + P3 V1, acc[64:128) += P1 V2 + P2 V2, acc[128:192) += P1 V3  (6 terms, 3 MMAs per k-step)
    ptx::tc_fence_after();
    uint32_t acc = 0;
#pragma unroll 1
    for (int ks = 0; ks < kAttNK / 16; ++ks) {                 // 16 keys per MMA: 8 TMEM columns of A, 32 B of each V^T row
      const uint64_t b_desc = ptx::make_kmajor_sw128_desc(v_pl + (ks >> 2) * (3 * kVtSub) + (ks & 3) * 32);
      mma_bf16_ts(t_o, t_s + (uint32_t)(ks * 8), b_desc, make_idesc_bf16(kAttMQ, 192), acc);
      mma_bf16_ts(t_o, t_s + (uint32_t)(kPCols + ks * 8), b_desc, make_idesc_bf16(kAttMQ, 128), 1u);
      mma_bf16_ts(t_o, t_s + (uint32_t)(2 * kPCols + ks * 8), b_desc, make_idesc_bf16(kAttMQ, 64), 1u);
      acc = 1;
    }
    ptx::mma_commit(bar_o);
  }
  {
    // the Q planes (and, at the end of a pair, the K planes) are dead since the S product
    if (q_tile + 1 < q_tiles) {
      convert_q(b, h, q_tile + 1);
    } else if (pair + (int)gridDim.x < total_pairs) {
      const int np = pair + gridDim.x;
      convert_qk(np / H, np % H);
    }
  }
  if (prof) ts[4] = clock64();
  ptx::mbar_wait(bar_o, par);
  // the next tile's S product reads the Q / K planes written above (generic proxy -> tensor core) and overwrites the
  // P columns of TMEM; it leaves the O columns alone, so it is issued BEFORE this tile's epilogue
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (q_tile + 1 < q_tiles || pair + (int)gridDim.x < total_pairs) issue_s();
  if (prof) ts[5] = clock64();

  // ---- phase D: normalise and store.  Column quarter cq of each lane quarter takes head-dim [16*cq, 16*cq + 16)
  if (rows_live) {
    uint32_t r[16], r2[16], r3[16];
    tmem_ld_32x16(t_o + lane_addr + (uint32_t)(cq * 16), r);
    tmem_ld_32x16(t_o + lane_addr + (uint32_t)(64 + cq * 16), r2);
    tmem_ld_32x16(t_o + lane_addr + (uint32_t)(128 + cq * 16), r3);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      r[j] = __float_as_uint((__uint_as_float(r3[j]) + __uint_as_float(r2[j])) + __uint_as_float(r[j]));   // small terms first
    if (dump) {
      float* d = dbg + ((((int64_t)b * H + h) * 256 + q0 + row) * 512) + 416;
#pragma unroll
      for (int j = 0; j < 16; ++j) d[cq * 16 + j] = __uint_as_float(r[j]);
      if (cq == 0) { d[64] = rs0 + rs1; d[65] = rs2 + rs3; d[66] = inv; }
    }
    const int t = q0 + row;
    if (t < T) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * inv;
      if (out) {
        float* dst = out + ((int64_t)b * T + t) * ((int64_t)H * kAttHd) + h * kAttHd + cq * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          stg_v4_b32(dst + 4 * j, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                     __float_as_uint(v[4 * j + 3]));
      }
      if (codes) {                                             // quantize_act of the consumer layer (quant_layers.py:356-381)
        int8_t* dst = codes + ((int64_t)b * T + t) * ld_codes + h * kAttHd + cq * 16;
        const uint4 w = sym_codes16(v, qp, qf, qfl);
        stg_v4_b32(dst, w.x, w.y, w.z, w.w);
      }
    }
  }
  if (prof) {
    ts[6] = clock64();
    for (int i = 0; i < 6; ++i) dbg[255 * 512 + 500 + (tile_no - 2) * 8 + i] = (float)(ts[i + 1] - ts[0]);
  }
  }  // query tiles of the pair
  }  // persistent loop over (batch, head)

  if (codes) {
    qfl = warp_or(qfl);
    if (qfl && q_flags && lane == 0) atomicOr(q_flags, qfl);
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

static int attention_launch(const float* qkv, int B, int T, int H, int head_dim, float scale, float* out, float* dbg,
                            int diag, qvit_stream_t stream, int8_t* codes = nullptr, int64_t ld_codes = 0,
                            const float* q_d = nullptr, const float* q_qm = nullptr, const float* q_t = nullptr,
                            int32_t* q_flags = nullptr) {
  QVIT_REQUIRE(qkv && (out || codes) && B > 0 && T > 0 && H > 0, "qvit_attention_f32: bad argument");
  QVIT_REQUIRE(!codes || (q_d && q_qm && ld_codes >= (int64_t)H * head_dim && (ld_codes & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(codes) & 15) == 0),
               "qvit_attention_quantize_sym: codes need 16-byte alignment, a pitch >= H * head_dim that is a multiple of 16, and d / q_m");
  if (head_dim != kAttHd || T > kAttNK) {
    set_error("qvit_attention_f32: supports head_dim == 64 and T <= 208 (got head_dim=%d, T=%d)", head_dim, T);
    return QVIT_ERR_UNSUPPORTED;
  }
  QVIT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "qvit_attention_f32: pointers must be 16-byte aligned");
  int dev = 0, maj = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  if (maj != 10) {
    set_error("qvit_attention_f32: needs sm_100 (tcgen05)");
    return QVIT_ERR_UNSUPPORTED;
  }
  static bool attr_set[64] = {false};
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem);
    if (e != cudaSuccess) {
      set_error("qvit_attention_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return QVIT_ERR_CUDA;
    }
    attr_set[dev] = true;
  }
  const int64_t total_pairs = (int64_t)H * B;                  // work unit: one (batch, head), all of its query tiles
  QVIT_REQUIRE(total_pairs < (1ll << 30), "qvit_attention_f32: problem too large");
  const int grid = (int)(total_pairs < sm_count() ? total_pairs : sm_count());
  attention_f32_kernel<<<grid, kAttThreads, kAttSmem, (cudaStream_t)stream>>>(qkv, out, T, H, (int)total_pairs,
                                                                             scale * 1.4426950408889634f, dbg, (dbg != nullptr && diag == 0) ? 1 : 0,
                                                                             codes, ld_codes, q_d, q_qm, q_t, q_flags);
  return check_launch("qvit_attention_f32");
}

extern "C" int qvit_attention_f32(const float* qkv, int B, int T, int H, int head_dim, float scale, float* out,
                                  qvit_stream_t stream) {
  return attention_launch(qkv, B, T, H, head_dim, scale, out, nullptr, 0, stream);
}

// Attention core with the consumer layer's activation quantizer fused into the epilogue: int8 codes [B*T, ld_codes]
// (columns h * 64 + i) of softmax(QK^T scale) V, optionally the fp32 context as well (out may be NULL).
extern "C" int qvit_attention_quantize_sym(const float* qkv, int B, int T, int H, int head_dim, float scale, const float* d,
                                           const float* q_m, const float* t, int8_t* codes, int64_t ld_codes, float* out,
                                           int32_t* flags, qvit_stream_t stream) {
  QVIT_REQUIRE(codes != nullptr, "qvit_attention_quantize_sym: codes is NULL");
  return attention_launch(qkv, B, T, H, head_dim, scale, out, nullptr, 0, stream, codes, ld_codes, d, q_m, t, flags);
}

// test hook: additionally dumps raw scores S (cols 0..207) and un-normalised probabilities P (cols 208..415) per query row
// into dbg [B, H, 256, 512] fp32 (cols 416..479 raw O, 480..482 row sums / 1/sum)
extern "C" int qvit_attention_f32_debug(const float* qkv, int B, int T, int H, int head_dim, float scale, float* out,
                                        float* dbg, int diag, qvit_stream_t stream) {
  return attention_launch(qkv, B, T, H, head_dim, scale, out, dbg, diag, stream);
}
