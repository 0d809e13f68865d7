// K3: QuantLinear as a dense int8 x int8 -> int32 tensor-core GEMM for sm_100a.
//
//   acc[m, n] = sum_k A[m, k] * W[n, k]        A: [M, lda] int8/uint8 codes (K-major), W: [N, ldw] int8 codes
//
// replaces F.linear(x_q, w_q, bias) on fake-quant values (reference quant_layers.py:499, quant_ultra.py:220).
// 4-bit codes are carried as int8 because Blackwell has no dense int4 MMA (BASELINE.json north_star (3)).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles (128-byte swizzle) of A [128 x 128 B] and
//               W [BN x 128 B] into a kStages-deep shared-memory ring, completion on `full` mbarriers
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::i8 (M=128, N=BN, K=32),
//               4 per stage; tcgen05.commit releases the smem slot (`empty`) and, after the last k-block,
//               publishes the accumulator (`tmem_full`).  Accumulators live in TMEM, double-buffered
//               (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//   warps 2..5  epilogue: tcgen05.ld 32x32b (thread = one output row, 32 columns per load) -> fused
//               dequant / bias / GELU / residual / re-quantise (epilogue.cuh) -> vector stores.
// Ragged M, N, K are handled by TMA out-of-bounds zero fill on loads and predicated stores.
#include <cuda.h>

#include "common.cuh"
#include "epilogue.cuh"
#include "tc_ptx.cuh"

namespace qvit {

constexpr int kBM = 128;          // rows of A per tile (= UMMA M, = TMEM lanes)
constexpr int kBK = 128;          // bytes (= int8 elements) of K per stage = one 128B swizzle atom
constexpr int kUmmaK = 32;        // K per tcgen05.mma for 8-bit operands
constexpr int kEpiWarps = 16;     // four warps per TMEM lane quarter ("quads"), each quad owns a quarter of the columns
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr int kBoxBytes = kBM * 128;                // staging buffer of one quad: 128 rows x (<=128) B, swizzled

// CG = 1: one CTA computes a [128 x BN] tile.  CG = 2: a CTA pair (tcgen05 cta_group::2) computes a [256 x BN] tile;
// each CTA stages its own 128 rows of A and HALF of the W rows, the tensor core reads both halves - 1.5x less
// L2->SM operand traffic per MAC and 32 KiB stages (5 deep) instead of 48 KiB (3 deep).
template <int BN, int CG = 1>
struct GemmSmem {
  static constexpr int kStages = (BN == 256 && CG == 1) ? 3 : 5;
  static constexpr int kABytes = kBM * kBK;
  static constexpr int kBBytes = (BN / CG) * kBK;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = 4 * kBoxBytes;     // one buffer per quad
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kOutBytes + kBarBytes + 1024;   // +1024 for manual alignment
};

__host__ __device__ constexpr int out_elem_size(int out_kind) {
  return out_kind == QVIT_OUT_BF16 ? 2 : (out_kind == QVIT_OUT_I8 ? 1 : 4);
}
// bytes one quad contributes per output row and tile, capped at the 128 B of a swizzle atom: the TMA-store box width
__host__ __device__ constexpr int out_box_bytes(int bn, int out_kind) {
  return (bn / 4) * out_elem_size(out_kind) > 128 ? 128 : (bn / 4) * out_elem_size(out_kind);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// y[32] for columns [n0, n0+32) of row m (dequant, col-scale, bias, activation, residual); guarded when ragged.
template <bool FACC>   // FACC: the accumulators are fp32 (bf16 GEMM) instead of int32
__device__ __forceinline__ void epi_compute32(EpiParams e, const uint32_t (&acc)[32], float scale, int64_t m, int n0,
                                              bool row_ok, float (&y)[32], bool skip_residual) {
  if (skip_residual) e.residual = nullptr;                  // the caller adds it from the TMA-staged tile
  const bool full = (n0 + 32 <= e.N);
  const bool vec_in = full && row_ok && (!e.bias || ((reinterpret_cast<uintptr_t>(e.bias + n0) & 15) == 0)) &&
                      (!e.col_scale || ((reinterpret_cast<uintptr_t>(e.col_scale + n0) & 15) == 0)) &&
                      (!e.residual || ((reinterpret_cast<uintptr_t>(e.residual + m * e.ld_res + n0) & 15) == 0));
  if (vec_in) {
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = (FACC ? __uint_as_float(acc[j]) : (float)(int32_t)acc[j]) * scale;
    if (e.col_scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(e.col_scale + n0) + j);
        y[4 * j] *= c.x; y[4 * j + 1] *= c.y; y[4 * j + 2] *= c.z; y[4 * j + 3] *= c.w;
      }
    }
    if (e.bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(e.bias + n0) + j);
        y[4 * j] += c.x; y[4 * j + 1] += c.y; y[4 * j + 2] += c.z; y[4 * j + 3] += c.w;
      }
    }
    if (e.act == QVIT_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = gelu_erf(y[j]);
    } else if (e.act == QVIT_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.0f);
    }
    if (e.residual) {
      const float* r = e.residual + m * e.ld_res + n0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 c = *reinterpret_cast<const float4*>(r + 4 * j);   // plain load: `out` may alias `residual`
        y[4 * j] += c.x; y[4 * j + 1] += c.y; y[4 * j + 2] += c.z; y[4 * j + 3] += c.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      y[j] = (row_ok && n0 + j < e.N) ? epi_value_f(e, FACC ? __uint_as_float(acc[j]) : (float)(int32_t)acc[j], scale, m, n0 + j) : 0.0f;
  }
}

// KIND 0: int8 x int8 -> int32 (tcgen05 kind::i8).  KIND 1: bf16 x bf16 -> fp32 (kind::f16) for the QAT gradient GEMMs:
// the fp32 gradient operand arrives as three exact bf16 planes concatenated along K, the integer codes as one bf16
// plane that is re-read for every A plane (`b_wrap` k-blocks), i.e. D = (A1 + A2 + A3) * B^T with fp32 accumulation.
template <int BN, int OUT, int CG, int KIND>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                  const EpiParams ep, const int K, const uint32_t idesc, const int tma_store, const int res_tma,
                  const int mma_only, const int b_wrap) {
  using S = GemmSmem<BN, CG>;
  constexpr int kStages = S::kStages;
  constexpr int kTmemCols = 2 * BN;   // 256 or 512: a power of two >= 32
  constexpr int kTileM = kBM * CG;    // rows of the (pair) tile
  const uint32_t rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs)

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t out_base = smem_base + kStages * S::kStageBytes;
  const uint32_t bar_base = out_base + S::kOutBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  auto res_bar = [&](int q) { return bar_base + 8u * (2 * kStages + 4 + q); };
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * S::kStageBytes + S::kOutBytes + 8 * (2 * kStages + 8));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (ep.M + kTileM - 1) / kTileM;
  const int n_tiles = (ep.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + kBK - 1) / kBK;
  const int tile_first = blockIdx.x / CG, tile_step = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    if (tma_store) ptx::prefetch_tmap(&tmap_out);
    if (res_tma) ptx::prefetch_tmap(&tmap_res);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(full_bar(s), 1);
        ptx::mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(tfull_bar(a), 1);
        ptx::mbar_init(tempty_bar(a), kEpiWarps * CG); // one arrive per epilogue warp (of both CTAs of a pair)
      }
      for (int q = 0; q < 4; ++q) ptx::mbar_init(res_bar(q), 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc<CG>(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();   // the peer's barriers are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        const int a_row = m_blk * kTileM + (int)rank * kBM;          // this CTA's 128 rows of A
        const int w_row = n_blk * BN + (int)rank * (BN / CG);        // this CTA's share of the W rows
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * S::kStageBytes;
          const uint32_t b_dst = a_dst + S::kABytes;
          if (mma_only && (tile != tile_first || kb >= kStages)) {
            // benchmark mode (QVIT_OUT_NONE only): operands stay whatever the first kStages loads brought in -
            // measures the tensor-core issue rate with no L2 / HBM traffic at all
            if (rank == 0) ptx::mbar_arrive(full_bar(stage));
          } else if (CG == 1) {
            ptx::mbar_expect_tx(full_bar(stage), S::kStageBytes);
            ptx::tma_load_2d(a_dst, &tmap_a, full_bar(stage), kb * kBK, a_row);
            ptx::tma_load_2d(b_dst, &tmap_w, full_bar(stage), (kb % b_wrap) * kBK, w_row);
          } else {
            // both CTAs signal the LEADER's barrier (peer bit of the shared::cluster address cleared); the leader arms it
            // for the bytes of both.  A peer completion that overtakes the leader's expect_tx only makes the pending
            // tx-count transiently negative: the phase cannot complete before the leader's (single) arrival.
            const uint32_t lead_bar = full_bar(stage) & 0xFEFFFFFFu;
            if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * S::kStageBytes);
            ptx::tma_load_2d_cg2(a_dst, &tmap_a, lead_bar, kb * kBK, a_row);
            ptx::tma_load_2d_cg2(b_dst, &tmap_w, lead_bar, (kb % b_wrap) * kBK, w_row);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA of a pair only)
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);     // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t a_src = smem_base + stage * S::kStageBytes;
          const uint64_t a_desc = ptx::make_kmajor_sw128_desc(a_src);
          const uint64_t b_desc = ptx::make_kmajor_sw128_desc(a_src + S::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance both descriptors by k*32 bytes inside the swizzle atom (address field is >>4)
            if (KIND == 0)
              ptx::mma_i8<CG>(d_tmem, a_desc + (uint64_t)(k * (kUmmaK >> 4)), b_desc + (uint64_t)(k * (kUmmaK >> 4)), idesc,
                              (uint32_t)((kb | k) != 0));
            else
              ptx::mma_bf16<CG>(d_tmem, a_desc + (uint64_t)(k * (kUmmaK >> 4)), b_desc + (uint64_t)(k * (kUmmaK >> 4)), idesc,
                                (uint32_t)((kb | k) != 0));
          }
          // smem slot free once these MMAs retire (in both CTAs of a pair)
          if (CG == 1) ptx::mma_commit(empty_bar(stage)); else ptx::mma_commit_cg2(empty_bar(stage), 0x3);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (CG == 1) ptx::mma_commit(tfull_bar(acc)); else ptx::mma_commit_cg2(tfull_bar(acc), 0x3);   // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..17)
    const int lane_grp = warp & 3;                           // TMEM lanes [32*lane_grp, +32) are this warp's
    const int quad = (warp - 2) >> 2;                        // 0..3: which quarter of the tile's columns
    const int row = lane_grp * 32 + lane;                    // row inside the tile
    const bool leader = (lane_grp == 2 && lane == 0);        // first warp of each quad (warps 2, 6, 10, 14)
    constexpr int kChunksPerQuad = BN / 128;                 // 32-column chunks per quad and tile (2 or 1)
    constexpr int kEsz = out_elem_size(OUT);
    constexpr int kBoxW = out_box_bytes(BN, OUT);            // 128, 64 or 32 bytes per staged row
    constexpr int kChunkBytes = 32 * kEsz;
    constexpr int kChunksPerBox = kBoxW / kChunkBytes;       // 1 or 2
    constexpr uint32_t kSwzMask = kBoxW / 16 - 1;            // TMA swizzle: 16B-chunk index ^= (byte offset >> 7) & mask
    const float scale = epi_scale(ep);
    SymParams nq;
    FastQ fq;
    if (OUT == QVIT_OUT_I8) {
      nq = load_sym_params(ep.next_d, ep.next_qm, ep.next_t);
      fq = make_fastq(nq);
    }
    int fl = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_phase = 0;
    const bool use_res_tma = (OUT == QVIT_OUT_F32) && res_tma;   // residual tile staged by TMA into the output buffer
    const uint32_t buf = out_base + (uint32_t)(quad * kBoxBytes);
    const uint32_t row_off = (uint32_t)(row * kBoxW);
    const uint32_t sw = (row_off >> 7) & kSwzMask;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const int row0 = m_blk * kTileM + (int)rank * kBM;      // first output row of this CTA's half of the tile
      const int64_t m = (int64_t)row0 + row;
      const bool row_ok = m < ep.M;
      const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
      for (int cq = 0; cq < kChunksPerQuad; ++cq) {
        const int c = quad * kChunksPerQuad + cq;            // chunk index inside the tile
        const int n0 = n_blk * BN + c * 32;
        if (use_res_tma && leader) {
          // buffer free again -> fetch the [128 x 32] fp32 residual tile (coalesced, async) while the math runs
          ptx::tma_store_wait_read<0>();
          ptx::mbar_expect_tx(res_bar(quad), (uint32_t)(kBM * 128));
          ptx::tma_load_2d(buf, &tmap_res, res_bar(quad), n0, row0);
        }
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_row + (uint32_t)(c * 32), r);
        ptx::tmem_ld_wait();
        if (cq == kChunksPerQuad - 1) {                      // all TMEM reads of this warp for this tile are done
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 1) ptx::mbar_arrive(tempty_bar(acc));
            else ptx::mbar_arrive_cluster(tempty_bar(acc), 0);   // the leader's MMA warp waits for both CTAs
          }
        }
        if (OUT == QVIT_OUT_NONE) continue;                  // main-loop benchmark mode
        if (!tma_store) {
          if (row_ok && KIND == 0) epi_store_chunk32(ep, &nq, r, scale, m, n0, fl);   // (KIND 1 always uses the TMA path)
          continue;
        }
        // ---- math first (registers only), so that the previous TMA store of this quad drains meanwhile
        uint32_t w[kChunkBytes / 4];
        if (OUT == QVIT_OUT_I32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j % (kChunkBytes / 4)] = r[j];      // kChunkBytes/4 == 32 here
        } else {
          float y[32];
          epi_compute32<KIND == 1>(ep, r, scale, m, n0, row_ok, y, use_res_tma);
          if (use_res_tma) {
            ptx::mbar_wait(res_bar(quad), res_phase);
            res_phase ^= 1u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 rv;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(rv.x), "=f"(rv.y), "=f"(rv.z), "=f"(rv.w)
                           : "r"(buf + row_off + ((((uint32_t)j) ^ sw) << 4)));
              y[4 * j] += rv.x; y[4 * j + 1] += rv.y; y[4 * j + 2] += rv.z; y[4 * j + 3] += rv.w;
            }
          }
          if (OUT == QVIT_OUT_F32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j % (kChunkBytes / 4)] = __float_as_uint(y[j]);
          } else if (OUT == QVIT_OUT_BF16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __nv_bfloat162 p2 = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
              w[j % (kChunkBytes / 4)] = *reinterpret_cast<const uint32_t*>(&p2);
            }
          } else {
            int doubt = fq.generic;
            if (!fq.generic) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                w[j % (kChunkBytes / 4)] = pack4_i8(sym_code_fast(y[4 * j], fq, doubt), sym_code_fast(y[4 * j + 1], fq, doubt),
                                                    sym_code_fast(y[4 * j + 2], fq, doubt), sym_code_fast(y[4 * j + 3], fq, doubt));
            }
            if (doubt) {                                     // rare: an element sits on a rounding boundary (or generic quantizer)
#pragma unroll
              for (int j = 0; j < 8; ++j)
                w[j % (kChunkBytes / 4)] = pack4_i8(sym_code(y[4 * j], nq, fl), sym_code(y[4 * j + 1], nq, fl),
                                                    sym_code(y[4 * j + 2], nq, fl), sym_code(y[4 * j + 3], nq, fl));
            }
          }
        }
        const int in_box = cq % kChunksPerBox;
        if (in_box == 0 && !use_res_tma) {
          if (leader) ptx::tma_store_wait_read<0>();         // the quad's previous box has left shared memory
          named_bar_sync(1 + quad, 128);
        }
#pragma unroll
        for (int j = 0; j < kChunkBytes / 16; ++j) {
          const uint32_t chunk16 = (uint32_t)(in_box * (kChunkBytes / 16) + j);
          sts_v4(buf + row_off + ((chunk16 ^ sw) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        }
        if (in_box == kChunksPerBox - 1) {
          ptx::fence_proxy_async_smem();                     // generic-proxy writes -> visible to the TMA engine
          named_bar_sync(1 + quad, 128);
          if (leader) {
            const int box_n0 = n_blk * BN + (c - in_box) * 32;
            ptx::tma_store_2d(&tmap_out, buf, box_n0, row0);
            ptx::tma_store_commit();
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (tma_store && leader) ptx::tma_store_wait<0>();
    fl = warp_or(fl);
    if (fl && ep.flags && lane == 0) atomicOr(ep.flags, fl);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();   // the peer may still be reading this CTA's operands / signalling its barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

// [rows, K] byte matrix with row pitch ld (bytes), box = [box_rows x 128 B], 128B swizzle
static int make_tmap_bytes(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return QVIT_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld K=%lld ld=%lld", (int)r, (long long)rows,
              (long long)K, (long long)ld);
    return QVIT_ERR_CUDA;
  }
  return QVIT_OK;
}

// output matrix [M, N] with row pitch ldo (elements): box = 128 rows x box_bytes, swizzle mode = box width
static int make_tmap_out(CUtensorMap* map, void* base, int64_t M, int64_t N, int64_t ldo, int out_kind, int box_bytes) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return QVIT_ERR_CUDA;
  }
  CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const int esz = out_elem_size(out_kind);
  if (out_kind == QVIT_OUT_I32) dt = CU_TENSOR_MAP_DATA_TYPE_INT32;
  else if (out_kind == QVIT_OUT_BF16) dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  else if (out_kind == QVIT_OUT_I8) dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
  const CUtensorMapSwizzle swz = box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                  : (box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)(ldo * esz)};
  cuuint32_t box[2] = {(cuuint32_t)(box_bytes / esz), (cuuint32_t)kBM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(out) failed (CUresult %d) M=%lld N=%lld ldo=%lld", (int)r, (long long)M, (long long)N,
              (long long)ldo);
    return QVIT_ERR_CUDA;
  }
  return QVIT_OK;
}

bool gemm_tc_supported(const void* a, int64_t lda, const void* w, int64_t ldw, int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return false;
  if ((lda & 15) || (ldw & 15)) return false;
  int dev = 0, maj = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  return maj == 10 && get_encode_fn() != nullptr;
}

static int g_force_cg = 0;   // 0 = automatic, 1 / 2 = force single-CTA / CTA-pair tiles (tests, benchmarks)
static int g_mma_only = 0;   // benchmark: skip operand loads after the first pipeline fill (QVIT_OUT_NONE only)
static int g_no_tma_store = 0;   // benchmark: per-thread vector stores instead of staged TMA stores
void gemm_tc_force_cta_group(int cg) {
  g_no_tma_store = (cg >= 20) ? 1 : 0;
  g_mma_only = (cg >= 10 && cg < 20) ? 1 : 0;
  g_force_cg = cg % 10;
}

struct TcMaps {
  CUtensorMap a, w, out, res;
  int tma_store, res_tma;
};

template <int BN, int OUT, int CG, int KIND = 0>
static int launch_tc(const TcMaps& tm, const EpiParams& ep, int K, bool a_unsigned, int max_ctas, cudaStream_t s,
                     int b_wrap = 1 << 30) {
  using S = GemmSmem<BN, CG>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_i8_tc_kernel<BN, OUT, CG, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", S::kTotal, cudaGetErrorString(e));
      return QVIT_ERR_CUDA;
    }
    attr_set[dev] = true;
  }
  const int m_tiles = (ep.M + kBM * CG - 1) / (kBM * CG), n_tiles = (ep.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles * CG;
  if (grid > max_ctas) grid = max_ctas;
  if (CG == 2) grid &= ~1;
  const uint32_t idesc = KIND == 0 ? ptx::make_idesc_i8(kBM * CG, BN, !a_unsigned, true) : ptx::make_idesc_bf16_f32(kBM * CG, BN);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_i8_tc_kernel<BN, OUT, CG, KIND>, tm.a, tm.w, tm.out, tm.res, ep, K, idesc,
                                     tm.tma_store, tm.res_tma, (OUT == QVIT_OUT_NONE) ? g_mma_only : 0, b_wrap);
  if (e != cudaSuccess) {
    set_error("gemm_i8_tc_kernel launch: %s", cudaGetErrorString(e));
    return QVIT_ERR_CUDA;
  }
  return check_launch("gemm_i8_tc_kernel");
}

template <int BN, int CG>
static int launch_tc_kind(const TcMaps& tm, const EpiParams& ep, int K, bool a_unsigned, int max_ctas, cudaStream_t s) {
  switch (ep.out_kind) {
    case QVIT_OUT_I32: return launch_tc<BN, QVIT_OUT_I32, CG>(tm, ep, K, a_unsigned, max_ctas, s);
    case QVIT_OUT_F32: return launch_tc<BN, QVIT_OUT_F32, CG>(tm, ep, K, a_unsigned, max_ctas, s);
    case QVIT_OUT_BF16: return launch_tc<BN, QVIT_OUT_BF16, CG>(tm, ep, K, a_unsigned, max_ctas, s);
    case QVIT_OUT_I8: return launch_tc<BN, QVIT_OUT_I8, CG>(tm, ep, K, a_unsigned, max_ctas, s);
    default: return launch_tc<BN, QVIT_OUT_NONE, CG>(tm, ep, K, a_unsigned, max_ctas, s);
  }
}


int gemm_tc_launch(const void* a, int64_t lda, int a_unsigned, const int8_t* w, int64_t ldw, const EpiParams& ep, int K,
                   cudaStream_t s) {
  const int M = ep.M, N = ep.N;
  // tile width: 256 for wide layers, 128 when that fills the machine better or N is small
  const int sms = sm_count();
  int bn = 256;
  if (N <= 128) bn = 128;
  else {
    const int64_t t256 = (int64_t)((M + kBM - 1) / kBM) * ((N + 255) / 256);
    if (t256 < sms) bn = 128;
  }
  // CTA pairs once there is at least one full wave of [256 x 256] pair tiles
  // Measured (tools/epi_bench.py, M = 50 432): pairs win when the epilogue is light (main loop only 129 -> 114 us, bf16
  // out 181 -> 167 us) and lose a little when it is the bottleneck (fp32 + residual 282 -> 310 us, int8 + GELU 310 -> 320 us),
  // so automatic mode uses them for the raw / bf16 kinds only.
  int cg = 1;
  if (bn == 256 && (int64_t)((M + 2 * kBM - 1) / (2 * kBM)) * ((N + 255) / 256) * 2 >= sms &&
      (ep.out_kind == QVIT_OUT_BF16 || ep.out_kind == QVIT_OUT_I32 || ep.out_kind == QVIT_OUT_NONE) && !ep.residual)
    cg = 2;
  if (g_force_cg == 1) cg = 1;
  if (g_force_cg == 2 && bn == 256) cg = 2;
  TcMaps tm;
  int rc = make_tmap_bytes(&tm.a, a, M, K, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bytes(&tm.w, w, N, K, ldw, bn / cg);
  if (rc) return rc;
  // Coalesced output through shared memory + TMA store when the output matrix is TMA-addressable;
  // predicated per-thread vector stores otherwise.
  const int esz = out_elem_size(ep.out_kind);
  tm.tma_store = (ep.out_kind != QVIT_OUT_NONE) && ((reinterpret_cast<uintptr_t>(ep.out) & 15) == 0) &&
                 (((ep.ldo * esz) & 15) == 0) && !g_no_tma_store;
  tm.out = tm.a;
  tm.res = tm.a;
  if (tm.tma_store) {
    rc = make_tmap_out(&tm.out, ep.out, M, N, ep.ldo, ep.out_kind, out_box_bytes(bn, ep.out_kind));
    if (rc) return rc;
  }
  // fp32 residual (Block.forward's "x + ...", vit_model.py:206-207) staged tile-wise by TMA instead of per-thread row reads
  tm.res_tma = tm.tma_store && ep.out_kind == QVIT_OUT_F32 && ep.residual != nullptr &&
               ((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0) && (((ep.ld_res * 4) & 15) == 0);
  if (tm.res_tma) {
    rc = make_tmap_out(&tm.res, const_cast<float*>(ep.residual), M, N, ep.ld_res, QVIT_OUT_F32, 128);
    if (rc) return rc;
  }
  if (cg == 2) return launch_tc_kind<256, 2>(tm, ep, K, a_unsigned != 0, sms, s);
  if (bn == 256) return launch_tc_kind<256, 1>(tm, ep, K, a_unsigned != 0, sms, s);
  return launch_tc_kind<128, 1>(tm, ep, K, a_unsigned != 0, sms, s);
}

// D[M, N] (fp32) = scale * sum_p A_p[M, K] * B[N, K]^T : A holds `planes` bf16 planes side by side ([M, planes * Kp]),
// B one bf16 plane [N, >= Kp]; Kp = K rounded up to 64 elements (one 128-byte k-block).
int gemm_tc_launch_bf16_split(const void* a, int64_t lda, int planes, const void* b, int64_t ldb, const EpiParams& ep, int K,
                              cudaStream_t s) {
  const int M = ep.M, N = ep.N;
  const int Kp = (K + 63) / 64 * 64;
  const int sms = sm_count();
  int bn = 256;
  if (N <= 128 || (int64_t)((M + kBM - 1) / kBM) * ((N + 255) / 256) < sms) bn = 128;
  TcMaps tm;
  // byte views: the A row holds planes * Kp bf16 = 2 * planes * Kp bytes; OOB rows / k are zero filled
  int rc = make_tmap_bytes(&tm.a, a, M, (int64_t)2 * planes * Kp, lda * 2, kBM);
  if (rc) return rc;
  rc = make_tmap_bytes(&tm.w, b, N, (int64_t)2 * Kp, ldb * 2, bn);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(ep.out) & 15) || ((ep.ldo * 4) & 15)) {
    set_error("qvit_gemm_bf16_split: output must be 16-byte aligned with a pitch that is a multiple of 4 floats");
    return QVIT_ERR_UNSUPPORTED;
  }
  tm.tma_store = 1;
  tm.res = tm.a;
  rc = make_tmap_out(&tm.out, ep.out, M, N, ep.ldo, QVIT_OUT_F32, out_box_bytes(bn, QVIT_OUT_F32));
  if (rc) return rc;
  tm.res_tma = ep.residual != nullptr && ((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0) && (((ep.ld_res * 4) & 15) == 0);
  if (tm.res_tma) {
    rc = make_tmap_out(&tm.res, const_cast<float*>(ep.residual), M, N, ep.ld_res, QVIT_OUT_F32, 128);
    if (rc) return rc;
  }
  const int k_bytes = 2 * planes * Kp;          // contraction length of the kernel's byte-wise K loop
  const int b_wrap = (2 * Kp) / kBK;            // k-blocks per plane: the B coordinate wraps, A runs through all planes
  if (bn == 256) return launch_tc<256, QVIT_OUT_F32, 1, 1>(tm, ep, k_bytes, false, sms, s, b_wrap);
  return launch_tc<128, QVIT_OUT_F32, 1, 1>(tm, ep, k_bytes, false, sms, s, b_wrap);
}

}  // namespace qvit
