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
constexpr int kGemmThreads = 192; // 6 warps

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = kBM * kBK;
  static constexpr int kBBytes = BN * kBK;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;   // +1024 for manual alignment
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const EpiParams ep, const int K, const uint32_t idesc) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  constexpr int kTmemCols = 2 * BN;   // 256 or 512: a power of two >= 32

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kStages * S::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * S::kStageBytes + 8 * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (ep.M + kBM - 1) / kBM;
  const int n_tiles = (ep.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(full_bar(s), 1);
        ptx::mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(tfull_bar(a), 1);
        ptx::mbar_init(tempty_bar(a), 4);      // one arrive per epilogue warp
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * S::kStageBytes;
          const uint32_t b_dst = a_dst + S::kABytes;
          ptx::mbar_expect_tx(full_bar(stage), S::kStageBytes);
          ptx::tma_load_2d(a_dst, &tmap_a, full_bar(stage), kb * kBK, m_blk * kBM);
          ptx::tma_load_2d(b_dst, &tmap_w, full_bar(stage), kb * kBK, n_blk * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
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
            ptx::mma_i8<1>(d_tmem, a_desc + (uint64_t)(k * (kUmmaK >> 4)), b_desc + (uint64_t)(k * (kUmmaK >> 4)), idesc,
                           (uint32_t)((kb | k) != 0));
          }
          ptx::mma_commit(empty_bar(stage));                 // smem slot free once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        ptx::mma_commit(tfull_bar(acc));                     // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int lane_grp = warp & 3;                           // TMEM lanes [32*lane_grp, +32) are this warp's
    const float scale = epi_scale(ep);
    SymParams nq;
    if (ep.out_kind == QVIT_OUT_I8) nq = load_sym_params(ep.next_d, ep.next_qm, ep.next_t);
    int fl = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const int64_t m = (int64_t)m_blk * kBM + lane_grp * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_row + (uint32_t)(c * 32), r);
        ptx::tmem_ld_wait();
        if (m < ep.M) epi_store_chunk32(ep, &nq, r, scale, m, n_blk * BN + c * 32, fl);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    fl = warp_or(fl);
    if (fl && ep.flags && lane == 0) atomicOr(ep.flags, fl);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
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

bool gemm_tc_supported(const void* a, int64_t lda, const void* w, int64_t ldw, int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return false;
  if ((lda & 15) || (ldw & 15)) return false;
  int dev = 0, maj = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  return maj == 10 && get_encode_fn() != nullptr;
}

template <int BN>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tw, const EpiParams& ep, int K, bool a_unsigned,
                     int max_ctas, cudaStream_t s) {
  using S = GemmSmem<BN>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_i8_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", S::kTotal, cudaGetErrorString(e));
      return QVIT_ERR_CUDA;
    }
    attr_set[dev] = true;
  }
  const int m_tiles = (ep.M + kBM - 1) / kBM, n_tiles = (ep.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles;
  if (grid > max_ctas) grid = max_ctas;
  const uint32_t idesc = ptx::make_idesc_i8(kBM, BN, !a_unsigned, true);
  gemm_i8_tc_kernel<BN><<<grid, kGemmThreads, S::kTotal, s>>>(ta, tw, ep, K, idesc);
  return check_launch("gemm_i8_tc_kernel");
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
  CUtensorMap ta, tw;
  int rc = make_tmap_bytes(&ta, a, M, K, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bytes(&tw, w, N, K, ldw, bn);
  if (rc) return rc;
  if (bn == 256) return launch_tc<256>(ta, tw, ep, K, a_unsigned != 0, sms, s);
  return launch_tc<128>(ta, tw, ep, K, a_unsigned != 0, sms, s);
}

}  // namespace qvit
