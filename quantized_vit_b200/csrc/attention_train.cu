// Attention core of the QAT step: forward with saved log-sum-exp, and the backward, on the tcgen05 tensor cores.
//
// Glue BESIDE the hot path (ViTAttention.forward, vit_model.py:133-149, under autograd - config 3):
// the library fp32 kernels this replaces (memory-efficient SDPA forward + backward) were 30 % of the step.
//
// Forward  : qkv (fp32) -> two fp16 planes (split2_f16, attention_f16.cu) -> attention_f16x2_kernel<kLse = true>.
// Backward : ONE launch, work item = (pass, batch, head); both passes are the same program with the operand roles swapped:
//                        rows R1   R2      columns C1  C2     S = R1 C1^T   dP = R2 C2^T   P = 2^(S s - L[query])   dS = P (dP - D[query])
//   pass 0 (dQ)          q_i  dO_i        k   v      (statistics per ROW)      out1 = dS C1          -> dq = scale * out1
//   pass 1 (dK, dV)      k_j  v_j         q   dO     (statistics per COLUMN)   out1 = dS C1, out2 = P C2 -> dk = scale * out1, dv = out2
//   (pass 1 computes S^T and dP^T directly, so no transposed copy of P / dS is ever needed: every "A" operand of the second
//   products sits in TMEM with the lane = the output row, TS-mode MMAs as in the forward.)
// Operands are converted in the kernel from fp32 to TWO bf16 planes (hi + lo = 16 significant bits, fp32 range - gradients need
// no scaling) written in the 128B-swizzled layout the MMAs read; each product is lo*hi' + hi*lo' + hi*hi' with fp32 accumulation.
// The key / query columns are walked in chunks of 64: S and dP of a chunk (2 x 64 TMEM columns, three buffers) are turned into
// the bf16 planes of P and dS in place, and the second products accumulate over the chunks into out1 / out2 (2 x 64 columns).
//   control warp    converged, one elected lane issues every MMA
//   16 compute warps (thread = row, four warps share a row's 64 chunk columns): operand conversion, P / dS, epilogue
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace qvit {

int attention_f16x2_launch(const void* planes, int64_t ld, int plane_off, int B, int T, int H, int head_dim, float scale, int exp_q,
                           int exp_k, int exp_v, const float* d, const float* q_m, const float* t, int8_t* codes, int64_t ld_codes,
                           float* out, int32_t* flags, long long* prof, float* lse, cudaStream_t stream);

namespace ab {
constexpr int kHd = 64;
constexpr int kMR = 128;               // rows per tile
constexpr int kNC = 208;               // columns per item (13 x 16)
constexpr int kStat = 256;             // pitch of the per-(batch, head) statistics vectors
constexpr int kComputeWarps = 16;
constexpr int kThreads = 32 * (kComputeWarps + 1);
constexpr int kRPlane = kMR * 128;     // 16 KiB
constexpr int kCPlane = kNC * 128;     // 26 KiB
constexpr int kOffR = 0;                                   // R1 hi, R1 lo, R2 hi, R2 lo
constexpr int kOffC = kOffR + 4 * kRPlane;                 // C1 hi, C1 lo, C2 hi, C2 lo
constexpr int kOffStat = kOffC + 4 * kCPlane;              // L[256], D[256] of the item (pass 1)
constexpr int kOffBar = kOffStat + 2 * kStat * 4;
constexpr int kSmem = kOffBar + 256 + 1024;
#ifndef QVIT_ATT_PREFETCH
#define QVIT_ATT_PREFETCH 1
#endif
constexpr bool kPrefetchRows = QVIT_ATT_PREFETCH != 0;
constexpr int kBufs = 3;               // chunk buffers in TMEM
constexpr int kBufCols = 128;          // one chunk buffer: S / P planes [0, 64), dP / dS planes [64, 128)
constexpr int kOut1 = kBufs * kBufCols, kOut2 = kOut1 + 64;      // 384, 448: all 512 columns are in use

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool b_mn_major) {
  return (1u << 4) /*D = f32*/ | (1u << 7) /*A = bf16*/ | (1u << 10) /*B = bf16*/ | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t idesc_rt(int N, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kMR >> 4) << 24);
}
// two fp32 values -> packed bf16 hi pair and lo pair (first value in the low half)
__device__ __forceinline__ void split2_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h2);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// One 16-byte piece (8 values) of an operand row -> the hi / lo bf16 planes of the tile: 128-byte rows, 16-byte pieces XOR-swizzled
// with the row number (what a SWIZZLE_128B tensor map would have written)
__device__ __forceinline__ void store_piece(uint8_t* hi, uint8_t* lo, int r, int ch, const float4& a, const float4& b) {
  uint4 h, l;
  split2_bf16(a.x, a.y, h.x, l.x);
  split2_bf16(a.z, a.w, h.y, l.y);
  split2_bf16(b.x, b.y, h.z, l.z);
  split2_bf16(b.z, b.w, h.w, l.w);
  const int off = r * 128 + ((ch ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(hi + off) = h;
  *reinterpret_cast<uint4*>(lo + off) = l;
}
// fp32 rows [valid_rows, 64] (pitch row_stride) -> planes of total_rows (<= 256) rows, rows past valid_rows zero.  All global loads
// of the thread are issued before the first conversion: one memory latency per call.
__device__ __forceinline__ void load_planes_c(uint8_t* hi, uint8_t* lo, const float* __restrict__ src, int64_t row_stride, int valid_rows,
                                              int total_rows, int tid) {
  float4 v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = tid + k * 32 * kComputeWarps, r = idx >> 3, ch = idx & 7;
    v[2 * k] = v[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < valid_rows) {
      const float4* p = reinterpret_cast<const float4*>(src + (int64_t)r * row_stride + ch * 8);
      v[2 * k] = __ldg(p);
      v[2 * k + 1] = __ldg(p + 1);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = tid + k * 32 * kComputeWarps, r = idx >> 3, ch = idx & 7;
    if (r < total_rows) store_piece(hi, lo, r, ch, v[2 * k], v[2 * k + 1]);
  }
}
}  // namespace ab

// qkv fp32 [B, T, 3, H, 64]; dout fp32 [B, T, H, 64]; lse / dstat fp32 [B, H, 256] (finite beyond T); dqkv fp32 like qkv
__global__ void __launch_bounds__(ab::kThreads, 1)
attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                     const float* __restrict__ dstat, float* __restrict__ dqkv, int T, int H, int total_items, float scale,
                     long long* __restrict__ prof) {
  using namespace ab;
  using ptx::mma_f16_ss;
  using ptx::mma_f16_ts;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  float* Ls = reinterpret_cast<float*>(gen + kOffStat);
  float* Ds = Ls + kStat;
  const uint32_t bar0 = base + kOffBar;
  const uint32_t ops_ready = bar0, mma_done = bar0 + 8u, out_done = bar0 + 16u;
  auto sdp_done = [&](int b) { return bar0 + 24u + 8u * b; };
  auto pds_ready = [&](int b) { return bar0 + 48u + 8u * b; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + kOffBar + 128);

  // warp-uniform values the compiler can SEE are uniform (shuffle from lane 0): the MMA-issuing code below then keeps its
  // descriptors in uniform registers and issues each tcgen05.mma directly.  (With a plain `lane == 0` guard every MMA was wrapped
  // in an elect / broadcast loop: tools/ubench/mma_rate.cu, 134 -> 83 cycles per issued MMA.)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int D = H * kHd;
  const int64_t D3 = 3 * (int64_t)D;
  const int Tc = (T + 15) & ~15;
  const int n_chunks = (Tc + 63) >> 6;
  const int r_tiles = (T + kMR - 1) / kMR;
  const int n_items = ((int)blockIdx.x < total_items) ? (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int n_tiles = n_items * r_tiles;
  const int half_grid = ((int)gridDim.x + 1) >> 1;
  // item -> (pair, pass); consecutive items of a CTA alternate between the passes (pass 1 is the longer one)
  auto decode = [&](int i, int& pair, int& pass) {
    const int id = (int)blockIdx.x + i * (int)gridDim.x;
    pair = id >> 1;
    pass = (id & 1) ^ ((pair / half_grid) & 1);
  };

  if (threadIdx.x == 0) {
    ptx::mbar_init(ops_ready, kComputeWarps);
    ptx::mbar_init(mma_done, 1);
    ptx::mbar_init(out_done, 1);
    for (int i = 0; i < kBufs; ++i) {
      ptx::mbar_init(sdp_done(i), 1);
      ptx::mbar_init(pds_ready(i), kComputeWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == kComputeWarps) {
    // ------------------------------------------------------------------------------------------ control warp (converged; one
    // elected lane issues)
    const uint32_t r1h = base + kOffR, r1l = r1h + kRPlane, r2h = r1l + kRPlane, r2l = r2h + kRPlane;
    const uint32_t c1h = base + kOffC, c1l = c1h + kCPlane, c2h = c1l + kCPlane, c2l = c2h + kCPlane;
    uint32_t ph_ops = 0, ph_mma = 0, ph_pds = 0;
    // S and dP of chunk c into buffer b: [128 rows] x [width columns], K = 64 head-dim values in 4 steps, three terms each
    auto issue_sdp = [&](int c, int b) {
      const int width = min(64, Tc - 64 * c);
      const uint32_t idesc = idesc_rt(width, false);
      const uint32_t coff = (uint32_t)c * 8192u;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint32_t ah = which ? r2h : r1h, al = which ? r2l : r1l, bh = (which ? c2h : c1h) + coff, bl = (which ? c2l : c1l) + coff;
        const uint32_t d = tmem + (uint32_t)(b * kBufCols + which * 64);
        const uint64_t dah = ptx::make_kmajor_sw128_desc(ah), dal = ptx::make_kmajor_sw128_desc(al);
        const uint64_t dbh = ptx::make_kmajor_sw128_desc(bh), dbl = ptx::make_kmajor_sw128_desc(bl);
#pragma unroll
        for (int ks = 0; ks < kHd / 16; ++ks) mma_f16_ss(d, dal + 2 * ks, dbh + 2 * ks, idesc, ks ? 1u : 0u);   // lo * hi'
#pragma unroll
        for (int ks = 0; ks < kHd / 16; ++ks) mma_f16_ss(d, dah + 2 * ks, dbl + 2 * ks, idesc, 1u);             // hi * lo'
#pragma unroll
        for (int ks = 0; ks < kHd / 16; ++ks) mma_f16_ss(d, dah + 2 * ks, dbh + 2 * ks, idesc, 1u);             // hi * hi'
      }
      ptx::mma_commit(sdp_done(b));
    };
    // out1 += dS_c C1_c (and out2 += P_c C2_c): A = the packed bf16 planes in TMEM (hi words [0, 32), lo words [32, 64) of the
    // 64-column half), B = the column operand's rows as they lie in memory ([token][head dim] = MN-major, 2048 B per 16 tokens)
    auto issue_out = [&](int c, int b, bool two, uint32_t acc0) {
      const int ksteps = min(64, Tc - 64 * c) >> 4;
      constexpr uint32_t idesc = idesc_bf16(kMR, kHd, true);
      const uint32_t coff = (uint32_t)c * 8192u;
      for (int which = 0; which < (two ? 2 : 1); ++which) {
        const uint32_t pa = tmem + (uint32_t)(b * kBufCols + (which ? 0 : 64));     // out1 <- dS (second half), out2 <- P (first half)
        const uint64_t dh0 = ptx::make_kmajor_sw128_desc((which ? c2h : c1h) + coff), dl0 = ptx::make_kmajor_sw128_desc((which ? c2l : c1l) + coff);
        const uint32_t d = tmem + (uint32_t)(which ? kOut2 : kOut1);
        uint32_t acc = acc0;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t dh = dh0 + (uint64_t)(ks * 128), dl = dl0 + (uint64_t)(ks * 128);
          mma_f16_ts(d, pa + (uint32_t)(32 + ks * 8), dh, idesc, acc);     // lo * hi'
          mma_f16_ts(d, pa + (uint32_t)(ks * 8), dl, idesc, 1u);           // hi * lo'
          mma_f16_ts(d, pa + (uint32_t)(ks * 8), dh, idesc, 1u);           // hi * hi'
          acc = 1;
        }
      }
    };
    for (int it = 0; it < n_items; ++it) {
      int pair, pass;
      decode(it, pair, pass);
      for (int rt = 0; rt < r_tiles; ++rt) {
        const int tt = it * r_tiles + rt;
        long long* pc = (prof && blockIdx.x == 0 && tt < 8 && lane == 0) ? prof + tt * 32 : nullptr;
        if (pc) pc[0] = clock64();
        ptx::mbar_wait(ops_ready, ph_ops);
        ph_ops ^= 1;
        ptx::tc_fence_after();
        if (pc) pc[1] = clock64();
        if (ptx::elect_one()) {
          for (int c = 0; c < kBufs && c < n_chunks; ++c) issue_sdp(c, c);
        }
        __syncwarp();
        if (pc) pc[2] = clock64();
        for (int c = 0; c < n_chunks; ++c) {
          const int b = c % kBufs;
          ptx::mbar_wait(pds_ready(b), (ph_pds >> b) & 1u);
          ph_pds ^= 1u << b;
          ptx::tc_fence_after();
          if (pc && c < 4) pc[3 + c * 3] = clock64();
          if (ptx::elect_one()) {
            issue_out(c, b, pass == 1, c > 0 ? 1u : 0u);
            if (c + kBufs < n_chunks) ptx::mma_commit(mma_done);
            if (c == n_chunks - 1) ptx::mma_commit(out_done);
          }
          __syncwarp();
          if (pc && c < 4) pc[4 + c * 3] = clock64();
          if (c + kBufs < n_chunks) {
            ptx::mbar_wait(mma_done, ph_mma);               // the planes of buffer b have been consumed: refill it with chunk c + kBufs
            ph_mma ^= 1;
            if (pc && c < 2) pc[5 + c * 3] = clock64();
            if (ptx::elect_one()) issue_sdp(c + kBufs, b);
            __syncwarp();
          }
        }
        if (pc) pc[15] = clock64();
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ compute warps
    const int tid = threadIdx.x;
    const int row = (warp & 3) * 32 + lane;
    const int cq = warp >> 2;
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
    const float s2 = scale * 1.4426950408889634f;
    uint32_t ph_sdp = 0, ph_out = 0;
    // row operands of tile g: the global loads (into registers: issued one tile ahead, before the wait for the current tile's last
    // MMAs) and, once those MMAs - the last readers of the old planes - are done, the conversion into the planes
    auto r_issue = [&](int g, float4 (&pf)[8]) {
      int pair, pass;
      decode(g / r_tiles, pair, pass);
      const int bi = pair / H, h = pair % H, r0 = (g % r_tiles) * kMR, vr = min(kMR, T - r0);
      const float* qb = qkv + ((int64_t)bi * T + r0) * D3 + h * kHd;
      const float* s1 = pass == 0 ? qb : qb + D;                                                       // q tile | k tile
      const float* s2p = pass == 0 ? dout + ((int64_t)bi * T + r0) * D + h * kHd : qb + 2 * D;         // dO tile | v tile
      const int64_t st2 = pass == 0 ? (int64_t)D : D3;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int idx = tid + k * 32 * kComputeWarps, r = idx >> 3, ch = idx & 7;
        pf[2 * k] = pf[2 * k + 1] = pf[4 + 2 * k] = pf[5 + 2 * k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < vr) {
          const float4* p1 = reinterpret_cast<const float4*>(s1 + (int64_t)r * D3 + ch * 8);
          const float4* p2 = reinterpret_cast<const float4*>(s2p + (int64_t)r * st2 + ch * 8);
          pf[2 * k] = __ldg(p1);
          pf[2 * k + 1] = __ldg(p1 + 1);
          pf[4 + 2 * k] = __ldg(p2);
          pf[5 + 2 * k] = __ldg(p2 + 1);
        }
      }
    };
    auto r_store = [&](const float4 (&pf)[8]) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int idx = tid + k * 32 * kComputeWarps, r = idx >> 3, ch = idx & 7;
        store_piece(gen + kOffR, gen + kOffR + kRPlane, r, ch, pf[2 * k], pf[2 * k + 1]);
        store_piece(gen + kOffR + 2 * kRPlane, gen + kOffR + 3 * kRPlane, r, ch, pf[4 + 2 * k], pf[5 + 2 * k]);
      }
    };
    // L2 prefetch of the NEXT tile's operands (one 128-byte line per thread and operand row half), issued at the start of a tile:
    // the bursts of register loads above then hit L2 instead of queueing on HBM behind every other CTA's burst
    auto l2_prefetch = [&](int g) {
      int pair, pass;
      decode(g / r_tiles, pair, pass);
      const int bi = pair / H, h = pair % H, rt = g % r_tiles, r0 = rt * kMR, vr = min(kMR, T - r0);
      const float* qb = qkv + (int64_t)bi * T * D3 + h * kHd;
      const float* gb = dout + (int64_t)bi * T * D + h * kHd;
      {
        const int op = tid >> 8, r = (tid & 255) >> 1, half = tid & 1;
        const float* src = op == 0 ? (pass == 0 ? qb : qb + D) : (pass == 0 ? gb : qb + 2 * D);
        const int64_t st = (op == 1 && pass == 0) ? (int64_t)D : D3;
        if (r < vr) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (int64_t)(r0 + r) * st + half * 32));
      }
      if (rt == 0) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int idx = tid + k * 32 * kComputeWarps, op = idx >= 2 * kNC ? 1 : 0, rem = idx - op * 2 * kNC, r = rem >> 1, half = rem & 1;
          const float* src = op == 0 ? (pass == 0 ? qb + D : qb) : (pass == 0 ? qb + 2 * D : gb);
          const int64_t st = (op == 1 && pass == 1) ? (int64_t)D : D3;
          if (idx < 4 * kNC && r < T) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (int64_t)r * st + half * 32));
        }
      }
    };
    if (kPrefetchRows && n_tiles > 0) {
      float4 pf[8];
      r_issue(0, pf);
      r_store(pf);
    }
#pragma unroll 1
    for (int g = 0; g < n_tiles; ++g) {
      const int it = g / r_tiles, rt = g - it * r_tiles;
      int pair, pass;
      decode(it, pair, pass);
      const int bi = pair / H, h = pair % H;
      const float* Lg = lse + (int64_t)pair * kStat;
      const float* Dg = dstat + (int64_t)pair * kStat;
      const int r0 = rt * kMR;
      const bool rows_live = r0 + (warp & 3) * 32 < T;
      long long* pk = (prof && blockIdx.x == 0 && tid == 0 && g < 8) ? prof + g * 32 + 16 : nullptr;
      if (pk) pk[0] = clock64();
      // ---- operands (the MMAs of the previous tile have completed: every warp waited for out_done in its epilogue)
      if (!kPrefetchRows) {
        float4 pf[8];
        r_issue(g, pf);
        r_store(pf);
      }
      if (rt == 0) {
        const float* qb = qkv + (int64_t)bi * T * D3 + h * kHd;             // q rows; k at + D, v at + 2 D
        const float* gb = dout + (int64_t)bi * T * D + h * kHd;
        // C1 = k | q, C2 = v | dO  (two rounds of eight 16-byte loads per thread: sixteen at once do not stay in registers)
        load_planes_c(gen + kOffC, gen + kOffC + kCPlane, pass == 0 ? qb + D : qb, D3, T, Tc, tid);
        load_planes_c(gen + kOffC + 2 * kCPlane, gen + kOffC + 3 * kCPlane, pass == 0 ? qb + 2 * D : gb, pass == 0 ? D3 : (int64_t)D, T, Tc, tid);
        if (pass == 1) {
          if (tid < kStat) {
            Ls[tid] = __ldg(Lg + tid);
            Ds[tid] = __ldg(Dg + tid);
          }
        }
      }
      ptx::fence_proxy_async_smem();
      if (rt == 0 && pass == 1) asm volatile("bar.sync 5, 512;" ::: "memory");     // Ls / Ds visible to every compute warp
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ops_ready);
      if (g + 1 < n_tiles) l2_prefetch(g + 1);
      if (pk) pk[1] = clock64();
      float l_row = 0.f, d_row = 0.f;
      if (pass == 0) {
        l_row = __ldg(Lg + r0 + row);
        d_row = __ldg(Dg + r0 + row);
      }
      // ---- chunks
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const int b = c % kBufs;
        const int width = min(64, Tc - 64 * c);
        const uint32_t buf = tmem + lane_addr + (uint32_t)(b * kBufCols);
        ptx::mbar_wait(sdp_done(b), (ph_sdp >> b) & 1u);
        ph_sdp ^= 1u << b;
        ptx::tc_fence_after();
        if (pk && c < 4) pk[2 + c * 3] = clock64();
        const bool active = rows_live && cq * 16 < width;
        uint32_t s[16], gq[16];
        if (active) {
          ptx::tmem_ld16(buf + (uint32_t)(cq * 16), s);
          ptx::tmem_ld16(buf + (uint32_t)(64 + cq * 16), gq);
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before();
        // every warp of the lane quarter has read its S / dP columns before the planes overwrite them
        if (rows_live) asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp & 3)) : "memory");
        ptx::tc_fence_after();
        if (pk && c < 4) pk[3 + c * 3] = clock64();
        if (active) {
          uint32_t ph[8], pl[8], dh[8], dl[8];
          const int cb = c * 64 + cq * 16;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float l0 = l_row, l1 = l_row, e0 = d_row, e1 = d_row;
            if (pass == 1) {
              l0 = Ls[cb + 2 * j];
              l1 = Ls[cb + 2 * j + 1];
              e0 = Ds[cb + 2 * j];
              e1 = Ds[cb + 2 * j + 1];
            }
            const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * j]), s2, -l0));
            const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * j + 1]), s2, -l1));
            const float ds0 = p0 * (__uint_as_float(gq[2 * j]) - e0), ds1 = p1 * (__uint_as_float(gq[2 * j + 1]) - e1);
            split2_bf16(p0, p1, ph[j], pl[j]);
            split2_bf16(ds0, ds1, dh[j], dl[j]);
          }
          ptx::tmem_st<8>(buf + (uint32_t)(cq * 8), ph);
          ptx::tmem_st<8>(buf + (uint32_t)(32 + cq * 8), pl);
          ptx::tmem_st<8>(buf + (uint32_t)(64 + cq * 8), dh);
          ptx::tmem_st<8>(buf + (uint32_t)(96 + cq * 8), dl);
          ptx::tmem_st_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(pds_ready(b));
        if (pk && c < 4) pk[4 + c * 3] = clock64();
      }
      // ---- next tile's row operands on their way while the last MMAs finish, then the epilogue
      if (kPrefetchRows) {
        // (unconditional: after the last tile the same rows are fetched and stored once more - a branch around the loads makes the
        // compiler keep them in local memory, and the store to it waits for the data right here)
        float4 pf[8];
        r_issue(min(g + 1, n_tiles - 1), pf);
        ptx::mbar_wait(out_done, ph_out);
        ptx::tc_fence_after();
        if (pk) pk[14] = clock64();
        r_store(pf);                                        // every MMA that read the old planes has completed
      } else {
        ptx::mbar_wait(out_done, ph_out);
        ptx::tc_fence_after();
        if (pk) pk[14] = clock64();
      }
      ph_out ^= 1;
      const int t = r0 + row;
      if (rows_live) {
        uint32_t o1[16], o2[16];
        ptx::tmem_ld16(tmem + lane_addr + (uint32_t)(kOut1 + cq * 16), o1);
        if (pass == 1) ptx::tmem_ld16(tmem + lane_addr + (uint32_t)(kOut2 + cq * 16), o2);
        ptx::tmem_ld_wait();
        if (t < T) {
          float* dst = dqkv + ((int64_t)bi * T + t) * D3 + (pass == 0 ? 0 : D) + h * kHd + cq * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j) o1[j] = __float_as_uint(__uint_as_float(o1[j]) * scale);
          stg_v8_b32(dst, o1);
          stg_v8_b32(dst + 8, o1 + 8);
          if (pass == 1) {
            stg_v8_b32(dst + D, o2);
            stg_v8_b32(dst + D + 8, o2 + 8);
          }
        }
      }
      ptx::tc_fence_before();
      if (pk) pk[15] = clock64();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem, 512);
  }
}

// dstat[(b * H + h) * 256 + t] = sum_d dout[b, t, h, d] * out[b, t, h, d]  (0 for t >= T): one warp per (b, t < 256)
__global__ void attention_dstat_kernel(const float* __restrict__ out, const float* __restrict__ dout, float* __restrict__ dstat, int B, int T,
                                       int H) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * ab::kStat) return;
  const int b = (int)(w / ab::kStat), t = (int)(w % ab::kStat);
  const int D = H * ab::kHd;
  for (int c0 = 0; c0 < D; c0 += 128) {
    const int col = c0 + lane * 4;
    float s = 0.f;
    if (t < T && col < D) {
      const float4 o = __ldg(reinterpret_cast<const float4*>(out + ((int64_t)b * T + t) * D + col));
      const float4 g = __ldg(reinterpret_cast<const float4*>(dout + ((int64_t)b * T + t) * D + col));
      s = (o.x * g.x + o.y * g.y) + (o.z * g.z + o.w * g.w);
    }
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((lane & 15) == 0 && col < D) dstat[((int64_t)b * H + col / ab::kHd) * ab::kStat + t] = s;
  }
}

}  // namespace qvit

using namespace qvit;

extern "C" int qvit_attention_train_bwd_prof(const float* qkv, const float* out, const float* dout, const float* lse, int B, int T, int H,
                                             int head_dim, float scale, float* dstat, float* dqkv, long long* prof, qvit_stream_t stream);
extern "C" int qvit_split2_f16(const float* x, int64_t rows, int cols, int64_t ldx, const int* col_exp, void* out, int64_t ld,
                               int plane_off, int32_t* flags, qvit_stream_t stream);

static int train_shape_ok(const char* who, int B, int T, int H, int head_dim) {
  if (B <= 0 || T <= 0 || H <= 0 || head_dim != ab::kHd || T > ab::kNC) {
    set_error("%s: supports head_dim == 64 and 1 <= T <= 208 (got B=%d T=%d H=%d head_dim=%d)", who, B, T, H, head_dim);
    return QVIT_ERR_UNSUPPORTED;
  }
  return QVIT_OK;
}

// Training forward of the attention core: out[b, t, h, :] = softmax(q k^T * scale) v from qkv fp32 [B, T, 3, H, 64] (the output of
// the qkv layer as it lies in memory), plus the row statistics the backward needs (lse fp32 [B, H, 256], base-2 log-sum-exp of the
// scaled scores).  planes: workspace, fp16 [B * T, 2 * 3 * H * 64].
extern "C" int qvit_attention_train_fwd(const float* qkv, int B, int T, int H, int head_dim, float scale, void* planes, float* out,
                                        float* lse, qvit_stream_t stream) {
  QVIT_REQUIRE(qkv && planes && out && lse, "qvit_attention_train_fwd: null pointer");
  int rc = train_shape_ok("qvit_attention_train_fwd", B, T, H, head_dim);
  if (rc) return rc;
  const int D3 = 3 * H * head_dim;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(lse, 0, sizeof(float) * (size_t)B * H * ab::kStat, s) != cudaSuccess) {
    set_error("qvit_attention_train_fwd: cudaMemsetAsync failed");
    return QVIT_ERR_CUDA;
  }
  rc = qvit_split2_f16(qkv, (int64_t)B * T, D3, D3, nullptr, planes, 2 * (int64_t)D3, D3, nullptr, stream);
  if (rc) return rc;
  return attention_f16x2_launch(planes, 2 * (int64_t)D3, D3, B, T, H, head_dim, scale, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, out,
                                nullptr, nullptr, lse, s);
}

// Backward of the attention core: dqkv (fp32, laid out like qkv; every element is written) from qkv, the forward's out / lse and
// dout = d loss / d out (fp32 [B, T, H, 64]).  dstat: workspace fp32 [B, H, 256].
extern "C" int qvit_attention_train_bwd(const float* qkv, const float* out, const float* dout, const float* lse, int B, int T, int H,
                                        int head_dim, float scale, float* dstat, float* dqkv, qvit_stream_t stream) {
  return qvit_attention_train_bwd_prof(qkv, out, dout, lse, B, T, H, head_dim, scale, dstat, dqkv, nullptr, stream);
}

// same, with the cycle stamps of CTA 0's first 8 tiles (prof: int64 [256]; tools/att_bwd_prof.py prints the timeline)
extern "C" int qvit_attention_train_bwd_prof(const float* qkv, const float* out, const float* dout, const float* lse, int B, int T, int H,
                                             int head_dim, float scale, float* dstat, float* dqkv, long long* prof, qvit_stream_t stream) {
  QVIT_REQUIRE(qkv && out && dout && lse && dstat && dqkv, "qvit_attention_train_bwd: null pointer");
  int rc = train_shape_ok("qvit_attention_train_bwd", B, T, H, head_dim);
  if (rc) return rc;
  QVIT_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) |
                 reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0, "qvit_attention_train_bwd: 16-byte aligned tensors");
  int dev = 0, maj = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  if (maj != 10) {
    set_error("qvit_attention_train_bwd: needs sm_100 (tcgen05)");
    return QVIT_ERR_UNSUPPORTED;
  }
  static bool attr_set[64] = {false};
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ab::kSmem);
    if (e != cudaSuccess) {
      set_error("qvit_attention_train_bwd: cudaFuncSetAttribute(%d): %s", ab::kSmem, cudaGetErrorString(e));
      return QVIT_ERR_CUDA;
    }
    attr_set[dev] = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t warps = (int64_t)B * ab::kStat;
  attention_dstat_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(out, dout, dstat, B, T, H);
  rc = check_launch("qvit_attention_train_bwd (dstat)");
  if (rc) return rc;
  const int64_t items = 2 * (int64_t)B * H;
  QVIT_REQUIRE(items < (1ll << 30), "qvit_attention_train_bwd: problem too large");
  const int grid = (int)(items < sm_count() ? items : sm_count());
  attention_bwd_kernel<<<grid, ab::kThreads, ab::kSmem, s>>>(qkv, dout, lse, dstat, dqkv, T, H, (int)items, scale, prof);
  return check_launch("qvit_attention_train_bwd");
}
