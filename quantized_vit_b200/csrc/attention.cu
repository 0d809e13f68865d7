// softmax(Q K^T * scale) V in fp32-equivalent precision on the tcgen05 tensor cores ("3xTF32").
//
// This is glue BESIDE the hot path: the reference never quantizes the attention core
// (vit_model.py:141-149) and runs it in fp32, and the 4-bit quantizer that follows (`proj`) turns any 1e-4-level
// error of its input into flipped activation codes.  A bf16 or single-pass TF32 kernel is therefore not usable;
// instead every fp32 operand x is split exactly into x = hi + lo (hi = upper 19 bits, a valid TF32 number;
// lo = x - hi, again truncated to TF32) and each product a*b is evaluated as a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with
// fp32 accumulation in TMEM: relative error ~2^-21, at tensor-core speed.
//
// One CTA = one (batch, head) and one tile of 128 queries; keys/values of the whole sequence (T <= 208) are resident.
//   phase 1  TMA loads Q [128 x 64] and K [208 x 64] fp32 (128B-swizzled 32-float sub-tiles, rows beyond T zero-filled)
//   phase 2  all threads split Q, K into hi / lo planes in shared memory
//   phase 3  S = Q K^T: 8 k-steps x 3 terms of tcgen05.mma kind::tf32 (M=128, N=208, K=8), SS mode, accumulator in TMEM
//   phase 4  softmax over the row held by each thread (thread = query row, two warps share a row's columns);
//            P = exp2((S - max) * scale * log2e) is written back to TMEM as P_hi (over S) and P_lo;
//            meanwhile TMA loads the raw V tile into the (dead) Q buffers
//   phase 5  transpose + split V into K-major V^T planes; O = P V: 26 k-steps x 3 terms, A = P from TMEM (TS mode)
//   phase 6  O / rowsum -> global [B, T, H*64]
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace qvit {

constexpr int kAttHd = 64;            // head dim (floats): 2 sub-tiles of 32 floats = 128 B rows
constexpr int kAttMQ = 128;           // queries per CTA
constexpr int kAttNK = 208;           // keys per CTA (13 x 16): UMMA N of the S tile, K extent of the PV product
constexpr int kAttThreads = 256;
constexpr int kQSub = kAttMQ * 128;   // bytes of one Q sub-tile  [128 rows x 128 B]
constexpr int kKSub = kAttNK * 128;   // bytes of one K/V sub-tile [208 rows x 128 B] (26 KiB, multiple of 1024)
constexpr int kVtSub = kAttHd * 128;  // bytes of one V^T sub-tile [64 head-dim rows x 32 keys]
constexpr int kVtSubs = 7;            // 7 x 32 = 224 >= 208 keys
constexpr int kPlane = kVtSubs * kVtSub;              // 56 KiB: holds a K plane (52 KiB) or a V^T plane
constexpr int kAttSmem = 4 * kQSub + 2 * kPlane + 4096 + 1024;  // Q hi/lo | K or V^T hi | lo | misc | alignment slack
static_assert(2 * kKSub <= kPlane && 2 * kKSub <= 4 * kQSub, "plane sizes");

// round-to-nearest TF32 (10-bit mantissa) and the exact remainder, itself rounded to TF32
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rna(x);
  lo = tf32_rna(x - __uint_as_float(hi));
}

// tcgen05.mma kind::tf32, A and B from shared memory
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// A from tensor memory (lane = row, column = k), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
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

// MN-major operand (V: keys = K dimension are the rows, head-dim floats contiguous), 128B swizzle.
// LBO = byte distance between the two 32-float column atoms, SBO = byte distance between 8-row (K) atoms.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool b_mn_major) {
  return (1u << 4) /*D = f32*/ | (2u << 7) /*A = tf32*/ | (2u << 10) /*B = tf32*/ | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// hi/lo split of a plane of fp32 values in shared memory (position preserving, so the TMA swizzle is irrelevant)
__device__ __forceinline__ void split_plane(uint8_t* hi_plane, uint8_t* lo_plane, int bytes) {
  for (int off = threadIdx.x * 16; off < bytes; off += kAttThreads * 16) {
    const float4 v = *reinterpret_cast<const float4*>(hi_plane + off);
    uint4 h, l;
    tf32_split(v.x, h.x, l.x);
    tf32_split(v.y, h.y, l.y);
    tf32_split(v.z, h.z, l.z);
    tf32_split(v.w, h.w, l.w);
    *reinterpret_cast<uint4*>(hi_plane + off) = h;
    *reinterpret_cast<uint4*>(lo_plane + off) = l;
  }
}

// V arrives as [key rows x 64 floats] (two 128B-swizzled sub-tiles of 32 floats); the PV product wants B = V^T K-major:
// [64 head-dim rows x keys], as 128B-swizzled sub-tiles of 32 keys.  A warp owns 32 head-dim rows and 4 consecutive
// keys per step: the four reads each sweep one (permuted) 128-byte row, the 128-bit writes hit 8 distinct 16-byte
// chunks per quarter warp - both free of bank conflicts.
__device__ __forceinline__ void transpose_split_v(const uint8_t* raw, uint8_t* hi_plane, uint8_t* lo_plane) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int item = warp; item < 2 * (kAttNK / 4); item += kAttThreads / 32) {
    const int hsub = item / (kAttNK / 4), kg = item - hsub * (kAttNK / 4);
    const int hd = hsub * 32 + lane;
    uint4 h, l;
    uint32_t* hp = &h.x;
    uint32_t* lp = &l.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int key = 4 * kg + i;
      const float x = *reinterpret_cast<const float*>(raw + hsub * kKSub + key * 128 + (((lane >> 2) ^ (key & 7)) << 4) +
                                                      ((lane & 3) << 2));
      tf32_split(x, hp[i], lp[i]);
    }
    const int off = (kg >> 3) * kVtSub + hd * 128 + (((kg & 7) ^ (hd & 7)) << 4);
    *reinterpret_cast<uint4*>(hi_plane + off) = h;
    *reinterpret_cast<uint4*>(lo_plane + off) = l;
  }
}

__global__ void __launch_bounds__(kAttThreads, 1)
attention_f32_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     float* __restrict__ out, int T, int H, float scale_log2e, float* __restrict__ dbg, int diag) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  // layout: Q_hi (2 sub-tiles) | Q_lo (2) | plane_hi | plane_lo | misc.  The planes hold K (hi/lo) for S = Q K^T and
  // later V^T (hi/lo) for O = P V; the raw V tile lands in the (by then dead) Q region.
  const uint32_t q_hi = base, q_lo = base + 2 * kQSub, kv_hi = base + 4 * kQSub, kv_lo = kv_hi + kPlane;
  uint8_t* g_q_hi = gen;
  uint8_t* g_q_lo = gen + 2 * kQSub;
  uint8_t* g_kv_hi = gen + 4 * kQSub;
  uint8_t* g_kv_lo = g_kv_hi + kPlane;
  uint8_t* misc = g_kv_lo + kPlane;
  const uint32_t misc_a = kv_lo + kPlane;
  const uint32_t bar_qk = misc_a, bar_v = misc_a + 8, bar_s = misc_a + 16, bar_o = misc_a + 24;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(misc + 32);
  float* red_max = reinterpret_cast<float*>(misc + 64);          // [2][128]
  float* red_sum = red_max + 2 * kAttMQ;                         // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_tile * kAttMQ;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_kv);
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_v, 1);
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
  const uint32_t t_s = tmem, t_plo = tmem + kAttNK, t_o = tmem + 2 * kAttNK;

  // ---- phase 1: Q and K tiles (column index in floats: q at (0*H + h)*64, k at (1*H + h)*64, v at (2*H + h)*64)
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(bar_qk, 2 * kQSub + 2 * kKSub);
    for (int s = 0; s < 2; ++s) {
      tma_load_3d(q_hi + s * kQSub, &tmap_q, bar_qk, (0 * H + h) * kAttHd + s * 32, q0, b);
      tma_load_3d(kv_hi + s * kKSub, &tmap_kv, bar_qk, (1 * H + h) * kAttHd + s * 32, 0, b);
    }
  }
  ptx::mbar_wait(bar_qk, 0);

  // ---- phase 2: hi / lo planes
  split_plane(g_q_hi, g_q_lo, 2 * kQSub);
  split_plane(g_kv_hi, g_kv_lo, 2 * kKSub);
  ptx::fence_proxy_async_smem();
  __syncthreads();

  // ---- phase 3: S = Q K^T  (3 x TF32)
  if (threadIdx.x == 32) {
    ptx::tc_fence_after();
    constexpr uint32_t idesc = make_idesc_tf32(kAttMQ, kAttNK, false);
    uint32_t acc = 0;
#pragma unroll 1
    for (int term = 0; term < 3; ++term) {
      const uint32_t a_base = (term == 1) ? q_lo : q_hi;       // hi*hi, lo*hi, hi*lo
      const uint32_t b_base = (term == 2) ? kv_lo : kv_hi;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {                         // 8 floats (32 B) of the head dim per MMA
        const uint32_t sub = ks >> 2, kk = ks & 3;
        const uint64_t a_desc = ptx::make_kmajor_sw128_desc(a_base + sub * kQSub + kk * 32);
        const uint64_t b_desc = ptx::make_kmajor_sw128_desc(b_base + sub * kKSub + kk * 32);
        mma_tf32_ss(t_s, a_desc, b_desc, idesc, acc);
        acc = 1;
      }
    }
    ptx::mma_commit(bar_s);
  }
  // the raw V tile may overwrite the Q planes as soon as the S MMAs have retired
  if (threadIdx.x == 0) {
    ptx::mbar_wait(bar_s, 0);
    ptx::mbar_expect_tx(bar_v, 2 * kKSub);
    for (int s = 0; s < 2; ++s) tma_load_3d(q_hi + s * kKSub, &tmap_kv, bar_v, (2 * H + h) * kAttHd + s * 32, 0, b);
  }
  ptx::mbar_wait(bar_s, 0);
  ptx::tc_fence_after();

  // ---- phase 4: softmax.  thread = row (lane quarter = warp & 3), column half = warp >> 2 (104 columns each)
  const int row = (warp & 3) * 32 + lane;
  const int half = warp >> 2;
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
  const int col0 = half * (kAttNK / 2);
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {                                // 3 chunks of 32 columns ...
    uint32_t r[32];
    ptx::tmem_ld_32x32(t_s + lane_addr + (uint32_t)(col0 + c * 32), r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + c * 32 + j < T) mx = fmaxf(mx, __uint_as_float(r[j]));
    if (dbg) {
      float* d = dbg + ((((int64_t)b * H + h) * 256 + q0 + ((warp & 3) * 32 + lane)) * 512) + col0 + c * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) d[j] = __uint_as_float(r[j]);
    }
  }
  {                                                            // ... and a tail of 8 (104 = 3 * 32 + 8)
    uint32_t r[8];
    tmem_ld_32x8(t_s + lane_addr + (uint32_t)(col0 + 96), r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (col0 + 96 + j < T) mx = fmaxf(mx, __uint_as_float(r[j]));
    if (dbg) {
      float* d = dbg + ((((int64_t)b * H + h) * 256 + q0 + ((warp & 3) * 32 + lane)) * 512) + col0 + 96;
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = __uint_as_float(r[j]);
    }
  }
  red_max[half * kAttMQ + row] = mx;
  __syncthreads();
  mx = fmaxf(red_max[row], red_max[kAttMQ + row]);
  const float mbias = mx * scale_log2e;
  float sum = 0.f;
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    const int cbase = col0 + c * 32;
    uint32_t r[32], lo[32];
    ptx::tmem_ld_32x32(t_s + lane_addr + (uint32_t)cbase, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float p = exp2f(fmaf(__uint_as_float(r[j]), scale_log2e, -mbias));
      if (cbase + j >= T) p = 0.f;
      sum += p;
      tf32_split(p, r[j], lo[j]);
      if (dbg) dbg[((((int64_t)b * H + h) * 256 + q0 + row) * 512) + 208 + cbase + j] = p;
    }
    tmem_st_32x32(t_s + lane_addr + (uint32_t)cbase, r);
    tmem_st_32x32(t_plo + lane_addr + (uint32_t)cbase, lo);
  }
  {
    const int cbase = col0 + 96;
    uint32_t r[8], lo[8];
    tmem_ld_32x8(t_s + lane_addr + (uint32_t)cbase, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p = exp2f(fmaf(__uint_as_float(r[j]), scale_log2e, -mbias));
      if (cbase + j >= T) p = 0.f;
      sum += p;
      tf32_split(p, r[j], lo[j]);
      if (dbg) dbg[((((int64_t)b * H + h) * 256 + q0 + row) * 512) + 208 + cbase + j] = p;
    }
    tmem_st_32x8(t_s + lane_addr + (uint32_t)cbase, r);
    tmem_st_32x8(t_plo + lane_addr + (uint32_t)cbase, lo);
  }
  tmem_st_wait();
  red_sum[half * kAttMQ + row] = sum;

  // ---- phase 5: V planes, then O = P V
  ptx::mbar_wait(bar_v, 0);
  transpose_split_v(g_q_hi, g_kv_hi, g_kv_lo);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 32 && diag == 1) {
    // diagnostic: O_diag[m][n] = sum_{k<64} P_hi[m][k] * Q_hi[n][k]  (TS mode against a K-major operand known to work)
    ptx::tc_fence_after();
    constexpr uint32_t idesc = make_idesc_tf32(kAttMQ, kAttHd, false);
    uint32_t acc = 0;
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t sub = ks >> 2, kk = ks & 3;
      const uint64_t b_desc = ptx::make_kmajor_sw128_desc(q_hi + sub * kQSub + kk * 32);
      mma_tf32_ts(t_o, t_s + (uint32_t)(ks * 8), b_desc, idesc, acc);
      acc = 1;
    }
    ptx::mma_commit(bar_o);
  } else if (threadIdx.x == 32) {
    ptx::tc_fence_after();
    constexpr uint32_t idesc = make_idesc_tf32(kAttMQ, kAttHd, false);
    uint32_t acc = 0;
#pragma unroll 1
    for (int term = 0; term < 3; ++term) {
      const uint32_t a_t = (term == 1) ? t_plo : t_s;          // P_hi*V_hi, P_lo*V_hi, P_hi*V_lo
      const uint32_t b_base = (term == 2) ? kv_lo : kv_hi;
#pragma unroll 2
      for (int ks = 0; ks < kAttNK / 8; ++ks) {                // 8 keys (32 B of a V^T row) per MMA
        const uint64_t b_desc = ptx::make_kmajor_sw128_desc(b_base + (ks >> 2) * kVtSub + (ks & 3) * 32);
        mma_tf32_ts(t_o, a_t + (uint32_t)(ks * 8), b_desc, idesc, acc);
        acc = 1;
      }
    }
    ptx::mma_commit(bar_o);
  }
  ptx::mbar_wait(bar_o, 0);
  ptx::tc_fence_after();

  // ---- phase 6: normalise and store.  warps 0-3: head-dim 0..31, warps 4-7: 32..63
  {
    const float inv = __fdiv_rn(1.0f, red_sum[row] + red_sum[kAttMQ + row]);
    uint32_t r[32];
    ptx::tmem_ld_32x32(t_o + lane_addr + (uint32_t)(half * 32), r);
    ptx::tmem_ld_wait();
    if (dbg) {
      float* d = dbg + ((((int64_t)b * H + h) * 256 + q0 + row) * 512) + 416;
#pragma unroll
      for (int j = 0; j < 32; ++j) d[half * 32 + j] = __uint_as_float(r[j]);
      if (half == 0) { d[64] = red_sum[row]; d[65] = red_sum[kAttMQ + row]; d[66] = inv; }
    }
    const int t = q0 + row;
    if (t < T) {
      float* dst = out + ((int64_t)b * T + t) * ((int64_t)H * kAttHd) + h * kAttHd + half * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        stg_v4_b32(dst + 4 * j, __float_as_uint(__uint_as_float(r[4 * j]) * inv), __float_as_uint(__uint_as_float(r[4 * j + 1]) * inv),
                   __float_as_uint(__uint_as_float(r[4 * j + 2]) * inv), __float_as_uint(__uint_as_float(r[4 * j + 3]) * inv));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn att_encode_fn() {
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

}  // namespace qvit

using namespace qvit;

static int attention_launch(const float* qkv, int B, int T, int H, int head_dim, float scale, float* out, float* dbg,
                            int diag, qvit_stream_t stream) {
  QVIT_REQUIRE(qkv && out && B > 0 && T > 0 && H > 0, "qvit_attention_f32: bad argument");
  if (head_dim != kAttHd || T > kAttNK) {
    set_error("qvit_attention_f32: supports head_dim == 64 and T <= 208 (got head_dim=%d, T=%d)", head_dim, T);
    return QVIT_ERR_UNSUPPORTED;
  }
  QVIT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "qvit_attention_f32: pointers must be 16-byte aligned");
  int dev = 0, maj = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  EncodeTiledFn enc = att_encode_fn();
  if (maj != 10 || !enc) {
    set_error("qvit_attention_f32: needs sm_100 and cuTensorMapEncodeTiled");
    return QVIT_ERR_UNSUPPORTED;
  }
  const uint64_t row_floats = 3ull * H * kAttHd;
  CUtensorMap tq, tkv;
  cuuint64_t dims[3] = {row_floats, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {row_floats * 4, row_floats * 4 * (cuuint64_t)T};
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint32_t box_q[3] = {32, (cuuint32_t)kAttMQ, 1};
  cuuint32_t box_kv[3] = {32, (cuuint32_t)kAttNK, 1};
  CUresult r1 = enc(&tq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(qkv), dims, strides, box_q, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(qkv), dims, strides, box_kv, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    set_error("qvit_attention_f32: cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
    return QVIT_ERR_CUDA;
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
  dim3 grid((unsigned)((T + kAttMQ - 1) / kAttMQ), (unsigned)H, (unsigned)B);
  attention_f32_kernel<<<grid, kAttThreads, kAttSmem, (cudaStream_t)stream>>>(tq, tkv, out, T, H,
                                                                             scale * 1.4426950408889634f, dbg, diag);
  return check_launch("qvit_attention_f32");
}

extern "C" int qvit_attention_f32(const float* qkv, int B, int T, int H, int head_dim, float scale, float* out,
                                  qvit_stream_t stream) {
  return attention_launch(qkv, B, T, H, head_dim, scale, out, nullptr, 0, stream);
}

// test hook: additionally dumps raw scores S (cols 0..207) and un-normalised probabilities P (cols 208..415) per query row
// into dbg [B, H, 256, 512] fp32 (cols 416..479 raw O, 480..482 row sums / 1/sum)
extern "C" int qvit_attention_f32_debug(const float* qkv, int B, int T, int H, int head_dim, float scale, float* out,
                                        float* dbg, int diag, qvit_stream_t stream) {
  return attention_launch(qkv, B, T, H, head_dim, scale, out, dbg, diag, stream);
}
