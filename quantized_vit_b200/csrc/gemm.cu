// qvit_gemm_i8: dispatch between the tcgen05 tensor-core kernel (gemm_tc.cu) and a CUDA-core dp4a kernel
// that accepts ANY pitch / alignment (pruned checkpoints give arbitrary dims) and
// doubles as the on-device cross-check of the tensor-core path.  Both share epilogue.cuh.
#include "common.cuh"
#include "epilogue.cuh"

namespace qvit {

bool gemm_tc_supported(const void* a, int64_t lda, const void* w, int64_t ldw, int M, int N, int K);
int gemm_tc_launch(const void* a, int64_t lda, int a_unsigned, const int8_t* w, int64_t ldw, const EpiParams& ep, int K,
                   cudaStream_t s);

void gemm_tc_force_cta_group(int cg);
int gemm_tc_read_profile(long long* host, int n);

constexpr int kST = 64;        // SIMT tile (rows and cols)
constexpr int kSK = 64;        // bytes of K per step

__device__ __forceinline__ int dp4a_ss(int a, int b, int c) { return __dp4a(a, b, c); }
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <bool A_UNSIGNED>
__global__ void __launch_bounds__(256)
gemm_i8_simt_kernel(const uint8_t* __restrict__ A, int64_t lda, const int8_t* __restrict__ W, int64_t ldw, int K,
                    const EpiParams ep) {
  __shared__ uint32_t As[kST][kSK / 4 + 1];
  __shared__ uint32_t Ws[kST][kSK / 4 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * kST;
  const int n0 = blockIdx.x * kST;
  int acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;

  for (int k0 = 0; k0 < K; k0 += kSK) {
    // each thread fills 4 words of A and 4 words of W (byte-wise: no alignment assumption)
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = threadIdx.x + it * 256;          // 0..1023 words
      const int r = idx >> 4, wcol = idx & 15;
      uint32_t va = 0, vw = 0;
      const int kk = k0 + wcol * 4;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (kk + b < K) {
          if (m0 + r < ep.M) va |= (uint32_t)A[(m0 + r) * lda + kk + b] << (8 * b);
          if (n0 + r < ep.N) vw |= (uint32_t)(uint8_t)W[(int64_t)(n0 + r) * ldw + kk + b] << (8 * b);
        }
      }
      As[r][wcol] = va;
      Ws[r][wcol] = vw;
    }
    __syncthreads();
#pragma unroll 4
    for (int w = 0; w < kSK / 4; ++w) {
      uint32_t a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][w];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[tx * 4 + j][w];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] = A_UNSIGNED ? dp4a_us(a[i], (int)b[j], acc[i][j]) : dp4a_ss((int)a[i], (int)b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float scale = epi_scale(ep);
  SymParams nq;
  if (ep.out_kind == QVIT_OUT_I8) nq = load_sym_params(ep.next_d, ep.next_qm, ep.next_t);
  int fl = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= ep.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < ep.N) epi_store_one(ep, &nq, acc[i][j], scale, m, n, fl);
    }
  }
  fl = warp_or(fl);
  if (fl && ep.flags && (threadIdx.x & 31) == 0) atomicOr(ep.flags, fl);
}

// K == 0: y = act(bias) + residual
__global__ void gemm_k0_kernel(const EpiParams ep) {
  const float scale = epi_scale(ep);
  SymParams nq;
  if (ep.out_kind == QVIT_OUT_I8) nq = load_sym_params(ep.next_d, ep.next_qm, ep.next_t);
  int fl = 0;
  const int64_t total = (int64_t)ep.M * ep.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    epi_store_one(ep, &nq, 0, scale, i / ep.N, (int)(i % ep.N), fl);
  fl = warp_or(fl);
  if (fl && ep.flags && (threadIdx.x & 31) == 0) atomicOr(ep.flags, fl);
}

}  // namespace qvit

using namespace qvit;

extern "C" int qvit_gemm_read_profile(long long* host, int n) { return gemm_tc_read_profile(host, n); }

extern "C" int qvit_gemm_set_cta_group(int cta_group) {
  QVIT_REQUIRE(cta_group >= 0 && cta_group % 10 <= 2 && (cta_group / 10) % 10 <= 9 && cta_group < 300,
               "qvit_gemm_set_cta_group: 0 (auto), 1 or 2; +10 / +20 / +30 / +40 select benchmark modes (see gemm_tc.cu)");
  gemm_tc_force_cta_group(cta_group);
  return QVIT_OK;
}

extern "C" int qvit_gemm_i8(const void* a, int64_t lda, int a_unsigned, const int8_t* w, int64_t ldw, int M, int N, int K,
                            void* out, int64_t ldo, const qvit_epilogue_t* epi, int backend, qvit_stream_t stream) {
  QVIT_REQUIRE(epi != nullptr, "qvit_gemm_i8: epilogue descriptor is NULL");
  QVIT_REQUIRE(M >= 0 && N >= 0 && K >= 0, "qvit_gemm_i8: negative dimension");
  if (M == 0 || N == 0) return QVIT_OK;
  QVIT_REQUIRE(epi->out_kind == QVIT_OUT_NONE || (out != nullptr && ldo >= N), "qvit_gemm_i8: bad output (ldo=%lld < N=%d?)",
               (long long)ldo, N);
  QVIT_REQUIRE(K == 0 || (a != nullptr && w != nullptr && lda >= K && ldw >= K), "qvit_gemm_i8: bad operands");
  QVIT_REQUIRE(epi->out_kind >= QVIT_OUT_I32 && epi->out_kind <= QVIT_OUT_F16X2, "qvit_gemm_i8: bad out_kind %d", epi->out_kind);
  QVIT_REQUIRE(epi->out_kind != QVIT_OUT_F16X2 || (N % 32 == 0 && ldo % 16 == 0 && ldo / 2 >= N && !epi->residual),
               "qvit_gemm_i8: QVIT_OUT_F16X2 needs N %% 32 == 0, ldo %% 16 == 0, ldo / 2 >= N and no residual");
  QVIT_REQUIRE(epi->act >= QVIT_ACT_NONE && epi->act <= QVIT_ACT_RELU, "qvit_gemm_i8: bad act %d", epi->act);
  QVIT_REQUIRE(epi->out_kind != QVIT_OUT_I8 || (epi->next_d && epi->next_qm), "qvit_gemm_i8: QVIT_OUT_I8 needs next_d/next_qm");
  QVIT_REQUIRE(!epi->residual || epi->ld_res >= N, "qvit_gemm_i8: ld_res < N");
  EpiParams ep;
  ep.out_kind = epi->out_kind;
  ep.act = epi->act;
  ep.scale_const = epi->scale_const;
  ep.acc_abs_max = epi->acc_abs_max;
  ep.scale_a = epi->scale_a;
  ep.scale_w = epi->scale_w;
  ep.col_scale = epi->col_scale;
  ep.bias = epi->bias;
  ep.residual = epi->residual;
  ep.ld_res = epi->ld_res;
  ep.next_d = epi->next_d;
  ep.next_qm = epi->next_qm;
  ep.next_t = epi->next_t;
  ep.flags = epi->flags;
  ep.out = out;
  ep.ldo = ldo;
  ep.M = M;
  ep.N = N;
  cudaStream_t s = (cudaStream_t)stream;
  if (epi->out_kind == QVIT_OUT_NONE && backend != QVIT_GEMM_TCGEN05) {
    set_error("qvit_gemm_i8: QVIT_OUT_NONE is a tensor-core benchmark mode (backend must be QVIT_GEMM_TCGEN05)");
    return QVIT_ERR_INVALID;
  }
  if (K == 0) {
    gemm_k0_kernel<<<div_up((int64_t)M * N, 256) > 1184 ? 1184 : div_up((int64_t)M * N, 256), 256, 0, s>>>(ep);
    return check_launch("gemm_k0_kernel");
  }
  const bool tc_ok = gemm_tc_supported(a, lda, w, ldw, M, N, K);
  if (backend == QVIT_GEMM_TCGEN05 && !tc_ok) {
    set_error("qvit_gemm_i8: tcgen05 backend needs sm_100, 16-byte aligned bases and lda/ldw multiples of 16");
    return QVIT_ERR_UNSUPPORTED;
  }
  if (backend == QVIT_GEMM_TCGEN05 || (backend == QVIT_GEMM_AUTO && tc_ok))
    return gemm_tc_launch(a, lda, a_unsigned, w, ldw, ep, K, s);
  QVIT_REQUIRE(backend == QVIT_GEMM_AUTO || backend == QVIT_GEMM_SIMT, "qvit_gemm_i8: bad backend %d", backend);
  dim3 grid((unsigned)div_up(N, kST), (unsigned)div_up(M, kST));
  QVIT_REQUIRE(grid.y <= 65535u, "qvit_gemm_i8: SIMT backend supports M <= %d", 65535 * kST);
  if (a_unsigned)
    gemm_i8_simt_kernel<true><<<grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(a), lda, w, ldw, K, ep);
  else
    gemm_i8_simt_kernel<false><<<grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(a), lda, w, ldw, K, ep);
  return check_launch("gemm_i8_simt_kernel");
}
