// tcgen05 / TMEM / TMA GEMM for sm_100a:   C[M,N] = act(alpha * A[M,K] . B[N,K]^T + bias[N])
//
// A and B are bf16, K-major (row-major with K contiguous: activations [tokens, features] and nn.Linear weights
// [out, in] as they are), accumulation in fp32 in Tensor Memory, C is bf16 or fp32.
//
// Structure (one 128 x BLOCK_N output tile per CTA, BLOCK_K = 64 bf16 = one 128-byte swizzle row):
//   warp 0    : TMA producer   -- cp.async.bulk.tensor.2d (SWIZZLE_128B) into a 4-stage smem ring, mbarrier expect_tx
//   warp 1    : MMA issuer     -- one elected lane issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M128 x N x K16)
//                                 per stage from smem descriptors, tcgen05.commit releases the stage / signals the epilogue
//   warp 2    : TMEM allocator -- tcgen05.alloc / dealloc of BLOCK_N fp32 accumulator columns
//   warps 4-7 : epilogue       -- tcgen05.ld 32x32b (lane quadrant = warp % 4), alpha/bias/GELU, vectorised global stores
// The same producer/consumer skeleton (with 4-D shifted-box TMA loads as the A operand) is the implicit-GEMM
// convolution in conv_tcgen05.cu.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace hfl {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 256;

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;   // 16 KB
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = GEMM_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, bool OUT_BF16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    void* __restrict__ C, const float* __restrict__ bias, float alpha, int gelu, int M, int N, int K,
                    int ldc) {
  extern __shared__ uint8_t smem_raw[];
  using S = GemmSmem<BLOCK_N>;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GEMM_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;
  const int num_k = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % GEMM_STAGES;
        const uint32_t ph = (kb / GEMM_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * S::STAGE_BYTES;
        uint8_t* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        tma_load_2d(sa, &tmap_a, &full_bar[s], kb * GEMM_BLOCK_K, m_blk * GEMM_BLOCK_M);
        tma_load_2d(sb, &tmap_b, &full_bar[s], kb * GEMM_BLOCK_K, n_blk * BLOCK_N);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, BLOCK_N);
    for (int kb = 0; kb < num_k; ++kb) {
      const int s = kb % GEMM_STAGES;
      const uint32_t ph = (kb / GEMM_STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
        const uint64_t adesc = make_smem_desc_sw128(a_addr);
        const uint64_t bdesc = make_smem_desc_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);                       // frees the smem stage once these MMAs retire
        if (kb == num_k - 1) umma_commit(tmem_full_bar);  // accumulator complete -> epilogue
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp - 4;                                // TMEM lane quadrant
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int row = m_blk * GEMM_BLOCK_M + q * 32 + lane;
    const bool row_ok = row < M;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      const int col0 = n_blk * BLOCK_N + c0;
      if (row_ok) {
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) * alpha;
          if (bias != nullptr && col0 + j < N) x += bias[col0 + j];
          if (gelu) x = 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
          f[j] = x;
        }
        if (OUT_BF16) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(C) + (size_t)row * ldc + col0;
          if (col0 + 32 <= N && (((uintptr_t)dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
              h[0] = __floats2bfloat162_rn(f[j], f[j + 1]);
              h[1] = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
              h[2] = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
              h[3] = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(dst + j) = u;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) dst[j] = __float2bfloat16(f[j]);
          }
        } else {
          float* dst = reinterpret_cast<float*>(C) + (size_t)row * ldc + col0;
          if (col0 + 32 <= N && (((uintptr_t)dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) dst[j] = f[j];
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// General-layout variant: either operand may be MN-major (given as [K, M] resp. [K, N] row-major), which is what
// the backward GEMMs need without materialising transposes:
//     dX[M,K] = dY[M,N] . W[N,K]        -> B operand = W seen as [red=N, out=K]:  MN-major
//     dW[N,K] = dY[M,N]^T . X[M,K]      -> A = dY seen as [red=M, out=N], B = X seen as [red=M, out=K]: both MN-major
// C is fp32 or bf16 [M, N].  BLOCK_N = 128 (or 64), BLOCK_M = 128, BLOCK_K = 64.
// ---------------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool A_MN, bool B_MN, bool OUT_BF16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_general_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                            void* __restrict__ C, float alpha, int M, int N, int K, int ldc) {
  extern __shared__ uint8_t smem_raw[];
  using S = GemmSmem<BLOCK_N>;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GEMM_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;
  const int num_k = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % GEMM_STAGES;
        const uint32_t ph = (kb / GEMM_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * S::STAGE_BYTES;
        uint8_t* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        if (A_MN) {   // two 64(m) x 64(k) boxes of the [K, M] matrix
          tma_load_2d(sa, &tmap_a, &full_bar[s], m_blk * GEMM_BLOCK_M, kb * GEMM_BLOCK_K);
          tma_load_2d(sa + 8192, &tmap_a, &full_bar[s], m_blk * GEMM_BLOCK_M + 64, kb * GEMM_BLOCK_K);
        } else {
          tma_load_2d(sa, &tmap_a, &full_bar[s], kb * GEMM_BLOCK_K, m_blk * GEMM_BLOCK_M);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[s], n_blk * BLOCK_N + j * 64, kb * GEMM_BLOCK_K);
        } else {
          tma_load_2d(sb, &tmap_b, &full_bar[s], kb * GEMM_BLOCK_K, n_blk * BLOCK_N);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, BLOCK_N, A_MN, B_MN);
    for (int kb = 0; kb < num_k; ++kb) {
      const int s = kb % GEMM_STAGES;
      const uint32_t ph = (kb / GEMM_STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
        const uint64_t adesc = A_MN ? make_smem_desc_sw128_mn(a_addr) : make_smem_desc_sw128(a_addr);
        const uint64_t bdesc = B_MN ? make_smem_desc_sw128_mn(b_addr) : make_smem_desc_sw128(b_addr);
        // K advance of 16 elements: K-major +32 B (inside the swizzle row); MN-major +16 k-rows = 2048 B
        constexpr uint64_t a_step = A_MN ? (2048 >> 4) : (32 >> 4);
        constexpr uint64_t b_step = B_MN ? (2048 >> 4) : (32 >> 4);
#pragma unroll
        for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
          umma_bf16(tmem_base, adesc + a_step * k, bdesc + b_step * k, idesc, (kb | k) != 0);
        umma_commit(&empty_bar[s]);
        if (kb == num_k - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int row = m_blk * GEMM_BLOCK_M + q * 32 + lane;
    const bool row_ok = row < M;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      const int col0 = n_blk * BLOCK_N + c0;
      if (row_ok) {
        for (int j = 0; j < 32; ++j) {
          if (col0 + j >= N) break;
          const float x = __uint_as_float(v[j]) * alpha;
          if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(C)[(size_t)row * ldc + col0 + j] = __float2bfloat16(x);
          else reinterpret_cast<float*>(C)[(size_t)row * ldc + col0 + j] = x;
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace hfl

using namespace hfl;

template <bool A_MN, bool B_MN, bool OUT_BF16>
static int launch_gemm_general(const CUtensorMap& ta, const CUtensorMap& tb, void* C, float alpha, int M, int N, int K,
                               int ldc, cudaStream_t stream) {
  auto kern = gemm_tcgen05_general_kernel<128, A_MN, B_MN, OUT_BF16>;
  constexpr int smem = GemmSmem<128>::TOTAL;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M, (N + 127) / 128);
  kern<<<grid, GEMM_THREADS, smem, stream>>>(ta, tb, C, alpha, M, N, K, ldc);
  count_launch();
  return (int)cudaGetLastError();
}

// General layouts.  a_mn: A is stored [K, M] (lda = row stride in elements) instead of [M, K]; b_mn: B is stored
// [K, N] instead of [N, K].  All leading dimensions and the contiguous extents must be multiples of 8 elements.
extern "C" int hfl_gemm_bf16_general(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc,
                                     int out_bf16, float alpha, int M, int N, int K, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  CUtensorMap ta, tb;
  int r = a_mn ? make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, GEMM_BLOCK_K)
               : make_tmap_2d_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, GEMM_BLOCK_K, GEMM_BLOCK_M);
  if (r) return 1000 + r;
  r = b_mn ? make_tmap_2d_bf16(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, GEMM_BLOCK_K)
           : make_tmap_2d_bf16(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, GEMM_BLOCK_K, 128);
  if (r) return 2000 + r;
#define HFL_GG(AM, BM)                                                                                       \
  if ((a_mn != 0) == AM && (b_mn != 0) == BM)                                                                \
    return out_bf16 ? launch_gemm_general<AM, BM, true>(ta, tb, C, alpha, M, N, K, ldc, stream)              \
                    : launch_gemm_general<AM, BM, false>(ta, tb, C, alpha, M, N, K, ldc, stream);
  HFL_GG(false, false)
  HFL_GG(false, true)
  HFL_GG(true, false)
  HFL_GG(true, true)
#undef HFL_GG
  return -1;
}

template <int BLOCK_N, bool OUT_BF16>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, void* C, const float* bias, float alpha, int gelu,
                       int M, int N, int K, int ldc, cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BLOCK_N, OUT_BF16>;
  constexpr int smem = GemmSmem<BLOCK_N>::TOTAL;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M, (N + BLOCK_N - 1) / BLOCK_N);
  kern<<<grid, GEMM_THREADS, smem, stream>>>(ta, tb, C, bias, alpha, gelu, M, N, K, ldc);
  count_launch();
  return (int)cudaGetLastError();
}

// A: [M, K] bf16 row-major (lda elements), B: [N, K] bf16 row-major (ldb elements).  Returns a cudaError / CUresult code.
extern "C" int hfl_gemm_bf16_tn(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int out_bf16,
                                const float* bias, float alpha, int gelu, int M, int N, int K, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int block_n = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  CUtensorMap ta, tb;
  int r = make_tmap_2d_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, GEMM_BLOCK_K, GEMM_BLOCK_M);
  if (r) return 1000 + r;
  r = make_tmap_2d_bf16(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, GEMM_BLOCK_K, block_n);
  if (r) return 1000 + r;
#define HFL_GEMM_CASE(BN)                                                                                     \
  if (block_n == BN)                                                                                          \
    return out_bf16 ? launch_gemm<BN, true>(ta, tb, C, bias, alpha, gelu, M, N, K, ldc, stream)                \
                    : launch_gemm<BN, false>(ta, tb, C, bias, alpha, gelu, M, N, K, ldc, stream);
  HFL_GEMM_CASE(32)
  HFL_GEMM_CASE(64)
  HFL_GEMM_CASE(128)
#undef HFL_GEMM_CASE
  return -1;
}
