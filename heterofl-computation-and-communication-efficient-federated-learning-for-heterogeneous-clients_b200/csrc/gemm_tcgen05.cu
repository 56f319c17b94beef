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
#include "gemm_epilogue.cuh"

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
  static constexpr int TOTAL = GEMM_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + BLOCK_N * 4 /*bias tile*/;
};


#ifdef HFL_GEMM_TIMING
// phase timestamps of CTA (0,0) (clock64), read back with hfl_gemm_timing(): 0 start, 1 after setup sync, 2 first TMA
// issued, 3 last TMA issued, 4 first full barrier seen, 5 last MMA committed, 6 accumulator ready (epilogue start),
// 7 epilogue end, 8 before dealloc
__device__ long long g_gemm_t[16];
#define GEMM_T(i) do { if (blockIdx.x == 0 && blockIdx.y == 0) g_gemm_t[i] = clock64(); } while (0)
#else
#define GEMM_T(i) do { } while (0)
#endif

// BIAS: 0 none, 1 fp32 bias added after alpha, 2 bf16 bias added before alpha.  The epilogue variants are compile-time:
// the four epilogue warps run one per scheduler with nothing to hide latency behind, so a runtime `gelu` flag (which
// the compiler if-converts into predicated erff code for every element) cost 10 000 cycles per tile.
template <int BLOCK_N, bool OUT_BF16, int BIAS, bool GELU>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    void* __restrict__ C, const float* __restrict__ bias, const __nv_bfloat16* __restrict__ bias_h,
                    float alpha, int M, int N, int K, int ldc) {
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
  if (threadIdx.x == 0) GEMM_T(0);

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
  pdl_sync();   // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel
  if (threadIdx.x == 0) GEMM_T(1);

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
        if (kb == 0) GEMM_T(2);
      }
      GEMM_T(3);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, BLOCK_N);
    for (int kb = 0; kb < num_k; ++kb) {
      const int s = kb % GEMM_STAGES;
      const uint32_t ph = (kb / GEMM_STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      if (lane == 0 && kb == 0) GEMM_T(4);
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
        if (kb == num_k - 1) GEMM_T(5);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp - 4;                                // TMEM lane quadrant
    float* sbias = reinterpret_cast<float*>(smem + GEMM_STAGES * S::STAGE_BYTES + 256);
    if (BIAS != 0) {
      // the tile's bias row goes to shared memory while the mainloop runs: in the element loop every thread (= one
      // accumulator row) needs all BLOCK_N values, which from global memory was a chain of dependent loads
      for (int t = threadIdx.x - 128; t < BLOCK_N; t += 128) {
        const int col = n_blk * BLOCK_N + t;
        sbias[t] = col < N ? (BIAS == 1 ? bias[col] : __bfloat162float(bias_h[col])) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");       // epilogue warps only
    }
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    if (threadIdx.x == 128) GEMM_T(6);
    auto xf = [&](float acc, int jc) -> float {              // jc = column inside the tile
      float x = acc;
      if (BIAS == 2) x += sbias[jc];                          // alpha * (acc + b)
      x *= alpha;
      if (BIAS == 1) x += sbias[jc];                          // alpha * acc + b
      if (GELU) x = 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
      return x;
    };
    epilogue_store_tile<BLOCK_N, OUT_BF16>(tmem_base, smem, q, lane, C, ldc, m_blk * GEMM_BLOCK_M, n_blk * BLOCK_N, M, N,
                                           xf);
    tcgen05_fence_before();
    if (threadIdx.x == 128) GEMM_T(7);
  }
  __syncthreads();
  if (threadIdx.x == 0) GEMM_T(8);
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
  // split-K: gridDim.z CTAs share one output tile, each reducing a contiguous range of k-blocks and accumulating its
  // fp32 partial into a zero-filled C with atomics (skinny problems: dX of the vocabulary projection is 64 x 256
  // with K = 33 278 -- two CTAs streaming 520 k-blocks each otherwise)
  extern __shared__ uint8_t smem_raw[];
  using S = GemmSmem<BLOCK_N>;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GEMM_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;
  const int num_k_all = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int per_split = (num_k_all + gridDim.z - 1) / gridDim.z;
  const int kb_lo = blockIdx.z * per_split;
  const int num_k = min(per_split, num_k_all - kb_lo);
  if (num_k <= 0) return;                                  // whole CTA: nothing allocated yet
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
  pdl_sync();   // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel
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
          tma_load_2d(sa, &tmap_a, &full_bar[s], m_blk * GEMM_BLOCK_M, (kb_lo + kb) * GEMM_BLOCK_K);
          tma_load_2d(sa + 8192, &tmap_a, &full_bar[s], m_blk * GEMM_BLOCK_M + 64, (kb_lo + kb) * GEMM_BLOCK_K);
        } else {
          tma_load_2d(sa, &tmap_a, &full_bar[s], (kb_lo + kb) * GEMM_BLOCK_K, m_blk * GEMM_BLOCK_M);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[s], n_blk * BLOCK_N + j * 64, (kb_lo + kb) * GEMM_BLOCK_K);
        } else {
          tma_load_2d(sb, &tmap_b, &full_bar[s], (kb_lo + kb) * GEMM_BLOCK_K, n_blk * BLOCK_N);
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
    auto xf = [&](float acc, int) -> float { return acc * alpha; };
    if (!OUT_BF16 && gridDim.z > 1)
      epilogue_store_tile<BLOCK_N, false, true>(tmem_base, smem, q, lane, C, ldc, m_blk * GEMM_BLOCK_M, n_blk * BLOCK_N,
                                                M, N, xf);
    else
      epilogue_store_tile<BLOCK_N, OUT_BF16>(tmem_base, smem, q, lane, C, ldc, m_blk * GEMM_BLOCK_M, n_blk * BLOCK_N, M,
                                             N, xf);
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
                               int ldc, int splits, cudaStream_t stream) {
  auto kern = gemm_tcgen05_general_kernel<128, A_MN, B_MN, OUT_BF16>;
  constexpr int smem = GemmSmem<128>::TOTAL;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M, (N + 127) / 128, splits);
  launch_k(kern, dim3(grid), dim3(GEMM_THREADS), smem, stream, ta, tb, C, alpha, M, N, K, ldc);
  count_launch();
  return (int)cudaGetLastError();
}

// General layouts.  a_mn: A is stored [K, M] (lda = row stride in elements) instead of [M, K]; b_mn: B is stored
// [K, N] instead of [N, K].  All leading dimensions and the contiguous extents must be multiples of 8 elements.
// How many K-splits a problem should use: only skinny grids with a long reduction (each split keeps >= 8 k-blocks).
extern "C" int hfl_gemm_general_splits(int M, int N, int K) {
  const int tiles = ((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M) * ((N + 127) / 128);
  const int num_k = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  if (tiles >= 32 || num_k < 32) return 1;
  int s = 148 / tiles;
  if (s > num_k / 8) s = num_k / 8;
  return s < 1 ? 1 : s;
}

// splits > 1: C must be fp32 and zero-filled (partials are accumulated with atomics).
extern "C" int hfl_gemm_bf16_general(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc,
                                     int out_bf16, float alpha, int M, int N, int K, int splits, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (splits < 1) splits = 1;
  if (splits > 1 && out_bf16) return -4;
  CUtensorMap ta, tb;
  int r = a_mn ? make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, GEMM_BLOCK_K)
               : make_tmap_2d_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, GEMM_BLOCK_K, GEMM_BLOCK_M);
  if (r) return 1000 + r;
  r = b_mn ? make_tmap_2d_bf16(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, GEMM_BLOCK_K)
           : make_tmap_2d_bf16(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, GEMM_BLOCK_K, 128);
  if (r) return 2000 + r;
#define HFL_GG(AM, BM)                                                                                       \
  if ((a_mn != 0) == AM && (b_mn != 0) == BM)                                                                \
    return out_bf16 ? launch_gemm_general<AM, BM, true>(ta, tb, C, alpha, M, N, K, ldc, splits, stream)      \
                    : launch_gemm_general<AM, BM, false>(ta, tb, C, alpha, M, N, K, ldc, splits, stream);
  HFL_GG(false, false)
  HFL_GG(false, true)
  HFL_GG(true, false)
  HFL_GG(true, true)
#undef HFL_GG
  return -1;
}

template <int BLOCK_N, bool OUT_BF16, int BIAS, bool GELU>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, void* C, const float* bias,
                       const __nv_bfloat16* bias_h, float alpha, int M, int N, int K, int ldc, cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BLOCK_N, OUT_BF16, BIAS, GELU>;
  constexpr int smem = GemmSmem<BLOCK_N>::TOTAL;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M, (N + BLOCK_N - 1) / BLOCK_N);
  launch_k(kern, dim3(grid), dim3(GEMM_THREADS), smem, stream, ta, tb, C, bias, bias_h, alpha, M, N, K, ldc);
  count_launch();
  return (int)cudaGetLastError();
}

// A: [M, K] bf16 row-major (lda elements), B: [N, K] bf16 row-major (ldb elements).  Returns a cudaError / CUresult code.
// bias (fp32) is added after alpha; bias_h (bf16, e.g. straight from a parameter arena) is added before it.
extern "C" int hfl_gemm_bf16_tn(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int out_bf16,
                                const float* bias, const void* bias_h_, float alpha, int gelu, int M, int N, int K,
                                cudaStream_t stream) {
  const __nv_bfloat16* bias_h = reinterpret_cast<const __nv_bfloat16*>(bias_h_);
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int block_n = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  CUtensorMap ta, tb;
  int r = make_tmap_2d_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, GEMM_BLOCK_K, GEMM_BLOCK_M);
  if (r) return 1000 + r;
  r = make_tmap_2d_bf16(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, GEMM_BLOCK_K, block_n);
  if (r) return 1000 + r;
  const int bias_mode = bias_h != nullptr ? 2 : (bias != nullptr ? 1 : 0);
#define HFL_GEMM_EPI(BN, OB, BI, GE)                                                                          \
  if (bias_mode == BI && (gelu != 0) == GE)                                                                   \
    return launch_gemm<BN, OB, BI, GE>(ta, tb, C, bias, bias_h, alpha, M, N, K, ldc, stream);
#define HFL_GEMM_OUT(BN, OB)                                                                                  \
  HFL_GEMM_EPI(BN, OB, 0, false) HFL_GEMM_EPI(BN, OB, 1, false) HFL_GEMM_EPI(BN, OB, 2, false)                \
  HFL_GEMM_EPI(BN, OB, 0, true) HFL_GEMM_EPI(BN, OB, 1, true) HFL_GEMM_EPI(BN, OB, 2, true)
#define HFL_GEMM_CASE(BN)                                                                                     \
  if (block_n == BN) {                                                                                        \
    if (out_bf16) { HFL_GEMM_OUT(BN, true) } else { HFL_GEMM_OUT(BN, false) }                                 \
  }
  HFL_GEMM_CASE(32)
  HFL_GEMM_CASE(64)
  HFL_GEMM_CASE(128)
#undef HFL_GEMM_EPI
#undef HFL_GEMM_OUT
#undef HFL_GEMM_CASE
  return -1;
}

#ifdef HFL_GEMM_TIMING
extern "C" void hfl_gemm_timing(long long* out16) { cudaMemcpyFromSymbol(out16, hfl::g_gemm_t, sizeof(long long) * 16); }
#else
extern "C" void hfl_gemm_timing(long long* out16) { for (int i = 0; i < 16; ++i) out16[i] = 0; }
#endif
