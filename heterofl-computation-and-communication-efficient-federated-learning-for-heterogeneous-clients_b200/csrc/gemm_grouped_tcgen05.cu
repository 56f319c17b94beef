// Grouped tcgen05 GEMMs for the Transformer's q / k / v projections (sm_100a).
//
// The federated masked-LM step is 64 tokens: its 78 linear-layer GEMMs are launch-latency bound (5 us each for
// 64 x 256 x 256), so the three projections that share an operand are issued as ONE launch
// (reference call sites: src/models/transformer.py:59-62, 80-84):
//
//   SEPARATE (gridDim.z = groups):  C_g = alpha_g * (A_g . B_g^T + bias_g)      forward  q,k,v = x Wq^T, x Wk^T, x Wv^T
//                                                                              backward dWq,dWk,dWv = dq^T x, dk^T x, dv^T x
//   ACCUM    (gridDim.z = 1):       C   = sum_g A_g . B_g^T                     backward dx = dq Wq + dk Wk + dv Wv
//
// Same pipeline as gemm_tcgen05.cu (TMA SWIZZLE_128B -> 4-stage mbarrier ring -> tcgen05.mma M128 x N128 x K16 ->
// TMEM -> smem-transposed epilogue); either operand may be MN-major so the backward forms read their transposed operands
// in place.  The per-group tensor maps / pointers travel in one __grid_constant__ struct.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "gemm_epilogue.cuh"

namespace hfl {

constexpr int GG_BLOCK_M = 128, GG_BLOCK_N = 128, GG_BLOCK_K = 64, GG_STAGES = 4, GG_THREADS = 256, GG_MAX = 3;
constexpr int GG_A_BYTES = GG_BLOCK_M * GG_BLOCK_K * 2, GG_B_BYTES = GG_BLOCK_N * GG_BLOCK_K * 2;
constexpr int GG_STAGE_BYTES = GG_A_BYTES + GG_B_BYTES;
constexpr int GG_SMEM = GG_STAGES * GG_STAGE_BYTES + 1024 + 256 + GG_BLOCK_N * 4;

struct GroupArgs {
  CUtensorMap ta[GG_MAX];
  CUtensorMap tb[GG_MAX];
  void* C[GG_MAX];
  const __nv_bfloat16* bias[GG_MAX];
  float alpha[GG_MAX];
  int ngroups;
};

template <bool A_MN, bool B_MN, bool ACCUM, bool BIAS>
__global__ void __launch_bounds__(GG_THREADS, 1)
gemm_grouped_tcgen05_kernel(const __grid_constant__ GroupArgs G, int M, int N, int K, int ldc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GG_STAGES * GG_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GG_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GG_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;
  const int g0 = ACCUM ? 0 : (int)blockIdx.z;
  const int g1 = ACCUM ? G.ngroups : g0 + 1;
  const int num_k = (K + GG_BLOCK_K - 1) / GG_BLOCK_K;
  const int total_k = num_k * (g1 - g0);
  if (warp == 0 && lane == 0)
    for (int g = g0; g < g1; ++g) { tma_prefetch_desc(&G.ta[g]); tma_prefetch_desc(&G.tb[g]); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, GG_BLOCK_N);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();   // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel
  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int g = g0; g < g1; ++g) {
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % GG_STAGES;
          const uint32_t ph = (it / GG_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * GG_STAGE_BYTES;
          uint8_t* sb = sa + GG_A_BYTES;
          mbar_expect_tx(&full_bar[s], GG_STAGE_BYTES);
          if (A_MN) {   // two 64(m) x 64(k) boxes of the [K, M] matrix
            tma_load_2d(sa, &G.ta[g], &full_bar[s], m_blk * GG_BLOCK_M, kb * GG_BLOCK_K);
            tma_load_2d(sa + 8192, &G.ta[g], &full_bar[s], m_blk * GG_BLOCK_M + 64, kb * GG_BLOCK_K);
          } else {
            tma_load_2d(sa, &G.ta[g], &full_bar[s], kb * GG_BLOCK_K, m_blk * GG_BLOCK_M);
          }
          if (B_MN) {
            tma_load_2d(sb, &G.tb[g], &full_bar[s], n_blk * GG_BLOCK_N, kb * GG_BLOCK_K);
            tma_load_2d(sb + 8192, &G.tb[g], &full_bar[s], n_blk * GG_BLOCK_N + 64, kb * GG_BLOCK_K);
          } else {
            tma_load_2d(sb, &G.tb[g], &full_bar[s], kb * GG_BLOCK_K, n_blk * GG_BLOCK_N);
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(GG_BLOCK_M, GG_BLOCK_N, A_MN, B_MN);
    for (int it = 0; it < total_k; ++it) {
      const int s = it % GG_STAGES;
      const uint32_t ph = (it / GG_STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * GG_STAGE_BYTES);
        const uint32_t b_addr = a_addr + GG_A_BYTES;
        const uint64_t adesc = A_MN ? make_smem_desc_sw128_mn(a_addr) : make_smem_desc_sw128(a_addr);
        const uint64_t bdesc = B_MN ? make_smem_desc_sw128_mn(b_addr) : make_smem_desc_sw128(b_addr);
        constexpr uint64_t a_step = A_MN ? (2048 >> 4) : (32 >> 4);
        constexpr uint64_t b_step = B_MN ? (2048 >> 4) : (32 >> 4);
#pragma unroll
        for (int k = 0; k < GG_BLOCK_K / 16; ++k)
          umma_bf16(tmem_base, adesc + a_step * k, bdesc + b_step * k, idesc, (it | k) != 0);
        umma_commit(&empty_bar[s]);
        if (it == total_k - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    float* sbias = reinterpret_cast<float*>(smem + GG_STAGES * GG_STAGE_BYTES + 256);
    if (BIAS) {
      const __nv_bfloat16* bp = G.bias[g0];
      for (int t = threadIdx.x - 128; t < GG_BLOCK_N; t += 128) {
        const int col = n_blk * GG_BLOCK_N + t;
        sbias[t] = (col < N && bp != nullptr) ? __bfloat162float(bp[col]) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const float alpha = G.alpha[g0];
    auto xf = [&](float acc, int jc) -> float { return BIAS ? (acc + sbias[jc]) * alpha : acc * alpha; };
    epilogue_store_tile<GG_BLOCK_N, true>(tmem_base, smem, q, lane, G.C[g0], ldc, m_blk * GG_BLOCK_M, n_blk * GG_BLOCK_N,
                                          M, N, xf);
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, GG_BLOCK_N);
  }
}

}  // namespace hfl

using namespace hfl;

template <bool A_MN, bool B_MN, bool ACCUM, bool BIAS>
static int launch_grouped(const GroupArgs& G, int M, int N, int K, int ldc, cudaStream_t stream) {
  auto kern = gemm_grouped_tcgen05_kernel<A_MN, B_MN, ACCUM, BIAS>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GG_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid((M + GG_BLOCK_M - 1) / GG_BLOCK_M, (N + GG_BLOCK_N - 1) / GG_BLOCK_N, ACCUM ? 1 : G.ngroups);
  launch_k(kern, dim3(grid), dim3(GG_THREADS), GG_SMEM, stream, G, M, N, K, ldc);
  count_launch();
  return (int)cudaGetLastError();
}

// A_g: [M, K] (a_mn = 0, lda) or [K, M] (a_mn = 1); B_g: [N, K] (b_mn = 0, ldb) or [K, N] (b_mn = 1); C (bf16) [M, N], ldc.
// accum = 1: one output C[0] = alpha[0] * sum_g A_g B_g^T;  accum = 0: C_g = alpha_g * (A_g B_g^T + bias_g).
extern "C" int hfl_gemm_grouped(int ngroups, const void* const* A, int lda, int a_mn, const void* const* B, int ldb,
                                int b_mn, void* const* C, int ldc, const void* const* bias, const float* alpha, int accum,
                                int M, int N, int K, cudaStream_t stream) {
  if (ngroups < 1 || ngroups > GG_MAX) return -5;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  GroupArgs G;
  G.ngroups = ngroups;
  bool any_bias = false;
  for (int g = 0; g < GG_MAX; ++g) {
    const int s = g < ngroups ? g : 0;
    int r = a_mn ? make_tmap_2d_bf16(&G.ta[g], A[s], (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, GG_BLOCK_K)
                 : make_tmap_2d_bf16(&G.ta[g], A[s], (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, GG_BLOCK_K, GG_BLOCK_M);
    if (r) return 1000 + r;
    r = b_mn ? make_tmap_2d_bf16(&G.tb[g], B[s], (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, GG_BLOCK_K)
             : make_tmap_2d_bf16(&G.tb[g], B[s], (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, GG_BLOCK_K, GG_BLOCK_N);
    if (r) return 2000 + r;
    G.C[g] = C[accum ? 0 : s];
    G.bias[g] = bias != nullptr ? reinterpret_cast<const __nv_bfloat16*>(bias[s]) : nullptr;
    G.alpha[g] = alpha[accum ? 0 : s];
    any_bias = any_bias || G.bias[g] != nullptr;
  }
  if (accum && any_bias) return -6;
#define HFL_GGC(AM, BM)                                                                                   \
  if ((a_mn != 0) == AM && (b_mn != 0) == BM) {                                                           \
    if (accum) return launch_grouped<AM, BM, true, false>(G, M, N, K, ldc, stream);                       \
    if (any_bias) return launch_grouped<AM, BM, false, true>(G, M, N, K, ldc, stream);                    \
    return launch_grouped<AM, BM, false, false>(G, M, N, K, ldc, stream);                                 \
  }
  HFL_GGC(false, false)
  HFL_GGC(false, true)
  HFL_GGC(true, true)
#undef HFL_GGC
  return -1;
}
