// Large-shape tcgen05 GEMM with cluster-level operand sharing:   C[M,N] = act(alpha * (A[M,K] . B[N,K]^T + bias))
//
// The single-CTA kernel (gemm_tcgen05.cu) is bound by L2 -> SM operand traffic on big problems: a 128 x 128 tile pulls
// 32 KB per 64-deep k-block, 11 TB/s over the whole chip at 8192 x 4096 x 4096, and the tensor pipe idles 60 % of the
// time (profiles/prof_gemm.md).  This variant halves the bytes per FLOP:
//
//   * 128 x 256 tiles (UMMA M128 x N256 x K16, 256 TMEM columns): one A tile feeds twice the math;
//   * thread-block clusters of 2 CTAs along M: both CTAs need the SAME 256 x 64 B tile, so each one loads only half of it
//     and the TMA unit multicasts that half into both CTAs' shared memory (`.multicast::cluster`, mask 0b11);
//     per CTA and k-block: 16 KB (own A) + 16 KB (half B) for 128 x 256 x 64 of MMA, i.e. 2x the intensity.
//
// Barrier protocol (4-stage ring, same roles as the single-CTA kernel: warp 0 TMA, warp 1 MMA, warp 2 TMEM, 4-7 epilogue):
//   full[s]   count 1 + tx: each CTA's producer arms its own barrier with the FULL stage byte count; the peer's
//             multicast half signals the same barrier offset in this CTA.
//   empty[s]  count 2: a stage is rewritten by BOTH producers (own A + own half-B, peer's half-B), so it may only be
//             re-armed when both CTAs' MMA warps are done with it -- `tcgen05.commit ... multicast::cluster` arrives on
//             the barrier at the same offset in both CTAs.
//   cluster barrier after barrier init (no remote arrive may hit an uninitialised mbarrier) and before exit (no CTA may
//   retire while its peer can still signal into its shared memory).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "gemm_epilogue.cuh"

namespace hfl {

constexpr int MC_BLOCK_M = 128;
constexpr int MC_BLOCK_N = 256;
constexpr int MC_BLOCK_K = 64;
constexpr int MC_STAGES = 4;
constexpr int MC_THREADS = 256;
constexpr int MC_A_BYTES = MC_BLOCK_M * MC_BLOCK_K * 2;      // 16 KB
constexpr int MC_B_BYTES = MC_BLOCK_N * MC_BLOCK_K * 2;      // 32 KB (each CTA loads 16 KB of it)
constexpr int MC_STAGE_BYTES = MC_A_BYTES + MC_B_BYTES;
constexpr int MC_SMEM = MC_STAGES * MC_STAGE_BYTES + 1024 + 256 + MC_BLOCK_N * 4;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

template <int BIAS, bool GELU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MC_THREADS, 1)
gemm_tcgen05_mc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       __nv_bfloat16* __restrict__ C, const float* __restrict__ bias,
                       const __nv_bfloat16* __restrict__ bias_h, float alpha, int M, int N, int K, int ldc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + MC_STAGES * MC_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MC_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MC_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* sbias = reinterpret_cast<float*>(smem + MC_STAGES * MC_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 / 1 inside the CTA pair
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;      // the pair (2p, 2p+1) shares n_blk
  const int num_k = (K + MC_BLOCK_K - 1) / MC_BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 2); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, MC_BLOCK_N);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                    // peer barriers are initialised before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();   // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % MC_STAGES;
        const uint32_t ph = (kb / MC_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);                // both CTAs have drained this stage
        uint8_t* sa = smem + s * MC_STAGE_BYTES;
        uint8_t* sb = sa + MC_A_BYTES;
        mbar_expect_tx(&full_bar[s], MC_STAGE_BYTES);    // own A + own half-B + the peer's half-B
        tma_load_2d(sa, &tmap_a, &full_bar[s], kb * MC_BLOCK_K, m_blk * MC_BLOCK_M);
        tma_load_2d_mc(sb + rank * (MC_B_BYTES / 2), &tmap_b, &full_bar[s], kb * MC_BLOCK_K,
                       n_blk * MC_BLOCK_N + (int)rank * (MC_BLOCK_N / 2), (uint16_t)0x3);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(MC_BLOCK_M, MC_BLOCK_N);
    for (int kb = 0; kb < num_k; ++kb) {
      const int s = kb % MC_STAGES;
      const uint32_t ph = (kb / MC_STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * MC_STAGE_BYTES);
        const uint64_t adesc = make_smem_desc_sw128(a_addr);
        const uint64_t bdesc = make_smem_desc_sw128(a_addr + MC_A_BYTES);
#pragma unroll
        for (int k = 0; k < MC_BLOCK_K / 16; ++k)
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        umma_commit_mc(&empty_bar[s], (uint16_t)0x3);    // frees the stage in BOTH CTAs once these MMAs retire
        if (kb == num_k - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    if (BIAS != 0) {
      for (int t = threadIdx.x - 128; t < MC_BLOCK_N; t += 128) {
        const int col = n_blk * MC_BLOCK_N + t;
        sbias[t] = col < N ? (BIAS == 1 ? bias[col] : __bfloat162float(bias_h[col])) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    // the peer may still be multicasting k-blocks into OUR pipeline stages (its MMA can lag ours), so the staging area
    // of the transposed write-out must not alias them: wait until the peer's last load has landed = our full barriers
    // have all completed (they have: tmem_full fires after the MMAs that consumed every stage) -- and the peer issues
    // no load for a stage before both CTAs freed it, i.e. nothing is in flight once our last MMA retired.
    auto xf = [&](float acc, int jc) -> float {
      float x = acc;
      if (BIAS == 2) x += sbias[jc];
      x *= alpha;
      if (BIAS == 1) x += sbias[jc];
      if (GELU) x = 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
      return x;
    };
    epilogue_store_tile<MC_BLOCK_N, true>(tmem_base, smem, q, lane, C, ldc, m_blk * MC_BLOCK_M, n_blk * MC_BLOCK_N, M, N,
                                          xf);
    tcgen05_fence_before();
  }
  __syncthreads();
  cluster_sync_all();                                    // nobody retires while the peer can still signal its barriers
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, MC_BLOCK_N);
  }
}

}  // namespace hfl

using namespace hfl;

template <int BIAS, bool GELU>
static int launch_mc(const CUtensorMap& ta, const CUtensorMap& tb, void* C, const float* bias, const __nv_bfloat16* bias_h,
                     float alpha, int M, int N, int K, int ldc, cudaStream_t stream) {
  auto kern = gemm_tcgen05_mc_kernel<BIAS, GELU>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MC_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int tiles_m = (M + MC_BLOCK_M - 1) / MC_BLOCK_M;
  dim3 grid((tiles_m + 1) / 2 * 2, (N + MC_BLOCK_N - 1) / MC_BLOCK_N);     // x is a multiple of the cluster size
  launch_k(kern, dim3(grid), dim3(MC_THREADS), MC_SMEM, stream, ta, tb, (__nv_bfloat16*)C, bias, bias_h, alpha, M, N, K, ldc);
  count_launch();
  return (int)cudaGetLastError();
}

// 1 when the cluster kernel should take this problem: big enough that operand traffic, not latency, is the bound
// (measured: 8192x4096x4096 790 -> 1250 TFLOP/s, but 1000x768x320 7.3 -> 10.5 us; bench/gemm_mc_check.py)
extern "C" int hfl_gemm_mc_wanted(int M, int N, int K) { return (M >= 1024 && N >= 1024 && K >= 256) ? 1 : 0; }

// bf16 output only.  A: [M, K] row-major (lda), B: [N, K] row-major (ldb); bias (fp32, after alpha) or bias_h (bf16,
// before alpha) or neither.
extern "C" int hfl_gemm_bf16_tn_mc(const void* A, int lda, const void* B, int ldb, void* C, int ldc, const float* bias,
                                   const void* bias_h_, float alpha, int gelu, int M, int N, int K, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const __nv_bfloat16* bias_h = reinterpret_cast<const __nv_bfloat16*>(bias_h_);
  CUtensorMap ta, tb;
  int r = make_tmap_2d_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, MC_BLOCK_K, MC_BLOCK_M);
  if (r) return 1000 + r;
  r = make_tmap_2d_bf16(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, MC_BLOCK_K, MC_BLOCK_N / 2);
  if (r) return 2000 + r;
  const int mode = bias_h != nullptr ? 2 : (bias != nullptr ? 1 : 0);
#define HFL_MC(BI, GE) \
  if (mode == BI && (gelu != 0) == GE) return launch_mc<BI, GE>(ta, tb, C, bias, bias_h, alpha, M, N, K, ldc, stream);
  HFL_MC(0, false) HFL_MC(1, false) HFL_MC(2, false) HFL_MC(0, true) HFL_MC(1, true) HFL_MC(2, true)
#undef HFL_MC
  return -1;
}
