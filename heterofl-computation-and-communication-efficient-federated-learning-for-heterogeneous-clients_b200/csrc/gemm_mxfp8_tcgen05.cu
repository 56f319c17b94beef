// Block-scaled FP8 (MXFP8: e4m3 elements, one ue8m0 scale per 32 elements along K) GEMM on the 5th-gen tensor cores.
//
//   C[M, N] = alpha * (A[M, K] . B[N, K]^T) (+ bias)        A, B: bf16 in HBM, quantised on the fly by mx_quantize
//
// Reference call sites: the Transformer's Linear layers (src/models/transformer.py:59-62,94-96,125-129); BASELINE.json
// names the MXFP8 path as the reduced-precision mode of the language-model benchmark.  Two kernels:
//
//   mx_quantize_kernel  bf16 [R, K] -> e4m3 [R, K] + ue8m0 scale factors, OCP-MX rule: shared exponent
//                       floor(log2(amax_32)) - 8 (8 = emax of e4m3), elements saturate-cast after the division.
//                       Scale factors are written directly in the layout tcgen05 reads them in: one 512-byte chunk per
//                       (128-row tile, 128-element k-block), byte (r % 32) * 16 + (r / 32) * 4 + (k / 32) % 4.
//   gemm_mxfp8_kernel   128 x 128 x 128 tiles; warp 0 = TMA producer (two 16 KB SWIZZLE_128B operand boxes + two 512-byte
//                       scale-factor chunks by cp.async.bulk per stage), warp 1 = issuer: `tcgen05.cp.32x128b.warpx4`
//                       moves the scale factors of the stage into TMEM (4 columns each), then four
//                       `tcgen05.mma.kind::mxf8f6f4.block_scale` (K = 32) select their scale byte through the sf_id
//                       fields of the instruction descriptor; warps 4-7 = epilogue (tcgen05.ld -> smem -> 16-byte
//                       stores, shared with the bf16 GEMM).  fp32 accumulation in TMEM.
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "gemm_epilogue.cuh"

namespace hfl {

constexpr int MX_BM = 128, MX_BN = 128, MX_BK = 128;      // BK in elements = bytes
constexpr int MX_STAGES = 5;
constexpr int MX_THREADS = 256;
constexpr int MX_A_BYTES = MX_BM * MX_BK, MX_B_BYTES = MX_BN * MX_BK, MX_SF_BYTES = 512;
constexpr int MX_STAGE_BYTES = MX_A_BYTES + MX_B_BYTES + 2 * MX_SF_BYTES;      // 33 792 (keeps 1024-byte alignment)
constexpr int MX_SMEM = MX_STAGES * MX_STAGE_BYTES + 1024 + 256 + MX_BN * 4;
constexpr int MX_SF_SLOTS = 4;                            // TMEM scale-factor buffers (k-block % 4)
constexpr uint32_t MX_TMEM_COLS = 256;                    // 128 accumulator + 4 x (4 + 4) scale-factor columns

// ---------------------------------------------------------------------------------------------------------
// quantiser: one thread = one 32-element block
// ---------------------------------------------------------------------------------------------------------
__global__ void mx_quantize_kernel(const __nv_bfloat16* __restrict__ x, int R, int K, int ldx, uint8_t* __restrict__ q,
                                   int ldq, uint8_t* __restrict__ sf, int num_kb) {
  pdl_sync();
  const int kblocks = K / 32;
  const long total = (long)R * kblocks;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / kblocks), kb32 = (int)(i - (long)r * kblocks);
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)r * ldx + (size_t)kb32 * 32);
    float v[32];
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 u = src[j];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[j * 8 + 2 * e] = __low2float(h[e]);
        v[j * 8 + 2 * e + 1] = __high2float(h[e]);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(v[j]));
    // shared exponent: floor(log2(amax)) - emax(e4m3 = 8); biased ue8m0 byte = exponent + 127
    int e = amax > 0.f ? (int)((__float_as_uint(amax) >> 23) & 0xFF) - 127 - 8 : -127;
    e = max(-127, min(127, e));
    const float inv = __uint_as_float((uint32_t)(127 - e) << 23);      // 2^-e (e in [-127, 127] -> exponent 0..254)
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t p = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        p |= (uint32_t)__nv_cvt_float_to_fp8(v[j * 4 + b] * inv, __NV_SATFINITE, __NV_E4M3) << (8 * b);
      w[j] = p;
    }
    uint4* dst = reinterpret_cast<uint4*>(q + (size_t)r * ldq + (size_t)kb32 * 32);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    const int rt = r >> 7, rr = r & 127, kb = kb32 >> 2, ks = kb32 & 3;
    sf[((size_t)rt * num_kb + kb) * MX_SF_BYTES + (rr & 31) * 16 + (rr >> 5) * 4 + ks] = (uint8_t)(e + 127);
  }
}

// ---------------------------------------------------------------------------------------------------------
// PTX helpers specific to the block-scaled path
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// 32 rows x 16 bytes, rows 16 B apart (no swizzle): core matrices of 8 rows = 128 B, SBO = 128 B
__device__ __forceinline__ uint64_t make_smem_desc_sf(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;       // leading byte offset (single core matrix along K: unused)
  d |= (uint64_t)(128 >> 4) << 32;      // stride byte offset between 8-row core matrices
  d |= (uint64_t)1 << 46;               // descriptor version
  return d;                             // layout type 0 = no swizzle
}
__device__ __forceinline__ void tmem_cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// cute::UMMA::InstrDescriptorBlockScaled: e4m3 x e4m3, ue8m0 scales, K-major operands, dense, K = 32
__device__ __forceinline__ uint32_t make_idesc_mxf8(int M, int N, uint32_t a_sf_id, uint32_t b_sf_id) {
  return (b_sf_id << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24) |
         (a_sf_id << 29);
}
__device__ __forceinline__ void umma_mxf8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t tsfa,
                                          uint32_t tsfb, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(tsfa), "r"(tsfb)
      : "memory");
}

template <bool OUT_BF16, bool BIAS>
__global__ void __launch_bounds__(MX_THREADS, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const uint8_t* __restrict__ sfa, const uint8_t* __restrict__ sfb, void* __restrict__ C,
                  const float* __restrict__ bias, float alpha, int M, int N, int K, int ldc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + MX_STAGES * MX_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MX_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* sbias = reinterpret_cast<float*>(smem + MX_STAGES * MX_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;
  const int num_k = K / MX_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, MX_TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();   // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel

  if (warp == 0) {
    if (lane == 0) {
      const uint8_t* ga = sfa + (size_t)m_blk * num_k * MX_SF_BYTES;
      const uint8_t* gb = sfb + (size_t)n_blk * num_k * MX_SF_BYTES;
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % MX_STAGES;
        const uint32_t ph = (kb / MX_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * MX_STAGE_BYTES;
        uint8_t* sb = sa + MX_A_BYTES;
        uint8_t* ssa = sb + MX_B_BYTES;
        mbar_expect_tx(&full_bar[s], MX_STAGE_BYTES);
        tma_load_2d(sa, &tmap_a, &full_bar[s], kb * MX_BK, m_blk * MX_BM);
        tma_load_2d(sb, &tmap_b, &full_bar[s], kb * MX_BK, n_blk * MX_BN);
        bulk_load_1d(ssa, ga + (size_t)kb * MX_SF_BYTES, MX_SF_BYTES, &full_bar[s]);
        bulk_load_1d(ssa + MX_SF_BYTES, gb + (size_t)kb * MX_SF_BYTES, MX_SF_BYTES, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    for (int kb = 0; kb < num_k; ++kb) {
      const int s = kb % MX_STAGES;
      const uint32_t ph = (kb / MX_STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * MX_STAGE_BYTES);
        const uint32_t b_addr = a_addr + MX_A_BYTES;
        const uint32_t sf_addr = b_addr + MX_B_BYTES;
        const uint32_t tsfa = tmem_base + MX_BN + (uint32_t)(kb % MX_SF_SLOTS) * 8;
        const uint32_t tsfb = tsfa + 4;
        tmem_cp_32x128b_warpx4(tsfa, make_smem_desc_sf(sf_addr));
        tmem_cp_32x128b_warpx4(tsfb, make_smem_desc_sf(sf_addr + MX_SF_BYTES));
        const uint64_t adesc = make_smem_desc_sw128(a_addr);
        const uint64_t bdesc = make_smem_desc_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < MX_BK / 32; ++k)          // 32 fp8 = 32 bytes along K: +2 in the (addr >> 4) field
          umma_mxf8(tmem_base, adesc + 2 * k, bdesc + 2 * k, make_idesc_mxf8(MX_BM, MX_BN, k, k), tsfa, tsfb,
                    (kb | k) != 0);
        umma_commit(&empty_bar[s]);
        if (kb == num_k - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    if (BIAS) {
      for (int t = threadIdx.x - 128; t < MX_BN; t += 128) {
        const int col = n_blk * MX_BN + t;
        sbias[t] = col < N ? bias[col] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    auto xf = [&](float acc, int jc) -> float { return BIAS ? acc * alpha + sbias[jc] : acc * alpha; };
    epilogue_store_tile<MX_BN, OUT_BF16>(tmem_base, smem, q, lane, C, ldc, m_blk * MX_BM, n_blk * MX_BN, M, N, xf);
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, MX_TMEM_COLS);
  }
}

inline int make_tmap_2d_u8(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                           uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return 1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 100 + (int)r;
}

}  // namespace hfl

using namespace hfl;

extern "C" int hfl_mx_sf_bytes(int R, int K) { return ((R + 127) / 128) * (K / 128) * MX_SF_BYTES; }

// x: bf16 [R, K] (row stride ldx elements, K % 128 == 0, rows 16-byte aligned); q: e4m3 [R, K] (row stride ldq bytes);
// sf: hfl_mx_sf_bytes(R, K) bytes, ZERO-FILLED by the caller when R % 128 != 0 (padding rows keep scale 2^-127)
extern "C" int hfl_mx_quantize(const void* x, int R, int K, int ldx, void* q, int ldq, void* sf, cudaStream_t st) {
  if (K % 128 != 0 || ldx % 8 != 0 || ldq % 16 != 0) return -1;
  const long total = (long)R * (K / 32);
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(mx_quantize_kernel, dim3((int)blocks), dim3(256), 0, st, (const __nv_bfloat16*)x, R, K, ldx, (uint8_t*)q, ldq, (uint8_t*)sf,
                                                 K / 128);
  count_launch();
  return (int)cudaGetLastError();
}

// C[M, N] = alpha * Aq . Bq^T (+ bias); Aq [M, K] / Bq [N, K] e4m3 with row strides lda / ldb bytes
extern "C" int hfl_gemm_mxfp8(const void* Aq, int lda, const void* sfa, const void* Bq, int ldb, const void* sfb, void* C,
                              int ldc, int out_bf16, const float* bias, float alpha, int M, int N, int K,
                              cudaStream_t st) {
  if (K % MX_BK != 0 || K <= 0) return -1;
  CUtensorMap ta, tb;
  int e = make_tmap_2d_u8(&ta, Aq, (uint64_t)K, (uint64_t)M, (uint64_t)lda, MX_BK, MX_BM);
  if (e) return e;
  e = make_tmap_2d_u8(&tb, Bq, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, MX_BK, MX_BN);
  if (e) return e;
  dim3 grid((M + MX_BM - 1) / MX_BM, (N + MX_BN - 1) / MX_BN);
  cudaError_t ce = cudaSuccess;
#define MX_LAUNCH(OB, BI)                                                                                      \
  do {                                                                                                         \
    auto kern = gemm_mxfp8_kernel<OB, BI>;                                                                     \
    ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MX_SMEM);                     \
    if (ce != cudaSuccess) return (int)ce;                                                                     \
    launch_k(kern, dim3(grid), dim3(MX_THREADS), MX_SMEM, st, ta, tb, (const uint8_t*)sfa, (const uint8_t*)sfb, C, bias, alpha, M, N, \
                                            K, ldc);                                                           \
  } while (0)
  if (out_bf16) { if (bias) MX_LAUNCH(true, true); else MX_LAUNCH(true, false); }
  else { if (bias) MX_LAUNCH(false, true); else MX_LAUNCH(false, false); }
#undef MX_LAUNCH
  count_launch();
  return (int)cudaGetLastError();
}
