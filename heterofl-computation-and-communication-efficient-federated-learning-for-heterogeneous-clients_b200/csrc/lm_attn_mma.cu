// Tensor-core attention for the federated masked-LM step (S <= 64 keys, head dim <= 32), forward and backward.
//
// Reference (src/models/transformer.py:45-51): scores = q k^T / sqrt(d) -> softmax -> attn @ v, per (row, head).
// A client batch is ONE row of 64 tokens x 8 heads, so every contraction here is a 64 x 64 x 32 (or 64 x 32 x 64) tile:
// far below anything a tcgen05 pipeline can amortise (TMEM allocation + barrier set-up cost more than the math), which is
// why these kernels use the warp-level tensor-core path (mma.sync.m16n8k16, bf16 in / fp32 accumulate) on operands staged
// once in shared memory.  One CTA = one (row, head) pair, 4 warps x 16 query rows (backward: x 16 keys in its second
// half).  The softmax stays in the accumulator registers and the probabilities are re-used as the A operand of P.V
// without leaving the register file (the m16n8 accumulator layout is the m16n8k16 A layout).
//
//   forward : S = Q K^T (16 mma / warp) -> softmax -> P (bf16, saved) -> O = P V (16 mma / warp)
//   backward: dP = dO V^T -> dS = P o (dP - delta) * scale -> dQ = dS K ; then, keys across warps,
//             dK = dS^T Q, dV = P^T dO from shared-memory copies of dS and P; bias gradients = column sums.
// The previous CUDA-core kernels (lm_kernels.cu) remain as the reference implementation (HFL_LM_ATTN_MMA=0).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "common.cuh"

namespace hfl {

typedef __nv_bfloat16 bf16;
constexpr int AM_S = 64, AM_DH = 32;
constexpr int AM_PF = 40;      // feature-tile row pitch (bf16): 80 B, conflict-free for the 32-bit fragment loads
constexpr int AM_PS = 72;      // S x S tile row pitch (bf16)

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t am_pack2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t am_ld32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }
// two bf16 values from different rows -> one fragment register (low half = *p0)
__device__ __forceinline__ uint32_t am_pack_rows(const bf16* p0, const bf16* p1) {
  return (uint32_t)__bfloat16_as_ushort(*p0) | ((uint32_t)__bfloat16_as_ushort(*p1) << 16);
}
__device__ __forceinline__ float am_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float am_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// [S, dh] head slice of a [N*S, E] tensor -> zero-padded [64][AM_PF] shared tile (128 threads)
__device__ __forceinline__ void am_load_tile(bf16 (*dst)[AM_PF], const bf16* __restrict__ src, size_t base, int S, int E,
                                             int dh) {
  const bool vec = (dh % 8 == 0) && (E % 8 == 0);
  for (int it = threadIdx.x; it < AM_S * (AM_DH / 8); it += blockDim.x) {
    const int row = it >> 2, c8 = (it & 3) * 8;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (row < S && c8 < dh) {
      const bf16* p = src + base + (size_t)row * E + c8;
      if (vec) {
        u = *reinterpret_cast<const uint4*>(p);
      } else {
        __align__(16) bf16 tmp[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) tmp[e] = (c8 + e < dh) ? p[e] : __float2bfloat16(0.f);
        u = *reinterpret_cast<const uint4*>(tmp);
      }
    }
    *reinterpret_cast<uint4*>(&dst[row][c8]) = u;
  }
}

// A fragment (16 rows x 16 k) of a row-major [rows][AM_PF] tile: rows r0.., k = k0..
__device__ __forceinline__ void am_a_frag(uint32_t (&a)[4], bf16 (*t)[AM_PF], int r0, int k0, int g, int tq) {
  a[0] = am_ld32(&t[r0 + g][k0 + 2 * tq]);
  a[1] = am_ld32(&t[r0 + g + 8][k0 + 2 * tq]);
  a[2] = am_ld32(&t[r0 + g][k0 + 8 + 2 * tq]);
  a[3] = am_ld32(&t[r0 + g + 8][k0 + 8 + 2 * tq]);
}

// ------------------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(128)
lm_attn_fwd_mma_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, int S, int H,
                       int dh, float scale, bf16* __restrict__ o, bf16* __restrict__ P) {
  pdl_sync();
  __shared__ __align__(16) bf16 sQ[AM_S][AM_PF];
  __shared__ __align__(16) bf16 sK[AM_S][AM_PF];
  __shared__ __align__(16) bf16 sV[AM_S][AM_PF];
  const int n = blockIdx.x / H, h = blockIdx.x % H, E = H * dh;
  const size_t base = (size_t)n * S * E + (size_t)h * dh;
  am_load_tile(sQ, q, base, S, E, dh);
  am_load_tile(sK, k, base, S, E, dh);
  am_load_tile(sV, v, base, S, E, dh);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int r0 = warp * 16;
  if (r0 >= S) return;
  // ---- scores: 16 query rows x 64 keys
  float sc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < AM_DH / 16; ++ks) {
    uint32_t a[4];
    am_a_frag(a, sQ, r0, ks * 16, g, tq);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      mma16816(sc[nt], a, am_ld32(&sK[8 * nt + g][ks * 16 + 2 * tq]), am_ld32(&sK[8 * nt + g][ks * 16 + 8 + 2 * tq]));
  }
  // ---- softmax over the keys of rows (r0 + g) and (r0 + g + 8); a row lives in the 4 lanes of a quad
  uint32_t pk[8][2];          // packed bf16 probabilities: [n-tile][row half]
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * nt + 2 * tq + e;
        float x = sc[nt][2 * hf + e] * scale;
        x = j < S ? x : -INFINITY;
        sc[nt][2 * hf + e] = x;
        mx = fmaxf(mx, x);
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float x = sc[nt][2 * hf + e];
        const float ex = x == -INFINITY ? 0.f : __expf(x - mx);
        sc[nt][2 * hf + e] = ex;
        sum += ex;
      }
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.f / sum;
    const int i = r0 + g + 8 * hf;
    bf16* Prow = P + (((size_t)n * H + h) * S + i) * S;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint32_t u = am_pack2(sc[nt][2 * hf] * inv, sc[nt][2 * hf + 1] * inv);
      pk[nt][hf] = u;                       // the saved (rounded) probabilities are the ones used for P.V
      const int j = 8 * nt + 2 * tq;
      if (i < S) {
        if ((S & 1) == 0) {
          if (j < S) *reinterpret_cast<uint32_t*>(Prow + j) = u;
        } else {
          if (j < S) Prow[j] = __ushort_as_bfloat16((unsigned short)(u & 0xffffu));
          if (j + 1 < S) Prow[j + 1] = __ushort_as_bfloat16((unsigned short)(u >> 16));
        }
      }
    }
  }
  // ---- O = P V : reduction over the 64 keys (4 k-steps), features in 4 n-tiles
  float oc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { oc[nt][0] = oc[nt][1] = oc[nt][2] = oc[nt][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t a[4] = {pk[2 * kk][0], pk[2 * kk][1], pk[2 * kk + 1][0], pk[2 * kk + 1][1]};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const uint32_t b0 = am_pack_rows(&sV[16 * kk + 2 * tq][8 * nt + g], &sV[16 * kk + 2 * tq + 1][8 * nt + g]);
      const uint32_t b1 = am_pack_rows(&sV[16 * kk + 8 + 2 * tq][8 * nt + g], &sV[16 * kk + 9 + 2 * tq][8 * nt + g]);
      mma16816(oc[nt], a, b0, b1);
    }
  }
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int i = r0 + g + 8 * hf;
    if (i >= S) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int d = 8 * nt + 2 * tq;
      bf16* dst = o + base + (size_t)i * E + d;
      if (d + 1 < dh && ((dh & 1) == 0)) {
        *reinterpret_cast<uint32_t*>(dst) = am_pack2(oc[nt][2 * hf], oc[nt][2 * hf + 1]);
      } else {
        if (d < dh) dst[0] = __float2bfloat16(oc[nt][2 * hf]);
        if (d + 1 < dh) dst[1] = __float2bfloat16(oc[nt][2 * hf + 1]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------ backward
// column sums of a 16-row accumulator block over its rows -> shared bias accumulator (features 8*nt + 2*tq + {0,1})
__device__ __forceinline__ void am_bias_accum(float* sb, const float (&c)[4][4], int g, int tq) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    float s0 = c[nt][0] + c[nt][2], s1 = c[nt][1] + c[nt][3];
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (g == 0) {
      atomicAdd(&sb[8 * nt + 2 * tq], s0);
      atomicAdd(&sb[8 * nt + 2 * tq + 1], s1);
    }
  }
}
__device__ __forceinline__ void am_store_rows(bf16* __restrict__ dst, size_t base, int E, int dh, int S, int r0, int g, int tq,
                                              const float (&c)[4][4], float alpha) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int i = r0 + g + 8 * hf;
    if (i >= S) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int d = 8 * nt + 2 * tq;
      bf16* p = dst + base + (size_t)i * E + d;
      if (d + 1 < dh && ((dh & 1) == 0)) {
        *reinterpret_cast<uint32_t*>(p) = am_pack2(c[nt][2 * hf] * alpha, c[nt][2 * hf + 1] * alpha);
      } else {
        if (d < dh) p[0] = __float2bfloat16(c[nt][2 * hf] * alpha);
        if (d + 1 < dh) p[1] = __float2bfloat16(c[nt][2 * hf + 1] * alpha);
      }
    }
  }
}

__global__ void __launch_bounds__(128)
lm_attn_bwd_mma_kernel(const bf16* __restrict__ dO, const bf16* __restrict__ q, const bf16* __restrict__ k,
                       const bf16* __restrict__ v, const bf16* __restrict__ O, const bf16* __restrict__ P, int S, int H,
                       int dh, float scale, bf16* __restrict__ dq, bf16* __restrict__ dk, bf16* __restrict__ dv,
                       float* __restrict__ dbq, float* __restrict__ dbk, float* __restrict__ dbv, float alpha_v,
                       int scale_dv) {
  pdl_sync();
  __shared__ __align__(16) bf16 sQ[AM_S][AM_PF];
  __shared__ __align__(16) bf16 sK[AM_S][AM_PF];
  __shared__ __align__(16) bf16 sV[AM_S][AM_PF];
  __shared__ __align__(16) bf16 sD[AM_S][AM_PF];          // dO
  __shared__ __align__(16) bf16 sP[AM_S][AM_PS];          // P  [query][key]
  __shared__ __align__(16) bf16 sS[AM_S][AM_PS];          // dS [query][key]
  __shared__ float sDelta[AM_S];
  __shared__ float sRS[AM_S];                             // fp32 row sums of dS (phase A) -> k-bias gradient
  __shared__ float sBias[3][AM_DH];
  const int n = blockIdx.x / H, h = blockIdx.x % H, E = H * dh;
  const size_t base = (size_t)n * S * E + (size_t)h * dh;
  const int tid = threadIdx.x;
  am_load_tile(sQ, q, base, S, E, dh);
  am_load_tile(sK, k, base, S, E, dh);
  am_load_tile(sV, v, base, S, E, dh);
  am_load_tile(sD, dO, base, S, E, dh);
  {
    const bf16* Pb = P + ((size_t)n * H + h) * S * S;
    for (int it = tid; it < AM_S * AM_S; it += blockDim.x) {
      const int i = it >> 6, j = it & 63;
      sP[i][j] = (i < S && j < S) ? Pb[(size_t)i * S + j] : __float2bfloat16(0.f);
    }
    if (tid < 3 * AM_DH) (&sBias[0][0])[tid] = 0.f;
    // delta_i = dO_i . O_i : two threads per query row, 16 features each
    const int i = tid >> 1, part = tid & 1;
    float a = 0.f;
    if (i < S) {
      for (int d = part * 16; d < min(dh, part * 16 + 16); ++d) {
        const size_t gi = base + (size_t)i * E + d;
        a += __bfloat162float(dO[gi]) * __bfloat162float(O[gi]);
      }
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (part == 0) sDelta[i] = a;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int r0 = warp * 16;
  // ---- phase A: this warp's 16 query rows -> dP, dS (kept as A fragments + written to shared memory), dQ
  {
    float dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < AM_DH / 16; ++ks) {
      uint32_t a[4];
      am_a_frag(a, sD, r0, ks * 16, g, tq);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)      // B[k = feature][n = key] = V[key][feature]
        mma16816(dp[nt], a, am_ld32(&sV[8 * nt + g][ks * 16 + 2 * tq]), am_ld32(&sV[8 * nt + g][ks * 16 + 8 + 2 * tq]));
    }
    uint32_t ds[8][2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int i = r0 + g + 8 * hf;
      const float dl = sDelta[i];
      float rsum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int j = 8 * nt + 2 * tq;
        const uint32_t pu = am_ld32(&sP[i][j]);
        const float s0 = am_lo(pu) * (dp[nt][2 * hf] - dl) * scale;
        const float s1 = am_hi(pu) * (dp[nt][2 * hf + 1] - dl) * scale;
        rsum += s0 + s1;
        const uint32_t u = am_pack2(s0, s1);
        ds[nt][hf] = u;
        *reinterpret_cast<uint32_t*>(&sS[i][j]) = u;
      }
      rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
      rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
      if (tq == 0) sRS[i] = rsum;
    }
    __syncwarp();
    {
      // k-bias gradient = sum_j dK[j][d] = sum_i Q[i][d] * (sum_j dS[i][j]).  A softmax gradient row sums to zero, which
      // the fp32 row sums keep to round-off; summing the dK tile (built from bf16-rounded dS) would not.
      float acc = 0.f;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) acc += sRS[r0 + rr] * __bfloat162float(sQ[r0 + rr][lane]);
      atomicAdd(&sBias[1][lane], acc);
    }
    float dqc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { dqc[nt][0] = dqc[nt][1] = dqc[nt][2] = dqc[nt][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {      // dQ = dS K : reduction over keys, B[k = key][n = feature] = K[key][feature]
      const uint32_t a[4] = {ds[2 * kk][0], ds[2 * kk][1], ds[2 * kk + 1][0], ds[2 * kk + 1][1]};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint32_t b0 = am_pack_rows(&sK[16 * kk + 2 * tq][8 * nt + g], &sK[16 * kk + 2 * tq + 1][8 * nt + g]);
        const uint32_t b1 = am_pack_rows(&sK[16 * kk + 8 + 2 * tq][8 * nt + g], &sK[16 * kk + 9 + 2 * tq][8 * nt + g]);
        mma16816(dqc[nt], a, b0, b1);
      }
    }
    am_store_rows(dq, base, E, dh, S, r0, g, tq, dqc, 1.f);
    am_bias_accum(sBias[0], dqc, g, tq);
  }
  __syncthreads();
  // ---- phase B: this warp's 16 keys -> dK = dS^T Q, dV = P^T dO  (A[m = key][k = query] read transposed from smem)
  {
    float dkc[4][4], dvc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      dkc[nt][0] = dkc[nt][1] = dkc[nt][2] = dkc[nt][3] = 0.f;
      dvc[nt][0] = dvc[nt][1] = dvc[nt][2] = dvc[nt][3] = 0.f;
    }
    const int j0 = r0;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int i0 = 16 * kk + 2 * tq;
      const uint32_t as[4] = {am_pack_rows(&sS[i0][j0 + g], &sS[i0 + 1][j0 + g]),
                              am_pack_rows(&sS[i0][j0 + g + 8], &sS[i0 + 1][j0 + g + 8]),
                              am_pack_rows(&sS[i0 + 8][j0 + g], &sS[i0 + 9][j0 + g]),
                              am_pack_rows(&sS[i0 + 8][j0 + g + 8], &sS[i0 + 9][j0 + g + 8])};
      const uint32_t ap[4] = {am_pack_rows(&sP[i0][j0 + g], &sP[i0 + 1][j0 + g]),
                              am_pack_rows(&sP[i0][j0 + g + 8], &sP[i0 + 1][j0 + g + 8]),
                              am_pack_rows(&sP[i0 + 8][j0 + g], &sP[i0 + 9][j0 + g]),
                              am_pack_rows(&sP[i0 + 8][j0 + g + 8], &sP[i0 + 9][j0 + g + 8])};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint32_t q0 = am_pack_rows(&sQ[i0][8 * nt + g], &sQ[i0 + 1][8 * nt + g]);
        const uint32_t q1 = am_pack_rows(&sQ[i0 + 8][8 * nt + g], &sQ[i0 + 9][8 * nt + g]);
        mma16816(dkc[nt], as, q0, q1);
        const uint32_t d0 = am_pack_rows(&sD[i0][8 * nt + g], &sD[i0 + 1][8 * nt + g]);
        const uint32_t d1 = am_pack_rows(&sD[i0 + 8][8 * nt + g], &sD[i0 + 9][8 * nt + g]);
        mma16816(dvc[nt], ap, d0, d1);
      }
    }
    am_store_rows(dk, base, E, dh, S, j0, g, tq, dkc, 1.f);
    am_store_rows(dv, base, E, dh, S, j0, g, tq, dvc, scale_dv ? alpha_v : 1.f);
    am_bias_accum(sBias[2], dvc, g, tq);            // (the k-bias gradient comes from phase A's fp32 row sums)
  }
  __syncthreads();
  if (tid < dh) {
    atomicAdd(&dbq[h * dh + tid], sBias[0][tid]);
    atomicAdd(&dbk[h * dh + tid], sBias[1][tid]);
    atomicAdd(&dbv[h * dh + tid], sBias[2][tid] * alpha_v);
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" void hfl_lm_attn_fwd_mma(const void* q, const void* k, const void* v, int N, int S, int H, int dh, float scale,
                                    void* o, void* P, cudaStream_t st) {
  launch_k(lm_attn_fwd_mma_kernel, dim3(N * H), dim3(128), 0, st, (const bf16*)q, (const bf16*)k, (const bf16*)v, S, H, dh,
           scale, (bf16*)o, (bf16*)P);
  count_launch();
}

extern "C" int hfl_lm_attn_bwd_mma(const void* dO, const void* q, const void* k, const void* v, const void* O,
                                   const void* P, int N, int S, int H, int dh, float scale, void* dq, void* dk, void* dv,
                                   float* dbq, float* dbk, float* dbv, float alpha_v, int scale_dv, cudaStream_t st) {
  launch_k(lm_attn_bwd_mma_kernel, dim3(N * H), dim3(128), 0, st, (const bf16*)dO, (const bf16*)q, (const bf16*)k,
           (const bf16*)v, (const bf16*)O, (const bf16*)P, S, H, dh, scale, (bf16*)dq, (bf16*)dk, (bf16*)dv, dbq, dbk, dbv,
           alpha_v, scale_dv);
  count_launch();
  return (int)cudaGetLastError();
}
