// Fused masked-LM Transformer kernels (everything around the GEMMs of the encoder step).
//
// The federated LM step is tiny -- one client batch is rows x 64 tokens with rows = 1 for WikiText2/100 users, so a step
// is ~64 tokens of E <= 256 -- and is bound by the NUMBER of kernels, not by their work (447 launches per step on the
// autograd path).  These kernels collapse the elementwise / normalisation / attention / bias-gradient work between
// the tcgen05 GEMMs into one launch per site (reference call sites: src/models/transformer.py:14-37 embedding +
// LayerNorm + dropout, :45-51 + :66-85 attention, :92-100 residual + dropout + LayerNorm, :149-151 Bernoulli token
// masking, :31/:94/:125 GELU):
//
//   lm_embed_fwd / bwd      in-kernel Bernoulli [MASK] replacement (counter-based RNG) + token/position gather + LayerNorm
//                           (Scaler folded into eps) + dropout;  bwd = dropout mask + LN backward + sparse row scatter-add
//   lm_add_ln_fwd / bwd     y = LN(x + dropout(h)); bwd sums up to four incoming gradient branches on load and emits the
//                           residual gradient, the branch gradient, gamma/beta gradients and the producing linear's bias
//                           gradient (column sums) in the same pass
//   lm_attn_fwd / bwd       softmax(q k^T * scale) v for S <= 64 keys, one CTA per (row, head); bwd also reduces the
//                           q/k/v bias gradients
//   lm_gelu_drop_fwd / bwd  h = dropout(GELU(u)); bwd emits du and the bias gradient
//   lm_colsum               bias gradient of the vocabulary projection (column sums of dlogits)
//
// Random bits: every draw is hash(seed, step, site, element) -- `seed`/`step` are device scalars (the step counter is
// bumped inside the captured CUDA graph), so a replayed graph draws fresh masks every step.  Masks are written out
// (uint8) for the backward pass and for the numerics tests.
#include <stdlib.h>
#include "common.cuh"

namespace hfl {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t rng_key(const long long* seed_p, const long long* step_p, uint32_t site) {
  const unsigned long long s = (unsigned long long)seed_p[0] * 0x9E3779B97F4A7C15ULL + (unsigned long long)step_p[0];
  return mix32(uint32_t(s) ^ mix32(uint32_t(s >> 32) ^ (site * 0x9E3779B9U + 0x632BE5ABU)));
}
__device__ __forceinline__ float rng_uniform(uint32_t key, uint32_t idx) {
  const uint32_t h = mix32(key ^ mix32(idx + 0x85EBCA6BU));
  return (h >> 8) * (1.0f / 16777216.0f);
}

constexpr int LN_WARPS = 8;          // tokens per CTA
constexpr int LN_MAX_PL = 8;         // elements per lane -> E <= 256

__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float u) {
  return 0.5f * (1.f + erff(u * 0.70710678118654752f)) + u * 0.3989422804014327f * __expf(-0.5f * u * u);
}

// ---------------------------------------------------------------------------------------------------------------
// Embedding + LayerNorm + dropout
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LN_WARPS * 32)
lm_embed_fwd_kernel(const long long* __restrict__ tok, const bf16* __restrict__ emb, const bf16* __restrict__ pos,
                    const bf16* __restrict__ gamma, const bf16* __restrict__ beta, int M, int S, int E, int mask_id,
                    float eps, float p_mask, float p_drop, const long long* seed_p, const long long* step_p, uint32_t site,
                    bf16* __restrict__ out, int* __restrict__ src, bf16* __restrict__ xhat, float* __restrict__ rstd,
                    uint8_t* __restrict__ dmask) {
  pdl_sync();
  const int m = blockIdx.x * LN_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  const uint32_t k_mask = rng_key(seed_p, step_p, site), k_drop = rng_key(seed_p, step_p, site + 1);
  const int s = m % S;
  const bool replaced = p_mask > 0.f && rng_uniform(k_mask, (uint32_t)m) < p_mask;
  const int id = replaced ? mask_id : (int)tok[m];
  if (lane == 0) src[m] = id;
  float v[LN_MAX_PL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PL; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < E ? __bfloat162float(emb[(size_t)id * E + c]) + __bfloat162float(pos[(size_t)s * E + c]) : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PL; ++i) {
    const int c = lane + 32 * i;
    const float d = c < E ? v[i] - mean : 0.f;
    sq += d * d;
  }
  const float r = rsqrtf(warp_sum(sq) / E + eps);
  if (lane == 0) rstd[m] = r;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PL; ++i) {
    const int c = lane + 32 * i;
    if (c < E) {
      const size_t o = (size_t)m * E + c;
      const float xh = (v[i] - mean) * r;
      float y = xh * __bfloat162float(gamma[c]) + __bfloat162float(beta[c]);
      xhat[o] = __float2bfloat16(xh);
      if (p_drop > 0.f) {
        const bool keep = rng_uniform(k_drop, (uint32_t)o) >= p_drop;
        dmask[o] = keep ? 1 : 0;
        y = keep ? y * keep_scale : 0.f;
      }
      out[o] = __float2bfloat16(y);
    }
  }
}

// Shared LayerNorm-backward core: g = dy * gamma; dz = rstd * (g - mean(g) - xhat * mean(g * xhat)).
// Column partials (dgamma, dbeta [, dbias]) are reduced across the CTA's warps in shared memory, then one atomicAdd
// per column per CTA.
struct DyParts { const bf16* p[4]; int n; const float* f; };   // up to 4 bf16 branches + one optional fp32 branch

__global__ void __launch_bounds__(LN_WARPS * 32)
lm_embed_bwd_kernel(DyParts dy, const uint8_t* __restrict__ dmask, float p_drop, const bf16* __restrict__ xhat,
                    const float* __restrict__ rstd, const bf16* __restrict__ gamma, const int* __restrict__ src, int M,
                    int S, int E, float* __restrict__ demb, float* __restrict__ dpos, float* __restrict__ dgamma,
                    float* __restrict__ dbeta) {
  pdl_sync();
  __shared__ float sm_g[LN_WARPS][LN_MAX_PL * 32];
  __shared__ float sm_b[LN_WARPS][LN_MAX_PL * 32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * LN_WARPS + w;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float dyv[LN_MAX_PL], xh[LN_MAX_PL];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PL; ++i) {
    const int c = lane + 32 * i;
    dyv[i] = 0.f; xh[i] = 0.f;
    if (m < M && c < E) {
      const size_t o = (size_t)m * E + c;
      float d = 0.f;
      for (int k = 0; k < dy.n; ++k) d += __bfloat162float(dy.p[k][o]);
      if (dy.f != nullptr) d += dy.f[o];
      if (p_drop > 0.f) d = dmask[o] ? d * keep_scale : 0.f;
      dyv[i] = d;
      xh[i] = __bfloat162float(xhat[o]);
      const float g = d * __bfloat162float(gamma[c]);
      s1 += g; s2 += g * xh[i];
    }
    sm_g[w][c] = dyv[i] * xh[i];
    sm_b[w][c] = dyv[i];
  }
  s1 = warp_sum(s1) / E; s2 = warp_sum(s2) / E;
  if (m < M) {
    const float r = rstd[m];
    const int id = src[m], s = m % S;
#pragma unroll
    for (int i = 0; i < LN_MAX_PL; ++i) {
      const int c = lane + 32 * i;
      if (c < E) {
        const float dx = r * (dyv[i] * __bfloat162float(gamma[c]) - s1 - xh[i] * s2);
        atomicAdd(&demb[(size_t)id * E + c], dx);
        atomicAdd(&dpos[(size_t)s * E + c], dx);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < LN_WARPS; ++k) { a += sm_g[k][c]; b += sm_b[k][c]; }
    atomicAdd(&dgamma[c], a);
    atomicAdd(&dbeta[c], b);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// y = LN(x + dropout(h))      (x optional, dropout optional)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LN_WARPS * 32)
lm_add_ln_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ h, const bf16* __restrict__ gamma,
                     const bf16* __restrict__ beta, int M, int E, float eps, float p_drop, const long long* seed_p,
                     const long long* step_p, uint32_t site, bf16* __restrict__ y, bf16* __restrict__ xhat,
                     float* __restrict__ rstd, uint8_t* __restrict__ dmask) {
  pdl_sync();
  const int m = blockIdx.x * LN_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  const uint32_t k_drop = p_drop > 0.f ? rng_key(seed_p, step_p, site) : 0u;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float v[LN_MAX_PL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PL; ++i) {
    const int c = lane + 32 * i;
    v[i] = 0.f;
    if (c < E) {
      const size_t o = (size_t)m * E + c;
      float hv = __bfloat162float(h[o]);
      if (p_drop > 0.f) {
        const bool keep = rng_uniform(k_drop, (uint32_t)o) >= p_drop;
        dmask[o] = keep ? 1 : 0;
        hv = keep ? hv * keep_scale : 0.f;
      }
      // the residual sum is rounded to bf16 like the unfused graph (x + dropout(h) materialised in bf16)
      v[i] = x != nullptr ? __bfloat162float(__float2bfloat16(__bfloat162float(x[o]) + hv)) : hv;
    }
    sum += v[i];
  }
  const float mean = warp_sum(sum) / E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PL; ++i) {
    const int c = lane + 32 * i;
    const float d = c < E ? v[i] - mean : 0.f;
    sq += d * d;
  }
  const float r = rsqrtf(warp_sum(sq) / E + eps);
  if (lane == 0) rstd[m] = r;
#pragma unroll
  for (int i = 0; i < LN_MAX_PL; ++i) {
    const int c = lane + 32 * i;
    if (c < E) {
      const size_t o = (size_t)m * E + c;
      const float xh = (v[i] - mean) * r;
      xhat[o] = __float2bfloat16(xh);
      y[o] = __float2bfloat16(xh * __bfloat162float(gamma[c]) + __bfloat162float(beta[c]));
    }
  }
}

__global__ void __launch_bounds__(LN_WARPS * 32)
lm_add_ln_bwd_kernel(DyParts dy, const bf16* __restrict__ xhat, const float* __restrict__ rstd,
                     const bf16* __restrict__ gamma, const uint8_t* __restrict__ dmask, float p_drop, int M, int E,
                     bf16* __restrict__ dz, bf16* __restrict__ dh, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dbias, float bias_alpha) {
  pdl_sync();
  __shared__ float sm_g[LN_WARPS][LN_MAX_PL * 32];
  __shared__ float sm_b[LN_WARPS][LN_MAX_PL * 32];
  __shared__ float sm_h[LN_WARPS][LN_MAX_PL * 32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * LN_WARPS + w;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float dyv[LN_MAX_PL], xh[LN_MAX_PL];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PL; ++i) {
    const int c = lane + 32 * i;
    dyv[i] = 0.f; xh[i] = 0.f;
    if (m < M && c < E) {
      const size_t o = (size_t)m * E + c;
      float d = 0.f;
      for (int k = 0; k < dy.n; ++k) d += __bfloat162float(dy.p[k][o]);
      if (dy.f != nullptr) d += dy.f[o];
      dyv[i] = d;
      xh[i] = __bfloat162float(xhat[o]);
      const float g = d * __bfloat162float(gamma[c]);
      s1 += g; s2 += g * xh[i];
    }
    sm_g[w][c] = dyv[i] * xh[i];
    sm_b[w][c] = dyv[i];
  }
  s1 = warp_sum(s1) / E; s2 = warp_sum(s2) / E;
  const float r = m < M ? rstd[m] : 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PL; ++i) {
    const int c = lane + 32 * i;
    float hv = 0.f;
    if (m < M && c < E) {
      const size_t o = (size_t)m * E + c;
      const float d = r * (dyv[i] * __bfloat162float(gamma[c]) - s1 - xh[i] * s2);
      dz[o] = __float2bfloat16(d);
      hv = d;
      if (p_drop > 0.f) hv = dmask[o] ? d * keep_scale : 0.f;
      if (dh != nullptr) dh[o] = __float2bfloat16(hv);
    }
    sm_h[w][c] = hv;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float a = 0.f, b = 0.f, hsum = 0.f;
#pragma unroll
    for (int k = 0; k < LN_WARPS; ++k) { a += sm_g[k][c]; b += sm_b[k][c]; hsum += sm_h[k][c]; }
    atomicAdd(&dgamma[c], a);
    atomicAdd(&dbeta[c], b);
    if (dbias != nullptr) atomicAdd(&dbias[c], hsum * bias_alpha);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// h = dropout(GELU(u))
// ---------------------------------------------------------------------------------------------------------------
__global__ void lm_gelu_drop_fwd_kernel(const bf16* __restrict__ u, long n, float p_drop, const long long* seed_p,
                                        const long long* step_p, uint32_t site, bf16* __restrict__ h,
                                        uint8_t* __restrict__ dmask) {
  pdl_sync();
  const uint32_t k_drop = p_drop > 0.f ? rng_key(seed_p, step_p, site) : 0u;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float g = __bfloat162float(__float2bfloat16(gelu_f(__bfloat162float(u[i]))));
    if (p_drop > 0.f) {
      const bool keep = rng_uniform(k_drop, (uint32_t)i) >= p_drop;
      dmask[i] = keep ? 1 : 0;
      g = keep ? g * keep_scale : 0.f;
    }
    h[i] = __float2bfloat16(g);
  }
}

constexpr int GB_ROWS = 32;   // rows per CTA in the column-reducing elementwise kernels
// du = dh * mask/(1-p) * GELU'(u);  dbias[c] += alpha * sum_rows du     grid (ceil(F/128), ceil(M/GB_ROWS)), 128 threads
__global__ void __launch_bounds__(128)
lm_gelu_drop_bwd_kernel(const bf16* __restrict__ dh, const bf16* __restrict__ u, const uint8_t* __restrict__ dmask,
                        float p_drop, int M, int F, bf16* __restrict__ du, float* __restrict__ dbias, float bias_alpha) {
  pdl_sync();
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= F) return;
  const int r0 = blockIdx.y * GB_ROWS, r1 = min(M, r0 + GB_ROWS);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) {
    const size_t o = (size_t)r * F + c;
    float d = __bfloat162float(dh[o]);
    if (p_drop > 0.f) d = dmask[o] ? d * keep_scale : 0.f;
    d *= gelu_grad(__bfloat162float(u[o]));
    du[o] = __float2bfloat16(d);
    acc += d;
  }
  if (dbias != nullptr) atomicAdd(&dbias[c], acc * bias_alpha);
}

// out[c] += alpha * sum_r x[r, c]
__global__ void __launch_bounds__(128)
lm_colsum_kernel(const bf16* __restrict__ x, int M, int N, int ld, float alpha, float* __restrict__ out) {
  pdl_sync();
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= N) return;
  const int r0 = blockIdx.y * GB_ROWS, r1 = min(M, r0 + GB_ROWS);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += __bfloat162float(x[(size_t)r * ld + c]);
  atomicAdd(&out[c], acc * alpha);
}

// ---------------------------------------------------------------------------------------------------------------
// Attention for S <= 64 keys, head dim <= 32.  q, k, v, o: [N*S, E] bf16 with head h at columns [h*dh, (h+1)*dh);
// P: [N, H, S, S] bf16 (saved for backward).  A federated client batch is ONE row of 64 tokens, i.e. only H = 8
// (row, head) pairs: the kernels therefore split every pair over several CTAs so the step does not crawl on 8 SMs.
//   forward : grid (N*H, 4)  -- CTA = 16 query rows x 16 lanes (4 keys and 2 features per lane)
//   backward: grid (N*H, 8)  -- y < 4: 16 query rows -> dQ;  y >= 4: 16 keys -> dK, dV.  Both roles recompute their
//             dP block and use delta_i = dO_i . O_i (so no role needs a full row/column of the other).
// ---------------------------------------------------------------------------------------------------------------
constexpr int AT_S = 64, AT_DH = 32, AT_BLK = 16;

__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256)
lm_attn_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, int S, int H,
                   int dh, float scale, bf16* __restrict__ o, bf16* __restrict__ P) {
  pdl_sync();
  __shared__ float sK[AT_S][AT_DH + 1];
  __shared__ float sV[AT_S][AT_DH + 1];
  __shared__ float sQ[AT_BLK][AT_DH + 1];
  __shared__ float sP[AT_BLK][AT_S + 1];
  const int n = blockIdx.x / H, h = blockIdx.x % H, E = H * dh;
  const int i0 = blockIdx.y * AT_BLK;
  if (i0 >= S) return;
  const int tid = threadIdx.x;
  const size_t base = (size_t)n * S * E + (size_t)h * dh;
  for (int t = tid; t < AT_S * AT_DH; t += 256) {
    const int j = t / AT_DH, d = t % AT_DH;
    const bool ok = j < S && d < dh;
    sK[j][d] = ok ? __bfloat162float(k[base + (size_t)j * E + d]) : 0.f;
    sV[j][d] = ok ? __bfloat162float(v[base + (size_t)j * E + d]) : 0.f;
  }
  for (int t = tid; t < AT_BLK * AT_DH; t += 256) {
    const int r = t / AT_DH, d = t % AT_DH;
    sQ[r][d] = (i0 + r < S && d < dh) ? __bfloat162float(q[base + (size_t)(i0 + r) * E + d]) : 0.f;
  }
  __syncthreads();
  const int r = tid >> 4, l = tid & 15, i = i0 + r;
  float sc[4];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int j = l + 16 * t;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < AT_DH; ++d) a += sQ[r][d] * sK[j][d];
    sc[t] = j < S ? a * scale : -INFINITY;
    mx = fmaxf(mx, sc[t]);
  }
  mx = half_warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    sc[t] = sc[t] == -INFINITY ? 0.f : __expf(sc[t] - mx);
    sum += sc[t];
  }
  const float inv = 1.f / half_warp_sum(sum);
  bf16* Prow = P + (((size_t)n * H + h) * S + i) * S;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int j = l + 16 * t;
    const bf16 pb = __float2bfloat16(sc[t] * inv);
    sP[r][j] = __bfloat162float(pb);            // the saved (rounded) probabilities are the ones used for P.V
    if (i < S && j < S) Prow[j] = pb;
  }
  __syncwarp();                                  // a row's 16 lanes live in one warp
  if (i < S) {
    for (int d = l; d < dh; d += 16) {
      float a = 0.f;
#pragma unroll 8
      for (int j = 0; j < AT_S; ++j) a += sP[r][j] * sV[j][d];
      o[base + (size_t)i * E + d] = __float2bfloat16(a);
    }
  }
}

__global__ void __launch_bounds__(256)
lm_attn_bwd_kernel(const bf16* __restrict__ dO, const bf16* __restrict__ q, const bf16* __restrict__ k,
                   const bf16* __restrict__ v, const bf16* __restrict__ O, const bf16* __restrict__ P, int S, int H,
                   int dh, float scale, bf16* __restrict__ dq, bf16* __restrict__ dk, bf16* __restrict__ dv,
                   float* __restrict__ dbq, float* __restrict__ dbk, float* __restrict__ dbv, float alpha_v,
                   int scale_dv /* 1: dv is written pre-multiplied by alpha_v (grouped q/k/v backward GEMMs) */) {
  pdl_sync();
  __shared__ float sA[AT_S][AT_DH + 1];          // role Q: K      role K: Q
  __shared__ float sB[AT_S][AT_DH + 1];          // role Q: V      role K: dO
  __shared__ float sC[AT_BLK][AT_DH + 1];        // role Q: dO_i   role K: V_j
  __shared__ float sS[AT_BLK][AT_S + 1];         // dS block (role K: transposed, [key][query])
  __shared__ float sPb[AT_BLK][AT_S + 1];        // role K: P block transposed
  __shared__ float sDelta[AT_S];
  __shared__ float sBias[2][AT_DH];
  const int n = blockIdx.x / H, h = blockIdx.x % H, E = H * dh;
  const bool role_q = blockIdx.y < 4;
  const int b0 = (blockIdx.y & 3) * AT_BLK;      // first query (role Q) / key (role K) of this CTA
  if (b0 >= S) return;
  const int tid = threadIdx.x;
  const size_t base = (size_t)n * S * E + (size_t)h * dh;
  const bf16* Pb = P + ((size_t)n * H + h) * S * S;
  const bf16* gA = role_q ? k : q;
  const bf16* gB = role_q ? v : dO;
  const bf16* gC = role_q ? dO : v;
  for (int t = tid; t < AT_S * AT_DH; t += 256) {
    const int j = t / AT_DH, d = t % AT_DH;
    const bool ok = j < S && d < dh;
    const size_t g = base + (size_t)j * E + d;
    sA[j][d] = ok ? __bfloat162float(gA[g]) : 0.f;
    sB[j][d] = ok ? __bfloat162float(gB[g]) : 0.f;
  }
  for (int t = tid; t < AT_BLK * AT_DH; t += 256) {
    const int r = t / AT_DH, d = t % AT_DH;
    sC[r][d] = (b0 + r < S && d < dh) ? __bfloat162float(gC[base + (size_t)(b0 + r) * E + d]) : 0.f;
  }
  if (tid < 2 * AT_DH) (&sBias[0][0])[tid] = 0.f;
  {  // delta_i = dO_i . O_i for every query: 4 threads per query, 8 features each
    const int i = tid >> 2, part = tid & 3;
    float a = 0.f;
    if (i < S) {
      for (int d = part * 8; d < min(dh, part * 8 + 8); ++d) {
        const size_t g = base + (size_t)i * E + d;
        a += __bfloat162float(dO[g]) * __bfloat162float(O[g]);
      }
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    if (part == 0) sDelta[i] = a;
  }
  __syncthreads();
  const int r = tid >> 4, l = tid & 15;
  if (role_q) {
    const int i = b0 + r;
    // dS_ij = P_ij (dO_i . V_j - delta_i) * scale for the lane's 4 keys
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = l + 16 * t;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < AT_DH; ++d) a += sC[r][d] * sB[j][d];
      const float p = (i < S && j < S) ? __bfloat162float(Pb[(size_t)i * S + j]) : 0.f;
      sS[r][j] = p * (a - sDelta[min(i, AT_S - 1)]) * scale;
    }
    __syncwarp();
    if (i < S) {
      for (int d = l; d < dh; d += 16) {
        float a = 0.f;
#pragma unroll 8
        for (int j = 0; j < AT_S; ++j) a += sS[r][j] * sA[j][d];
        dq[base + (size_t)i * E + d] = __float2bfloat16(a);
        atomicAdd(&sBias[0][d], a);
      }
    }
    __syncthreads();
    if (tid < dh) atomicAdd(&dbq[h * dh + tid], sBias[0][tid]);
  } else {
    const int j = b0 + r;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = l + 16 * t;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < AT_DH; ++d) a += sB[i][d] * sC[r][d];           // dO_i . V_j
      const float p = (i < S && j < S) ? __bfloat162float(Pb[(size_t)i * S + j]) : 0.f;
      sPb[r][i] = p;
      sS[r][i] = p * (a - sDelta[i]) * scale;
    }
    __syncwarp();
    if (j < S) {
      for (int d = l; d < dh; d += 16) {
        float ak = 0.f, av = 0.f;
#pragma unroll 8
        for (int i = 0; i < AT_S; ++i) {
          ak += sS[r][i] * sA[i][d];             // dS_ij Q_i
          av += sPb[r][i] * sB[i][d];            // P_ij dO_i
        }
        const size_t g = base + (size_t)j * E + d;
        dk[g] = __float2bfloat16(ak);
        dv[g] = __float2bfloat16(scale_dv ? av * alpha_v : av);
        atomicAdd(&sBias[0][d], ak);
        atomicAdd(&sBias[1][d], av);
      }
    }
    __syncthreads();
    if (tid < dh) {
      atomicAdd(&dbk[h * dh + tid], sBias[0][tid]);
      atomicAdd(&dbv[h * dh + tid], sBias[1][tid] * alpha_v);
    }
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" int hfl_lm_max_e() { return LN_MAX_PL * 32; }
extern "C" int hfl_lm_max_s() { return AT_S; }
extern "C" int hfl_lm_max_dh() { return AT_DH; }

extern "C" void hfl_lm_embed_fwd(const long long* tok, const void* emb, const void* pos, const void* gamma,
                                 const void* beta, int M, int S, int E, int mask_id, float eps, float p_mask,
                                 float p_drop, const long long* seed_p, const long long* step_p, unsigned site, void* out,
                                 int* src, void* xhat, float* rstd, unsigned char* dmask, cudaStream_t st) {
  launch_k(lm_embed_fwd_kernel, dim3((M + LN_WARPS - 1) / LN_WARPS), dim3(LN_WARPS * 32), 0, st, 
      tok, (const bf16*)emb, (const bf16*)pos, (const bf16*)gamma, (const bf16*)beta, M, S, E, mask_id, eps, p_mask,
      p_drop, seed_p, step_p, site, (bf16*)out, src, (bf16*)xhat, rstd, dmask);
  count_launch();
}

static DyParts make_parts(const void* const* dy, int n, const float* f = nullptr) {
  DyParts d;
  d.n = n;
  d.f = f;
  for (int i = 0; i < 4; ++i) d.p[i] = (const bf16*)(i < n ? dy[i] : nullptr);
  return d;
}

extern "C" void hfl_lm_embed_bwd(const void* const* dy, int n_dy, const unsigned char* dmask, float p_drop,
                                 const void* xhat, const float* rstd, const void* gamma, const int* src, int M, int S,
                                 int E, float* demb, float* dpos, float* dgamma, float* dbeta, cudaStream_t st) {
  launch_k(lm_embed_bwd_kernel, dim3((M + LN_WARPS - 1) / LN_WARPS), dim3(LN_WARPS * 32), 0, st, 
      make_parts(dy, n_dy), dmask, p_drop, (const bf16*)xhat, rstd, (const bf16*)gamma, src, M, S, E, demb, dpos, dgamma,
      dbeta);
  count_launch();
}

extern "C" void hfl_lm_add_ln_fwd(const void* x, const void* h, const void* gamma, const void* beta, int M, int E,
                                  float eps, float p_drop, const long long* seed_p, const long long* step_p,
                                  unsigned site, void* y, void* xhat, float* rstd, unsigned char* dmask,
                                  cudaStream_t st) {
  launch_k(lm_add_ln_fwd_kernel, dim3((M + LN_WARPS - 1) / LN_WARPS), dim3(LN_WARPS * 32), 0, st, 
      (const bf16*)x, (const bf16*)h, (const bf16*)gamma, (const bf16*)beta, M, E, eps, p_drop, seed_p, step_p, site,
      (bf16*)y, (bf16*)xhat, rstd, dmask);
  count_launch();
}

extern "C" void hfl_lm_add_ln_bwd(const void* const* dy, int n_dy, const void* xhat, const float* rstd,
                                  const void* gamma, const unsigned char* dmask, float p_drop, int M, int E, void* dz,
                                  void* dh, float* dgamma, float* dbeta, float* dbias, float bias_alpha,
                                  const float* dy_f32, cudaStream_t st) {
  launch_k(lm_add_ln_bwd_kernel, dim3((M + LN_WARPS - 1) / LN_WARPS), dim3(LN_WARPS * 32), 0, st, 
      make_parts(dy, n_dy, dy_f32), (const bf16*)xhat, rstd, (const bf16*)gamma, dmask, p_drop, M, E, (bf16*)dz, (bf16*)dh,
      dgamma, dbeta, dbias, bias_alpha);
  count_launch();
}

extern "C" void hfl_lm_gelu_drop_fwd(const void* u, long n, float p_drop, const long long* seed_p,
                                     const long long* step_p, unsigned site, void* h, unsigned char* dmask,
                                     cudaStream_t st) {
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  launch_k(lm_gelu_drop_fwd_kernel, dim3(blocks), dim3(256), 0, st, (const bf16*)u, n, p_drop, seed_p, step_p, site, (bf16*)h, dmask);
  count_launch();
}

extern "C" void hfl_lm_gelu_drop_bwd(const void* dh, const void* u, const unsigned char* dmask, float p_drop, int M,
                                     int F, void* du, float* dbias, float bias_alpha, cudaStream_t st) {
  dim3 grid((F + 127) / 128, (M + GB_ROWS - 1) / GB_ROWS);
  launch_k(lm_gelu_drop_bwd_kernel, dim3(grid), dim3(128), 0, st, (const bf16*)dh, (const bf16*)u, dmask, p_drop, M, F, (bf16*)du, dbias,
                                                bias_alpha);
  count_launch();
}

extern "C" void hfl_lm_colsum(const void* x, int M, int N, int ld, float alpha, float* out, cudaStream_t st) {
  dim3 grid((N + 127) / 128, (M + GB_ROWS - 1) / GB_ROWS);
  launch_k(lm_colsum_kernel, dim3(grid), dim3(128), 0, st, (const bf16*)x, M, N, ld, alpha, out);
  count_launch();
}

// tensor-core (mma.sync) attention kernels, csrc/lm_attn_mma.cu; HFL_LM_ATTN_MMA=0 selects the CUDA-core kernels above
extern "C" void hfl_lm_attn_fwd_mma(const void*, const void*, const void*, int, int, int, int, float, void*, void*,
                                    cudaStream_t);
extern "C" int hfl_lm_attn_bwd_mma(const void*, const void*, const void*, const void*, const void*, const void*, int, int,
                                   int, int, float, void*, void*, void*, float*, float*, float*, float, int, cudaStream_t);
// Dispatch (measured, profiles/pdl_conv_ab.md + profiles/prof_attn_mma.md): the federated step has N*H = 8 (row, head) pairs --
// there the CUDA-core kernels, which split a pair over 4-8 CTAs, finish in 7.3 / 9.3 us against 10.3 / 17.5 us for the
// one-CTA-per-pair mma kernels; from a few dozen pairs up (evaluation: 6 400, centralised training) the machine is full
// either way and the contraction belongs on the tensor cores.  HFL_LM_ATTN_MMA=1 / 0 forces one path.
static bool attn_mma_enabled(int pairs) {
  static const int mode = []() { const char* e = getenv("HFL_LM_ATTN_MMA"); return e == nullptr ? -1 : (e[0] != '0' ? 1 : 0); }();
  return mode < 0 ? pairs >= 32 : mode == 1;
}

extern "C" void hfl_lm_attn_fwd(const void* q, const void* k, const void* v, int N, int S, int H, int dh, float scale,
                                void* o, void* P, cudaStream_t st) {
  if (attn_mma_enabled(N * H)) { hfl_lm_attn_fwd_mma(q, k, v, N, S, H, dh, scale, o, P, st); return; }
  launch_k(lm_attn_fwd_kernel, dim3(dim3(N * H, 4)), dim3(256), 0, st, (const bf16*)q, (const bf16*)k, (const bf16*)v, S, H, dh, scale,
                                                     (bf16*)o, (bf16*)P);
  count_launch();
}

extern "C" int hfl_lm_attn_bwd(const void* dO, const void* q, const void* k, const void* v, const void* O, const void* P,
                               int N, int S, int H, int dh, float scale, void* dq, void* dk, void* dv, float* dbq,
                               float* dbk, float* dbv, float alpha_v, int scale_dv, cudaStream_t st) {
  if (attn_mma_enabled(N * H))
    return hfl_lm_attn_bwd_mma(dO, q, k, v, O, P, N, S, H, dh, scale, dq, dk, dv, dbq, dbk, dbv, alpha_v, scale_dv, st);
  launch_k(lm_attn_bwd_kernel, dim3(dim3(N * H, 8)), dim3(256), 0, st, (const bf16*)dO, (const bf16*)q, (const bf16*)k, (const bf16*)v,
                                                     (const bf16*)O, (const bf16*)P, S, H, dh, scale, (bf16*)dq,
                                                     (bf16*)dk, (bf16*)dv, dbq, dbk, dbv, alpha_v, scale_dv);
  count_launch();
  return (int)cudaGetLastError();
}
