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
struct DyParts { const bf16* p[4]; int n; };

__global__ void __launch_bounds__(LN_WARPS * 32)
lm_embed_bwd_kernel(DyParts dy, const uint8_t* __restrict__ dmask, float p_drop, const bf16* __restrict__ xhat,
                    const float* __restrict__ rstd, const bf16* __restrict__ gamma, const int* __restrict__ src, int M,
                    int S, int E, float* __restrict__ demb, float* __restrict__ dpos, float* __restrict__ dgamma,
                    float* __restrict__ dbeta) {
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
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= N) return;
  const int r0 = blockIdx.y * GB_ROWS, r1 = min(M, r0 + GB_ROWS);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += __bfloat162float(x[(size_t)r * ld + c]);
  atomicAdd(&out[c], acc * alpha);
}

// ---------------------------------------------------------------------------------------------------------------
// Attention for S <= 64 keys: one CTA per (row n, head h), 256 threads = 64 query rows x 4 key/feature parts.
// q, k, v, o: [N*S, E] bf16 with head h at columns [h*dh, (h+1)*dh).  P: [N, H, S, S] bf16 (saved for backward).
// ---------------------------------------------------------------------------------------------------------------
constexpr int AT_S = 64, AT_DH = 32;

__global__ void __launch_bounds__(256)
lm_attn_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, int S, int H,
                   int dh, float scale, bf16* __restrict__ o, bf16* __restrict__ P) {
  __shared__ float sK[AT_S][AT_DH + 1];
  __shared__ float sV[AT_S][AT_DH + 1];
  __shared__ float sP[AT_S][AT_S + 1];
  const int n = blockIdx.x / H, h = blockIdx.x % H, E = H * dh;
  const int tid = threadIdx.x;
  const size_t base = (size_t)n * S * E + (size_t)h * dh;
  for (int t = tid; t < AT_S * AT_DH; t += 256) {
    const int j = t / AT_DH, d = t % AT_DH;
    const bool ok = j < S && d < dh;
    sK[j][d] = ok ? __bfloat162float(k[base + (size_t)j * E + d]) : 0.f;
    sV[j][d] = ok ? __bfloat162float(v[base + (size_t)j * E + d]) : 0.f;
  }
  __syncthreads();
  const int i = tid >> 2, part = tid & 3;
  float qv[AT_DH];
#pragma unroll
  for (int d = 0; d < AT_DH; ++d) qv[d] = (i < S && d < dh) ? __bfloat162float(q[base + (size_t)i * E + d]) : 0.f;
  float sc[AT_S / 4];
  float mx = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < AT_S / 4; ++jj) {
    const int j = part * (AT_S / 4) + jj;
    float a = 0.f;
    if (j < S) {
#pragma unroll
      for (int d = 0; d < AT_DH; ++d) a += qv[d] * sK[j][d];      // sK columns >= dh are never read as non-zero q
      a *= scale;
      mx = fmaxf(mx, a);
    }
    sc[jj] = j < S ? a : -INFINITY;
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float sum = 0.f;
#pragma unroll
  for (int jj = 0; jj < AT_S / 4; ++jj) {
    sc[jj] = sc[jj] == -INFINITY ? 0.f : __expf(sc[jj] - mx);
    sum += sc[jj];
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  const float inv = 1.f / sum;
  bf16* Prow = P + (((size_t)n * H + h) * S + i) * S;
#pragma unroll
  for (int jj = 0; jj < AT_S / 4; ++jj) {
    const int j = part * (AT_S / 4) + jj;
    if (i < S && j < S) {
      const bf16 pb = __float2bfloat16(sc[jj] * inv);
      Prow[j] = pb;
      sP[i][j] = __bfloat162float(pb);        // the saved (rounded) probabilities are the ones used for P.V
    }
  }
  __syncthreads();
  if (i < S) {
    for (int d = part; d < dh; d += 4) {
      float a = 0.f;
      for (int j = 0; j < S; ++j) a += sP[i][j] * sV[j][d];
      o[base + (size_t)i * E + d] = __float2bfloat16(a);
    }
  }
}

// dynamic smem: sQ, sK, sV, sO [S][DH+1] floats, sP, sS [S][S+1] floats
__global__ void __launch_bounds__(256)
lm_attn_bwd_kernel(const bf16* __restrict__ dO, const bf16* __restrict__ q, const bf16* __restrict__ k,
                   const bf16* __restrict__ v, const bf16* __restrict__ P, int S, int H, int dh, float scale,
                   bf16* __restrict__ dq, bf16* __restrict__ dk, bf16* __restrict__ dv, float* __restrict__ dbq,
                   float* __restrict__ dbk, float* __restrict__ dbv, float alpha_v) {
  extern __shared__ float smf[];
  float (*sQ)[AT_DH + 1] = reinterpret_cast<float (*)[AT_DH + 1]>(smf);
  float (*sK)[AT_DH + 1] = sQ + AT_S;
  float (*sV)[AT_DH + 1] = sK + AT_S;
  float (*sO)[AT_DH + 1] = sV + AT_S;
  float (*sP)[AT_S + 1] = reinterpret_cast<float (*)[AT_S + 1]>(&sO[AT_S][0]);
  float (*sS)[AT_S + 1] = sP + AT_S;
  float* sB = &sS[AT_S][0];                              // [3][AT_DH] bias partials
  const int n = blockIdx.x / H, h = blockIdx.x % H, E = H * dh;
  const int tid = threadIdx.x;
  const size_t base = (size_t)n * S * E + (size_t)h * dh;
  for (int t = tid; t < AT_S * AT_DH; t += 256) {
    const int j = t / AT_DH, d = t % AT_DH;
    const bool ok = j < S && d < dh;
    const size_t g = base + (size_t)j * E + d;
    sQ[j][d] = ok ? __bfloat162float(q[g]) : 0.f;
    sK[j][d] = ok ? __bfloat162float(k[g]) : 0.f;
    sV[j][d] = ok ? __bfloat162float(v[g]) : 0.f;
    sO[j][d] = ok ? __bfloat162float(dO[g]) : 0.f;
  }
  const bf16* Pb = P + ((size_t)n * H + h) * S * S;
  for (int t = tid; t < AT_S * AT_S; t += 256) {
    const int i = t / AT_S, j = t % AT_S;
    sP[i][j] = (i < S && j < S) ? __bfloat162float(Pb[(size_t)i * S + j]) : 0.f;
  }
  if (tid < 3 * AT_DH) sB[tid] = 0.f;
  __syncthreads();
  const int i = tid >> 2, part = tid & 3;
  {
    float dp[AT_S / 4];
    float delta = 0.f;
#pragma unroll
    for (int jj = 0; jj < AT_S / 4; ++jj) {
      const int j = part * (AT_S / 4) + jj;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < AT_DH; ++d) a += sO[i][d] * sV[j][d];
      dp[jj] = a;
      delta += a * sP[i][j];
    }
    delta += __shfl_xor_sync(0xffffffffu, delta, 1);
    delta += __shfl_xor_sync(0xffffffffu, delta, 2);
#pragma unroll
    for (int jj = 0; jj < AT_S / 4; ++jj) {
      const int j = part * (AT_S / 4) + jj;
      sS[i][j] = sP[i][j] * (dp[jj] - delta) * scale;
    }
  }
  __syncthreads();
  // row index r = i plays "query i" for dQ and "key j" for dK / dV
  if (i < S) {
    for (int d = part; d < dh; d += 4) {
      float aq = 0.f, ak = 0.f, av = 0.f;
      for (int j = 0; j < S; ++j) {
        aq += sS[i][j] * sK[j][d];
        ak += sS[j][i] * sQ[j][d];
        av += sP[j][i] * sO[j][d];
      }
      const size_t g = base + (size_t)i * E + d;
      dq[g] = __float2bfloat16(aq);
      dk[g] = __float2bfloat16(ak);
      dv[g] = __float2bfloat16(av);
      atomicAdd(&sB[d], aq);
      atomicAdd(&sB[AT_DH + d], ak);
      atomicAdd(&sB[2 * AT_DH + d], av);
    }
  }
  __syncthreads();
  if (tid < dh) {
    atomicAdd(&dbq[h * dh + tid], sB[tid]);
    atomicAdd(&dbk[h * dh + tid], sB[AT_DH + tid]);
    atomicAdd(&dbv[h * dh + tid], sB[2 * AT_DH + tid] * alpha_v);
  }
}

constexpr int ATTN_BWD_SMEM = (4 * AT_S * (AT_DH + 1) + 2 * AT_S * (AT_S + 1) + 3 * AT_DH) * 4;

}  // namespace hfl

using namespace hfl;

extern "C" int hfl_lm_max_e() { return LN_MAX_PL * 32; }
extern "C" int hfl_lm_max_s() { return AT_S; }
extern "C" int hfl_lm_max_dh() { return AT_DH; }

extern "C" void hfl_lm_embed_fwd(const long long* tok, const void* emb, const void* pos, const void* gamma,
                                 const void* beta, int M, int S, int E, int mask_id, float eps, float p_mask,
                                 float p_drop, const long long* seed_p, const long long* step_p, unsigned site, void* out,
                                 int* src, void* xhat, float* rstd, unsigned char* dmask, cudaStream_t st) {
  lm_embed_fwd_kernel<<<(M + LN_WARPS - 1) / LN_WARPS, LN_WARPS * 32, 0, st>>>(
      tok, (const bf16*)emb, (const bf16*)pos, (const bf16*)gamma, (const bf16*)beta, M, S, E, mask_id, eps, p_mask,
      p_drop, seed_p, step_p, site, (bf16*)out, src, (bf16*)xhat, rstd, dmask);
  count_launch();
}

static DyParts make_parts(const void* const* dy, int n) {
  DyParts d;
  d.n = n;
  for (int i = 0; i < 4; ++i) d.p[i] = (const bf16*)(i < n ? dy[i] : nullptr);
  return d;
}

extern "C" void hfl_lm_embed_bwd(const void* const* dy, int n_dy, const unsigned char* dmask, float p_drop,
                                 const void* xhat, const float* rstd, const void* gamma, const int* src, int M, int S,
                                 int E, float* demb, float* dpos, float* dgamma, float* dbeta, cudaStream_t st) {
  lm_embed_bwd_kernel<<<(M + LN_WARPS - 1) / LN_WARPS, LN_WARPS * 32, 0, st>>>(
      make_parts(dy, n_dy), dmask, p_drop, (const bf16*)xhat, rstd, (const bf16*)gamma, src, M, S, E, demb, dpos, dgamma,
      dbeta);
  count_launch();
}

extern "C" void hfl_lm_add_ln_fwd(const void* x, const void* h, const void* gamma, const void* beta, int M, int E,
                                  float eps, float p_drop, const long long* seed_p, const long long* step_p,
                                  unsigned site, void* y, void* xhat, float* rstd, unsigned char* dmask,
                                  cudaStream_t st) {
  lm_add_ln_fwd_kernel<<<(M + LN_WARPS - 1) / LN_WARPS, LN_WARPS * 32, 0, st>>>(
      (const bf16*)x, (const bf16*)h, (const bf16*)gamma, (const bf16*)beta, M, E, eps, p_drop, seed_p, step_p, site,
      (bf16*)y, (bf16*)xhat, rstd, dmask);
  count_launch();
}

extern "C" void hfl_lm_add_ln_bwd(const void* const* dy, int n_dy, const void* xhat, const float* rstd,
                                  const void* gamma, const unsigned char* dmask, float p_drop, int M, int E, void* dz,
                                  void* dh, float* dgamma, float* dbeta, float* dbias, float bias_alpha,
                                  cudaStream_t st) {
  lm_add_ln_bwd_kernel<<<(M + LN_WARPS - 1) / LN_WARPS, LN_WARPS * 32, 0, st>>>(
      make_parts(dy, n_dy), (const bf16*)xhat, rstd, (const bf16*)gamma, dmask, p_drop, M, E, (bf16*)dz, (bf16*)dh,
      dgamma, dbeta, dbias, bias_alpha);
  count_launch();
}

extern "C" void hfl_lm_gelu_drop_fwd(const void* u, long n, float p_drop, const long long* seed_p,
                                     const long long* step_p, unsigned site, void* h, unsigned char* dmask,
                                     cudaStream_t st) {
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  lm_gelu_drop_fwd_kernel<<<blocks, 256, 0, st>>>((const bf16*)u, n, p_drop, seed_p, step_p, site, (bf16*)h, dmask);
  count_launch();
}

extern "C" void hfl_lm_gelu_drop_bwd(const void* dh, const void* u, const unsigned char* dmask, float p_drop, int M,
                                     int F, void* du, float* dbias, float bias_alpha, cudaStream_t st) {
  dim3 grid((F + 127) / 128, (M + GB_ROWS - 1) / GB_ROWS);
  lm_gelu_drop_bwd_kernel<<<grid, 128, 0, st>>>((const bf16*)dh, (const bf16*)u, dmask, p_drop, M, F, (bf16*)du, dbias,
                                                bias_alpha);
  count_launch();
}

extern "C" void hfl_lm_colsum(const void* x, int M, int N, int ld, float alpha, float* out, cudaStream_t st) {
  dim3 grid((N + 127) / 128, (M + GB_ROWS - 1) / GB_ROWS);
  lm_colsum_kernel<<<grid, 128, 0, st>>>((const bf16*)x, M, N, ld, alpha, out);
  count_launch();
}

extern "C" void hfl_lm_attn_fwd(const void* q, const void* k, const void* v, int N, int S, int H, int dh, float scale,
                                void* o, void* P, cudaStream_t st) {
  lm_attn_fwd_kernel<<<N * H, 256, 0, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, S, H, dh, scale, (bf16*)o,
                                            (bf16*)P);
  count_launch();
}

extern "C" int hfl_lm_attn_bwd(const void* dO, const void* q, const void* k, const void* v, const void* P, int N, int S,
                               int H, int dh, float scale, void* dq, void* dk, void* dv, float* dbq, float* dbk,
                               float* dbv, float alpha_v, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lm_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_BWD_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  lm_attn_bwd_kernel<<<N * H, 256, ATTN_BWD_SMEM, st>>>((const bf16*)dO, (const bf16*)q, (const bf16*)k, (const bf16*)v,
                                                        (const bf16*)P, S, H, dh, scale, (bf16*)dq, (bf16*)dk, (bf16*)dv,
                                                        dbq, dbk, dbv, alpha_v);
  count_launch();
  return (int)cudaGetLastError();
}
