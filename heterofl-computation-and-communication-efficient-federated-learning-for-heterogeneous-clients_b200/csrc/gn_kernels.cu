// Fused GroupNorm / InstanceNorm / LayerNorm (+ residual add + ReLU) for the ResNet train step, forward and backward.
//
// Reference: `nn.GroupNorm(4, C)` / `GroupNorm(C, C)` / `GroupNorm(1, C)` behind a Scaler and followed by ReLU
// (src/models/resnet.py:18-26,44-50).  Statistics are per (sample, channel group) over (C/G channels x H x W), so unlike
// BatchNorm nothing crosses samples: one CTA owns one (sample, group) pair, makes two passes over its slab (the slab of
// a batch-10 step is L1/L2 resident) and writes s = x (+ res), y = relu(gamma * xhat + beta), mean and rstd.  The
// backward joins up to two incoming gradients and the identity-path gradient on load, applies the ReLU mask, reduces
// (sum dxhat, sum dxhat * xhat) per (sample, group), writes ds and leaves its per-sample dgamma / dbeta share for a tiny
// deterministic reduction kernel (no global atomics, no zero-fill).
// Activations are NHWC bf16; the Scaler is folded into epsilon by the caller (ops/nn.py).
#include "common.cuh"

namespace hfl {

constexpr int GN_THREADS = 256;

// gamma / beta live in the compute copy of the arena: bf16 under mixed precision, fp32 otherwise
__device__ __forceinline__ float ldp(const void* p, int i, int pb) {
  return pb ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}

// element index inside the (sample, group) slab -> (pixel, channel)
__device__ __forceinline__ void gn_decode(int i, int Cg, int& p, int& c) { p = i / Cg; c = i - p * Cg; }

__global__ void __launch_bounds__(GN_THREADS)
gn_act_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                  const void* __restrict__ gamma, const void* __restrict__ beta, int pb,
                  __nv_bfloat16* __restrict__ s_out, __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
                  float* __restrict__ rstd, int HW, int C, int G, float eps, int relu) {
  pdl_sync();
  __shared__ float sm[33];
  const int n = blockIdx.x / G, g = blockIdx.x - n * G;
  const int Cg = C / G, cnt = HW * Cg;
  const size_t base = (size_t)n * HW * C + (size_t)g * Cg;
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < cnt; i += GN_THREADS) {
    int p, c;
    gn_decode(i, Cg, p, c);
    const size_t o = base + (size_t)p * C + c;
    float v = __bfloat162float(x[o]);
    if (res != nullptr) {
      v = __bfloat162float(__float2bfloat16(v + __bfloat162float(res[o])));
      s_out[o] = __float2bfloat16(v);
    }
    a += v;
    b += v * v;
  }
  a = block_sum(a, sm);
  b = block_sum(b, sm);
  const float mu = a / (float)cnt;
  const float var = fmaxf(b / (float)cnt - mu * mu, 0.f);
  const float rs = rsqrtf(var + eps);
  if (threadIdx.x == 0) { mean[blockIdx.x] = mu; rstd[blockIdx.x] = rs; }
  const __nv_bfloat16* src = res != nullptr ? s_out : x;
  for (int i = threadIdx.x; i < cnt; i += GN_THREADS) {
    int p, c;
    gn_decode(i, Cg, p, c);
    const size_t o = base + (size_t)p * C + c;
    const int ch = g * Cg + c;
    float v = (__bfloat162float(src[o]) - mu) * rs * ldp(gamma, ch, pb) + ldp(beta, ch, pb);
    if (relu) v = fmaxf(v, 0.f);
    y[o] = __float2bfloat16(v);
  }
}

__global__ void __launch_bounds__(GN_THREADS)
gn_act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                  const __nv_bfloat16* __restrict__ dres, const __nv_bfloat16* __restrict__ s,
                  const float* __restrict__ mean, const float* __restrict__ rstd,
                  const void* __restrict__ gamma, const void* __restrict__ beta, int pb,
                  __nv_bfloat16* __restrict__ ds, float* __restrict__ part, int HW, int C, int G, int relu, int hw_bcast,
                  float dy_scale) {
  pdl_sync();
  // part: fp32 [N][2][C] per-sample shares of dgamma / dbeta (summed over samples by gn_param_grad_kernel)
  // hw_bcast > 0: dy is [N, C] and is broadcast over the hw_bcast positions of every sample (gradient of the global
  // average pool), scaled by dy_scale
  __shared__ float sm[33];
  extern __shared__ float chan[];          // [2][Cg] per-channel partial dgamma / dbeta of this (sample, group)
  const int n = blockIdx.x / G, g = blockIdx.x - n * G;
  const int Cg = C / G, cnt = HW * Cg;
  const size_t base = (size_t)n * HW * C + (size_t)g * Cg;
  const float mu = mean[blockIdx.x], rs = rstd[blockIdx.x];
  for (int c = threadIdx.x; c < 2 * Cg; c += GN_THREADS) chan[c] = 0.f;
  __syncthreads();
  // when Cg divides the block size a thread stays on one channel: keep its dgamma / dbeta share in registers
  const bool fixed_c = (GN_THREADS % Cg) == 0;
  float a = 0.f, b = 0.f, tg = 0.f, tb = 0.f;
  for (int i = threadIdx.x; i < cnt; i += GN_THREADS) {
    int p, c;
    gn_decode(i, Cg, p, c);
    const size_t o = base + (size_t)p * C + c;
    const int ch = g * Cg + c;
    const float xh = (__bfloat162float(s[o]) - mu) * rs;
    const float ga = ldp(gamma, ch, pb);
    float gr = hw_bcast > 0 ? __bfloat162float(dy[(size_t)n * C + ch]) * dy_scale : __bfloat162float(dy[o]);
    if (dy2 != nullptr) gr += __bfloat162float(dy2[o]);
    if (relu && !(xh * ga + ldp(beta, ch, pb) > 0.f)) gr = 0.f;
    if (fixed_c) {
      tg += gr * xh;
      tb += gr;
    } else {
      atomicAdd(&chan[c], gr * xh);
      atomicAdd(&chan[Cg + c], gr);
    }
    const float dxh = gr * ga;
    a += dxh;
    b += dxh * xh;
  }
  if (fixed_c && threadIdx.x < cnt) {
    const int c = threadIdx.x % Cg;
    atomicAdd(&chan[c], tg);
    atomicAdd(&chan[Cg + c], tb);
  }
  a = block_sum(a, sm);
  b = block_sum(b, sm);
  const float inv = 1.f / (float)cnt;
  for (int i = threadIdx.x; i < cnt; i += GN_THREADS) {
    int p, c;
    gn_decode(i, Cg, p, c);
    const size_t o = base + (size_t)p * C + c;
    const int ch = g * Cg + c;
    const float xh = (__bfloat162float(s[o]) - mu) * rs;
    const float ga = ldp(gamma, ch, pb);
    float gr = hw_bcast > 0 ? __bfloat162float(dy[(size_t)n * C + ch]) * dy_scale : __bfloat162float(dy[o]);
    if (dy2 != nullptr) gr += __bfloat162float(dy2[o]);
    if (relu && !(xh * ga + ldp(beta, ch, pb) > 0.f)) gr = 0.f;
    float v = rs * (gr * ga - a * inv - xh * b * inv);
    if (dres != nullptr) v += __bfloat162float(dres[o]);
    ds[o] = __float2bfloat16(v);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cg; c += GN_THREADS) {
    part[((size_t)n * 2 + 0) * C + g * Cg + c] = chan[c];
    part[((size_t)n * 2 + 1) * C + g * Cg + c] = chan[Cg + c];
  }
}

// dgamma[c] = sum_n part[n][0][c], dbeta[c] = sum_n part[n][1][c]; written in the parameter dtype
__global__ void gn_param_grad_kernel(const float* __restrict__ part, void* __restrict__ dgamma, void* __restrict__ dbeta,
                                     int N, int C, int pb) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) {
    a += part[((size_t)n * 2 + 0) * C + c];
    b += part[((size_t)n * 2 + 1) * C + c];
  }
  if (pb) {
    ((__nv_bfloat16*)dgamma)[c] = __float2bfloat16(a);
    ((__nv_bfloat16*)dbeta)[c] = __float2bfloat16(b);
  } else {
    ((float*)dgamma)[c] = a;
    ((float*)dbeta)[c] = b;
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" int hfl_gn_act_fwd(const void* x, const void* res, const void* gamma, const void* beta, int pb, void* s_out,
                              void* y, float* mean, float* rstd, int N, int HW, int C, int G, float eps, int relu,
                              cudaStream_t st) {
  if (G < 1 || C % G != 0) return -1;
  launch_k(gn_act_fwd_kernel, dim3(N * G), dim3(GN_THREADS), 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)res, gamma, beta, pb,
                                                 (__nv_bfloat16*)s_out, (__nv_bfloat16*)y, mean, rstd, HW, C, G, eps,
                                                 relu);
  count_launch();
  return (int)cudaGetLastError();
}

// part: fp32 scratch [N][2][C] (fully overwritten); dgamma / dbeta in the dtype of gamma
extern "C" int hfl_gn_act_bwd(const void* dy, const void* dy2, const void* dres, const void* s, const float* mean,
                              const float* rstd, const void* gamma, const void* beta, int pb, void* ds, float* part,
                              void* dgamma, void* dbeta, int N, int HW, int C, int G, int relu, int hw_bcast,
                              float dy_scale, cudaStream_t st) {
  if (G < 1 || C % G != 0) return -1;
  const int smem = 2 * (C / G) * (int)sizeof(float);
  launch_k(gn_act_bwd_kernel, dim3(N * G), dim3(GN_THREADS), smem, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2,
                                                    (const __nv_bfloat16*)dres, (const __nv_bfloat16*)s, mean, rstd,
                                                    gamma, beta, pb, (__nv_bfloat16*)ds, part, HW, C, G, relu, hw_bcast,
                                                    dy_scale);
  count_launch();
  launch_k(gn_param_grad_kernel, dim3((C + 127) / 128), dim3(128), 0, st, part, dgamma, dbeta, N, C, pb);
  count_launch();
  return (int)cudaGetLastError();
}
