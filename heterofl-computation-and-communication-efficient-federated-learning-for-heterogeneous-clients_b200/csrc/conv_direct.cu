// Direct NHWC convolutions for the NARROW sub-networks (levels c/d/e: 4..64 channels), forward / dgrad / wgrad.
//
// EXPERIMENTAL -- gated behind HFL_CONV_DIRECT=1 (ops/fused.py), not on any default path yet.
//
// Why: a level-e ResNet-18 step launches ~50 convolutions with C_in, C_out in {4, 8, 16, 32}; the whole layer is
// 1.5-12 MFLOP, far below one tensor-core tile, and the library serves it with generic kernels at 4-10 us apiece plus
// layout-conversion and memset nodes (profiles/step_census.md: ~65 % of the step).  At this size the right tool is the
// CUDA-core FMA pipe with maximal parallelism and zero set-up: no TMA descriptors, no TMEM allocation, one launch.
//
//   fprop : thread = (output pixel, KT output channels); the filter is transposed into shared memory as [tap][c][k] so a
//           thread's KT weights are one 8/16-byte shared load shared by all pixels of the CTA; inputs come straight from
//           global memory through L1 (neighbouring pixels and the K/KT threads of a pixel hit the same lines).
//   dgrad : same scheme with the roles of C and K swapped and the tap test `(h + pad - r) % stride == 0` (handles the
//           stride-2 layers and the 1x1 shortcuts); the filter is used in its native [k][tap][c] order.
//   wgrad : grid (pixel chunks, taps); a CTA owns all K x C products of one tap, stages 32 pixels of dy and of the
//           shifted x in shared memory per iteration and accumulates in registers; partials are combined with fp32
//           atomics only when there is more than one chunk.
//
// Layouts: activations [N, H, W, C] bf16 (channels-last storage), filters [K, R, S, C] bf16 (the arena layout),
// dW fp32 [K, R, S, C].  Reference semantics: F.conv2d(x, w, stride, padding = R // 2), src/models/resnet.py:33-48.
#include "common.cuh"

namespace hfl {

using bf16 = __nv_bfloat16;

struct DConvGeom {
  int N, H, W, C, K, OH, OW, R, S, stride, pad;
};

constexpr int DC_THREADS = 256;

// loads V consecutive bf16 (V in {1, 4, 8}) and widens them
template <int V>
__device__ __forceinline__ void load_bf16_vec(const bf16* p, float* out);
template <>
__device__ __forceinline__ void load_bf16_vec<8>(const bf16* p, float* out) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { out[2 * i] = __low2float(h[i]); out[2 * i + 1] = __high2float(h[i]); }
}
template <>
__device__ __forceinline__ void load_bf16_vec<4>(const bf16* p, float* out) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 2; ++i) { out[2 * i] = __low2float(h[i]); out[2 * i + 1] = __high2float(h[i]); }
}
template <>
__device__ __forceinline__ void load_bf16_vec<1>(const bf16* p, float* out) { out[0] = __bfloat162float(p[0]); }

template <int V>
__device__ __forceinline__ void store_bf16_vec(bf16* p, const float* v);
template <>
__device__ __forceinline__ void store_bf16_vec<8>(bf16* p, const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
template <>
__device__ __forceinline__ void store_bf16_vec<4>(bf16* p, const float* v) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 2; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---------------------------------------------------------------------------------------------------------------
// forward:  y[n,oh,ow,k] = sum_{r,s,c} x[n, oh*st + r - pad, ow*st + s - pad, c] * w[k,r,s,c]  (+ residual)
// OT = output channels per thread (4 or 8, divides K); IV = input-channel vector width (8, 4 or 1; divides C)
// ---------------------------------------------------------------------------------------------------------------
template <int OT, int IV>
__global__ void __launch_bounds__(DC_THREADS)
conv_direct_fprop_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ res,
                         bf16* __restrict__ y, const DConvGeom g) {
  extern __shared__ __align__(16) uint8_t dc_smem[];
  bf16* w_s = reinterpret_cast<bf16*>(dc_smem);                 // [tap*C + c][K]
  const int RS = g.R * g.S, RSC = RS * g.C;
  for (int i = threadIdx.x; i < g.K * RSC; i += DC_THREADS) {
    const int k = i / RSC, rem = i - k * RSC;
    w_s[rem * g.K + k] = w[i];
  }
  __syncthreads();
  const int kgroups = g.K / OT;
  const long total = (long)g.N * g.OH * g.OW * kgroups;
  const long idx = (long)blockIdx.x * DC_THREADS + threadIdx.x;
  if (idx >= total) return;
  const int kg = (int)(idx % kgroups);
  const long pix = idx / kgroups;
  const int ow = (int)(pix % g.OW);
  const long t = pix / g.OW;
  const int oh = (int)(t % g.OH), n = (int)(t / g.OH);
  float acc[OT];
#pragma unroll
  for (int j = 0; j < OT; ++j) acc[j] = 0.f;
  for (int r = 0; r < g.R; ++r) {
    const int ih = oh * g.stride + r - g.pad;
    if (ih < 0 || ih >= g.H) continue;
    for (int s = 0; s < g.S; ++s) {
      const int iw = ow * g.stride + s - g.pad;
      if (iw < 0 || iw >= g.W) continue;
      const bf16* xp = x + (((size_t)n * g.H + ih) * g.W + iw) * g.C;
      const bf16* wp = w_s + (size_t)((r * g.S + s) * g.C) * g.K + kg * OT;
      for (int c = 0; c < g.C; c += IV) {
        float xv[IV];
        load_bf16_vec<IV>(xp + c, xv);
#pragma unroll
        for (int cc = 0; cc < IV; ++cc) {
          float wv[OT];
          load_bf16_vec<OT>(wp + (size_t)(c + cc) * g.K, wv);
#pragma unroll
          for (int j = 0; j < OT; ++j) acc[j] = fmaf(xv[cc], wv[j], acc[j]);
        }
      }
    }
  }
  const size_t o = (size_t)pix * g.K + kg * OT;
  if (res != nullptr) {
    float rv[OT];
    load_bf16_vec<OT>(res + o, rv);
#pragma unroll
    for (int j = 0; j < OT; ++j) acc[j] += rv[j];
  }
  store_bf16_vec<OT>(y + o, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// dgrad:  dx[n,h,w,c] = sum_{r,s,k} dy[n, (h+pad-r)/st, (w+pad-s)/st, k] * w[k,r,s,c]   over taps with exact division
// OT = dx channels per thread (divides C); IV = dy-channel vector width (divides K)
// ---------------------------------------------------------------------------------------------------------------
template <int OT, int IV>
__global__ void __launch_bounds__(DC_THREADS)
conv_direct_dgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ w, bf16* __restrict__ dx,
                         const DConvGeom g) {
  extern __shared__ __align__(16) uint8_t dc_smem[];
  bf16* w_s = reinterpret_cast<bf16*>(dc_smem);                 // native order [k][tap][c]
  const int RS = g.R * g.S;
  for (int i = threadIdx.x; i < g.K * RS * g.C; i += DC_THREADS) w_s[i] = w[i];
  __syncthreads();
  const int cgroups = g.C / OT;
  const long total = (long)g.N * g.H * g.W * cgroups;
  const long idx = (long)blockIdx.x * DC_THREADS + threadIdx.x;
  if (idx >= total) return;
  const int cg = (int)(idx % cgroups);
  const long pix = idx / cgroups;
  const int wq = (int)(pix % g.W);
  const long t = pix / g.W;
  const int h = (int)(t % g.H), n = (int)(t / g.H);
  float acc[OT];
#pragma unroll
  for (int j = 0; j < OT; ++j) acc[j] = 0.f;
  for (int r = 0; r < g.R; ++r) {
    const int th = h + g.pad - r;
    if (th < 0 || (th % g.stride) != 0) continue;
    const int oh = th / g.stride;
    if (oh >= g.OH) continue;
    for (int s = 0; s < g.S; ++s) {
      const int tw = wq + g.pad - s;
      if (tw < 0 || (tw % g.stride) != 0) continue;
      const int ow = tw / g.stride;
      if (ow >= g.OW) continue;
      const bf16* dyp = dy + (((size_t)n * g.OH + oh) * g.OW + ow) * g.K;
      const int tap = r * g.S + s;
      for (int k = 0; k < g.K; k += IV) {
        float dv[IV];
        load_bf16_vec<IV>(dyp + k, dv);
#pragma unroll
        for (int kk = 0; kk < IV; ++kk) {
          float wv[OT];
          load_bf16_vec<OT>(w_s + ((size_t)(k + kk) * RS + tap) * g.C + cg * OT, wv);
#pragma unroll
          for (int j = 0; j < OT; ++j) acc[j] = fmaf(dv[kk], wv[j], acc[j]);
        }
      }
    }
  }
  store_bf16_vec<OT>(dx + (size_t)pix * g.C + cg * OT, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad:  dw[k,r,s,c] = sum_{n,oh,ow} dy[n,oh,ow,k] * x[n, oh*st + r - pad, ow*st + s - pad, c]
// grid (chunks, R*S); a CTA accumulates the K x C products of one tap over its pixel range.
// ---------------------------------------------------------------------------------------------------------------
constexpr int WG_PC = 32;        // pixels staged per iteration
constexpr int WG_MAXO = 16;      // outputs per thread: K*C <= 256 * 16 = 4096

__global__ void __launch_bounds__(DC_THREADS)
conv_direct_wgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, float* __restrict__ dw,
                         const DConvGeom g, int pix_per_chunk, int accumulate) {
  extern __shared__ __align__(16) uint8_t dc_smem[];
  float* dy_s = reinterpret_cast<float*>(dc_smem);              // [WG_PC][K]
  float* x_s = dy_s + WG_PC * g.K;                              // [WG_PC][C]
  const int tap = blockIdx.y, r = tap / g.S, s = tap - r * g.S;
  const long P = (long)g.N * g.OH * g.OW;
  const long p0 = (long)blockIdx.x * pix_per_chunk;
  const long p1 = min(P, p0 + pix_per_chunk);
  const int KC = g.K * g.C;
  float acc[WG_MAXO];
#pragma unroll
  for (int j = 0; j < WG_MAXO; ++j) acc[j] = 0.f;
  for (long pb = p0; pb < p1; pb += WG_PC) {
    const int np = (int)min((long)WG_PC, p1 - pb);
    for (int i = threadIdx.x; i < WG_PC * g.K; i += DC_THREADS) {
      const int p = i / g.K, k = i - p * g.K;
      dy_s[i] = p < np ? __bfloat162float(dy[(size_t)(pb + p) * g.K + k]) : 0.f;
    }
    for (int i = threadIdx.x; i < WG_PC * g.C; i += DC_THREADS) {
      const int p = i / g.C, c = i - p * g.C;
      float v = 0.f;
      if (p < np) {
        const long pix = pb + p;
        const int ow = (int)(pix % g.OW);
        const long t = pix / g.OW;
        const int oh = (int)(t % g.OH), n = (int)(t / g.OH);
        const int ih = oh * g.stride + r - g.pad, iw = ow * g.stride + s - g.pad;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
          v = __bfloat162float(x[(((size_t)n * g.H + ih) * g.W + iw) * g.C + c]);
      }
      x_s[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < WG_MAXO; ++j) {
      const int o = threadIdx.x + j * DC_THREADS;
      if (o < KC) {
        const int k = o / g.C, c = o - k * g.C;
        float a = acc[j];
#pragma unroll 8
        for (int p = 0; p < WG_PC; ++p) a = fmaf(dy_s[p * g.K + k], x_s[p * g.C + c], a);
        acc[j] = a;
      }
    }
    __syncthreads();
  }
  const int RS = g.R * g.S;
#pragma unroll
  for (int j = 0; j < WG_MAXO; ++j) {
    const int o = threadIdx.x + j * DC_THREADS;
    if (o < KC) {
      const int k = o / g.C, c = o - k * g.C;
      float* dst = dw + ((size_t)k * RS + tap) * g.C + c;
      if (accumulate) atomicAdd(dst, acc[j]);
      else *dst = acc[j];
    }
  }
}

}  // namespace hfl

using namespace hfl;

static DConvGeom make_geom(int N, int H, int W, int C, int K, int R, int S, int stride) {
  DConvGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.K = K; g.R = R; g.S = S; g.stride = stride; g.pad = R / 2;
  g.OH = (H + 2 * g.pad - R) / stride + 1;
  g.OW = (W + 2 * g.pad - S) / stride + 1;
  return g;
}

// 0 when the direct kernels cover the problem
extern "C" int hfl_conv_direct_supported(int N, int H, int W, int C, int K, int R, int S, int stride) {
  if (!((R == 3 && S == 3) || (R == 1 && S == 1))) return 1;
  if (stride != 1 && stride != 2) return 2;
  if (C > 64 || K > 64 || (K % 4) != 0) return 3;
  if ((C % 4) != 0 && C != 3 && C != 1) return 4;                  // stem: 3 (CIFAR) or 1 (MNIST) input channels
  if ((size_t)R * S * C * K * 2 > 96 * 1024) return 5;           // filter must fit in shared memory
  if ((size_t)K * C > 256 * WG_MAXO) return 6;
  if (N <= 0 || H <= 0 || W <= 0) return 7;
  return 0;
}

template <int OT, int IV, typename KernT>
static int set_smem(KernT kern, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

#define HFL_DC_DISPATCH(OTV, IVV, CALL)                                         \
  if (ot == OTV && iv == IVV) {                                                 \
    constexpr int OT = OTV, IV = IVV;                                           \
    CALL;                                                                       \
  }

extern "C" int hfl_conv_direct_fprop(const void* x, const void* w, const void* res, void* y, int N, int H, int W, int C,
                                     int K, int R, int S, int stride, cudaStream_t st) {
  if (hfl_conv_direct_supported(N, H, W, C, K, R, S, stride) != 0) return -2;
  const DConvGeom g = make_geom(N, H, W, C, K, R, S, stride);
  const int ot = (K % 8 == 0) ? 8 : 4;
  const int iv = (C % 8 == 0) ? 8 : ((C % 4 == 0) ? 4 : 1);
  const size_t smem = (size_t)R * S * C * K * 2;
  const long total = (long)N * g.OH * g.OW * (K / ot);
  const int blocks = (int)((total + DC_THREADS - 1) / DC_THREADS);
  int e = -1;
#define HFL_DC_FPROP                                                                                           \
  {                                                                                                            \
    auto kern = conv_direct_fprop_kernel<OT, IV>;                                                              \
    e = set_smem<OT, IV>(kern, smem);                                                                          \
    if (e == 0) kern<<<blocks, DC_THREADS, smem, st>>>((const bf16*)x, (const bf16*)w, (const bf16*)res, (bf16*)y, g); \
  }
  HFL_DC_DISPATCH(8, 8, HFL_DC_FPROP) HFL_DC_DISPATCH(8, 4, HFL_DC_FPROP) HFL_DC_DISPATCH(8, 1, HFL_DC_FPROP)
  HFL_DC_DISPATCH(4, 8, HFL_DC_FPROP) HFL_DC_DISPATCH(4, 4, HFL_DC_FPROP) HFL_DC_DISPATCH(4, 1, HFL_DC_FPROP)
#undef HFL_DC_FPROP
  if (e != 0) return e;
  count_launch();
  return (int)cudaGetLastError();
}

// dy: [N, OH, OW, K]; dx: [N, H, W, C]
extern "C" int hfl_conv_direct_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int C, int K, int R,
                                     int S, int stride, cudaStream_t st) {
  if (hfl_conv_direct_supported(N, H, W, C, K, R, S, stride) != 0 || (C % 4) != 0) return -2;
  const DConvGeom g = make_geom(N, H, W, C, K, R, S, stride);
  const int ot = (C % 8 == 0) ? 8 : 4;
  const int iv = (K % 8 == 0) ? 8 : 4;
  const size_t smem = (size_t)R * S * C * K * 2;
  const long total = (long)N * H * W * (C / ot);
  const int blocks = (int)((total + DC_THREADS - 1) / DC_THREADS);
  int e = -1;
#define HFL_DC_DGRAD                                                                                           \
  {                                                                                                            \
    auto kern = conv_direct_dgrad_kernel<OT, IV>;                                                              \
    e = set_smem<OT, IV>(kern, smem);                                                                          \
    if (e == 0) kern<<<blocks, DC_THREADS, smem, st>>>((const bf16*)dy, (const bf16*)w, (bf16*)dx, g);          \
  }
  HFL_DC_DISPATCH(8, 8, HFL_DC_DGRAD) HFL_DC_DISPATCH(8, 4, HFL_DC_DGRAD)
  HFL_DC_DISPATCH(4, 8, HFL_DC_DGRAD) HFL_DC_DISPATCH(4, 4, HFL_DC_DGRAD)
#undef HFL_DC_DGRAD
  if (e != 0) return e;
  count_launch();
  return (int)cudaGetLastError();
}

// dw: fp32 [K, R, S, C]; returns the number of pixel chunks (> 1: dw was zero-filled here and accumulated atomically)
extern "C" int hfl_conv_direct_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int C, int K, int R,
                                     int S, int stride, cudaStream_t st) {
  if (hfl_conv_direct_supported(N, H, W, C, K, R, S, stride) != 0) return -2;
  const DConvGeom g = make_geom(N, H, W, C, K, R, S, stride);
  const long P = (long)N * g.OH * g.OW;
  int chunks = (int)((P + 511) / 512);
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  int per = (int)((P + chunks - 1) / chunks);
  per = (per + WG_PC - 1) / WG_PC * WG_PC;
  chunks = (int)((P + per - 1) / per);
  if (chunks > 1) cudaMemsetAsync(dw, 0, (size_t)K * R * S * C * sizeof(float), st);
  const size_t smem = (size_t)WG_PC * (K + C) * sizeof(float);
  dim3 grid(chunks, R * S);
  conv_direct_wgrad_kernel<<<grid, DC_THREADS, smem, st>>>((const bf16*)dy, (const bf16*)x, dw, g, per, chunks > 1 ? 1 : 0);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? chunks : -(int)e;
}
