// 2x2 / stride-2 max pooling on NHWC bf16 activations, forward and backward (MNIST CNN train step).
//
// Reference: `nn.MaxPool2d(2)` after every conv block but the last (src/models/conv.py:29-45); odd extents floor
// (7 -> 3), the dropped last row / column receives no gradient.  The backward pass recomputes the arg-max from the saved
// pre-pool activation instead of storing indices (the first maximum in scan order wins a tie, as in ATen); one thread
// owns one pooled (pixel, 8-channel) vector so every global access is a 16-byte transaction.
#include "common.cuh"

namespace hfl {

__device__ __forceinline__ void unpack8h(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ uint4 pack8h(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// VEC = 8 (C % 8 == 0, 16-byte accesses) or 1 (any C)
template <int VEC>
__global__ void maxpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                    int W, int C) {
  pdl_sync();
  const int OH = H / 2, OW = W / 2, CV = C / VEC;
  const long total = (long)N * OH * OW * CV;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long t = i / CV;
    const int ow = (int)(t % OW);
    t /= OW;
    const int oh = (int)(t % OH), n = (int)(t / OH);
    const size_t src = (((size_t)n * H + 2 * oh) * W + 2 * ow) * C + (size_t)cv * VEC;
    const size_t dst = (((size_t)n * OH + oh) * OW + ow) * C + (size_t)cv * VEC;
    if (VEC == 8) {
      float a[8], b[8];
      unpack8h(*reinterpret_cast<const uint4*>(x + src), a);
      unpack8h(*reinterpret_cast<const uint4*>(x + src + C), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], b[e]);
      unpack8h(*reinterpret_cast<const uint4*>(x + src + (size_t)W * C), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], b[e]);
      unpack8h(*reinterpret_cast<const uint4*>(x + src + (size_t)W * C + C), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], b[e]);
      *reinterpret_cast<uint4*>(y + dst) = pack8h(a);
    } else {
      const float m = fmaxf(fmaxf(__bfloat162float(x[src]), __bfloat162float(x[src + C])),
                            fmaxf(__bfloat162float(x[src + (size_t)W * C]), __bfloat162float(x[src + (size_t)W * C + C])));
      y[dst] = __float2bfloat16(m);
    }
  }
}

// dx[n, h, w, c] = dy[n, h/2, w/2, c] at the window's first maximum, 0 elsewhere (and on a dropped odd row / column).
// One thread per INPUT-resolution 2x2 window (incl. the partial windows of an odd edge, which only write zeros).
template <int VEC>
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                    __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C) {
  pdl_sync();
  const int OH = H / 2, OW = W / 2, CV = C / VEC;
  const int WH = (H + 1) / 2, WW = (W + 1) / 2;      // windows incl. the partial ones
  const long total = (long)N * WH * WW * CV;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long t = i / CV;
    const int ww = (int)(t % WW);
    t /= WW;
    const int wh = (int)(t % WH), n = (int)(t / WH);
    const size_t p00 = (((size_t)n * H + 2 * wh) * W + 2 * ww) * C + (size_t)cv * VEC;
    const bool full = wh < OH && ww < OW;
    if (!full) {   // dropped edge: zero whatever part of the window exists
      for (int dh = 0; dh < 2; ++dh)
        for (int dw = 0; dw < 2; ++dw) {
          if (2 * wh + dh >= H || 2 * ww + dw >= W) continue;
          const size_t o = p00 + ((size_t)dh * W + dw) * C;
          if (VEC == 8) *reinterpret_cast<uint4*>(dx + o) = make_uint4(0, 0, 0, 0);
          else dx[o] = __float2bfloat16(0.f);
        }
      continue;
    }
    const size_t q = (((size_t)n * OH + wh) * OW + ww) * C + (size_t)cv * VEC;
    const size_t offs[4] = {0, (size_t)C, (size_t)W * C, (size_t)W * C + C};
    if (VEC == 8) {
      float v[4][8], g[8], o[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) unpack8h(*reinterpret_cast<const uint4*>(x + p00 + offs[k]), v[k]);
      unpack8h(*reinterpret_cast<const uint4*>(dy + q), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int arg = 0;
        float m = v[0][e];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (v[k][e] > m) { m = v[k][e]; arg = k; }
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k][e] = (k == arg) ? g[e] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dx + p00 + offs[k]) = pack8h(o[k]);
    } else {
      int arg = 0;
      float m = __bfloat162float(x[p00]);
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        const float vk = __bfloat162float(x[p00 + offs[k]]);
        if (vk > m) { m = vk; arg = k; }
      }
      const __nv_bfloat16 gq = dy[q];
#pragma unroll
      for (int k = 0; k < 4; ++k) dx[p00 + offs[k]] = (k == arg) ? gq : __float2bfloat16(0.f);
    }
  }
}

}  // namespace hfl

static int pool_grid(long total) {
  long b = (total + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int hfl_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, cudaStream_t st) {
  const bool v8 = (C % 8) == 0;
  const long total = (long)N * (H / 2) * (W / 2) * (v8 ? C / 8 : C);
  if (v8)
    hfl::launch_k(hfl::maxpool2_fwd_kernel<8>, dim3(pool_grid(total)), dim3(256), 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C);
  else
    hfl::launch_k(hfl::maxpool2_fwd_kernel<1>, dim3(pool_grid(total)), dim3(256), 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C);
  hfl::count_launch();
  return (int)cudaGetLastError();
}

extern "C" int hfl_maxpool2_bwd(const void* dy, const void* x, void* dx, int N, int H, int W, int C, cudaStream_t st) {
  const bool v8 = (C % 8) == 0;
  const long total = (long)N * ((H + 1) / 2) * ((W + 1) / 2) * (v8 ? C / 8 : C);
  if (v8)
    hfl::launch_k(hfl::maxpool2_bwd_kernel<8>, dim3(pool_grid(total)), dim3(256), 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                                 (__nv_bfloat16*)dx, N, H, W, C);
  else
    hfl::launch_k(hfl::maxpool2_bwd_kernel<1>, dim3(pool_grid(total)), dim3(256), 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                                 (__nv_bfloat16*)dx, N, H, W, C);
  hfl::count_launch();
  return (int)cudaGetLastError();
}
