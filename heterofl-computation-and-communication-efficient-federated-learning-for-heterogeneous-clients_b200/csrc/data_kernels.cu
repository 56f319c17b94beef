// Batch assembly: gather + RandomCrop(pad 4, zero fill) + horizontal flip + ToTensor + Normalize in one kernel,
// emitting channels-last activations in the compute dtype (reference: CPU PIL pipeline + H2D copy per step,
// src/data.py:19-27,113-119).
#include "common.cuh"

namespace hfl {

template <typename T>
__global__ void augment_kernel(const unsigned char* __restrict__ src, const long long* __restrict__ idx,
                               const int* __restrict__ params, const float* __restrict__ mean,
                               const float* __restrict__ inv_std, int B, int H, int W, int C, int pad,
                               T* __restrict__ out) {
  pdl_sync();
  const int total = B * H * W * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C;
    int r = i / C;
    const int x = r % W; r /= W;
    const int y = r % H;
    const int b = r / H;
    int sy = y, sx = x;
    if (params != nullptr) {
      const int dy = params[b * 3 + 0], dx = params[b * 3 + 1], flip = params[b * 3 + 2];
      sy = y + dy - pad;
      sx = (flip ? (W - 1 - x) : x) + dx - pad;
    }
    float v = 0.f;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
      const long long n = idx ? idx[b] : (long long)b;
      v = (float)src[((n * H + sy) * W + sx) * C + c] * (1.f / 255.f);
    }
    v = (v - mean[c]) * inv_std[c];
    if constexpr (sizeof(T) == 2) out[i] = __float2bfloat16(v); else out[i] = v;
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" void hfl_augment(const unsigned char* src, const long long* idx, const int* params, const float* mean,
                            const float* inv_std, int B, int H, int W, int C, int pad, void* out, int out_bf16,
                            cudaStream_t stream) {
  const int total = B * H * W * C;
  if (total <= 0) return;
  int blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (out_bf16)
    launch_k(augment_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, src, idx, params, mean, inv_std, B, H, W, C, pad,
                                                             (__nv_bfloat16*)out);
  else
    launch_k(augment_kernel<float>, dim3(blocks), dim3(256), 0, stream, src, idx, params, mean, inv_std, B, H, W, C, pad, (float*)out);
  count_launch();
}
