// Masked softmax cross-entropy, forward + backward + metrics in one pass over the logits.
//
// Reference (src/models/resnet.py:152-157, src/models/transformer.py:154-161): build a label mask, masked_fill the
// logits of absent classes with 0 (NOT -inf: they still sit in the softmax denominator with value 0), then
// F.cross_entropy; accuracy/perplexity are computed afterwards with extra kernels and a host sync per step
// (src/metrics/metrics.py:7-25).  Here one block per row produces loss, d(loss)/d(logits) and the correct-prediction
// flag; a 1-block finaliser reduces the rows and bumps an on-device meter [sum loss*n, sum correct, sum n, sum exp(loss)*n].
#include "common.cuh"

namespace hfl {

template <typename T> __device__ __forceinline__ float ldf(const T* p, size_t i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T* p, size_t i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, size_t i, float v) { p[i] = __float2bfloat16(v); }

template <typename T>
__global__ void ce_rows_kernel(const T* __restrict__ logits, const long long* __restrict__ target,
                               const unsigned char* __restrict__ mask, int rows, int V, T* __restrict__ dlogits,
                               float* __restrict__ row_loss, float* __restrict__ row_correct) {
  __shared__ float sm[33];
  __shared__ int s_arg;
  const int r = blockIdx.x;
  const T* z = logits + (size_t)r * V;
  const int tgt = (int)target[r];
  // pass 1: max and argmax of the masked logits
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    const float v = (mask == nullptr || mask[j]) ? ldf(z, j) : 0.f;
    if (v > mx) { mx = v; arg = j; }
  }
  const float bmx = block_max(mx, sm);
  if (threadIdx.x == 0) s_arg = 0x7fffffff;
  __syncthreads();
  if (mx == bmx) atomicMin(&s_arg, arg);
  // pass 2: sum exp
  float se = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    const float v = (mask == nullptr || mask[j]) ? ldf(z, j) : 0.f;
    se += __expf(v - bmx);
  }
  se = block_sum(se, sm);
  const float lse = __logf(se) + bmx;
  if (threadIdx.x == 0) {
    const float zt = (mask == nullptr || mask[tgt]) ? ldf(z, tgt) : 0.f;
    row_loss[r] = lse - zt;
    row_correct[r] = (s_arg == tgt) ? 1.f : 0.f;
  }
  // pass 3: gradient of the MEAN loss
  if (dlogits != nullptr) {
    const float inv_rows = 1.f / (float)rows;
    T* dz = dlogits + (size_t)r * V;
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      const bool on = (mask == nullptr || mask[j]);
      float gsm = 0.f;
      if (on) gsm = (__expf(ldf(z, j) - lse) - (j == tgt ? 1.f : 0.f)) * inv_rows;
      stf(dz, j, gsm);
    }
  }
}

__global__ void ce_finalize_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_correct, int rows,
                                   float n_weight, float* __restrict__ loss_out, double* __restrict__ meter) {
  __shared__ float sm[33];
  float a = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) { a += row_loss[i]; c += row_correct[i]; }
  a = block_sum(a, sm);
  c = block_sum(c, sm);
  if (threadIdx.x == 0) {
    const float loss = a / (float)rows;
    *loss_out = loss;
    if (meter != nullptr) {
      meter[0] += (double)loss * n_weight;
      meter[1] += (double)c;
      meter[2] += (double)n_weight;
      meter[3] += (double)__expf(loss) * n_weight;
    }
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" void hfl_masked_ce(const void* logits, int is_bf16, const long long* target, const unsigned char* mask,
                              int rows, int V, void* dlogits, float* row_loss, float* row_correct, float n_weight,
                              float* loss_out, double* meter, cudaStream_t stream) {
  const int threads = V <= 64 ? 32 : (V <= 1024 ? 128 : 512);
  if (is_bf16)
    ce_rows_kernel<__nv_bfloat16><<<rows, threads, 0, stream>>>((const __nv_bfloat16*)logits, target, mask, rows, V,
                                                                (__nv_bfloat16*)dlogits, row_loss, row_correct);
  else
    ce_rows_kernel<float><<<rows, threads, 0, stream>>>((const float*)logits, target, mask, rows, V, (float*)dlogits,
                                                        row_loss, row_correct);
  ce_finalize_kernel<<<1, 256, 0, stream>>>(row_loss, row_correct, rows, n_weight, loss_out, meter);
  count_launch(2);
}
