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
                               const unsigned char* __restrict__ mask, int rows, int V, int ld,
                               T* __restrict__ dlogits, float* __restrict__ row_loss,
                               float* __restrict__ row_correct) {
  pdl_sync();
  __shared__ float sm[33];
  __shared__ int s_arg;
  const int r = blockIdx.x;
  const T* z = logits + (size_t)r * ld;          // ld >= V: row stride (padded vocabularies)
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
    T* dz = dlogits + (size_t)r * ld;
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
  pdl_sync();
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
                              int rows, int V, int ld, void* dlogits, float* row_loss, float* row_correct, float n_weight,
                              float* loss_out, double* meter, cudaStream_t stream) {
  const int threads = V <= 64 ? 32 : (V <= 1024 ? 128 : 512);
  if (is_bf16)
    launch_k(ce_rows_kernel<__nv_bfloat16>, dim3(rows), dim3(threads), 0, stream, (const __nv_bfloat16*)logits, target, mask, rows, V, ld,
                                                                (__nv_bfloat16*)dlogits, row_loss, row_correct);
  else
    launch_k(ce_rows_kernel<float>, dim3(rows), dim3(threads), 0, stream, (const float*)logits, target, mask, rows, V, ld,
                                                        (float*)dlogits, row_loss, row_correct);
  launch_k(ce_finalize_kernel, dim3(1), dim3(256), 0, stream, row_loss, row_correct, rows, n_weight, loss_out, meter);
  count_launch(2);
}

// ---------------------------------------------------------------------------------------------------------
// Classifier head, forward + backward in ONE single-CTA launch:
//   pooled = mean_hw(y)  ->  logits = pooled W^T + b  ->  masked CE (loss, metrics)  ->  dW, db, dpooled
// (reference: adaptive_avg_pool2d, view, linear, masked_fill, cross_entropy + their 5 backward kernels,
//  src/models/resnet.py:149-157).  Sizes are tiny (N=10, C<=512, 10 classes): everything lives in shared memory.
// ---------------------------------------------------------------------------------------------------------
namespace hfl {

// pass 1: global average pool, one thread per (n, c); 20 small CTAs instead of one long loop
__global__ void head_pool_kernel(const __nv_bfloat16* __restrict__ y, int N, int HW, int C, float* __restrict__ pooled) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  const __nv_bfloat16* p = y + ((size_t)n * HW) * C + c;
  float s = 0.f;
  for (int h = 0; h < HW; ++h) s += __bfloat162float(p[(size_t)h * C]);
  pooled[i] = __bfloat162float(__float2bfloat16(s / (float)HW));   // bf16-rounded like the unfused mean()
}

// pass 2: every CTA recomputes the (tiny) logits + CE, then produces its channel slice of dW / dpooled
template <typename PT>
__global__ void __launch_bounds__(512)
head_fwd_bwd_kernel(const float* __restrict__ pooled_g, int N, int C, const PT* __restrict__ W,
                    const PT* __restrict__ bias, int Kc, const long long* __restrict__ target,
                    const unsigned char* __restrict__ mask, float n_weight, float* __restrict__ loss_out,
                    double* __restrict__ meter, PT* __restrict__ dW, PT* __restrict__ db,
                    __nv_bfloat16* __restrict__ dpooled, const int* __restrict__ nvalid) {
  pdl_sync();
  // nvalid != nullptr: only the first *nvalid rows are real samples (partial batch inside a captured step); the loss
  // is their mean, the padding rows get zero gradient and the meter weighs the step by *nvalid.
  extern __shared__ float sm[];
  const int Nv = nvalid ? max(min(*nvalid, N), 1) : N;
  if (nvalid) n_weight = (float)Nv;
  float* pooled = sm;                 // [N, C]
  float* logit = pooled + N * C;      // [N, Kc]  (becomes dlogits)
  float* red = logit + N * Kc;        // [N] row losses, [N] correct
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < N * C; i += nt) pooled[i] = pooled_g[i];
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int o = warp; o < N * Kc; o += nw) {
    const int n = o / Kc, k = o - n * Kc;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += pooled[n * C + c] * (float)W[(size_t)k * C + c];
    s = warp_sum(s);
    if (lane == 0) {
      float z = __bfloat162float(__float2bfloat16(s + (float)bias[k]));
      if (mask != nullptr && !mask[k]) z = 0.f;
      logit[o] = z;
    }
  }
  __syncthreads();
  for (int n = warp; n < N; n += nw) {
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    for (int k = lane; k < Kc; k += 32) { const float z = logit[n * Kc + k]; if (z > mx) { mx = z; arg = k; } }
    const float bmx = warp_max(mx);
    int cand = (mx == bmx) ? arg : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    float se = 0.f;
    for (int k = lane; k < Kc; k += 32) se += __expf(logit[n * Kc + k] - bmx);
    se = warp_sum(se);
    const float lse = __logf(se) + bmx;
    const int t = (int)target[n];
    const bool real = n < Nv;
    if (lane == 0) { red[n] = real ? lse - logit[n * Kc + t] : 0.f; red[N + n] = (real && cand == t) ? 1.f : 0.f; }
    __syncwarp();
    for (int k = lane; k < Kc; k += 32) {
      const bool on = real && (mask == nullptr || mask[k]);
      const float g = on ? (__expf(logit[n * Kc + k] - lse) - (k == t ? 1.f : 0.f)) / (float)Nv : 0.f;
      logit[n * Kc + k] = __bfloat162float(__float2bfloat16(g));
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid == 0) {
      float a = 0.f, c = 0.f;
      for (int n = 0; n < N; ++n) { a += red[n]; c += red[N + n]; }
      const float loss = a / (float)Nv;
      *loss_out = loss;
      if (meter != nullptr) {
        meter[0] += (double)loss * n_weight;
        meter[1] += (double)c;
        meter[2] += (double)n_weight;
        meter[3] += (double)__expf(loss) * n_weight;
      }
    }
    for (int k = tid; k < Kc; k += nt) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += logit[n * Kc + k];
      db[k] = (PT)s;
    }
  }
  // channel slice of this CTA
  const int cs = (C + gridDim.x - 1) / gridDim.x;
  const int c0 = blockIdx.x * cs, c1 = min(c0 + cs, C);
  const int wsl = c1 - c0;
  if (wsl <= 0) return;
  for (int i = tid; i < Kc * wsl; i += nt) {
    const int k = i / wsl, c = c0 + (i - k * wsl);
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += logit[n * Kc + k] * pooled[n * C + c];
    dW[(size_t)k * C + c] = (PT)s;
  }
  for (int i = tid; i < N * wsl; i += nt) {
    const int n = i / wsl, c = c0 + (i - n * wsl);
    float s = 0.f;
    for (int k = 0; k < Kc; ++k) s += logit[n * Kc + k] * (float)W[(size_t)k * C + c];
    dpooled[(size_t)n * C + c] = __float2bfloat16(s);
  }
}

}  // namespace hfl

extern "C" int hfl_head_smem_bytes(int N, int C, int Kc) { return (N * C + N * Kc + 2 * N) * (int)sizeof(float); }

extern "C" int hfl_head_fwd_bwd(const void* y, int N, int HW, int C, const void* W, const void* bias, int param_bf16,
                                int Kc, const long long* target, const unsigned char* mask, float n_weight,
                                float* loss_out, double* meter, void* dW, void* db, void* dpooled, float* pooled_ws,
                                const int* nvalid, cudaStream_t stream) {
  const int smem = hfl_head_smem_bytes(N, C, Kc);
  if (smem > 200 * 1024) return -1;
  hfl::launch_k(hfl::head_pool_kernel, dim3((N * C + 255) / 256), dim3(256), 0, stream, (const __nv_bfloat16*)y, N, HW, C, pooled_ws);
  const int grid = C >= 128 ? C / 64 : 1;
  cudaError_t e = cudaSuccess;
  if (param_bf16) {
    auto kern = hfl::head_fwd_bwd_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    launch_k(kern, dim3(grid), dim3(512), smem, stream, pooled_ws, N, C, (const __nv_bfloat16*)W, (const __nv_bfloat16*)bias, Kc, target,
                                     mask, n_weight, loss_out, meter, (__nv_bfloat16*)dW, (__nv_bfloat16*)db,
                                     (__nv_bfloat16*)dpooled, nvalid);
  } else {
    auto kern = hfl::head_fwd_bwd_kernel<float>;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    launch_k(kern, dim3(grid), dim3(512), smem, stream, pooled_ws, N, C, (const float*)W, (const float*)bias, Kc, target, mask, n_weight,
                                     loss_out, meter, (float*)dW, (float*)db, (__nv_bfloat16*)dpooled, nvalid);
  }
  hfl::count_launch(2);
  return (int)cudaGetLastError();
}
