// Shared helpers for the hfl_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace hfl {

// Host-side launch accounting (bench.py reports how many of OUR kernels ran inside the timed region).
void count_launch(int n = 1);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (result valid in every thread).
__device__ __forceinline__ float block_sum(float v, float* smem /* >= 33 floats */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.f;
  if (wid == 0) {
    r = warp_sum(r);
    if (lane == 0) smem[32] = r;
  }
  __syncthreads();
  r = smem[32];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_max(float v, float* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? smem[threadIdx.x] : -INFINITY;
  if (wid == 0) {
    r = warp_max(r);
    if (lane == 0) smem[32] = r;
  }
  __syncthreads();
  r = smem[32];
  __syncthreads();
  return r;
}

}  // namespace hfl
