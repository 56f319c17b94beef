// Shared helpers for the hfl_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <utility>

namespace hfl {

// Host-side launch accounting (bench.py reports how many of OUR kernels ran inside the timed region).
void count_launch(int n = 1);

// ---------------------------------------------------------------------------------- programmatic dependent launch
// A train step is a chain of ~100-200 small dependent kernels (3-10 us each); between two kernel nodes of a stream /
// captured graph the GPU otherwise drains kernel N completely before it starts fetching, scheduling and running the
// prologue of kernel N+1 (barrier init, TMEM allocation, tensor-map prefetch).  Every kernel of this library
//   * is launched through launch_k(), which sets cudaLaunchAttributeProgrammaticStreamSerialization (captured into
//     CUDA graphs as a programmatic edge), and
//   * executes pdl_wait() (griddepcontrol.wait: returns once ALL prerequisite grids have completed and their memory is
//     visible) before its first access to global data, followed by pdl_trigger() (griddepcontrol.launch_dependents) so
//     the next kernel may become resident and run ITS prologue while this one computes.
// Transitivity: a grid completes only after every one of its CTAs has exited, and every CTA that reaches its main body
// has passed pdl_wait(); so when a dependent's wait returns, the whole chain before it has completed.  The instructions
// are no-ops in a kernel that was launched without the attribute (HFL_PDL=0).
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

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
