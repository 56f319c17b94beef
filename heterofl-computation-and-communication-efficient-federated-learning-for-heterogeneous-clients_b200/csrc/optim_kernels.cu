// Fused clip_grad_norm_(.,max_norm) + SGD(momentum, weight_decay) over a client's FLAT parameter arena.
//
// Reference per step (src/train_classifier_fed.py:202-206): zero_grad, backward, clip_grad_norm_ (one norm kernel per
// tensor + stack + norm + per-tensor scale), SGD.step (foreach over 56-74 tensors).  Here: 2 launches total.
//   pass 1: ||g||^2 (block partials -> one atomicAdd each)
//   pass 2: coef = min(1, max_norm/(||g||+1e-6));  g' = coef*g + wd*w;  m = mu*m + g';  w -= lr*m;
//           also refreshes the bf16 compute copy of w and zeroes g for the next backward.
#include "common.cuh"

namespace hfl {

template <typename G> struct GradVec;
template <> struct GradVec<float> {
  static __device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void zero4(float* p) { *reinterpret_cast<float4*>(p) = make_float4(0, 0, 0, 0); }
};
template <> struct GradVec<__nv_bfloat16> {
  static __device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void zero4(__nv_bfloat16* p) { *reinterpret_cast<uint2*>(p) = make_uint2(0, 0); }
};

template <typename G>
__global__ void __launch_bounds__(256) grad_norm_sq_kernel(const G* __restrict__ g, int n4, float* __restrict__ scratch) {
  pdl_sync();
  __shared__ float sm[33];
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float v[4];
    GradVec<G>::load4(g + 4 * (size_t)i, v);
    acc += v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) atomicAdd(scratch, acc);
}

template <typename G>
__global__ void __launch_bounds__(256)
sgd_update_kernel(float* __restrict__ w, __nv_bfloat16* __restrict__ wc, G* __restrict__ g, float* __restrict__ mom,
                  int n4, const float* __restrict__ lr_p, float mu, float wd, float max_norm,
                  float* __restrict__ scratch /* [0]=norm^2, [1]=done-block counter (as uint) */) {
  pdl_sync();
  const float total = sqrtf(*reinterpret_cast<volatile float*>(scratch));
  const float coef = fminf(1.f, max_norm / (total + 1e-6f));
  const float lr = *lr_p;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const size_t o = 4 * (size_t)i;
    float gv[4];
    GradVec<G>::load4(g + o, gv);
    float4 wv = *reinterpret_cast<const float4*>(w + o);
    float4 mv = *reinterpret_cast<const float4*>(mom + o);
    float* wp = reinterpret_cast<float*>(&wv);
    float* mp = reinterpret_cast<float*>(&mv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gg = gv[k] * coef + wd * wp[k];
      mp[k] = mu * mp[k] + gg;
      wp[k] -= lr * mp[k];
    }
    *reinterpret_cast<float4*>(w + o) = wv;
    *reinterpret_cast<float4*>(mom + o) = mv;
    if (wc) {
      __nv_bfloat162 a = __floats2bfloat162_rn(wp[0], wp[1]);
      __nv_bfloat162 b = __floats2bfloat162_rn(wp[2], wp[3]);
      uint2 t;
      t.x = *reinterpret_cast<unsigned*>(&a);
      t.y = *reinterpret_cast<unsigned*>(&b);
      *reinterpret_cast<uint2*>(wc + o) = t;
    }
    GradVec<G>::zero4(g + o);
  }
  // the last block to finish re-arms the scratch for the next step (every block has read scratch[0] by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned* cnt = reinterpret_cast<unsigned*>(scratch + 1);
    const unsigned done = atomicAdd(cnt, 1u);
    if (done == gridDim.x - 1) {
      scratch[0] = 0.f;
      *cnt = 0u;
      scratch[2] = total;  // last observed grad norm (diagnostics)
    }
  }
}

}  // namespace hfl

using namespace hfl;

static int grid_for(int n4) {
  int b = (n4 + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return b;
}

extern "C" void hfl_sgd_clip_step(float* w, void* wc, void* g, int g_is_bf16, float* mom, long n,
                                  const float* lr, float mu, float wd, float max_norm, float* scratch,
                                  cudaStream_t stream) {
  const int n4 = (int)(n / 4);
  const int grid = grid_for(n4);
  if (g_is_bf16) {
    launch_k(grad_norm_sq_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, stream, (const __nv_bfloat16*)g, n4, scratch);
    launch_k(sgd_update_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, stream, w, (__nv_bfloat16*)wc, (__nv_bfloat16*)g, mom, n4, lr,
                                                              mu, wd, max_norm, scratch);
  } else {
    launch_k(grad_norm_sq_kernel<float>, dim3(grid), dim3(256), 0, stream, (const float*)g, n4, scratch);
    launch_k(sgd_update_kernel<float>, dim3(grid), dim3(256), 0, stream, w, (__nv_bfloat16*)wc, (float*)g, mom, n4, lr, mu, wd,
                                                      max_norm, scratch);
  }
  count_launch(2);
}

// ---------------------------------------------------------------------------------------------------------
// Multi-tensor variant: gradients live wherever their producer left them (cuDNN wgrad outputs inside the CUDA-graph
// pool, fused-norm kernels' dgamma/dbeta, ...) and are described by a segment table
//     seg[i] = { grad pointer, arena offset, numel, grad-is-bf16 }
// while w / wc / mom stay flat arenas.  No gradient copy, no accumulate, no zeroing pass.
// ---------------------------------------------------------------------------------------------------------
namespace hfl {

struct Seg { const void* g; long off; long n; long bf16; };
constexpr int MS_CHUNK = 4096;

// s.g == nullptr: a segment whose gradient is applied separately (row-sparse embedding table, below); the dense pass
// runs its weight-decay / momentum part with a zero gradient and the norm pass skips it
__device__ __forceinline__ float seg_grad(const Seg& s, long i) {
  if (s.g == nullptr) return 0.f;
  return s.bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(s.g)[i])
                : reinterpret_cast<const float*>(s.g)[i];
}

__global__ void __launch_bounds__(256) multi_norm_sq_kernel(const Seg* __restrict__ segs, const int2* __restrict__ blocks,
                                                            float* __restrict__ scratch) {
  pdl_sync();
  __shared__ float sm[33];
  const int2 b = blocks[blockIdx.x];
  const Seg s = segs[b.x];
  if (s.g == nullptr) return;                       // whole CTA: no gradient array to read
  const long end = min((long)b.y + MS_CHUNK, s.n);
  float acc = 0.f;
  for (long i = b.y + threadIdx.x; i < end; i += 256) {
    const float v = seg_grad(s, i);
    acc += v * v;
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) atomicAdd(scratch, acc);
}

__global__ void __launch_bounds__(256)
multi_sgd_update_kernel(const Seg* __restrict__ segs, const int2* __restrict__ blocks, float* __restrict__ w,
                        __nv_bfloat16* __restrict__ wc, float* __restrict__ mom, const float* __restrict__ lr_p,
                        float mu, float wd, float max_norm, float* __restrict__ scratch) {
  pdl_sync();
  const float total = sqrtf(*reinterpret_cast<volatile float*>(scratch));
  const float coef = fminf(1.f, max_norm / (total + 1e-6f));
  const float lr = *lr_p;
  const int2 b = blocks[blockIdx.x];
  const Seg s = segs[b.x];
  const long end = min((long)b.y + MS_CHUNK, s.n);
  for (long i = b.y + threadIdx.x; i < end; i += 256) {
    const long o = s.off + i;
    const float wv = w[o];
    const float gg = seg_grad(s, i) * coef + wd * wv;
    const float mv = mu * mom[o] + gg;
    const float nw = wv - lr * mv;
    mom[o] = mv;
    w[o] = nw;
    if (wc) wc[o] = __float2bfloat16(nw);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned* cnt = reinterpret_cast<unsigned*>(scratch + 1);
    const unsigned done = atomicAdd(cnt, 1u);
    if (done == gridDim.x - 1) {
      scratch[0] = 0.f;
      *cnt = 0u;
      scratch[2] = total;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Row-sparse gradient of an embedding table (masked-LM step: <= tokens-per-step rows of the 33 279 x E table are touched).
// SGD with momentum and weight decay is affine in the gradient:
//     mom' = mu*mom + wd*w + c*g,   w' = w - lr*mom'      (c = clip coefficient)
// so the dense pass above runs with g = 0 for this segment (it never reads, and nobody has to zero, the 34 MB gradient
// array) and the touched rows receive the missing  mom += c*g,  w -= lr*c*g  afterwards -- the same update up to fp32
// rounding order.  `src[m]` = table row read by token m (duplicates allowed: the first occurrence owns the row).
//   sparse_norm : adds sum g^2 of the unique touched rows to scratch[0]      (BEFORE the dense norm / update)
//   sparse_apply: applies the correction, refreshes the bf16 copy, zeroes the rows of g   (AFTER the dense update, which
//                 leaves the total gradient norm in scratch[2])
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool first_occurrence(const int* __restrict__ src, int m) {
  const int r = src[m];
  int dup = 0;
  for (int t = threadIdx.x; t < m; t += blockDim.x) dup |= (src[t] == r);
  return __syncthreads_or(dup) == 0;
}

__global__ void __launch_bounds__(128)
emb_sparse_norm_kernel(const int* __restrict__ src, const float* __restrict__ g, int E, float* __restrict__ scratch) {
  pdl_sync();
  __shared__ float sm[33];
  const int m = blockIdx.x;
  if (!first_occurrence(src, m)) return;
  const float* row = g + (size_t)src[m] * E;
  float acc = 0.f;
  for (int c = threadIdx.x; c < E; c += blockDim.x) acc += row[c] * row[c];
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) atomicAdd(scratch, acc);
}

__global__ void __launch_bounds__(128)
emb_sparse_apply_kernel(const int* __restrict__ src, float* __restrict__ g, int E, long off, float* __restrict__ w,
                        __nv_bfloat16* __restrict__ wc, float* __restrict__ mom, const float* __restrict__ lr_p,
                        float max_norm, const float* __restrict__ scratch) {
  pdl_sync();
  const int m = blockIdx.x;
  if (!first_occurrence(src, m)) return;
  const float coef = fminf(1.f, max_norm / (scratch[2] + 1e-6f));
  const float lr = *lr_p;
  const size_t r = (size_t)src[m] * E;
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    const float gv = g[r + c] * coef;
    const size_t o = (size_t)off + r + c;
    mom[o] += gv;
    const float nw = w[o] - lr * gv;
    w[o] = nw;
    if (wc) wc[o] = __float2bfloat16(nw);
    g[r + c] = 0.f;
  }
}

}  // namespace hfl

extern "C" int hfl_ms_chunk() { return hfl::MS_CHUNK; }

extern "C" void hfl_emb_sparse_norm(const int* src, int M, const float* g, int E, float* scratch, cudaStream_t stream) {
  if (M <= 0) return;
  hfl::launch_k(hfl::emb_sparse_norm_kernel, dim3(M), dim3(128), 0, stream, src, g, E, scratch);
  hfl::count_launch();
}
extern "C" void hfl_emb_sparse_apply(const int* src, int M, float* g, int E, long off, float* w, void* wc, float* mom,
                                     const float* lr, float max_norm, const float* scratch, cudaStream_t stream) {
  if (M <= 0) return;
  hfl::launch_k(hfl::emb_sparse_apply_kernel, dim3(M), dim3(128), 0, stream, src, g, E, off, w, (__nv_bfloat16*)wc, mom, lr,
                max_norm, scratch);
  hfl::count_launch();
}

extern "C" void hfl_multi_sgd_clip_step(const void* segs, const void* blocks, int n_blocks, float* w, void* wc,
                                        float* mom, const float* lr, float mu, float wd, float max_norm,
                                        float* scratch, cudaStream_t stream) {
  if (n_blocks <= 0) return;
  hfl::launch_k(hfl::multi_norm_sq_kernel, dim3(n_blocks), dim3(256), 0, stream, (const hfl::Seg*)segs, (const int2*)blocks, scratch);
  hfl::launch_k(hfl::multi_sgd_update_kernel, dim3(n_blocks), dim3(256), 0, stream, (const hfl::Seg*)segs, (const int2*)blocks, w,
                                                            (__nv_bfloat16*)wc, mom, lr, mu, wd, max_norm, scratch);
  hfl::count_launch(2);
}
