// Fused HeteroFL communication kernels (sm_100a).
//
//   distribute : W_g  --slice-->  every active client's sub-network arena   (reference src/fed.py:161-178)
//   aggregate  : heterogeneous FedAvg of all client arenas into W_g          (reference src/fed.py:180-297)
//
// Both run as ONE launch over all tensors and all clients.  Slices are prefixes / per-head prefixes, so the
// local<->global coordinate map of a tensor is 8 integers per (tensor, level) (plan/layout.py) instead of
// materialised index tensors, and the coverage count is computed on the fly instead of a second `count` tensor.
// Client arenas and output replicas are raw pointers: when they are peer-mapped (parallel/peer.py) the same kernels
// pull client values over NVLink (P2P loads) and publish the new W_g shard into every rank's replica (P2P stores),
// i.e. gather-reduce + all-gather fused into the aggregation pass.
#include "common.cuh"

namespace hfl {

constexpr int FED_THREADS = 256;
constexpr int FED_CHUNK = 2048;     // elements per block
constexpr int MAX_CLIENTS_SMEM = 256;

struct TensorDesc { int g_off, R, C, inner, rH, cH, label_rows, col_inner; };
struct LevelDesc  { int l_off, r_loc, c_loc, r_dl, r_dg, c_dl, c_dg, pad; };

// element e of a tensor (physical order [R][C][inner] or [R][inner][C]) -> (row, col, t)
__device__ __forceinline__ void decode(const TensorDesc& d, int C, int e, int& row, int& col, int& t) {
  const int per_row = C * d.inner;
  row = e / per_row;
  const int rem = e - row * per_row;
  if (d.col_inner) { t = rem / C; col = rem - t * C; }
  else             { col = rem / d.inner; t = rem - col * d.inner; }
}

// global axis index -> local axis index under a (per-head) prefix slice; returns -1 when not covered
__device__ __forceinline__ int to_local(int g, int heads, int dl, int dg) {
  if (heads == 1) return g < dl ? g : -1;
  const int h = g / dg, t = g - h * dg;
  return t < dl ? h * dl + t : -1;
}
// local axis index -> global axis index
__device__ __forceinline__ int to_global(int l, int heads, int dl, int dg) {
  if (heads == 1) return l;
  const int h = l / dl, t = l - h * dl;
  return h * dg + t;
}

// ---------------------------------------------------------------------------------------------------------
// aggregate
// ---------------------------------------------------------------------------------------------------------
struct ClientSm { const float* base; const unsigned char* mask; int l_off, r_dl, r_dg, c_dl, c_dg, c_loc; };

__global__ void __launch_bounds__(FED_THREADS)
aggregate_kernel(const TensorDesc* __restrict__ tensors, const LevelDesc* __restrict__ levels, int T,
                 const int2* __restrict__ block_tab, int block_lo,
                 const unsigned long long* __restrict__ client_ptr, const int* __restrict__ client_level,
                 const unsigned char* __restrict__ client_mask, int mask_stride, int M,
                 const unsigned long long* __restrict__ out_ptr, int n_out) {
  pdl_sync();
  __shared__ ClientSm cs[MAX_CLIENTS_SMEM];
  const int2 bt = block_tab[block_lo + blockIdx.x];
  const TensorDesc d = tensors[bt.x];
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const LevelDesc l = levels[client_level[m] * T + bt.x];
    ClientSm c;
    c.base = reinterpret_cast<const float*>(client_ptr[m]);
    c.mask = d.label_rows ? client_mask + (size_t)m * mask_stride : nullptr;
    c.l_off = l.l_off; c.r_dl = l.r_dl; c.r_dg = l.r_dg; c.c_dl = l.c_dl; c.c_dg = l.c_dg; c.c_loc = l.c_loc;
    cs[m] = c;
  }
  __syncthreads();
  const int numel = d.R * d.C * d.inner;
  const int end = min(bt.y + FED_CHUNK, numel);
  for (int e = bt.y + threadIdx.x; e < end; e += FED_THREADS) {
    int row, col, t;
    decode(d, d.C, e, row, col, t);
    float acc = 0.f;
    int cnt = 0;
    for (int m = 0; m < M; ++m) {
      const ClientSm& c = cs[m];
      const int lr = to_local(row, d.rH, c.r_dl, c.r_dg);
      const int lc = to_local(col, d.cH, c.c_dl, c.c_dg);
      bool in = (lr >= 0) & (lc >= 0);
      if (in && c.mask) in = c.mask[row] != 0;
      if (in) {
        const int li = d.col_inner ? (lr * d.inner + t) * c.c_loc + lc : (lr * c.c_loc + lc) * d.inner + t;
        acc += c.base[c.l_off + li];          // local or NVLink peer load
        ++cnt;
      }
    }
    if (cnt > 0) {
      const float v = acc / (float)cnt;
      for (int o = 0; o < n_out; ++o)         // local store + peer stores (fused all-gather of the new W_g)
        reinterpret_cast<float*>(out_ptr[o])[d.g_off + e] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// aggregate v2: row-wise, 16-byte accesses, clients grouped by level, no client-count limit.
//
// A thread owns 4 consecutive coordinates along a tensor's innermost axis.  Slices are (per-head) prefixes whose
// extents are multiples of 4 on every level (checked on the host: `vec4[t]`), so the 4 coordinates are covered by
// exactly the same clients and map to 4 consecutive local coordinates: the (row, col, tap) decode and the
// global -> local map run once per LEVEL (<= 8 per model) instead of once per client and element, and every client
// value arrives as one float4 (local or NVLink peer load).  Clients are sorted by level (`lvl_start`); the per-client
// state is just (pointer, mask row) read through the read-only path, so M is unbounded.  Tensors that do not qualify
// (3 input channels, odd label rows, ...) take the scalar loop below with the same tables.
// An optional small all-reduce (every rank's metric slab -> `red_dst`) rides in block 0 of the same launch, so the
// round's metric merge needs no extra kernel and no extra flag barrier.
// ---------------------------------------------------------------------------------------------------------
constexpr int FED2_CHUNK = 4096;     // elements per block (4 float4 passes of 256 threads)

__global__ void __launch_bounds__(FED_THREADS)
aggregate_v2_kernel(const TensorDesc* __restrict__ tensors, const LevelDesc* __restrict__ levels, int T,
                    const int* __restrict__ vec4, const int2* __restrict__ block_tab, int block_lo,
                    const unsigned long long* __restrict__ client_ptr, const int* __restrict__ lvl_start, int L,
                    const unsigned char* __restrict__ client_mask, int mask_stride,
                    const unsigned long long* __restrict__ out_ptr, int n_out,
                    const unsigned long long* __restrict__ red_src, int red_n_src, float* __restrict__ red_dst,
                    int red_n) {
  pdl_sync();
  if (blockIdx.x == 0 && red_n > 0) {
    for (int i = threadIdx.x; i < red_n; i += FED_THREADS) {
      float acc = 0.f;
      for (int s = 0; s < red_n_src; ++s) acc += reinterpret_cast<const float*>(red_src[s])[i];
      red_dst[i] = acc;
    }
  }
  const int2 bt = block_tab[block_lo + blockIdx.x];
  const TensorDesc d = tensors[bt.x];
  const int numel = d.R * d.C * d.inner;
  const int end = min(bt.y + FED2_CHUNK, numel);
  if (vec4[bt.x]) {
    for (int e = bt.y + 4 * threadIdx.x; e < end; e += 4 * FED_THREADS) {
      int row, col, t;
      decode(d, d.C, e, row, col, t);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      for (int lv = 0; lv < L; ++lv) {
        const int m0 = lvl_start[lv], m1 = lvl_start[lv + 1];
        if (m0 == m1) continue;
        const LevelDesc l = levels[lv * T + bt.x];
        const int lr = to_local(row, d.rH, l.r_dl, l.r_dg);
        const int lc = to_local(col, d.cH, l.c_dl, l.c_dg);
        if ((lr < 0) | (lc < 0)) continue;
        const int li = l.l_off + (d.col_inner ? (lr * d.inner + t) * l.c_loc + lc : (lr * l.c_loc + lc) * d.inner + t);
        if (d.label_rows) {
#pragma unroll 4
          for (int m = m0; m < m1; ++m) {
            if (client_mask[(size_t)m * mask_stride + row]) {
              const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(client_ptr[m]) + li);
              acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
              ++cnt;
            }
          }
        } else {
#pragma unroll 4
          for (int m = m0; m < m1; ++m) {
            const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(client_ptr[m]) + li);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          cnt += m1 - m0;
        }
      }
      if (cnt > 0) {
        const float inv = 1.f / (float)cnt;
        const float4 v = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        for (int o = 0; o < n_out; ++o)
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_ptr[o]) + d.g_off + e) = v;
      }
    }
    return;
  }
  for (int e = bt.y + threadIdx.x; e < end; e += FED_THREADS) {
    int row, col, t;
    decode(d, d.C, e, row, col, t);
    float acc = 0.f;
    int cnt = 0;
    for (int lv = 0; lv < L; ++lv) {
      const int m0 = lvl_start[lv], m1 = lvl_start[lv + 1];
      if (m0 == m1) continue;
      const LevelDesc l = levels[lv * T + bt.x];
      const int lr = to_local(row, d.rH, l.r_dl, l.r_dg);
      const int lc = to_local(col, d.cH, l.c_dl, l.c_dg);
      if ((lr < 0) | (lc < 0)) continue;
      const int li = l.l_off + (d.col_inner ? (lr * d.inner + t) * l.c_loc + lc : (lr * l.c_loc + lc) * d.inner + t);
      for (int m = m0; m < m1; ++m) {
        if (d.label_rows && !client_mask[(size_t)m * mask_stride + row]) continue;
        acc += reinterpret_cast<const float*>(client_ptr[m])[li];
        ++cnt;
      }
    }
    if (cnt > 0) {
      const float v = acc / (float)cnt;
      for (int o = 0; o < n_out; ++o) reinterpret_cast<float*>(out_ptr[o])[d.g_off + e] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// distribute: one launch prepares every client's optimiser-ready state
//   w  (fp32 master)  <- slice of W_g        wc (bf16 compute copy, optional) <- same
//   mom (fp32 momentum, optional) <- 0
// grid = (max blocks per level, M clients)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FED_THREADS)
distribute_kernel(const TensorDesc* __restrict__ tensors, const LevelDesc* __restrict__ levels, int T,
                  const int2* __restrict__ lvl_block_tab, const int* __restrict__ lvl_block_start,
                  const float* __restrict__ global_w,
                  const unsigned long long* __restrict__ w_ptr, const unsigned long long* __restrict__ wc_ptr,
                  const unsigned long long* __restrict__ mom_ptr, const int* __restrict__ client_level) {
  pdl_sync();
  const int m = blockIdx.y;
  const int lv = client_level[m];
  const int nb = lvl_block_start[lv + 1] - lvl_block_start[lv];
  if ((int)blockIdx.x >= nb) return;
  const int2 bt = lvl_block_tab[lvl_block_start[lv] + blockIdx.x];
  const TensorDesc d = tensors[bt.x];
  const LevelDesc l = levels[lv * T + bt.x];
  float* w = reinterpret_cast<float*>(w_ptr[m]);
  __nv_bfloat16* wc = wc_ptr ? reinterpret_cast<__nv_bfloat16*>(wc_ptr[m]) : nullptr;
  float* mom = mom_ptr ? reinterpret_cast<float*>(mom_ptr[m]) : nullptr;
  const int numel = l.r_loc * l.c_loc * d.inner;
  const int end = min(bt.y + FED_CHUNK, numel);
  for (int e = bt.y + threadIdx.x; e < end; e += FED_THREADS) {
    int lr, lc, t;
    decode(d, l.c_loc, e, lr, lc, t);
    const int gr = to_global(lr, d.rH, l.r_dl, l.r_dg);
    const int gc = to_global(lc, d.cH, l.c_dl, l.c_dg);
    const int gi = d.col_inner ? (gr * d.inner + t) * d.C + gc : (gr * d.C + gc) * d.inner + t;
    const float v = global_w[d.g_off + gi];
    w[l.l_off + e] = v;
    if (wc) wc[l.l_off + e] = __float2bfloat16(v);
    if (mom) mom[l.l_off + e] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// peer_sum: one-shot all-reduce(sum) of small vectors over peer pointers (sBN statistic merge, metric merge).
// Every rank reads all n_src buffers and writes the sum to its own dst.
// ---------------------------------------------------------------------------------------------------------
__global__ void peer_sum_kernel(const unsigned long long* __restrict__ src_ptr, int n_src, float* __restrict__ dst,
                                int n) {
  pdl_sync();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < n_src; ++s) acc += reinterpret_cast<const float*>(src_ptr[s])[i];
    dst[i] = acc;
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" void hfl_aggregate(const void* tensors, const void* levels, int T, const void* block_tab, int block_lo,
                              int n_blocks, const void* client_ptr, const void* client_level,
                              const void* client_mask, int mask_stride, int M, const void* out_ptr, int n_out,
                              cudaStream_t stream) {
  if (n_blocks <= 0 || M <= 0) return;
  launch_k(aggregate_kernel, dim3(n_blocks), dim3(FED_THREADS), 0, stream, 
      (const TensorDesc*)tensors, (const LevelDesc*)levels, T, (const int2*)block_tab, block_lo,
      (const unsigned long long*)client_ptr, (const int*)client_level, (const unsigned char*)client_mask,
      mask_stride, M, (const unsigned long long*)out_ptr, n_out);
  count_launch();
}

extern "C" void hfl_distribute(const void* tensors, const void* levels, int T, const void* lvl_block_tab,
                               const void* lvl_block_start, int max_blocks, const float* global_w,
                               const void* w_ptr, const void* wc_ptr, const void* mom_ptr,
                               const void* client_level, int M, cudaStream_t stream) {
  if (M <= 0 || max_blocks <= 0) return;
  dim3 grid(max_blocks, M);
  launch_k(distribute_kernel, dim3(grid), dim3(FED_THREADS), 0, stream, 
      (const TensorDesc*)tensors, (const LevelDesc*)levels, T, (const int2*)lvl_block_tab,
      (const int*)lvl_block_start, global_w, (const unsigned long long*)w_ptr,
      (const unsigned long long*)wc_ptr, (const unsigned long long*)mom_ptr, (const int*)client_level);
  count_launch();
}

extern "C" void hfl_peer_sum(const void* src_ptr, int n_src, float* dst, int n, cudaStream_t stream) {
  if (n <= 0) return;
  int blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  launch_k(peer_sum_kernel, dim3(blocks), dim3(256), 0, stream, (const unsigned long long*)src_ptr, n_src, dst, n);
  count_launch();
}

extern "C" void hfl_aggregate_v2(const void* tensors, const void* levels, int T, const int* vec4, const void* block_tab,
                                 int block_lo, int n_blocks, const void* client_ptr, const int* lvl_start, int L,
                                 const void* client_mask, int mask_stride, const void* out_ptr, int n_out,
                                 const void* red_src, int red_n_src, float* red_dst, int red_n, cudaStream_t stream) {
  if (n_blocks <= 0) return;
  launch_k(aggregate_v2_kernel, dim3(n_blocks), dim3(FED_THREADS), 0, stream, 
      (const TensorDesc*)tensors, (const LevelDesc*)levels, T, vec4, (const int2*)block_tab, block_lo,
      (const unsigned long long*)client_ptr, lvl_start, L, (const unsigned char*)client_mask, mask_stride,
      (const unsigned long long*)out_ptr, n_out, (const unsigned long long*)red_src, red_n_src, red_dst, red_n);
  count_launch();
}

extern "C" int hfl_fed2_chunk() { return FED2_CHUNK; }
extern "C" int hfl_fed_chunk() { return FED_CHUNK; }
extern "C" int hfl_fed_max_clients() { return MAX_CLIENTS_SMEM; }
