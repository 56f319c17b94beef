// Fused (residual-add +) BatchNorm(batch statistics) + ReLU, forward and backward, one launch each (sm_100a).
//
// Reference per norm layer (src/models/resnet.py:44-50): Scaler div, batch_norm (stats + apply), relu, and the
// residual `out += shortcut` of the previous block are separate kernels in forward, and relu_backward,
// batch_norm_backward (2 reductions + apply), add (gradient join) again in backward.  Here:
//
//   bn_act_fwd : s = x (+ res);  mean/var over the batch;  y = relu((s-mean)*invstd*gamma + beta)
//   bn_act_bwd : g = (dy (+dy2)) * [y>0];  dbeta = sum g; dgamma = sum g*xhat;
//                ds = gamma*invstd*(g - dbeta/n - xhat*dgamma/n) (+ dres)
//
// One thread-block CLUSTER (<= 8 CTAs, DSMEM all-reduce of the per-channel partials, cluster.sync between the
// reduction pass and the apply pass) handles one statistics group, so statistics + apply is a single launch
// instead of 2-4; the second pass re-reads the tile from L2.  `groups` > 1 evaluates ghost batches (the sBN
// statistics pass: G train-mode batches per launch).  The Scaler is folded into `eps` by the caller.
// Activations are NHWC bf16, viewed as [rows, C]; C is a power of two (4..4096).
#include <cooperative_groups.h>
#include <cstdlib>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace hfl {

constexpr int NT = 512;  // threads per CTA

struct Vec8 { uint4 u; };
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ float ldp(const __nv_bfloat16* p, int i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ float ldp(const float* p, int i) { return p[i]; }
__device__ __forceinline__ void stp(__nv_bfloat16* p, int i, float v) { p[i] = __float2bfloat16(v); }
__device__ __forceinline__ void stp(float* p, int i, float v) { p[i] = v; }

// Reduce per-thread channel partials a[8], b[8] into smem sa[C], sb[C] (no atomics: warp shuffles, then one smem
// slab per warp, then a 16-way column sum).  Thread t owns channels (8*t+e) % C.  `part` holds 2 * NW * min(C,256) floats.
constexpr int NW = NT / 32;
__device__ __forceinline__ void cta_channel_reduce(float (&a)[8], float (&b)[8], int C, float* sa, float* sb,
                                                   float* part) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (C < 8) {  // C == 4: slots e and e+4 are the same channel
#pragma unroll
    for (int e = 0; e < 4; ++e) { a[e] += a[e + 4]; b[e] += b[e + 4]; }
  }
  const int cols = C >= 8 ? C / 8 : 1;  // lanes l, l+cols, ... own the same channels
  if (cols < 32) {
    for (int o = 16; o >= cols; o >>= 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a[e] += __shfl_xor_sync(0xffffffffu, a[e], o);
        b[e] += __shfl_xor_sync(0xffffffffu, b[e], o);
      }
    }
  }
  const int CW = C < 256 ? C : 256;          // channels covered by one warp
  float* pa = part;
  float* pb = part + NW * CW;
  if (cols >= 32 || lane < cols) {
    const int ne = C < 8 ? C : 8;
    const int base = (8 * lane) % CW;
    for (int e = 0; e < ne; ++e) {
      pa[w * CW + base + e] = a[e];
      pb[w * CW + base + e] = b[e];
    }
  }
  __syncthreads();
  const int nrep = C > 256 ? C / 256 : 1;    // warps w, w+nrep, ... cover the same 256-channel window
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int win = c / CW, cl = c - win * CW;
    float x = 0.f, y = 0.f;
    for (int ww = win % nrep; ww < NW; ww += nrep) { x += pa[ww * CW + cl]; y += pb[ww * CW + cl]; }
    sa[c] = x;
    sb[c] = y;
  }
}

// Sum the CTA partials of every CTA in the cluster (DSMEM) into local totals ta[C], tb[C].  All remote loads of a
// thread are issued before the first add (one DSMEM latency instead of 2K of them).
__device__ __forceinline__ void cluster_channel_reduce(cg::cluster_group& cluster, int C, float* sa, float* sb,
                                                       float* ta, float* tb) {
  const unsigned K = cluster.num_blocks();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float vx[16], vy[16];
#pragma unroll
    for (unsigned r = 0; r < 16; ++r) {
      vx[r] = r < K ? cluster.map_shared_rank(sa, r)[c] : 0.f;
      vy[r] = r < K ? cluster.map_shared_rank(sb, r)[c] : 0.f;
    }
    float x = 0.f, y = 0.f;
#pragma unroll
    for (unsigned r = 0; r < 16; ++r) { x += vx[r]; y += vy[r]; }
    ta[c] = x;
    tb[c] = y;
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward
// grid = (K, G), cluster (K,1,1).  Group g covers rows [g*Mg, (g+1)*Mg); CTA k of the cluster covers a slice.
// smem: 6*C floats
// ---------------------------------------------------------------------------------------------------------
template <typename PT, int NV>
__global__ void __launch_bounds__(NT)
bn_act_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                  __nv_bfloat16* __restrict__ s_out, __nv_bfloat16* __restrict__ y, const PT* __restrict__ gamma,
                  const PT* __restrict__ beta, float eps, int Mg, int C, int relu, float* __restrict__ mean_out,
                  float* __restrict__ invstd_out, float* __restrict__ varu_out, const int* __restrict__ nvalid,
                  int rps) {
  pdl_sync();
  // nvalid != nullptr (training step with a partial last batch inside a captured graph): only the first
  // *nvalid samples (rps rows each) are real; the statistics ignore the padding rows and s / y are written as zeros
  // there, so the padding never reaches a weight gradient.
  // NV > 0: the CTA's slice (NV vectors per thread) stays in registers between the statistics pass and the apply
  //         pass -- one global read, one write, all loads in flight at once.
  // NV == 0: looping fallback for tensors larger than 16 CTAs x 512 threads x 10 vectors (second pass re-reads L2).
  extern __shared__ float smem[];
  float *sa = smem, *sb = smem + C, *ta = smem + 2 * C, *tb = smem + 3 * C, *sc_a = smem + 4 * C, *sc_b = smem + 5 * C;
  cg::cluster_group cluster = cg::this_cluster();
  const int K = cluster.num_blocks(), k = cluster.block_rank(), g = blockIdx.y;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) smem[c] = 0.f;
  __syncthreads();
  const long gvec0 = (long)g * Mg * C / 8;
  const int nvec = Mg * C / 8;
  const int per = NV > 0 ? NV * NT : ((nvec + K - 1) / K + NT - 1) / NT * NT;   // multiple of NT
  const int v0 = min(k * per, nvec), v1 = min(v0 + per, nvec);
  const int Mv = nvalid ? min(*nvalid * rps, Mg) : Mg;
  const int vv = min(v1, Mv * C / 8);          // vectors below vv hold real rows
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const uint4* xv = reinterpret_cast<const uint4*>(x) + gvec0;
  const uint4* rv = res ? reinterpret_cast<const uint4*>(res) + gvec0 : nullptr;
  uint4* sv = s_out ? reinterpret_cast<uint4*>(s_out) + gvec0 : nullptr;
  uint4* yv = reinterpret_cast<uint4*>(y) + gvec0;
  constexpr int NR = NV > 0 ? NV : 1;
  uint4 keep[NR];
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
  if (NV > 0) {
    uint4 rr[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int v = v0 + j * NT + threadIdx.x;
      keep[j] = v < vv ? xv[v] : zero4;
      if (rv) rr[j] = v < vv ? rv[v] : zero4;
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int v = v0 + j * NT + threadIdx.x;
      float f[8];
      unpack8(keep[j], f);
      if (rv) {
        float r[8];
        unpack8(rr[j], r);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += r[e];
        keep[j] = pack8(f);
        if (v < v1) sv[v] = keep[j];
        unpack8(keep[j], f);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { a[e] += f[e]; b[e] += f[e] * f[e]; }
    }
  } else {
    for (int v = v0 + threadIdx.x; v < v1; v += NT) {
      float f[8];
      if (v >= vv) {
        if (rv) sv[v] = zero4;
        continue;
      }
      unpack8(xv[v], f);
      if (rv) {
        float r[8];
        unpack8(rv[v], r);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += r[e];
        const uint4 packed = pack8(f);
        sv[v] = packed;
        unpack8(packed, f);   // statistics of the bf16-rounded sum (what a separate add kernel would have produced)
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { a[e] += f[e]; b[e] += f[e] * f[e]; }
    }
  }
  cta_channel_reduce(a, b, C, sa, sb, smem + 6 * C);
  __syncthreads();
  cluster.sync();
  cluster_channel_reduce(cluster, C, sa, sb, ta, tb);
  __syncthreads();
  const float inv_n = 1.f / (float)max(Mv, 1);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = ta[c] * inv_n;
    const float var = fmaxf(tb[c] * inv_n - mean * mean, 0.f);
    const float invstd = rsqrtf(var + eps);
    const float ga = ldp(gamma, c), be = ldp(beta, c);
    sc_a[c] = invstd * ga;
    sc_b[c] = be - mean * invstd * ga;
    if (k == 0) {
      mean_out[(long)g * C + c] = mean;
      invstd_out[(long)g * C + c] = invstd;
      if (varu_out) varu_out[(long)g * C + c] = var * ((float)Mv / fmaxf((float)Mv - 1.f, 1.f));
    }
  }
  __syncthreads();
  float pa[8], pb[8];
  {
    const int base = (8 * threadIdx.x) % C;
#pragma unroll
    for (int e = 0; e < 8; ++e) { pa[e] = sc_a[(base + e) % C]; pb[e] = sc_b[(base + e) % C]; }
  }
  if (NV > 0) {
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int v = v0 + j * NT + threadIdx.x;
      float f[8];
      unpack8(keep[j], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        f[e] = f[e] * pa[e] + pb[e];
        if (relu) f[e] = fmaxf(f[e], 0.f);
      }
      if (v < v1) yv[v] = v < vv ? pack8(f) : zero4;
    }
  } else {
    const uint4* in2 = rv ? reinterpret_cast<const uint4*>(sv) : xv;
    for (int v = v0 + threadIdx.x; v < v1; v += NT) {
      float f[8];
      if (v >= vv) { yv[v] = zero4; continue; }
      unpack8(in2[v], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        f[e] = f[e] * pa[e] + pb[e];
        if (relu) f[e] = fmaxf(f[e], 0.f);
      }
      yv[v] = pack8(f);
    }
  }
  cluster.sync();   // keep this CTA's smem alive until every peer has read it
}

// eval / generic: y = relu((x + res) * a[c] + b[c]) with precomputed per-channel a, b
__global__ void affine_act_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                  __nv_bfloat16* __restrict__ s_out, __nv_bfloat16* __restrict__ y,
                                  const float* __restrict__ a, const float* __restrict__ b, long nvec, int C, int relu) {
  pdl_sync();
  const uint4* xv = reinterpret_cast<const uint4*>(x);
  const uint4* rv = reinterpret_cast<const uint4*>(res);
  for (long v = blockIdx.x * (long)blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(xv[v], f);
    if (res) {
      float r[8];
      unpack8(rv[v], r);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += r[e];
      const uint4 p = pack8(f);
      reinterpret_cast<uint4*>(s_out)[v] = p;
      unpack8(p, f);
    }
    const int base = (int)((8 * v) % C);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (base + e) % C;
      f[e] = f[e] * a[c] + b[c];
      if (relu) f[e] = fmaxf(f[e], 0.f);
    }
    reinterpret_cast<uint4*>(y)[v] = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward (one statistics group: the training batch).  grid = (K), cluster (K,1,1).  smem: 6*C floats
// dy may be a [N, C] tensor broadcast over the HW positions of each sample (hw > 0: global-average-pool grad).
// ---------------------------------------------------------------------------------------------------------
template <typename PT, int NV>
__global__ void __launch_bounds__(NT)
bn_act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                  const __nv_bfloat16* __restrict__ dres, const __nv_bfloat16* __restrict__ s,
                  const float* __restrict__ mean, const float* __restrict__ invstd, const PT* __restrict__ gamma,
                  const PT* __restrict__ beta, int M, int C, int hw, float dy_scale, int relu,
                  __nv_bfloat16* __restrict__ ds, PT* __restrict__ dgamma, PT* __restrict__ dbeta,
                  const int* __restrict__ nvalid, int rps) {
  pdl_sync();
  extern __shared__ float smem[];
  float *sa = smem, *sb = smem + C, *ta = smem + 2 * C, *tb = smem + 3 * C, *c_m = smem + 4 * C, *c_i = smem + 5 * C;
  cg::cluster_group cluster = cg::this_cluster();
  const int K = cluster.num_blocks(), k = cluster.block_rank();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) smem[c] = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { c_m[c] = mean[c]; c_i[c] = invstd[c]; }
  __syncthreads();
  const int nvec = M * C / 8;
  const int per = NV > 0 ? NV * NT : ((nvec + K - 1) / K + NT - 1) / NT * NT;
  const int v0 = min(k * per, nvec), v1 = min(v0 + per, nvec);
  const int Mv = nvalid ? min(*nvalid * rps, M) : M;
  const int vv = min(v1, Mv * C / 8);          // padding rows: no contribution to the sums, ds = 0
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const uint4* sv = reinterpret_cast<const uint4*>(s);
  const uint4* gv = reinterpret_cast<const uint4*>(dy);
  const uint4* g2v = reinterpret_cast<const uint4*>(dy2);
  const uint4* rv = reinterpret_cast<const uint4*>(dres);
  uint4* dv = reinterpret_cast<uint4*>(ds);
  float pm[8], pi[8], pg[8], pb[8];
  {
    const int base = (8 * threadIdx.x) % C;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (base + e) % C;
      pm[e] = c_m[c]; pi[e] = c_i[c]; pg[e] = ldp(gamma, c); pb[e] = ldp(beta, c);
    }
  }
  const int vec_per_sample = hw > 0 ? hw * C / 8 : 1;
  auto load_g_raw = [&](int v) -> uint4 {   // packed bf16 gradient w.r.t. y for vector v (before the ReLU mask)
    if (hw > 0) {
      const int n = v / vec_per_sample;
      const int base = (int)((8L * v) % C);
      float g[8];
      if (C >= 8) {
        unpack8(reinterpret_cast<const uint4*>(dy + (long)n * C + base)[0], g);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = __bfloat162float(dy[(long)n * C + (base + e) % C]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] *= dy_scale;
      return pack8(g);
    }
    uint4 u = gv[v];
    if (dy2) {
      float g[8], t[8];
      unpack8(u, g);
      unpack8(g2v[v], t);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] += t[e];
      u = pack8(g);
    }
    return u;
  };
  constexpr int NR = NV > 0 ? NV : 1;
  constexpr bool KEEP_G = NV > 0 && NV <= 4;   // wide slices keep only s in registers and re-read the gradient
  constexpr int NG = KEEP_G ? NR : 1;
  uint4 ks[NR], kg[NG];
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
  auto accumulate = [&](const uint4& xs, uint4& gs) {
    float x[8], g[8];
    unpack8(xs, x);
    unpack8(gs, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (x[e] - pm[e]) * pi[e];
      const float pre = x[e] * (pi[e] * pg[e]) + (pb[e] - pm[e] * pi[e] * pg[e]);   // same expression as forward
      if (relu && !(pre > 0.f)) g[e] = 0.f;
      a[e] += g[e];
      b[e] += g[e] * xh;
    }
    gs = pack8(g);   // masked gradient (exact: zeroing does not change the bf16 values)
  };
  if (NV > 0) {
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int v = v0 + j * NT + threadIdx.x;
      ks[j] = v < vv ? sv[v] : zero4;
    }
    if (KEEP_G) {
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int v = v0 + j * NT + threadIdx.x;
        kg[j % NG] = v < vv ? load_g_raw(v) : zero4;
      }
#pragma unroll
      for (int j = 0; j < NR; ++j) accumulate(ks[j], kg[j % NG]);
    } else {
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int v = v0 + j * NT + threadIdx.x;
        uint4 gs = v < vv ? load_g_raw(v) : zero4;
        accumulate(ks[j], gs);
      }
    }
  } else {
    for (int v = v0 + threadIdx.x; v < vv; v += NT) {
      uint4 gs = load_g_raw(v);
      accumulate(sv[v], gs);
    }
  }
  cta_channel_reduce(a, b, C, sa, sb, smem + 6 * C);
  __syncthreads();
  cluster.sync();
  cluster_channel_reduce(cluster, C, sa, sb, ta, tb);
  __syncthreads();
  if (k == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) { stp(dbeta, c, ta[c]); stp(dgamma, c, tb[c]); }
  }
  const float inv_n = 1.f / (float)max(Mv, 1);
  float qa[8], qb[8];
  {
    const int base = (8 * threadIdx.x) % C;
#pragma unroll
    for (int e = 0; e < 8; ++e) { qa[e] = ta[(base + e) % C] * inv_n; qb[e] = tb[(base + e) % C] * inv_n; }
  }
  auto apply = [&](int v, const uint4& xs, const uint4& gs) {
    float x[8], g[8], o[8], r[8];
    unpack8(xs, x);
    unpack8(gs, g);
    if (dres) unpack8(rv[v], r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (x[e] - pm[e]) * pi[e];
      o[e] = pg[e] * pi[e] * (g[e] - qa[e] - xh * qb[e]);
      if (dres) o[e] += r[e];
    }
    dv[v] = pack8(o);
  };
  auto masked_g = [&](int v, const uint4& xs) -> uint4 {
    float x[8], g[8];
    unpack8(xs, x);
    unpack8(load_g_raw(v), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float pre = x[e] * (pi[e] * pg[e]) + (pb[e] - pm[e] * pi[e] * pg[e]);
      if (relu && !(pre > 0.f)) g[e] = 0.f;
    }
    return pack8(g);
  };
  if (NV > 0) {
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int v = v0 + j * NT + threadIdx.x;
      if (v < vv) {
        if (KEEP_G) apply(v, ks[j], kg[j % NG]);
        else apply(v, ks[j], masked_g(v, ks[j]));
      } else if (v < v1) {
        dv[v] = zero4;
      }
    }
  } else {
    for (int v = v0 + threadIdx.x; v < v1; v += NT) {
      if (v >= vv) { dv[v] = zero4; continue; }
      const uint4 xs = sv[v];
      apply(v, xs, masked_g(v, xs));
    }
  }
  cluster.sync();
}


// ---------------------------------------------------------------------------------------------------------
// Split path for LARGE single-group tensors (> 16 CTAs x 512 thr x 4 vectors): a wide statistics kernel (global
// atomics into ws[2][C]) followed by a wide apply kernel.  Two fully parallel launches beat one 16-CTA cluster here.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
bn_stats_wide_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                     __nv_bfloat16* __restrict__ s_out, int nvec, int C, float* __restrict__ ws,
                     const int* __restrict__ nvalid, int rps) {
  pdl_sync();
  // blockIdx.y = statistics group (ghost batch); nvec = vectors per group; ws = [G][2][C]
  const int vv = nvalid ? min(*nvalid * rps * C / 8, nvec) : nvec;
  extern __shared__ float smem[];
  float *sa = smem, *sb = smem + C;
  const long goff = (long)blockIdx.y * nvec;
  ws += (long)blockIdx.y * 2 * C;
  const uint4* xv = reinterpret_cast<const uint4*>(x) + goff;
  const uint4* rv = res ? reinterpret_cast<const uint4*>(res) + goff : nullptr;
  uint4* sv = s_out ? reinterpret_cast<uint4*>(s_out) + goff : nullptr;
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
  for (int v = blockIdx.x * NT + threadIdx.x; v < nvec; v += gridDim.x * NT) {
    float f[8];
    if (v >= vv) {
      if (rv) sv[v] = make_uint4(0, 0, 0, 0);
      continue;
    }
    unpack8(xv[v], f);
    if (rv) {
      float r[8];
      unpack8(rv[v], r);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += r[e];
      const uint4 p = pack8(f);
      sv[v] = p;
      unpack8(p, f);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { a[e] += f[e]; b[e] += f[e] * f[e]; }
  }
  cta_channel_reduce(a, b, C, sa, sb, smem + 2 * C);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) { atomicAdd(&ws[c], sa[c]); atomicAdd(&ws[C + c], sb[c]); }
}

// Persistent scratch protocol: ws = [2C sums][1 ticket]; every CTA of the consumer kernel takes a ticket after it has
// copied the sums it needs; the last one zeroes sums + ticket, so the next producer launch starts from zero without a
// memset node in the captured graph.
__device__ __forceinline__ void ws_release(float* ws, int C, unsigned nblocks) {
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned* ticket = reinterpret_cast<unsigned*>(ws + 2 * C);
    const unsigned t = atomicAdd(ticket, 1u);
    if (t == nblocks - 1) {
      for (int c = 0; c < 2 * C; ++c) ws[c] = 0.f;
      *ticket = 0u;
      __threadfence();
    }
  }
}

template <typename PT>
__global__ void __launch_bounds__(NT)
bn_apply_wide_kernel(const __nv_bfloat16* __restrict__ s, __nv_bfloat16* __restrict__ y, const PT* __restrict__ gamma,
                     const PT* __restrict__ beta, float eps, int M, int C, int relu, float* __restrict__ ws,
                     float* __restrict__ mean_out, float* __restrict__ invstd_out, float* __restrict__ varu_out,
                     int self_clean, const int* __restrict__ nvalid, int rps) {
  pdl_sync();
  extern __shared__ float smem[];
  float *sc_a = smem, *sc_b = smem + C;
  const int Mv = nvalid ? max(min(*nvalid * rps, M), 1) : M;
  const float inv_n = 1.f / (float)Mv;
  {   // blockIdx.y = statistics group
    const long g = blockIdx.y;
    ws += g * 2 * C;
    mean_out += g * C;
    invstd_out += g * C;
    if (varu_out) varu_out += g * C;
    s += g * (long)M * C;
    y += g * (long)M * C;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = ws[c] * inv_n;
    const float var = fmaxf(ws[C + c] * inv_n - mean * mean, 0.f);
    const float invstd = rsqrtf(var + eps);
    const float ga = ldp(gamma, c), be = ldp(beta, c);
    sc_a[c] = invstd * ga;
    sc_b[c] = be - mean * invstd * ga;
    if (blockIdx.x == 0) {
      mean_out[c] = mean;
      invstd_out[c] = invstd;
      if (varu_out) varu_out[c] = var * ((float)Mv / fmaxf((float)Mv - 1.f, 1.f));
    }
  }
  __syncthreads();
  if (self_clean) ws_release(ws, C, gridDim.x);   // last CTA to have read the sums re-zeroes them for the next launch
  float pa[8], pb[8];
  {
    const int base = (8 * threadIdx.x) % C;
#pragma unroll
    for (int e = 0; e < 8; ++e) { pa[e] = sc_a[(base + e) % C]; pb[e] = sc_b[(base + e) % C]; }
  }
  const int nvec = M * C / 8;
  const int vv = nvalid ? min(*nvalid * rps * C / 8, nvec) : nvec;
  const uint4* sv = reinterpret_cast<const uint4*>(s);
  uint4* yv = reinterpret_cast<uint4*>(y);
  for (int v = blockIdx.x * NT + threadIdx.x; v < nvec; v += gridDim.x * NT) {
    float f[8];
    if (v >= vv) { yv[v] = make_uint4(0, 0, 0, 0); continue; }
    unpack8(sv[v], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      f[e] = f[e] * pa[e] + pb[e];
      if (relu) f[e] = fmaxf(f[e], 0.f);
    }
    yv[v] = pack8(f);
  }
}

template <typename PT>
__global__ void __launch_bounds__(NT)
bn_bwd_reduce_wide_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                          const __nv_bfloat16* __restrict__ s, const float* __restrict__ mean,
                          const float* __restrict__ invstd, const PT* __restrict__ gamma, const PT* __restrict__ beta,
                          int nvec_all, int C, int relu, float* __restrict__ ws, const int* __restrict__ nvalid,
                          int rps) {
  pdl_sync();
  extern __shared__ float smem[];
  float *sa = smem, *sb = smem + C;
  const int nvec = nvalid ? min(*nvalid * rps * C / 8, nvec_all) : nvec_all;   // padding rows carry no gradient
  float pm[8], pi[8], pg[8], pb[8];
  {
    const int base = (8 * threadIdx.x) % C;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (base + e) % C;
      pm[e] = mean[c]; pi[e] = invstd[c]; pg[e] = ldp(gamma, c); pb[e] = ldp(beta, c);
    }
  }
  const uint4* sv = reinterpret_cast<const uint4*>(s);
  const uint4* gv = reinterpret_cast<const uint4*>(dy);
  const uint4* g2v = reinterpret_cast<const uint4*>(dy2);
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
  for (int v = blockIdx.x * NT + threadIdx.x; v < nvec; v += gridDim.x * NT) {
    float x[8], g[8];
    unpack8(sv[v], x);
    unpack8(gv[v], g);
    if (dy2) {
      float t[8];
      unpack8(g2v[v], t);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] += t[e];
      unpack8(pack8(g), g);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (x[e] - pm[e]) * pi[e];
      const float pre = x[e] * (pi[e] * pg[e]) + (pb[e] - pm[e] * pi[e] * pg[e]);
      const float ge = (!relu || pre > 0.f) ? g[e] : 0.f;
      a[e] += ge;
      b[e] += ge * xh;
    }
  }
  cta_channel_reduce(a, b, C, sa, sb, smem + 2 * C);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) { atomicAdd(&ws[c], sa[c]); atomicAdd(&ws[C + c], sb[c]); }
}

template <typename PT>
__global__ void __launch_bounds__(NT)
bn_bwd_apply_wide_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                         const __nv_bfloat16* __restrict__ dres, const __nv_bfloat16* __restrict__ s,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const PT* __restrict__ gamma, const PT* __restrict__ beta, int M, int C, int relu,
                         float* __restrict__ ws, __nv_bfloat16* __restrict__ ds, PT* __restrict__ dgamma,
                         PT* __restrict__ dbeta, int self_clean, const int* __restrict__ nvalid, int rps) {
  pdl_sync();
  float pm[8], pi[8], pg[8], pb[8], qa[8], qb[8];
  const int Mv = nvalid ? max(min(*nvalid * rps, M), 1) : M;
  const float inv_n = 1.f / (float)Mv;
  {
    const int base = (8 * threadIdx.x) % C;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (base + e) % C;
      pm[e] = mean[c]; pi[e] = invstd[c]; pg[e] = ldp(gamma, c); pb[e] = ldp(beta, c);
      qa[e] = ws[c] * inv_n; qb[e] = ws[C + c] * inv_n;
    }
  }
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) { stp(dbeta, c, ws[c]); stp(dgamma, c, ws[C + c]); }
  __syncthreads();
  if (self_clean) ws_release(ws, C, gridDim.x);
  const int nvec = M * C / 8;
  const uint4* sv = reinterpret_cast<const uint4*>(s);
  const uint4* gv = reinterpret_cast<const uint4*>(dy);
  const uint4* g2v = reinterpret_cast<const uint4*>(dy2);
  const uint4* rv = reinterpret_cast<const uint4*>(dres);
  uint4* dv = reinterpret_cast<uint4*>(ds);
  const int vv = nvalid ? min(*nvalid * rps * C / 8, nvec) : nvec;
  for (int v = blockIdx.x * NT + threadIdx.x; v < nvec; v += gridDim.x * NT) {
    float x[8], g[8], o[8], r[8];
    if (v >= vv) { dv[v] = make_uint4(0, 0, 0, 0); continue; }
    unpack8(sv[v], x);
    unpack8(gv[v], g);
    if (dy2) {
      float t[8];
      unpack8(g2v[v], t);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] += t[e];
      unpack8(pack8(g), g);
    }
    if (dres) unpack8(rv[v], r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (x[e] - pm[e]) * pi[e];
      const float pre = x[e] * (pi[e] * pg[e]) + (pb[e] - pm[e] * pi[e] * pg[e]);
      const float ge = (!relu || pre > 0.f) ? g[e] : 0.f;
      o[e] = pg[e] * pi[e] * (ge - qa[e] - xh * qb[e]);
      if (dres) o[e] += r[e];
    }
    dv[v] = pack8(o);
  }
}

}  // namespace hfl

using namespace hfl;

// Choose (cluster size K, vectors per thread NV) so that K * NT * NV vectors cover one statistics group.
// NV == 0 selects the looping kernel (very large groups only).
static int bn_shape_policy() {
  static const int p = []() { const char* e = getenv("HFL_BN_SHAPE"); return e ? atoi(e) : 0; }();
  return p;
}
static void pick_shape(long elems, int& K, int& NV) {
  const long nvec = elems / 8;
  const int nvs[4] = {1, 2, 4, 10};   // NV=10 only reachable for multi-group (ghost batch) launches
  if (bn_shape_policy() == 1) {
    // fewest CTAs first: tiny tensors are latency-bound, a 1-2 CTA launch skips most of the cluster scheduling and
    // DSMEM round trips of an 8-CTA cluster
    for (int ki = 0; ki < 5; ++ki) {
      const int kk = 1 << ki;
      for (int ni = 0; ni < 4; ++ni)
        if ((long)kk * NT * nvs[ni] >= nvec) { K = kk; NV = nvs[ni]; return; }
    }
    K = 8;
    NV = 0;
    return;
  }
  for (int ki = 0; ki < 5; ++ki) {
    const int kk = 1 << ki;                       // 1, 2, 4, 8, 16
    for (int ni = 0; ni < 4; ++ni) {
      // prefer spreading over more CTAs before holding more vectors per thread
      if (nvs[ni] > 2 && kk < 16) continue;
      if ((long)kk * NT * nvs[ni] >= nvec) { K = kk; NV = nvs[ni]; return; }
    }
  }
  K = 8;
  NV = 0;
}

template <typename Kern, typename... Args>
static cudaError_t launch_cluster(Kern kern, dim3 grid, int cluster_x, size_t smem, cudaStream_t stream, Args... args) {
  if (cluster_x > 8) {
    cudaError_t e0 = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e0 != cudaSuccess) return e0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

#define HFL_NV_DISPATCH(NVVAL, CALL) \
  switch (NVVAL) {                   \
    case 1: { constexpr int NVC = 1; CALL; break; }   \
    case 2: { constexpr int NVC = 2; CALL; break; }   \
    case 4: { constexpr int NVC = 4; CALL; break; }   \
    case 10: { constexpr int NVC = 10; CALL; break; } \
    default: { constexpr int NVC = 0; CALL; break; }  \
  }

// per-group elements above which the split (2-launch) path wins over the single cluster kernel (measured on one GPU with
// 10-16 concurrent clients); HFL_BN_WIDE_ELEMS overrides for A/B (the cluster kernel holds at most 16 x 512 x 10 vectors)
static const long WIDE_ELEMS = []() { const char* e = getenv("HFL_BN_WIDE_ELEMS"); return e ? atol(e) : 48L * 1024; }();
static size_t wide_smem(int C) { return (2 * (size_t)C + 2 * 16 * (size_t)(C < 256 ? C : 256)) * sizeof(float); }
static int wide_grid(long nvec) {
  long b = (nvec + NT * 2 - 1) / (NT * 2);
  if (b > 148 * 2) b = 148 * 2;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int hfl_bn_act_fwd(const void* x, const void* res, void* s_out, void* y, const void* gamma,
                              const void* beta, int param_bf16, float eps, int groups, int Mg, int C, int relu,
                              float* mean_out, float* invstd_out, float* varu_out, float* ws, int ws_persistent,
                              const int* nvalid, int rps, cudaStream_t stream) {
  if (groups > 1) nvalid = nullptr;               // ghost batches are always full
  // ghost batches (sBN pass): the wide path by default; HFL_BN_GHOST_CLUSTER=1 keeps every ghost batch in the registers of
  // one cluster instead (one read of x instead of two) when it fits 16 CTAs x 512 threads x 10 vectors
  static const bool ghost_cluster = []() { const char* e = getenv("HFL_BN_GHOST_CLUSTER"); return e && e[0] == '1'; }();
  const bool ghost_wide = groups > 1 && !(ghost_cluster && (long)Mg * C <= 16L * NT * 10 * 8);
  if (((long)Mg * C > WIDE_ELEMS || ghost_wide) && ws != nullptr && !(groups > 1 && !ghost_wide)) {
    const int nvec = Mg * C / 8;
    int gx = wide_grid(nvec);
    if (groups > 1 && gx > 16) gx = 16;           // many groups already fill the machine
    dim3 grid(gx, groups);
    const int self_clean = (ws_persistent && groups == 1) ? 1 : 0;
    if (!self_clean) cudaMemsetAsync(ws, 0, 2 * (size_t)C * groups * sizeof(float), stream);
    launch_k(bn_stats_wide_kernel, dim3(grid), dim3(NT), wide_smem(C), stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)res,
                                                             (__nv_bfloat16*)s_out, nvec, C, ws, nvalid, rps);
    const void* sin = res ? s_out : x;
    if (param_bf16)
      launch_k(bn_apply_wide_kernel<__nv_bfloat16>, dim3(grid), dim3(NT), 2 * C * sizeof(float), stream, 
          (const __nv_bfloat16*)sin, (__nv_bfloat16*)y, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, eps, Mg, C,
          relu, ws, mean_out, invstd_out, varu_out, self_clean, nvalid, rps);
    else
      launch_k(bn_apply_wide_kernel<float>, dim3(grid), dim3(NT), 2 * C * sizeof(float), stream, 
          (const __nv_bfloat16*)sin, (__nv_bfloat16*)y, (const float*)gamma, (const float*)beta, eps, Mg, C, relu, ws,
          mean_out, invstd_out, varu_out, self_clean, nvalid, rps);
    count_launch(2);
    return (int)cudaGetLastError();
  }
  int K, NV;
  pick_shape((long)Mg * C, K, NV);
  const size_t smem = (6 * (size_t)C + 2 * 16 * (size_t)(C < 256 ? C : 256)) * sizeof(float);
  dim3 grid(K, groups);
  cudaError_t e = cudaSuccess;
  if (param_bf16) {
    HFL_NV_DISPATCH(NV, (e = launch_cluster(bn_act_fwd_kernel<__nv_bfloat16, NVC>, grid, K, smem, stream,
                                            (const __nv_bfloat16*)x, (const __nv_bfloat16*)res, (__nv_bfloat16*)s_out,
                                            (__nv_bfloat16*)y, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta,
                                            eps, Mg, C, relu, mean_out, invstd_out, varu_out, nvalid, rps)))
  } else {
    HFL_NV_DISPATCH(NV, (e = launch_cluster(bn_act_fwd_kernel<float, NVC>, grid, K, smem, stream,
                                            (const __nv_bfloat16*)x, (const __nv_bfloat16*)res, (__nv_bfloat16*)s_out,
                                            (__nv_bfloat16*)y, (const float*)gamma, (const float*)beta, eps, Mg, C, relu,
                                            mean_out, invstd_out, varu_out, nvalid, rps)))
  }
  count_launch();
  return (int)e;
}

// Apply-only half of the split path for producers that already left (sum, sum of squares) in `ws` = [2C sums][ticket]:
// the tcgen05 convolution's statistics epilogue (conv_tcgen05.cu).  The kernel re-zeroes `ws` itself (ticket protocol).
extern "C" int hfl_bn_apply_sums(const void* s, void* y, const void* gamma, const void* beta, int param_bf16, float eps,
                                 int M, int C, int relu, float* ws, float* mean_out, float* invstd_out, float* varu_out,
                                 const int* nvalid, int rps, cudaStream_t stream) {
  const int nvec = M * C / 8;
  dim3 grid(wide_grid(nvec), 1);
  if (param_bf16)
    launch_k(bn_apply_wide_kernel<__nv_bfloat16>, grid, dim3(NT), 2 * C * sizeof(float), stream,
             (const __nv_bfloat16*)s, (__nv_bfloat16*)y, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, eps, M, C,
             relu, ws, mean_out, invstd_out, varu_out, 1, nvalid, rps);
  else
    launch_k(bn_apply_wide_kernel<float>, grid, dim3(NT), 2 * C * sizeof(float), stream, (const __nv_bfloat16*)s,
             (__nv_bfloat16*)y, (const float*)gamma, (const float*)beta, eps, M, C, relu, ws, mean_out, invstd_out,
             varu_out, 1, nvalid, rps);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" void hfl_affine_act(const void* x, const void* res, void* s_out, void* y, const float* a, const float* b,
                               long n, int C, int relu, cudaStream_t stream) {
  const long nvec = n / 8;
  long blocks = (nvec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  launch_k(affine_act_kernel, dim3((int)blocks), dim3(256), 0, stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)res,
                                                     (__nv_bfloat16*)s_out, (__nv_bfloat16*)y, a, b, nvec, C, relu);
  count_launch();
}

extern "C" int hfl_bn_act_bwd(const void* dy, const void* dy2, const void* dres, const void* s, const float* mean,
                              const float* invstd, const void* gamma, const void* beta, int param_bf16, int M, int C,
                              int hw, float dy_scale, int relu, void* ds, void* dgamma, void* dbeta, float* ws,
                              int ws_persistent, const int* nvalid, int rps, cudaStream_t stream) {
  if (hw == 0 && (long)M * C > WIDE_ELEMS && ws != nullptr) {
    const int nvec = M * C / 8;
    const int grid = wide_grid(nvec);
    const int self_clean = ws_persistent ? 1 : 0;
    if (!self_clean) cudaMemsetAsync(ws, 0, 2 * (size_t)C * sizeof(float), stream);
    if (param_bf16) {
      launch_k(bn_bwd_reduce_wide_kernel<__nv_bfloat16>, dim3(grid), dim3(NT), wide_smem(C), stream, 
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2, (const __nv_bfloat16*)s, mean, invstd,
          (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, nvec, C, relu, ws, nvalid, rps);
      launch_k(bn_bwd_apply_wide_kernel<__nv_bfloat16>, dim3(grid), dim3(NT), 0, stream, 
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2, (const __nv_bfloat16*)dres, (const __nv_bfloat16*)s, mean,
          invstd, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, M, C, relu, ws, (__nv_bfloat16*)ds,
          (__nv_bfloat16*)dgamma, (__nv_bfloat16*)dbeta, self_clean, nvalid, rps);
    } else {
      launch_k(bn_bwd_reduce_wide_kernel<float>, dim3(grid), dim3(NT), wide_smem(C), stream, 
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2, (const __nv_bfloat16*)s, mean, invstd, (const float*)gamma,
          (const float*)beta, nvec, C, relu, ws, nvalid, rps);
      launch_k(bn_bwd_apply_wide_kernel<float>, dim3(grid), dim3(NT), 0, stream, 
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2, (const __nv_bfloat16*)dres, (const __nv_bfloat16*)s, mean,
          invstd, (const float*)gamma, (const float*)beta, M, C, relu, ws, (__nv_bfloat16*)ds, (float*)dgamma,
          (float*)dbeta, self_clean, nvalid, rps);
    }
    count_launch(2);
    return (int)cudaGetLastError();
  }
  int K, NV;
  pick_shape((long)M * C, K, NV);
  const size_t smem = (6 * (size_t)C + 2 * 16 * (size_t)(C < 256 ? C : 256)) * sizeof(float);
  dim3 grid(K);
  cudaError_t e = cudaSuccess;
  if (param_bf16) {
    HFL_NV_DISPATCH(NV, (e = launch_cluster(bn_act_bwd_kernel<__nv_bfloat16, NVC>, grid, K, smem, stream,
                                            (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2,
                                            (const __nv_bfloat16*)dres, (const __nv_bfloat16*)s, mean, invstd,
                                            (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, M, C, hw, dy_scale,
                                            relu, (__nv_bfloat16*)ds, (__nv_bfloat16*)dgamma, (__nv_bfloat16*)dbeta,
                                            nvalid, rps)))
  } else {
    HFL_NV_DISPATCH(NV, (e = launch_cluster(bn_act_bwd_kernel<float, NVC>, grid, K, smem, stream,
                                            (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2,
                                            (const __nv_bfloat16*)dres, (const __nv_bfloat16*)s, mean, invstd,
                                            (const float*)gamma, (const float*)beta, M, C, hw, dy_scale, relu,
                                            (__nv_bfloat16*)ds, (float*)dgamma, (float*)dbeta, nvalid, rps)))
  }
  count_launch();
  return (int)e;
}
