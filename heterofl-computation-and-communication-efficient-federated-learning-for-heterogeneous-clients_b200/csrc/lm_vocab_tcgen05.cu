// Fused vocabulary head of the masked-LM Transformer on tcgen05 (sm_100a): decoder.linear2 + zero-fill label mask +
// softmax cross-entropy, forward AND backward, without ever writing the [tokens, vocabulary] logits.
//
// Reference (src/models/transformer.py:125-129,154-161): out = linear2(x) -> masked_fill(label_mask == 0, 0) ->
// F.cross_entropy; autograd then materialises logits, softmax and dlogits ([64, 33 278] each) and runs two more GEMMs.
// Here the vocabulary is cut into 128-column tiles, one CTA per tile, and a step is three launches:
//
//   lm_vocab_fwd_kernel   logits tile = d . W2_tile^T in TMEM (M128 x N128, K = E); the epilogue adds the bias, applies the
//                         label mask and reduces the tile to per-row (max, sum exp, arg max) partials + the target logit
//   lm_vocab_lse_kernel   combines the partials of a row (online log-sum-exp), emits loss / accuracy / perplexity meter
//   lm_vocab_bwd_kernel   recomputes the logits tile (4 k-blocks: cheaper than reading it back), forms
//                         g = (softmax - onehot) / rows in registers, writes it as bf16 into shared memory in the
//                         128-byte-swizzled UMMA layout and feeds it straight back to the tensor core twice:
//                           dW2_tile[128 voc, E] = g^T . d       (g read MN-major, d read MN-major: same bytes TMA loaded)
//                           dX     [tok, E]    += g . W2_tile    (g read K-major, W2 tile read MN-major: same bytes)
//                         plus the bias gradient (column sums of g).  dX partials of the 260 tiles are combined with
//                         16-byte fp32 reductions into a [tokens, E] buffer the next kernel reads.
//
// TMEM: logits accumulator 128 columns; after the first epilogue its columns are dead and dW2 reuses them:
//   [0, E) dW2 tile, [256, 256 + E) dX partial  (E <= 256 -> 512 columns).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace hfl {

constexpr int VH_THREADS = 256;
constexpr int VH_BV = 128;                      // vocabulary columns per tile
constexpr int VH_KB_BYTES = 128 * 64 * 2;       // one [128 rows x 64 cols] bf16 k-block: 16 KB
constexpr int VH_MAX_KB = 4;                    // E <= 256
constexpr int VH_OFF_D = 0;
constexpr int VH_OFF_W = VH_MAX_KB * VH_KB_BYTES;             // 64 KB
constexpr int VH_OFF_G = 2 * VH_MAX_KB * VH_KB_BYTES;         // 128 KB: g tile, two [128 tok x 64 voc] sub-tiles
constexpr int VH_OFF_MISC = VH_OFF_G + 2 * VH_KB_BYTES;       // 160 KB
constexpr int VH_SMEM = VH_OFF_MISC + 2048 + 1024;            // misc (barriers, bias / mask tile) + alignment slack

// MN-major SWIZZLE_128B operand whose 64-element mn-blocks are `lbo` bytes apart (tc_common's variant fixes 8192)
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn_lbo(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct VhMisc {
  uint64_t bar_load, bar_acc1, bar_g, bar_acc2;
  uint32_t tmem_ptr;
  uint32_t pad;
  float bias[VH_BV];
  uint8_t on[VH_BV];        // 1: column exists and is not masked off; 0: masked (logit = 0); 2: beyond the vocabulary
};

// common prologue: barriers, TMEM, operand loads, logits MMA.  Returns the TMEM base.
template <uint32_t TMEM_COLS>
__device__ __forceinline__ uint32_t vh_setup_and_logits(uint8_t* smem, VhMisc* mi, const CUtensorMap* tmap_d,
                                                        const CUtensorMap* tmap_w, int v0, int nkb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(tmap_d); tma_prefetch_desc(tmap_w); }
  if (warp == 1 && lane == 0) {
    mbar_init(&mi->bar_load, 1); mbar_init(&mi->bar_acc1, 1); mbar_init(&mi->bar_g, 1); mbar_init(&mi->bar_acc2, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&mi->tmem_ptr, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = mi->tmem_ptr;
  pdl_sync();
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(&mi->bar_load, (uint32_t)(2 * nkb * VH_KB_BYTES));
    for (int kb = 0; kb < nkb; ++kb) {
      tma_load_2d(smem + VH_OFF_D + kb * VH_KB_BYTES, tmap_d, &mi->bar_load, kb * 64, 0);      // rows >= M: zero fill
      tma_load_2d(smem + VH_OFF_W + kb * VH_KB_BYTES, tmap_w, &mi->bar_load, kb * 64, v0);     // rows >= V: zero fill
    }
  } else if (warp == 1) {
    mbar_wait(&mi->bar_load, 0);
    tcgen05_fence_after();
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, VH_BV);
      for (int kb = 0; kb < nkb; ++kb) {
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem + VH_OFF_D + kb * VH_KB_BYTES));
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + VH_OFF_W + kb * VH_KB_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
      }
      umma_commit(&mi->bar_acc1);
    }
    __syncwarp();
  }
  return tmem_base;
}

// epilogue warps: the tile's bias and column state into shared memory (overlaps the loads / MMA)
__device__ __forceinline__ void vh_stage_columns(VhMisc* mi, const __nv_bfloat16* __restrict__ bias,
                                                 const unsigned char* __restrict__ mask, int v0, int V) {
  const int t = threadIdx.x - 128;
  if (t >= 0 && t < VH_BV) {
    const int col = v0 + t;
    mi->bias[t] = col < V ? __bfloat162float(bias[col]) : 0.f;
    mi->on[t] = col < V ? ((mask == nullptr || mask[col]) ? 1 : 0) : 2;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

// --------------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(VH_THREADS, 1)
lm_vocab_fwd_kernel(const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_w,
                    const __nv_bfloat16* __restrict__ bias, const unsigned char* __restrict__ mask,
                    const long long* __restrict__ target, int M, int V, int E, int T, float* __restrict__ part_max,
                    float* __restrict__ part_sum, int* __restrict__ part_arg, float* __restrict__ ztgt,
                    float* __restrict__ zero_buf, int zero_n) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  VhMisc* mi = reinterpret_cast<VhMisc*>(smem + VH_OFF_MISC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, v0 = tile * VH_BV;
  const uint32_t tmem_base = vh_setup_and_logits<128>(smem, mi, &tmap_d, &tmap_w, v0, E / 64);
  if (warp >= 4) {
    // the backward kernel accumulates its dX partials into zero_buf: clear it here, one slice per CTA
    if (zero_buf != nullptr) {
      const int per = (zero_n / 4 + gridDim.x - 1) / gridDim.x;
      float4* zb = reinterpret_cast<float4*>(zero_buf);
      for (int i = threadIdx.x - 128; i < per; i += 128) {
        const int j = tile * per + i;
        if (j < zero_n / 4) zb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    vh_stage_columns(mi, bias, mask, v0, V);
    const int q = warp - 4;
    const int m = q * 32 + lane;
    mbar_wait(&mi->bar_acc1, 0);
    tcgen05_fence_after();
    const int tgt = m < M ? (int)target[m] : -1;
    float mx = -INFINITY, se = 0.f;
    int arg = 0x7fffffff;
#pragma unroll 1
    for (int c0 = 0; c0 < VH_BV; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      float z[32];
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint8_t on = mi->on[c0 + j];
        z[j] = on == 1 ? __uint_as_float(v[j]) + mi->bias[c0 + j] : (on == 0 ? 0.f : -INFINITY);
        cm = fmaxf(cm, z[j]);
      }
      if (cm == -INFINITY) continue;                      // chunk entirely beyond the vocabulary
      if (cm > mx) {
#pragma unroll
        for (int j = 31; j >= 0; --j) if (z[j] == cm) arg = v0 + c0 + j;     // first column holding the chunk maximum
        se *= __expf(mx - cm);                             // mx = -inf on the first chunk: se is 0 anyway
        mx = cm;
      }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) s += __expf(z[j] - mx);  // exp(-inf) = 0 for columns beyond the vocabulary
      se += s;
      if (tgt >= v0 + c0 && tgt < v0 + c0 + 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (v0 + c0 + j == tgt) ztgt[m] = z[j];
      }
    }
    if (m < M) {
      part_max[(size_t)m * T + tile] = mx;
      part_sum[(size_t)m * T + tile] = se;
      part_arg[(size_t)m * T + tile] = arg;
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------- log-sum-exp + loss + meter
// One block; warp w handles rows w, w + nwarps, ...; lanes stride over the T tile partials of a row.
__global__ void __launch_bounds__(1024)
lm_vocab_lse_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                    const int* __restrict__ part_arg, const float* __restrict__ ztgt,
                    const long long* __restrict__ target, int M, int T, float n_weight, float* __restrict__ lse_out,
                    float* __restrict__ loss_out, double* __restrict__ meter) {
  pdl_sync();
  __shared__ float sm[33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float loss_acc = 0.f, corr_acc = 0.f;
  for (int r = warp; r < M; r += nw) {
    float mx = -INFINITY, se = 0.f;
    int arg = 0x7fffffff;
    for (int t = lane; t < T; t += 32) {
      const float pm = part_max[(size_t)r * T + t], ps = part_sum[(size_t)r * T + t];
      const int pa = part_arg[(size_t)r * T + t];
      if (pm > mx) { se = se * __expf(mx - pm) + ps; mx = pm; arg = pa; }
      else if (pm > -INFINITY) { se += ps * __expf(pm - mx); if (pm == mx && pa < arg) arg = pa; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float omx = __shfl_xor_sync(0xffffffffu, mx, o), ose = __shfl_xor_sync(0xffffffffu, se, o);
      const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
      const float nmx = fmaxf(mx, omx);
      if (nmx > -INFINITY) {
        se = se * __expf(mx - nmx) + ose * __expf(omx - nmx);
        if (omx > mx || (omx == mx && oarg < arg)) arg = oarg;
        mx = nmx;
      }
    }
    const float lse = __logf(se) + mx;
    if (lane == 0) {
      lse_out[r] = lse;
      loss_acc += lse - ztgt[r];
      corr_acc += (arg == (int)target[r]) ? 1.f : 0.f;
    }
  }
  loss_acc = block_sum(loss_acc, sm);
  corr_acc = block_sum(corr_acc, sm);
  if (threadIdx.x == 0) {
    const float loss = loss_acc / (float)M;
    *loss_out = loss;
    if (meter != nullptr) {
      meter[0] += (double)loss * n_weight;
      meter[1] += (double)corr_acc;
      meter[2] += (double)n_weight;
      meter[3] += (double)__expf(loss) * n_weight;
    }
  }
}

// -------------------------------------------------------------------------------------------------- backward
__device__ __forceinline__ void vh_red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void vh_col_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

__global__ void __launch_bounds__(VH_THREADS, 1)
lm_vocab_bwd_kernel(const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_w,
                    const __nv_bfloat16* __restrict__ bias, const unsigned char* __restrict__ mask,
                    const long long* __restrict__ target, const float* __restrict__ lse, int M, int V, int E,
                    __nv_bfloat16* __restrict__ dW, int lddw, float* __restrict__ dbias, float* __restrict__ dX) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  VhMisc* mi = reinterpret_cast<VhMisc*>(smem + VH_OFF_MISC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, v0 = tile * VH_BV;
  const int nkb = E / 64;
  const uint32_t tmem_base = vh_setup_and_logits<512>(smem, mi, &tmap_d, &tmap_w, v0, nkb);
  const uint32_t acc_dw = tmem_base, acc_dx = tmem_base + 256;
  if (warp == 1) {
    // ---- second round of MMAs once the epilogue has published g in shared memory
    mbar_wait(&mi->bar_g, 0);
    tcgen05_fence_after();
    if (lane == 0) {
      const uint32_t g_addr = smem_u32(smem + VH_OFF_G), d_addr = smem_u32(smem + VH_OFF_D), w_addr = smem_u32(smem + VH_OFF_W);
      // dW2 tile [128 voc, E] = g^T . d : reduction over tokens (rows >= M of g and d are zero); both operands MN-major
      {
        const uint32_t idesc = make_idesc_bf16(128, E, true, true);
        const uint64_t adesc = make_smem_desc_sw128_mn_lbo(g_addr, VH_KB_BYTES);
        const uint64_t bdesc = make_smem_desc_sw128_mn_lbo(d_addr, VH_KB_BYTES);
        const int ksteps = (M + 15) / 16;
        for (int k = 0; k < ksteps; ++k)
          umma_bf16(acc_dw, adesc + (uint64_t)(2048 >> 4) * k, bdesc + (uint64_t)(2048 >> 4) * k, idesc, k != 0);
      }
      // dX partial [tok, E] = g . W2_tile : reduction over the tile's 128 vocabulary rows; g K-major, W2 MN-major
      {
        const uint32_t idesc = make_idesc_bf16(128, E, false, true);
        const uint64_t bdesc = make_smem_desc_sw128_mn_lbo(w_addr, VH_KB_BYTES);
        for (int k = 0; k < 8; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(g_addr + (k >> 2) * VH_KB_BYTES) + 2 * (k & 3);
          umma_bf16(acc_dx, adesc, bdesc + (uint64_t)(2048 >> 4) * k, idesc, k != 0);
        }
      }
      umma_commit(&mi->bar_acc2);
    }
    __syncwarp();
  } else if (warp >= 4) {
    vh_stage_columns(mi, bias, mask, v0, V);
    const int q = warp - 4;
    const int m = q * 32 + lane;
    const int tgt = m < M ? (int)target[m] : -1;
    const float row_lse = m < M ? lse[m] : 0.f;
    const float inv_rows = 1.f / (float)M;
    mbar_wait(&mi->bar_acc1, 0);
    tcgen05_fence_after();
    uint8_t* g_row = smem + VH_OFF_G + (m >> 3) * 1024 + (m & 7) * 128;
#pragma unroll 1
    for (int c0 = 0; c0 < VH_BV; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      float g[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint8_t on = mi->on[c0 + j];
        float gj = 0.f;
        if (on == 1 && m < M) gj = (__expf(__uint_as_float(v[j]) + mi->bias[c0 + j] - row_lse) - (v0 + c0 + j == tgt ? 1.f : 0.f)) * inv_rows;
        g[j] = gj;
      }
      // bf16 g -> the 128-byte-swizzled K-major tile [128 tok x 64 voc] x 2: row m, 16-byte chunk ((c0 % 64) / 8 + i)
      uint8_t* sub = g_row + (c0 >> 6) * VH_KB_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          h[e] = __floats2bfloat162_rn(g[8 * i + 2 * e], g[8 * i + 2 * e + 1]);
          g[8 * i + 2 * e] = __low2float(h[e]);
          g[8 * i + 2 * e + 1] = __high2float(h[e]);
        }
        const int chunk = ((c0 & 63) >> 3) + i;
        *reinterpret_cast<uint4*>(sub + ((chunk ^ (m & 7)) << 4)) = u;
      }
      // bias gradient: column sums of the (rounded) g over this warp's rows
      if (q * 32 < M) {
        vh_col_reduce32(g, lane);
        if (mi->on[c0 + lane] == 1) atomicAdd(&dbias[v0 + c0 + lane], g[0]);
      }
    }
    // generic-proxy smem writes -> visible to the tensor core (async proxy); TMEM reads ordered before the next MMAs
    fence_proxy_async();
    tcgen05_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 128) mbar_arrive(&mi->bar_g);
    mbar_wait(&mi->bar_acc2, 0);
    tcgen05_fence_after();
    // dW2 tile: accumulator row = vocabulary row
    {
      const int vr = v0 + m;
#pragma unroll 1
      for (int c0 = 0; c0 < E; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(acc_dw + (uint32_t(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (vr < V) {
          __nv_bfloat16* dst = dW + (size_t)vr * lddw + c0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              h[e] = __floats2bfloat162_rn(__uint_as_float(v[8 * i + 2 * e]), __uint_as_float(v[8 * i + 2 * e + 1]));
            *reinterpret_cast<uint4*>(dst + 8 * i) = u;
          }
        }
      }
    }
    // dX partial: accumulator row = token
    if (q * 32 < M) {
#pragma unroll 1
      for (int c0 = 0; c0 < E; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(acc_dx + (uint32_t(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (m < M) {
          float* dst = dX + (size_t)m * E + c0;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            vh_red_add4(dst + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                        __uint_as_float(v[4 * i + 3]));
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" int hfl_lm_vocab_supported(int M, int V, int E) {
  return (M >= 1 && M <= 128 && E % 64 == 0 && E >= 64 && E <= 256 && V >= 1) ? 0 : 1;
}
extern "C" int hfl_lm_vocab_tiles(int V) { return (V + VH_BV - 1) / VH_BV; }

static int vh_maps(CUtensorMap* td, CUtensorMap* tw, const void* d, int ldd, const void* W, int ldw, int M, int V, int E) {
  int r = make_tmap_2d_bf16(td, d, (uint64_t)E, (uint64_t)M, (uint64_t)ldd * 2, 64, 128);
  if (r) return 1000 + r;
  r = make_tmap_2d_bf16(tw, W, (uint64_t)E, (uint64_t)V, (uint64_t)ldw * 2, 64, 128);
  if (r) return 2000 + r;
  return 0;
}

// forward partials + (optionally) loss / meter.  `zero_buf` (fp32, zero_n elements, multiple of 4): cleared for the
// backward kernel's dX accumulation.
extern "C" int hfl_lm_vocab_fwd(const void* d, int ldd, const void* W, int ldw, const void* bias, const unsigned char* mask,
                                const long long* target, int M, int V, int E, float* part_max, float* part_sum,
                                int* part_arg, float* ztgt, float* lse, float n_weight, float* loss_out, double* meter,
                                float* zero_buf, int zero_n, cudaStream_t stream) {
  if (hfl_lm_vocab_supported(M, V, E) != 0) return -2;
  CUtensorMap td, tw;
  int r = vh_maps(&td, &tw, d, ldd, W, ldw, M, V, E);
  if (r) return r;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lm_vocab_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VH_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int T = hfl_lm_vocab_tiles(V);
  cudaError_t e = launch_k(lm_vocab_fwd_kernel, dim3(T), dim3(VH_THREADS), VH_SMEM, stream, td, tw,
                           (const __nv_bfloat16*)bias, mask, target, M, V, E, T, part_max, part_sum, part_arg, ztgt,
                           zero_buf, zero_n);
  if (e != cudaSuccess) return (int)e;
  e = launch_k(lm_vocab_lse_kernel, dim3(1), dim3(1024), 0, stream, (const float*)part_max, (const float*)part_sum,
               (const int*)part_arg, (const float*)ztgt, target, M, T, n_weight, lse, loss_out, meter);
  count_launch(2);
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

extern "C" int hfl_lm_vocab_bwd(const void* d, int ldd, const void* W, int ldw, const void* bias, const unsigned char* mask,
                                const long long* target, const float* lse, int M, int V, int E, void* dW, int lddw,
                                float* dbias, float* dX, cudaStream_t stream) {
  if (hfl_lm_vocab_supported(M, V, E) != 0) return -2;
  CUtensorMap td, tw;
  int r = vh_maps(&td, &tw, d, ldd, W, ldw, M, V, E);
  if (r) return r;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lm_vocab_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VH_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int T = hfl_lm_vocab_tiles(V);
  cudaError_t e = launch_k(lm_vocab_bwd_kernel, dim3(T), dim3(VH_THREADS), VH_SMEM, stream, td, tw,
                           (const __nv_bfloat16*)bias, mask, target, lse, M, V, E, (__nv_bfloat16*)dW, lddw, dbias, dX);
  count_launch();
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}
