// Resident client-training megakernel (sm_100a).
//
// The reference trains a client with ~300 kernel launches and 2 host syncs per batch-10 step
// (src/train_classifier_fed.py:190-210); round 1 of this repo cut that to ~160 graph nodes per step and then sat at the
// GPU front-end's node issue rate.  Here a client's WHOLE local training (all epochs x steps) is ONE persistent
// cooperative launch: the kernel walks a phase program (csrc/mega.h) -- batch assembly, every convolution of the
// forward and backward pass as tcgen05 implicit GEMMs fed by TMA, BatchNorm statistics / apply / backward, the
// classifier head with masked cross-entropy, clip + SGD -- separating dependent phases with a device-side grid barrier
// (~1 us) instead of a kernel boundary.  Several clients run concurrently as independent cooperative grids on disjoint
// SM subsets.  The same kernel runs the forward-only programs of the sBN statistics pass (ghost batches) and of
// evaluation.
//
//   OP_CONV   128-pixel x BLOCK_N tile per work item; A = 4-D TMA box of the NHWC activation at the tap's halo offset
//             (zero fill = padding, traversal stride = conv stride), B = 3-D TMA box of the K,R,S,C filter (K-major for
//             fprop, MN-major for dgrad: no transposed weight copy); 3-stage mbarrier ring; tcgen05.mma M128 x N x K16
//             into TMEM; epilogue: residual add, bf16 store (strided for the parity classes of a stride-2 dgrad) and
//             per-channel sum / sum-of-squares of the tile (BatchNorm batch statistics) folded into the same pass.
//             Channel counts below 64 use the same 128-byte-swizzle path: the TMA unit zero-fills the missing channels.
//   OP_WGRAD  both operands MN-major (pixel index is the reduction), split over pixel bricks, fp32 vector atomics straight
//             into the flat gradient arena the optimiser phase reads.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "mega.h"

namespace hfl {
namespace mega {

constexpr int MG_THREADS = 256;
constexpr int MG_STAGES = 5;
constexpr int MG_A_BYTES = 16384;
constexpr int MG_STAGE_BYTES = 32768;
constexpr int MG_WG_PITCH = 64 * 4 + 16;                       // wgrad epilogue staging row pitch (bytes)
constexpr int MG_WG_STAGING = 4 * 32 * MG_WG_PITCH;            // dedicated: the TMA producer runs ahead of the epilogue
constexpr int MG_SMEM = MG_STAGES * MG_STAGE_BYTES + MG_WG_STAGING + 1024 + 256;
constexpr uint32_t MG_TMEM_COLS = 256;                         // two 128-column accumulators (epilogue / MMA overlap)
constexpr int HDR_WORDS = 96;   // op..pv = 384 bytes of a Phase (the tensor maps are read in place from global memory)

struct Ctx {
  uint8_t* smem;
  uint64_t *full, *empty, *tfull, *tempty;   // smem ring; TMEM accumulator full / empty (x2)
  uint8_t* wg_stage;
  uint32_t tmem;
  unsigned* ctl;          // launch-wide: barrier counter, abort flag, diagnostics
  unsigned* jctl;         // per job: gradient-norm accumulator
  uint32_t kctr, tctr;   // running k-block / accumulator-tile counters: every warp advances them identically
  int nvalid, step, phase;
};

// ------------------------------------------------------------------------------------------- bounded waits / barrier
__device__ __forceinline__ bool aborted(unsigned* ctl) { return *reinterpret_cast<volatile unsigned*>(ctl + CTL_ABORT) != 0u; }

__device__ __noinline__ bool mbar_wait_slow(uint64_t* bar, uint32_t parity, unsigned* ctl) {
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i)
      if (mbar_try_wait(bar, parity)) return true;
    if (aborted(ctl)) return false;
    if (clock64() - t0 > 3000000000LL) { atomicExch(ctl + CTL_ABORT, 2u); return false; }
  }
}
__device__ __forceinline__ bool mbar_wait_b(uint64_t* bar, uint32_t parity, unsigned* ctl) {
  if (mbar_try_wait(bar, parity)) return true;
  return mbar_wait_slow(bar, parity, ctl);
}

// all CTAs of the (co-resident, cooperative) grid; `epoch` counts the barriers this CTA has passed
__device__ __forceinline__ void grid_sync(unsigned* ctl, unsigned& epoch) {
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    // release: every global write of this CTA (made visible to thread 0 by the bar.sync above) happens-before the add
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctl + CTL_ARRIVE), "r"(1u) : "memory");
    const unsigned target = epoch * gridDim.x;
    const long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctl + CTL_ARRIVE) : "memory");
      if (v >= target) break;
      if (((++spins) & 255u) == 0u) {
        if (aborted(ctl)) break;
        if (clock64() - t0 > 6000000000LL) { atomicExch(ctl + CTL_ABORT, 1u); break; }
      }
    }
  }
  __syncthreads();
  fence_proxy_async();
}

// Producer-side sanity check before the first TMA of a tile: the header this CTA works from must be the record the
// tensor maps are read from, and the map must have been encoded (a zero descriptor is an illegal instruction).  Turns a
// would-be device fault into an abort code + diagnostics in the control block (words 8..13).
__device__ __noinline__ bool tile_guard(const Phase* __restrict__ Pg, const Phase& P, Ctx& cx, int tile, int map) {
  const int gop = __ldcg(reinterpret_cast<const int*>(&Pg->op));
  const unsigned long long d0 = __ldcg(reinterpret_cast<const unsigned long long*>(&Pg->tm[map]));
  if (gop == P.op && d0 != 0ull) return true;
  if (atomicCAS(cx.ctl + CTL_ABORT, 0u, 4u) == 0u) {
    cx.ctl[8] = (unsigned)cx.phase; cx.ctl[9] = (unsigned)tile; cx.ctl[10] = (unsigned)P.op; cx.ctl[11] = (unsigned)gop;
    cx.ctl[12] = blockIdx.x; cx.ctl[13] = (unsigned)cx.step;
  }
  return false;
}

// ----------------------------------------------------------------------------------------------------- small helpers
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { f[2 * e] = __low2float(h[e]); f[2 * e + 1] = __high2float(h[e]); }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return u;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&f)[8]) { unpack8(__ldcg(reinterpret_cast<const uint4*>(p)), f); }
__device__ __forceinline__ void st8(bf16* p, const float (&f)[8]) { *reinterpret_cast<uint4*>(p) = pack8(f); }
__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ float ldbf(const bf16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}
// fire-and-forget fp32 reductions on GLOBAL addresses (atomicAdd on a generic pointer compiles to a shared/global
// dispatch with a CAS loop for the shared case)
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, const float4& v) {
  asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float bfr(float v) { return __bfloat162float(__float2bfloat16(v)); }

// Column sums over the 32 lanes of a warp: on return lane j holds sum_over_lanes(v[j]) in v[0]  (31 shuffles).
__device__ __forceinline__ void warp_col_reduce(float (&v)[32], int lane) {
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

// ------------------------------------------------------------------------------------------------------------ OP_CONV
// iv: 0 N  1 OH  2 OW  3 bn  4 bh  5 bw  6 tiles_h  7 m_tiles  8 n_tiles  9 block_n  10 Cy  11 out_H  12 out_W
//     13 out_s  14 out_ph  15 out_pw  16 in_stride  17 b_mn  18 ntaps  19 has_res  20 has_stats  21 group_n
//     22 num_k  23 a_bytes  24 b_bytes  25 use_nvalid  26..43 taps (2 words each: 9 taps);  header word `flags` = ksplit
// pv: 0 out  1 residual  2 stats [G][2][Cy]  3 split-K scratch (fp32, same indexing as out)  11 split-K tile counters
//     4 y  5 S  6 gamma  7 beta  8 T   (dgrad only, pv[4] != 0: the BatchNorm backward reductions of the layer that
//     produced y run in this epilogue -- out = dz * [bn(y) > 0], T += (sum out, sum out * xhat));  fv: 0 eps
// tm: 0,1 activation maps   2,3 filter maps
//
// The three roles (TMA producer = warp 0, MMA issuer = warp 1, epilogue = warps 4-7) each walk the CTA's tile sequence
// of the phase on their own: the producer runs ahead through the smem ring, the MMA warp alternates between two TMEM
// accumulators, so the epilogue of tile i overlaps the loads and MMAs of tile i+1.
// Split-K (flags > 1): one SM pulls ~60-80 GB/s from L2 / HBM, so a 128 x 128 tile with 72 k-blocks (2.3 MB of operands,
// the 512-channel layers at batch 10) takes 36-70 us on its own.  The k-blocks of such a tile are dealt to `ksplit` work
// items; each stores its partial accumulator into its own fp32 slice with 16-byte stores and the OP_CFIN phase that
// follows sums the slices with every CTA of the grid (coalesced, no atomics, no serial last-arrival reduction:
// bench/mega_conv_probe.py has the measurements that ruled those out).
struct ConvTileId { int n_blk, n0, oh0, kb0, kb1, mn; };

__device__ __forceinline__ ConvTileId conv_decode(const Phase& P, int lt) {
  ConvTileId c;
  const int* iv = P.iv;
  const int ksplit = P.flags > 1 ? P.flags : 1;
  const int ks = lt % ksplit;
  c.mn = lt / ksplit;
  const int n_tiles = iv[8];
  c.n_blk = c.mn % n_tiles;
  const int m_tile = c.mn / n_tiles;
  const int tiles_h = iv[6];
  const int tile_n = m_tile / tiles_h, tile_h = m_tile - tile_n * tiles_h;
  c.n0 = tile_n * iv[3];
  c.oh0 = tile_h * iv[4];
  const int num_k = iv[22];
  const int per = (num_k + ksplit - 1) / ksplit;
  c.kb0 = ks * per;
  c.kb1 = min(num_k, c.kb0 + per);
  return c;
}

// warp 0, lane 0
__device__ __noinline__ bool conv_produce(const Phase* __restrict__ Pg, const Phase& P, Ctx& cx, int lt) {
  const int* iv = P.iv;
  const ConvTileId c = conv_decode(P, lt);
  const int BN = iv[9], b_mn = iv[17], ntaps = iv[18], in_stride = iv[16];
  const uint32_t tx = (uint32_t)(iv[23] + iv[24]);
  if (cx.step == 0 && !(tile_guard(Pg, P, cx, lt, 0) && tile_guard(Pg, P, cx, lt, 2))) return false;
  int kb = 0;
  for (int t = 0; t < ntaps; ++t) {
    const int w0 = iv[26 + 2 * t], w1 = iv[27 + 2 * t];
    const int a_map = w0 & 15, b_map = (w0 >> 4) & 15, b_tap = (w0 >> 8) & 255, kb_t = (w0 >> 16) & 0xffff;
    if (kb + kb_t <= c.kb0 || kb >= c.kb1) { kb += kb_t; continue; }
    const int dh = (int)(short)(w1 & 0xffff), dw = (int)(short)((w1 >> 16) & 0xffff);
    const CUtensorMap* am = &Pg->tm[a_map];
    const CUtensorMap* bm = &Pg->tm[2 + b_map];
    for (int cb = 0; cb < kb_t; ++cb, ++kb) {
      if (kb < c.kb0 || kb >= c.kb1) continue;
      const uint32_t kc = cx.kctr++;
      const int st = kc % MG_STAGES;
      const uint32_t ph = (kc / MG_STAGES) & 1;
      if (!mbar_wait_b(&cx.empty[st], ph ^ 1, cx.ctl)) return false;
      uint8_t* sa = cx.smem + st * MG_STAGE_BYTES;
      uint8_t* sb = sa + MG_A_BYTES;
      mbar_expect_tx(&cx.full[st], tx);
      tma_load_4d(sa, am, &cx.full[st], cb * 64, dw, c.oh0 * in_stride + dh, c.n0);
      if (!b_mn) {
        tma_load_3d(sb, bm, &cx.full[st], cb * 64, b_tap, c.n_blk * BN);
      } else {
        for (int j = 0; j < BN / 64; ++j)
          tma_load_3d(sb + j * 8192, bm, &cx.full[st], c.n_blk * BN + j * 64, b_tap, cb * 64);
      }
    }
  }
  return true;
}

// warp 1 (all lanes; lane 0 issues)
__device__ __noinline__ bool conv_mma(const Phase& P, Ctx& cx, int lt) {
  const int* iv = P.iv;
  const ConvTileId c = conv_decode(P, lt);
  const int lane = threadIdx.x & 31;
  const int BN = iv[9], b_mn = iv[17];
  const int nk = c.kb1 - c.kb0;
  const uint32_t idesc = make_idesc_bf16(128, BN, false, b_mn != 0);
  const uint32_t tc = cx.tctr++;
  const uint32_t acc = tc & 1u, use = tc >> 1;
  if (!mbar_wait_b(&cx.tempty[acc], (use & 1u) ^ 1u, cx.ctl)) return false;   // epilogue drained this accumulator
  tcgen05_fence_after();
  const uint32_t tmem_d = cx.tmem + acc * 128u;
  for (int i = 0; i < nk; ++i) {
    const uint32_t kc = cx.kctr++;
    const int st = kc % MG_STAGES;
    const uint32_t ph = (kc / MG_STAGES) & 1;
    if (!mbar_wait_b(&cx.full[st], ph, cx.ctl)) return false;
    tcgen05_fence_after();
    if (lane == 0) {
      const uint32_t a_addr = smem_u32(cx.smem + st * MG_STAGE_BYTES);
      const uint64_t adesc = make_smem_desc_sw128(a_addr);
      const uint64_t bdesc = b_mn ? make_smem_desc_sw128_mn(a_addr + MG_A_BYTES) : make_smem_desc_sw128(a_addr + MG_A_BYTES);
      const uint64_t b_step = b_mn ? (uint64_t)(2048 >> 4) : (uint64_t)(32 >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + b_step * k, idesc, (i | k) != 0);
      umma_commit(&cx.empty[st]);
      if (i == nk - 1) umma_commit(&cx.tfull[acc]);
    }
    __syncwarp();
  }
  return true;
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// warps 4..7
__device__ __noinline__ bool conv_epilogue(const Phase& P, Ctx& cx, int lt, int* s_flag, float* s_bn /* [4][4][32] */) {
  const int* iv = P.iv;
  const ConvTileId c = conv_decode(P, lt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp - 4;
  const int BN = iv[9];
  const int bn = iv[3], bh = iv[4], bw = iv[5];
  const uint32_t tc = cx.tctr++;
  const uint32_t acc = tc & 1u, use = tc >> 1;
  if (!mbar_wait_b(&cx.tfull[acc], use & 1u, cx.ctl)) return false;
  tcgen05_fence_after();
  const uint32_t tmem_d = cx.tmem + acc * 128u;
  const int m = q * 32 + lane;                       // pixel within the brick: ((n_l*bh + h_l)*bw + w_l)
  const int w_l = m % bw;
  const int t = m / bw;
  const int h_l = t % bh, n_l = t / bh;
  const int n = c.n0 + n_l, oh = c.oh0 + h_l;
  const int nlim = iv[25] ? min(iv[0], cx.nvalid) : iv[0];
  const bool ok = (m < bn * bh * bw) && (n < nlim) && (oh < iv[1]);
  const int Cy = iv[10];
  const size_t pix = ((size_t)n * iv[11] + (size_t)(oh * iv[13] + iv[14])) * iv[12] + (size_t)(w_l * iv[13] + iv[15]);
  bf16* Y = reinterpret_cast<bf16*>(P.pv[0]);
  const bf16* res = iv[19] ? reinterpret_cast<const bf16*>(P.pv[1]) : nullptr;
  float* stats = nullptr;
  if (iv[20]) {
    const int g = iv[21] > 0 ? c.n0 / iv[21] : 0;
    stats = reinterpret_cast<float*>(P.pv[2]) + (size_t)g * 2 * Cy;
  }
  const int ksplit = P.flags > 1 ? P.flags : 1;
  float* scratch = reinterpret_cast<float*>(P.pv[3]);
  const bf16* bn_y = reinterpret_cast<const bf16*>(P.pv[4]);
  const float* bn_S = reinterpret_cast<const float*>(P.pv[5]);
  const bf16* bn_g = reinterpret_cast<const bf16*>(P.pv[6]);
  const bf16* bn_b = reinterpret_cast<const bf16*>(P.pv[7]);
  float* bn_T = reinterpret_cast<float*>(P.pv[8]);
  const float bn_cnt = (float)nlim * (float)(iv[11] * iv[12]);
  const bool finalize = true;
  const size_t slice = (size_t)iv[0] * iv[11] * iv[12] * Cy;      // one fp32 partial slice = the output tensor's extent
  if (ksplit > 1) {
    // ---- partial accumulator -> this split's fp32 slice (plain 16-byte stores: no contended atomics)
    const int ks = lt % ksplit;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      const int col0 = c.n_blk * BN + c0;
      if (col0 >= Cy) break;
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_d + (uint32_t(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (ok) {
        uint4* dst = reinterpret_cast<uint4*>(scratch + (size_t)ks * slice + pix * Cy + col0);
        const int ncol = min(32, Cy - col0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (4 * j < ncol) __stcg(dst + j, make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
      }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&cx.tempty[acc]);
    return true;                                       // the OP_CFIN phase that follows sums the slices
  }
  if (finalize) {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      const int col0 = c.n_blk * BN + c0;
      if (col0 >= Cy) break;                           // warp-uniform
      const int ncol = min(32, Cy - col0);             // multiple of 8
      float f[32];
      {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_d + (uint32_t(q * 32) << 16) + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = ok ? __uint_as_float(v[j]) : 0.f;
      }
      if (bn_y != nullptr) {
        // BatchNorm-backward reductions fused into the dgrad epilogue
        float* tb = s_bn + q * 128;
        {
          const int col = col0 + lane;
          float a = 0.f, bsh = 0.f, mean = 0.f, inv = 0.f;
          if (lane < ncol && bn_cnt > 0.f) {
            const float sm = __ldcg(bn_S + col), sq = __ldcg(bn_S + Cy + col);
            mean = sm / bn_cnt;
            inv = rsqrtf(fmaxf(sq / bn_cnt - mean * mean, 0.f) + P.fv[0]);
            a = ldbf(bn_g + col) * inv;
            bsh = ldbf(bn_b + col) - mean * a;
          }
          __syncwarp();
          tb[lane] = a; tb[32 + lane] = bsh; tb[64 + lane] = mean; tb[96 + lane] = inv;
          __syncwarp();
        }
        float sx[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) sx[j] = 0.f;
        if (ok) {
          const bf16* yp = bn_y + pix * Cy + col0;
          bf16* dst = Y + pix * Cy + col0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (8 * j < ncol) {
              float y8[8], o8[8];
              ld8(yp + 8 * j, y8);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int c = 8 * j + e;
                const float pre = fmaf(tb[c], y8[e], tb[32 + c]);
                o8[e] = pre > 0.f ? f[c] : 0.f;
              }
              const uint4 u = pack8(o8);
              *reinterpret_cast<uint4*>(dst + 8 * j) = u;
              unpack8(u, o8);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int c = 8 * j + e;
                f[c] = o8[e];
                sx[c] = o8[e] * (y8[e] - tb[64 + c]) * tb[96 + c];
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[8 * j + e] = 0.f;
            }
          }
        }
        warp_col_reduce(f, lane);
        warp_col_reduce(sx, lane);
        if (lane < ncol) {
          red_add(&bn_T[col0 + lane], f[0]);
          red_add(&bn_T[Cy + col0 + lane], sx[0]);
        }
        continue;
      }
      if (ok) {
        bf16* dst = Y + pix * Cy + col0;
        if (res != nullptr) {
          const bf16* rp = res + pix * Cy + col0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (8 * j < ncol) {
              float r8[8];
              ld8(rp + 8 * j, r8);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[8 * j + e] += r8[e];
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o8[e] = f[8 * j + e];
          const uint4 u = pack8(o8);
          if (8 * j < ncol) *reinterpret_cast<uint4*>(dst + 8 * j) = u;
          float b8[8];
          unpack8(u, b8);                              // statistics of the bf16-rounded values (what a reader sees)
#pragma unroll
          for (int e = 0; e < 8; ++e) f[8 * j + e] = b8[e];
        }
      }
      if (stats != nullptr) {
        float s2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) s2[j] = f[j] * f[j];
        warp_col_reduce(f, lane);
        warp_col_reduce(s2, lane);
        if (lane < ncol) {
          red_add(&stats[col0 + lane], f[0]);
          red_add(&stats[Cy + col0 + lane], s2[0]);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&cx.tempty[acc]);
  return true;
}

// ----------------------------------------------------------------------------------------------------------- OP_WGRAD
// iv: 0 N  1 OH  2 OW  3 bn  4 bh  5 bw  6 bricks_h  7 bricks  8 splits  9 C_real  10 Kout_real  11 taps  12 S
//     13 stride  14 pad  15 k_blks  16 c_blks  17 a_boxes
// pv: 0 dW (fp32 [Kout][taps][C_real]);   tm: 0 dY map, 1 X map
struct WgTileId { int k_blk, c_blk, tap, r, s, b0, num_k; };

__device__ __forceinline__ WgTileId wg_decode(const int* iv, int lt) {
  WgTileId w;
  const int splits = iv[8], taps = iv[11], c_blks = iv[16];
  int t = lt;
  const int split = t % splits; t /= splits;
  w.tap = t % taps; t /= taps;
  w.c_blk = t % c_blks;
  w.k_blk = t / c_blks;
  w.r = w.tap / iv[12];
  w.s = w.tap - w.r * iv[12];
  const int bricks = iv[7];
  const int per = (bricks + splits - 1) / splits;
  w.b0 = split * per;
  w.num_k = max(0, min(w.b0 + per, bricks) - w.b0);
  return w;
}

__device__ __noinline__ bool wgrad_produce(const Phase* __restrict__ Pg, const Phase& P, Ctx& cx, int lt) {
  const int* iv = P.iv;
  const WgTileId w = wg_decode(iv, lt);
  if (w.num_k <= 0) return true;
  if (cx.step == 0 && !(tile_guard(Pg, P, cx, lt, 0) && tile_guard(Pg, P, cx, lt, 1))) return false;
  const int a_boxes = iv[17];
  const int bricks_h = iv[6], bn = iv[3], bh = iv[4], stride = iv[13], pad = iv[14];
  const uint32_t tx = (uint32_t)(a_boxes * 8192 + 8192);
  for (int kb = 0; kb < w.num_k; ++kb) {
    const uint32_t kc = cx.kctr++;
    const int st = kc % MG_STAGES;
    const uint32_t ph = (kc / MG_STAGES) & 1;
    const int brick = w.b0 + kb;
    const int tn = brick / bricks_h, th = brick - tn * bricks_h;
    const int n0 = tn * bn, oh0 = th * bh;
    if (!mbar_wait_b(&cx.empty[st], ph ^ 1, cx.ctl)) return false;
    uint8_t* sa = cx.smem + st * MG_STAGE_BYTES;
    uint8_t* sb = sa + MG_A_BYTES;
    mbar_expect_tx(&cx.full[st], tx);
    tma_load_4d(sa, &Pg->tm[0], &cx.full[st], w.k_blk * 128, 0, oh0, n0);
    if (a_boxes == 2) tma_load_4d(sa + 8192, &Pg->tm[0], &cx.full[st], w.k_blk * 128 + 64, 0, oh0, n0);
    tma_load_4d(sb, &Pg->tm[1], &cx.full[st], w.c_blk * 64, w.s - pad, oh0 * stride + w.r - pad, n0);
  }
  return true;
}

__device__ __noinline__ bool wgrad_mma(const Phase& P, Ctx& cx, int lt) {
  const WgTileId w = wg_decode(P.iv, lt);
  if (w.num_k <= 0) return true;
  const int lane = threadIdx.x & 31;
  const uint32_t idesc = make_idesc_bf16(128, 64, true, true);
  const uint32_t tc = cx.tctr++;
  const uint32_t acc = tc & 1u, use = tc >> 1;
  if (!mbar_wait_b(&cx.tempty[acc], (use & 1u) ^ 1u, cx.ctl)) return false;
  tcgen05_fence_after();
  const uint32_t tmem_d = cx.tmem + acc * 128u;
  for (int kb = 0; kb < w.num_k; ++kb) {
    const uint32_t kc = cx.kctr++;
    const int st = kc % MG_STAGES;
    const uint32_t ph = (kc / MG_STAGES) & 1;
    if (!mbar_wait_b(&cx.full[st], ph, cx.ctl)) return false;
    tcgen05_fence_after();
    if (lane == 0) {
      const uint32_t a_addr = smem_u32(cx.smem + st * MG_STAGE_BYTES);
      const uint64_t adesc = make_smem_desc_sw128_mn(a_addr);
      const uint64_t bdesc = make_smem_desc_sw128_mn(a_addr + MG_A_BYTES);
#pragma unroll
      for (int k = 0; k < 4; ++k)      // 64 pixels = 4 x UMMA_K(16); +16 k-rows = +2048 B
        umma_bf16(tmem_d, adesc + (2048 >> 4) * k, bdesc + (2048 >> 4) * k, idesc, (kb | k) != 0);
      umma_commit(&cx.empty[st]);
      if (kb == w.num_k - 1) umma_commit(&cx.tfull[acc]);
    }
    __syncwarp();
  }
  return true;
}

__device__ __noinline__ bool wgrad_epilogue(const Phase& P, Ctx& cx, int lt) {
  const int* iv = P.iv;
  const WgTileId w = wg_decode(iv, lt);
  if (w.num_k <= 0) return true;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp - 4;
  const uint32_t tc = cx.tctr++;
  const uint32_t acc = tc & 1u, use = tc >> 1;
  if (!mbar_wait_b(&cx.tfull[acc], use & 1u, cx.ctl)) return false;
  tcgen05_fence_after();
  const uint32_t tmem_d = cx.tmem + acc * 128u;
  // stage the warp's 32 x 64 fp32 block in (dedicated) shared memory, then accumulate whole filter rows with 16-byte
  // reductions (a row per thread would touch 32 different rows per instruction)
  uint8_t* my = cx.wg_stage + (size_t)q * 32 * MG_WG_PITCH;
  {
    uint8_t* myrow = my + lane * MG_WG_PITCH;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_d + (uint32_t(q * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(myrow + (c0 + j) * 4) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&cx.tempty[acc]);          // the accumulator is free: the next tile's MMAs may start
  const int C = iv[9], Kout = iv[10], taps = iv[11];
  float* dW = reinterpret_cast<float*>(P.pv[0]);
  const int rsub = lane >> 4, cseg = (lane & 15) * 4;
  const bool vec = (C & 3) == 0;
#pragma unroll 4
  for (int r0 = 0; r0 < 32; r0 += 2) {
    const int k = w.k_blk * 128 + q * 32 + r0 + rsub;
    const int cc = w.c_blk * 64 + cseg;
    if (k < Kout && cc < C) {
      const float4 val = *reinterpret_cast<const float4*>(my + (r0 + rsub) * MG_WG_PITCH + cseg * 4);
      float* dst = dW + ((size_t)k * taps + w.tap) * C + cc;
      if (vec) {
        red_add4(dst, val);
      } else {
        red_add(dst, val.x);
        if (cc + 1 < C) red_add(dst + 1, val.y);
        if (cc + 2 < C) red_add(dst + 2, val.z);
        if (cc + 3 < C) red_add(dst + 3, val.w);
      }
    }
  }
  __syncwarp();                                          // staging rows are rewritten by the next tile
  return true;
}

// ------------------------------------------------------------------------------------------- normalisation phases
// shared conventions: activations are [rows, Cp] bf16 with Cp a power of two >= 8; thread (cv, r0) owns 8 channels
// iv: 0 M_alloc  1 Cp  2 C_real  3 HW  4 mode  5 relu  6 group_n  7 acc_running  8 rows_per_item  9 use_nvalid  10 N
struct NormCol {
  float a[8], b[8], mean[8], inv[8], gam[8];
  float cnt;
};

// per-channel affine of the 8 channels at c0 from the (sum, sumsq) statistics of group g
__device__ __forceinline__ void norm_setup(const Phase& P, int c0, int g, float cnt, NormCol& nc) {
  const int Cp = P.iv[1], Cr = P.iv[2];
  const float* S = reinterpret_cast<const float*>(P.pv[2]) + (size_t)g * 2 * Cp;
  float sm[8], sq[8], ga[8], be[8];
  ld8f(S + c0, sm);
  ld8f(S + Cp + c0, sq);
  const bf16* gp = reinterpret_cast<const bf16*>(P.pv[3]);
  const bf16* bp = reinterpret_cast<const bf16*>(P.pv[4]);
  const float eps = P.fv[0];
  nc.cnt = cnt;
  const float ic = cnt > 0.f ? 1.f / cnt : 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const bool real = (c0 + e) < Cr && cnt > 0.f;
    ga[e] = real ? __bfloat162float(gp[c0 + e]) : 0.f;    // parameters are tiny and L2-resident: scalar loads
    be[e] = real ? __bfloat162float(bp[c0 + e]) : 0.f;
    const float mean = sm[e] * ic;
    const float var = fmaxf(sq[e] * ic - mean * mean, 0.f);
    const float inv = rsqrtf(var + eps);
    nc.mean[e] = mean;
    nc.inv[e] = real ? inv : 0.f;
    nc.gam[e] = ga[e];
    nc.a[e] = ga[e] * nc.inv[e];
    nc.b[e] = be[e] - mean * nc.a[e];
  }
}

// OP_NAPPLY   pv: 0 y  1 z  2 stats  3 gamma  4 beta  5 acc_mean  6 acc_var  7 a  8 b;   fv: 0 eps 1 inv_s 2 inv_s2
__device__ __noinline__ void napply_item(const Phase& P, Ctx& cx, int item) {
  const int* iv = P.iv;
  const int M = iv[0], Cp = iv[1], HW = iv[3];
  const int nvec = Cp >> 3, rstep = MG_THREADS / nvec;
  const int cv = threadIdx.x % nvec, r0 = threadIdx.x / nvec;
  const int c0 = cv * 8;
  const int rpi = iv[8];
  const int row_lo = item * rpi, row_hi = min(row_lo + rpi, M);
  const int nlim = iv[9] ? min(iv[10], cx.nvalid) : iv[10];
  float a[8], b[8];
  if (iv[4] == 0) {
    const int gn = iv[6];
    const int g = gn > 0 ? (row_lo / HW) / gn : 0;
    const int nv = gn > 0 ? max(0, min(nlim - g * gn, gn)) : nlim;
    NormCol nc;
    norm_setup(P, c0, g, (float)nv * (float)HW, nc);
#pragma unroll
    for (int e = 0; e < 8; ++e) { a[e] = nc.a[e]; b[e] = nc.b[e]; }
    if (iv[7] && nv > 0 && r0 == 0 && (row_lo % (max(gn, 1) * HW)) == 0) {
      // sBN: running statistics are the mean over ghost batches of (batch mean, UNBIASED batch variance), taken on the
      // un-scaled activations (Scaler is folded into epsilon): reference src/train_classifier_fed.py:127-138
      float* am = reinterpret_cast<float*>(P.pv[5]);
      float* av = reinterpret_cast<float*>(P.pv[6]);
      const float cnt = nc.cnt;
      const float* S = reinterpret_cast<const float*>(P.pv[2]) + (size_t)g * 2 * Cp;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (c0 + e < iv[2]) {
          const float mean = nc.mean[e];
          const float var = fmaxf(__ldcg(S + Cp + c0 + e) / cnt - mean * mean, 0.f);
          red_add(am + c0 + e, mean * P.fv[1]);
          red_add(av + c0 + e, var * (cnt / fmaxf(cnt - 1.f, 1.f)) * P.fv[2]);
        }
      }
    }
  } else {
    ld8f(reinterpret_cast<const float*>(P.pv[7]) + c0, a);
    ld8f(reinterpret_cast<const float*>(P.pv[8]) + c0, b);
  }
  const bf16* Y = reinterpret_cast<const bf16*>(P.pv[0]);
  bf16* Z = reinterpret_cast<bf16*>(P.pv[1]);
  const bool relu = iv[5] != 0;
  for (int row = row_lo + r0; row < row_hi; row += 4 * rstep) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                       // 4 independent 16-byte loads in flight per thread
      const int r = row + k * rstep;
      u[k] = r < row_hi ? __ldcg(reinterpret_cast<const uint4*>(Y + (size_t)r * Cp + c0)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = row + k * rstep;
      if (r < row_hi) {
        float y[8];
        unpack8(u[k], y);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float z = fmaf(a[e], y[e], b[e]);
          y[e] = relu ? fmaxf(z, 0.f) : z;
        }
        st8(Z + (size_t)r * Cp + c0, y);
      }
    }
  }
}

// OP_NBWD_RED / OP_NBWD_APPLY
// pv: 0 y  1 dz  2 stats S  3 gamma  4 beta  5 dz2  6 T [2][Cp]  7 dres  8 dy  9 dgamma  10 dbeta;  fv: 0 eps
__device__ __noinline__ void nbwd_item(const Phase& P, Ctx& cx, int item, bool apply) {
  const int* iv = P.iv;
  const int M = iv[0], Cp = iv[1], Cr = iv[2], HW = iv[3];
  const int nvec = Cp >> 3, rstep = MG_THREADS / nvec;
  const int cv = threadIdx.x % nvec, r0 = threadIdx.x / nvec;
  const int c0 = cv * 8;
  const int rpi = iv[8];
  const int row_lo = item * rpi, row_hi = min(row_lo + rpi, M);
  const int nlim = iv[9] ? min(iv[10], cx.nvalid) : iv[10];
  NormCol nc;
  norm_setup(P, c0, 0, (float)nlim * (float)HW, nc);
  const bf16* Y = reinterpret_cast<const bf16*>(P.pv[0]);
  const bf16* DZ = reinterpret_cast<const bf16*>(P.pv[1]);
  const bf16* DZ2 = reinterpret_cast<const bf16*>(P.pv[5]);
  float* T = reinterpret_cast<float*>(P.pv[6]);
  const bool relu = iv[5] != 0;
  const int row_lim = nlim * HW;
  if (!apply) {
    float sb[8], sg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sb[e] = 0.f; sg[e] = 0.f; }
#pragma unroll 2
    for (int row = row_lo + r0; row < min(row_hi, row_lim); row += rstep) {
      float y[8], dz[8];
      ld8(Y + (size_t)row * Cp + c0, y);
      ld8(DZ + (size_t)row * Cp + c0, dz);
      if (DZ2 != nullptr) {
        float d2[8];
        ld8(DZ2 + (size_t)row * Cp + c0, d2);
#pragma unroll
        for (int e = 0; e < 8; ++e) dz[e] += d2[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float pre = fmaf(nc.a[e], y[e], nc.b[e]);
        const float g = (relu && !(pre > 0.f)) ? 0.f : dz[e];
        const float xh = (y[e] - nc.mean[e]) * nc.inv[e];
        sb[e] += g;
        sg[e] = fmaf(g, xh, sg[e]);
      }
    }
    // reduce over the r0 dimension through shared memory (idle pipeline stages)
    float* red = reinterpret_cast<float*>(cx.smem);       // [256][16]
#pragma unroll
    for (int e = 0; e < 8; ++e) { red[threadIdx.x * 16 + e] = sb[e]; red[threadIdx.x * 16 + 8 + e] = sg[e]; }
    __syncthreads();
    for (int half = rstep >> 1; half >= 1; half >>= 1) {
      if (r0 < half) {
        float* mine = red + threadIdx.x * 16;
        const float* other = red + (threadIdx.x + half * nvec) * 16;
#pragma unroll
        for (int e = 0; e < 16; ++e) mine[e] += other[e];
      }
      __syncthreads();
    }
    if (r0 == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (c0 + e < Cr) {
          red_add(T + c0 + e, red[threadIdx.x * 16 + e]);
          red_add(T + Cp + c0 + e, red[threadIdx.x * 16 + 8 + e]);
        }
      }
    }
    __syncthreads();
  } else {
    float sb[8], sg[8];
    ld8f(T + c0, sb);
    ld8f(T + Cp + c0, sg);
    const float inv_n = nc.cnt > 0.f ? 1.f / nc.cnt : 0.f;
    const bf16* DR = reinterpret_cast<const bf16*>(P.pv[7]);
    bf16* DY = reinterpret_cast<bf16*>(P.pv[8]);
    if (item == 0 && r0 == 0) {
      float* dga = reinterpret_cast<float*>(P.pv[9]);
      float* dbe = reinterpret_cast<float*>(P.pv[10]);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c0 + e < Cr) { dga[c0 + e] = sg[e]; dbe[c0 + e] = sb[e]; }
    }
    for (int row = row_lo + r0; row < row_hi; row += 2 * rstep) {
      uint4 uy[2], ud[2], u2[2], ur[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {                     // all loads of two rows in flight before any is consumed
        const int r = row + k * rstep;
        const bool in = r < row_hi && r < row_lim;
        const size_t off = (size_t)r * Cp + c0;
        uy[k] = in ? __ldcg(reinterpret_cast<const uint4*>(Y + off)) : make_uint4(0, 0, 0, 0);
        ud[k] = in ? __ldcg(reinterpret_cast<const uint4*>(DZ + off)) : make_uint4(0, 0, 0, 0);
        u2[k] = (in && DZ2 != nullptr) ? __ldcg(reinterpret_cast<const uint4*>(DZ2 + off)) : make_uint4(0, 0, 0, 0);
        ur[k] = (in && DR != nullptr) ? __ldcg(reinterpret_cast<const uint4*>(DR + off)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int r = row + k * rstep;
        if (r >= row_hi) continue;
        float o[8];
        if (r < row_lim) {
          float y[8], dz[8], d2[8], dr[8];
          unpack8(uy[k], y);
          unpack8(ud[k], dz);
          unpack8(u2[k], d2);
          unpack8(ur[k], dr);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float dzz = dz[e] + d2[e];
            const float pre = fmaf(nc.a[e], y[e], nc.b[e]);
            const float g = (relu && !(pre > 0.f)) ? 0.f : dzz;
            const float xh = (y[e] - nc.mean[e]) * nc.inv[e];
            o[e] = nc.a[e] * (g - sb[e] * inv_n - xh * sg[e] * inv_n) + dr[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = 0.f;
        }
        st8(DY + (size_t)r * Cp + c0, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------ OP_CFIN
// iv: 0 M_alloc  1 Cp  2 C_real  3 HW  4 ksplit  5 has_res  6 bnred  7 has_stats  8 rows_per_item  9 use_nvalid  10 N
// pv: 0 out  1 residual  2 stats [2][Cp]  3 slices (fp32 [ksplit][M][Cp])  4 y  5 S  6 gamma  7 beta  8 T;  fv: 0 eps
__device__ __noinline__ void cfin_item(const Phase& P, Ctx& cx, int item) {
  const int* iv = P.iv;
  const int M = iv[0], Cp = iv[1], HW = iv[3], ksplit = iv[4];
  const int nvec = Cp >> 3, rstep = MG_THREADS / nvec;
  const int cv = threadIdx.x % nvec, r0 = threadIdx.x / nvec;
  const int c0 = cv * 8;
  const int rpi = iv[8];
  const int row_lo = item * rpi, row_hi = min(row_lo + rpi, M);
  const int nlim = iv[9] ? min(iv[10], cx.nvalid) : iv[10];
  const int row_lim = nlim * HW;
  const float* SL = reinterpret_cast<const float*>(P.pv[3]);
  const size_t slice = (size_t)M * Cp;
  bf16* Y = reinterpret_cast<bf16*>(P.pv[0]);
  const bf16* RES = iv[5] ? reinterpret_cast<const bf16*>(P.pv[1]) : nullptr;
  const bool bnred = iv[6] != 0, has_stats = iv[7] != 0;
  NormCol nc;
  const bf16* BY = nullptr;
  if (bnred) {
    // BatchNorm-backward reductions of the layer whose input is pv[4]: out = dz * [bn(y) > 0];  T += (sum, sum * xhat)
    Phase const& Q = P;       // norm_setup reads S / gamma / beta from pv[2..4]: remap below
    (void)Q;
    BY = reinterpret_cast<const bf16*>(P.pv[4]);
    const float cnt = (float)nlim * (float)HW;
    const float* S = reinterpret_cast<const float*>(P.pv[5]);
    const bf16* gp = reinterpret_cast<const bf16*>(P.pv[6]);
    const bf16* bp = reinterpret_cast<const bf16*>(P.pv[7]);
    float sm[8], sq[8];
    ld8f(S + c0, sm);
    ld8f(S + Cp + c0, sq);
    const float ic = cnt > 0.f ? 1.f / cnt : 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const bool real = (c0 + e) < iv[2] && cnt > 0.f;
      const float ga = real ? ldbf(gp + c0 + e) : 0.f, be = real ? ldbf(bp + c0 + e) : 0.f;
      const float mean = sm[e] * ic;
      const float inv = real ? rsqrtf(fmaxf(sq[e] * ic - mean * mean, 0.f) + P.fv[0]) : 0.f;
      nc.mean[e] = mean; nc.inv[e] = inv; nc.a[e] = ga * inv; nc.b[e] = be - mean * nc.a[e];
    }
  }
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
  for (int row = row_lo + r0; row < row_hi; row += rstep) {
    if (row >= row_lim) continue;
    const size_t off = (size_t)row * Cp + c0;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    for (int s0 = 0; s0 < ksplit; s0 += 6) {            // up to 12 independent 16-byte loads in flight
      float4 a[12];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const bool in = s0 + j < ksplit;
        const float4* sp = reinterpret_cast<const float4*>(SL + (size_t)(s0 + j) * slice + off);
        a[2 * j] = in ? __ldcg(sp) : make_float4(0.f, 0.f, 0.f, 0.f);
        a[2 * j + 1] = in ? __ldcg(sp + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        v[0] += a[2 * j].x; v[1] += a[2 * j].y; v[2] += a[2 * j].z; v[3] += a[2 * j].w;
        v[4] += a[2 * j + 1].x; v[5] += a[2 * j + 1].y; v[6] += a[2 * j + 1].z; v[7] += a[2 * j + 1].w;
      }
    }
    if (RES != nullptr) {
      float r8[8];
      ld8(RES + off, r8);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += r8[e];
    }
    if (bnred) {
      float y8[8];
      ld8(BY + off, y8);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(nc.a[e], y8[e], nc.b[e]) > 0.f ? v[e] : 0.f;
      const uint4 u = pack8(v);
      *reinterpret_cast<uint4*>(Y + off) = u;
      unpack8(u, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s1[e] += v[e]; s2[e] = fmaf(v[e], (y8[e] - nc.mean[e]) * nc.inv[e], s2[e]); }
    } else {
      const uint4 u = pack8(v);
      *reinterpret_cast<uint4*>(Y + off) = u;
      unpack8(u, v);                                    // statistics of the bf16-rounded values
#pragma unroll
      for (int e = 0; e < 8; ++e) { s1[e] += v[e]; s2[e] = fmaf(v[e], v[e], s2[e]); }
    }
  }
  if (!(bnred || has_stats)) return;                    // uniform over the CTA
  float* red = reinterpret_cast<float*>(cx.smem);       // [256][16]
#pragma unroll
  for (int e = 0; e < 8; ++e) { red[threadIdx.x * 16 + e] = s1[e]; red[threadIdx.x * 16 + 8 + e] = s2[e]; }
  __syncthreads();
  for (int half = rstep >> 1; half >= 1; half >>= 1) {
    if (r0 < half) {
      float* mine = red + threadIdx.x * 16;
      const float* other = red + (threadIdx.x + half * nvec) * 16;
#pragma unroll
      for (int e = 0; e < 16; ++e) mine[e] += other[e];
    }
    __syncthreads();
  }
  if (r0 == 0) {
    float* dst = reinterpret_cast<float*>(bnred ? P.pv[8] : P.pv[2]);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red_add(dst + c0 + e, red[threadIdx.x * 16 + e]);
      red_add(dst + Cp + c0 + e, red[threadIdx.x * 16 + 8 + e]);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------ OP_HEAD
// iv: 0 N  1 HW  2 Cp  3 C_real  4 Kc  5 train  6 step-strided logits (eval)  9 use_nvalid
// pv: 0 z  1 W (bf16 [Kc][C_real])  2 bias  3 labels (int32)  4 mask (u8)  5 meter (double[4])  6 dW  7 db  8 dz out
//     9 logits out (fp32 [N][Kc], eval)  10 loss out (float)
__device__ __noinline__ void head_train(const Phase& P, Ctx& cx) {
  const int* iv = P.iv;
  const int N = iv[9] ? min(iv[0], cx.nvalid) : iv[0], Nal = iv[0], HW = iv[1], Cp = iv[2], C = iv[3], Kc = iv[4];
  const int tid = threadIdx.x, nt = MG_THREADS, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  // everything the head touches more than once is staged in shared memory with 16-byte loads: the pooled activations,
  // the classifier weights, the logits / dlogits
  bf16* Ws = reinterpret_cast<bf16*>(cx.smem);              // [Kc][Cp] classifier weights (16-byte aligned rows)
  float* pooled = reinterpret_cast<float*>(cx.smem + (((size_t)Kc * Cp * 2 + 15) & ~(size_t)15));   // [Nal][Cp]
  float* logit = pooled + Nal * Cp;                         // [Nal][Kc]  (becomes dlogits)
  float* red = logit + Nal * Kc;                            // [2*Nal]
  float* biasS = red + 2 * Nal;                             // [Kc]
  int* maskS = reinterpret_cast<int*>(biasS + Kc);          // [Kc]
  int* labS = maskS + Kc;                                   // [Nal]
  const bf16* Z = reinterpret_cast<const bf16*>(P.pv[0]);
  const bf16* W = reinterpret_cast<const bf16*>(P.pv[1]);
  const bf16* Bv = reinterpret_cast<const bf16*>(P.pv[2]);
  const int* lab = reinterpret_cast<const int*>(P.pv[3]);
  const unsigned char* mask = reinterpret_cast<const unsigned char*>(P.pv[4]);
  const float inv_hw = 1.f / (float)HW;
  const int nvec = Cp >> 3;
  for (int k = tid; k < Kc; k += nt) {
    biasS[k] = ldbf(Bv + k);
    maskS[k] = (mask == nullptr || mask[k]) ? 1 : 0;
  }
  for (int n = tid; n < Nal; n += nt) labS[n] = n < N ? __ldcg(lab + n) : 0;
  // classifier weights -> shared memory (16-byte loads when the rows allow it)
  if ((C & 7) == 0) {
    const int wv = C >> 3;
    for (int i = tid; i < Kc * wv; i += nt) {
      const int k = i / wv, v = i - k * wv;
      *reinterpret_cast<uint4*>(Ws + (size_t)k * Cp + v * 8) = __ldcg(reinterpret_cast<const uint4*>(W + (size_t)k * C) + v);
    }
  } else {
    for (int i = tid; i < Kc * Cp; i += nt) {
      const int k = i / Cp, c = i - k * Cp;
      Ws[i] = c < C ? __ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(W + (size_t)k * C + c)))
                    : __float2bfloat16(0.f);
    }
  }
  for (int i = tid; i < Nal * nvec; i += nt) {
    const int n = i / nvec, c0 = (i - n * nvec) * 8;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    if (n < N) {
      const bf16* pz = Z + ((size_t)n * HW) * Cp + c0;
      for (int h0 = 0; h0 < HW; h0 += 16) {
        uint4 u[16];
#pragma unroll
        for (int h = 0; h < 16; ++h)
          u[h] = (h0 + h < HW) ? __ldcg(reinterpret_cast<const uint4*>(pz + (size_t)(h0 + h) * Cp)) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          float v[8];
          unpack8(u[h], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += v[e];
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) pooled[n * Cp + c0 + e] = bfr(acc[e] * inv_hw);
  }
  __syncthreads();
  for (int o = warp; o < N * Kc; o += nw) {
    const int n = o / Kc, k = o - n * Kc;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += pooled[n * Cp + c] * __bfloat162float(Ws[k * Cp + c]);
    s = warp_sum(s);
    if (lane == 0) {
      float z = bfr(s + biasS[k]);
      if (!maskS[k]) z = 0.f;
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
    const int tg = labS[n];
    if (lane == 0) { red[n] = lse - logit[n * Kc + tg]; red[Nal + n] = (cand == tg) ? 1.f : 0.f; }
    __syncwarp();
    for (int k = lane; k < Kc; k += 32) {
      const float g = maskS[k] ? (__expf(logit[n * Kc + k] - lse) - (k == tg ? 1.f : 0.f)) / (float)N : 0.f;
      logit[n * Kc + k] = bfr(g);
    }
  }
  __syncthreads();
  if (tid < 4 && N > 0) {     // the 4 meter words are updated by 4 threads (only this CTA touches them in this phase)
    float a = 0.f, c = 0.f;
    for (int n = 0; n < N; ++n) { a += red[n]; c += red[Nal + n]; }
    const float loss = a / (float)N;
    if (tid == 0 && P.pv[10]) *reinterpret_cast<float*>(P.pv[10]) = loss;
    double* meter = reinterpret_cast<double*>(P.pv[5]);
    if (meter != nullptr) {
      const double add = tid == 0 ? (double)loss * (double)N : tid == 1 ? (double)c : tid == 2 ? (double)N
                                                                                        : (double)__expf(loss) * (double)N;
      meter[tid] += add;
    }
  }
  float* dW = reinterpret_cast<float*>(P.pv[6]);
  float* db = reinterpret_cast<float*>(P.pv[7]);
  for (int k = tid; k < Kc; k += nt) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += logit[n * Kc + k];
    db[k] = s;
  }
  for (int i = tid; i < Kc * C; i += nt) {
    const int k = i / C, c = i - k * C;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += logit[n * Kc + k] * pooled[n * Cp + c];
    dW[i] = s;
  }
  // dpooled -> dz (broadcast over the HW positions), 8 channels per thread, 16-byte stores
  bf16* DZ = reinterpret_cast<bf16*>(P.pv[8]);
  for (int i = tid; i < Nal * nvec; i += nt) {
    const int n = i / nvec, c0 = (i - n * nvec) * 8;
    float o8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float sacc = 0.f;
      if (n < N && c0 + e < C) {
        for (int k = 0; k < Kc; ++k) sacc += logit[n * Kc + k] * __bfloat162float(Ws[k * Cp + c0 + e]);
        sacc = bfr(sacc) * inv_hw;
      }
      o8[e] = sacc;
    }
    const uint4 u = pack8(o8);
    bf16* pd = DZ + ((size_t)n * HW) * Cp + c0;
    for (int h = 0; h < HW; ++h) *reinterpret_cast<uint4*>(pd + (size_t)h * Cp) = u;
  }
  __syncthreads();
}

// eval: one work item = one image; logits[n][k] (fp32)
__device__ __noinline__ void head_eval_item(const Phase& P, Ctx& cx, int n) {
  const int* iv = P.iv;
  const int HW = iv[1], Cp = iv[2], C = iv[3], Kc = iv[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = MG_THREADS >> 5;
  float* pooled = reinterpret_cast<float*>(cx.smem);
  const bf16* Z = reinterpret_cast<const bf16*>(P.pv[0]);
  const bf16* W = reinterpret_cast<const bf16*>(P.pv[1]);
  const bf16* Bv = reinterpret_cast<const bf16*>(P.pv[2]);
  float* out = reinterpret_cast<float*>(P.pv[9]) + (iv[6] ? (size_t)cx.step * iv[0] * Kc : (size_t)0);
  for (int c = tid; c < C; c += MG_THREADS) {
    const bf16* p = Z + ((size_t)n * HW) * Cp + c;
    float s = 0.f;
    for (int h = 0; h < HW; ++h) s += __bfloat162float(p[(size_t)h * Cp]);
    pooled[c] = bfr(s / (float)HW);
  }
  __syncthreads();
  for (int k = warp; k < Kc; k += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += pooled[c] * __bfloat162float(W[(size_t)k * C + c]);
    s = warp_sum(s);
    if (lane == 0) out[(size_t)n * Kc + k] = bfr(s + __bfloat162float(Bv[k]));
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------- optimiser
// OP_SGD_NORM   iv: 0 n4;  pv: 0 g
__device__ __noinline__ void sgd_norm_item(const Phase& P, Ctx& cx, int item) {
  __shared__ float sm[33];
  const int n4 = P.iv[0];
  const float4* g = reinterpret_cast<const float4*>(P.pv[0]);
  float acc = 0.f;
  const int lo = item * (MG_THREADS * 8);
  float4 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = lo + k * MG_THREADS + threadIdx.x;
    v[k] = (i < n4) ? __ldcg(g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) red_add(reinterpret_cast<float*>(cx.jctl + CTL_GNORM2), acc);
}

// OP_SGD_UPDATE  iv: 0 n4  1 upd_items  2 zero_n4;  pv: 0 w  1 wc  2 g  3 mom  4 zero_ptr;  fv: 0 lr 1 mu 2 wd 3 max_norm
__device__ __noinline__ void sgd_update_item(const Phase& P, Ctx& cx, int item) {
  const int upd_items = P.iv[1];
  if (item >= upd_items) {
    float4* z = reinterpret_cast<float4*>(P.pv[4]);
    const int zn4 = P.iv[2];
    const int lo = (item - upd_items) * (MG_THREADS * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = lo + k * MG_THREADS + threadIdx.x;
      if (i < zn4) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const int n4 = P.iv[0];
  const float total = sqrtf(__ldcg(reinterpret_cast<const float*>(cx.jctl + CTL_GNORM2)));
  const float coef = fminf(1.f, P.fv[3] / (total + 1e-6f));
  const float lr = P.fv[0], mu = P.fv[1], wd = P.fv[2];
  float4* w = reinterpret_cast<float4*>(P.pv[0]);
  uint2* wc = reinterpret_cast<uint2*>(P.pv[1]);
  float4* g = reinterpret_cast<float4*>(P.pv[2]);
  float4* mom = reinterpret_cast<float4*>(P.pv[3]);
  const int lo = item * (MG_THREADS * 8);
#pragma unroll
  for (int kk = 0; kk < 8; kk += 4) {
    float4 gv[4], wv[4], mv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {       // 12 independent 16-byte loads in flight per thread
      const int i = lo + (kk + k) * MG_THREADS + threadIdx.x;
      if (i < n4) { gv[k] = __ldcg(g + i); wv[k] = __ldcg(w + i); mv[k] = __ldcg(mom + i); }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = lo + (kk + k) * MG_THREADS + threadIdx.x;
      if (i < n4) {
        const float* gp = reinterpret_cast<const float*>(&gv[k]);
        float* wp = reinterpret_cast<float*>(&wv[k]);
        float* mp = reinterpret_cast<float*>(&mv[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float gg = gp[e] * coef + wd * wp[e];
          mp[e] = mu * mp[e] + gg;
          wp[e] -= lr * mp[e];
        }
        w[i] = wv[k];
        mom[i] = mv[k];
        const __nv_bfloat162 a = __floats2bfloat162_rn(wp[0], wp[1]), b = __floats2bfloat162_rn(wp[2], wp[3]);
        uint2 tt;
        tt.x = *reinterpret_cast<const unsigned*>(&a);
        tt.y = *reinterpret_cast<const unsigned*>(&b);
        wc[i] = tt;
        g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  if (item == 0 && threadIdx.x == 0) reinterpret_cast<float*>(cx.jctl)[CTL_LAST_GNORM] = total;
}

// ------------------------------------------------------------------------------------------------------------ OP_PREP
// iv: 0 B  1 H  2 W  3 C_src  4 pad  5 npack  6 use_params  7 Cx (channel stride of x, 8)
// pv: 0 src u8 [n][H][W][C_src]  1 idx_tab (int64 [steps][B])  2 par_tab (int32 [steps][B][3])  3 mean  4 inv_std
//     5 x out  6 labels src (int64)  7 labels out (int32)  8 pack table (int64 [npack][5])  9 w (fp32)  10 packbuf
__device__ __noinline__ void prep_item(const Phase& P, Ctx& cx, int item) {
  const int* iv = P.iv;
  const int B = iv[0];
  if (item == 0 && threadIdx.x == 0) reinterpret_cast<float*>(cx.jctl)[CTL_GNORM2] = 0.f;
  if (item < B) {
    const int H = iv[1], W = iv[2], Cs = iv[3], pad = iv[4], Cx = iv[7];
    const unsigned char* src = reinterpret_cast<const unsigned char*>(P.pv[0]);
    const long long* idx = reinterpret_cast<const long long*>(P.pv[1]);
    const int* par = iv[6] ? reinterpret_cast<const int*>(P.pv[2]) : nullptr;
    const float* mean = reinterpret_cast<const float*>(P.pv[3]);
    const float* istd = reinterpret_cast<const float*>(P.pv[4]);
    bf16* X = reinterpret_cast<bf16*>(P.pv[5]);
    const size_t slot = (size_t)cx.step * B + item;
    const bool valid = item < cx.nvalid;
    const long long n = valid ? idx[slot] : 0;
    int dy = pad, dx = pad, flip = 0;
    if (par != nullptr && valid) { dy = par[slot * 3 + 0]; dx = par[slot * 3 + 1]; flip = par[slot * 3 + 2]; }
    if (threadIdx.x == 0 && P.pv[7]) {
      const long long* ls = reinterpret_cast<const long long*>(P.pv[6]);
      reinterpret_cast<int*>(P.pv[7])[item] = valid ? (int)ls[n] : 0;
    }
    for (int p = threadIdx.x; p < H * W; p += MG_THREADS) {
      const int y = p / W, x = p - y * W;
      const int sy = y + dy - pad, sx = (flip ? (W - 1 - x) : x) + dx - pad;
      const bool in = valid && sy >= 0 && sy < H && sx >= 0 && sx < W;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = 0.f;
      for (int c = 0; c < Cs && c < 8; ++c) {
        float v = 0.f;
        if (in) v = (float)src[(((size_t)n * H + sy) * W + sx) * Cs + c] * (1.f / 255.f);
        f[c] = valid ? (v - mean[c]) * istd[c] : 0.f;
      }
      st8(X + ((size_t)item * H * W + p) * Cx, f);      // Cx == 8: one 16-byte pixel
    }
  } else {
    const long long* tab = reinterpret_cast<const long long*>(P.pv[8]) + (size_t)(item - B) * 5;
    const float* w = reinterpret_cast<const float*>(P.pv[9]) + tab[0];
    bf16* dst = reinterpret_cast<bf16*>(P.pv[10]) + tab[1];
    const int rows = (int)tab[2], C = (int)tab[3], Cpd = (int)tab[4];
    for (int i = threadIdx.x; i < rows * Cpd; i += MG_THREADS) {
      const int r = i / Cpd, c = i - r * Cpd;
      dst[i] = __float2bfloat16(c < C ? __ldcg(w + (size_t)r * C + c) : 0.f);
    }
  }
}

// Headers of phase `pn` for every job -> shared memory (warps 2-3: 64 threads; 24 uint4 per header; every load of a
// thread is in flight before its first store, i.e. one L2 round trip per phase).
__device__ __noinline__ void prefetch_headers(const Job* sJob, int njobs, int pn, uint32_t* dst /* [MAX_JOBS][HDR_WORDS] */) {
  constexpr int Q = HDR_WORDS / 4, R = (MAX_JOBS * Q + 63) / 64;
  const int i0 = threadIdx.x - 64, nvq = njobs * Q;
  uint4 hv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = i0 + r * 64;
    if (i < nvq) {
      const int k = i / Q, q4 = i - k * Q;
      hv[r] = __ldcg(reinterpret_cast<const uint4*>(sJob[k].prog + pn) + q4);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = i0 + r * 64;
    if (i < nvq) {
      const int k = i / Q, q4 = i - k * Q;
      reinterpret_cast<uint4*>(dst + k * HDR_WORDS)[q4] = hv[r];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(MG_THREADS, 1)
mega_kernel(const Job* __restrict__ jobs, int njobs, int nphase, int nsteps, unsigned* __restrict__ ctl,
            unsigned long long* __restrict__ prof) {
  extern __shared__ uint8_t smem_raw[];
  // headers of the current / next phase for every job (tensor maps stay in global memory), double buffered
  __shared__ __align__(16) uint32_t sHdr[2][MAX_JOBS][HDR_WORDS];
  __shared__ Job sJob[MAX_JOBS];
  __shared__ int sPre[MAX_JOBS + 1];
  __shared__ int sNv[MAX_JOBS];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_flag;
  __shared__ float s_bn[4 * 128];
  Ctx cx;
  cx.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  cx.wg_stage = cx.smem + MG_STAGES * MG_STAGE_BYTES;
  cx.full = reinterpret_cast<uint64_t*>(cx.wg_stage + MG_WG_STAGING);
  cx.empty = cx.full + MG_STAGES;
  cx.tfull = cx.empty + MG_STAGES;
  cx.tempty = cx.tfull + 2;
  cx.ctl = ctl;
  cx.kctr = 0;
  cx.tctr = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (njobs > MAX_JOBS) njobs = MAX_JOBS;
  if (threadIdx.x < njobs) sJob[threadIdx.x] = jobs[threadIdx.x];
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MG_STAGES; ++s) { mbar_init(&cx.full[s], 1); mbar_init(&cx.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&cx.tfull[a], 1); mbar_init(&cx.tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, MG_TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cx.tmem = tmem_slot;
  // headers of phase 0
  for (int i = threadIdx.x; i < njobs * HDR_WORDS; i += MG_THREADS) {
    const int k = i / HDR_WORDS, wd = i - k * HDR_WORDS;
    sHdr[0][k][wd] = __ldcg(reinterpret_cast<const uint32_t*>(sJob[k].prog) + wd);
  }
  __syncthreads();

  unsigned epoch = 0;
  unsigned rot = 0;
  const unsigned G = gridDim.x;
  bool dead = false, synced = true;
  int buf = 0;
  const bool do_prof = (prof != nullptr) && blockIdx.x == 0 && threadIdx.x == 0;
  for (int step = 0; step < nsteps && !dead; ++step) {
    cx.step = step;
    if (threadIdx.x < njobs) {
      const Job& jb = sJob[threadIdx.x];
      sNv[threadIdx.x] = (jb.nvalid_tab != nullptr && step < jb.nsteps) ? __ldcg(jb.nvalid_tab + step) : 0x7fffffff;
    }
    __syncthreads();
    for (int p = 0; p < nphase; ++p) {
      cx.phase = p;
      // prefetch the next phase's headers into the other buffer (its last readers passed the sync that ended the
      // previous phase); consumed after the sync at the end of this iteration
      {
        const int pn = (p + 1 < nphase) ? p + 1 : 0;
        if (warp >= 2 && warp < 4) prefetch_headers(sJob, njobs, pn, &sHdr[buf ^ 1][0][0]);
      }
      // tile space of this phase: the active jobs' tiles back to back (jobs are sorted heaviest first on the host)
      if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < njobs; ++k) {
          sPre[k] = acc;
          if (step < sJob[k].nsteps) acc += reinterpret_cast<const Phase*>(sHdr[buf][k])->ntiles;
        }
        sPre[njobs] = acc;
      }
      __syncthreads();
      long long t0 = 0;
      if (do_prof) t0 = clock64();
      const int nt = sPre[njobs];
      const unsigned start = (blockIdx.x + G - (rot % G)) % G;
      const int op0 = reinterpret_cast<const Phase*>(sHdr[buf][0])->op;
      if (op0 == OP_CONV || op0 == OP_WGRAD) {
        // tensor-core phases: every role walks the CTA's tile sequence on its own (software pipeline across tiles)
        const bool role = (warp == 0 && lane == 0) || warp == 1 || warp >= 4;
        if (role) {
          int k = 0;
          bool live = true;
          for (int t = (int)start; t < nt && live; t += (int)G) {
            while (t >= sPre[k + 1]) ++k;
            const Phase& sP = *reinterpret_cast<const Phase*>(sHdr[buf][k]);
            const Phase* Pg = sJob[k].prog + p;
            const int lt = t - sPre[k];
            const bool conv = sP.op == OP_CONV;
            if (warp == 0) {
              live = conv ? conv_produce(Pg, sP, cx, lt) : wgrad_produce(Pg, sP, cx, lt);
            } else if (warp == 1) {
              live = conv ? conv_mma(sP, cx, lt) : wgrad_mma(sP, cx, lt);
            } else {
              cx.jctl = sJob[k].jctl;
              cx.nvalid = sNv[k];
              live = conv ? conv_epilogue(sP, cx, lt, &s_flag, s_bn) : wgrad_epilogue(sP, cx, lt);
            }
          }
        }
      } else {
        int k = 0;
        for (int t = (int)start; t < nt; t += (int)G) {
          while (t >= sPre[k + 1]) ++k;
          const Phase& sP = *reinterpret_cast<const Phase*>(sHdr[buf][k]);
          const int lt = t - sPre[k];
          cx.jctl = sJob[k].jctl;
          cx.nvalid = sNv[k];
          switch (sP.op) {
            case OP_PREP: prep_item(sP, cx, lt); break;
            case OP_NAPPLY: napply_item(sP, cx, lt); break;
            case OP_CFIN: cfin_item(sP, cx, lt); break;
            case OP_NBWD_RED: nbwd_item(sP, cx, lt, false); break;
            case OP_NBWD_APPLY: nbwd_item(sP, cx, lt, true); break;
            case OP_HEAD:
              if (sP.iv[5]) head_train(sP, cx); else head_eval_item(sP, cx, lt);
              break;
            case OP_SGD_NORM: sgd_norm_item(sP, cx, lt); break;
            case OP_SGD_UPDATE: sgd_update_item(sP, cx, lt); break;
            default: break;
          }
        }
      }
      rot = (rot + (unsigned)nt) % G;
      long long t1 = 0;
      if (do_prof) t1 = clock64();
      const int sync_after = reinterpret_cast<const Phase*>(sHdr[buf][0])->sync_after;
      // an empty phase right after a barrier has nothing to order: its own barrier is skipped (structural placeholders)
      if (sync_after && !(nt == 0 && synced)) {
        grid_sync(ctl, epoch);
        if (aborted(ctl)) { dead = true; break; }
        synced = true;
      } else {
        __syncthreads();
        if (nt != 0) synced = false;
      }
      if (do_prof) {
        const long long t2 = clock64();
        prof[2 * p] += (unsigned long long)(t1 - t0);
        prof[2 * p + 1] += (unsigned long long)(t2 - t1);
      }
      buf ^= 1;
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl[CTL_STEP] = (unsigned)nsteps;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(cx.tmem, MG_TMEM_COLS);
  }
}

}  // namespace mega
}  // namespace hfl

using namespace hfl::mega;

extern "C" int hfl_mega_phase_bytes() { return (int)sizeof(Phase); }
extern "C" int hfl_mega_smem_bytes() { return MG_SMEM; }

static int mega_configure() {
  static int configured = -1;
  if (configured < 0) {
    cudaError_t e = cudaFuncSetAttribute(mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
    // two CTAs per SM (2 x ~100 KB): ask for the largest shared-memory carve-out
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mega_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = (e == cudaSuccess) ? 0 : (int)e;
  }
  return configured;
}

// CTAs of the megakernel that can be co-resident on the current device (the cooperative-launch limit)
extern "C" int hfl_mega_max_grid() {
  if (mega_configure() != 0) return 0;
  int per_sm = 0, dev = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mega_kernel, MG_THREADS, MG_SMEM) != cudaSuccess) return 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return per_sm * sms;
}

// debugging aid: kernel attributes + occupancy for a few dynamic shared-memory sizes
extern "C" void hfl_mega_occupancy_probe(int* out /* [16] */) {
  mega_configure();
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, mega_kernel);
  out[0] = a.numRegs; out[1] = (int)a.sharedSizeBytes; out[2] = a.maxDynamicSharedSizeBytes; out[3] = a.maxThreadsPerBlock;
  out[4] = a.preferredShmemCarveout; out[5] = (int)a.localSizeBytes;
  const int sizes[5] = {0, 48 * 1024, 96 * 1024, 110 * 1024, MG_SMEM};
  for (int i = 0; i < 5; ++i) {
    int n = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mega_kernel, MG_THREADS, sizes[i]);
    out[6 + i] = n;
  }
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev); out[11] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxRegistersPerMultiprocessor, dev); out[12] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxBlocksPerMultiprocessor, dev); out[13] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrReservedSharedMemoryPerBlock, dev); out[14] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); out[15] = v;
}

// jobs: device array of `Job` records (all sharing one phase structure); ctl: launch-wide control block (zeroed here).
extern "C" int hfl_mega_launch(const void* jobs, int njobs, int nphase, int nsteps, void* ctl, int grid, void* prof,
                               cudaStream_t stream) {
  int e0 = mega_configure();
  if (e0 != 0) return e0;
  if (njobs < 1 || njobs > MAX_JOBS) return -7;
  cudaError_t e = cudaMemsetAsync(ctl, 0, CTL_WORDS * sizeof(unsigned), stream);
  if (e != cudaSuccess) return (int)e;
  const Job* j = reinterpret_cast<const Job*>(jobs);
  unsigned* c = reinterpret_cast<unsigned*>(ctl);
  unsigned long long* pr = reinterpret_cast<unsigned long long*>(prof);
  void* args[] = {(void*)&j, (void*)&njobs, (void*)&nphase, (void*)&nsteps, (void*)&c, (void*)&pr};
  e = cudaLaunchCooperativeKernel((const void*)mega_kernel, dim3(grid), dim3(MG_THREADS), args, (size_t)MG_SMEM, stream);
  hfl::count_launch();
  return (int)e;
}

extern "C" int hfl_mega_max_jobs() { return MAX_JOBS; }

// Encode a bf16 tiled tensor map (rank 2..5, 128-byte swizzle, zero OOB fill) into `out` (128 bytes, host memory).
extern "C" int hfl_mega_encode_tmap(void* out, const void* base, int rank, const unsigned long long* dims,
                                    const unsigned long long* strides_bytes, const unsigned* box, const unsigned* estr) {
  CUtensorMap m;
  uint64_t d[5], s[4];
  uint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  int r = hfl::make_tmap_bf16_generic(&m, base, rank, d, s, b, e);
  if (r) return r;
  memcpy(out, &m, sizeof(CUtensorMap));
  return 0;
}
