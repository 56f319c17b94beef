// 3x3 / stride-1 implicit-GEMM convolution that loads every input pixel ONCE per channel block.
//
// EXPERIMENTAL -- gated behind HFL_CONV_HALO=1 (ops/fused.py); written against the measured diagnosis of
// conv_tcgen05.cu and not validated on hardware yet.
//
// conv_tcgen05.cu re-loads the shifted 128-pixel A tile by TMA for each of the nine filter taps: 9 x 16 KB per
// 64-channel block and CTA, which makes the kernel L2 -> SM bandwidth bound at 25-45 % of the tensor peak on the
// large-batch (sBN / test) shapes.  Here a CTA owns a 16-row x 8-pixel output tile of ONE image and loads the haloed
// input brick {64 ch, 16 w, 18 h} once (36 KB, zero-filled outside the image = the padding).  The brick is written by
// TMA with the 128-byte swizzle at a row pitch of 16 pixels, i.e. exactly two 1 024-byte swizzle atoms per image row.
// The A operand of tap (r, s) is then the SAME shared memory seen through a different UMMA descriptor:
//
//     start address = brick + ((r * 16 + s) * 128) B      row (h + r, w + s) of the brick = GEMM row (h, w)
//     SBO           = 2 048 B                             the 8-pixel groups of consecutive image rows
//     base_offset   = s                                   swizzle phase of a start that is not 1 024-byte aligned
//
// so the nine taps cost nine descriptor changes and no memory traffic.  Weights stream through their own 4-stage ring
// (one BLOCK_N x 64 slab per tap).  Roles: warp 0 = A-brick TMA, warp 3 = weight TMA, warp 1 = MMA issue,
// warp 2 = TMEM allocation, warps 4-7 = epilogue (residual add / bf16 store / BN statistics, compile-time variants).
//
// Limits of this first version: R = S = 3, stride 1, C and K_out multiples of 64, H % 16 == 0, W % 8 == 0.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace hfl {

constexpr int CH_TILE_H = 16, CH_TILE_W = 8;                 // 128 output pixels
constexpr int CH_BRICK_H = CH_TILE_H + 2, CH_BRICK_W = 16;   // haloed rows; row pitch padded to two swizzle atoms
constexpr int CH_A_BYTES = CH_BRICK_H * CH_BRICK_W * 128;    // 36 864 B per 64-channel block
constexpr int CH_A_STAGES = 2;
constexpr int CH_B_STAGES = 4;
constexpr int CH_THREADS = 256;

template <int BLOCK_N>
struct HaloSmem {
  static constexpr int B_BYTES = BLOCK_N * 128;              // BLOCK_N rows x 64 channels
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = CH_A_STAGES * CH_A_BYTES;     // 73 728: 1 024-aligned
  static constexpr int BAR_OFF = B_OFF + CH_B_STAGES * B_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

struct HaloGeom {
  int N, H, W, C, Kout;
  int tiles_w, tiles_h;        // W / 8, H / 16
};

// K-major SW128 descriptor with an explicit group stride and swizzle phase (see file header)
__device__ __forceinline__ uint64_t make_smem_desc_sw128_ex(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address
  d |= (uint64_t)1 << 16;                                // LBO (unused for swizzled K-major)
  d |= (uint64_t)(sbo_bytes >> 4) << 32;                 // stride between 8-row groups
  d |= (uint64_t)1 << 46;                                // descriptor version
  d |= (uint64_t)((smem_addr >> 7) & 7u) << 49;          // matrix base offset: phase inside the 1 024-byte pattern
  d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
  return d;
}

template <int BLOCK_N, bool HAS_RES, bool HAS_STATS>
__global__ void __launch_bounds__(CH_THREADS, 1)
conv_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         __nv_bfloat16* __restrict__ Y, const __nv_bfloat16* __restrict__ residual,
                         float* __restrict__ stats, const HaloGeom g) {
  extern __shared__ uint8_t smem_raw[];
  using S = HaloSmem<BLOCK_N>;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* a_empty = a_full + CH_A_STAGES;
  uint64_t* b_full = a_empty + CH_A_STAGES;
  uint64_t* b_empty = b_full + CH_B_STAGES;
  uint64_t* tmem_full_bar = b_empty + CH_B_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = g.tiles_w * g.tiles_h;
  const int n = blockIdx.x / tiles_per_img;
  const int t_in = blockIdx.x - n * tiles_per_img;
  const int th = t_in / g.tiles_w, tw = t_in - th * g.tiles_w;
  const int h0 = th * CH_TILE_H, w0 = tw * CH_TILE_W;
  const int n_blk = blockIdx.y;
  const int cblocks = g.C / 64;
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_x); tma_prefetch_desc(&tmap_w); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CH_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < CH_B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== A producer: one haloed brick per 64-channel block =====
    if (lane == 0) {
      for (int cb = 0; cb < cblocks; ++cb) {
        const int st = cb % CH_A_STAGES;
        const uint32_t ph = (cb / CH_A_STAGES) & 1;
        mbar_wait(&a_empty[st], ph ^ 1);
        mbar_expect_tx(&a_full[st], CH_A_BYTES);
        // box {64 ch, 16 w, 18 h, 1 n} at (c, w, h, n) = (cb*64, w0 - 1, h0 - 1, n); outside the image -> zeros
        tma_load_4d(smem + S::A_OFF + st * CH_A_BYTES, &tmap_x, &a_full[st], cb * 64, w0 - 1, h0 - 1, n);
      }
    }
  } else if (warp == 3) {
    // ===== B producer: one BLOCK_N x 64 weight slab per (channel block, tap) =====
    if (lane == 0) {
      const int total = cblocks * 9;
      for (int it = 0; it < total; ++it) {
        const int st = it % CH_B_STAGES;
        const uint32_t ph = (it / CH_B_STAGES) & 1;
        const int cb = it / 9, tap = it - cb * 9;
        mbar_wait(&b_empty[st], ph ^ 1);
        mbar_expect_tx(&b_full[st], S::B_BYTES);
        tma_load_2d(smem + S::B_OFF + st * S::B_BYTES, &tmap_w, &b_full[st], tap * g.C + cb * 64, n_blk * BLOCK_N);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_bf16(128, BLOCK_N);
    int it = 0;
    for (int cb = 0; cb < cblocks; ++cb) {
      const int sa = cb % CH_A_STAGES;
      const uint32_t pha = (cb / CH_A_STAGES) & 1;
      mbar_wait(&a_full[sa], pha);
      tcgen05_fence_after();
      const uint32_t brick = smem_u32(smem + S::A_OFF + sa * CH_A_BYTES);
      for (int tap = 0; tap < 9; ++tap, ++it) {
        const int sb = it % CH_B_STAGES;
        const uint32_t phb = (it / CH_B_STAGES) & 1;
        mbar_wait(&b_full[sb], phb);
        tcgen05_fence_after();
        if (lane == 0) {
          const int r = tap / 3, s = tap - r * 3;
          const uint64_t adesc = make_smem_desc_sw128_ex(brick + (uint32_t)((r * CH_BRICK_W + s) * 128), 2048);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + S::B_OFF + sb * S::B_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)      // 64 channels = 4 x K16: +32 B inside the 128-byte row
            umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
          umma_commit(&b_empty[sb]);
          if (tap == 8) umma_commit(&a_empty[sa]);
          if (cb == cblocks - 1 && tap == 8) umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: accumulator row m = (h_l, w_l) = (m / 8, m % 8) =====
    const int q = warp - 4;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int m = q * 32 + lane;
    const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
    const size_t pix = ((size_t)n * g.H + oh) * g.W + ow;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      const int col0 = n_blk * BLOCK_N + c0;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      __nv_bfloat16* dst = Y + pix * g.Kout + col0;
      if (HAS_RES) {
        const uint4* rp = reinterpret_cast<const uint4*>(residual + pix * g.Kout + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = rp[j];
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) { f[8 * j + 2 * e] += __low2float(h[e]); f[8 * j + 2 * e + 1] += __high2float(h[e]); }
        }
      }
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
        h[0] = __floats2bfloat162_rn(f[j], f[j + 1]);
        h[1] = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
        h[2] = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
        h[3] = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
        *reinterpret_cast<uint4*>(dst + j) = u;
        if (HAS_STATS) {
          f[j] = __low2float(h[0]); f[j + 1] = __high2float(h[0]); f[j + 2] = __low2float(h[1]); f[j + 3] = __high2float(h[1]);
          f[j + 4] = __low2float(h[2]); f[j + 5] = __high2float(h[2]); f[j + 6] = __low2float(h[3]); f[j + 7] = __high2float(h[3]);
        }
      }
      if (HAS_STATS) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float sx = f[j], sq = f[j] * f[j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
          }
          if (lane == j) {
            atomicAdd(&stats[col0 + j], sx);
            atomicAdd(&stats[g.Kout + col0 + j], sq);
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" int hfl_conv_halo_supported(int N, int H, int W, int C, int Kout, int R, int S, int stride) {
  if (R != 3 || S != 3 || stride != 1) return 1;
  if (C % 64 != 0 || Kout % 64 != 0) return 2;
  if (H % CH_TILE_H != 0 || W % CH_TILE_W != 0) return 3;
  if (N <= 0) return 4;
  return 0;
}

template <int BLOCK_N, bool HAS_RES, bool HAS_STATS>
static int launch_halo(const CUtensorMap& tx, const CUtensorMap& tw, void* Y, const void* residual, float* stats,
                       const HaloGeom& g, cudaStream_t stream) {
  auto kern = conv_halo_tcgen05_kernel<BLOCK_N, HAS_RES, HAS_STATS>;
  constexpr int smem = HaloSmem<BLOCK_N>::TOTAL;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid(g.N * g.tiles_w * g.tiles_h, g.Kout / BLOCK_N);
  kern<<<grid, CH_THREADS, smem, stream>>>(tx, tw, (__nv_bfloat16*)Y, (const __nv_bfloat16*)residual, stats, g);
  count_launch();
  return (int)cudaGetLastError();
}

template <int BLOCK_N>
static int dispatch_halo(const CUtensorMap& tx, const CUtensorMap& tw, void* Y, const void* residual, float* stats,
                         const HaloGeom& g, cudaStream_t stream) {
  if (residual != nullptr) {
    if (stats != nullptr) return launch_halo<BLOCK_N, true, true>(tx, tw, Y, residual, stats, g, stream);
    return launch_halo<BLOCK_N, true, false>(tx, tw, Y, residual, stats, g, stream);
  }
  if (stats != nullptr) return launch_halo<BLOCK_N, false, true>(tx, tw, Y, residual, stats, g, stream);
  return launch_halo<BLOCK_N, false, false>(tx, tw, Y, residual, stats, g, stream);
}

// X: [N, H, W, C] bf16, Wt: [Kout, 3, 3, C] bf16, Y: [N, H, W, Kout] bf16; residual / stats optional
extern "C" int hfl_conv_halo_fprop(const void* X, const void* Wt, void* Y, const void* residual, float* stats, int N,
                                   int H, int W, int C, int Kout, cudaStream_t stream) {
  if (hfl_conv_halo_supported(N, H, W, C, Kout, 3, 3, 1) != 0) return -2;
  if ((reinterpret_cast<uintptr_t>(Y) & 15) != 0 || (residual && (reinterpret_cast<uintptr_t>(residual) & 15) != 0)) return -3;
  HaloGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.Kout = Kout;
  g.tiles_w = W / CH_TILE_W; g.tiles_h = H / CH_TILE_H;
  const int block_n = (Kout % 128 == 0) ? 128 : 64;
  CUtensorMap tx, tw;
  int r = make_tmap_nhwc_bf16(&tx, X, (uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N, 64, CH_BRICK_W, CH_BRICK_H, 1, 1, 1,
                              CU_TENSOR_MAP_SWIZZLE_128B);
  if (r) return 1000 + r;
  r = make_tmap_2d_bf16(&tw, Wt, (uint64_t)9 * C, (uint64_t)Kout, (uint64_t)9 * C * 2, 64, block_n);
  if (r) return 2000 + r;
  if (block_n == 128) return dispatch_halo<128>(tx, tw, Y, residual, stats, g, stream);
  return dispatch_halo<64>(tx, tw, Y, residual, stats, g, stream);
}
