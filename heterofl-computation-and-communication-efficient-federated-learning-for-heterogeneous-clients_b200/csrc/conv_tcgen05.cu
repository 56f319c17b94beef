// Implicit-GEMM convolution forward / data gradient / weight gradient on tcgen05 / TMEM, operands fed by TMA (sm_100a).
//
//   Y[n, oh, ow, k] = sum_{r,s,c} X[n, oh*stride + r - pad, ow*stride + s - pad, c] * Wt[k, r, s, c]   (+ residual)
//
// X, Y are NHWC bf16, Wt is K,R,S,C bf16 (the layout the parameter arenas store, plan/layout.py), fp32 accumulation.
// Reference ops: every nn.Conv2d of src/models/resnet.py:44-50,140-158 (and their autograd backward).
// GEMM view: M = N*OH*OW output pixels, N = K_out, K = R*S*C.  One CTA computes a 128-pixel x BLOCK_N tile:
//
//   * the 128 pixels are a (bn images) x (bh rows) x (OW columns) brick, so the A-operand slice for filter tap (r,s)
//     and channel block cb is ONE 4-D TMA box {64 ch, OW, bh, bn} of X starting at the tap's halo offset -- padding is
//     the TMA unit's out-of-bounds zero fill, stride-2 is the tensor map's traversal stride.  No im2col buffer, no
//     index arithmetic in the kernel: the producer warp only changes box coordinates.
//   * B-operand slices are 3-D boxes {64 c, 1 tap, BLOCK_N k} of the filter seen as {C, R*S, K_out}; both land in
//     128B-swizzled smem and are consumed by tcgen05.mma (M128 x BLOCK_N x K16, cta_group::1) issued by one elected
//     thread; a 4-stage mbarrier ring overlaps TMA with MMA; the accumulator lives in TMEM and is drained by 4 epilogue
//     warps with tcgen05.ld.
//   * channel counts: any multiple of 8 (16-byte global strides).  A partial 64-channel block -- the narrow HeteroFL
//     levels have 8..32 channels -- is TMA zero fill on both operands; BLOCK_N is 32 / 64 / 128 and the epilogue masks
//     the columns beyond K_out.
//   * epilogue (fprop): optional residual add, bf16 store, optional per-channel sum / sum-of-squares of the rounded
//     outputs (BatchNorm batch statistics of the NEXT norm: it becomes an apply-only pass, norm_kernels.cu
//     hfl_bn_apply_sums), optional device-side `nvalid` (padding samples of a partial batch: zeros, out of the sums).
//   * data gradient: the same kernel with mirrored taps and the filter read MN-major in place (no transposed / flipped
//     weight copy); a stride-2 gradient runs as four output-parity classes inside one launch (ConvGeom::taps).
//   * split-K (deep reductions, few tiles: the 256- / 512-channel layers of a batch-10 step): thread-block clusters
//     {1,1,ksplit}; every CTA reduces a k-range of the SAME tile into its own TMEM accumulator, stages it as fp32 in its
//     shared memory, and after a cluster barrier CTA z sums rows [z*128/ksplit, ...) of all peers through DSMEM and runs
//     the normal epilogue on them.  No atomics, no zero-filled scratch, no second kernel.
//   * every launch carries the programmatic-dependent-launch attribute; griddepcontrol.wait sits after the barrier /
//     TMEM / descriptor set-up (common.cuh).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace hfl {

constexpr int CV_BLOCK_M = 128;
constexpr int CV_BLOCK_K = 64;
constexpr int CV_THREADS = 256;

struct ConvGeom {
  int N, H, W, C, Kout, OH, OW, R, S, stride, pad;
  int bn, bh, bw;          // pixel brick of one tile (bn*bh*bw == 128, bw == OW)
  int tiles_h;             // ceil(OH / bh)
  // data gradient only: parity classes of a strided convolution.  The dX pixels (out_s*a + ph, out_s*b + pw) of class
  // cls = ph*2 + pw receive contributions from the filter taps r = (ph + pad) mod out_s (+ out_s ...) only, each a plain
  // shifted read of dY -- so every class is a small stride-1 problem over the dY grid with its own tap list, written
  // with stride out_s.  Stride 1: one class, all taps.  taps[cls][t] = filter tap | (dh + 8) << 8 | (dw + 8) << 16.
  int tiles_m_cls;         // M tiles per class (blockIdx.x = cls * tiles_m_cls + tile)
  int out_s, out_H, out_W; // output pixel stride / extent (dgrad: dX; fprop: out_s = 1, OH, OW)
  int ntaps[4];
  int taps[4][9];
  // forward only: a second operand pair accumulated into the same tile -- the 1x1 stride-`stride2` shortcut convolution
  // of a ResNet transition block (src/models/resnet.py:44-50: out = conv2(..) + shortcut(x)): kb2 extra k-blocks whose
  // A boxes come from tmap_x2 (the block input) and whose B boxes come from tmap_w2 (the shortcut filter)
  int kb2, stride2;
};

// Column sums over the 32 lanes of a warp: on return lane j holds sum_over_lanes(v[j]) in v[0]  (31 shuffles).
__device__ __forceinline__ void warp_col_reduce32(float (&v)[32], int lane) {
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
// thread-block-cluster helpers (split-K reduction through distributed shared memory)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int BLOCK_N>
struct ConvSmem {
  static constexpr int A_BYTES = CV_BLOCK_M * CV_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * CV_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // ring depth: 4 stages.  A 6-9 stage ring (-DHFL_CONV_DEEP_RING: the whole 9-18 k-block reduction of a batch-10 tile in
  // flight at once, ~190 KB) was measured on the same box (gpurun_out/r2s): isolated steps unchanged (921 vs 924 us level a,
  // 495 vs 498 us level c -- the TMA round trips are not what bounds these tiles), 4 concurrent slots 7 % slower (one
  // CTA per SM instead of two), round 476 vs 457 ms.
#ifdef HFL_CONV_DEEP_RING
  static constexpr int STAGES = BLOCK_N >= 128 ? 6 : (BLOCK_N >= 64 ? 8 : 9);
#else
  static constexpr int STAGES = 4;
#endif
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};

// DGRAD = true turns the same kernel into the stride-1 data-gradient:
//   dX[n,h,w,c] = sum_{r,s,k} dY[n, h+pad-r, w+pad-s, k] * Wt[k,r,s,c]
// "x" is then dY (reduction over its K_out channels, geometry fields C/Kout are swapped by the host), the filter taps
// are visited with mirrored halo offsets, and the B operand -- Wt[k, tap, c-range] with c contiguous -- is consumed
// MN-major straight from the K,R,S,C filter matrix (no transposed / flipped weight copy).
// HAS_RES / HAS_STATS are compile-time: the four epilogue warps run one per scheduler with nothing to hide latency
// behind, so code for an absent residual / statistics pass must not even be issued (predicated-off instructions of a
// runtime flag cost issue slots; the same lesson as the GEMM epilogue, csrc/gemm_tcgen05.cu).
template <int BLOCK_N, bool DGRAD, bool HAS_RES, bool HAS_STATS>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv_fprop_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                          const __grid_constant__ CUtensorMap tmap_x2, const __grid_constant__ CUtensorMap tmap_w2,
                          __nv_bfloat16* __restrict__ Y, const __nv_bfloat16* __restrict__ residual,
                          float* __restrict__ stats /* [2, Kout] or null */, const int* __restrict__ nvalid, const ConvGeom g) {
  extern __shared__ uint8_t smem_raw[];
  using S = ConvSmem<BLOCK_N>;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int CV_STAGES = S::STAGES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CV_STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + CV_STAGES;
  uint64_t* tmem_full_bar = empty_bar + CV_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cls = DGRAD ? blockIdx.x / g.tiles_m_cls : 0;      // parity class of a strided data gradient
  const int bx = DGRAD ? blockIdx.x - cls * g.tiles_m_cls : blockIdx.x;
  const int tile_n = bx / g.tiles_h;                   // image-group index
  const int tile_h = bx - tile_n * g.tiles_h;          // row-group index
  const int n0 = tile_n * g.bn, oh0 = tile_h * g.bh;
  const int n_blk = blockIdx.y;
  const int cblocks = (g.C + CV_BLOCK_K - 1) / CV_BLOCK_K;   // a partial last block is zero-filled by the TMA unit
  const int num_k_main = (DGRAD ? g.ntaps[cls] : g.R * g.S) * cblocks;  // 0: a class no filter tap reaches (zeros)
  const int num_k_all = num_k_main + (DGRAD ? 0 : g.kb2);               // + k-blocks of the folded shortcut convolution
  // split-K (gridDim.z > 1, launched as thread-block clusters {1, 1, ksplit}): the deep-K layers of a batch-10 step
  // have 2-8 output tiles and 36-72 k-blocks each; the CTAs of a cluster reduce disjoint k-ranges of ONE tile into
  // their own TMEM accumulators and combine them through distributed shared memory (no atomics, no zero-filled
  // scratch, no second kernel).
  const int ksplit = gridDim.z;
  const int k_per = (num_k_all + ksplit - 1) / ksplit;
  const int kb0 = blockIdx.z * k_per;
  const int num_k = max(min(k_per, num_k_all - kb0), 0);
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CV_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();   // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int st = kb % CV_STAGES;
        const uint32_t ph = (kb / CV_STAGES) & 1;
        const int kg = kb0 + kb;
        const int tap = kg / cblocks, cb = kg - tap * cblocks;
        const int r = tap / g.S, s = tap - r * g.S;
        mbar_wait(&empty_bar[st], ph ^ 1);
        uint8_t* sa = smem + st * S::STAGE_BYTES;
        uint8_t* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full_bar[st], S::STAGE_BYTES);
        if (!DGRAD && kg >= num_k_main) {
          // folded 1x1 shortcut: A = the block input at the strided output positions, B = the shortcut filter
          const int cb2 = kg - num_k_main;
          tma_load_4d(sa, &tmap_x2, &full_bar[st], cb2 * CV_BLOCK_K, 0, oh0 * g.stride2, n0);
          tma_load_3d(sb, &tmap_w2, &full_bar[st], cb2 * CV_BLOCK_K, 0, n_blk * BLOCK_N);
          continue;
        }
        if (!DGRAD) {
          // A: box {64ch, bw, bh, bn} at (c, w, h, n) = (cb*64, s - pad, oh0*stride + r - pad, n0); halo -> zero fill
          tma_load_4d(sa, &tmap_x, &full_bar[st], cb * CV_BLOCK_K, s - g.pad, oh0 * g.stride + r - g.pad, n0);
          // B: box {64 c, 1 tap, BLOCK_N k_out} of the filter seen as {C, R*S, K_out}; channels / rows beyond the
          // tensor are zero-filled, so any C and K_out that are multiples of 8 ride the same 128-byte-swizzle path
          tma_load_3d(sb, &tmap_w, &full_bar[st], cb * CV_BLOCK_K, tap, n_blk * BLOCK_N);
        } else {
          // A: dY box at the tap's (mirrored) offset; `tap` indexes this class's tap list
          const int tw = g.taps[cls][tap];
          const int ftap = tw & 0xff, dh = ((tw >> 8) & 0xff) - 8, dw = ((tw >> 16) & 0xff) - 8;
          tma_load_4d(sa, &tmap_x, &full_bar[st], cb * CV_BLOCK_K, dw, oh0 + dh, n0);
          // B (MN-major): 64(c) x 64(k_out) boxes of the {C, R*S, K_out} filter; here g.Kout is the number of dX
          // channels (= filter C) and g.C the reduction length (= filter K_out)
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_3d(sb + j * 8192, &tmap_w, &full_bar[st], n_blk * BLOCK_N + j * 64, ftap, cb * CV_BLOCK_K);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(CV_BLOCK_M, BLOCK_N, false, DGRAD);
    for (int kb = 0; kb < num_k; ++kb) {
      const int st = kb % CV_STAGES;
      const uint32_t ph = (kb / CV_STAGES) & 1;
      mbar_wait(&full_bar[st], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + st * S::STAGE_BYTES);
        const uint64_t adesc = make_smem_desc_sw128(a_addr);
        const uint64_t bdesc = DGRAD ? make_smem_desc_sw128_mn(a_addr + S::A_BYTES) : make_smem_desc_sw128(a_addr + S::A_BYTES);
        constexpr uint64_t b_step = DGRAD ? (2048 >> 4) : (32 >> 4);
#pragma unroll
        for (int k = 0; k < CV_BLOCK_K / 16; ++k)
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + b_step * k, idesc, (kb | k) != 0);
        umma_commit(&empty_bar[st]);
        if (kb == num_k - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && ksplit > 1) {
    // ---- split-K, step 1: this CTA's partial accumulator -> fp32 staging tile in its own shared memory (the stage
    // ring is free: the accumulator barrier fires after the last MMA has read it)
    const int q = warp - 4;
    if (num_k > 0) mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    float* stg = reinterpret_cast<float*>(smem) + (size_t)(q * 32 + lane) * (BLOCK_N + 4);
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      if (num_k > 0) {
        tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(stg + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    tcgen05_fence_before();
  } else if (warp >= 4) {
    const int q = warp - 4;
    if (num_k > 0) mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    // One accumulator row (= one output pixel, K_out contiguous bf16) per thread: 64-byte runs per thread and chunk.
    // (A shared-memory transposed write-out like the GEMM's was measured 5-10 % SLOWER here -- the fp32 staging traffic
    // costs more than the row-strided 16-byte stores -- so the direct form stays; profiles/conv_bench.md.)
    const int m = q * 32 + lane;                          // pixel within the brick: ((n_l*bh + h_l)*bw + w_l)
    const int w_l = m % g.bw;
    const int t = m / g.bw;
    const int h_l = t % g.bh, n_l = t / g.bh;
    const int n = n0 + n_l, oh = oh0 + h_l;
    // nvalid (training step with a padded partial batch): samples >= *nvalid are padding -- their outputs are written
    // as zeros (whatever the inputs / residual hold there) and stay out of the statistics
    const int nlim = nvalid ? min(g.N, __ldg(nvalid)) : g.N;
    const bool inr = (n < g.N) && (oh < g.OH);
    const bool ok = inr && (n < nlim);
    const size_t pix = ((size_t)n * g.out_H + (size_t)(oh * g.out_s + (cls >> 1))) * g.out_W + (size_t)(w_l * g.out_s + (cls & 1));
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      const int col0 = n_blk * BLOCK_N + c0;
      const int ncol = min(32, g.Kout - col0);              // produced channels in this chunk (multiple of 8)
      if (ncol <= 0) break;                                 // warp-uniform
      uint32_t v[32];
      if (num_k > 0) {
        tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = (ok && j < ncol) ? __uint_as_float(v[j]) : 0.f;
      if (inr) {
        __nv_bfloat16* dst = Y + pix * g.Kout + col0;
        if (HAS_RES && ok) {
          const uint4* rp = reinterpret_cast<const uint4*>(residual + pix * g.Kout + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (8 * j < ncol) {
              const uint4 u = rp[j];
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) { f[8 * j + 2 * e] += __low2float(h[e]); f[8 * j + 2 * e + 1] += __high2float(h[e]); }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (j < ncol) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
            h[0] = __floats2bfloat162_rn(f[j], f[j + 1]);
            h[1] = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
            h[2] = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
            h[3] = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
            *reinterpret_cast<uint4*>(dst + j) = u;
            if (HAS_STATS) {   // statistics of the bf16-rounded outputs (what a separate BN kernel would read)
              f[j] = __low2float(h[0]); f[j + 1] = __high2float(h[0]); f[j + 2] = __low2float(h[1]); f[j + 3] = __high2float(h[1]);
              f[j + 4] = __low2float(h[2]); f[j + 5] = __high2float(h[2]); f[j + 6] = __low2float(h[3]); f[j + 7] = __high2float(h[3]);
            }
          }
        }
      }
      if (HAS_STATS) {
        // column sums over the warp's 32 rows (rows outside the image / batch hold zeros): 2 x 31 shuffles, lane j
        // ends up owning column j
        float sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) sq[j] = f[j] * f[j];
        warp_col_reduce32(f, lane);
        warp_col_reduce32(sq, lane);
        if (lane < ncol) {
          red_add_f32(&stats[col0 + lane], f[0]);
          red_add_f32(&stats[g.Kout + col0 + lane], sq[0]);
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (ksplit > 1) {
    // ---- split-K, step 2: CTA z of the cluster sums rows [z*RPC, (z+1)*RPC) of all ksplit staging tiles through
    // DSMEM and runs the normal epilogue (residual, bf16 store, BatchNorm statistics) on them
    cluster_sync_all();
    if (warp >= 4) {
      constexpr int PITCH = BLOCK_N + 4;
      constexpr int CG = BLOCK_N / 4;                        // float4 column groups per row
      const int t = threadIdx.x - 128;
      const int cg = t % CG;                                  // this thread's 4 columns (fixed), rows strided
      const int rpc = CV_BLOCK_M / ksplit;
      const int row0 = blockIdx.z * rpc;
      const int col = n_blk * BLOCK_N + cg * 4;
      const bool col_ok = col < g.Kout;                       // K_out is a multiple of 8: the 4 columns share the answer
      const uint32_t stg0 = smem_u32(smem);
      const int nlim2 = nvalid ? min(g.N, __ldg(nvalid)) : g.N;
      float sx[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
      for (int rr = t / CG; rr < rpc; rr += 128 / CG) {
        const int m = row0 + rr;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t off = (uint32_t)(m * PITCH + cg * 4) * 4u;
        for (int z = 0; z < ksplit; ++z) {
          const float4 p = ld_dsmem_f4(mapa_u32(stg0 + off, (uint32_t)z));
          acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
        const int w_l = m % g.bw;
        const int tt = m / g.bw;
        const int h_l = tt % g.bh, n_l = tt / g.bh;
        const int n = n0 + n_l, oh = oh0 + h_l;
        if (!((n < g.N) && (oh < g.OH) && col_ok)) continue;
        const size_t pix = ((size_t)n * g.out_H + (size_t)(oh * g.out_s + (cls >> 1))) * g.out_W + (size_t)(w_l * g.out_s + (cls & 1));
        if (n >= nlim2) {                                     // padding sample: zeros, no statistics
          *reinterpret_cast<uint2*>(Y + pix * g.Kout + col) = make_uint2(0u, 0u);
          continue;
        }
        if (HAS_RES) {
          const uint2 u = *reinterpret_cast<const uint2*>(residual + pix * g.Kout + col);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
          acc.x += __low2float(h[0]); acc.y += __high2float(h[0]); acc.z += __low2float(h[1]); acc.w += __high2float(h[1]);
        }
        uint2 o;
        __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
        ho[0] = __floats2bfloat162_rn(acc.x, acc.y);
        ho[1] = __floats2bfloat162_rn(acc.z, acc.w);
        *reinterpret_cast<uint2*>(Y + pix * g.Kout + col) = o;
        if (HAS_STATS) {
          const float a0 = __low2float(ho[0]), a1 = __high2float(ho[0]), a2 = __low2float(ho[1]), a3 = __high2float(ho[1]);
          sx[0] += a0; sx[1] += a1; sx[2] += a2; sx[3] += a3;
          sq[0] += a0 * a0; sq[1] += a1 * a1; sq[2] += a2 * a2; sq[3] += a3 * a3;
        }
      }
      if (HAS_STATS) {
        // lanes that share a column group (CG < 32) combine first
#pragma unroll
        for (int o = 16; o >= CG; o >>= 1) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            sx[e] += __shfl_xor_sync(0xffffffffu, sx[e], o);
            sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], o);
          }
        }
        if (lane < CG && col_ok) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            red_add_f32(&stats[col + e], sx[e]);
            red_add_f32(&stats[g.Kout + col + e], sq[e]);
          }
        }
      }
    }
    cluster_sync_all();     // peers may still be reading this CTA's staging tile
  }
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace hfl

using namespace hfl;

// K,R,S,C filter as a 3-D tensor {C, R*S, K_out} (innermost first); box {box_c, 1, box_k}
static int make_tmap_filter_bf16(CUtensorMap* m, const void* Wt, int C, int taps, int Kout, int box_c, int box_k) {
  const uint64_t dims[3] = {(uint64_t)C, (uint64_t)taps, (uint64_t)Kout};
  const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)taps * C * 2};
  const uint32_t box[3] = {(uint32_t)box_c, 1u, (uint32_t)box_k};
  const uint32_t estr[3] = {1u, 1u, 1u};
  return make_tmap_bf16_generic(m, Wt, 3, dims, strides, box, estr);
}

template <int BLOCK_N, bool DGRAD, bool HAS_RES, bool HAS_STATS>
static int launch_conv_v(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tx2, const CUtensorMap& tw2, void* Y,
                         const void* residual, float* stats, const int* nvalid, const ConvGeom& g, int tiles_m, int n_tiles,
                         int ksplit, cudaStream_t stream) {
  auto kern = conv_fprop_tcgen05_kernel<BLOCK_N, DGRAD, HAS_RES, HAS_STATS>;
  constexpr int smem = ConvSmem<BLOCK_N>::TOTAL;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles_m, n_tiles, ksplit);
  cfg.blockDim = dim3(CV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (ksplit > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = ksplit;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tx, tw, tx2, tw2, (__nv_bfloat16*)Y, (const __nv_bfloat16*)residual, stats,
                                     nvalid, g);
  count_launch();
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

// split-K factor of a tile grid: only when the grid leaves most SMs idle and the reduction is long; powers of two
// (the cluster splits the 128 tile rows evenly), every split keeps >= 4 k-blocks.
// Policy (measured, profiles/pdl_conv_ab.md): splitting shortens the isolated step (the 8-GPU regime, one or two clients
// per GPU, chain-bound) but on one GPU, where 10-16 clients already fill the SMs, the 8x larger grids of 200 KB CTAs
// cost more than they save (round 516 vs 480 ms).  Default: on from 4 ranks up, off below; HFL_CONV_KSPLIT=<cap>
// overrides (1 = off).
static int g_ksplit_override = -1;          // tests / probes: hfl_conv_set_ksplit(cap); -1 = policy below
extern "C" void hfl_conv_set_ksplit(int cap) { g_ksplit_override = cap; }

static int conv_ksplit(int tiles, int num_k) {
  static const int policy_cap = []() {
    const char* e = getenv("HFL_CONV_KSPLIT");
    if (e) return atoi(e);
    const char* w = getenv("WORLD_SIZE");
    return (w && atoi(w) >= 4) ? 8 : 1;
  }();
  const int cap = g_ksplit_override >= 1 ? g_ksplit_override : policy_cap;
  // only the deep reductions (C >= 256: 36 / 72 k-blocks) split, and every split keeps >= 8 k-blocks: with 9-18 k-blocks the
  // cluster launch + DSMEM reduction cost more than the shorter mainloop saves (isolated narrow-level steps were 40-50 us
  // SLOWER when their 9-k-block convolutions were split in two, gpurun_out/r2i)
  if (num_k < 32) return 1;
  int ks = 1;
  while (ks * 2 <= cap && tiles * ks * 2 <= 148 && num_k / (ks * 2) >= 8) ks *= 2;
  return ks;
}

template <int BLOCK_N, bool DGRAD = false>
static int launch_conv(const CUtensorMap& tx, const CUtensorMap& tw, void* Y, const void* residual, float* stats,
                       const int* nvalid, const ConvGeom& g, int tiles_m, int n_tiles, cudaStream_t stream,
                       int num_k_hint = -1, const CUtensorMap* tx2p = nullptr, const CUtensorMap* tw2p = nullptr) {
  const CUtensorMap& tx2 = tx2p ? *tx2p : tx;      // unused copies when no second operand pair is folded in
  const CUtensorMap& tw2 = tw2p ? *tw2p : tw;
  if (residual != nullptr && (reinterpret_cast<uintptr_t>(residual) & 15) != 0) return -3;
  if ((reinterpret_cast<uintptr_t>(Y) & 15) != 0) return -3;
  const int num_k = num_k_hint >= 0 ? num_k_hint : g.R * g.S * ((g.C + CV_BLOCK_K - 1) / CV_BLOCK_K);
  const int ks = conv_ksplit(tiles_m * n_tiles, num_k);
  if (residual != nullptr) {
    if (stats != nullptr) return launch_conv_v<BLOCK_N, DGRAD, true, true>(tx, tw, tx2, tw2, Y, residual, stats, nvalid, g, tiles_m, n_tiles, ks, stream);
    return launch_conv_v<BLOCK_N, DGRAD, true, false>(tx, tw, tx2, tw2, Y, residual, stats, nvalid, g, tiles_m, n_tiles, ks, stream);
  }
  if (stats != nullptr) return launch_conv_v<BLOCK_N, DGRAD, false, true>(tx, tw, tx2, tw2, Y, residual, stats, nvalid, g, tiles_m, n_tiles, ks, stream);
  return launch_conv_v<BLOCK_N, DGRAD, false, false>(tx, tw, tx2, tw2, Y, residual, stats, nvalid, g, tiles_m, n_tiles, ks, stream);
}

// 0 if this problem can run on the tcgen05 path
extern "C" int hfl_conv_tc_supported(int N, int H, int W, int C, int Kout, int R, int S, int stride) {
  if (C % 8 != 0 || Kout % 8 != 0) return 1;     // 16-byte global strides for TMA; partial 64-blocks are zero-filled
  if (!((R == 3 && S == 3) || (R == 1 && S == 1))) return 2;
  if (stride != 1 && stride != 2) return 3;
  const int pad = R / 2;
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  if (OW > 128 || (128 % OW) != 0) return 4;
  const int rows = 128 / OW;                   // bh * bn
  if (rows > OH && (rows % OH) != 0) return 5;
  if (rows < OH && (OH % rows) != 0) return 6;
  if (OW * stride > 256 || (rows < OH ? rows : OH) * stride > 256) return 7;
  return 0;
}

// X2 / Wt2 (optional): the block input [N, stride2*OH, stride2*OW, C2] and the 1x1 shortcut filter [Kout, C2] whose
// stride-`stride2` convolution is accumulated into the same output tile (ResNet transition blocks).
extern "C" int hfl_conv_fprop_tc2(const void* X, const void* Wt, void* Y, const void* residual, float* stats,
                                  const int* nvalid, int N, int H, int W, int C, int Kout, int R, int S, int stride,
                                  const void* X2, const void* Wt2, int C2, int stride2, cudaStream_t stream) {
  if (hfl_conv_tc_supported(N, H, W, C, Kout, R, S, stride) != 0) return -2;
  ConvGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.Kout = Kout; g.R = R; g.S = S; g.stride = stride; g.pad = R / 2;
  g.OH = (H + 2 * g.pad - R) / stride + 1;
  g.OW = (W + 2 * g.pad - S) / stride + 1;
  g.bw = g.OW;
  const int rows = 128 / g.OW;
  g.bh = rows < g.OH ? rows : g.OH;
  g.bn = rows / g.bh;
  g.tiles_h = (g.OH + g.bh - 1) / g.bh;
  const int tiles_m = ((N + g.bn - 1) / g.bn) * g.tiles_h;
  g.tiles_m_cls = tiles_m; g.out_s = 1; g.out_H = g.OH; g.out_W = g.OW;
  for (int c = 0; c < 4; ++c) g.ntaps[c] = 0;
  g.kb2 = 0; g.stride2 = 1;
  const int block_n = (Kout % 128 == 0) ? 128 : (Kout > 32 ? 64 : 32);
  CUtensorMap tx, tw, tx2, tw2;
  // traversal strides: box extents are given in traversed elements (ceil(box/stride) elements are loaded)
  int r = make_tmap_nhwc_bf16(&tx, X, (uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N, CV_BLOCK_K,
                              (uint32_t)(g.bw * stride), (uint32_t)(g.bh * stride), (uint32_t)g.bn, (uint32_t)stride,
                              (uint32_t)stride, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r) return 1000 + r;
  r = make_tmap_filter_bf16(&tw, Wt, C, R * S, Kout, CV_BLOCK_K, block_n);
  if (r) return 2000 + r;
  if (X2 != nullptr) {
    if (C2 % 8 != 0 || (stride2 != 1 && stride2 != 2) || g.bw * stride2 > 256 || g.bh * stride2 > 256) return -4;
    g.kb2 = (C2 + CV_BLOCK_K - 1) / CV_BLOCK_K;
    g.stride2 = stride2;
    r = make_tmap_nhwc_bf16(&tx2, X2, (uint64_t)C2, (uint64_t)g.OW * stride2, (uint64_t)g.OH * stride2, (uint64_t)N, CV_BLOCK_K,
                            (uint32_t)(g.bw * stride2), (uint32_t)(g.bh * stride2), (uint32_t)g.bn, (uint32_t)stride2,
                            (uint32_t)stride2, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return 3000 + r;
    r = make_tmap_filter_bf16(&tw2, Wt2, C2, 1, Kout, CV_BLOCK_K, block_n);
    if (r) return 4000 + r;
  }
  const int n_tiles = (Kout + block_n - 1) / block_n;
  const int num_k = R * S * ((C + CV_BLOCK_K - 1) / CV_BLOCK_K) + g.kb2;
  const CUtensorMap* p2 = X2 != nullptr ? &tx2 : nullptr;
  const CUtensorMap* q2 = X2 != nullptr ? &tw2 : nullptr;
  if (block_n == 128) return launch_conv<128>(tx, tw, Y, residual, stats, nvalid, g, tiles_m, n_tiles, stream, num_k, p2, q2);
  if (block_n == 64) return launch_conv<64>(tx, tw, Y, residual, stats, nvalid, g, tiles_m, n_tiles, stream, num_k, p2, q2);
  return launch_conv<32>(tx, tw, Y, residual, stats, nvalid, g, tiles_m, n_tiles, stream, num_k, p2, q2);
}

extern "C" int hfl_conv_fprop_tc(const void* X, const void* Wt, void* Y, const void* residual, float* stats,
                                 const int* nvalid, int N, int H, int W, int C, int Kout, int R, int S, int stride,
                                 cudaStream_t stream) {
  return hfl_conv_fprop_tc2(X, Wt, Y, residual, stats, nvalid, N, H, W, C, Kout, R, S, stride, nullptr, nullptr, 0, 1, stream);
}

// Data gradient: dX[N,H,W,C] from dY[N,OH,OW,Kout] and the K,R,S,C filter, stride 1 or 2 (H = stride * OH).  `residual` is
// an optional tensor added to dX (gradient arriving through a skip connection).  A strided gradient runs as `stride^2`
// parity classes inside ONE launch (see ConvGeom); classes no tap reaches (1x1 filters) are written as zeros.
extern "C" int hfl_conv_dgrad_tc_supported(int N, int H, int W, int C, int Kout, int R, int S, int stride) {
  if (stride != 1 && stride != 2) return 3;
  if (H % stride != 0 || W % stride != 0) return 8;
  return hfl_conv_tc_supported(N, H / stride, W / stride, Kout, C, R, S, 1);   // tiling over the dY grid, roles swapped
}

extern "C" int hfl_conv_dgrad_tc(const void* dY, const void* Wt, void* dX, const void* residual, int N, int H, int W,
                                 int C, int Kout, int R, int S, int stride, cudaStream_t stream) {
  if (hfl_conv_dgrad_tc_supported(N, H, W, C, Kout, R, S, stride) != 0) return -2;
  const int OH = H / stride, OW = W / stride;
  ConvGeom g;
  g.N = N; g.H = OH; g.W = OW; g.R = R; g.S = S; g.stride = 1; g.pad = R / 2;
  g.C = Kout;            // reduction channels (dY)
  g.Kout = C;            // produced channels (dX)
  g.OH = OH; g.OW = OW;
  g.bw = OW;
  const int rows = 128 / OW;
  g.bh = rows < OH ? rows : OH;
  g.bn = rows / g.bh;
  g.tiles_h = (OH + g.bh - 1) / g.bh;
  const int tiles_m = ((N + g.bn - 1) / g.bn) * g.tiles_h;
  g.tiles_m_cls = tiles_m; g.out_s = stride; g.out_H = H; g.out_W = W;
  g.kb2 = 0; g.stride2 = 1;
  const int ncls = stride * stride;
  int max_taps = 0;
  for (int cls = 0; cls < 4; ++cls) {
    g.ntaps[cls] = 0;
    if (cls >= ncls) continue;
    const int ph = stride == 1 ? 0 : (cls >> 1), pw = stride == 1 ? 0 : (cls & 1);
    for (int r = 0; r < R; ++r) {
      if ((ph + g.pad - r) % stride != 0) continue;          // (ph + pad - r) may be negative: C % keeps the sign, 0 stays 0
      for (int s2 = 0; s2 < S; ++s2) {
        if ((pw + g.pad - s2) % stride != 0) continue;
        const int dh = (ph + g.pad - r) / stride, dw = (pw + g.pad - s2) / stride;     // exact divisions
        g.taps[cls][g.ntaps[cls]++] = (r * S + s2) | ((dh + 8) << 8) | ((dw + 8) << 16);
      }
    }
    if (g.ntaps[cls] > max_taps) max_taps = g.ntaps[cls];
  }
  const int block_n = (C % 128 == 0) ? 128 : 64;      // MN-major B: whole 64-channel boxes (zero-filled past C)
  CUtensorMap tx, tw;
  int r = make_tmap_nhwc_bf16(&tx, dY, (uint64_t)Kout, (uint64_t)OW, (uint64_t)OH, (uint64_t)N, CV_BLOCK_K, (uint32_t)g.bw,
                              (uint32_t)g.bh, (uint32_t)g.bn, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r) return 1000 + r;
  r = make_tmap_filter_bf16(&tw, Wt, C, R * S, Kout, 64, CV_BLOCK_K);
  if (r) return 2000 + r;
  const int n_tiles = (C + block_n - 1) / block_n;
  const int grid_m = tiles_m * ncls;
  const int num_k = max_taps * ((Kout + CV_BLOCK_K - 1) / CV_BLOCK_K);
  if (block_n == 128) return launch_conv<128, true>(tx, tw, dX, residual, nullptr, nullptr, g, grid_m, n_tiles, stream, num_k);
  return launch_conv<64, true>(tx, tw, dX, residual, nullptr, nullptr, g, grid_m, n_tiles, stream, num_k);
}

// ---------------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05:   dW[k, r, s, c] = sum_{n,oh,ow} dY[n,oh,ow,k] * X[n, oh*st + r - pad, ow*st + s - pad, c]
//
// GEMM view per filter tap: M = K_out, N = C, reduction = output pixels.  BOTH operands are MN-major (the reduction
// index -- the pixel -- is the strided one in NHWC), which UMMA reads directly: per K-block one 64-pixel brick
// {bn x bh x OW} of dY (two 64-channel boxes -> M = 128) and the tap-shifted brick of X (one 64-channel box), both
// fetched by 4-D TMA.  grid = (K_out/128 tiles, C/64 tiles, taps * splits); split-K partial sums are reduced with fp32
// atomics into a zero-initialised [K,R,S,C] fp32 buffer (which the multi-tensor optimiser consumes as is).
// ---------------------------------------------------------------------------------------------------------
namespace hfl {

struct WgradGeom {
  int N, H, W, C, Kout, OH, OW, R, S, stride, pad;
  int bn, bh, bw;          // 64-pixel brick (bn*bh*bw == 64, bw == OW)
  int bricks_h, bricks;    // bricks per image-group along H, total bricks
  int splits;
};

constexpr int WG_STAGES = 4;
constexpr int WG_A_BYTES = 128 * 64 * 2;    // two 64(k_out) x 64(pixel) boxes
constexpr int WG_B_BYTES = 64 * 64 * 2;     // one 64(c) x 64(pixel) box
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(CV_THREADS, 1)
conv_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                          float* __restrict__ dW, const WgradGeom g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WG_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_blk = blockIdx.x, c_blk = blockIdx.y;
  const int tap = blockIdx.z / g.splits, split = blockIdx.z - tap * g.splits;
  const int r = tap / g.S, s = tap - r * g.S;
  const int per = (g.bricks + g.splits - 1) / g.splits;
  const int b0 = split * per, b1 = min(b0 + per, g.bricks);
  const int num_k = b1 - b0;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_dy); tma_prefetch_desc(&tmap_x); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();   // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel
  if (num_k > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int kb = 0; kb < num_k; ++kb) {
          const int st = kb % WG_STAGES;
          const uint32_t ph = (kb / WG_STAGES) & 1;
          const int brick = b0 + kb;
          const int tn = brick / g.bricks_h, th = brick - tn * g.bricks_h;
          const int n0 = tn * g.bn, oh0 = th * g.bh;
          mbar_wait(&empty_bar[st], ph ^ 1);
          uint8_t* sa = smem + st * WG_STAGE_BYTES;
          uint8_t* sb = sa + WG_A_BYTES;
          // a K_out tail of <= 64 channels loads one dY box only: accumulator rows 64..127 then hold whatever the
          // stale half-stage multiplies to and are never stored (k < K_out mask in the epilogue)
          const bool two = k_blk * 128 + 64 < g.Kout;
          mbar_expect_tx(&full_bar[st], two ? WG_STAGE_BYTES : WG_STAGE_BYTES - 8192);
          tma_load_4d(sa, &tmap_dy, &full_bar[st], k_blk * 128, 0, oh0, n0);
          if (two) tma_load_4d(sa + 8192, &tmap_dy, &full_bar[st], k_blk * 128 + 64, 0, oh0, n0);
          tma_load_4d(sb, &tmap_x, &full_bar[st], c_blk * 64, s - g.pad, oh0 * g.stride + r - g.pad, n0);
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc_bf16(128, 64, true, true);
      for (int kb = 0; kb < num_k; ++kb) {
        const int st = kb % WG_STAGES;
        const uint32_t ph = (kb / WG_STAGES) & 1;
        mbar_wait(&full_bar[st], ph);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + st * WG_STAGE_BYTES);
          const uint64_t adesc = make_smem_desc_sw128_mn(a_addr);
          const uint64_t bdesc = make_smem_desc_sw128_mn(a_addr + WG_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)      // 64 pixels = 4 x UMMA_K(16); +16 k-rows = +2048 B
            umma_bf16(tmem_base, adesc + (2048 >> 4) * k, bdesc + (2048 >> 4) * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[st]);
          if (kb == num_k - 1) umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
    } else if (warp >= 4) {
      const int q = warp - 4;
      mbar_wait(tmem_full_bar, 0);
      tcgen05_fence_after();
      // stage the warp's 32 x 64 fp32 block in shared memory, then write / accumulate whole 256-byte filter rows
      // with 16-byte accesses (a row per thread made every store or atomic touch 32 different rows)
      constexpr int PITCH = 64 * 4 + 16;
      uint8_t* my = smem + (size_t)q * 32 * PITCH;
      {
        uint8_t* myrow = my + lane * PITCH;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(myrow + (c0 + j) * 4) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      __syncwarp();
      const int rsub = lane >> 4, cseg = (lane & 15) * 4;     // 16 lanes x 4 floats per 64-channel row, 2 rows / instr
#pragma unroll 4
      for (int r0 = 0; r0 < 32; r0 += 2) {
        const int k = k_blk * 128 + q * 32 + r0 + rsub;
        if (k < g.Kout && c_blk * 64 + cseg < g.C) {
          const float4 val = *reinterpret_cast<const float4*>(my + (r0 + rsub) * PITCH + cseg * 4);
          float* dst = dW + ((size_t)k * g.R * g.S + tap) * g.C + c_blk * 64 + cseg;
          if (g.splits == 1) {
            *reinterpret_cast<float4*>(dst) = val;
          } else {
            atomicAdd(reinterpret_cast<float4*>(dst), val);
          }
        }
      }
      tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace hfl

extern "C" int hfl_conv_wgrad_tc_supported(int N, int H, int W, int C, int Kout, int R, int S, int stride) {
  if (C % 8 != 0 || Kout % 8 != 0) return 1;
  if (!((R == 3 && S == 3) || (R == 1 && S == 1))) return 2;
  if (stride != 1 && stride != 2) return 3;
  const int pad = R / 2;
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  if (OW > 64 || (64 % OW) != 0) return 4;
  const int rows = 64 / OW;
  if (rows > OH && (rows % OH) != 0) return 5;
  if (rows < OH && (OH % rows) != 0) return 6;
  const int bh = rows < OH ? rows : OH, bn = rows / bh;
  (void)bn;                         // a partial last image group is fine: both TMA boxes zero-fill the images beyond N
  return 0;
}

// dW (fp32, [K,R,S,C]) must be zero-initialised by the caller when the returned split count is > 1; returns < 0 on
// error, otherwise the number of K-splits used.
extern "C" int hfl_conv_wgrad_tc(const void* dY, const void* X, float* dW, int N, int H, int W, int C, int Kout, int R,
                                 int S, int stride, int zero_first, cudaStream_t stream) {
  if (hfl_conv_wgrad_tc_supported(N, H, W, C, Kout, R, S, stride) != 0) return -2;
  hfl::WgradGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.Kout = Kout; g.R = R; g.S = S; g.stride = stride; g.pad = R / 2;
  g.OH = (H + 2 * g.pad - R) / stride + 1;
  g.OW = (W + 2 * g.pad - S) / stride + 1;
  g.bw = g.OW;
  const int rows = 64 / g.OW;
  g.bh = rows < g.OH ? rows : g.OH;
  g.bn = rows / g.bh;
  g.bricks_h = g.OH / g.bh;
  g.bricks = ((N + g.bn - 1) / g.bn) * g.bricks_h;
  const int base_ctas = ((Kout + 127) / 128) * ((C + 63) / 64) * R * S;
  int splits = 1;
  while (base_ctas * splits < 96 && splits * 8 <= g.bricks) splits *= 2;   // fill the machine, keep >= 8 bricks per CTA
  g.splits = splits;
  CUtensorMap tdy, tx;
  int r = hfl::make_tmap_nhwc_bf16(&tdy, dY, (uint64_t)Kout, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)N, 64, (uint32_t)g.bw,
                                   (uint32_t)g.bh, (uint32_t)g.bn, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r) return -(1000 + r);
  r = hfl::make_tmap_nhwc_bf16(&tx, X, (uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N, 64, (uint32_t)(g.bw * stride),
                               (uint32_t)(g.bh * stride), (uint32_t)g.bn, (uint32_t)stride, (uint32_t)stride,
                               CU_TENSOR_MAP_SWIZZLE_128B);
  if (r) return -(2000 + r);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hfl::conv_wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         hfl::WG_SMEM);
    if (e != cudaSuccess) return -(int)e;
    configured = true;
  }
  if (splits > 1 && zero_first) cudaMemsetAsync(dW, 0, (size_t)Kout * R * S * C * sizeof(float), stream);
  dim3 grid((Kout + 127) / 128, (C + 63) / 64, R * S * splits);
  hfl::launch_k(hfl::conv_wgrad_tcgen05_kernel, dim3(grid), dim3(hfl::CV_THREADS), hfl::WG_SMEM, stream, tdy, tx, dW, g);
  hfl::count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? splits : -(int)e;
}
