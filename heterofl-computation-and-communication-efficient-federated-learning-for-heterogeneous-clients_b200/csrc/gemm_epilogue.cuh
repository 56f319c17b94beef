// Shared-memory transposed epilogue for 128-row tcgen05 accumulator tiles (used by gemm_tcgen05.cu / gemm_mc_tcgen05.cu).
#pragma once
#include "tc_common.cuh"

namespace hfl {

// ---------------------------------------------------------------------------------------------------------
// Epilogue of one 128 x BLOCK_N fp32 accumulator tile held in TMEM.  Warp q (0..3) owns TMEM lanes [32q, 32q+32),
// i.e. one accumulator ROW per thread.  Writing C straight from that mapping makes every store instruction touch
// 32 different rows (32 L1 wavefronts per instruction) -- measured 14 000 cycles per tile, as long as the whole
// 64-k-block mainloop.  Instead each warp transposes through shared memory (the drained pipeline stages; row pitch
// +16 B so the 16-byte staging writes are conflict-free) and writes whole row segments with 16-byte stores.
// `xf(acc, tile_col)` applies alpha / bias / activation.
// ---------------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool OUT_BF16, bool ATOMIC = false, typename XF>
__device__ __forceinline__ void epilogue_store_tile(uint32_t tmem_base, uint8_t* smem_stage, int q, int lane, void* C,
                                                    int ldc, int row0, int col0, int M, int N, XF xf) {
  constexpr int ELT = OUT_BF16 ? 2 : 4;
  constexpr int ROW_BYTES = BLOCK_N * ELT;
  constexpr int PITCH = ROW_BYTES + 16;
  constexpr int LPR = ROW_BYTES / 16;            // lanes per output row in the write-out
  static_assert(LPR >= 1 && LPR <= 32, "tile row must be 16..512 bytes");
  constexpr int RPI = 32 / LPR;                  // rows per store instruction
  uint8_t* my = smem_stage + (size_t)q * 32 * PITCH;
  uint8_t* myrow = my + lane * PITCH;
#pragma unroll 1
  for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = xf(__uint_as_float(v[j]), c0 + j);
    if (OUT_BF16) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
        h[0] = __floats2bfloat162_rn(f[j], f[j + 1]);
        h[1] = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
        h[2] = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
        h[3] = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
        *reinterpret_cast<uint4*>(myrow + (c0 + j) * 2) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(myrow + (c0 + j) * 4) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    }
  }
  __syncwarp();
  const int rsub = lane / LPR, cb = (lane % LPR) * 16;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (((size_t)ldc * ELT) % 16 == 0);
  const int gcol = col0 + cb / ELT;
#pragma unroll 4
  for (int r0 = 0; r0 < 32; r0 += RPI) {
    const int r = r0 + rsub;
    const int grow = row0 + q * 32 + r;
    if (grow < M && gcol < N) {
      const uint8_t* src = my + r * PITCH + cb;
      uint8_t* dst = reinterpret_cast<uint8_t*>(C) + ((size_t)grow * ldc + gcol) * ELT;
      if (ATOMIC) {                       // split-K partial: fp32 accumulate into a zero-filled C
        for (int e = 0; e < 4 && gcol + e < N; ++e)
          atomicAdd(reinterpret_cast<float*>(dst) + e, reinterpret_cast<const float*>(src)[e]);
      } else if (vec_ok && gcol + 16 / ELT <= N) {
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
      } else {
        for (int e = 0; e < 16 / ELT && gcol + e < N; ++e) {
          if (OUT_BF16) reinterpret_cast<uint16_t*>(dst)[e] = reinterpret_cast<const uint16_t*>(src)[e];
          else reinterpret_cast<uint32_t*>(dst)[e] = reinterpret_cast<const uint32_t*>(src)[e];
        }
      }
    }
  }
}

}  // namespace hfl
