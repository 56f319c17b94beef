// Phase-program ABI of the resident client-training megakernel (csrc/mega_kernel.cu).
//
// A program is an array of fixed-size `Phase` records in device memory; the persistent cooperative kernel walks it
// `nsteps` times, separating phases by a device-side grid barrier (unless `sync_after == 0`).  The Python side
// (models/mega_resnet.py) assembles the records with numpy using the same field indices.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace hfl {
namespace mega {

constexpr int NIV = 44, NFV = 16, NPV = 16, NTM = 4;

struct alignas(64) Phase {
  int op, ntiles, sync_after, flags;
  int iv[NIV];
  float fv[NFV];
  unsigned long long pv[NPV];
  CUtensorMap tm[NTM];
};
static_assert(sizeof(Phase) == 896, "Phase ABI");

enum Op : int {
  OP_NOP = 0,
  OP_PREP = 1,        // batch assembly (gather + crop + flip + normalise), weight re-packing, per-step resets
  OP_CONV = 2,        // tcgen05 implicit-GEMM conv: fprop / dgrad (stride 1, or one parity class of stride 2)
  OP_WGRAD = 3,       // tcgen05 weight gradient, fp32 atomics into the flat gradient arena
  OP_NAPPLY = 4,      // z = relu(norm(y)) from batch statistics (or a given affine)
  OP_NBWD_RED = 5,    // BN backward reductions
  OP_NBWD_APPLY = 6,  // BN backward apply (+ identity-path gradient), dgamma / dbeta
  OP_HEAD = 7,        // pool + linear + masked CE (+ backward) / eval logits
  OP_SGD_NORM = 8,    // ||g||^2 over the flat arena
  OP_SGD_UPDATE = 9,  // clip + weight decay + momentum + update, bf16 copy, gradient / scratch zeroing
  OP_CFIN = 10,       // split-K finalisation: sum the fp32 partial slices, residual, bf16 store, BN statistics
};

// One client (or one evaluation pass) of a launch.  All jobs of a launch share the phase structure (same ops / barriers
// per phase index); a job is idle once `step >= nsteps`.
struct alignas(16) Job {
  const Phase* prog;
  const int* nvalid_tab;      // [nsteps] valid samples per step (nullptr: all)
  unsigned* jctl;             // per-job scalars: CTL_GNORM2 / CTL_LAST_GNORM words
  int nsteps;
  int pad;
};
static_assert(sizeof(Job) == 32, "Job ABI");
constexpr int MAX_JOBS = 16;

// control block (uint32 words)
constexpr int CTL_ARRIVE = 0, CTL_ABORT = 1, CTL_STEP = 2, CTL_PHASE = 3, CTL_GNORM2 = 4, CTL_LAST_GNORM = 5,
              CTL_WORDS = 16;

}  // namespace mega
}  // namespace hfl
