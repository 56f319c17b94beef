// Cross-GPU stream-ordered barrier over peer-mapped flag words (NVLink P2P), plus launch accounting.
//
// Each rank owns `flags[world]` (uint32, zero-initialised) inside its peer-mapped pool.  Barrier #epoch:
//   rank r stores `epoch` into flags_of_peer[p][r] for every p   (st.release.sys -> visible across NVLink)
//   then spins until its own flags[q] >= epoch for every q       (ld.acquire.sys)
// The kernel is enqueued on the compute stream, so "all ranks finished local training" / "new W_g published" are
// device-side events: no host synchronisation and no NCCL call sits between the phases of a round.
// A bounded spin (~60 s: a peer may legitimately be busy capturing CUDA graphs on a cold start) turns a lost peer into
// an error flag (PeerPool.check) instead of a hung GPU.
#include <atomic>
#include <cstdlib>
#include "common.cuh"

namespace hfl {

static std::atomic<long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// HFL_PDL=0 launches every kernel with plain stream serialisation (A/B switch; read once)
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("HFL_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

__global__ void peer_barrier_kernel(const unsigned long long* __restrict__ flag_ptrs, int rank, int world,
                                    unsigned epoch, int* __restrict__ err) {
  const int p = threadIdx.x;
  // "this rank's earlier kernels have completed" is what the flag publishes.  No early launch_dependents here: the kernel may
  // spin for as long as the slowest peer trains, and a dependent launched early (the aggregation grid: ~2 800 CTAs parked in
  // griddepcontrol.wait) would occupy every SM slot meanwhile -- starving the evaluation stream of the asynchronous mode.
  pdl_wait();
  if (p < world) {
    unsigned* remote = reinterpret_cast<unsigned*>(flag_ptrs[p]) + rank;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const unsigned* mine = reinterpret_cast<const unsigned*>(flag_ptrs[rank]) + p;
    const long long t0 = clock64();
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > 120000000000LL) { if (err) atomicExch(err, 1); break; }
      __nanosleep(200);
    } while (true);
  }
  __syncthreads();
}

}  // namespace hfl

using namespace hfl;

extern "C" long hfl_launch_count() { return g_launches.load(std::memory_order_relaxed); }
extern "C" void hfl_launch_count_add(long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" void hfl_peer_barrier(const void* flag_ptrs, int rank, int world, unsigned epoch, int* err,
                                 cudaStream_t stream) {
  launch_k(peer_barrier_kernel, dim3(1), dim3(32 * ((world + 31) / 32)), 0, stream, (const unsigned long long*)flag_ptrs, rank, world,
                                                                  epoch, err);
  count_launch();
}
