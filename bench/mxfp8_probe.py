"""MXFP8 GEMM throughput vs the bf16 tcgen05 GEMM and cuBLAS (CUDA events, warm, L2 flushed between timed calls)."""
import sys

import torch

sys.path.insert(0, ".")
from hfl_b200.ops import cuda as cu
from hfl_b200.ops import mxfp8 as mx


def timeit(fn, flush, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.add_(1)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    dev = torch.device("cuda")
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    print("| M | N | K | MXFP8 us | TFLOP/s | quantise A us | tcgen05 bf16 us | TFLOP/s | cuBLAS bf16 us | TFLOP/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for (M, N, K) in [(8192, 8192, 8192), (4096, 4096, 4096), (6400, 1024, 256), (6400, 256, 1024), (6400, 33280, 256)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        aq, sfa = mx.quantize(a)
        bq, sfb = mx.quantize(b)
        fl = 2.0 * M * N * K
        t_mx = timeit(lambda: mx.mx_gemm(aq, sfa, bq, sfb), flush)
        t_q = timeit(lambda: mx.quantize(a), flush)
        t_tc = timeit(lambda: cu.C().gemm_bf16_tn(a, b, None, 1.0, False, False, False, None), flush)
        t_cb = timeit(lambda: torch.matmul(a, b.t()), flush)
        print("| %d | %d | %d | %.1f | %.0f | %.1f | %.1f | %.0f | %.1f | %.0f |" % (
            M, N, K, t_mx * 1e6, fl / t_mx * 1e-12, t_q * 1e6, t_tc * 1e6, fl / t_tc * 1e-12, t_cb * 1e6,
            fl / t_cb * 1e-12))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
