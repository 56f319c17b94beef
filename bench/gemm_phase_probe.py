"""Phase timestamps (clock64) of the tcgen05 GEMM on a tiny problem; build with HFL_NVCC_EXTRA=-DHFL_GEMM_TIMING."""
import sys

import torch

sys.path.insert(0, ".")
from hfl_b200.ops import cuda as cu

C = cu.C()
dev = torch.device("cuda")
for (M, K, N) in ((64, 256, 256), (64, 64, 256), (128, 1024, 128)):
    x = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    for _ in range(5):
        C.gemm_bf16_tn(x, w, None, 1.0, False, False, False, None)
    torch.cuda.synchronize()
    t = C.gemm_timing()
    names = ["start", "setup_done", "tma_first", "tma_last", "full_first", "mma_last", "acc_ready", "epi_end", "pre_dealloc"]
    print((M, K, N), {n: t[i] - t[0] for i, n in enumerate(names)})
