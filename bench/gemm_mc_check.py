"""Correctness + throughput of the cluster/multicast GEMM (csrc/gemm_mc_tcgen05.cu) vs the single-CTA kernel and cuBLAS."""
import os
import sys

import torch

sys.path.insert(0, ".")
from hfl_b200.ops import cuda as cu

C = cu.C()
dev = torch.device("cuda")
out = []
for (M, N, K) in ((512, 512, 256), (1000, 768, 320), (640, 33278, 256), (8192, 4096, 4096), (6400, 33278, 256)):
    torch.manual_seed(M + N)
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    Np = (N + 7) // 8 * 8
    buf = torch.zeros(M, Np, device=dev).bfloat16()
    y = C.gemm_bf16_tn(x, w, b, 0.5, True, False, True, buf[:, :N])
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(0.5 * (x.float() @ w.float().t() + b.float()))
    err = float((y.float() - ref).abs().max())
    y2 = C.gemm_bf16_tn(x, w, None, 1.0, False, False, False, None)
    err2 = float((y2.float() - x.float() @ w.float().t()).abs().max())

    def t(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps
    t_mc = t(lambda: C.gemm_bf16_tn(x, w, None, 1.0, False, False, False, None))
    t_lib = t(lambda: x @ w.t())
    fl = 2.0 * M * N * K
    out.append({"shape": [M, N, K], "max_err_gelu_bias": err, "max_err_plain": err2, "us": round(t_mc, 1),
                "tflops": round(fl / t_mc / 1e6, 1), "cublas_us": round(t_lib, 1), "cublas_tflops": round(fl / t_lib / 1e6, 1),
                "mc": os.environ.get("HFL_GEMM_MC", "1")})
    print(out[-1], flush=True)
