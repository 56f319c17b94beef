"""tcgen05 GEMM (csrc/gemm_tcgen05.cu) vs cuBLAS, bf16: TN forward, MN-major backward forms, split-K skinny case."""
import json
import sys

import torch

sys.path.insert(0, ".")
from hfl_b200.ops import cuda as cu

C = cu.C()
dev = torch.device("cuda")


def timeit(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


rows = []
for (M, N, K, tag) in ((8192, 4096, 4096, "square"), (6400, 33278, 256, "vocab fwd, centralised batch"),
                       (64, 33278, 256, "vocab fwd, federated batch"), (6400, 512, 256, "ffn1 centralised"),
                       (64, 256, 256, "qkv federated")):
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    reps = 20 if M * N * K > 1e10 else 200
    t_tc = timeit(lambda: C.gemm_bf16_tn(x, w, b, 1.0, False, False, True, None), reps)
    t_lib = timeit(lambda: torch.addmm(b, x, w.t()), reps)
    fl = 2.0 * M * N * K
    rows.append({"op": "y = x w^T + b", "shape": [M, N, K], "what": tag, "tc_us": round(t_tc, 2), "cublas_us": round(t_lib, 2),
                 "tc_tflops": round(fl / t_tc / 1e6, 1), "cublas_tflops": round(fl / t_lib / 1e6, 1)})
    dy = torch.randn(M, (N + 7) // 8 * 8, device=dev).bfloat16()[:, :N]     # row-padded like the fused LM step's logits
    t_dx = timeit(lambda: C.gemm_bf16_general(dy, False, w, True, 1.0, False), reps)
    t_dx_lib = timeit(lambda: dy @ w, reps)
    t_dw = timeit(lambda: C.gemm_bf16_general(dy, True, x, True, 1.0, False), reps)
    t_dw_lib = timeit(lambda: dy.t() @ x, reps)
    rows.append({"op": "dX = dY w (w MN-major)", "shape": [M, N, K], "what": tag, "tc_us": round(t_dx, 2),
                 "cublas_us": round(t_dx_lib, 2), "tc_tflops": round(fl / t_dx / 1e6, 1),
                 "cublas_tflops": round(fl / t_dx_lib / 1e6, 1)})
    rows.append({"op": "dW = dY^T x (both MN-major)", "shape": [M, N, K], "what": tag, "tc_us": round(t_dw, 2),
                 "cublas_us": round(t_dw_lib, 2), "tc_tflops": round(fl / t_dw / 1e6, 1),
                 "cublas_tflops": round(fl / t_dw_lib / 1e6, 1)})
print(json.dumps(rows, indent=1))
