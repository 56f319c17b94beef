"""Why does a 64x256x256 tcgen05 GEMM take ~14 us?  Back-to-back timing of the tiny-GEMM variants, alone and
interleaved with a small-shared-memory kernel (carve-out switch hypothesis), vs cuBLAS."""
import sys

import torch

sys.path.insert(0, ".")
from hfl_b200.ops import cuda as cu

C = cu.C()
dev = torch.device("cuda")
M, K, N = 64, 256, 256
x = torch.randn(M, K, device=dev).bfloat16()
w = torch.randn(N, K, device=dev).bfloat16()
b = torch.randn(N, device=dev).bfloat16()
dy = torch.randn(M, N, device=dev).bfloat16()
g = torch.ones(K, device=dev).bfloat16()
seed = torch.zeros(1, dtype=torch.int64, device=dev)


def timeit(fn, reps=300):
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr, stream=s):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


res = {}
res["tn_64x256x256"] = timeit(lambda: C.gemm_bf16_tn(x, w, b, 1.0, False, False, True, None))
res["general_dx"] = timeit(lambda: C.gemm_bf16_general(dy, False, w, True, 1.0, False))
res["general_dw"] = timeit(lambda: C.gemm_bf16_general(dy, True, x, True, 1.0, False))
res["cublas_fwd"] = timeit(lambda: torch.addmm(b, x, w.t()))
res["cublas_dw"] = timeit(lambda: dy.t() @ x)
res["ln_only"] = timeit(lambda: C.lm_add_ln_fwd(None, x, g, g, 1e-5, 0.0, seed, seed, 0))


def pair():
    C.gemm_bf16_tn(x, w, b, 1.0, False, False, True, None)
    C.lm_add_ln_fwd(None, x, g, g, 1e-5, 0.0, seed, seed, 0)


res["tn_plus_ln_pair"] = timeit(pair)
xl = torch.randn(64, 256, device=dev).bfloat16()
wl = torch.randn(33278, 256, device=dev).bfloat16()
dl = torch.randn(64, 33280, device=dev).bfloat16()[:, :33278]
res["vocab_fwd"] = timeit(lambda: C.gemm_bf16_tn(xl, wl, None, 1.0, False, False, False, None), 50)
res["vocab_dx"] = timeit(lambda: C.gemm_bf16_general(dl, False, wl, True, 1.0, False), 50)
res["vocab_dw"] = timeit(lambda: C.gemm_bf16_general(dl, True, xl, True, 1.0, False), 50)
res["cublas_vocab_dx"] = timeit(lambda: dl @ wl, 50)
for k, v in res.items():
    print("%-20s %8.2f us" % (k, v))
