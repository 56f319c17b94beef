"""tcgen05 dgrad / wgrad vs the library (aten.convolution_backward -> cuDNN) on ResNet-18 level-a training shapes."""
import json
import sys

import torch

sys.path.insert(0, ".")
from hfl_b200.ops import cuda as cu

dev = torch.device("cuda")
C_ = cu.C()
rows = []


def timeit(fn, n=100):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / n


for (N, C, H, K, R, stride) in [(10, 64, 32, 64, 3, 1), (10, 64, 32, 128, 3, 2), (10, 128, 16, 128, 3, 1),
                                (10, 128, 16, 256, 3, 2), (10, 256, 8, 256, 3, 1), (10, 256, 8, 512, 3, 2),
                                (12, 512, 4, 512, 3, 1)]:
    x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(K, C, R, R, device=dev) * 0.05).bfloat16().contiguous(memory_format=torch.channels_last)
    OH = (H + 2 * (R // 2) - R) // stride + 1
    dy = torch.randn(N, K, OH, OH, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    pad = R // 2
    rec = {"N": N, "C": C, "H": H, "K": K, "R": R, "stride": stride}
    rec["lib_dgrad_us"] = round(timeit(lambda: torch.ops.aten.convolution_backward(
        dy, x, w, None, (stride, stride), (pad, pad), (1, 1), False, (0, 0), 1, (True, False, False))), 2)
    rec["lib_wgrad_us"] = round(timeit(lambda: torch.ops.aten.convolution_backward(
        dy, x, w, None, (stride, stride), (pad, pad), (1, 1), False, (0, 0), 1, (False, True, False))), 2)
    if stride == 1:
        rec["tc_dgrad_us"] = round(timeit(lambda: C_.conv_dgrad_tc(dy, w, None)), 2)
    if C_.conv_wgrad_tc_supported(N, H, H, C, K, R, R, stride):
        rec["tc_wgrad_us"] = round(timeit(lambda: C_.conv_wgrad_tc(dy, x, R, stride)), 2)
    rows.append(rec)
print(json.dumps(rows, indent=1))
