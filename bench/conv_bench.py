"""tcgen05 implicit-GEMM conv vs the library conv (cuDNN) on the ResNet-18 level-a/b layer shapes.
CUDA events, 20 warm-up + 100 timed launches each, inputs rotated through > 126 MB of buffers (L2 flush)."""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from hfl_b200.ops import cuda as cu

dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
C_ = cu.C()
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {}
rows = []
for batch in (10, 2560):
    for (C, H, K, R, stride) in [(64, 32, 64, 3, 1), (64, 32, 128, 3, 2), (64, 32, 128, 1, 2), (128, 16, 128, 3, 1),
                                 (128, 16, 256, 3, 2), (256, 8, 256, 3, 1), (256, 8, 512, 3, 2), (512, 4, 512, 3, 1)]:
        nbuf = max(2, int(200e6 // (batch * C * H * H * 2)) + 1) if batch > 100 else 4
        nbuf = min(nbuf, 8)
        xs = [torch.randn(batch, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
              for _ in range(nbuf)]
        w = (torch.randn(K, C, R, R, device=dev) * 0.05).bfloat16().contiguous(memory_format=torch.channels_last)
        if R == 1:
            w = w.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        OH = (H + 2 * (R // 2) - R) // stride + 1
        flops = 2.0 * batch * OH * OH * K * C * R * R

        def timeit(fn):
            for i in range(20):
                fn(xs[i % nbuf])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(100):
                fn(xs[i % nbuf])
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 10.0   # us per launch

        t_lib = timeit(lambda x: F.conv2d(x, w, None, stride, R // 2))
        t_tc = timeit(lambda x: C_.conv_fprop_tc(x, w, None, None, stride))
        rows.append({"batch": batch, "C": C, "H": H, "K": K, "R": R, "stride": stride, "lib_us": round(t_lib, 2),
                     "tc_us": round(t_tc, 2), "tc_tflops": round(flops / t_tc / 1e6, 1),
                     "lib_tflops": round(flops / t_lib / 1e6, 1),
                     "tc_frac_of_measured_bf16_peak": round(flops / t_tc / 1e6 / peaks.get("bf16_tflops", 1737.9), 3)})
print(json.dumps(rows, indent=1))
