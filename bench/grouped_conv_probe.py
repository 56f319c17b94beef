"""Would client-stacked (channel-grouped) convolutions pay?  cuDNN conv fwd / dgrad / wgrad at batch 10 for one client
vs G same-width clients stacked along channels (groups=G), bf16 NHWC, CUDA-event timing inside a captured graph."""
import sys

import torch
import torch.nn.functional as F

torch.backends.cudnn.benchmark = True


def graph_time(fn, iters=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    dev = torch.device("cuda")
    print("| C | H | G | fwd us | dgrad us | wgrad us | fwd us / client | bwd us / client |")
    print("|---|---|---|---|---|---|---|---|")
    for (C, H) in [(64, 32), (128, 16), (256, 8), (512, 4), (32, 32), (16, 32), (4, 32)]:
        for G in (1, 2, 4):
            x = torch.randn(10, G * C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
            w = torch.randn(G * C, C, 3, 3, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
            dy = torch.randn_like(x)
            tf = graph_time(lambda: F.conv2d(x, w, None, 1, 1, 1, G))
            td = graph_time(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (1, 1), (1, 1), False,
                                                                        (0, 0), G, (True, False, False)))
            tw = graph_time(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (1, 1), (1, 1), False,
                                                                        (0, 0), G, (False, True, False)))
            print("| %d | %d | %d | %.1f | %.1f | %.1f | %.1f | %.1f |" % (C, H, G, tf, td, tw, tf / G, (td + tw) / G))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
