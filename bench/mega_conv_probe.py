"""Single-phase probes of the megakernel's conv tiles: us per phase for the deep, few-tile layers at the training batch
(batch 10) as a function of split-K, measured as `steps` repetitions of a one-phase program (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hfl_b200.models import mega_resnet as mr      # noqa: E402

DEV = "cuda"


def act(N, H, C):
    t = torch.zeros(N, H, H, mr.rup(C, 8), dtype=torch.bfloat16, device=DEV)
    t[..., :C] = torch.randn(N, H, H, C, device=DEV).to(torch.bfloat16)
    return t


def run(prog, grid, steps):
    pd = prog.finalize(DEV)
    ctl = torch.zeros(16, dtype=torch.int32, device=DEV)
    jctl = torch.zeros(16, dtype=torch.int32, device=DEV)
    mr.launch([(pd, None, jctl, 20)], len(prog), grid, ctl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mr.launch([(pd, None, jctl, steps)], len(prog), grid, ctl)
    e1.record()
    torch.cuda.synchronize()
    assert ctl.tolist()[1] == 0, ctl.tolist()
    return 1e3 * e0.elapsed_time(e1) / steps


def main():
    torch.manual_seed(0)
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    for (N, H, Cin, Cout, R, stride) in [(10, 4, 512, 512, 3, 1), (10, 8, 256, 256, 3, 1), (10, 16, 128, 128, 3, 1),
                                         (10, 32, 64, 64, 3, 1)]:
        x = act(N, H, Cin)
        w = (torch.randn(Cout, R * R, Cin, device=DEV) * 0.02).to(torch.bfloat16)
        out = torch.zeros(N, H // stride, H // stride, Cout, dtype=torch.bfloat16, device=DEV)
        wd = mr.WDesc(w.data_ptr(), Cout, R * R, Cin, Cin)
        # flush weights out of L2 between repetitions? The real step streams ~300 MB between two uses of a layer's
        # weights, so also time a variant whose program touches a 256 MB buffer in a second phase.
        big = torch.zeros(64 << 20, dtype=torch.float32, device=DEV)
        for split in (0, 2 * 8, 3 * 8, 6 * 8, 9 * 8, 18 * 8):
            for flush in (0, 1):
                prog = mr.Program()
                mr.conv_fprop(prog, x, wd, out, R, stride, split_target=split, splittable=True)
                if flush:
                    mr.zero_phase(prog, big)
                t = run(prog, grid, 50)
                rec = prog.recs[0]
                print("N%d %dx%d %d->%d  tiles %3d ksplit %d num_k %3d  flush %d: %7.1f us/phase%s" % (
                    N, H, H, Cin, Cout, int(rec["ntiles"]), max(int(rec["flags"]), 1), int(rec["iv"][22]), flush, t,
                    "  (incl. ~%.0f us flush)" % (big.numel() * 4 / 6.4e6) if flush else ""), flush=True)


if __name__ == "__main__":
    main()
