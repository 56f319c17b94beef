"""Abort-and-resume supervisor (the reference has no failure handling beyond ``--resume_mode 1``).

``python -m hfl_b200.cli.supervise [--max_restarts N] [--backoff S] -- <training command ...>``

runs the command (typically ``python -m torch.distributed.run ... train_classifier_fed.py ...``) and, when it exits
non-zero -- a rank died, the NCCL watchdog aborted a collective (``TORCH_NCCL_ASYNC_ERROR_HANDLING``), a device-side peer
barrier timed out (``parallel/peer.py::PeerPool.check``) -- re-launches it with ``--resume_mode 1`` so the run continues
from the last per-round checkpoint (``fed/engine.py::save_checkpoint``; the checkpoint is written every round).  The
child runs in its own process group and is stopped by pid/pgid only.
"""
from __future__ import annotations

import argparse
import os
import signal
import subprocess
import sys
import time
from typing import List, Optional


def with_resume(cmd: List[str]) -> List[str]:
    """Return ``cmd`` with ``--resume_mode 1`` (replacing an existing value)."""
    out, skip = [], False
    for a in cmd:
        if skip:
            skip = False
            continue
        if a == "--resume_mode":
            skip = True
            continue
        if a.startswith("--resume_mode="):
            continue
        out.append(a)
    return out + ["--resume_mode", "1"]


def supervise(cmd: List[str], max_restarts: int = 3, backoff_s: float = 2.0, timeout_s: Optional[float] = None,
              env: Optional[dict] = None) -> int:
    env = dict(os.environ if env is None else env)
    env.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")     # a stuck collective aborts the process, not the job
    env.setdefault("TORCH_NCCL_BLOCKING_WAIT", "0")
    attempt, cur = 0, list(cmd)
    while True:
        t0 = time.time()
        proc = subprocess.Popen(cur, env=env, start_new_session=True)
        try:
            rc = proc.wait(timeout=timeout_s)
        except subprocess.TimeoutExpired:
            os.killpg(proc.pid, signal.SIGTERM)              # exactly the group we started
            try:
                proc.wait(timeout=20)
            except subprocess.TimeoutExpired:
                os.killpg(proc.pid, signal.SIGKILL)
                proc.wait()
            rc = 124
        except KeyboardInterrupt:
            os.killpg(proc.pid, signal.SIGTERM)
            raise
        if rc == 0:
            return 0
        attempt += 1
        print("[supervise] attempt %d exited with %d after %.0fs" % (attempt, rc, time.time() - t0), file=sys.stderr)
        if attempt > max_restarts:
            return rc
        time.sleep(backoff_s)
        cur = with_resume(cmd)
        print("[supervise] resuming: %s" % " ".join(cur), file=sys.stderr)


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if "--" in argv:
        i = argv.index("--")
        mine, cmd = argv[:i], argv[i + 1:]
    else:
        mine, cmd = [], argv
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--max_restarts", type=int, default=3)
    ap.add_argument("--backoff", type=float, default=2.0)
    ap.add_argument("--timeout", type=float, default=0.0, help="per-attempt wall-clock limit in seconds (0 = none)")
    a = ap.parse_args(mine)
    if not cmd:
        ap.error("no command given")
    return supervise(cmd, a.max_restarts, a.backoff, a.timeout or None)


if __name__ == "__main__":
    sys.exit(main())
