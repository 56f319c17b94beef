"""Phase programs of the resident client-training megakernel (``csrc/mega_kernel.cu``) for pre-activation ResNets.

A *program* is a device array of fixed-size phase records (``csrc/mega.h``): the persistent cooperative kernel walks it
once per optimisation step, separating dependent phases by a device-side grid barrier.  This module lays out every
activation / gradient buffer of a client sub-network at one width and emits the records for

* ``train``: batch assembly -> forward (tcgen05 convolutions with the BatchNorm batch statistics folded into their
  epilogue, one apply pass per norm site) -> head with masked cross-entropy -> backward (dgrad / wgrad convolutions,
  BN backward) -> clip + SGD.  A whole local training of a client (``epochs x steps``) is ONE kernel launch;
* ``stats``: the sBN statistics pass (reference ``src/train_classifier_fed.py:127-138``) over ghost batches;
* ``eval``: inference with running statistics, logits out.

Network semantics follow reference ``src/models/resnet.py:10-50,104-158`` (Block: ``relu(n1(scaler(x)))`` ->
[shortcut] -> conv1 -> ``relu(n2(scaler(.)))`` -> conv2 -> add); the Scaler is folded into epsilon (``ops/nn.py``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from ..ops import cuda as cu
from ..ops.nn import BN_EPS

NIV, NFV, NPV, NTM = 44, 16, 16, 4
PHASE_DT = np.dtype([("op", "<i4"), ("ntiles", "<i4"), ("sync", "<i4"), ("flags", "<i4"), ("iv", "<i4", (NIV,)),
                     ("fv", "<f4", (NFV,)), ("pv", "<u8", (NPV,)), ("tm", "u1", (NTM, 128))])
assert PHASE_DT.itemsize == 896

(OP_NOP, OP_PREP, OP_CONV, OP_WGRAD, OP_NAPPLY, OP_NBWD_RED, OP_NBWD_APPLY, OP_HEAD, OP_SGD_NORM, OP_SGD_UPDATE,
 OP_CFIN) = range(11)
OP_NAMES = ["nop", "prep", "conv", "wgrad", "napply", "nbwd_red", "nbwd_apply", "head", "sgd_norm", "sgd_update", "cfin"]
CTL_WORDS = 16
ITEM_ELEMS4 = 256 * 8          # float4 elements per optimiser work item


def rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _i32(v: int) -> int:
    return int(np.array(v & 0xFFFFFFFF, dtype=np.uint32).view(np.int32))


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else int(t.data_ptr())


def _encode_tmap(ptr: int, dims, strides, box, estr) -> np.ndarray:
    """128-byte CUtensorMap; zeros on machines without a GPU driver (program structure is still testable there)."""
    assert ptr % 16 == 0 and all(s % 16 == 0 for s in strides), (ptr, strides)
    assert all(1 <= b <= 256 for b in box), box
    if not cu.available():
        return np.zeros(128, dtype=np.uint8)
    return cu.C().mega_encode_tmap(int(ptr), [int(d) for d in dims], [int(s) for s in strides], [int(b) for b in box],
                                   [int(e) for e in estr]).numpy()


def act_map(t: torch.Tensor, bw: int, bh: int, bn: int, stride: int = 1) -> np.ndarray:
    """Tensor map of a ``[N, H, W, Cp]`` bf16 activation: box = 64 channels x (bw, bh, bn) pixels, traversal ``stride``."""
    N, H, W, Cp = t.shape
    return _encode_tmap(t.data_ptr(), [Cp, W, H, N], [Cp * 2, W * Cp * 2, H * W * Cp * 2],
                        [64, bw * stride, bh * stride, bn], [1, stride, stride, 1])


class WDesc:
    """A conv filter stored ``[Kout][taps][Cin_p]`` bf16 (the K,R,S,C arena order; ``Cin_p`` = channel stride)."""

    def __init__(self, ptr: int, Kout: int, taps: int, Cin: int, Cin_p: int):
        self.ptr, self.Kout, self.taps, self.Cin, self.Cin_p = int(ptr), Kout, taps, Cin, Cin_p

    def kmajor_map(self, block_n: int) -> np.ndarray:       # fprop B operand: rows = output channels
        return _encode_tmap(self.ptr, [self.Cin_p, self.taps, self.Kout], [self.Cin_p * 2, self.taps * self.Cin_p * 2],
                            [64, 1, block_n], [1, 1, 1])

    def mn_map(self) -> np.ndarray:                          # dgrad B operand: 64(c) x 64(k) boxes
        return _encode_tmap(self.ptr, [self.Cin_p, self.taps, self.Kout], [self.Cin_p * 2, self.taps * self.Cin_p * 2],
                            [64, 1, 64], [1, 1, 1])


def brick(OH: int, OW: int, rows: int, group_n: int = 0) -> Tuple[int, int, int]:
    """(bn, bh, bw) pixel brick of at most ``rows`` pixels with bw == OW; group-aligned when ``group_n`` is given."""
    assert rows % OW == 0, (OW, rows)
    r = rows // OW
    if r <= OH:
        assert OH % r == 0, (OH, r)
        return 1, r, OW
    assert r % OH == 0, (OH, r)
    bn = r // OH
    if group_n:
        while group_n % bn:
            bn -= 1
    return bn, OH, OW


class Program:
    def __init__(self):
        self.recs: List[np.ndarray] = []
        self.keep: List[torch.Tensor] = []
        self.names: List[str] = []

    def add(self, op: int, ntiles: int, sync: int = 1, iv=(), fv=(), pv=(), tm=(), name: str = "", flags: int = 0) -> None:
        r = np.zeros((), dtype=PHASE_DT)
        r["op"], r["ntiles"], r["sync"], r["flags"] = op, int(ntiles), int(sync), int(flags)
        assert len(iv) <= NIV and len(fv) <= NFV and len(pv) <= NPV and len(tm) <= NTM
        for i, v in enumerate(iv):
            r["iv"][i] = int(v)
        for i, v in enumerate(fv):
            r["fv"][i] = float(v)
        for i, v in enumerate(pv):
            r["pv"][i] = int(v)
        for i, m in enumerate(tm):
            if m is not None:
                r["tm"][i] = m
        self.recs.append(r)
        self.names.append(name or OP_NAMES[op])

    def __len__(self) -> int:
        return len(self.recs)

    def n_barriers(self) -> int:
        return sum(int(r["sync"]) for r in self.recs)

    def hold(self, *objs) -> None:
        """Keep buffers referenced by raw pointers alive as long as the program."""
        self.keep.extend(objs)

    def finalize(self, device) -> torch.Tensor:
        arr = np.stack(self.recs) if self.recs else np.zeros((0,), dtype=PHASE_DT)
        t = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy())
        return t.to(device)


def launch(jobs, nphase: int, grid: int, ctl: torch.Tensor, prof: Optional[torch.Tensor] = None) -> None:
    """One cooperative launch of ``csrc/mega_kernel.cu`` over ``jobs`` = [(prog_dev, nvalid_tab | None, jctl, nsteps)]:
    the jobs advance through their (structurally identical) phase programs in lock step and share every grid barrier;
    a job idles once its own ``nsteps`` are done.  Heaviest job first gives the best tile balance."""
    rows = [[int(pd.data_ptr()), 0 if nv is None else int(nv.data_ptr()), int(jc.data_ptr()), int(ns)]
            for (pd, nv, jc, ns) in jobs]
    jt = torch.tensor(rows, dtype=torch.int64).to(ctl.device, non_blocking=True)
    cu.C().mega_launch(jt, int(nphase), max(int(r[3]) for r in rows), ctl, int(grid), prof)


# --------------------------------------------------------------------------------------------------------------
# phase emitters
# --------------------------------------------------------------------------------------------------------------
def _tap(a_map: int, b_map: int, b_tap: int, kblocks: int, dh: int, dw: int) -> Tuple[int, int]:
    assert 0 <= b_tap < 256 and 0 < kblocks < 65536
    return (_i32(a_map | (b_map << 4) | (b_tap << 8) | (kblocks << 16)), _i32((dh & 0xFFFF) | ((dw & 0xFFFF) << 16)))


def _conv_phase(prog: Program, *, N: int, OH: int, OW: int, out: torch.Tensor, taps: List[Tuple[int, int]],
                a_maps: List[np.ndarray], b_maps: List[np.ndarray], bnhw: Tuple[int, int, int], block_n: int, b_mn: int,
                in_stride: int, out_s: int = 1, out_ph: int = 0, out_pw: int = 0, res: Optional[torch.Tensor] = None,
                stats: Optional[torch.Tensor] = None, group_n: int = 0, use_nvalid: int = 0, sync: int = 1,
                name: str = "conv", bnred: Optional[Dict] = None, ksplit: int = 1,
                slices: Optional[torch.Tensor] = None) -> None:
    bn, bh, bw = bnhw
    Cy = out.shape[3]
    out_H, out_W = out.shape[1], out.shape[2]
    tiles_h = OH // bh
    m_tiles = ((N + bn - 1) // bn) * tiles_h
    n_tiles = (Cy + block_n - 1) // block_n
    num_k = sum((w0 >> 16) & 0xFFFF for w0, _ in taps)
    a_bytes = bn * bh * bw * 128
    b_bytes = block_n * 128
    assert len(taps) <= 9 and a_bytes <= 16384 and b_bytes <= 16384
    iv = [0] * NIV
    iv[0:26] = [N, OH, OW, bn, bh, bw, tiles_h, m_tiles, n_tiles, block_n, Cy, out_H, out_W, out_s, out_ph, out_pw,
                in_stride, b_mn, len(taps), int(res is not None), int(stats is not None), group_n, num_k, a_bytes,
                b_bytes, use_nvalid]
    for i, (w0, w1) in enumerate(taps):
        iv[26 + 2 * i], iv[27 + 2 * i] = w0, w1
    if res is not None:
        assert res.shape == out.shape
    tm = [a_maps[0], a_maps[1] if len(a_maps) > 1 else None, b_maps[0], b_maps[1] if len(b_maps) > 1 else None]
    base = m_tiles * n_tiles
    pv = [_ptr(out), _ptr(res), _ptr(stats), _ptr(slices), 0, 0, 0, 0, 0, 0, 0, 0]
    fv = []
    if ksplit > 1:               # partial sums go to `slices`; residual / statistics / BN-backward run in OP_CFIN
        assert res is None and stats is None and bnred is None and slices is not None
        per = -(-num_k // ksplit)
        assert (ksplit - 1) * per < num_k, (num_k, ksplit)
    if bnred is not None:       # BatchNorm-backward reductions of the layer that produced bnred["y"], in this epilogue
        assert res is None and stats is None and bnred["y"].shape == out.shape
        pv[4:9] = [_ptr(bnred["y"]), _ptr(bnred["S"]), bnred["gamma"], bnred["beta"], _ptr(bnred["T"])]
        fv = [bnred["eps"]]
    prog.add(OP_CONV, base * ksplit, sync, iv=iv, fv=fv, pv=pv, tm=tm, name=name, flags=ksplit)


def pick_ksplit(base_tiles: int, num_ks: List[int], split_target: int) -> int:
    """Split-K factor for a (group of) conv phase(s) with `base_tiles` output tiles and `num_ks` k-blocks per tile."""
    import os as _os
    if not split_target or base_tiles * 2 > split_target or min(num_ks) < 8:
        return 1
    k = min(-(-split_target // base_tiles), min(num_ks) // 2, int(_os.environ.get("HFL_MEGA_KSPLIT_MAX", "12")))
    while k > 1 and any((k - 1) * (-(-nk // k)) >= nk for nk in num_ks):
        k -= 1
    return max(k, 1)


def conv_finalize(prog: Program, slices: torch.Tensor, ksplit: int, out: torch.Tensor, C: int, *, res=None, stats=None,
                  bnred: Optional[Dict] = None, use_nvalid: int = 0, sync: int = 1, name: str = "cfin") -> None:
    """Sum the split-K partial slices into ``out`` (+ residual), with the BatchNorm statistics (or the BatchNorm-backward
    reductions) the conv epilogue would have produced."""
    N, H, W_, Cp = out.shape
    M, HW = N * H * W_, H * W_
    if ksplit <= 1:              # structural placeholder (phase programs of all widths must line up): no work, and the
        prog.add(OP_CFIN, 0, sync, name=name)      # kernel skips the barrier of an empty phase that follows a barrier
        return
    rstep = 256 // (Cp // 8)
    rpi = max(rstep, -(-M // 96) // rstep * rstep)
    iv = [M, Cp, C, HW, ksplit, int(res is not None), int(bnred is not None), int(stats is not None), rpi, use_nvalid, N]
    pv = [_ptr(out), _ptr(res), _ptr(stats), _ptr(slices), 0, 0, 0, 0, 0]
    fv = [0.0]
    if bnred is not None:
        pv[4:9] = [_ptr(bnred["y"]), _ptr(bnred["S"]), bnred["gamma"], bnred["beta"], _ptr(bnred["T"])]
        fv = [bnred["eps"]]
    prog.add(OP_CFIN, (M + rpi - 1) // rpi, sync, iv=iv, fv=fv, pv=pv, name=name)


def _block_n(Cy: int) -> int:
    return 128 if Cy >= 128 else 64


def conv_fprop(prog: Program, x: torch.Tensor, w: WDesc, out: torch.Tensor, R: int, stride: int,
               res: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None, group_n: int = 0,
               use_nvalid: int = 0, sync: int = 1, rows: int = 128, name: str = "fprop", split_target: int = 0,
               splittable: bool = False) -> None:
    """``out = conv(x, w) (+ res)``, optional per-channel (sum, sumsq) of the bf16 outputs into ``stats``."""
    N, H, W_, Cxp = x.shape
    _, OH, OW, Cy = out.shape
    pad = R // 2
    assert OH == (H + 2 * pad - R) // stride + 1 and w.Cin_p == Cxp and w.taps == R * R
    bnhw = brick(OH, OW, rows, group_n)
    bn, bh, bw = bnhw
    bnk = _block_n(Cy)
    kb = (Cxp + 63) // 64
    taps = [_tap(0, 0, r * R + s, kb, r - pad, s - pad) for r in range(R) for s in range(R)]
    base = ((N + bn - 1) // bn) * (OH // bh) * ((Cy + bnk - 1) // bnk)
    ksplit = pick_ksplit(base, [kb * R * R], split_target) if splittable else 1
    common = dict(N=N, OH=OH, OW=OW, out=out, taps=taps, a_maps=[act_map(x, bw, bh, bn, stride)],
                  b_maps=[w.kmajor_map(bnk)], bnhw=bnhw, block_n=bnk, b_mn=0, in_stride=stride, group_n=group_n,
                  use_nvalid=use_nvalid, name=name)
    if ksplit == 1:
        _conv_phase(prog, res=res, stats=stats, sync=sync, **common)
        if splittable:
            assert sync == 1
            conv_finalize(prog, None, 1, out, w.Kout, sync=1, name=name + ".fin")
        return
    assert splittable
    slices = torch.zeros(ksplit, N * OH * OW * Cy, dtype=torch.float32, device=out.device)
    prog.keep.append(slices)
    _conv_phase(prog, sync=1, ksplit=ksplit, slices=slices, **common)
    conv_finalize(prog, slices, ksplit, out, w.Kout, res=res, stats=stats, use_nvalid=use_nvalid, sync=sync,
                  name=name + ".fin")


def conv_dgrad(prog: Program, dy: torch.Tensor, w: WDesc, dx: torch.Tensor, R: int, stride: int,
               dy2: Optional[torch.Tensor] = None, w2: Optional[WDesc] = None, sync: int = 1,
               name: str = "dgrad", split_target: int = 0, bnred: Optional[Dict] = None, use_nvalid: int = 1,
               splittable: bool = False) -> None:
    """``dx = conv_transpose(dy, w)`` for a 3x3 / pad-1 (or 1x1 / pad-0) filter of stride 1 or 2; for stride 2 an optional
    second gradient ``dy2`` through a 1x1 stride-2 filter ``w2`` (the block shortcut) is accumulated in the same pass."""
    N, OH, OW, Kp = dy.shape
    _, H, W_, Cxp = dx.shape
    pad = R // 2
    assert w.Cin_p == Cxp and H == OH * stride and W_ == OW * stride
    bnk = _block_n(Cxp)
    kb = (Kp + 63) // 64
    n_tiles = (Cxp + bnk - 1) // bnk
    if stride == 1:
        bnhw = brick(H, W_, 128)
        bn, bh, bw = bnhw
        taps = [_tap(0, 0, r * R + s, kb, pad - r, pad - s) for r in range(R) for s in range(R)]
        base = ((N + bn - 1) // bn) * (H // bh) * n_tiles
        ksplit = pick_ksplit(base, [kb * R * R], split_target) if splittable else 1
        common = dict(N=N, OH=H, OW=W_, out=dx, taps=taps, a_maps=[act_map(dy, bw, bh, bn)], b_maps=[w.mn_map()],
                      bnhw=bnhw, block_n=bnk, b_mn=1, in_stride=1, use_nvalid=use_nvalid, name=name)
        if ksplit == 1:
            _conv_phase(prog, sync=sync, bnred=bnred, **common)
            if splittable:
                assert sync == 1
                conv_finalize(prog, None, 1, dx, w.Cin, sync=1, name=name + ".fin")
        else:
            assert splittable
            slices = torch.zeros(ksplit, dx.numel(), dtype=torch.float32, device=dx.device)
            prog.keep.append(slices)
            _conv_phase(prog, sync=1, ksplit=ksplit, slices=slices, **common)
            conv_finalize(prog, slices, ksplit, dx, w.Cin, bnred=bnred, use_nvalid=use_nvalid, sync=sync,
                          name=name + ".fin")
        return
    assert stride == 2 and R == 3
    bnhw = brick(OH, OW, 128)
    bn, bh, bw = bnhw
    a_maps = [act_map(dy, bw, bh, bn)]
    b_maps = [w.mn_map()]
    if dy2 is not None:
        assert dy2.shape[:3] == dy.shape[:3] and w2 is not None and w2.Cin_p == Cxp and w2.taps == 1
        a_maps.append(act_map(dy2, bw, bh, bn))
        b_maps.append(w2.mn_map())
    sel = {0: [(1, 0)], 1: [(0, 1), (2, 0)]}      # parity -> [(filter index, dY offset)]
    groups = []
    for ph in (0, 1):
        for pw in (0, 1):
            taps = [_tap(0, 0, r * 3 + s, kb, di, dj) for (r, di) in sel[ph] for (s, dj) in sel[pw]]
            if dy2 is not None and ph == 0 and pw == 0:
                taps.append(_tap(1, 1, 0, (dy2.shape[3] + 63) // 64, 0, 0))
            groups.append((ph, pw, taps))
    base = ((N + bn - 1) // bn) * (OH // bh) * n_tiles
    # the four parity classes write disjoint pixels of dx: one split factor, one slice buffer, one finalisation
    ksplit = pick_ksplit(4 * base, [sum((t[0] >> 16) & 0xFFFF for t in g[2]) for g in groups], split_target) \
        if splittable else 1
    slices = None
    if ksplit > 1:
        slices = torch.zeros(ksplit, dx.numel(), dtype=torch.float32, device=dx.device)
        prog.keep.append(slices)
    for gi, (ph, pw, taps) in enumerate(groups):
        last = gi == 3
        _conv_phase(prog, N=N, OH=OH, OW=OW, out=dx, taps=taps, a_maps=a_maps, b_maps=b_maps, bnhw=bnhw,
                    block_n=bnk, b_mn=1, in_stride=1, out_s=2, out_ph=ph, out_pw=pw,
                    sync=((1 if ksplit > 1 else sync) if last else 0), name="%s.p%d%d" % (name, ph, pw),
                    bnred=bnred if ksplit == 1 else None, use_nvalid=use_nvalid, ksplit=ksplit, slices=slices)
    if ksplit > 1 or splittable:
        assert ksplit == 1 or splittable
        conv_finalize(prog, slices, ksplit, dx, w.Cin, bnred=bnred, use_nvalid=use_nvalid, sync=sync, name=name + ".fin")


def conv_wgrad(prog: Program, dy: torch.Tensor, x: torch.Tensor, dW_ptr: int, Kout: int, C: int, R: int, stride: int,
               target_items: int, sync: int = 0, name: str = "wgrad") -> None:
    """``dW[k, tap, c] += sum_pixels dy[.., k] * x[.. shifted by tap .., c]`` (fp32 atomics into the gradient arena)."""
    N, OH, OW, Kp = dy.shape
    _, H, W_, Cxp = x.shape
    pad = R // 2
    bn, bh, bw = brick(OH, OW, 64)
    bricks_h = OH // bh
    bricks = ((N + bn - 1) // bn) * bricks_h
    k_blks, c_blks = (Kout + 127) // 128, (C + 63) // 64
    base = k_blks * c_blks * R * R
    splits = 1
    while base * splits < target_items and bricks // (splits * 2) >= 2:
        splits *= 2
    iv = [N, OH, OW, bn, bh, bw, bricks_h, bricks, splits, C, Kout, R * R, R, stride, pad, k_blks, c_blks,
          2 if Kout > 64 else 1]
    prog.add(OP_WGRAD, base * splits, sync, iv=iv, pv=[dW_ptr],
             tm=[act_map(dy, bw, bh, bn), act_map(x, bw, bh, bn, stride)], name=name)


def _rows_per_item(rows_total: int, Cp: int, group_rows: int = 0) -> int:
    rstep = 256 // (Cp // 8)
    want = max(rstep * 8, 64)
    if group_rows:
        k = max(1, -(-group_rows // want))
        while group_rows % k:
            k += 1
        return group_rows // k
    return min(want, max(rows_total, 1))


def norm_apply(prog: Program, y: torch.Tensor, z: torch.Tensor, C: int, *, stats=None, gamma: int = 0, beta: int = 0,
               eps: float = BN_EPS, relu: bool = True, group_n: int = 0, acc=None, inv_s: float = 1.0, affine=None,
               use_nvalid: int = 0, sync: int = 1, name: str = "napply") -> None:
    N, H, W_, Cp = y.shape
    assert Cp & (Cp - 1) == 0 and 8 <= Cp <= 2048, Cp
    M, HW = N * H * W_, H * W_
    rpi = _rows_per_item(M, Cp, group_n * HW if group_n else 0)
    iv = [M, Cp, C, HW, 0 if affine is None else 1, int(relu), group_n, int(acc is not None), rpi, use_nvalid, N]
    pv = [_ptr(y), _ptr(z), _ptr(stats), gamma, beta, _ptr(acc[0]) if acc else 0, _ptr(acc[1]) if acc else 0,
          _ptr(affine[0]) if affine else 0, _ptr(affine[1]) if affine else 0]
    prog.add(OP_NAPPLY, (M + rpi - 1) // rpi, sync, iv=iv, fv=[eps, inv_s, inv_s * inv_s], pv=pv, name=name)


def norm_bwd(prog: Program, y: torch.Tensor, dz: torch.Tensor, dy: torch.Tensor, C: int, *, stats, gamma: int,
             beta: int, T: torch.Tensor, dgamma: int, dbeta: int, eps: float, dz2=None, dres=None, relu: bool = True,
             use_nvalid: int = 1, name: str = "nbwd", fused_red: bool = False) -> None:
    N, H, W_, Cp = y.shape
    M, HW = N * H * W_, H * W_
    rpi = _rows_per_item(M, Cp)
    iv = [M, Cp, C, HW, 0, int(relu), 0, 0, rpi, use_nvalid, N]
    pv = [_ptr(y), _ptr(dz), _ptr(stats), gamma, beta, _ptr(dz2), _ptr(T), _ptr(dres), _ptr(dy), dgamma, dbeta]
    items = (M + rpi - 1) // rpi
    if not fused_red:        # otherwise the producing dgrad's epilogue already masked dz and accumulated T
        prog.add(OP_NBWD_RED, items, 1, iv=iv, fv=[eps], pv=pv, name=name + ".red")
    prog.add(OP_NBWD_APPLY, items, 1, iv=iv, fv=[eps], pv=pv, name=name + ".apply")


def zero_phase(prog: Program, buf: torch.Tensor, sync: int = 1, name: str = "zero") -> None:
    assert buf.dtype == torch.float32 and buf.numel() % 4 == 0
    n4 = buf.numel() // 4
    items = (n4 + ITEM_ELEMS4 - 1) // ITEM_ELEMS4
    prog.add(OP_SGD_UPDATE, items, sync, iv=[0, 0, n4], pv=[0, 0, 0, 0, _ptr(buf)], name=name)


# --------------------------------------------------------------------------------------------------------------
class MegaResNet:
    """Buffers + phase programs of one ResNet sub-network (one complexity level) on one device."""

    def __init__(self, model, device, grid: int):
        spec = model.spec
        assert spec.family == "resnet" and spec.meta["norm"] == "bn" and spec.meta["kind"] == "block"
        self.model, self.spec, self.device, self.grid = model, spec, torch.device(device), int(grid)
        self.lay = model.layout
        assert self.lay.channels_last, "device arenas store filters K,R,S,C"
        self.blocks = spec.meta["blocks"]
        self.scale_rate = float(model.scale_rate)
        self.eps = BN_EPS * self.scale_rate * self.scale_rate
        ds = spec.meta["data_shape"]
        self.Cin, self.H, self.W = int(ds[0]), int(ds[1]), int(ds[2])
        self.classes = int(spec.meta["classes"])
        # conv filters whose input-channel count is not a multiple of 8 get a re-packed bf16 copy (16-byte TMA strides)
        self.pack_rows: List[Tuple[int, int, int, int, int]] = []
        self.pack_off: Dict[str, int] = {}
        off = 0
        for name, e in self.lay.entries.items():
            if len(e.shape) == 4 and e.shape[1] % 8:
                K, C, R, S = e.shape
                Cp = rup(C, 8)
                self.pack_rows.append((e.offset, off, K * R * S, C, Cp))
                self.pack_off[name] = off
                off += rup(K * R * S * Cp, 8)
        self.pack_size = max(off, 8)

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _act(self, N: int, H: int, W: int, C: int) -> torch.Tensor:
        return torch.zeros(N, H, W, rup(C, 8), dtype=torch.bfloat16, device=self.device)

    def _wdesc(self, wc: torch.Tensor, packbuf: torch.Tensor, name: str) -> WDesc:
        e = self.lay.entries[name]
        K, C, R, S = e.shape
        if name in self.pack_off:
            return WDesc(packbuf.data_ptr() + 2 * self.pack_off[name], K, R * S, C, rup(C, 8))
        return WDesc(wc.data_ptr() + 2 * e.offset, K, R * S, C, C)

    def _pptr(self, arena: torch.Tensor, name: str) -> int:
        return int(arena.data_ptr()) + arena.element_size() * self.lay.entries[name].offset

    def widths(self) -> List[int]:
        return [self.lay.entries[b["prefix"] + ".conv1.weight"].shape[0] for b in self.blocks]

    def norm_sites(self) -> List[Tuple[str, int]]:
        out = []
        for b in self.blocks:
            out.append((b["prefix"] + ".n1", self.lay.entries[b["prefix"] + ".n1.weight"].shape[0]))
            out.append((b["prefix"] + ".n2", self.lay.entries[b["prefix"] + ".n2.weight"].shape[0]))
        out.append(("n4", self.lay.entries["n4.weight"].shape[0]))
        return out

    # ---- forward emitter shared by the three programs -----------------------------------------------------------
    def _forward(self, prog: Program, st: Dict, N: int, mode: str, group_n: int, use_nvalid: int) -> torch.Tensor:
        """Emit stem + blocks + n4.  ``st`` carries wc / packbuf / stats slabs / accumulators; returns z4."""
        wc, packbuf = st["wc"], st["packbuf"]
        G = (N // group_n) if group_n else 1
        keep = st.setdefault("acts", {})

        def stats_for(site: str, C: int) -> Optional[torch.Tensor]:
            if mode == "eval":
                return None
            return st["S"][site]

        def napply(site: str, y: torch.Tensor, z: torch.Tensor, C: int):
            if mode == "eval":
                norm_apply(prog, y, z, C, affine=st["affine"][site], name=site + ".affine")
            else:
                norm_apply(prog, y, z, C, stats=st["S"][site], gamma=self._pptr(wc, site + ".weight"),
                           beta=self._pptr(wc, site + ".bias"), eps=self.eps, group_n=group_n,
                           acc=st["acc"][site] if mode == "stats" else None, inv_s=1.0 / self.scale_rate,
                           use_nvalid=use_nvalid, name=site + ".apply")

        x = st["x"]
        H, W = self.H, self.W
        C0 = self.lay.entries["conv1.weight"].shape[0]
        first = self.blocks[0]["prefix"] + ".n1"
        y = keep.setdefault("y0", self._act(N, H, W, C0))
        split = int(st.get("split", 0)) if mode == "train" else 0
        conv_fprop(prog, x, self._wdesc(wc, packbuf, "conv1.weight"), y, 3, 1, stats=stats_for(first, C0),
                   group_n=group_n, use_nvalid=use_nvalid, name="stem")
        cin = C0
        for bi, blk in enumerate(self.blocks):
            pre, stride = blk["prefix"], blk["stride"]
            planes = self.lay.entries[pre + ".conv1.weight"].shape[0]
            OH, OW = H // stride, W // stride
            z1 = keep.setdefault(pre + ".z1", self._act(N, H, W, cin))
            napply(pre + ".n1", y, z1, cin)
            if blk["shortcut"]:
                sc = keep.setdefault(pre + ".sc", self._act(N, OH, OW, planes))
                conv_fprop(prog, z1, self._wdesc(wc, packbuf, pre + ".shortcut.weight"), sc, 1, stride, sync=0,
                           group_n=group_n, name=pre + ".shortcut")
                res = sc
            else:
                res = y
            c1 = keep.setdefault(pre + ".c1", self._act(N, OH, OW, planes))
            conv_fprop(prog, z1, self._wdesc(wc, packbuf, pre + ".conv1.weight"), c1, 3, stride,
                       stats=stats_for(pre + ".n2", planes), group_n=group_n, use_nvalid=use_nvalid, name=pre + ".conv1",
                       split_target=split, splittable=(mode == "train" and OH <= 16))
            z2 = keep.setdefault(pre + ".z2", self._act(N, OH, OW, planes))
            napply(pre + ".n2", c1, z2, planes)
            nxt = (self.blocks[bi + 1]["prefix"] + ".n1") if bi + 1 < len(self.blocks) else "n4"
            yo = keep.setdefault(pre + ".out", self._act(N, OH, OW, planes))
            conv_fprop(prog, z2, self._wdesc(wc, packbuf, pre + ".conv2.weight"), yo, 3, 1, res=res,
                       stats=stats_for(nxt, planes), group_n=group_n, use_nvalid=use_nvalid, name=pre + ".conv2",
                       split_target=split, splittable=(mode == "train" and OH <= 16))
            st.setdefault("yin", {})[pre] = y
            y, cin, H, W = yo, planes, OH, OW
        z4 = keep.setdefault("z4", self._act(N, H, W, cin))
        napply("n4", y, z4, cin)
        st["y_last"], st["c_last"], st["hw_last"] = y, cin, (H, W)
        return z4

    def _prep(self, prog: Program, st: Dict, N: int, augment: bool, with_pack: bool, sync: int = 1) -> None:
        iv = [N, self.H, self.W, self.Cin, 4, len(self.pack_rows) if with_pack else 0, int(augment), 8]
        pv = [_ptr(st["src"]), _ptr(st["idx"]), _ptr(st.get("par")) if augment else 0, _ptr(st["mean"]), _ptr(st["istd"]),
              _ptr(st["x"]), _ptr(st.get("labels")), _ptr(st.get("lab")), _ptr(st.get("pack_tab")), _ptr(st.get("w")),
              _ptr(st["packbuf"])]
        prog.add(OP_PREP, N + (len(self.pack_rows) if with_pack else 0), sync, iv=iv, pv=pv, name="prep")

    def _slabs(self, G: int) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        sites = self.norm_sites()
        tot = sum(2 * G * rup(C, 8) for _, C in sites)
        buf = torch.zeros(rup(tot, 4), dtype=torch.float32, device=self.device)
        out, off = {}, 0
        for name, C in sites:
            n = 2 * G * rup(C, 8)
            out[name] = buf[off:off + n]
            off += n
        return buf, out

    # ---- programs ------------------------------------------------------------------------------------------------
    def build_train(self, st: Dict, N: int, with_opt: bool = True) -> Program:
        """``st`` must provide: w, wc, g (fp32 grad arena), mom, lr (unused: lr is baked per launch through fv), src,
        labels, idx, par, mean, istd, mask, meter; buffers for activations are created here."""
        prog = Program()
        dev = self.device
        st["x"] = torch.zeros(N, self.H, self.W, 8, dtype=torch.bfloat16, device=dev)
        st["lab"] = torch.zeros(max(N, 4), dtype=torch.int32, device=dev)
        st["packbuf"] = st.get("packbuf", torch.zeros(self.pack_size, dtype=torch.bfloat16, device=dev))
        st["pack_tab"] = torch.tensor(self.pack_rows or [(0, 0, 0, 1, 8)], dtype=torch.int64, device=dev)
        sbuf, st["S"] = self._slabs(1)
        tbuf, st["T"] = self._slabs(1)
        st["scr"] = (sbuf, tbuf)
        wc, w, g = st["wc"], st["w"], st["g"]
        tgt = 2 * self.grid
        import os as _os
        st.setdefault("split", int(_os.environ.get("HFL_MEGA_SPLIT", self.grid)))
        st.setdefault("fuse_red", _os.environ.get("HFL_MEGA_FUSE_RED", "1") != "0")
        split = int(st["split"])
        self._prep(prog, st, N, augment=True, with_pack=True)
        z4 = self._forward(prog, st, N, "train", 0, 1)
        y_last, C4, (H4, W4) = st["y_last"], st["c_last"], st["hw_last"]
        # ---- head ----
        dz4 = self._act(N, H4, W4, C4)
        st["loss"] = torch.zeros(4, dtype=torch.float32, device=dev)
        iv = [N, H4 * W4, z4.shape[3], C4, self.classes, 1, 0, 0, 0, 1]
        pv = [_ptr(z4), self._pptr(wc, "linear.weight"), self._pptr(wc, "linear.bias"), _ptr(st["lab"]),
              _ptr(st.get("mask")), _ptr(st.get("meter")), self._pptr(g, "linear.weight"), self._pptr(g, "linear.bias"),
              _ptr(dz4), 0, _ptr(st["loss"])]
        Cp4 = z4.shape[3]
        assert (N * Cp4 + N * self.classes + 3 * N + 2 * self.classes + 8) * 4 + self.classes * Cp4 * 2 <= 150 * 1024, \
            "head does not fit"
        prog.add(OP_HEAD, 1, 1, iv=iv, pv=pv, name="head")
        # ---- backward ----
        acts = st["acts"]
        bk = st.setdefault("bwd", {})
        ds = bk.setdefault("n4.ds", self._act(N, H4, W4, C4))      # every buffer a phase points at must outlive the build
        norm_bwd(prog, y_last, dz4, ds, C4, stats=st["S"]["n4"], gamma=self._pptr(wc, "n4.weight"),
                 beta=self._pptr(wc, "n4.bias"), T=st["T"]["n4"], dgamma=self._pptr(g, "n4.weight"),
                 dbeta=self._pptr(g, "n4.bias"), eps=self.eps, name="n4.bwd")
        H, W = H4, W4
        for blk in reversed(self.blocks):
            pre, stride = blk["prefix"], blk["stride"]
            e1 = self.lay.entries[pre + ".conv1.weight"]
            planes, cin = e1.shape[0], e1.shape[1]
            z1, c1, z2 = acts[pre + ".z1"], acts[pre + ".c1"], acts[pre + ".z2"]
            y_in = st["yin"][pre]
            IH, IW = H * stride, W * stride
            # conv2: wgrad (off the critical path) + dgrad
            conv_wgrad(prog, ds, z2, self._pptr(g, pre + ".conv2.weight"), planes, planes, 3, 1, tgt, name=pre + ".conv2.wgrad")
            dz2 = bk.setdefault(pre + ".dz2", self._act(N, H, W, planes))
            fuse = bool(st.get("fuse_red", True))

            def red_of(site: str, ybuf: torch.Tensor):
                return dict(y=ybuf, S=st["S"][site], gamma=self._pptr(wc, site + ".weight"),
                            beta=self._pptr(wc, site + ".bias"), T=st["T"][site], eps=self.eps) if fuse else None

            conv_dgrad(prog, ds, self._wdesc(wc, st["packbuf"], pre + ".conv2.weight"), dz2, 3, 1, name=pre + ".conv2.dgrad",
                       split_target=split, bnred=red_of(pre + ".n2", c1), splittable=(H <= 16))
            dc1 = bk.setdefault(pre + ".dc1", self._act(N, H, W, planes))
            norm_bwd(prog, c1, dz2, dc1, planes, stats=st["S"][pre + ".n2"], gamma=self._pptr(wc, pre + ".n2.weight"),
                     beta=self._pptr(wc, pre + ".n2.bias"), T=st["T"][pre + ".n2"],
                     dgamma=self._pptr(g, pre + ".n2.weight"), dbeta=self._pptr(g, pre + ".n2.bias"), eps=self.eps,
                     name=pre + ".n2.bwd", fused_red=fuse)
            conv_wgrad(prog, dc1, z1, self._pptr(g, pre + ".conv1.weight"), planes, cin, 3, stride, tgt,
                       name=pre + ".conv1.wgrad")
            dz1 = bk.setdefault(pre + ".dz1", self._act(N, IH, IW, cin))
            if blk["shortcut"]:
                conv_wgrad(prog, ds, z1, self._pptr(g, pre + ".shortcut.weight"), planes, cin, 1, stride, tgt,
                           name=pre + ".shortcut.wgrad")
                wsc = self._wdesc(wc, st["packbuf"], pre + ".shortcut.weight")
                if stride == 2:
                    conv_dgrad(prog, dc1, self._wdesc(wc, st["packbuf"], pre + ".conv1.weight"), dz1, 3, 2, dy2=ds,
                               w2=wsc, name=pre + ".conv1.dgrad", split_target=split, bnred=red_of(pre + ".n1", y_in),
                               splittable=(IH <= 16))
                    dz1b = None
                    fuse1 = fuse
                else:        # widening block without down-sampling (not in resnet18/34, kept for completeness)
                    conv_dgrad(prog, dc1, self._wdesc(wc, st["packbuf"], pre + ".conv1.weight"), dz1, 3, 1, sync=0,
                               name=pre + ".conv1.dgrad")
                    dz1b = bk.setdefault(pre + ".dz1b", self._act(N, IH, IW, cin))
                    conv_dgrad(prog, ds, wsc, dz1b, 1, 1, name=pre + ".shortcut.dgrad")
                    fuse1 = False
                dres = None
            else:
                conv_dgrad(prog, dc1, self._wdesc(wc, st["packbuf"], pre + ".conv1.weight"), dz1, 3, stride,
                           name=pre + ".conv1.dgrad", split_target=split, bnred=red_of(pre + ".n1", y_in),
                           splittable=(IH <= 16))
                dz1b, dres, fuse1 = None, ds, fuse
            ds_prev = bk.setdefault(pre + ".ds", self._act(N, IH, IW, cin))
            norm_bwd(prog, y_in, dz1, ds_prev, cin, stats=st["S"][pre + ".n1"], gamma=self._pptr(wc, pre + ".n1.weight"),
                     beta=self._pptr(wc, pre + ".n1.bias"), T=st["T"][pre + ".n1"],
                     dgamma=self._pptr(g, pre + ".n1.weight"), dbeta=self._pptr(g, pre + ".n1.bias"), eps=self.eps,
                     dz2=dz1b, dres=dres, name=pre + ".n1.bwd", fused_red=fuse1)
            ds, H, W = ds_prev, IH, IW
        C0 = self.lay.entries["conv1.weight"].shape[0]
        conv_wgrad(prog, ds, st["x"], self._pptr(g, "conv1.weight"), C0, self.Cin, 3, 1, tgt, sync=1, name="stem.wgrad")
        prog.keep += [sbuf, tbuf, dz4, st]
        if not with_opt:         # debugging / tests: gradients stay in ``g``, statistics slabs are not re-armed
            return prog
        # ---- optimiser ----
        n4 = self.lay.size // 4
        items = (n4 + ITEM_ELEMS4 - 1) // ITEM_ELEMS4
        prog.add(OP_SGD_NORM, items, 1, iv=[n4], pv=[_ptr(g)], name="sgd.norm")
        # statistics slabs are zeroed for the next step in the same phase (two slabs -> two zero ranges)
        zitems_s = (sbuf.numel() // 4 + ITEM_ELEMS4 - 1) // ITEM_ELEMS4
        prog.add(OP_SGD_UPDATE, items + zitems_s, 0, iv=[n4, items, sbuf.numel() // 4],
                 fv=[st["lr"], st["momentum"], st["wd"], st["max_norm"]],
                 pv=[_ptr(w), _ptr(wc), _ptr(g), _ptr(st["mom"]), _ptr(sbuf)], name="sgd.update")
        zero_phase(prog, tbuf, sync=1, name="zero.T")
        return prog

    def build_forward(self, st: Dict, N: int, mode: str, group_n: int = 0, augment: bool = False,
                      steps: int = 1) -> Program:
        """``mode``: "stats" (sBN pass over ghost batches of ``group_n`` images; accumulates running mean / unbiased
        variance sums into ``st['acc']``) or "eval" (affine from running statistics, logits into ``st['logits']``)."""
        assert mode in ("stats", "eval")
        prog = Program()
        dev = self.device
        st["x"] = torch.zeros(N, self.H, self.W, 8, dtype=torch.bfloat16, device=dev)
        st["packbuf"] = st.get("packbuf", torch.zeros(self.pack_size, dtype=torch.bfloat16, device=dev))
        st["pack_tab"] = torch.tensor(self.pack_rows or [(0, 0, 0, 1, 8)], dtype=torch.int64, device=dev)
        sites = self.norm_sites()
        if mode == "stats":
            assert group_n > 0 and N % group_n == 0
            sbuf, st["S"] = self._slabs(N // group_n)
            st["acc"] = {name: (torch.zeros(rup(C, 8), dtype=torch.float32, device=dev),
                                torch.zeros(rup(C, 8), dtype=torch.float32, device=dev)) for name, C in sites}
            zero_phase(prog, sbuf, sync=0, name="zero.S")
            prog.keep.append(sbuf)
        else:
            st["affine"] = {name: (torch.zeros(rup(C, 8), dtype=torch.float32, device=dev),
                                   torch.zeros(rup(C, 8), dtype=torch.float32, device=dev)) for name, C in sites}
        self._prep(prog, st, N, augment=augment, with_pack=True)
        z4 = self._forward(prog, st, N, mode, group_n if mode == "stats" else 0, 1 if mode == "stats" else 0)
        if mode == "eval":
            C4, (H4, W4) = st["c_last"], st["hw_last"]
            st["logits"] = torch.zeros(steps * N, self.classes, dtype=torch.float32, device=dev)   # row = step * N + n
            iv = [N, H4 * W4, z4.shape[3], C4, self.classes, 0, int(steps > 1)]
            pv = [_ptr(z4), self._pptr(st["wc"], "linear.weight"), self._pptr(st["wc"], "linear.bias"), 0, 0, 0, 0, 0, 0,
                  _ptr(st["logits"]), 0]
            prog.add(OP_HEAD, N, 1, iv=iv, pv=pv, name="head.eval")
        prog.hold(st)
        return prog
