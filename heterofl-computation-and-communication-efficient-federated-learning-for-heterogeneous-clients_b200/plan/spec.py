"""Slice algebra: declarative parameter specs for the three model families.

The reference derives every client's index sets by walking ``state_dict()`` keys and threading the previous
layer's output index into the next layer's input (``src/fed.py:26-159``).  Here each parameter instead *declares*
what its two axes are -- a fixed axis (data channels, classes, vocabulary, positions), a width axis (kept as a
prefix ``[:ceil(n*s)]``) or a per-head width axis (prefix inside each of H heads) -- and everything else (local
shapes, flat-arena offsets, index maps, coverage counts, kernel descriptor tables) is computed from that
declaration.  Key names and order equal the reference modules' ``state_dict()`` (``src/models/conv.py``,
``src/models/resnet.py``, ``src/models/transformer.py``) so checkpoints interchange.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

FIXED, PREFIX, HEADS = "fixed", "prefix", "heads"


def _ceil(x: float) -> int:
    return int(math.ceil(x))


@dataclass(frozen=True)
class Axis:
    """One sliceable axis of a parameter, described at *unit* (level-a) width."""
    kind: str          # FIXED | PREFIX | HEADS
    size: int          # size at rate 1.0 (for FIXED: the size, rate independent)
    heads: int = 1     # number of heads for HEADS axes

    def local(self, rate: float) -> int:
        """Size of this axis in a model built at absolute width ``rate``."""
        if self.kind == FIXED:
            return self.size
        if self.kind == PREFIX:
            return _ceil(rate * self.size)
        # per-head: the model is built with E=ceil(rate*E0) and d_head=E//H
        return self.heads * (_ceil(rate * self.size) // self.heads)

    def head_dim(self, rate: float) -> int:
        return self.local(rate) // self.heads


@dataclass(frozen=True)
class ParamSpec:
    name: str
    out: Axis                       # axis 0
    inp: Optional[Axis] = None      # axis 1 (None for vectors)
    inner: int = 1                  # product of trailing dims (kh*kw for convs)
    tail: Tuple[int, ...] = ()      # trailing dims themselves
    label_rows: bool = False        # aggregation restricted to the client's label rows (classifier/vocab)
    init: str = "default"           # initialisation recipe (see models.init)
    group: str = "param"            # "param" | "buffer"

    @property
    def ndim(self) -> int:
        return (1 if self.inp is None else 2) + len(self.tail)

    def shape(self, rate: float) -> Tuple[int, ...]:
        if self.inp is None:
            return (self.out.local(rate),)
        return (self.out.local(rate), self.inp.local(rate)) + tuple(self.tail)

    def numel(self, rate: float) -> int:
        n = 1
        for s in self.shape(rate):
            n *= s
        return n


@dataclass
class ModelSpec:
    family: str                     # "conv" | "resnet" | "transformer"
    name: str                       # "conv", "resnet18", ..., "transformer"
    params: List[ParamSpec]
    meta: Dict = field(default_factory=dict)
    federable: bool = True          # Bottleneck resnets are centralised-only (reference src/fed.py:74-89)

    def by_name(self) -> Dict[str, ParamSpec]:
        return {p.name: p for p in self.params}

    def num_params(self, rate: float) -> int:
        return sum(p.numel(rate) for p in self.params)


# ----------------------------------------------------------------------------------------------------------
# model families
# ----------------------------------------------------------------------------------------------------------
def _norm_specs(prefix: str, width: Axis, norm: str) -> List[ParamSpec]:
    if norm == "none":
        return []
    return [ParamSpec(prefix + ".weight", width, init="ones"), ParamSpec(prefix + ".bias", width, init="zeros")]


def conv_spec(data_shape: Sequence[int], hidden: Sequence[int], classes: int, norm: str) -> ModelSpec:
    """MNIST CNN -- key names of ``nn.Sequential`` blocks in reference ``src/models/conv.py:29-60``."""
    ps: List[ParamSpec] = []
    prev = Axis(FIXED, data_shape[0])
    for i, h in enumerate(hidden):
        w = Axis(PREFIX, h)
        base = 5 * i
        ps.append(ParamSpec("blocks.%d.weight" % base, w, prev, 9, (3, 3), init="conv"))
        ps.append(ParamSpec("blocks.%d.bias" % base, w, init="conv_bias"))
        ps += _norm_specs("blocks.%d" % (base + 2), w, norm)
        prev = w
    lin = 5 * len(hidden) - 1 + 2
    ps.append(ParamSpec("blocks.%d.weight" % lin, Axis(FIXED, classes), prev, label_rows=True, init="linear"))
    ps.append(ParamSpec("blocks.%d.bias" % lin, Axis(FIXED, classes), label_rows=True, init="zeros"))
    return ModelSpec("conv", "conv", ps, dict(data_shape=list(data_shape), hidden=list(hidden), classes=classes,
                                              norm=norm))


RESNET_BLOCKS = {"resnet18": ("block", [2, 2, 2, 2]), "resnet34": ("block", [3, 4, 6, 3]),
                 "resnet50": ("bottleneck", [3, 4, 6, 3]), "resnet101": ("bottleneck", [3, 4, 23, 3]),
                 "resnet152": ("bottleneck", [3, 8, 36, 3])}


def resnet_spec(name: str, data_shape: Sequence[int], hidden: Sequence[int], classes: int,
                norm: str) -> ModelSpec:
    """Pre-activation ResNet -- module registration order of reference ``src/models/resnet.py:10-50,104-139``
    (n1, conv1, n2, conv2, [n3, conv3], shortcut)."""
    kind, nblocks = RESNET_BLOCKS[name]
    exp = 1 if kind == "block" else 4
    ps: List[ParamSpec] = [ParamSpec("conv1.weight", Axis(PREFIX, hidden[0]), Axis(FIXED, data_shape[0]), 9, (3, 3),
                                     init="conv")]
    in_ax = Axis(PREFIX, hidden[0])
    blocks_meta = []
    for li, (planes, nb) in enumerate(zip(hidden, nblocks)):
        for bi in range(nb):
            stride = (1 if li == 0 else 2) if bi == 0 else 1
            pre = "layer%d.%d" % (li + 1, bi)
            p_ax = Axis(PREFIX, planes)
            o_ax = Axis(PREFIX, planes * exp)
            has_sc = stride != 1 or in_ax.size != planes * exp
            ps += _norm_specs(pre + ".n1", in_ax, norm)
            if kind == "block":
                ps.append(ParamSpec(pre + ".conv1.weight", p_ax, in_ax, 9, (3, 3), init="conv"))
                ps += _norm_specs(pre + ".n2", p_ax, norm)
                ps.append(ParamSpec(pre + ".conv2.weight", p_ax, p_ax, 9, (3, 3), init="conv"))
            else:
                ps.append(ParamSpec(pre + ".conv1.weight", p_ax, in_ax, 1, (1, 1), init="conv"))
                ps += _norm_specs(pre + ".n2", p_ax, norm)
                ps.append(ParamSpec(pre + ".conv2.weight", p_ax, p_ax, 9, (3, 3), init="conv"))
                ps += _norm_specs(pre + ".n3", p_ax, norm)
                ps.append(ParamSpec(pre + ".conv3.weight", o_ax, p_ax, 1, (1, 1), init="conv"))
            if has_sc:
                ps.append(ParamSpec(pre + ".shortcut.weight", o_ax, in_ax, 1, (1, 1), init="conv"))
            blocks_meta.append(dict(prefix=pre, stride=stride, shortcut=has_sc, in_planes=in_ax.size,
                                    planes=planes, kind=kind))
            in_ax = o_ax
    ps += _norm_specs("n4", in_ax, norm)
    ps.append(ParamSpec("linear.weight", Axis(FIXED, classes), in_ax, label_rows=True, init="linear"))
    ps.append(ParamSpec("linear.bias", Axis(FIXED, classes), label_rows=True, init="zeros"))
    return ModelSpec("resnet", name, ps, dict(data_shape=list(data_shape), hidden=list(hidden), classes=classes,
                                              norm=norm, blocks=blocks_meta, kind=kind, expansion=exp),
                     federable=(kind == "block"))


def transformer_spec(num_tokens: int, bptt: int, emb: int, heads: int, hidden: int, layers: int) -> ModelSpec:
    """Masked-LM encoder -- ``state_dict()`` order of reference ``src/models/transformer.py:11-143``.

    Slicing rules (reference ``src/fed.py:104-156``): embedding tables keep all rows / prefix of columns;
    q,k,v outputs are per-head prefixes; ``linear_o`` consumes the per-head index; ``decoder.linear2`` keeps all
    vocabulary rows.  Label-row aggregation applies to the *token* embedding and ``decoder.linear2``
    (``src/fed.py:263-274,282-286``), not to the positional table.
    """
    E, Hd = Axis(PREFIX, emb), Axis(PREFIX, hidden)
    EH = Axis(HEADS, emb, heads)
    ps: List[ParamSpec] = [
        ParamSpec("transformer_embedding.positional_embedding.positional_embedding.weight", Axis(FIXED, bptt), E,
                  init="embedding"),
        ParamSpec("transformer_embedding.embedding.weight", Axis(FIXED, num_tokens + 1), E, label_rows=True,
                  init="embedding"),
        ParamSpec("transformer_embedding.norm.weight", E, init="ones"),
        ParamSpec("transformer_embedding.norm.bias", E, init="zeros"),
    ]
    for l in range(layers):
        pre = "transformer_encoder.layers.%d" % l
        for nm in ("q", "k", "v"):
            ps.append(ParamSpec("%s.mha.linear_%s.weight" % (pre, nm), EH, E, init="linear"))
            ps.append(ParamSpec("%s.mha.linear_%s.bias" % (pre, nm), EH, init="zeros"))
        ps.append(ParamSpec(pre + ".mha.linear_o.weight", E, EH, init="linear"))
        ps.append(ParamSpec(pre + ".mha.linear_o.bias", E, init="zeros"))
        ps.append(ParamSpec(pre + ".norm1.weight", E, init="ones"))
        ps.append(ParamSpec(pre + ".norm1.bias", E, init="zeros"))
        ps.append(ParamSpec(pre + ".linear1.weight", Hd, E, init="normal02"))
        ps.append(ParamSpec(pre + ".linear1.bias", Hd, init="zeros"))
        ps.append(ParamSpec(pre + ".linear2.weight", E, Hd, init="normal02"))
        ps.append(ParamSpec(pre + ".linear2.bias", E, init="zeros"))
        ps.append(ParamSpec(pre + ".norm2.weight", E, init="ones"))
        ps.append(ParamSpec(pre + ".norm2.bias", E, init="zeros"))
    ps += [
        ParamSpec("decoder.linear1.weight", E, E, init="linear"),
        ParamSpec("decoder.linear1.bias", E, init="zeros"),
        ParamSpec("decoder.norm1.weight", E, init="ones"),
        ParamSpec("decoder.norm1.bias", E, init="zeros"),
        ParamSpec("decoder.linear2.weight", Axis(FIXED, num_tokens), E, label_rows=True, init="linear"),
        ParamSpec("decoder.linear2.bias", Axis(FIXED, num_tokens), label_rows=True, init="zeros"),
    ]
    return ModelSpec("transformer", "transformer", ps,
                     dict(num_tokens=num_tokens, bptt=bptt, emb=emb, heads=heads, hidden=hidden, layers=layers))


def build_spec(cfg: Dict) -> ModelSpec:
    """ModelSpec for ``cfg['model_name']`` (needs ``classes_size`` / ``num_tokens`` already set by the data layer)."""
    mn = cfg["model_name"]
    if mn == "conv":
        return conv_spec(cfg["data_shape"], cfg["conv"]["hidden_size"], cfg["classes_size"], cfg["norm"])
    if mn in RESNET_BLOCKS:
        return resnet_spec(mn, cfg["data_shape"], cfg["resnet"]["hidden_size"], cfg["classes_size"], cfg["norm"])
    if mn == "transformer":
        t = cfg["transformer"]
        return transformer_spec(cfg["num_tokens"], cfg["bptt"], t["embedding_size"], t["num_heads"],
                                t["hidden_size"], t["num_layers"])
    raise ValueError("Not valid model name")


def bn_buffer_names(spec: ModelSpec) -> List[Tuple[str, Axis]]:
    """(prefix, width axis) of every BatchNorm layer, in module order -- the sBN statistics live here."""
    if spec.meta.get("norm") != "bn":
        return []
    out = []
    for p in spec.params:
        if p.inp is None and p.name.endswith(".weight") and p.init == "ones":
            out.append((p.name[:-len(".weight")], p.out))
    return out
