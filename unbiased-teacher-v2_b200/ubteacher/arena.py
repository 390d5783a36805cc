"""Flat HBM arenas for parameters, gradients, momentum and the packed bf16 tensor-core operands.

Why: the reference updates ~320 tensors one by one (EMA loop trainer.py:468-486, per-parameter SGD,
DDP's 25 MB buckets). Here every model owns ONE contiguous fp32 arena laid out as

    [ trainable, weight-decayed | trainable, no decay (norm layers) | frozen params | buffers ]

so that EMA is one kernel over the whole arena, SGD is one kernel per decay group, the gradient
all-reduce is one NCCL call on a contiguous buffer, and ``state_dict()`` is a dict of views with the
reference's key names / logical shapes (conv weights are channels-last views, i.e. physically
[Cout, R, S, Cin] — the K-major layout the implicit-GEMM kernels want).
"""
import struct
from collections import OrderedDict

import torch

GROUPS = ("decay", "nodecay", "frozen", "buffer")


class Spec:
    __slots__ = ("name", "shape", "group", "align", "persistent")

    def __init__(self, name, shape, group, align=4, persistent=True):
        self.name, self.shape, self.group, self.align, self.persistent = name, tuple(shape), group, align, persistent

    @property
    def numel(self):
        n = 1
        for s in self.shape:
            n *= s
        return n


def _view(flat, off, shape):
    n = 1
    for s in shape:
        n *= s
    v = flat[off:off + n]
    if len(shape) == 4:  # logical [Cout, Cin, R, S], physical [Cout, R, S, Cin]
        co, ci, r, s = shape
        return v.view(co, r, s, ci).permute(0, 3, 1, 2)
    return v.view(shape)


class ParamArena:
    def __init__(self, specs, device):
        self.device = device
        self.specs = OrderedDict()
        self.offset = {}
        off = 0
        self.group_range = {}
        for g in GROUPS:
            start = off
            for sp in specs:
                if sp.group != g:
                    continue
                if off % sp.align:
                    off += sp.align - off % sp.align
                self.specs[sp.name] = sp
                self.offset[sp.name] = off
                off += sp.numel
            if off % 4:
                off += 4 - off % 4
            self.group_range[g] = (start, off)
        self.total = off
        self.n_trainable = self.group_range["nodecay"][1]
        self.data = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(self.n_trainable, dtype=torch.float32, device=device)
        self.mom = torch.zeros(self.n_trainable, dtype=torch.float32, device=device)
        self.views = OrderedDict((n, _view(self.data, self.offset[n], sp.shape)) for n, sp in self.specs.items())
        self.gviews = OrderedDict((n, _view(self.grad, self.offset[n], sp.shape)) for n, sp in self.specs.items()
                                  if sp.group in ("decay", "nodecay"))
        self.export_shape = {}     # name -> checkpoint shape where it differs from the logical view (fc1: [1024, 12544])

    def flat(self, name):
        """Flat fp32 slice of one entry (physical order)."""
        o = self.offset[name]
        return self.data[o:o + self.specs[name].numel]

    def gflat(self, name):
        o = self.offset[name]
        return self.grad[o:o + self.specs[name].numel]

    def state_dict(self, prefix=""):
        return OrderedDict((prefix + n, v.reshape(self.export_shape[n]) if n in self.export_shape else v)
                           for n, v in self.views.items() if self.specs[n].persistent and not n.startswith("_"))

    def load_state_dict(self, sd, strict=True):
        missing = []
        for n, v in self.views.items():
            if not self.specs[n].persistent or n.startswith("_"):
                continue
            if n in sd:
                v.copy_(sd[n].to(v.device, torch.float32).reshape(v.shape))
            else:
                missing.append(n)
        if strict and missing:
            raise KeyError(f"missing keys in state_dict: {missing[:5]}...")
        return missing

    def copy_from(self, other):
        self.data.copy_(other.data)


class PackPlan:
    """bf16 operand arena + the descriptor table consumed by ut2_pack_conv_weights_batched."""

    def __init__(self, arena):
        self.arena = arena
        self.entries = []      # (src_off, wf_off, wt_off, scale_off, Cout, Cin, R, S, CoutT, n_off)
        self.scales = None     # fp32 tensor the scale offsets index (the folded FrozenBN scales)
        self.size = 0
        self.packed = None
        self.table = None
        self.total = 0

    def alloc(self, numel):
        off = self.size
        self.size += (numel + 63) // 64 * 64     # 128-byte aligned operands for TMA
        return off

    def add(self, name, wf_off, wt_off, cout, cin, R, S, coutT, n_off=0, scale_off=-1):
        self.entries.append((self.arena.offset[name], wf_off, wt_off, scale_off, cout, cin, R, S, coutT, n_off))

    def finalize(self):
        dev = self.arena.device
        self.packed = torch.zeros(max(self.size, 64), dtype=torch.bfloat16, device=dev)
        raw = bytearray()
        begin = 0
        for (src, wf, wt, sc, cout, cin, R, S, coutT, n_off) in self.entries:
            raw += struct.pack("<qqqqqiiiiii", src, wf, wt, begin, sc, cout, cin, R, S, coutT, n_off)
            begin += R * S * ((cout + 63) // 64) * ((cin + 31) // 32)      # 64 x 32 tiles per tap (csrc/optim.cu)
        self.total = begin
        self.table = torch.frombuffer(raw, dtype=torch.uint8).clone().to(dev)
        return self

    def view(self, off, shape):
        n = 1
        for s in shape:
            n *= s
        return self.packed[off:off + n].view(shape)

    def run(self, dgrad=True):
        from . import _C
        from ._C import i64
        _C.counted_call("ut2_pack_conv_weights_batched", self.table, len(self.entries), i64(self.total),
                        self.arena.data, self.scales, self.packed, int(dgrad))
