"""nn.Module facades over sub-trees of an engine's parameter arena.

The reference composes its detectors from Detectron2 modules looked up in string registries (backbone, proposal
generator, ROI heads, output layers). Here one engine (fcos_engine.py / rcnn_engine.py) owns every parameter of a
replica in a flat arena; the classes built on ``ArenaView`` give each reference module its name, constructor
signature, ``forward`` contract and ``state_dict`` sub-tree (same keys as the reference module's), executing on the
engine's kernels. A view built without ``engine=`` owns a private engine, so ``REGISTRY.get(name)(cfg, input_shape)``
works on its own as well as inside the meta-architecture (which passes its engine to all of its views)."""
from collections import OrderedDict

import torch
from torch import nn


class ArenaView(nn.Module):
    def __init__(self, engine, prefix):
        super().__init__()
        object.__setattr__(self, "engine", engine)       # not an nn.Module: a plain attribute
        self._prefix = prefix
        self._param_cache = None

    @property
    def device(self):
        return self.engine.device

    def _entries(self):
        A = self.engine.arena
        for n, sp in A.specs.items():
            if n.startswith(self._prefix) and sp.persistent and not n.startswith("_"):
                yield n, sp

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        A = self.engine.arena
        full = A.state_dict()
        sd = OrderedDict((prefix + n[len(self._prefix):], full[n]) for n, _ in self._entries())
        if destination is not None:
            destination.update(sd)
            return destination
        return sd

    def load_state_dict(self, state_dict, strict=True):
        A = self.engine.arena
        missing = []
        for n, _ in self._entries():
            k = n[len(self._prefix):]
            if k in state_dict:
                v = A.views[n]
                v.copy_(state_dict[k].to(v.device, torch.float32).reshape(v.shape))
            else:
                missing.append(k)
        if strict and missing:
            raise KeyError(f"missing keys in state_dict: {missing[:5]}")
        self.engine.refresh_operands()
        return missing

    def named_parameters(self, prefix="", recurse=True, remove_duplicate=True):
        if self._param_cache is None:
            A = self.engine.arena
            self._param_cache = []
            for n, sp in self._entries():
                if sp.group == "buffer":
                    continue
                p = nn.Parameter(A.views[n], requires_grad=sp.group in ("decay", "nodecay"))
                if p.requires_grad:
                    p.grad = A.gviews[n]
                self._param_cache.append((n[len(self._prefix):], p))
        for n, p in self._param_cache:
            yield prefix + n, p

    def parameters(self, recurse=True):
        for _, p in self.named_parameters():
            yield p


def nchw(t):
    """[N, H, W, C] device tensor -> logical [N, C, H, W] (a channels-last view, no copy)."""
    return t.permute(0, 3, 1, 2)
