"""Shared machinery of the B200 execution engines (FCOS: fcos_engine.py, Faster R-CNN: rcnn_engine.py):
the ResNet-50 trunk ([D2] BasicStem / BottleneckBlock(STRIDE_IN_1X1) / FrozenBatchNorm2d, FREEZE_AT=2 — SURVEY.md B.1),
convolution handles over the C-ABI kernels, the flat parameter arena and the packed bf16 operand plan.

One engine = one model replica (student or teacher): NHWC bf16 activations, fp32 accumulation, an explicit
forward / backward schedule (no autograd tape, no host synchronisation).
"""
import math
import os

import torch

from .. import _C, ops
from ..arena import PackPlan, ParamArena, Spec

BF16 = torch.bfloat16
RES_STAGES = [("res2", 3, 64, 256, 1), ("res3", 4, 128, 512, 2), ("res4", 6, 256, 1024, 2), ("res5", 3, 512, 2048, 2)]
BN_EPS = 1e-5
STEM_FUSED = os.environ.get("UT2_STEM_FUSED", "1") != "0"     # 0: separate stem and max-pool kernels (A/B runs)


class Conv:
    """Runtime handle of one convolution (or FC layer seen as a 1x1 conv): geometry + views of its operands / gradients."""

    def __init__(self, name, cin, cout, k, stride, pad, bn=False, bias=False, trainable=True, need_dgrad=True):
        self.name, self.cin, self.cout, self.k, self.stride, self.pad = name, cin, cout, k, stride, pad
        self.bn, self.bias, self.trainable, self.need_dgrad = bn, bias, trainable, need_dgrad
        self.wf = self.wt = self.scale = self.shift = self.dw = self.db = self.bn_scale = None
        self.cout_store = 0

    def fwd(self, x, residual=None, relu=False, res_up2=False, out=None):
        return ops.conv2d(x, self.wf, self.cout, self.k, self.k, self.stride, self.pad, self.scale, self.shift,
                          residual, relu, out, res_up2)

    def wgrad(self, x, g, bias_done=False):
        if not self.trainable:
            return
        ops.conv2d_wgrad(x, g, self.cout, self.k, self.k, self.stride, self.pad, self.dw, self.bn_scale, self.cout_store)
        if self.bias and not bias_done:
            ops.colsum(g.view(-1, self.cout), self.db)

    def dgrad(self, g, in_hw, residual=None, relu_mask=None):
        """dX for stride 1 (direct) or stride 2 (1x1: compact GEMM then zero-stuff; 3x3: zero-stuff then s1 conv)."""
        k, H, W = self.k, in_hw[0], in_hw[1]
        if self.stride == 1:
            return ops.conv2d(g, self.wt, self.cin, k, k, 1, k - 1 - self.pad, None, None, residual, False, None, False,
                              relu_mask)
        assert self.stride == 2
        if k == 1:
            xc = ops.conv2d(g, self.wt, self.cin, 1, 1, 1, 0, None, None, residual)   # residual is compact here
            return ops.zero_stuff_s2(xc, H, W)
        gz = ops.zero_stuff_s2(g, H, W)
        alg = 2.0 * g.shape[0] * g.shape[1] * g.shape[2] * self.cout * self.cin * k * k   # zeros are not work
        return ops.conv2d(gz, self.wt, self.cin, k, k, 1, k - 1 - self.pad, None, None, residual, False, None, False,
                          relu_mask, alg)

    # ---- level-major pyramid variants (stride-1 convs whose weights are shared by all levels)
    def fwd_levels(self, x, geom, N, relu=False, out=None):
        return ops.conv2d_levels(x, geom, N, self.wf, self.cout, self.k, self.k, self.pad, self.scale, self.shift, None, relu, out)

    def wgrad_levels(self, x, g, geom, N, bias_done=False):
        ops.conv2d_wgrad_levels(x, g, geom, N, self.cout, self.k, self.k, self.pad, self.dw, self.bn_scale, self.cout_store)
        if self.bias and not bias_done:
            ops.colsum(g.view(-1, self.cout), self.db)

    def dgrad_levels(self, g, geom, N, residual=None, relu_mask=None):
        return ops.conv2d_levels(g, geom, N, self.wt, self.cin, self.k, self.k, self.k - 1 - self.pad, None, None, residual,
                                 False, None, relu_mask)

    def dgrad_compact(self, g, residual=None):
        """stride-2 1x1 only: the un-stuffed [N, P, Q, Cin] gradient (so two of them can be summed first)."""
        return ops.conv2d(g, self.wt, self.cin, 1, 1, 1, 0, None, None, residual)


class EngineBase:
    # ------------------------------------------------------------------------------------ construction helpers
    def _begin(self, cfg, device):
        self.cfg = cfg
        self.device = torch.device(device)
        self.pixel_mean = [float(v) for v in cfg.MODEL.PIXEL_MEAN]
        self.pixel_std = [float(v) for v in cfg.MODEL.PIXEL_STD]
        assert cfg.MODEL.BACKBONE.FREEZE_AT == 2, "the B200 engines implement the shipped recipes (FREEZE_AT 2)"
        self._specs, self.convs, self.bn_list, self._fused = [], {}, [], []

    def add_conv(self, name, cin, cout, k, stride, pad, bn, bias, trainable, need_dgrad=True, group=None, shape=None):
        c = Conv(name, cin, cout, k, stride, pad, bn, bias, trainable, need_dgrad)
        grp = group or ("decay" if trainable else "frozen")
        self._specs.append(Spec(name + ".weight", shape or (cout, cin, k, k), grp))
        if bias:
            self._specs.append(Spec(name + ".bias", (cout,), grp))
        if bn:
            self.bn_list.append((name + ".norm", cout))
        self.convs[name] = c
        return c

    def add_fused(self, name, cin, k, pad, parts, cout_pad, linear=False):
        """Several predictors sharing one input, run as ONE zero-padded operator: `parts` = [(param prefix, rows)].
        Their master weights (then their biases) are laid out back to back in the arena in that order, so the wgrad /
        bias-grad of the fused operator land directly in the individual parameters' gradients."""
        rows = sum(r for _, r in parts)
        for i, (p, r) in enumerate(parts):
            shape = (r, cin) if linear else (r, cin, k, k)
            self._specs.append(Spec(p + ".weight", shape, "decay", align=4 if i == 0 else 1))
        for i, (p, r) in enumerate(parts):
            self._specs.append(Spec(p + ".bias", (r,), "decay", align=4 if i == 0 else 1))
        if cout_pad > rows:
            self._specs.append(Spec(f"_{name}_bias_pad", (cout_pad - rows,), "decay", align=1, persistent=False))
        c = Conv(name, cin, cout_pad, k, 1, pad, False, True, True)
        c.cout_store = rows
        self._fused.append((c, parts))
        return c

    def build_trunk(self):
        bu = "backbone.bottom_up."
        self.stem = self.add_conv(bu + "stem.conv1", 3, 64, 7, 2, 3, True, False, False, False)
        self.blocks = []      # (stage, [dict(conv1, conv2, conv3, shortcut)])
        cin = 64
        for stage, nblk, mid, cout, stride in RES_STAGES:
            trainable = stage != "res2"
            blks = []
            for i in range(nblk):
                s = stride if i == 0 else 1
                p = f"{bu}{stage}.{i}."
                first_trainable_input = trainable and not (stage == "res3" and i == 0)
                b = {"stride": s}
                if cin != cout:
                    b["shortcut"] = self.add_conv(p + "shortcut", cin, cout, 1, s, 0, True, False, trainable,
                                                  first_trainable_input)
                b["conv1"] = self.add_conv(p + "conv1", cin, mid, 1, s, 0, True, False, trainable, first_trainable_input)
                b["conv2"] = self.add_conv(p + "conv2", mid, mid, 3, 1, 1, True, False, trainable)
                b["conv3"] = self.add_conv(p + "conv3", mid, cout, 1, 1, 0, True, False, trainable)
                b["need_dx"] = first_trainable_input
                blks.append(b)
                cin = cout
            self.blocks.append((stage, blks))

    def finish(self, extra_buffers=()):
        """Buffers (FrozenBN blocks, pixel stats), the arena, the packed-operand plan and every Conv's views."""
        specs, convs, bn_list = self._specs, self.convs, self.bn_list
        for field in ("weight", "bias", "running_mean", "running_var"):
            for j, (n, c) in enumerate(bn_list):
                specs.append(Spec(f"{n}.{field}", (c,), "buffer", align=4 if j == 0 else 1))
        for sp in extra_buffers:
            specs.append(sp)
        self.arena = A = ParamArena(specs, self.device)
        self.bn_total = sum(c for _, c in bn_list)
        self.bn_fold = torch.zeros(2, self.bn_total, dtype=torch.float32, device=self.device)
        self.bn_base = A.offset[bn_list[0][0] + ".weight"]
        assert A.offset[bn_list[0][0] + ".bias"] == self.bn_base + self.bn_total, "FrozenBN blocks must be contiguous"
        self.plan = plan = PackPlan(A)
        bn_off, o = {}, 0
        for n, c in bn_list:
            bn_off[n] = o
            o += c
        for name, c in convs.items():
            K = c.k * c.k * c.cin
            if name == self.stem.name:
                continue  # the stem kernel packs its own filter
            wf = plan.alloc(c.cout * K)
            wt = -1
            coutT = (c.cout + 7) // 8 * 8
            if c.trainable and c.need_dgrad:
                wt = plan.alloc(c.cin * c.k * c.k * coutT)
            plan.add(name + ".weight", wf, wt, c.cout, c.cin, c.k, c.k, coutT,
                     scale_off=bn_off[name + ".norm"] if c.bn else -1)
            c._wf_off, c._wt_off, c._coutT = wf, wt, coutT
        for c, parts in self._fused:
            K = c.k * c.k * c.cin
            c._wf_off = plan.alloc(c.cout * K)
            c._wt_off = plan.alloc(c.cin * c.k * c.k * c.cout)
            c._coutT = c.cout
            row = 0
            for p, rows in parts:
                plan.add(p + ".weight", c._wf_off + row * K, c._wt_off, rows, c.cin, c.k, c.k, c.cout, row)
                row += rows
        plan.finalize()
        plan.scales = self.bn_fold[0]
        for name, c in list(convs.items()) + [(c.name, c) for c, _ in self._fused]:
            if name == self.stem.name:
                continue
            c.wf = plan.view(c._wf_off, (c.cout, c.k, c.k, c.cin))
            if c._wt_off >= 0:
                c.wt = plan.view(c._wt_off, (c.cin, c.k, c.k, c._coutT))
        for name, c in convs.items():
            if c.bn:
                # the FrozenBN scale is folded into the packed bf16 weights (forward and dgrad operands); the epilogue
                # only adds the shift. The weight gradient still needs the scale: d/dW = scale * (dY^T X).
                o = bn_off[name + ".norm"]
                c.bn_scale, c.shift = self.bn_fold[0, o:o + c.cout], self.bn_fold[1, o:o + c.cout]
                if name == self.stem.name:
                    c.scale = c.bn_scale      # the stem kernel packs its own fp32 filter
            elif c.bias:
                c.shift = A.flat(name + ".bias")
                c.db = A.gflat(name + ".bias")
            if c.trainable:
                c.dw = A.gflat(name + ".weight")
        for c, parts in self._fused:
            K = c.k * c.k * c.cin
            w0, b0 = A.offset[parts[0][0] + ".weight"], A.offset[parts[0][0] + ".bias"]
            c.shift = A.data[b0:b0 + c.cout]
            c.db = A.grad[b0:b0 + c.cout]
            c.dw = A.grad[w0:w0 + c.cout_store * K]
            row = 0
            for p, rows in parts:        # the fused operator's rows must be contiguous in the arena
                assert A.offset[p + ".weight"] == w0 + row * K and A.offset[p + ".bias"] == b0 + row, p
                row += rows
        self.stem_w = torch.zeros(7, 7, 3, 64, dtype=torch.float32, device=self.device)
        # Gradient-arena segments in the order the backward pass completes them: everything above the trunk ("top": FPN, heads,
        # GroupNorm), then res5, res4, res3. `grad_hook(lo, hi)` — set by the trainer for the LAST backward of a step — is called
        # as soon as a segment's gradients are final, so that its all-reduce overlaps the rest of the backward pass (the
        # reference's DDP does the same with 25 MB buckets, engine/trainer.py:60-63).
        self.grad_hook = None
        bounds, prev = {}, None
        for n, sp in A.specs.items():
            if sp.group not in ("decay", "nodecay"):
                continue
            key = next((st for st in ("res3", "res4", "res5") if n.startswith(f"backbone.bottom_up.{st}.")), "top")
            if key != prev:
                assert key not in bounds, f"gradient arena: segment {key} is not contiguous"
                bounds[key] = A.offset[n]
                prev = key
        order = sorted(bounds, key=bounds.get)
        assert order == ["res3", "res4", "res5", "top"], order
        ends = [bounds[k] for k in order[1:]] + [A.n_trainable]
        self.grad_segments = {k: (bounds[k], e) for k, e in zip(order, ends)}

    # ------------------------------------------------------------------------------------ init
    def init_trunk_entry(self, name, v, g):
        """Seeded [D2]-style init of trunk / FPN / FrozenBN / pixel-stat entries. Returns a CPU tensor or None."""
        t = torch.zeros(v.shape)
        if name.startswith("backbone.bottom_up.") and name.endswith(".weight") and ".norm." not in name:
            fan_out = v.shape[0] * v.shape[2] * v.shape[3]
            return t.normal_(0, math.sqrt(2.0 / fan_out), generator=g)           # MSRA fill
        if ".norm.weight" in name:
            # FrozenBN gamma. The pretrained R-50 (MSRA pickle, unreachable offline) carries statistics that keep the
            # residual stream O(1); with identity FrozenBN a randomly initialised trunk grows to ~3e3 by res5 and one SGD
            # step diverges. Offline stand-in: damp the stem (pixel-scale input) and each block's residual branch.
            return t.fill_(0.04 if "stem." in name else 0.25 if ".conv3." in name else 1.0)
        if ".norm.running_var" in name:
            return t.fill_(1.0 - BN_EPS)
        if ".norm." in name:
            return t
        if (name.startswith("backbone.fpn_") or name.startswith("backbone.top_block.")) and name.endswith(".weight"):
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            bound = math.sqrt(3.0 / fan_in)          # c2_xavier_fill = kaiming_uniform_(a=1)
            return t.uniform_(-bound, bound, generator=g)
        if name == "pixel_mean":
            return torch.tensor(self.pixel_mean).view(3, 1, 1)
        if name == "pixel_std":
            return torch.tensor(self.pixel_std).view(3, 1, 1)
        return None

    def refresh_operands(self, dgrad=True):
        """Re-derive everything the kernels read from the fp32 arena: packed bf16 weights (one launch),
        FrozenBN scale/shift (one launch), the stem filter in [R,S,C,K] order. dgrad=False skips the transposed
        (data-gradient) operands: the EMA teacher never back-propagates."""
        A = self.arena
        b, n = self.bn_base, self.bn_total
        _C.counted_call("ut2_frozen_bn_fold", A.data[b:b + n], A.data[b + n:b + 2 * n], A.data[b + 2 * n:b + 3 * n],
                        A.data[b + 3 * n:b + 4 * n], _C.f32(BN_EPS), self.bn_fold[0], self.bn_fold[1], n)
        self.plan.run(dgrad)   # after the fold: the packer multiplies the FrozenBN scale into the bf16 weights
        self.stem_w.copy_(A.views[self.stem.name + ".weight"].permute(2, 3, 1, 0))

    # ------------------------------------------------------------------------------------ geometry
    @staticmethod
    def padded_size(image_sizes, div=32):
        H = max(s[0] for s in image_sizes)
        W = max(s[1] for s in image_sizes)
        return (H + div - 1) // div * div, (W + div - 1) // div * div

    # ------------------------------------------------------------------------------------ trunk forward / backward
    def trunk_forward(self, images, train, tape):
        """uint8 CHW images -> {"res2".."res5"} NHWC bf16 features (tape[stage] filled when `train`)."""
        N = len(images)
        sizes = [(int(im.shape[1]), int(im.shape[2])) for im in images]
        Hp, Wp = self.padded_size(sizes)
        if STEM_FUSED:      # stem + max-pool in one pass: the 64-channel stem activation never reaches HBM (csrc/stem_pool.cu)
            x = ops.stem_pool_batched(images, self.stem_w, self.stem.scale, self.stem.shift, self.pixel_mean, self.pixel_std, Hp, Wp)
        else:
            P, Q = Hp // 2, Wp // 2
            x = torch.empty((N, P, Q, 64), dtype=BF16, device=self.device)
            ops.stem_conv_batched(images, self.stem_w, self.stem.scale, self.stem.shift, self.pixel_mean, self.pixel_std, x, P, Q)
            x = ops.maxpool3x3s2(x)
        feats = {}
        for stage, blks in self.blocks:
            saved = []
            for b in blks:
                x, ctx = self._block_fwd(b, x, train and stage != "res2")
                saved.append(ctx)
            feats[stage] = x
            if train:
                tape[stage] = saved
        return feats, sizes, (Hp, Wp)

    def trunk_backward(self, tape, lateral):
        """lateral: {"res5": (conv, dl5, hw), "res4": ..., "res3": ...} — the FPN lateral convolutions with the gradient of their
        OUTPUT. Their data-gradient is launched here, when the gradient flowing down from the later stage is known, so that the
        block-output ReLU backward g3 = (dC + d_next) * (y > 0) happens in that dgrad's epilogue (residual + mask tiles) and
        never as a pass of its own. Likewise every bottleneck hands the previous block an already masked gradient."""
        from_next = None          # d(loss)/d(stage output) flowing down from the later stage (not yet masked)
        for stage, blks in reversed(self.blocks):
            if stage == "res2":
                break
            conv, dl, hw = lateral[stage]
            saved = tape[stage]
            g3 = conv.dgrad(dl, hw, residual=from_next, relu_mask=saved[-1][3])
            for i in range(len(blks) - 1, -1, -1):
                g3 = self._block_bwd(blks[i], saved[i], g3, mask_input=i > 0)
            from_next = g3
            if self.grad_hook is not None:
                self.grad_hook(*self.grad_segments[stage])

    def _block_fwd(self, b, x, save):
        a = b["conv1"].fwd(x, relu=True)
        m = b["conv2"].fwd(a, relu=True)
        sc = b["shortcut"].fwd(x) if "shortcut" in b else x
        y = b["conv3"].fwd(m, residual=sc, relu=True)
        return y, ((x, a, m, y) if save else None)

    def _block_bwd(self, b, ctx, g3, mask_input=False):
        """g3 = d(loss)/d(conv3 + shortcut) (the block-output ReLU is already applied to it). Returns d(loss)/d(block input);
        with mask_input it is multiplied by (input > 0), i.e. it IS the g3 of the previous block of the stage (whose output
        this input is)."""
        x, a, m, y = ctx
        H, W = x.shape[1], x.shape[2]
        b["conv3"].wgrad(m, g3)
        g2 = b["conv3"].dgrad(g3, (m.shape[1], m.shape[2]), relu_mask=m)
        b["conv2"].wgrad(a, g2)
        g1 = b["conv2"].dgrad(g2, (a.shape[1], a.shape[2]), relu_mask=a)
        b["conv1"].wgrad(x, g1)
        if "shortcut" in b:
            b["shortcut"].wgrad(x, g3)
        if not b["need_dx"]:
            return None
        mask = x if mask_input else None
        if "shortcut" in b:
            if b["stride"] == 2:
                assert mask is None       # strided blocks open a stage: their input gradient is masked by the lateral dgrad
                xc = b["conv1"].dgrad_compact(g1)
                xs = b["shortcut"].dgrad_compact(g3, residual=xc)
                return ops.zero_stuff_s2(xs, H, W)
            xs = b["shortcut"].dgrad(g3, (H, W))
            return b["conv1"].dgrad(g1, (H, W), residual=xs, relu_mask=mask)
        return b["conv1"].dgrad(g1, (H, W), residual=g3, relu_mask=mask)

    # ------------------------------------------------------------------------------------ optimiser hooks
    def sgd_step(self, lr, momentum, wd, wd_norm, first_step, grad_scale=1.0, lr_dev=None):
        A = self.arena
        d0, d1 = A.group_range["decay"]
        n0, n1 = A.group_range["nodecay"]
        ops.sgd_step(A.data[d0:d1], A.grad[d0:d1], A.mom[d0:d1], lr, momentum, wd, first_step, True, grad_scale, lr_dev)
        if n1 > n0:
            ops.sgd_step(A.data[n0:n1], A.grad[n0:n1], A.mom[n0:n1], lr, momentum, wd_norm, first_step, True, grad_scale,
                         lr_dev)

    def ema_from(self, student, keep_rate):
        ops.ema_update(student.arena.data, self.arena.data, keep_rate)
