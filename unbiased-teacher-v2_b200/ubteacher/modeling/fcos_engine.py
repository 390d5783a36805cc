"""B200-native execution engine for the FCOS R50-FPN detector of Unbiased Teacher v2.

One object = one model replica (student or teacher): a flat parameter arena (arena.py), packed bf16
tensor-core operands, and an explicit forward / backward schedule over the C-ABI kernels — NHWC bf16
activations, fp32 accumulation, no autograd tape, no host synchronisation.

Reference modules restated here (paths under /root/reference/ubteacher, [D2] = Detectron2 v0.6,
not on disk — SURVEY.md appendix B):
  backbone   modeling/backbone/fpn.py:51-78 (build_fcos_resnet_fpn_backbone), :11-29 (LastLevelP6P7),
             [D2] BasicStem / BottleneckBlock (STRIDE_IN_1X1) / FrozenBatchNorm2d / FPN
  head       modeling/fcos/fcos.py:220-376 (FCOSHead: 4+4 tower convs w/ GN(32)+ReLU, cls/bbox/std/ctr convs, Scale)
  losses     modeling/fcos/fcos_outputs.py:212-631  (via ops.fcos_loss_*)
  proposals  modeling/fcos/fcos_outputs.py:1046-1320 (via ops.fcos_predict_proposals)
"""
import math

import torch

from .. import _C, ops
from ..arena import PackPlan, ParamArena, Spec

BF16 = torch.bfloat16
RES_STAGES = [("res2", 3, 64, 256, 1), ("res3", 4, 128, 512, 2), ("res4", 6, 256, 1024, 2), ("res5", 3, 512, 2048, 2)]
BN_EPS = 1e-5


class Conv:
    """Runtime handle of one convolution: geometry + views of its operands / gradients."""

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

    def dgrad_compact(self, g, residual=None):
        """stride-2 1x1 only: the un-stuffed [N, P, Q, Cin] gradient (so two of them can be summed first)."""
        return ops.conv2d(g, self.wt, self.cin, 1, 1, 1, 0, None, None, residual)


class FcosEngine:
    def __init__(self, cfg, device="cuda", seed=0, init=True):
        self.cfg = cfg
        self.device = torch.device(device)
        f = cfg.MODEL.FCOS
        self.num_classes = f.NUM_CLASSES
        self.strides = list(f.FPN_STRIDES)
        self.soi = list(f.SIZES_OF_INTEREST)
        assert f.NUM_CLASSES == 80 and f.REG_DISCRETE and f.REG_MAX == 16 and f.KL_LOSS and f.NORM == "GN", \
            "the B200 engine implements the shipped UT2 FCOS recipe (80 classes, REG_DISCRETE(16), KL_LOSS, GN)"
        assert f.NUM_CLS_CONVS == 4 and f.NUM_BOX_CONVS == 4 and f.NUM_SHARE_CONVS == 0 and f.TOP_LEVELS == 2
        assert not f.CENTER_SAMPLE and f.USE_SCALE and cfg.MODEL.BACKBONE.FREEZE_AT == 2
        self.pixel_mean = [float(v) for v in cfg.MODEL.PIXEL_MEAN]
        self.pixel_std = [float(v) for v in cfg.MODEL.PIXEL_STD]
        self._build()
        if init:
            self.init_weights(seed)
        self.refresh_operands()

    # ------------------------------------------------------------------------------------ construction
    def _build(self):
        specs, convs, bn_list = [], {}, []

        def add_conv(name, cin, cout, k, stride, pad, bn, bias, trainable, need_dgrad=True, group=None):
            c = Conv(name, cin, cout, k, stride, pad, bn, bias, trainable, need_dgrad)
            grp = group or ("decay" if trainable else "frozen")
            specs.append(Spec(name + ".weight", (cout, cin, k, k), grp))
            if bias:
                specs.append(Spec(name + ".bias", (cout,), grp))
            if bn:
                bn_list.append((name + ".norm", cout))
            convs[name] = c
            return c

        bu = "backbone.bottom_up."
        self.stem = add_conv(bu + "stem.conv1", 3, 64, 7, 2, 3, True, False, False, False)
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
                    b["shortcut"] = add_conv(p + "shortcut", cin, cout, 1, s, 0, True, False, trainable,
                                             first_trainable_input)
                b["conv1"] = add_conv(p + "conv1", cin, mid, 1, s, 0, True, False, trainable, first_trainable_input)
                b["conv2"] = add_conv(p + "conv2", mid, mid, 3, 1, 1, True, False, trainable)
                b["conv3"] = add_conv(p + "conv3", mid, cout, 1, 1, 0, True, False, trainable)
                b["need_dx"] = first_trainable_input
                blks.append(b)
                cin = cout
            self.blocks.append((stage, blks))
        self.fpn_lat, self.fpn_out = {}, {}
        for lvl, c in ((3, 512), (4, 1024), (5, 2048)):
            self.fpn_lat[lvl] = add_conv(f"backbone.fpn_lateral{lvl}", c, 256, 1, 1, 0, False, True, True)
            self.fpn_out[lvl] = add_conv(f"backbone.fpn_output{lvl}", 256, 256, 3, 1, 1, False, True, True)
        self.p6 = add_conv("backbone.top_block.p6", 256, 256, 3, 2, 1, False, True, True)
        self.p7 = add_conv("backbone.top_block.p7", 256, 256, 3, 2, 1, False, True, True)
        hd = "proposal_generator.fcos_head."
        self.towers = {}
        for t in ("cls_tower", "bbox_tower"):
            layers = []
            for i in range(4):
                c = add_conv(f"{hd}{t}.{3 * i}", 256, 256, 3, 1, 1, False, True, True)
                specs.append(Spec(f"{hd}{t}.{3 * i + 1}.weight", (256,), "nodecay"))
                specs.append(Spec(f"{hd}{t}.{3 * i + 1}.bias", (256,), "nodecay"))
                layers.append((c, f"{hd}{t}.{3 * i + 1}"))
            self.towers[t] = layers
        self.cls_logits = add_conv(hd + "cls_logits", 256, 80, 3, 1, 1, False, True, True)
        # fused box predictor: rows 0..67 bbox_pred | 68..71 bbox_pred_std | 72 ctrness | 73..79 zero padding
        # (weights and biases are laid out back to back in the arena in that order, so the wgrad / bias-grad of
        #  the fused 80-row operator land directly in the three parameters' gradients)
        w_specs = [Spec(hd + "bbox_pred.weight", (68, 256, 3, 3), "decay", align=4),
                   Spec(hd + "bbox_pred_std.weight", (4, 256, 3, 3), "decay", align=1),
                   Spec(hd + "ctrness.weight", (1, 256, 3, 3), "decay", align=1),
                   Spec(hd + "bbox_pred.bias", (68,), "decay", align=4),
                   Spec(hd + "bbox_pred_std.bias", (4,), "decay", align=1),
                   Spec(hd + "ctrness.bias", (1,), "decay", align=1),
                   Spec("_boxpred_bias_pad", (7,), "decay", align=1, persistent=False)]
        specs.extend(w_specs)
        self.box_pred = Conv(hd + "box_fused", 256, 80, 3, 1, 1, False, True, True)
        self.box_pred.cout_store = 73
        for i in range(5):
            specs.append(Spec(f"{hd}scales.{i}.scale", (1,), "decay", align=4 if i == 0 else 1))
        # buffers: FrozenBN blocks [all weight | all bias | all running_mean | all running_var]
        for field in ("weight", "bias", "running_mean", "running_var"):
            for j, (n, c) in enumerate(bn_list):
                specs.append(Spec(f"{n}.{field}", (c,), "buffer", align=4 if j == 0 else 1))
        specs.append(Spec("pixel_mean", (3, 1, 1), "buffer"))
        specs.append(Spec("pixel_std", (3, 1, 1), "buffer"))
        specs.append(Spec("proposal_generator.fcos_outputs.integral.project", (17,), "buffer"))

        self.convs = convs
        self.bn_list = bn_list
        self.arena = A = ParamArena(specs, self.device)
        self.bn_total = sum(c for _, c in bn_list)
        self.bn_fold = torch.zeros(2, self.bn_total, dtype=torch.float32, device=self.device)
        self.bn_base = A.offset[bn_list[0][0] + ".weight"]
        assert A.offset[bn_list[0][0] + ".bias"] == self.bn_base + self.bn_total, "FrozenBN blocks must be contiguous"
        # packed operands
        self.plan = plan = PackPlan(A)
        bn_off = {}
        o = 0
        for n, c in bn_list:
            bn_off[n] = o
            o += c
        for name, c in convs.items():
            K = c.k * c.k * c.cin
            if name == self.stem.name:
                continue  # stem runs on the fp32 CUDA-core kernel
            wf = plan.alloc(c.cout * K)
            wt = -1
            coutT = (c.cout + 7) // 8 * 8
            if c.trainable and c.need_dgrad:
                wt = plan.alloc(c.cin * c.k * c.k * coutT)
            plan.add(name + ".weight", wf, wt, c.cout, c.cin, c.k, c.k, coutT,
                     scale_off=bn_off[name + ".norm"] if c.bn else -1)
            c._wf_off, c._wt_off, c._coutT = wf, wt, coutT
        # fused box predictor operands
        K = 9 * 256
        bp = self.box_pred
        bp._wf_off = plan.alloc(80 * K)
        bp._wt_off = plan.alloc(256 * 9 * 80)
        bp._coutT = 80
        row = 0
        for nm, rows in (("bbox_pred", 68), ("bbox_pred_std", 4), ("ctrness", 1)):
            plan.add(hd + nm + ".weight", bp._wf_off + row * K, bp._wt_off, rows, 256, 3, 3, 80, row)
            row += rows
        plan.finalize()
        plan.scales = self.bn_fold[0]
        for name, c in list(convs.items()) + [(bp.name, bp)]:
            if name == self.stem.name:
                continue
            K = c.k * c.k * c.cin
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
        bp.shift = A.data[A.offset[hd + "bbox_pred.bias"]: A.offset[hd + "bbox_pred.bias"] + 80]
        bp.db = A.grad[A.offset[hd + "bbox_pred.bias"]: A.offset[hd + "bbox_pred.bias"] + 80]
        bp.dw = A.grad[A.offset[hd + "bbox_pred.weight"]: A.offset[hd + "bbox_pred.weight"] + 73 * K]
        assert A.offset[hd + "bbox_pred_std.weight"] == A.offset[hd + "bbox_pred.weight"] + 68 * K
        assert A.offset[hd + "ctrness.bias"] == A.offset[hd + "bbox_pred.bias"] + 72
        self.scales = A.data[A.offset[hd + "scales.0.scale"]: A.offset[hd + "scales.0.scale"] + 5]
        self.dscales = A.grad[A.offset[hd + "scales.0.scale"]: A.offset[hd + "scales.0.scale"] + 5]
        self.gn = {}
        for t, layers in self.towers.items():
            for _, gname in layers:
                self.gn[gname] = (A.flat(gname + ".weight"), A.flat(gname + ".bias"), A.gflat(gname + ".weight"),
                                  A.gflat(gname + ".bias"))
        self.stem_w = torch.zeros(7, 7, 3, 64, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------------------------ init
    def init_weights(self, seed=0):
        """Seeded [D2]-style init: MSRA for the trunk, c2_xavier for FPN/top block, N(0, .01) heads,
        prior-prob bias on cls_logits (fcos.py:312-336), FrozenBN = identity, Scale = 1."""
        g = torch.Generator().manual_seed(seed)
        V = self.arena.views
        hd = "proposal_generator.fcos_head."
        for name, v in V.items():
            if name.startswith("_"):
                continue
            t = torch.zeros(v.shape)
            if name.startswith("backbone.bottom_up.") and name.endswith(".weight") and ".norm." not in name:
                fan_out = v.shape[0] * v.shape[2] * v.shape[3]
                t.normal_(0, math.sqrt(2.0 / fan_out), generator=g)
            elif ".norm.weight" in name:
                t.fill_(1.0)
            elif ".norm.running_var" in name:
                t.fill_(1.0 - BN_EPS)
            elif (name.startswith("backbone.fpn_") or name.startswith("backbone.top_block.")) and name.endswith(".weight"):
                fan_in = v.shape[1] * v.shape[2] * v.shape[3]
                bound = math.sqrt(3.0 / fan_in)          # kaiming_uniform_(a=1)
                t.uniform_(-bound, bound, generator=g)
            elif name.startswith(hd) and name.endswith(".weight") and t.dim() == 4:
                t.normal_(0, 1e-4 if "bbox_pred_std" in name else 0.01, generator=g)
            elif name.startswith(hd) and name.endswith(".weight") and t.dim() == 1:
                t.fill_(1.0)                              # GroupNorm gamma
            elif name.endswith(".scale"):
                t.fill_(1.0)
            elif name == hd + "cls_logits.bias":
                prior = self.cfg.MODEL.FCOS.PRIOR_PROB
                t.fill_(-math.log((1 - prior) / prior))
            elif name == "pixel_mean":
                t = torch.tensor(self.pixel_mean).view(3, 1, 1)
            elif name == "pixel_std":
                t = torch.tensor(self.pixel_std).view(3, 1, 1)
            elif name.endswith("integral.project"):
                t = torch.linspace(0, 16, 17)
            v.copy_(t.to(self.device))

    def refresh_operands(self):
        """Re-derive everything the kernels read from the fp32 arena: packed bf16 weights (one launch),
        FrozenBN scale/shift (one launch), the stem filter in [R,S,C,K] order."""
        A = self.arena
        b, n = self.bn_base, self.bn_total
        _C.counted_call("ut2_frozen_bn_fold", A.data[b:b + n], A.data[b + n:b + 2 * n], A.data[b + 2 * n:b + 3 * n],
                        A.data[b + 3 * n:b + 4 * n], _C.f32(BN_EPS), self.bn_fold[0], self.bn_fold[1], n)
        self.plan.run()        # after the fold: the packer multiplies the FrozenBN scale into the bf16 weights
        self.stem_w.copy_(A.views[self.stem.name + ".weight"].permute(2, 3, 1, 0))

    # ------------------------------------------------------------------------------------ geometry
    @staticmethod
    def padded_size(image_sizes, div=32):
        H = max(s[0] for s in image_sizes)
        W = max(s[1] for s in image_sizes)
        return (H + div - 1) // div * div, (W + div - 1) // div * div

    def level_geom(self, Hp, Wp):
        hw = []
        h, w = Hp // 8, Wp // 8
        for i in range(5):
            hw.append((h, w))
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        return ops.LevelGeom(hw, self.strides, self.soi)

    # ------------------------------------------------------------------------------------ forward
    def forward(self, images, train):
        """images: list of uint8 [3, h, w] CUDA tensors (BGR). Returns a dict with the dense head outputs
        (cls_out / box_out, level-major [P, 80] bf16), the level geometry and — when `train` — the tape."""
        N = len(images)
        sizes = [(int(im.shape[1]), int(im.shape[2])) for im in images]
        Hp, Wp = self.padded_size(sizes)
        geom = self.level_geom(Hp, Wp)
        tape = {} if train else None
        # stem + max-pool + res2 (frozen, no tape)
        P, Q = Hp // 2, Wp // 2
        x = torch.empty((N, P, Q, 64), dtype=BF16, device=self.device)
        for i, im in enumerate(images):
            ops.stem_conv(im, self.stem_w, self.stem.scale, self.stem.shift, self.pixel_mean, self.pixel_std, x[i], P, Q)
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
        # FPN
        c3, c4, c5 = feats["res3"], feats["res4"], feats["res5"]
        lat5 = self.fpn_lat[5].fwd(c5)
        lat4 = self.fpn_lat[4].fwd(c4, residual=lat5, res_up2=True)
        lat3 = self.fpn_lat[3].fwd(c3, residual=lat4, res_up2=True)
        p5 = self.fpn_out[5].fwd(lat5)
        p4 = self.fpn_out[4].fwd(lat4)
        p3 = self.fpn_out[3].fwd(lat3)
        p6 = self.p6.fwd(p5)
        p6r = ops.relu_bwd(p6, p6)              # relu(p6) = p6 * (p6 > 0)
        p7 = self.p7.fwd(p6r)
        if train:
            tape["fpn"] = (c3, c4, c5, lat3, lat4, lat5, p5, p6, p6r)
        # head
        Ptot = geom.L * N
        cls_out = torch.empty((Ptot, 80), dtype=BF16, device=self.device)
        box_out = torch.empty((Ptot, 80), dtype=BF16, device=self.device)
        head_tape = []
        for l, feat in enumerate((p3, p4, p5, p6, p7)):
            h, w = geom.hw[l]
            lo, hi = geom.off[l] * N, geom.off[l + 1] * N
            ctx = {"feat": feat}
            for t, pred, out in (("cls_tower", self.cls_logits, cls_out), ("bbox_tower", self.box_pred, box_out)):
                xx = feat
                saved = []
                for conv, gname in self.towers[t]:
                    c = conv.fwd(xx)
                    gam, bet, _, _ = self.gn[gname]
                    y, stats = ops.groupnorm_relu_fwd(c, gam, bet)
                    saved.append((xx, c, stats))
                    xx = y
                pred.fwd(xx, out=out[lo:hi].view(N, h, w, 80))
                ctx[t] = (saved, xx)
            head_tape.append(ctx)
        if train:
            tape["head"] = head_tape
            tape["N"] = N
        return {"cls_out": cls_out, "box_out": box_out, "geom": geom, "N": N, "image_sizes": sizes, "tape": tape,
                "padded": (Hp, Wp), "scales": self.scales}

    def _block_fwd(self, b, x, save):
        a = b["conv1"].fwd(x, relu=True)
        m = b["conv2"].fwd(a, relu=True)
        sc = b["shortcut"].fwd(x) if "shortcut" in b else x
        y = b["conv3"].fwd(m, residual=sc, relu=True)
        return y, ((x, a, m, y) if save else None)

    # ------------------------------------------------------------------------------------ backward
    def backward(self, fwd, dcls, dbox):
        """Back-propagate d(loss)/d(cls_out), d(loss)/d(box_out) through head, FPN and res5..res3,
        accumulating into the gradient arena (wgrad uses fp32 atomics, so repeated calls add up)."""
        tape, geom, N = fwd["tape"], fwd["geom"], fwd["N"]
        dfeat = []
        for l, ctx in enumerate(tape["head"]):
            h, w = geom.hw[l]
            lo, hi = geom.off[l] * N, geom.off[l + 1] * N
            acc = None
            for t, pred, dout in (("cls_tower", self.cls_logits, dcls), ("bbox_tower", self.box_pred, dbox)):
                saved, top = ctx[t]
                g = dout[lo:hi].view(N, h, w, 80)
                pred.wgrad(top, g)
                dx = pred.dgrad(g, (h, w))
                for i in range(3, -1, -1):
                    conv, gname = self.towers[t][i]
                    xin, c, stats = saved[i]
                    gam, bet, dgam, dbet = self.gn[gname]
                    dc = ops.groupnorm_relu_bwd(dx, c, stats, gam, bet, dgam, dbet, dbias_prev=conv.db)
                    conv.wgrad(xin, dc, bias_done=True)       # bias gradient came out of the GroupNorm backward
                    dx = conv.dgrad(dc, (h, w), residual=acc if i == 0 else None)
                acc = dx
            dfeat.append(acc)
        d3, d4, d5, d6, d7 = dfeat
        c3, c4, c5, lat3, lat4, lat5, p5, p6, p6r = tape["fpn"]
        hw = geom.hw
        self.p7.wgrad(p6r, d7)
        d6m = self.p7.dgrad(d7, hw[3], relu_mask=p6)
        d6t = ops.add_bf16(d6, d6m)
        self.p6.wgrad(p5, d6t)
        d5t = self.p6.dgrad(d6t, hw[2], residual=d5)
        self.fpn_out[5].wgrad(lat5, d5t)
        self.fpn_out[4].wgrad(lat4, d4)
        self.fpn_out[3].wgrad(lat3, d3)
        dl5 = self.fpn_out[5].dgrad(d5t, hw[2])
        dl4 = self.fpn_out[4].dgrad(d4, hw[1])
        dl3 = self.fpn_out[3].dgrad(d3, hw[0])
        dl4 = ops.downsample2x_sum(dl3, dl4)
        dl5 = ops.downsample2x_sum(dl4, dl5)
        self.fpn_lat[3].wgrad(c3, dl3)
        self.fpn_lat[4].wgrad(c4, dl4)
        self.fpn_lat[5].wgrad(c5, dl5)
        dc3 = self.fpn_lat[3].dgrad(dl3, hw[0])
        dc4 = self.fpn_lat[4].dgrad(dl4, hw[1])
        dc5 = self.fpn_lat[5].dgrad(dl5, hw[2])
        # trunk: res5 <- dc5 ; res4 <- dc4 + d(res5 input) ; res3 <- dc3 + d(res4 input)
        extra = {"res5": dc5, "res4": dc4, "res3": dc3}
        from_next = None          # d(loss)/d(stage output) flowing down from the later stage
        for stage, blks in reversed(self.blocks):
            if stage == "res2":
                break
            dy, dy2 = extra[stage], from_next
            for b, ctx in zip(reversed(blks), reversed(tape[stage])):
                dy = self._block_bwd(b, ctx, dy, dy2)
                dy2 = None
            from_next = dy
        return None

    def _block_bwd(self, b, ctx, dy, dy2=None):
        x, a, m, y = ctx
        H, W = x.shape[1], x.shape[2]
        g3 = ops.relu_bwd(dy, y, dy2)
        b["conv3"].wgrad(m, g3)
        g2 = b["conv3"].dgrad(g3, (m.shape[1], m.shape[2]), relu_mask=m)
        b["conv2"].wgrad(a, g2)
        g1 = b["conv2"].dgrad(g2, (a.shape[1], a.shape[2]), relu_mask=a)
        b["conv1"].wgrad(x, g1)
        if "shortcut" in b:
            b["shortcut"].wgrad(x, g3)
        if not b["need_dx"]:
            return None
        if "shortcut" in b:
            if b["stride"] == 2:
                xc = b["conv1"].dgrad_compact(g1)
                xs = b["shortcut"].dgrad_compact(g3, residual=xc)
                return ops.zero_stuff_s2(xs, H, W)
            xs = b["shortcut"].dgrad(g3, (H, W))
            return b["conv1"].dgrad(g1, (H, W), residual=xs)
        return b["conv1"].dgrad(g1, (H, W), residual=g3)

    # ------------------------------------------------------------------------------------ optimiser hooks
    def sgd_step(self, lr, momentum, wd, wd_norm, first_step, grad_scale=1.0, lr_dev=None):
        A = self.arena
        d0, d1 = A.group_range["decay"]
        n0, n1 = A.group_range["nodecay"]
        ops.sgd_step(A.data[d0:d1], A.grad[d0:d1], A.mom[d0:d1], lr, momentum, wd, first_step, True, grad_scale, lr_dev)
        ops.sgd_step(A.data[n0:n1], A.grad[n0:n1], A.mom[n0:n1], lr, momentum, wd_norm, first_step, True, grad_scale,
                     lr_dev)

    def ema_from(self, student, keep_rate):
        ops.ema_update(student.arena.data, self.arena.data, keep_rate)
