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
from ..arena import Spec
from .engine_base import BF16, Conv, EngineBase


class FcosEngine(EngineBase):
    def __init__(self, cfg, device="cuda", seed=0, init=True):
        self._begin(cfg, device)
        f = cfg.MODEL.FCOS
        self.num_classes = f.NUM_CLASSES
        self.strides = list(f.FPN_STRIDES)
        self.soi = list(f.SIZES_OF_INTEREST)
        assert f.NUM_CLASSES == 80 and f.REG_DISCRETE and f.REG_MAX == 16 and f.KL_LOSS and f.NORM == "GN", \
            "the B200 engine implements the shipped UT2 FCOS recipe (80 classes, REG_DISCRETE(16), KL_LOSS, GN)"
        assert f.NUM_CLS_CONVS == 4 and f.NUM_BOX_CONVS == 4 and f.NUM_SHARE_CONVS == 0 and f.TOP_LEVELS == 2
        assert f.USE_SCALE      # CENTER_SAMPLE / POS_RADIUS are target-assignment options (FCOSOutputs), not engine ones
        self._build()
        if init:
            self.init_weights(seed)
        self.refresh_operands()

    # ------------------------------------------------------------------------------------ construction
    def _build(self):
        specs, add_conv = self._specs, self.add_conv
        self.build_trunk()
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
        self.box_pred = self.add_fused(hd + "box_fused", 256, 3, 1,
                                       [(hd + "bbox_pred", 68), (hd + "bbox_pred_std", 4), (hd + "ctrness", 1)], 80)
        for i in range(5):
            specs.append(Spec(f"{hd}scales.{i}.scale", (1,), "decay", align=4 if i == 0 else 1))
        self.finish([Spec("pixel_mean", (3, 1, 1), "buffer"), Spec("pixel_std", (3, 1, 1), "buffer"),
                     Spec("proposal_generator.fcos_outputs.integral.project", (17,), "buffer")])
        A = self.arena
        self.scales = A.data[A.offset[hd + "scales.0.scale"]: A.offset[hd + "scales.0.scale"] + 5]
        self.dscales = A.grad[A.offset[hd + "scales.0.scale"]: A.offset[hd + "scales.0.scale"] + 5]
        self.gn = {}
        for t, layers in self.towers.items():
            for _, gname in layers:
                self.gn[gname] = (A.flat(gname + ".weight"), A.flat(gname + ".bias"), A.gflat(gname + ".weight"),
                                  A.gflat(gname + ".bias"))

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
            t = self.init_trunk_entry(name, v, g)
            if t is None:
                t = torch.zeros(v.shape)
                if name.startswith(hd) and name.endswith(".weight") and t.dim() == 4:
                    t.normal_(0, 1e-4 if "bbox_pred_std" in name else 0.01, generator=g)
                elif name.startswith(hd) and name.endswith(".weight") and t.dim() == 1:
                    t.fill_(1.0)                              # GroupNorm gamma
                elif name.endswith(".scale"):
                    t.fill_(1.0)
                elif name == hd + "cls_logits.bias":
                    prior = self.cfg.MODEL.FCOS.PRIOR_PROB
                    t.fill_(-math.log((1 - prior) / prior))
                elif name.endswith("integral.project"):
                    t = torch.linspace(0, 16, 17)
            v.copy_(t.to(self.device))

    # ------------------------------------------------------------------------------------ geometry
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
        tape = {} if train else None
        feats, sizes, (Hp, Wp) = self.trunk_forward(images, train, tape)
        geom = self.level_geom(Hp, Wp)
        feat = self.fpn_forward(feats, geom, N, tape)
        cls_out, box_out = self.head_forward(feat, geom, N, tape)
        return {"cls_out": cls_out, "box_out": box_out, "geom": geom, "N": N, "image_sizes": sizes, "tape": tape,
                "padded": (Hp, Wp), "scales": self.scales, "feat": feat}

    def fpn_forward(self, feats, geom, N, tape=None):
        """FPN over res3..res5 + LastLevelP6P7 (backbone/fpn.py:11-78). Returns the level-major pyramid [N * L, 256]."""
        c3, c4, c5 = feats["res3"], feats["res4"], feats["res5"]
        lat5 = self.fpn_lat[5].fwd(c5)
        lat4 = self.fpn_lat[4].fwd(c4, residual=lat5, res_up2=True)
        lat3 = self.fpn_lat[3].fwd(c3, residual=lat4, res_up2=True)
        # FPN outputs p3..p7 go straight into ONE level-major buffer [N * 22400, 256] (level, image, h, w): the towers
        # and predictors share their weights across levels (fcos.py:338-376), so every head layer below is a single
        # launch over the whole pyramid instead of five (the 273- and 77-location levels cannot fill 148 SMs alone).
        feat = torch.empty((geom.L * N, 256), dtype=BF16, device=self.device)
        lv = self.level_views(feat, geom, N, 256)
        self.fpn_out[3].fwd(lat3, out=lv[0])
        self.fpn_out[4].fwd(lat4, out=lv[1])
        p5 = self.fpn_out[5].fwd(lat5, out=lv[2])
        p6 = self.p6.fwd(p5, out=lv[3])
        p6r = ops.relu_bwd(p6, p6)              # relu(p6) = p6 * (p6 > 0)
        self.p7.fwd(p6r, out=lv[4])
        if tape is not None:
            tape["fpn"] = (c3, c4, c5, lat3, lat4, lat5, p5, p6, p6r)
        return feat

    @staticmethod
    def level_views(buf, geom, N, C):
        return [buf[geom.off[l] * N: geom.off[l + 1] * N].view(N, geom.hw[l][0], geom.hw[l][1], C) for l in range(geom.num)]

    def head_forward(self, feat, geom, N, tape=None):
        """FCOSHead (fcos/fcos.py:338-376) over the level-major pyramid: (cls_out [P, 80], box_out [P, 80] = bbox_pred(68,
        BEFORE Scale_l) | bbox_pred_std(4) | ctrness(1) | 7 zero columns)."""
        Ptot = geom.L * N
        cls_out = torch.empty((Ptot, 80), dtype=BF16, device=self.device)
        box_out = torch.empty((Ptot, 80), dtype=BF16, device=self.device)
        head_tape = {"feat": feat}
        for t, pred, out in (("cls_tower", self.cls_logits, cls_out), ("bbox_tower", self.box_pred, box_out)):
            xx = feat
            saved = []
            for conv, gname in self.towers[t]:
                c = conv.fwd_levels(xx, geom, N)
                gam, bet, _, _ = self.gn[gname]
                y, stats = ops.groupnorm_relu_levels_fwd(c, geom, N, gam, bet)
                saved.append((xx, c, stats))
                xx = y
            pred.fwd_levels(xx, geom, N, out=out)
            head_tape[t] = (saved, xx)
        if tape is not None:
            tape["head"] = head_tape
            tape["N"] = N
        return cls_out, box_out

    # ------------------------------------------------------------------------------------ backward
    def backward(self, fwd, dcls, dbox):
        """Back-propagate d(loss)/d(cls_out), d(loss)/d(box_out) through head, FPN and res5..res3,
        accumulating into the gradient arena (wgrad uses fp32 atomics, so repeated calls add up)."""
        tape, geom, N = fwd["tape"], fwd["geom"], fwd["N"]
        ops.COLSUM_QUEUE = []           # conv bias gradients: queued, one batched launch at the end
        try:
            dfeat = self.head_backward(tape, geom, N, dcls, dbox)
            lateral = self.fpn_backward(tape, geom, N, dfeat)
        finally:
            ops.colsum_flush()          # the trunk convolutions have no bias
        if self.grad_hook is not None:
            self.grad_hook(*self.grad_segments["top"])
        # trunk: res5 <- dC5 ; res4 <- dC4 + d(res5 input) ; res3 <- dC3 + d(res4 input)
        self.trunk_backward(tape, lateral)
        return None

    def head_backward(self, tape, geom, N, dcls, dbox):
        """-> d(loss)/d(pyramid) [N * L, 256] (the two towers' input gradients summed in the last dgrad epilogue)."""
        ctx = tape["head"]
        acc = None
        for t, pred, dout in (("cls_tower", self.cls_logits, dcls), ("bbox_tower", self.box_pred, dbox)):
            saved, top = ctx[t]
            pred.wgrad_levels(top, dout, geom, N)
            dx = pred.dgrad_levels(dout, geom, N)
            for i in range(3, -1, -1):
                conv, gname = self.towers[t][i]
                xin, c, stats = saved[i]
                gam, bet, dgam, dbet = self.gn[gname]
                dc = ops.groupnorm_relu_levels_bwd(dx, c, geom, N, stats, gam, bet, dgam, dbet, dbias_prev=conv.db)
                conv.wgrad_levels(xin, dc, geom, N, bias_done=True)   # bias gradient came out of the GroupNorm backward
                dx = conv.dgrad_levels(dc, geom, N, residual=acc if i == 0 else None)
            acc = dx
        return acc

    def fpn_backward(self, tape, geom, N, dfeat):
        """-> {"res5": (lateral conv, d(lateral output), (H, W)), ...}: what trunk_backward needs to send the gradients down."""
        d3, d4, d5, d6, d7 = self.level_views(dfeat, geom, N, 256)
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
        # the laterals' data-gradients are launched by trunk_backward (fused with the stage-output ReLU backward)
        return {"res5": (self.fpn_lat[5], dl5, hw[2]), "res4": (self.fpn_lat[4], dl4, hw[1]), "res3": (self.fpn_lat[3], dl3, hw[0])}
