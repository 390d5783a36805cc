"""B200-native execution engine for the Faster R-CNN R50-FPN detector of Unbiased Teacher v2 (SURVEY.md §8 rows
a20-a24). Same design as fcos_engine.py: one object = one replica (student or teacher), a flat parameter arena,
packed bf16 tensor-core operands, an explicit forward / backward schedule over the C-ABI kernels.

Reference modules restated here (paths under /root/reference/ubteacher, [D2] = Detectron2 v0.6):
  meta-arch   modeling/meta_arch/rcnn.py:8-72 (three branches: supervised, unsup_data_weak, unsup_data_train)
  backbone    configs/Faster-RCNN/Base-RCNN-FPN.yaml:4 -> [D2] build_resnet_fpn_backbone (p2..p5 + LastLevelMaxPool p6)
  RPN         modeling/proposal_generator/rpn.py:21-225 ([D2] StandardRPNHead: conv3x3+ReLU, 1x1 -> 3, 1x1 -> 12)
  ROI heads   modeling/roi_heads/roi_heads.py:75-270 (sampling, [D2] ROIPooler, FastRCNNConvFCHead 2 x FC1024)
  predictor   modeling/roi_heads/fast_rcnn.py:715-1225 (cls_score 81 | bbox_pred 4 | bbox_pred_std 4, class-agnostic)

Tensor-core work: every conv, and the three FC layers as 1x1 "convs" over a [1, 1, R, K] activation (fc1: K = 12544 =
7*7*256 in (ph, pw, c) order — the ROIAlign kernel writes exactly that layout, and the fc1 master weight is stored
[1024, 7, 7, 256], exported to checkpoints as the reference's [1024, 256*7*7]). The two RPN predictors run as one
16-row operator and the three box predictors as one 96-row operator (zero-padded, see EngineBase.add_fused).
"""
import math
import os

import torch

from .. import ops, ops_rcnn as R
from ..arena import Spec
from .engine_base import BF16, EngineBase


class RcnnEngine(EngineBase):
    def __init__(self, cfg, device="cuda", seed=0, init=True):
        self._begin(cfg, device)
        m = cfg.MODEL
        assert m.ROI_HEADS.NUM_CLASSES == 80 and m.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG, \
            "the B200 engine implements the shipped UT2 Faster R-CNN recipe (80 classes, class-agnostic box regression)"
        assert m.ROI_HEADS.LOSS == "FocalLoss_BoundaryVar" and m.ROI_BOX_HEAD.BBOX_REG_LOSS_TYPE == "nlloss"
        assert m.ROI_BOX_HEAD.BBOX_PSEUDO_REG_LOSS_TYPE == "tsbetter" and m.ROI_BOX_HEAD.POOLER_RESOLUTION == 7
        assert m.ROI_BOX_HEAD.NUM_FC == 2 and m.ROI_BOX_HEAD.FC_DIM == 1024 and m.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO == 0
        assert list(m.RPN.IN_FEATURES) == ["p2", "p3", "p4", "p5", "p6"] and list(m.ROI_HEADS.IN_FEATURES) == ["p2", "p3", "p4", "p5"]
        self.num_classes = 80
        self.rpn_strides = [4, 8, 16, 32, 64]
        self.anchor_sizes = [s[0] for s in m.ANCHOR_GENERATOR.SIZES]
        self.anchor_ratios = list(m.ANCHOR_GENERATOR.ASPECT_RATIOS[0])
        r = m.RPN
        self.rpn_batch, self.rpn_pos_frac = r.BATCH_SIZE_PER_IMAGE, r.POSITIVE_FRACTION
        self.rpn_thr = tuple(r.IOU_THRESHOLDS)
        self.rpn_pre, self.rpn_post, self.rpn_nms = r.PRE_NMS_TOPK_TRAIN, r.POST_NMS_TOPK_TRAIN, r.NMS_THRESH
        self.rpn_pre_test, self.rpn_post_test = r.PRE_NMS_TOPK_TEST, r.POST_NMS_TOPK_TEST
        h = m.ROI_HEADS
        self.roi_batch, self.roi_pos_frac, self.roi_iou = h.BATCH_SIZE_PER_IMAGE, h.POSITIVE_FRACTION, h.IOU_THRESHOLDS[0]
        self.append_gt = h.PROPOSAL_APPEND_GT
        self.test_score, self.test_nms, self.test_topk = h.SCORE_THRESH_TEST, h.NMS_THRESH_TEST, cfg.TEST.DETECTIONS_PER_IMAGE
        self.box_w = tuple(m.ROI_BOX_HEAD.BBOX_REG_WEIGHTS)[:2]
        self.ts_better, self.t_cert = cfg.SEMISUPNET.TS_BETTER, cfg.SEMISUPNET.T_CERT
        self.seed = int(seed)
        self.draws = 0           # sampling-key stream position ([D2] subsample_labels draws from the global torch RNG)
        # device-resident word mixed into every sampling seed: a captured CUDA graph freezes the host-side `draws`
        # values, the trainer bumps this word before each replay so the draws keep changing (include/ut2.h, seed_dev)
        self.seed_dev = torch.zeros(1, dtype=torch.int32, device=device)
        self._image_hw = {}
        self._build()
        if init:
            self.init_weights(seed)
        self.refresh_operands()

    # ------------------------------------------------------------------------------------ construction
    def _build(self):
        add_conv = self.add_conv
        self.build_trunk()
        self.fpn_lat, self.fpn_out = {}, {}
        for lvl, c in ((2, 256), (3, 512), (4, 1024), (5, 2048)):
            self.fpn_lat[lvl] = add_conv(f"backbone.fpn_lateral{lvl}", c, 256, 1, 1, 0, False, True, True, need_dgrad=lvl > 2)
            self.fpn_out[lvl] = add_conv(f"backbone.fpn_output{lvl}", 256, 256, 3, 1, 1, False, True, True)
        rp = "proposal_generator.rpn_head."
        self.rpn_conv = add_conv(rp + "conv", 256, 256, 3, 1, 1, False, True, True)
        self.rpn_pred = self.add_fused(rp + "pred_fused", 256, 1, 0, [(rp + "objectness_logits", 3), (rp + "anchor_deltas", 12)], 16)
        bh = "roi_heads.box_head."
        self.fc1 = add_conv(bh + "fc1", 12544, 1024, 1, 1, 0, False, True, True, shape=(1024, 256, 7, 7))
        self.fc2 = add_conv(bh + "fc2", 1024, 1024, 1, 1, 0, False, True, True, shape=(1024, 1024))
        bp = "roi_heads.box_predictor."
        self.box_pred = self.add_fused(bp + "fused", 1024, 1, 0, [(bp + "cls_score", 81), (bp + "bbox_pred", 4), (bp + "bbox_pred_std", 4)],
                                       96, linear=True)
        # [D2] v0.6: pixel_mean / pixel_std are non-persistent buffers of GeneralizedRCNN
        self.finish([Spec("pixel_mean", (3, 1, 1), "buffer", persistent=False), Spec("pixel_std", (3, 1, 1), "buffer", persistent=False)])
        self.arena.export_shape = {bh + "fc1.weight": (1024, 12544)}

    def init_weights(self, seed=0):
        """Seeded [D2]-style init: MSRA trunk, c2_xavier FPN and box-head FCs, N(0, .01) RPN, N(0, .01 / .001 / .0001)
        box predictors (fast_rcnn.py:766-771), zero biases, FrozenBN = identity."""
        g = torch.Generator().manual_seed(seed)
        for name, v in self.arena.views.items():
            if name.startswith("_"):
                continue
            t = self.init_trunk_entry(name, v, g)
            if t is None:
                t = torch.zeros(v.shape)
                if name.endswith(".weight"):
                    if name.startswith("proposal_generator.rpn_head."):
                        t.normal_(0, 0.01, generator=g)
                    elif name.startswith("roi_heads.box_head."):
                        fan_in = t[0].numel()
                        bound = math.sqrt(3.0 / fan_in)
                        t.uniform_(-bound, bound, generator=g)
                    elif name.endswith("cls_score.weight"):
                        t.normal_(0, 0.01, generator=g)
                    elif name.endswith("bbox_pred.weight"):
                        t.normal_(0, 0.001, generator=g)
                    elif name.endswith("bbox_pred_std.weight"):
                        t.normal_(0, 0.0001, generator=g)
            v.copy_(t.to(self.device))

    # ------------------------------------------------------------------------------------ geometry
    def level_geom(self, Hp, Wp):
        hw = []
        h, w = Hp // 4, Wp // 4
        for i in range(5):
            hw.append((h, w))
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        return R.RpnGeom(hw, self.rpn_strides, self.anchor_sizes, self.anchor_ratios), \
            R.RoiGeom(hw[:4], [1.0 / s for s in self.rpn_strides[:4]])

    def _next_seed(self):
        self.draws += 1
        return (self.seed * 0x9E3779B1 + self.draws * 0x85EBCA77) & 0xFFFFFFFF

    # ------------------------------------------------------------------------------------ forward
    def forward_features(self, images, train):
        """Backbone + FPN + RPN head. Returns the fwd dict: rpn_out (level-major [P,16] bf16), FPN levels, geometry."""
        N = len(images)
        tape = {} if train else None
        feats, sizes, (Hp, Wp) = self.trunk_forward(images, train, tape)
        geom, rgeom = self.level_geom(Hp, Wp)
        feat, levels = self.fpn_forward(feats, geom, N, tape)
        rpn_out = self.rpn_head_forward(feat, geom, N, tape)
        return {"rpn_out": rpn_out, "levels": levels, "geom": geom, "rgeom": rgeom, "N": N, "image_sizes": sizes,
                "image_hw": self.image_hw(sizes, images), "tape": tape, "padded": (Hp, Wp), "feat": feat}

    @staticmethod
    def level_views(buf, geom, N, C):
        return [buf[geom.off[l] * N: geom.off[l + 1] * N].view(N, geom.hw[l][0], geom.hw[l][1], C) for l in range(geom.num)]

    def fpn_forward(self, feats, geom, N, tape=None):
        """[D2] FPN over res2..res5 + LastLevelMaxPool. FPN outputs p2..p6 live in ONE level-major buffer
        [N * sum(H_l W_l), 256]: the RPN head shares its weights across levels (rpn.py:31), so its 3x3 conv and the fused
        predictor are one launch each over the pyramid. Returns (buffer, per-level NHWC views)."""
        c2, c3, c4, c5 = feats["res2"], feats["res3"], feats["res4"], feats["res5"]
        lat5 = self.fpn_lat[5].fwd(c5)
        lat4 = self.fpn_lat[4].fwd(c4, residual=lat5, res_up2=True)
        lat3 = self.fpn_lat[3].fwd(c3, residual=lat4, res_up2=True)
        lat2 = self.fpn_lat[2].fwd(c2, residual=lat3, res_up2=True)
        feat = torch.empty((geom.L * N, 256), dtype=BF16, device=self.device)
        levels = self.level_views(feat, geom, N, 256)
        self.fpn_out[5].fwd(lat5, out=levels[3])
        self.fpn_out[4].fwd(lat4, out=levels[2])
        self.fpn_out[3].fwd(lat3, out=levels[1])
        self.fpn_out[2].fwd(lat2, out=levels[0])
        R.subsample2x(levels[3], out=levels[4])      # [D2] LastLevelMaxPool: max_pool2d(kernel 1, stride 2)
        if tape is not None:
            tape["fpn"] = (c2, c3, c4, c5, lat2, lat3, lat4, lat5)
        return feat, levels

    def rpn_head_forward(self, feat, geom, N, tape=None):
        """[D2] StandardRPNHead over the level-major pyramid -> rpn_out [N * L, 16] = 3 objectness | 12 deltas | pad."""
        hidden = self.rpn_conv.fwd_levels(feat, geom, N, relu=True)
        rpn_out = torch.empty((geom.L * N, R.RPN_LD), dtype=BF16, device=self.device)
        self.rpn_pred.fwd_levels(hidden, geom, N, out=rpn_out)
        if tape is not None:
            tape["rpn_hidden"] = hidden
            tape["feat"] = feat
        return rpn_out

    def image_hw(self, sizes, images=None):
        if images is not None:                      # static inputs of a CUDA-graph step: the sizes live in device memory
            hw = ops.lookup_static(ops.STATIC_HW, images[0].data_ptr())
            if hw is not None:
                return hw
        key = tuple(map(tuple, sizes))
        if key not in self._image_hw:               # cached: no pinned temporary inside a graph capture
            self._image_hw[key] = torch.tensor(sizes, dtype=torch.float32).pin_memory().to(self.device, non_blocking=True)
        return self._image_hw[key]

    def proposals(self, fwd, test=False):
        """[D2] find_top_rpn_proposals (rpn.py:72-74). The teacher is never put in eval mode (trainer.py:628-629,
        SURVEY A.3 #8), so inside a training step both replicas use the *_TRAIN top-k values; `test` selects the *_TEST
        values of a model in eval mode."""
        pre, post = (self.rpn_pre_test, self.rpn_post_test) if test else (self.rpn_pre, self.rpn_post)
        return R.rpn_predict_proposals(fwd["geom"], fwd["N"], fwd["rpn_out"], fwd["image_hw"], pre, post, self.rpn_nms)

    def box_head(self, fwd, rois, roi_cnt, train):
        """ROIAlign over p2..p5 -> fc1 -> ReLU -> fc2 -> ReLU -> fused predictor. rois [N, Rcap, 4]."""
        N, Rcap = rois.shape[:2]
        pooled = R.roi_align_fwd(fwd["rgeom"], fwd["levels"][:4], rois, roi_cnt)          # [N*Rcap, 7, 7, 256]
        x = pooled.view(1, 1, N * Rcap, 12544)
        h1 = self.fc1.fwd(x, relu=True)
        h2 = self.fc2.fwd(h1, relu=True)
        pred = self.box_pred.fwd(h2).view(N * Rcap, R.PRED_LD)
        if train:
            fwd["tape"]["box"] = (x, h1, h2)
        return pred

    def forward_losses(self, fwd, gt, pseudo):
        """supervised / unsup_data_train branches (rcnn.py:26-40, :57-72). gt: BoxSet (boxes, classes, counts and, for
        pseudo labels, scores + reg_pred_std = the teacher's pred_boxes_std). Returns (losses float[4] parts, ctx)."""
        # anchor labelling + RPN losses (latency-bound: one CTA per image for the selections) do not feed the proposal -> ROI
        # sampling -> box head chain: they run on a second stream (forked / joined here, capturable) next to it
        side = self._side_stream()
        main = torch.cuda.current_stream()
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                rpn_losses, ctx = self.rpn_losses(fwd, gt, pseudo)
        else:
            rpn_losses, ctx = self.rpn_losses(fwd, gt, pseudo)
        props = self.proposals(fwd)
        roi_losses, rctx = self.roi_losses(fwd, props, gt, pseudo)
        if side is not None:
            main.wait_stream(side)
        ctx.update(rctx)
        ctx["proposals"] = props
        return rpn_losses, roi_losses, ctx

    def _side_stream(self):
        if os.environ.get("UT2_RPN_SIDE_STREAM", "1") == "0":
            return None
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def rpn_losses(self, fwd, gt, pseudo):
        """PseudoLabRPN: anchor labelling + sampling, objectness BCE (x teacher score for pseudo labels) and L1 box loss."""
        N, geom = fwd["N"], fwd["geom"]
        scores = gt.scores if pseudo else None
        dk = getattr(self, "debug_keys", None) or {}       # parity tests inject the sampling draws (else: hashed seed)
        labels, matched = R.rpn_label_anchors(geom, N, gt.boxes, gt.counts, keys=dk.get("rpn"), seed=self._next_seed(), batch=self.rpn_batch,
                                              pos_frac=self.rpn_pos_frac, lo=self.rpn_thr[0], hi=self.rpn_thr[1], seed_dev=self.seed_dev)
        losses = R.rpn_loss_fwd(geom, N, fwd["rpn_out"], labels, matched, gt.boxes, scores, gt.counts, self.rpn_batch)
        return losses, {"labels": labels, "matched": matched, "gt": gt, "scores": scores}

    def roi_losses(self, fwd, props, gt, pseudo):
        """StandardROIHeadsPseudoLab: append GT, match, sample 512 ROIs, ROIAlign, box head, focal + box losses."""
        dk = getattr(self, "debug_keys", None) or {}
        scores = gt.scores if pseudo else None
        s = R.roi_sample(props["proposal_boxes"], props["count"], gt.boxes, gt.classes, gt.counts, scores,
                         gt.reg_pred_std if pseudo else None, keys=dk.get("roi"), seed=self._next_seed(), batch=self.roi_batch,
                         pos_frac=self.roi_pos_frac, iou_thr=self.roi_iou, num_classes=self.num_classes, append_gt=self.append_gt,
                         seed_dev=self.seed_dev)
        pred = self.box_head(fwd, s["proposal_boxes"], s["count"], True)
        mode = 1 if pseudo else 0
        losses = R.fastrcnn_loss_fwd(pred, s, mode, self.box_w, ts_better=self.ts_better, t_cert=self.t_cert)
        return losses, {"sample": s, "pred": pred, "mode": mode}

    def forward_inference(self, fwd, test=False):
        """unsup_data_weak branch (rcnn.py:42-55): proposals -> box head on all of them -> fast_rcnn_inference."""
        props = self.proposals(fwd, test)
        pred = self.box_head(fwd, props["proposal_boxes"], props["count"], False)
        dets = R.fastrcnn_inference(pred, props["proposal_boxes"], props["count"], fwd["image_hw"], self.test_score,
                                    self.test_nms, self.test_topk, self.box_w)
        dets["image_sizes"] = fwd["image_sizes"]
        return props, dets, pred

    # ------------------------------------------------------------------------------------ backward
    def backward(self, fwd, ctx, gout_rpn, gout_roi):
        """gout_rpn / gout_roi: float[2] device tensors = d(total)/d{loss_rpn_cls, loss_rpn_loc} / d{loss_cls, loss_box_reg}.
        Accumulates into the gradient arena (wgrad uses fp32 atomics, so repeated calls add up)."""
        tape, geom, N = fwd["tape"], fwd["geom"], fwd["N"]
        ops.COLSUM_QUEUE = []           # conv bias gradients: queued, one batched launch at the end
        try:
            dlev = self.roi_backward(fwd, ctx, gout_roi)
            dfe = self.rpn_backward(fwd, ctx, gout_rpn)
            lateral = self.fpn_backward(tape, geom, N, dfe, dlev)
        finally:
            ops.colsum_flush()          # the trunk convolutions have no bias
        if self.grad_hook is not None:
            self.grad_hook(*self.grad_segments["top"])
        self.trunk_backward(tape, lateral)

    def roi_backward(self, fwd, ctx, gout_roi):
        """Box-head backward -> fp32 gradient maps of p2..p5 (ROIAlign backward accumulates with atomics)."""
        tape = fwd["tape"]
        s = ctx["sample"]
        Rtot = s["gt_classes"].numel()
        dpred = R.fastrcnn_loss_bwd(ctx["pred"], s, ctx["mode"], gout_roi, self.box_w, ts_better=self.ts_better,
                                    t_cert=self.t_cert).view(1, 1, Rtot, R.PRED_LD)
        dpool = self.box_head_backward(tape, dpred)
        dlev = [torch.zeros(f.shape, dtype=torch.float32, device=self.device) for f in fwd["levels"][:4]]
        R.roi_align_bwd(fwd["rgeom"], dlev, s["proposal_boxes"], s["count"], dpool)
        return dlev

    def box_head_backward(self, tape, dpred):
        """fused predictor <- fc2 <- fc1: weight gradients into the arena, returns d(pooled) [1, 1, R, 12544]."""
        Rtot = dpred.shape[2]
        x, h1, h2 = tape["box"]
        self.box_pred.wgrad(h2, dpred)
        dh2 = self.box_pred.dgrad(dpred, (1, Rtot), relu_mask=h2)
        self.fc2.wgrad(h1, dh2)
        dh1 = self.fc2.dgrad(dh2, (1, Rtot), relu_mask=h1)
        self.fc1.wgrad(x, dh1)
        return self.fc1.dgrad(dh1, (1, Rtot))

    def rpn_backward(self, fwd, ctx, gout_rpn):
        """RPN losses -> level-major gradient of the pyramid [N * L, 256]."""
        tape, geom, N = fwd["tape"], fwd["geom"], fwd["N"]
        gt = ctx["gt"]
        drpn = R.rpn_loss_bwd(geom, N, fwd["rpn_out"], ctx["labels"], ctx["matched"], gt.boxes, ctx["scores"], gt.counts,
                              gout_rpn, self.rpn_batch)
        t = tape["rpn_hidden"]
        self.rpn_pred.wgrad_levels(t, drpn, geom, N)
        dt = self.rpn_pred.dgrad_levels(drpn, geom, N, relu_mask=t)
        self.rpn_conv.wgrad_levels(tape["feat"], dt, geom, N)
        return self.rpn_conv.dgrad_levels(dt, geom, N)

    def fpn_backward(self, tape, geom, N, dfe, dlev):
        """dfe: level-major bf16 gradient of p2..p6 (RPN), dlev: fp32 gradients of p2..p5 (ROI heads) or None.
        -> {"res5": (lateral conv, d(lateral output), (H, W)), ...} for trunk_backward."""
        dP = []
        for l, g in enumerate(self.level_views(dfe, geom, N, 256)):
            dP.append(R.add_f32_bf16(dlev[l], g) if (l < 4 and dlev is not None) else g)
        d2, d3, d4, d5, d6 = dP
        hw = geom.hw
        d5 = ops.add_bf16(d5, ops.zero_stuff_s2(d6, hw[3][0], hw[3][1]))              # p6 = p5[:, ::2, ::2]
        c2, c3, c4, c5, lat2, lat3, lat4, lat5 = tape["fpn"]
        dl = {}
        for lvl, lat, d in ((5, lat5, d5), (4, lat4, d4), (3, lat3, d3), (2, lat2, d2)):
            self.fpn_out[lvl].wgrad(lat, d)
            dl[lvl] = self.fpn_out[lvl].dgrad(d, hw[lvl - 2])
        dl[3] = ops.downsample2x_sum(dl[2], dl[3])
        dl[4] = ops.downsample2x_sum(dl[3], dl[4])
        dl[5] = ops.downsample2x_sum(dl[4], dl[5])
        self.fpn_lat[2].wgrad(c2, dl[2])                                              # res2 is frozen: no dgrad below
        for lvl, c in ((3, c3), (4, c4), (5, c5)):
            self.fpn_lat[lvl].wgrad(c, dl[lvl])
        # the laterals' data-gradients are launched by trunk_backward (fused with the stage-output ReLU backward)
        return {"res5": (self.fpn_lat[5], dl[5], hw[3]), "res4": (self.fpn_lat[4], dl[4], hw[2]), "res3": (self.fpn_lat[3], dl[3], hw[1])}
