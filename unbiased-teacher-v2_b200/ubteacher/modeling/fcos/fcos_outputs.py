"""Loss / proposal front-end over the dense FCOS head outputs — the B200 counterpart of the reference's
``FCOSOutputs`` (ubteacher/modeling/fcos/fcos_outputs.py:132). Method names follow the reference
(``losses``, ``pseudo_losses``, ``predict_proposals``); the arithmetic lives in csrc/fcos_loss.cu and
csrc/proposals.cu. Everything is device-resident: ground truth / pseudo-label sets are fixed-capacity
``[N, G, ...]`` tensors + a count vector, losses are a 4-vector (cls, loc, ctr, teacher_better_student).
"""
import torch

from ... import ops
from ...d2compat import comm
from ...d2compat.structures import Boxes, Instances

GT_CAP = 128


class BoxSet:
    """Fixed-capacity per-image box sets on the device (ground truth or pseudo labels)."""

    def __init__(self, boxes, classes, counts, reg_pred_std=None, scores=None, extra=None):
        self.boxes, self.classes, self.counts = boxes, classes, counts
        self.reg_pred_std, self.scores = reg_pred_std, scores
        self.extra = extra or {}

    @property
    def N(self):
        return self.boxes.shape[0]

    @staticmethod
    def from_instances(instances, device, cap=GT_CAP):
        """Packs list[Instances] (fields gt_boxes, gt_classes[, scores, reg_pred_std]) on the host and ships it
        with one pinned H2D copy per field."""
        N = len(instances)
        n_max = max([len(x) for x in instances] + [1])
        G = max(cap, (n_max + 15) // 16 * 16)
        b = torch.zeros((N, G, 4), dtype=torch.float32).pin_memory()
        c = torch.zeros((N, G), dtype=torch.int64).pin_memory()
        cnt = torch.zeros(N, dtype=torch.int32).pin_memory()
        has_std = N > 0 and all(x.has("reg_pred_std") for x in instances)
        s = torch.zeros((N, G, 4), dtype=torch.float32).pin_memory() if has_std else None
        for i, inst in enumerate(instances):
            n = len(inst)
            cnt[i] = n
            if n:
                b[i, :n] = inst.gt_boxes.tensor.detach().float().cpu()
                c[i, :n] = inst.gt_classes.detach().cpu()
                if has_std:
                    s[i, :n] = inst.reg_pred_std.detach().float().cpu()
        to = lambda t: t.to(device, non_blocking=True) if t is not None else None
        return BoxSet(to(b), to(c), to(cnt), to(s))

    def to_instances(self, image_sizes):
        """Device -> list[Instances] (forces a host sync; API-compat path only, not used by the trainer)."""
        cnt = self.counts.cpu().tolist()
        out = []
        for i, n in enumerate(cnt):
            inst = Instances(tuple(image_sizes[i]))
            inst.gt_boxes = Boxes(self.boxes[i, :n])
            inst.gt_classes = self.classes[i, :n]
            if self.scores is not None:
                inst.scores = self.scores[i, :n]
            if self.reg_pred_std is not None:
                inst.reg_pred_std = self.reg_pred_std[i, :n]
            for k, v in self.extra.items():
                inst.set(k, v[i, :n])
            out.append(inst)
        return out


def as_boxset(gt, device):
    if isinstance(gt, BoxSet):
        return gt
    return BoxSet.from_instances(gt, device)


class LossCtx:
    __slots__ = ("mode", "tg", "acc", "losses")

    def __init__(self, mode, tg, acc, losses):
        self.mode, self.tg, self.acc, self.losses = mode, tg, acc, losses


class FCOSOutputs:
    def __init__(self, cfg):
        f = cfg.MODEL.FCOS
        self.alpha, self.gamma = f.LOSS_ALPHA, f.LOSS_GAMMA
        self.num_classes = f.NUM_CLASSES
        self.strides = f.FPN_STRIDES
        self.kl_w = f.KLLOSS_WEIGHT
        self.pre_nms_thresh_train, self.pre_nms_topk_train = f.INFERENCE_TH_TRAIN, f.PRE_NMS_TOPK_TRAIN
        self.post_nms_topk_train = f.POST_NMS_TOPK_TRAIN
        self.pre_nms_thresh_test, self.pre_nms_topk_test = f.INFERENCE_TH_TEST, f.PRE_NMS_TOPK_TEST
        self.post_nms_topk_test = f.POST_NMS_TOPK_TEST
        self.nms_thresh = f.NMS_TH
        self.ts_better, self.ts_cert = cfg.SEMISUPNET.TS_BETTER, cfg.SEMISUPNET.TS_BETTER_CERT
        self.reg_unsup_loss = cfg.SEMISUPNET.CONSIST_REG_LOSS
        self.center_radius = float(f.POS_RADIUS) if f.CENTER_SAMPLE else 0.0     # get_sample_region (fcos_outputs.py:700-770)
        assert f.KL_LOSS_TYPE in ("nlloss", "klloss") and f.LOC_LOSS_TYPE == "giou" and f.QUALITY_EST == "centerness" and \
            cfg.SEMISUPNET.CLS_LOSS_METHOD == "focal" and not f.THRESH_WITH_CTR and not cfg.SEMISUPNET.SOFT_CLS_LABEL, \
            "the B200 loss kernels implement the shipped UT2 recipe (focal / centerness / nlloss|klloss + giou)"
        # supervised uncertainty term (fcos_outputs.py:377-416): NLLoss, or KLLoss reduced by LOC_FUN_ALL
        self.kl_mode = ops.KL_MODES["nlloss" if f.KL_LOSS_TYPE == "nlloss" else f.LOC_FUN_ALL]
        assert self.reg_unsup_loss == "ts_locvar_better_nms_nll_l1"
        self.training = True

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    # ---------------------------------------------------------------- losses
    def _targets(self, fwd, boxset, ignore_near=False):
        tg = ops.fcos_assign_targets(fwd["geom"], fwd["N"], boxset.boxes, boxset.classes, boxset.counts,
                                     boxset.reg_pred_std, self.num_classes, self.center_radius, ignore_near)
        world = comm.get_world_size()
        if world > 1:  # fcos_outputs.py:319-321,362 — normalisers are world-averaged; one tiny all-reduce
            torch.distributed.all_reduce(tg["norm"])
        return tg, float(world)

    def _fwd(self, fwd, scales, boxset, mode, ignore_near=False):
        tg, world = self._targets(fwd, boxset, ignore_near)
        losses, acc = ops.fcos_loss_fwd(fwd["geom"], fwd["N"], fwd["cls_out"], fwd["box_out"], scales, tg, mode,
                                        self.alpha, self.gamma, self.kl_w, self.ts_better, self.ts_cert, world,
                                        self.num_classes, self.kl_mode)
        return LossCtx(mode, tg, acc, losses)

    def losses(self, fwd, scales, gt, ignore_near=False):
        """Supervised branch (fcos_outputs.py:212-305 + :307-444). Returns (dict of loss scalars, ctx). ``ignore_near`` drops
        the locations inside a box but off its sample region from the classification loss (keep_locations, :310-311)."""
        ctx = self._fwd(fwd, scales, gt, 0, ignore_near)
        L = ctx.losses
        return {"loss_fcos_cls": L[0], "loss_fcos_loc": L[1], "loss_fcos_ctr": L[2]}, [ctx]

    def pseudo_losses(self, fwd, scales, gt_cls, gt_reg, ignore_near=False):
        """Unsupervised branch (fcos_outputs.py:447-490 + :492-631). ``ignore_near`` (SEMISUPNET.PSEUDO_CLS_IGNORE_NEAR) only
        changes the keep_locations the targets carry: fcos_pseudo_losses never reads them (:492-631), here as there."""
        c1 = self._fwd(fwd, scales, gt_cls, 1, ignore_near)
        c2 = self._fwd(fwd, scales, gt_reg, 2, ignore_near)
        return {"loss_fcos_cls": c1.losses[0], "loss_fcos_ctr": c1.losses[2], "teacher_better_student": c2.losses[3],
                "loss_fcos_loc": c2.losses[1]}, [c1, c2]

    def backward(self, fwd, scales, dscales, ctxs, gouts):
        """d(sum_i gout_i . losses_i) / d(cls_out, box_out). `gouts`: one float[4] device tensor per ctx."""
        dcls = torch.empty_like(fwd["cls_out"])
        dbox = torch.empty_like(fwd["box_out"])
        world = float(comm.get_world_size())
        first = True
        for ctx, gout in zip(ctxs, gouts):
            ops.fcos_loss_bwd(fwd["geom"], fwd["N"], fwd["cls_out"], fwd["box_out"], scales, ctx.tg, ctx.mode, self.alpha,
                              self.gamma, self.kl_w, self.ts_better, self.ts_cert, world, ctx.acc, gout,
                              dcls if ctx.mode != 2 else None, dbox, dscales, self.num_classes, accumulate=not first,
                              kl_mode=self.kl_mode)
            first = False
        return dcls, dbox

    # ---------------------------------------------------------------- proposals
    def predict_proposals(self, fwd, scales, nms_method="cls_n_ctr", out_cap=128):
        """fcos_outputs.py:1046-1132: returns the device-resident detections dict ([N, out_cap, ...] + count)."""
        if self.training:
            thr, pre, post = self.pre_nms_thresh_train, self.pre_nms_topk_train, self.post_nms_topk_train
        else:
            thr, pre, post = self.pre_nms_thresh_test, self.pre_nms_topk_test, self.post_nms_topk_test
        dets = ops.fcos_predict_proposals(fwd["geom"], fwd["N"], fwd["cls_out"], fwd["box_out"], scales, nms_method, thr,
                                          pre, self.nms_thresh, post, out_cap, self.num_classes)
        dets["image_sizes"] = fwd["image_sizes"]
        return dets


def dets_to_instances(dets):
    """Detections dict -> list[Instances] with the reference's field names (host sync; compat path)."""
    cnt = dets["count"].cpu().tolist()
    out = []
    for i, n in enumerate(cnt):
        inst = Instances(tuple(dets["image_sizes"][i]))
        inst.pred_boxes = Boxes(dets["pred_boxes"][i, :n])
        for k in ("scores", "pred_classes", "centerness", "cls_confid", "reg_pred_std", "locations", "fpn_levels"):
            inst.set(k, dets[k][i, :n])
        out.append(inst)
    return out
