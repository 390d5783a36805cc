"""``StandardROIHeadsPseudoLab`` with the reference's call contract (ubteacher/modeling/roi_heads/roi_heads.py:23-270):

    roi_heads = ROI_HEADS_REGISTRY.get("StandardROIHeadsPseudoLab")(cfg, input_shape)
    training and compute_loss:  proposals_sampled, losses = roi_heads(images, features, proposals, targets, True, branch)
    otherwise                :  pred_instances, predictions = roi_heads(images, features, proposals, None, False, branch)

features: {"p2".."p5": [N, 256, H_l, W_l]}; proposals: list[Instances] (proposal_boxes); targets: list[Instances]
(gt_boxes, gt_classes; pseudo labels also ``scores`` and ``pred_boxes_std`` -> gt_confid / gt_loc_std, roi_heads.py:92-98).
``box_predictor`` is the ``FastRCNNFocaltLossBoundaryVarOutputLayers`` view (roi_heads.py:60-62). Executed by the engine:
append-GT + IoU matching + fg/bg sampling (csrc/roi.cu), ROIAlign over p2..p5, FC x 2 + fused predictor on the tensor
cores, focal / box losses or fast_rcnn_inference. The losses hang on one autograd node whose backward (loss backward, FC
dgrad / wgrad, ROIAlign backward) returns the feature-map gradients.
"""
import torch

from ... import ops_rcnn as R
from ...d2compat.registry import ROI_HEADS_REGISTRY
from ...d2compat.structures import Boxes, Instances
from ..views import ArenaView, nchw
from .fast_rcnn import FastRCNNFocaltLossBoundaryVarOutputLayers, _own_engine

BF16 = torch.bfloat16


def pack_proposals(proposals, device, cap=1000):
    N = len(proposals)
    cnt = [len(p) for p in proposals]
    Pcap = max(cap, max(cnt + [1]))
    boxes = torch.zeros((N, Pcap, 4), dtype=torch.float32, device=device)
    for i, p in enumerate(proposals):
        if cnt[i]:
            boxes[i, :cnt[i]] = p.proposal_boxes.tensor.to(device).float()
    return boxes, torch.tensor(cnt, dtype=torch.int32).to(device)


class _RoiLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, heads, fwd, lctx, losses, trigger, *xs):
        ctx.heads, ctx.fwd, ctx.lctx = heads, fwd, lctx
        ctx.dtypes = [x.dtype for x in xs]
        return losses.clone()

    @staticmethod
    def backward(ctx, g):
        eng, fwd = ctx.heads.engine, ctx.fwd
        dlev = eng.roi_backward(fwd, ctx.lctx, g.contiguous().float())
        fwd["tape"] = None
        return (None, None, None, None, None) + tuple(nchw(d).to(dt) for d, dt in zip(dlev, ctx.dtypes))


@ROI_HEADS_REGISTRY.register()
class StandardROIHeadsPseudoLab(ArenaView):
    def __init__(self, cfg, input_shape=None, engine=None):
        super().__init__(engine if engine is not None else _own_engine(cfg), "roi_heads.")
        self.box_in_features = list(cfg.MODEL.ROI_HEADS.IN_FEATURES)
        self.box_predictor = FastRCNNFocaltLossBoundaryVarOutputLayers(cfg, None, engine=self.engine)
        self.train_on_pred_boxes = cfg.MODEL.ROI_BOX_HEAD.TRAIN_ON_PRED_BOXES
        assert not self.train_on_pred_boxes
        self._trigger = torch.zeros(1, device=self.engine.device, requires_grad=True)

    def _fwd(self, features, sizes, train):
        eng = self.engine
        xs = [features[f] for f in self.box_in_features]
        levels = [x.detach().permute(0, 2, 3, 1).to(BF16).contiguous() for x in xs]       # no copy for the backbone view's maps
        rgeom = R.RoiGeom([tuple(x.shape[2:]) for x in xs], [1.0 / s for s in eng.rpn_strides[:4]])
        return xs, {"levels": levels, "rgeom": rgeom, "N": len(sizes), "image_sizes": sizes, "image_hw": eng.image_hw(sizes),
                    "tape": {} if train else None}

    def forward(self, images, features, proposals, targets=None, compute_loss=True, branch=""):
        from ..fcos.fcos_outputs import BoxSet
        from ..meta_arch.rcnn import detections_to_instances, rcnn_boxset
        eng = self.engine
        sizes = [tuple(p.image_size) for p in proposals]
        train = self.training and compute_loss
        xs, fwd = self._fwd(features, sizes, train)
        pb, pc = pack_proposals(proposals, eng.device)
        if train:
            assert targets
            gt = targets if isinstance(targets, BoxSet) else rcnn_boxset(targets, eng.device)
            pseudo = gt.scores is not None                       # roi_heads.py:92: targets carrying `scores` are pseudo labels
            vec, lctx = eng.roi_losses(fwd, {"proposal_boxes": pb, "count": pc}, gt, pseudo)
            if torch.is_grad_enabled():
                vec = _RoiLossFn.apply(self, fwd, lctx, vec, self._trigger, *xs)
            w = self.box_predictor.loss_weight
            losses = {"loss_cls": vec[0] * w.get("loss_cls", 1.0), "loss_box_reg": vec[1] * w.get("loss_box_reg", 1.0)}
            s = lctx["sample"]
            out = []
            for i, n in enumerate(s["count"].cpu().tolist()):  # the sampled proposals with their matched targets (roi_heads.py:156-199)
                inst = Instances(sizes[i])
                inst.proposal_boxes = Boxes(s["proposal_boxes"][i, :n])
                inst.gt_classes = s["gt_classes"][i, :n]
                inst.gt_boxes = Boxes(s["gt_boxes"][i, :n])
                if pseudo:
                    inst.gt_confid = s["gt_confid"][i, :n]
                    inst.gt_loc_std = s["gt_loc_std"][i, :n]
                out.append(inst)
            return out, losses
        pred = eng.box_head(fwd, pb, pc, False)
        dets = R.fastrcnn_inference(pred, pb, pc, fwd["image_hw"], eng.test_score, eng.test_nms, eng.test_topk, eng.box_w)
        dets["image_sizes"] = sizes
        rows = torch.cat([torch.arange(int(n), device=eng.device) + i * pb.shape[1] for i, n in enumerate(pc.cpu().tolist())])
        p = pred[rows].float()
        return detections_to_instances(dets), (p[:, :81], p[:, 81:85], p[:, 85:89])
