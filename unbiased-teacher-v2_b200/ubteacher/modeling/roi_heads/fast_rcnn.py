"""``FastRCNNFocaltLossBoundaryVarOutputLayers`` with the reference's interface
(ubteacher/modeling/roi_heads/fast_rcnn.py:715-1225): three linear predictors on the box-head feature
(cls_score 1024 -> 81, class-agnostic bbox_pred 1024 -> 4, bbox_pred_std 1024 -> 4), run as ONE zero-padded 96-row
tensor-core operator; ``losses`` = focal loss (gamma 1.5) + 'nlloss' box loss (supervised) / 'tsbetter' masked L1
(pseudo labels); ``inference`` = softmax, Box2BoxXYXYTransform.apply_deltas, clip, score threshold, per-class NMS, top-k,
plus ``pred_boxes_std``.

    layers = FastRCNNFocaltLossBoundaryVarOutputLayers(cfg, input_shape)
    scores, deltas, deltas_std = layers(x)                       # x: [R, 1024]
    losses = layers.losses((scores, deltas, deltas_std), proposals, branch)      # proposals: list[Instances]
    instances, kept = layers.inference((scores, deltas, deltas_std), proposals)

forward and losses are differentiable (one autograd node each: the engine's wgrad / dgrad and loss-backward kernels).
"""
import torch

from ... import ops_rcnn as R
from ...d2compat.structures import Boxes, Instances
from ..box_regression import Box2BoxXYXYTransform
from ..views import ArenaView

BF16 = torch.bfloat16
PREFIX = "roi_heads.box_predictor."


def _own_engine(cfg):
    from ..rcnn_engine import RcnnEngine
    dev = cfg.MODEL.DEVICE
    if not str(dev).startswith("cuda"):
        raise RuntimeError("the UT2 B200 path has no CPU fallback: MODEL.DEVICE must be a CUDA device")
    return RcnnEngine(cfg, device=dev, seed=max(cfg.SEED, 0))


def pack_rois(proposals, device, pseudo):
    """list[Instances] (proposal_boxes, gt_classes[, gt_boxes, gt_loc_std]) -> the fixed-capacity sample dict of the loss
    kernels ([N, Rcap, ...] + count) and the row index of every proposal inside it."""
    N = len(proposals)
    cnt = [len(p) for p in proposals]
    Rcap = max(max(cnt + [1]), 1)
    s = {"proposal_boxes": torch.zeros((N, Rcap, 4), dtype=torch.float32, device=device),
         "gt_classes": torch.zeros((N, Rcap), dtype=torch.int64, device=device),
         "gt_boxes": torch.zeros((N, Rcap, 4), dtype=torch.float32, device=device),
         "gt_loc_std": torch.zeros((N, Rcap, 4), dtype=torch.float32, device=device),
         "count": torch.tensor(cnt, dtype=torch.int32).to(device)}
    rows = []
    for i, p in enumerate(proposals):
        n = cnt[i]
        rows.append(torch.arange(n, device=device) + i * Rcap)
        if not n:
            continue
        pb = p.proposal_boxes.tensor.to(device).float()
        s["proposal_boxes"][i, :n] = pb
        if p.has("gt_classes"):
            s["gt_classes"][i, :n] = p.gt_classes.to(device)
        s["gt_boxes"][i, :n] = p.gt_boxes.tensor.to(device).float() if p.has("gt_boxes") else pb     # fast_rcnn.py:876-882
        if pseudo and p.has("gt_loc_std"):
            s["gt_loc_std"][i, :n] = p.gt_loc_std.to(device).float()
    return s, torch.cat(rows) if rows else torch.zeros(0, dtype=torch.int64, device=device), Rcap


def pack_pred(predictions, rows, total_rows):
    scores, deltas, deltas_std = predictions
    pred = torch.zeros((total_rows, R.PRED_LD), dtype=BF16, device=scores.device)
    pred[rows, :81] = scores.detach().to(BF16)
    pred[rows, 81:85] = deltas.detach().to(BF16)
    pred[rows, 85:89] = deltas_std.detach().to(BF16)
    return pred


class _PredFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, x, trigger):
        eng = layer.engine
        xb = x.detach().reshape(1, 1, x.shape[0], -1).to(BF16).contiguous()
        pred = eng.box_pred.fwd(xb).view(x.shape[0], R.PRED_LD)
        ctx.layer, ctx.xb, ctx.dtype = layer, xb, x.dtype
        p = pred.float()
        return p[:, :81], p[:, 81:85], p[:, 85:89]

    @staticmethod
    def backward(ctx, g_sc, g_dl, g_st):
        eng = ctx.layer.engine
        Rn = ctx.xb.shape[2]
        dpred = torch.zeros((1, 1, Rn, R.PRED_LD), dtype=BF16, device=eng.device)
        v = dpred.view(Rn, R.PRED_LD)
        if g_sc is not None:
            v[:, :81] = g_sc.to(BF16)
        if g_dl is not None:
            v[:, 81:85] = g_dl.to(BF16)
        if g_st is not None:
            v[:, 85:89] = g_st.to(BF16)
        eng.box_pred.wgrad(ctx.xb, dpred)
        dx = eng.box_pred.dgrad(dpred, (1, Rn))
        return None, dx.view(Rn, -1).to(ctx.dtype), None


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, s, mode, rows, Rtot, trigger, scores, deltas, deltas_std):
        pred = pack_pred((scores, deltas, deltas_std), rows, Rtot)
        losses = R.fastrcnn_loss_fwd(pred, s, mode, layer.box2box_transform.weights[:2], gamma=layer.focal_gamma,
                                     ts_better=layer.ts_better, t_cert=layer.t_cert)
        ctx.layer, ctx.s, ctx.mode, ctx.rows, ctx.pred = layer, s, mode, rows, pred
        ctx.dtypes = (scores.dtype, deltas.dtype, deltas_std.dtype)
        return losses.clone()

    @staticmethod
    def backward(ctx, g):
        layer = ctx.layer
        dpred = R.fastrcnn_loss_bwd(ctx.pred, ctx.s, ctx.mode, g.contiguous().float(), layer.box2box_transform.weights[:2],
                                    gamma=layer.focal_gamma, ts_better=layer.ts_better, t_cert=layer.t_cert)
        d = dpred[ctx.rows].float()
        return (None, None, None, None, None, None, d[:, :81].to(ctx.dtypes[0]), d[:, 81:85].to(ctx.dtypes[1]),
                d[:, 85:89].to(ctx.dtypes[2]))


class FastRCNNFocaltLossBoundaryVarOutputLayers(ArenaView):
    def __init__(self, cfg, input_shape=None, engine=None):
        super().__init__(engine if engine is not None else _own_engine(cfg), PREFIX)
        m = cfg.MODEL
        assert m.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG and m.ROI_HEADS.NUM_CLASSES == 80
        self.num_classes = m.ROI_HEADS.NUM_CLASSES
        self.box2box_transform = Box2BoxXYXYTransform(weights=m.ROI_BOX_HEAD.BBOX_REG_WEIGHTS)
        self.smooth_l1_beta = m.ROI_BOX_HEAD.SMOOTH_L1_BETA
        self.test_score_thresh, self.test_nms_thresh = m.ROI_HEADS.SCORE_THRESH_TEST, m.ROI_HEADS.NMS_THRESH_TEST
        self.test_topk_per_image = cfg.TEST.DETECTIONS_PER_IMAGE
        self.box_reg_loss_type = m.ROI_BOX_HEAD.BBOX_REG_LOSS_TYPE
        self.box_pseudo_reg_loss_type = m.ROI_BOX_HEAD.BBOX_PSEUDO_REG_LOSS_TYPE
        self.loss_weight = {"loss_box_reg": m.ROI_BOX_HEAD.BBOX_REG_LOSS_WEIGHT}
        self.ts_better, self.t_cert = cfg.SEMISUPNET.TS_BETTER, cfg.SEMISUPNET.T_CERT
        self.focal_gamma = 1.5                           # fast_rcnn.py:925-936 (FocalLoss(gamma=1.5))
        self._trigger = torch.zeros(1, device=self.engine.device, requires_grad=True)

    def forward(self, x):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        if torch.is_grad_enabled():
            return _PredFn.apply(self, x, self._trigger)
        xb = x.reshape(1, 1, x.shape[0], -1).to(BF16).contiguous()
        p = self.engine.box_pred.fwd(xb).view(x.shape[0], R.PRED_LD).float()
        return p[:, :81], p[:, 81:85], p[:, 85:89]

    def losses(self, predictions, proposals, branch):
        dev = self.engine.device
        pseudo = branch == "unsup_data_train"
        s, rows, Rcap = pack_rois(proposals, dev, pseudo)
        mode = 1 if pseudo else 0
        Rtot = len(proposals) * Rcap
        if torch.is_grad_enabled():
            vec = _LossFn.apply(self, s, mode, rows, Rtot, self._trigger, *predictions)
        else:
            vec = R.fastrcnn_loss_fwd(pack_pred(predictions, rows, Rtot), s, mode, self.box2box_transform.weights[:2],
                                      gamma=self.focal_gamma, ts_better=self.ts_better, t_cert=self.t_cert)
        losses = {"loss_cls": vec[0], "loss_box_reg": vec[1]}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    @torch.no_grad()
    def inference(self, predictions, proposals):
        dev = self.engine.device
        s, rows, Rcap = pack_rois(proposals, dev, False)
        N = len(proposals)
        pred = pack_pred(predictions, rows, N * Rcap)
        sizes = [tuple(p.image_size) for p in proposals]
        dets = R.fastrcnn_inference(pred, s["proposal_boxes"], s["count"], self.engine.image_hw(sizes), self.test_score_thresh,
                                    self.test_nms_thresh, self.test_topk_per_image, self.box2box_transform.weights[:2])
        cnt = dets["count"].cpu().tolist()
        out, kept = [], []
        for i, n in enumerate(cnt):
            inst = Instances(sizes[i])
            inst.pred_boxes = Boxes(dets["pred_boxes"][i, :n])
            inst.scores = dets["scores"][i, :n]
            inst.pred_classes = dets["pred_classes"][i, :n]
            inst.pred_boxes_std = dets["reg_pred_std"][i, :n]
            out.append(inst)
            kept.append(dets["roi_index"][i, :n].long())
        return out, kept

    def predict_boxes(self, predictions, proposals):
        _, deltas, _ = predictions
        boxes = torch.cat([p.proposal_boxes.tensor for p in proposals]).to(self.engine.device)
        return self.box2box_transform.apply_deltas(deltas, boxes).split([len(p) for p in proposals])

    def predict_boxes_std(self, predictions, proposals):
        return predictions[2].split([len(p) for p in proposals])

    def predict_probs(self, predictions, proposals):
        return torch.softmax(predictions[0].float(), dim=-1).split([len(p) for p in proposals])
