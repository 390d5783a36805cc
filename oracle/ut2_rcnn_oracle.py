"""TEST INFRASTRUCTURE ONLY — CPU oracle for the Faster R-CNN half of the Unbiased-Teacher-v2 hot path
(SURVEY.md §8 rows a2, a20-a24). Plain torch-CPU / numpy; every function cites the reference file:line it
follows (paths relative to /root/reference/ubteacher) or, for Detectron2 v0.6 pieces that are not on disk,
the SURVEY.md appendix-B item it restates ([D2]).

Pinning: the reference-owned logic (PseudoLabRPN.label_and_sample_anchors_pseudo / losses,
StandardROIHeadsPseudoLab.label_and_sample_proposals[_pseudo], FastRCNNFocaltLossBoundaryVarOutputLayers.losses /
inference, UBRCNNTeacherTrainer.threshold_bbox) is executed from the reference's own files by
oracle/gen_golden_rcnn.py (the [D2] base classes stubbed with the restatements below) and compared in
tests/test_oracle_rcnn_golden.py. The [D2] pieces themselves (Matcher, subsample_labels, anchors,
Box2BoxTransform, find_top_rpn_proposals, ROIPooler, fast_rcnn_inference) are "unpinned by the reference's own
tests": restated from published behaviour, roi_align / nms cross-checked against the installed torchvision.ops.

Randomness: [D2] subsample_labels draws torch.randperm on the device. Here (and in the CUDA kernels) the draw is
replaced by per-element 32-bit keys: the n elements with the smallest (key, index) are taken — the same
distribution as randperm()[:n] for i.i.d. keys, and it lets parity tests inject identical draws on both sides.
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .ut2_oracle import batched_nms, box2box_xyxy_apply_deltas, box2box_xyxy_get_deltas, matched_boxlist_iou, \
    rcnn_focal_loss, rcnn_nl_loss

SCALE_CLAMP = math.log(1000.0 / 16)          # [D2] Box2BoxTransform default
ANCHOR_SIZES = (32, 64, 128, 256, 512)
ANCHOR_RATIOS = (0.5, 1.0, 2.0)
RPN_STRIDES = (4, 8, 16, 32, 64)


# ------------------------------------------------------------------------------------------ anchors / transforms
def cell_anchors(size: float, ratios: Sequence[float] = ANCHOR_RATIOS) -> torch.Tensor:
    """[D2] DefaultAnchorGenerator.generate_cell_anchors: w = sqrt(area / r), h = r * w."""
    out = []
    area = float(size) ** 2.0
    for r in ratios:
        w = math.sqrt(area / r)
        h = r * w
        out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(out, dtype=torch.float32)


def level_anchors(h: int, w: int, stride: int, size: float) -> torch.Tensor:
    """[D2] DefaultAnchorGenerator._grid_anchors, offset 0: order (h, w, a)."""
    sx = torch.arange(0, w * stride, step=stride, dtype=torch.float32)
    sy = torch.arange(0, h * stride, step=stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
    return (shifts.view(-1, 1, 4) + cell_anchors(size).view(1, -1, 4)).reshape(-1, 4)


def generate_anchors(level_hw: Sequence[Tuple[int, int]], strides=RPN_STRIDES, sizes=ANCHOR_SIZES) -> List[torch.Tensor]:
    return [level_anchors(h, w, s, z) for (h, w), s, z in zip(level_hw, strides, sizes)]


def box2box_get_deltas(src: torch.Tensor, tgt: torch.Tensor, weights=(1.0, 1.0, 1.0, 1.0)) -> torch.Tensor:
    """[D2] Box2BoxTransform.get_deltas (RPN)."""
    sw = src[:, 2] - src[:, 0]
    sh = src[:, 3] - src[:, 1]
    sx = src[:, 0] + 0.5 * sw
    sy = src[:, 1] + 0.5 * sh
    tw = tgt[:, 2] - tgt[:, 0]
    th = tgt[:, 3] - tgt[:, 1]
    tx = tgt[:, 0] + 0.5 * tw
    ty = tgt[:, 1] + 0.5 * th
    wx, wy, ww, wh = weights
    return torch.stack((wx * (tx - sx) / sw, wy * (ty - sy) / sh, ww * torch.log(tw / sw), wh * torch.log(th / sh)), dim=1)


def box2box_apply_deltas(deltas: torch.Tensor, boxes: torch.Tensor, weights=(1.0, 1.0, 1.0, 1.0)) -> torch.Tensor:
    """[D2] Box2BoxTransform.apply_deltas (RPN decode)."""
    deltas = deltas.float()
    boxes = boxes.to(deltas.dtype)
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    wx, wy, ww, wh = weights
    dx, dy = deltas[:, 0] / wx, deltas[:, 1] / wy
    dw = torch.clamp(deltas[:, 2] / ww, max=SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3] / wh, max=SCALE_CLAMP)
    pcx = dx * w + cx
    pcy = dy * h + cy
    pw = torch.exp(dw) * w
    ph = torch.exp(dh) * h
    return torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=1)


def pairwise_iou(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """[D2] pairwise_iou (rpn.py:117, roi_heads.py:156,220): [M, N]; zero where the intersection is empty."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = (torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])).clamp(min=0)
    inter = wh.prod(dim=2)
    return torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros(1, dtype=inter.dtype))


def matcher(iou: torch.Tensor, thresholds: Sequence[float], labels: Sequence[int], allow_low_quality: bool):
    """[D2] Matcher.__call__ + set_low_quality_matches_. iou is [M gt, N candidates]."""
    n = iou.shape[1]
    if iou.numel() == 0:
        return torch.zeros(n, dtype=torch.int64), torch.full((n,), labels[0], dtype=torch.int8)
    vals, idx = iou.max(dim=0)
    lab = torch.full((n,), 1, dtype=torch.int8)
    th = [-float("inf")] + list(thresholds) + [float("inf")]
    for l, lo, hi in zip(labels, th[:-1], th[1:]):
        lab[(vals >= lo) & (vals < hi)] = l
    if allow_low_quality:
        best, _ = iou.max(dim=1)
        _, cols = torch.nonzero(iou == best[:, None], as_tuple=True)
        lab[cols] = 1
    return idx, lab


def _smallest_by_key(cand: torch.Tensor, keys: torch.Tensor, n: int) -> torch.Tensor:
    """The n entries of `cand` (index tensor) with the smallest (key, index), in that order."""
    if n <= 0 or cand.numel() == 0:
        return cand[:0]
    return cand[torch.argsort(keys[cand].to(torch.int64), stable=True)[:n]]      # cand is ascending: ties -> lower index


def subsample_labels(labels: torch.Tensor, num: int, pos_frac: float, bg_label: int, keys: torch.Tensor):
    """[D2] subsample_labels with the randperm draw replaced by injected keys (module docstring)."""
    pos = torch.nonzero((labels != -1) & (labels != bg_label), as_tuple=True)[0]
    neg = torch.nonzero(labels == bg_label, as_tuple=True)[0]
    n_pos = min(pos.numel(), int(num * pos_frac))
    n_neg = min(neg.numel(), num - n_pos)
    return _smallest_by_key(pos, keys, n_pos), _smallest_by_key(neg, keys, n_neg)


# ------------------------------------------------------------------------------------------ RPN (a21)
def rpn_label_and_sample(anchors: torch.Tensor, gt_boxes: List[torch.Tensor], gt_scores: Optional[List[torch.Tensor]],
                         keys: List[torch.Tensor], batch: int = 256, pos_frac: float = 0.25):
    """modeling/proposal_generator/rpn.py:78-150 (pseudo variant; with gt_scores=None it is [D2]
    RPN.label_and_sample_anchors). Returns per image: labels {-1,0,1} [A], matched gt index [A], matched boxes
    [A,4], confidences [A] (None without scores)."""
    out_l, out_i, out_b, out_c = [], [], [], []
    for i, gb in enumerate(gt_boxes):
        iou = pairwise_iou(gb, anchors)
        midx, lab = matcher(iou, (0.3, 0.7), (0, -1, 1), True)
        pos, neg = subsample_labels(lab, batch, pos_frac, 0, keys[i])
        lab = torch.full_like(lab, -1)
        lab[pos] = 1
        lab[neg] = 0
        if gb.shape[0] == 0:
            mb = torch.zeros_like(anchors)
            conf = torch.zeros(anchors.shape[0])
        else:
            mb = gb[midx]
            conf = gt_scores[i][midx] if gt_scores is not None else None
        out_l.append(lab)
        out_i.append(midx)
        out_b.append(mb)
        out_c.append(conf)
    return out_l, out_i, out_b, (out_c if gt_scores is not None else None)


def rpn_losses(anchors: torch.Tensor, logits: torch.Tensor, deltas: torch.Tensor, labels: List[torch.Tensor],
               matched_boxes: List[torch.Tensor], confids: Optional[List[torch.Tensor]], batch: int = 256) -> Dict:
    """rpn.py:153-225 (+ [D2] _dense_box_regression_loss, smooth-L1 beta 0 = L1). logits [N, A], deltas [N, A, 4]."""
    N = len(labels)
    lab = torch.stack(labels)
    pos = lab == 1
    tgt = torch.stack([box2box_get_deltas(anchors, mb) for mb in matched_boxes])
    loc = torch.abs(deltas[pos] - tgt[pos]).sum()
    valid = lab >= 0
    w = torch.stack(confids)[valid].float() if confids else None
    obj = F.binary_cross_entropy_with_logits(logits[valid], lab[valid].to(torch.float32), weight=w, reduction="sum")
    norm = batch * N
    return {"loss_rpn_cls": obj / norm, "loss_rpn_loc": loc / norm}


def rpn_select_decode(anchors: List[torch.Tensor], logits: List[torch.Tensor], deltas: List[torch.Tensor],
                      image_sizes, pre_topk: int = 2000):
    """First half of [D2] find_top_rpn_proposals: per level top-k by logit (ties: lower anchor index),
    decode, clip; returns per image (boxes [M,4], scores [M], level ids [M], valid mask [M])."""
    N = logits[0].shape[0]
    res = []
    for n in range(N):
        bs, ss, ls = [], [], []
        for l, (a, lg, dl) in enumerate(zip(anchors, logits, deltas)):
            k = min(lg.shape[1], pre_topk)
            s, idx = torch.sort(lg[n].float(), descending=True, stable=True)
            s, idx = s[:k], idx[:k]
            bs.append(box2box_apply_deltas(dl[n][idx].float(), a[idx]))
            ss.append(s)
            ls.append(torch.full((k,), l, dtype=torch.int64))
        b, s, lv = torch.cat(bs), torch.cat(ss), torch.cat(ls)
        valid = torch.isfinite(b).all(dim=1) & torch.isfinite(s)
        h, w = image_sizes[n]
        b = torch.stack((b[:, 0].clamp(0, w), b[:, 1].clamp(0, h), b[:, 2].clamp(0, w), b[:, 3].clamp(0, h)), dim=1)
        valid &= ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
        res.append((b, s, lv, valid))
    return res


def rpn_nms(cands, nms_thr: float = 0.7, post_topk: int = 1000, device_type: str = "cuda"):
    """Second half of [D2] find_top_rpn_proposals: batched_nms per level id, first post_topk in score order."""
    out = []
    for b, s, lv, valid in cands:
        b, s, lv = b[valid], s[valid], lv[valid]
        keep = batched_nms(b, s, lv, nms_thr, device_type)[:post_topk]
        out.append({"proposal_boxes": b[keep], "objectness_logits": s[keep]})
    return out


def rpn_find_top_proposals(anchors, logits, deltas, image_sizes, pre_topk=2000, post_topk=1000, nms_thr=0.7,
                           device_type="cuda"):
    return rpn_nms(rpn_select_decode(anchors, logits, deltas, image_sizes, pre_topk), nms_thr, post_topk, device_type)


# ------------------------------------------------------------------------------------------ ROI heads (a22)
def roi_label_and_sample(prop_boxes: torch.Tensor, gt_boxes: torch.Tensor, gt_classes: torch.Tensor,
                         gt_scores: Optional[torch.Tensor], gt_std: Optional[torch.Tensor], keys: torch.Tensor,
                         batch: int = 512, pos_frac: float = 0.25, num_classes: int = 80, append_gt: bool = True) -> Dict:
    """modeling/roi_heads/roi_heads.py:138-270 for one image ([D2] add_ground_truth_to_proposals, Matcher(.5),
    _sample_proposals). `keys` index the concatenated [proposals | gt] list."""
    boxes = torch.cat([prop_boxes, gt_boxes]) if append_gt else prop_boxes
    has_gt = gt_boxes.shape[0] > 0
    iou = pairwise_iou(gt_boxes, boxes)
    midx, mlab = matcher(iou, (0.5,), (0, 1), False)
    if has_gt:
        cls = gt_classes[midx].clone()
        cls[mlab == 0] = num_classes
        cls[mlab == -1] = -1
    else:
        cls = torch.full((boxes.shape[0],), num_classes, dtype=torch.int64)
    fg, bg = subsample_labels(cls, batch, pos_frac, num_classes, keys)
    sel = torch.cat([fg, bg])
    out = {"sampled_idxs": sel, "proposal_boxes": boxes[sel], "gt_classes": cls[sel]}
    if has_gt:
        t = midx[sel]
        out["gt_boxes"] = gt_boxes[t]
        if gt_scores is not None:
            out["gt_confid"] = gt_scores[t]
        if gt_std is not None:
            out["gt_loc_std"] = gt_std[t]
    else:
        out["gt_boxes"] = torch.zeros(sel.numel(), 4)
        if gt_scores is not None:
            out["gt_confid"] = torch.zeros(sel.numel())
        if gt_std is not None:
            out["gt_loc_std"] = torch.zeros(sel.numel(), 4)
    return out


def roi_levels(boxes: torch.Tensor, min_level=2, max_level=5, canonical_size=224, canonical_level=4) -> torch.Tensor:
    """[D2] assign_boxes_to_levels."""
    sizes = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    lv = torch.floor(canonical_level + torch.log2(sizes / canonical_size + 1e-8))
    return torch.clamp(lv, min=min_level, max=max_level).to(torch.int64) - min_level


def roi_align_ref(feat: torch.Tensor, rois: torch.Tensor, out: int, scale: float) -> torch.Tensor:
    """[tv] roi_align(aligned=True, sampling_ratio=0) restated in numpy-style loops (small cases; the tests also
    check this function against torchvision.ops.roi_align). feat [N,C,H,W]; rois [K,5] (batch, x1,y1,x2,y2)."""
    K = rois.shape[0]
    N, C, H, W = feat.shape
    res = torch.zeros(K, C, out, out, dtype=feat.dtype)
    for k in range(K):
        b = int(rois[k, 0])
        x1, y1, x2, y2 = [float(v) * scale - 0.5 for v in rois[k, 1:]]
        rw, rh = x2 - x1, y2 - y1
        bw, bh = rw / out, rh / out
        gh, gw = int(math.ceil(rh / out)), int(math.ceil(rw / out))
        cnt = max(gh * gw, 1)
        for ph in range(out):
            for pw in range(out):
                acc = torch.zeros(C, dtype=feat.dtype)
                for iy in range(gh):
                    y = y1 + ph * bh + (iy + 0.5) * bh / gh
                    for ix in range(gw):
                        x = x1 + pw * bw + (ix + 0.5) * bw / gw
                        if y < -1.0 or y > H or x < -1.0 or x > W:
                            continue
                        yy, xx = max(y, 0.0), max(x, 0.0)
                        yl, xl = int(yy), int(xx)
                        if yl >= H - 1:
                            yh = yl = H - 1
                            yy = float(yl)
                        else:
                            yh = yl + 1
                        if xl >= W - 1:
                            xh = xl = W - 1
                            xx = float(xl)
                        else:
                            xh = xl + 1
                        ly, lx = yy - yl, xx - xl
                        hy, hx = 1.0 - ly, 1.0 - lx
                        acc += hy * hx * feat[b, :, yl, xl] + hy * lx * feat[b, :, yl, xh] + \
                            ly * hx * feat[b, :, yh, xl] + ly * lx * feat[b, :, yh, xh]
                res[k, :, ph, pw] = acc / cnt
    return res


def roi_pool(feats: List[torch.Tensor], boxes: List[torch.Tensor], out: int = 7,
             scales=(1 / 4, 1 / 8, 1 / 16, 1 / 32)) -> torch.Tensor:
    """[D2] ROIPooler (ROIAlignV2, sampling_ratio 0) — roi_heads.py:118. Uses torchvision.ops.roi_align."""
    from torchvision.ops import roi_align
    allb = torch.cat(boxes)
    bidx = torch.cat([torch.full((len(b), 1), i, dtype=allb.dtype) for i, b in enumerate(boxes)])
    rois = torch.cat([bidx, allb], dim=1)
    lv = roi_levels(allb)
    res = torch.zeros(rois.shape[0], feats[0].shape[1], out, out, dtype=feats[0].dtype)
    for l, (f, s) in enumerate(zip(feats, scales)):
        inds = torch.nonzero(lv == l, as_tuple=True)[0]
        if inds.numel():
            res[inds] = roi_align(f, rois[inds], out, s, 0, True)
    return res


# ------------------------------------------------------------------------------------------ box predictor (a23)
def fast_rcnn_losses(scores, deltas, deltas_std, proposal_boxes, gt_boxes, gt_classes, branch, gt_loc_std=None,
                     num_classes=80, weights=(10.0, 10.0, 5.0, 5.0), ts_better=0.1, t_cert=0.5) -> Dict:
    """modeling/roi_heads/fast_rcnn.py:834-1084 under the shipped recipe: FocalLoss(gamma 1.5)/R; supervised
    box loss = [L1 + 0.05 * NLL * IoU(pred, gt)] / R ('nlloss', :941-1000); pseudo box loss = teacher-better L1 / R
    ('tsbetter', :1046-1069). Class-agnostic deltas."""
    R = gt_classes.numel()
    if R == 0:
        return {"loss_cls": 0.0 * scores.sum(), "loss_box_reg": 0.0 * deltas.sum()}
    loss_cls = rcnn_focal_loss(scores, gt_classes, 1.5) / R
    fg = torch.nonzero((gt_classes >= 0) & (gt_classes < num_classes), as_tuple=True)[0]
    d, s = deltas[fg], deltas_std[fg]
    tgt = box2box_xyxy_get_deltas(proposal_boxes[fg], gt_boxes[fg], weights)
    if branch == "unsup_data_train":
        ct = 1 - torch.sigmoid(gt_loc_std[fg])
        cs = 1 - torch.sigmoid(s)
        m = (ct > cs + ts_better) & (ct > t_cert)
        box = torch.abs(d[m] - tgt[m]).sum()
    else:
        pred = box2box_xyxy_apply_deltas(d, proposal_boxes[fg], weights)
        iou = matched_boxlist_iou(gt_boxes[fg], pred)
        box = torch.abs(d - tgt).sum() + 0.05 * rcnn_nl_loss(d, s, tgt, iou, "sum")
    return {"loss_cls": loss_cls, "loss_box_reg": box / max(R, 1.0)}


def fast_rcnn_candidates(scores, deltas, proposal_boxes, image_size, score_thresh=0.05,
                         weights=(10.0, 10.0, 5.0, 5.0)):
    """First half of [D2] fast_rcnn_inference_single_image: softmax, class-agnostic decode, clip,
    score > thresh. Returns boxes [n,4], scores [n], classes [n], roi row index [n] (nonzero order: row-major)."""
    probs = F.softmax(scores.float(), dim=-1)
    boxes = box2box_xyxy_apply_deltas(deltas.float(), proposal_boxes, weights)
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(probs).all(dim=1)
    probs = probs[:, :-1]
    h, w = image_size
    boxes = torch.stack((boxes[:, 0].clamp(0, w), boxes[:, 1].clamp(0, h), boxes[:, 2].clamp(0, w),
                         boxes[:, 3].clamp(0, h)), dim=1)
    mask = (probs > score_thresh) & valid[:, None]
    inds = mask.nonzero()
    return boxes[inds[:, 0]], probs[mask], inds[:, 1], inds[:, 0]


def fast_rcnn_inference(scores, deltas, deltas_std, proposal_boxes, image_size, score_thresh=0.05, nms_thresh=0.5,
                        topk=100, device_type="cuda") -> Dict:
    """fast_rcnn.py:1086-1125 (+ [D2] fast_rcnn_inference): detections + pred_boxes_std gathered by ROI row."""
    b, s, c, r = fast_rcnn_candidates(scores, deltas, proposal_boxes, image_size, score_thresh)
    keep = batched_nms(b, s, c, nms_thresh, device_type)[:topk]
    return {"pred_boxes": b[keep], "scores": s[keep], "pred_classes": c[keep], "pred_boxes_std": deltas_std[r[keep]].float(),
            "roi_index": r[keep]}


def threshold_bbox_rcnn(inst: Dict, thres: float = 0.7) -> Dict:
    """engine/trainer.py:727-751 (UBRCNNTeacherTrainer.threshold_bbox, proposal_type 'roih')."""
    valid = inst["scores"] > thres
    out = {"gt_boxes": inst["pred_boxes"][valid], "gt_classes": inst["pred_classes"][valid], "scores": inst["scores"][valid]}
    if "pred_boxes_std" in inst:
        out["pred_boxes_std"] = inst["pred_boxes_std"][valid]
    return out
