"""Thin Python wrappers over the Faster R-CNN C-ABI kernels (include/ut2.h, csrc/{nms,rpn,roi}.cu). Tensors are torch
CUDA tensors used as device-memory handles; all arithmetic runs in the sm_100a kernels."""
import ctypes
import math

import torch

from . import _C
from ._C import f32, i64

BF16 = torch.bfloat16
SCALE_CLAMP = math.log(1000.0 / 16)       # [D2] Box2BoxTransform (RPN)
XYXY_CLAMP = 1000.0 / 16                  # ubteacher/modeling/box_regression.py:5 (no log)
PRED_LD = 96                              # fused box predictor: 81 scores | 4 deltas | 4 std | pad
RPN_LD = 16                               # fused RPN predictor: 3 objectness | 12 deltas | pad


class RpnGeom:
    """Host-side launch parameters of the RPN kernels: level sizes, strides and the [D2] cell anchors."""

    def __init__(self, hw, strides, sizes=(32, 64, 128, 256, 512), ratios=(0.5, 1.0, 2.0)):
        self.hw = [tuple(x) for x in hw]
        self.strides = list(strides)
        self.num = len(self.hw)
        self.L = sum(h * w for h, w in self.hw)
        self.A = 3 * self.L
        assert len(ratios) == 3
        cell = []
        for s in sizes[: self.num]:
            area = float(s) ** 2.0
            for r in ratios:
                w = math.sqrt(area / r)
                h = r * w
                cell += [-w / 2.0, -h / 2.0, w / 2.0, h / 2.0]
        self.c_hw = (ctypes.c_int * (2 * self.num))(*[v for x in self.hw for v in x])
        self.c_strides = (ctypes.c_int * self.num)(*self.strides)
        self.c_cell = (ctypes.c_float * len(cell))(*cell)
        off = [0]
        for h, w in self.hw:
            off.append(off[-1] + h * w)
        self.off = off

    def cand_slots(self, pre_topk):
        return sum(min(3 * h * w, pre_topk) for h, w in self.hw)


def _lib_ll(name):
    fn = getattr(_C.lib(), name)
    fn.restype = ctypes.c_longlong
    return fn


def nms_batched(boxes, scores, cls, cnt, thr, max_keep, tie=None, trick_limit=20000):
    """boxes [N,M,4] f32, scores [N,M] f32, cls [N,M] i32, cnt [N] i32 -> keep_idx [N,max_keep] i32, keep_cnt [N]."""
    N, M = scores.shape
    dev = scores.device
    wsb = _lib_ll("ut2_nms_workspace_bytes")(N, M)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    keep = torch.empty((N, max_keep), dtype=torch.int32, device=dev)
    kcnt = torch.empty(N, dtype=torch.int32, device=dev)
    _C.counted_call("ut2_nms_batched", N, M, boxes, scores, tie, cls, cnt, f32(thr), trick_limit, max_keep, ws, i64(wsb), keep, kcnt)
    _C.launch_count += 2
    return keep, kcnt


def nms_segmented(boxes, scores, tie, seg_off, thr, max_keep, trick_limit=20000):
    """Per-segment NMS + merge (csrc/nms.cu, ut2_nms_segmented): boxes [N,M,4], scores [N,M], tie [N,M] i32 | None, seg_off:
    list of S+1 slot offsets -> keep_idx [N,max_keep] i32 (slots in the image's list), keep_cnt [N]."""
    N, M = scores.shape
    dev = scores.device
    S = len(seg_off) - 1
    mseg = max(max(seg_off[i + 1] - seg_off[i] for i in range(S)), 1)
    fn = _lib_ll("ut2_nms_segmented_workspace_bytes")
    wsb = fn(N, S, mseg, max_keep)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    keep = torch.empty((N, max_keep), dtype=torch.int32, device=dev)
    kcnt = torch.empty(N, dtype=torch.int32, device=dev)
    c_off = (ctypes.c_int * (S + 1))(*seg_off)
    _C.counted_call("ut2_nms_segmented", N, M, S, c_off, boxes, scores, tie, f32(thr), trick_limit, max_keep, ws, i64(wsb), keep, kcnt)
    _C.launch_count += 4
    return keep, kcnt


def gather_rows(src, idx, cnt, width):
    """src [N,M,width] (4- or 8-byte elements) -> [N,K,width] gathered by idx [N,K] (zero beyond cnt)."""
    N, K = idx.shape
    M = src.shape[1]
    out = torch.empty((N, K) + tuple(src.shape[2:]), dtype=src.dtype, device=src.device)
    _C.counted_call("ut2_gather_rows", N, M, K, width, src.element_size(), src, idx, cnt, out)
    return out


def rpn_label_anchors(geom, N, gt_boxes, gt_cnt, keys=None, seed=0, batch=256, pos_frac=0.25, lo=0.3, hi=0.7, seed_dev=None):
    dev = gt_boxes.device
    G = gt_boxes.shape[1]
    wsb = _lib_ll("ut2_rpn_label_workspace_bytes")(N, ctypes.c_longlong(geom.A), G)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    labels = torch.empty((N, geom.A), dtype=torch.int8, device=dev)
    matched = torch.empty((N, geom.A), dtype=torch.int32, device=dev)
    _C.counted_call("ut2_rpn_label_anchors", geom.num, geom.c_hw, geom.c_strides, geom.c_cell, N, G, gt_boxes, gt_cnt, keys,
                    ctypes.c_uint(seed & 0xFFFFFFFF), seed_dev, batch, f32(pos_frac), f32(lo), f32(hi), ws, i64(wsb), labels, matched)
    _C.launch_count += 2
    return labels, matched


def rpn_loss_fwd(geom, N, rpn_out, labels, matched, gt_boxes, gt_scores, gt_cnt, batch=256):
    dev = rpn_out.device
    acc = torch.empty(2, dtype=torch.float64, device=dev)
    losses = torch.empty(2, dtype=torch.float32, device=dev)
    _C.counted_call("ut2_rpn_loss_fwd", geom.num, geom.c_hw, geom.c_strides, geom.c_cell, N, gt_boxes.shape[1], rpn_out, labels,
                    matched, gt_boxes, gt_scores, gt_cnt, batch, acc, losses)
    _C.launch_count += 1
    return losses


def rpn_loss_bwd(geom, N, rpn_out, labels, matched, gt_boxes, gt_scores, gt_cnt, gout, batch=256):
    drpn = torch.empty_like(rpn_out)
    _C.counted_call("ut2_rpn_loss_bwd", geom.num, geom.c_hw, geom.c_strides, geom.c_cell, N, gt_boxes.shape[1], rpn_out, labels,
                    matched, gt_boxes, gt_scores, gt_cnt, batch, gout, drpn)
    return drpn


def rpn_select_decode(geom, N, rpn_out, image_hw, pre_topk):
    dev = rpn_out.device
    M = geom.cand_slots(pre_topk)
    out = {"boxes": torch.empty((N, M, 4), dtype=torch.float32, device=dev),
           "scores": torch.empty((N, M), dtype=torch.float32, device=dev),
           "canon": torch.empty((N, M), dtype=torch.int32, device=dev),
           "levels": torch.empty((N, M), dtype=torch.int32, device=dev),
           "count": torch.empty(N, dtype=torch.int32, device=dev)}
    _C.counted_call("ut2_rpn_select_decode", geom.num, geom.c_hw, geom.c_strides, geom.c_cell, N, rpn_out, image_hw, pre_topk,
                    f32(SCALE_CLAMP), M, out["boxes"], out["scores"], out["canon"], out["levels"], out["count"])
    _C.launch_count += 1
    return out


def rpn_predict_proposals(geom, N, rpn_out, image_hw, pre_topk=2000, post_topk=1000, nms_thr=0.7, segmented=True):
    """[D2] find_top_rpn_proposals: -> proposal_boxes [N,post_topk,4], objectness_logits [N,post_topk], count [N]."""
    c = rpn_select_decode(geom, N, rpn_out, image_hw, pre_topk)
    seg_off = [0]
    for h, w in geom.hw:                    # the candidate list is laid out level by level (rpn_select_decode)
        seg_off.append(seg_off[-1] + min(3 * h * w, pre_topk))
    if segmented and pre_topk <= 2048 and len(geom.hw) <= 8:
        keep, kcnt = nms_segmented(c["boxes"], c["scores"], c["canon"], seg_off, nms_thr, post_topk)
    else:
        keep, kcnt = nms_batched(c["boxes"], c["scores"], c["levels"], c["count"], nms_thr, post_topk, tie=c["canon"])
    return {"proposal_boxes": gather_rows(c["boxes"], keep, kcnt, 4),
            "objectness_logits": gather_rows(c["scores"].unsqueeze(-1), keep, kcnt, 1).squeeze(-1), "count": kcnt}


def roi_sample(prop_boxes, prop_cnt, gt_boxes, gt_classes, gt_cnt, gt_scores=None, gt_std=None, keys=None, seed=0,
               batch=512, pos_frac=0.25, iou_thr=0.5, num_classes=80, append_gt=True, seed_dev=None):
    N, Pcap = prop_boxes.shape[:2]
    G = gt_boxes.shape[1]
    dev = prop_boxes.device
    out = {"proposal_boxes": torch.empty((N, batch, 4), dtype=torch.float32, device=dev),
           "gt_classes": torch.empty((N, batch), dtype=torch.int64, device=dev),
           "gt_boxes": torch.empty((N, batch, 4), dtype=torch.float32, device=dev),
           "gt_confid": torch.empty((N, batch), dtype=torch.float32, device=dev),
           "gt_loc_std": torch.empty((N, batch, 4), dtype=torch.float32, device=dev),
           "sampled_idxs": torch.empty((N, batch), dtype=torch.int32, device=dev),
           "count": torch.empty(N, dtype=torch.int32, device=dev)}
    _C.counted_call("ut2_roi_sample", N, Pcap, G, batch, prop_boxes, prop_cnt, gt_boxes, gt_classes, gt_cnt, gt_scores, gt_std,
                    keys, keys.shape[1] if keys is not None else 0, ctypes.c_uint(seed & 0xFFFFFFFF), seed_dev, f32(pos_frac),
                    f32(iou_thr), num_classes, int(append_gt), out["proposal_boxes"], out["gt_classes"], out["gt_boxes"], out["gt_confid"],
                    out["gt_loc_std"], out["sampled_idxs"], out["count"])
    return out


class RoiGeom:
    def __init__(self, hw, scales):
        self.num = len(hw)
        self.hw = [tuple(x) for x in hw]
        self.c_hw = (ctypes.c_int * (2 * self.num))(*[v for x in hw for v in x])
        self.c_scales = (ctypes.c_float * self.num)(*scales)

    @staticmethod
    def ptrs(tensors):
        return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def roi_align_fwd(rgeom, feats, rois, roi_cnt):
    """feats: list of [N,H_l,W_l,C] bf16; rois [N,Rcap,4] f32 -> [N*Rcap, 7, 7, C] bf16."""
    N, Rcap = rois.shape[:2]
    C = feats[0].shape[-1]
    out = torch.empty((N * Rcap, 7, 7, C), dtype=BF16, device=rois.device)
    _C.counted_call("ut2_roi_align_fwd", rgeom.num, RoiGeom.ptrs(feats), rgeom.c_hw, rgeom.c_scales, N, C, Rcap, rois, roi_cnt, out)
    return out


def roi_align_bwd(rgeom, dfeats, rois, roi_cnt, dout):
    """dfeats: list of zero-initialised fp32 [N,H_l,W_l,C] accumulators (atomically added to)."""
    N, Rcap = rois.shape[:2]
    C = dfeats[0].shape[-1]
    _C.counted_call("ut2_roi_align_bwd", rgeom.num, RoiGeom.ptrs(dfeats), rgeom.c_hw, rgeom.c_scales, N, C, Rcap, rois, roi_cnt, dout)


def fastrcnn_loss_fwd(pred, s, mode, weights=(10.0, 10.0), gamma=1.5, nll_w=0.05, ts_better=0.1, t_cert=0.5):
    """pred [N*Rcap, 96] bf16; s: dict from roi_sample. Returns float[2] = {loss_cls, loss_box_reg}."""
    N, Rcap = s["gt_classes"].shape
    dev = pred.device
    acc = torch.empty(2, dtype=torch.float64, device=dev)
    losses = torch.empty(2, dtype=torch.float32, device=dev)
    _C.counted_call("ut2_fastrcnn_loss_fwd", N, Rcap, pred, s["proposal_boxes"], s["gt_classes"], s["gt_boxes"], s["gt_loc_std"],
                    s["count"], mode, f32(weights[0]), f32(weights[1]), f32(XYXY_CLAMP), f32(gamma), f32(nll_w), f32(ts_better),
                    f32(t_cert), acc, losses)
    _C.launch_count += 1
    return losses


def fastrcnn_loss_bwd(pred, s, mode, gout, weights=(10.0, 10.0), gamma=1.5, nll_w=0.05, ts_better=0.1, t_cert=0.5):
    N, Rcap = s["gt_classes"].shape
    dpred = torch.empty_like(pred)
    _C.counted_call("ut2_fastrcnn_loss_bwd", N, Rcap, pred, s["proposal_boxes"], s["gt_classes"], s["gt_boxes"], s["gt_loc_std"],
                    s["count"], mode, f32(weights[0]), f32(weights[1]), f32(XYXY_CLAMP), f32(gamma), f32(nll_w), f32(ts_better),
                    f32(t_cert), gout, dpred)
    return dpred


def fastrcnn_inference(pred, rois, roi_cnt, image_hw, score_thr=0.05, nms_thr=0.5, topk=100, weights=(10.0, 10.0),
                       cand_cap=8192):
    """fast_rcnn.py:1086-1125: -> dict(pred_boxes [N,topk,4], scores, pred_classes i64, reg_pred_std [N,topk,4] (the
    reference's pred_boxes_std), roi_index, count). The fields the shared threshold-scatter kernel also carries
    (centerness, cls_confid) are zero-filled."""
    N, Rcap = rois.shape[:2]
    dev = pred.device
    cb = torch.empty((N, cand_cap, 4), dtype=torch.float32, device=dev)
    cs = torch.empty((N, cand_cap), dtype=torch.float32, device=dev)
    cc = torch.empty((N, cand_cap), dtype=torch.int32, device=dev)
    cn = torch.empty((N, cand_cap), dtype=torch.int32, device=dev)
    cnt = torch.empty(N, dtype=torch.int32, device=dev)
    ovf = torch.empty(1, dtype=torch.int32, device=dev)
    _C.counted_call("ut2_fastrcnn_candidates", N, Rcap, pred, rois, roi_cnt, image_hw, f32(weights[0]), f32(weights[1]),
                    f32(XYXY_CLAMP), f32(score_thr), cand_cap, cb, cs, cc, cn, cnt, ovf)
    _C.launch_count += 1
    keep, kcnt = nms_batched(cb, cs, cc, cnt, nms_thr, topk, tie=cn)
    K = topk
    out = {"pred_boxes": torch.empty((N, K, 4), dtype=torch.float32, device=dev),
           "scores": torch.empty((N, K), dtype=torch.float32, device=dev),
           "pred_classes": torch.empty((N, K), dtype=torch.int64, device=dev),
           "reg_pred_std": torch.empty((N, K, 4), dtype=torch.float32, device=dev),
           "roi_index": torch.empty((N, K), dtype=torch.int32, device=dev),
           "count": torch.empty(N, dtype=torch.int32, device=dev), "overflow": ovf,
           "candidates": {"boxes": cb, "scores": cs, "classes": cc, "canon": cn, "count": cnt}}
    _C.counted_call("ut2_fastrcnn_gather", N, cand_cap, K, Rcap, keep, kcnt, cb, cs, cn, pred, out["pred_boxes"], out["scores"],
                    out["pred_classes"], out["reg_pred_std"], out["roi_index"], out["count"])
    out["centerness"] = torch.zeros((N, K), dtype=torch.float32, device=dev)
    out["cls_confid"] = torch.zeros((N, K), dtype=torch.float32, device=dev)
    return out


def add_f32_bf16(a_f32, b_bf16=None):
    out = torch.empty(a_f32.shape, dtype=BF16, device=a_f32.device)
    _C.counted_call("ut2_add_f32_bf16", a_f32, b_bf16, out, i64(a_f32.numel()))
    return out


def subsample2x(x, out=None):
    N, H, W, C = x.shape
    y = out if out is not None else torch.empty((N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=BF16, device=x.device)
    _C.counted_call("ut2_subsample2x_nhwc", x, y, N, H, W, C)
    return y
