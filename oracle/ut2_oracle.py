"""TEST INFRASTRUCTURE ONLY — CPU oracle for the Unbiased-Teacher-v2 hot path.

A plain torch-CPU / numpy restatement of the reference's algorithms, each function citing the
reference file:line it follows (paths relative to /root/reference). Only tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() may import this module; the product
path (unbiased-teacher-v2_b200/) never does.

Pinning: the functions below are checked in tests/test_oracle_golden.py against fixtures minted by
oracle/gen_golden.py from the reference's own Python (tests/golden/*.pt). The Detectron2 / fvcore /
torchvision pieces (not on disk) are restated from their published behaviour and
cross-checked against torchvision.ops where it offers the same primitive — for those pieces parity
is "unpinned by the reference's own tests" (the reference has none).
"""
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

INF = 100000000  # ubteacher/modeling/fcos/fcos_outputs.py:19


# ------------------------------------------------------------------------------------------
# locations / targets
# ------------------------------------------------------------------------------------------
def compute_locations(h: int, w: int, stride: int) -> torch.Tensor:
    """ubteacher/utils/comm.py:34-45 — (x, y) = stride * idx + stride // 2, row-major over (h, w)."""
    ys = torch.arange(h, dtype=torch.float32) * stride
    xs = torch.arange(w, dtype=torch.float32) * stride
    x = xs.repeat(h)
    y = ys.repeat_interleave(w)
    return torch.stack([x, y], dim=1) + stride // 2


def sizes_of_interest(soi: Sequence[int]) -> List[List[float]]:
    """fcos_outputs.py:199-206."""
    out, prev = [], -1
    for s in soi:
        out.append([prev, s])
        prev = s
    out.append([prev, INF])
    return out


def center_sample_region(loc: torch.Tensor, b: torch.Tensor, strides: Sequence[int], num_loc: Sequence[int],
                         radius: float) -> torch.Tensor:
    """fcos_outputs.py:700-770 (get_sample_region, box centres; no bitmasks): [L, n] mask of the locations strictly inside a
    box's centre region — centre -/+ stride * radius of the location's level, clipped to the box. The reference's "no gt"
    test (:736) also fires when the FIRST box has centre x == 0 and then nothing is sampled."""
    cx = (b[:, 0] + b[:, 2]) * 0.5
    cy = (b[:, 1] + b[:, 3]) * 0.5
    if b.shape[0] == 0 or float(cx[0]) == 0.0:
        return torch.zeros(loc.shape[0], b.shape[0], dtype=torch.bool)
    sr = torch.cat([torch.full((n, 1), float(strides[i] * radius)) for i, n in enumerate(num_loc)])
    x, y = loc[:, 0:1], loc[:, 1:2]
    rx1 = torch.maximum(cx[None] - sr, b[None, :, 0])
    ry1 = torch.maximum(cy[None] - sr, b[None, :, 1])
    rx2 = torch.minimum(cx[None] + sr, b[None, :, 2])
    ry2 = torch.minimum(cy[None] + sr, b[None, :, 3])
    return torch.stack([x - rx1, y - ry1, rx2 - x, ry2 - y], 2).min(2)[0] > 0


def fcos_assign_targets(locations: List[torch.Tensor], boxes: List[torch.Tensor], classes: List[torch.Tensor],
                        strides: Sequence[int], soi: Sequence[int] = (64, 128, 256, 512), num_classes: int = 80,
                        scores: Optional[List[torch.Tensor]] = None, reg_pred_std: Optional[List[torch.Tensor]] = None,
                        soft_cls_label: bool = False, center_sample: bool = False, radius: float = 1.5,
                        ignore_near: bool = False) -> Dict[str, List[torch.Tensor]]:
    """fcos_outputs.py:649-698 (_get_ground_truth) + :772-906 (compute_targets_for_locations).
    Returns level-first lists (image-major inside a level). center_sample / radius: MODEL.FCOS.CENTER_SAMPLE / POS_RADIUS
    (:823-838); ignore_near: keep_locations per :841-848.
    """
    ranges = sizes_of_interest(soi)
    num_loc = [len(l) for l in locations]
    loc = torch.cat(locations, 0)
    lo = torch.cat([torch.full((n,), float(ranges[i][0])) for i, n in enumerate(num_loc)])
    hi = torch.cat([torch.full((n,), float(ranges[i][1])) for i, n in enumerate(num_loc)])
    L = loc.shape[0]
    per_im = {k: [] for k in ["labels", "box_weights", "reg_targets", "target_inds", "keep_locations", "boundary_vars"]}
    num_targets = 0
    for i in range(len(boxes)):
        b = boxes[i].float()
        c = classes[i].long()
        n = b.shape[0]
        use_score = scores is not None and soft_cls_label
        bw_src = scores[i] if use_score else torch.ones_like(c)
        bv_src = reg_pred_std[i] if reg_pred_std is not None else torch.zeros_like(b)
        if n == 0:  # :804-815
            per_im["labels"].append(torch.full((L,), num_classes, dtype=torch.long))
            per_im["box_weights"].append(torch.zeros(L, dtype=bw_src.dtype))
            per_im["reg_targets"].append(torch.zeros(L, 4))
            per_im["boundary_vars"].append(torch.zeros(L, 4))
            per_im["target_inds"].append(torch.full((L,), -1, dtype=torch.long))
            per_im["keep_locations"].append(torch.zeros(L, dtype=torch.bool))
            continue
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        labels = torch.empty(L, dtype=torch.long)
        inds = torch.empty(L, dtype=torch.long)
        regs = torch.empty(L, 4)
        isbg = torch.empty(L, dtype=torch.bool)
        keep = torch.ones(L, dtype=torch.bool)
        region = center_sample_region(loc, b, strides, num_loc, radius) if center_sample else None
        for p in range(L):  # explicit per-location scan (small cases only)
            x, y = loc[p, 0], loc[p, 1]
            best, best_area = 0, float(INF)
            any_inside = any_region = False
            for j in range(n):
                l_, t_, r_, b_ = x - b[j, 0], y - b[j, 1], b[j, 2] - x, b[j, 3] - y
                mn = min(l_, t_, r_, b_)
                mx = max(l_, t_, r_, b_)
                a = float(area[j])
                inside = bool(mn > 0)
                sampled = bool(region[p, j]) if center_sample else inside
                any_inside |= inside
                any_region |= sampled
                if not sampled:
                    a = float(INF)
                if not (mx >= lo[p] and mx <= hi[p]):
                    a = float(INF)
                if a < best_area:  # first minimum wins (torch.min returns the first index on CPU)
                    best, best_area = j, a
            if ignore_near:  # :841-848 background outside every box + everything inside a sample region
                keep[p] = (not any_inside) or any_region
            inds[p] = best
            isbg[p] = best_area == float(INF)
            regs[p] = torch.stack([x - b[best, 0], y - b[best, 1], b[best, 2] - x, b[best, 3] - y])
            labels[p] = c[best]
        labels[isbg] = num_classes
        bw = bw_src[inds].clone()
        bw[isbg] = 1.0
        bv = bv_src[inds].clone()
        bv[isbg] = 99999.0
        per_im["labels"].append(labels)
        per_im["box_weights"].append(bw)
        per_im["reg_targets"].append(regs)
        per_im["target_inds"].append(inds + num_targets)
        per_im["keep_locations"].append(keep)
        per_im["boundary_vars"].append(bv)
        num_targets += n
    per_im["locations"] = [loc.clone() for _ in boxes]
    per_im["im_inds"] = [torch.full((L,), i, dtype=torch.long) for i in range(len(boxes))]
    out = {}
    for k, v in per_im.items():  # :634-647 image-first -> level-first
        split = [torch.split(t, num_loc, dim=0) for t in v]
        out[k] = [torch.cat(per_level, 0) for per_level in zip(*split)]
    out["fpn_levels"] = [torch.full((len(l),), i, dtype=torch.long) for i, l in enumerate(out["locations"])]
    out["reg_targets"] = [t / float(strides[i]) for i, t in enumerate(out["reg_targets"])]  # :691-696
    return out


def fcos_assign_targets_fast(locations, boxes, classes, strides, soi=(64, 128, 256, 512), num_classes=80,
                             scores=None, reg_pred_std=None, soft_cls_label=False, center_sample=False, radius=1.5,
                             ignore_near=False):
    """Vectorised twin of fcos_assign_targets for full-size inputs (same tie rule: first minimum)."""
    ranges = sizes_of_interest(soi)
    num_loc = [len(l) for l in locations]
    loc = torch.cat(locations, 0)
    lo = torch.cat([torch.full((n, 1), float(ranges[i][0])) for i, n in enumerate(num_loc)])
    hi = torch.cat([torch.full((n, 1), float(ranges[i][1])) for i, n in enumerate(num_loc)])
    L = loc.shape[0]
    keys = ["labels", "box_weights", "reg_targets", "target_inds", "keep_locations", "boundary_vars"]
    per_im = {k: [] for k in keys}
    num_targets = 0
    for i in range(len(boxes)):
        b, c = boxes[i].float(), classes[i].long()
        n = b.shape[0]
        bw_src = scores[i] if (scores is not None and soft_cls_label) else torch.ones_like(c)
        bv_src = reg_pred_std[i] if reg_pred_std is not None else torch.zeros_like(b)
        if n == 0:
            per_im["labels"].append(torch.full((L,), num_classes, dtype=torch.long))
            per_im["box_weights"].append(torch.zeros(L, dtype=bw_src.dtype))
            per_im["reg_targets"].append(torch.zeros(L, 4))
            per_im["boundary_vars"].append(torch.zeros(L, 4))
            per_im["target_inds"].append(torch.full((L,), -1, dtype=torch.long))
            per_im["keep_locations"].append(torch.zeros(L, dtype=torch.bool))
            continue
        x, y = loc[:, 0:1], loc[:, 1:2]
        ltrb = torch.stack([x - b[None, :, 0], y - b[None, :, 1], b[None, :, 2] - x, b[None, :, 3] - y], 2)
        inside = ltrb.min(2)[0] > 0
        sampled = center_sample_region(loc, b, strides, num_loc, radius) if center_sample else inside
        keep = (~inside.any(1)) | sampled.any(1) if ignore_near else torch.ones(L, dtype=torch.bool)
        mx = ltrb.max(2)[0]
        cared = (mx >= lo) & (mx <= hi)
        area = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None].repeat(L, 1)
        area[~sampled] = INF
        area[~cared] = INF
        amin, inds = area.min(1)
        isbg = amin == INF
        labels = c[inds].clone()
        labels[isbg] = num_classes
        bw = bw_src[inds].clone()
        bw[isbg] = 1.0
        bv = bv_src[inds].clone()
        bv[isbg] = 99999.0
        per_im["labels"].append(labels)
        per_im["box_weights"].append(bw)
        per_im["reg_targets"].append(ltrb[torch.arange(L), inds])
        per_im["target_inds"].append(inds + num_targets)
        per_im["keep_locations"].append(keep)
        per_im["boundary_vars"].append(bv)
        num_targets += n
    per_im["locations"] = [loc.clone() for _ in boxes]
    per_im["im_inds"] = [torch.full((L,), i, dtype=torch.long) for i in range(len(boxes))]
    out = {}
    for k, v in per_im.items():
        split = [torch.split(t, num_loc, dim=0) for t in v]
        out[k] = [torch.cat(per_level, 0) for per_level in zip(*split)]
    out["fpn_levels"] = [torch.full((len(l),), i, dtype=torch.long) for i, l in enumerate(out["locations"])]
    out["reg_targets"] = [t / float(strides[i]) for i, t in enumerate(out["reg_targets"])]
    return out


# ------------------------------------------------------------------------------------------
# loss pieces
# ------------------------------------------------------------------------------------------
def integral(x: torch.Tensor, reg_max: int = 16) -> torch.Tensor:
    """fcos_outputs.py:44-77 — softmax over reg_max+1 bins, expectation with 0..reg_max."""
    p = torch.softmax(x.reshape(-1, reg_max + 1), dim=1)
    proj = torch.arange(reg_max + 1, dtype=x.dtype)
    return (p * proj).sum(1).reshape(-1, 4)


def ctrness_targets(t: torch.Tensor) -> torch.Tensor:
    """fcos_outputs.py:80-88."""
    if len(t) == 0:
        return t.new_zeros(0)
    lr = t[:, [0, 2]]
    tb = t[:, [1, 3]]
    return torch.sqrt((lr.min(1)[0] / lr.max(1)[0]) * (tb.min(1)[0] / tb.max(1)[0]))


def _ltrb_iou_terms(pred, tgt):
    ta = (tgt[:, 0] + tgt[:, 2]) * (tgt[:, 1] + tgt[:, 3])
    pa = (pred[:, 0] + pred[:, 2]) * (pred[:, 1] + pred[:, 3])
    wi = torch.min(pred[:, 0], tgt[:, 0]) + torch.min(pred[:, 2], tgt[:, 2])
    hi = torch.min(pred[:, 3], tgt[:, 3]) + torch.min(pred[:, 1], tgt[:, 1])
    gw = torch.max(pred[:, 0], tgt[:, 0]) + torch.max(pred[:, 2], tgt[:, 2])
    gh = torch.max(pred[:, 3], tgt[:, 3]) + torch.max(pred[:, 1], tgt[:, 1])
    inter = wi * hi
    union = ta + pa - inter
    return inter, union, gw * gh


def iou_targets(pred: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """fcos_outputs.py:91-129 — ltrb IoU with +1 smoothing."""
    if len(tgt) == 0:
        return tgt.new_zeros(0)
    inter, union, _ = _ltrb_iou_terms(pred, tgt)
    return (inter + 1.0) / (union + 1.0)


def iou_loss(pred, tgt, weight=None, kind="giou"):
    """ubteacher/layers/iou_loss.py:23-76."""
    inter, union, ac = _ltrb_iou_terms(pred, tgt)
    ious = (inter + 1.0) / (union + 1.0)
    gious = ious - (ac - union) / ac
    if kind == "iou":
        losses = -torch.log(ious)
    elif kind == "linear_iou":
        losses = 1 - ious
    elif kind == "giou":
        losses = 1 - gious
    else:
        raise NotImplementedError
    return (losses * weight).sum() if weight is not None else losses.sum()


def nl_loss_fcos(mean, std, tgt, iou_weight):
    """ubteacher/layers/kl_loss.py:75-105 — NLLoss: ignores weight/denorm/method; mean over boxes."""
    sigma_sq = torch.square(torch.sigmoid(std))
    s = (torch.square(tgt - mean) / (2 * sigma_sq) + 0.5 * torch.log(sigma_sq)).sum(1)
    s = s + 2 * torch.log(2 * torch.tensor([math.pi]))
    return (s * iou_weight).mean()


def kl_loss(inp, inp_std, tgt, weight=None, beta=1.0, loss_denorm=None, method="weight_ctr_sum"):
    """ubteacher/layers/kl_loss.py:17-66 (beta >= 1e-5 branch, the only one that returns)."""
    n = torch.abs(inp - tgt)
    l1s = torch.where(n < beta, 0.5 * n**2 / beta, n - 0.5 * beta)
    loss = torch.exp(-inp_std) * l1s + 0.5 * inp_std
    if method == "weight_ctr_sum":
        return (loss.sum(1) * weight).sum()
    if method == "weight_ctr_mean":
        return (loss.sum(1) * weight).sum() / loss_denorm
    if method == "sum":
        return loss.sum()
    if method == "mean":
        return loss.mean()
    raise ValueError


def sigmoid_focal_loss(x, t, alpha=0.25, gamma=2.0):
    """[fvcore] sigmoid_focal_loss_jit, reduction='none'."""
    p = torch.sigmoid(x)
    ce = F.binary_cross_entropy_with_logits(x, t, reduction="none")
    p_t = p * t + (1 - p) * (1 - t)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * t + (1 - alpha) * (1 - t)) * loss
    return loss


def _flatten_level_first(xs: List[torch.Tensor], C: int) -> torch.Tensor:
    """fcos_outputs.py:257-296 / :986-1030 — [N,C,H,W] per level -> [(level, image, h, w), C]."""
    return torch.cat([x.permute(0, 2, 3, 1).reshape(-1, C) for x in xs], 0)


def fcos_losses_labeled(logits, reg, std, ctr, locations, boxes, classes, strides=(8, 16, 32, 64, 128),
                        num_classes=80, kl_w=0.05, alpha=0.25, gamma=2.0, world_size=1,
                        allreduce=lambda t: t, assign=fcos_assign_targets_fast, kl_loss_type="nlloss", loc_fun_all="mean"):
    """fcos_outputs.py:212-305 (losses) + :307-444 (fcos_losses), sup1 recipe:
    focal cls, centerness quality, REG_DISCRETE, KL_LOSS nlloss, giou. Returns (losses, extras).
    kl_loss_type "klloss" (:380-397, config.py:198 default) swaps the NLL term for KLLoss with the LOC_FUN_ALL reduction."""
    tg = assign(locations, boxes, classes, strides, num_classes=num_classes)
    labels = torch.cat(tg["labels"])
    reg_t = torch.cat(tg["reg_targets"])
    keep = torch.cat(tg["keep_locations"])
    lg = _flatten_level_first(logits, num_classes)
    rg = _flatten_level_first(reg, 68)
    sd = _flatten_level_first(std, 4)
    ct = _flatten_level_first(ctr, 1).reshape(-1)
    if keep.sum() > 0:  # :310
        labels, reg_t, lg, rg, sd, ct, keep = labels[keep], reg_t[keep], lg[keep], rg[keep], sd[keep], ct[keep], keep[keep]
    pos = torch.nonzero(labels != num_classes).squeeze(1)
    num_pos_avg = max(float(allreduce(torch.tensor([float(pos.numel())]))) / world_size, 1.0)
    onehot = torch.zeros_like(lg)
    onehot[pos, labels[pos]] = 1
    cls_loss = sigmoid_focal_loss(lg, onehot, alpha, gamma).sum(1).sum() / num_pos_avg
    rt = reg_t[pos]
    pred = integral(rg[pos]) if pos.numel() > 0 else rg[pos]
    ctr_t = ctrness_targets(rt)
    denorm = max(float(allreduce(ctr_t.sum().reshape(1))) / world_size, 1e-6)
    if pos.numel() > 0:
        iou_t = iou_targets(pred.detach(), rt)
        ctr_loss = F.binary_cross_entropy_with_logits(ct[pos], ctr_t, reduction="sum") / num_pos_avg
        if kl_loss_type == "klloss":                                  # :380-389
            nll = kl_w * kl_loss(pred, sd[pos], rt, ctr_t, 1.0, denorm, loc_fun_all)
        else:
            nll = kl_w * nl_loss_fcos(pred, sd[pos], rt, iou_t)      # :400-408
        giou = iou_loss(pred, rt, ctr_t, "giou") / denorm            # :410-415
        reg_loss = kl_w * nll + giou                                 # :416 (weight applied twice)
    else:
        reg_loss = torch.tensor(0.0)
        ctr_loss = torch.tensor(0.0)
        denorm = 1.0
    if keep[pos].sum() == 0:  # :430-434 — evaluated on the positives-only view
        cls_loss, reg_loss, ctr_loss, denorm = cls_loss * 0, reg_loss * 0, ctr_loss * 0, 1.0
    return {"loss_fcos_cls": cls_loss, "loss_fcos_loc": reg_loss, "loss_fcos_ctr": ctr_loss}, {"loss_denorm": denorm}


def fcos_losses_pseudo(logits, reg, std, ctr, locations, cls_set, reg_set, strides=(8, 16, 32, 64, 128),
                       num_classes=80, alpha=0.25, gamma=2.0, ts_better=0.1, ts_cert=0.8, world_size=1,
                       allreduce=lambda t: t, assign=fcos_assign_targets_fast):
    """fcos_outputs.py:447-490 (pseudo_losses) + :492-608 (fcos_pseudo_losses) + :610-631 (class_loss),
    CONSIST_REG_LOSS = ts_locvar_better_nms_nll_l1. `cls_set` / `reg_set`: dicts with boxes, classes,
    scores, reg_pred_std lists."""
    lg = _flatten_level_first(logits, num_classes)
    rg = _flatten_level_first(reg, 68)
    sd = _flatten_level_first(std, 4)
    ct = _flatten_level_first(ctr, 1).reshape(-1)
    losses = {}
    for kind, S in (("cls", cls_set), ("reg", reg_set)):
        tg = assign(locations, S["boxes"], S["classes"], strides, num_classes=num_classes,
                    scores=S.get("scores"), reg_pred_std=S.get("reg_pred_std"))
        labels = torch.cat(tg["labels"])
        reg_t = torch.cat(tg["reg_targets"])
        bvar = torch.cat(tg["boundary_vars"])
        pos = torch.nonzero(labels != num_classes).squeeze(1)
        num_pos_avg = max(float(allreduce(torch.tensor([float(pos.numel())]))) / world_size, 1.0)
        if kind == "cls":
            onehot = torch.zeros_like(lg)
            onehot[pos, labels[pos]] = 1
            losses["loss_fcos_cls"] = sigmoid_focal_loss(lg, onehot, alpha, gamma).sum(1).sum() / num_pos_avg
        rt = reg_t[pos]
        ctr_t = ctrness_targets(rt)
        _ = max(float(allreduce(ctr_t.sum().reshape(1))) / world_size, 1e-6)
        if pos.numel() > 0:
            if kind == "cls":
                losses["loss_fcos_ctr"] = F.binary_cross_entropy_with_logits(ct[pos], ctr_t, reduction="sum") / num_pos_avg
            else:
                pred = integral(rg[pos])
                conf_s = 1 - torch.sigmoid(sd[pos])
                conf_t = 1 - torch.sigmoid(bvar[pos])
                select = (conf_t > ts_cert) & (conf_t > conf_s + ts_better)
                losses["teacher_better_student"] = select.sum()
                if select.sum() > 0:
                    losses["loss_fcos_loc"] = torch.abs(pred[select] - rt[select]).mean()  # smooth_l1(beta=0) :565
                else:
                    losses["loss_fcos_loc"] = torch.tensor(0.0)
        else:
            if kind == "cls":
                losses["loss_fcos_ctr"] = torch.tensor(0.0)
            else:
                losses["loss_fcos_loc"] = torch.tensor(0.0)
                losses["teacher_better_student"] = torch.tensor(0.0)
    return losses


# ------------------------------------------------------------------------------------------
# NMS / proposals
# ------------------------------------------------------------------------------------------
def nms_plain(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """[tv] nms: stable descending sort; suppress j when
    inter / (area_i + area_j - inter) > thr, fp32 arithmetic, widths clamped at 0, no +1."""
    boxes = boxes.astype(np.float32)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = ((x2 - x1) * (y2 - y1)).astype(np.float32)
    n = len(order)
    dead = np.zeros(n, dtype=bool)
    keep = []
    for a in range(n):
        i = order[a]
        if dead[a]:
            continue
        keep.append(i)
        rest = order[a + 1:]
        w = np.maximum(np.float32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
        h = np.maximum(np.float32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
        inter = (w * h).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / ((areas[i] + areas[rest]).astype(np.float32) - inter)
        dead[a + 1:] |= ovr > np.float32(thr)
    return np.asarray(keep, dtype=np.int64)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, thr: float,
                device_type: str = "cuda") -> torch.Tensor:
    """[D2] batched_nms -> [tv] batched_nms (layers/ml_nms.py:27): boxes.float(); coordinate trick
    (offset = idx * (max_coord + 1), fp32) unless numel() exceeds 4000 (cpu) / 20000 (cuda), in which
    case a per-class loop. `device_type` selects which of the reference's two behaviours to restate:
    the reference trains on cuda, the fixtures were minted on cpu."""
    boxes = boxes.float()
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    limit = 4000 if device_type == "cpu" else 20000
    if boxes.numel() > limit:
        keep_mask = torch.zeros_like(scores, dtype=torch.bool)
        for cid in torch.unique(idxs):
            ci = torch.where(idxs == cid)[0]
            k = nms_plain(boxes[ci].numpy(), scores[ci].numpy(), thr)
            keep_mask[ci[torch.from_numpy(k)]] = True
        ki = torch.where(keep_mask)[0]
        return ki[scores[ki].sort(descending=True, stable=True)[1]]
    max_c = boxes.max()
    offs = idxs.to(boxes) * (max_c + torch.tensor(1).to(boxes))
    b = boxes + offs[:, None]
    return torch.from_numpy(nms_plain(b.numpy(), scores.numpy(), thr))


def fcos_predict_proposals(logits, reg, std, ctr, locations, image_sizes, method="cls_n_ctr",
                           strides=(8, 16, 32, 64, 128), thr=0.05, pre_topk=1000, post_topk=100, nms_thr=0.6,
                           device_type="cuda"):
    """fcos_outputs.py:1046-1132 (predict_proposals) + :1134-1298 (forward_for_single_feature_map)
    + :1300-1320 (select_over_all_levels). Returns per image a dict of tensors."""
    if method not in ("cls", "cls_n_ctr", "cls_n_loc"):
        raise ValueError("Undefined nms criteria")  # "ctr" raises inside the reference (:1193/:1202)
    N = logits[0].shape[0]
    per_image = [[] for _ in range(N)]
    for lvl, (lg, rg, sd, ct, loc, s) in enumerate(zip(logits, reg, std, ctr, locations, strides)):
        C = lg.shape[1]
        prob = lg.permute(0, 2, 3, 1).reshape(N, -1, C).sigmoid()
        rg_cls = rg.permute(0, 2, 3, 1).reshape(N, -1, 68)
        dist = integral(rg_cls.reshape(-1, 68)).reshape(N, -1, 4) * s
        sdv = sd.permute(0, 2, 3, 1).reshape(N, -1, 4)
        ctv = ct.permute(0, 2, 3, 1).reshape(N, -1).sigmoid()
        cand = prob > thr
        if method == "cls_n_ctr":
            rank = prob * ctv[:, :, None]
        elif method == "cls":
            rank = prob
        else:
            rank = prob * (1 - sdv.sigmoid()).mean(2)[:, :, None]
        for i in range(N):
            nz = cand[i].nonzero()
            pl, pc = nz[:, 0], nz[:, 1]
            sc = rank[i][cand[i]]
            conf = prob[i][cand[i]]
            k = min(int(cand[i].sum()), pre_topk)
            if int(cand[i].sum()) > k:
                sc, top = sc.topk(k, sorted=False)
                pl, pc, conf = pl[top], pc[top], conf[top]
            d = dist[i][pl]
            l = loc[pl]
            det = torch.stack([l[:, 0] - d[:, 0], l[:, 1] - d[:, 1], l[:, 0] + d[:, 2], l[:, 1] + d[:, 3]], 1)
            per_image[i].append({
                "pred_boxes": det, "scores": torch.sqrt(sc) if method in ("cls_n_ctr", "cls_n_loc") else sc,
                "reg_pred_cls": rg_cls[i][pl], "reg_pred_cls_scalar": torch.ones(len(pl)) * s,
                "reg_pred_std": sdv[i][pl], "pred_classes": pc, "locations": l, "centerness": ctv[i][pl],
                "cls_confid": conf, "fpn_levels": torch.full((len(pl),), lvl, dtype=torch.long)})
    results = []
    for i in range(N):
        allv = {k: torch.cat([d[k] for d in per_image[i]], 0) for k in per_image[i][0]}
        keep = batched_nms(allv["pred_boxes"], allv["scores"], allv["pred_classes"], nms_thr, device_type)
        res = {k: v[keep] for k, v in allv.items()}
        n = len(keep)
        if n > post_topk > 0:
            kth, _ = torch.kthvalue(res["scores"], n - post_topk + 1)
            sel = torch.nonzero(res["scores"] >= kth.item()).squeeze(1)
            res = {k: v[sel] for k, v in res.items()}
        res["image_size"] = image_sizes[i]
        results.append(res)
    return results


def threshold_bbox(inst: Dict[str, torch.Tensor], thres: float) -> Dict[str, torch.Tensor]:
    """ubteacher/modeling/pseudo_generator.py:62-105 (proposal_type == "roih")."""
    valid = inst["scores"] > thres
    out = {"gt_boxes": inst["pred_boxes"][valid], "gt_classes": inst["pred_classes"][valid],
           "scores": inst["scores"][valid], "centerness": inst["centerness"][valid],
           "cls_confid": inst["cls_confid"][valid]}
    if "reg_pred_std" in inst:
        out["reg_pred_std"] = inst["reg_pred_std"][valid]
    return out


def threshold_cls_ctr_bbox(inst, thres):
    """pseudo_generator.py:107-131."""
    valid = (inst["cls_confid"] > thres[0]) & (inst["centerness"] > thres[1])
    out = {"gt_boxes": inst["pred_boxes"][valid], "gt_classes": inst["pred_classes"][valid],
           "scores": inst["scores"][valid], "centerness": inst["centerness"][valid],
           "cls_confid": inst["cls_confid"][valid]}
    if "reg_pred_std" in inst:
        out["reg_pred_std"] = inst["reg_pred_std"][valid]
    return out


# ------------------------------------------------------------------------------------------
# EMA / SGD
# ------------------------------------------------------------------------------------------
def ema_update(student: torch.Tensor, teacher: torch.Tensor, keep: float) -> torch.Tensor:
    """ubteacher/engine/trainer.py:468-486 — new = student*(1-keep) + teacher*keep in fp32 with both
    products rounded separately (python scalars are rounded to fp32 by torch)."""
    a = np.float32(1 - keep)
    b = np.float32(keep)
    s = student.numpy().astype(np.float32)
    t = teacher.numpy().astype(np.float32)
    return torch.from_numpy((s * a).astype(np.float32) + (t * b).astype(np.float32))


def sgd_step(p, g, buf, lr, momentum=0.9, wd=1e-4, first=False):
    """torch.optim.SGD (dampening 0, no nesterov) as built by [D2] build_optimizer (trainer.py:422-429):
    g += wd*p ; buf = g (first step) or momentum*buf + g ; p -= lr*buf."""
    g = g + wd * p
    buf = g.clone() if first else momentum * buf + g
    return p - lr * buf, buf


# ------------------------------------------------------------------------------------------
# R-CNN pieces (a23 / a24)
# ------------------------------------------------------------------------------------------
def box2box_xyxy_get_deltas(src, tgt, weights=(10.0, 10.0, 5.0, 5.0)):
    """ubteacher/modeling/box_regression.py:36-73: divides by size+1; wy = weights[1]."""
    sw = src[:, 2] - src[:, 0] + 1.0
    sh = src[:, 3] - src[:, 1] + 1.0
    wx, wy = weights[0], weights[1]
    dl = wx * (tgt[:, 0] - src[:, 0]) / sw
    dr = wx * (tgt[:, 2] - src[:, 2]) / sw
    dd = wy * (tgt[:, 1] - src[:, 1]) / sh
    du = wy * (tgt[:, 3] - src[:, 3]) / sh
    return torch.stack((dl, dr, dd, du), dim=1)


def box2box_xyxy_apply_deltas(deltas, boxes, weights=(10.0, 10.0, 5.0, 5.0), clamp=1000.0 / 16):
    """box_regression.py:75-129: multiplies by size (no +1); scale_clamp is 1000/16 = 62.5 (:5, no log) —
    see reference: dl,dr,dd,du are divided by weights then clamped to [-clamp, clamp]."""
    deltas = deltas.float()
    boxes = boxes.to(deltas.dtype)
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    wx, wy = weights[0], weights[1]
    dl = deltas[:, 0::4] / wx
    dr = deltas[:, 1::4] / wx
    dd = deltas[:, 2::4] / wy
    du = deltas[:, 3::4] / wy
    dl, dr, dd, du = [torch.clamp(v, max=clamp, min=-clamp) for v in (dl, dr, dd, du)]
    out = torch.zeros_like(deltas)
    out[:, 0::4] = dl * w[:, None] + boxes[:, 0:1]
    out[:, 1::4] = dd * h[:, None] + boxes[:, 1:2]
    out[:, 2::4] = dr * w[:, None] + boxes[:, 2:3]
    out[:, 3::4] = du * h[:, None] + boxes[:, 3:4]
    return out


def rcnn_focal_loss(x, y, gamma=1.5, confid=None):
    """ubteacher/modeling/roi_heads/fast_rcnn.py:1405-1429 — sum over rows of (1-p)^gamma * CE."""
    ce = F.cross_entropy(x, y, reduction="none")
    loss = (1 - torch.exp(-ce)) ** gamma * ce
    if confid is not None:
        loss = loss * confid
    return loss.sum()


def rcnn_nl_loss(mean, std, tgt, iou_weight, reduction="sum"):
    """fast_rcnn.py:1228-1292."""
    sigma_sq = torch.square(torch.sigmoid(std))
    s = (torch.square(tgt - mean) / (2 * sigma_sq) + 0.5 * torch.log(sigma_sq)).sum(1)
    s = (s + 2 * torch.log(2 * torch.tensor([math.pi]))) * iou_weight
    return s.mean() if reduction == "mean" else s.sum()


def matched_boxlist_iou(b1, b2):
    """fast_rcnn.py:20-44 ([D2]-style matched IoU, no +1)."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, :2], b2[:, :2])
    rb = torch.min(b1[:, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    return inter / (a1 + a2 - inter)
