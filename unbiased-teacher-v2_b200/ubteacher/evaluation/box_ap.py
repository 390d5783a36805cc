"""COCO-style box AP on the host (the metric [D2] ``COCOEvaluator`` / pycocotools ``COCOeval`` report as bbox AP, AP50,
AP75; pycocotools is not installed here, so this is a restatement of the published protocol and is unpinned): per category
and IoU threshold 0.50:0.05:0.95, detections (<= 100 per image) are visited in descending score order and greedily matched
to the unmatched ground-truth box of highest IoU >= threshold; precision is made monotone and sampled at 101 recall points.
No crowd regions and no area ranges (UT2 reports bbox AP only)."""
import numpy as np
import torch

IOU_THRS = np.linspace(0.5, 0.95, 10)
REC_THRS = np.linspace(0.0, 1.0, 101)


def _iou(d, g):
    """COCO box IoU (continuous coordinates, no +1): d [n, 4], g [m, 4] xyxy -> [n, m]."""
    if len(d) == 0 or len(g) == 0:
        return np.zeros((len(d), len(g)))
    lt = np.maximum(d[:, None, :2], g[None, :, :2])
    rb = np.minimum(d[:, None, 2:], g[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    ad = (d[:, 2] - d[:, 0]) * (d[:, 3] - d[:, 1])
    ag = (g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1])
    return inter / np.maximum(ad[:, None] + ag[None, :] - inter, 1e-12)


def coco_box_ap(dets, gts, max_dets=100):
    """dets / gts: {image_id: (boxes [n, 4], classes [n][, scores [n]])}. Returns {"AP", "AP50", "AP75"} in percent."""
    cats = sorted({int(c) for _, (_, cls) in gts.items() for c in np.asarray(cls).tolist()})
    ap = np.full((len(IOU_THRS), len(cats)), -1.0)
    for ci, cat in enumerate(cats):
        scores, matched, n_gt = [], [[] for _ in IOU_THRS], 0
        for img, (gb, gc) in gts.items():
            gb = np.asarray(gb, dtype=np.float64).reshape(-1, 4)[np.asarray(gc) == cat]
            n_gt += len(gb)
            if img not in dets:
                continue
            db, dc, ds = (np.asarray(x) for x in dets[img])
            sel = dc == cat
            db, ds = db.reshape(-1, 4)[sel].astype(np.float64), ds[sel].astype(np.float64)
            order = np.argsort(-ds, kind="mergesort")[:max_dets]
            db, ds = db[order], ds[order]
            iou = _iou(db, gb)
            scores.append(ds)
            for ti, thr in enumerate(IOU_THRS):
                taken = np.zeros(len(gb), dtype=bool)
                tp = np.zeros(len(db), dtype=bool)
                for di in range(len(db)):
                    best, bj = min(thr, 1 - 1e-10), -1
                    for gj in range(len(gb)):
                        if taken[gj] or iou[di, gj] < best:
                            continue
                        best, bj = iou[di, gj], gj
                    if bj >= 0:
                        taken[bj] = True
                        tp[di] = True
                matched[ti].append(tp)
        if n_gt == 0:
            continue
        s = np.concatenate(scores) if scores else np.zeros(0)
        order = np.argsort(-s, kind="mergesort")
        for ti in range(len(IOU_THRS)):
            tp = (np.concatenate(matched[ti]) if matched[ti] else np.zeros(0, dtype=bool))[order]
            ctp, cfp = np.cumsum(tp), np.cumsum(~tp)
            rc = ctp / n_gt
            pr = ctp / np.maximum(ctp + cfp, np.spacing(1))
            for i in range(len(pr) - 1, 0, -1):          # precision envelope
                pr[i - 1] = max(pr[i - 1], pr[i])
            q = np.zeros(len(REC_THRS))
            idx = np.searchsorted(rc, REC_THRS, side="left")
            ok = idx < len(pr)
            q[ok] = pr[idx[ok]]
            ap[ti, ci] = q.mean()
    valid = ap > -1
    mean = lambda a, v: float(a[v].mean() * 100) if v.any() else float("nan")        # noqa: E731
    return {"AP": mean(ap, valid), "AP50": mean(ap[0], valid[0]), "AP75": mean(ap[5], valid[5])}


class BoxAPEvaluator:
    """[D2] DatasetEvaluator protocol. Ground truth comes from the inputs themselves: dict["instances"] (gt_boxes,
    gt_classes, in the network-input frame; rescaled to dict["height"] x dict["width"] like the detections) or
    dict["annotations"] (original-image coordinates, XYXY_ABS or XYWH_ABS)."""

    def reset(self):
        self._dets, self._gts = {}, {}

    def process(self, inputs, outputs):
        for inp, out in zip(inputs, outputs):
            key = inp.get("image_id", len(self._gts))
            inst = out["instances"]
            self._dets[key] = (inst.pred_boxes.tensor.detach().float().cpu().numpy(), inst.pred_classes.detach().cpu().numpy(),
                               inst.scores.detach().float().cpu().numpy())
            if "annotations" in inp:            # dataset annotations: original-image coordinates, any absolute bbox_mode
                from ..data.dataset_mapper import bbox_xyxy
                ann = [a for a in inp["annotations"] if not a.get("iscrowd", 0)]
                self._gts[key] = (np.array([bbox_xyxy(a) for a in ann], dtype=np.float64).reshape(-1, 4),
                                  np.array([a["category_id"] for a in ann], dtype=np.int64))
            else:                               # Instances live in the network-input frame: bring them to the frame the
                g = inp["instances"]            # detections were post-processed to (dict height x width)
                b = g.gt_boxes.tensor.detach().float().cpu().numpy().astype(np.float64)
                ih, iw = g.image_size
                sy, sx = inp.get("height", ih) / float(ih), inp.get("width", iw) / float(iw)
                self._gts[key] = (b * np.array([sx, sy, sx, sy]), g.gt_classes.detach().cpu().numpy())

    def evaluate(self):
        return {"bbox": coco_box_ap(self._dets, self._gts)}
