"""TEST INFRASTRUCTURE ONLY. Mints the Faster R-CNN golden fixtures (tests/golden/rcnn_*.pt) by executing the
reference's own Python — modeling/proposal_generator/rpn.py, modeling/roi_heads/{roi_heads,fast_rcnn}.py and the
``threshold_bbox`` method of engine/trainer.py — read-only from /root/reference, on seeded synthetic inputs on CPU.

Detectron2 is not installed: the [D2] base classes (RPN, StandardROIHeads, FastRCNNOutputLayers) are empty stubs and
the [D2] helpers the reference methods call (pairwise_iou, Matcher, subsample_labels, _dense_box_regression_loss,
add_ground_truth_to_proposals, fast_rcnn_inference) are bound to the restatements in oracle/ut2_rcnn_oracle.py, so
the fixtures pin exactly the logic the reference owns (score / std gathers, loss weighting, tsbetter masks, ...).

Run here (the build container) only:  python oracle/gen_golden_rcnn.py
"""
import ast
import os
import sys
import types

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

LEVEL_HW = [(32, 40), (16, 20), (8, 10), (4, 5), (2, 3)]
IMG_HW = (128, 160)


def rand_boxes(g, n, H, W, min_size=8.0):
    x1 = torch.rand(n, generator=g) * (W - min_size - 1)
    y1 = torch.rand(n, generator=g) * (H - min_size - 1)
    w = min_size + torch.rand(n, generator=g) * (W - x1 - min_size)
    h = min_size + torch.rand(n, generator=g) * (H - y1 - min_size)
    return torch.stack([x1, y1, x1 + w, y1 + h], dim=1)


class _Storage:
    def __init__(self):
        self.scalars = {}

    def put_scalar(self, k, v):
        self.scalars[k] = float(v)


def main():
    import refshim
    from oracle import ut2_rcnn_oracle as R

    load, st = refshim.install()
    Boxes, Instances = st.Boxes, st.Instances
    storage = _Storage()

    class _Registry:
        def register(self, obj=None):
            return obj if obj is not None else (lambda o: o)

    # ---- [D2] stubs bound to the oracle restatements
    def d2_dense_box_regression_loss(anchors, box2box_transform, pred_anchor_deltas, gt_boxes, fg_mask,
                                     box_reg_loss_type="smooth_l1", smooth_l1_beta=0.0):
        a = Boxes.cat(anchors).tensor
        tgt = torch.stack([R.box2box_get_deltas(a, k) for k in gt_boxes])
        return torch.abs(st.cat(pred_anchor_deltas, dim=1)[fg_mask] - tgt[fg_mask]).sum()

    def d2_pairwise_iou(b1, b2):
        return R.pairwise_iou(b1.tensor, b2.tensor)

    def d2_add_gt(gt_boxes, proposals):
        out = []
        for gb, p in zip(gt_boxes, proposals):
            gi = Instances(p.image_size)
            gi.proposal_boxes = gb
            gi.objectness_logits = torch.full((len(gb),), 23.0258509299)
            out.append(Instances.cat([p, gi]))
        return out

    def d2_fast_rcnn_inference(boxes, scores, image_shapes, score_thresh, nms_thresh, topk):
        res, idx = [], []
        for b, s, shp in zip(boxes, scores, image_shapes):
            h, w = shp
            valid = torch.isfinite(b).all(1) & torch.isfinite(s).all(1)
            s2 = s[:, :-1]
            bb = torch.stack((b[:, 0].clamp(0, w), b[:, 1].clamp(0, h), b[:, 2].clamp(0, w), b[:, 3].clamp(0, h)), 1)
            mask = (s2 > score_thresh) & valid[:, None]
            inds = mask.nonzero()
            cb, cs = bb[inds[:, 0]], s2[mask]
            keep = R.batched_nms(cb, cs, inds[:, 1], nms_thresh, "cpu")[:topk]
            r = Instances(shp)
            r.pred_boxes = Boxes(cb[keep])
            r.scores = cs[keep]
            r.pred_classes = inds[keep, 1]
            res.append(r)
            idx.append(inds[keep, 0])
        return res, idx

    m = sys.modules
    refshim._mod("detectron2.modeling.box_regression", Box2BoxTransform=object,
                 _dense_box_regression_loss=d2_dense_box_regression_loss)
    refshim._mod("detectron2.modeling.proposal_generator", RPN=type("RPN", (nn.Module,), {}))
    refshim._mod("detectron2.modeling.proposal_generator.build", PROPOSAL_GENERATOR_REGISTRY=_Registry())
    refshim._mod("detectron2.modeling.proposal_generator.proposal_utils", add_ground_truth_to_proposals=d2_add_gt)
    m["detectron2.structures"].pairwise_iou = d2_pairwise_iou
    m["detectron2.utils.events"].get_event_storage = lambda: storage
    refshim._mod("detectron2.utils.memory", retry_if_cuda_oom=lambda f: f)
    refshim._mod("detectron2.modeling.poolers", ROIPooler=object)
    rh = m["detectron2.modeling.roi_heads"]
    rh.ROI_HEADS_REGISTRY = _Registry()
    rh.StandardROIHeads = type("StandardROIHeads", (nn.Module,), {})
    refshim._mod("detectron2.modeling.roi_heads.box_head", build_box_head=None)
    frm = m["detectron2.modeling.roi_heads.fast_rcnn"]
    frm.FastRCNNOutputLayers = type("FastRCNNOutputLayers", (nn.Module,), {})
    frm.fast_rcnn_inference = d2_fast_rcnn_inference
    for name, rel in [("ubteacher.modeling.proposal_generator", "ubteacher/modeling/proposal_generator")]:
        mm = refshim._mod(name)
        mm.__path__ = [os.path.join(refshim.REF, rel)]

    load("ubteacher/modeling/box_regression.py", "ubteacher.modeling.box_regression")
    fr = load("ubteacher/modeling/roi_heads/fast_rcnn.py", "ubteacher.modeling.roi_heads.fast_rcnn")
    rpn_mod = load("ubteacher/modeling/proposal_generator/rpn.py", "ubteacher.modeling.proposal_generator.rpn")
    roi_mod = load("ubteacher/modeling/roi_heads/roi_heads.py", "ubteacher.modeling.roi_heads.roi_heads")

    g = torch.Generator().manual_seed(20261)
    H, W = IMG_HW
    anchors = R.generate_anchors(LEVEL_HW)
    A = sum(a.shape[0] for a in anchors)
    N = 3

    # ================================================================ RPN: labels + losses (rpn.py:78-225)
    gt_boxes = [rand_boxes(g, 5, H, W), rand_boxes(g, 2, H, W), torch.zeros(0, 4)]
    gt_scores = [torch.rand(5, generator=g) * 0.3 + 0.7, torch.rand(2, generator=g) * 0.3 + 0.7, torch.zeros(0)]
    keys = [torch.randint(0, 2 ** 32, (A,), generator=g, dtype=torch.int64) for _ in range(N)]
    logits = torch.randn(N, A, generator=g)
    deltas = torch.randn(N, A, 4, generator=g) * 0.5

    rpn = object.__new__(rpn_mod.PseudoLabRPN)
    nn.Module.__init__(rpn)
    rpn.anchor_boundary_thresh = -1
    rpn.batch_size_per_image = 256
    rpn.box2box_transform = None
    rpn.box_reg_loss_type = "smooth_l1"
    rpn.smooth_l1_beta = 0.0
    rpn.loss_weight = {"loss_rpn_cls": 1.0, "loss_rpn_loc": 1.0}
    rpn.anchor_matcher = lambda q: R.matcher(q, (0.3, 0.7), (0, -1, 1), True)
    counter = {"i": 0}

    def _subsample(label):
        pos, neg = R.subsample_labels(label, 256, 0.25, 0, keys[counter["i"]])
        counter["i"] += 1
        label.fill_(-1)
        label.scatter_(0, pos, 1)
        label.scatter_(0, neg, 0)
        return label

    rpn._subsample_labels = _subsample
    insts = []
    for b, s in zip(gt_boxes, gt_scores):
        it = Instances(IMG_HW)
        it.gt_boxes = Boxes(b)
        it.scores = s
        insts.append(it)
    anchor_boxes = [Boxes(a) for a in anchors]
    lab, mb, conf = rpn.label_and_sample_anchors_pseudo(anchor_boxes, insts)
    offs = [0]
    for a in anchors:
        offs.append(offs[-1] + a.shape[0])
    lg = logits.clone().requires_grad_(True)
    dl = deltas.clone().requires_grad_(True)
    split = lambda t: [t[:, offs[i]:offs[i + 1]] for i in range(len(anchors))]
    rec = {"level_hw": LEVEL_HW, "image_size": IMG_HW, "gt_boxes": gt_boxes, "gt_scores": gt_scores, "keys": keys,
           "logits": logits, "deltas": deltas, "labels": lab, "matched_boxes": mb, "confids": conf}
    for tag, cf in (("pseudo", conf), ("sup", None)):
        lg.grad = dl.grad = None
        ls = rpn.losses(anchor_boxes, split(lg), lab, split(dl), mb, cf)
        (ls["loss_rpn_cls"] * 1.7 + ls["loss_rpn_loc"] * 0.6).backward()
        rec[tag] = {"loss_rpn_cls": ls["loss_rpn_cls"].detach(), "loss_rpn_loc": ls["loss_rpn_loc"].detach(),
                    "dlogits": lg.grad.clone(), "ddeltas": dl.grad.clone()}
    torch.save(rec, os.path.join(OUT, "rcnn_rpn_losses.pt"))

    # ================================================================ ROI sampling (roi_heads.py:138-270)
    roi = object.__new__(roi_mod.StandardROIHeadsPseudoLab)
    nn.Module.__init__(roi)
    roi.proposal_append_gt = True
    roi.num_classes = 80
    roi.proposal_matcher = lambda q: R.matcher(q, (0.5,), (0, 1), False)
    P = 300
    gt_cls = [torch.randint(0, 80, (5,), generator=g), torch.randint(0, 80, (2,), generator=g),
              torch.zeros(0, dtype=torch.int64)]
    gt_std = [torch.randn(5, 4, generator=g), torch.randn(2, 4, generator=g), torch.zeros(0, 4)]
    props, rkeys = [], []
    for i in range(N):
        pb = rand_boxes(g, P, H, W)
        if gt_boxes[i].shape[0]:       # make some proposals overlap the ground truth
            j = torch.randint(0, gt_boxes[i].shape[0], (P // 3,), generator=g)
            pb[:P // 3] = gt_boxes[i][j] + torch.randn(P // 3, 4, generator=g) * 4
        props.append(pb)
        rkeys.append(torch.randint(0, 2 ** 32, (P + gt_boxes[i].shape[0],), generator=g, dtype=torch.int64))
    kc = {"i": 0}

    def _sample_proposals(matched_idxs, matched_labels, gt_classes):
        has_gt = gt_classes.numel() > 0
        if has_gt:
            c = gt_classes[matched_idxs]
            c[matched_labels == 0] = 80
            c[matched_labels == -1] = -1
        else:
            c = torch.zeros_like(matched_idxs) + 80
        fg, bg = R.subsample_labels(c, 512, 0.25, 80, rkeys[kc["i"]])
        kc["i"] += 1
        sel = torch.cat([fg, bg])
        return sel, c[sel]

    roi._sample_proposals = _sample_proposals

    def mk(pseudo):
        ps, ts = [], []
        for i in range(N):
            p = Instances(IMG_HW)
            p.proposal_boxes = Boxes(props[i])
            p.objectness_logits = torch.zeros(P)
            t = Instances(IMG_HW)
            t.gt_boxes = Boxes(gt_boxes[i])
            t.gt_classes = gt_cls[i]
            if pseudo:
                t.scores = gt_scores[i]
                t.pred_boxes_std = gt_std[i]
            ps.append(p)
            ts.append(t)
        return ps, ts

    fields = lambda x: {k: (v.tensor if isinstance(v, Boxes) else v) for k, v in x.get_fields().items()}
    rec = {"image_size": IMG_HW, "proposals": props, "gt_boxes": gt_boxes, "gt_classes": gt_cls, "gt_scores": gt_scores,
           "gt_std": gt_std, "keys": rkeys}
    kc["i"] = 0
    ps, ts = mk(True)
    rec["pseudo"] = [fields(x) for x in roi.label_and_sample_proposals_pseudo(ps, ts, branch="unsup_data_train")]
    kc["i"] = 0
    ps, ts = mk(False)
    rec["sup"] = [fields(x) for x in roi.label_and_sample_proposals(ps, ts, branch="supervised")]
    torch.save(rec, os.path.join(OUT, "rcnn_roi_sampling.pt"))

    # ================================================================ box predictor losses (fast_rcnn.py:834-1084)
    from ubteacher.modeling.box_regression import Box2BoxXYXYTransform
    layer = fr.FastRCNNFocaltLossBoundaryVarOutputLayers(
        st.ShapeSpec(channels=32), box2box_transform=Box2BoxXYXYTransform(weights=(10.0, 10.0, 5.0, 5.0)), num_classes=80,
        test_score_thresh=0.05, test_nms_thresh=0.5, test_topk_per_image=100, cls_agnostic_bbox_reg=True,
        smooth_l1_beta=0.0, box_reg_loss_type="nlloss", box_pseudo_reg_loss_type="tsbetter",
        loss_weight={"loss_box_reg": 1.0}, ts_better=0.1, t_cert=0.5)
    sampled = rec["pseudo"]
    Rr = sum(len(x["gt_classes"]) for x in sampled)
    scores = torch.randn(Rr, 81, generator=g) * 2
    dlt = torch.randn(Rr, 4, generator=g) * 0.8
    dstd = torch.randn(Rr, 4, generator=g) * 1.5
    out = {"scores": scores, "deltas": dlt, "deltas_std": dstd, "sampled": sampled}
    for branch in ("supervised", "unsup_data_train"):
        plist = []
        for x in sampled:
            p = Instances(IMG_HW)
            p.proposal_boxes = Boxes(x["proposal_boxes"])
            p.gt_classes = x["gt_classes"]
            p.gt_boxes = Boxes(x["gt_boxes"])
            p.gt_loc_std = x["gt_loc_std"]
            plist.append(p)
        s_, d_, t_ = [v.clone().requires_grad_(True) for v in (scores, dlt, dstd)]
        ls = layer.losses((s_, d_, t_), plist, branch)
        (ls["loss_cls"] * 1.3 + ls["loss_box_reg"] * 0.7).backward()
        zero = lambda v: v.grad.clone() if v.grad is not None else torch.zeros_like(v)
        out[branch] = {"loss_cls": ls["loss_cls"].detach(), "loss_box_reg": ls["loss_box_reg"].detach(),
                       "dscores": zero(s_), "ddeltas": zero(d_), "dstd": zero(t_)}
    torch.save(out, os.path.join(OUT, "rcnn_box_losses.pt"))

    # ================================================================ inference + threshold_bbox
    plist, pboxes = [], []
    for i in range(2):
        p = Instances(IMG_HW)
        pb = rand_boxes(g, 200, H, W)
        p.proposal_boxes = Boxes(pb)
        plist.append(p)
        pboxes.append(pb)
    sc = torch.randn(400, 81, generator=g) * 3
    sc[:, 80] += 1.0
    dd = torch.randn(400, 4, generator=g)
    ds = torch.randn(400, 4, generator=g)
    with torch.no_grad():
        res, keep_idx = layer.inference((sc, dd, ds), plist)
    src = open(os.path.join(refshim.REF, "ubteacher/engine/trainer.py")).read()
    fn = None
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and node.name == "UBRCNNTeacherTrainer":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == "threshold_bbox":
                    fn = sub
    ns = {"Instances": Instances, "Boxes": Boxes}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "trainer.py:threshold_bbox", "exec"), ns)
    thr = [ns["threshold_bbox"](None, r, thres=0.7, proposal_type="roih") for r in res]
    torch.save({"image_size": IMG_HW, "proposal_boxes": pboxes, "scores": sc, "deltas": dd, "deltas_std": ds,
                "detections": [fields(r) for r in res], "keep_idx": keep_idx, "thresholded": [fields(t) for t in thr]},
               os.path.join(OUT, "rcnn_inference.pt"))
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT)) if f.startswith("rcnn_")}
    print(sizes, "total", sum(sizes.values()))


if __name__ == "__main__":
    main()
