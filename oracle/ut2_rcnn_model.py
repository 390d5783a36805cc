"""TEST INFRASTRUCTURE ONLY — plain PyTorch fp32 CPU restatement of the Faster R-CNN R50-FPN detector of Unbiased
Teacher v2 and of its training step, driven by a ``state_dict`` with the reference's key names.

Follows (paths under /root/reference/ubteacher; [D2] = Detectron2 v0.6, restated from SURVEY.md app. B):
  modeling/meta_arch/rcnn.py:8-72, [D2] build_resnet_fpn_backbone + LastLevelMaxPool, proposal_generator/rpn.py:21-225,
  roi_heads/roi_heads.py:75-270, roi_heads/fast_rcnn.py:715-1225, engine/trainer.py:786-912 (run_step_full_semisup),
  :950-968 (EMA), [D2] SGD. Used by tests (parity of the CUDA path) and by bench.py's cpu_baseline / --impl reference.
"""
import torch
import torch.nn.functional as F

from . import ut2_model as M
from . import ut2_oracle as O
from . import ut2_rcnn_oracle as R

STRIDES = (4, 8, 16, 32, 64)


def backbone(sd, x):
    """trunk -> FPN p2..p5, p6 = max_pool2d(p5, kernel 1, stride 2) ([D2] LastLevelMaxPool)."""
    o = M.fpn_topdown(sd, M.trunk(sd, x), (5, 4, 3, 2))
    return [o[2], o[3], o[4], o[5], F.max_pool2d(o[5], kernel_size=1, stride=2, padding=0)]


def rpn_head(sd, feats):
    """[D2] StandardRPNHead + the permutes of rpn.py:33-46: per level logits [N, HWA], deltas [N, HWA, 4]."""
    p = "proposal_generator.rpn_head."
    logits, deltas = [], []
    for f in feats:
        t = M.q(F.relu(M.conv(f, sd[p + "conv.weight"], sd[p + "conv.bias"], 1, 1)))
        lg = M.q(M.conv(t, sd[p + "objectness_logits.weight"], sd[p + "objectness_logits.bias"]))
        dl = M.q(M.conv(t, sd[p + "anchor_deltas.weight"], sd[p + "anchor_deltas.bias"]))
        N, _, H, W = lg.shape
        logits.append(lg.permute(0, 2, 3, 1).flatten(1))
        deltas.append(dl.view(N, -1, 4, H, W).permute(0, 3, 4, 1, 2).flatten(1, -2))
    return logits, deltas


def box_head(sd, pooled):
    """[D2] FastRCNNConvFCHead (2 x FC 1024 + ReLU) + the three predictors (fast_rcnn.py:760-766, :818-832)."""
    q, qw = M.q, M.qw                  # bf16 rounding points (identity outside ``with M.bf16_points()``)
    x = q(pooled.flatten(1))           # ROIAlign writes bf16
    x = q(F.relu(F.linear(x, qw(sd["roi_heads.box_head.fc1.weight"]), sd["roi_heads.box_head.fc1.bias"])))
    x = q(F.relu(F.linear(x, qw(sd["roi_heads.box_head.fc2.weight"]), sd["roi_heads.box_head.fc2.bias"])))
    bp = "roi_heads.box_predictor."
    return (q(F.linear(x, qw(sd[bp + "cls_score.weight"]), sd[bp + "cls_score.bias"])),
            q(F.linear(x, qw(sd[bp + "bbox_pred.weight"]), sd[bp + "bbox_pred.bias"])),
            q(F.linear(x, qw(sd[bp + "bbox_pred_std.weight"]), sd[bp + "bbox_pred_std.bias"])))


def preprocess(sd, images):
    mean = sd["pixel_mean"] if "pixel_mean" in sd else torch.tensor([103.530, 116.280, 123.675]).view(3, 1, 1)
    std = sd["pixel_std"] if "pixel_std" in sd else torch.ones(3, 1, 1)
    return M.preprocess({"pixel_mean": mean, "pixel_std": std}, images)


def forward_features(sd, images):
    x, sizes = preprocess(sd, images)
    feats = backbone(sd, x)
    logits, deltas = rpn_head(sd, feats)
    anchors = R.generate_anchors([(f.shape[2], f.shape[3]) for f in feats], STRIDES)
    return {"feats": feats, "logits": logits, "deltas": deltas, "anchors": anchors, "image_sizes": sizes}


def forward_train(sd, images, gt, branch, keys_rpn, keys_roi, proposals=None, device_type="cuda"):
    """rcnn.py:26-40 / :57-72. gt: dict(boxes, classes[, scores, std]) of per-image lists. keys_*: per-image sampling
    keys. `proposals` optionally injects the device's proposal boxes (parity tests: NMS borderlines)."""
    f = forward_features(sd, images)
    pseudo = "scores" in gt
    A = torch.cat(f["anchors"])
    lab, midx, mb, conf = R.rpn_label_and_sample(A, gt["boxes"], gt["scores"] if pseudo else None, keys_rpn)
    rl = R.rpn_losses(A, torch.cat(f["logits"], 1), torch.cat(f["deltas"], 1), lab, mb, conf)
    if proposals is None:
        with torch.no_grad():
            props = R.rpn_find_top_proposals(f["anchors"], [x.detach() for x in f["logits"]], [x.detach() for x in f["deltas"]],
                                             f["image_sizes"], device_type=device_type)
        proposals = [p["proposal_boxes"] for p in props]
    samp = [R.roi_label_and_sample(proposals[i], gt["boxes"][i], gt["classes"][i], gt["scores"][i] if pseudo else None,
                                   gt["std"][i] if pseudo else None, keys_roi[i]) for i in range(len(images))]
    pooled = R.roi_pool(f["feats"][:4], [s["proposal_boxes"] for s in samp])
    sc, dl, st = box_head(sd, pooled)
    cat = lambda k: torch.cat([s[k] for s in samp])
    bl = R.fast_rcnn_losses(sc, dl, st, cat("proposal_boxes"), cat("gt_boxes"), cat("gt_classes"), branch,
                            cat("gt_loc_std") if pseudo else None)
    losses = dict(rl)
    losses.update(bl)
    return losses, {"proposals": proposals, "samples": samp, "labels": lab}


def forward_teacher(sd, images, device_type="cuda", proposals=None):
    """rcnn.py:42-55 (unsup_data_weak): RPN proposals -> box head on all of them -> fast_rcnn_inference."""
    with torch.no_grad():
        f = forward_features(sd, images)
        if proposals is None:
            props = R.rpn_find_top_proposals(f["anchors"], f["logits"], f["deltas"], f["image_sizes"], device_type=device_type)
            proposals = [p["proposal_boxes"] for p in props]
        pooled = R.roi_pool(f["feats"][:4], proposals)
        sc, dl, st = box_head(sd, pooled)
        dets, off = [], 0
        for i, pb in enumerate(proposals):
            n = pb.shape[0]
            dets.append(R.fast_rcnn_inference(sc[off:off + n], dl[off:off + n], st[off:off + n], pb, f["image_sizes"][i],
                                              device_type=device_type))
            off += n
    return proposals, dets, (sc, dl, st)


def trainable_keys(sd):
    keys = []
    for k in sd:
        if ".norm." in k or k in ("pixel_mean", "pixel_std"):
            continue
        if k.startswith("backbone.bottom_up.stem.") or k.startswith("backbone.bottom_up.res2."):
            continue
        keys.append(k)
    return keys


def ut2_rcnn_step(student, teacher, mom, batch, cfg, first_step, keys, device_type="cuda", pseudo_override=None,
                  proposals_override=None):
    """One post-burn-in UT2 Faster R-CNN step (trainer.py:786-912) on CPU in fp32. batch = (label_q, label_k, unlabel_q,
    unlabel_k) lists of {"image": u8 CHW, "boxes", "classes"}; keys = dict(rpn_sup, roi_sup, rpn_unsup, roi_unsup) of
    per-image sampling keys. Returns (record, grads, pseudo-label set)."""
    lam, mu = cfg["UNSUP_LOSS_WEIGHT"], cfg["UNSUP_REG_LOSS_WEIGHT"]
    for k in teacher:                                             # EMA over the whole state_dict (:950-968)
        teacher[k] = O.ema_update(student[k].reshape(-1), teacher[k].reshape(-1), cfg["EMA_KEEP_RATE"]).reshape(teacher[k].shape)
    label_q, label_k, unlabel_q, unlabel_k = batch
    _, dets, _ = forward_teacher(teacher, [d["image"] for d in unlabel_k], device_type)
    th = [R.threshold_bbox_rcnn(d, cfg["BBOX_THRESHOLD"]) for d in dets]
    pseudo = {"boxes": [x["gt_boxes"] for x in th], "classes": [x["gt_classes"] for x in th],
              "scores": [x["scores"] for x in th], "std": [x["pred_boxes_std"] for x in th]}
    if pseudo_override is not None:
        pseudo = pseudo_override
    tk = trainable_keys(student)
    params = {k: student[k].detach().clone().requires_grad_(True) for k in tk}
    sd = dict(student)
    sd.update(params)
    lab = label_q + label_k
    po = proposals_override or {}
    sup, _ = forward_train(sd, [d["image"] for d in lab], {"boxes": [d["boxes"] for d in lab], "classes": [d["classes"] for d in lab]},
                           "supervised", keys["rpn_sup"], keys["roi_sup"], po.get("sup"), device_type)
    uns, _ = forward_train(sd, [d["image"] for d in unlabel_q], pseudo, "unsup_data_train", keys["rpn_unsup"], keys["roi_unsup"],
                           po.get("unsup"), device_type)
    record = dict(sup)
    record.update({k + "_pseudo": v for k, v in uns.items()})
    total = sum(sup.values()) + lam * (uns["loss_cls"] + uns["loss_rpn_cls"]) + mu * uns["loss_box_reg"] + 0.0 * uns["loss_rpn_loc"]
    total.backward()
    grads = {}
    for k in tk:
        g = params[k].grad if params[k].grad is not None else torch.zeros_like(params[k])
        grads[k] = g
        p, b = O.sgd_step(student[k], g, mom.get(k, torch.zeros_like(g)), cfg["LR"], cfg["MOMENTUM"], cfg["WEIGHT_DECAY"], first_step)
        student[k], mom[k] = p.detach(), b.detach()
    record = {k: (v.detach() if isinstance(v, torch.Tensor) else torch.tensor(float(v))) for k, v in record.items()}
    return record, grads, pseudo


def init_state_dict(seed=0):
    """A seeded random state_dict with the reference's Faster R-CNN key names / shapes — for CPU-only
    runs of this restatement (tests, cpu_baseline) when no device model is around to export one."""
    import math
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k, bn=False, bias=False, std=None):
        fan_out = cout * k * k
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * (std if std is not None else math.sqrt(2.0 / fan_out))
        if bias:
            sd[name + ".bias"] = torch.zeros(cout)
        if bn:
            sd[name + ".norm.weight"] = torch.full((cout,), 0.04 if "stem." in name else 0.25 if name.endswith("conv3") else 1.0)
            sd[name + ".norm.bias"] = torch.zeros(cout)
            sd[name + ".norm.running_mean"] = torch.zeros(cout)
            sd[name + ".norm.running_var"] = torch.ones(cout) - M.BN_EPS

    bu = "backbone.bottom_up."
    conv(bu + "stem.conv1", 64, 3, 7, bn=True)
    cin = 64
    for (stage, n, _), (mid, cout) in zip(M.STAGES, ((64, 256), (128, 512), (256, 1024), (512, 2048))):
        for i in range(n):
            p = f"{bu}{stage}.{i}."
            if cin != cout:
                conv(p + "shortcut", cout, cin, 1, bn=True)
            conv(p + "conv1", mid, cin, 1, bn=True)
            conv(p + "conv2", mid, mid, 3, bn=True)
            conv(p + "conv3", cout, mid, 1, bn=True)
            cin = cout
    for l, c in ((2, 256), (3, 512), (4, 1024), (5, 2048)):
        conv(f"backbone.fpn_lateral{l}", 256, c, 1, bias=True, std=math.sqrt(1.0 / c))      # c2_xavier-like
        conv(f"backbone.fpn_output{l}", 256, 256, 3, bias=True, std=math.sqrt(1.0 / 2304))
    rp = "proposal_generator.rpn_head."
    conv(rp + "conv", 256, 256, 3, bias=True, std=0.01)
    conv(rp + "objectness_logits", 3, 256, 1, bias=True, std=0.01)
    conv(rp + "anchor_deltas", 12, 256, 1, bias=True, std=0.01)
    for n, o, i, std in (("roi_heads.box_head.fc1", 1024, 12544, math.sqrt(1.0 / 12544)), ("roi_heads.box_head.fc2", 1024, 1024, math.sqrt(1.0 / 1024)),
                         ("roi_heads.box_predictor.cls_score", 81, 1024, 0.01), ("roi_heads.box_predictor.bbox_pred", 4, 1024, 0.001),
                         ("roi_heads.box_predictor.bbox_pred_std", 4, 1024, 0.0001)):
        sd[n + ".weight"] = torch.randn(o, i, generator=g) * std
        sd[n + ".bias"] = torch.zeros(o)
    return sd
