"""TEST INFRASTRUCTURE ONLY. Mints the golden fixtures under tests/golden/ by executing the
reference's own Python (read-only, from /root/reference) on seeded synthetic inputs on CPU.

Run here (the build container) only:  python oracle/gen_golden.py
The GPU box has no /root/reference, so the resulting small .pt files are committed.
Every fixture stores inputs and the reference outputs (and input gradients where a loss is involved).
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "unbiased-teacher-v2_b200")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def build_cfg():
    """Phase 1: resolve the FCOS sup1 recipe with the product's config loader, then drop the
    product package from sys.modules so that the name `ubteacher` can be re-bound to the reference."""
    sys.path.insert(0, PKG)
    from ubteacher.config import add_ubteacher_config
    from ubteacher.d2compat.config import get_cfg

    cfg = get_cfg()
    add_ubteacher_config(cfg)
    cfg.merge_from_file(os.path.join(PKG, "configs/FCOS/coco-standard/fcos_R_50_ut2_sup1_run0.yaml"))
    rcfg = get_cfg()
    add_ubteacher_config(rcfg)
    rcfg.merge_from_file(os.path.join(PKG, "configs/Faster-RCNN/coco-standard/faster_rcnn_R_50_FPN_ut2_sup1_run0.yaml"))
    sys.path.remove(PKG)
    for k in [k for k in sys.modules if k == "ubteacher" or k.startswith("ubteacher.")]:
        del sys.modules[k]
    return cfg, rcfg


LEVEL_HW = [(16, 20), (8, 10), (4, 5), (2, 3), (1, 2)]
STRIDES = [8, 16, 32, 64, 128]
IMG_HW = (128, 160)


def rand_boxes(g, n, H, W, min_size=6.0):
    x1 = torch.rand(n, generator=g) * (W - min_size - 1)
    y1 = torch.rand(n, generator=g) * (H - min_size - 1)
    w = min_size + torch.rand(n, generator=g) * (W - x1 - min_size)
    h = min_size + torch.rand(n, generator=g) * (H - y1 - min_size)
    return torch.stack([x1, y1, x1 + w * torch.rand(n, generator=g).clamp(min=0.15),
                        y1 + h * torch.rand(n, generator=g).clamp(min=0.15)], dim=1)


def dense_preds(g, N, scale_logit=1.5, shift_logit=-3.0):
    logits, reg, std, ctr = [], [], [], []
    for (h, w) in LEVEL_HW:
        logits.append(torch.randn(N, 80, h, w, generator=g) * scale_logit + shift_logit)
        reg.append(torch.randn(N, 68, h, w, generator=g) * 1.5)
        std.append(torch.randn(N, 4, h, w, generator=g) * 2.0)
        ctr.append(torch.randn(N, 1, h, w, generator=g))
    return logits, reg, std, ctr


def main():
    cfg, rcfg = build_cfg()
    import refshim

    load, st = refshim.install()
    fo = load("ubteacher/modeling/fcos/fcos_outputs.py", "ubteacher.modeling.fcos.fcos_outputs")
    pg = load("ubteacher/modeling/pseudo_generator.py", "ubteacher.modeling.pseudo_generator")
    br = load("ubteacher/modeling/box_regression.py", "ubteacher.modeling.box_regression")
    fr = load("ubteacher/modeling/roi_heads/fast_rcnn.py", "ubteacher.modeling.roi_heads.fast_rcnn")
    comm = sys.modules["ubteacher.utils.comm"]
    Instances, Boxes = st.Instances, st.Boxes
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)

    # ---------------------------------------------------------------- locations
    locs = {}
    for (h, w), s in zip(LEVEL_HW + [(100, 168), (7, 11)], STRIDES + [8, 128]):
        locs[f"{h}x{w}s{s}"] = comm.compute_locations(h, w, s, torch.device("cpu"))
    torch.save(locs, os.path.join(OUT, "compute_locations.pt"))

    outputs = fo.FCOSOutputs(cfg)
    outputs.train()
    locations = [comm.compute_locations(h, w, s, torch.device("cpu")) for (h, w), s in zip(LEVEL_HW, STRIDES)]

    # ---------------------------------------------------------------- targets (labeled + pseudo sets)
    g = torch.Generator().manual_seed(20260)
    N = 3
    gt = []
    for i, n in enumerate([5, 0, 9]):
        inst = Instances(IMG_HW)
        inst.gt_boxes = Boxes(rand_boxes(g, n, *IMG_HW))
        inst.gt_classes = torch.randint(0, 80, (n,), generator=g)
        gt.append(inst)
    tt = outputs._get_ground_truth(locations, gt, False)
    torch.save({"boxes": [x.gt_boxes.tensor for x in gt], "classes": [x.gt_classes for x in gt],
                "targets": {k: [t.clone() for t in v] for k, v in tt.items()}},
               os.path.join(OUT, "fcos_targets_labeled.pt"))

    pseudo = []
    for i, n in enumerate([4, 7, 0]):
        inst = Instances(IMG_HW)
        inst.gt_boxes = Boxes(rand_boxes(g, n, *IMG_HW))
        inst.gt_classes = torch.randint(0, 80, (n,), generator=g)
        inst.scores = torch.rand(n, generator=g)
        inst.reg_pred_std = torch.randn(n, 4, generator=g) * 2
        pseudo.append(inst)
    tp = outputs._get_ground_truth(locations, pseudo, False)
    torch.save({"boxes": [x.gt_boxes.tensor for x in pseudo], "classes": [x.gt_classes for x in pseudo],
                "scores": [x.scores for x in pseudo], "reg_pred_std": [x.reg_pred_std for x in pseudo],
                "targets": {k: [t.clone() for t in v] for k, v in tp.items()}},
               os.path.join(OUT, "fcos_targets_pseudo.pt"))

    # ---------------------------------------------------------------- targets with centre sampling (config.py:151-152 defaults;
    # the shipped recipes switch it off) and the ignore_near keep mask, on the two box sets above (no new random draws)
    outputs.center_sample, outputs.radius = True, 1.5
    rec = {"radius": 1.5}
    for name, insts in (("labeled", gt), ("pseudo", pseudo)):
        for ign in (False, True):
            t = outputs._get_ground_truth(locations, insts, ign)
            rec[f"{name}_ignore_near{int(ign)}"] = {k: [x.clone() for x in v] for k, v in t.items()}
    torch.save(rec, os.path.join(OUT, "fcos_targets_center_sample.pt"))
    outputs.center_sample, outputs.radius = cfg.MODEL.FCOS.CENTER_SAMPLE, cfg.MODEL.FCOS.POS_RADIUS

    # ---------------------------------------------------------------- supervised losses (+ grads)
    def leafs(ts):
        return [t.clone().requires_grad_(True) for t in ts]

    logits, reg, std, ctr = dense_preds(g, N)
    L, R, S, C = leafs(logits), leafs(reg), leafs(std), leafs(ctr)
    extras, losses = outputs.losses(L, R, C, locations, gt, S, [], False, branch="labeled")
    total = sum(v * (i + 1) for i, v in enumerate(losses.values()))
    total.backward()
    torch.save({"logits": logits, "reg": reg, "std": std, "ctr": ctr,
                "boxes": [x.gt_boxes.tensor for x in gt], "classes": [x.gt_classes for x in gt],
                "losses": {k: v.detach() for k, v in losses.items()}, "loss_denorm": float(extras["loss_denorm"]),
                "grads": {"logits": [t.grad for t in L], "reg": [t.grad for t in R],
                          "std": [t.grad for t in S], "ctr": [t.grad for t in C]}},
               os.path.join(OUT, "fcos_losses_labeled.pt"))

    # the same inputs through the KL_LOSS_TYPE "klloss" branch (config.py:198 default; the shipped recipes use "nlloss") for the
    # four LOC_FUN_ALL reductions (fcos_outputs.py:380-397, layers/kl_loss.py:11-73); no new random draws
    kl_mod = sys.modules["ubteacher.layers.kl_loss"]
    saved = (outputs.kl_loss_type, outputs.kl_loc_loss_func, outputs.loc_fun_all)
    outputs.kl_loss_type, outputs.kl_loc_loss_func = "klloss", kl_mod.KLLoss()
    rec_kl = {}
    for method in ("mean", "sum", "weight_ctr_sum", "weight_ctr_mean"):
        outputs.loc_fun_all = method
        L2, R2, S2, C2 = leafs(logits), leafs(reg), leafs(std), leafs(ctr)
        _, lk = outputs.losses(L2, R2, C2, locations, gt, S2, [], False, branch="labeled")
        sum(v * (i + 1) for i, v in enumerate(lk.values())).backward()
        rec_kl[method] = {"losses": {k: v.detach() for k, v in lk.items()}}
        if method in ("mean", "weight_ctr_mean"):       # gradients of the regression / uncertainty maps (the others do not change)
            rec_kl[method]["grads"] = {"reg": [t.grad for t in R2], "std": [t.grad for t in S2]}
    outputs.kl_loss_type, outputs.kl_loc_loss_func, outputs.loc_fun_all = saved
    torch.save(rec_kl, os.path.join(OUT, "fcos_losses_labeled_klloss.pt"))

    # supervised loss with no positive location at all (all-background batch -> every loss * 0)
    gt0 = []
    for _ in range(2):
        inst = Instances(IMG_HW)
        inst.gt_boxes = Boxes(torch.zeros(0, 4))
        inst.gt_classes = torch.zeros(0, dtype=torch.long)
        gt0.append(inst)
    l0, r0, s0, c0 = dense_preds(g, 2)
    _, losses0 = outputs.losses(l0, r0, c0, locations, gt0, s0, [], False, branch="labeled")
    torch.save({"logits": l0, "reg": r0, "std": s0, "ctr": c0,
                "losses": {k: torch.as_tensor(v).float() for k, v in losses0.items()}},
               os.path.join(OUT, "fcos_losses_labeled_empty.pt"))

    # ---------------------------------------------------------------- pseudo losses (+ grads)
    logits, reg, std, ctr = dense_preds(g, N)
    L, R, S, C = leafs(logits), leafs(reg), leafs(std), leafs(ctr)
    pseudo_reg = []
    for i, n in enumerate([6, 0, 3]):
        inst = Instances(IMG_HW)
        inst.gt_boxes = Boxes(rand_boxes(g, n, *IMG_HW))
        inst.gt_classes = torch.randint(0, 80, (n,), generator=g)
        inst.scores = torch.rand(n, generator=g)
        inst.reg_pred_std = torch.randn(n, 4, generator=g) * 2 - 2.0
        pseudo_reg.append(inst)
    extras, losses = outputs.pseudo_losses(L, R, C, locations, {"cls": pseudo, "reg": pseudo_reg}, S, [], False,
                                           branch="unlabeled")
    total = sum(v * (i + 1) for i, (k, v) in enumerate(losses.items()) if k.startswith("loss"))
    total.backward()

    def pack(insts):
        return {"boxes": [x.gt_boxes.tensor for x in insts], "classes": [x.gt_classes for x in insts],
                "scores": [x.scores for x in insts], "reg_pred_std": [x.reg_pred_std for x in insts]}

    torch.save({"logits": logits, "reg": reg, "std": std, "ctr": ctr, "cls_set": pack(pseudo),
                "reg_set": pack(pseudo_reg),
                "losses": {k: torch.as_tensor(v).detach().float() for k, v in losses.items()},
                "loss_order": list(losses.keys()),
                "grads": {"logits": [t.grad for t in L], "reg": [t.grad for t in R],
                          "std": [t.grad if t.grad is not None else torch.zeros_like(t) for t in S],
                          "ctr": [t.grad for t in C]}},
               os.path.join(OUT, "fcos_losses_pseudo.pt"))

    # ---------------------------------------------------------------- predict_proposals, 4 criteria x {small, large}
    for tag, (sc, sh) in {"small": (1.0, -5.2), "large": (1.5, -2.0)}.items():
        logits, reg, std, ctr = dense_preds(g, 2, sc, sh)
        rec = {"logits": logits, "reg": reg, "std": std, "ctr": ctr, "image_sizes": [IMG_HW, (120, 150)], "out": {}}
        for method in ["cls", "cls_n_ctr", "cls_n_loc"]:  # "ctr" raises IndexError in the reference (fcos_outputs.py:1193,1202)
            with torch.no_grad():
                res = outputs.predict_proposals(logits, reg, ctr, locations, rec["image_sizes"], std, [], method)
            rec["out"][method] = [
                {k: (v.tensor if isinstance(v, Boxes) else v) for k, v in r.get_fields().items()} for r in res]
        rec["n_pre_nms"] = [[int((lv[i].sigmoid() > 0.05).sum().clamp(max=1000)) for lv in logits] for i in range(2)]
        torch.save(rec, os.path.join(OUT, f"fcos_predict_{tag}.pt"))

    # ---------------------------------------------------------------- pseudo-label thresholding
    gen = pg.PseudoGenerator(cfg)
    n = 40
    inst = Instances(IMG_HW)
    inst.pred_boxes = Boxes(rand_boxes(g, n, *IMG_HW))
    inst.scores = torch.rand(n, generator=g)
    inst.pred_classes = torch.randint(0, 80, (n,), generator=g)
    inst.centerness = torch.rand(n, generator=g)
    inst.cls_confid = torch.rand(n, generator=g)
    inst.reg_pred_std = torch.randn(n, 4, generator=g)
    empty = inst[torch.zeros(n, dtype=torch.bool)]
    a, na = gen.process_pseudo_label([inst, empty], 0.5, "roih", "thresholding")
    b, nb = gen.process_pseudo_label([inst, empty], (0.5, 0.3), "roih", "thresholding_cls_ctr")

    def fields(x):
        return {k: (v.tensor if isinstance(v, Boxes) else v) for k, v in x.get_fields().items()}

    torch.save({"inst": fields(inst), "thresholding": [fields(x) for x in a], "n_thresholding": na,
                "thresholding_cls_ctr": [fields(x) for x in b], "n_cls_ctr": nb},
               os.path.join(OUT, "pseudo_threshold.pt"))

    # ---------------------------------------------------------------- small loss pieces
    P = 37
    pred = torch.rand(P, 4, generator=g) * 6
    tgt = torch.rand(P, 4, generator=g) * 6
    wgt = torch.rand(P, generator=g)
    stdp = torch.randn(P, 4, generator=g)
    dist = torch.randn(P, 68, generator=g) * 2
    rec = {"pred": pred, "tgt": tgt, "w": wgt, "std": stdp, "dist": dist}
    rec["integral"] = outputs.integral(dist)
    rec["ctrness"] = fo.compute_ctrness_targets(tgt)
    rec["iou_targets"] = fo.compute_iou_targets(pred, tgt)
    for kind in ["iou", "linear_iou", "giou"]:
        rec["iouloss_" + kind] = sys.modules["ubteacher.layers.iou_loss"].IOULoss(kind)(pred, tgt, wgt)
    kl = sys.modules["ubteacher.layers.kl_loss"]
    rec["nlloss"] = kl.NLLoss()(pred, stdp, tgt, weight=wgt, iou_weight=rec["iou_targets"], loss_denorm=3.0, method="mean")
    rec["klloss_mean"] = kl.KLLoss()(pred, stdp, tgt, weight=wgt, iou_weight=None, loss_denorm=3.0, method="mean")
    rec["klloss_wsum"] = kl.KLLoss()(pred, stdp, tgt, weight=wgt, iou_weight=None, loss_denorm=3.0, method="weight_ctr_sum")
    torch.save(rec, os.path.join(OUT, "loss_pieces.pt"))

    # ---------------------------------------------------------------- Box2BoxXYXYTransform (R-CNN)
    tr = br.Box2BoxXYXYTransform(weights=(10.0, 10.0, 5.0, 5.0))
    src = rand_boxes(g, 50, 600, 800)
    dst = rand_boxes(g, 50, 600, 800)
    deltas = tr.get_deltas(src, dst)
    rnd = torch.randn(50, 4, generator=g) * 3
    torch.save({"src": src, "dst": dst, "deltas": deltas, "rand_deltas": rnd,
                "applied": tr.apply_deltas(rnd, src), "applied_roundtrip": tr.apply_deltas(deltas, src)},
               os.path.join(OUT, "box2box_xyxy.pt"))

    # ---------------------------------------------------------------- R-CNN loss pieces
    x = torch.randn(64, 81, generator=g)
    y = torch.randint(0, 81, (64,), generator=g)
    fl = fr.FocalLoss(gamma=1.5, num_classes=80)
    b1 = rand_boxes(g, 30, 600, 800)
    b2 = b1 + torch.randn(30, 4, generator=g) * 4
    rec = {"x": x, "y": y, "focal_sum": fl(x, y), "b1": b1, "b2": b2,
           "matched_iou": fr.matched_boxlist_iou(Boxes(b1), Boxes(b2))}
    mu, sg, tg, iw = (torch.randn(30, 4, generator=g), torch.randn(30, 4, generator=g),
                      torch.randn(30, 4, generator=g), torch.rand(30, generator=g))
    rec.update({"nl_mu": mu, "nl_std": sg, "nl_tgt": tg, "nl_iou": iw, "nl_loss": fr.nl_loss(mu, sg, tg, 0.0, iou_weight=iw, reduction="sum")})
    torch.save(rec, os.path.join(OUT, "rcnn_loss_pieces.pt"))

    # ---------------------------------------------------------------- EMA update rule (trainer.py:468-486)
    s = torch.randn(1000, generator=g)
    t = torch.randn(1000, generator=g)
    rec = {"student": s, "teacher": t}
    for keep in [0.0, 0.9996, 0.9999]:
        rec[f"keep{keep}"] = s * (1 - keep) + t * keep
    torch.save(rec, os.path.join(OUT, "ema.pt"))

    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))}
    print(sizes, "total", sum(sizes.values()))


if __name__ == "__main__":
    main()
