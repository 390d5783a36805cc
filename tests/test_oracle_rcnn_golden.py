"""Pins the Faster R-CNN CPU oracle (oracle/ut2_rcnn_oracle.py) against fixtures minted from the reference's own
rpn.py / roi_heads.py / fast_rcnn.py / trainer.py (oracle/gen_golden_rcnn.py). Index outputs exact, fp32 rtol 1e-5.
Also cross-checks the restated ROIAlign and NMS against the installed torchvision.ops."""
import os

import pytest
import torch

from oracle import ut2_rcnn_oracle as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def close(a, b, rtol=1e-5, atol=1e-6):
    assert a.shape == b.shape
    assert torch.allclose(a.float(), b.float(), rtol=rtol, atol=atol), (a.float() - b.float()).abs().max()


def test_rpn_labels_and_losses():
    g = load("rcnn_rpn_losses.pt")
    anchors = torch.cat(R.generate_anchors(g["level_hw"]))
    lab, midx, mb, conf = R.rpn_label_and_sample(anchors, g["gt_boxes"], g["gt_scores"], g["keys"])
    for i in range(len(lab)):
        assert torch.equal(lab[i], g["labels"][i])
        assert torch.equal(mb[i], g["matched_boxes"][i])
        assert torch.equal(conf[i].float(), g["confids"][i].float())
        assert int((lab[i] == 1).sum()) <= 64 and int((lab[i] >= 0).sum()) == 256
    for tag, cf in (("pseudo", conf), ("sup", None)):
        lg = g["logits"].clone().requires_grad_(True)
        dl = g["deltas"].clone().requires_grad_(True)
        ls = R.rpn_losses(anchors, lg, dl, lab, mb, cf)
        (ls["loss_rpn_cls"] * 1.7 + ls["loss_rpn_loc"] * 0.6).backward()
        close(ls["loss_rpn_cls"], g[tag]["loss_rpn_cls"])
        close(ls["loss_rpn_loc"], g[tag]["loss_rpn_loc"])
        close(lg.grad, g[tag]["dlogits"])
        close(dl.grad, g[tag]["ddeltas"])


@pytest.mark.parametrize("mode", ["pseudo", "sup"])
def test_roi_sampling(mode):
    g = load("rcnn_roi_sampling.pt")
    for i in range(len(g["proposals"])):
        out = R.roi_label_and_sample(g["proposals"][i], g["gt_boxes"][i], g["gt_classes"][i],
                                     g["gt_scores"][i] if mode == "pseudo" else None,
                                     g["gt_std"][i] if mode == "pseudo" else None, g["keys"][i])
        ref = g[mode][i]
        assert torch.equal(out["proposal_boxes"], ref["proposal_boxes"])
        assert torch.equal(out["gt_classes"], ref["gt_classes"])
        assert torch.equal(out["gt_boxes"], ref["gt_boxes"])
        if mode == "pseudo":
            assert torch.equal(out["gt_confid"].float(), ref["gt_confid"].float())
            assert torch.equal(out["gt_loc_std"], ref["gt_loc_std"])
        n_fg = int((out["gt_classes"] < 80).sum())
        assert n_fg <= 128 and len(out["gt_classes"]) <= 512


@pytest.mark.parametrize("branch", ["supervised", "unsup_data_train"])
def test_box_losses(branch):
    g = load("rcnn_box_losses.pt")
    cat = lambda k: torch.cat([x[k] for x in g["sampled"]])
    s, d, t = [g[k].clone().requires_grad_(True) for k in ("scores", "deltas", "deltas_std")]
    ls = R.fast_rcnn_losses(s, d, t, cat("proposal_boxes"), cat("gt_boxes"), cat("gt_classes"), branch, cat("gt_loc_std"))
    (ls["loss_cls"] * 1.3 + ls["loss_box_reg"] * 0.7).backward()
    ref = g[branch]
    close(ls["loss_cls"], ref["loss_cls"])
    close(ls["loss_box_reg"], ref["loss_box_reg"])
    close(s.grad, ref["dscores"])
    close(d.grad if d.grad is not None else torch.zeros_like(d), ref["ddeltas"], atol=1e-6)
    close(t.grad if t.grad is not None else torch.zeros_like(t), ref["dstd"], atol=1e-6)


def test_inference_and_threshold():
    g = load("rcnn_inference.pt")
    off = 0
    for i, pb in enumerate(g["proposal_boxes"]):
        n = pb.shape[0]
        out = R.fast_rcnn_inference(g["scores"][off:off + n], g["deltas"][off:off + n], g["deltas_std"][off:off + n], pb,
                                    g["image_size"], device_type="cpu")
        off += n
        ref = g["detections"][i]
        assert torch.equal(out["roi_index"], g["keep_idx"][i])
        assert torch.equal(out["pred_classes"], ref["pred_classes"])
        close(out["pred_boxes"], ref["pred_boxes"])
        close(out["scores"], ref["scores"])
        close(out["pred_boxes_std"], ref["pred_boxes_std"])
        assert len(out["scores"]) <= 100
        th = R.threshold_bbox_rcnn(out, 0.7)
        rt = g["thresholded"][i]
        assert torch.equal(th["gt_classes"], rt["gt_classes"])
        close(th["gt_boxes"], rt["gt_boxes"])
        close(th["pred_boxes_std"], rt["pred_boxes_std"])
        assert (th["scores"] > 0.7).all()


def test_roi_align_restatement_matches_torchvision():
    from torchvision.ops import roi_align
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(2, 5, 13, 17, generator=g)
    rois = torch.tensor([[0, 3.2, 4.1, 40.7, 30.3], [1, -5.0, -3.0, 20.0, 60.0], [0, 10.0, 10.0, 10.5, 10.2],
                         [1, 30.0, 20.0, 70.0, 55.0]])
    for scale in (0.25, 0.125):
        close(R.roi_align_ref(feat, rois, 7, scale), roi_align(feat, rois, 7, scale, 0, True), rtol=1e-4, atol=1e-5)


def test_nms_restatement_matches_torchvision():
    from torchvision.ops import batched_nms
    g = torch.Generator().manual_seed(5)
    xy = torch.rand(600, 2, generator=g) * 200
    wh = torch.rand(600, 2, generator=g) * 60 + 4
    boxes = torch.cat([xy, xy + wh], 1)
    scores = torch.rand(600, generator=g)
    cls = torch.randint(0, 4, (600,), generator=g)
    assert torch.equal(R.batched_nms(boxes, scores, cls, 0.5, "cpu"), batched_nms(boxes, scores, cls, 0.5))


def test_rpn_proposals_shape_and_order():
    g = torch.Generator().manual_seed(9)
    hw = [(16, 20), (8, 10), (4, 5), (2, 3), (1, 2)]
    anchors = R.generate_anchors(hw)
    logits = [torch.randn(2, a.shape[0], generator=g) for a in anchors]
    deltas = [torch.randn(2, a.shape[0], 4, generator=g) * 0.3 for a in anchors]
    out = R.rpn_find_top_proposals(anchors, logits, deltas, [(64, 80), (60, 70)], pre_topk=200, post_topk=50)
    for o, (h, w) in zip(out, [(64, 80), (60, 70)]):
        b, s = o["proposal_boxes"], o["objectness_logits"]
        assert len(s) <= 50 and (s[:-1] >= s[1:]).all()
        assert (b[:, 0] >= 0).all() and (b[:, 2] <= w).all() and (b[:, 3] <= h).all()
        assert ((b[:, 2] - b[:, 0]) > 0).all()


def test_oracle_model_runs_end_to_end_on_cpu():
    """The restated detector + step (oracle/ut2_rcnn_model.py) on a tiny input: losses finite, gradients reach the trunk,
    the teacher branch yields <= 100 detections per image."""
    from oracle import ut2_rcnn_model as RM
    sd = RM.init_state_dict(3)
    n = sum(v.numel() for k, v in sd.items() if ".norm." not in k)
    assert abs(n - 41.38e6) < 0.05e6
    g = torch.Generator().manual_seed(4)
    imgs = [torch.randint(0, 256, (3, 64, 96), generator=g, dtype=torch.uint8)]
    gt = {"boxes": [torch.tensor([[8.0, 10.0, 60.0, 50.0], [30.0, 5.0, 90.0, 40.0]])], "classes": [torch.tensor([3, 17])]}
    A = 3 * (16 * 24 + 8 * 12 + 4 * 6 + 2 * 3 + 1 * 2)
    keys = [torch.randint(0, 2 ** 32, (A,), generator=g, dtype=torch.int64)]
    kroi = [torch.randint(0, 2 ** 32, (1200,), generator=g, dtype=torch.int64)]
    w = sd["backbone.bottom_up.res3.0.conv1.weight"].requires_grad_(True)
    losses, aux = RM.forward_train(sd, imgs, gt, "supervised", keys, kroi)
    assert set(losses) == {"loss_rpn_cls", "loss_rpn_loc", "loss_cls", "loss_box_reg"}
    sum(losses.values()).backward()
    assert all(torch.isfinite(v) for v in losses.values()) and w.grad is not None and float(w.grad.abs().sum()) > 0
    sd["backbone.bottom_up.res3.0.conv1.weight"] = w.detach()
    props, dets, _ = RM.forward_teacher(sd, imgs)
    assert len(dets[0]["scores"]) <= 100 and props[0].shape[0] <= 1000
