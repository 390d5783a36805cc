"""Pins the CPU oracle (oracle/ut2_oracle.py) against the golden fixtures minted from the
reference's own Python by oracle/gen_golden.py. Integer / index outputs: exact. fp32: rtol 1e-5."""
import os

import pytest
import torch

from oracle import ut2_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LEVEL_HW = [(16, 20), (8, 10), (4, 5), (2, 3), (1, 2)]
STRIDES = [8, 16, 32, 64, 128]


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def locs():
    return [O.compute_locations(h, w, s) for (h, w), s in zip(LEVEL_HW, STRIDES)]


def test_compute_locations():
    g = load("compute_locations.pt")
    for key, ref in g.items():
        hw, s = key.split("s")
        h, w = map(int, hw.split("x"))
        assert torch.equal(O.compute_locations(h, w, int(s)), ref)


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("name", ["fcos_targets_labeled.pt", "fcos_targets_pseudo.pt"])
def test_fcos_targets(name, fast):
    g = load(name)
    fn = O.fcos_assign_targets_fast if fast else O.fcos_assign_targets
    out = fn(locs(), g["boxes"], g["classes"], STRIDES, scores=g.get("scores"), reg_pred_std=g.get("reg_pred_std"))
    for k, ref in g["targets"].items():
        for lvl, (a, b) in enumerate(zip(out[k], ref)):
            assert a.shape == b.shape, (k, lvl)
            if a.dtype.is_floating_point:
                assert torch.equal(a, b.to(a.dtype)), (k, lvl)
            else:
                assert torch.equal(a.long(), b.long()), (k, lvl)


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("ignore_near", [False, True])
@pytest.mark.parametrize("which", ["labeled", "pseudo"])
def test_fcos_targets_center_sample(which, ignore_near, fast):
    """CENTER_SAMPLE / POS_RADIUS (get_sample_region) and ignore_near against the reference's _get_ground_truth."""
    g = load(f"fcos_targets_{which}.pt")
    ref = load("fcos_targets_center_sample.pt")
    fn = O.fcos_assign_targets_fast if fast else O.fcos_assign_targets
    out = fn(locs(), g["boxes"], g["classes"], STRIDES, scores=g.get("scores"), reg_pred_std=g.get("reg_pred_std"),
             center_sample=True, radius=ref["radius"], ignore_near=ignore_near)
    for k, r in ref[f"{which}_ignore_near{int(ignore_near)}"].items():
        for lvl, (a, b) in enumerate(zip(out[k], r)):
            assert a.shape == b.shape, (k, lvl)
            if a.dtype.is_floating_point:
                assert torch.equal(a, b.to(a.dtype)), (k, lvl)
            else:
                assert torch.equal(a.long(), b.long()), (k, lvl)


def test_loss_pieces():
    g = load("loss_pieces.pt")
    tc = lambda a, b: torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    tc(O.integral(g["dist"]), g["integral"])
    tc(O.ctrness_targets(g["tgt"]), g["ctrness"])
    tc(O.iou_targets(g["pred"], g["tgt"]), g["iou_targets"])
    for kind in ["iou", "linear_iou", "giou"]:
        tc(O.iou_loss(g["pred"], g["tgt"], g["w"], kind), g["iouloss_" + kind])
    tc(O.nl_loss_fcos(g["pred"], g["std"], g["tgt"], g["iou_targets"]), g["nlloss"])
    tc(O.kl_loss(g["pred"], g["std"], g["tgt"], g["w"], loss_denorm=3.0, method="mean"), g["klloss_mean"])
    tc(O.kl_loss(g["pred"], g["std"], g["tgt"], g["w"], loss_denorm=3.0, method="weight_ctr_sum"), g["klloss_wsum"])


def test_fcos_losses_labeled_and_grads():
    g = load("fcos_losses_labeled.pt")
    leaf = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    L, R, S, C = leaf(g["logits"]), leaf(g["reg"]), leaf(g["std"]), leaf(g["ctr"])
    losses, extras = O.fcos_losses_labeled(L, R, S, C, locs(), g["boxes"], g["classes"])
    for k, ref in g["losses"].items():
        torch.testing.assert_close(losses[k].detach(), ref, rtol=1e-5, atol=1e-6)
    assert abs(extras["loss_denorm"] - g["loss_denorm"]) < 1e-4
    total = sum(losses[k] * (i + 1) for i, k in enumerate(g["losses"].keys()))
    total.backward()
    for name, leaves in (("logits", L), ("reg", R), ("std", S), ("ctr", C)):
        for a, b in zip(leaves, g["grads"][name]):
            torch.testing.assert_close(a.grad, b, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("method", ["mean", "sum", "weight_ctr_sum", "weight_ctr_mean"])
def test_fcos_losses_labeled_klloss(method):
    """KL_LOSS_TYPE "klloss" with the four LOC_FUN_ALL reductions, on the inputs of fcos_losses_labeled.pt."""
    g, ref = load("fcos_losses_labeled.pt"), load("fcos_losses_labeled_klloss.pt")[method]
    leaf = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    L, R, S, C = leaf(g["logits"]), leaf(g["reg"]), leaf(g["std"]), leaf(g["ctr"])
    losses, _ = O.fcos_losses_labeled(L, R, S, C, locs(), g["boxes"], g["classes"], kl_loss_type="klloss", loc_fun_all=method)
    for k, r in ref["losses"].items():
        torch.testing.assert_close(losses[k].detach(), r, rtol=1e-5, atol=1e-6)
    if "grads" in ref:
        sum(losses[k] * (i + 1) for i, k in enumerate(ref["losses"].keys())).backward()
        for name, leaves in (("reg", R), ("std", S)):
            for a, b in zip(leaves, ref["grads"][name]):
                torch.testing.assert_close(a.grad, b, rtol=1e-4, atol=1e-7)


def test_fcos_losses_labeled_no_positive():
    g = load("fcos_losses_labeled_empty.pt")
    empty = [torch.zeros(0, 4)] * 2
    cls = [torch.zeros(0, dtype=torch.long)] * 2
    losses, _ = O.fcos_losses_labeled(g["logits"], g["reg"], g["std"], g["ctr"], locs(), empty, cls)
    for k, ref in g["losses"].items():
        assert float(losses[k]) == float(ref) == 0.0


def test_fcos_losses_pseudo_and_grads():
    g = load("fcos_losses_pseudo.pt")
    leaf = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    L, R, S, C = leaf(g["logits"]), leaf(g["reg"]), leaf(g["std"]), leaf(g["ctr"])
    losses = O.fcos_losses_pseudo(L, R, S, C, locs(), g["cls_set"], g["reg_set"])
    for k, ref in g["losses"].items():
        torch.testing.assert_close(torch.as_tensor(losses[k]).detach().float(), ref, rtol=1e-5, atol=1e-6)
    total = sum(losses[k] * (i + 1) for i, k in enumerate(g["loss_order"]) if k.startswith("loss"))
    total.backward()
    for name, leaves in (("logits", L), ("reg", R), ("ctr", C)):
        for a, b in zip(leaves, g["grads"][name]):
            torch.testing.assert_close(a.grad, b, rtol=1e-4, atol=1e-7)
    for a in S:  # the shipped pseudo regression loss does not back-propagate into the std head
        assert a.grad is None or float(a.grad.abs().max()) == 0.0


def _as_sets(res):
    """Order inside a level is arbitrary in the reference (topk sorted=False) before NMS, but after NMS
    rows are in descending-score order; compare row-aligned after a canonical sort."""
    key = torch.stack([res["fpn_levels"].float(), res["locations"][:, 0], res["locations"][:, 1],
                       res["pred_classes"].float()], 1)
    order = sorted(range(len(key)), key=lambda i: tuple(key[i].tolist()))
    return {k: v[order] for k, v in res.items() if isinstance(v, torch.Tensor)}


@pytest.mark.parametrize("tag", ["small", "large"])
@pytest.mark.parametrize("method", ["cls", "cls_n_ctr", "cls_n_loc"])
def test_fcos_predict_proposals(tag, method):
    g = load(f"fcos_predict_{tag}.pt")
    out = O.fcos_predict_proposals(g["logits"], g["reg"], g["std"], g["ctr"], locs(), g["image_sizes"], method,
                                   device_type="cpu")  # fixtures were minted on cpu tensors
    for res, ref in zip(out, g["out"][method]):
        a, b = _as_sets(res), _as_sets(ref)
        assert len(a["scores"]) == len(b["scores"])
        for k in b:
            if b[k].dtype.is_floating_point:
                torch.testing.assert_close(a[k], b[k], rtol=1e-5, atol=1e-5)
            else:
                assert torch.equal(a[k], b[k]), k


def test_predict_ctr_method_is_rejected():
    g = load("fcos_predict_small.pt")
    with pytest.raises(ValueError):
        O.fcos_predict_proposals(g["logits"], g["reg"], g["std"], g["ctr"], locs(), g["image_sizes"], "ctr")


def test_nms_matches_torchvision():
    import torchvision

    gen = torch.Generator().manual_seed(3)
    for n in (1, 17, 300, 1500):
        xy = torch.rand(n, 2, generator=gen) * 200
        wh = torch.rand(n, 2, generator=gen) * 60 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=gen)
        idxs = torch.randint(0, 5, (n,), generator=gen)
        ref = torchvision.ops.nms(boxes, scores, 0.6)
        assert torch.equal(torch.from_numpy(O.nms_plain(boxes.numpy(), scores.numpy(), 0.6)), ref)
        refb = torchvision.ops.batched_nms(boxes, scores, idxs, 0.6)
        assert torch.equal(O.batched_nms(boxes, scores, idxs, 0.6, device_type="cpu"), refb)
        # the coordinate trick (what the reference gets on cuda for <= 5000 boxes) vs torchvision's own helper
        trick = torchvision.ops.boxes._batched_nms_coordinate_trick(boxes, scores, idxs, 0.6)
        assert torch.equal(O.batched_nms(boxes, scores, idxs, 0.6, device_type="cuda"), trick)


def test_pseudo_threshold():
    g = load("pseudo_threshold.pt")
    inst = g["inst"]
    empty = {k: v[:0] for k, v in inst.items()}
    for src, ref in zip([inst, empty], g["thresholding"]):
        out = O.threshold_bbox(src, 0.5)
        assert set(out) == set(ref)
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
    for src, ref in zip([inst, empty], g["thresholding_cls_ctr"]):
        out = O.threshold_cls_ctr_bbox(src, (0.5, 0.3))
        for k in ref:
            assert torch.equal(out[k], ref[k]), k


def test_ema_bit_exact():
    g = load("ema.pt")
    for keep in [0.0, 0.9996, 0.9999]:
        assert torch.equal(O.ema_update(g["student"], g["teacher"], keep), g[f"keep{keep}"])


def test_box2box_xyxy():
    g = load("box2box_xyxy.pt")
    torch.testing.assert_close(O.box2box_xyxy_get_deltas(g["src"], g["dst"]), g["deltas"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(O.box2box_xyxy_apply_deltas(g["rand_deltas"], g["src"]), g["applied"], rtol=1e-6, atol=1e-4)
    torch.testing.assert_close(O.box2box_xyxy_apply_deltas(g["deltas"], g["src"]), g["applied_roundtrip"], rtol=1e-6, atol=1e-4)


def test_rcnn_loss_pieces():
    g = load("rcnn_loss_pieces.pt")
    torch.testing.assert_close(O.rcnn_focal_loss(g["x"], g["y"]), g["focal_sum"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(O.matched_boxlist_iou(g["b1"], g["b2"]), g["matched_iou"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(O.rcnn_nl_loss(g["nl_mu"], g["nl_std"], g["nl_tgt"], g["nl_iou"]).reshape(()),
                               g["nl_loss"].reshape(()), rtol=1e-5, atol=1e-5)
