"""The reference's module-level plugin surface on the device: every class / builder the reference
registers is resolved through its registry, called with the reference's arguments (NCHW feature maps, list[Instances]) and
checked against the oracle. Floating-point tolerances are stated at each check; Box2BoxXYXYTransform is bit-exact."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.float().cpu().double(), b.float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-12))


def cosr(a, b):
    a, b = a.float().cpu().double().flatten(), b.float().cpu().double().flatten()
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30)), float(a.norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def fcos_model():
    from test_model_gpu import diversify
    from util_cfg import fcos_cfg
    import ubteacher.modeling  # noqa: F401
    from ubteacher.d2compat.registry import META_ARCH_REGISTRY
    m = META_ARCH_REGISTRY.get("OneStageDetector")(fcos_cfg())
    diversify(m)
    return m


@pytest.fixture(scope="module")
def rcnn_model():
    from test_rcnn_model_gpu import diversify
    from util_cfg import rcnn_cfg
    import ubteacher.modeling  # noqa: F401
    from ubteacher.d2compat.registry import META_ARCH_REGISTRY
    m = META_ARCH_REGISTRY.get("TwoStagePseudoLabGeneralizedRCNN")(rcnn_cfg())
    diversify(m)
    m.train()
    return m


def test_registries_resolve_the_reference_names():
    import ubteacher.modeling  # noqa: F401
    from ubteacher.d2compat.registry import BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, ROI_HEADS_REGISTRY
    from ubteacher.modeling.backbone.fpn import LastLevelP6P7, build_fcos_resnet_fpn_backbone
    from ubteacher.modeling.box_regression import Box2BoxXYXYTransform  # noqa: F401
    from ubteacher.modeling.fcos.fcos import FCOS, FCOSHead, Scale  # noqa: F401
    from ubteacher.modeling.proposal_generator.rpn import PseudoLabRPN
    from ubteacher.modeling.roi_heads.fast_rcnn import FastRCNNFocaltLossBoundaryVarOutputLayers  # noqa: F401
    from ubteacher.modeling.roi_heads.roi_heads import StandardROIHeadsPseudoLab
    assert PROPOSAL_GENERATOR_REGISTRY.get("FCOS") is FCOS and PROPOSAL_GENERATOR_REGISTRY.get("PseudoLabRPN") is PseudoLabRPN
    assert ROI_HEADS_REGISTRY.get("StandardROIHeadsPseudoLab") is StandardROIHeadsPseudoLab
    assert BACKBONE_REGISTRY.get("build_fcos_resnet_fpn_backbone") is build_fcos_resnet_fpn_backbone
    for n in ("OneStageDetector", "PseudoProposalNetwork", "TwoStagePseudoLabGeneralizedRCNN"):
        META_ARCH_REGISTRY.get(n)
    assert LastLevelP6P7.__name__ == "LastLevelP6P7"


def test_meta_arch_exposes_reference_submodules(fcos_model, rcnn_model):
    from ubteacher.modeling.fcos.fcos import FCOS, FCOSHead
    from ubteacher.modeling.roi_heads.fast_rcnn import FastRCNNFocaltLossBoundaryVarOutputLayers
    m = fcos_model
    assert isinstance(m.proposal_generator, FCOS) and isinstance(m.proposal_generator.fcos_head, FCOSHead)
    assert m.proposal_generator.fcos_outputs is m.fcos_outputs
    full = m.state_dict()
    head = m.proposal_generator.fcos_head.state_dict()
    # the sub-tree carries the reference's keys (fcos.py:221-336) and aliases the same storage
    for k in ("cls_tower.0.weight", "cls_tower.1.bias", "bbox_tower.9.weight", "cls_logits.bias", "bbox_pred.weight",
              "bbox_pred_std.weight", "ctrness.bias", "scales.4.scale"):
        assert head[k].data_ptr() == full["proposal_generator.fcos_head." + k].data_ptr(), k
    assert set(m.backbone.top_block.state_dict()) == {"p6.weight", "p6.bias", "p7.weight", "p7.bias"}
    assert set(m.backbone.output_shape()) == {"p3", "p4", "p5", "p6", "p7"}
    r = rcnn_model
    assert isinstance(r.roi_heads.box_predictor, FastRCNNFocaltLossBoundaryVarOutputLayers)
    assert set(r.roi_heads.box_predictor.state_dict()) == {f"{a}.{b}" for a in ("cls_score", "bbox_pred", "bbox_pred_std") for b in ("weight", "bias")}
    assert r.proposal_generator.state_dict()["rpn_head.conv.weight"].shape == (256, 256, 3, 3)
    assert len(list(r.roi_heads.box_predictor.parameters())) == 6


def _pyramid(N, Hp, Wp, strides, seed, c=256):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(N, c, (Hp // s), (Wp // s), generator=g).relu() * 0.5) for s in strides]


def test_fcos_head_forward_backward_vs_oracle(fcos_model):
    """FCOSHead(x) (fcos.py:338-376): 6-tuple of per-level lists from NCHW maps, and autograd through it."""
    from oracle import ut2_model as M
    head = fcos_model.proposal_generator.fcos_head
    hw = [(20, 28), (10, 14), (5, 7), (3, 4), (2, 2)]
    g = torch.Generator().manual_seed(5)
    xs = [(torch.randn(2, 256, h, w, generator=g).relu() * 0.5) for h, w in hw]
    xd = [x.cuda().requires_grad_(True) for x in xs]
    eng = fcos_model.engine
    eng.arena.grad.zero_()
    logits, reg, std, ctr, top_feats, towers = head(xd, None, True)
    assert len(logits) == 5 and logits[0].shape == (2, 80, 20, 28) and reg[1].shape == (2, 68, 10, 14)
    assert std[2].shape == (2, 4, 5, 7) and ctr[4].shape == (2, 1, 2, 2) and top_feats == [] and towers[0].shape == (2, 256, 20, 28)
    sd = {k: v.detach().cpu().clone() for k, v in fcos_model.state_dict().items()}
    hd = "proposal_generator.fcos_head."
    tk = [k for k in sd if k.startswith(hd)]
    params = {k: sd[k].clone().requires_grad_(True) for k in tk}
    sdp = dict(sd)
    sdp.update(params)
    xr = [x.clone().requires_grad_(True) for x in xs]
    with M.bf16_points():
        rl, rr, rs, rc = M.fcos_head(sdp, xr)
    bias = sd[hd + "cls_logits.bias"].view(1, -1, 1, 1)
    gs = []
    tot_d, tot_r = 0, 0
    for l in range(5):      # five bf16 conv + GroupNorm layers: 1.5 %
        assert rel(logits[l].detach().cpu() - bias, rl[l].detach() - bias) < 0.015, l
        assert rel(reg[l], rr[l]) < 0.015 and rel(std[l], rs[l]) < 0.015 and rel(ctr[l], rc[l]) < 0.02, l
        w = [torch.randn(t.shape, generator=g) * 0.05 for t in (rl[l], rr[l], rs[l], rc[l])]
        gs.append(w)
        tot_d = tot_d + sum((a * b.cuda()).sum() for a, b in zip((logits[l], reg[l], std[l], ctr[l]), w))
        tot_r = tot_r + sum((a * b).sum() for a, b in zip((rl[l], rr[l], rs[l], rc[l]), w))
    tot_d.backward()
    tot_r.backward()
    G = eng.arena.gviews
    for k in tk:
        if params[k].grad is None:
            continue
        cos, ratio = cosr(G[k], params[k].grad)
        assert cos > 0.995 and abs(ratio - 1) < 0.03, (k, cos, ratio)
    for l in range(5):
        cos, ratio = cosr(xd[l].grad, xr[l].grad)
        assert cos > 0.995 and abs(ratio - 1) < 0.03, (l, cos, ratio)
    # parameters() of the view expose the arena gradients as .grad
    p = dict(head.named_parameters())
    assert p["cls_logits.weight"].grad.data_ptr() == G[hd + "cls_logits.weight"].data_ptr()
    eng.arena.grad.zero_()


def test_fcos_generator_and_backbone_views_vs_meta_arch(fcos_model):
    """backbone(images) -> FCOS(images, features, gt, ...) composes to exactly what OneStageDetector computes
    (one_stage_detector.py:165-214), in training (losses, raw output) and eval (proposals)."""
    from test_model_gpu import make_batch
    from oracle import ut2_model as M
    m = fcos_model
    batch = make_batch(2, [(160, 224), (128, 192)], 12)
    m.train()
    images = m.preprocess_image(batch)
    assert images.image_sizes == [(160, 224), (128, 192)] and images.tensor.shape == (2, 3, 160, 224)
    feats = m.backbone(images)
    assert list(feats) == ["p3", "p4", "p5", "p6", "p7"] and feats["p3"].shape == (2, 256, 20, 28)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    ref = M.backbone(sd, M.preprocess(sd, [b["image"] for b in batch])[0])
    for k, r in zip(feats, ref):
        assert rel(feats[k], r) < 0.03, k            # bf16 chain of ~55 convs vs fp32
    gt = [b["instances"] for b in batch]
    results, losses, raw = m.proposal_generator(images, feats, gt, output_raw=True, branch="labeled")
    want, _ = m.forward_train(batch, "labeled")
    for k in want:                                    # same kernels on the same pyramid
        torch.testing.assert_close(losses[k].detach(), want[k], rtol=1e-5, atol=1e-6)
    assert raw["logits_pred"][0].shape == (2, 80, 20, 28) and raw["locations"][1].shape == (10 * 14, 2)
    assert raw["image_sizes"] == [(160, 224), (128, 192)]
    # the losses are differentiable w.r.t. the parameters (arena gradients) through the view
    m.engine.arena.grad.zero_()
    sum(losses.values()).backward()
    assert float(m.engine.arena.gviews["proposal_generator.fcos_head.cls_logits.weight"].abs().sum()) > 0
    m.engine.arena.grad.zero_()
    m.eval()
    with torch.no_grad():
        res, _ = m.proposal_generator(images, m.backbone(images), nms_method="cls_n_ctr")
        dets, _ = m(batch, output_raw=True)          # "output raw will not rescale" (one_stage_detector.py:131-133)
    for i, a in enumerate(res):
        n = int(dets["count"][i])
        assert len(a) == n and torch.equal(a.pred_boxes.tensor, dets["pred_boxes"][i, :n]) and torch.equal(a.scores, dets["scores"][i, :n])
    m.train()
    with pytest.raises(ValueError):
        m.proposal_generator(images, feats, gt, branch="bogus")


def test_last_level_p6p7_view(fcos_model):
    from oracle import ut2_model as M
    m = fcos_model
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    p5 = torch.randn(2, 256, 10, 14, generator=torch.Generator().manual_seed(2))
    p6, p7 = m.backbone.top_block(p5.cuda())
    r6 = F.conv2d(p5, sd["backbone.top_block.p6.weight"], sd["backbone.top_block.p6.bias"], 2, 1)
    r7 = F.conv2d(F.relu(r6), sd["backbone.top_block.p7.weight"], sd["backbone.top_block.p7.bias"], 2, 1)
    assert rel(p6, r6) < 0.01 and rel(p7, r7) < 0.015 and p7.shape == (2, 256, 3, 4)


def test_box2box_xyxy_transform_bit_exact():
    from ubteacher.modeling.box_regression import Box2BoxXYXYTransform
    g = torch.load(os.path.join(GOLD, "box2box_xyxy.pt"))
    t = Box2BoxXYXYTransform(weights=(10.0, 10.0, 5.0, 5.0))
    assert torch.equal(t.get_deltas(g["src"].cuda(), g["dst"].cuda()).cpu(), g["deltas"])
    assert torch.equal(t.apply_deltas(g["rand_deltas"].cuda(), g["src"].cuda()).cpu(), g["applied"])
    assert torch.equal(t.apply_deltas(g["deltas"].cuda(), g["src"].cuda()).cpu(), g["applied_roundtrip"])


def _roi_case(seed, n_img=2, pseudo=False):
    from oracle import ut2_rcnn_oracle as R
    from ubteacher.d2compat.structures import Boxes, Instances
    g = torch.Generator().manual_seed(seed)
    props, flat = [], {"pb": [], "gb": [], "gc": [], "std": []}
    for i in range(n_img):
        n = 40 + 13 * i
        xy = torch.rand(n, 2, generator=g) * 100
        wh = torch.rand(n, 2, generator=g) * 60 + 8
        pb = torch.cat([xy, xy + wh], 1)
        gb = pb + torch.randn(n, 4, generator=g) * 3
        gc = torch.randint(0, 81, (n,), generator=g)
        inst = Instances((160, 224))
        inst.proposal_boxes, inst.gt_boxes, inst.gt_classes = Boxes(pb), Boxes(gb), gc
        st = torch.randn(n, 4, generator=g) * 2
        if pseudo:
            inst.gt_loc_std = st
        props.append(inst)
        for k, v in zip(("pb", "gb", "gc", "std"), (pb, gb, gc, st)):
            flat[k].append(v)
    return props, {k: torch.cat(v) for k, v in flat.items()}


@pytest.mark.parametrize("branch", ["supervised", "unsup_data_train"])
def test_output_layers_forward_losses_inference_vs_oracle(rcnn_model, branch):
    """FastRCNNFocaltLossBoundaryVarOutputLayers.forward / losses / inference (fast_rcnn.py:812-829, :831-923, :1094-1125)."""
    from oracle import ut2_rcnn_oracle as R
    layers = rcnn_model.roi_heads.box_predictor
    eng = rcnn_model.engine
    sd = {k: v.detach().cpu().clone() for k, v in layers.state_dict().items()}
    props, flat = _roi_case(3, pseudo=branch == "unsup_data_train")
    Rn = flat["pb"].shape[0]
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(Rn, 1024, generator=g).relu() * 0.7).bfloat16().float()
    xd = x.cuda().requires_grad_(True)
    eng.arena.grad.zero_()
    sc, dl, st = layers(xd)
    assert sc.shape == (Rn, 81) and dl.shape == (Rn, 4) and st.shape == (Rn, 4)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    q = lambda w: w + (w.bfloat16().float() - w).detach()          # bf16 operand, straight-through gradient
    rsc = F.linear(xr, q(params["cls_score.weight"]), params["cls_score.bias"])
    rdl = F.linear(xr, q(params["bbox_pred.weight"]), params["bbox_pred.bias"])
    rst = F.linear(xr, q(params["bbox_pred_std.weight"]), params["bbox_pred_std.bias"])
    for a, b in ((sc, rsc), (dl, rdl), (st, rst)):          # bf16 output rounding
        torch.testing.assert_close(a.detach().cpu(), b.detach(), rtol=1 / 128, atol=2e-3)
    losses = layers.losses((sc, dl, st), [p.to("cuda") if hasattr(p, "to") else p for p in props], branch)
    ref = R.fast_rcnn_losses(rsc, rdl, rst, flat["pb"], flat["gb"], flat["gc"], branch, flat["std"] if branch == "unsup_data_train" else None)
    for k in ("loss_cls", "loss_box_reg"):                  # fp32 loss arithmetic on bf16-rounded predictions
        torch.testing.assert_close(losses[k].detach().cpu(), ref[k].detach(), rtol=1e-2, atol=1e-4)
    (losses["loss_cls"] * 1.0 + losses["loss_box_reg"] * 0.7).backward()
    (ref["loss_cls"] * 1.0 + ref["loss_box_reg"] * 0.7).backward()
    G = eng.arena.gviews
    for k in params:
        if params[k].grad is None or float(params[k].grad.abs().max()) == 0:
            continue
        cos, ratio = cosr(G["roi_heads.box_predictor." + k], params[k].grad)
        assert cos > 0.995 and abs(ratio - 1) < 0.03, (k, cos, ratio)
    cos, ratio = cosr(xd.grad, xr.grad)
    assert cos > 0.995 and abs(ratio - 1) < 0.03, ("dx", cos, ratio)
    eng.arena.grad.zero_()
    if branch == "supervised":
        inst, kept = layers.inference((sc.detach(), dl.detach(), st.detach()), props)
        off = 0
        for i, p in enumerate(props):
            n = len(p)
            o = R.fast_rcnn_inference(sc[off:off + n].detach().cpu(), dl[off:off + n].detach().cpu(), st[off:off + n].detach().cpu(),
                                      flat["pb"][off:off + n], (160, 224))
            off += n
            # identical inputs (the device's own bf16 predictions): the kept set is exact, order included
            assert torch.equal(kept[i].cpu(), o["roi_index"]) and torch.equal(inst[i].pred_classes.cpu(), o["pred_classes"])
            torch.testing.assert_close(inst[i].pred_boxes.tensor.cpu(), o["pred_boxes"], rtol=0, atol=1e-4)
            torch.testing.assert_close(inst[i].scores.cpu(), o["scores"], rtol=1e-5, atol=1e-6)
            assert torch.equal(inst[i].pred_boxes_std.cpu(), o["pred_boxes_std"])


def test_rpn_and_roi_heads_views_vs_meta_arch(rcnn_model):
    """backbone(images) -> PseudoLabRPN(images, features, gt) -> StandardROIHeadsPseudoLab(images, features, proposals, gt,
    branch=...) composes to what TwoStagePseudoLabGeneralizedRCNN computes (rcnn.py:26-40), with the same sampling draws."""
    from test_rcnn_model_gpu import _inject_keys, make_batch
    m = rcnn_model
    eng = m.engine
    batch = make_batch(2, [(160, 224), (128, 192)], 31)
    geom, _ = eng.level_geom(160, 224)
    _inject_keys(m, 2, geom.A, 5)
    try:
        want, pending = m.forward_train(batch, "supervised")
        images = m.preprocess_image(batch)
        feats = m.backbone(images)
        assert list(feats) == ["p2", "p3", "p4", "p5", "p6"] and feats["p2"].shape == (2, 256, 40, 56)
        gt = [b["instances"] for b in batch]
        props, rl = m.proposal_generator(images, feats, gt)
        for k in ("loss_rpn_cls", "loss_rpn_loc"):
            torch.testing.assert_close(rl[k].detach(), want[k], rtol=1e-5, atol=1e-6)
        assert len(props) == 2 and props[0].has("proposal_boxes") and props[0].has("objectness_logits") and len(props[0]) <= 1000
        sampled, hl = m.roi_heads(images, feats, props, gt, True, "supervised")
        for k in ("loss_cls", "loss_box_reg"):
            torch.testing.assert_close(hl[k].detach(), want[k], rtol=1e-5, atol=1e-6)
        assert all(len(s) <= 512 and s.has("gt_classes") and s.has("gt_boxes") for s in sampled)
        # autograd through the views reaches the parameters (arena gradients)
        eng.arena.grad.zero_()
        (rl["loss_rpn_cls"] + rl["loss_rpn_loc"] + hl["loss_cls"] + hl["loss_box_reg"]).backward()
        G = eng.arena.gviews
        assert float(G["proposal_generator.rpn_head.conv.weight"].abs().sum()) > 0
        assert float(G["roi_heads.box_head.fc1.weight"].abs().sum()) > 0
        eng.arena.grad.zero_()
        # inference through the views == the unsup_data_weak branch of the meta-architecture
        with torch.no_grad():
            pred_inst, predictions = m.roi_heads(images, feats, props, None, False, "unsup_data_weak")
            _, _, dets, _ = m(batch, branch="unsup_data_weak")
        cnt = dets["count"].cpu().tolist()
        for i, n in enumerate(cnt):
            assert len(pred_inst[i]) == n and torch.equal(pred_inst[i].pred_boxes.tensor, dets["pred_boxes"][i, :n])
            assert pred_inst[i].has("pred_boxes_std")
        assert predictions[0].shape[1] == 81 and predictions[1].shape[1] == 4
    finally:
        eng.debug_keys = None
