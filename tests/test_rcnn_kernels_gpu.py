"""GPU parity of the Faster R-CNN kernels (csrc/{nms,rpn,roi}.cu) against the CPU oracle (oracle/ut2_rcnn_oracle.py),
through the C ABI. Index / label / keep-list outputs: bit-exact. Floating point: tolerance stated at each check."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LEVEL_HW = [(32, 40), (16, 20), (8, 10), (4, 5), (2, 3)]
STRIDES = [4, 8, 16, 32, 64]
IMG_HW = (128, 160)


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def rb(x):
    return x.bfloat16().float()


def rand_boxes(g, n, H, W, min_size=4.0):
    x1 = torch.rand(n, generator=g) * (W - min_size - 1)
    y1 = torch.rand(n, generator=g) * (H - min_size - 1)
    w = min_size + torch.rand(n, generator=g) * (W - x1 - min_size)
    h = min_size + torch.rand(n, generator=g) * (H - y1 - min_size)
    return torch.stack([x1, y1, x1 + w, y1 + h], dim=1)


def pack_gt(boxes, G=16, classes=None, scores=None, std=None):
    N = len(boxes)
    b = torch.zeros(N, G, 4)
    c = torch.zeros(N, G, dtype=torch.int64)
    s = torch.zeros(N, G)
    d = torch.zeros(N, G, 4)
    cnt = torch.zeros(N, dtype=torch.int32)
    for i in range(N):
        n = len(boxes[i])
        cnt[i] = n
        b[i, :n] = boxes[i]
        if classes is not None:
            c[i, :n] = classes[i]
        if scores is not None:
            s[i, :n] = scores[i]
        if std is not None:
            d[i, :n] = std[i]
    return b.cuda(), c.cuda(), s.cuda(), d.cuda(), cnt.cuda()


def pack_rpn(logits, deltas, level_hw):
    """[N, A] logits + [N, A, 4] deltas (anchor order (h, w, a) per level) -> level-major [P, 16] bf16."""
    N = logits.shape[0]
    rows, off = [], 0
    for (h, w) in level_hw:
        n = h * w * 3
        lg = logits[:, off:off + n].reshape(N, h * w, 3)
        dl = deltas[:, off:off + n].reshape(N, h * w, 12)
        rows.append(torch.cat([lg, dl, torch.zeros(N, h * w, 1)], dim=2).reshape(-1, 16))
        off += n
    return torch.cat(rows).bfloat16().cuda()


def unpack_rpn(t, N, level_hw):
    """inverse of pack_rpn for a [P,16] tensor -> ([N,A], [N,A,4]) fp32 CPU"""
    t = t.float().cpu()
    lgs, dls, off = [], [], 0
    for (h, w) in level_hw:
        blk = t[off:off + N * h * w].reshape(N, h * w, 16)
        lgs.append(blk[:, :, :3].reshape(N, -1))
        dls.append(blk[:, :, 3:15].reshape(N, -1, 4))
        off += N * h * w
    return torch.cat(lgs, 1), torch.cat(dls, 1)


def rgeom():
    from ubteacher import ops_rcnn as R
    return R.RpnGeom(LEVEL_HW, STRIDES)


# --------------------------------------------------------------------------------------- NMS
@pytest.mark.parametrize("n,ncls,thr,trick", [(700, 5, 0.7, True), (3000, 5, 0.7, True), (6000, 4, 0.5, False),
                                              (9000, 5, 0.7, False), (40, 80, 0.5, True)])
def test_nms_batched_bit_exact(n, ncls, thr, trick):
    from oracle import ut2_oracle as O
    from ubteacher import ops_rcnn as R
    g = torch.Generator().manual_seed(n)
    N = 2
    M = n + 37
    boxes = torch.zeros(N, M, 4)
    scores = torch.full((N, M), float("-inf"))
    cls = torch.zeros(N, M, dtype=torch.int32)
    cnt = torch.tensor([n, n - 11], dtype=torch.int32)
    for i in range(N):
        k = int(cnt[i])
        xy = torch.rand(k, 2, generator=g) * 300
        wh = torch.rand(k, 2, generator=g) * 80 + 2
        boxes[i, :k] = torch.cat([xy, xy + wh], 1)
        scores[i, :k] = rb(torch.randn(k, generator=g))          # bf16 rounding => plenty of score ties
        cls[i, :k] = torch.randint(0, ncls, (k,), generator=g, dtype=torch.int32)
        scores[i, 5] = float("-inf")                             # a dropped (non-finite) candidate
    assert (4 * n <= 20000) == trick
    keep, kcnt = R.nms_batched(boxes.cuda(), scores.cuda(), cls.cuda(), cnt.cuda(), thr, 1000)
    torch.cuda.synchronize()
    for i in range(N):
        k = int(cnt[i])
        valid = torch.isfinite(scores[i, :k]).nonzero().squeeze(1)
        ref = O.batched_nms(boxes[i, valid], scores[i, valid], cls[i, valid].long(), thr, "cuda")[:1000]
        ref = valid[ref]
        got = keep[i, :int(kcnt[i])].cpu().long()
        assert torch.equal(got, ref), (i, len(got), len(ref))
        # idempotence at full size: NMS of the survivors keeps every one of them
        kb = boxes[i, got].unsqueeze(0).cuda().contiguous()
        ks = scores[i, got].unsqueeze(0).cuda().contiguous()
        kc = cls[i, got].unsqueeze(0).cuda().contiguous()
        k2, c2 = R.nms_batched(kb, ks, kc, torch.tensor([len(got)], dtype=torch.int32).cuda(), thr, 1000)
        assert int(c2[0]) == len(got) and torch.equal(k2[0, :len(got)].cpu().long(), torch.arange(len(got)))


# --------------------------------------------------------------------------------------- RPN labels / losses
def test_rpn_label_anchors_golden_bit_exact():
    from oracle import ut2_rcnn_oracle as OR
    from ubteacher import ops_rcnn as R
    g = load("rcnn_rpn_losses.pt")
    gm = rgeom()
    N = len(g["gt_boxes"])
    gb, _, gs, _, cnt = pack_gt(g["gt_boxes"], scores=g["gt_scores"])
    keys = torch.stack(g["keys"]).to(torch.int64).cuda().to(torch.int32)       # same low 32 bits
    labels, matched = R.rpn_label_anchors(gm, N, gb, cnt, keys=keys)
    torch.cuda.synchronize()
    for i in range(N):
        assert torch.equal(labels[i].cpu(), g["labels"][i]), i
        if len(g["gt_boxes"][i]):
            mb = g["gt_boxes"][i][matched[i].cpu().long()]
            assert torch.equal(mb, g["matched_boxes"][i])
    # hashed keys (product path): class counts follow [D2] subsample_labels
    labels2, _ = R.rpn_label_anchors(gm, N, gb, cnt, keys=None, seed=123)
    for i in range(N):
        l2 = labels2[i].cpu()
        assert int((l2 >= 0).sum()) == 256 and int((l2 == 1).sum()) == int((g["labels"][i] == 1).sum())
        anchors = torch.cat(OR.generate_anchors(LEVEL_HW))
        iou = OR.pairwise_iou(g["gt_boxes"][i], anchors)
        _, pre = OR.matcher(iou, (0.3, 0.7), (0, -1, 1), True)
        assert ((l2 == 1) <= (pre == 1)).all() and ((l2 == 0) <= (pre == 0)).all()
    # device-resident draw counter (CUDA-graph replay): word 0 = the plain seed, another word = another draw
    word = torch.zeros(1, dtype=torch.int32, device="cuda")
    labels3, _ = R.rpn_label_anchors(gm, N, gb, cnt, keys=None, seed=123, seed_dev=word)
    assert torch.equal(labels3, labels2)
    word.fill_(5)
    labels4, _ = R.rpn_label_anchors(gm, N, gb, cnt, keys=None, seed=123, seed_dev=word)
    assert not torch.equal(labels4, labels2)
    for i in range(N):
        assert int((labels4[i] >= 0).sum()) == 256 and int((labels4[i] == 1).sum()) == int((labels2[i] == 1).sum())


@pytest.mark.parametrize("tag", ["pseudo", "sup"])
def test_rpn_losses_fwd_bwd(tag):
    from oracle import ut2_rcnn_oracle as OR
    from ubteacher import ops_rcnn as R
    g = load("rcnn_rpn_losses.pt")
    gm = rgeom()
    N = len(g["gt_boxes"])
    gb, _, gs, _, cnt = pack_gt(g["gt_boxes"], scores=g["gt_scores"])
    keys = torch.stack(g["keys"]).to(torch.int64).cuda().to(torch.int32)
    labels, matched = R.rpn_label_anchors(gm, N, gb, cnt, keys=keys)
    rpn_out = pack_rpn(g["logits"], g["deltas"], LEVEL_HW)
    lg, dl = unpack_rpn(rpn_out, N, LEVEL_HW)                  # the bf16-rounded values the device sees
    lg.requires_grad_(True)
    dl.requires_grad_(True)
    anchors = torch.cat(OR.generate_anchors(LEVEL_HW))
    ref = OR.rpn_losses(anchors, lg, dl, g["labels"], g["matched_boxes"], g["confids"] if tag == "pseudo" else None)
    (ref["loss_rpn_cls"] * 1.7 + ref["loss_rpn_loc"] * 0.6).backward()
    sc = gs if tag == "pseudo" else None
    losses = R.rpn_loss_fwd(gm, N, rpn_out, labels, matched, gb, sc, cnt)
    gout = torch.tensor([1.7, 0.6], device="cuda")
    drpn = R.rpn_loss_bwd(gm, N, rpn_out, labels, matched, gb, sc, cnt, gout)
    torch.cuda.synchronize()
    got = losses.cpu()
    # fp32 sums over 768 anchors: relative 2e-5
    assert abs(got[0] - ref["loss_rpn_cls"]) <= 2e-5 * abs(ref["loss_rpn_cls"]) + 1e-7
    assert abs(got[1] - ref["loss_rpn_loc"]) <= 2e-5 * abs(ref["loss_rpn_loc"]) + 1e-7
    dlg, ddl = unpack_rpn(drpn, N, LEVEL_HW)
    # gradients are stored in bf16: relative 2^-8
    assert torch.allclose(dlg, lg.grad, rtol=1 / 128, atol=1e-7)
    assert torch.allclose(ddl, dl.grad, rtol=1 / 128, atol=1e-7)


# --------------------------------------------------------------------------------------- RPN proposals
@pytest.mark.parametrize("pre,post", [(200, 50), (2000, 1000)])
def test_rpn_proposals(pre, post):
    from oracle import ut2_rcnn_oracle as OR
    from ubteacher import ops_rcnn as R
    g = torch.Generator().manual_seed(pre)
    gm = rgeom()
    N = 3
    A = gm.A
    logits = torch.randn(N, A, generator=g) * 2
    deltas = torch.randn(N, A, 4, generator=g) * 0.4
    rpn_out = pack_rpn(logits, deltas, LEVEL_HW)
    lg, dl = unpack_rpn(rpn_out, N, LEVEL_HW)
    sizes = [(128, 160), (120, 150), (96, 160)]
    image_hw = torch.tensor(sizes, dtype=torch.float32).cuda()
    c = R.rpn_select_decode(gm, N, rpn_out, image_hw, pre)
    out = R.rpn_predict_proposals(gm, N, rpn_out, image_hw, pre, post, 0.7)
    torch.cuda.synchronize()
    anchors = OR.generate_anchors(LEVEL_HW)
    offs = [0]
    for a in anchors:
        offs.append(offs[-1] + a.shape[0])
    split = lambda t: [t[:, offs[i]:offs[i + 1]] for i in range(len(anchors))]
    ref = OR.rpn_select_decode(anchors, split(lg), split(dl), sizes, pre)
    for i in range(N):
        rbx, rs, rl, rvalid = ref[i]
        canon = c["canon"][i].cpu().long()
        order = torch.argsort(canon)
        # oracle candidates -> canonical anchor index
        rcanon = []
        for l, (lgl) in enumerate(split(lg)):
            k = min(lgl.shape[1], pre)
            _, idx = torch.sort(lgl[i], descending=True, stable=True)
            rcanon.append(idx[:k] + offs[l])
        rcanon = torch.cat(rcanon)
        rorder = torch.argsort(rcanon)
        assert torch.equal(canon[order], rcanon[rorder])                       # selected set: exact
        kb = c["boxes"][i].cpu()[order]
        ks = c["scores"][i].cpu()[order]
        assert torch.equal(torch.isfinite(ks), rvalid[rorder])
        assert torch.allclose(kb, rbx[rorder], rtol=1e-5, atol=2e-3)           # expf: CUDA vs CPU libm
        fin = torch.isfinite(ks)
        assert torch.equal(ks[fin], rs[rorder][fin])
        # NMS half: exact on the device-decoded candidates
        cb, cs, cl = c["boxes"][i].cpu(), c["scores"][i].cpu(), c["levels"][i].cpu().long()
        valid = torch.isfinite(cs)
        # oracle visits in (score desc, position) order; make position == canonical order
        pos = torch.argsort(canon)
        cbp, csp, clp, vp = cb[pos], cs[pos], cl[pos], valid[pos]
        keep = OR.rpn_nms([(cbp, csp, clp, vp)], 0.7, post)[0]
        n = int(out["count"][i])
        assert n == len(keep["objectness_logits"])
        assert torch.equal(out["proposal_boxes"][i, :n].cpu(), keep["proposal_boxes"])
        assert torch.equal(out["objectness_logits"][i, :n].cpu(), keep["objectness_logits"])


# --------------------------------------------------------------------------------------- ROI sampling
@pytest.mark.parametrize("mode", ["pseudo", "sup"])
def test_roi_sample_golden_bit_exact(mode):
    from ubteacher import ops_rcnn as R
    g = load("rcnn_roi_sampling.pt")
    N = len(g["proposals"])
    P = g["proposals"][0].shape[0]
    props = torch.zeros(N, P + 20, 4)
    for i in range(N):
        props[i, :P] = g["proposals"][i]
    pcnt = torch.full((N,), P, dtype=torch.int32)
    gb, gc, gs, gd, cnt = pack_gt(g["gt_boxes"], classes=g["gt_classes"], scores=g["gt_scores"], std=g["gt_std"])
    keys = torch.zeros(N, P + 16, dtype=torch.int64)
    for i in range(N):
        keys[i, :len(g["keys"][i])] = g["keys"][i]
    keys = keys.cuda().to(torch.int32)
    out = R.roi_sample(props.cuda(), pcnt.cuda(), gb, gc, cnt, gs if mode == "pseudo" else None,
                       gd if mode == "pseudo" else None, keys=keys)
    torch.cuda.synchronize()
    for i in range(N):
        ref = g[mode][i]
        n = int(out["count"][i])
        assert n == len(ref["gt_classes"])
        assert torch.equal(out["proposal_boxes"][i, :n].cpu(), ref["proposal_boxes"])
        assert torch.equal(out["gt_classes"][i, :n].cpu(), ref["gt_classes"])
        assert torch.equal(out["gt_boxes"][i, :n].cpu(), ref["gt_boxes"])
        if mode == "pseudo":
            assert torch.equal(out["gt_confid"][i, :n].cpu(), ref["gt_confid"].float())
            assert torch.equal(out["gt_loc_std"][i, :n].cpu(), ref["gt_loc_std"])
        assert (out["gt_classes"][i, n:].cpu() == -1).all()


# --------------------------------------------------------------------------------------- ROIAlign
def test_roi_align_fwd_bwd():
    from oracle import ut2_rcnn_oracle as OR
    from ubteacher import ops_rcnn as R
    g = torch.Generator().manual_seed(11)
    N, C = 2, 64
    hw = [(64, 80), (32, 40), (16, 20), (8, 10)]
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    feats = [rb(torch.randn(N, C, h, w, generator=g)) for h, w in hw]
    Rcap = 40
    rois = torch.zeros(N, Rcap, 4)
    cnt = torch.tensor([40, 33], dtype=torch.int32)
    for i in range(N):
        b = rand_boxes(g, Rcap, 256, 320, 2.0)
        b[:8] = b[:8] * 0.2 + 5                                    # small boxes -> level 0
        b[8] = torch.tensor([-20.0, -10.0, 400.0, 300.0])          # out of bounds + largest level
        b[9] = torch.tensor([-500.0, 100.0, 800.0, 104.0])         # 47 samples per bin along x: the per-sample path of the backward
        b[10] = torch.tensor([30.0, 20.0, 33.0, 140.0])            # thin and tall
        rois[i] = b
    rg = R.RoiGeom(hw, scales)
    nhwc = [f.permute(0, 2, 3, 1).contiguous().bfloat16().cuda() for f in feats]
    out = R.roi_align_fwd(rg, nhwc, rois.cuda(), cnt.cuda())
    torch.cuda.synchronize()
    fr = [f.clone().requires_grad_(True) for f in feats]
    boxes = [rois[i, :int(cnt[i])] for i in range(N)]
    ref = OR.roi_pool(fr, boxes)                                     # [R, C, 7, 7]
    got = out.float().cpu().reshape(N, Rcap, 7, 7, C).permute(0, 1, 4, 2, 3)
    off = 0
    for i in range(N):
        n = int(cnt[i])
        # bf16 output of an fp32 bilinear average: 2^-8 relative
        assert torch.allclose(got[i, :n], ref[off:off + n].detach(), rtol=1 / 128, atol=2e-3), i
        assert (got[i, n:] == 0).all()
        off += n
    dout = rb(torch.randn(N, Rcap, 7, 7, C, generator=g))
    dref = torch.cat([dout[i, :int(cnt[i])] for i in range(N)]).permute(0, 3, 1, 2)
    ref.backward(dref)
    dfe = [torch.zeros(N, h, w, C, dtype=torch.float32, device="cuda") for h, w in hw]
    R.roi_align_bwd(rg, dfe, rois.cuda(), cnt.cuda(), dout.reshape(N * Rcap, 7, 7, C).bfloat16().cuda())
    torch.cuda.synchronize()
    for l in range(4):
        a, b = dfe[l].cpu().permute(0, 3, 1, 2), (fr[l].grad if fr[l].grad is not None else torch.zeros_like(fr[l]))
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4), (l, (a - b).abs().max())    # fp32 atomics, order differs


# --------------------------------------------------------------------------------------- box predictor losses
def pack_pred(scores, deltas, std):
    R_ = scores.shape[0]
    return torch.cat([scores, deltas, std, torch.zeros(R_, 7)], dim=1).bfloat16()


@pytest.mark.parametrize("branch", ["supervised", "unsup_data_train"])
def test_fastrcnn_losses_fwd_bwd(branch):
    from oracle import ut2_rcnn_oracle as OR
    from ubteacher import ops_rcnn as R
    g = load("rcnn_box_losses.pt")
    sampled = g["sampled"]
    N = len(sampled)
    Rcap = 512
    s = {"proposal_boxes": torch.zeros(N, Rcap, 4), "gt_classes": torch.full((N, Rcap), -1, dtype=torch.int64),
         "gt_boxes": torch.zeros(N, Rcap, 4), "gt_loc_std": torch.zeros(N, Rcap, 4), "count": torch.zeros(N, dtype=torch.int32)}
    pred = torch.zeros(N, Rcap, 96)
    off = 0
    rows = []
    for i, x in enumerate(sampled):
        n = len(x["gt_classes"])
        s["count"][i] = n
        for k in ("proposal_boxes", "gt_classes", "gt_boxes", "gt_loc_std"):
            s[k][i, :n] = x[k]
        pred[i, :n] = pack_pred(g["scores"][off:off + n], g["deltas"][off:off + n], g["deltas_std"][off:off + n]).float()
        rows.append(pred[i, :n])
        off += n
    s = {k: v.cuda() for k, v in s.items()}
    pd = pred.reshape(N * Rcap, 96).bfloat16().cuda()
    mode = 0 if branch == "supervised" else 1
    losses = R.fastrcnn_loss_fwd(pd, s, mode)
    gout = torch.tensor([1.3, 0.7], device="cuda")
    dpred = R.fastrcnn_loss_bwd(pd, s, mode, gout)
    torch.cuda.synchronize()
    allp = torch.cat(rows)                                          # bf16-rounded values
    sc, dl, st = [allp[:, a:b].clone().requires_grad_(True) for a, b in ((0, 81), (81, 85), (85, 89))]
    cat = lambda k: torch.cat([x[k] for x in sampled])
    ref = OR.fast_rcnn_losses(sc, dl, st, cat("proposal_boxes"), cat("gt_boxes"), cat("gt_classes"), branch, cat("gt_loc_std"))
    (ref["loss_cls"] * 1.3 + ref["loss_box_reg"] * 0.7).backward()
    got = losses.cpu()
    assert abs(got[0] - ref["loss_cls"]) <= 3e-5 * abs(ref["loss_cls"]) + 1e-7          # fp32 sums, ~900 rows
    assert abs(got[1] - ref["loss_box_reg"]) <= 3e-5 * abs(ref["loss_box_reg"]) + 1e-7
    dp = dpred.float().cpu().reshape(N, Rcap, 96)
    drows = torch.cat([dp[i, :int(s["count"][i])] for i in range(N)])
    z = lambda v: v.grad if v.grad is not None else torch.zeros_like(v)
    # bf16-stored gradients: 2^-8 relative (+ small absolute floor for near-cancelling softmax terms)
    assert torch.allclose(drows[:, :81], z(sc), rtol=1 / 128, atol=2e-7)
    assert torch.allclose(drows[:, 81:85], z(dl), rtol=1 / 128, atol=2e-6), (drows[:, 81:85] - z(dl)).abs().max()
    assert torch.allclose(drows[:, 85:89], z(st), rtol=1 / 128, atol=2e-6), (drows[:, 85:89] - z(st)).abs().max()
    for i in range(N):
        assert (dp[i, int(s["count"][i]):] == 0).all()


# --------------------------------------------------------------------------------------- inference
def test_fastrcnn_inference_and_threshold():
    from oracle import ut2_rcnn_oracle as OR
    from oracle import ut2_oracle as O
    from ubteacher import ops, ops_rcnn as R
    g = load("rcnn_inference.pt")
    N = len(g["proposal_boxes"])
    Rcap = 256
    rois = torch.zeros(N, Rcap, 4)
    pred = torch.zeros(N, Rcap, 96)
    cnt = torch.zeros(N, dtype=torch.int32)
    off = 0
    for i, pb in enumerate(g["proposal_boxes"]):
        n = pb.shape[0]
        rois[i, :n] = pb
        pred[i, :n] = pack_pred(g["scores"][off:off + n], g["deltas"][off:off + n], g["deltas_std"][off:off + n]).float()
        cnt[i] = n
        off += n
    image_hw = torch.tensor([list(g["image_size"])] * N, dtype=torch.float32).cuda()
    pd = pred.reshape(N * Rcap, 96).bfloat16().cuda()
    out = R.fastrcnn_inference(pd, rois.cuda(), cnt.cuda(), image_hw)
    thr = ops.threshold_scatter(out, 0, 0.7)
    torch.cuda.synchronize()
    assert int(out["overflow"][0]) == 0
    for i in range(N):
        n = int(cnt[i])
        p = pred[i, :n].bfloat16().float()
        cb, cs, cc, cr = OR.fast_rcnn_candidates(p[:, :81], p[:, 81:85], rois[i, :n], g["image_size"])
        cand = out["candidates"]
        m = int(cand["count"][i])
        canon = cand["canon"][i, :m].cpu().long()
        order = torch.argsort(canon)
        rcanon = cr * 80 + cc
        near = (cs - 0.05).abs() < 1e-6                                # threshold borderline: expf ulps
        got_set, ref_set = set(canon.tolist()), set(rcanon.tolist())
        assert (got_set ^ ref_set) <= set(rcanon[near].tolist())
        if got_set == ref_set:
            assert torch.allclose(cand["boxes"][i, :m].cpu()[order], cb, rtol=1e-6, atol=1e-4)
            assert torch.allclose(cand["scores"][i, :m].cpu()[order], cs, rtol=1e-5, atol=1e-7)
        # NMS + gather: exact on the device's own candidates (visited in (score desc, canon asc) order)
        dbx, dsc, dcl = cand["boxes"][i, :m].cpu()[order], cand["scores"][i, :m].cpu()[order], cand["classes"][i, :m].cpu()[order]
        keep = O.batched_nms(dbx, dsc, dcl.long(), 0.5, "cuda")[:100]
        k = int(out["count"][i])
        assert k == len(keep)
        assert torch.equal(out["pred_boxes"][i, :k].cpu(), dbx[keep])
        assert torch.equal(out["scores"][i, :k].cpu(), dsc[keep])
        assert torch.equal(out["pred_classes"][i, :k].cpu(), dcl[keep].long())
        roi_idx = canon[order][keep] // 80
        assert torch.equal(out["roi_index"][i, :k].cpu().long(), roi_idx)
        assert torch.equal(out["reg_pred_std"][i, :k].cpu(), p[roi_idx, 85:89])
        # UBRCNNTeacherTrainer.threshold_bbox (trainer.py:727-751) on the detections
        ref_t = OR.threshold_bbox_rcnn({"pred_boxes": out["pred_boxes"][i, :k].cpu(), "scores": out["scores"][i, :k].cpu(),
                                        "pred_classes": out["pred_classes"][i, :k].cpu(),
                                        "pred_boxes_std": out["reg_pred_std"][i, :k].cpu()}, 0.7)
        t = int(thr["count"][i])
        assert t == len(ref_t["scores"])
        assert torch.equal(thr["pred_boxes"][i, :t].cpu(), ref_t["gt_boxes"])
        assert torch.equal(thr["pred_classes"][i, :t].cpu(), ref_t["gt_classes"])
        assert torch.equal(thr["reg_pred_std"][i, :t].cpu(), ref_t["pred_boxes_std"])


def test_helpers():
    from ubteacher import ops_rcnn as R
    g = torch.Generator().manual_seed(1)
    a = torch.randn(3, 5, 7, 16, generator=g)
    b = rb(torch.randn(3, 5, 7, 16, generator=g))
    out = R.add_f32_bf16(a.cuda(), b.bfloat16().cuda())
    assert torch.equal(out.float().cpu(), (a + b).bfloat16().float())
    x = rb(torch.randn(2, 25, 42, 32, generator=g))
    y = R.subsample2x(x.bfloat16().cuda())
    assert torch.equal(y.float().cpu(), x[:, ::2, ::2])


@pytest.mark.parametrize("trick_limit", [20000, 100])
def test_nms_segmented_equals_joint_nms(trick_limit):
    """Per-(image, level) NMS + merge == the joint batched NMS with idxs = level (both torchvision strategies: coordinate
    trick for small calls, per class otherwise), including -inf scores, an empty segment and a max_keep that cuts."""
    from ubteacher import ops_rcnn as R
    g = torch.Generator().manual_seed(31)
    N, seg_sizes = 3, [700, 300, 0, 90, 17]
    seg_off = [0]
    for n in seg_sizes:
        seg_off.append(seg_off[-1] + n)
    M = seg_off[-1] + 5                       # a few unused slots after the last segment
    ctr = torch.rand(N, M, 2, generator=g) * 300
    wh = torch.rand(N, M, 2, generator=g) * 120 + 4
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], dim=-1)
    scores = torch.randn(N, M, generator=g)
    scores[:, ::7] = scores[:, 3:4]           # ties across and inside segments
    scores[0, 5:40] = float("-inf")
    lvl = torch.zeros(N, M, dtype=torch.int32)
    for s in range(len(seg_sizes)):
        lvl[:, seg_off[s]:seg_off[s + 1]] = s
    tie = torch.stack([torch.randperm(M, generator=g) for _ in range(N)]).to(torch.int32)
    cnt = torch.full((N,), seg_off[-1], dtype=torch.int32)
    for max_keep in (1000, 60):
        k_ref, c_ref = R.nms_batched(boxes.cuda(), scores.cuda(), lvl.cuda(), cnt.cuda(), 0.7, max_keep, tie=tie.cuda(), trick_limit=trick_limit)
        k_seg, c_seg = R.nms_segmented(boxes.cuda(), scores.cuda(), tie.cuda(), seg_off, 0.7, max_keep, trick_limit=trick_limit)
        torch.cuda.synchronize()
        assert torch.equal(c_ref, c_seg), (c_ref, c_seg)
        for i in range(N):
            n = int(c_ref[i])
            assert torch.equal(k_ref[i, :n], k_seg[i, :n]), (i, max_keep)
