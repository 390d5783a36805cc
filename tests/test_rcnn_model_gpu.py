"""End-to-end GPU parity of the Faster R-CNN engine / trainer against the fp32 CPU oracle model
(oracle/ut2_rcnn_model.py) on the same seeded weights, inputs and sampling keys. bf16 tensor-core activations vs
fp32: tolerances are relative L2 errors / cosine similarities, stated at each check."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu().double(), b.float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-12))


def make_batch(n, sizes, seed, nbox=5):
    from ubteacher.data.synthetic import synth_instances
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        h, w = sizes[i % len(sizes)]
        inst = synth_instances(g, h, w, nbox)
        out.append({"image": torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8), "instances": inst,
                    "boxes": inst.gt_boxes.tensor, "classes": inst.gt_classes})
    return out


def diversify(model, seed=3):
    """Random-init heads give near-constant outputs; widen them so proposals / detections are non-trivial."""
    g = torch.Generator().manual_seed(seed)
    V = model.engine.arena.views
    for name, scale in (("proposal_generator.rpn_head.objectness_logits.weight", 0.05),
                        ("proposal_generator.rpn_head.anchor_deltas.weight", 0.02),
                        ("roi_heads.box_predictor.cls_score.weight", 0.08), ("roi_heads.box_predictor.bbox_pred.weight", 0.04),
                        ("roi_heads.box_predictor.bbox_pred_std.weight", 0.04)):
        V[name].copy_((torch.randn(V[name].shape, generator=g) * scale).to(V[name].device))
    model.engine.refresh_operands()


@pytest.fixture(scope="module")
def model():
    from util_cfg import rcnn_cfg
    from ubteacher.modeling import TwoStagePseudoLabGeneralizedRCNN
    m = TwoStagePseudoLabGeneralizedRCNN(rcnn_cfg())
    diversify(m)
    m.train()
    return m


def cpu_sd(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def test_state_dict_keys_and_shapes(model):
    sd = model.state_dict()
    assert sd["backbone.fpn_lateral2.weight"].shape == (256, 256, 1, 1)
    assert sd["proposal_generator.rpn_head.conv.weight"].shape == (256, 256, 3, 3)
    assert sd["proposal_generator.rpn_head.objectness_logits.weight"].shape == (3, 256, 1, 1)
    assert sd["proposal_generator.rpn_head.anchor_deltas.bias"].shape == (12,)
    assert sd["roi_heads.box_head.fc1.weight"].shape == (1024, 12544)
    assert sd["roi_heads.box_head.fc2.weight"].shape == (1024, 1024)
    assert sd["roi_heads.box_predictor.cls_score.weight"].shape == (81, 1024)
    assert sd["roi_heads.box_predictor.bbox_pred_std.bias"].shape == (4,)
    assert "pixel_mean" not in sd                       # non-persistent in [D2] v0.6
    n = sum(v.numel() for k, v in sd.items() if ".norm." not in k)
    assert abs(n - 41.38e6) < 0.05e6                    # 41.38 M parameters
    # load_state_dict round trip through the reference layout of fc1 ([1024, 256*7*7], (c, h, w) order)
    w = sd["roi_heads.box_head.fc1.weight"].clone()
    model.load_state_dict({k: v.clone() for k, v in sd.items()})
    assert torch.equal(model.state_dict()["roi_heads.box_head.fc1.weight"], w)


def test_features_and_rpn_match_oracle(model):
    from oracle import ut2_rcnn_model as RM
    batch = make_batch(2, [(150, 200), (128, 180)], 1)
    sd = cpu_sd(model)
    fwd = model.engine.forward_features([b["image"].cuda() for b in batch], train=False)
    ref = RM.forward_features(sd, [b["image"] for b in batch])
    N, geom = 2, fwd["geom"]
    for l in range(5):      # FPN outputs: bf16 chain of ~55 convs, rel L2 < 3 %
        assert rel(fwd["levels"][l].permute(0, 3, 1, 2), ref["feats"][l]) < 0.03, l
        h, w = geom.hw[l]
        lo, hi = geom.off[l] * N, geom.off[l + 1] * N
        out = fwd["rpn_out"][lo:hi].view(N, h * w, 16).float().cpu()
        assert rel(out[:, :, :3].reshape(N, -1), ref["logits"][l]) < 0.04, l
        assert rel(out[:, :, 3:15].reshape(N, -1, 4), ref["deltas"][l]) < 0.04, l
        assert (out[:, :, 15] == 0).all()


def _inject_keys(model, N, A, seed):
    g = torch.Generator().manual_seed(seed)
    kr = torch.randint(0, 2 ** 32, (N, A), generator=g, dtype=torch.int64)
    ko = torch.randint(0, 2 ** 32, (N, 1200), generator=g, dtype=torch.int64)
    model.engine.debug_keys = {"rpn": kr.cuda().to(torch.int32), "roi": ko.cuda().to(torch.int32)}
    return kr, ko


GRAD_KEYS = ["roi_heads.box_predictor.cls_score.weight", "roi_heads.box_predictor.bbox_pred.weight",
             "roi_heads.box_predictor.bbox_pred_std.weight", "roi_heads.box_predictor.cls_score.bias",
             "roi_heads.box_head.fc2.weight", "roi_heads.box_head.fc1.weight", "roi_heads.box_head.fc1.bias",
             "proposal_generator.rpn_head.objectness_logits.weight", "proposal_generator.rpn_head.anchor_deltas.weight",
             "proposal_generator.rpn_head.anchor_deltas.bias", "proposal_generator.rpn_head.conv.weight",
             "backbone.fpn_output2.weight", "backbone.fpn_output4.weight", "backbone.fpn_lateral2.weight",
             "backbone.fpn_lateral5.weight", "backbone.fpn_lateral3.bias", "backbone.bottom_up.res5.2.conv3.weight",
             "backbone.bottom_up.res4.0.conv1.weight", "backbone.bottom_up.res3.1.conv2.weight"]


@pytest.mark.parametrize("branch", ["supervised", "unsup_data_train"])
def test_losses_and_gradients_match_oracle(model, branch):
    from oracle import ut2_rcnn_model as RM
    from ubteacher.modeling.fcos.fcos_outputs import BoxSet
    batch = make_batch(3, [(160, 224), (128, 192)], 2 if branch == "supervised" else 5)
    sd = cpu_sd(model)
    eng = model.engine
    N = len(batch)
    geom, _ = eng.level_geom(160, 224)
    kr, ko = _inject_keys(model, N, geom.A, 11)
    gt = {"boxes": [b["boxes"] for b in batch], "classes": [b["classes"] for b in batch]}
    if branch == "unsup_data_train":
        g = torch.Generator().manual_seed(8)
        gt["scores"] = [torch.rand(len(b), generator=g) * 0.3 + 0.7 for b in gt["boxes"]]
        gt["std"] = [torch.randn(len(b), 4, generator=g) * 2 for b in gt["boxes"]]
        G = 128
        bx, cl, sc, st = torch.zeros(N, G, 4), torch.zeros(N, G, dtype=torch.int64), torch.zeros(N, G), torch.zeros(N, G, 4)
        cnt = torch.zeros(N, dtype=torch.int32)
        for i in range(N):
            n = len(gt["boxes"][i])
            bx[i, :n], cl[i, :n], sc[i, :n], st[i, :n], cnt[i] = gt["boxes"][i], gt["classes"][i], gt["scores"][i], gt["std"][i], n
        bs = BoxSet(bx.cuda(), cl.cuda(), cnt.cuda(), st.cuda(), sc.cuda())
        batch = [dict(b, instances=bs) for b in batch]
    eng.arena.grad.zero_()
    losses, pending = model.forward_train(batch, branch)
    w = [1.3, 0.6, 1.0, 0.8]
    props = pending["ctx"]["proposals"]
    pcnt = props["count"].cpu().tolist()
    dev_props = [props["proposal_boxes"][i, :pcnt[i]].cpu() for i in range(N)]
    model.backward_pending(pending, w)
    torch.cuda.synchronize()
    eng.debug_keys = None
    tk = RM.trainable_keys(sd)
    params = {k: sd[k].clone().requires_grad_(True) for k in tk}
    sdp = dict(sd)
    sdp.update(params)
    keys_roi = [ko[i, :pcnt[i] + len(gt["boxes"][i])] for i in range(N)]
    ref, aux = RM.forward_train(sdp, [b["image"] for b in batch], gt, branch, [kr[i] for i in range(N)], keys_roi, dev_props)
    # the device's own proposals come from an independent fp32 pass in the oracle too: the sets agree up to NMS borderlines
    for k in ("loss_rpn_cls", "loss_rpn_loc", "loss_cls", "loss_box_reg"):   # fp32 loss arithmetic on bf16 activations: 4 %
        torch.testing.assert_close(losses[k].cpu(), ref[k].detach(), rtol=4e-2, atol=2e-3)
    (ref["loss_rpn_cls"] * w[0] + ref["loss_rpn_loc"] * w[1] + ref["loss_cls"] * w[2] + ref["loss_box_reg"] * w[3]).backward()
    # second oracle run with the device's bf16 rounding points (oracle/ut2_model.py: bf16_points): rounding vs logic
    from oracle import ut2_model as M
    params_q = {k: sd[k].clone().requires_grad_(True) for k in tk}
    sdq = dict(sd)
    sdq.update(params_q)
    with M.bf16_points():
        refq, _ = RM.forward_train(sdq, [b["image"] for b in batch], gt, branch, [kr[i] for i in range(N)], keys_roi, dev_props)
        (refq["loss_rpn_cls"] * w[0] + refq["loss_rpn_loc"] * w[1] + refq["loss_cls"] * w[2] + refq["loss_box_reg"] * w[3]).backward()
    for k in ("loss_rpn_cls", "loss_rpn_loc", "loss_cls", "loss_box_reg"):
        torch.testing.assert_close(losses[k].cpu(), refq[k].detach(), rtol=4e-2, atol=2e-3)
    G = {k: v for k, v in model.named_parameters()}
    sdg = eng.arena.gviews
    report = []
    for k in GRAD_KEYS:
        gv = sdg[k]
        if k == "roi_heads.box_head.fc1.weight":
            gv = gv.reshape(1024, 12544)
        if params[k].grad is None or float(params[k].grad.abs().max()) == 0.0:
            # e.g. bbox_pred_std under 'tsbetter' (the teacher-better mask is not differentiable): no gradient on either side
            assert float(gv.float().abs().max()) == 0.0, k
            continue
        a, b = gv.float().cpu().double().flatten(), params[k].grad.double().flatten()
        cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
        ratio = float(a.norm() / (b.norm() + 1e-30))
        # same depth-aware bounds as the FCOS model test (bf16 activations and bf16 back-propagated gradients vs fp32)
        # measured on B200 (profiles/r02_gradient_parity.txt): >= 0.9997 everywhere except fpn_output4 (0.998: a ROI on a
        # pooler-level borderline) — bf16 rounding of the whole chain; logic is bounded per block in test_model_gpu.py
        need = 0.995
        assert cos > need and 0.95 < ratio < 1.05, (k, cos, ratio)
        bq = params_q[k].grad.double().flatten()
        report.append((k, cos, ratio, float((a * bq).sum() / (a.norm() * bq.norm() + 1e-30)), float(a.norm() / (bq.norm() + 1e-30))))
        # bf16-point oracle: the non-differentiable 'tsbetter' mask (c_t > c_s + 0.1) flips on borderline elements
        assert report[-1][3] > 0.99 and 0.95 < report[-1][4] < 1.05, report[-1]
    print("\n".join(f"{k:58s} fp32: cos {c:.4f} ratio {r:.3f} | bf16 points: cos {cq:.5f} ratio {rq:.4f}" for k, c, r, cq, rq in report))
    eng.arena.grad.zero_()


def test_teacher_inference_matches_oracle(model):
    from oracle import ut2_rcnn_model as RM
    batch = make_batch(2, [(160, 224), (128, 192)], 9)
    sd = cpu_sd(model)
    props, dets, pred = model.forward_teacher(batch)
    torch.cuda.synchronize()
    pcnt = props["count"].cpu().tolist()
    assert all(0 < n <= 1000 for n in pcnt)
    dev_props = [props["proposal_boxes"][i, :pcnt[i]].cpu() for i in range(2)]
    _, rdets, (sc, dl, st) = RM.forward_teacher(sd, [b["image"] for b in batch], proposals=dev_props)
    off = 0
    for i in range(2):
        n = pcnt[i]
        p = pred.view(2, -1, 96)[i, :n].float().cpu()
        assert rel(p[:, :81], sc[off:off + n]) < 0.05 and rel(p[:, 81:85], dl[off:off + n]) < 0.05
        off += n
        k = int(dets["count"][i])
        # detections: same count up to score-threshold / NMS borderlines, and every device detection has an oracle twin
        assert abs(k - len(rdets[i]["scores"])) <= max(3, k // 10), (k, len(rdets[i]["scores"]))
        if k and len(rdets[i]["scores"]):
            from oracle import ut2_rcnn_oracle as R
            iou = R.pairwise_iou(dets["pred_boxes"][i, :k].cpu(), rdets[i]["pred_boxes"])
            same = dets["pred_classes"][i, :k].cpu()[:, None] == rdets[i]["pred_classes"][None, :]
            # (the bit-exact check of the inference kernels on identical inputs is tests/test_rcnn_kernels_gpu.py; here bf16 vs
            #  fp32 scores straddle the 0.05 threshold, which changes the candidate sets NMS sees)
            assert float(((iou > 0.9) & same).any(dim=1).float().mean()) > 0.6


def test_autograd_bridge_equals_explicit_backward(model):
    batch = make_batch(2, [(128, 160)], 4)
    A = model.engine.arena
    A.grad.zero_()
    d0 = model.engine.draws
    losses, pending = model.forward_train(batch, "supervised")
    model.backward_pending(pending, [1.0, 2.0, 3.0, 0.5])
    g1 = A.grad.clone()
    A.grad.zero_()
    model.engine.draws = d0                       # replay the same sampling draws
    out, _, _, _ = model(batch, branch="supervised")
    (out["loss_rpn_cls"] * 1.0 + out["loss_rpn_loc"] * 2.0 + out["loss_cls"] * 3.0 + out["loss_box_reg"] * 0.5).backward()
    g2 = A.grad.clone()
    A.grad.zero_()
    assert rel(g2, g1) < 2e-3     # same kernels; only the fp32 atomic accumulation order differs


def test_trainer_steps():
    """Two UBRCNNTeacherTrainer steps at small resolution: finite losses with the reference's names, the student moves,
    the teacher is the bit-exact EMA of the student state (trainer.py:950-968)."""
    from oracle import ut2_oracle as O
    from util_cfg import rcnn_cfg
    from ubteacher.engine import UBRCNNTeacherTrainer

    class Loader:
        def __init__(self):
            self.i = 0

        def __iter__(self):
            return self

        def __next__(self):
            self.i += 1
            mk = lambda n, seed: make_batch(n, [(128, 160), (160, 192)], seed, nbox=4)
            lq, uq = mk(1, 100 + self.i), mk(2, 200 + self.i)
            lk = [dict(d, image=torch.flip(d["image"], [0])) for d in lq]
            uk = [dict(d) for d in uq]
            return lq, lk, uq, uk

    tr = UBRCNNTeacherTrainer(rcnn_cfg(), data_loader=Loader())
    diversify(tr.model)
    assert tr.model_teacher.training and tr.pseudo_generator is None
    s0 = tr.model.engine.arena.data.clone()
    tr.iter = 0
    tr.run_step_full_semisup()               # iter == BURN_UP_STEP: copy then EMA
    names, vec = tr.last_losses
    assert set(names) >= {"loss_cls", "loss_box_reg", "loss_rpn_cls", "loss_rpn_loc", "loss_cls_pseudo", "loss_box_reg_pseudo",
                          "loss_rpn_cls_pseudo", "loss_rpn_loc_pseudo"}
    assert torch.isfinite(vec).all()
    s1 = tr.model.engine.arena.data.clone()
    t1 = tr.model_teacher.engine.arena.data.clone()
    assert torch.allclose(t1, s0, rtol=1e-6, atol=1e-8)   # keep-0 copy, then EMA of two equal states (a*(1-k) + a*k)
    assert not torch.equal(s1, s0)
    tr.iter = 1
    tr.run_step_full_semisup()
    t2 = tr.model_teacher.engine.arena.data
    ref = O.ema_update(s1.cpu(), t1.cpu(), tr.cfg.SEMISUPNET.EMA_KEEP_RATE)
    assert torch.equal(t2.cpu(), ref)
    assert torch.isfinite(tr.last_losses[1]).all()


def _rcnn_steps_vs_oracle(sizes, n_steps, n_label, n_unlabel, nbox):
    """UBRCNNTeacherTrainer.run_step_full_semisup against oracle ut2_rcnn_step (trainer.py:786-912) on the same weights,
    images and injected sampling keys. The oracle is driven with the device's pseudo-label set and proposal boxes (score /
    NMS borderlines differ between bf16 and fp32), everything else — anchor labelling + sampling, RPN losses, ROI sampling,
    ROIAlign, box head, focal / L1 + NLL / tsbetter losses, the loss weights (lambda, 0, lambda, mu), SGD, EMA — is independent."""
    from oracle import ut2_rcnn_model as RM
    from util_cfg import rcnn_cfg
    from ubteacher.arena import _view
    from ubteacher.d2compat.events import EventStorage
    from ubteacher.engine import UBRCNNTeacherTrainer

    class Loader:
        def __init__(self):
            self.i = 0

        def __iter__(self):
            return self

        def __next__(self):
            self.i += 1
            mk = lambda n, seed: make_batch(n, sizes, seed, nbox=nbox)
            lq, uq = mk(n_label, 100 + self.i), mk(n_unlabel, 200 + self.i)
            lk = [dict(d, image=torch.flip(d["image"], [0])) for d in lq]
            uk = [dict(d) for d in uq]
            return lq, lk, uq, uk

    # a fifth of the recipe's learning rate: with randomly initialised, widened predictors one step at 0.01 can push the
    # box deltas far enough that the NLL * IoU term of the NEXT step is NaN (in the oracle too, fed the same weights)
    cfg = rcnn_cfg(**{"SOLVER.BASE_LR": 0.002})
    tr = UBRCNNTeacherTrainer(cfg, data_loader=Loader())
    diversify(tr.model)
    tr.scheduler.warmup_iters = 0
    tr.scheduler.last_epoch = -1
    tr.scheduler.step()
    eng = tr.model.engine
    Hp, Wp = eng.padded_size(sizes)
    geom, _ = eng.level_geom(Hp, Wp)
    nmax = max(2 * n_label, n_unlabel)
    kr, ko = _inject_keys(tr.model, nmax, geom.A, 21)
    # a pseudo-label threshold that the randomly initialised teacher actually clears: the median detection score
    probe = next(Loader())[3]
    tr.model.train()
    _, _, dets, _ = tr.model(probe, branch="unsup_data_weak")
    sc = torch.cat([dets["scores"][i, :int(n)] for i, n in enumerate(dets["count"].cpu().tolist())])
    assert sc.numel() > 4
    cfg.defrost()
    cfg.SEMISUPNET.BBOX_THRESHOLD = float(sc.median())
    student = cpu_sd(tr.model)
    init = {k: v.clone() for k, v in student.items()}
    before = {k: v.clone() for k, v in student.items()}
    teacher = {k: v.clone() for k, v in student.items()}      # iter == BURN_UP_STEP copies the student first (trainer.py:815-817)
    mom = {}
    ref_loader = Loader()
    ss = cfg.SEMISUPNET
    with EventStorage(0) as tr.storage:
        for it in range(n_steps):
            tr.iter = it
            lr = tr.optimizer.param_groups[0]["lr"]
            pend, pseudo_dev = [], []
            f_orig, p_orig = tr.model.forward_train, tr.process_pseudo_label

            def f_spy(*a, **k):
                out = f_orig(*a, **k)
                pend.append(out[1])
                return out

            def p_spy(*a, **k):
                out = p_orig(*a, **k)
                pseudo_dev.append(out[0])
                return out

            tr.model.forward_train, tr.process_pseudo_label = f_spy, p_spy
            tr.run_step_full_semisup()
            tr.model.forward_train, tr.process_pseudo_label = f_orig, p_orig
            names, vec = tr.last_losses
            got = dict(zip(names, vec.cpu().tolist()))
            bs = pseudo_dev[0]
            cnt = bs.counts.cpu().tolist()
            assert sum(cnt) > 0, "test needs a non-empty pseudo-label set"
            pseudo = {"boxes": [bs.boxes[i, :n].cpu() for i, n in enumerate(cnt)], "classes": [bs.classes[i, :n].cpu() for i, n in enumerate(cnt)],
                      "scores": [bs.scores[i, :n].cpu() for i, n in enumerate(cnt)], "std": [bs.reg_pred_std[i, :n].cpu() for i, n in enumerate(cnt)]}
            batch = next(ref_loader)
            props, keys = {}, {}
            ngt = {"sup": [len(d["boxes"]) for d in batch[0] + batch[1]], "unsup": cnt}
            for tag, pd in zip(("sup", "unsup"), pend):
                pc = pd["ctx"]["proposals"]["count"].cpu().tolist()
                props[tag] = [pd["ctx"]["proposals"]["proposal_boxes"][i, :n].cpu() for i, n in enumerate(pc)]
                keys["rpn_" + tag] = [kr[i] for i in range(len(pc))]
                keys["roi_" + tag] = [ko[i, :pc[i] + ngt[tag][i]] for i in range(len(pc))]
            ocfg = {"UNSUP_LOSS_WEIGHT": ss.UNSUP_LOSS_WEIGHT, "UNSUP_REG_LOSS_WEIGHT": ss.UNSUP_REG_LOSS_WEIGHT,
                    "EMA_KEEP_RATE": ss.EMA_KEEP_RATE, "BBOX_THRESHOLD": ss.BBOX_THRESHOLD, "WEIGHT_DECAY": cfg.SOLVER.WEIGHT_DECAY,
                    "MOMENTUM": cfg.SOLVER.MOMENTUM, "LR": lr}
            rec, grads, _ = RM.ut2_rcnn_step(student, teacher, mom, batch, ocfg, it == 0, keys, pseudo_override=pseudo,
                                             proposals_override=props)
            assert set(got) >= {k for k in rec if k.startswith("loss")}
            for k, v in rec.items():
                if k.startswith("loss"):       # fp32 loss arithmetic on bf16 activations: 4 %
                    assert abs(got[k] - float(v)) <= 4e-2 * abs(float(v)) + 2e-3, (it, k, got[k], float(v))
            # this step's parameter UPDATE (loss weights lambda, 0, lambda, mu; SGD with momentum and weight decay) points the
            # same way ...
            sd = cpu_sd(tr.model)
            for k in ["roi_heads.box_predictor.cls_score.weight", "roi_heads.box_head.fc2.weight", "proposal_generator.rpn_head.conv.weight",
                      "proposal_generator.rpn_head.anchor_deltas.weight", "backbone.fpn_output3.weight",
                      "backbone.bottom_up.res5.1.conv2.weight", "backbone.bottom_up.res3.2.conv1.weight"]:
                a = (sd[k] - before[k]).double().flatten()
                b = (student[k] - before[k]).double().flatten()
                assert float((a * b).sum() / (a.norm() * b.norm())) > 0.97, (it, k)
                assert 0.9 < float(a.norm() / b.norm()) < 1.1, (it, k)
            # ... frozen parameters did not move and the EMA teacher tracks the oracle's
            assert torch.equal(sd["backbone.bottom_up.res2.0.conv1.weight"], init["backbone.bottom_up.res2.0.conv1.weight"])
            td = cpu_sd(tr.model_teacher)
            for k in ["roi_heads.box_predictor.cls_score.weight", "backbone.bottom_up.stem.conv1.weight", "backbone.fpn_lateral4.weight"]:
                assert rel(td[k], teacher[k]) < 1e-5, (it, k)
            # every step is checked from the SAME starting state: re-seed the oracle with the device's student, teacher and
            # momentum (otherwise step 2 compares two trajectories that have already drifted apart by one bf16 update)
            A = eng.arena
            student, teacher = sd, td
            mom = {k: _view(A.mom, A.offset[k], A.specs[k].shape).detach().cpu().clone().reshape(student[k].shape)
                   for k in RM.trainable_keys(student)}
            before = {k: v.clone() for k, v in student.items()}
            tr.scheduler.step()
            tr.storage.step()
    eng.debug_keys = None


def test_trainer_step_vs_oracle_step():
    _rcnn_steps_vs_oracle([(128, 160), (160, 192)], 2, 1, 2, 4)


def test_full_size_step_vs_oracle():
    """BASELINE configs #3 / #5 geometry: ONE full-size step (1 labeled + 1 unlabeled 3 x 800 x 1333 image, 268 569 anchors
    per image) against the oracle step."""
    _rcnn_steps_vs_oracle([(800, 1333)], 1, 1, 1, 7)


def test_cuda_graph_step_matches_eager():
    """The captured-and-replayed R-CNN step (device-resident sampling counter) against the eager schedule on the same
    injected sampling keys. The learning rate is tiny so that the discrete choices downstream of the weights (top-k, NMS,
    fg/bg matching) do not amplify the fp32-atomic accumulation-order noise two eager runs already show; the parameter
    UPDATE (not the parameters) is what gets compared."""
    from util_cfg import rcnn_cfg
    from ubteacher.d2compat.events import EventStorage
    from ubteacher.data.synthetic import SyntheticTwoCropLoader
    from ubteacher.engine import UBRCNNTeacherTrainer

    def run(graph):
        tr = UBRCNNTeacherTrainer(rcnn_cfg(**{"SOLVER.BASE_LR": 1e-5}),
                                  data_loader=SyntheticTwoCropLoader(1, 2, h=128, w=160, boxes_per_image=3, pool=2))
        diversify(tr.model)
        s0 = tr.model.engine.arena.data.clone()
        geom, _ = tr.model.engine.level_geom(128, 160)
        _inject_keys(tr.model, 2, geom.A, 11)
        tr.enable_cuda_graph(graph)
        out = []
        with EventStorage(0) as tr.storage:
            for it in range(4):
                tr.iter = it
                tr.run_step_full_semisup()
                out.append(tr.last_losses[1].cpu().clone())
                tr.scheduler.step()
        return out, tr.model.engine.arena.data - s0, tr

    eager, ps, _ = run(False)
    graph, gs, tr = run(True)
    assert any(e["graph"] is not None for e in tr._graphs.values()), "the step was never captured"
    assert int(tr.model.engine.seed_dev.item()) == 4      # advanced before every replay
    for i, (a, b) in enumerate(zip(eager, graph)):
        assert torch.isfinite(b).all()
        torch.testing.assert_close(a, b, rtol=[1e-5, 5e-3, 5e-3, 5e-3][i], atol=1e-3)
    assert rel(gs, ps) < 2e-2


def test_cuda_graph_padded_key_serves_mixed_image_sizes():
    """Mixed image sizes inside a batch (a real dataset): the R-CNN step's graph is keyed by the padded size of the forward groups,
    the image sizes are device data for the stem (hw_dev) and for the proposal / detection clipping (image_hw), and one captured
    graph follows the eager schedule over batches whose images differ in size."""
    import random
    from util_cfg import rcnn_cfg
    from ubteacher.d2compat.events import EventStorage
    from ubteacher.d2compat.structures import Boxes, Instances
    from ubteacher.data.synthetic import SyntheticTwoCropLoader
    from ubteacher.engine import UBRCNNTeacherTrainer

    class MixedSizes:
        def __iter__(self):
            rng = random.Random(9)
            for batch in SyntheticTwoCropLoader(1, 2, h=128, w=160, boxes_per_image=3, pool=2):
                out = []
                for part in batch:
                    new = []
                    for d in part:
                        h, w = rng.randint(97, 128), rng.randint(129, 160)
                        d = dict(d)
                        d["image"] = d["image"][:, :h, :w].contiguous()
                        if "instances" in d:
                            b = d["instances"].gt_boxes.tensor.clone()
                            b[:, 0::2].clamp_(0, w)
                            b[:, 1::2].clamp_(0, h)
                            keep = (b[:, 2] - b[:, 0] > 2) & (b[:, 3] - b[:, 1] > 2)
                            inst = Instances((h, w))
                            inst.gt_boxes, inst.gt_classes = Boxes(b[keep]), d["instances"].gt_classes[keep]
                            d["instances"] = inst
                        new.append(d)
                    out.append(new)
                yield tuple(out)

    def run(graph):
        tr = UBRCNNTeacherTrainer(rcnn_cfg(**{"SOLVER.BASE_LR": 1e-5}), data_loader=MixedSizes())
        diversify(tr.model)
        geom, _ = tr.model.engine.level_geom(128, 160)
        _inject_keys(tr.model, 2, geom.A, 11)
        tr.enable_cuda_graph(graph)
        out = []
        with EventStorage(0) as tr.storage:
            for it in range(6):
                tr.iter = it
                tr.run_step_full_semisup()
                out.append(tr.last_losses[1].cpu().clone())
                tr.scheduler.step()
        return out, tr

    eager, _ = run(False)
    graph, tr = run(True)
    (key, e), = tr._graphs.items()
    assert key[0] == "padded" and e["graph"] is not None, "one captured graph for the six mixed-size batches"
    # steps 0 / 1 are eager in both runs, step 2 is the capture + first replay, step 3 a replay on images of other sizes: measured
    # bit-identical to the eager schedule. From step 4 on the two runs drift like two eager runs do (fp32-atomic order flips a
    # thresholded pseudo box or an NMS tie: 0.2 % .. 8 % on single loss terms), so only finiteness is asserted there.
    for i, (a, b) in enumerate(zip(eager, graph)):
        assert torch.isfinite(b).all()
        if i < 4:
            torch.testing.assert_close(a, b, rtol=[1e-5, 5e-3, 5e-3, 5e-3][i], atol=1e-3)


def test_eval_mode_inference_contract():
    """model.eval(): [D2] GeneralizedRCNN.inference -> [{"instances": Instances(pred_boxes, scores, pred_classes,
    pred_boxes_std)}] rescaled to the dict's height / width ([D2] detector_postprocess), RPN *_TEST top-k."""
    from util_cfg import rcnn_cfg
    import ubteacher.modeling  # noqa: F401  (registers the meta-architectures)
    from ubteacher.d2compat.registry import META_ARCH_REGISTRY
    m = META_ARCH_REGISTRY.get("TwoStagePseudoLabGeneralizedRCNN")(rcnn_cfg())
    diversify(m)
    m.eval()
    img = torch.randint(0, 256, (3, 128, 160), dtype=torch.uint8, generator=torch.Generator().manual_seed(4))
    a = m([{"image": img}])[0]["instances"]
    b = m([{"image": img, "height": 256, "width": 320}])[0]["instances"]
    assert a.image_size == (128, 160) and b.image_size == (256, 320)
    assert a.has("pred_boxes") and a.has("scores") and a.has("pred_classes") and a.has("pred_boxes_std") and len(a) <= 100
    ref = a.pred_boxes.tensor.cpu() * 2.0
    keep = ((ref[:, 2] - ref[:, 0]) > 0) & ((ref[:, 3] - ref[:, 1]) > 0)
    torch.testing.assert_close(b.pred_boxes.tensor.cpu(), ref[keep], rtol=0, atol=1e-3)
    assert torch.equal(b.scores.cpu(), a.scores.cpu()[keep])
