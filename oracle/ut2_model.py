"""TEST INFRASTRUCTURE ONLY — plain PyTorch fp32 CPU restatement of the FCOS R50-FPN detector and of the
Unbiased-Teacher-v2 training step, driven by a ``state_dict`` with the reference's key names.

Follows (paths under /root/reference/ubteacher; [D2] = Detectron2 v0.6, restated from SURVEY.md app. B):
  one_stage_detector.py:155-240, backbone/fpn.py:11-78, [D2] ResNet/FPN, fcos/fcos.py:83-376,
  engine/trainer.py:181-429 (run_step_full_semisup) and :468-486 (EMA), [D2] SGD.
Used by tests (parity of the CUDA path), by bench.py's cpu_baseline and `--impl reference` arm.
"""
import torch
import torch.nn.functional as F

from . import ut2_oracle as O

STAGES = [("res2", 3, 1), ("res3", 4, 2), ("res4", 6, 2), ("res5", 3, 2)]
BN_EPS = 1e-5

# ---------------------------------------------------------------------------------------------------------------------
# bf16 rounding points. The CUDA path keeps every activation AND every back-propagated activation gradient as a bf16
# tensor in HBM (fp32 accumulation inside a kernel, one round-to-nearest-even when the tensor is written) and multiplies
# bf16 copies of the fp32 master weights (FrozenBN scale folded in before the rounding). Inside ``with bf16_points():``
# this restatement rounds at exactly those places — q() where a tensor is produced and where it is consumed (forward
# value and incoming gradient), qw() on the tensor-core weight operand (forward only: weight gradients stay fp32) — so
# that a comparison with the device separates accumulated ROUNDING from LOGIC errors (tests/test_model_gpu.py,
# tests/test_rcnn_model_gpu.py). Outside the context the functions are the plain fp32 reference arithmetic.
_BF16 = [False]


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


class bf16_points:
    def __enter__(self):
        self.prev = _BF16[0]
        _BF16[0] = True

    def __exit__(self, *a):
        _BF16[0] = self.prev


def q(x):
    return _RoundBoth.apply(x) if _BF16[0] else x


def qw(w):
    return _RoundFwd.apply(w) if _BF16[0] else w


def conv(x, w, b=None, stride=1, pad=0):
    """conv2d through the rounding points: bf16 input / weight operands, fp32 accumulation + bias."""
    return F.conv2d(q(x), qw(w), b, stride, pad)


def frozen_bn(x, sd, p):
    """[D2] FrozenBatchNorm2d: y = x * (w * rsqrt(var + eps)) + (b - mean * scale)."""
    scale = sd[p + ".weight"] * (sd[p + ".running_var"] + BN_EPS).rsqrt()
    shift = sd[p + ".bias"] - sd[p + ".running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def conv_bn(x, sd, p, stride=1, pad=0):
    if _BF16[0]:   # the device folds the FrozenBN scale into the packed bf16 weights; the epilogue adds the shift
        n = p + ".norm"
        scale = sd[n + ".weight"] * (sd[n + ".running_var"] + BN_EPS).rsqrt()
        shift = sd[n + ".bias"] - sd[n + ".running_mean"] * scale
        return F.conv2d(q(x), qw(sd[p + ".weight"] * scale.view(-1, 1, 1, 1)), None, stride, pad) + shift.view(1, -1, 1, 1)
    return frozen_bn(F.conv2d(x, sd[p + ".weight"], None, stride, pad), sd, p + ".norm")


def stem(x, sd, p):
    if _BF16[0]:   # csrc/stem_tc.cu: bf16 pixels x bf16 (unscaled) filter, FrozenBN scale + shift in the fp32 epilogue
        return frozen_bn(F.conv2d(q(x), qw(sd[p + ".weight"]), None, 2, 3), sd, p + ".norm")
    return conv_bn(x, sd, p, 2, 3)


def trunk(sd, x):
    """[D2] BasicStem + 4 bottleneck stages (STRIDE_IN_1X1) -> {"res2".."res5"}."""
    bu = "backbone.bottom_up."
    x = q(F.relu(stem(x, sd, bu + "stem.conv1")))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = {}
    for stage, n, stride in STAGES:
        for i in range(n):
            p = f"{bu}{stage}.{i}."
            s = stride if i == 0 else 1
            sc = q(conv_bn(x, sd, p + "shortcut", s)) if (p + "shortcut.weight") in sd else x
            out = q(F.relu(conv_bn(x, sd, p + "conv1", s)))
            out = q(F.relu(conv_bn(out, sd, p + "conv2", 1, 1)))
            out = conv_bn(out, sd, p + "conv3")
            x = q(F.relu(out + sc))
        feats[stage] = x
    return feats


def fpn_topdown(sd, feats, levels):
    """[D2] FPN (fuse "sum", nearest 2x top-down): returns {level: output} for the given level numbers (descending)."""
    lat = lambda l, c: conv(c, sd[f"backbone.fpn_lateral{l}.weight"], sd[f"backbone.fpn_lateral{l}.bias"])
    outc = lambda l, t: conv(t, sd[f"backbone.fpn_output{l}.weight"], sd[f"backbone.fpn_output{l}.bias"], 1, 1)
    prev, outs = None, {}
    for l in levels:
        cur = lat(l, feats[f"res{l}"])
        prev = q(cur if prev is None else cur + F.interpolate(prev, scale_factor=2.0, mode="nearest"))
        outs[l] = q(outc(l, prev))
    return outs


def backbone(sd, x):
    """FCOS backbone: trunk -> FPN p3..p5 + LastLevelP6P7(p5) (backbone/fpn.py:11-78)."""
    o = fpn_topdown(sd, trunk(sd, x), (5, 4, 3))
    p5, p4, p3 = o[5], o[4], o[3]
    p6 = q(conv(p5, sd["backbone.top_block.p6.weight"], sd["backbone.top_block.p6.bias"], 2, 1))
    p7 = q(conv(F.relu(p6), sd["backbone.top_block.p7.weight"], sd["backbone.top_block.p7.bias"], 2, 1))
    return [p3, p4, p5, p6, p7]


def fcos_head(sd, feats):
    """fcos/fcos.py:338-376."""
    hd = "proposal_generator.fcos_head."
    logits, reg, std, ctr = [], [], [], []
    for l, f in enumerate(feats):
        towers = {}
        for t in ("cls_tower", "bbox_tower"):
            x = f
            for i in range(4):
                x = q(conv(x, sd[f"{hd}{t}.{3 * i}.weight"], sd[f"{hd}{t}.{3 * i}.bias"], 1, 1))
                x = q(F.relu(F.group_norm(x, 32, sd[f"{hd}{t}.{3 * i + 1}.weight"], sd[f"{hd}{t}.{3 * i + 1}.bias"], 1e-5)))
            towers[t] = x
        # the predictors' outputs are bf16 tensors; Scale_l is applied to the rounded value inside the loss / decode kernels
        c = lambda n, x: q(conv(x, sd[hd + n + ".weight"], sd[hd + n + ".bias"], 1, 1))
        logits.append(c("cls_logits", towers["cls_tower"]))
        ctr.append(c("ctrness", towers["bbox_tower"]))
        reg.append(c("bbox_pred", towers["bbox_tower"]) * sd[f"{hd}scales.{l}.scale"])
        std.append(c("bbox_pred_std", towers["bbox_tower"]))
    return logits, reg, std, ctr


def preprocess(sd, images, div=32):
    """one_stage_detector.py:165-167: (x - mean) / std then zero-pad to the batch max rounded up to /32."""
    mean, std = sd["pixel_mean"].view(3, 1, 1), sd["pixel_std"].view(3, 1, 1)
    xs = [(im.float() - mean) / std for im in images]
    H = (max(x.shape[1] for x in xs) + div - 1) // div * div
    W = (max(x.shape[2] for x in xs) + div - 1) // div * div
    out = torch.zeros(len(xs), 3, H, W)
    for i, x in enumerate(xs):
        out[i, :, :x.shape[1], :x.shape[2]] = x
    return out, [(x.shape[1], x.shape[2]) for x in xs]


def forward_dense(sd, images):
    x, sizes = preprocess(sd, images)
    feats = backbone(sd, x)
    logits, reg, std, ctr = fcos_head(sd, feats)
    locs = [O.compute_locations(f.shape[2], f.shape[3], s) for f, s in zip(feats, (8, 16, 32, 64, 128))]
    return {"logits": logits, "reg": reg, "std": std, "ctr": ctr, "locations": locs, "image_sizes": sizes}


def trainable_keys(sd):
    """FREEZE_AT=2: stem + res2 frozen; FrozenBN / pixel stats / integral.project are buffers."""
    keys = []
    for k in sd:
        if ".norm." in k or k in ("pixel_mean", "pixel_std") or k.endswith("integral.project"):
            continue
        if k.startswith("backbone.bottom_up.stem.") or k.startswith("backbone.bottom_up.res2."):
            continue
        keys.append(k)
    return keys


def is_norm_param(k):
    hd = "proposal_generator.fcos_head."
    return k.startswith(hd) and ("_tower." in k) and int(k.split(".")[3]) % 3 == 1


def ut2_step(student, teacher, mom, batch, cfg, first_step, device_type="cuda", sets_override=None):
    """One post-burn-in UT2 FCOS step (trainer.py:181-429) on CPU in fp32.
    student / teacher: dict name -> tensor (updated in place); mom: dict of momentum buffers;
    batch: (label_q, label_k, unlabel_q, unlabel_k) lists of {"image": u8 CHW, "boxes", "classes"}.
    Returns the record dict of loss scalars."""
    lam, mu = cfg["UNSUP_LOSS_WEIGHT"], cfg["UNSUP_REG_LOSS_WEIGHT"]
    keep = cfg["EMA_KEEP_RATE"] if not cfg.get("copy_teacher") else 0.0
    for k in teacher:                                             # EMA over the whole state_dict (:468-486)
        teacher[k] = O.ema_update(student[k].reshape(-1), teacher[k].reshape(-1), keep).reshape(teacher[k].shape)
    label_q, label_k, unlabel_q, unlabel_k = batch
    with torch.no_grad():
        t = forward_dense(teacher, [d["image"] for d in unlabel_k])
        pt = O.fcos_predict_proposals(t["logits"], t["reg"], t["std"], t["ctr"], t["locations"], t["image_sizes"],
                                      cfg["NMS_CRITERIA_TRAIN"], device_type=device_type)
        ptl = O.fcos_predict_proposals(t["logits"], t["reg"], t["std"], t["ctr"], t["locations"], t["image_sizes"],
                                       cfg["NMS_CRITERIA_REG_TRAIN"], device_type=device_type)
        sets = []
        for preds, thr in ((pt, cfg["BBOX_THRESHOLD"]), (ptl, cfg["BBOX_THRESHOLD_REG"])):
            th = [O.threshold_bbox(p, thr) for p in preds]
            sets.append({"boxes": [x["gt_boxes"] for x in th], "classes": [x["gt_classes"] for x in th],
                         "scores": [x["scores"] for x in th], "reg_pred_std": [x["reg_pred_std"] for x in th]})
    if sets_override is not None:       # parity tests inject the device's pseudo-label sets (threshold borderlines)
        sets = sets_override
    tk = trainable_keys(student)
    params = {k: student[k].detach().clone().requires_grad_(True) for k in tk}
    sd = dict(student)
    sd.update(params)
    lab = label_q + label_k
    s = forward_dense(sd, [d["image"] for d in lab])
    sup, _ = O.fcos_losses_labeled(s["logits"], s["reg"], s["std"], s["ctr"], s["locations"], [d["boxes"] for d in lab],
                                   [d["classes"] for d in lab])
    u = forward_dense(sd, [d["image"] for d in unlabel_q])
    uns = O.fcos_losses_pseudo(u["logits"], u["reg"], u["std"], u["ctr"], u["locations"], sets[0], sets[1])
    record = dict(sup)
    record.update({k + "_pseudo": v for k, v in uns.items()})
    total = (sup["loss_fcos_cls"] + sup["loss_fcos_ctr"]) / (lam + 1) + sup["loss_fcos_loc"] / (mu + 1) + \
        (uns["loss_fcos_cls"] + uns["loss_fcos_ctr"]) * lam / (lam + 1) + uns["loss_fcos_loc"] * mu / (mu + 1)
    total.backward()
    grads = {}
    for k in tk:                                                  # [D2] SGD: momentum .9, wd 1e-4 (0 for norm layers)
        g = params[k].grad if params[k].grad is not None else torch.zeros_like(params[k])
        grads[k] = g
        wd = cfg["WEIGHT_DECAY_NORM"] if is_norm_param(k) else cfg["WEIGHT_DECAY"]
        p, b = O.sgd_step(student[k], g, mom.get(k, torch.zeros_like(g)), cfg["LR"], cfg["MOMENTUM"], wd, first_step)
        student[k], mom[k] = p.detach(), b.detach()
    record = {k: (v.detach() if isinstance(v, torch.Tensor) else torch.tensor(float(v))) for k, v in record.items()}
    return record, grads, sets
