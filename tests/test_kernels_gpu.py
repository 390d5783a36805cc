"""GPU parity of the bandwidth kernels against the CPU oracle / golden fixtures (through the C-ABI).
Integer / index / byte results: bit-exact. Floating point: tolerance stated at each check."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LEVEL_HW = [(16, 20), (8, 10), (4, 5), (2, 3), (1, 2)]
STRIDES = [8, 16, 32, 64, 128]


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def rb(x):  # round to bf16 and back: what the device path sees
    return x.bfloat16().float()


def geom():
    from ubteacher import ops
    return ops.LevelGeom(LEVEL_HW, STRIDES, [64, 128, 256, 512])


def locs():
    from oracle import ut2_oracle as O
    return [O.compute_locations(h, w, s) for (h, w), s in zip(LEVEL_HW, STRIDES)]


def pack_head(logits, reg, std, ctr):
    cls = torch.cat([x.permute(0, 2, 3, 1).reshape(-1, 80) for x in logits]).bfloat16()
    box = torch.cat([torch.cat([r.permute(0, 2, 3, 1).reshape(-1, 68), s.permute(0, 2, 3, 1).reshape(-1, 4),
                                c.permute(0, 2, 3, 1).reshape(-1, 1),
                                torch.zeros(r.shape[0] * r.shape[2] * r.shape[3], 7)], 1)
                     for r, s, c in zip(reg, std, ctr)]).bfloat16()
    return cls.cuda(), box.cuda()


def pack_gt(boxes, classes, std=None, G=16):
    N = len(boxes)
    b = torch.zeros(N, G, 4)
    c = torch.zeros(N, G, dtype=torch.int64)
    s = torch.zeros(N, G, 4)
    cnt = torch.zeros(N, dtype=torch.int32)
    for i in range(N):
        n = len(boxes[i])
        b[i, :n], c[i, :n], cnt[i] = boxes[i], classes[i], n
        if std is not None:
            s[i, :n] = std[i]
    return b.cuda(), c.cuda(), cnt.cuda(), (s.cuda() if std is not None else None)


# --------------------------------------------------------------------------------------- EMA / SGD
def test_ema_bit_exact():
    from ubteacher import ops
    g = load("ema.pt")
    for keep in [0.0, 0.9996, 0.9999]:
        t = g["teacher"].clone().cuda()
        ops.ema_update(g["student"].cuda(), t, keep)
        assert torch.equal(t.cpu(), g[f"keep{keep}"])
    # odd length + properties at full size: keep=1 is the identity, keep=0 copies
    n = 32_400_003
    s = torch.randn(n, device="cuda")
    t = torch.randn(n, device="cuda")
    t0 = t.clone()
    ops.ema_update(s, t, 1.0)
    assert torch.equal(t, t0)
    ops.ema_update(s, t, 0.0)
    assert torch.equal(t, s)


def test_sgd_matches_oracle():
    from oracle import ut2_oracle as O
    from ubteacher import ops
    gen = torch.Generator().manual_seed(0)
    p, g, buf = torch.randn(10007, generator=gen), torch.randn(10007, generator=gen), torch.randn(10007, generator=gen)
    for first in (True, False):
        pc, gc, bc = p.clone().cuda(), g.clone().cuda(), buf.clone().cuda()
        ops.sgd_step(pc, gc, bc, 0.01, 0.9, 1e-4, first, zero_grad=True)
        pr, br = O.sgd_step(p, g, buf, 0.01, 0.9, 1e-4, first)
        torch.testing.assert_close(pc.cpu(), pr, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(bc.cpu(), br, rtol=1e-6, atol=1e-7)
        assert float(gc.abs().max()) == 0.0


# --------------------------------------------------------------------------------------- elementwise
def test_groupnorm_relu_fwd_bwd():
    from ubteacher import ops
    gen = torch.Generator().manual_seed(1)
    N, H, W, C = 3, 13, 21, 256
    x = (torch.randn(N, H, W, C, generator=gen) * 2 + 0.3).bfloat16()
    gamma = torch.rand(C, generator=gen) + 0.5
    beta = torch.randn(C, generator=gen) * 0.2
    dy = torch.randn(N, H, W, C, generator=gen).bfloat16()
    y, stats = ops.groupnorm_relu_fwd(x.cuda(), gamma.cuda(), beta.cuda())
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.relu(F.group_norm(xr, 32, gr, br, 1e-5))
    torch.testing.assert_close(y.float().cpu(), yr.permute(0, 2, 3, 1), rtol=1e-2, atol=1e-2)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    dgam = torch.zeros(C, device="cuda")
    dbet = torch.zeros(C, device="cuda")
    dbias = torch.zeros(C, device="cuda")
    dx = ops.groupnorm_relu_bwd(dy.cuda(), x.cuda(), stats, gamma.cuda(), beta.cuda(), dgam, dbet, dbias_prev=dbias)
    torch.testing.assert_close(dbias.cpu(), dx.float().cpu().sum((0, 1, 2)), rtol=1e-4, atol=1e-3)   # fused bias grad
    torch.testing.assert_close(dx.float().cpu(), xr.grad.permute(0, 2, 3, 1), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dgam.cpu(), gr.grad, rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(dbet.cpu(), br.grad, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("tensor_core", [False, True])
def test_stem_maxpool(tensor_core):
    from ubteacher import ops
    gen = torch.Generator().manual_seed(2)
    h, w = 67, 293
    Hp, Wp = 96, 320
    img = torch.randint(0, 256, (3, h, w), generator=gen, dtype=torch.uint8)
    wgt = torch.randn(64, 3, 7, 7, generator=gen) * 0.05
    scale = torch.rand(64, generator=gen) + 0.5
    shift = torch.randn(64, generator=gen) * 0.1
    mean, std = [103.53, 116.28, 123.675], [1.0, 1.0, 1.0]
    P, Q = Hp // 2, Wp // 2
    out = torch.empty(1, P, Q, 64, dtype=torch.bfloat16, device="cuda")
    ops.stem_conv(img.cuda(), wgt.permute(2, 3, 1, 0).contiguous().cuda(), scale.cuda(), shift.cuda(), mean, std,
                  out[0], P, Q, tensor_core=tensor_core)
    x = (img.float() - torch.tensor(mean).view(3, 1, 1)) / torch.tensor(std).view(3, 1, 1)
    xp = torch.zeros(1, 3, Hp, Wp)
    xp[0, :, :h, :w] = x
    ref = F.relu(F.conv2d(xp, wgt, stride=2, padding=3) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    # tensor-core path rounds the normalised pixels and the filter to bf16 (2^-9 each) before the fp32 accumulate
    torch.testing.assert_close(out.float().cpu(), ref.permute(0, 2, 3, 1), rtol=1e-2, atol=1.5 if tensor_core else 5e-2)
    assert float((out.float().cpu() - ref.permute(0, 2, 3, 1)).norm() / ref.norm()) < (6e-3 if tensor_core else 3e-3)
    mp = ops.maxpool3x3s2(out)
    refp = F.max_pool2d(out.float().cpu().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(mp.float().cpu(), refp)


def test_stem_batched_ragged():
    """One launch over a batch of differently sized images (ImageList padding = zeros after normalisation), odd P / Q."""
    from ubteacher import ops
    gen = torch.Generator().manual_seed(12)
    sizes = [(67, 293), (90, 250), (33, 300)]
    Hp, Wp = 102, 302            # P = 51 (odd: ragged row pair), Q = 151 (ragged 64-pixel column tile)
    imgs = [torch.randint(0, 256, (3, h, w), generator=gen, dtype=torch.uint8) for h, w in sizes]
    wgt = torch.randn(64, 3, 7, 7, generator=gen) * 0.05
    scale = torch.rand(64, generator=gen) + 0.5
    shift = torch.randn(64, generator=gen) * 0.1
    mean, std = [103.53, 116.28, 123.675], [57.375, 57.12, 58.395]
    P, Q = Hp // 2, Wp // 2
    out = torch.full((len(imgs), P, Q, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.stem_conv_batched([i.cuda() for i in imgs], wgt.permute(2, 3, 1, 0).contiguous().cuda(), scale.cuda(), shift.cuda(),
                          mean, std, out, P, Q)
    xp = torch.zeros(len(imgs), 3, Hp, Wp)
    for i, (im, (h, w)) in enumerate(zip(imgs, sizes)):
        xp[i, :, :h, :w] = (im.float() - torch.tensor(mean).view(3, 1, 1)) / torch.tensor(std).view(3, 1, 1)
    ref = F.relu(F.conv2d(xp, wgt, stride=2, padding=3) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    assert ref.shape == out.shape and torch.isfinite(out.float()).all()
    assert float((out.float().cpu() - ref).norm() / ref.norm()) < 6e-3
    torch.testing.assert_close(out.float().cpu(), ref, rtol=1e-2, atol=5e-2)


def test_fpn_and_grad_helpers():
    from ubteacher import ops
    gen = torch.Generator().manual_seed(4)
    lat = torch.randn(2, 10, 14, 64, generator=gen).bfloat16()
    top = torch.randn(2, 5, 7, 64, generator=gen).bfloat16()
    out = ops.upsample2x_add(lat.cuda(), top.cuda())
    up = top.float().repeat_interleave(2, 1).repeat_interleave(2, 2)
    assert torch.equal(out.cpu(), (lat.float() + up).bfloat16())
    add = torch.randn(2, 5, 7, 64, generator=gen).bfloat16()
    ds = ops.downsample2x_sum(lat.cuda(), add.cuda())
    ref = lat.float().reshape(2, 5, 2, 7, 2, 64).sum((2, 4)) + add.float()
    torch.testing.assert_close(ds.float().cpu(), ref, rtol=1e-2, atol=1e-2)
    y = torch.randn(2, 10, 14, 64, generator=gen).relu().bfloat16()
    g = ops.relu_bwd(lat.cuda(), y.cuda())
    assert torch.equal(g.cpu(), torch.where(y.float() > 0, lat.float(), torch.zeros(())).bfloat16())
    z = ops.zero_stuff_s2(top.cuda(), 10, 14, 1, 1)
    refz = torch.zeros(2, 10, 14, 64)
    refz[:, 1::2, 1::2] = top.float()
    assert torch.equal(z.float().cpu(), refz)
    db = torch.zeros(64, device="cuda")
    ops.colsum(lat.reshape(-1, 64).cuda(), db)
    torch.testing.assert_close(db.cpu(), lat.float().reshape(-1, 64).sum(0), rtol=1e-4, atol=1e-3)
    s = ops.add_bf16(lat.cuda(), y.cuda())
    assert torch.equal(s.cpu(), (lat.float() + y.float()).bfloat16())


def test_pack_weight():
    from ubteacher import ops
    gen = torch.Generator().manual_seed(5)
    Cout, Cin, R = 24, 64, 3
    w = torch.randn(Cout, Cin, R, R, generator=gen)
    master = w.permute(0, 2, 3, 1).contiguous().cuda()     # channels-last physical layout
    wf = torch.zeros(Cout, R, R, Cin, dtype=torch.bfloat16, device="cuda")
    wt = torch.zeros(Cin, R, R, 32, dtype=torch.bfloat16, device="cuda")
    ops.pack_conv_weight(master, wf, wt, Cout, Cin, R, R, 32)
    assert torch.equal(wf.cpu(), w.permute(0, 2, 3, 1).bfloat16())
    ref_t = w.flip(2, 3).permute(1, 2, 3, 0).bfloat16()          # [Cin, R, S, Cout]
    assert torch.equal(wt.cpu()[..., :Cout], ref_t)
    assert float(wt.cpu()[..., Cout:].float().abs().max()) == 0.0


def test_pack_plan_batched():
    """Table-driven tiled packer (PackPlan / ut2_pack_conv_weights_batched): ragged channel counts, a FrozenBN scale
    folded in, fused predictors sharing one dgrad operand, and write_dgrad=0 leaving the transposed operands alone."""
    from ubteacher.arena import PackPlan, ParamArena, Spec
    gen = torch.Generator().manual_seed(9)
    shapes = {"a.weight": (100, 72, 3, 3), "b.weight": (24, 64, 1, 1), "p0.weight": (3, 40, 3, 3), "p1.weight": (12, 40, 3, 3)}
    A = ParamArena([Spec(n, sh, "decay") for n, sh in shapes.items()], torch.device("cuda"))
    w = {n: torch.randn(sh, generator=gen) for n, sh in shapes.items()}
    for n in shapes:
        A.views[n].copy_(w[n].cuda())
    scales = torch.rand(128, generator=gen) + 0.5
    plan = PackPlan(A)
    offs = {}
    for n, coutT in (("a.weight", 104), ("b.weight", 24)):
        co, ci, R, _ = shapes[n]
        offs[n] = (plan.alloc(co * ci * R * R), plan.alloc(ci * R * R * coutT) if n == "a.weight" else -1, coutT)
        plan.add(n, offs[n][0], offs[n][1], co, ci, R, R, coutT, scale_off=8 if n == "a.weight" else -1)
    wf_f, wt_f = plan.alloc(16 * 9 * 40), plan.alloc(40 * 9 * 16)
    plan.add("p0.weight", wf_f, wt_f, 3, 40, 3, 3, 16, 0)
    plan.add("p1.weight", wf_f + 3 * 360, wt_f, 12, 40, 3, 3, 16, 3)
    plan.finalize()
    plan.scales = scales.cuda()
    plan.run(dgrad=False)
    assert float(plan.view(offs["a.weight"][1], (72, 3, 3, 104)).float().abs().max()) == 0.0
    plan.run()
    wa = w["a.weight"] * scales[8:108].view(-1, 1, 1, 1)
    assert torch.equal(plan.view(offs["a.weight"][0], (100, 3, 3, 72)).cpu(), wa.permute(0, 2, 3, 1).bfloat16())
    ta = plan.view(offs["a.weight"][1], (72, 3, 3, 104)).cpu()
    assert torch.equal(ta[..., :100], wa.flip(2, 3).permute(1, 2, 3, 0).bfloat16()) and float(ta[..., 100:].float().abs().max()) == 0
    assert torch.equal(plan.view(offs["b.weight"][0], (24, 1, 1, 64)).cpu(), w["b.weight"].permute(0, 2, 3, 1).bfloat16())
    wp = torch.cat([w["p0.weight"], w["p1.weight"]])
    assert torch.equal(plan.view(wf_f, (15, 3, 3, 40)).cpu(), wp.permute(0, 2, 3, 1).bfloat16())
    tp = plan.view(wt_f, (40, 3, 3, 16)).cpu()
    assert torch.equal(tp[..., :15], wp.flip(2, 3).permute(1, 2, 3, 0).bfloat16()) and float(tp[..., 15].float().abs().max()) == 0


# --------------------------------------------------------------------------------------- FCOS targets / losses
@pytest.mark.parametrize("name", ["fcos_targets_labeled.pt", "fcos_targets_pseudo.pt"])
def test_assign_targets_golden(name):
    from ubteacher import ops
    g = load(name)
    b, c, cnt, s = pack_gt(g["boxes"], g["classes"], g.get("reg_pred_std"))
    out = ops.fcos_assign_targets(geom(), len(g["boxes"]), b, c, cnt, s)
    ref = {k: torch.cat([t.reshape(t.shape[0], -1) if t.dim() > 1 else t for t in v]) for k, v in g["targets"].items()}
    assert torch.equal(out["labels"].cpu(), ref["labels"])
    assert torch.equal(out["target_inds"].cpu(), ref["target_inds"])
    assert torch.equal(out["reg_targets"].cpu(), ref["reg_targets"])
    assert torch.equal(out["boundary_vars"].cpu(), ref["boundary_vars"].float())
    assert torch.equal(out["keep_locations"].cpu().bool(), ref["keep_locations"])


def test_assign_targets_full_size_vs_oracle():
    from oracle import ut2_oracle as O
    from ubteacher import ops
    hw = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    gm = ops.LevelGeom(hw, STRIDES, [64, 128, 256, 512])
    gen = torch.Generator().manual_seed(6)
    N = 4
    boxes, classes = [], []
    for i in range(N):
        n = [7, 0, 30, 100][i]
        xy = torch.rand(n, 2, generator=gen) * torch.tensor([1000.0, 600.0])
        wh = torch.exp(torch.rand(n, 2, generator=gen) * 2.77 + 3.46)
        boxes.append(torch.cat([xy, xy + wh], 1))
        classes.append(torch.randint(0, 80, (n,), generator=gen))
    b, c, cnt, _ = pack_gt(boxes, classes, None, G=128)
    out = ops.fcos_assign_targets(gm, N, b, c, cnt, None)
    L = [O.compute_locations(h, w, s) for (h, w), s in zip(hw, STRIDES)]
    ref = O.fcos_assign_targets_fast(L, boxes, classes, STRIDES)
    assert torch.equal(out["labels"].cpu(), torch.cat(ref["labels"]))
    assert torch.equal(out["target_inds"].cpu(), torch.cat(ref["target_inds"]))
    assert torch.equal(out["reg_targets"].cpu(), torch.cat(ref["reg_targets"]))
    pos = torch.cat(ref["labels"]) != 80
    ctr = O.ctrness_targets(torch.cat(ref["reg_targets"])[pos])
    torch.testing.assert_close(out["norm"].cpu(), torch.stack([pos.sum().float(), ctr.sum()]), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("ignore_near", [False, True])
@pytest.mark.parametrize("which", ["labeled", "pseudo"])
def test_assign_targets_center_sample_golden(which, ignore_near):
    """MODEL.FCOS.CENTER_SAMPLE True / POS_RADIUS 1.5 (get_sample_region) and the ignore_near keep mask: bit-exact against
    the reference's _get_ground_truth on the two golden box sets."""
    from ubteacher import ops
    g = load(f"fcos_targets_{which}.pt")
    ref = load("fcos_targets_center_sample.pt")
    tg = ref[f"{which}_ignore_near{int(ignore_near)}"]
    b, c, cnt, s = pack_gt(g["boxes"], g["classes"], g.get("reg_pred_std"))
    out = ops.fcos_assign_targets(geom(), len(g["boxes"]), b, c, cnt, s, center_radius=ref["radius"], ignore_near=ignore_near)
    cat = {k: torch.cat([t.reshape(t.shape[0], -1) if t.dim() > 1 else t for t in v]) for k, v in tg.items()}
    assert torch.equal(out["labels"].cpu(), cat["labels"])
    assert torch.equal(out["target_inds"].cpu(), cat["target_inds"])
    assert torch.equal(out["reg_targets"].cpu(), cat["reg_targets"])
    assert torch.equal(out["boundary_vars"].cpu(), cat["boundary_vars"].float())
    assert torch.equal(out["keep_locations"].cpu().bool(), cat["keep_locations"])
    plain = load(f"fcos_targets_{which}.pt")["targets"]
    assert not torch.equal(cat["labels"], torch.cat(plain["labels"]))          # the option changes the assignment on this set
    if ignore_near:
        assert not bool(cat["keep_locations"].all())


def test_assign_targets_center_sample_full_size_vs_oracle():
    from oracle import ut2_oracle as O
    from ubteacher import ops
    hw = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    gm = ops.LevelGeom(hw, STRIDES, [64, 128, 256, 512])
    gen = torch.Generator().manual_seed(16)
    N = 3
    boxes, classes = [], []
    for i in range(N):
        n = [40, 0, 100][i]
        xy = torch.rand(n, 2, generator=gen) * torch.tensor([1000.0, 600.0])
        wh = torch.exp(torch.rand(n, 2, generator=gen) * 2.77 + 3.46)
        boxes.append(torch.cat([xy, xy + wh], 1))
        classes.append(torch.randint(0, 80, (n,), generator=gen))
    b, c, cnt, _ = pack_gt(boxes, classes, None, G=128)
    out = ops.fcos_assign_targets(gm, N, b, c, cnt, None, center_radius=1.5, ignore_near=True)
    L = [O.compute_locations(h, w, s) for (h, w), s in zip(hw, STRIDES)]
    ref = O.fcos_assign_targets_fast(L, boxes, classes, STRIDES, center_sample=True, radius=1.5, ignore_near=True)
    assert torch.equal(out["labels"].cpu(), torch.cat(ref["labels"]))
    assert torch.equal(out["target_inds"].cpu(), torch.cat(ref["target_inds"]))
    assert torch.equal(out["reg_targets"].cpu(), torch.cat(ref["reg_targets"]))
    assert torch.equal(out["keep_locations"].cpu().bool(), torch.cat(ref["keep_locations"]))
    assert 0 < int(out["keep_locations"].sum()) < out["keep_locations"].numel()


def _run_loss(g, mode, tg_kw, gout, scales=None, kl_mode=0):
    from ubteacher import ops
    N = g["logits"][0].shape[0]
    cls, box = pack_head(g["logits"], g["reg"], g["std"], g["ctr"])
    b, c, cnt, s = pack_gt(tg_kw["boxes"], tg_kw["classes"], tg_kw.get("reg_pred_std"))
    gm = geom()
    tg = ops.fcos_assign_targets(gm, N, b, c, cnt, s)
    sc = (scales if scales is not None else torch.ones(5)).cuda()
    losses, acc = ops.fcos_loss_fwd(gm, N, cls, box, sc, tg, mode, 0.25, 2.0, 0.05, 0.1, 0.8, 1.0, kl_mode=kl_mode)
    dcls = torch.full_like(cls, float("nan")) if mode != 2 else None
    dbox = torch.full_like(box, float("nan"))
    dsc = torch.zeros(5, device="cuda")
    ops.fcos_loss_bwd(gm, N, cls, box, sc, tg, mode, 0.25, 2.0, 0.05, 0.1, 0.8, 1.0, acc, gout.cuda(), dcls, dbox, dsc, kl_mode=kl_mode)
    return losses.cpu(), dcls, dbox, dsc.cpu()


def _unpack_grads(dcls, dbox, N):
    outs = {"logits": [], "reg": [], "std": [], "ctr": []}
    off = 0
    for (h, w) in LEVEL_HW:
        n = N * h * w
        if dcls is not None:
            outs["logits"].append(dcls[off:off + n].float().cpu().reshape(N, h, w, 80).permute(0, 3, 1, 2))
        blk = dbox[off:off + n].float().cpu().reshape(N, h, w, 80).permute(0, 3, 1, 2)
        outs["reg"].append(blk[:, :68])
        outs["std"].append(blk[:, 68:72])
        outs["ctr"].append(blk[:, 72:73])
        off += n
    return outs


def test_fcos_losses_labeled_vs_oracle_and_golden():
    from oracle import ut2_oracle as O
    g = load("fcos_losses_labeled.pt")
    losses, dcls, dbox, _ = _run_loss(g, 0, {"boxes": g["boxes"], "classes": g["classes"]}, torch.tensor([1.0, 2.0, 3.0, 0.0]))
    # oracle on the bf16-rounded inputs the device saw: fp32 tolerance
    leaf = lambda ts: [rb(t).requires_grad_(True) for t in ts]
    L, R, S, C = leaf(g["logits"]), leaf(g["reg"]), leaf(g["std"]), leaf(g["ctr"])
    ref, _ = O.fcos_losses_labeled(L, R, S, C, locs(), g["boxes"], g["classes"])
    got = {"loss_fcos_cls": losses[0], "loss_fcos_loc": losses[1], "loss_fcos_ctr": losses[2]}
    for k in ref:
        torch.testing.assert_close(got[k], ref[k].detach(), rtol=2e-5, atol=1e-5)
        torch.testing.assert_close(got[k], g["losses"][k], rtol=2e-2, atol=1e-2)   # vs fp32-input reference
    (ref["loss_fcos_cls"] * 1 + ref["loss_fcos_loc"] * 2 + ref["loss_fcos_ctr"] * 3).backward()
    gr = _unpack_grads(dcls, dbox, 3)
    for name, leaves in (("logits", L), ("reg", R), ("std", S), ("ctr", C)):
        for a, b in zip(gr[name], leaves):
            scale = float(b.grad.abs().max()) + 1e-12
            torch.testing.assert_close(a, b.grad, rtol=1e-2, atol=1e-2 * scale)   # bf16 gradient storage


@pytest.mark.parametrize("method", ["mean", "sum", "weight_ctr_sum", "weight_ctr_mean"])
def test_fcos_losses_labeled_klloss(method):
    """MODEL.FCOS.KL_LOSS_TYPE "klloss" (config.py:198 default, not used by the shipped recipes) with the four LOC_FUN_ALL
    reductions: the fused loss kernels against the oracle on the device's inputs and against the reference's numbers."""
    from oracle import ut2_oracle as O
    from ubteacher import ops
    g, gold = load("fcos_losses_labeled.pt"), load("fcos_losses_labeled_klloss.pt")[method]
    losses, dcls, dbox, _ = _run_loss(g, 0, {"boxes": g["boxes"], "classes": g["classes"]}, torch.tensor([1.0, 2.0, 3.0, 0.0]),
                                      kl_mode=ops.KL_MODES[method])
    leaf = lambda ts: [rb(t).requires_grad_(True) for t in ts]
    L, R, S, C = leaf(g["logits"]), leaf(g["reg"]), leaf(g["std"]), leaf(g["ctr"])
    ref, _ = O.fcos_losses_labeled(L, R, S, C, locs(), g["boxes"], g["classes"], kl_loss_type="klloss", loc_fun_all=method)
    got = {"loss_fcos_cls": losses[0], "loss_fcos_loc": losses[1], "loss_fcos_ctr": losses[2]}
    for k in ref:
        torch.testing.assert_close(got[k], ref[k].detach(), rtol=2e-5, atol=1e-5)
        torch.testing.assert_close(got[k], gold["losses"][k], rtol=2e-2, atol=1e-2)   # vs the reference on fp32 inputs
    (ref["loss_fcos_cls"] * 1 + ref["loss_fcos_loc"] * 2 + ref["loss_fcos_ctr"] * 3).backward()
    gr = _unpack_grads(dcls, dbox, 3)
    for name, leaves in (("reg", R), ("std", S)):
        for a, b in zip(gr[name], leaves):
            scale = float(b.grad.abs().max()) + 1e-12
            torch.testing.assert_close(a, b.grad, rtol=1e-2, atol=1e-2 * scale)   # bf16 gradient storage


def test_fcos_losses_labeled_no_positive():
    g = load("fcos_losses_labeled_empty.pt")
    empty = {"boxes": [torch.zeros(0, 4)] * 2, "classes": [torch.zeros(0, dtype=torch.long)] * 2}
    losses, dcls, dbox, _ = _run_loss(g, 0, empty, torch.tensor([1.0, 1.0, 1.0, 0.0]))
    assert float(losses[:3].abs().max()) == 0.0
    assert float(dcls.float().abs().max()) == 0.0 and float(dbox.float().abs().max()) == 0.0


def test_fcos_losses_pseudo_vs_oracle_and_golden():
    from oracle import ut2_oracle as O
    g = load("fcos_losses_pseudo.pt")
    leaf = lambda ts: [rb(t).requires_grad_(True) for t in ts]
    L, R, S, C = leaf(g["logits"]), leaf(g["reg"]), leaf(g["std"]), leaf(g["ctr"])
    ref = O.fcos_losses_pseudo(L, R, S, C, locs(), g["cls_set"], g["reg_set"])
    (ref["loss_fcos_cls"] * 1 + ref["loss_fcos_ctr"] * 2 + ref["loss_fcos_loc"] * 4).backward()
    lc, dcls, dbox_c, _ = _run_loss(g, 1, g["cls_set"], torch.tensor([1.0, 0.0, 2.0, 0.0]))
    lr_, _, dbox_r, _ = _run_loss(g, 2, g["reg_set"], torch.tensor([0.0, 4.0, 0.0, 0.0]))
    torch.testing.assert_close(lc[0], ref["loss_fcos_cls"].detach(), rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(lc[2], ref["loss_fcos_ctr"].detach(), rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(lr_[1], ref["loss_fcos_loc"].detach(), rtol=2e-5, atol=1e-5)
    assert float(lr_[3]) == float(ref["teacher_better_student"])
    torch.testing.assert_close(lc[0], g["losses"]["loss_fcos_cls"], rtol=2e-2, atol=1e-2)
    gc = _unpack_grads(dcls, dbox_c, 3)
    gr = _unpack_grads(None, dbox_r, 3)
    for a, b in zip(gc["logits"], L):
        torch.testing.assert_close(a, b.grad, rtol=1e-2, atol=1e-2 * float(b.grad.abs().max()))
    for a, b in zip(gc["ctr"], C):
        torch.testing.assert_close(a, b.grad, rtol=1e-2, atol=1e-2 * float(b.grad.abs().max()) + 1e-9)
    for a, b in zip(gr["reg"], R):
        torch.testing.assert_close(a, b.grad, rtol=1e-2, atol=1e-2 * float(b.grad.abs().max()) + 1e-9)
    for a in gr["std"]:
        assert float(a.abs().max()) == 0.0


def test_fcos_loss_scale_gradient():
    """The learnable per-level Scale (fcos.py:367) is applied inside the loss kernels; check d/dscale."""
    from oracle import ut2_oracle as O
    g = load("fcos_losses_labeled.pt")
    sc = torch.tensor([1.3, 0.8, 1.1, 0.9, 1.2])
    losses, _, _, dsc = _run_loss(g, 0, {"boxes": g["boxes"], "classes": g["classes"]}, torch.tensor([0.0, 1.0, 0.0, 0.0]), sc)
    s = sc.clone().requires_grad_(True)
    reg = [rb(r) * s[i] for i, r in enumerate(g["reg"])]
    ref, _ = O.fcos_losses_labeled([rb(t) for t in g["logits"]], reg, [rb(t) for t in g["std"]],
                                   [rb(t) for t in g["ctr"]], locs(), g["boxes"], g["classes"])
    ref["loss_fcos_loc"].backward()
    torch.testing.assert_close(losses[1], ref["loss_fcos_loc"].detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dsc, s.grad, rtol=1e-3, atol=1e-4)


# --------------------------------------------------------------------------------------- proposals / NMS
def _canon(res, n=None):
    f = {k: (v[:n] if n is not None else v) for k, v in res.items() if isinstance(v, torch.Tensor) and v.dim() >= 1}
    key = torch.stack([f["fpn_levels"].double(), f["locations"][:, 0].double(), f["locations"][:, 1].double(),
                       f["pred_classes"].double()], 1)
    order = sorted(range(len(key)), key=lambda i: tuple(key[i].tolist()))
    return {k: v[order] for k, v in f.items()}


@pytest.mark.parametrize("tag", ["small", "large"])
@pytest.mark.parametrize("method", ["cls", "cls_n_ctr", "cls_n_loc"])
def test_predict_proposals_vs_oracle(tag, method):
    from oracle import ut2_oracle as O
    from ubteacher import ops
    g = load(f"fcos_predict_{tag}.pt")
    cls, box = pack_head(g["logits"], g["reg"], g["std"], g["ctr"])
    out = ops.fcos_predict_proposals(geom(), 2, cls, box, torch.ones(5).cuda(), method, 0.05, 1000, 0.6, 100)
    ref = O.fcos_predict_proposals([rb(t) for t in g["logits"]], [rb(t) for t in g["reg"]], [rb(t) for t in g["std"]],
                                   [rb(t) for t in g["ctr"]], locs(), g["image_sizes"], method, device_type="cuda")
    for i in range(2):
        n = int(out["count"][i])
        assert n == len(ref[i]["scores"])
        a = _canon({k: v[i].cpu() for k, v in out.items() if k != "count"}, n)
        b = _canon(ref[i])
        assert torch.equal(a["pred_classes"], b["pred_classes"])          # the kept SET is exact
        assert torch.equal(a["fpn_levels"], b["fpn_levels"])
        assert torch.equal(a["locations"], b["locations"])
        for k in ["pred_boxes", "scores", "centerness", "cls_confid", "reg_pred_std"]:
            torch.testing.assert_close(a[k], b[k], rtol=1e-5, atol=1e-4)
        # rows come out in descending-score order like the reference's NMS output
        s = out["scores"][i, :n].cpu()
        assert bool((s[:-1] >= s[1:]).all())


def test_predict_rejects_ctr():
    from ubteacher import ops
    g = load("fcos_predict_small.pt")
    cls, box = pack_head(g["logits"], g["reg"], g["std"], g["ctr"])
    with pytest.raises(ValueError):
        ops.fcos_predict_proposals(geom(), 2, cls, box, torch.ones(5).cuda(), "ctr", 0.05, 1000, 0.6, 100)


def test_threshold_scatter_vs_oracle():
    from oracle import ut2_oracle as O
    from ubteacher import ops
    g = load("pseudo_threshold.pt")
    inst = g["inst"]
    n, cap = len(inst["scores"]), 128
    dets = {"pred_boxes": torch.zeros(2, cap, 4), "scores": torch.zeros(2, cap),
            "pred_classes": torch.zeros(2, cap, dtype=torch.int64), "centerness": torch.zeros(2, cap),
            "cls_confid": torch.zeros(2, cap), "reg_pred_std": torch.zeros(2, cap, 4)}
    for k in dets:
        dets[k][0, :n] = inst[k]
    dets = {k: v.cuda() for k, v in dets.items()}
    dets["count"] = torch.tensor([n, 0], dtype=torch.int32).cuda()
    for mode, thr, ref in ((0, (0.5, 0.0), O.threshold_bbox(inst, 0.5)),
                           (1, (0.5, 0.3), O.threshold_cls_ctr_bbox(inst, (0.5, 0.3)))):
        out = ops.threshold_scatter(dets, mode, *thr)
        m = int(out["count"][0])
        assert m == len(ref["scores"]) and int(out["count"][1]) == 0
        for a, b in (("pred_boxes", "gt_boxes"), ("pred_classes", "gt_classes"), ("scores", "scores"),
                     ("centerness", "centerness"), ("cls_confid", "cls_confid"), ("reg_pred_std", "reg_pred_std")):
            assert torch.equal(out[a][0, :m].cpu(), ref[b]), a


@pytest.mark.parametrize("N", [1, 4])
def test_groupnorm_levels_full_size_vs_torch(N):
    """The single-pass cooperative GroupNorm kernels (csrc/groupnorm.cu: chunks staged in shared memory, cross-CTA statistics
    through fp64 atomics + arrival counters) on the full-size level-major pyramid (N x 22 400 locations: up to 700 chunks, 129
    CTAs per level-0 slab) against torch's fp32 group_norm + relu, forward and backward."""
    import torch.nn.functional as F
    from ubteacher import ops
    hw = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    geom = ops.LevelGeom(hw, [8, 16, 32, 64, 128])
    g = torch.Generator().manual_seed(40 + N)
    x = (torch.randn(geom.L * N, 256, generator=g) * 1.5 + 0.3).bfloat16()
    dy = torch.randn(geom.L * N, 256, generator=g).bfloat16()
    gam = torch.rand(256, generator=g) + 0.5
    bet = torch.randn(256, generator=g) * 0.5
    y, stats = ops.groupnorm_relu_levels_fwd(x.cuda(), geom, N, gam.cuda(), bet.cuda())
    dgam, dbet, dbias = [torch.zeros(256, device="cuda") for _ in range(3)]
    dx = ops.groupnorm_relu_levels_bwd(dy.cuda(), x.cuda(), geom, N, stats, gam.cuda(), bet.cuda(), dgam, dbet, dbias_prev=dbias)
    torch.cuda.synchronize()
    rg, rb, rbias, off = torch.zeros(256), torch.zeros(256), torch.zeros(256), 0
    for h, w in hw:
        n = N * h * w
        xl = x[off:off + n].float().view(N, h * w, 256).permute(0, 2, 1).clone().requires_grad_(True)
        gp, bp = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
        yl = F.relu(F.group_norm(xl, 32, gp, bp, 1e-5))
        yl.backward(dy[off:off + n].float().view(N, h * w, 256).permute(0, 2, 1))
        ref_y = yl.detach().permute(0, 2, 1).reshape(n, 256)
        ref_dx = xl.grad.permute(0, 2, 1).reshape(n, 256)
        torch.testing.assert_close(y[off:off + n].float().cpu(), ref_y, rtol=1 / 128, atol=1e-2)
        # ReLU-mask borderlines (y == 0 within rounding) move single elements; compare in the L2 sense
        e = (dx[off:off + n].float().cpu() - ref_dx).norm() / ref_dx.norm()
        assert float(e) < 6e-3, (h, w, float(e))
        rg += gp.grad
        rb += bp.grad
        rbias += dx[off:off + n].float().cpu().sum(0)
        off += n
    for a, b in ((dgam, rg), (dbet, rb), (dbias, rbias)):
        e = (a.cpu() - b).norm() / b.norm()
        assert float(e) < 3e-3, float(e)


@pytest.mark.parametrize("sizes,pad", [([(64, 96), (50, 77)], (64, 96)), ([(160, 224), (128, 192), (157, 223)], (160, 224)),
                                       ([(800, 1333), (771, 1201)], (800, 1344))])
def test_stem_pool_fused_vs_torch_and_split_kernels(sizes, pad):
    """csrc/stem_pool.cu (normalise + space-to-depth, then the 7x7 s2 stem as a 4x4 s1 tcgen05 GEMM with FrozenBN / ReLU /
    3x3 s2 max-pool in the epilogue) against fp32 torch on the same bf16-rounded pixels and filter, and against the separate
    stem + max-pool kernels it replaces. Images smaller than the padded batch size exercise the ImageList zero padding; the
    full-size case runs 11 column blocks x 10 row blocks per image (warm-up tiles, ring reuse across work items)."""
    import torch.nn.functional as F
    from ubteacher import ops
    Hp, Wp = pad
    g = torch.Generator().manual_seed(Hp + len(sizes))
    imgs = [torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8) for h, w in sizes]
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.05
    scale = torch.rand(64, generator=g) * 0.1 + 0.02
    shift = torch.randn(64, generator=g) * 0.3
    mean, std = [103.53, 116.28, 123.675], [1.0, 1.0, 1.0]
    w_rsck = w.permute(2, 3, 1, 0).contiguous().cuda()
    dev = [im.cuda() for im in imgs]
    y = ops.stem_pool_batched(dev, w_rsck, scale.cuda(), shift.cuda(), mean, std, Hp, Wp)
    N = len(imgs)
    x1 = torch.empty((N, Hp // 2, Wp // 2, 64), dtype=torch.bfloat16, device="cuda")
    ops.stem_conv_batched(dev, w_rsck, scale.cuda(), shift.cuda(), mean, std, x1, Hp // 2, Wp // 2)
    y_split = ops.maxpool3x3s2(x1)
    torch.cuda.synchronize()
    assert y.shape == (N, Hp // 4, Wp // 4, 64)
    xin = torch.zeros(N, 3, Hp, Wp)
    for i, im in enumerate(imgs):
        xin[i, :, :im.shape[1], :im.shape[2]] = ((im.float() - torch.tensor(mean).view(3, 1, 1)) / torch.tensor(std).view(3, 1, 1)).bfloat16().float()
    ref = F.conv2d(xin, w.bfloat16().float(), None, 2, 3) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    ref = F.max_pool2d(F.relu(ref).bfloat16().float(), 3, 2, 1).permute(0, 2, 3, 1)
    torch.testing.assert_close(y.float().cpu(), ref, rtol=1 / 128, atol=1e-2)
    # same arithmetic, different fp32 accumulation order: at most one bf16 ulp on a few elements
    d = (y.float() - y_split.float()).abs()
    assert float(d.max()) <= 1 / 64 * float(y_split.float().abs().max()) and float((d > 0).float().mean()) < 0.02
