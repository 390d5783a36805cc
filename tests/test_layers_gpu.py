"""``ubteacher.layers`` (IOULoss, NLLoss, KLLoss, ml_nms) on the device against the oracle restatements of
ubteacher/layers/*.py (pinned by tests/golden/loss_pieces.pt) and torchvision's batched_nms: values 1e-5, gradients 1e-4
(fp32 arithmetic on both sides), keep lists exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ltrb(g, n):
    return torch.rand(n, 4, generator=g) * 60 + 1


@pytest.mark.parametrize("kind", ["iou", "linear_iou", "giou"])
@pytest.mark.parametrize("weighted", [False, True])
def test_iou_loss(kind, weighted):
    from oracle import ut2_oracle as O
    from ubteacher.layers import IOULoss
    g = torch.Generator().manual_seed(3)
    pred, tgt = _ltrb(g, 777), _ltrb(g, 777)
    w = torch.rand(777, generator=g) if weighted else None
    pr = pred.clone().requires_grad_(True)
    ref = O.iou_loss(pr, tgt, w, kind)
    ref.backward()
    pd = pred.cuda().requires_grad_(True)
    got = IOULoss(kind)(pd, tgt.cuda(), w.cuda() if weighted else None)
    (got * 0.7).backward()
    torch.testing.assert_close(got.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(pd.grad.cpu(), 0.7 * pr.grad, rtol=1e-4, atol=1e-6)
    with pytest.raises(NotImplementedError):
        IOULoss("diou")(pd, tgt.cuda())


def test_nl_loss():
    from oracle import ut2_oracle as O
    from ubteacher.layers import NLLoss
    g = torch.Generator().manual_seed(4)
    mu, sd, tgt, iw = torch.randn(500, 4, generator=g) * 3, torch.randn(500, 4, generator=g), torch.randn(500, 4, generator=g) * 3, torch.rand(500, generator=g)
    a, b = mu.clone().requires_grad_(True), sd.clone().requires_grad_(True)
    ref = O.nl_loss_fcos(a, b, tgt, iw)
    ref.backward()
    x, s = mu.cuda().requires_grad_(True), sd.cuda().requires_grad_(True)
    got = NLLoss()(x, s, tgt.cuda(), iou_weight=iw.cuda())
    got.backward()
    torch.testing.assert_close(got.detach().cpu(), ref.detach().reshape(()), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad.cpu(), a.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(s.grad.cpu(), b.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("method", ["weight_ctr_sum", "weight_ctr_mean", "sum", "mean"])
def test_kl_loss(method):
    from oracle import ut2_oracle as O
    from ubteacher.layers import KLLoss
    g = torch.Generator().manual_seed(5)
    x0, sd, tgt, w = torch.randn(300, 4, generator=g) * 2, torch.randn(300, 4, generator=g), torch.randn(300, 4, generator=g) * 2, torch.rand(300, generator=g)
    a, b = x0.clone().requires_grad_(True), sd.clone().requires_grad_(True)
    ref = O.kl_loss(a, b, tgt, w, 1.0, 7.5, method)
    ref.backward()
    x, s = x0.cuda().requires_grad_(True), sd.cuda().requires_grad_(True)
    got = KLLoss()(x, s, tgt.cuda(), weight=w.cuda(), beta=1.0, loss_denorm=7.5, method=method)
    got.backward()
    torch.testing.assert_close(got.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad.cpu(), a.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(s.grad.cpu(), b.grad, rtol=1e-4, atol=1e-7)
    with pytest.raises(NotImplementedError):
        KLLoss()(x, s, tgt.cuda(), weight=w.cuda(), beta=0.0)
    with pytest.raises(ValueError):
        KLLoss()(x, s, tgt.cuda(), weight=w.cuda(), method="median")


def test_ml_nms_equals_torchvision_batched_nms():
    import torchvision
    from ubteacher.d2compat.structures import Boxes, Instances
    from ubteacher.layers import ml_nms
    g = torch.Generator().manual_seed(6)
    n = 1500
    xy = torch.rand(n, 2, generator=g) * 300
    wh = torch.rand(n, 2, generator=g) * 80 + 4
    inst = Instances((400, 400))
    inst.pred_boxes = Boxes(torch.cat([xy, xy + wh], 1).cuda())
    inst.scores = torch.rand(n, generator=g).cuda()
    inst.pred_classes = torch.randint(0, 5, (n,), generator=g).cuda()
    out = ml_nms(inst, 0.6)
    keep = torchvision.ops.batched_nms(inst.pred_boxes.tensor.cpu(), inst.scores.cpu(), inst.pred_classes.cpu(), 0.6)
    assert len(out) == len(keep)
    assert torch.equal(out.scores.cpu(), inst.scores.cpu()[keep]) and torch.equal(out.pred_boxes.tensor.cpu(), inst.pred_boxes.tensor.cpu()[keep])
    top = ml_nms(inst, 0.6, max_proposals=50)
    assert len(top) == 50 and torch.equal(top.scores.cpu(), inst.scores.cpu()[keep[:50]])
    assert ml_nms(inst, 0.0) is inst
