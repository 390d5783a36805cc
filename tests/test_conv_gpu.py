"""GPU parity of the tcgen05 implicit-GEMM convolution (fwd / dgrad-as-fwd / wgrad) against a plain
PyTorch fp32 reference of the same op (floating-point kernel: tolerance stated below)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# bf16 output rounding is 2^-9 relative; accumulation is fp32 in both paths.
RTOL, ATOL = 1.0 / 128, 2e-2


def _mk(N, H, W, Cin, Cout, R, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, H, W, Cin, generator=g).bfloat16()
    w = (torch.randn(Cout, R, R, Cin, generator=g) / (R * R * Cin) ** 0.5).bfloat16()
    return x, w


def _ref_fwd(x, w, stride, pad, scale=None, shift=None, res=None, relu=False):
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), stride=stride, padding=pad)
    y = y.permute(0, 2, 3, 1)
    if scale is not None:
        y = y * scale
    if shift is not None:
        y = y + shift
    if res is not None:
        y = y + res.float()
    if relu:
        y = y.relu()
    return y


CASES = [
    # N, H, W, Cin, Cout, R, stride, pad
    (2, 13, 21, 64, 64, 3, 1, 1),
    (1, 25, 42, 256, 256, 3, 1, 1),
    (3, 9, 11, 128, 512, 1, 1, 0),
    (2, 13, 21, 80, 256, 3, 1, 1),
    (2, 26, 42, 256, 128, 1, 2, 0),
    (2, 25, 42, 256, 256, 3, 2, 1),
    (2, 13, 21, 256, 80, 3, 1, 1),
    (1, 7, 11, 64, 16, 3, 1, 1),
    (1, 20, 20, 1024, 2048, 1, 1, 0),
]


def test_im2col_probe_semantics():
    from ubteacher import _C

    N, H, W, C = 2, 5, 7, 64
    x = torch.arange(N * H * W * C, dtype=torch.float32).reshape(N, H, W, C) % 251
    xb = x.bfloat16().cuda()
    pixels = 64
    for (stride, pad, R) in [(1, 1, 3), (2, 1, 3), (1, 0, 1), (2, 0, 1)]:
        P = (H + 2 * pad - R) // stride + 1
        Q = (W + 2 * pad - R) // stride + 1
        for (m0, r, s) in [(0, 0, 0), (3, 1, 2 % R), (Q + 2, R - 1, R - 1), (P * Q - 3, 0, R - 1)]:
            img, rem = divmod(m0, P * Q)
            p0, q0 = divmod(rem, Q)
            out = torch.zeros(pixels * 128, dtype=torch.uint8, device="cuda")
            _C.call("ut2_debug_im2col_probe", xb, N, H, W, C, R, R, stride, pad, pixels, 0,
                    q0 * stride - pad, p0 * stride - pad, img, s, r, out)
            torch.cuda.synchronize()
            raw = out.cpu().view(torch.bfloat16).reshape(pixels, 8, 8)
            got = torch.empty_like(raw)
            for row in range(pixels):
                for j in range(8):
                    got[row, j] = raw[row, j ^ (row % 8)]
            got = got.reshape(pixels, 64).float()
            exp = torch.zeros(pixels, 64)
            for i in range(pixels):
                m = m0 + i
                if m >= N * P * Q:
                    continue
                im, rem = divmod(m, P * Q)
                p, q = divmod(rem, Q)
                hh, ww = p * stride - pad + r, q * stride - pad + s
                if 0 <= hh < H and 0 <= ww < W:
                    exp[i] = xb[im, hh, ww].float().cpu()
            assert torch.equal(got, exp), (stride, pad, R, m0, r, s)


@pytest.mark.parametrize("case", CASES)
def test_conv_fwd(case):
    from ubteacher import _C

    N, H, W, Cin, Cout, R, stride, pad = case
    x, w = _mk(N, H, W, Cin, Cout, R)
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    y = torch.full((N, P, Q, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    _C.call("ut2_conv2d_nhwc_bf16_fwd", x.cuda(), N, H, W, Cin, w.cuda(), Cout, R, R, stride, pad,
            None, None, None, 0, None, 0, y)
    torch.cuda.synchronize()
    ref = _ref_fwd(x, w, stride, pad)
    torch.testing.assert_close(y.float().cpu(), ref, rtol=RTOL, atol=ATOL)


def test_conv_fwd_epilogue():
    from ubteacher import _C

    N, H, W, Cin, Cout, R, stride, pad = 2, 13, 21, 128, 256, 3, 1, 1
    x, w = _mk(N, H, W, Cin, Cout, R, seed=3)
    g = torch.Generator().manual_seed(7)
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g)
    res = torch.randn(N, H, W, Cout, generator=g).bfloat16()
    y = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device="cuda")
    _C.call("ut2_conv2d_nhwc_bf16_fwd", x.cuda(), N, H, W, Cin, w.cuda(), Cout, R, R, stride, pad,
            scale.cuda(), shift.cuda(), res.cuda(), 0, None, 1, y)
    torch.cuda.synchronize()
    ref = _ref_fwd(x, w, stride, pad, scale, shift, res, True)
    torch.testing.assert_close(y.float().cpu(), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("case", CASES)
def test_conv_wgrad(case):
    from ubteacher import _C

    N, H, W, Cin, Cout, R, stride, pad = case
    if Cin % 64:
        pytest.skip("wgrad tiles input channels by 64 (ragged Cin only occurs as a dgrad input)")
    x, w = _mk(N, H, W, Cin, Cout, R, seed=11)
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    g = torch.Generator().manual_seed(5)
    dy = torch.randn(N, P, Q, Cout, generator=g).bfloat16()
    dw = torch.zeros(Cout, R, R, Cin, dtype=torch.float32, device="cuda")
    _C.call("ut2_conv2d_nhwc_bf16_wgrad", x.cuda(), N, H, W, Cin, dy.cuda(), Cout, R, R, stride, pad,
            None, dw, 0)
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wr = w.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=stride, padding=pad)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    ref = wr.grad.permute(0, 2, 3, 1)
    scale = ref.abs().max().item()
    torch.testing.assert_close(dw.cpu(), ref, rtol=2e-3, atol=2e-3 * scale)


# --------------------------------------------------------------------------------------- level-major launches
LEVEL_HW = [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)]


def _level_geom():
    from ubteacher import ops
    return ops.LevelGeom(LEVEL_HW, [8, 16, 32, 64, 128])


def _split(t, N, C):
    out, off = [], 0
    for h, w in LEVEL_HW:
        out.append(t[off:off + N * h * w].view(N, h, w, C))
        off += N * h * w
    return out


@pytest.mark.parametrize("cin,cout,aux", [(256, 256, None), (256, 80, None), (80, 256, "res"), (256, 256, "mask")])
def test_conv_levels_equals_per_level_launches(cin, cout, aux):
    """One launch over a level-major pyramid == five single-level launches of the same kernel (bit-exact)."""
    from ubteacher import ops
    geom, N = _level_geom(), 3
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(geom.L * N, cin, generator=g).bfloat16().cuda()
    w = (torch.randn(cout, 3, 3, cin, generator=g) / (9 * cin) ** 0.5).bfloat16().cuda()
    shift = torch.randn(cout, generator=g).cuda()
    extra = torch.randn(geom.L * N, cout, generator=g).bfloat16().cuda()
    res = extra if aux == "res" else None
    mask = extra if aux == "mask" else None
    y = ops.conv2d_levels(x, geom, N, w, cout, 3, 3, 1, None, shift, res, aux is None, None, mask)
    xs, es = _split(x, N, cin), _split(extra, N, cout)
    ys = _split(y, N, cout)
    for l in range(5):
        ref = ops.conv2d(xs[l].contiguous(), w, cout, 3, 3, 1, 1, None, shift, es[l] if aux == "res" else None, aux is None,
                         None, False, es[l] if aux == "mask" else None)
        assert torch.equal(ys[l], ref), l
    # and against the fp32 reference of the op
    ref0 = _ref_fwd(xs[0].cpu(), w.cpu(), 1, 1, None, shift.cpu(), es[0].cpu() if aux == "res" else None, aux is None)
    if aux == "mask":
        ref0 = ref0 * (es[0].cpu().float() > 0)
    assert torch.allclose(ys[0].float().cpu(), ref0, rtol=RTOL, atol=ATOL)


def test_wgrad_levels_equals_sum_of_levels():
    from ubteacher import ops
    geom, N = _level_geom(), 3
    g = torch.Generator().manual_seed(5)
    x = torch.randn(geom.L * N, 256, generator=g).bfloat16().cuda()
    dy = torch.randn(geom.L * N, 80, generator=g).bfloat16().cuda()
    dw = torch.zeros(80, 3, 3, 256, device="cuda")
    ops.conv2d_wgrad_levels(x, dy, geom, N, 80, 3, 3, 1, dw, None, 73)
    ref = torch.zeros_like(dw)
    for xl, dl in zip(_split(x, N, 256), _split(dy, N, 80)):
        ops.conv2d_wgrad(xl.contiguous(), dl.contiguous(), 80, 3, 3, 1, 1, ref, None, 73)
    torch.cuda.synchronize()
    assert (dw[73:] == 0).all()
    # same products, different fp32 atomic accumulation order
    assert torch.allclose(dw, ref, rtol=1e-3, atol=1e-3), (dw - ref).abs().max()
    # fp32 reference of the op on the largest level
    xl, dl = _split(x, N, 256)[0].float().cpu(), _split(dy, N, 80)[0].float().cpu()
    xt = xl.permute(0, 3, 1, 2).requires_grad_(False)
    wt = torch.zeros(80, 256, 3, 3, requires_grad=True)
    F.conv2d(xt, wt, padding=1).backward(dl.permute(0, 3, 1, 2))
    dw0 = torch.zeros(80, 3, 3, 256, device="cuda")
    ops.conv2d_wgrad(_split(x, N, 256)[0].contiguous(), _split(dy, N, 80)[0].contiguous(), 80, 3, 3, 1, 1, dw0, None, 80)
    assert torch.allclose(dw0.cpu(), wt.grad.permute(0, 2, 3, 1), rtol=2e-3, atol=2e-2)


def test_groupnorm_levels_equals_per_level():
    from ubteacher import ops
    geom, N = _level_geom(), 3
    g = torch.Generator().manual_seed(6)
    x = torch.randn(geom.L * N, 256, generator=g).bfloat16().cuda()
    dy = torch.randn(geom.L * N, 256, generator=g).bfloat16().cuda()
    gam = (torch.rand(256, generator=g) + 0.5).cuda()
    bet = torch.randn(256, generator=g).cuda()
    y, stats = ops.groupnorm_relu_levels_fwd(x, geom, N, gam, bet)
    dgam, dbet, dbias = [torch.zeros(256, device="cuda") for _ in range(3)]
    dx = ops.groupnorm_relu_levels_bwd(dy, x, geom, N, stats, gam, bet, dgam, dbet, dbias_prev=dbias)
    rg, rb, rbias = [torch.zeros(256, device="cuda") for _ in range(3)]
    for l, (xl, dl, yl, dxl) in enumerate(zip(_split(x, N, 256), _split(dy, N, 256), _split(y, N, 256), _split(dx, N, 256))):
        yr, st = ops.groupnorm_relu_fwd(xl.contiguous(), gam, bet)
        dr = ops.groupnorm_relu_bwd(dl.contiguous(), xl.contiguous(), st, gam, bet, rg, rb, dbias_prev=rbias)
        # same arithmetic; only the order of the fp64 / fp32 atomics differs -> at most one bf16 ulp
        assert torch.allclose(yl.float(), yr.float(), rtol=1 / 128, atol=1e-3), l
        assert torch.allclose(dxl.float(), dr.float(), rtol=1 / 64, atol=2e-3), l
    assert torch.allclose(dgam, rg, rtol=1e-3, atol=1e-2) and torch.allclose(dbet, rb, rtol=1e-3, atol=1e-2)
    assert torch.allclose(dbias, rbias, rtol=1e-3, atol=1e-2)
