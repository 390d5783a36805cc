"""GPU parity of the tcgen05 implicit-GEMM convolution (fwd / dgrad-as-fwd / wgrad) against a plain
PyTorch fp32 reference of the same op (floating-point kernel: tolerance stated below)."""
import os

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


# --------------------------------------------------------------------------------------- benchmark-scale launches
# The kernels are persistent (grid <= 148 CTAs): the cases above never give a CTA a second tile. These launches have 4-7x
# as many tiles as SMs, i.e. every CTA walks the operand ring across tile boundaries, alternates the two TMEM accumulator
# buffers, flips the aux-tile phases ((it >> 1) & 1 with aux_dbl) and restages scale / shift between n-tiles — the regime
# bench.py runs (16 x 100 x 168 -> 2100 tiles). Reference: fp32 F.conv2d on the CPU, same tolerances as above.
BIG_FWD = [
    # id, (N, H, W, Cin, Cout, R, stride, pad), aux, relu
    ("head3x3_256", (5, 100, 168, 256, 256, 3, 1, 1), None, True),                 # 657 tiles, tensor-bound head / FPN output shape
    ("res2_conv3_aux_dbl", (2, 200, 336, 64, 256, 1, 1, 0), "res", True),          # 1050 tiles, K = 64: double-buffered aux tiles
    ("res4_conv3_res", (5, 50, 84, 256, 1024, 1, 1, 0), "res", True),              # 165 x 4 = 660 tiles, single aux tile, 4 n-tiles
    ("res5_conv3_res_8ntiles", (8, 25, 42, 512, 2048, 1, 1, 0), "res", True),      # 66 x 8 = 528 tiles, scale / shift restaging
    ("dgrad_mask_128", (4, 100, 168, 128, 128, 3, 1, 1), "mask", False),           # 525 tiles, ReLU-mask aux tile (dgrad epilogue)
    ("dgrad_mask_64_aux_dbl", (2, 200, 336, 64, 64, 1, 1, 0), "mask", False),      # 1050 tiles, mask + aux_dbl, N = 64 tile
    ("fpn_lateral_res_up2", (4, 100, 168, 512, 256, 1, 1, 0), "res_up2", False),   # 525 tiles, manual epilogue (nearest-2x residual)
    ("res_and_mask_128", (4, 100, 168, 64, 128, 1, 1, 0), "res+mask", False),      # 525 tiles, residual AND mask tiles (aux_kind 3)
    ("res_and_mask_1024", (5, 50, 84, 256, 1024, 1, 1, 0), "res+mask", False),     # 660 tiles: block-output ReLU backward fused in conv1.dgrad
    ("res_and_mask_2048", (8, 25, 42, 512, 2048, 1, 1, 0), "res+mask", False),     # 528 tiles, K = 512, two operand stages
    ("cls_logits_80", (4, 100, 168, 256, 80, 3, 1, 1), None, False),               # 525 tiles, Cout = 80
    ("res3_conv1_s2", (4, 200, 336, 256, 128, 1, 2, 0), None, True),               # 525 tiles, strided 1x1 (STRIDE_IN_1X1)
    ("p6_3x3_s2", (8, 100, 168, 256, 256, 3, 2, 1), None, False),                  # 263 tiles, strided 3x3
    ("dgrad_ragged_cin", (4, 100, 168, 80, 256, 3, 1, 1), None, False),            # 525 tiles, Cin = 80 (zero-filled ragged chunk)
    # narrow 3x3 / stride 1 convolutions: 2-D patch kernel (conv3x3_halo.cu), like "dgrad_mask_128" above
    ("halo_res2_conv2", (2, 200, 336, 64, 64, 3, 1, 1), None, True),               # 2 x 25 x 21 = 1050 full patches, N = 64
    ("halo_res3_conv2", (3, 100, 168, 128, 128, 3, 1, 1), None, True),             # 13 x 11 patches per image, ragged in both axes
    ("halo_ragged_mask", (3, 77, 101, 64, 128, 3, 1, 1), "mask", False),           # 10 x 7 patches per image, odd sizes, mask loads
    ("halo_cin256", (2, 90, 120, 256, 128, 3, 1, 1), None, True),                  # four 64-channel slices per patch
]
HALO_CASES = {"dgrad_mask_128", "halo_res2_conv2", "halo_res3_conv2", "halo_ragged_mask", "halo_cin256", "cls_logits_80"}


def _halo_launches():
    import ctypes
    from ubteacher import _C
    fn = _C.lib().ut2_conv3x3_halo_launches
    fn.restype = ctypes.c_longlong
    return int(fn())


@pytest.mark.parametrize("name,case,aux,relu", BIG_FWD, ids=[c[0] for c in BIG_FWD])
def test_conv_fwd_many_tiles_per_cta(name, case, aux, relu):
    from ubteacher import _C

    N, H, W, Cin, Cout, R, stride, pad = case
    x, w = _mk(N, H, W, Cin, Cout, R, seed=len(name))
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    tiles = ((N * P * Q + 127) // 128) * ((Cout + 255) // 256 if Cout > 256 else 1)
    assert tiles > 148, "the point of this test is more tiles than SMs"
    g = torch.Generator().manual_seed(13)
    shift = torch.randn(Cout, generator=g)
    res = mask = None
    if aux in ("res", "res+mask"):
        res = torch.randn(N, P, Q, Cout, generator=g).bfloat16()
    if aux == "res_up2":
        res = torch.randn(N, P // 2, Q // 2, Cout, generator=g).bfloat16()
    if aux in ("mask", "res+mask"):
        mask = torch.randn(N, P, Q, Cout, generator=g).relu().bfloat16()      # half of it exactly zero, like a ReLU output
    y = torch.full((N, P, Q, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    cu = lambda t: t.cuda() if t is not None else None
    halo0 = _halo_launches()
    _C.call("ut2_conv2d_nhwc_bf16_fwd", x.cuda(), N, H, W, Cin, w.cuda(), Cout, R, R, stride, pad,
            None, shift.cuda(), cu(res), int(aux == "res_up2"), cu(mask), int(relu), y)
    torch.cuda.synchronize()
    if os.environ.get("UT2_HALO3", "1") != "0":
        assert _halo_launches() - halo0 == (1 if name in HALO_CASES else 0), "2-D patch kernel taken exactly where expected"
    r = res
    if aux == "res_up2":
        r = res.float().repeat_interleave(2, 1).repeat_interleave(2, 2)
    ref = _ref_fwd(x, w, stride, pad, None, shift, r, relu)
    if mask is not None:
        ref = ref * (mask.float() > 0)
    got = y.float().cpu()
    assert torch.isfinite(got).all(), "unwritten output rows"
    torch.testing.assert_close(got, ref, rtol=RTOL, atol=ATOL)


BIG_WGRAD = [
    # id, (N, H, W, Cin, Cout, R, stride, pad)          stage width chosen by the launcher
    ("head3x3_256_pix96", (4, 100, 168, 256, 256, 3, 1, 1)),          # block_n 256 -> 96-pixel stages, split-K 8 x 88 blocks
    ("res2_1x1_64_pix128", (2, 200, 336, 64, 256, 1, 1, 0)),          # block_n 64 -> 128-pixel stages, 65-way split-K
    ("res3_3x3_128_pix128", (4, 100, 168, 128, 128, 3, 1, 1)),        # block_n 128
    ("res4_1x1_1024_256_pix96", (5, 50, 84, 1024, 256, 1, 1, 0)),     # 4 input-channel tiles x 2 output tiles
    ("cls_logits_80", (4, 100, 168, 256, 80, 3, 1, 1)),               # ragged Cout
    ("res3_conv1_s2", (4, 200, 336, 256, 128, 1, 2, 0)),              # strided
]


@pytest.mark.parametrize("name,case", BIG_WGRAD, ids=[c[0] for c in BIG_WGRAD])
def test_conv_wgrad_long_reductions(name, case):
    """Weight gradients at reduction lengths (tens of thousands of pixels per output tile, split-K over the whole grid)
    where each CTA wraps its operand ring many times."""
    from ubteacher import _C

    N, H, W, Cin, Cout, R, stride, pad = case
    x, w = _mk(N, H, W, Cin, Cout, R, seed=len(name))
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    g = torch.Generator().manual_seed(5)
    dy = torch.randn(N, P, Q, Cout, generator=g).bfloat16()
    dw = torch.zeros(Cout, R, R, Cin, dtype=torch.float32, device="cuda")
    _C.call("ut2_conv2d_nhwc_bf16_wgrad", x.cuda(), N, H, W, Cin, dy.cuda(), Cout, R, R, stride, pad, None, dw, 0)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, R, R), dy.float().permute(0, 3, 1, 2),
                                      stride=stride, padding=pad).permute(0, 2, 3, 1)
    scale = ref.abs().max().item()
    torch.testing.assert_close(dw.cpu(), ref, rtol=2e-3, atol=2e-3 * scale)


@pytest.mark.parametrize("pix", [64, 96, 128])
def test_conv_wgrad_every_stage_width(pix, monkeypatch):
    """All three instantiations of conv_wgrad_kernel<PIX> on one long reduction (UT2_WG_PIX is read once per process, so the
    override runs in a child interpreter)."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path[:0] = [%r, %r]
from ubteacher import _C
N, H, W, Cin, Cout, R = 3, 100, 168, 256, 256, 3
g = torch.Generator().manual_seed(1)
x = torch.randn(N, H, W, Cin, generator=g).bfloat16()
dy = torch.randn(N, H, W, Cout, generator=g).bfloat16()
dw = torch.zeros(Cout, R, R, Cin, device="cuda")
_C.call("ut2_conv2d_nhwc_bf16_wgrad", x.cuda(), N, H, W, Cin, dy.cuda(), Cout, R, R, 1, 1, None, dw, 0)
torch.cuda.synchronize()
ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, R, R), dy.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
torch.testing.assert_close(dw.cpu(), ref, rtol=2e-3, atol=2e-3 * ref.abs().max().item())
print("OK")
''' % (os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "unbiased-teacher-v2_b200"),
       os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, UT2_WG_PIX=str(pix))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]


FULL_HW = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]       # the FCOS pyramid of an 800 x 1344 batch


def _split_hw(t, N, C, hws):
    out, off = [], 0
    for h, w in hws:
        out.append(t[off:off + N * h * w].view(N, h, w, C))
        off += N * h * w
    return out


@pytest.mark.parametrize("cout,relu", [(256, False), (80, False)])
def test_conv_levels_full_size_pyramid(cout, relu):
    """ut2_conv2d_levels_bf16_fwd / _wgrad over the level-major pyramid of FOUR full-size images (89 600 rows, 702 tiles: tiles
    of all five levels interleave inside every CTA) against fp32 F.conv2d per level."""
    from ubteacher import ops
    geom, N = ops.LevelGeom(FULL_HW, [8, 16, 32, 64, 128]), 4
    g = torch.Generator().manual_seed(cout)
    x = torch.randn(geom.L * N, 256, generator=g).bfloat16()
    w = (torch.randn(cout, 3, 3, 256, generator=g) / (9 * 256) ** 0.5).bfloat16()
    shift = torch.randn(cout, generator=g)
    dy = torch.randn(geom.L * N, cout, generator=g).bfloat16()
    halo0 = _halo_launches()
    y = ops.conv2d_levels(x.cuda(), geom, N, w.cuda(), cout, 3, 3, 1, None, shift.cuda(), None, relu)
    if os.environ.get("UT2_HALO3", "1") != "0":     # the 80-channel predictors run on 2-D patches, level-major (conv3x3_halo.cu)
        assert _halo_launches() - halo0 == (1 if cout == 80 else 0)
    dw = torch.zeros(cout, 3, 3, 256, device="cuda")
    ops.conv2d_wgrad_levels(x.cuda(), dy.cuda(), geom, N, cout, 3, 3, 1, dw)
    torch.cuda.synchronize()
    ref_dw = torch.zeros(cout, 3, 3, 256)
    for xl, yl, dl in zip(_split_hw(x, N, 256, FULL_HW), _split_hw(y.cpu(), N, cout, FULL_HW), _split_hw(dy, N, cout, FULL_HW)):
        ref = _ref_fwd(xl, w, 1, 1, None, shift, None, relu)
        torch.testing.assert_close(yl.float(), ref, rtol=RTOL, atol=ATOL)
        ref_dw += torch.nn.grad.conv2d_weight(xl.float().permute(0, 3, 1, 2), (cout, 256, 3, 3), dl.float().permute(0, 3, 1, 2),
                                              padding=1).permute(0, 2, 3, 1)
    torch.testing.assert_close(dw.cpu(), ref_dw, rtol=2e-3, atol=2e-3 * ref_dw.abs().max().item())
