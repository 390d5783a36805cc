"""Thin Python wrappers over the C-ABI kernels (include/ut2.h). Tensors are torch CUDA tensors used as
device-memory handles only; every arithmetic op below is one of our sm_100a kernels."""
import ctypes

import torch

from . import _C
from ._C import f32, f64, i64

BF16 = torch.bfloat16

# bench.py sets PROFILE = {"conv_fwd": [], "conv_wgrad": []} to bracket every tensor-core launch with CUDA events
# on the launching stream: entries are (start_event, end_event, algorithmic_flops).
PROFILE = None


def _timed(kind, flops, name, *args, nbytes=0.0):
    if PROFILE is None:
        tag = None
        if _C.EVENT_PROFILE is not None:   # per-shape breakdown: N,H,W,Cin -> Cout RxS /stride
            a = args
            tag = f"{kind} {a[1]}x{a[2]}x{a[3]}x{a[4]}->{a[6]} {a[7]}x{a[8]}/{a[9]}"
        _C.counted_call(name, *args, tag=tag)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _C.counted_call(name, *args)
    e1.record()
    PROFILE[kind].append((e0, e1, flops, nbytes))


def conv_out_hw(H, W, R, S, stride, pad):
    return (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1


def conv2d(x, w, cout, R, S, stride, pad, scale=None, shift=None, residual=None, relu=False, out=None,
           res_up2=False, relu_mask=None, alg_flops=None):
    """x: [N,H,W,Cin] bf16 NHWC; w: bf16 [>=cout, R, S, Cin]; returns [N,P,Q,cout] bf16."""
    N, H, W, Cin = x.shape
    P, Q = conv_out_hw(H, W, R, S, stride, pad)
    if out is None:
        out = torch.empty((N, P, Q, cout), dtype=BF16, device=x.device)
    flops = alg_flops if alg_flops is not None else 2.0 * N * P * Q * cout * Cin * R * S
    # algorithmic bytes (bf16): the input pixels the filter touches, the weights, the output, residual and mask tiles
    zero_frac = (alg_flops / (2.0 * N * P * Q * cout * Cin * R * S)) if alg_flops is not None else 1.0
    nbytes = 2.0 * ((N * P * Q if R * S == 1 else N * H * W * zero_frac) * Cin + cout * Cin * R * S +
                    N * P * Q * cout * (1 + (0.25 if res_up2 else 1) * (residual is not None) + (relu_mask is not None)))
    _timed("conv_fwd", flops, "ut2_conv2d_nhwc_bf16_fwd", x, N, H, W, Cin, w, cout, R, S, stride, pad, scale, shift,
           residual, int(res_up2), relu_mask, int(relu), out, nbytes=nbytes)
    return out


def conv2d_wgrad(x, dy, cout, R, S, stride, pad, dw, scale=None, cout_store=0):
    """Accumulates dW (fp32, [cout, R, S, Cin]) += dY^T * im2col(X)."""
    N, H, W, Cin = x.shape
    flops = 2.0 * dy.shape[0] * dy.shape[1] * dy.shape[2] * (cout_store or cout) * Cin * R * S
    px = dy.shape[0] * dy.shape[1] * dy.shape[2]
    nbytes = 2.0 * ((px if R * S == 1 else N * H * W) * Cin + px * cout) + 4.0 * cout * Cin * R * S
    _timed("conv_wgrad", flops, "ut2_conv2d_nhwc_bf16_wgrad", x, N, H, W, Cin, dy, cout, R, S, stride, pad, scale, dw,
           cout_store, nbytes=nbytes)


def conv2d_levels(x, geom, N, w, cout, R, S, pad, scale=None, shift=None, residual=None, relu=False, out=None, relu_mask=None):
    """x: level-major [N*L, Cin] bf16 (levels of `geom` back to back); returns level-major [N*L, cout] bf16."""
    Cin = x.shape[-1]
    rows = geom.L * N
    if out is None:
        out = torch.empty((rows, cout), dtype=BF16, device=x.device)
    flops = 2.0 * rows * cout * Cin * R * S
    args = ("ut2_conv2d_levels_bf16_fwd", x, geom.num, geom.c_hw, N, Cin, w, cout, R, S, pad, scale, shift, residual, relu_mask,
            int(relu), out)
    if PROFILE is None:
        tag = f"conv_fwd {N}xL{geom.L}x{Cin}->{cout} {R}x{S}/1" if _C.EVENT_PROFILE is not None else None
        _C.counted_call(*args, tag=tag)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _C.counted_call(*args)
        e1.record()
        PROFILE["conv_fwd"].append((e0, e1, flops, 2.0 * (rows * (Cin + cout * (1 + (residual is not None) + (relu_mask is not None))) + cout * Cin * R * S)))
    return out


def conv2d_wgrad_levels(x, dy, geom, N, cout, R, S, pad, dw, scale=None, cout_store=0):
    Cin = x.shape[-1]
    flops = 2.0 * geom.L * N * (cout_store or cout) * Cin * R * S
    args = ("ut2_conv2d_levels_bf16_wgrad", x, geom.num, geom.c_hw, N, Cin, dy, cout, R, S, pad, scale, dw, cout_store)
    if PROFILE is None:
        tag = f"conv_wgrad {N}xL{geom.L}x{Cin}->{cout} {R}x{S}/1" if _C.EVENT_PROFILE is not None else None
        _C.counted_call(*args, tag=tag)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _C.counted_call(*args)
        e1.record()
        PROFILE["conv_wgrad"].append((e0, e1, flops, 2.0 * geom.L * N * (Cin + cout) + 4.0 * cout * Cin * R * S))


def groupnorm_relu_levels_fwd(x, geom, N, gamma, beta, eps=1e-5, relu=True):
    y = torch.empty_like(x)
    stats = torch.empty((geom.num * N, 32, 2), dtype=torch.float64, device=x.device)
    _C.counted_call("ut2_groupnorm_relu_levels_fwd", x, gamma, beta, f32(eps), y, stats, geom.num, geom.c_hws, N, x.shape[-1], 32,
                    int(relu))
    _C.launch_count += 1  # two kernels
    return y, stats


def groupnorm_relu_levels_bwd(dy, x, geom, N, stats, gamma, beta, dgamma, dbeta, eps=1e-5, relu=True, dbias_prev=None):
    dx = torch.empty_like(x)
    ws = torch.empty((geom.num * N, 32, 2), dtype=torch.float64, device=x.device)
    _C.counted_call("ut2_groupnorm_relu_levels_bwd", dy, x, stats, gamma, beta, f32(eps), dx, dgamma, dbeta, dbias_prev, ws,
                    geom.num, geom.c_hws, N, x.shape[-1], 32, int(relu))
    _C.launch_count += 1
    return dx


def groupnorm_relu_fwd(x, gamma, beta, eps=1e-5, relu=True):
    N, H, W, C = x.shape
    y = torch.empty_like(x)
    stats = torch.empty((N, 32, 2), dtype=torch.float64, device=x.device)
    _C.counted_call("ut2_groupnorm_relu_fwd", x, gamma, beta, f32(eps), y, stats, N, H * W, C, 32, int(relu))
    _C.launch_count += 1  # two kernels
    return y, stats


def groupnorm_relu_bwd(dy, x, stats, gamma, beta, dgamma, dbeta, eps=1e-5, relu=True, dbias_prev=None):
    N, H, W, C = x.shape
    dx = torch.empty_like(x)
    ws = torch.empty((N, 32, 2), dtype=torch.float64, device=x.device)
    _C.counted_call("ut2_groupnorm_relu_bwd", dy, x, stats, gamma, beta, f32(eps), dx, dgamma, dbeta, dbias_prev, ws, N,
                    H * W, C, 32, int(relu))
    _C.launch_count += 1
    return dx


def stem_conv(img_u8_chw, w_rsck, scale, shift, mean, std, out_slice, P, Q, tensor_core=True):
    """tensor_core=True: tcgen05 implicit GEMM (bf16 inputs); False: the fp32 CUDA-core kernel."""
    _, h, w = img_u8_chw.shape
    _C.counted_call("ut2_stem_conv_u8_tc" if tensor_core else "ut2_stem_conv_u8", img_u8_chw, h, w, w_rsck, scale, shift, f32(mean[0]), f32(mean[1]),
                    f32(mean[2]), f32(std[0]), f32(std[1]), f32(std[2]), out_slice, P, Q)


def stem_conv_batched(images, w_rsck, scale, shift, mean, std, out, P, Q):
    """images: list of uint8 CHW device tensors (sizes may differ); out [N, P, Q, 64] bf16. One launch per 32 images."""
    n = len(images)
    ptrs = (ctypes.c_void_p * n)(*[im.data_ptr() for im in images])
    hs = (ctypes.c_int * n)(*[int(im.shape[1]) for im in images])
    ws = (ctypes.c_int * n)(*[int(im.shape[2]) for im in images])
    _C.counted_call("ut2_stem_conv_u8_tc_batched", ptrs, hs, ws, n, w_rsck, scale, shift, f32(mean[0]), f32(mean[1]), f32(mean[2]),
                    f32(std[0]), f32(std[1]), f32(std[2]), out, P, Q)


# data_ptr of the first image of a static input group -> device int32 [n, 2] with the (h, w) of the group's images. Filled by the
# trainer's CUDA-graph staging for batches of mixed image sizes: the stem then reads the sizes from device memory, and one
# captured graph serves every batch with the same padded size (engine/trainer.py: _stage_inputs).
# Entries are (weak reference to the static buffer, tensor): once the trainer that owns the buffer is gone its address may be
# handed to an unrelated image, so a dead reference invalidates the entry.
STATIC_SIZES = {}
STATIC_HW = {}       # the same sizes as float32 [n, 2] (the R-CNN proposal / detection clipping kernels read image_hw)


def register_static(table, buf, value):
    import weakref
    table[buf.data_ptr()] = (weakref.ref(buf), value)


def lookup_static(table, ptr):
    hit = table.get(ptr)
    if hit is None:
        return None
    if hit[0]() is None:
        del table[ptr]
        return None
    return hit[1]


def stem_pool_batched(images, w_rsck, scale, shift, mean, std, Hp, Wp):
    """Fused stem + max-pool: list of uint8 CHW device images -> [N, Hp/4, Wp/4, 64] bf16 (csrc/stem_pool.cu)."""
    n = len(images)
    dev = images[0].device
    lib = _C.lib()
    lib.ut2_stem_pool_workspace_bytes.restype = ctypes.c_longlong
    wsb = lib.ut2_stem_pool_workspace_bytes(n, Hp, Wp)
    if wsb <= 0:
        raise RuntimeError("stem_pool: the padded size must be a multiple of 4")
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = torch.empty((n, Hp // 4, Wp // 4, 64), dtype=BF16, device=dev)
    ptrs = (ctypes.c_void_p * n)(*[im.data_ptr() for im in images])
    hs = (ctypes.c_int * n)(*[int(im.shape[1]) for im in images])
    ws_ = (ctypes.c_int * n)(*[int(im.shape[2]) for im in images])
    _C.counted_call("ut2_stem_pool_u8_batched", ptrs, hs, ws_, n, w_rsck, scale, shift, f32(mean[0]), f32(mean[1]), f32(mean[2]),
                    f32(std[0]), f32(std[1]), f32(std[2]), ws, i64(wsb), out, Hp, Wp, lookup_static(STATIC_SIZES, images[0].data_ptr()))
    _C.launch_count += 1  # two kernels
    return out


def maxpool3x3s2(x):
    N, H, W, C = x.shape
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((N, P, Q, C), dtype=BF16, device=x.device)
    _C.counted_call("ut2_maxpool3x3s2_nhwc", x, y, N, H, W, C)
    return y


def upsample2x_add(lat, top):
    N, H, W, C = lat.shape
    out = torch.empty_like(lat)
    _C.counted_call("ut2_upsample2x_add_nhwc", lat, top, out, N, H, W, C)
    return out


def downsample2x_sum(g, addend=None):
    N, H, W, C = g.shape
    out = torch.empty((N, H // 2, W // 2, C), dtype=BF16, device=g.device)
    _C.counted_call("ut2_downsample2x_sum_nhwc", g, addend, out, N, H // 2, W // 2, C)
    return out


def relu_bwd(dy, y, dy2=None):
    g = torch.empty_like(y)
    _C.counted_call("ut2_relu_bwd_bf16", dy, dy2, y, g, i64(y.numel()))
    return g


def add_bf16(a, b):
    out = torch.empty_like(a)
    _C.counted_call("ut2_add_bf16", a, b, out, i64(a.numel()))
    return out


def zero_stuff_s2(g, H, W, oh=0, ow=0):
    N, P, Q, C = g.shape
    out = torch.empty((N, H, W, C), dtype=BF16, device=g.device)
    _C.counted_call("ut2_zero_stuff_s2_nhwc", g, out, N, P, Q, H, W, C, oh, ow)
    return out


# set to a list by an engine's backward: conv bias gradients are then queued and flushed as ONE batched launch
COLSUM_QUEUE = None


def colsum(g2d, db):
    M, C = g2d.shape
    if COLSUM_QUEUE is not None and C <= 256:
        COLSUM_QUEUE.append((g2d, db))        # keeps the gradient tensor alive until the flush
        return
    _C.counted_call("ut2_colsum_bf16", g2d, db, M, C)


def colsum_flush():
    """Launch the queued column sums (16 per launch) and leave queueing mode."""
    global COLSUM_QUEUE
    q, COLSUM_QUEUE = COLSUM_QUEUE, None
    for i in range(0, len(q or []), 16):
        part = q[i:i + 16]
        n = len(part)
        gs = (ctypes.c_void_p * n)(*[g.data_ptr() for g, _ in part])
        dbs = (ctypes.c_void_p * n)(*[d.data_ptr() for _, d in part])
        Ms = (ctypes.c_int * n)(*[int(g.shape[0]) for g, _ in part])
        Cs = (ctypes.c_int * n)(*[int(g.shape[1]) for g, _ in part])
        _C.counted_call("ut2_colsum_bf16_batched", gs, dbs, Ms, Cs, n)


def pack_conv_weight(w_master, wf, wt, cout, cin, R, S, coutT):
    _C.counted_call("ut2_pack_conv_weight", w_master, wf, wt, cout, cin, R, S, coutT)


def ema_update(student_flat, teacher_flat, keep_rate):
    _C.counted_call("ut2_ema_update", student_flat, teacher_flat, i64(student_flat.numel()), f64(keep_rate))


def sgd_step(p, g, buf, lr, momentum, wd, first_step, zero_grad=True, grad_scale=1.0, lr_dev=None):
    _C.counted_call("ut2_sgd_step", p, g, buf, i64(p.numel()), f32(lr), lr_dev, f32(momentum), f32(wd), int(first_step),
                    int(zero_grad), f32(grad_scale))


class LevelGeom:
    """Host-side description of the FPN levels a batch was run on (launch parameters for the loss,
    target-assignment and proposal kernels)."""

    def __init__(self, hw, strides, sizes_of_interest=None):
        self.hw = [tuple(x) for x in hw]
        self.strides = list(strides)
        self.num = len(self.hw)
        self.L = sum(h * w for h, w in self.hw)
        self.c_hw = (ctypes.c_int * (2 * self.num))(*[v for x in self.hw for v in x])
        self.c_strides = (ctypes.c_int * self.num)(*self.strides)
        self.c_hws = (ctypes.c_int * self.num)(*[h * w for h, w in self.hw])
        if sizes_of_interest is not None:
            INF = 100000000.0
            rng, prev = [], -1.0
            for s in sizes_of_interest:
                rng += [prev, float(s)]
                prev = float(s)
            rng += [prev, INF]
            self.c_ranges = (ctypes.c_float * (2 * self.num))(*rng[: 2 * self.num])
        else:
            self.c_ranges = None
        off = [0]
        for h, w in self.hw:
            off.append(off[-1] + h * w)
        self.off = off


def fcos_assign_targets(geom, N, boxes, classes, counts, bvar, num_classes=80, center_radius=0.0, ignore_near=False):
    """boxes [N,G,4] f32, classes [N,G] i64, counts [N] i32, bvar [N,G,4] f32 or None. center_radius > 0: CENTER_SAMPLE
    with that POS_RADIUS; ignore_near: keep_locations per fcos_outputs.py:841-848."""
    dev = boxes.device
    P = geom.L * N
    G = boxes.shape[1]
    out = {
        "labels": torch.empty(P, dtype=torch.int64, device=dev),
        "target_inds": torch.empty(P, dtype=torch.int64, device=dev),
        "reg_targets": torch.empty((P, 4), dtype=torch.float32, device=dev),
        "boundary_vars": torch.empty((P, 4), dtype=torch.float32, device=dev),
        "keep_locations": torch.empty(P, dtype=torch.uint8, device=dev),
        "norm": torch.empty(2, dtype=torch.float32, device=dev),
    }
    _C.counted_call("ut2_fcos_assign_targets", geom.num, geom.c_hw, geom.c_strides, geom.c_ranges, N, G, boxes,
                    classes, counts, bvar, num_classes, f32(center_radius), int(bool(ignore_near)), out["labels"], out["target_inds"], out["reg_targets"],
                    out["boundary_vars"], out["keep_locations"], out["norm"])
    return out


KL_MODES = {"nlloss": 0, "mean": 1, "sum": 2, "weight_ctr_sum": 3, "weight_ctr_mean": 4}     # kl_mode of ut2_fcos_loss_*


def fcos_loss_fwd(geom, N, cls_out, box_out, scales, tg, mode, alpha, gamma, kl_w, ts_better, ts_cert, world,
                  num_classes=80, kl_mode=0):
    dev = box_out.device
    acc = torch.empty(8, dtype=torch.float64, device=dev)
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    _C.counted_call("ut2_fcos_loss_fwd", geom.num, geom.c_hw, geom.c_strides, N, cls_out, box_out, box_out.shape[-1],
                    scales, tg["labels"], tg["keep_locations"], tg["reg_targets"], tg["boundary_vars"], num_classes, mode, f32(alpha),
                    f32(gamma), f32(kl_w), f32(ts_better), f32(ts_cert), tg["norm"], f32(world), int(kl_mode), acc, losses)
    _C.launch_count += 2 if mode != 2 else 1
    return losses, acc


def fcos_loss_bwd(geom, N, cls_out, box_out, scales, tg, mode, alpha, gamma, kl_w, ts_better, ts_cert, world, acc,
                  gout, dcls, dbox, dscales, num_classes=80, accumulate=False, kl_mode=0):
    _C.counted_call("ut2_fcos_loss_bwd", geom.num, geom.c_hw, geom.c_strides, N, cls_out, box_out, box_out.shape[-1],
                    scales, tg["labels"], tg["keep_locations"], tg["reg_targets"], tg["boundary_vars"], num_classes, mode, f32(alpha),
                    f32(gamma), f32(kl_w), f32(ts_better), f32(ts_cert), tg["norm"], f32(world), int(kl_mode), acc, gout, dcls, dbox,
                    dscales, int(accumulate))
    if mode != 2:
        _C.launch_count += 1


_NMS_METHOD = {"cls": 0, "cls_n_ctr": 1, "cls_n_loc": 2}


def fcos_predict_proposals(geom, N, cls_out, box_out, scales, method, pre_thr, pre_topk, nms_thr, post_topk,
                           out_cap=128, num_classes=80):
    if method not in _NMS_METHOD:
        raise ValueError("Undefined nms criteria")
    dev = box_out.device
    lib = _C.lib()
    lib.ut2_fcos_predict_workspace_bytes.restype = ctypes.c_longlong
    wsb = lib.ut2_fcos_predict_workspace_bytes(geom.num, N, ctypes.c_longlong(geom.L), num_classes, pre_topk)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = {
        "pred_boxes": torch.zeros((N, out_cap, 4), dtype=torch.float32, device=dev),
        "scores": torch.zeros((N, out_cap), dtype=torch.float32, device=dev),
        "pred_classes": torch.zeros((N, out_cap), dtype=torch.int64, device=dev),
        "centerness": torch.zeros((N, out_cap), dtype=torch.float32, device=dev),
        "cls_confid": torch.zeros((N, out_cap), dtype=torch.float32, device=dev),
        "reg_pred_std": torch.zeros((N, out_cap, 4), dtype=torch.float32, device=dev),
        "locations": torch.zeros((N, out_cap, 2), dtype=torch.float32, device=dev),
        "fpn_levels": torch.zeros((N, out_cap), dtype=torch.int64, device=dev),
        "count": torch.zeros(N, dtype=torch.int32, device=dev),
    }
    _C.counted_call("ut2_fcos_predict_proposals", geom.num, geom.c_hw, geom.c_strides, N, num_classes, cls_out, box_out,
                    box_out.shape[-1], scales, _NMS_METHOD[method], f32(pre_thr), pre_topk, f32(nms_thr), post_topk,
                    out_cap, ws, i64(wsb), out["pred_boxes"], out["scores"], out["pred_classes"], out["centerness"],
                    out["cls_confid"], out["reg_pred_std"], out["locations"], out["fpn_levels"], out["count"])
    _C.launch_count += 5
    return out


def threshold_scatter(dets, mode, thr0, thr1=0.0):
    """dets: dict from fcos_predict_proposals. Returns the compacted pseudo-label set (same capacity)."""
    N, cap = dets["scores"].shape
    out = {k: torch.zeros_like(dets[k]) for k in ["pred_boxes", "scores", "pred_classes", "centerness", "cls_confid",
                                                  "reg_pred_std"]}
    out["count"] = torch.zeros_like(dets["count"])
    _C.counted_call("ut2_threshold_scatter", N, cap, mode, f32(thr0), f32(thr1), dets["count"], dets["pred_boxes"],
                    dets["scores"], dets["pred_classes"], dets["centerness"], dets["cls_confid"], dets["reg_pred_std"],
                    out["count"], out["pred_boxes"], out["scores"], out["pred_classes"], out["centerness"],
                    out["cls_confid"], out["reg_pred_std"])
    return out
