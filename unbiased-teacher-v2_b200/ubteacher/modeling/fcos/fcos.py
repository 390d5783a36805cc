"""``FCOS`` proposal generator, ``FCOSHead`` and ``Scale`` with the reference's names and call contracts
(ubteacher/modeling/fcos/fcos.py:22-29, :44-217, :220-376), executed by the B200 engine's head kernels.

    FCOSHead(cfg, input_shape)(x, top_module=None, yield_bbox_towers=False)
        x: list of per-level [N, 256, H_l, W_l] feature maps (any float dtype / memory format)
        -> (logits, bbox_reg, bbox_reg_std, ctrness, top_feats, bbox_towers): per-level lists of [N, {80, 68, 4, 1}, H_l, W_l]
    FCOS(cfg, input_shape)(images, features, gt_instances=None, top_module=None, output_raw=False,
                           nms_method="cls_n_ctr", ignore_near=False, branch="labeled")
        -> (results, losses) | (results, losses, raw_output)

Both are differentiable: the outputs hang on ONE autograd node whose backward runs the engine's explicit head backward
(GroupNorm / dgrad / wgrad kernels), accumulates the parameter gradients in the arena (``p.grad`` of ``parameters()``)
and returns the gradients of the input feature maps.
"""
import torch
from torch import nn

from ... import ops
from ...d2compat.registry import PROPOSAL_GENERATOR_REGISTRY
from ...utils.comm import compute_locations
from ..views import ArenaView, nchw
from .fcos_outputs import BoxSet, FCOSOutputs, as_boxset, dets_to_instances

__all__ = ["FCOS", "FCOSHead", "Scale"]
HEAD_PREFIX = "proposal_generator.fcos_head."
BF16 = torch.bfloat16


def _own_engine(cfg):
    from ..fcos_engine import FcosEngine
    dev = cfg.MODEL.DEVICE
    if not str(dev).startswith("cuda"):
        raise RuntimeError("the UT2 B200 path has no CPU fallback: MODEL.DEVICE must be a CUDA device")
    return FcosEngine(cfg, device=dev, seed=max(cfg.SEED, 0))


class Scale(nn.Module):
    """fcos.py:22-29: ``input * scale`` with a learnable scalar; here a view of one arena entry."""

    def __init__(self, init_value=1.0, engine=None, index=None):
        super().__init__()
        if engine is None:
            self.scale = nn.Parameter(torch.tensor([float(init_value)]))
        else:
            object.__setattr__(self, "engine", engine)
            n = f"{HEAD_PREFIX}scales.{index}.scale"
            self.scale = nn.Parameter(engine.arena.views[n])
            self.scale.grad = engine.arena.gviews[n]

    def forward(self, input):
        return input * self.scale.to(input.dtype)


def pack_levels(engine, xs, strides, soi):
    """list of [N, C, H_l, W_l] -> (level-major [N * L, C] bf16, geometry). A no-op when the maps already are the NCHW
    views of one level-major buffer (the backbone view's output)."""
    N, C = xs[0].shape[0], xs[0].shape[1]
    geom = ops.LevelGeom([(int(x.shape[2]), int(x.shape[3])) for x in xs], strides, soi)
    base = getattr(xs[0], "_ut2_level_major", None)
    if base is not None and all(getattr(x, "_ut2_level_major", None) is base for x in xs) and base.shape == (geom.L * N, C):
        return base, geom
    feat = torch.empty((geom.L * N, C), dtype=BF16, device=engine.device)
    for x, lv in zip(xs, engine.level_views(feat, geom, N, C)):
        lv.copy_(x.detach().permute(0, 2, 3, 1))
    return feat, geom


def split_head_outputs(engine, cls_out, box_out, geom, N, scales):
    """level-major (cls_out, box_out) -> per-level NCHW fp32 lists (logits, bbox_reg [Scale_l applied], std, ctrness)."""
    logits, reg, std, ctr = [], [], [], []
    for l, (c, b) in enumerate(zip(engine.level_views(cls_out, geom, N, 80), engine.level_views(box_out, geom, N, 80))):
        bf = nchw(b).float()
        logits.append(nchw(c).float())
        reg.append(bf[:, :68] * scales[l])
        std.append(bf[:, 68:72])
        ctr.append(bf[:, 72:73])
    return logits, reg, std, ctr


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, head, geom, N, feat, want_towers, trigger, *xs):
        eng = head.engine
        tape = {}
        cls_out, box_out = eng.head_forward(feat, geom, N, tape)
        logits, reg, std, ctr = split_head_outputs(eng, cls_out, box_out, geom, N, [eng.scales[l] for l in range(geom.num)])
        towers = []
        if want_towers:
            towers = [nchw(t).float() for t in eng.level_views(tape["head"]["bbox_tower"][1], geom, N, 256)]
            ctx.mark_non_differentiable(*towers)
        ctx.head, ctx.geom, ctx.N, ctx.tape, ctx.box_out = head, geom, N, tape, box_out
        ctx.in_dtypes = [x.dtype for x in xs]
        return tuple(logits + reg + std + ctr + towers)

    @staticmethod
    def backward(ctx, *grads):
        eng, geom, N = ctx.head.engine, ctx.geom, ctx.N
        n = geom.num
        P = geom.L * N
        dcls = torch.zeros((P, 80), dtype=BF16, device=eng.device)
        dbox = torch.zeros((P, 80), dtype=BF16, device=eng.device)
        dc, db = eng.level_views(dcls, geom, N, 80), eng.level_views(dbox, geom, N, 80)
        raw = eng.level_views(ctx.box_out, geom, N, 80)
        for l in range(n):
            g_log, g_reg, g_std, g_ctr = grads[l], grads[n + l], grads[2 * n + l], grads[3 * n + l]
            if g_log is not None:
                dc[l].copy_(g_log.permute(0, 2, 3, 1))
            if g_reg is not None:          # bbox_reg = Scale_l(bbox_pred): d bbox_pred = g * s_l, d s_l = sum(g * bbox_pred)
                g = g_reg.permute(0, 2, 3, 1).float()
                db[l][..., :68].copy_(g * eng.scales[l])
                eng.dscales[l] += (g * raw[l][..., :68].float()).sum()
            if g_std is not None:
                db[l][..., 68:72].copy_(g_std.permute(0, 2, 3, 1))
            if g_ctr is not None:
                db[l][..., 72:73].copy_(g_ctr.permute(0, 2, 3, 1))
        dfeat = eng.head_backward(ctx.tape, geom, N, dcls, dbox)
        ctx.tape = None
        dx = [nchw(d).to(dt) for d, dt in zip(eng.level_views(dfeat, geom, N, 256), ctx.in_dtypes)]
        return (None, None, None, None, None, None) + tuple(dx)


class FCOSHead(ArenaView):
    def __init__(self, cfg, input_shape=None, engine=None):
        super().__init__(engine if engine is not None else _own_engine(cfg), HEAD_PREFIX)
        f = cfg.MODEL.FCOS
        self.num_classes, self.fpn_strides = f.NUM_CLASSES, list(f.FPN_STRIDES)
        self.reg_max, self.reg_discrete, self.kl_loss = f.REG_MAX, f.REG_DISCRETE, f.KL_LOSS
        self.num_levels = len(self.fpn_strides)
        self.in_channels_to_top_module = 256
        self._soi = list(f.SIZES_OF_INTEREST)
        self.scales = nn.ModuleList([Scale(engine=self.engine, index=i) for i in range(self.num_levels)])
        # a leaf that requires grad: the autograd node must fire (parameter gradients!) even when the inputs do not
        self._trigger = torch.zeros(1, device=self.engine.device, requires_grad=True)

    def forward(self, x, top_module=None, yield_bbox_towers=False):
        assert len(x) == self.num_levels and all(t.shape[1] == 256 for t in x), "FCOSHead: five 256-channel levels"
        eng = self.engine
        N = int(x[0].shape[0])
        feat, geom = pack_levels(eng, x, self.fpn_strides, self._soi)
        want_towers = bool(yield_bbox_towers or top_module is not None)
        n = self.num_levels
        if torch.is_grad_enabled():       # one autograd node: backward = the engine's explicit head backward
            out = _HeadFn.apply(self, geom, N, feat, want_towers, self._trigger, *x)
        else:
            tape = {} if want_towers else None
            cls_out, box_out = eng.head_forward(feat, geom, N, tape)
            lists = split_head_outputs(eng, cls_out, box_out, geom, N, [eng.scales[l] for l in range(n)])
            towers = [nchw(t).float() for t in eng.level_views(tape["head"]["bbox_tower"][1], geom, N, 256)] if want_towers else []
            out = tuple(lists[0] + lists[1] + lists[2] + lists[3] + towers)
        logits, reg, std, ctr = list(out[:n]), list(out[n:2 * n]), list(out[2 * n:3 * n]), list(out[3 * n:4 * n])
        towers = list(out[4 * n:5 * n]) if want_towers else []
        top_feats = [top_module(t) for t in towers] if top_module is not None else []
        bbox_towers = towers if yield_bbox_towers else []
        if self.kl_loss:
            return logits, reg, std, ctr, top_feats, bbox_towers
        return logits, reg, ctr, top_feats, bbox_towers


class RawOutput(dict):
    """The dense teacher output handed to ``PseudoGenerator.nms_from_dense`` (fcos.py:131-138). Holds the engine's
    level-major tensors (what the kernels consume) and materialises the reference's per-level NCHW lists on first use."""
    LAZY = ("logits_pred", "reg_pred", "reg_pred_std", "ctrness_pred", "locations")

    def __missing__(self, key):
        if key not in self.LAZY:
            raise KeyError(key)
        eng, geom, N = self["_engine"], self["geom"], self["N"]
        if key == "locations":
            val = [compute_locations(h, w, s, eng.device) for (h, w), s in zip(geom.hw, geom.strides)]
        else:
            lists = split_head_outputs(eng, self["cls_out"], self["box_out"], geom, N, [self["scales"][l] for l in range(geom.num)])
            val = lists[("logits_pred", "reg_pred", "reg_pred_std", "ctrness_pred").index(key)]
        self[key] = val
        return val


class _FcosLossFn(torch.autograd.Function):
    """Losses of one FCOS.forward call as ONE autograd node: backward = loss-kernel backward + head backward; returns the
    feature-map gradients (parameter gradients accumulate in the arena)."""

    @staticmethod
    def forward(ctx, fcos, fwd, ctxs, trigger, *xs):
        ctx.fcos, ctx.fwd, ctx.ctxs = fcos, fwd, ctxs
        ctx.in_dtypes = [x.dtype for x in xs]
        return torch.cat([c.losses for c in ctxs]).clone()

    @staticmethod
    def backward(ctx, grad):
        fcos, fwd = ctx.fcos, ctx.fwd
        eng, geom, N = fcos.engine, fwd["geom"], fwd["N"]
        gouts = [grad[4 * i:4 * i + 4].contiguous() for i in range(len(ctx.ctxs))]
        dcls, dbox = fcos.fcos_outputs.backward(fwd, eng.scales, eng.dscales, ctx.ctxs, gouts)
        dfeat = eng.head_backward(fwd["tape"], geom, N, dcls, dbox)
        fwd["tape"] = None
        dx = [nchw(d).to(dt) for d, dt in zip(eng.level_views(dfeat, geom, N, 256), ctx.in_dtypes)]
        return (None, None, None, None) + tuple(dx)


@PROPOSAL_GENERATOR_REGISTRY.register()
class FCOS(nn.Module):
    def __init__(self, cfg, input_shape=None, engine=None, fcos_outputs=None):
        super().__init__()
        f = cfg.MODEL.FCOS
        self.in_features = list(f.IN_FEATURES)
        self.fpn_strides = list(f.FPN_STRIDES)
        self.yield_proposal = f.YIELD_PROPOSAL
        self.kl_loss, self.kl_loss_type = f.KL_LOSS, f.KL_LOSS_TYPE
        self.fcos_head = FCOSHead(cfg, None if input_shape is None else [input_shape[k] for k in self.in_features], engine=engine)
        object.__setattr__(self, "engine", self.fcos_head.engine)
        self.in_channels_to_top_module = self.fcos_head.in_channels_to_top_module
        self.fcos_outputs = fcos_outputs if fcos_outputs is not None else FCOSOutputs(cfg)

    def train(self, mode=True):
        super().train(mode)
        self.fcos_outputs.train(mode)
        return self

    def compute_locations(self, features):
        return [compute_locations(f.shape[-2], f.shape[-1], self.fpn_strides[l], f.device) for l, f in enumerate(features)]

    def forward_head(self, features, top_module=None):
        out = self.fcos_head([features[f] for f in self.in_features], top_module, self.yield_proposal)
        return out[0], out[1], out[3], out[4], out[5]

    def _run_head(self, xs, train):
        eng = self.engine
        N = int(xs[0].shape[0])
        feat, geom = pack_levels(eng, xs, self.fpn_strides, self.fcos_head._soi)
        tape = {} if train else None
        cls_out, box_out = eng.head_forward(feat, geom, N, tape)
        return {"cls_out": cls_out, "box_out": box_out, "geom": geom, "N": N, "tape": tape, "scales": eng.scales, "feat": feat}

    def forward(self, images, features, gt_instances=None, top_module=None, output_raw=False, nms_method="cls_n_ctr",
                ignore_near=False, branch="labeled"):
        if top_module is not None:
            raise NotImplementedError("top_module (AdelaiDet mask branches) is not part of the UT2 recipes")
        xs = [features[f] for f in self.in_features]
        eng = self.engine
        need_loss = self.training and branch in ("labeled", "unlabeled")
        fwd = self._run_head(xs, need_loss)
        fwd["image_sizes"] = [tuple(int(v) for v in s) for s in images.image_sizes]
        raw_output = RawOutput(fwd)
        raw_output["_engine"] = eng
        raw_output["top_feats"], raw_output["bbox_towers"] = [], []
        results, losses = {}, {}
        if self.training:
            if branch == "labeled":
                gt = as_boxset(gt_instances, eng.device)
                vals, ctxs = self.fcos_outputs.losses(fwd, eng.scales, gt, ignore_near)
                names = [[(0, "loss_fcos_cls"), (1, "loss_fcos_loc"), (2, "loss_fcos_ctr")]]
            elif branch == "unlabeled":
                # one_stage_detector.py:170-181: {"cls": instances_class, "reg": instances_reg}
                gt_cls, gt_reg = (gt_instances["cls"], gt_instances["reg"]) if isinstance(gt_instances, dict) else gt_instances
                vals, ctxs = self.fcos_outputs.pseudo_losses(fwd, eng.scales, as_boxset(gt_cls, eng.device), as_boxset(gt_reg, eng.device),
                                                             ignore_near)
                names = [[(0, "loss_fcos_cls"), (2, "loss_fcos_ctr")], [(3, "teacher_better_student"), (1, "loss_fcos_loc")]]
            elif branch == "raw":
                ctxs, names = [], []
            else:
                raise ValueError("Unknown branch")
            if ctxs:
                vec = _FcosLossFn.apply(self, fwd, ctxs, self.fcos_head._trigger, *xs) if torch.is_grad_enabled() else \
                    torch.cat([c.losses for c in ctxs])
                for i, ctx_names in enumerate(names):
                    for j, name in ctx_names:
                        losses[name] = vec[4 * i + j]
            if self.yield_proposal:
                results["proposals"] = dets_to_instances(self.fcos_outputs.predict_proposals(fwd, eng.scales, nms_method))
        else:
            results = dets_to_instances(self.fcos_outputs.predict_proposals(fwd, eng.scales, nms_method))
        if output_raw:
            return results, losses, raw_output
        return results, losses
