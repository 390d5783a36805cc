"""``PseudoLabRPN`` with the reference's call contract (ubteacher/modeling/proposal_generator/rpn.py:15-225):

    rpn = PROPOSAL_GENERATOR_REGISTRY.get("PseudoLabRPN")(cfg, input_shape)
    proposals, losses = rpn(images, features, gt_instances=None, compute_loss=True, compute_val_loss=False)

features: {"p2".."p6": [N, 256, H_l, W_l]}; images: anything with ``.image_sizes``; gt_instances: list[Instances]
(gt_boxes; pseudo labels also carry ``scores`` -> objectness BCE weighted by the matched teacher score, rpn.py:49-57).
Returns list[Instances] (proposal_boxes, objectness_logits) and {"loss_rpn_cls", "loss_rpn_loc"}. Executed by the engine:
[D2] StandardRPNHead as two level-major tensor-core launches, anchor labelling / sampling / losses / top-k + NMS kernels
(csrc/rpn.cu, csrc/nms.cu). The losses hang on one autograd node whose backward returns the feature-map gradients.
"""
import torch

from ...d2compat.registry import PROPOSAL_GENERATOR_REGISTRY
from ..views import ArenaView, nchw

BF16 = torch.bfloat16


def _own_engine(cfg):
    from ..roi_heads.fast_rcnn import _own_engine as mk
    return mk(cfg)


def pack_pyramid(engine, xs, geom):
    N, C = xs[0].shape[0], xs[0].shape[1]
    base = getattr(xs[0], "_ut2_level_major", None)
    if base is not None and all(getattr(x, "_ut2_level_major", None) is base for x in xs) and base.shape == (geom.L * N, C):
        return base
    feat = torch.empty((geom.L * N, C), dtype=BF16, device=engine.device)
    for x, lv in zip(xs, engine.level_views(feat, geom, N, C)):
        lv.copy_(x.detach().permute(0, 2, 3, 1))
    return feat


class _RpnLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rpn, fwd, lctx, losses, trigger, *xs):
        ctx.rpn, ctx.fwd, ctx.lctx = rpn, fwd, lctx
        ctx.dtypes = [x.dtype for x in xs]
        return losses.clone()

    @staticmethod
    def backward(ctx, g):
        eng, fwd = ctx.rpn.engine, ctx.fwd
        dfe = eng.rpn_backward(fwd, ctx.lctx, g.contiguous().float())
        fwd["tape"] = None
        dx = [nchw(d).to(dt) for d, dt in zip(eng.level_views(dfe, fwd["geom"], fwd["N"], 256), ctx.dtypes)]
        return (None, None, None, None, None) + tuple(dx)


@PROPOSAL_GENERATOR_REGISTRY.register()
class PseudoLabRPN(ArenaView):
    def __init__(self, cfg, input_shape=None, engine=None):
        super().__init__(engine if engine is not None else _own_engine(cfg), "proposal_generator.")
        self.in_features = list(cfg.MODEL.RPN.IN_FEATURES)
        self.loss_weight = {"loss_rpn_cls": cfg.MODEL.RPN.LOSS_WEIGHT, "loss_rpn_loc": cfg.MODEL.RPN.BBOX_REG_LOSS_WEIGHT * cfg.MODEL.RPN.LOSS_WEIGHT}
        self._trigger = torch.zeros(1, device=self.engine.device, requires_grad=True)

    def forward(self, images, features, gt_instances=None, compute_loss=True, compute_val_loss=False):
        from ..meta_arch.rcnn import proposals_to_instances, rcnn_boxset
        from ..fcos.fcos_outputs import BoxSet
        eng = self.engine
        xs = [features[f] for f in self.in_features]
        N = int(xs[0].shape[0])
        sizes = [tuple(int(v) for v in s) for s in images.image_sizes]
        Hp, Wp = int(xs[0].shape[2]) * eng.rpn_strides[0], int(xs[0].shape[3]) * eng.rpn_strides[0]
        geom, rgeom = eng.level_geom(Hp, Wp)
        assert [tuple(x.shape[2:]) for x in xs] == geom.hw, "PseudoLabRPN: p2..p6 of one /32-padded batch expected"
        need_loss = (self.training and compute_loss) or compute_val_loss
        tape = {} if need_loss else None
        feat = pack_pyramid(eng, xs, geom)
        rpn_out = eng.rpn_head_forward(feat, geom, N, tape)
        fwd = {"rpn_out": rpn_out, "geom": geom, "rgeom": rgeom, "N": N, "image_sizes": sizes, "image_hw": eng.image_hw(sizes),
               "tape": tape, "feat": feat}
        losses = {}
        if need_loss:
            gt = gt_instances if isinstance(gt_instances, BoxSet) else rcnn_boxset(gt_instances, eng.device)
            vec, lctx = eng.rpn_losses(fwd, gt, gt.scores is not None)
            if torch.is_grad_enabled():
                vec = _RpnLossFn.apply(self, fwd, lctx, vec, self._trigger, *xs)
            losses = {"loss_rpn_cls": vec[0] * self.loss_weight["loss_rpn_cls"], "loss_rpn_loc": vec[1] * self.loss_weight["loss_rpn_loc"]}
        props = eng.proposals(fwd, test=not self.training)
        self.last_proposals = props                      # device-resident fixed-capacity form (no host sync)
        return proposals_to_instances(props, sizes), losses
