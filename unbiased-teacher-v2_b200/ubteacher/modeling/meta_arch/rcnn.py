"""``TwoStagePseudoLabGeneralizedRCNN`` with the reference's call contract
(ubteacher/modeling/meta_arch/rcnn.py:6-72), executed by the B200 engine (rcnn_engine.py).

    model(batched_inputs, branch="supervised" | "unsup_data_weak" | "unsup_data_train") -> 4-tuple
      supervised / unsup_data_train : (losses {loss_cls, loss_box_reg, loss_rpn_cls, loss_rpn_loc}, [], [], None)
      unsup_data_weak               : ({}, proposals_rpn, proposals_roih, ROI_predictions)

Loss tensors are attached to autograd through one custom Function per forward call, so
``sum(losses.values()).backward()`` works like in the reference; the trainer uses the explicit
``forward_train`` / ``backward_pending`` fast path (same kernels, no autograd bookkeeping). Proposals and detections
stay on the device as fixed-capacity sets + counts (``to_instances`` converts, at the price of a host sync).
"""
import torch
from torch import nn

from ...d2compat.registry import META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, ROI_HEADS_REGISTRY
from ...d2compat.structures import Boxes, Instances, detector_postprocess
from ..fcos.fcos_outputs import GT_CAP, BoxSet
from ..rcnn_engine import RcnnEngine

LOSS_NAMES = ("loss_rpn_cls", "loss_rpn_loc", "loss_cls", "loss_box_reg")


def rcnn_boxset(instances, device, cap=GT_CAP):
    """list[Instances] -> BoxSet. Ground truth carries gt_boxes / gt_classes; pseudo labels (trainer.py:727-751) also
    ``scores`` and ``pred_boxes_std`` (stored in BoxSet.reg_pred_std)."""
    N = len(instances)
    n_max = max([len(x) for x in instances] + [1])
    G = max(cap, (n_max + 15) // 16 * 16)
    b = torch.zeros((N, G, 4), dtype=torch.float32).pin_memory()
    c = torch.zeros((N, G), dtype=torch.int64).pin_memory()
    cnt = torch.zeros(N, dtype=torch.int32).pin_memory()
    pseudo = N > 0 and instances[0].has("scores")
    s = torch.zeros((N, G), dtype=torch.float32).pin_memory() if pseudo else None
    d = torch.zeros((N, G, 4), dtype=torch.float32).pin_memory() if pseudo else None
    for i, inst in enumerate(instances):
        n = len(inst)
        cnt[i] = n
        if n:
            b[i, :n] = inst.gt_boxes.tensor.detach().float().cpu()
            c[i, :n] = inst.gt_classes.detach().cpu()
            if pseudo:
                s[i, :n] = inst.scores.detach().float().cpu()
                if inst.has("pred_boxes_std"):
                    d[i, :n] = inst.pred_boxes_std.detach().float().cpu()
    to = lambda t: t.to(device, non_blocking=True) if t is not None else None
    return BoxSet(to(b), to(c), to(cnt), to(d), to(s))


def proposals_to_instances(props, image_sizes):
    cnt = props["count"].cpu().tolist()
    out = []
    for i, n in enumerate(cnt):
        inst = Instances(tuple(image_sizes[i]))
        inst.proposal_boxes = Boxes(props["proposal_boxes"][i, :n])
        inst.objectness_logits = props["objectness_logits"][i, :n]
        out.append(inst)
    return out


def detections_to_instances(dets):
    cnt = dets["count"].cpu().tolist()
    out = []
    for i, n in enumerate(cnt):
        inst = Instances(tuple(dets["image_sizes"][i]))
        inst.pred_boxes = Boxes(dets["pred_boxes"][i, :n])
        inst.scores = dets["scores"][i, :n]
        inst.pred_classes = dets["pred_classes"][i, :n]
        inst.pred_boxes_std = dets["reg_pred_std"][i, :n]
        out.append(inst)
    return out


class _LossGraph(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trigger, model, pending):
        ctx.model, ctx.pending = model, pending
        return pending["vec"].clone()

    @staticmethod
    def backward(ctx, grad):
        g = grad.contiguous()
        ctx.model._run_backward(ctx.pending, g[0:2].contiguous(), g[2:4].contiguous())
        return None, None, None


@META_ARCH_REGISTRY.register()
class TwoStagePseudoLabGeneralizedRCNN(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        assert cfg.MODEL.PROPOSAL_GENERATOR.NAME == "PseudoLabRPN" and cfg.MODEL.ROI_HEADS.NAME == "StandardROIHeadsPseudoLab"
        self.cfg = cfg
        dev = cfg.MODEL.DEVICE
        if not str(dev).startswith("cuda"):
            raise RuntimeError("the UT2 B200 path has no CPU fallback: MODEL.DEVICE must be a CUDA device")
        self.engine = RcnnEngine(cfg, device=dev, seed=max(cfg.SEED, 0))
        # the reference's sub-modules, looked up by their registry names ([D2] GeneralizedRCNN.from_config), as views that
        # share this replica's engine: model.proposal_generator(images, features, gt) / model.roi_heads(images, features, ...)
        from ..proposal_generator import rpn as _rpn  # noqa: F401  (registers PseudoLabRPN)
        from ..roi_heads import roi_heads as _rh  # noqa: F401  (registers StandardROIHeadsPseudoLab)
        from ..backbone.fpn import RcnnResnetFpnBackbone
        self.backbone = RcnnResnetFpnBackbone(cfg, None, engine=self.engine)
        self.proposal_generator = PROPOSAL_GENERATOR_REGISTRY.get(cfg.MODEL.PROPOSAL_GENERATOR.NAME)(cfg, None, engine=self.engine)
        self.roi_heads = ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, None, engine=self.engine)
        self._trigger = torch.zeros(1, device=dev, requires_grad=True)
        self._params = None
        self._gout_cache = {}

    # ---- nn.Module surface backed by the arena ---------------------------------------------------
    @property
    def device(self):
        return self.engine.device

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        sd = self.engine.arena.state_dict(prefix)
        if destination is not None:
            destination.update(sd)
            return destination
        return sd

    def load_state_dict(self, state_dict, strict=True):
        missing = self.engine.arena.load_state_dict(state_dict, strict)
        self.engine.refresh_operands()
        return missing

    def named_parameters(self, prefix="", recurse=True, remove_duplicate=True):
        if self._params is None:
            A = self.engine.arena
            self._params = []
            for n, sp in A.specs.items():
                if n.startswith("_") or sp.group == "buffer":
                    continue
                p = nn.Parameter(A.views[n], requires_grad=sp.group in ("decay", "nodecay"))
                if p.requires_grad:
                    p.grad = A.gviews[n]
                self._params.append((n, p))
        for n, p in self._params:
            yield prefix + n, p

    def parameters(self, recurse=True):
        for _, p in self.named_parameters():
            yield p

    # ---- forward ---------------------------------------------------------------------------------
    def _images(self, batched_inputs):
        return [x["image"].to(self.device, non_blocking=True) for x in batched_inputs]

    def preprocess_image(self, batched_inputs):
        from ..one_stage_detector import U8Images
        return U8Images(self._images(batched_inputs), self.engine.pixel_mean, self.engine.pixel_std)

    def _gt(self, batched_inputs):
        g = batched_inputs[0]["instances"]
        return g if isinstance(g, BoxSet) else rcnn_boxset([x["instances"] for x in batched_inputs], self.device)

    def forward(self, batched_inputs, branch="supervised", given_proposals=None, val_mode=False):
        if (not self.training) and (not val_mode):
            return self.inference(batched_inputs)
        if branch in ("supervised", "unsup_data_train"):
            losses, pending = self.forward_train(batched_inputs, branch)
            vec = _LossGraph.apply(self._trigger, self, pending)
            return {k: vec[i] for i, k in enumerate(LOSS_NAMES)}, [], [], None
        if branch == "unsup_data_weak":
            props, dets, pred = self.forward_teacher(batched_inputs)
            return {}, props, dets, pred
        raise ValueError(f"unknown branch {branch}")

    def inference(self, batched_inputs):
        """[D2] GeneralizedRCNN.inference (eval mode): RPN with the *_TEST top-k values, box head, fast_rcnn_inference,
        then [D2] detector_postprocess to the dataset dict's height / width -> [{"instances": Instances}]."""
        eng = self.engine
        fwd = eng.forward_features(self._images(batched_inputs), train=False)
        _, dets, _ = eng.forward_inference(fwd, test=True)
        out = []
        for d, r in zip(batched_inputs, detections_to_instances(dets)):
            out.append({"instances": detector_postprocess(r, d.get("height", r.image_size[0]), d.get("width", r.image_size[1]))})
        return out

    def forward_teacher(self, batched_inputs):
        """unsup_data_weak: (proposals_rpn, proposals_roih, ROI_predictions), all device-resident."""
        eng = self.engine
        fwd = eng.forward_features(self._images(batched_inputs), train=False)
        return eng.forward_inference(fwd)

    def forward_train(self, batched_inputs, branch):
        """Explicit (autograd-free) training forward: (dict of detached loss scalars, pending ctx)."""
        if "instances" not in batched_inputs[0]:
            raise ValueError("training branches need ground truth or pseudo labels under 'instances'")
        eng = self.engine
        gt = self._gt(batched_inputs)
        pseudo = gt.scores is not None           # roi_heads.py:92, rpn.py:49: targets carrying `scores` are pseudo labels
        fwd = eng.forward_features(self._images(batched_inputs), train=True)
        rpn_l, roi_l, ctx = eng.forward_losses(fwd, gt, pseudo)
        vec = torch.cat([rpn_l, roi_l])
        losses = {k: vec[i] for i, k in enumerate(LOSS_NAMES)}
        return losses, {"fwd": fwd, "ctx": ctx, "vec": vec}

    def _run_backward(self, pending, gout_rpn, gout_roi):
        self.engine.backward(pending["fwd"], pending["ctx"], gout_rpn, gout_roi)
        pending["fwd"]["tape"] = None

    def backward_pending(self, pending, weights):
        """weights: [w_rpn_cls, w_rpn_loc, w_cls, w_box_reg] (the trainer's loss weighting, trainer.py:880-905)."""
        key = tuple(float(v) for v in weights)
        if key not in self._gout_cache:
            t = torch.tensor(key, dtype=torch.float32, device=self.device)
            self._gout_cache[key] = (t[0:2].contiguous(), t[2:4].contiguous())
        g_rpn, g_roi = self._gout_cache[key]
        self._run_backward(pending, g_rpn, g_roi)
