"""``OneStageDetector`` / ``PseudoProposalNetwork`` with the reference's call contract
(ubteacher/modeling/one_stage_detector.py:46-240), executed by the B200 engine (fcos_engine.py).

    model(batched_inputs, output_raw=False, nms_method="cls_n_ctr", ignore_near=False, branch="labeled")

train mode  -> losses dict                       | (losses, raw_pred, proposals) when output_raw
eval mode   -> [{"instances": Instances}, ...]    | (proposals, raw_pred)        when output_raw

Loss tensors are attached to autograd through one custom Function per forward call, so
``sum(losses.values()).backward()`` works like in the reference; the trainer uses the explicit
``backward_pending`` fast path instead (same kernels, no autograd bookkeeping).
"""
import torch
from torch import nn

from ..d2compat.registry import BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY
from ..d2compat.structures import Instances, detector_postprocess
from .backbone import fpn as _fpn  # noqa: F401  (registers build_fcos_resnet_fpn_backbone)
from .fcos import fcos as _fcos  # noqa: F401  (registers FCOS)
from .fcos.fcos_outputs import BoxSet, FCOSOutputs, as_boxset, dets_to_instances
from .fcos_engine import FcosEngine


class U8Images:
    """What ``preprocess_image`` returns here: the batch's uint8 BGR images on the device + their sizes. The reference
    builds a normalised, zero-padded float ``ImageList`` (one_stage_detector.py:165-167); normalisation and padding are
    fused into the stem kernel, so only ``image_sizes`` (and, lazily, ``tensor``) of that interface are kept."""

    def __init__(self, images_u8, pixel_mean, pixel_std, div=32):
        self.images_u8 = images_u8
        self.image_sizes = [(int(im.shape[1]), int(im.shape[2])) for im in images_u8]
        self._mean, self._std, self._div = pixel_mean, pixel_std, div

    def __len__(self):
        return len(self.images_u8)

    @property
    def tensor(self):
        """The reference's normalised, zero-padded [N, 3, H, W] float batch (compatibility only: not used by the engine)."""
        from ..d2compat.structures import ImageList
        dev = self.images_u8[0].device
        m = torch.tensor(self._mean, device=dev).view(3, 1, 1)
        s = torch.tensor(self._std, device=dev).view(3, 1, 1)
        return ImageList.from_tensors([(im.float() - m) / s for im in self.images_u8], self._div).tensor


class _LossGraph(torch.autograd.Function):
    """Autograd bridge: outputs the concatenated loss vectors of one forward; backward runs the engine's
    explicit backward schedule with the incoming gradient as the loss weights."""

    @staticmethod
    def forward(ctx, trigger, model, pending):
        ctx.model, ctx.pending = model, pending
        return torch.cat([c.losses for c in pending["ctxs"]]).clone()

    @staticmethod
    def backward(ctx, grad):
        gouts = [grad[4 * i:4 * i + 4].contiguous() for i in range(len(ctx.pending["ctxs"]))]
        ctx.model._run_backward(ctx.pending, gouts)
        return None, None, None


@META_ARCH_REGISTRY.register()
class PseudoProposalNetwork(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        assert cfg.MODEL.PROPOSAL_GENERATOR.NAME == "FCOS" and cfg.MODEL.BACKBONE.NAME == "build_fcos_resnet_fpn_backbone"
        self.cfg = cfg
        dev = cfg.MODEL.DEVICE
        if not str(dev).startswith("cuda"):
            raise RuntimeError("the UT2 B200 path has no CPU fallback: MODEL.DEVICE must be a CUDA device")
        self.engine = FcosEngine(cfg, device=dev, seed=max(cfg.SEED, 0))
        self.fcos_outputs = FCOSOutputs(cfg)
        self.yield_proposal = cfg.MODEL.FCOS.YIELD_PROPOSAL
        # the reference's sub-modules, looked up by their registry names (one_stage_detector.py:52-55), as views that
        # share this replica's engine: model.backbone(images) / model.proposal_generator(images, features, gt, ...)
        self.backbone = BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, None, engine=self.engine)
        self.proposal_generator = PROPOSAL_GENERATOR_REGISTRY.get(cfg.MODEL.PROPOSAL_GENERATOR.NAME)(
            cfg, self.backbone.output_shape(), engine=self.engine, fcos_outputs=self.fcos_outputs)
        self._trigger = torch.zeros(1, device=dev, requires_grad=True)
        self._params = None
        self._gout_cache = {}
        self.last_proposals = None

    # ---- nn.Module surface backed by the arena -------------------------------------------------
    @property
    def device(self):
        return self.engine.device

    def train(self, mode=True):
        super().train(mode)
        self.fcos_outputs.train(mode)
        return self

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

    # ---- forward ----------------------------------------------------------------------------------
    def _images(self, batched_inputs):
        return [x["image"].to(self.device, non_blocking=True) for x in batched_inputs]

    def preprocess_image(self, batched_inputs):
        return U8Images(self._images(batched_inputs), self.engine.pixel_mean, self.engine.pixel_std)

    def _scales(self):
        return self.engine.scales

    def _predict(self, fwd, nms_method, output_raw, batched_inputs=None):
        dets = self.fcos_outputs.predict_proposals(fwd, self._scales(), nms_method)
        if output_raw:
            return dets, fwd          # "output raw will not rescale" (one_stage_detector.py:131-133)
        out = []
        for i, r in enumerate(dets_to_instances(dets)):      # standard output rescales to the dataset dict's height / width
            d = batched_inputs[i] if batched_inputs is not None else {}
            out.append({"proposals": detector_postprocess(r, d.get("height", r.image_size[0]), d.get("width", r.image_size[1]))})
        return out

    def forward(self, batched_inputs, output_raw=False, nms_method="cls_n_ctr", ignore_near=False, branch="labeled"):
        fwd = self.engine.forward(self._images(batched_inputs), train=False)
        return self._predict(fwd, nms_method, output_raw, batched_inputs)

    def _run_backward(self, pending, gouts):
        eng = self.engine
        dcls, dbox = self.fcos_outputs.backward(pending["fwd"], eng.scales, eng.dscales, pending["ctxs"], gouts)
        eng.backward(pending["fwd"], dcls, dbox)
        pending["fwd"]["tape"] = None       # free the activations


@META_ARCH_REGISTRY.register()
class OneStageDetector(PseudoProposalNetwork):
    def forward(self, batched_inputs, output_raw=False, nms_method="cls_n_ctr", ignore_near=False, branch="labeled"):
        if not self.training:
            if output_raw:
                return super().forward(batched_inputs, output_raw, nms_method, branch=branch)
            return [{"instances": r["proposals"]} for r in super().forward(batched_inputs, False, nms_method, branch=branch)]
        losses, pending = self.forward_train(batched_inputs, branch, ignore_near)
        vec = _LossGraph.apply(self._trigger, self, pending)
        out, i = {}, 0
        for ctx_names in pending["names"]:
            for j, name in ctx_names:
                out[name] = vec[4 * i + j]
            i += 1
        if output_raw:
            return out, pending["fwd"], self.last_proposals
        return out

    def forward_train(self, batched_inputs, branch, ignore_near=False):
        """Explicit (autograd-free) training forward: returns (dict of detached loss scalars, pending ctx)."""
        eng = self.engine
        fwd = eng.forward(self._images(batched_inputs), train=True)
        b0 = batched_inputs[0]
        if "instances_class" in b0 and "instances_reg" in b0:
            gt_cls = b0["instances_class"] if isinstance(b0["instances_class"], BoxSet) else \
                as_boxset([x["instances_class"] for x in batched_inputs], self.device)
            gt_reg = b0["instances_reg"] if isinstance(b0["instances_reg"], BoxSet) else \
                as_boxset([x["instances_reg"] for x in batched_inputs], self.device)
            if branch != "unlabeled":
                raise ValueError("Incorrect branch name")
            losses, ctxs = self.fcos_outputs.pseudo_losses(fwd, eng.scales, gt_cls, gt_reg, ignore_near)
            names = [[(0, "loss_fcos_cls"), (2, "loss_fcos_ctr")], [(3, "teacher_better_student"), (1, "loss_fcos_loc")]]
        elif "instances" in b0 and branch != "teacher_weak":
            if branch != "labeled":
                raise ValueError("Incorrect branch name")
            gt = b0["instances"] if isinstance(b0["instances"], BoxSet) else \
                as_boxset([x["instances"] for x in batched_inputs], self.device)
            losses, ctxs = self.fcos_outputs.losses(fwd, eng.scales, gt, ignore_near)
            names = [[(0, "loss_fcos_cls"), (1, "loss_fcos_loc"), (2, "loss_fcos_ctr")]]
        else:
            raise ValueError("Unknown branch")
        if self.yield_proposal:   # fcos.py:176-187 — the reference also decodes proposals on the student (no_grad)
            self.last_proposals = self.fcos_outputs.predict_proposals(fwd, eng.scales, "cls_n_ctr")
        return losses, {"fwd": fwd, "ctxs": ctxs, "names": names}

    def backward_pending(self, pending, weights):
        """weights: one [w_cls, w_loc, w_ctr, 0] float list per loss ctx (the trainer's loss weighting)."""
        gouts = []
        for w in weights:       # loss weights are config constants: keep them resident (also graph-capture safe)
            key = tuple(float(v) for v in w)
            if key not in self._gout_cache:
                self._gout_cache[key] = torch.tensor(key, dtype=torch.float32, device=self.device)
            gouts.append(self._gout_cache[key])
        self._run_backward(pending, gouts)
