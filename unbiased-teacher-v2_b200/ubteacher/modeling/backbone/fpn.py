"""``build_fcos_resnet_fpn_backbone`` / ``LastLevelP6P7`` with the reference's names
(ubteacher/modeling/backbone/fpn.py:11-29, :50-78): ResNet-50 (FrozenBN, stride in the 1x1, frozen at res2) + FPN over
res3..res5 + P6 = conv3x3s2(p5), P7 = conv3x3s2(relu(P6)), executed by the B200 engine's trunk / FPN kernels.

    backbone = BACKBONE_REGISTRY.get("build_fcos_resnet_fpn_backbone")(cfg, input_shape)
    backbone(images) -> {"p3": [N, 256, H/8, W/8], ..., "p7": ...}     (bf16, channels-last views of one level-major buffer)

``images``: the batch's uint8 BGR images — a list of [3, h, w] tensors or an object with ``.images_u8`` /
``.image_sizes`` (``OneStageDetector.preprocess_image``). Pixel normalisation and the /32 zero padding of [D2]
``ImageList.from_tensors`` are fused into the stem kernel, so the backbone consumes the raw images rather than a
normalised float tensor (the one deviation from the reference call: documented in INTEGRATION.md). Inference-only view:
training back-propagates through the meta-architecture's explicit schedule.
"""
import torch

from ...d2compat.registry import BACKBONE_REGISTRY
from ...d2compat.structures import ShapeSpec
from ..views import ArenaView, nchw


class LastLevelP6P7(ArenaView):
    """backbone/fpn.py:11-29. forward(p5) -> [p6, p7] (NCHW views)."""

    def __init__(self, in_channels=256, out_channels=256, in_features="p5", engine=None, cfg=None):
        if engine is None:
            from ..fcos.fcos import _own_engine
            engine = _own_engine(cfg)
        super().__init__(engine, "backbone.top_block.")
        assert in_channels == 256 and out_channels == 256
        self.num_levels, self.in_feature = 2, in_features

    @torch.no_grad()
    def forward(self, x):
        from ... import ops
        eng = self.engine
        p5 = x.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        p6 = eng.p6.fwd(p5)
        p7 = eng.p7.fwd(ops.relu_bwd(p6, p6))
        return [nchw(p6), nchw(p7)]


class FcosResnetFpnBackbone(ArenaView):
    def __init__(self, cfg, input_shape=None, engine=None):
        if engine is None:
            from ..fcos.fcos import _own_engine
            engine = _own_engine(cfg)
        super().__init__(engine, "backbone.")
        self._out_features = ["p3", "p4", "p5", "p6", "p7"]
        self._strides = dict(zip(self._out_features, (8, 16, 32, 64, 128)))
        self.top_block = LastLevelP6P7(engine=engine)
        self.size_divisibility = 32

    def output_shape(self):
        return {k: ShapeSpec(channels=256, stride=s) for k, s in self._strides.items()}

    @torch.no_grad()
    def forward(self, images):
        eng = self.engine
        imgs = getattr(images, "images_u8", images)
        if isinstance(imgs, torch.Tensor):
            if imgs.dtype != torch.uint8:
                raise TypeError("the B200 backbone fuses pixel normalisation into the stem kernel: pass the uint8 images "
                                "(list of [3, h, w] tensors or OneStageDetector.preprocess_image(batched_inputs)), not a normalised tensor")
            imgs = list(imgs)
        imgs = [im.to(eng.device, non_blocking=True) for im in imgs]
        N = len(imgs)
        feats, sizes, (Hp, Wp) = eng.trunk_forward(imgs, False, None)
        geom = eng.level_geom(Hp, Wp)
        feat = eng.fpn_forward(feats, geom, N, None)
        out = {}
        for k, lv in zip(self._out_features, eng.level_views(feat, geom, N, 256)):
            v = nchw(lv)
            v._ut2_level_major = feat          # lets FCOSHead / FCOS consume the pyramid without re-packing it
            out[k] = v
        return out


@BACKBONE_REGISTRY.register()
def build_fcos_resnet_fpn_backbone(cfg, input_shape=None, engine=None):
    return FcosResnetFpnBackbone(cfg, input_shape, engine)


class RcnnResnetFpnBackbone(ArenaView):
    """[D2] ``build_resnet_fpn_backbone`` (configs/Faster-RCNN/Base-RCNN-FPN.yaml:4): ResNet-50 + FPN over res2..res5 +
    LastLevelMaxPool, as a view over an RcnnEngine. backbone(images) -> {"p2".."p6"} (same input contract as above)."""

    def __init__(self, cfg, input_shape=None, engine=None):
        if engine is None:
            from ..roi_heads.fast_rcnn import _own_engine
            engine = _own_engine(cfg)
        super().__init__(engine, "backbone.")
        self._out_features = ["p2", "p3", "p4", "p5", "p6"]
        self._strides = dict(zip(self._out_features, (4, 8, 16, 32, 64)))
        self.size_divisibility = 32

    def output_shape(self):
        return {k: ShapeSpec(channels=256, stride=s) for k, s in self._strides.items()}

    @torch.no_grad()
    def forward(self, images):
        eng = self.engine
        imgs = getattr(images, "images_u8", images)
        if isinstance(imgs, torch.Tensor):
            if imgs.dtype != torch.uint8:
                raise TypeError("the B200 backbone fuses pixel normalisation into the stem kernel: pass the uint8 images")
            imgs = list(imgs)
        imgs = [im.to(eng.device, non_blocking=True) for im in imgs]
        N = len(imgs)
        feats, sizes, (Hp, Wp) = eng.trunk_forward(imgs, False, None)
        geom, _ = eng.level_geom(Hp, Wp)
        feat, levels = eng.fpn_forward(feats, geom, N, None)
        out = {}
        for k, lv in zip(self._out_features, levels):
            v = nchw(lv)
            v._ut2_level_major = feat
            out[k] = v
        return out


@BACKBONE_REGISTRY.register()
def build_resnet_fpn_backbone(cfg, input_shape=None, engine=None):
    return RcnnResnetFpnBackbone(cfg, input_shape, engine)
