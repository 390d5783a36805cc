"""A small yacs-compatible ``CfgNode`` plus the subset of Detectron2 v0.6 defaults that the UT2 hot
path reads. ``get_cfg()`` mirrors ``detectron2.config.get_cfg``; YAML files support
``_BASE_`` inheritance and ``KEY VALUE`` list overrides like the reference CLI (train_net.py:15-26).
"""
import copy
import os
from ast import literal_eval

import yaml

BASE_KEY = "_BASE_"


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        object.__setattr__(self, "_frozen", False)
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # attribute access -----------------------------------------------------------------
    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if object.__getattribute__(self, "_frozen"):
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = value

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def is_frozen(self):
        return object.__getattribute__(self, "_frozen")

    def _set_frozen(self, flag):
        object.__setattr__(self, "_frozen", flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    # merging ----------------------------------------------------------------------------
    @staticmethod
    def load_yaml_with_base(filename):
        with open(filename, "r") as f:
            cfg = yaml.safe_load(f) or {}
        cfg = _tuplify(cfg)
        if BASE_KEY in cfg:
            base = cfg.pop(BASE_KEY)
            if base.startswith("~"):
                base = os.path.expanduser(base)
            if not any(map(base.startswith, ["/", "https://", "http://"])):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = CfgNode.load_yaml_with_base(base)
            _merge_dict(cfg, base_cfg)
            return base_cfg
        return cfg

    def merge_from_file(self, cfg_filename, allow_unsafe=False):
        loaded = CfgNode.load_yaml_with_base(cfg_filename)
        self.merge_from_other_cfg(CfgNode(loaded))

    def merge_from_other_cfg(self, other):
        _merge_into(other, self, [])

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0, "Override list has odd length: {}".format(cfg_list)
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            d = self
            keys = full_key.split(".")
            for sub in keys[:-1]:
                assert sub in d, "Non-existent key: {}".format(full_key)
                d = d[sub]
            assert keys[-1] in d, "Non-existent key: {}".format(full_key)
            value = _decode(v)
            value = _coerce(value, d[keys[-1]], full_key)
            dict.__setitem__(d, keys[-1], value)

    def dump(self, **kwargs):
        def to_dict(n):
            return {k: to_dict(v) if isinstance(v, CfgNode) else (list(v) if isinstance(v, tuple) else v)
                    for k, v in n.items()}

        return yaml.safe_dump(to_dict(self), **kwargs)


def _tuplify(x):
    """yacs decodes string values such as "(1, 2)" with literal_eval; emulate for YAML scalars."""
    if isinstance(x, dict):
        return {k: _tuplify(v) for k, v in x.items()}
    if isinstance(x, str):
        return _decode(x)
    return x


def _decode(v):
    if not isinstance(v, str):
        return v
    try:
        return literal_eval(v)
    except (ValueError, SyntaxError):
        return v


def _coerce(new, old, key):
    if old is None or new is None or type(new) == type(old):
        return new
    for a, b in ((list, tuple), (tuple, list)):
        if isinstance(new, a) and isinstance(old, b):
            return b(new)
    if isinstance(old, float) and isinstance(new, int):
        return float(new)
    if isinstance(old, CfgNode) and isinstance(new, dict):
        return CfgNode(new)
    raise ValueError(f"Type mismatch ({type(old)} vs. {type(new)}) for config key: {key}")


def _merge_dict(src, dst):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_dict(v, dst[k])
        else:
            dst[k] = v


def _merge_into(a, b, key_list):
    for k, v_ in a.items():
        full_key = ".".join(key_list + [k])
        v = copy.deepcopy(v_)
        if k in b:
            if isinstance(v, CfgNode) and isinstance(b[k], CfgNode):
                _merge_into(v, b[k], key_list + [k])
            else:
                dict.__setitem__(b, k, _coerce(v, b[k], full_key))
        else:
            raise KeyError("Non-existent config key: {}".format(full_key))


# ------------------------------------------------------------------------------------------
# Detectron2 v0.6 defaults (subset; values per SURVEY.md A.1 — D2 itself is not on disk).
_D2_DEFAULTS = {
    "VERSION": 2,
    "MODEL": {
        "LOAD_PROPOSALS": False, "MASK_ON": False, "KEYPOINT_ON": False, "DEVICE": "cuda",
        "META_ARCHITECTURE": "GeneralizedRCNN", "WEIGHTS": "",
        "PIXEL_MEAN": [103.530, 116.280, 123.675], "PIXEL_STD": [1.0, 1.0, 1.0],
        "BACKBONE": {"NAME": "build_resnet_backbone", "FREEZE_AT": 2},
        "FPN": {"IN_FEATURES": [], "OUT_CHANNELS": 256, "NORM": "", "FUSE_TYPE": "sum"},
        "PROPOSAL_GENERATOR": {"NAME": "RPN", "MIN_SIZE": 0},
        "ANCHOR_GENERATOR": {"NAME": "DefaultAnchorGenerator", "SIZES": [[32, 64, 128, 256, 512]],
                             "ASPECT_RATIOS": [[0.5, 1.0, 2.0]], "ANGLES": [[-90, 0, 90]], "OFFSET": 0.0},
        "RPN": {"HEAD_NAME": "StandardRPNHead", "IN_FEATURES": ["res4"], "BOUNDARY_THRESH": -1,
                "IOU_THRESHOLDS": [0.3, 0.7], "IOU_LABELS": [0, -1, 1], "BATCH_SIZE_PER_IMAGE": 256,
                "POSITIVE_FRACTION": 0.5, "BBOX_REG_LOSS_TYPE": "smooth_l1", "BBOX_REG_LOSS_WEIGHT": 1.0,
                "BBOX_REG_WEIGHTS": (1.0, 1.0, 1.0, 1.0), "SMOOTH_L1_BETA": 0.0, "LOSS_WEIGHT": 1.0,
                "PRE_NMS_TOPK_TRAIN": 12000, "PRE_NMS_TOPK_TEST": 6000, "POST_NMS_TOPK_TRAIN": 2000,
                "POST_NMS_TOPK_TEST": 1000, "NMS_THRESH": 0.7, "CONV_DIMS": [-1]},
        "ROI_HEADS": {"NAME": "Res5ROIHeads", "NUM_CLASSES": 80, "IN_FEATURES": ["res4"],
                      "IOU_THRESHOLDS": [0.5], "IOU_LABELS": [0, 1], "BATCH_SIZE_PER_IMAGE": 512,
                      "POSITIVE_FRACTION": 0.25, "SCORE_THRESH_TEST": 0.05, "NMS_THRESH_TEST": 0.5,
                      "PROPOSAL_APPEND_GT": True},
        "ROI_BOX_HEAD": {"NAME": "", "BBOX_REG_LOSS_TYPE": "smooth_l1", "BBOX_REG_LOSS_WEIGHT": 1.0,
                         "BBOX_REG_WEIGHTS": (10.0, 10.0, 5.0, 5.0), "SMOOTH_L1_BETA": 0.0,
                         "POOLER_RESOLUTION": 14, "POOLER_SAMPLING_RATIO": 0, "POOLER_TYPE": "ROIAlignV2",
                         "NUM_FC": 0, "FC_DIM": 1024, "NUM_CONV": 0, "CONV_DIM": 256, "NORM": "",
                         "CLS_AGNOSTIC_BBOX_REG": False, "TRAIN_ON_PRED_BOXES": False},
        "RESNETS": {"DEPTH": 50, "OUT_FEATURES": ["res4"], "NUM_GROUPS": 1, "NORM": "FrozenBN",
                    "WIDTH_PER_GROUP": 64, "STRIDE_IN_1X1": True, "RES5_DILATION": 1,
                    "RES2_OUT_CHANNELS": 256, "STEM_OUT_CHANNELS": 64,
                    "DEFORM_ON_PER_STAGE": [False, False, False, False], "DEFORM_MODULATED": False,
                    "DEFORM_NUM_GROUPS": 1},
    },
    "INPUT": {"MIN_SIZE_TRAIN": (800,), "MIN_SIZE_TRAIN_SAMPLING": "choice", "MAX_SIZE_TRAIN": 1333,
              "MIN_SIZE_TEST": 800, "MAX_SIZE_TEST": 1333, "RANDOM_FLIP": "horizontal", "FORMAT": "BGR",
              "MASK_FORMAT": "polygon",
              "CROP": {"ENABLED": False, "TYPE": "relative_range", "SIZE": [0.9, 0.9]}},
    "DATASETS": {"TRAIN": (), "PROPOSAL_FILES_TRAIN": (), "PRECOMPUTED_PROPOSAL_TOPK_TRAIN": 2000,
                 "TEST": (), "PROPOSAL_FILES_TEST": (), "PRECOMPUTED_PROPOSAL_TOPK_TEST": 1000},
    "DATALOADER": {"NUM_WORKERS": 4, "ASPECT_RATIO_GROUPING": True, "SAMPLER_TRAIN": "TrainingSampler",
                   "REPEAT_THRESHOLD": 0.0, "FILTER_EMPTY_ANNOTATIONS": True},
    "SOLVER": {"LR_SCHEDULER_NAME": "WarmupMultiStepLR", "MAX_ITER": 40000, "BASE_LR": 0.001,
               "MOMENTUM": 0.9, "NESTEROV": False, "WEIGHT_DECAY": 0.0001, "WEIGHT_DECAY_NORM": 0.0,
               "GAMMA": 0.1, "STEPS": (30000,), "WARMUP_FACTOR": 1.0 / 1000, "WARMUP_ITERS": 1000,
               "WARMUP_METHOD": "linear", "CHECKPOINT_PERIOD": 5000, "IMS_PER_BATCH": 16,
               "REFERENCE_WORLD_SIZE": 0, "BIAS_LR_FACTOR": 1.0, "WEIGHT_DECAY_BIAS": None,
               "CLIP_GRADIENTS": {"ENABLED": False, "CLIP_TYPE": "value", "CLIP_VALUE": 1.0, "NORM_TYPE": 2.0},
               "AMP": {"ENABLED": False}},
    "TEST": {"EXPECTED_RESULTS": [], "EVAL_PERIOD": 0, "KEYPOINT_OKS_SIGMAS": [], "DETECTIONS_PER_IMAGE": 100,
             "AUG": {"ENABLED": False}, "PRECISE_BN": {"ENABLED": False, "NUM_ITER": 200}},
    "OUTPUT_DIR": "./output", "SEED": -1, "CUDNN_BENCHMARK": False, "VIS_PERIOD": 0,
    "GLOBAL": {"HACK": 1.0},
}


def get_cfg():
    return CfgNode(copy.deepcopy(_D2_DEFAULTS))
