"""``DetectionTSCheckpointer`` (reference: ubteacher/checkpoint/detection_checkpoint.py:10-89 on top of [D2]
``DetectionCheckpointer`` / fvcore ``Checkpointer``): SURVEY.md §8(f) rank 2.

Same behaviour for the two cases the reference distinguishes:
  * a Caffe2 pickle (``detectron2://ImageNetPretrained/MSRA/R-50.pkl``, ``__author__ == "Caffe2"``): blob names are
    converted to Detectron2 names and loaded into the STUDENT only, by suffix matching against its state_dict
    ([D2] ``align_and_update_state_dicts``); missing ``pixel_mean`` / ``pixel_std`` are not reported;
  * anything else: a whole ``EnsembleTSModel`` state with ``modelTeacher.`` / ``modelStudent.`` prefixes.
Checkpoints written by ``save`` hold {"model", "optimizer", "scheduler", "iteration", ...} like fvcore's, so files are
interchangeable with the reference (tensor names and shapes are the reference's: ParamArena.state_dict)."""
import logging
import os
import pickle
import re
from collections import namedtuple

import numpy as np
import torch

logger = logging.getLogger(__name__)
IncompatibleKeys = namedtuple("IncompatibleKeys", ["missing_keys", "unexpected_keys", "incorrect_shapes"])


def convert_c2_resnet_names(blobs):
    """Caffe2 ResNet blob names -> Detectron2 module names ([D2] c2_model_loading.convert_basic_c2_names + the
    backbone part of convert_c2_detectron_names), e.g.
        conv1_w -> stem.conv1.weight            res_conv1_bn_s -> stem.conv1.norm.weight
        res2_0_branch2a_w -> res2.0.conv1.weight     res2_0_branch1_bn_b -> res2.0.shortcut.norm.bias
    FrozenBN statistics of an affine-only C2 model: running_mean = 0, running_var = 1 (added by the caller where the
    pickle has none). Non-backbone blobs (fc1000, momentum) are dropped."""
    out = {}
    for k, v in blobs.items():
        if k.endswith("_momentum") or k.startswith("fc1000") or k.startswith("pred_"):
            continue
        n = k
        n = re.sub(r"_bn_s$", ".norm.weight", n)
        n = re.sub(r"_bn_b$", ".norm.bias", n)
        n = re.sub(r"_bn_rm$", ".norm.running_mean", n)
        n = re.sub(r"_bn_riv$", ".norm.running_var", n)
        n = re.sub(r"_w$", ".weight", n)
        n = re.sub(r"_b$", ".bias", n)
        n = re.sub(r"^res_conv1", "conv1", n)
        n = re.sub(r"^conv1", "stem.conv1", n)
        m = re.match(r"^res(\d)_(\d+)_branch(1|2a|2b|2c)(.*)$", n)
        if m:
            conv = {"1": "shortcut", "2a": "conv1", "2b": "conv2", "2c": "conv3"}[m.group(3)]
            n = f"res{m.group(1)}.{m.group(2)}.{conv}{m.group(4)}"
        out[n] = v
    return out


def align_and_update_state_dicts(model_sd, ckpt_sd):
    """[D2] align_and_update_state_dicts, the part that matters here: each checkpoint key is assigned to the model key
    it is the longest ``.``-delimited suffix of; returns {model key: tensor} for the matched keys only."""
    matched = {}
    for mk in model_sd:
        best = None
        for ck in ckpt_sd:
            if mk == ck or mk.endswith("." + ck):
                if best is None or len(ck) > len(best):
                    best = ck
        if best is not None:
            matched[mk] = ckpt_sd[best]
    return matched


class DetectionTSCheckpointer:
    def __init__(self, model, save_dir="", *, save_to_disk=True, **checkpointables):
        self.model = model                      # EnsembleTSModel
        self.save_dir = save_dir
        self.save_to_disk = save_to_disk
        self.checkpointables = dict(checkpointables)     # optimizer=..., scheduler=...

    # ------------------------------------------------------------------ fvcore Checkpointer surface
    def save(self, name, **kwargs):
        if not self.save_dir or not self.save_to_disk:
            return None
        data = {"model": {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}}
        for key, obj in self.checkpointables.items():
            data[key] = _to_cpu(obj.state_dict())
        data.update(kwargs)
        os.makedirs(self.save_dir, exist_ok=True)
        path = os.path.join(self.save_dir, f"{name}.pth")
        torch.save(data, path)
        with open(os.path.join(self.save_dir, "last_checkpoint"), "w") as f:
            f.write(os.path.basename(path))
        return path

    def has_checkpoint(self):
        return bool(self.save_dir) and os.path.exists(os.path.join(self.save_dir, "last_checkpoint"))

    def get_checkpoint_file(self):
        with open(os.path.join(self.save_dir, "last_checkpoint")) as f:
            return os.path.join(self.save_dir, f.read().strip())

    def resume_or_load(self, path, *, resume=True):
        if resume and self.has_checkpoint():
            return self.load(self.get_checkpoint_file())
        return self.load(path, checkpointables=[])

    def load(self, path, checkpointables=None):
        if not path:
            return {}
        ckpt = self._load_file(path)
        incompatible = self._load_model(ckpt)
        if incompatible.missing_keys or incompatible.unexpected_keys or incompatible.incorrect_shapes:
            logger.warning("checkpoint %s: %d missing, %d unexpected, %d wrong shape", path, len(incompatible.missing_keys),
                           len(incompatible.unexpected_keys), len(incompatible.incorrect_shapes))
        self.last_incompatible = incompatible
        for key in (self.checkpointables if checkpointables is None else checkpointables):
            if key in ckpt:
                self.checkpointables[key].load_state_dict(ckpt.pop(key))
        return ckpt       # whatever is left: "iteration", ...

    # ------------------------------------------------------------------ [D2] DetectionCheckpointer._load_file
    def _load_file(self, filename):
        if filename.endswith(".pkl"):
            with open(filename, "rb") as f:
                data = pickle.load(f, encoding="latin1")
            if "model" in data and "__author__" in data:
                return data
            if "blobs" in data:         # Detectron (C2) model zoo format
                data = data["blobs"]
            data = {k: v for k, v in data.items() if not k.endswith("_momentum")}
            return {"model": data, "__author__": "Caffe2", "matching_heuristics": True}
        loaded = torch.load(filename, map_location="cpu", weights_only=False)
        if "model" not in loaded:
            loaded = {"model": loaded}
        return loaded

    # ------------------------------------------------------------------ detection_checkpoint.py:11-60
    def _load_model(self, checkpoint):
        if checkpoint.get("__author__", None) == "Caffe2":
            student = self.model.modelStudent
            if checkpoint.get("matching_heuristics", False):
                sd = {k: _as_tensor(v) for k, v in convert_c2_resnet_names(checkpoint["model"]).items()}
                checkpoint["model"] = align_and_update_state_dicts(student.state_dict(), sd)
            inc = self._load_student_model(checkpoint)
            missing = [k for k in inc.missing_keys if k not in ("pixel_mean", "pixel_std")]
            return IncompatibleKeys(missing, inc.unexpected_keys, inc.incorrect_shapes)
        sd = {k[len("module."):] if k.startswith("module.") else k: _as_tensor(v) for k, v in checkpoint.pop("model").items()}
        model_sd = self.model.state_dict()
        wrong = [(k, tuple(v.shape), tuple(model_sd[k].shape)) for k, v in sd.items()
                 if k in model_sd and tuple(v.shape) != tuple(model_sd[k].shape)]
        for k, _, _ in wrong:
            sd.pop(k)
        self.model.load_state_dict(sd, strict=False)
        missing = [k for k in model_sd if k not in sd and k.split(".", 1)[-1] not in ("pixel_mean", "pixel_std")]
        return IncompatibleKeys(missing, [k for k in sd if k not in model_sd], wrong)

    # ------------------------------------------------------------------ detection_checkpoint.py:62-89
    def _load_student_model(self, checkpoint):
        sd = {k[len("module."):] if k.startswith("module.") else k: _as_tensor(v) for k, v in checkpoint.pop("model").items()}
        student = self.model.modelStudent
        model_sd = student.state_dict()
        wrong = []
        for k in list(sd.keys()):
            if k in model_sd and tuple(model_sd[k].shape) != tuple(sd[k].shape):
                wrong.append((k, tuple(sd[k].shape), tuple(model_sd[k].shape)))
                sd.pop(k)
        student.load_state_dict(sd, strict=False)
        return IncompatibleKeys([k for k in model_sd if k not in sd], [k for k in sd if k not in model_sd], wrong)


def _as_tensor(v):
    if isinstance(v, np.ndarray):
        return torch.from_numpy(v.copy())
    return v


def _to_cpu(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj
