"""``EnsembleTSModel`` (reference: ubteacher/modeling/meta_arch/ts_ensemble.py:6-16): the container whose
attribute names give checkpoints their ``modelTeacher.`` / ``modelStudent.`` key prefixes."""
from torch import nn
from torch.nn.parallel import DataParallel, DistributedDataParallel

_PREFIXES = ("modelTeacher.", "modelStudent.")


def _unwrap(model):
    """The replicas are registered without a (Distributed)DataParallel wrapper, so the key prefixes never carry `module.`."""
    return model.module if isinstance(model, (DistributedDataParallel, DataParallel)) else model


class EnsembleTSModel(nn.Module):
    def __init__(self, modelTeacher, modelStudent):
        super().__init__()
        self.modelTeacher, self.modelStudent = _unwrap(modelTeacher), _unwrap(modelStudent)

    def _members(self):
        return zip(_PREFIXES, (self.modelTeacher, self.modelStudent))

    def state_dict(self, *args, **kwargs):
        out = {}
        for prefix, member in self._members():
            out.update(member.state_dict(prefix=prefix))
        return out

    def load_state_dict(self, sd, strict=True):
        for prefix, member in self._members():
            part = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
            if part:
                member.load_state_dict(part, strict)
