"""``EnsembleTSModel`` (reference: ubteacher/modeling/meta_arch/ts_ensemble.py:6-16): the container whose
attribute names give checkpoints their ``modelTeacher.`` / ``modelStudent.`` key prefixes."""
from torch import nn
from torch.nn.parallel import DataParallel, DistributedDataParallel


class EnsembleTSModel(nn.Module):
    def __init__(self, modelTeacher, modelStudent):
        super().__init__()
        if isinstance(modelTeacher, (DistributedDataParallel, DataParallel)):
            modelTeacher = modelTeacher.module
        if isinstance(modelStudent, (DistributedDataParallel, DataParallel)):
            modelStudent = modelStudent.module
        self.modelTeacher = modelTeacher
        self.modelStudent = modelStudent

    def state_dict(self, *args, **kwargs):
        sd = {}
        sd.update(self.modelTeacher.state_dict(prefix="modelTeacher."))
        sd.update(self.modelStudent.state_dict(prefix="modelStudent."))
        return sd

    def load_state_dict(self, sd, strict=True):
        for pre, m in (("modelTeacher.", self.modelTeacher), ("modelStudent.", self.modelStudent)):
            sub = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
            if sub:
                m.load_state_dict(sub, strict)
