import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unbiased-teacher-v2_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_rcnn_model_gpu as T
from util_cfg import rcnn_cfg
from ubteacher.engine import UBRCNNTeacherTrainer

class Loader:
    def __init__(self): self.i = 0
    def __iter__(self): return self
    def __next__(self):
        self.i += 1
        mk = lambda n, seed: T.make_batch(n, [(128, 160), (160, 192)], seed, nbox=4)
        lq, uq = mk(1, 100 + self.i), mk(2, 200 + self.i)
        lk = [dict(d, image=torch.flip(d["image"], [0])) for d in lq]
        uk = [dict(d) for d in uq]
        return lq, lk, uq, uk

tr = UBRCNNTeacherTrainer(rcnn_cfg(), data_loader=Loader())
T.diversify(tr.model)
eng = tr.model.engine
print("fpn_out2 std", float(eng.arena.views["backbone.fpn_output2.weight"].std()))
data = next(iter(Loader()))
lq, lk, uq, uk = data
fwd = eng.forward_features([d["image"].cuda() for d in lq + lk], train=False)
for l, f in enumerate(fwd["levels"]): print("student level", l, float(f.float().std()), bool(torch.isfinite(f.float()).all()))
print("rpn_out std", float(fwd["rpn_out"].float().std()))
tr.iter = 0
tr.run_step_full_semisup()
print(tr.last_losses)
fwd = eng.forward_features([d["image"].cuda() for d in lq + lk], train=False)
for l, f in enumerate(fwd["levels"]): print("after step: student level", l, float(f.float().std()), bool(torch.isfinite(f.float()).all()))
print("lr", tr.optimizer.param_groups[0]["lr"], "grad absmax", float(eng.arena.grad.abs().max()))
