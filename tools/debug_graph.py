import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200"), os.path.join(ROOT, "tests")]
from util_cfg import fcos_cfg
from test_model_gpu import diversify
from ubteacher.d2compat.events import EventStorage
from ubteacher.data.synthetic import SyntheticTwoCropLoader
from ubteacher.engine import UBTeacherTrainer
def run(graph):
    tr = UBTeacherTrainer(fcos_cfg(), data_loader=SyntheticTwoCropLoader(1, 2, h=128, w=160, boxes_per_image=3, pool=2))
    diversify(tr.model); tr.enable_cuda_graph(graph)
    out = []
    with EventStorage(0) as tr.storage:
        for it in range(5):
            tr.iter = it; tr.run_step_full_semisup(); out.append(tr.last_losses[1].cpu().clone()); tr.scheduler.step()
    return out
a, b, c = run(False), run(False), run(True)
for i in range(5):
    print(i, "eager ", [f"{v:.5f}" for v in a[i].tolist()])
    print(i, "eager2", [f"{v:.5f}" for v in b[i].tolist()])
    print(i, "graph ", [f"{v:.5f}" for v in c[i].tolist()])
