"""In-step CUDA-event breakdown of one UT2 training step by C-ABI entry point and conv shape (run under gpurun)."""
import os, sys, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200")]
import bench
from ubteacher import _C
from ubteacher.d2compat.events import EventStorage
from ubteacher.data.synthetic import SyntheticTwoCropLoader
from ubteacher.engine import UBRCNNTeacherTrainer, UBTeacherTrainer
B = int(os.environ.get("B", 8))
ARCH = os.environ.get("ARCH", "fcos")
cfg = bench.build_cfg(B, B, arch=ARCH)
tr = (UBTeacherTrainer if ARCH == "fcos" else UBRCNNTeacherTrainer)(cfg, data_loader=SyntheticTwoCropLoader(B, B, device=torch.device("cuda")))
tr.storage = EventStorage(0); tr.metrics_period = 10**9; tr.iter = -1
for _ in range(3):
    tr.iter += 1; tr.run_step_full_semisup()
torch.cuda.synchronize()
_C.EVENT_PROFILE = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.perf_counter(); e0.record()
STEPS = 3
for _ in range(STEPS):
    tr.iter += 1; tr.run_step_full_semisup()
e1.record(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
prof, _C.EVENT_PROFILE = _C.EVENT_PROFILE, None
tot = {k: sum(a.elapsed_time(b) for a, b in v) / STEPS for k, v in prof.items()}
T = sum(tot.values())
print(f"step {e0.elapsed_time(e1)/STEPS:.2f} ms (host wall {wall/STEPS*1e3:.2f} ms), sum of kernel brackets {T:.2f} ms")
grp = collections.defaultdict(float)
for k, v in tot.items():
    grp[k.split(" ")[0]] += v
for k, v in sorted(grp.items(), key=lambda x: -x[1]):
    print(f"  {v:8.3f} ms {100*v/T:5.1f}%  {k}")
print("top conv shapes:")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:45]:
    if k.startswith("conv"):
        n = len(prof[k]) / STEPS
        kind, shp, flt = k.split(" ")
        dims, cout = shp.split("->"); R = int(flt[0]); st = int(flt.split("/")[1])
        if "L" in dims:      # level-major launch: N x L<locations> x C
            N_, L_, C = dims.split("x"); N_, C, PQ = int(N_), int(C), int(L_[1:])
        else:
            N_, H, W, C = map(int, dims.split("x"))
            PQ = ((H + 2*(R//2) - R)//st + 1) * ((W + 2*(R//2) - R)//st + 1)
        fl = 2.0 * N_ * PQ * int(cout) * C * R * R * n
        print(f"  {v:8.3f} ms n={n:5.1f} {fl/v/1e9:7.1f} TF/s  {k}")
