"""In-step CUDA-event breakdown of one UT2 training step by C-ABI entry point and conv shape (run under gpurun)."""
import os, sys, collections, torch
os.environ.setdefault("UT2_CONCURRENT_TEACHER", "0")   # per-launch events need the launches to own the device
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
if os.environ.get("REGIME") == "saturated":
    bench.saturate(tr, ARCH)
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
print("conv shapes (ms/step, launches/step, TF/s, min-traffic GB/s [in + out + weights, no residual], loss vs max(1450 TF/s, 6.0 TB/s) ideal):")
rows = []
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
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
        by = 2.0 * n * (N_ * PQ * (C * (1 if st == 1 else st * st if R > 1 else 1) + int(cout)) + int(cout) * C * R * R * (2 if kind == "conv_wgrad" else 1))
        ideal = max(fl / 1450e9, by / 6.0e9)
        rows.append((v - ideal, v, n, fl / v / 1e9, by / v / 1e6, ideal, k))
for loss, v, n, tf, gbs, ideal, k in sorted(rows, key=lambda r: -r[0])[:int(os.environ.get("TOP", 60))]:
    print(f"  {v:8.3f} ms n={n:5.1f} {tf:7.1f} TF/s {gbs:7.0f} GB/s  ideal {ideal:6.3f} loss {loss:6.3f}  {k}")
print(f"  sum over conv shapes: {sum(r[1] for r in rows):.2f} ms, ideal {sum(r[5] for r in rows):.2f} ms")
