"""Fused stem + max-pool micro-benchmark (run under gpurun; ncu target)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200")]
from ubteacher import ops
N = int(os.environ.get("B", 16))
imgs = [torch.randint(0, 256, (3, 800, 1333), dtype=torch.uint8, device="cuda") for _ in range(N)]
w = (torch.randn(7, 7, 3, 64, device="cuda") * 0.05)
sc, sh = torch.rand(64, device="cuda") * 0.1, torch.randn(64, device="cuda")
mean, std = [103.53, 116.28, 123.675], [1.0, 1.0, 1.0]
def fused(): return ops.stem_pool_batched(imgs, w, sc, sh, mean, std, 800, 1344)
x1 = torch.empty((N, 400, 672, 64), dtype=torch.bfloat16, device="cuda")
def split():
    ops.stem_conv_batched(imgs, w, sc, sh, mean, std, x1, 400, 672)
    return ops.maxpool3x3s2(x1)
for name, f in (("fused", fused), ("split", split)):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us for {N} images")
