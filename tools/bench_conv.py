"""Micro-benchmark of the tensor-core conv kernels over the FCOS R50-FPN layer shapes (run under gpurun)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200")]
from ubteacher import ops

B = int(os.environ.get("B", 16))
# (name, H, W, Cin, Cout, k, stride, count per image-forward)
SHAPES = [
    ("head3x3 p3", 100, 168, 256, 256, 3, 1, 9), ("head3x3 p4", 50, 84, 256, 256, 3, 1, 9), ("head3x3 p5", 25, 42, 256, 256, 3, 1, 9),
    ("head3x3 p6", 13, 21, 256, 256, 3, 1, 8), ("head3x3 p7", 7, 11, 256, 256, 3, 1, 8),
    ("pred80 p3", 100, 168, 256, 80, 3, 1, 2), ("pred80 p4", 50, 84, 256, 80, 3, 1, 2),
    ("res2 1x1 64-64", 200, 336, 64, 64, 1, 1, 1), ("res2 3x3 64", 200, 336, 64, 64, 3, 1, 3), ("res2 1x1 64-256", 200, 336, 64, 256, 1, 1, 4),
    ("res2 1x1 256-64", 200, 336, 256, 64, 1, 1, 2),
    ("res3 1x1s2 256-128", 200, 336, 256, 128, 1, 2, 1), ("res3 3x3 128", 100, 168, 128, 128, 3, 1, 4), ("res3 1x1 128-512", 100, 168, 128, 512, 1, 1, 4),
    ("res3 1x1s2 256-512", 200, 336, 256, 512, 1, 2, 1), ("res3 1x1 512-128", 100, 168, 512, 128, 1, 1, 3),
    ("res4 1x1s2 512-256", 100, 168, 512, 256, 1, 2, 1), ("res4 3x3 256", 50, 84, 256, 256, 3, 1, 6), ("res4 1x1 256-1024", 50, 84, 256, 1024, 1, 1, 6),
    ("res4 1x1s2 512-1024", 100, 168, 512, 1024, 1, 2, 1), ("res4 1x1 1024-256", 50, 84, 1024, 256, 1, 1, 5),
    ("res5 1x1s2 1024-512", 50, 84, 1024, 512, 1, 2, 1), ("res5 3x3 512", 25, 42, 512, 512, 3, 1, 3), ("res5 1x1 512-2048", 25, 42, 512, 2048, 1, 1, 3),
    ("res5 1x1s2 1024-2048", 50, 84, 1024, 2048, 1, 2, 1), ("res5 1x1 2048-512", 25, 42, 2048, 512, 1, 1, 2),
    ("fpn lat3", 100, 168, 512, 256, 1, 1, 1), ("fpn lat4", 50, 84, 1024, 256, 1, 1, 1), ("fpn lat5", 25, 42, 2048, 256, 1, 1, 1),
]

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

tot_f = tot_w = tot_fl = 0.0
print(f"batch {B}")
for name, H, W, Cin, Cout, k, s, cnt in SHAPES:
    pad = k // 2
    x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
    w = torch.randn(Cout, k, k, Cin, device="cuda").bfloat16()
    P, Q = ops.conv_out_hw(H, W, k, k, s, pad)
    y = torch.empty(B, P, Q, Cout, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(B, P, Q, Cout, device="cuda").bfloat16()
    dw = torch.zeros(Cout, k, k, Cin, device="cuda")
    fl = 2.0 * B * P * Q * Cout * Cin * k * k
    tf = timeit(lambda: ops.conv2d(x, w, Cout, k, k, s, pad, out=y))
    tw = timeit(lambda: ops.conv2d_wgrad(x, dy, Cout, k, k, s, pad, dw)) if Cin % 64 == 0 else float("nan")
    byts = (x.numel() + y.numel()) * 2
    print(f"{name:22s} M={B*P*Q:7d} N={Cout:4d} K={Cin*k*k:5d}  fwd {tf*1e3:8.1f} us {fl/tf/1e9:7.1f} TF/s ({byts/tf/1e6:6.0f} GB/s)   wgrad {tw*1e3:8.1f} us {fl/tw/1e9:7.1f} TF/s   x{cnt}")
    tot_f += tf * cnt; tot_w += tw * cnt; tot_fl += fl * cnt
print(f"weighted per-forward: fwd {tot_f:.2f} ms ({tot_fl/tot_f/1e9:.0f} TF/s)  wgrad {tot_w:.2f} ms ({tot_fl/tot_w/1e9:.0f} TF/s)")
