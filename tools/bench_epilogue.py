"""Micro-benchmark of the memory-bound 1x1 convolutions with their epilogue variants (run under gpurun):
none | res (fwd conv3) | mask | res+mask (conv1.dgrad with the fused block-output ReLU backward) vs a separate relu_bwd pass."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200")]
from ubteacher import ops
B = int(os.environ.get("B", 16))
SHAPES = [("res3 128->512", 100, 168, 128, 512), ("res4 256->1024", 50, 84, 256, 1024), ("res5 512->2048", 25, 42, 512, 2048)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=10):
    for _ in range(2): fn()
    t = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
    return t / iters * 1e3

for name, H, W, Cin, Cout in SHAPES:
    x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
    w = torch.randn(Cout, 1, 1, Cin, device="cuda").bfloat16()
    r = torch.randn(B, H, W, Cout, device="cuda").bfloat16()
    m = torch.randn(B, H, W, Cout, device="cuda").relu().bfloat16()
    y = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    out_mb = y.numel() * 2 / 1e6
    t0 = timeit(lambda: ops.conv2d(x, w, Cout, 1, 1, 1, 0, out=y))
    t1 = timeit(lambda: ops.conv2d(x, w, Cout, 1, 1, 1, 0, residual=r, relu=True, out=y))
    t2 = timeit(lambda: ops.conv2d(x, w, Cout, 1, 1, 1, 0, relu_mask=m, out=y))
    t3 = timeit(lambda: ops.conv2d(x, w, Cout, 1, 1, 1, 0, residual=r, relu_mask=m, out=y))
    t4 = timeit(lambda: ops.relu_bwd(r, m))
    print(f"{name:16s} out {out_mb:6.1f} MB | none {t0:6.1f} us ({out_mb/t0*1e3:5.0f} GB/s out) | res {t1:6.1f} ({2*out_mb/t1*1e3:5.0f}) | mask {t2:6.1f} | res+mask {t3:6.1f} ({3*out_mb/t3*1e3:5.0f}) | res + separate relu_bwd {t1+t4:6.1f} (relu_bwd {t4:5.1f})")
