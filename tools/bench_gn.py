"""GroupNorm micro-benchmark on the full-size level-major pyramid (run under gpurun): UT2_GN_COOP=0|1 python tools/bench_gn.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200")]
from ubteacher import ops
hw = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
geom = ops.LevelGeom(hw, [8, 16, 32, 64, 128])
for N in (2, 8, 16):
    x = torch.randn(geom.L * N, 256, device="cuda").bfloat16()
    dy = torch.randn(geom.L * N, 256, device="cuda").bfloat16()
    gam, bet = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
    dg, db, dbias = [torch.zeros(256, device="cuda") for _ in range(3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def run(f, n=20):
        f(); torch.cuda.synchronize()
        t = 0.0
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            t += e0.elapsed_time(e1)
        return t / n
    y, stats = ops.groupnorm_relu_levels_fwd(x, geom, N, gam, bet)
    tf = run(lambda: ops.groupnorm_relu_levels_fwd(x, geom, N, gam, bet))
    tb = run(lambda: ops.groupnorm_relu_levels_bwd(dy, x, geom, N, stats, gam, bet, dg, db, dbias_prev=dbias))
    el = geom.L * N * 256
    print(f"N={N:2d} coop={os.environ.get('UT2_GN_COOP', '1')}  fwd {tf*1e3:7.1f} us = {4*el/tf/1e6:6.0f} GB/s at 4 B/elem   bwd {tb*1e3:7.1f} us = {6*el/tb/1e6:6.0f} GB/s at 6 B/elem")
