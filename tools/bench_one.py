"""Run one conv configuration a few times (ncu target). usage: bench_one.py B H W Cin Cout k stride [res|mask|none] [fwd|wgrad]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200")]
from ubteacher import ops
B, H, W, Cin, Cout, k, s = map(int, sys.argv[1:8])
aux = sys.argv[8] if len(sys.argv) > 8 else "none"
kind = sys.argv[9] if len(sys.argv) > 9 else "fwd"
pad = k // 2
x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
w = torch.randn(Cout, k, k, Cin, device="cuda").bfloat16()
P, Q = ops.conv_out_hw(H, W, k, k, s, pad)
y = torch.empty(B, P, Q, Cout, device="cuda", dtype=torch.bfloat16)
r = torch.randn(B, P, Q, Cout, device="cuda").bfloat16()
r2 = torch.randn(B, P, Q, Cout, device="cuda").bfloat16()
sc = (torch.rand(Cout, device="cuda") + 0.5) if os.environ.get("SCALE") else None   # product path: FrozenBN scale folded into the weights
sh = torch.randn(Cout, device="cuda")
dw = torch.zeros(Cout, k, k, Cin, device="cuda")
def run():
    if kind == "wgrad":
        ops.conv2d_wgrad(x, r, Cout, k, k, s, pad, dw)
    else:
        ops.conv2d(x, w, Cout, k, k, s, pad, sc, sh, r if aux in ("res", "resmask") else None, aux != "resmask", y, False, r2 if aux in ("mask", "resmask") else None)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
fl = 2.0 * B * P * Q * Cout * Cin * k * k
byts = (x.numel() + y.numel() * (1 + {"none": 0, "res": 1, "mask": 1, "resmask": 2}[aux])) * 2
print(f"{kind} {aux}: {t*1e3:.1f} us  {fl/t/1e9:.1f} TF/s  {byts/t/1e6:.0f} GB/s")
