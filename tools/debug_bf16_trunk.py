"""Block-by-block comparison of the trunk with the bf16-rounding-point oracle (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200"), os.path.join(ROOT, "tests")]
import torch
import torch.nn.functional as F
from oracle import ut2_model as M
from test_model_gpu import make_batch, diversify, rel
from util_cfg import fcos_cfg
from ubteacher.modeling import OneStageDetector
from ubteacher import ops

m = OneStageDetector(fcos_cfg())
diversify(m)
batch = make_batch(3, [(160, 224), (128, 192)], 2)
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
eng = m.engine
images = [b["image"].cuda() for b in batch]
nchw = lambda t: t.float().cpu().permute(0, 3, 1, 2)
N = 3
Hp, Wp = 160, 224
P, Q = Hp // 2, Wp // 2
x = torch.empty((N, P, Q, 64), dtype=torch.bfloat16, device="cuda")
ops.stem_conv_batched(images, eng.stem_w, eng.stem.scale, eng.stem.shift, eng.pixel_mean, eng.pixel_std, x, P, Q)
xp = ops.maxpool3x3s2(x)
with M.bf16_points():
    xo, _ = M.preprocess(sd, [b["image"] for b in batch])
    bu = "backbone.bottom_up."
    so = M.q(F.relu(M.stem(xo, sd, bu + "stem.conv1")))
    print("stem", rel(nchw(x), so), "max abs", float((nchw(x) - so).abs().max()), "ref max", float(so.abs().max()))
    d = (nchw(x) - so).abs()
    print("  frac differing", float((d > 0).float().mean()))
    po = F.max_pool2d(so, 3, 2, 1)
    print("pool", rel(nchw(xp), po))
    xd, xr = xp, po
    for stage, blks in eng.blocks:
        for i, b in enumerate(blks):
            p = f"{bu}{stage}.{i}."
            s = b["stride"]
            # device, op by op
            a_d = b["conv1"].fwd(xd, relu=True)
            m_d = b["conv2"].fwd(a_d, relu=True)
            sc_d = b["shortcut"].fwd(xd) if "shortcut" in b else xd
            y_d = b["conv3"].fwd(m_d, residual=sc_d, relu=True)
            # oracle on the DEVICE's input of this block (isolates the block)
            xin = nchw(xd)
            sc = M.q(M.conv_bn(xin, sd, p + "shortcut", s)) if (p + "shortcut.weight") in sd else xin
            a = M.q(F.relu(M.conv_bn(xin, sd, p + "conv1", s)))
            mm = M.q(F.relu(M.conv_bn(nchw(a_d), sd, p + "conv2", 1, 1)))
            out = M.conv_bn(nchw(m_d), sd, p + "conv3")
            y = M.q(F.relu(out + nchw(sc_d)))
            print(stage, i, "conv1 %.5f conv2 %.5f shortcut %.5f conv3+res %.5f" % (rel(nchw(a_d), a), rel(nchw(m_d), mm), rel(nchw(sc_d), sc), rel(nchw(y_d), y)))
            xd = y_d
