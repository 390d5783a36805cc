"""Layer-by-layer comparison of the FCOS engine with the bf16-rounding-point oracle (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200"), os.path.join(ROOT, "tests")]
import torch
import torch.nn.functional as F
from oracle import ut2_model as M, ut2_oracle as O
from test_model_gpu import make_batch, diversify, rel
from util_cfg import fcos_cfg
from ubteacher.modeling import OneStageDetector

m = OneStageDetector(fcos_cfg())
diversify(m)
m.train()
batch = make_batch(3, [(160, 224), (128, 192)], 2)
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
eng = m.engine
fwd = eng.forward([b["image"].cuda() for b in batch], train=True)
tape = fwd["tape"]
N, geom = 3, fwd["geom"]
nchw = lambda t: t.float().cpu().permute(0, 3, 1, 2)
for mode in ("fp32", "bf16"):
    ctx = M.bf16_points() if mode == "bf16" else None
    if ctx: ctx.__enter__()
    x, _ = M.preprocess(sd, [b["image"] for b in batch])
    feats = M.trunk(sd, x)
    print(mode)
    for st in ("res3", "res4", "res5"):
        y = tape[st][-1][3]
        print("  ", st, rel(nchw(y), feats[st]))
    o = M.fpn_topdown(sd, feats, (5, 4, 3))
    c3, c4, c5, lat3, lat4, lat5, p5, p6, p6r = tape["fpn"]
    print("   p5", rel(nchw(p5), o[5]))
    fe = M.backbone(sd, x)
    feat = tape["head"]["feat"]
    for l in range(5):
        h, w = geom.hw[l]
        d = feat[geom.off[l] * N: geom.off[l + 1] * N].view(N, h, w, 256)
        print("   p%d" % (l + 3), rel(nchw(d), fe[l]))
    hd = "proposal_generator.fcos_head."
    for t in ("cls_tower", "bbox_tower"):
        saved, top = tape["head"][t]
        xs = list(fe)
        for i in range(4):
            xin, c, stats = saved[i]
            cs, ys = [], []
            for l in range(5):
                cc = M.q(M.conv(xs[l], sd[f"{hd}{t}.{3 * i}.weight"], sd[f"{hd}{t}.{3 * i}.bias"], 1, 1))
                yy = M.q(F.relu(F.group_norm(cc, 32, sd[f"{hd}{t}.{3 * i + 1}.weight"], sd[f"{hd}{t}.{3 * i + 1}.bias"], 1e-5)))
                cs.append(cc); ys.append(yy)
            e = []
            for l in range(5):
                h, w = geom.hw[l]
                d = c[geom.off[l] * N: geom.off[l + 1] * N].view(N, h, w, 256)
                e.append(rel(nchw(d), cs[l]))
            print("  ", t, i, "conv out per level", ["%.4f" % v for v in e])
            xs = ys
        e = []
        for l in range(5):
            h, w = geom.hw[l]
            d = top[geom.off[l] * N: geom.off[l + 1] * N].view(N, h, w, 256)
            e.append(rel(nchw(d), xs[l]))
        print("  ", t, "top per level", ["%.4f" % v for v in e])
    s = M.forward_dense(sd, [b["image"] for b in batch])
    ref, _ = O.fcos_losses_labeled(s["logits"], s["reg"], s["std"], s["ctr"], s["locations"], [b["boxes"] for b in batch], [b["classes"] for b in batch])
    bias = sd[hd + "cls_logits.bias"].view(1, -1, 1, 1)
    for l in range(5):
        h, w = geom.hw[l]
        lo, hi = geom.off[l] * N, geom.off[l + 1] * N
        cls = fwd["cls_out"][lo:hi].view(N, h, w, 80).permute(0, 3, 1, 2).float().cpu()
        box = fwd["box_out"][lo:hi].view(N, h, w, 80).permute(0, 3, 1, 2).float().cpu()
        sc = float(sd[f"{hd}scales.{l}.scale"])
        print("   level", l, "logits", rel(cls - bias, s["logits"][l] - bias), "reg", rel(box[:, :68] * sc, s["reg"][l]),
              "std", rel(box[:, 68:72], s["std"][l]), "ctr", rel(box[:, 72:73], s["ctr"][l]))
    print("   oracle losses", {k: float(v) for k, v in ref.items()})
    if ctx: ctx.__exit__()
losses, pending = m.forward_train(batch, "labeled")
print("device losses", {k: float(v) for k, v in losses.items()})
