"""Oracle loss evaluated on the DEVICE's dense head outputs (debug aid): isolates the loss kernels on real data."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200"), os.path.join(ROOT, "tests")]
import torch
from oracle import ut2_model as M, ut2_oracle as O
from test_model_gpu import make_batch, diversify, rel
from util_cfg import fcos_cfg
from ubteacher.modeling import OneStageDetector

m = OneStageDetector(fcos_cfg())
diversify(m)
m.train()
batch = make_batch(3, [(160, 224), (128, 192)], 2)
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
losses, pending = m.forward_train(batch, "labeled")
fwd = pending["fwd"]
N, geom = 3, fwd["geom"]
hd = "proposal_generator.fcos_head."
logits, reg, std, ctr, leaves = [], [], [], [], []
scales = [sd[f"{hd}scales.{l}.scale"].clone().requires_grad_(True) for l in range(5)]
for l in range(5):
    h, w = geom.hw[l]
    lo, hi = geom.off[l] * N, geom.off[l + 1] * N
    cls = fwd["cls_out"][lo:hi].view(N, h, w, 80).permute(0, 3, 1, 2).float().cpu().requires_grad_(True)
    box = fwd["box_out"][lo:hi].view(N, h, w, 80).permute(0, 3, 1, 2).float().cpu().requires_grad_(True)
    leaves.append((cls, box))
    logits.append(cls); reg.append(box[:, :68] * scales[l]); std.append(box[:, 68:72]); ctr.append(box[:, 72:73])
locs = [O.compute_locations(h, w, s) for (h, w), s in zip(geom.hw, (8, 16, 32, 64, 128))]
ref, _ = O.fcos_losses_labeled(logits, reg, std, ctr, locs, [b["boxes"] for b in batch], [b["classes"] for b in batch])
print("device", {k: float(v) for k, v in losses.items()})
print("oracle on device outputs", {k: float(v) for k, v in ref.items()})
w = [1.0, 0.7, 1.3, 0.0]
(ref["loss_fcos_cls"] * w[0] + ref["loss_fcos_loc"] * w[1] + ref["loss_fcos_ctr"] * w[2]).backward()
eng = m.engine
eng.arena.grad.zero_()
gout = torch.tensor(w, device="cuda")
dcls, dbox = m.fcos_outputs.backward(fwd, eng.scales, eng.dscales, pending["ctxs"], [gout])
torch.cuda.synchronize()
for l in range(5):
    h, w_ = geom.hw[l]
    lo, hi = geom.off[l] * N, geom.off[l + 1] * N
    dc = dcls[lo:hi].view(N, h, w_, 80).permute(0, 3, 1, 2).float().cpu()
    db = dbox[lo:hi].view(N, h, w_, 80).permute(0, 3, 1, 2).float().cpu()
    gc, gb = leaves[l][0].grad, leaves[l][1].grad
    print("level", l, "dcls rel", rel(dc, gc), "dbox[:68]", rel(db[:, :68], gb[:, :68]), "dstd", rel(db[:, 68:72], gb[:, 68:72]), "dctr", rel(db[:, 72:73], gb[:, 72:73]),
          "dscale dev", float(eng.dscales[l]), "oracle", float(scales[l].grad))
