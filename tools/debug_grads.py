import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200"), os.path.join(ROOT, "tests")]
from util_cfg import fcos_cfg
from test_model_gpu import make_batch, diversify
from ubteacher.modeling import OneStageDetector
from oracle import ut2_model as M, ut2_oracle as O
m = OneStageDetector(fcos_cfg()); diversify(m); m.train()
batch = make_batch(3, [(160, 224), (128, 192)], 2)
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
m.engine.arena.grad.zero_()
losses, pending = m.forward_train(batch, "labeled")
w = [1.0, 0.7, 1.3, 0.0]
m.backward_pending(pending, [w])
tk = M.trainable_keys(sd)
params = {k: sd[k].clone().requires_grad_(True) for k in tk}
sdp = dict(sd); sdp.update(params)
s = M.forward_dense(sdp, [b["image"] for b in batch])
ref, _ = O.fcos_losses_labeled(s["logits"], s["reg"], s["std"], s["ctr"], s["locations"], [b["boxes"] for b in batch], [b["classes"] for b in batch])
(ref["loss_fcos_cls"] * w[0] + ref["loss_fcos_loc"] * w[1] + ref["loss_fcos_ctr"] * w[2]).backward()
G = m.engine.arena.gviews
for k in tk:
    a, b = G[k].float().cpu().double().flatten(), params[k].grad.double().flatten()
    cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30)); ratio = float(a.norm() / (b.norm() + 1e-30))
    flag = "" if (cos > 0.97 and 0.9 < ratio < 1.1) else "   <<<<"
    print(f"{k:70s} cos={cos:.4f} ratio={ratio:.4f} |g|={float(b.norm()):.3e}{flag}")
