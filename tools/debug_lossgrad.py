import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200"), os.path.join(ROOT, "tests")]
import test_kernels_gpu as T
from oracle import ut2_oracle as O
g = T.load("fcos_losses_labeled.pt")
for w in ([1.0,0,0,0],[0,1.0,0,0],[0,0,1.0,0]):
    losses, dcls, dbox, _ = T._run_loss(g, 0, {"boxes": g["boxes"], "classes": g["classes"]}, torch.tensor(w))
    leaf = lambda ts: [T.rb(t).requires_grad_(True) for t in ts]
    L, R, S, C = leaf(g["logits"]), leaf(g["reg"]), leaf(g["std"]), leaf(g["ctr"])
    ref, _ = O.fcos_losses_labeled(L, R, S, C, T.locs(), g["boxes"], g["classes"])
    (ref["loss_fcos_cls"] * w[0] + ref["loss_fcos_loc"] * w[1] + ref["loss_fcos_ctr"] * w[2]).backward()
    gr = T._unpack_grads(dcls, dbox, 3)
    for name, leaves in (("logits", L), ("reg", R), ("std", S), ("ctr", C)):
        a = torch.cat([x.flatten() for x in gr[name]]).double()
        b = torch.cat([(x.grad if x.grad is not None else torch.zeros_like(x)).flatten() for x in leaves]).double()
        if b.norm() == 0: continue
        print(w, name, "ratio", float(a.norm()/b.norm()), "cos", float((a*b).sum()/(a.norm()*b.norm())), "maxabs", float(b.abs().max()))
print(losses, {k: float(v) for k,v in ref.items()})
