import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unbiased-teacher-v2_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ut2_rcnn_oracle as OR
from ubteacher import ops_rcnn as R
import test_rcnn_kernels_gpu as T
g = T.load("rcnn_rpn_losses.pt")
gm = T.rgeom()
N = len(g["gt_boxes"])
gb, _, gs, _, cnt = T.pack_gt(g["gt_boxes"], scores=g["gt_scores"])
anchors = torch.cat(OR.generate_anchors(T.LEVEL_HW))
pre, _ = R.rpn_label_anchors(gm, N, gb, cnt, keys=None, seed=1, batch=4 * gm.A, pos_frac=0.5)
keys = torch.stack(g["keys"]).to(torch.int64).cuda().to(torch.int32)
lab, matched = R.rpn_label_anchors(gm, N, gb, cnt, keys=keys)
torch.cuda.synchronize()
for i in range(N):
    iou = OR.pairwise_iou(g["gt_boxes"][i], anchors)
    midx, ref = OR.matcher(iou, (0.3, 0.7), (0, -1, 1), True)
    p = pre[i].cpu()
    print("img", i, "pre mismatches", int((p != ref).sum()), "ref counts", [(int((ref == v).sum())) for v in (-1, 0, 1)],
          "got", [(int((p == v).sum())) for v in (-1, 0, 1)])
    bad = (p != ref).nonzero().squeeze(1)[:10]
    print("   bad idx", bad.tolist(), "ref", ref[bad].tolist(), "got", p[bad].tolist())
    l = lab[i].cpu(); r = g["labels"][i]
    print("   sampled mismatches", int((l != r).sum()), "got counts", [(int((l == v).sum())) for v in (-1, 0, 1)],
          "ref", [(int((r == v).sum())) for v in (-1, 0, 1)])
    if len(g["gt_boxes"][i]):
        print("   matched mism", int((matched[i].cpu().long() != midx).sum()))
