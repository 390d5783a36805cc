"""``ml_nms`` (ubteacher/layers/ml_nms.py:8-31): multi-class NMS on an ``Instances`` ([D2] ``batched_nms``: boxes of different
classes never suppress each other), optional top ``max_proposals``; returns ``boxlist[keep]`` with ``keep`` sorted by score.
Device operator ``ut2_nms_batched`` (the keep list equals torchvision's, bit for bit: tests/test_layers_gpu.py)."""
import torch

from .. import ops_rcnn as R


def ml_nms(boxlist, nms_thresh, max_proposals=-1, score_field="scores", label_field="labels"):
    if nms_thresh <= 0:
        return boxlist
    boxes = boxlist.pred_boxes.tensor
    if not boxes.is_cuda:
        raise RuntimeError("ml_nms runs on the device (no CPU fallback)")
    n = boxes.shape[0]
    if n == 0:
        return boxlist
    scores = boxlist.scores
    labels = boxlist.pred_classes
    cap = n if max_proposals <= 0 else min(n, int(max_proposals))
    cnt = torch.full((1,), n, dtype=torch.int32, device=boxes.device)
    keep, kcnt = R.nms_batched(boxes.detach().float().reshape(1, n, 4).contiguous(), scores.detach().float().reshape(1, n).contiguous(),
                               labels.to(torch.int32).reshape(1, n).contiguous(), cnt, float(nms_thresh), cap)
    k = int(kcnt[0])
    return boxlist[keep[0, :k].long()]
