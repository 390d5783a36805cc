"""TEST INFRASTRUCTURE (lives beside __graft_entry__.py, outside the product package: it imports the oracle).
__graft_entry__.smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle.

1+1 images at 128x160: EMA copy, teacher forward + two NMS criteria + pseudo-label thresholding, student
forward/backward on the labeled and unlabeled batches, SGD — then the supervised losses are compared with
oracle/ut2_model.py on the same weights and inputs (bf16 activations vs fp32: 5 %)."""
import os
import sys

import torch


def run():
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs cuda:0 (the UT2 B200 path has no CPU fallback)")
    root = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.join(root, "unbiased-teacher-v2_b200")
    for p in (root, pkg):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import ut2_model as M
    from oracle import ut2_oracle as O

    from ubteacher import _C
    from ubteacher.config import add_ubteacher_config
    from ubteacher.d2compat.config import get_cfg
    from ubteacher.d2compat.events import EventStorage
    from ubteacher.data.synthetic import SyntheticTwoCropLoader
    from ubteacher.engine import UBTeacherTrainer

    cfg = get_cfg()
    add_ubteacher_config(cfg)
    cfg.merge_from_file(os.path.join(pkg, "configs/FCOS/coco-standard/fcos_R_50_ut2_sup1_run0.yaml"))
    cfg.merge_from_list(["SEMISUPNET.BURN_UP_STEP", 0, "SOLVER.IMG_PER_BATCH_LABEL", 1, "SOLVER.IMG_PER_BATCH_UNLABEL", 1,
                         "MODEL.DEVICE", "cuda:0", "SEED", 7])
    loader = SyntheticTwoCropLoader(1, 1, h=128, w=160, boxes_per_image=3, pool=1)
    tr = UBTeacherTrainer(cfg, data_loader=loader)
    sd = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    # peek at the batch the trainer is going to consume (the pool is deterministic)
    ref_loader = SyntheticTwoCropLoader(1, 1, h=128, w=160, boxes_per_image=3, pool=1)
    lq, lk, _, _ = next(ref_loader)
    with EventStorage(0) as tr.storage:
        tr.iter = 0
        tr.run_step_full_semisup()
    names, vec = tr.last_losses
    got = dict(zip(names, vec.cpu().tolist()))
    lab = lq + lk
    s = M.forward_dense(sd, [d["image"] for d in lab])
    ref, _ = O.fcos_losses_labeled(s["logits"], s["reg"], s["std"], s["ctr"], s["locations"],
                                   [d["instances"].gt_boxes.tensor for d in lab], [d["instances"].gt_classes for d in lab])
    for k, v in ref.items():
        v = float(v)
        if abs(got[k] - v) > 5e-2 * abs(v) + 2e-3:
            raise AssertionError(f"smoke: {k} = {got[k]} but the oracle says {v}")
    torch.cuda.synchronize()
    print(f"smoke ok: {_C.launch_count} kernel launches; losses " + ", ".join(f"{k}={got[k]:.4f}" for k in ref))
