import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "unbiased-teacher-v2_b200")


def fcos_cfg(**overrides):
    from ubteacher.config import add_ubteacher_config
    from ubteacher.d2compat.config import get_cfg

    cfg = get_cfg()
    add_ubteacher_config(cfg)
    cfg.merge_from_file(os.path.join(PKG, "configs/FCOS/coco-standard/fcos_R_50_ut2_sup1_run0.yaml"))
    opts = ["SEMISUPNET.BURN_UP_STEP", 0, "SEED", 7]
    for k, v in overrides.items():
        opts += [k, v]
    cfg.merge_from_list(opts)
    return cfg


def oracle_step_cfg(cfg, lr):
    s = cfg.SEMISUPNET
    return {"UNSUP_LOSS_WEIGHT": s.UNSUP_LOSS_WEIGHT, "UNSUP_REG_LOSS_WEIGHT": s.UNSUP_REG_LOSS_WEIGHT,
            "EMA_KEEP_RATE": s.EMA_KEEP_RATE, "NMS_CRITERIA_TRAIN": cfg.MODEL.FCOS.NMS_CRITERIA_TRAIN,
            "NMS_CRITERIA_REG_TRAIN": cfg.MODEL.FCOS.NMS_CRITERIA_REG_TRAIN, "BBOX_THRESHOLD": s.BBOX_THRESHOLD,
            "BBOX_THRESHOLD_REG": s.BBOX_THRESHOLD_REG, "WEIGHT_DECAY": cfg.SOLVER.WEIGHT_DECAY,
            "WEIGHT_DECAY_NORM": cfg.SOLVER.WEIGHT_DECAY_NORM, "MOMENTUM": cfg.SOLVER.MOMENTUM, "LR": lr}


def rcnn_cfg(**overrides):
    from ubteacher.config import add_ubteacher_config
    from ubteacher.d2compat.config import get_cfg

    cfg = get_cfg()
    add_ubteacher_config(cfg)
    cfg.merge_from_file(os.path.join(PKG, "configs/Faster-RCNN/coco-standard/faster_rcnn_R_50_FPN_ut2_sup1_run0.yaml"))
    opts = ["SEMISUPNET.BURN_UP_STEP", 0, "SEED", 7]
    for k, v in overrides.items():
        opts += [k, v]
    cfg.merge_from_list(opts)
    return cfg
