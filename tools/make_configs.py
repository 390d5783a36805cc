"""Writes the YAML recipes under unbiased-teacher-v2_b200/configs/ from the tables below.

The key/value content matches the reference recipes (reference configs/FCOS/*, configs/Faster-RCNN/*;
effective values listed in SURVEY.md A.2) so that `train_net.py --config-file ...` resolves to the
same configuration; the files themselves are generated, not copied.
"""
import os

import yaml

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "unbiased-teacher-v2_b200", "configs")

BASE_FCOS = {
    "VERSION": 2,
    "MODEL": {
        "META_ARCHITECTURE": "OneStageDetector",
        "BACKBONE": {"NAME": "build_fcos_resnet_fpn_backbone"},
        "RESNETS": {"OUT_FEATURES": ["res3", "res4", "res5"]},
        "FPN": {"IN_FEATURES": ["res3", "res4", "res5"]},
        "PROPOSAL_GENERATOR": {"NAME": "FCOS"},
    },
    "DATASETS": {"TRAIN": "('coco_2017_train',)", "TEST": "('coco_2017_val',)"},
    "SOLVER": {"IMS_PER_BATCH": 16, "BASE_LR": 0.01, "STEPS": "(60000, 80000)", "MAX_ITER": 90000},
    "INPUT": {"MIN_SIZE_TRAIN": "(640, 672, 704, 736, 768, 800)"},
}

BASE_RCNN = {
    "VERSION": 2,
    "MODEL": {
        "META_ARCHITECTURE": "GeneralizedRCNN",
        "BACKBONE": {"NAME": "build_resnet_fpn_backbone"},
        "RESNETS": {"OUT_FEATURES": ["res2", "res3", "res4", "res5"]},
        "FPN": {"IN_FEATURES": ["res2", "res3", "res4", "res5"]},
        "ANCHOR_GENERATOR": {"SIZES": [[32], [64], [128], [256], [512]], "ASPECT_RATIOS": [[0.5, 1.0, 2.0]]},
        "RPN": {"IN_FEATURES": ["p2", "p3", "p4", "p5", "p6"], "PRE_NMS_TOPK_TRAIN": 2000,
                "PRE_NMS_TOPK_TEST": 1000, "POST_NMS_TOPK_TRAIN": 1000, "POST_NMS_TOPK_TEST": 1000},
        "ROI_HEADS": {"NAME": "StandardROIHeads", "IN_FEATURES": ["p2", "p3", "p4", "p5"]},
        "ROI_BOX_HEAD": {"NAME": "FastRCNNConvFCHead", "NUM_FC": 2, "POOLER_RESOLUTION": 7},
    },
    "DATASETS": {"TRAIN": "('coco_2017_train',)", "TEST": "('coco_2017_val',)"},
    "SOLVER": {"IMS_PER_BATCH": 16, "BASE_LR": 0.02, "STEPS": "(60000, 80000)", "MAX_ITER": 90000},
    "INPUT": {"MIN_SIZE_TRAIN": "(640, 672, 704, 736, 768, 800)"},
}


def fcos_recipe(sup, burn, unsup_w, steps="(180000,)"):
    return {
        "_BASE_": "../Base-FCOS.yaml",
        "MODEL": {
            "WEIGHTS": "detectron2://ImageNetPretrained/MSRA/R-50.pkl",
            "RESNETS": {"DEPTH": 50},
            "FCOS": {"CENTER_SAMPLE": False, "NMS_CRITERIA_TRAIN": "cls", "NMS_CRITERIA_TEST": "cls_n_ctr",
                     "NMS_CRITERIA_REG_TRAIN": "cls_n_loc", "POS_RADIUS": 1.5, "YIELD_PROPOSAL": True,
                     "UNIFY_CTRCLS": False, "DFL_WEIGHT": 0.0, "REG_DISCRETE": True, "KL_LOSS": True,
                     "KLLOSS_WEIGHT": 0.05, "KL_LOSS_TYPE": "nlloss"},
        },
        "SOLVER": {"IMG_PER_BATCH_LABEL": 8, "IMG_PER_BATCH_UNLABEL": 8, "BASE_LR": 0.01, "STEPS": steps,
                   "MAX_ITER": 180000, "AMP": {"ENABLED": True}},
        "DATALOADER": {"SUP_PERCENT": sup, "RANDOM_DATA_SEED": 1},
        "DATASETS": {"CROSS_DATASET": False, "TRAIN": "('coco_2017_train',)", "TEST": "('coco_2017_val',)"},
        "SEMISUPNET": {
            "Trainer": "ubteacher", "PSEUDO_BBOX_SAMPLE": "thresholding", "PSEUDO_BBOX_SAMPLE_REG": "thresholding",
            "BBOX_THRESHOLD": 0.5, "BBOX_THRESHOLD_REG": 0.5, "BBOX_CTR_THRESHOLD": 0.0,
            "BBOX_CTR_THRESHOLD_REG": 0.0, "TEACHER_UPDATE_ITER": 1, "BURN_UP_STEP": burn,
            "EMA_KEEP_RATE": 0.9999, "UNSUP_LOSS_WEIGHT": unsup_w, "UNSUP_REG_LOSS_WEIGHT": 0.2,
            "CONSIST_CTR_LOSS": "pseudo", "PSEUDO_CLS_IGNORE_NEAR": False, "PSEUDO_CTR_THRES": 0.5,
            "SOFT_CLS_LABEL": False, "CLS_LOSS_METHOD": "focal", "CLS_LOSS_PSEUDO_METHOD": "focal",
            "TS_BETTER": 0.1, "CONSIST_REG_LOSS": "ts_locvar_better_nms_nll_l1", "ANALYSIS_PRINT_FRE": 5000,
            "ANALYSIS_ACCUMLATE_FRE": 50},
        "TEST": {"EVAL_PERIOD": 2000, "EVALUATOR": "COCOeval", "VAL_LOSS": False},
        "INPUT": {"MIN_SIZE_TRAIN": "(400, 1200)", "MIN_SIZE_TRAIN_SAMPLING": "range"},
    }


def rcnn_recipe(sup, unsup_w, eval_period=3000):
    return {
        "_BASE_": "../Base-RCNN-FPN.yaml",
        "MODEL": {
            "META_ARCHITECTURE": "TwoStagePseudoLabGeneralizedRCNN",
            "WEIGHTS": "detectron2://ImageNetPretrained/MSRA/R-50.pkl",
            "MASK_ON": False,
            "RESNETS": {"DEPTH": 50},
            "PROPOSAL_GENERATOR": {"NAME": "PseudoLabRPN"},
            "RPN": {"POSITIVE_FRACTION": 0.25, "LOSS": "CrossEntropy"},
            "ROI_HEADS": {"NAME": "StandardROIHeadsPseudoLab", "LOSS": "FocalLoss_BoundaryVar"},
            "ROI_BOX_HEAD": {"BBOX_REG_LOSS_TYPE": "nlloss", "CLS_AGNOSTIC_BBOX_REG": True},
        },
        "SOLVER": {"LR_SCHEDULER_NAME": "WarmupMultiStepLR", "STEPS": "(180000,)", "MAX_ITER": 180000,
                   "IMG_PER_BATCH_LABEL": 32, "IMG_PER_BATCH_UNLABEL": 32, "BASE_LR": 0.01},
        "DATALOADER": {"SUP_PERCENT": sup, "RANDOM_DATA_SEED": 0},
        "DATASETS": {"CROSS_DATASET": False, "TRAIN": "('coco_2017_train',)", "TEST": "('coco_2017_val',)"},
        "SEMISUPNET": {"Trainer": "ubteacher_rcnn", "BBOX_THRESHOLD": 0.7, "TEACHER_UPDATE_ITER": 1,
                       "BURN_UP_STEP": 2000, "EMA_KEEP_RATE": 0.9996, "UNSUP_LOSS_WEIGHT": unsup_w,
                       "UNSUP_REG_LOSS_WEIGHT": 1.0},
        "TEST": {"EVAL_PERIOD": eval_period, "EVALUATOR": "COCOeval", "VAL_LOSS": False},
        "INPUT": {"MIN_SIZE_TRAIN": "(400, 1200)", "MIN_SIZE_TRAIN_SAMPLING": "range"},
    }


FILES = {
    "FCOS/Base-FCOS.yaml": BASE_FCOS,
    "Faster-RCNN/Base-RCNN-FPN.yaml": BASE_RCNN,
    "FCOS/coco-standard/fcos_R_50_ut2_sup1_run0.yaml": fcos_recipe(1.0, 10000, 3.0),
    "FCOS/coco-standard/fcos_R_50_ut2_sup2_run0.yaml": fcos_recipe(2.0, 15000, 3.0, "(179995, 179999)"),
    "FCOS/coco-standard/fcos_R_50_ut2_sup5_run0.yaml": fcos_recipe(5.0, 20000, 2.0),
    "FCOS/coco-standard/fcos_R_50_ut2_sup10_run0.yaml": fcos_recipe(10.0, 30000, 2.0),
    "Faster-RCNN/coco-standard/faster_rcnn_R_50_FPN_ut2_sup1_run0.yaml": rcnn_recipe(1.0, 4.0),
    "Faster-RCNN/coco-standard/faster_rcnn_R_50_FPN_ut2_sup2_run0.yaml": rcnn_recipe(2.0, 3.0),
    "Faster-RCNN/coco-standard/faster_rcnn_R_50_FPN_ut2_sup5_run0.yaml": rcnn_recipe(5.0, 2.0),
    "Faster-RCNN/coco-standard/faster_rcnn_R_50_FPN_ut2_sup10_run0.yaml": rcnn_recipe(10.0, 4.0, 2000),
}

if __name__ == "__main__":
    for rel, content in FILES.items():
        path = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write("# generated by tools/make_configs.py — edit the table there\n")
            yaml.safe_dump(content, f, sort_keys=True, default_flow_style=None, width=100)
    print("wrote", len(FILES), "files to", OUT)
