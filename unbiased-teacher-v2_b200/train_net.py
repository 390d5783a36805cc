#!/usr/bin/env python3
"""Entry point with the role of the reference's ``train_net.py`` (lines 15-73): config -> trainer picked by
``cfg.SEMISUPNET.Trainer`` -> ``resume_or_load`` -> ``train()`` (or ``--eval-only`` -> ``Trainer.test`` on the teacher).

Detectron2's ``default_argument_parser`` / ``launch`` are replaced by argparse and ``torch.distributed.run``: one process
per GPU, e.g.

    python train_net.py --config-file configs/FCOS/coco-standard/fcos_R_50_ut2_sup1_run0.yaml SOLVER.MAX_ITER 100
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train_net.py --config-file ... [--cuda-graph]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def argument_parser():
    p = argparse.ArgumentParser(description="Unbiased Teacher v2 on the B200 kernels")
    p.add_argument("--config-file", default="", metavar="FILE")
    p.add_argument("--resume", action="store_true", help="resume from the last checkpoint of OUTPUT_DIR")
    p.add_argument("--eval-only", action="store_true")
    p.add_argument("--cuda-graph", action="store_true", help="replay the semi-supervised step as one CUDA graph")
    p.add_argument("--dataset-json", action="append", default=[], metavar="NAME=FILE",
                   help="register a dataset (JSON list of Detectron2-style dataset dicts) under NAME; NAME must appear in "
                        "cfg.DATASETS.TRAIN / TEST (the reference registers the COCO splits in data/datasets/builtin.py)")
    p.add_argument("--synthetic", action="store_true", help="train / evaluate on synthetic random images when no dataset is registered")
    p.add_argument("opts", default=None, nargs=argparse.REMAINDER, help="KEY VALUE pairs merged into the config")
    return p


def setup(args):
    from ubteacher import add_ubteacher_config
    from ubteacher.d2compat.config import get_cfg
    cfg = get_cfg()
    add_ubteacher_config(cfg)
    if args.config_file:
        cfg.merge_from_file(args.config_file)
    cfg.merge_from_list(args.opts or [])
    return cfg


def pick_trainer(cfg):
    from ubteacher.engine import UBRCNNTeacherTrainer, UBTeacherTrainer
    if cfg.SEMISUPNET.Trainer == "ubteacher":
        return UBTeacherTrainer
    if cfg.SEMISUPNET.Trainer == "ubteacher_rcnn":
        return UBRCNNTeacherTrainer
    raise ValueError("Trainer Name is not found.")


def main(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        os.environ.setdefault("NCCL_MAX_CTAS", "8")      # segment all-reduces run under the backward pass on a few SMs
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
        args.opts = list(args.opts or []) + ["MODEL.DEVICE", f"cuda:{local}"]
    cfg = setup(args)
    Trainer = pick_trainer(cfg)
    from ubteacher.d2compat.catalog import register_json
    for spec in args.dataset_json:
        name, _, path = spec.partition("=")
        register_json(name, path)
    Trainer.allow_synthetic = bool(args.synthetic)
    if args.eval_only:
        from ubteacher.checkpoint import DetectionTSCheckpointer
        from ubteacher.modeling.meta_arch.ts_ensemble import EnsembleTSModel
        model, model_teacher = Trainer.build_model(cfg), Trainer.build_model(cfg)
        ensemble = EnsembleTSModel(model_teacher, model)
        weights = cfg.MODEL.WEIGHTS if os.path.exists(str(cfg.MODEL.WEIGHTS)) else ""
        DetectionTSCheckpointer(ensemble, cfg.OUTPUT_DIR).resume_or_load(weights, resume=args.resume)
        return Trainer.test(cfg, ensemble.modelTeacher)
    trainer = Trainer(cfg)
    trainer.resume_or_load(resume=args.resume)
    if args.cuda_graph:
        trainer.enable_cuda_graph(True)
    try:
        return trainer.train()
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    cli = argument_parser().parse_args()
    print("Command Line Args:", cli)
    main(cli)
