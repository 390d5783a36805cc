#!/usr/bin/env python
"""Benchmark of the UT2 FCOS R50-FPN training step on B200 (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--label B] [--unlabel B]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)
  python bench.py --impl reference ...    CPU oracle port of the reference step on the host cores

One step = UBTeacherTrainer.run_step_full_semisup(): EMA teacher update, teacher forward on the weak unlabeled
views, two NMS criteria + pseudo-label thresholding, student forward/backward on 2*B_l labeled and B_u unlabeled
images, gradient all-reduce (N > 1), SGD. images/step = world * (B_l + B_u) unique images.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "unbiased-teacher-v2_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

FLOP_PER_11 = {"fcos": 4063.6e9, "rcnn": 3888.2e9}   # algorithmic FLOPs of one step at B_l = B_u = 1, 800x1344
RECIPE = {"fcos": "configs/FCOS/coco-standard/fcos_R_50_ut2_sup1_run0.yaml",
          "rcnn": "configs/Faster-RCNN/coco-standard/faster_rcnn_R_50_FPN_ut2_sup1_run0.yaml"}
ARCH_NAME = {"fcos": "FCOS R50-FPN", "rcnn": "Faster-RCNN R50-FPN"}
METRIC = "images/sec per UT2 train step (teacher fwd + student fwd/bwd)"
FULL_PIXELS = 800 * 1344


def build_cfg(n_label, n_unlabel, device="cuda", arch="fcos"):
    from ubteacher.config import add_ubteacher_config
    from ubteacher.d2compat.config import get_cfg
    cfg = get_cfg()
    add_ubteacher_config(cfg)
    cfg.merge_from_file(os.path.join(PKG, RECIPE[arch]))
    cfg.merge_from_list(["SEMISUPNET.BURN_UP_STEP", 0, "SOLVER.IMG_PER_BATCH_LABEL", n_label,
                         "SOLVER.IMG_PER_BATCH_UNLABEL", n_unlabel, "MODEL.DEVICE", device, "SEED", 7])
    return cfg


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops_sustained"], p["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained)"
    except Exception:  # noqa: BLE001
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                   "reasons": sorted(reasons)}
        os.unlink(self.f.name)
        return out


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_state(cfg, seed=7):
    """Seeded weights with the reference key names (built by the product's initialiser on the CPU is not possible —
    the engine is CUDA-only — so the oracle gets an equivalent seeded init of its own)."""
    import math
    import torch
    from oracle import ut2_model as M  # noqa: F401
    g = torch.Generator().manual_seed(seed)
    sd = {}
    bu = "backbone.bottom_up."

    def conv(name, co, ci, k, std=None, bias=False, bn=False, uniform=False):
        w = torch.empty(co, ci, k, k)
        if uniform:
            b = math.sqrt(3.0 / (ci * k * k))
            w.uniform_(-b, b, generator=g)
        else:
            w.normal_(0, std if std is not None else math.sqrt(2.0 / (co * k * k)), generator=g)
        sd[name + ".weight"] = w
        if bias:
            sd[name + ".bias"] = torch.zeros(co)
        if bn:
            sd[name + ".norm.weight"] = torch.ones(co)
            sd[name + ".norm.bias"] = torch.zeros(co)
            sd[name + ".norm.running_mean"] = torch.zeros(co)
            sd[name + ".norm.running_var"] = torch.ones(co) - 1e-5
    conv(bu + "stem.conv1", 64, 3, 7, bn=True)
    cin = 64
    for stage, n, mid, cout in (("res2", 3, 64, 256), ("res3", 4, 128, 512), ("res4", 6, 256, 1024), ("res5", 3, 512, 2048)):
        for i in range(n):
            p = f"{bu}{stage}.{i}."
            if cin != cout:
                conv(p + "shortcut", cout, cin, 1, bn=True)
            conv(p + "conv1", mid, cin, 1, bn=True)
            conv(p + "conv2", mid, mid, 3, bn=True)
            conv(p + "conv3", cout, mid, 1, bn=True)
            cin = cout
    for l, c in ((3, 512), (4, 1024), (5, 2048)):
        conv(f"backbone.fpn_lateral{l}", 256, c, 1, bias=True, uniform=True)
        conv(f"backbone.fpn_output{l}", 256, 256, 3, bias=True, uniform=True)
    conv("backbone.top_block.p6", 256, 256, 3, bias=True, uniform=True)
    conv("backbone.top_block.p7", 256, 256, 3, bias=True, uniform=True)
    hd = "proposal_generator.fcos_head."
    for t in ("cls_tower", "bbox_tower"):
        for i in range(4):
            conv(f"{hd}{t}.{3 * i}", 256, 256, 3, std=0.01, bias=True)
            sd[f"{hd}{t}.{3 * i + 1}.weight"] = torch.ones(256)
            sd[f"{hd}{t}.{3 * i + 1}.bias"] = torch.zeros(256)
    conv(hd + "cls_logits", 80, 256, 3, std=0.01, bias=True)
    sd[hd + "cls_logits.bias"].fill_(-math.log(99.0))
    conv(hd + "bbox_pred", 68, 256, 3, std=0.01, bias=True)
    conv(hd + "bbox_pred_std", 4, 256, 3, std=1e-4, bias=True)
    conv(hd + "ctrness", 1, 256, 3, std=0.01, bias=True)
    for i in range(5):
        sd[f"{hd}scales.{i}.scale"] = torch.ones(1)
    sd["pixel_mean"] = torch.tensor(cfg.MODEL.PIXEL_MEAN).view(3, 1, 1)
    sd["pixel_std"] = torch.tensor(cfg.MODEL.PIXEL_STD).view(3, 1, 1)
    sd["proposal_generator.fcos_outputs.integral.project"] = torch.linspace(0, 16, 17)
    return sd


def cpu_batch(n_label, n_unlabel, h, w, seed):
    import torch
    from ubteacher.data.synthetic import synth_instances
    g = torch.Generator().manual_seed(seed)

    def mk(n, with_gt):
        out = []
        for _ in range(n):
            d = {"image": torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8)}
            if with_gt:
                inst = synth_instances(g, h, w, 7)
                d["boxes"], d["classes"] = inst.gt_boxes.tensor, inst.gt_classes
            out.append(d)
        return out
    lq = mk(n_label, True)
    lk = [dict(d, image=torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8)) for d in lq]
    return lq, lk, mk(n_unlabel, False), mk(n_unlabel, False)


def cpu_step_runner(cfg, h, w, n_label=1, n_unlabel=1):
    """Returns a closure running one oracle UT2 step (oracle/ut2_model.py:ut2_step) on the host cores."""
    import torch
    from oracle import ut2_model as M
    student = cpu_state(cfg)
    teacher = {k: v.clone() for k, v in student.items()}
    mom = {}
    s = cfg.SEMISUPNET
    ocfg = {"UNSUP_LOSS_WEIGHT": s.UNSUP_LOSS_WEIGHT, "UNSUP_REG_LOSS_WEIGHT": s.UNSUP_REG_LOSS_WEIGHT,
            "EMA_KEEP_RATE": s.EMA_KEEP_RATE, "NMS_CRITERIA_TRAIN": cfg.MODEL.FCOS.NMS_CRITERIA_TRAIN,
            "NMS_CRITERIA_REG_TRAIN": cfg.MODEL.FCOS.NMS_CRITERIA_REG_TRAIN, "BBOX_THRESHOLD": s.BBOX_THRESHOLD,
            "BBOX_THRESHOLD_REG": s.BBOX_THRESHOLD_REG, "WEIGHT_DECAY": cfg.SOLVER.WEIGHT_DECAY,
            "WEIGHT_DECAY_NORM": cfg.SOLVER.WEIGHT_DECAY_NORM, "MOMENTUM": cfg.SOLVER.MOMENTUM, "LR": 1e-5}
    state = {"i": 0}

    def step():
        batch = cpu_batch(n_label, n_unlabel, h, w, 20260 + state["i"])
        M.ut2_step(student, teacher, mom, batch, ocfg, state["i"] == 0)
        state["i"] += 1
    return step, torch.get_num_threads()


def cpu_step_runner_rcnn(cfg, h, w, n_label=1, n_unlabel=1):
    """One oracle UT2 Faster R-CNN step (oracle/ut2_rcnn_model.py:ut2_rcnn_step) on the host cores."""
    import torch
    from oracle import ut2_rcnn_model as RM
    student = RM.init_state_dict(7)
    teacher = {k: v.clone() for k, v in student.items()}
    mom = {}
    s = cfg.SEMISUPNET
    ocfg = {"UNSUP_LOSS_WEIGHT": s.UNSUP_LOSS_WEIGHT, "UNSUP_REG_LOSS_WEIGHT": s.UNSUP_REG_LOSS_WEIGHT,
            "EMA_KEEP_RATE": s.EMA_KEEP_RATE, "BBOX_THRESHOLD": s.BBOX_THRESHOLD, "WEIGHT_DECAY": cfg.SOLVER.WEIGHT_DECAY,
            "MOMENTUM": cfg.SOLVER.MOMENTUM, "LR": 1e-5}
    state = {"i": 0}
    Hp, Wp = (h + 31) // 32 * 32, (w + 31) // 32 * 32
    A, hh, ww = 0, Hp // 4, Wp // 4
    for _ in range(5):
        A += 3 * hh * ww
        hh, ww = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1

    def step():
        g = torch.Generator().manual_seed(77 + state["i"])
        batch = cpu_batch(n_label, n_unlabel, h, w, 20260 + state["i"])
        rk = lambda n, m: [torch.randint(0, 2 ** 32, (m,), generator=g, dtype=torch.int64) for _ in range(n)]
        keys = {"rpn_sup": rk(2 * n_label, A), "roi_sup": rk(2 * n_label, 1200), "rpn_unsup": rk(n_unlabel, A),
                "roi_unsup": rk(n_unlabel, 1200)}
        RM.ut2_rcnn_step(student, teacher, mom, batch, ocfg, state["i"] == 0, keys)
        state["i"] += 1
    return step, torch.get_num_threads()


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the step cannot run (Detectron2 absent, hard-coded
    .cuda(): BASELINE.md §2), so this arm times the oracle port of it on the host cores. Each step is a bounded sample of
    the workload: ONE labeled + ONE unlabeled FULL-SIZE image (3 x 800 x 1333 -> 800 x 1344, BASELINE config #1) at every
    N, so that the per-N series is comparable. Only if the first (warm-up) step says that K + W such steps would not fit
    REF_BUDGET_S on this host is the resolution lowered, and then throughput is reported in full-size-image equivalents."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = build_cfg(1, 1, device="cpu", arch=args.arch)
    runner = cpu_step_runner if args.arch == "fcos" else cpu_step_runner_rcnn
    port = "oracle/ut2_model.py:ut2_step" if args.arch == "fcos" else "oracle/ut2_rcnn_model.py:ut2_rcnn_step"
    budget = float(os.environ.get("REF_BUDGET_S", "600"))
    total = args.steps + max(args.warmup, 1)
    h, w = 800, 1333
    step, cores = runner(cfg, h, w)
    t0 = time.perf_counter()
    step()                                              # warm-up step 1 doubles as the calibration
    t_full = time.perf_counter() - t0
    done_warm = 1
    if t_full * total > budget:                         # a very slow host: shrink the sample (cost ~ linear in pixels)
        for scale in (0.75, 0.5, 0.375, 0.25, 0.1875, 0.125):
            h, w = int(800 * scale) // 32 * 32, int(1333 * scale) // 32 * 32
            if t_full * scale * scale * total <= budget:
                break
        step, cores = runner(cfg, h, w)
        done_warm = 0
    for _ in range(max(args.warmup - done_warm, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    eq_images = 2.0 * (((h + 31) // 32 * 32) * ((w + 31) // 32 * 32)) / FULL_PIXELS
    v = eq_images * args.steps / dt
    sample = (f"oracle port ({port}, fp32 torch CPU), 1 labeled + 1 unlabeled image of {h}x{w} per step "
              f"({eq_images:.3f} full-size-image equivalents), {args.steps} timed steps after {max(args.warmup, 1)} warm-up")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{ARCH_NAME[args.arch]} UT2 run_step_full_semisup, label {args.label} + unlabel {args.unlabel} per GPU, "
                                   "synthetic 3x800x1333 (reference arm: bounded CPU sample, see cpu_baseline.sample)",
                       "sample_hw": [h, w], "full_size_sample": (h, w) == (800, 1333)},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def saturate(trainer, arch):
    """`--regime saturated`: widen the student's prediction heads (the teacher copies them at the burn-in boundary) so that
    every level is full of candidates and pseudo labels survive the thresholds."""
    import torch
    g = torch.Generator().manual_seed(3)
    V = trainer.model.engine.arena.views
    if arch == "fcos":
        hd = "proposal_generator.fcos_head."
        spec = ((hd + "cls_logits.weight", 0.03), (hd + "bbox_pred.weight", 0.05), (hd + "ctrness.weight", 0.05),
                (hd + "bbox_pred_std.weight", 0.05))
        V[hd + "cls_logits.bias"].fill_(-3.8)
    else:
        spec = (("proposal_generator.rpn_head.objectness_logits.weight", 0.05), ("proposal_generator.rpn_head.anchor_deltas.weight", 0.02),
                ("roi_heads.box_predictor.cls_score.weight", 0.3), ("roi_heads.box_predictor.bbox_pred.weight", 0.04),
                ("roi_heads.box_predictor.bbox_pred_std.weight", 0.04))
    for name, scale in spec:
        V[name].copy_((torch.randn(V[name].shape, generator=g) * scale).to(V[name].device))
    trainer.model.engine.refresh_operands()


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--label", type=int, default=8, help="labeled images per GPU per step (B_l)")
    ap.add_argument("--unlabel", type=int, default=8, help="unlabeled images per GPU per step (B_u)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--arch", default="fcos", choices=["fcos", "rcnn"],
                    help="fcos = BASELINE config #2 (the headline workload); rcnn = the Faster R-CNN recipe (configs #3 / #5)")
    ap.add_argument("--augment", action="store_true", help="produce the strong views with the device two-crop augmentation "
                    "(SURVEY 8(f) rank 1) inside every timed step")
    ap.add_argument("--regime", default="cold", choices=["cold", "saturated"],
                    help="cold: seeded [D2]-style init (teacher scores ~0.01: no pseudo labels, the empty-GT path); saturated: the "
                         "prediction heads are widened so that every level is full of candidates and pseudo labels survive the "
                         "thresholds (worst-case top-k / NMS / target-assignment load, SURVEY.md 8d)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` arms of the JSON line (saturated regime, the other "
                    "detector, the recipes' 2 + 2 images per GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        # the segment all-reduces run under the backward pass on a few SMs (engine/trainer.py:_begin_overlap)
        os.environ.setdefault("NCCL_MAX_CTAS", "8")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ubteacher import _C, ops
    from ubteacher.d2compat.events import EventStorage
    from ubteacher.data.synthetic import SyntheticTwoCropLoader
    from ubteacher.engine import UBRCNNTeacherTrainer, UBTeacherTrainer

    Trainer = UBTeacherTrainer if args.arch == "fcos" else UBRCNNTeacherTrainer
    cfg = build_cfg(args.label * world, args.unlabel * world, arch=args.arch)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(trainer, steps, read_back, profile=False):
        """K steps bracketed by barrier + synchronize, device-timed with CUDA events, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _C.launch_count
        if profile:
            ops.PROFILE = {"conv_fwd": [], "conv_wgrad": []}
        t0 = time.perf_counter()
        e0.record()
        d2h = 0
        for i in range(steps):
            trainer.iter = trainer.iter + 1
            trainer.run_step_full_semisup()
            if read_back:
                host = trainer.last_losses[1].cpu()       # D2H read of the step's result
                d2h = host.numel() * host.element_size()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        prof, ops.PROFILE = ops.PROFILE, None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / 1e3, _C.launch_count - l0, d2h, prof, wall

    def pseudo_stats(trainer):
        if trainer.last_pseudo is None:
            return None
        return [float(p.counts.float().mean()) for p in trainer.last_pseudo]

    images_per_step = world * (args.label + args.unlabel)
    # ---- arm 1: inputs resident in HBM ---------------------------------------------------------------------
    loader = SyntheticTwoCropLoader(args.label, args.unlabel, rank=rank, device=dev, strong_augment=args.augment)
    tr = Trainer(cfg, data_loader=loader)
    tr.storage = EventStorage(0)
    tr.metrics_period = 10 ** 9
    tr.iter = -1
    if args.regime == "saturated":
        saturate(tr, args.arch)
    # roofline pass: a few eager steps with every tensor-core launch bracketed by CUDA events
    timed(tr, max(args.warmup, 3), False)
    # per-launch CUDA events need the launches to own the device: no second stream during the two profiled steps
    prev_ct = os.environ.get("UT2_CONCURRENT_TEACHER")
    os.environ["UT2_CONCURRENT_TEACHER"] = "0"
    _, _, _, prof, _ = timed(tr, 2, False, profile=True)
    if prev_ct is None:
        del os.environ["UT2_CONCURRENT_TEACHER"]
    else:
        os.environ["UT2_CONCURRENT_TEACHER"] = prev_ct
    prof_steps = 2
    if not args.no_graph:
        tr.enable_cuda_graph(True)
    timed(tr, args.warmup, False)
    clocks = ClockSampler(local) if rank == 0 else None
    secs, launches, _, _, wall = timed(tr, args.steps, False)
    clk = clocks.stop() if clocks else {}
    value = images_per_step * args.steps / secs
    pseudo_per_image = pseudo_stats(tr)
    peak_tf, peak_bw, peak_src = peaks()
    roof = None
    if prof:
        fl = sum(r[2] for r in prof["conv_fwd"])
        ms = sum(r[0].elapsed_time(r[1]) for r in prof["conv_fwd"])
        flw = sum(r[2] for r in prof["conv_wgrad"])
        msw = sum(r[0].elapsed_time(r[1]) for r in prof["conv_wgrad"])
        # every conv_fwd launch against ITS OWN bound: max(algorithmic FLOPs / tensor peak, algorithmic bytes / HBM peak)
        ideal = [max(r[2] / (peak_tf * 1e12), r[3] / (peak_bw * 1e9)) * 1e3 for r in prof["conv_fwd"]]
        tb = [r[2] / (peak_tf * 1e12) >= r[3] / (peak_bw * 1e9) for r in prof["conv_fwd"]]
        ms_each = [r[0].elapsed_time(r[1]) for r in prof["conv_fwd"]]
        ms_tb = sum(m for m, t in zip(ms_each, tb) if t)
        fl_tb = sum(r[2] for r, t in zip(prof["conv_fwd"], tb) if t)
        by_mb = sum(r[3] for r, t in zip(prof["conv_fwd"], tb) if not t)
        ach = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        traffic, traffic_note = None, None
        try:        # DRAM bytes of the dominant launch shape from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_note = (f"NOT measured live: read from the committed ncu --set full capture profiles/r02_traffic.json; "
                            f"bytes per launch of the dominant shape {tj['shape']}: algorithmic {tj['algorithmic_bytes']} B; "
                            f"`achieved` averages all {len(prof['conv_fwd']) // prof_steps} conv_fwd launches of a step")
        except Exception:  # noqa: BLE001
            pass
        roof = {"bound": "tensor", "kernel": "conv_fwd_kernel (implicit-GEMM fwd + dgrad, tcgen05; the ~29 narrow 3x3 launches per step that ut2_conv2d_* hands to conv3x3_halo_kernel are counted here too)", "achieved": ach,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peak_src,
                "launches_per_step": len(prof["conv_fwd"]) / prof_steps, "kernel_ms_per_step": ms / prof_steps,
                "per_launch_bound": {
                    "note": "the same kernel runs tensor-bound (3x3, head) and HBM-bound (narrow-K 1x1 + fused residual / mask epilogues) "
                            "launches; each launch is held against max(FLOPs / tensor peak, algorithmic bytes / HBM peak)",
                    "frac": sum(ideal) / ms if ms > 0 else 0.0, "ideal_ms_per_step": sum(ideal) / prof_steps,
                    "tensor_bound": {"launches_per_step": sum(tb) / prof_steps, "ms_per_step": ms_tb / prof_steps,
                                     "achieved_tflops": fl_tb / (ms_tb * 1e-3) / 1e12 if ms_tb > 0 else 0.0,
                                     "frac": fl_tb / (ms_tb * 1e-3) / 1e12 / peak_tf if ms_tb > 0 else 0.0},
                    "hbm_bound": {"launches_per_step": (len(tb) - sum(tb)) / prof_steps, "ms_per_step": (ms - ms_tb) / prof_steps,
                                  "achieved_gbs": by_mb / ((ms - ms_tb) * 1e-3) / 1e9 if ms > ms_tb else 0.0,
                                  "frac": by_mb / ((ms - ms_tb) * 1e-3) / 1e9 / peak_bw if ms > ms_tb else 0.0}},
                "flops_per_launch_avg": fl / max(len(prof["conv_fwd"]), 1),
                "wgrad": {"achieved": flw / (msw * 1e-3) / 1e12 if msw > 0 else 0.0, "kernel_ms_per_step": msw / prof_steps,
                          "launches_per_step": len(prof["conv_wgrad"]) / prof_steps},
                "step_flops_frac": FLOP_PER_11[args.arch] * (args.label + args.unlabel) / 2.0 * world / secs * args.steps / (world * peak_tf * 1e12)
                if args.label == args.unlabel else None}
    del tr, loader
    torch.cuda.empty_cache()

    def extra_arm(arch, label, unlabel, regime):
        """What the driver cannot see otherwise: the same device-timed measurement (inputs resident in HBM,
        graph replay, max over ranks) for another regime / detector / per-GPU batch, reduced to ms, images/s, step_flops_frac."""
        T = UBTeacherTrainer if arch == "fcos" else UBRCNNTeacherTrainer
        c = build_cfg(label * world, unlabel * world, arch=arch)
        ld = SyntheticTwoCropLoader(label, unlabel, rank=rank, device=dev)
        t = T(c, data_loader=ld)
        t.storage = EventStorage(0)
        t.metrics_period = 10 ** 9
        t.iter = -1
        if regime == "saturated":
            saturate(t, arch)
        timed(t, 3, False)                       # eager steps before the capture
        if not args.no_graph:
            t.enable_cuda_graph(True)
        timed(t, max(args.warmup, 3), False)
        sec, nl, _, _, _ = timed(t, args.steps, False)
        ms = 1e3 * sec / args.steps
        out = {"workload": f"{ARCH_NAME[arch]} UT2 step, label {label} + unlabel {unlabel} per GPU, {regime} regime, n_gpus {world}",
               "ms_per_step": ms, "value": world * (label + unlabel) / (ms * 1e-3), "unit": "images/s",
               "step_flops_frac": FLOP_PER_11[arch] * (label + unlabel) / 2.0 / (ms * 1e-3) / (peak_tf * 1e12) if label == unlabel else None,
               "pseudo_boxes_per_image": pseudo_stats(t), "gpu_launches": nl}
        del t, ld
        torch.cuda.empty_cache()
        return out

    extra = None
    if not args.no_extras:
        other = "rcnn" if args.arch == "fcos" else "fcos"
        extra = {"saturated" if args.regime == "cold" else "cold":
                 extra_arm(args.arch, args.label, args.unlabel, "saturated" if args.regime == "cold" else "cold"),
                 other: extra_arm(other, args.label, args.unlabel, args.regime),
                 "per_gpu_2": extra_arm(args.arch, 2, 2, args.regime),
                 other + "_per_gpu_2": extra_arm(other, 2, 2, args.regime)}
        # BASELINE config #5: the Faster R-CNN per-GPU batch sweep 2 / 4 / 8 / 16 (2 and 8 are above)
        for b in (4, 16):
            extra[f"rcnn_per_gpu_{b}"] = extra_arm("rcnn", b, b, args.regime)
    # ---- arm 2: end to end through the public API: pinned host inputs, H2D inside the step, loss read back --
    e2e = None
    if not args.no_e2e:
        loader = SyntheticTwoCropLoader(args.label, args.unlabel, rank=rank, device=None, strong_augment=args.augment)
        tr = Trainer(cfg, data_loader=loader)
        tr.storage = EventStorage(0)
        tr.metrics_period = 10 ** 9
        tr.iter = -1
        if args.regime == "saturated":
            saturate(tr, args.arch)
        if not args.no_graph:
            tr.enable_cuda_graph(True)
        timed(tr, max(args.warmup, 3), True)
        secs2, _, d2h, _, _ = timed(tr, args.steps, True)
        n_up = (args.label + args.unlabel) if args.augment else (2 * args.label + 2 * args.unlabel)   # augment: only the weak views cross PCIe
        h2d = n_up * 3 * 800 * 1333 + 2 * args.label * (128 * 4 * 4 + 128 * 8 + 4)
        e2e = {"value": images_per_step * args.steps / secs2, "unit": "images/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * secs2 / args.steps}
        del tr, loader
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        runner = cpu_step_runner if args.arch == "fcos" else cpu_step_runner_rcnn
        port = "oracle/ut2_model.py:ut2_step" if args.arch == "fcos" else "oracle/ut2_rcnn_model.py:ut2_rcnn_step"
        torch.set_num_threads(os.cpu_count() or 1)
        step, cores = runner(build_cfg(1, 1, device="cpu", arch=args.arch), 800, 1333)
        step()                                   # warm-up
        n_cpu = 3
        t0 = time.perf_counter()
        for _ in range(n_cpu):
            step()
        dt = (time.perf_counter() - t0) / n_cpu
        cpu = {"value": 2.0 / dt, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"oracle port ({port}, fp32 torch CPU): full steps with 1 labeled + 1 unlabeled 3x800x1333 image "
                         f"(BASELINE config #1), 1 warm-up + {n_cpu} timed, {dt:.2f} s per step"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{ARCH_NAME[args.arch]} UT2 run_step_full_semisup, IMG_PER_BATCH_LABEL={args.label} "
                                       f"UNLABEL={args.unlabel} per GPU, synthetic uint8 3x800x1333 (padded 800x1344), "
                                       f"BURN_UP_STEP=0, random init ({args.regime} pseudo-label regime)",
                           "regime": args.regime, "pseudo_boxes_per_image": pseudo_per_image,
                           "global_batch": images_per_step, "parallelism": f"dp{world}",
                           "l2_policy": "no flush needed: each step streams >10 GB of activations (>> 126 MB L2)",
                           "launch_mode": "eager" if args.no_graph else "whole step replayed as one CUDA graph",
                           "two_crop_augmentation": "device (strong = aug(weak) every step, outside the graph)" if args.augment else "none (pre-made views)"},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "extra": extra, "gpu_launches": launches, "clocks": clk,
                "host_wall_s": wall}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
