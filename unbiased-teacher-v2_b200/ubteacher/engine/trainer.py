"""``UBTeacherTrainer`` — the FCOS Unbiased-Teacher-v2 training loop with the reference's surface
(ubteacher/engine/trainer.py:38-610): ctor(cfg), train(), run_step_full_semisup(), _update_teacher_model(),
_write_metrics(), attributes model / model_teacher / optimizer / pseudo_generator / iter / storage.

What changed underneath (B200-first, same arithmetic):
  * teacher and student are FcosEngine replicas (flat arenas, tensor-core implicit-GEMM convs);
  * pseudo labels never leave the device (no .item()/nonzero syncs: fixed-capacity sets + counters);
  * the labeled and the unlabeled student passes are back-propagated right after their forward with the
    reference's loss weights as the incoming gradient (identical sum of gradients, half the live activations);
  * data parallel = one NCCL all-reduce of the contiguous gradient arena, mean folded into the SGD kernel;
  * EMA is one kernel over the whole state (trainer.py:468-486 is a ~320-tensor Python loop);
  * metrics are stacked on the device and read back once every `metrics_period` iterations.
"""
import contextlib
import logging
import os
import time

import numpy as np
import torch

from .. import ops as ops_mod
from ..checkpoint import DetectionTSCheckpointer
from ..d2compat import comm
from ..d2compat.events import EventStorage
from ..d2compat.registry import META_ARCH_REGISTRY
from ..modeling.meta_arch.ts_ensemble import EnsembleTSModel
from ..modeling.pseudo_generator import PseudoGenerator
from ..solver.lr_scheduler import WarmupMultiStepLR

logger = logging.getLogger(__name__)
_NVTX = os.environ.get("UT2_NVTX", "1") != "0"


class nvtx_range:
    """NVTX phase range (tracing). Host-side markers only: legal during CUDA-graph capture, visible in an
    Nsight Systems timeline of the eager steps; the replayed step shows up as one `ut2.step.graph_replay` range."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if _NVTX:
            torch.cuda.nvtx.range_pop()


class ArenaSGD:
    """torch.optim-like facade over the fused SGD kernel ([D2] build_optimizer: SGD, momentum 0.9, wd 1e-4,
    WEIGHT_DECAY_NORM 0 for the GroupNorm parameters, no gradient clipping)."""

    def __init__(self, cfg, model):
        s = cfg.SOLVER
        self.model = model
        self.momentum = s.MOMENTUM
        self.param_groups = [
            {"name": "decay", "lr": s.BASE_LR, "initial_lr": s.BASE_LR, "weight_decay": s.WEIGHT_DECAY},
            {"name": "norm", "lr": s.BASE_LR, "initial_lr": s.BASE_LR, "weight_decay": s.WEIGHT_DECAY_NORM},
        ]
        self.steps = 0
        self.grad_scale = 1.0
        dev = model.engine.device
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)      # read by the SGD kernel (graph replay)

    def zero_grad(self, set_to_none=False):
        pass  # the SGD kernel clears the gradient arena in the same pass that consumes it

    def push_lr(self):
        # a fill kernel with the value as a launch argument: stream-ordered, nothing the host could overwrite before the
        # device reads it (a pinned scalar + async copy would race with the next iteration's write)
        self.lr_dev.fill_(float(self.param_groups[0]["lr"]))

    def step(self, use_device_lr=False):
        eng = self.model.engine
        eng.sgd_step(self.param_groups[0]["lr"], self.momentum, self.param_groups[0]["weight_decay"],
                     self.param_groups[1]["weight_decay"], self.steps == 0, self.grad_scale,
                     self.lr_dev if use_device_lr else None)
        eng.refresh_operands()
        self.steps += 1

    def state_dict(self):
        return {"steps": self.steps, "momentum_buffer": self.model.engine.arena.mom,
                "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        if "steps" in sd:                       # this trainer's own format
            self.steps = sd["steps"]
            self.model.engine.arena.mom.copy_(sd["momentum_buffer"])
            return
        # a torch.optim.SGD state dict (what the reference's checkpoints hold): {"state": {i: {"momentum_buffer"}},
        # "param_groups": [{"params": [i, ...]}, ...]}. [D2] builds one group per parameter in named_parameters() order of the
        # trainable parameters, which is this model's order too; shapes are checked, anything else is skipped loudly.
        A = self.model.engine.arena
        names = [n for n, p in self.model.named_parameters() if p.requires_grad]
        ids = [i for g in sd.get("param_groups", []) for i in g.get("params", [])]
        state = sd.get("state", {})
        if len(ids) != len(names):
            logger.warning("optimizer state has %d parameters, the model %d: momentum buffers not loaded", len(ids), len(names))
            return
        from ..arena import _view
        loaded = 0
        for i, n in zip(ids, names):
            buf = state.get(i, {}).get("momentum_buffer")
            if buf is None:
                continue
            v = _view(A.mom, A.offset[n], A.specs[n].shape)
            if buf.numel() != v.numel():
                logger.warning("momentum buffer of %s has %d elements, expected %d: skipped", n, buf.numel(), v.numel())
                continue
            v.copy_(buf.to(v.device, torch.float32).reshape(v.shape))
            loaded += 1
        self.steps = 1 if loaded else 0


class UBTeacherTrainer:
    def __init__(self, cfg, data_loader=None):
        self.cfg = cfg
        if comm.get_world_size() > 1:
            torch.cuda.set_device(comm.get_local_rank())
        model = self.build_model(cfg)
        self.optimizer = self.build_optimizer(cfg, model)
        model_teacher = self.build_model(cfg)
        self.model_teacher = model_teacher
        self._set_teacher_mode(model_teacher)
        self.model = model
        self.model.train()
        self.data_loader = data_loader if data_loader is not None else self.build_train_loader(cfg)
        self._data_loader_iter = iter(self.data_loader)
        self.scheduler = self.build_lr_scheduler(cfg, self.optimizer)
        self.ensem_ts_model = EnsembleTSModel(model_teacher, model)
        self.pseudo_generator = self._make_pseudo_generator(cfg)
        # trainer.py:71-76: the checkpointer owns the teacher+student ensemble, the optimizer and the scheduler
        self.checkpointer = DetectionTSCheckpointer(self.ensem_ts_model, cfg.OUTPUT_DIR, save_to_disk=comm.is_main_process(),
                                                    optimizer=self.optimizer, scheduler=self.scheduler)
        self.start_iter = 0
        self.iter = 0
        self.max_iter = cfg.SOLVER.MAX_ITER
        self.storage = None
        self._hooks = []
        self._last_eval_results_teacher = self._last_eval_results_student = None
        self.metrics_period = 20                        # PeriodicWriter period (trainer.py:551)
        self._metric_names, self._metric_buf = None, []
        self.last_losses = None
        self.last_pseudo = None
        self._prefetched = None                         # next batch, when its H2D copy was started early (graph mode)
        self.use_cuda_graph = False                     # enable_cuda_graph(): replay the whole semi-sup step
        # CUDA graphs of the step, one per batch geometry (multi-scale training pads every batch to its own size): key = image
        # shapes of the four views -> {"static", "graph", "staged", ...}; all captures share one memory pool
        self._graphs = {}
        self._seen_once = {}
        self._graph_pool = None
        self._cur = None
        # gradient all-reduce overlapped with the last backward pass of a step (UT2_OVERLAP_ALLREDUCE=0: one call after it)
        self.overlap_allreduce = comm.get_world_size() > 1 and os.environ.get("UT2_OVERLAP_ALLREDUCE", "1") != "0"
        self._comm_stream = None
        self._grads_reduced = False
        if comm.get_world_size() > 1:                   # DDP's initial parameter broadcast
            torch.distributed.broadcast(model.engine.arena.data, 0)
            model.engine.refresh_operands()
            self.optimizer.grad_scale = 1.0 / comm.get_world_size()
            for m in (model, model_teacher):            # every rank draws its own anchor / proposal samples (the reference's
                if hasattr(m.engine, "seed"):           # ranks have independent torch RNG streams)
                    m.engine.seed = (int(m.engine.seed) * 1000003 + comm.get_rank() + 1) & 0x7FFFFFFF

    def _make_pseudo_generator(self, cfg):
        return PseudoGenerator(cfg)

    def _set_teacher_mode(self, model_teacher):
        model_teacher.eval()                            # trainer.py:55

    # ---------------------------------------------------------------- builders ([D2] DefaultTrainer API)
    @classmethod
    def build_model(cls, cfg):
        return META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)

    @classmethod
    def build_optimizer(cls, cfg, model):
        return ArenaSGD(cfg, model)

    @classmethod
    def build_lr_scheduler(cls, cfg, optimizer):
        name = cfg.SOLVER.LR_SCHEDULER_NAME
        if name != "WarmupMultiStepLR":        # solver/lr_scheduler.py:9-53 (the reference's recipes use nothing else)
            raise ValueError("Unknown LR scheduler: {}".format(name))
        return WarmupMultiStepLR(optimizer, cfg.SOLVER.STEPS, cfg.SOLVER.GAMMA, cfg.SOLVER.WARMUP_FACTOR,
                                 cfg.SOLVER.WARMUP_ITERS, cfg.SOLVER.WARMUP_METHOD)

    # no dataset ships with this repo: the synthetic loaders are used only when asked for (train_net.py --synthetic,
    # bench.py and the tests pass their loader explicitly); otherwise cfg.DATASETS.* must be registered in the DatasetCatalog
    allow_synthetic = False

    @classmethod
    def build_train_loader(cls, cfg):
        from ..d2compat.catalog import DatasetCatalog
        names = [n for n in cfg.DATASETS.TRAIN if n in DatasetCatalog]
        if names:       # data/build.py:144-272: label / unlabel split by SUP_PERCENT + RANDOM_DATA_SEED, two-crop device mapper
            from ..data.build import build_detection_semisup_train_loader_two_crops
            dicts = [d for n in names for d in DatasetCatalog.get(n)]
            return build_detection_semisup_train_loader_two_crops(cfg, dataset_dicts=dicts)
        if not cls.allow_synthetic:
            raise RuntimeError(f"none of cfg.DATASETS.TRAIN = {tuple(cfg.DATASETS.TRAIN)} is registered in the DatasetCatalog "
                               "(train_net.py --dataset-json NAME=FILE); pass --synthetic to train on random images")
        from ..data.synthetic import SyntheticTwoCropLoader
        logger.warning("--synthetic: training on the synthetic two-crop loader (random images, SURVEY.md §8d)")
        w = comm.get_world_size()
        return SyntheticTwoCropLoader(cfg.SOLVER.IMG_PER_BATCH_LABEL // w, cfg.SOLVER.IMG_PER_BATCH_UNLABEL // w,
                                      rank=comm.get_rank())

    @classmethod
    def build_evaluator(cls, cfg, dataset_name=None, output_folder=None):
        """trainer.py:104-127 builds [D2]'s COCOEvaluator; pycocotools is not available here, so the host-side COCO-style box
        AP (ubteacher/evaluation/box_ap.py) stands in: same DatasetEvaluator protocol, ground truth taken from the inputs."""
        from ..evaluation import BoxAPEvaluator
        return BoxAPEvaluator()

    @classmethod
    def build_test_loader(cls, cfg, dataset_name=None, num_batches=4, batch_size=2):
        """A fixed-length list of batches: the registered cfg.DATASETS.TEST split through the weak (test-time) mapper, or —
        only when synthetic data was asked for — synthetic labeled batches."""
        from ..d2compat.catalog import DatasetCatalog
        name = dataset_name or (cfg.DATASETS.TEST[0] if len(cfg.DATASETS.TEST) else None)
        if name is not None and name in DatasetCatalog:
            from ..data.build import read_image
            from ..data.dataset_mapper import DatasetMapperTwoCropSeparate
            mapper = DatasetMapperTwoCropSeparate(cfg, False)
            dicts = DatasetCatalog.get(name)
            out = []
            for i in range(0, len(dicts), batch_size):
                chunk = [dict(d, image=read_image(d["file_name"], cfg.INPUT.FORMAT)) for d in dicts[i:i + batch_size]]
                _, weak = mapper(chunk)
                for d, src in zip(weak, dicts[i:i + batch_size]):
                    d["annotations"] = src.get("annotations", [])
                out.append(weak)
            return out
        if not cls.allow_synthetic:
            raise RuntimeError(f"test dataset {name!r} is not registered in the DatasetCatalog; pass --synthetic for random images")
        from ..data.synthetic import SyntheticTwoCropLoader
        ld = SyntheticTwoCropLoader(batch_size, 1, rank=0, pin=False)
        out = []
        for i in range(num_batches):
            _, lk, _, _ = next(ld)
            for j, d in enumerate(lk):
                d["image_id"] = i * batch_size + j
            out.append(lk)
        return out

    @classmethod
    def test(cls, cfg, model, evaluators=None):
        """[D2] DefaultTrainer.test through the reference's inference_on_dataset (trainer.py:129-159, evaluation/evaluator.py)."""
        from ..evaluation import inference_on_dataset
        loader = cls.build_test_loader(cfg)
        evaluator = evaluators if evaluators is not None else cls.build_evaluator(cfg)
        return inference_on_dataset(model, loader, evaluator, cfg)

    def resume_or_load(self, resume=True):
        """trainer.py:86-102 ([D2] DefaultTrainer.resume_or_load): last checkpoint of OUTPUT_DIR when resuming, else
        MODEL.WEIGHTS (a Caffe2 pickle initialises the student backbone only). ``detectron2://`` URLs need the network
        and are skipped with a warning when no local copy exists."""
        import os
        path = self.cfg.MODEL.WEIGHTS
        if path and "://" in path and not os.path.exists(path):
            logger.warning("MODEL.WEIGHTS %s is not reachable offline: keeping the seeded initialisation", path)
            path = ""
        ckpt = self.checkpointer.resume_or_load(path, resume=resume)
        if resume and self.checkpointer.has_checkpoint():
            self.start_iter = ckpt.get("iteration", -1) + 1
            self.iter = self.start_iter
        return ckpt

    # ---------------------------------------------------------------- loop
    def train(self):
        """trainer.py:177-179 + [D2] DefaultTrainer.train: run the loop, return the last teacher evaluation."""
        self.train_loop(self.start_iter, self.max_iter)
        if self._last_eval_results_teacher is not None and comm.is_main_process():
            return self._last_eval_results_teacher

    def build_hooks(self):
        """trainer.py:503-552: periodic checkpoints (rank 0), evaluation of the student and of the teacher every
        TEST.EVAL_PERIOD iterations, periodic metric writers (rank 0; metrics.json + log line, every 20 iterations)."""
        from . import hooks
        cfg = self.cfg
        ret = []
        if comm.is_main_process():
            ret.append(hooks.PeriodicCheckpointer(self.checkpointer, cfg.SOLVER.CHECKPOINT_PERIOD))

        def test_and_save_results_student():
            self._last_eval_results_student = self.test(self.cfg, self.model)
            return {k + "_student": v for k, v in self._last_eval_results_student.items()}

        def test_and_save_results_teacher():
            self._last_eval_results_teacher = self.test(self.cfg, self.model_teacher)
            return self._last_eval_results_teacher

        ret.append(hooks.EvalHook(cfg.TEST.EVAL_PERIOD, test_and_save_results_student))
        ret.append(hooks.EvalHook(cfg.TEST.EVAL_PERIOD, test_and_save_results_teacher))
        if comm.is_main_process():
            ret.append(hooks.PeriodicWriter(self.build_writers(), period=self.metrics_period))
        return ret

    def build_writers(self):
        from . import hooks
        return [hooks.CommonMetricPrinter(self.max_iter), hooks.JSONWriter(os.path.join(self.cfg.OUTPUT_DIR, "metrics.json"))]

    def register_hooks(self, hook_list):
        for h in hook_list:
            if h is not None:
                h.trainer = self
                self._hooks.append(h)

    def train_loop(self, start_iter, max_iter):
        self.iter = self.start_iter = start_iter
        self.max_iter = max_iter
        if not self._hooks:
            self.register_hooks(self.build_hooks())
        with EventStorage(start_iter) as self.storage:
            try:
                for h in self._hooks:
                    h.before_train()
                for self.iter in range(start_iter, max_iter):
                    for h in self._hooks:
                        h.before_step()
                    self.run_step_full_semisup()
                    self.storage.put_scalar("lr", self.optimizer.param_groups[0]["lr"], smoothing_hint=False)
                    self.scheduler.step()                # [D2] hooks.LRScheduler.after_step
                    for h in self._hooks:
                        h.after_step()
                    self.storage.step()
            finally:
                for h in self._hooks:
                    h.after_train()

    # ---------------------------------------------------------------- pseudo-labeling helpers (trainer.py:161-175)
    def remove_label(self, label_data):
        for d in label_data:
            if "instances" in d:
                del d["instances"]
        return label_data

    def add_label(self, unlabled_data, label, labeltype=""):
        key = {"class": "instances_class", "reg": "instances_reg"}.get(labeltype, "instances")
        for d in unlabled_data:
            d[key] = label       # the whole batch's device-resident BoxSet (one object shared by the dicts)
        return unlabled_data

    # ---------------------------------------------------------------- the step (trainer.py:181-429)
    def run_step_full_semisup(self):
        assert self.model.training, "[UBTeacherTrainer] model was changed to eval mode!"
        ss = self.cfg.SEMISUPNET
        start = time.perf_counter()
        if self._prefetched is not None:             # fetched (and already on its way to the device) during the last step
            data, self._prefetched = self._prefetched, None
        else:
            data = next(self._data_loader_iter)
        data_time = time.perf_counter() - start
        if self.use_cuda_graph and self.iter > ss.BURN_UP_STEP and self.optimizer.steps > 0 and \
                (self.iter - ss.BURN_UP_STEP) % ss.TEACHER_UPDATE_ITER == 0:
            return self._graph_step(data, data_time)
        self._step_body(data, data_time)

    # ---------------------------------------------------------------- CUDA-graph replay of the step
    def enable_cuda_graph(self, flag=True):
        """The semi-supervised step has no host synchronisation and static shapes (fixed-capacity pseudo-label sets),
        so after the first eager steps it is captured and replayed: ~4000 launches per step cost one cudaGraphLaunch.
        Inputs are copied into static buffers; the learning rate lives in device memory. One graph per batch geometry
        (multi-scale training: INPUT.MIN_SIZE_TRAIN draws a size per image, so padded batches come in a few dozen shapes):
        a geometry runs eagerly the first time it is seen (its lazily built per-shape buffers must exist before a capture), is
        captured the second time, and replayed from then on; all captures share one memory pool (they never run
        concurrently). UT2_GRAPH_CACHE caps the number of cached geometries (default 48); beyond it new ones stay eager.
        Uniform batches are keyed by the image shape. Batches of mixed image sizes are keyed by the padded size of their three
        forward groups: the images' own sizes are device data (ops.STATIC_SIZES -> hw_dev of the stem, ops.STATIC_HW -> image_hw
        of the R-CNN clipping kernels), so one graph serves every batch of that padded size."""
        self.use_cuda_graph = flag
        if not flag:
            self._graphs, self._seen_once, self._graph_pool, self._cur = {}, {}, None, None
            # (entries of the dropped static buffers die with them: ops.lookup_static checks the weak reference)

    # True: batches of mixed image sizes are keyed by the PADDED size of their three forward groups, the images' own sizes travel
    # as device data (the fused stem reads them from memory: ops.STATIC_SIZES; the R-CNN proposal / detection clipping reads
    # ops.STATIC_HW). UT2_GRAPH_PADDED_KEY=0: exact image shapes only.
    graph_padded_key = os.environ.get("UT2_GRAPH_PADDED_KEY", "1") != "0"

    def _batch_key(self, data):
        lq, lk, uq, uk = data
        shapes = tuple(tuple(d["image"].shape) for d in lq + lk + uq + uk)
        if not self.graph_padded_key or len(set(shapes)) == 1:
            return shapes                       # uniform batches (and the R-CNN step): the images' own shapes
        pad = self.model.engine.padded_size
        return ("padded",) + tuple((len(g),) + pad([tuple(d["image"].shape[1:]) for d in g]) for g in (lq + lk, uq, uk))

    def _graph_entry(self, data, create=True):
        """Entry of this batch's geometry, or None (first sight, or cache full). Geometries seen once are only remembered as
        keys (a bounded LRU): batches of a dataset with free aspect ratios never repeat and must not fill the cache."""
        key = self._batch_key(data)
        e = self._graphs.get(key)
        if e is not None or not create:
            return e
        if key not in self._seen_once:
            self._seen_once[key] = True
            while len(self._seen_once) > 1024:
                self._seen_once.pop(next(iter(self._seen_once)))
            return None
        if len(self._graphs) >= int(os.environ.get("UT2_GRAPH_CACHE", "48")):
            return None
        self._seen_once.pop(key, None)
        e = self._graphs[key] = {"key": key, "static": None, "graph": None, "staged": None}
        return e

    def _stage_inputs(self, e, data):
        from ..modeling.fcos.fcos_outputs import BoxSet, as_boxset
        dev = self.model.device
        lq, lk, uq, uk = data
        lab = lq + lk
        gt = lab[0]["instances"] if isinstance(lab[0]["instances"], BoxSet) else as_boxset([d["instances"] for d in lab], dev)
        imgs = [d["image"] for d in lq + lk + uq + uk]
        n = [len(lq), len(lk), len(uq), len(uk)]
        o = [0, n[0], n[0] + n[1], n[0] + n[1] + n[2], sum(n)]
        padded = e["key"][0] == "padded"
        if e["static"] is None:
            st = {"gt_shape": tuple(gt.boxes.shape),
                  "gt": BoxSet(torch.empty_like(gt.boxes), torch.empty_like(gt.classes), torch.empty_like(gt.counts))}
            if padded:
                # flat buffers at the padded capacity of the image's forward group + one device int32 [n, 2] of (h, w) per group;
                # the stem finds the sizes through the group's first buffer (ops.STATIC_SIZES)
                caps = [e["key"][1][1:]] * (n[0] + n[1]) + [e["key"][2][1:]] * n[2] + [e["key"][3][1:]] * n[3]
                st["flat"] = [torch.empty(3 * hp * wp, dtype=torch.uint8, device=dev) for hp, wp in caps]
                st["sizes"] = [torch.empty((k, 2), dtype=torch.int32, device=dev) for k in (n[0] + n[1], n[2], n[3])]
                st["hw_f"] = [torch.empty((k, 2), dtype=torch.float32, device=dev) for k in (n[0] + n[1], n[2], n[3])]
                for first, sz, hwf in zip((o[0], o[2], o[3]), st["sizes"], st["hw_f"]):
                    ops_mod.register_static(ops_mod.STATIC_SIZES, st["flat"][first], sz)
                    ops_mod.register_static(ops_mod.STATIC_HW, st["flat"][first], hwf)
            else:
                st["imgs"] = [torch.empty(i.shape, dtype=torch.uint8, device=dev) for i in imgs]
            e["static"] = st
        st = e["static"]
        if tuple(gt.boxes.shape) != st["gt_shape"]:
            return None
        if padded:
            st["imgs"] = [f[:i.numel()].view(i.shape) for f, i in zip(st["flat"], imgs)]      # this batch's views of the buffers
            hw = [[int(i.shape[1]), int(i.shape[2])] for i in imgs]
            for sz, hwf, (a, b) in zip(st["sizes"], st["hw_f"], ((o[0], o[2]), (o[2], o[3]), (o[3], o[4]))):
                sz.copy_(torch.tensor(hw[a:b], dtype=torch.int32))
                hwf.copy_(sz)
        if padded or "data" not in st:
            mk = lambda a, b, with_gt: [dict({"image": t}, **({"instances": st["gt"]} if with_gt else {})) for t in st["imgs"][a:b]]
            st["data"] = (mk(o[0], o[1], True), mk(o[1], o[2], True), mk(o[2], o[3], False), mk(o[3], o[4], False))
        for dst, src in zip(st["imgs"], imgs):
            dst.copy_(src, non_blocking=True)
        st["gt"].boxes.copy_(gt.boxes, non_blocking=True)
        st["gt"].classes.copy_(gt.classes, non_blocking=True)
        st["gt"].counts.copy_(gt.counts, non_blocking=True)
        return st["data"]

    def _prefetch_next(self):
        """Called right after the graph launch of step i: fetch batch i+1 and, when it lives in (pinned) host memory and its
        geometry already has static buffers, start its host->device copies on a side stream into staging buffers so that they
        overlap step i's kernels; step i+1 then only does device-to-device copies into the graph's static inputs.
        Device-resident batches are just fetched."""
        from ..modeling.fcos.fcos_outputs import BoxSet, as_boxset
        try:
            data = next(self._data_loader_iter)
        except StopIteration:
            return
        self._prefetched = data
        lq, lk, uq, uk = data
        imgs = [d["image"] for d in lq + lk + uq + uk]
        e = self._graph_entry(data, create=False)
        if e is None or e["static"] is None or e["key"][0] == "padded" or any(i.is_cuda for i in imgs):
            return      # (mixed-size batches are staged by _stage_inputs only: their views change with every batch)
        dev = self.model.device
        if e["staged"] is None:
            e["staged"] = {"imgs": [torch.empty_like(t) for t in e["static"]["imgs"]], "stream": torch.cuda.Stream(device=dev),
                           "copied": torch.cuda.Event(), "consumed": torch.cuda.Event()}
            e["staged"]["consumed"].record()
        sg = e["staged"]
        with torch.cuda.stream(sg["stream"]):
            sg["stream"].wait_event(sg["consumed"])          # the previous staging contents were copied out
            for dst, src in zip(sg["imgs"], imgs):
                dst.copy_(src, non_blocking=True)
            lab = lq + lk
            gt = lab[0]["instances"] if isinstance(lab[0]["instances"], BoxSet) else as_boxset([d["instances"] for d in lab], dev)
            sg["gt"] = gt
            sg["copied"].record()
        sg["for"] = data

    def _stage_prefetched(self, e, data):
        """Static inputs <- staging buffers (device to device) for a batch whose H2D copies were started by _prefetch_next."""
        sg, st = e["staged"], e["static"]
        if sg is None or st is None or sg.get("for") is not data or tuple(sg["gt"].boxes.shape) != st["gt_shape"]:
            return None
        sg["for"] = None
        cur = torch.cuda.current_stream()
        cur.wait_event(sg["copied"])
        for dst, src in zip(st["imgs"], sg["imgs"]):
            dst.copy_(src, non_blocking=True)
        st["gt"].boxes.copy_(sg["gt"].boxes, non_blocking=True)
        st["gt"].classes.copy_(sg["gt"].classes, non_blocking=True)
        st["gt"].counts.copy_(sg["gt"].counts, non_blocking=True)
        sg["consumed"].record()
        return st["data"]

    def _graph_step(self, data, data_time):
        e = self._graph_entry(data)
        if e is None:       # first sight of this geometry (the eager step also builds its lazily allocated buffers), or cache full
            return self._step_body(data, data_time)
        static = self._stage_prefetched(e, data) or self._stage_inputs(e, data)
        if static is None:                      # a differently shaped label set: run it eagerly
            return self._step_body(data, data_time)
        self.optimizer.push_lr()
        if e["graph"] is None:
            from .. import _C
            torch.cuda.synchronize()
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            g = torch.cuda.CUDAGraph()
            l0 = _C.launch_count
            with torch.cuda.graph(g, pool=self._graph_pool):
                # fresh dict views for every capture: remove_label/add_label mutate the dicts
                self._step_body(tuple([dict(d) for d in part] for part in static), 0.0, device_lr=True, bookkeeping=False)
            e["graph"] = g
            e["launches"] = _C.launch_count - l0
            _C.launch_count = l0
            e["names"], e["vec"] = self.last_losses
        self._cur = e
        with nvtx_range("ut2.step.graph_replay"):
            e["graph"].replay()
        self._prefetch_next()
        from .. import _C
        _C.launch_count += e["launches"]      # kernels inside the replayed graph
        self.optimizer.steps += 1
        self.last_losses = (e["names"], e["vec"])
        self._host_metrics({"data_time": data_time}, e["names"], e["vec"])

    def _step_body(self, data, data_time, device_lr=False, bookkeeping=True):
        cfg, ss = self.cfg, self.cfg.SEMISUPNET
        label_data_q, label_data_k, unlabel_data_q, unlabel_data_k = data
        record = {}
        if self.iter < ss.BURN_UP_STEP:
            losses, pending = self.model.forward_train(label_data_q + label_data_k, "labeled")
            record.update(losses)
            self.model.backward_pending(pending, [[1.0, 1.0, 1.0, 0.0]])
        else:
            if self.iter == ss.BURN_UP_STEP:
                self._update_teacher_model(keep_rate=0.00)
                ema_keep_rate = ss.EMA_KEEP_RATE
            elif (self.iter - ss.BURN_UP_STEP) % ss.TEACHER_UPDATE_ITER == 0:
                ema_keep_rate = ss.EMA_KEEP_RATE
                self._update_teacher_model(keep_rate=ema_keep_rate)
            else:
                ema_keep_rate = ss.EMA_KEEP_RATE
            record["ema_rate_1000x"] = ema_keep_rate * 1000
            # teacher on the weak views (trainer.py:231-237) + second NMS criterion (:240-242); on a second stream next to the
            # student's labeled pass when the per-GPU batch is small (_teacher_stream)
            side = self._teacher_stream(len(unlabel_data_k))
            main = torch.cuda.current_stream()
            if side is not None:
                side.wait_stream(main)
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                with nvtx_range("ut2.teacher_forward"):
                    pred_teacher, raw_pred_teacher = self.model_teacher(
                        unlabel_data_k, output_raw=True, nms_method=cfg.MODEL.FCOS.NMS_CRITERIA_TRAIN, branch="teacher_weak")
                with nvtx_range("ut2.pseudo_labels"):
                    raw_pred_teacher["scales"] = self.model_teacher.engine.scales
                    pred_teacher_loc = self.pseudo_generator.nms_from_dense(raw_pred_teacher, cfg.MODEL.FCOS.NMS_CRITERIA_REG_TRAIN)
                    thr = self._threshold(ss.PSEUDO_BBOX_SAMPLE, ss.BBOX_THRESHOLD, ss.BBOX_CTR_THRESHOLD)
                    thr_reg = self._threshold(ss.PSEUDO_BBOX_SAMPLE_REG, ss.BBOX_THRESHOLD_REG, ss.BBOX_CTR_THRESHOLD_REG)
                    pseudo_cls, _ = self.pseudo_generator.process_pseudo_label(pred_teacher, thr, "roih", ss.PSEUDO_BBOX_SAMPLE)
                    pseudo_reg, _ = self.pseudo_generator.process_pseudo_label(pred_teacher_loc, thr_reg, "roih",
                                                                               ss.PSEUDO_BBOX_SAMPLE_REG)
            self.last_pseudo = (pseudo_cls, pseudo_reg)          # device-resident; read by benchmarks / analysis only
            unlabel_data_q = self.remove_label(unlabel_data_q)
            unlabel_data_q = self.add_label(unlabel_data_q, pseudo_cls, "class")
            unlabel_data_q = self.add_label(unlabel_data_q, pseudo_reg, "reg")
            lam, mu = ss.UNSUP_LOSS_WEIGHT, ss.UNSUP_REG_LOSS_WEIGHT
            # student: labeled strong + weak (trainer.py:315-322), weights :378-416
            with nvtx_range("ut2.student_labeled_forward"):
                losses, pending = self.model.forward_train(label_data_q + label_data_k, "labeled")
            record.update(losses)
            with nvtx_range("ut2.student_labeled_backward"):
                self.model.backward_pending(pending, [[1.0 / (lam + 1.0), 1.0 / (mu + 1.0), 1.0 / (lam + 1.0), 0.0]])
            if side is not None:
                main.wait_stream(side)         # the unlabeled pass needs the pseudo labels
            # student: unlabeled strong with the two pseudo-label sets (trainer.py:331-349)
            with nvtx_range("ut2.student_unlabeled_forward"):
                losses_u, pending_u = self.model.forward_train(unlabel_data_q, "unlabeled", ss.PSEUDO_CLS_IGNORE_NEAR)    # trainer.py:340,347
            record.update({k + "_pseudo": v for k, v in losses_u.items()})
            with nvtx_range("ut2.student_unlabeled_backward"):
                self._begin_overlap()
                try:
                    self.model.backward_pending(pending_u, [[lam / (lam + 1.0), 0.0, lam / (lam + 1.0), 0.0],
                                                            [0.0, mu / (mu + 1.0), 0.0, 0.0]])
                finally:
                    self._end_overlap()
        record["data_time"] = data_time
        self._write_metrics(record, bookkeeping)
        self._reduce_and_step(device_lr, bookkeeping)

    # ---------------------------------------------------------------- teacher forward on a second stream
    def _teacher_stream(self, n_unlabel):
        """The teacher forward + pseudo-labelling do not depend on the student's labeled pass (only the unlabeled pass needs the
        pseudo labels). With small per-GPU batches (the README recipes run 2 + 2 images per GPU) most launches cannot fill 148
        SMs: the two branches run on two streams — forked and joined inside the captured graph — and share the device.
        UT2_CONCURRENT_TEACHER = auto (default: per-GPU unlabeled batch <= 8; neutral at 8 + 8, -2 ... -5 % at 2 + 2) | 1 | 0. Returns (side stream | None)."""
        mode = os.environ.get("UT2_CONCURRENT_TEACHER", "auto")
        on = mode == "1" or (mode == "auto" and n_unlabel <= 8)
        if not on:
            return None
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.model.device)
        return self._side_stream

    # ---------------------------------------------------------------- overlapped gradient all-reduce
    def _begin_overlap(self):
        """Arm the engine's segment hook for the LAST backward of the step: as soon as the gradients of a segment (FPN + heads,
        res5, res4, res3) are final, their NCCL all-reduce is enqueued on a side stream and runs under the rest of the
        backward pass. The persistent conv kernels are sized to (SMs - NCCL CTAs) meanwhile: a grid that fills the device
        could not finish before the collective's CTAs release their SMs."""
        if not self.overlap_allreduce:
            return
        from .. import _C
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.model.device)
            self._sm_total = int(_C.lib().ut2_device_sm_count())
            self._nccl_ctas = int(os.environ.get("NCCL_MAX_CTAS", "8"))
        grad = self.model.engine.arena.grad

        def hook(lo, hi):
            if hi <= lo:
                return
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream), nvtx_range("ut2.grad_allreduce.segment"):
                torch.distributed.all_reduce(grad[lo:hi])
            _C.lib().ut2_set_sm_limit(self._sm_total - self._nccl_ctas)
        self.model.engine.grad_hook = hook

    def _end_overlap(self):
        if not self.overlap_allreduce:
            return
        from .. import _C
        self.model.engine.grad_hook = None
        _C.lib().ut2_set_sm_limit(0)
        torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._grads_reduced = True

    def _reduce_and_step(self, device_lr, bookkeeping):
        if comm.get_world_size() > 1 and not self._grads_reduced:   # one contiguous buffer (mean folded into the SGD kernel)
            with nvtx_range("ut2.grad_allreduce"):
                torch.distributed.all_reduce(self.model.engine.arena.grad)
        self._grads_reduced = False
        with nvtx_range("ut2.sgd_and_repack"):
            self.optimizer.zero_grad()
            self.optimizer.step(use_device_lr=device_lr)
        if not bookkeeping:
            self.optimizer.steps -= 1       # capture does not execute; _graph_step counts the replays

    @staticmethod
    def _threshold(method, t0, t1):
        if method == "thresholding":
            return t0
        if method == "thresholding_cls_ctr":
            return (t0, t1)
        raise ValueError

    # ---------------------------------------------------------------- metrics (trainer.py:431-466)
    def _write_metrics(self, metrics_dict, bookkeeping=True):
        names = [k for k, v in metrics_dict.items() if isinstance(v, torch.Tensor)]
        vec = torch.stack([metrics_dict[k].detach().float().reshape(()) for k in names])
        self.last_losses = (names, vec)
        if bookkeeping:
            self._host_metrics({k: v for k, v in metrics_dict.items() if not isinstance(v, torch.Tensor)}, names, vec)

    def _host_metrics(self, scalars, names, vec):
        if self.storage is None or (self.iter + 1) % self.metrics_period:
            return
        host = vec.cpu().tolist()                                   # the only D2H of the step
        md = {k: float(v) for k, v in scalars.items()}
        md.update(dict(zip(names, host)))
        all_md = comm.gather(md)
        if comm.is_main_process():
            if "data_time" in all_md[0]:
                self.storage.put_scalar("data_time", np.max([x.pop("data_time") for x in all_md]))
            md = {k: np.mean([x[k] for x in all_md]) for k in all_md[0].keys()}
            self.storage.put_scalar("total_loss", sum(v for k, v in md.items() if k[:4] == "loss"))
            if len(md) > 1:
                self.storage.put_scalars(**md)

    # ---------------------------------------------------------------- EMA (trainer.py:468-486)
    @torch.no_grad()
    def _update_teacher_model(self, keep_rate=0.996):
        t = self.model_teacher.engine
        with nvtx_range("ut2.ema_teacher_update"):
            t.ema_from(self.model.engine, keep_rate)
            t.refresh_operands(dgrad=False)        # the teacher is never back-propagated

    @torch.no_grad()
    def _copy_main_model(self):
        self.model_teacher.engine.arena.copy_from(self.model.engine.arena)
        self.model_teacher.engine.refresh_operands()


class UBRCNNTeacherTrainer(UBTeacherTrainer):
    """The Faster R-CNN Unbiased-Teacher-v2 loop (reference: ubteacher/engine/trainer.py:613-968): same surface as
    ``UBTeacherTrainer``; ``run_step_full_semisup`` follows :786-912 (teacher on the weak views -> threshold_bbox at
    BBOX_THRESHOLD -> student on labeled (strong + weak) and on strongly augmented unlabeled images with the pseudo
    labels; loss weights :880-905). Teacher and student are ``RcnnEngine`` replicas; pseudo labels stay on the device."""

    def _make_pseudo_generator(self, cfg):
        return None           # the R-CNN trainer thresholds the ROI-head detections itself (trainer.py:727-769)

    def _set_teacher_mode(self, model_teacher):
        model_teacher.train()  # the reference never switches the R-CNN teacher to eval (trainer.py:628-629)

    def _graph_step(self, data, data_time):
        # the sampling seeds baked into the captured launches are constants: the draw counter lives in device memory
        # (RcnnEngine.seed_dev, mixed into every anchor / proposal sampling key) and is advanced before each replay
        self.model.engine.seed_dev.fill_(int((self.iter + 1) & 0x7FFFFFFF))     # value travels as a launch argument: no host race
        return super()._graph_step(data, data_time)

    # ---------------------------------------------------------------- pseudo-labeling (trainer.py:727-769)
    def threshold_bbox(self, dets, thres=0.7, proposal_type="roih"):
        from .. import ops
        from ..modeling.fcos.fcos_outputs import BoxSet
        if proposal_type != "roih":
            raise ValueError("Error in proposal type.")
        o = ops.threshold_scatter(dets, 0, float(thres))
        return BoxSet(o["pred_boxes"], o["pred_classes"], o["count"], o["reg_pred_std"], o["scores"])

    def process_pseudo_label(self, proposals_rpn_unsup_k, cur_threshold, proposal_type, psedo_label_method=""):
        if psedo_label_method != "thresholding":
            raise ValueError("Unkown pseudo label boxes methods")
        out = self.threshold_bbox(proposals_rpn_unsup_k, thres=cur_threshold, proposal_type=proposal_type)
        return out, out.counts.float().mean()

    def add_label(self, unlabled_data, label, labeltype=""):
        for d in unlabled_data:
            d["instances"] = label          # the whole batch's device-resident pseudo-label BoxSet
        return unlabled_data

    # ---------------------------------------------------------------- the step (trainer.py:786-912)
    def _step_body(self, data, data_time, device_lr=False, bookkeeping=True):
        ss = self.cfg.SEMISUPNET
        label_data_q, label_data_k, unlabel_data_q, unlabel_data_k = data
        all_label_data = label_data_q + label_data_k if ss.USE_SUP_STRONG == "both" else label_data_k
        record = {}
        if self.iter < ss.BURN_UP_STEP:
            losses, pending = self.model.forward_train(all_label_data, "supervised")
            record.update(losses)
            self.model.backward_pending(pending, [1.0, 1.0, 1.0, 1.0])
        else:
            if self.iter == ss.BURN_UP_STEP:
                self._update_teacher_model(keep_rate=0.0)                     # copy, then the EMA below
            if (self.iter - ss.BURN_UP_STEP) % ss.TEACHER_UPDATE_ITER == 0:
                self._update_teacher_model(keep_rate=ss.EMA_KEEP_RATE)
            record["EMA_rate"] = ss.EMA_KEEP_RATE
            # teacher on the weak views; never switched to eval (trainer.py:830-837)
            side = self._teacher_stream(len(unlabel_data_k))
            main = torch.cuda.current_stream()
            if side is not None:
                side.wait_stream(main)
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                with nvtx_range("ut2.teacher_forward"):
                    _, proposals_rpn_unsup_k, proposals_roih_unsup_k, _ = self.model_teacher(unlabel_data_k, branch="unsup_data_weak")
                with nvtx_range("ut2.pseudo_labels"):
                    pseudo, _ = self.process_pseudo_label(proposals_roih_unsup_k, ss.BBOX_THRESHOLD, "roih", "thresholding")
            self.last_pseudo = (pseudo,)
            unlabel_data_q = self.add_label(self.remove_label(unlabel_data_q), pseudo)
            unlabel_data_k = self.add_label(self.remove_label(unlabel_data_k), pseudo)
            lam, mu = ss.UNSUP_LOSS_WEIGHT, ss.UNSUP_REG_LOSS_WEIGHT
            with nvtx_range("ut2.student_labeled_forward"):
                losses, pending = self.model.forward_train(all_label_data, "supervised")
            record.update(losses)
            with nvtx_range("ut2.student_labeled_backward"):
                self.model.backward_pending(pending, [1.0, 1.0, 1.0, 1.0])
            if side is not None:
                main.wait_stream(side)         # the unlabeled pass needs the pseudo labels
            with nvtx_range("ut2.student_unlabeled_forward"):
                losses_u, pending_u = self.model.forward_train(unlabel_data_q, "unsup_data_train")
            record.update({k + "_pseudo": v for k, v in losses_u.items()})
            # loss_rpn_loc_pseudo * 0, loss_box_reg_pseudo * UNSUP_REG_LOSS_WEIGHT, the two classification terms * UNSUP_LOSS_WEIGHT
            with nvtx_range("ut2.student_unlabeled_backward"):
                self._begin_overlap()
                try:
                    self.model.backward_pending(pending_u, [lam, 0.0, lam, mu])
                finally:
                    self._end_overlap()
        record["data_time"] = data_time
        self._write_metrics(record, bookkeeping)
        self._reduce_and_step(device_lr, bookkeeping)
